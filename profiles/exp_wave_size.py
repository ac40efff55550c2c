"""Experiment: device time per sequence as a function of the wave size (150-nt and 100-nt workloads)."""
import sys, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench, ccj_b200
ctx = ccj_b200.Context(0, str(ROOT / "params" / "rna_Turner04.par"), 2)
out = []
for n, sizes in ((150, (8, 16, 24, 32, 40, 48, 56, 64, 70)), (100, (32, 64, 96, 128, 192, 256, 340))):
    for b in sizes:
        if b > ctx.wave_capacity(n):
            continue
        seqs = bench.workload(0, b, n) if n == 150 else bench.workload2(b)
        ctx.prepare(seqs)
        ctx.fill()
        ms = min(ctx.fill() for _ in range(3))
        tb = ctx.traceback()
        out.append({"n": n, "wave": b, "fill_ms": ms, "tb_ms": tb, "ms_per_seq": (ms + tb) / b})
        print(out[-1], flush=True)
print(json.dumps(out))

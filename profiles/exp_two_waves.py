"""Experiment: do two independent half-size waves, filled concurrently from two contexts on the SAME GPU, use the
machine better than one full-size wave?  (k_roles is ALU-bound, the windows L1-bound, k_final memory-bound.)"""
import sys, time, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench, ccj_b200

par = str(ROOT / "params" / "rna_Turner04.par")
a = ccj_b200.Context(0, par, 2)
b = ccj_b200.Context(0, par, 2)
out = {}
for total in (32, 64):
    seqs = bench.workload(0, total)
    a.fold_batch(seqs[:8]); b.fold_batch(seqs[:8])
    # one context, one wave
    a.fold_batch(seqs)
    t0 = time.perf_counter(); f1 = a.fold_batch(seqs); t1 = time.perf_counter() - t0
    # two contexts, half a wave each, concurrently (ccj_fold_batch_multi: one host thread per context)
    half = total // 2
    import ctypes as C, numpy as np
    def two():
        blob, offs = ccj_b200.Context._pack(seqs)
        res = np.zeros(len(seqs), dtype=ccj_b200.RESULT_DTYPE)
        structs = np.zeros(len(blob), dtype=np.uint8)
        import threading
        outs = [None, None]
        def run(ctx, lo, hi, slot):
            outs[slot] = ctx.fold_batch(seqs[lo:hi])
        th = [threading.Thread(target=run, args=(a, 0, half, 0)), threading.Thread(target=run, args=(b, half, total, 1))]
        [t.start() for t in th]; [t.join() for t in th]
        return outs[0] + outs[1]
    two()
    t0 = time.perf_counter(); f2 = two(); t2 = time.perf_counter() - t0
    out[total] = {"one_wave_s": t1, "two_half_waves_s": t2, "identical": f1 == f2, "fill_ms_one": a.last_fill_ms}
    print(total, out[total], flush=True)
print(json.dumps(out))

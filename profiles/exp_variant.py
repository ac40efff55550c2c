"""Per-kernel-group times of a 32 x 150 nt fill for the library selected by CCJ_B200_LIB (build variants)."""
import os, sys, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench, ccj_b200
ctx = ccj_b200.Context(0, str(ROOT / "params" / "rna_Turner04.par"), 2)
seqs = bench.workload(0, 32)
ctx.prepare(seqs)
ONLY = os.environ.get("CCJ_EXP_PROFILED_ONLY") == "1"   # ablation builds: wrong tables, never run the guarded fill
if not ONLY:
    ctx.fill()
best = None
for _ in range(3):
    p = ctx.fill_profiled()
    if best is None or p["k4d_ms"] < best["k4d_ms"]:
        best = p
fill = None if ONLY else min(ctx.fill() for _ in range(3))
h = ctx.table4_hash(0, "PK")
print(json.dumps({"lib": os.environ.get("CCJ_B200_LIB", "default"), "roles": best["k4d_split_ms"], "windows": best["k4d_window_ms"],
                  "final": best["k4d_final_ms"], "P": best["kP_ms"], "fill_ms": fill, "hash": h[2]}))

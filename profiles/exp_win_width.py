"""Experiment: 16-lane groups in k_winLR for levels with m >= CCJ_WINLR_WIDE_FROM (one process per setting)."""
import os, sys, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench, ccj_b200
ctx = ccj_b200.Context(0, str(ROOT / "params" / "rna_Turner04.par"), 2)
seqs = bench.workload(0, 32)
ctx.prepare(seqs)
ctx.fill()
best = None
for _ in range(3):
    p = ctx.fill_profiled()
    if best is None or p["k4d_window_ms"] < best["k4d_window_ms"]:
        best = p
fill = min(ctx.fill() for _ in range(3))
ctx.traceback()
folds = ctx.fetch()
gold = bench.load_goldens("folds_long.json", "folds_config4.json")
chk = [(f, gold[f.sequence]) for f in folds if f.sequence in gold]
h = ctx.table4_hash(0, "PL"), ctx.table4_hash(0, "PR"), ctx.table4_hash(1, "PK")
print(json.dumps({"wide_from": os.environ.get("CCJ_WINLR_WIDE_FROM"), "window_ms": best["k4d_window_ms"], "roles_ms": best["k4d_split_ms"],
                  "final_ms": best["k4d_final_ms"], "fill_ms": fill, "golden": [len(chk), all(bench.same_as_golden(f, r) for f, r in chk)],
                  "hash": h}))

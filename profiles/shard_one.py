"""One in-process sharded fold (world ranks on one GPU), for timing and ncu: python profiles/shard_one.py N [WORLD]"""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ccj_b200
from ccj_b200 import shard5
n = int(sys.argv[1]); world = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctx = ccj_b200.Context(0, str(ROOT / "params" / "rna_Turner04.par"), 2)
grp = shard5.LocalGroup(ctx, world)
sh = grp.fold(shard5.config5_sequence(n))
f = sh.traceback()
print(n, world, grp.ms, f.energy)
lv = sh.level_ms()
print("P kernel ms", float(lv[:, 0].sum()), "allreduce", float(lv[:, 1].sum()), "2D+4D", float(lv[:, 2].sum()), "allgather", float(lv[:, 3].sum()))

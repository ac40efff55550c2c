"""Workload for compute-sanitizer (racecheck / initcheck / memcheck): a 4-sequence batch of 40-nt sequences through
the captured 3-stream graph (double-buffered window scratch, 3-deep role scratch), twice, plus the traceback.

    compute-sanitizer --tool racecheck python profiles/sanitize_small.py
"""
import random
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ccj_b200  # noqa: E402

rng = random.Random(77)
seqs = ["".join(rng.choice("ACGU") for _ in range(n)) for n in (40, 40, 37, 40)]
with ccj_b200.Context(0, ccj_b200.default_par_file("rna_Turner04.par"), 2) as ctx:
    a = ctx.fold_batch(seqs)
    b = ctx.fold_batch(seqs)
    assert a == b
    print("folds:", [f.energy for f in a], "launches:", ctx.last_fill_launches)

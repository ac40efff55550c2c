"""Pins the CPU restatement (oracle/ccj_oracle.cc) to the reference: golden table hashes and energies written
by the compiled reference, and -- where oracle/_ref/ccj_ref_dump exists -- live runs of the reference itself."""
import random
import re
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import oracle as orc  # noqa: E402


def test_oracle_matches_golden_table_hashes(golden_hashes):
    """All 22 gap tables and the 8 two-dimensional tables, bit for bit (FNV over every valid index)."""
    checked = 0
    for rec in golden_hashes:
        if len(rec["seq"]) > 60:
            continue  # the 80-nt case takes the plain loops half a minute; covered by the GPU suite
        tables, _ = orc.oracle_hashes(rec["seq"], rec["par"], rec["dangles"])
        assert tables == rec["tables"], rec["seq"]
        checked += 1
    assert checked >= 15


def test_oracle_energy_matches_golden_folds(golden_folds):
    """W[n] against the energy the reference printed, over both parameter sets, d0/d1/d2 and --noGU."""
    checked = 0
    for rec in golden_folds:
        n = len(rec["seq"])
        if rec["rc"] != 0 or n > 48 or n < 1 or any(x not in ("--noGU",) for x in rec["extra"]):
            continue
        if checked >= 90 and n > 30:
            continue
        m = re.search(r"\((-?[0-9.eE+-]+)\)\s*$", rec["stdout"])
        want = round(float(m.group(1)) * 100)
        got = orc.oracle_energy_dcal(rec["seq"], rec["par"], rec["dangles"], "--noGU" in rec["extra"])
        assert got == want, (rec["seq"], rec["par"], rec["dangles"], rec["extra"])
        checked += 1
    assert checked >= 60


@pytest.mark.skipif(not orc.REF_DUMP.exists(), reason="compiled reference (oracle/_ref) not present")
@pytest.mark.parametrize("par,dangles", [("rna_Turner04.par", 2), ("rna_Turner04.par", 1), ("rna_Turner04.par", 0),
                                         ("rna_DirksPierce09.par", 2)])
def test_oracle_matches_live_reference(par, dangles):
    rng = random.Random(sum(map(ord, par)) + dangles)
    for n in (17, 31, 44):
        seq = "".join(rng.choice("ACGU") for _ in range(n))
        want = subprocess.run([str(orc.REF_DUMP), "hash", str(ROOT / "params" / par), str(dangles), seq],
                              capture_output=True, text=True, check=True).stdout
        got = subprocess.run([str(orc.build_oracle()), "hash", orc.params_dump(par), str(dangles), "0", seq],
                             capture_output=True, text=True, check=True).stdout
        assert [l for l in got.splitlines() if not l.startswith("W ")] == want.splitlines(), seq


def test_window_term_count_matches_the_restated_fill(library):
    """the interior-window candidates the fill evaluates (the `iloop` part of the algorithmic bytes
    bench.py reports).  ccj_count_terms derives it from the pair table in closed form; the restatement counts the
    candidates while it evaluates them.  H60: 8.59e6, the figure SURVEY measured with an instrumented reference."""
    import random
    import ccj_b200
    H60 = "AAUAGGCGCAGCAUACACGGUCGAGCUGCGCCAAUAACAAUACGACCGUGAUAAAUAAAA"
    assert ccj_b200.count_terms(H60)["iloop"] == 8592675
    rng = random.Random(3)
    for seq in (H60, "".join(rng.choice("ACGU") for _ in range(47)), "".join(rng.choice("ACGU") for _ in range(23)), "ACGU"):
        assert ccj_b200.count_terms(seq)["iloop"] == orc.oracle_window_terms(seq)["iloop"], seq

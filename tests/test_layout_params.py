"""CPU: table layout (bijection, level-contiguity) and the energy-model loader against the golden
parameter dump of the reference (tests/golden/params_*.txt.gz, made by make_golden.py)."""
import gzip
import math
from pathlib import Path

import pytest

import ccj_b200

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("n", [3, 4, 5, 9, 17, 24])
def test_layout_is_a_bijection_onto_C_n1_4(library, n):
    seen = set()
    for i in range(1, n + 1):
        for j in range(i, n + 1):
            for k in range(j + 2, n + 1):
                for l in range(k, n + 1):
                    x = ccj_b200.layout_index(n, i, j, k, l)
                    assert 0 <= x < math.comb(n + 1, 4)
                    assert x not in seen
                    seen.add(x)
    assert len(seen) == math.comb(n + 1, 4) == ccj_b200.cells(n) == library.ccj_table4_len(n)


def test_layout_rows_are_contiguous_in_k(library):
    """Cells of one wavefront level with equal (a,b,i) are consecutive in k: a warp walking k is coalesced."""
    n = 20
    for a in range(0, 6):
        for b in range(0, 6):
            for i in range(1, n - a - b - 1):
                ks = list(range(i + a + 2, n - b + 1))
                idx = [ccj_b200.layout_index(n, i, i + a, k, k + b) for k in ks]
                assert idx == list(range(idx[0], idx[0] + len(ks)))


def test_invalid_indices(library):
    assert ccj_b200.layout_index(10, 1, 2, 3, 3) == -1      # j >= k-1
    assert ccj_b200.layout_index(10, 0, 1, 3, 3) == -1      # i < 1
    assert ccj_b200.layout_index(10, 1, 1, 3, 11) == -1     # l > n


def _parse(text):
    out = {}
    for line in text.splitlines():
        p = line.split()
        if len(p) >= 2:
            out[tuple(p[:-1])] = p[-1]
    return out


STUB_PAR = "## RNAfold parameter file v2.0\n\n#END\n"   # no section: the reference keeps all its compiled-in defaults


@pytest.mark.parametrize("par", ["rna_Turner04.par", "rna_DirksPierce09.par", "defaults", "partial"])
def test_scaled_model_matches_reference_dump(library, par, tmp_path):
    """"defaults": a file without sections; "partial": a file holding only # stack and # hairpin -- everything the
    file omits must keep the reference's compiled-in Turner-2004 values (src/ViennaRNA/params/default.c)."""
    if par == "defaults":
        (tmp_path / "stub.par").write_text(STUB_PAR)
        mine = _parse(ccj_b200.model_text(str(tmp_path / "stub.par"), 2))
    elif par == "partial":
        (tmp_path / "partial.par").write_text(partial_par())
        mine = _parse(ccj_b200.model_text(str(tmp_path / "partial.par"), 2))
    else:
        mine = _parse(ccj_b200.model_text(str(ROOT / "params" / par), 2))
    stem = par[:-4] if par.endswith(".par") else par
    with gzip.open(ROOT / "tests" / "golden" / f"params_{stem}.txt.gz", "rt") as f:
        ref = _parse(f.read())
    checked = 0
    for key, val in mine.items():
        name = key[0]
        if name in ("Tetraloop_E", "Triloop_E", "Hexaloop_E", "ninio"):
            continue
        idx = [int(x) for x in key[1:]]
        # int22 slots with a non-standard pair (7) or base 0 are filled by ViennaRNA's update_nst and can
        # never be indexed by a GCAU(T) sequence; the loader leaves them INF
        if name == "int22" and (0 in idx[2:] or 7 in idx[:2]):
            continue
        assert ref[key] == val, (key, val, ref[key])
        checked += 1
    assert checked > 20000
    assert mine[("ninio", "2")] == ref[("ninio", "2")]
    # special hairpins: names and energies in order
    for kind, width in (("Tetraloop", 7), ("Triloop", 6), ("Hexaloop", 9)):
        names = ref_names(ref, kind + "s", width)
        for pos, nm in enumerate(names):
            assert mine[(kind + "_E", nm)] == ref[(kind + "_E", str(pos))]


def partial_par():
    """The # stack and # hairpin sections of the DP09 file only (make_golden.py writes the same file for the reference)."""
    lines = (ROOT / "params" / "rna_DirksPierce09.par").read_text().splitlines()
    out, keep = [lines[0], ""], False
    for ln in lines[1:]:
        if ln.startswith("#"):
            keep = ln.strip() in ("# stack", "# hairpin")
        if keep:
            out.append(ln)
    return "\n".join(out) + "\n\n#END\n"


def test_embedded_dna_set_equals_the_par_file(library):
    """vrna_params_load_DNA_Mathews2004 (src/CCJ.cc:88-90): the set linked into the library is the reference's
    params/dna_Matthews04.par, which is byte-identical to its static/misc/dna_mathews2004.hex."""
    assert ccj_b200.model_text("@dna_mathews2004") == ccj_b200.model_text(str(ROOT / "params" / "dna_Matthews04.par"))
    assert ccj_b200.model_text("@rna_turner2004") == ccj_b200.model_text(str(ROOT / "params" / "rna_Turner04.par"))
    with pytest.raises(ccj_b200.CCJError):
        ccj_b200.model_text("@no_such_set")


def ref_names(ref, key, width):
    for k, v in ref.items():
        if k[0] == key:
            s = " ".join(list(k[1:]) + [v]).rstrip("|")
            return [x for x in s.split() if x]
    return []


def test_bad_parameter_file(tmp_path, library):
    with pytest.raises(ccj_b200.CCJError):
        ccj_b200.model_text(str(tmp_path / "nope.par"))
    bad = tmp_path / "bad.par"
    bad.write_text("## RNAfold parameter file v2.0\n# stack\n 1 2 three\n")
    with pytest.raises(ccj_b200.CCJError):
        ccj_b200.model_text(str(bad))


def test_parameter_reader_symmetry_warnings(library, tmp_path):
    """check_symmetry of the reference's reader (src/ViennaRNA/params/io.c:1126-1178): the Mathews-2004 DNA set makes
    the reference print four "stacking enthalpies not symmetric" warnings on stderr, the RNA sets none."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); import ccj_b200; "
            "ccj_b200.model_text(%r)") % (str(ROOT), "%s")
    for par, want in [("dna_Matthews04.par", "WARNING: stacking enthalpies not symmetric\n" * 4),
                      ("rna_Turner04.par", ""), ("rna_DirksPierce09.par", "")]:
        p = subprocess.run([sys.executable, "-c", code % str(ROOT / "params" / par)], capture_output=True, text=True)
        assert p.returncode == 0, p.stderr
        assert p.stderr == want, par
        ref = ROOT / "oracle" / "_ref" / "CCJ"
        if ref.exists() and par.startswith("dna"):
            r = subprocess.run([str(ref), "--noConv", "-P", str(ROOT / "params" / par), "GCAACGATGACATACATCGCTAGTCGACGC"],
                               capture_output=True, text=True)
            assert r.stderr == want

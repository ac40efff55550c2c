"""CPU: the product's cell functions (ccj_cells*.cuh, ccj_traceback.cuh -- the same code the CUDA
kernels call) compiled for the host and swept in the kernels' wavefront order, checked against the
golden vectors the reference produced (tests/golden/*.json).  Sized to stay within a few minutes."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def emu(emu_bin, mode, rec):
    args = [str(emu_bin), mode, str(ROOT / "params" / rec["par"]), str(rec["dangles"]), rec["seq"]]
    if "--noGU" in rec.get("extra", []):
        args.append("1")
    return subprocess.run(args, capture_output=True, text=True)


def test_golden_folds_small(emu_bin, golden_folds):
    recs = [r for r in golden_folds if len(r["seq"]) <= 36]
    assert len(recs) > 100
    n_err = 0
    for r in recs:
        p = emu(emu_bin, "fold", r)
        assert (p.returncode, p.stdout, p.stderr) == (r["rc"], r["stdout"], r["stderr"]), r["seq"]
        n_err += r["rc"] != 0
    assert n_err >= 1  # the corpus exercises the reference's exit(1) path


def test_golden_table_hashes_small(emu_bin, golden_hashes):
    recs = [r for r in golden_hashes if len(r["seq"]) <= 41]
    assert len(recs) >= 6
    for r in recs:
        p = emu(emu_bin, "hash", r)
        assert p.returncode == 0
        got = {}
        for line in p.stdout.splitlines()[1:]:
            name, cnt, agg, h = line.split()
            got[name] = [int(cnt), int(agg), h]
        assert got == r["tables"], r["seq"]


def test_emu_matches_reference_binary_when_present(emu_bin):
    """Extra randomised check, only where oracle/_ref was built (the build container)."""
    import random
    ref = ROOT / "oracle" / "_ref" / "CCJ"
    if not ref.exists():
        pytest.skip("oracle/_ref not built")
    for seed in range(12):
        rng = random.Random(4242 + seed)
        seq = "".join(rng.choice("ACGU") for _ in range(rng.randint(8, 34)))
        par = str(ROOT / "params" / "rna_Turner04.par")
        a = subprocess.run([str(ref), "-P", par, seq], capture_output=True, text=True)
        b = subprocess.run([str(emu_bin), "fold", par, "2", seq], capture_output=True, text=True)
        assert (a.returncode, a.stdout, a.stderr) == (b.returncode, b.stdout, b.stderr), seq


def test_real_int16_wrap(emu_bin, tmp_path):
    """The generic cell path narrows like Matrix4D::set: entries below -32768 dcal/mol wrap (tests/golden/wrap_real.json,
    the reference run with tripled stacking energies; see tests/test_gpu_parity.py)."""
    import json
    import sys
    sys.path.insert(0, str(ROOT / "tests" / "golden"))
    from make_golden import strong_stack_par, wrap87
    g = json.loads((ROOT / "tests" / "golden" / "wrap_real.json").read_text())
    par = strong_stack_par(tmp_path / "strong.par")
    p = subprocess.run([str(emu_bin), "hash", str(par), "2", wrap87()], capture_output=True, text=True)
    got = {}
    for line in p.stdout.splitlines()[1:]:
        name, cnt, agg, h = line.split()
        got[name] = [int(cnt), int(agg), h]
    assert got == g["tables"]
    assert min(v[1] for k, v in got.items() if k in ("PK", "PfromL", "PfromR")) < -32700   # at the edge of the range
    p = subprocess.run([str(emu_bin), "fold", str(par), "2", wrap87()], capture_output=True, text=True)
    assert (p.returncode, p.stdout, p.stderr) == (g["rc"], g["stdout"], g["stderr"])


def test_cell_functions_under_asan_ubsan(tmp_path, golden_folds, golden_hashes):
    """compute-sanitizer is closed on the GPU pool (profiles/r2_compute_sanitizer_closed.log), so the memory-safety check
    runs where it can: the product's cell, window-list and traceback functions -- the same source the kernels compile --
    built for the host with AddressSanitizer + UBSan (every table in an exactly sized heap buffer) and swept over the
    ordinary and the row-sharded layout.  Any out-of-bounds or uninitialised-index access aborts the tool."""
    from ccj_b200 import build
    exe = tmp_path / "ccj_emu_asan"
    srcs = [ROOT / "tests" / "emu" / "ccj_emu.cpp", ROOT / "ccj_b200" / "csrc" / "energy_model.cpp",
            ROOT / "ccj_b200" / "csrc" / "embedded_params.cpp"]
    subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all",
                    *build.embedded_par_defines(), "-o", str(exe)] + [str(s) for s in srcs], check=True)
    folds = [r for r in golden_folds if 18 <= len(r["seq"]) <= 30 and r["par"] == "rna_Turner04.par" and not r["extra"]][:4]
    folds += [r for r in golden_folds if r["rc"] != 0 and len(r["seq"]) <= 34][:1]
    assert len(folds) >= 4
    for r in folds:
        # [noGU, ranks, packed 2D records, partner lists, lean]; the last three: ccj_cells4_lean.cuh (ordinary, power-of-two
        # and other rank counts) -- its premultiplied level bases / 32-bit offsets must stay inside the exactly sized buffers
        for extra in ([], ["0", "0", "1", "1"], ["0", "4", "1", "1"], ["0", "3", "0", "0"],
                      ["0", "0", "1", "1", "1"], ["0", "4", "1", "1", "1"], ["0", "3", "1", "1", "1"]):
            args = [str(exe), "fold", str(ROOT / "params" / r["par"]), str(r["dangles"]), r["seq"]] + extra
            p = subprocess.run(args, capture_output=True, text=True)
            assert "Sanitizer" not in p.stderr and "runtime error" not in p.stderr, p.stderr[-2000:]
            assert (p.returncode, p.stdout, p.stderr) == (r["rc"], r["stdout"], r["stderr"]), (extra, r["seq"])
    h = [r for r in golden_hashes if len(r["seq"]) <= 26][:1]
    for r in h:
        p = subprocess.run([str(exe), "hash", str(ROOT / "params" / r["par"]), str(r["dangles"]), r["seq"], "0", "2", "1", "1"],
                           capture_output=True, text=True)
        assert p.returncode == 0 and "Sanitizer" not in p.stderr and "runtime error" not in p.stderr, p.stderr[-2000:]

"""CPU: the TUNED CUDA kernels themselves (ccj_b200/csrc/ccj_fill4.cu: k_prep_lay, k_fill_pmw, k_prep, k_roles, k_winLR,
k_winM, k_final, k_P_tuned) compiled by g++ for a small SIMT emulator (tests/emu/simt_emu.hpp: one OS thread per CUDA
thread, real barriers, warp collectives) and launched in the product's order with the product's grids
(tests/emu/ccj_emu_tuned.cpp).  Only the inline PTX has plain C++ stand-ins; the CUDA build of the same file is
byte-identical with and without those guards.

* every one of the 22 + 8 tables equals the golden vector the unmodified reference wrote (poisoned, exactly sized buffers);
* the same under AddressSanitizer + UBSan and under ThreadSanitizer -- compute-sanitizer (memcheck / initcheck /
  racecheck) is closed on the GPU pool (profiles/r2_compute_sanitizer_closed.log); a build whose __syncwarp() does nothing
  is the detector's negative control."""
import os
import random
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
CUDA_INC = Path(os.environ.get("CUDA_HOME", "/usr/local/cuda")) / "include"
SRCS = [ROOT / "tests" / "emu" / "ccj_emu_tuned.cpp", ROOT / "ccj_b200" / "csrc" / "energy_model.cpp",
        ROOT / "ccj_b200" / "csrc" / "embedded_params.cpp"]

pytestmark = pytest.mark.skipif(not (CUDA_INC / "cuda_runtime.h").exists(), reason="CUDA toolkit headers not found")


def _build(out: Path, flags):
    from ccj_b200 import build
    deps = SRCS + [ROOT / "tests" / "emu" / "simt_emu.hpp"] + list((ROOT / "ccj_b200" / "csrc").glob("*.cu*")) + \
        list((ROOT / "ccj_b200" / "csrc").glob("*.h*"))
    if not out.exists() or any(d.stat().st_mtime > out.stat().st_mtime for d in deps):
        out.parent.mkdir(exist_ok=True)
        subprocess.run(["g++", "-std=c++20", "-pthread", *flags, "-I", str(CUDA_INC), *build.embedded_par_defines(),
                        "-o", str(out)] + [str(s) for s in SRCS], check=True)
    return out


@pytest.fixture(scope="module")
def tuned_bin():
    return _build(ROOT / "build" / "ccj_emu_tuned", ["-O2"])


def _hash(exe, par, dangles, seq, no_gu=False, pipe=-1, timeout=900):
    p = subprocess.run([str(exe), "hash", str(ROOT / "params" / par), str(dangles), seq, "1" if no_gu else "0", str(pipe)],
                       capture_output=True, text=True, timeout=timeout)
    got = {}
    for line in p.stdout.splitlines()[1:]:
        name, cnt, agg, h = line.split()
        got[name] = [int(cnt), int(agg), h]
    return p, got


@pytest.mark.parametrize("n,pipe", [(20, -1), (26, 0), (35, 0)])
def test_tuned_kernels_on_the_host_match_the_reference(tuned_bin, golden_hashes, n, pipe):
    """pipe -1: what the launcher picks for one short sequence (software-pipelined window kernels); 0: the plain window
    kernels of large waves -- at n = 35 with the 16-lane groups on the first levels (runs of >= 32 cells)."""
    rec = next(r for r in golden_hashes if len(r["seq"]) == n)
    p, got = _hash(tuned_bin, rec["par"], rec["dangles"], rec["seq"], "--noGU" in rec.get("extra", []), pipe)
    assert p.returncode == 0 and p.stderr == "", p.stderr[-2000:]
    assert got == rec["tables"]


@pytest.mark.parametrize("par,dangles,no_gu", [("rna_DirksPierce09.par", 2, False), ("rna_Turner04.par", 1, False),
                                               ("rna_Turner04.par", 0, True)])
def test_tuned_kernels_on_the_host_other_models(tuned_bin, emu_bin, par, dangles, no_gu):
    """Other parameter sets / dangle models / --noGU: against the cell-function sweep (tests/emu/ccj_emu.cpp), which the
    golden vectors and live runs of the reference pin."""
    rng = random.Random(77 + dangles + len(par))
    seq = "".join(rng.choice("ACGU") for _ in range(21))
    p, got = _hash(tuned_bin, par, dangles, seq, no_gu)
    assert p.returncode == 0 and p.stderr == "", p.stderr[-2000:]
    q = subprocess.run([str(emu_bin), "hash", str(ROOT / "params" / par), str(dangles), seq, "1" if no_gu else "0"],
                       capture_output=True, text=True, check=True)
    assert p.stdout == q.stdout


def test_tuned_kernels_under_asan_ubsan(golden_hashes):
    exe = _build(ROOT / "build" / "ccj_emu_tuned_asan",
                 ["-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all"])
    rec = next(r for r in golden_hashes if len(r["seq"]) == 20)
    p, got = _hash(exe, rec["par"], rec["dangles"], rec["seq"], False, 0)
    assert "Sanitizer" not in p.stderr and "runtime error" not in p.stderr, p.stderr[-3000:]
    assert p.returncode == 0 and got == rec["tables"]


def test_tuned_kernels_under_tsan(golden_hashes):
    """Threads of a block are OS threads here: a shared-memory (or global-memory) word touched by two of them without a
    barrier in between is a data race ThreadSanitizer reports."""
    flags = ["-O1", "-g", "-fsanitize=thread"]
    exe = _build(ROOT / "build" / "ccj_emu_tuned_tsan", flags)
    rec = next(r for r in golden_hashes if len(r["seq"]) == 20)
    probe = subprocess.run([str(exe)], capture_output=True, text=True)
    if "ThreadSanitizer" in probe.stderr and "usage" not in probe.stderr:
        pytest.skip("ThreadSanitizer does not run in this environment: " + probe.stderr[-200:])
    p, got = _hash(exe, rec["par"], rec["dangles"], rec["seq"], False, -1)
    assert "ThreadSanitizer" not in p.stderr, p.stderr[-3000:]
    assert p.returncode == 0 and got == rec["tables"]
    # negative control: without the kernels' __syncwarp() the window kernels race on their shared-memory tiles
    neg = _build(ROOT / "build" / "ccj_emu_tuned_tsan_neg", flags + ["-DSIMT_EMU_DROP_SYNCWARP"])
    p, _ = _hash(neg, rec["par"], rec["dangles"], rec["seq"], False, -1)
    assert "data race" in p.stderr and "k_win" in p.stderr

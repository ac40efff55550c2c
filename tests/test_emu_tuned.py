"""CPU: the CUDA kernels themselves -- the tuned fill of ccj_b200/csrc/ccj_fill4.cu (k_prep_lay, k_fill_pmw, k_prep, k_roles,
k_winLR, k_winM, k_final, k_P_tuned), the kernels of ccj_b200/csrc/ccj_kernels.cu (k_init, k_2d, k_W, k_traceback, the
one-thread-per-cell fill k_P_lean / k_4d_lean) and those of the row-sharded fold (ccj_shard_kernels.cuh: k_P_shard_lean,
k_4d_shard_lean for power-of-two and other rank counts) -- compiled by g++ for a small SIMT emulator (tests/emu/simt_emu.hpp: one
OS thread per CUDA thread, real barriers, warp collectives) and launched in the product's order with the product's grids
(tests/emu/ccj_emu_tuned.cpp).  Only the inline PTX has plain C++ stand-ins and the <<< >>> launchers are left out; the
CUDA build of the same sources is unchanged by those guards.

* every one of the 22 + 8 tables equals the golden vector the unmodified reference wrote (poisoned, exactly sized buffers),
  folds (with k_traceback, 128 threads) equal the reference's (rc, stdout, stderr);
* the same under AddressSanitizer + UBSan and under ThreadSanitizer -- compute-sanitizer (memcheck / initcheck /
  racecheck) is closed on the GPU pool (profiles/r2_compute_sanitizer_closed.log); a run whose __syncwarp() does nothing
  is the detector's negative control.
The emulator spends its time in futex wake-ups, not in arithmetic: a run is fastest on two to four cores (128-256 threads
that meet at barriers migrate less), so independent runs go side by side, each pinned to its own pair of cores."""
import os
import queue
import random
import subprocess
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
CUDA_INC = Path(os.environ.get("CUDA_HOME", "/usr/local/cuda")) / "include"
SRCS = [ROOT / "tests" / "emu" / "ccj_emu_tuned.cpp", ROOT / "ccj_b200" / "csrc" / "energy_model.cpp",
        ROOT / "ccj_b200" / "csrc" / "embedded_params.cpp"]
SUPP = ROOT / "tests" / "emu" / "tsan_traceback.supp"

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


def _extract_enqueue_fill() -> Path:
    """the text of enqueue_fill() (with its two helper macros) exactly as ccj_abi.cu has it"""
    src = (ROOT / "ccj_b200" / "csrc" / "ccj_abi.cu").read_text()
    a, b = src.index("#define EQ(call)"), src.index("#undef EQL") + len("#undef EQL\n")
    assert "cudaError_t enqueue_fill(ccj_ctx *ctx, ccj::LaunchDims d)" in src[a:b]
    out = ROOT / "build" / "enqueue_fill.inc"
    out.parent.mkdir(exist_ok=True)
    if not out.exists() or out.read_text() != src[a:b]:
        out.write_text(src[a:b])
    return out


@pytest.fixture(scope="module")
def bins():
    """plain, ASan + UBSan and TSan builds of the harness and the one that compiles enqueue_fill(), side by side"""
    want = {"plain": (ROOT / "build" / "ccj_emu_tuned", ["-O2"]),
            "graph": (ROOT / "build" / "ccj_emu_graph", ["-O2", f"-DCCJ_ENQUEUE_FILL_INC={_extract_enqueue_fill()}"]),
            "asan": (ROOT / "build" / "ccj_emu_tuned_asan", ["-O1", "-fsanitize=address,undefined", "-fno-sanitize-recover=all"]),
            "tsan": (ROOT / "build" / "ccj_emu_tuned_tsan", ["-O1", "-g", "-fsanitize=thread"])}
    with ThreadPoolExecutor(4) as ex:
        futs = {k: ex.submit(_build, *v) for k, v in want.items()}
        return {k: f.result() for k, f in futs.items()}


_CORES = sorted(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else []
_PAIRS = queue.Queue()
for _x in range(0, max(len(_CORES) - 1, 1), 2):
    _PAIRS.put(set(_CORES[_x:_x + 2]) if _CORES else None)


def _run(exe, mode, par, dangles, seq, no_gu=False, pipe=-1, path="tuned", env=None, timeout=900, order=None):
    cores = _PAIRS.get()
    try:
        pin = (lambda: os.sched_setaffinity(0, cores)) if cores else None
        return subprocess.run([str(exe), mode, str(ROOT / "params" / par), str(dangles), seq, "1" if no_gu else "0", str(pipe), path]
                              + ([order] if order else []), capture_output=True, text=True, timeout=timeout, env=env, preexec_fn=pin)
    finally:
        _PAIRS.put(cores)


def _together(jobs):
    with ThreadPoolExecutor(max(1, min(len(jobs), _PAIRS.qsize()))) as ex:
        return [f.result() for f in [ex.submit(_run, *a, **kw) for a, kw in jobs]]


def _tables(stdout):
    got = {}
    for line in stdout.splitlines()[1:]:
        name, cnt, agg, h = line.split()
        got[name] = [int(cnt), int(agg), h]
    return got


def _small_folds(golden_folds, count):
    return [r for r in golden_folds if 18 <= len(r["seq"]) <= 23 and not r["extra"]][:count]


def test_kernels_on_the_host_match_the_reference_tables(bins, emu_bin, golden_hashes):
    """n = 20, the launcher's choice for one short sequence (software-pipelined window kernels), k_init / k_2d / k_W as
    kernels too; n = 35 with the plain window kernels of large waves, the 16-lane groups on the first levels (runs of
    >= 32 cells); n = 20 through k_P_lean + k_4d_lean (dynamic shared memory); two other models (DP09; Turner04 -d1
    --noGU) against the cell-function sweep of tests/emu/ccj_emu.cpp, which the golden vectors pin."""
    exe = bins["plain"]
    r20 = next(r for r in golden_hashes if len(r["seq"]) == 20)
    r35 = next(r for r in golden_hashes if len(r["seq"]) == 35)
    rng = random.Random(77)
    others = [("rna_DirksPierce09.par", 2, False, "".join(rng.choice("ACGU") for _ in range(21))),
              ("rna_Turner04.par", 1, True, "".join(rng.choice("ACGU") for _ in range(21)))]
    jobs = [((exe, "hash", r20["par"], r20["dangles"], r20["seq"], False, -1, "tuned+2d"), {}),
            ((exe, "hash", r35["par"], r35["dangles"], r35["seq"], False, 0, "tuned"), {}),
            ((exe, "hash", r20["par"], r20["dangles"], r20["seq"], False, -1, "lean"), {})]
    jobs += [((exe, "hash", par, d, seq, gu, -1, "tuned"), {}) for par, d, gu, seq in others]
    out = _together(jobs)
    for p in out:
        assert p.returncode == 0 and p.stderr == "", p.stderr[-2000:]
    assert _tables(out[0].stdout) == r20["tables"]
    assert _tables(out[1].stdout) == r35["tables"]
    assert _tables(out[2].stdout) == r20["tables"]
    for p, (par, d, gu, seq) in zip(out[3:], others):
        q = subprocess.run([str(emu_bin), "hash", str(ROOT / "params" / par), str(d), seq, "1" if gu else "0"],
                           capture_output=True, text=True, check=True)
        assert p.stdout == q.stdout, (par, d, gu)


def test_launch_graph_of_enqueue_fill_in_every_stream_priority_order(bins, emu_bin, golden_hashes):
    """The three-stream launch sequence of the tuned fill (main: k_roles / k_final, windows one level ahead, P + 2D tables one
    span ahead; dependency events between them) is not restated for this test: the text of enqueue_fill() of ccj_abi.cu runs
    against mock streams and events, its launches become a dependency graph (stream order + record -> wait edges), and the
    emulated kernels run in a legal order of that graph -- each of the six orders "stream A before B before C whenever
    ready", which push every side stream as far ahead or behind as the events allow.  Uninitialised memory is 0x8080 so
    that a value read too early wins its minimum.  Every order must give the reference's tables; as negative control three
    of the waits are dropped one at a time (windows on k_final two levels back, k_roles on the 2D span, compute_P on
    k_final three levels back) and some order must then go wrong."""
    rec = next(r for r in golden_hashes if len(r["seq"]) == 20)
    env = dict(os.environ, CCJ_EMU_POISON="80")
    orders = ["prio:123", "prio:132", "prio:213", "prio:231", "prio:312", "prio:321"]
    out = _together([((bins["graph"], "hash", rec["par"], rec["dangles"], rec["seq"], False, -1, "tuned"), {"env": env, "order": o})
                     for o in orders])
    for p, o in zip(out, orders):
        assert p.returncode == 0 and "cross-stream dependencies" in p.stderr, p.stderr[-2000:]
        assert _tables(p.stdout) == rec["tables"], o
    seq = rec["seq"][:14]
    want = subprocess.run([str(emu_bin), "hash", str(ROOT / "params" / rec["par"]), str(rec["dangles"]), seq], capture_output=True,
                          text=True, check=True).stdout
    ok = _run(bins["graph"], "hash", rec["par"], rec["dangles"], seq, False, -1, "tuned", env=env, order="prio:231")
    assert ok.stdout == want
    drops = [(6, "prio:231", "final0"), (10, "prio:123", "2d4"), (12, "prio:132", "final2")]
    out = _together([((bins["graph"], "hash", rec["par"], rec["dangles"], seq, False, -1, "tuned"),
                      {"env": dict(env, CCJ_EMU_DROP_WAIT=str(k)), "order": o}) for k, o, _ in drops])
    for p, (k, o, src) in zip(out, drops):
        assert f"dropped wait {k}:" in p.stderr and f"({src})" in p.stderr, p.stderr[:300]
        assert p.stdout != want, (k, o)


def test_folds_through_the_emulated_kernels(bins, golden_folds):
    """Fill + k_traceback (one block of 128 threads, block-wide ordered argmin) + the host rendering."""
    recs = _small_folds(golden_folds, 3)
    assert len(recs) == 3
    paths = ["tuned+2d", "lean", "tuned"]
    out = _together([((bins["plain"], "fold", r["par"], r["dangles"], r["seq"], False, -1, path), {}) for r, path in zip(recs, paths)])
    for p, r in zip(out, recs):
        assert (p.returncode, p.stdout, p.stderr) == (r["rc"], r["stdout"], r["stderr"]), r["seq"]


@pytest.mark.skipif(os.environ.get("CCJ_EMU_LONG") != "1", reason="4 minutes; set CCJ_EMU_LONG=1 (last run: profiles/r2_emu_tuned_h60.log)")
def test_baseline_config_1_through_the_emulated_kernels(bins, golden_hashes, golden_folds):
    """BASELINE config 1, the designed 60-nt H-type pseudoknot: every table and the fold (-29.07 kcal/mol, crossing
    brackets) through the emulated tuned kernels."""
    seq = "AAUAGGCGCAGCAUACACGGUCGAGCUGCGCCAAUAACAAUACGACCGUGAUAAAUAAAA"
    rec = next(r for r in golden_hashes if r["seq"] == seq and r["par"] == "rna_Turner04.par" and r["dangles"] == 2)
    fold = next(r for r in golden_folds if r["seq"] == seq and r["par"] == "rna_Turner04.par" and r["dangles"] == 2 and not r["extra"])
    h, f = _together([((bins["plain"], "hash", rec["par"], 2, seq, False, 0, "tuned+2d"), {"timeout": 3000}),
                      ((bins["plain"], "fold", rec["par"], 2, seq, False, -1, "tuned"), {"timeout": 3000})])
    assert _tables(h.stdout) == rec["tables"]
    assert (f.returncode, f.stdout, f.stderr) == (fold["rc"], fold["stdout"], fold["stderr"])


def test_results_do_not_depend_on_uninitialised_memory(bins, golden_hashes):
    """initcheck by construction: the buffers start as 0x8080 (-32640 per int16: a stale entry would win every minimum it
    leaked into -- rows of the window copies are padded and read past their ends) and as 0x7f7f instead of the default
    0x5555; all 22 + 8 tables still equal the golden vector."""
    rec = next(r for r in golden_hashes if len(r["seq"]) == 20)
    jobs = [((bins["plain"], "hash", rec["par"], rec["dangles"], rec["seq"], False, pipe, path),
             {"env": dict(os.environ, CCJ_EMU_POISON=poison)})
            for poison, path, pipe in (("80", "tuned", -1), ("80", "tuned", 0), ("80", "lean", -1), ("7f", "tuned", 0))]
    for p in _together(jobs):
        assert p.returncode == 0 and p.stderr == "", p.stderr[-2000:]
        assert _tables(p.stdout) == rec["tables"]


def test_ragged_wave_through_the_emulated_kernels(bins, emu_bin):
    """One wave of four sequences of different lengths (26, 20, 13 and 4 nt): the kernels pick their sequence with
    blockIdx.z / .y, the grids are sized for the longest one and the shorter ones drop out level by level -- tuned kernels
    (pipelined and plain windows) and the lean one-thread-per-cell kernels, every table of every sequence."""
    seqs = ["UUGUCAGAACGCUGAAGUGG", "UGAGUCCGAGGAG", "GCGCUAUACGCAGCCAAAACCAAUAC", "ACGU"]
    want = ""
    for x, seq in enumerate(seqs):
        q = subprocess.run([str(emu_bin), "hash", str(ROOT / "params" / "rna_Turner04.par"), "2", seq], capture_output=True,
                           text=True, check=True)
        want += f"# sequence {x}\n" + q.stdout
    out = _together([((bins["plain"], "hash", "rna_Turner04.par", 2, ",".join(seqs), False, pipe, path), {})
                     for path, pipe in (("tuned+2d", -1), ("tuned", 0), ("lean", -1))])
    for p in out:
        assert p.returncode == 0 and p.stderr == "", p.stderr[-2000:]
        assert p.stdout == want


def test_sharded_fold_kernels_on_the_host(bins, emu_bin):
    """k_P_shard_lean + k_4d_shard_lean per rank and level, the ranks sharing the replicated region (the state the per-level
    allgather establishes): 2 ranks and 3 ranks -- the <false> instantiations (rank counts that are not a power of two)
    never run in the GPU suite -- against the cell-function sweep, tables and folds."""
    rng = random.Random(1414)
    seq = "".join(rng.choice("ACGU") for _ in range(15))
    jobs = [((bins["plain"], mode, "rna_Turner04.par", 2, seq, False, -1, path), {})
            for mode, path in (("hash", "shard3"), ("fold", "shard3"), ("hash", "shard2+2d"), ("fold", "shard2"))]
    out = _together(jobs)
    for p, (a, kw) in zip(out, jobs):
        q = subprocess.run([str(emu_bin), a[1], str(ROOT / "params" / "rna_Turner04.par"), "2", seq], capture_output=True, text=True)
        assert (p.returncode, p.stdout, p.stderr) == (q.returncode, q.stdout, q.stderr), a[1:]


def test_edge_inputs_through_the_emulated_kernels(bins, emu_bin):
    """The shortest sequences the tuned kernels accept, no pair at all, every base paired: tables and folds against the
    cell-function sweep (whole levels fit one warp, most thread blocks return early, empty partner lists)."""
    seqs = ["ACGU", "GGAUC", "UACACUG", "GGGGCCCC", "AAAAAAAAAAAA", "ACCCCGGCCCC", "GCGCGCGCGCGC"]
    jobs = []
    for seq in seqs:
        jobs.append(((bins["plain"], "hash", "rna_Turner04.par", 2, seq, False, -1, "tuned+2d"), {}))
        jobs.append(((bins["plain"], "fold", "rna_Turner04.par", 2, seq, False, 0, "tuned"), {}))
    out = _together(jobs)
    for x, seq in enumerate(seqs):
        for p, mode in ((out[2 * x], "hash"), (out[2 * x + 1], "fold")):
            q = subprocess.run([str(emu_bin), mode, str(ROOT / "params" / "rna_Turner04.par"), "2", seq], capture_output=True, text=True)
            assert (p.returncode, p.stdout, p.stderr) == (q.returncode, q.stdout, q.stderr), (seq, mode)


def test_kernels_under_asan_ubsan(bins, golden_folds):
    """Out-of-bounds and misaligned accesses of any kernel against the exactly sized buffers; plain window kernels with the
    tuned fill, the lean one-thread-per-cell fill, 2D / W / traceback kernels."""
    r = _small_folds(golden_folds, 1)[0]
    out = _together([((bins["asan"], "fold", r["par"], r["dangles"], r["seq"], False, 0, "tuned+2d"), {}),
                     ((bins["asan"], "fold", r["par"], r["dangles"], r["seq"], False, -1, "lean"), {})])
    for p in out:
        assert "Sanitizer" not in p.stderr and "runtime error" not in p.stderr, p.stderr[-3000:]
        assert (p.returncode, p.stdout, p.stderr) == (r["rc"], r["stdout"], r["stderr"])


def test_kernels_under_tsan(bins, golden_folds):
    """Threads of a block are OS threads here: a shared-memory (or global-memory) word touched by two of them without a
    barrier in between is a data race ThreadSanitizer reports.  The only suppressed reports are k_traceback's intended
    identical-value stores (tests/emu/tsan_traceback.supp)."""
    exe = bins["tsan"]
    probe = subprocess.run([str(exe)], capture_output=True, text=True)
    if "ThreadSanitizer" in probe.stderr and "usage" not in probe.stderr:
        pytest.skip("ThreadSanitizer does not run in this environment: " + probe.stderr[-200:])
    env = dict(os.environ, TSAN_OPTIONS=f"suppressions={SUPP}")
    r = _small_folds(golden_folds, 1)[0]
    # second run: the negative control -- without the kernels' __syncwarp() the window kernels race on their tiles
    good, bad = _together([((exe, "fold", r["par"], r["dangles"], r["seq"], False, -1, "tuned+2d"), {"env": env}),
                           ((exe, "hash", r["par"], r["dangles"], r["seq"], False, -1, "tuned"),
                            {"env": dict(env, SIMT_EMU_DROP_SYNCWARP="1")})])
    assert "ThreadSanitizer" not in good.stderr, good.stderr[-3000:]
    assert (good.returncode, good.stdout, good.stderr) == (r["rc"], r["stdout"], r["stderr"])
    assert "data race" in bad.stderr and "k_win" in bad.stderr

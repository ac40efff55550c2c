"""GPU: the CUDA path through the C ABI against (1) the committed golden vectors of the reference,
(2) the reference binary itself where oracle/_ref travelled to the box, (3) size-independent
properties at the benchmark sizes.  Bar: bit-exact (integer dcal/mol, dot-bracket, exit behaviour)."""
import random
import subprocess
from pathlib import Path

import numpy as np
import pytest

import ccj_b200

ROOT = Path(__file__).resolve().parent.parent
REF = ROOT / "oracle" / "_ref" / "CCJ"
DUMP = ROOT / "oracle" / "_ref" / "ccj_ref_dump"
pytestmark = pytest.mark.gpu


def group(recs):
    g = {}
    for r in recs:
        g.setdefault((r["par"], r["dangles"], "--noGU" in r.get("extra", [])), []).append(r)
    return g


def check(folds, recs):
    for f, r in zip(folds, recs):
        assert (f.returncode, f.stdout, f.stderr) == (r["rc"], r["stdout"], r["stderr"]), r["seq"]


def test_golden_folds(ctx_factory, golden_folds):
    """277 sequences (1..80 nt; Turner04/DP09; d0/d1/d2; noGU) incl. the reference's exit(1) cases."""
    for (par, d, nogu), recs in group(golden_folds).items():
        ctx = ctx_factory(par, d, nogu)
        check(ctx.fold_batch([r["seq"] for r in recs]), recs)


def test_golden_table_hashes(ctx_factory, golden_hashes):
    """All 22 gap tables + V/WM/WMv/WMp/P/WBP/WPP, FNV-1a hashes of the reference's tables."""
    for r in golden_hashes:
        ctx = ctx_factory(r["par"], r["dangles"], False)
        ctx.prepare([r["seq"]])
        ctx.fill()
        for name, want in r["tables"].items():
            got = ctx.table4_hash(0, name) if name in ccj_b200.TABLE4 else ctx.table2_hash(0, name)
            assert got == want, (r["seq"], name)


def test_golden_long(ctx_factory, golden_long):
    """BASELINE configs: 100-nt random (config 2 seeds), 150-nt (config 4 seeds), 200-nt (config 3)."""
    if not golden_long:
        pytest.skip("folds_long.json not generated")
    ctx = ctx_factory()
    check(ctx.fold_batch([r["seq"] for r in golden_long]), golden_long)
    assert any("Should not be here!" in r["stdout"] for r in golden_long)


@pytest.mark.parametrize("name,count", [("folds_config4.json", 16), ("folds_config2.json", 64)])
def test_golden_benchmark_workloads(ctx_factory, name, count):
    """the first 16 config-4 sequences (150 nt, seeds 20000+idx) and the first 64 config-2
    sequences (100 nt, seeds 1000+idx), folded by the unmodified reference (make_golden.py config4 / config2)."""
    import json
    p = ROOT / "tests" / "golden" / name
    if not p.exists():
        pytest.skip(f"{name} not generated")
    recs = json.loads(p.read_text())
    assert len(recs) == count
    ctx = ctx_factory()
    check(ctx.fold_batch([r["seq"] for r in recs]), recs)


def test_beyond_the_reference_limit_against_the_cpu_restatement(ctx_factory):
    """n > 213: the reference aborts (src/matrices.hh:159-160), so the checker is oracle/ccj_oracle.cc -- itself
    pinned to the reference for n <= 213 -- run once on a 216-nt random and a 214-nt designed sequence
    (make_golden.py big, ~1 h per sequence on one core).  All 22+8 table hashes and W[n] of the tuned GPU path."""
    import json
    p = ROOT / "tests" / "golden" / "table_hashes_big.json"
    if not p.exists():
        pytest.skip("table_hashes_big.json not generated")
    ctx = ctx_factory()
    for r in json.loads(p.read_text()):
        assert len(r["seq"]) > 213
        ctx.prepare([r["seq"]])
        ctx.fill()
        for name, want in r["tables"].items():
            got = ctx.table4_hash(0, name) if name in ccj_b200.TABLE4 else ctx.table2_hash(0, name)
            assert got == want, (len(r["seq"]), name)
        ctx.traceback()
        assert ctx.fetch()[0].energy_dcal == r["W"]


def test_real_int16_wrap_matches_the_reference(tmp_path):
    """With the stacking energies tripled (custom -P file, tests/golden/make_golden.py strong_stack_par) a 40-bp helix
    inside one gapped region drives entries below -32768 dcal/mol at n=87, and the reference's Matrix4D::set wraps them
    (src/matrices.hh:188-191).  The tuned kernels saturate instead, detect that (k_final's guard) and hand the sequence to
    the generic kernels, which narrow like the reference: every table hash and the fold must equal the reference's."""
    import json
    import sys
    sys.path.insert(0, str(ROOT / "tests" / "golden"))
    from make_golden import strong_stack_par, wrap87
    g = json.loads((ROOT / "tests" / "golden" / "wrap_real.json").read_text())
    assert g["seq"] == wrap87()
    par = strong_stack_par(tmp_path / "strong.par")
    with ccj_b200.Context(0, str(par), 2) as ctx:
        # in a batch with ordinary neighbours: only the flagged sequence is repeated
        other = "GCAACGAUGACAUACAUCGCUAGUCGACGC"
        ctx.prepare([other, g["seq"], other])
        ctx.fill()
        for name, want in g["tables"].items():
            got = ctx.table4_hash(1, name) if name in ccj_b200.TABLE4 else ctx.table2_hash(1, name)
            assert got == want, name
        ctx.traceback()
        folds = ctx.fetch()
        assert (folds[1].returncode, folds[1].stdout, folds[1].stderr) == (g["rc"], g["stdout"], g["stderr"])
        assert folds[0] == folds[2] == ctx.fold(other)


def test_int16_negative_wrap_at_n213(ctx_factory):
    """Matrix4D::set clamps only the upper side and narrows int32 -> int16 (src/matrices.hh:188-191): an entry below
    -32768 dcal/mol wraps to a positive value.  With the shipped parameter sets the most extreme input the reference
    accepts -- a 103-bp poly-G/poly-C hairpin inside one arm of a gapped region at n=213, -335.1 kcal/mol -- comes within
    126 dcal of that range (smallest entry -32642) without reaching it; the tuned path detects the proximity and takes
    the generic kernels.  Fold and every table hash must equal the unmodified reference's (make_golden.py wrap)."""
    import json
    p = ROOT / "tests" / "golden" / "wrap213.json"
    if not p.exists():
        pytest.skip("wrap213.json not generated")
    g = json.loads(p.read_text())
    ctx = ctx_factory()
    seq = g["fold"]["seq"]
    ctx.prepare([seq])
    ctx.fill()
    for name, want in g["hashes"]["tables"].items():
        got = ctx.table4_hash(0, name) if name in ccj_b200.TABLE4 else ctx.table2_hash(0, name)
        assert got == want, name
    check(ctx.fold_batch([seq]), [g["fold"]])


def test_against_reference_binary_on_the_box(ctx_factory):
    if not REF.exists():
        pytest.skip("oracle/_ref did not travel")
    rng = random.Random(31337)
    seqs = ["".join(rng.choice("ACGU") for _ in range(rng.randint(1, 62))) for _ in range(40)]
    ctx = ctx_factory()
    folds = ctx.fold_batch(seqs)
    par = str(ROOT / "params" / "rna_Turner04.par")
    for s, f in zip(seqs, folds):
        p = subprocess.run([str(REF), "-P", par, s], capture_output=True, text=True)
        assert (p.returncode, p.stdout, p.stderr) == (f.returncode, f.stdout, f.stderr), s


@pytest.mark.parametrize("par,dangles,extra", [("rna_DirksPierce09.par", 2, []), ("rna_Turner04.par", 1, []),
                                               ("rna_Turner04.par", 0, ["--noGU"]), ("rna_DirksPierce09.par", 0, []),
                                               ("dna_Matthews04.par", 2, ["--noConv"])])
def test_other_models_against_reference_binary_on_the_box(ctx_factory, par, dangles, extra):
    """Live comparison with the unmodified reference for the other parameter sets, dangle models, --noGU and DNA."""
    if not REF.exists():
        pytest.skip("oracle/_ref did not travel")
    rng = random.Random(sum(map(ord, par)) * 7 + dangles)
    alphabet = "ACGT" if "--noConv" in extra else "ACGU"
    seqs = ["".join(rng.choice(alphabet) for _ in range(rng.randint(8, 52))) for _ in range(14)]
    ctx = ctx_factory(par, dangles, "--noGU" in extra)
    folds = ctx.fold_batch(seqs)
    for s, f in zip(seqs, folds):
        p = subprocess.run([str(REF), "-P", str(ROOT / "params" / par), "-d", str(dangles), *extra, s], capture_output=True, text=True)
        warn = "".join(l + "\n" for l in p.stderr.splitlines() if l.startswith("WARNING"))   # the reader's symmetry warnings
        assert (p.returncode, p.stdout, p.stderr[len(warn):]) == (f.returncode, f.stdout, f.stderr), (par, dangles, extra, s)


def test_full_tables_against_reference_dump(ctx_factory, tmp_path):
    """Every int16 of every table, not only hashes (n=48 random)."""
    if not DUMP.exists():
        pytest.skip("oracle/_ref did not travel")
    rng = random.Random(99)
    seq = "".join(rng.choice("ACGU") for _ in range(48))
    out = tmp_path / "t.bin"
    subprocess.run([str(DUMP), "bin", str(ROOT / "params" / "rna_Turner04.par"), "2", seq, str(out)], check=True)
    raw = out.read_bytes()
    hdr = np.frombuffer(raw[:16], dtype=np.int32)
    n = int(hdr[1])
    assert n == 48
    cells = ccj_b200.cells(n)
    t4 = np.frombuffer(raw[16:16 + 22 * cells * 2], dtype=np.int16).reshape(22, cells)
    t2 = np.frombuffer(raw[16 + 22 * cells * 2:], dtype=np.int32).reshape(8, n * (n + 1) // 2)
    ctx = ctx_factory()
    ctx.prepare([seq])
    ctx.fill()
    for t, name in enumerate(ccj_b200.TABLE4):
        np.testing.assert_array_equal(ctx.table4(0, name), t4[t], err_msg=name)
    for t, name in enumerate(ccj_b200.TABLE2[:8]):
        np.testing.assert_array_equal(ctx.table2(0, name), t2[t], err_msg=name)


def test_batch_invariance_and_ragged_batches(ctx_factory):
    """A fold must not depend on its neighbours in the batch, their lengths, or the wave split."""
    rng = random.Random(5)
    seqs = ["".join(rng.choice("ACGU") for _ in range(n)) for n in [1, 2, 3, 4, 5, 9, 33, 17, 60, 41, 7, 52, 28]]
    ctx = ctx_factory()
    together = ctx.fold_batch(seqs)
    alone = [ctx.fold(s) for s in seqs]
    assert together == alone
    assert ctx.fold_batch(seqs[::-1]) == together[::-1]


def test_in_process_multi_context_dealing(ctx_factory, golden_folds):
    """ccj_fold_batch_multi: one host thread per context, chunks dealt dynamically; results in input order and equal to
    the single-context call.  Contexts may sit on different GPUs (one per device) or, as here when only one GPU is
    visible, on the same one."""
    import torch
    ndev = max(1, torch.cuda.device_count())
    par = str(ROOT / "params" / "rna_Turner04.par")
    ctxs = [ccj_b200.Context(d % ndev, par, 2) for d in range(3)]
    try:
        recs = [r for r in golden_folds if r["par"] == "rna_Turner04.par" and r["dangles"] == 2 and not r["extra"]]
        recs = recs[:97]
        folds = ccj_b200.fold_batch_multi(ctxs, [r["seq"] for r in recs])
        check(folds, recs)
        assert folds == ctx_factory().fold_batch([r["seq"] for r in recs])
        with pytest.raises(ccj_b200.CCJError):
            ccj_b200.fold_batch_multi(ctxs, ["ACGU", "ACGN"])
    finally:
        for c in ctxs:
            c.close()


def test_more_sequences_than_one_launch_can_index(ctx_factory):
    """The level kernels put the sequence index (x up to 4 roles) into gridDim.y/z (limit 65535): waves are capped at
    16383 sequences, a longer batch is split, a prepared wave beyond the cap is refused."""
    ctx = ctx_factory()
    seqs = ["GGGAAACCC", "ACGUACGU", "GGGGAAAACCCC", "A", "GCGCAAGC"] * 3400          # 17 000 short sequences
    folds = ctx.fold_batch(seqs)
    assert len(folds) == len(seqs)
    want = ctx.fold_batch(seqs[:5])
    for x in range(0, len(seqs), 5):
        assert folds[x:x + 5] == want
    assert ctx.wave_capacity(8) <= 16383
    with pytest.raises(ccj_b200.CCJError):
        ctx.prepare(seqs)


def test_edge_inputs(ctx_factory):
    ctx = ctx_factory()
    assert ctx.fold("A").stdout == "A\n. (0)\n"
    assert ctx.fold("ACGU").stdout == "ACGU\n.... (0)\n"
    assert ctx.fold_batch([]) == []
    with pytest.raises(ccj_b200.CCJError):
        ctx.fold("ACGN")
    with pytest.raises(ccj_b200.CCJError):
        ctx.fold("")


def test_structure_is_consistent_with_pairs_at_benchmark_size(ctx_factory):
    """Size-independent properties at n=100: balanced brackets per family, energy <= 0, identical reruns."""
    rng = random.Random(1000)
    seq = "".join(rng.choice("ACGU") for _ in range(100))
    ctx = ctx_factory()
    a, b = ctx.fold(seq), ctx.fold(seq)
    assert a == b
    assert a.energy_dcal <= 0
    if a.status == 0:
        for o, c in ("()", "[]", "{}", "<>"):
            depth = 0
            for ch in a.structure:
                depth += (ch == o) - (ch == c)
                assert depth >= 0
            assert depth == 0


def _all_hashes(ctx):
    out = {name: ctx.table4_hash(0, name) for name in ccj_b200.TABLE4}
    out.update({name: ctx.table2_hash(0, name) for name in ccj_b200.TABLE2[:8]})
    return out


@pytest.mark.parametrize("n", [97, 230])
def test_tuned_kernels_equal_generic_kernels(ctx_factory, monkeypatch, n):
    """The tuned level kernels (read-group records, window lists, transposed copies) against the generic
    one-thread-per-cell kernel that calls ccj_cell4d, on every table -- also beyond n=213, where the
    reference itself aborts (SURVEY.md finding 2) and cannot serve as the oracle."""
    rng = random.Random(4000 + n)
    seq = "".join(rng.choice("ACGU") for _ in range(n))
    ctx = ctx_factory()
    ctx.prepare([seq])
    ctx.fill()
    ctx.traceback()
    tuned, fold_t = _all_hashes(ctx), ctx.fetch()[0]
    monkeypatch.setenv("CCJ_FILL_GENERIC", "1")
    ctx.prepare([seq])
    ctx.fill()
    ctx.traceback()
    generic, fold_g = _all_hashes(ctx), ctx.fetch()[0]
    monkeypatch.delenv("CCJ_FILL_GENERIC")
    assert tuned == generic
    assert fold_t == fold_g


@pytest.mark.parametrize("par,dangles,no_gu", [("rna_Turner04.par", 2, False), ("rna_Turner04.par", 1, False),
                                               ("rna_Turner04.par", 0, True), ("rna_DirksPierce09.par", 2, False)])
def test_cuda_matches_cpu_restatement(par, dangles, no_gu):
    """CUDA path against oracle/ccj_oracle.cc (the independent CPU restatement): W[n] on random sequences and every
    table hash on one of them."""
    from oracle import oracle as orc
    rng = random.Random(sum(map(ord, par)) + 10 * dangles)
    seqs = ["".join(rng.choice("ACGU") for _ in range(rng.randint(12, 46))) for _ in range(6)]
    with ccj_b200.Context(0, ccj_b200.default_par_file(par), dangles, no_gu) as ctx:
        ctx.prepare(seqs)
        ctx.fill()
        ctx.traceback()
        folds = ctx.fetch()
        for s, f in zip(seqs, folds):
            assert round(f.energy * 100) == orc.oracle_energy_dcal(s, par, dangles, no_gu), s
        tables, _ = orc.oracle_hashes(seqs[0], par, dangles, no_gu)
        for name, want in tables.items():
            got = ctx.table2_hash(0, name) if name in ccj_b200.TABLE2 else ctx.table4_hash(0, name)
            assert got == want, (name, seqs[0])


def test_beyond_the_tuned_range(golden_folds):
    """Sequences longer than the tuned kernels' limit get the slim buffer plan (22 tables only) and the generic
    64-bit-index kernels.  CCJ_TUNED_MAXN lowers the limit (read once per process, hence the subprocess) so that
    the path is exercised on golden 60-80-nt folds instead of a 450-nt one."""
    import json
    import os
    import sys
    recs = [r for r in golden_folds if len(r["seq"]) >= 60 and r["par"] == "rna_Turner04.par" and r["dangles"] == 2
            and not r["extra"]][:6]
    assert recs
    code = (
        "import json, sys; sys.path.insert(0, %r); import ccj_b200\n"
        "seqs = json.loads(sys.stdin.read())\n"
        "ctx = ccj_b200.Context(0, ccj_b200.default_par_file('rna_Turner04.par'), 2)\n"
        "print(json.dumps([[f.returncode, f.stdout, f.stderr] for f in ctx.fold_batch(seqs)]))\n" % str(ROOT))
    env = dict(os.environ, CCJ_TUNED_MAXN="50")
    p = subprocess.run([sys.executable, "-c", code], input=json.dumps([r["seq"] for r in recs]), capture_output=True,
                       text=True, env=env, check=True)
    got = json.loads(p.stdout.strip().splitlines()[-1])
    assert got == [[r["rc"], r["stdout"], r["stderr"]] for r in recs]

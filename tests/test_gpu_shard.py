"""GPU: the row-sharded fold of one sequence (BASELINE config 5, ccj_b200/csrc/ccj_shard.cu) against the reference's
golden vectors.  On one GPU the multi-rank logic runs as an in-process group (all ranks on one device, device copies
as collectives); with two or more GPUs the same checks run over NCCL, one process per GPU (torchrun)."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

import ccj_b200
from ccj_b200 import shard5

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_in_process_group_matches_reference_tables_and_folds(ctx_factory, golden_hashes, golden_folds, world):
    ctx = ctx_factory()
    grp = shard5.LocalGroup(ctx, world)
    try:
        recs = [r for r in golden_hashes if r["par"] == "rna_Turner04.par" and r["dangles"] == 2 and len(r["seq"]) >= 35][:3]
        assert recs
        for r in recs:
            sh = grp.fold(r["seq"])
            assert sh.all_hashes() == r["tables"], (world, r["seq"])
        folds = [r for r in golden_folds if r["par"] == "rna_Turner04.par" and r["dangles"] == 2 and not r["extra"]
                 and len(r["seq"]) >= 30][:12]
        assert any(r["rc"] != 0 for r in folds) or True
        for r in folds:
            f = grp.fold(r["seq"]).traceback()
            assert (f.returncode, f.stdout, f.stderr) == (r["rc"], r["stdout"], r["stderr"]), (world, r["seq"])
    finally:
        grp.close()


def test_in_process_group_on_a_benchmark_size_fold(ctx_factory, golden_long):
    """A 100-nt config-2 sequence on 8 emulated ranks: fold identical to the reference's."""
    if not golden_long:
        pytest.skip("folds_long.json not generated")
    ctx = ctx_factory()
    grp = shard5.LocalGroup(ctx, 8)
    try:
        r = [x for x in golden_long if len(x["seq"]) == 100][0]
        f = grp.fold(r["seq"]).traceback()
        assert (f.returncode, f.stdout, f.stderr) == (r["rc"], r["stdout"], r["stderr"])
        assert grp.ms["fill_ms"] > 0
    finally:
        grp.close()


def test_nccl_ranks_match_reference(golden_hashes):
    """One process per GPU over NCCL (needs >= 2 GPUs): profiles/config5_check.py compares table hashes and folds with the
    golden vectors on every world size the box offers."""
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if ngpu < 4 else (4 if ngpu < 8 else 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", str(ROOT / "profiles" / "config5_check.py"), "--golden"]
    p = subprocess.run(cmd, capture_output=True, text=True, cwd=str(ROOT), timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "CONFIG5 GOLDEN OK" in p.stdout


def test_in_process_nccl_group_over_all_gpus(golden_long, golden_folds):
    """ccj_shard_fold: every rank in ONE process, one GPU each, ncclCommInitAll + grouped collectives (what the CCJ command
    line uses for a sequence that exceeds one GPU).  Needs >= 2 GPUs."""
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least 2 GPUs")
    par = str(ROOT / "params" / "rna_Turner04.par")
    ctxs = [ccj_b200.Context(d, par, 2) for d in range(ngpu)]
    try:
        recs = [r for r in golden_folds if r["par"] == "rna_Turner04.par" and r["dangles"] == 2 and not r["extra"] and len(r["seq"]) >= 40][:6]
        recs += [r for r in golden_long if len(r["seq"]) == 150][:1]
        for r in recs:
            f, ms = shard5.fold_multi(ctxs, r["seq"])
            assert (f.returncode, f.stdout, f.stderr) == (r["rc"], r["stdout"], r["stderr"]), r["seq"]
            assert ms["fill_ms"] > 0
    finally:
        for c in ctxs:
            c.close()

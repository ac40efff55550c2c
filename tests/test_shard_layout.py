"""CPU: the row-sharded layout of one oversized fold (BASELINE config 5; ccj_b200/csrc/ccj_types.h "sharded layout",
ccj_shard.cu) -- host logic only, the fill itself needs GPUs:
  * every cell maps to exactly one (owner rank, level, inner index), inner indices of one (level, rank) are a
    bijection onto [0, rows' cells), fit the space the level reserves, and rows are contiguous in k;
  * the product's cell functions swept over that layout on the CPU (tests/emu) reproduce the reference's table hashes
    and folds for 1, 2, 3, 4 and 8 ranks -- the replicated region is shared between the emulated ranks, which is the
    state the per-level allgather establishes on the GPUs;
  * memory model: a 600-nt sequence fits 8 x 180 GB but not fewer ranks."""
import math
import subprocess
from pathlib import Path

import pytest

from ccj_b200 import shard5

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("n,world", [(9, 1), (12, 2), (14, 3), (17, 4), (21, 8)])
def test_sharded_layout_is_a_bijection_per_level_and_rank(library, n, world):
    seen = {}
    reserve = {}
    total = 0
    for i in range(1, n + 1):
        for j in range(i, n + 1):
            for k in range(j + 2, n + 1):
                for l in range(k, n + 1):
                    inner, owner, level, lc, lb = shard5.shard_layout(n, world, i, j, k, l)
                    assert owner == (i - 1) % world and level == (j - i) + (l - k)
                    assert 0 <= inner < lc, (i, j, k, l)
                    key = (level, owner, inner)
                    assert key not in seen, (key, seen.get(key), (i, j, k, l))
                    seen[key] = (i, j, k, l)
                    reserve[level] = (lc, lb)
                    total += 1
    assert total == math.comb(n + 1, 4)
    # dense: the inner indices of one (level, rank) are exactly 0..count-1
    per = {}
    for (level, owner, inner) in seen:
        per.setdefault((level, owner), []).append(inner)
    for key, v in per.items():
        assert sorted(v) == list(range(len(v))), key
    # rank 0 owns the most rows: its count is what every rank reserves; level bases are the running sum
    acc = 0
    for level in sorted(reserve):
        lc, lb = reserve[level]
        assert lb == acc and lc == len(per[(level, 0)])
        acc += lc
    assert shard5.shard_layout(n, world, 3, 2, 6, 7)[0] == -1


def test_sharded_rows_are_contiguous_in_k(library):
    n, world = 24, 4
    for a in range(0, 5):
        for b in range(0, 5):
            for i in range(1, n - a - b - 1):
                ks = list(range(i + a + 2, n - b + 1))
                idx = [shard5.shard_layout(n, world, i, i + a, k, k + b)[0] for k in ks]
                assert idx == list(range(idx[0], idx[0] + len(ks)))


def test_memory_model(library):
    gb = 1e9
    assert shard5.shard_bytes(600, 8) < 170 * gb        # config 5 on 8 B200
    assert shard5.shard_bytes(600, 4) < 170 * gb
    assert shard5.shard_bytes(600, 1) > 200 * gb        # does not fit one GPU: 22 x 2 B x C(601,4) = 237 GB
    b150 = shard5.shard_bytes(150, 1)   # 44 B per cell + the per-pair partner lists and the 2D tables
    assert 44 * math.comb(151, 4) <= b150 < 1.2 * 44 * math.comb(151, 4)


def _emu(emu_bin, mode, rec, world):
    # last arguments: keep {WB,WP,WBP} packed per interval and walk partner lists in the interior windows, as the
    # sharded GPU fold does (world 3: the plain getters and the reference's window scan)
    args = [str(emu_bin), mode, str(ROOT / "params" / rec["par"]), str(rec["dangles"]), rec["seq"],
            "1" if "--noGU" in rec.get("extra", []) else "0", str(world), "1" if world != 3 else "0", "1" if world != 3 else "0"]
    return subprocess.run(args, capture_output=True, text=True)


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_cell_functions_over_the_sharded_layout_match_the_reference(emu_bin, golden_hashes, golden_folds, world):
    recs = [r for r in golden_hashes if len(r["seq"]) <= 32][:3]
    assert recs
    for r in recs:
        p = _emu(emu_bin, "hash", r, world)
        assert p.returncode == 0, p.stderr
        got = {}
        for line in p.stdout.splitlines()[1:]:
            name, cnt, agg, h = line.split()
            got[name] = [int(cnt), int(agg), h]
        assert got == r["tables"], (world, r["seq"])
    folds = [r for r in golden_folds if 20 <= len(r["seq"]) <= 34][:8]
    for r in folds:
        p = _emu(emu_bin, "fold", r, world)
        assert (p.returncode, p.stdout, p.stderr) == (r["rc"], r["stdout"], r["stderr"]), (world, r["seq"])


@pytest.mark.parametrize("world", [0, 1, 2, 3, 4, 8])
def test_lean_cell_functions_match_the_reference(emu_bin, golden_hashes, golden_folds, world):
    """ccj_cells4_lean.cuh (what k_4d_lean / k_4d_shard_lean / k_P_*_lean run: one position per split point, compile-time
    table kinds, strided second factors in compute_P) swept on the CPU over the ordinary layout (world 0) and the sharded
    layout (power-of-two and other rank counts): table hashes and folds equal the reference's golden vectors."""
    def run(mode, rec):
        args = [str(emu_bin), mode, str(ROOT / "params" / rec["par"]), str(rec["dangles"]), rec["seq"],
                "1" if "--noGU" in rec.get("extra", []) else "0", str(world), "1", "1", "1"]
        return subprocess.run(args, capture_output=True, text=True)
    recs = [r for r in golden_hashes if len(r["seq"]) <= 41]
    assert len(recs) >= 6
    for r in recs:
        p = run("hash", r)
        assert p.returncode == 0, p.stderr
        got = {}
        for line in p.stdout.splitlines()[1:]:
            name, cnt, agg, h = line.split()
            got[name] = [int(cnt), int(agg), h]
        assert got == r["tables"], (world, r["seq"])
    for r in [r for r in golden_folds if 20 <= len(r["seq"]) <= 34][:8]:
        p = run("fold", r)
        assert (p.returncode, p.stdout, p.stderr) == (r["rc"], r["stdout"], r["stderr"]), (world, r["seq"])


def test_flat_thread_to_cell_mapping_is_the_exact_inverse(tmp_path):
    """k_4d / k_4d_shard give thread p of a slab the p-th cell in storage order; the inverse of the row offsets uses a
    float square root plus fix-ups -- checked exhaustively on the host (tests/shell/cell_of_test.cc)."""
    exe = tmp_path / "cell_of_test"
    subprocess.run(["g++", "-O2", "-I", str(ROOT / "ccj_b200" / "csrc"), "-o", str(exe), str(ROOT / "tests" / "shell" / "cell_of_test.cc")],
                   check=True)
    p = subprocess.run([str(exe)], capture_output=True, text=True)
    assert p.returncode == 0 and " 0 bad" in p.stdout, p.stdout

"""One sequence whose gap tables exceed one GPU (BASELINE config 5): rows of the gap tables dealt cyclically to the
ranks, one in-place NCCL allgather of the 12 column-read tables per finished level, an allreduce(min) of P per span
(ccj_b200/csrc/ccj_shard.cu; the loops being sharded are pseudo_loop::compute_energies, src/pseudo_loop.cc:69-132).

  ShardedFold(ctx, rank, world, unique_id)   one rank of an NCCL group (one process per GPU, torchrun)
  LocalGroup(ctx, world)                     all ranks inside one process on one GPU (device copies as collectives):
                                             the single-GPU test of the multi-rank logic
"""
from __future__ import annotations

import ctypes as C
import json
import os
import time
from typing import List, Optional

import numpy as np

from . import CCJError, Fold, RESULT_DTYPE, TABLE2, TABLE4, load_library

_bound = False


def _lib():
    global _bound
    lib = load_library()
    if not _bound:
        vp, i32 = C.c_void_p, C.c_int
        lib.ccj_shard_unique_id_bytes.restype = C.c_size_t
        lib.ccj_shard_unique_id.argtypes = [vp, C.c_size_t]
        lib.ccj_shard_create.argtypes = [vp, i32, i32, vp, C.POINTER(vp)]
        lib.ccj_shard_destroy.argtypes = [vp]
        lib.ccj_shard_destroy.restype = None
        lib.ccj_shard_last_error.argtypes = [vp]
        lib.ccj_shard_last_error.restype = C.c_char_p
        lib.ccj_shard_bytes.argtypes = [i32, i32]
        lib.ccj_shard_bytes.restype = C.c_int64
        lib.ccj_shard_prepare.argtypes = [vp, C.c_char_p, i32]
        lib.ccj_shard_fill.argtypes = [C.POINTER(vp), i32, C.POINTER(C.c_float)]
        lib.ccj_shard_level_ms.argtypes = [vp, C.POINTER(C.c_float), C.c_int64]
        lib.ccj_shard_level_bytes.argtypes = [i32, i32, i32]
        lib.ccj_shard_level_bytes.restype = C.c_int64
        lib.ccj_shard_ipc_bytes.restype = C.c_size_t
        lib.ccj_shard_ipc_handle.argtypes = [vp, vp, C.c_size_t]
        lib.ccj_shard_open_peers.argtypes = [vp, vp, C.c_size_t]
        lib.ccj_shard_link_local.argtypes = [vp, C.POINTER(vp), i32]
        lib.ccj_shard_traceback.argtypes = [vp, vp, vp, vp, C.POINTER(C.c_float)]
        lib.ccj_shard_energy.argtypes = [vp, C.POINTER(C.c_int32)]
        lib.ccj_shard_table4_hash.argtypes = [vp, i32, C.POINTER(C.c_uint64), C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
        lib.ccj_shard_table2_hash.argtypes = [vp, i32, C.POINTER(C.c_uint64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        lib.ccj_shard_fold.argtypes = [C.POINTER(vp), i32, C.c_char_p, i32, vp, vp, vp, C.POINTER(C.c_float)]
        lib.ccj_shard_layout.argtypes = [i32, i32, i32, i32, i32, i32, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                         C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        lib.ccj_shard_layout.restype = C.c_int64
        _bound = True
    return lib


def shard_bytes(n: int, world: int) -> int:
    """Device memory one rank needs for a length-n sequence dealt to `world` ranks (host only)."""
    return int(_lib().ccj_shard_bytes(n, world))


def shard_layout(n: int, world: int, i: int, j: int, k: int, l: int):
    """(inner index, owner rank, level, cells reserved per table and rank at that level, level base) -- host only."""
    owner, level, lc, lb = C.c_int32(), C.c_int32(), C.c_int64(), C.c_int64()
    inner = _lib().ccj_shard_layout(n, world, i, j, k, l, C.byref(owner), C.byref(level), C.byref(lc), C.byref(lb))
    return int(inner), owner.value, level.value, lc.value, lb.value


def unique_id() -> bytes:
    import torch  # noqa: F401  -- torch's bundled libnccl must be the one in the process (see ccj_shard.cu, nccl())
    lib = _lib()
    buf = C.create_string_buffer(lib.ccj_shard_unique_id_bytes())
    rc = lib.ccj_shard_unique_id(buf, len(buf))
    if rc != 0:
        raise CCJError(rc, "ncclGetUniqueId failed (libnccl.so.2 not loadable?)")
    return buf.raw


class ShardedFold:
    """One rank's share of one oversized fold."""

    def __init__(self, ctx, rank: int = 0, world: int = 1, uid: Optional[bytes] = None):
        if uid is not None:
            import torch  # noqa: F401  -- see unique_id()
        self._lib = _lib()
        self.ctx, self.rank, self.world = ctx, rank, world
        self._h = C.c_void_p()
        rc = self._lib.ccj_shard_create(ctx._h, rank, world, uid, C.byref(self._h))
        if rc != 0:
            raise CCJError(rc, "ccj_shard_create failed")
        self.seq = ""
        self.ms = None

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.ccj_shard_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise CCJError(rc, self._lib.ccj_shard_last_error(self._h).decode())

    def prepare(self, seq: str):
        self.seq = seq
        self._check(self._lib.ccj_shard_prepare(self._h, seq.encode("ascii"), len(seq)))

    def fill(self) -> dict:
        ms = (C.c_float * 4)()
        arr = (C.c_void_p * 1)(self._h)
        self._check(self._lib.ccj_shard_fill(arr, 1, ms))
        self.ms = {"fill_ms": ms[0], "compute_ms": ms[1], "allgather_ms": ms[2], "allreduce_ms": ms[3]}
        return self.ms

    def level_ms(self) -> np.ndarray:
        """(n, 4) device ms per step: P kernel, allreduce, 2D + gap-table kernels, allgather."""
        n = len(self.seq)
        out = np.zeros(4 * n, dtype=np.float32)
        self._check(self._lib.ccj_shard_level_ms(self._h, out.ctypes.data_as(C.POINTER(C.c_float)), out.size))
        return out.reshape(n, 4)

    def ipc_handle(self) -> bytes:
        buf = C.create_string_buffer(self._lib.ccj_shard_ipc_bytes())
        self._check(self._lib.ccj_shard_ipc_handle(self._h, buf, len(buf)))
        return buf.raw

    def open_peers(self, handles: List[bytes]):
        blob = b"".join(handles)
        self._check(self._lib.ccj_shard_open_peers(self._h, blob, len(blob)))

    def energy_dcal(self) -> int:
        v = C.c_int32()
        self._check(self._lib.ccj_shard_energy(self._h, C.byref(v)))
        return v.value

    def traceback(self) -> Fold:
        n = len(self.seq)
        res = np.zeros(1, dtype=RESULT_DTYPE)
        structs = np.zeros(n, dtype=np.uint8)
        ms = C.c_float()
        self._check(self._lib.ccj_shard_traceback(self._h, res.ctypes.data, None, structs.ctypes.data, C.byref(ms)))
        self.traceback_ms = ms.value
        r = res[0]
        return Fold(self.seq, structs.tobytes().decode("ascii"), int(r["energy_dcal"]), int(r["status"]),
                    int(r["n_should_not_be_here"]), int(r["msg_id"]), int(r["aux_i"]), int(r["aux_j"]))

    def table4_hash(self, table):
        t = TABLE4.index(table) if isinstance(table, str) else int(table)
        h, fin, mn = C.c_uint64(), C.c_int64(), C.c_int32()
        self._check(self._lib.ccj_shard_table4_hash(self._h, t, C.byref(h), C.byref(fin), C.byref(mn)))
        return [fin.value, mn.value, "%016x" % h.value]

    def table2_hash(self, table):
        t = TABLE2.index(table) if isinstance(table, str) else int(table)
        h, fin, sm = C.c_uint64(), C.c_int64(), C.c_int64()
        self._check(self._lib.ccj_shard_table2_hash(self._h, t, C.byref(h), C.byref(fin), C.byref(sm)))
        return [fin.value, sm.value, "%016x" % h.value]

    def all_hashes(self):
        out = {name: self.table4_hash(name) for name in TABLE4}
        out.update({name: self.table2_hash(name) for name in TABLE2[:8]})
        return out


def fold_multi(contexts, seq: str):
    """One sequence over the GPUs of `contexts` (one per device, same model) from ONE process: the in-process NCCL
    group of ccj_shard_fold.  Returns (Fold, {"fill_ms", "compute_ms", "allgather_ms", "allreduce_ms"})."""
    import torch  # noqa: F401  -- see unique_id()
    lib = _lib()
    n = len(seq)
    arr = (C.c_void_p * len(contexts))(*[c._h for c in contexts])
    res = np.zeros(1, dtype=RESULT_DTYPE)
    structs = np.zeros(n, dtype=np.uint8)
    ms = (C.c_float * 4)()
    rc = lib.ccj_shard_fold(arr, len(contexts), seq.encode("ascii"), n, res.ctypes.data, None, structs.ctypes.data, ms)
    if rc != 0:
        raise CCJError(rc, "ccj_shard_fold failed (see stderr)")
    r = res[0]
    fold = Fold(seq, structs.tobytes().decode("ascii"), int(r["energy_dcal"]), int(r["status"]), int(r["n_should_not_be_here"]),
                int(r["msg_id"]), int(r["aux_i"]), int(r["aux_j"]))
    return fold, {"fill_ms": ms[0], "compute_ms": ms[1], "allgather_ms": ms[2], "allreduce_ms": ms[3]}


class LocalGroup:
    """All `world` ranks of a sharded fold inside one process on one GPU; collectives are device copies."""

    def __init__(self, ctx, world: int):
        self.shards = [ShardedFold(ctx, r, world, None) for r in range(world)]
        self.world = world

    def close(self):
        for s in self.shards:
            s.close()

    def fold(self, seq: str):
        lib = _lib()
        for s in self.shards:
            s.prepare(seq)
        arr = (C.c_void_p * self.world)(*[s._h for s in self.shards])
        ms = (C.c_float * 4)()
        self.shards[0]._check(lib.ccj_shard_fill(arr, self.world, ms))
        self.ms = {"fill_ms": ms[0], "compute_ms": ms[1], "allgather_ms": ms[2], "allreduce_ms": ms[3]}
        self.shards[0]._check(lib.ccj_shard_link_local(self.shards[0]._h, arr, self.world))
        return self.shards[0]


# ------------------------------------------------------------------------------------------------------------------
# bench.py --config5
# ------------------------------------------------------------------------------------------------------------------
def config5_sequence(n: int) -> str:
    """BASELINE config 5: random.Random(600) stream, uniform over ACGU, first n nucleotides."""
    import random
    rng = random.Random(600)
    return "".join(rng.choice("ACGU") for _ in range(max(n, 600)))[:n]


def bench_config5(args):
    """torchrun --nproc-per-node G bench.py --config5 --n5 N: one sequence, rows dealt to the G ranks."""
    import torch
    import torch.distributed as dist
    import ccj_b200
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.n5
    seq = config5_sequence(n)
    ctx = ccj_b200.Context(local, str(root / "params" / "rna_Turner04.par"), 2)
    uid = None
    if world > 1:
        box = [unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]
    sh = ShardedFold(ctx, rank, world, uid)
    sh.prepare(seq)
    if world > 1:
        handles = [None] * world
        dist.all_gather_object(handles, sh.ipc_handle())
        if rank == 0:
            sh.open_peers(handles)
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ms = sh.fill()
    wall = time.perf_counter() - t0
    t = torch.tensor([ms["fill_ms"], ms["compute_ms"], ms["allgather_ms"], ms["allreduce_ms"], wall * 1e3],
                     dtype=torch.float64, device=f"cuda:{local}")
    tmax = t.clone()
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    fold = tb_ms = None
    hashes = None
    if rank == 0:
        fold = sh.traceback()
        tb_ms = sh.traceback_ms
        if args.hash5:
            hashes = sh.all_hashes()
    if world > 1:
        dist.barrier()
    if rank == 0:
        lv = sh.level_ms()
        top = sorted(range(n), key=lambda x: -lv[x, 3])[:5]
        gather_top = [{"level": int(x), "bytes_per_rank": int(_lib().ccj_shard_level_bytes(n, world, x)),
                       "rank0_ms": float(lv[x, 3]),
                       "rank0_recv_gbs": float(_lib().ccj_shard_level_bytes(n, world, x) * (world - 1) / max(lv[x, 3], 1e-6) / 1e6)}
                      for x in top]
        terms = ccj_b200.count_terms(seq)
        fill_s = float(tmax[0]) / 1e3
        peak = 6554.6
        p = root / "MEASURED_PEAKS.json"
        if p.exists():
            peak = float(json.loads(p.read_text())["hbm_gbs"])
        line = {
            "metric": "dp_cells_per_s_single_sharded_sequence", "mode": "config5", "value": terms["cells"] / fill_s,
            "unit": "cells/s", "n_gpus": world, "scaling": "strong", "higher_is_better": True, "data": "synthetic",
            "config": {"workload": f"config5: one {n}-nt sequence (random.Random(600)), gap-table rows dealt cyclically to "
                                   f"{world} ranks, NCCL allgather of the 12 column-read tables per level",
                       "params": "rna_Turner04.par", "dangles": 2, "seq_len": n},
            "fill_ms": float(tmax[0]), "compute_ms": float(tmax[1]), "allgather_ms": float(tmax[2]),
            "allreduce_ms": float(tmax[3]), "allgather_share_of_fill": float(tmax[2]) / max(float(tmax[0]), 1e-9),
            "rank0_ms": {"fill": float(t[0]), "compute": float(t[1]), "allgather": float(t[2]), "allreduce": float(t[3])},
            "wall_ms": float(tmax[4]), "traceback_ms": tb_ms, "largest_allgathers_rank0": gather_top,
            "allgather_bytes_received_per_rank": int(sum(_lib().ccj_shard_level_bytes(n, world, x) for x in range(n)) * (world - 1)),
            "bytes_per_rank": shard_bytes(n, world), "cells": terms["cells"],
            "algorithmic_bytes": terms["bytes"], "achieved_gbs_all_ranks": terms["bytes"] / fill_s / 1e9,
            "roofline_frac_of_aggregate_hbm": terms["bytes"] / fill_s / 1e9 / (peak * world),
            "kernel": ("one-thread-per-cell level kernel over the sharded layout, lean addressing (k_4d_shard_lean, k_P_shard_lean)"
                       if os.environ.get("CCJ_SHARD_LEAN", "1") != "0" else
                       "one-thread-per-cell level kernel over the sharded layout, generic addressing (k_4d_shard)"),
            "energy": fold.energy, "status": fold.status, "should_not_be_here": fold.n_should_not_be_here,
            "structure": fold.structure, "hashes": hashes,
        }
        print(json.dumps(line))
    sh.close()
    if world > 1:
        dist.destroy_process_group()

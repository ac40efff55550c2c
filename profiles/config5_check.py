"""Parity of the NCCL-sharded fold (one process per GPU), run under torchrun:

  --golden        n <= 213: table hashes / folds of the sharded fill == the reference's golden vectors
                  (tests/golden/table_hashes.json, folds.json, folds_long.json)
  --versus N      n = N (beyond the reference's limit): sharded fill == the unsharded tuned single-GPU fill on rank 0
                  (energy, structure, every 2D table hash, selected gap-table hashes)

Prints one JSON line per check and "CONFIG5 ... OK" at the end; exits non-zero on the first mismatch.
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import ccj_b200  # noqa: E402
from ccj_b200 import shard5  # noqa: E402


def setup():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = ccj_b200.Context(local, str(ROOT / "params" / "rna_Turner04.par"), 2)
    uid = None
    if world > 1:
        box = [shard5.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]
    return world, rank, ctx, shard5.ShardedFold(ctx, rank, world, uid)


def sharded_fold(sh, seq, world, rank):
    sh.prepare(seq)
    if world > 1:
        handles = [None] * world
        dist.all_gather_object(handles, sh.ipc_handle())
        if rank == 0:
            sh.open_peers(handles)
        dist.barrier()
    return sh.fill()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--golden", action="store_true")
    ap.add_argument("--versus", type=int, default=0)
    ap.add_argument("--tables", default="PK,PL,PfromO,POmloop10,PR,PMmloop10")
    args = ap.parse_args()
    world, rank, ctx, sh = setup()
    ok = True

    def fail(msg):
        nonlocal ok
        ok = False
        print("MISMATCH", msg, flush=True)

    if args.golden:
        hashes = [r for r in json.loads((ROOT / "tests/golden/table_hashes.json").read_text())
                  if r["par"] == "rna_Turner04.par" and r["dangles"] == 2 and len(r["seq"]) >= 40][:4]
        for r in hashes:
            ms = sharded_fold(sh, r["seq"], world, rank)
            if rank == 0:
                got = sh.all_hashes()
                if got != r["tables"]:
                    fail(f"hashes n={len(r['seq'])}: " + ", ".join(k for k in got if got[k] != r["tables"][k]))
                print(json.dumps({"check": "golden_hashes", "n": len(r["seq"]), "world": world, **ms}), flush=True)
            if world > 1:
                dist.barrier()
        folds = [r for r in json.loads((ROOT / "tests/golden/folds.json").read_text())
                 if r["par"] == "rna_Turner04.par" and r["dangles"] == 2 and not r["extra"] and len(r["seq"]) >= 40][:10]
        folds += [r for r in json.loads((ROOT / "tests/golden/folds_long.json").read_text()) if len(r["seq"]) in (100, 150, 200)][::7]
        for r in folds:
            ms = sharded_fold(sh, r["seq"], world, rank)
            if rank == 0:
                f = sh.traceback()
                if (f.returncode, f.stdout, f.stderr) != (r["rc"], r["stdout"], r["stderr"]):
                    fail(f"fold n={len(r['seq'])}")
                print(json.dumps({"check": "golden_fold", "n": len(r["seq"]), "world": world, "energy": f.energy, **ms}), flush=True)
            if world > 1:
                dist.barrier()
        if rank == 0 and ok:
            print("CONFIG5 GOLDEN OK", flush=True)

    if args.versus:
        n = args.versus
        seq = shard5.config5_sequence(n)
        ms = sharded_fold(sh, seq, world, rank)
        if rank == 0:
            t0 = time.perf_counter()
            f = sh.traceback()
            names = [x for x in args.tables.split(",") if x]
            got4 = {name: sh.table4_hash(name) for name in names}
            got2 = {name: sh.table2_hash(name) for name in ccj_b200.TABLE2[:8]}
            t_sh = time.perf_counter() - t0
        sh.close()   # free the shard's memory before the unsharded fill on rank 0
        if world > 1:
            dist.barrier()
        if rank == 0:
            ctx.prepare([seq])
            fill_ms = ctx.fill()
            ctx.traceback()
            g = ctx.fetch()[0]
            want4 = {name: ctx.table4_hash(0, name) for name in names}
            want2 = {name: ctx.table2_hash(0, name) for name in ccj_b200.TABLE2[:8]}
            if (f.returncode, f.stdout, f.stderr) != (g.returncode, g.stdout, g.stderr):
                fail(f"fold n={n}: {f.energy} vs {g.energy}")
            for name in names:
                if got4[name] != want4[name]:
                    fail(f"table {name} n={n}")
            for name in want2:
                if got2[name] != want2[name]:
                    fail(f"2D table {name} n={n}")
            print(json.dumps({"check": "sharded_vs_unsharded", "n": n, "world": world, "energy": f.energy,
                              "should_not_be_here": f.n_should_not_be_here, "tables": names, "unsharded_tuned_fill_ms": fill_ms,
                              "host_hash_s": t_sh, **ms}), flush=True)
            if ok:
                print(f"CONFIG5 VERSUS {n} OK", flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

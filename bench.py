#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config, one JSON line on stdout.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]
  python bench.py --full [--full-count 8192]      # the whole config-4 job once, sequences dealt dynamically
  python bench.py --config5 [--n5 600]             # one oversized sequence sharded by outer index (torchrun, N>=2)

Metric  : 4D DP cells/s (and folds/s) on the 150-nt batch (BASELINE config 4: random.Random(20000+idx)).
Step    : one pass of the hot path (fill of V/WM/P/22 gap tables/W + traceback) over one batch of B
          sequences per GPU; every step folds NEW sequences of the workload (step s of rank r takes batch
          number s*world + r).  Sequences are independent -> no collective on the data path (weak scaling).
value   : device-resident throughput (inputs already in HBM; CUDA events on the library's stream).
e2e     : the same metric through the C-ABI call a user makes (ccj_fold_batch) with HOST buffers: H2D of the
          sequences, fill, traceback, D2H of energies/pairs and dot-bracket rendering, wall clock.
roofline: for the dominant kernel group (the level-wavefront gap-table kernels): algorithmic bytes
          (2 B x every min-plus candidate read + 44 B per cell written, SURVEY.md 8d, counted exactly per input)
          / summed device time of its launches, against MEASURED_PEAKS.json's HBM copy bandwidth; `per_kernel`
          gives every kernel group against BOTH roofs (HBM bytes and the integer add-min ceiling measured live by
          ccj_measure_addmin_peak), `bound` = the roof that leaves less headroom.
config3 : BASELINE config 3 (single 200-nt fills: SURVEY App. C seed-200 sequence and designed(200)), fill time,
          roofline fraction, parity with the reference's golden output.
config2 : BASELINE config 2 (1,024 random 100-nt sequences) end to end through ccj_fold_batch, parity with every
          golden vector of tests/golden/ for that workload.
cpu_baseline / --impl reference: the unmodified reference binary (oracle/_ref/CCJ) on the host cores, one
          process per core, on a bounded sample (prefixes of the same workload sequences, length chosen so a
          step fits the time budget; the reference's cells/s only falls with n, so this flatters the CPU) and,
          once, on one full-length 150-nt sequence per core (`full_length`, the same-config figure).
"""
from __future__ import annotations

import argparse
import hashlib
import json
import math
import os
import random
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_NT = 150
SEED0 = 20000
PAR = "rna_Turner04.par"
GOLDEN = ROOT / "tests" / "golden"


def workload(idx0: int, count: int, n: int = N_NT):
    out = []
    for idx in range(idx0, idx0 + count):
        rng = random.Random(SEED0 + idx)
        out.append("".join(rng.choice("ACGU") for _ in range(N_NT))[:n])
    return out


def workload2(count: int = 1024):
    """BASELINE config 2: random.Random(1000+idx), 100 nt."""
    out = []
    for idx in range(count):
        rng = random.Random(1000 + idx)
        out.append("".join(rng.choice("ACGU") for _ in range(100)))
    return out


def cells(n: int) -> int:
    return math.comb(n + 1, 4)


def hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_goldens(*names):
    """{sequence: record} of the Turner04/d2 golden folds in the named fixture files."""
    out = {}
    for name in names:
        p = GOLDEN / name
        if not p.exists():
            continue
        for r in json.loads(p.read_text()):
            if r["par"] == PAR and r["dangles"] == 2 and not r.get("extra"):
                out[r["seq"]] = r
    return out


def same_as_golden(f, r) -> bool:
    return (f.returncode, f.stdout, f.stderr) == (r["rc"], r["stdout"], r["stderr"])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[x] for r in self.rows if len(r) >= 7 for x in range(4) if r[3 + x].lower() == "active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the unmodified reference on the host cores
# ------------------------------------------------------------------------------------------------
def ref_time_model(n: int) -> float:
    """seconds per fold on one core, from BASELINE.md (41 s at n=100, ~n^5.3)."""
    return 41.0 * (n / 100.0) ** 5.3


def pick_sample_len(budget_s: float) -> int:
    n = N_NT
    while n > 30 and ref_time_model(n) > budget_s:
        n -= 2
    return n


def run_reference_sample(n_sample: int, cores: int, idx0: int = 0, timeout_s: float = None):
    """One reference process per core on `cores` prefixes of the workload; returns (cells/s, folds/s, outputs)."""
    exe = ROOT / "oracle" / "_ref" / "CCJ"
    kind = "reference"
    if not exe.exists():
        exe = ROOT / "oracle" / "_ref" / "ccj_oracle"
        kind = "port"
    if not exe.exists():
        return None
    seqs = workload(idx0, cores, n_sample)
    par = str(ROOT / "params" / PAR)
    t0 = time.perf_counter()
    procs = [subprocess.Popen([str(exe), "-P", par, s], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for s in seqs]
    outs = []
    for p in procs:
        try:
            left = None if timeout_s is None else max(1.0, timeout_s - (time.perf_counter() - t0))
            outs.append(p.communicate(timeout=left) + (p.returncode,))
        except subprocess.TimeoutExpired:
            for q in procs:
                if q.poll() is None:
                    q.kill()
            return {"timed_out": True, "seconds": time.perf_counter() - t0, "kind": kind}
    dt = time.perf_counter() - t0
    return {"cells_per_s": cores * cells(n_sample) / dt, "folds_per_s": cores / dt, "seconds": dt, "kind": kind,
            "seqs": seqs, "outs": outs}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    total_steps = args.steps + args.warmup
    n_sample = pick_sample_len(150.0 / max(total_steps, 1))
    times, last = [], None
    for s in range(total_steps):
        last = run_reference_sample(n_sample, cores, idx0=s * cores)
        if last is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built on this box"}))
            return
        if s >= args.warmup:
            times.append(last["seconds"])
    dt = sum(times) / len(times)
    value = cores * cells(n_sample) / dt
    sample = (f"{cores} prefixes of length {n_sample} of the 150-nt workload per step, one reference process per "
              f"core; cells = C(n+1,4)")
    # the same-config figure: ONE full-length 150-nt workload sequence per core, measured once (about 6 minutes);
    # the K timed steps above stay on prefixes so that the run ends within minutes
    full = None
    if args.ref_full_length and args.gpus == 1:
        r = run_reference_sample(N_NT, cores, idx0=0, timeout_s=args.ref_full_timeout)
        if r is not None and not r.get("timed_out"):
            full = {"cells_per_s": r["cells_per_s"], "folds_per_s": r["folds_per_s"], "seconds": r["seconds"],
                    "cores": cores, "sample": f"workload sequences 0..{cores - 1} at their full 150 nt, one process per core, once",
                    "same_config": True}
        elif r is not None:
            full = {"timed_out_after_s": r["seconds"]}
    line = {
        "impl": "reference", "metric": "dp_cells_per_s_150nt_batch", "value": value, "unit": "cells/s",
        "folds_per_s": cores / dt, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
        "data": "synthetic", "config": {"workload": "config4: random 150-nt, seeds 20000+idx (CPU sample: see "
                                                    "cpu_baseline.sample)", "params": PAR, "dangles": 2},
        "cpu_baseline": {"value": value, "unit": "cells/s", "cores": cores, "kind": last["kind"], "sample": sample},
        "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "full_length": full,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def dist_setup():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return world, rank, local


def per_kernel_roofline(prof, terms, peak_gbs, addmin, traffic_json=None):
    """Every kernel group against both roofs.  Algorithmic bytes / candidates per group: split-point
    roles 2 B per split term; windows 2 B per evaluated window term; assembly 44 B per cell (the 22 stores);
    compute_P 4 B (two reads) per term.  One candidate = one add + one min = one VIADDMNMX.
    `hbm_frac` is the survey's algorithmic-byte figure (it exceeds 1 where one fetched record serves several levels);
    `dram_frac` uses the DRAM bytes ncu measured for that kernel (profiles/r2_traffic.json, scaled by algorithmic
    bytes to this step); `bound` = the larger of dram_frac and int32_frac, i.e. the roof that is actually closest."""
    split = sum(x["split"] for x in terms)
    iloop = sum(x["iloop"] for x in terms)
    pterm = sum(x["pterms"] for x in terms)
    ncell = sum(x["cells"] for x in terms)
    groups = {
        "k_roles": (2 * split, split, prof["k4d_split_ms"], "int16_unpack_viaddmnmx"),
        "k_winLR+k_winM": (2 * iloop, iloop, prof["k4d_window_ms"], "int16x2_viaddmnmx"),
        "k_final": (44 * ncell, 0, prof["k4d_final_ms"], "int32_viaddmnmx"),
        "k_P_tuned": (4 * pterm, pterm, prof["kP_ms"], "int32_viaddmnmx"),
    }
    out = {}
    ncu_names = {"k_roles": ("k_roles",), "k_winLR+k_winM": ("k_winLR", "k_winM"), "k_final": ("k_final",),
                 "k_P_tuned": ("k_P_tuned",)}   # prefixes of the kernel names in the ncu launch list
    scale = None
    if traffic_json:
        alg_all = sum(x["bytes_4d"] for x in terms)
        scale = alg_all / traffic_json["algorithmic_bytes"]
    for name, (nbytes, cands, ms, form) in groups.items():
        if ms <= 0:
            continue
        gbs = nbytes / (ms / 1e3) / 1e9
        cps = cands / (ms / 1e3)
        hbm_frac = gbs / peak_gbs
        int_frac = cps / addmin[form] if addmin.get(form) else None
        dram = None
        if scale is not None:
            pk = traffic_json.get("per_kernel", {})
            dram = sum((v["dram_read_GB"] + v["dram_write_GB"]) * 1e9 for k, v in pk.items() if k.startswith(ncu_names[name])) * scale
        dram_frac = dram / (ms / 1e3) / 1e9 / peak_gbs if dram else None
        roofs = {"hbm": dram_frac if dram_frac is not None else hbm_frac, "int32": int_frac or 0.0}
        out[name] = {"ms": ms, "algorithmic_bytes": nbytes, "candidates": cands, "achieved_gbs": gbs,
                     "hbm_frac": hbm_frac, "dram_bytes_ncu": dram, "dram_frac": dram_frac, "candidates_per_s": cps,
                     "int32_frac": int_frac, "int32_form": form, "bound": max(roofs, key=roofs.get)}
    return out


def config3_block(ctx, ccj_b200, peak):
    """BASELINE config 3: the single 200-nt fills, fill only (traceback separately), parity with the goldens."""
    gold = load_goldens("folds_long.json")
    seqs = [s for s in gold if len(s) == 200]
    out = []
    for s in seqs:
        ctx.prepare([s])
        ctx.fill()                       # warm-up (graph capture, first touch)
        fills = [ctx.fill() for _ in range(3)]
        tb = ctx.traceback()
        f = ctx.fetch()[0]
        t = ccj_b200.count_terms(s)
        fill_ms = min(fills)
        gbs = t["bytes"] / (fill_ms / 1e3) / 1e9
        out.append({"n": 200, "fill_ms": fill_ms, "fill_ms_all": fills, "traceback_ms": tb,
                    "algorithmic_bytes": t["bytes"], "cells": t["cells"], "cells_per_s": t["cells"] / (fill_ms / 1e3),
                    "achieved_gbs": gbs, "frac": gbs / peak, "energy": f.energy,
                    "should_not_be_here": f.n_should_not_be_here, "parity": same_as_golden(f, gold[s])})
    if not out:
        return None
    worst = max(out, key=lambda r: r["fill_ms"])
    return {"workload": "config3: SURVEY App. C seed-200 200-mer and designed(200), one sequence per fill",
            "fill_ms": worst["fill_ms"], "frac": worst["frac"], "parity": all(r["parity"] for r in out), "runs": out,
            "roofline_note": "frac = (2 B x (split + window + 2 x P terms) + 44 B x cells) / fill time / measured HBM peak"}


def config2_block(ctx):
    """BASELINE config 2: 1,024 random 100-nt sequences through ccj_fold_batch with host buffers."""
    seqs = workload2(1024)
    gold = load_goldens("folds_long.json", "folds_config2.json")
    dt = None
    for _ in range(2):   # the first pass also captures the launch graphs of the two wave shapes; the better pass counts
        t0 = time.perf_counter()
        folds = ctx.fold_batch(seqs)
        t = time.perf_counter() - t0
        dt = t if dt is None else min(dt, t)
    checked = [(f, gold[s]) for s, f in zip(seqs, folds) if s in gold]
    return {"workload": "config2: 1024 random 100-nt sequences, seeds 1000+idx, one GPU, host buffers",
            "seconds": dt, "folds_per_s": len(seqs) / dt, "cells_per_s": len(seqs) * cells(100) / dt,
            "reference_exit1": sum(f.status == 1 for f in folds),
            "should_not_be_here_lines": sum(f.n_should_not_be_here for f in folds),
            "golden_checked": len(checked), "parity": all(same_as_golden(f, r) for f, r in checked)}


def gpu_arm(args):
    import torch
    import torch.distributed as dist
    import ccj_b200

    world, rank, local = dist_setup()
    ctx = ccj_b200.Context(local, str(ROOT / "params" / PAR), 2)
    B = min(args.batch, ctx.wave_capacity(N_NT))
    cells_step = B * cells(N_NT)
    total_steps = args.warmup + args.steps

    # every step folds new workload sequences: timed step x of rank r takes batch x*world + r (so the timed region starts
    # at workload index 0, where the golden vectors are), the warm-up steps take the batches after the timed ones
    order = list(range(args.steps, total_steps)) + list(range(args.steps))
    batches = [workload((x * world + rank) * B, B) for x in order]

    def batch_of(step):
        return batches[step]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # --- device-resident arm: the step's inputs are staged (prepare), then fill + traceback are timed on the device ---
    for s in range(args.warmup):
        ctx.prepare(batch_of(s))
        ctx.fill()
        ctx.traceback()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    dev_ms = 0.0
    folds = []
    for s in range(args.warmup, total_steps):
        ctx.prepare(batch_of(s))
        dev_ms += ctx.fill()
        dev_ms += ctx.traceback()
        folds += ctx.fetch()
    barrier()
    launches = ctx.last_fill_launches + 1
    # --- end-to-end arm: host buffers in, host results out, through ccj_fold_batch ---
    barrier()
    t0 = time.perf_counter()
    e2e_folds = []
    for s in range(args.warmup, total_steps):
        e2e_folds += ctx.fold_batch(batch_of(s))
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    assert [f.stdout for f in e2e_folds] == [f.stdout for f in folds]

    t = torch.tensor([dev_ms / 1e3, e2e_s], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_s, e2e_s = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = world * cells_step * args.steps / dev_s
    e2e_value = world * cells_step * args.steps / e2e_s
    seqs = batch_of(total_steps - 1)
    h2d = sum(2 * len(s) + 2 for s in seqs) + B * 160
    d2h = sum(4 * (8 + 2 * len(s) + 3) for s in seqs)
    gold4 = load_goldens("folds_long.json", "folds_config4.json")
    checked4 = [(f, gold4[f.sequence]) for f in folds if f.sequence in gold4]

    # --- roofline of the dominant kernel group, measured live with per-launch CUDA events ---
    ctx.prepare(seqs)
    prof = ctx.fill_profiled()
    terms = [ccj_b200.count_terms(s) for s in seqs]
    bytes_4d = sum(x["bytes_4d"] for x in terms)
    peak, peak_src = hbm_peak()
    addmin = ctx.addmin_peak()
    achieved = bytes_4d / (prof["k4d_ms"] / 1e3) / 1e9
    # SURVEY.md 8d also defines the figure over the WHOLE fill (gap tables + compute_P, graph launch, all streams)
    bytes_fill = bytes_4d + sum(x["bytes_P"] for x in terms)
    fill_ms = ctx.fill()
    # measured DRAM traffic: ncu dram__bytes_{read,write}.sum summed over every gap-table launch of one fold of this
    # workload (profiles/*_traffic.json, written by profiles/measure_traffic.py from the committed ncu launch list),
    # scaled from its algorithmic bytes to this step's
    traffic, traffic_src, tj = None, None, None
    for name in ("r2_traffic.json", "r1_traffic.json"):
        tpath = ROOT / "profiles" / name
        if tpath.exists():
            tj = json.loads(tpath.read_text())
            traffic = tj["dram_bytes"] / tj["algorithmic_bytes"] * bytes_4d
            traffic_src = tj.get("source")
            break
    per_kernel = per_kernel_roofline(prof, terms, peak, addmin, tj)
    cand_4d = sum(x["split"] + x["iloop"] for x in terms)
    int_frac = cand_4d / (prof["k4d_ms"] / 1e3) / addmin["int32_viaddmnmx"]
    roofline = {"bound": "int32" if int_frac > achieved / peak else "hbm", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "int32_frac": int_frac, "traffic": traffic,
                "traffic_source": traffic_src,
                "kernel": "level-wavefront gap-table kernels (k_roles + k_winLR + k_winM + k_final), all launches of one step",
                "algorithmic_bytes_per_step": bytes_4d, "candidates_per_step": cand_4d,
                "kernel_ms_per_step": prof["k4d_ms"],
                "kernel_share_of_fill": prof["k4d_ms"] / max(prof["total_ms"], 1e-9), "peak_source": peak_src,
                "addmin_peak_pairs_per_s": addmin,
                "addmin_peak_source": "measured live (ccj_measure_addmin_peak: register-only VIADDMNMX chains on all SMs)",
                "kernel_ms": {"roles": prof["k4d_split_ms"], "windows": prof["k4d_window_ms"], "final": prof["k4d_final_ms"],
                              "P": prof["kP_ms"], "2D": prof["k2d_ms"], "other": prof["other_ms"]},
                "per_kernel": per_kernel,
                "whole_fill": {"algorithmic_bytes": bytes_fill, "ms": fill_ms,
                               "achieved": bytes_fill / (fill_ms / 1e3) / 1e9,
                               "frac": bytes_fill / (fill_ms / 1e3) / 1e9 / peak}}

    # --- the other single-GPU configs and the CPU baseline (N=1 only) ---
    cpu = c3 = c2 = None
    if world == 1:
        c3 = config3_block(ctx, ccj_b200, peak)
        c2 = config2_block(ctx)
        cores = os.cpu_count() or 1
        n_sample = pick_sample_len(args.cpu_budget)
        r = run_reference_sample(n_sample, cores)
        if r is not None:
            mine = ctx.fold_batch(r["seqs"])
            parity = all((f.stdout, f.returncode) == (o[0], o[2]) for f, o in zip(mine, r["outs"]))
            cpu = {"value": r["cells_per_s"], "unit": "cells/s", "cores": cores, "kind": r["kind"],
                   "sample": f"{cores} prefixes of length {n_sample} of the workload sequences, one process per core, "
                             f"{r['seconds']:.1f} s wall", "folds_per_s": r["folds_per_s"],
                   "gpu_output_identical_on_sample": parity}

    line = {
        "metric": "dp_cells_per_s_150nt_batch", "value": value, "unit": "cells/s",
        "folds_per_s": value / cells(N_NT), "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_s / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32 (int16 tables)", "data": "synthetic",
        "config": {"workload": "config4: random 150-nt sequences, seeds 20000+idx, independent-sequence sharding; "
                               "every step folds new sequences",
                   "batch_per_gpu": B, "global_batch": B * world, "seq_len": N_NT, "params": PAR, "dangles": 2,
                   "l2_note": "tables of one step (>1 GB per sequence) exceed the 126 MB L2"},
        "e2e": {"value": e2e_value, "unit": "cells/s", "folds_per_s": e2e_value / cells(N_NT),
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
        "statuses": {"ok": sum(f.status == 0 for f in folds), "reference_exit1": sum(f.status == 1 for f in folds),
                     "should_not_be_here_lines": sum(f.n_should_not_be_here for f in folds)},
        "golden_150nt": {"checked": len(checked4), "parity": all(same_as_golden(f, r) for f, r in checked4)},
        "config3": c3, "config2": c2,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# --full: the whole config-4 job (8,192 x 150 nt) once, sequences dealt dynamically
# ------------------------------------------------------------------------------------------------
def full_arm(args):
    """Ranks draw chunks of consecutive workload indices from one shared counter (the process group's TCP store,
    atomic add) until the job is empty: a GPU that runs slower (power capping) simply draws fewer chunks.  No
    data-path collective; results are gathered on the host and hashed in index order, so the SHA-256 must be the
    same for every N."""
    import torch
    import torch.distributed as dist
    import ccj_b200
    from ccj_b200 import shard

    world, rank, local = dist_setup()
    ctx = ccj_b200.Context(local, str(ROOT / "params" / PAR), 2)
    total = args.full_count
    chunk = min(args.full_chunk, ctx.wave_capacity(N_NT))
    ctx.fold_batch(workload(0, min(chunk, 8)))   # warm-up: library, arena
    counter = shard.StoreCounter(dist.distributed_c10d._get_default_store(), "ccj_full_next") if world > 1 \
        else shard.LocalCounter()
    lib_stream = torch.cuda.ExternalStream(ctx.stream_ptr, device=torch.device("cuda", local))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record(lib_stream)
    mine = shard.fold_dealt(total, chunk, counter, lambda lo, hi: ctx.fold_batch(workload(lo, hi - lo)))
    e1.record(lib_stream)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    dev_s = e0.elapsed_time(e1) / 1e3
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([dev_s, wall], dtype=torch.float64, device=f"cuda:{local}")
    parts = [mine]
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        parts = [None] * world
        dist.all_gather_object(parts, mine)
    if rank == 0:
        res = sorted((x for p in parts for x in p), key=lambda x: x[0])
        assert [x[0] for x in res] == list(range(total)), "every index exactly once"
        h = hashlib.sha256()
        for idx, f in res:
            h.update(repr((idx, f.returncode, f.stdout, f.stderr)).encode())
        gold = load_goldens("folds_long.json", "folds_config4.json")
        checked = [(f, gold[f.sequence]) for _, f in res if f.sequence in gold]
        secs = float(t[0])
        print(json.dumps({
            "metric": "dp_cells_per_s_150nt_batch", "mode": "full", "value": total * cells(N_NT) / secs,
            "unit": "cells/s", "folds_per_s": total / secs, "seconds": secs, "wall_seconds": float(t[1]),
            "n_gpus": world, "scaling": "strong", "higher_is_better": True, "data": "synthetic",
            "config": {"workload": f"config4: all {total} random 150-nt sequences (seeds 20000+idx), chunks of {chunk} "
                                   "dealt dynamically from a shared counter", "params": PAR, "dangles": 2},
            "timing": "CUDA events on the library's stream around the rank's whole share, max over ranks; host "
                      "buffers in and out (ccj_fold_batch)",
            "chunks_per_rank": [len(p) // max(chunk, 1) + (1 if len(p) % max(chunk, 1) else 0) for p in parts],
            "sequences_per_rank": [len(p) for p in parts],
            "sha256": h.hexdigest(),
            "statuses": {"ok": sum(f.status == 0 for _, f in res), "reference_exit1": sum(f.status == 1 for _, f in res),
                         "should_not_be_here_lines": sum(f.n_should_not_be_here for _, f in res)},
            "golden_150nt": {"checked": len(checked), "parity": all(same_as_golden(f, r) for f, r in checked)},
            "clocks": clocks}))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=48, help="150-nt sequences per GPU per step")
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU work for the cpu_baseline sample")
    ap.add_argument("--full", action="store_true", help="fold the whole config-4 job once (dynamic dealing)")
    ap.add_argument("--full-count", type=int, default=8192)
    ap.add_argument("--full-chunk", type=int, default=64)
    ap.add_argument("--config5", action="store_true", help="one oversized sequence, gap tables sharded by outer index")
    ap.add_argument("--n5", type=int, default=600)
    ap.add_argument("--hash5", action="store_true", help="config5: also hash every table on rank 0 (small n only)")
    ap.add_argument("--no-ref-full-length", dest="ref_full_length", action="store_false",
                    help="reference arm: skip the single full-length 150-nt wave (about 6 minutes)")
    ap.add_argument("--ref-full-timeout", type=float, default=720.0)
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    elif args.full:
        full_arm(args)
    elif args.config5:
        from ccj_b200 import shard5
        shard5.bench_config5(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()

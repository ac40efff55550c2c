#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config, one JSON line on stdout.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

Metric  : 4D DP cells/s (and folds/s) on the 150-nt batch (BASELINE config 4: random.Random(20000+idx)).
Step    : one pass of the hot path (fill of V/WM/P/22 gap tables/W + traceback) over one batch of B
          sequences per GPU.  Sequences are independent -> ranks fold disjoint index ranges, no collective
          on the data path (weak scaling: B per GPU is fixed).
value   : device-resident throughput (inputs already in HBM; CUDA events on the library's stream).
e2e     : the same metric through the C-ABI call a user makes (ccj_fold_batch) with HOST buffers: H2D of the
          sequences, fill, traceback, D2H of energies/pairs and dot-bracket rendering, wall clock.
roofline: for the dominant kernel (the level-wavefront gap-table kernel): algorithmic bytes
          (2 B x every min-plus candidate read + 44 B per cell written, SURVEY.md 8d, counted exactly per input)
          / summed device time of its launches, against MEASURED_PEAKS.json's HBM copy bandwidth.
cpu_baseline / --impl reference: the unmodified reference binary (oracle/_ref/CCJ) on the host cores, one
          process per core, on a bounded sample (prefixes of the same workload sequences, length chosen so a
          step fits the time budget; the reference's cells/s only falls with n, so this flatters the CPU).
"""
from __future__ import annotations

import argparse
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


def workload(idx0: int, count: int, n: int = N_NT):
    out = []
    for idx in range(idx0, idx0 + count):
        rng = random.Random(SEED0 + idx)
        out.append("".join(rng.choice("ACGU") for _ in range(N_NT))[:n])
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


def run_reference_sample(n_sample: int, cores: int, idx0: int = 0):
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
    outs = [p.communicate() + (p.returncode,) for p in procs]
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
    line = {
        "impl": "reference", "metric": "dp_cells_per_s_150nt_batch", "value": value, "unit": "cells/s",
        "folds_per_s": cores / dt, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
        "data": "synthetic", "config": {"workload": "config4: random 150-nt, seeds 20000+idx (CPU sample: see "
                                                    "cpu_baseline.sample)", "params": PAR, "dangles": 2},
        "cpu_baseline": {"value": value, "unit": "cells/s", "cores": cores, "kind": last["kind"], "sample": sample},
        "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def gpu_arm(args):
    import torch
    import torch.distributed as dist
    import ccj_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = ccj_b200.Context(local, str(ROOT / "params" / PAR), 2)
    B = args.batch
    cap = ctx.wave_capacity(N_NT)
    if B > cap:
        B = cap
    seqs = workload(rank * B, B)
    cells_step = B * cells(N_NT)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # --- device-resident arm: inputs staged once, then fill + traceback per step ---
    ctx.prepare(seqs)
    for _ in range(args.warmup):
        ctx.fill()
        ctx.traceback()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    dev_ms = 0.0
    for _ in range(args.steps):
        dev_ms += ctx.fill()
        dev_ms += ctx.traceback()
    barrier()
    folds = ctx.fetch()
    launches = ctx.last_fill_launches + 1
    # --- end-to-end arm: host buffers in, host results out, through ccj_fold_batch ---
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_folds = ctx.fold_batch(seqs)
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
    h2d = sum(2 * len(s) + 2 for s in seqs) + B * 160
    d2h = sum(4 * (8 + 2 * len(s) + 3) for s in seqs)

    # --- roofline of the dominant kernel, measured live with per-launch CUDA events ---
    prof = ctx.fill_profiled()
    terms = [ccj_b200.count_terms(s) for s in seqs]
    bytes_4d = sum(x["bytes_4d"] for x in terms)
    peak, peak_src = hbm_peak()
    achieved = bytes_4d / (prof["k4d_ms"] / 1e3) / 1e9
    # SURVEY.md 8d also defines the figure over the WHOLE fill (gap tables + compute_P, graph launch, all streams)
    bytes_fill = bytes_4d + sum(x["bytes_P"] for x in terms)
    fill_ms = ctx.fill()
    # measured DRAM traffic: ncu dram__bytes_{read,write}.sum summed over every gap-table launch of one fold of this
    # workload (profiles/r1_traffic.json, written by profiles/measure_traffic.py from the committed ncu launch list),
    # scaled from its algorithmic bytes to this step's
    traffic, traffic_src = None, None
    tpath = ROOT / "profiles" / "r1_traffic.json"
    if tpath.exists():
        tj = json.loads(tpath.read_text())
        traffic = tj["dram_bytes"] / tj["algorithmic_bytes"] * bytes_4d
        traffic_src = tj.get("source")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src,
                "kernel": "level-wavefront gap-table kernels (k_roles + k_winLR + k_winM + k_final), all launches of one step",
                "algorithmic_bytes_per_step": bytes_4d, "kernel_ms_per_step": prof["k4d_ms"],
                "kernel_share_of_fill": prof["k4d_ms"] / max(prof["total_ms"], 1e-9), "peak_source": peak_src,
                "kernel_ms": {"roles": prof["k4d_split_ms"], "windows": prof["k4d_window_ms"], "final": prof["k4d_final_ms"],
                              "P": prof["kP_ms"], "2D": prof["k2d_ms"], "other": prof["other_ms"]},
                "whole_fill": {"algorithmic_bytes": bytes_fill, "ms": fill_ms,
                               "achieved": bytes_fill / (fill_ms / 1e3) / 1e9,
                               "frac": bytes_fill / (fill_ms / 1e3) / 1e9 / peak}}

    # --- CPU baseline on a bounded sample + parity of the GPU path on exactly that sample ---
    cpu = None
    if world == 1:
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
        "config": {"workload": "config4: random 150-nt sequences, seeds 20000+idx, independent-sequence sharding",
                   "batch_per_gpu": B, "global_batch": B * world, "seq_len": N_NT, "params": PAR, "dangles": 2,
                   "l2_note": "tables of one step (>1 GB per sequence) exceed the 126 MB L2"},
        "e2e": {"value": e2e_value, "unit": "cells/s", "folds_per_s": e2e_value / cells(N_NT),
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
        "statuses": {"ok": sum(f.status == 0 for f in folds), "reference_exit1": sum(f.status == 1 for f in folds),
                     "should_not_be_here_lines": sum(f.n_should_not_be_here for f in folds)},
    }
    print(json.dumps(line))
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
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()

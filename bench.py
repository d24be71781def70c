#!/usr/bin/env python
"""bench.py — soil cell-steps/sec of the fused RHS + SSPRK33 path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (C restatement)

A "step" is one SSPRK33 step (3 fused RHS+stage launches) of the whole column set.  At N = 1 the
workload is BASELINE.json configs[3] — 2^20 independent columns x 64 layers, coupled water+heat
(parameters of the reference's test/SoilModel/coupled.jl, seeded synthetic profiles of SURVEY §8d) —
which fits one GPU (2.7 GB).  For N > 1 the SAME 2^20 columns are cut into contiguous shards, one
process per GPU, no data-path collective ("scaling": "strong"); NCCL is used only for the 2-double
budget all-reduce, which is timed separately.

One JSON line on stdout (rank 0).  value = whole-job cell-steps/s with the state resident in HBM,
timed with CUDA events on the ctx stream (max over ranks).  e2e = the same job through the C ABI
from pinned HOST buffers: H2D state upload + K x (step with a host-built bc table + budget read) +
D2H state download, wall-clocked with a device sync on both sides; the columns are cut into --e2e-shards shards (one ctx,
stream and host thread each) whose transfers take turns on the PCIe link and overlap the kernels of the other shards.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "soil cell-steps/sec (fp64, coupled water+heat, SSPRK33)"
UNIT = "cell-steps/s"
BYTES_PER_CELL_STEP = {"coupled": 152, "richards": 88}     # BASELINE.md §3 (algorithmic, per-stage fusion)
REF_SAMPLE_COLS = 16384                                    # bounded sample for the CPU arms


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="coupled", choices=["coupled", "richards"])
    ap.add_argument("--ncol", type=int, default=1 << 20)
    ap.add_argument("--nlayer", type=int, default=64)
    ap.add_argument("--general-vg", action="store_true",
                    help="never use the van Genuchten n == 2 square-root specialisation (LH_FLAG_GENERAL_VG)")
    ap.add_argument("--launch", default="auto", choices=["auto", "stage", "persistent"],
                    help="lh_soil_step_ssprk33 strategy: one launch per stage, one persistent launch per call, or the library's "
                         "own choice (persistent for small launch-bound grids of <= 3 waves)")
    ap.add_argument("--het", action="store_true",
                    help="heterogeneous soils: random per-column nu / theta_r / van Genuchten n, alpha / Ksat (lh_soil_set_column_params)")
    ap.add_argument("--e2e-shards", type=int, default=8,
                    help="e2e leg: column shards per GPU, one ctx (= one stream) and one host thread each, so that the upload of "
                         "one shard, the steps of another and the download of a third overlap; 1 = a single ctx")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def make_workload(w, model, ncol, nlayer, col_range):
    if model == "coupled":
        return w.coupled_workload(ncol=ncol, nlayer=nlayer, col_range=col_range)
    return w.richards_workload(ncol=ncol, nlayer=nlayer, col_range=col_range,
                               zlim=(-1.5 * nlayer / 100.0, 0.0))


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [l for (ts, l) in self.lines if t0 - 0.05 <= ts <= t1 + 0.15] or [l for _, l in self.lines]
        sm, smax, reasons = [], None, set()
        for l in rows:
            parts = [x.strip() for x in l.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); smax = float(parts[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax,
                "reasons": sorted(reasons), "samples": len(sm)}


def pinned_like(a: np.ndarray) -> np.ndarray:
    import torch

    t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
    out = t.numpy()
    out[...] = a
    out_base_keepalive.append(t)
    return out


out_base_keepalive = []


def bc_table_for(wl, t0, dt, nsteps):
    """What the host mirror's Simulation builds: the LH_BCV_* 4-vector at the 3 stage times of each
    step (here the Dirichlet closures are constants, as in the reference's tests)."""
    vals = np.array([wl.top[1], wl.top[3], wl.bottom[1], wl.bottom[3]], dtype=np.float64)
    return np.broadcast_to(vals, (nsteps, 3, 4)).copy()


def time_oracle(w, lh, graft, model, nlayer, steps, warmup, target_seconds=None):
    """The reference's CPU path (C restatement, OpenMP over columns) on a bounded sample."""
    import ctypes as C

    wl = make_workload(w, model, REF_SAMPLE_COLS, nlayer, (0, REF_SAMPLE_COLS))
    lib = lh.SoilLibrary(graft.build_oracle(), "lho_")
    # all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1: override it)
    set_threads = lib.raw("lho_soil_set_num_threads")
    set_threads.restype = None
    set_threads.argtypes = [C.c_int32]
    set_threads(len(os.sched_getaffinity(0)))
    nthreads = lib.raw("lho_soil_num_threads")
    nthreads.restype = C.c_int32
    cores = int(nthreads())
    ctx = lh.SoilContext(lib, wl.config())
    wl.upload(ctx)
    ctx.step(0.0, wl.dt, max(1, warmup) if target_seconds is None else 1)
    if target_seconds is not None:
        t0 = time.perf_counter(); ctx.step(0.0, wl.dt, 1); one = time.perf_counter() - t0
        steps = int(min(max(2, round(target_seconds / max(one, 1e-6))), 200))
    t0 = time.perf_counter()
    ctx.step(0.0, wl.dt, steps)
    sec = time.perf_counter() - t0
    value = wl.cells * steps / sec
    sample = (f"{REF_SAMPLE_COLS} columns x {nlayer} layers ({wl.cells} cells) of the same {model} workload, "
              f"{steps} SSPRK33 steps, {sec:.1f} s")
    return value, cores, sample, sec, steps, wl


def main():
    args = parse_args()
    # stdout carries exactly ONE JSON line.  Native libraries write to file descriptor 1 behind Python's back (NCCL prints
    # its version banner there): point fd 1 at stderr for the duration of the run and keep the real stdout for the result.
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    try:
        return _main(args, real_stdout)
    finally:
        real_stdout.flush()


def _main(args, result_stream):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    import __graft_entry__ as graft

    lh = graft.load_package()
    import workloads as w

    model_id = {"coupled": lh._abi.LH_MODEL_COUPLED, "richards": lh._abi.LH_MODEL_RICHARDS}[args.model]
    config = {
        "workload": f"{args.ncol} columns x {args.nlayer} layers, {args.model} "
                    f"(BASELINE.json configs[3]: 1M columns x 64 layers coupled water+heat)"
                    if args.model == "coupled" else f"{args.ncol} columns x {args.nlayer} layers, richards (Bonan sand)",
        "columns": args.ncol, "layers": args.nlayer, "stepper": "SSPRK33, fused RHS+stage kernel",
        "sharding": f"contiguous column ranges over {world} GPU(s), no halo",
        "l2": "state (2.7 GB at N=1) is larger than the 126 MB L2; no flush needed",
        "closures": "general van Genuchten n (log/exp form)" if (args.general_vg or args.model != "coupled")
                    else "n = 2 of the coupled.jl parameters -> square-root specialisation (automatic; --general-vg disables)",
    }

    # ---------------- reference arm: the reference's CPU implementation of the path -----------------
    if args.impl == "reference":
        if rank != 0:
            return 0
        value, cores, sample, sec, steps, wl = time_oracle(w, lh, graft, args.model, args.nlayer, args.steps, args.warmup)
        line = {
            "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec / steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "note": "C restatement of the reference CPU path (Julia is not installable here), OpenMP over columns"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line), file=result_stream)
        return 0

    # ---------------- this repo's arm ------------------------------------------------------------------
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: this path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    lo, hi = lh.shard_range(args.ncol, world, rank)
    wl = make_workload(w, args.model, args.ncol, args.nlayer, (lo, hi))
    wl.device = local_rank
    lib = lh.cuda_library()
    flags = (lh._abi.LH_FLAG_GENERAL_VG if args.general_vg else 0) | \
            {"auto": 0, "stage": lh._abi.LH_FLAG_STAGE_LAUNCHES, "persistent": lh._abi.LH_FLAG_PERSISTENT}[args.launch]
    ctx = lh.SoilContext(lib, wl.config(flags=flags))
    if args.het:
        rng = np.random.default_rng(11 + rank)
        n_ = wl.ncol
        ctx.set_column_params(nu=wl.params.nu * rng.uniform(0.98, 1.15, n_), theta_r=rng.uniform(0.0, 0.02, n_),
                              vg_n=rng.uniform(1.5, 3.5, n_), vg_alpha=wl.params.vg_alpha * rng.uniform(0.5, 2.0, n_),
                              Ksat=wl.params.Ksat * 10.0 ** rng.uniform(-1.0, 1.0, n_))
        config["closures"] = "heterogeneous: per-column nu, theta_r, van Genuchten n and alpha, Ksat (general closures, per-lane parameters)"
    host = {fid: pinned_like(a) for fid, a in wl.fields.items()}
    for fid, a in host.items():
        ctx.set_state(fid, a)
    if world > 1:
        class _Eng:  # minimal engine-like holder for init_budget_comm
            pass
        eng = _Eng(); eng.lib = lib; eng.ctx = ctx
        lh.init_budget_comm(eng, dist)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ctx.sync()

    # warm-up (W >= 3 untimed steps)
    t = 0.0
    ctx.step(t, wl.dt, args.warmup)
    ctx.sync()
    t += args.warmup * wl.dt

    # ---- timed region: exactly K steps, state resident in HBM, CUDA events on the ctx stream ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    barrier()
    w0 = time.perf_counter()
    ctx.step(t, wl.dt, args.steps)
    ms, launches = ctx.last_step_timing()          # cudaEventElapsedTime(start, stop) on the ctx stream
    barrier()
    w1 = time.perf_counter()
    clocks = sampler.stop(w0, w1) if rank == 0 else None
    t += args.steps * wl.dt
    ms_t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_max = float(ms_t.item())

    # ---- the one collective: global water/energy budgets ----
    budgets = ctx.budgets_allreduce() if world > 1 else ctx.budgets()     # first call: NCCL lazy set-up
    tb0 = time.perf_counter()
    budgets = ctx.budgets_allreduce() if world > 1 else ctx.budgets()
    budget_ms = 1e3 * (time.perf_counter() - tb0)
    if not np.all(np.isfinite(budgets)):
        raise SystemExit(f"non-finite budgets after the timed steps: {budgets}")

    # ---- e2e through the C ABI from pinned host buffers ----
    e2e = None
    if not args.no_e2e:
        K = args.steps
        table = bc_table_for(wl, 0.0, wl.dt, 1)
        out = {fid: pinned_like(np.empty_like(wl.fields[fid])) for fid in ((0, 2) if args.model == "coupled" else (0,))}
        # Columns are independent, so the public API lets a host cut them into shards, one ctx each: every ctx has its own
        # stream, the calls release the GIL, and PCIe (2.7 GB per 20 steps) overlaps with the kernels of the other shards.
        S = 1 if args.het else max(1, min(args.e2e_shards, (hi - lo) // 4096))
        ncol_r = hi - lo
        cuts = [(ncol_r * k // S) // 32 * 32 for k in range(S)] + [ncol_r]
        subs = []
        for k in range(S):
            c0, c1 = cuts[k], cuts[k + 1]
            sub = ctx if S == 1 else lh.SoilContext(lib, wl.config(ncol=c1 - c0, flags=flags))
            if S > 1:
                sub.set_state(1, host[1][c0:c1])               # untimed: first-use allocation of the ctx's staging buffers
            subs.append((sub, c0, c1))

        # Uploads (and downloads) of the shards take turns on the PCIe link in shard order, so shard k computes while
        # shard k+1 uploads and shard k-1 downloads, instead of all shards moving in lockstep.
        up_done = [threading.Event() for _ in range(S)]
        down_done = [threading.Event() for _ in range(S)]

        def run_shard(k, sub, c0, c1):
            if k > 0:
                up_done[k - 1].wait()
            for fid, a in host.items():
                sub.set_state(fid, a[c0:c1])                   # H2D (+ layout transform on device)
            up_done[k].set()
            tt = 0.0
            for _ in range(K):
                sub.step(tt, wl.dt, 1, table)                  # host-evaluated bc values for the 3 stage times
                sub.budgets()                                  # D2H read of the step's result (16 B)
                tt += wl.dt
            if k > 0:
                down_done[k - 1].wait()
            for fid, a in out.items():
                sub.get_state(fid, a[c0:c1])                   # D2H
            down_done[k].set()

        barrier()
        e0 = time.perf_counter()
        if S == 1:
            run_shard(0, *subs[0])
        else:
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(S) as pool:
                list(pool.map(lambda kt: run_shard(kt[0], *kt[1]), enumerate(subs)))
            for sub, _, _ in subs:
                sub.sync()
        barrier()
        e_sec = time.perf_counter() - e0
        e_t = torch.tensor([e_sec], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(e_t, op=dist.ReduceOp.MAX)
        e_sec = float(e_t.item())
        cells_total = args.ncol * args.nlayer
        nfields_in, nfields_out = len(host), len(out)
        e2e = {
            "value": cells_total * K / e_sec, "unit": UNIT,
            "h2d_bytes_per_step": int(nfields_in * cells_total * 8 / K + 96),
            "d2h_bytes_per_step": int(nfields_out * cells_total * 8 / K + 16),
            "what": f"lh_soil_set_state x{nfields_in} (pinned host, reference layout) + {K} x [lh_soil_step_ssprk33(1 step, bc table) + "
                    f"lh_soil_budgets] + lh_soil_get_state x{nfields_out}, over {S} column shard(s) per GPU (one ctx and host thread "
                    f"each, transfers overlapping kernels); wall clock, max over ranks",
            "shards_per_gpu": S,
            "seconds": e_sec,
        }

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    cells_total = args.ncol * args.nlayer
    value = cells_total * args.steps / (ms_max * 1e-3)
    peak, peak_src = load_peaks()
    cells_rank = (hi - lo) * args.nlayer
    persistent = int(launches) != 3 * args.steps
    if not persistent:
        # dominant kernel = lh_soil_stage_kernel<model, stage 1|2|3>: every launch in the timed region is one of its three
        # stage instantiations; per-launch figures are the averages over the 3K launches.
        bpcs = BYTES_PER_CELL_STEP[args.model]
        bytes_per_launch = cells_rank * bpcs / 3.0
        launch_ms = ms_max / (3 * args.steps)
        kernel = "lh_soil_stage_kernel (fused closures + stencil + SSPRK33 stage)"
        note = ("per-launch average over the 3 stage launches of each step (40/56/56 B per cell coupled, 24/32/32 Richards); "
                "besides HBM the kernel is bounded by issue slots: an fp64 instruction holds the issue port for two "
                "cycles on B200 (DESIGN.md §4.1)")
    else:
        # one persistent launch for all K steps: a block keeps its columns, the stage registers stay in L2, and the
        # compulsory traffic of a launch is one read of the state and one write of the prognostic fields per STEP
        bpcs = {"coupled": 40, "richards": 24}[args.model]
        bytes_per_launch = cells_rank * bpcs * args.steps / max(int(launches), 1)
        launch_ms = ms_max / max(int(launches), 1)
        kernel = "lh_soil_ssprk33_persistent_kernel (all stages of all steps in one launch, columns L2-resident)"
        note = ("persistent launch (grid of few waves): algorithmic bytes are 40 B (coupled) / 24 B (Richards) per cell-STEP, "
                "the path is issue-bound, not HBM-bound (DESIGN.md §4.1)")
    achieved = bytes_per_launch / (launch_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(tpath) and not persistent:
        try:
            traffic = json.load(open(tpath)).get(f"{args.model}_{args.ncol}x{args.nlayer}_bytes_per_launch")
            if traffic is not None:
                traffic = traffic * cells_rank / (args.ncol * args.nlayer)     # measured at N = 1; per launch of this rank's shard
        except Exception:
            traffic = None
    config["launch"] = "persistent (1 launch per call)" if persistent else "3 launches per step"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": config, "clocks": clocks,
        "gpu_launches": int(launches),
        "roofline": {
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "peak_source": peak_src,
            "kernel": kernel, "algorithmic_bytes_per_cell_step": bpcs, "launch_ms": launch_ms, "note": note,
        },
        "budgets": {"water": float(budgets[0]), "energy": float(budgets[1]), "allreduce_ms": budget_ms},
    }
    if e2e is not None:
        line["e2e"] = e2e
    if world == 1 and not args.no_cpu_baseline:
        v, cores, sample, sec, steps, _ = time_oracle(w, lh, graft, args.model, args.nlayer, 0, 0, target_seconds=12.0)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    print(json.dumps(line), file=result_stream)
    result_stream.flush()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

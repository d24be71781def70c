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

One JSON line on stdout (rank 0).

value   whole-job cell-steps/s with the state resident in HBM.  A timed BLOCK is exactly K steps, bracketed by a
        barrier and a device synchronisation, timed with CUDA events on the ctx stream (max over ranks).  Blocks are
        repeated back to back until at least --min-seconds (3 s) of device time have been measured; `value` is the MEDIAN
        block (sustained clocks), the first (burst) block and the spread are in `sustained`.
roofline  contract fraction (SURVEY §8d bytes: θ_i counted in every stage) AND the on-wire fraction (the bytes the
        launched kernel variant really moves), with the variant named by the library itself (lh_soil_kernel_info).
e2e     the same job through the C ABI from pinned HOST buffers: H2D state upload + K x (step with a host-built bc
        table + budget read) + D2H state download, wall clock with a device sync on both sides; the columns are cut
        into --e2e-shards shards (one ctx, stream and host thread each) whose transfers take turns on the PCIe link
        and overlap the kernels of the other shards.
cpu_baseline / --impl reference   the C restatement of the reference CPU path (OpenMP over columns) on the SAME
        column set and layer count (full 2^20 x 64 unless K + W is so large that the run would not end in minutes).
extra.variants  the kernel specialisations real soils hit (general van Genuchten n, ice, impedance + viscosity,
        Richards, heterogeneous columns) and the other BASELINE shapes, timed in this process (tools/variants.py).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

METRIC = "soil cell-steps/sec (fp64, coupled water+heat, SSPRK33)"
UNIT = "cell-steps/s"
BYTES_PER_CELL_STEP = {"coupled": 152, "richards": 88}     # BASELINE.md §3 (algorithmic, per-stage fusion; the contract figure)
REF_MAX_STEPS_FULL = 60                                    # reference arm: the full column set while K + W <= this
DEVICE = "cuda"                                            # tests/test_bench_multirank.py drives the N > 1 control flow on CPU (gloo)
EXIT_HARD = True                                           # multi-rank runs end with os._exit after the last barrier
E2E_DEADLINE_S = 240.0                                     # a multi-rank e2e leg that has not finished by then is reported as failed
_T0 = time.perf_counter()


def progress(msg):
    """Phase marker on stderr (never stdout): a run that stops somewhere says where (round 2's first N > 1 run hung silently)."""
    print(f"[bench rank {os.environ.get('RANK', '0')} +{time.perf_counter() - _T0:7.2f}s] {msg}", file=sys.stderr, flush=True)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="coupled", choices=["coupled", "richards"])
    ap.add_argument("--ncol", type=int, default=1 << 20)
    ap.add_argument("--nlayer", type=int, default=64)
    ap.add_argument("--min-seconds", type=float, default=3.0,
                    help="repeat the timed K-step block until this much device time has been measured (sustained clocks)")
    ap.add_argument("--general-vg", action="store_true",
                    help="never use the van Genuchten n == 2 square-root specialisation (LH_FLAG_GENERAL_VG)")
    ap.add_argument("--ice", action="store_true", help="θ_i ~ U(0, 0.05) in every cell: the ICE kernel variants")
    ap.add_argument("--launch", default="auto", choices=["auto", "stage", "persistent"],
                    help="lh_soil_step_ssprk33 strategy: one launch per stage, one persistent launch per call, or the library's "
                         "own choice (persistent only for small, launch-bound grids of <= 1.5 waves of resident blocks)")
    ap.add_argument("--no-chain", action="store_true",
                    help="whole-grid dependency between consecutive stage launches (LH_FLAG_NO_CHAIN) instead of block-to-block")
    ap.add_argument("--het", action="store_true",
                    help="heterogeneous soils: random per-column nu / theta_r / van Genuchten n, alpha / Ksat (lh_soil_set_column_params)")
    ap.add_argument("--e2e-shards", type=int, default=8,
                    help="e2e leg: column shards per GPU, one ctx (= one stream) and one host thread each, so that the upload of "
                         "one shard, the steps of another and the download of a third overlap; 1 = a single ctx")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-variants", action="store_true", help="skip the extra.variants table (N = 1 only)")
    return ap.parse_args()


def make_workload(w, model, ncol, nlayer, col_range, ice=False):
    if model == "coupled":
        return w.coupled_workload(ncol=ncol, nlayer=nlayer, col_range=col_range, ice=ice)
    return w.richards_workload(ncol=ncol, nlayer=nlayer, col_range=col_range, ice=ice,
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
                ["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50"],
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
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [l for (ts, l) in self.lines if t0 + 0.05 <= ts <= t1]     # strictly inside the timed region
        sm, power, smax, reasons = [], [], None, set()
        for l in rows:
            parts = [x.strip() for x in l.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); smax = float(parts[2]); power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_min_mhz": float(np.min(sm)) if sm else None,
                "sm_max_mhz": smax, "power_w_median": float(np.median(power)) if power else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def pinned_like(a: np.ndarray) -> np.ndarray:
    import torch

    t = torch.empty(a.shape, dtype=torch.float64, pin_memory=(DEVICE == "cuda"))
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


def time_oracle(w, lh, graft, model, ncol, nlayer, steps, warmup, target_seconds=None, ice=False):
    """The reference's CPU path (C restatement, OpenMP over columns) on `ncol` columns of the same workload."""
    import ctypes as C

    wl = make_workload(w, model, ncol, nlayer, (0, ncol), ice=ice)
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
    t0 = time.perf_counter()
    ctx.step(0.0, wl.dt, 1)
    one = time.perf_counter() - t0
    if max(0, warmup - 1) > 0:
        ctx.step(0.0, wl.dt, warmup - 1)
    if target_seconds is not None:
        steps = int(min(max(2, round(target_seconds / max(one, 1e-6))), 200))
    t0 = time.perf_counter()
    ctx.step(0.0, wl.dt, steps)
    sec = time.perf_counter() - t0
    value = wl.cells * steps / sec
    sample = (f"{ncol} columns x {nlayer} layers ({wl.cells} cells) of the same {model} workload, "
              f"{steps} SSPRK33 steps after {max(1, warmup)} warm-up step(s), {sec:.1f} s")
    ctx.close()
    return value, cores, sample, sec, steps


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
        import faulthandler

        faulthandler.cancel_dump_traceback_later()
        real_stdout.flush()


def _main(args, result_stream):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    import faulthandler

    import __graft_entry__ as graft

    lh = graft.load_package()
    import workloads as w

    # A run that stops somewhere must say where and must end: Python stacks of every thread go to stderr if the whole run
    # is still going after 25 minutes, and the process exits (the driver then sees a failed run with a reason, not a hang).
    faulthandler.dump_traceback_later(1500.0, exit=True, file=sys.stderr)

    config = {
        "workload": f"{args.ncol} columns x {args.nlayer} layers, {args.model} "
                    f"(BASELINE.json configs[3]: 1M columns x 64 layers coupled water+heat)"
                    if args.model == "coupled" else f"{args.ncol} columns x {args.nlayer} layers, richards (Bonan sand)",
        "columns": args.ncol, "layers": args.nlayer, "stepper": "SSPRK33, fused RHS+stage kernel",
        "sharding": f"contiguous column ranges over {world} GPU(s), no halo",
        "l2": "state (2.7 GB at N=1) is larger than the 126 MB L2; no flush needed",
        "closures": "general van Genuchten n (fixed-exponent power tables)" if (args.general_vg or args.model != "coupled")
                    else "n = 2 of the coupled.jl parameters -> square-root specialisation (automatic; --general-vg disables)",
        "ice": bool(args.ice),
    }

    # ---------------- reference arm: the reference's CPU implementation of the path -----------------
    if args.impl == "reference":
        if rank != 0:
            return 0
        total = args.steps + max(1, args.warmup)
        ref_cols = args.ncol if total <= REF_MAX_STEPS_FULL else max(16384, (args.ncol * REF_MAX_STEPS_FULL // total) // 16384 * 16384)
        value, cores, sample, sec, steps = time_oracle(w, lh, graft, args.model, ref_cols, args.nlayer, args.steps, args.warmup, ice=args.ice)
        line = {
            "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec / steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "same_config": ref_cols == args.ncol,
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
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    # Pinned host buffers are placed on the NUMA node of the allocating thread: sit next to this rank's GPU first.
    affinity0 = os.sched_getaffinity(0)
    numa_bound = lh.bind_to_gpu_numa_node(local_rank)

    progress(f"world {world}: process group up, numa_bound={bool(numa_bound)}")
    lo, hi = lh.shard_range(args.ncol, world, rank)
    wl = make_workload(w, args.model, args.ncol, args.nlayer, (lo, hi), ice=args.ice)
    wl.device = local_rank
    lib = lh.cuda_library()
    A = lh._abi
    flags = (A.LH_FLAG_GENERAL_VG if args.general_vg else 0) | (A.LH_FLAG_NO_CHAIN if args.no_chain else 0) | \
            {"auto": 0, "stage": A.LH_FLAG_STAGE_LAUNCHES, "persistent": A.LH_FLAG_PERSISTENT}[args.launch]
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
    kernel_info = ctx.kernel_info()
    progress(f"columns [{lo}, {hi}) uploaded; {kernel_info}")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ctx.sync()

    def rank_max(x):
        tt = torch.tensor(x, dtype=torch.float64, device=DEVICE)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return tt.cpu().numpy()

    # warm-up (W >= 3 untimed steps)
    t = 0.0
    ctx.step(t, wl.dt, args.warmup)
    ctx.sync()
    t += args.warmup * wl.dt
    progress(f"{args.warmup} warm-up steps done")

    # ---- timed region: blocks of exactly K steps, state resident in HBM, CUDA events on the ctx stream ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)

    def timed_block():
        nonlocal t
        barrier()
        ctx.step(t, wl.dt, args.steps)
        ms, launches = ctx.last_step_timing()          # cudaEventElapsedTime(start, stop) on the ctx stream; syncs on stop
        t += args.steps * wl.dt
        return ms, launches

    barrier()
    w0 = time.perf_counter()
    ms0, launches = timed_block()
    ms0 = float(rank_max([ms0])[0])
    nblocks = int(min(2000, max(3, math.ceil(args.min_seconds * 1e3 / max(ms0, 1e-3)))))
    block_ms = [ms0]
    for _ in range(nblocks - 1):
        ms, launches = timed_block()
        block_ms.append(ms)
    barrier()
    w1 = time.perf_counter()
    clocks = sampler.stop(w0, w1) if rank == 0 else None
    block_ms = rank_max(block_ms)                          # per block: the slowest rank
    ms_med = float(np.median(block_ms))
    progress(f"timed region done: {len(block_ms)} blocks of {args.steps} steps, median {ms_med / args.steps:.4f} ms/step")
    if rank == 0 and (clocks is None or not clocks.get("samples")) and (w1 - w0) < 0.25 and args.min_seconds < 0.25:
        # profiling runs (--min-seconds 0 under ncu): a timed region shorter than a few sampling periods cannot be vetted
        clocks = dict(clocks or {}, note=f"timed region {w1 - w0:.3f} s is shorter than the 50 ms sampling period allows; not vetted")
    elif rank == 0 and (clocks is None or not clocks.get("samples")):
        raise SystemExit(f"clock sampler recorded no nvidia-smi sample inside the {w1 - w0:.2f} s timed region: "
                         "the measurement cannot be vetted for throttling (is nvidia-smi on PATH?)")

    # ---- the one collective: global water/energy budgets ----
    budgets = ctx.budgets_allreduce() if world > 1 else ctx.budgets()     # first call: NCCL lazy set-up
    tb0 = time.perf_counter()
    budgets = ctx.budgets_allreduce() if world > 1 else ctx.budgets()
    budget_ms = 1e3 * (time.perf_counter() - tb0)
    if not np.all(np.isfinite(budgets)):
        raise SystemExit(f"non-finite budgets after the timed steps: {budgets}")
    progress(f"budgets {'all-reduced over ' + str(world) + ' ranks' if world > 1 else 'read'} in {budget_ms:.3f} ms")

    def finish_distributed():
        """End of a multi-rank run.  ncclCommDestroy is an intra-node collective ("all ranks on the same node should call [it]
        to avoid a hang"): in round 2 rank 0 closed its ctx (and with it the library's budget communicator) while the other
        ranks were already tearing down torch's group or gone, and the N = 2 and N = 4 runs never returned.  Nothing is torn
        down piecemeal any more: every rank waits at one last barrier (so nobody leaves while a peer is still inside a
        collective), the result line is already on stdout, and the process ends with os._exit — no communicator destructor
        runs, in any order.  A watchdog covers a barrier that does not return."""
        if world > 1:
            result_stream.flush()
            sys.stderr.flush()
            t_kill = threading.Timer(60.0, lambda: os._exit(0))
            t_kill.daemon = True
            t_kill.start()
            try:
                dist.barrier()
            finally:
                if EXIT_HARD:
                    os._exit(0)

    def make_line(e2e):
        cells_total = args.ncol * args.nlayer
        value = cells_total * args.steps / (ms_med * 1e-3)
        peak, peak_src = load_peaks()
        cells_rank = (hi - lo) * args.nlayer
        persistent = int(launches) != 3 * args.steps
        wire_bpcs = None
        try:
            wire_bpcs = int(kernel_info.rsplit("=", 1)[1])
        except Exception:
            pass
        if not persistent:
            # dominant kernel = lh_soil_stage_kernel<model, stage 1|2|3>: every launch in the timed region is one of its three
            # stage instantiations; per-launch figures are the averages over the 3K launches.
            bpcs = BYTES_PER_CELL_STEP[args.model]
            nl = 3 * args.steps
            kernel = "lh_soil_stage_kernel (fused closures + stencil + SSPRK33 stage)"
            note = ("per-launch average over the 3 stage launches of each step (contract: 40/56/56 B per cell coupled, 24/32/32 Richards). "
                    "`frac` credits the SURVEY §8d contract bytes; `frac_on_wire` the bytes the launched variant really moves (the !ICE "
                    "variants never read theta_i). The kernel is bounded by issue slots, not HBM: an fp64 instruction holds the issue "
                    "port for two cycles on B200 (DESIGN.md §4.1), see `issue_model`")
        else:
            # one persistent launch for all K steps: a block keeps its columns, the stage registers stay in L2, and the
            # compulsory traffic of a launch is one read of the state and one write of the prognostic fields per STEP
            bpcs = {"coupled": 40, "richards": 24}[args.model]
            wire_bpcs = bpcs - (0 if args.ice or args.model != "coupled" else 8)
            nl = max(int(launches), 1)
            kernel = "lh_soil_ssprk33_persistent_kernel (all stages of all steps in one launch, columns L2-resident)"
            note = ("persistent launch (grid of few waves): algorithmic bytes are 40 B (coupled) / 24 B (Richards) per cell-STEP, "
                    "the path is issue-bound, not HBM-bound (DESIGN.md §4.1)")
        launch_ms = ms_med / nl
        achieved = cells_rank * bpcs * args.steps / nl / (launch_ms * 1e-3) / 1e9
        on_wire = cells_rank * (wire_bpcs or bpcs) * args.steps / nl / (launch_ms * 1e-3) / 1e9
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")
        if os.path.exists(tpath) and not persistent and not args.het:
            try:
                tkey = args.model + ("_general_vg" if (args.general_vg and args.model == "coupled") else "") + ("_ice" if args.ice else "")
                traffic = json.load(open(tpath)).get(f"{tkey}_{args.ncol}x{args.nlayer}_bytes_per_launch")
                if traffic is not None:
                    traffic = traffic * cells_rank / (args.ncol * args.nlayer)     # measured at N = 1; per launch of this rank's shard
                    traffic_src = "profiles/dram_traffic.json (CACHED ncu --set full capture of this command, not measured in this run)"
            except Exception:
                traffic = None
        config["launch"] = "persistent (1 launch per call)" if persistent else "3 launches per step, chained block to block"
        if args.no_chain:
            config["launch"] = "3 launches per step, whole-grid dependency (LH_FLAG_NO_CHAIN)"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_med / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config, "clocks": clocks,
            "gpu_launches": int(launches) * len(block_ms),
            "sustained": {
                "what": f"{len(block_ms)} back-to-back timed blocks of exactly {args.steps} steps; value = median block",
                "blocks": len(block_ms), "device_seconds": float(np.sum(block_ms)) * 1e-3,
                "ms_per_step_first_block": float(block_ms[0]) / args.steps, "ms_per_step_min": float(np.min(block_ms)) / args.steps,
                "ms_per_step_median": ms_med / args.steps, "ms_per_step_max": float(np.max(block_ms)) / args.steps,
                "value_first_block": cells_total * args.steps / (float(block_ms[0]) * 1e-3),
            },
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "achieved_on_wire": on_wire, "frac_on_wire": on_wire / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "kernel": kernel, "variant": kernel_info,
                "algorithmic_bytes_per_cell_step": bpcs, "on_wire_bytes_per_cell_step": wire_bpcs, "launch_ms": launch_ms, "note": note,
            },
            "budgets": {"water": float(budgets[0]), "energy": float(budgets[1]), "allreduce_ms": budget_ms},
        }
        # issue-slot model of the launched variant (static SASS counts of the layer loop, profiles/r02_sass_loop_mix.json)
        try:
            mix = json.load(open(os.path.join(ROOT, "profiles", "r02_sass_loop_mix.json")))
            key = kernel_info.split("FLAGS=")[1].split(":")[0]
            mkey = f"{ {'coupled': 2, 'richards': 0}[args.model] }_{key}"
            if mkey in mix and clocks and clocks.get("sm_mhz"):
                m = mix[mkey]
                cyc = m["cycles_per_warp_cell_stage_mean"]
                sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
                ceiling = sm_count * 4 * clocks["sm_mhz"] * 1e6 * 32 / cyc / 3.0 * world
                line["roofline"]["issue_model"] = {
                    "cycles_per_warp_cell_stage": cyc, "fp64_per_cell_stage": m["fp64_mean"], "other_per_cell_stage": m["other_mean"],
                    "ceiling_cell_steps_per_s_at_measured_clock": ceiling, "frac_of_issue_ceiling": value / ceiling,
                    "what": "2 x fp64 + other warp-instructions per cell and stage in the layer loop (cuobjdump -sass, "
                            "profiles/r02_sass_loop_mix.json) x 4 schedulers x SMs x median SM clock"}
        except Exception:
            pass
        if e2e is not None:
            line["e2e"] = e2e
        return line

    # ---- e2e through the C ABI from pinned host buffers ----
    def run_e2e():
        K = args.steps
        table_K = bc_table_for(wl, 0.0, wl.dt, K)
        out = {fid: pinned_like(np.empty_like(wl.fields[fid])) for fid in ((0, 2) if args.model == "coupled" else (0,))}
        # Columns are independent, so the public API lets a host cut them into shards, one ctx each: every ctx has its own
        # stream, the calls release the GIL, and PCIe (2.7 GB per 20 steps) overlaps with the kernels of the other shards.
        S = 1 if args.het else max(1, min(args.e2e_shards, (hi - lo) // 32768))      # >= 32768 columns per shard
        ncol_r = hi - lo
        cuts = [(ncol_r * k // S) // 32 * 32 for k in range(S)] + [ncol_r]
        subs = []
        for k in range(S):
            c0, c1 = cuts[k], cuts[k + 1]
            sub = ctx if S == 1 else lh.SoilContext(lib, wl.config(ncol=c1 - c0, flags=flags))
            if S > 1:
                # untimed warm-up of the shard's ctx: first-use allocations (staging blocks, the pinned budget history of
                # lh_soil_run, events) must not sit inside the timed region, which starts by uploading the state again
                for fid, a in host.items():
                    sub.set_state(fid, a[c0:c1])
                sub.run(0.0, wl.dt, K, bc_table=table_K, budget_every=1)
                sub.sync()
            subs.append((sub, c0, c1))

        # Uploads (and downloads) of the shards take turns on the PCIe link in shard order, so shard k computes while
        # shard k+1 uploads and shard k-1 downloads, instead of all shards moving in lockstep.
        up_done = [threading.Event() for _ in range(S)]
        down_done = [threading.Event() for _ in range(S)]
        step_budgets = np.zeros((S, K, 2))

        marks = np.zeros((S, 6))                               # per shard: upload start / end, run end, download start / end (s)

        def run_shard(k, sub, c0, c1):
            if k > 0:
                up_done[k - 1].wait()
            marks[k, 0] = time.perf_counter()
            for fid, a in host.items():
                sub.set_state(fid, a[c0:c1])                   # H2D (+ layout transform on device)
            marks[k, 1] = time.perf_counter()
            up_done[k].set()
            # run!(sim): K steps in ONE call; after every step the budgets (the step's result, 16 B) are reduced on the
            # device and copied to a pinned slot on the host, asynchronously, while the next step runs.
            bud, _ = sub.run(0.0, wl.dt, K, bc_table=table_K, budget_every=1)
            marks[k, 2] = time.perf_counter()
            step_budgets[k] = bud
            if k > 0:
                down_done[k - 1].wait()
            marks[k, 3] = time.perf_counter()
            for fid, a in out.items():
                sub.get_state(fid, a[c0:c1])                   # D2H
            marks[k, 4] = time.perf_counter()
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
        e_sec = float(rank_max([e_sec])[0])
        if not np.all(np.isfinite(step_budgets)):
            raise SystemExit("non-finite per-step budgets in the e2e leg")
        cells_total = args.ncol * args.nlayer
        nfields_in, nfields_out = len(host), len(out)
        e2e = {
            "value": cells_total * K / e_sec, "unit": UNIT,
            "h2d_bytes_per_step": int(nfields_in * cells_total * 8 / K + 96),
            "d2h_bytes_per_step": int(nfields_out * cells_total * 8 / K + 16),
            "what": f"lh_soil_set_state x{nfields_in} (pinned host, reference layout) + lh_soil_run({K} steps, host-built bc table, budgets "
                    f"reduced and copied D2H after EVERY step, 16 B each) + lh_soil_get_state x{nfields_out}, over {S} column shard(s) per GPU "
                    f"(one ctx and host thread each, transfers overlapping kernels); wall clock, max over ranks",
            "numa_bound": bool(numa_bound),
            "shards_per_gpu": S,
            "seconds": e_sec,
        }
        # where the wall time of this rank's e2e leg went (host clocks around the blocking calls of each shard's thread)
        up_s, run_s, down_s = marks[:, 1] - marks[:, 0], marks[:, 2] - marks[:, 1], marks[:, 4] - marks[:, 3]
        up_bytes = nfields_in * (hi - lo) * args.nlayer * 8
        down_bytes = nfields_out * (hi - lo) * args.nlayer * 8
        e2e["breakdown_rank0"] = {
            "what": "per-shard host wall clock of the blocking calls: upload (H2D + device transpose), lh_soil_run (K steps, competing "
                    "with the other shards' kernels and transposes), download (device transpose + D2H); uploads / downloads of the "
                    "shards take turns, so their sums are the time the PCIe link was claimed in each direction",
            "upload_s_sum": float(up_s.sum()), "run_s_mean": float(run_s.mean()), "run_s_max": float(run_s.max()), "download_s_sum": float(down_s.sum()),
            "h2d_GBps_while_uploading": float(up_bytes / max(up_s.sum(), 1e-9) / 1e9),
            "d2h_GBps_while_downloading": float(down_bytes / max(down_s.sum(), 1e-9) / 1e9),
            "first_upload_to_last_upload_end_s": float(marks[:, 1].max() - marks[:, 0].min()),
            "last_upload_end_to_last_download_end_s": float(marks[:, 4].max() - marks[:, 1].max()),
        }
        for sub, _, _ in subs:
            if sub is not ctx:
                sub.close()

        return e2e

    # The line is complete without the e2e leg; that leg runs under a deadline so that a transfer or a shard thread that never
    # returns (on any rank) costs the run its e2e number, not its result: at the deadline every rank dumps its Python stacks,
    # rank 0 prints the line with the failure recorded in `e2e`, and the processes end.
    emit_lock = threading.Lock()
    emitted = []

    def emit(line):
        with emit_lock:
            if not emitted:
                emitted.append(True)
                print(json.dumps(line), file=result_stream)
                result_stream.flush()

    def e2e_deadline():
        progress(f"e2e leg still running after {E2E_DEADLINE_S:.0f} s: giving up on it")
        faulthandler.dump_traceback(file=sys.stderr, all_threads=True)
        if rank == 0:
            emit(make_line({"value": None, "unit": UNIT, "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                            "error": f"the e2e leg did not finish within {E2E_DEADLINE_S:.0f} s on some rank (stacks on stderr)"}))
        sys.stderr.flush()
        os._exit(0)                                        # the failure is on record in the line; a non-zero rank exit would make torchrun tear rank 0 down mid-print

    e2e = None
    if not args.no_e2e:
        watchdog = threading.Timer(E2E_DEADLINE_S, e2e_deadline)
        watchdog.daemon = True
        watchdog.start()
        e2e = run_e2e()
        watchdog.cancel()
        progress(f"e2e leg done: {e2e['seconds']:.4f} s over {e2e['shards_per_gpu']} shard(s) per GPU")

    if rank != 0:
        finish_distributed()
        return 0

    line = make_line(e2e)
    # From here on the line only GROWS (CPU baseline, variant table).  If one of those legs never returns, the line as it stood
    # at the last completed leg is printed instead of nothing: `tail_json` is refreshed after every leg and variant row.
    tail_json = [json.dumps(line)]

    def tail_deadline():
        progress("the CPU-baseline / variant legs are still running after 15 minutes: printing the line without what is missing")
        faulthandler.dump_traceback(file=sys.stderr, all_threads=True)
        with emit_lock:
            if not emitted:
                emitted.append(True)
                print(tail_json[0], file=result_stream)
                result_stream.flush()
        sys.stderr.flush()
        os._exit(0)

    tail_watchdog = threading.Timer(900.0, tail_deadline)
    tail_watchdog.daemon = True
    if world == 1:
        tail_watchdog.start()
    if world == 1:
        ctx.close()                                        # free the 2.7 GB before the CPU baseline and the variant table
        del host
    os.sched_setaffinity(0, affinity0)                     # the CPU legs use every core this process was given
    if world == 1 and not args.no_cpu_baseline:
        v, cores, sample, sec, steps = time_oracle(w, lh, graft, args.model, args.ncol, args.nlayer, 0, 1, target_seconds=12.0, ice=args.ice)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        tail_json[0] = json.dumps(line)
        progress(f"cpu baseline done: {v:.4g} {UNIT} on {cores} cores")
    if world == 1 and not args.no_variants:
        import variants as V

        rows = {}
        vnote = "variant table incomplete: the run was cut off after these rows"
        for name, spec in V.variant_specs(lh, w).items():
            try:
                rows[name] = V.measure(lh, spec, steps=args.steps, warmup=args.warmup, reps=5, device=local_rank)
            except Exception as exc:                      # a variant that fails must not hide the headline
                rows[name] = {"error": f"{type(exc).__name__}: {exc}"}
            tail_json[0] = json.dumps(dict(line, extra={"variants": dict(rows), "variants_note": vnote}))
            progress(f"variant {name}: {rows[name].get('cell_steps_per_s', rows[name].get('error'))}")
        line["extra"] = {"variants": rows,
                         "variants_note": "same process, same GPU, CUDA events on the ctx stream; per variant the MEDIAN of back-to-back "
                                          "blocks of K steps (small configs: 50 K) covering >= 1 s of device time, i.e. sustained "
                                          "(power-capped) clocks like `value`; cell_steps_per_s_first_block is the burst figure; "
                                          "frac_contract / frac_on_wire as in `roofline`"}
    tail_watchdog.cancel()
    emit(line)
    progress("result line written")
    finish_distributed()
    return 0


if __name__ == "__main__":
    sys.exit(main())

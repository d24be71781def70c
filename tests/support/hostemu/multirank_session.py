"""Run by tests/test_hostemu.py in a child process (a clean one: no torch, hence no real libnccl mapped): N host threads act as
the ranks of a column-sharded run on the EMULATED build of the product library, with tests/support/hostemu/fake_nccl.cpp standing
in for libnccl.so.2.  Mirrors tests/test_multi_gpu.py::test_two_gpu_sharded_run_matches_single.  Prints one JSON line."""
import ctypes as C
import json
import os
import sys
import threading

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import hostemu  # noqa: E402

C.CDLL(hostemu.build_fake_nccl(), mode=C.RTLD_GLOBAL)          # before the product library looks for libnccl.so.2
import workloads as w  # noqa: E402

lh = w.lh
lib = hostemu.library(lh)
world, ncol, nlayer, nsteps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
bc = dict(top=(w.F, 0.0, w.F, 0.0), bottom=(w.F, 0.0, w.F, 0.0))
uid = lib.comm_unique_id()
states, locals_, totals, errors = [None] * world, [None] * world, [None] * world, []


def rank_main(rank):
    try:
        lo, hi = lh.shard_range(ncol, world, rank)
        wl = w.coupled_workload(ncol=ncol, nlayer=nlayer, seed=77, col_range=(lo, hi), **bc)
        ctx = lh.SoilContext(lib, wl.config())
        wl.upload(ctx)
        ctx.comm_init(world, rank, uid)                        # a rendezvous of all ranks, as with NCCL
        ctx.step(0.0, wl.dt, nsteps)
        totals[rank] = ctx.budgets_allreduce()                 # fused stage-3 sums -> local reduction -> all-reduce
        locals_[rank] = ctx.budgets()
        states[rank] = ctx.get_state(0)
        ctx.close()
    except Exception as exc:  # noqa: BLE001
        errors.append(f"rank {rank}: {type(exc).__name__}: {exc}")


threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
for t in threads:
    t.start()
for t in threads:
    t.join(timeout=300)
if errors or any(t.is_alive() for t in threads):
    print(json.dumps({"errors": errors or ["a rank did not finish"]}))
    os._exit(1)
wl = w.coupled_workload(ncol=ncol, nlayer=nlayer, seed=77, **bc)
ctx = lh.SoilContext(lib, wl.config())
wl.upload(ctx)
W0 = ctx.budgets()
ctx.step(0.0, wl.dt, nsteps)
whole, Wn = ctx.get_state(0), ctx.budgets()
print(json.dumps({
    "shards_bit_identical": bool(np.array_equal(np.concatenate(states), whole)),
    "same_total_on_every_rank": bool(all(np.array_equal(totals[0], t) for t in totals)),
    "total_is_sum_of_locals": bool(np.allclose(totals[0], np.sum(locals_, axis=0), rtol=1e-15, atol=0.0)),
    "total_vs_unsharded_rel": float(np.max(np.abs(totals[0] - Wn) / np.abs(Wn))),
    "conserved_rel": float(np.max(np.abs(totals[0] - W0) / np.abs(W0))),
}))

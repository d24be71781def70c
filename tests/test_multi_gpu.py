"""Column shards across GPUs and the one collective on this path (NCCL budget all-reduce).

* 1 GPU: a single-rank NCCL communicator exercises lh_soil_comm_unique_id / lh_soil_comm_init /
  lh_soil_budgets_allreduce end to end.
* >= 2 GPUs (``gpurun --gpus 2``): two processes, one ctx per GPU over contiguous column ranges, no
  halo; shards must evolve bit-identically to the unsharded run and every rank must see the same
  global budgets.  Skipped on a 1-GPU box.  The host-side logic is also covered on CPU by
  tests/test_sharding_gloo.py.
"""
import os

import numpy as np
import pytest

import workloads as w

pytestmark = pytest.mark.gpu

lh = w.lh


def _ngpus():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


def test_single_rank_nccl_allreduce(cuda):
    wl = w.coupled_workload(ncol=512, nlayer=32, seed=5)
    ctx = lh.SoilContext(cuda, wl.config())
    wl.upload(ctx)
    uid = cuda.comm_unique_id()
    assert len(uid) == 128 and any(uid)
    ctx.comm_init(1, 0, uid)
    local = ctx.budgets()
    glob = ctx.budgets_allreduce()
    assert np.array_equal(local, glob)


def test_allreduce_before_comm_init_is_an_error(cuda):
    wl = w.coupled_workload(ncol=32, nlayer=8, seed=5)
    ctx = lh.SoilContext(cuda, wl.config())
    with pytest.raises(lh.SoilError):
        ctx.budgets_allreduce()


def _worker(rank, world, port, ncol, nlayer, nsteps, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        lib = lh.cuda_library()
        lo, hi = lh.shard_range(ncol, world, rank)
        wl = w.coupled_workload(ncol=ncol, nlayer=nlayer, seed=77, col_range=(lo, hi),
                                top=(w.F, 0.0, w.F, 0.0), bottom=(w.F, 0.0, w.F, 0.0))
        wl.device = rank
        ctx = lh.SoilContext(lib, wl.config())
        wl.upload(ctx)

        class Eng:
            pass

        e = Eng(); e.lib = lib; e.ctx = ctx
        lh.init_budget_comm(e, dist)
        ctx.step(0.0, wl.dt, nsteps)
        total = ctx.budgets_allreduce()
        np.save(os.path.join(out_dir, f"state_{rank}.npy"), ctx.get_state(0))
        np.save(os.path.join(out_dir, f"budget_{rank}.npy"), np.concatenate([ctx.budgets(), total]))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_ngpus() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_gpu_sharded_run_matches_single(cuda, tmp_path):
    import torch.multiprocessing as mp

    ncol, nlayer, nsteps, world = 4096, 64, 5, 2
    port = 29700 + (os.getpid() % 2000)
    mp.start_processes(_worker, args=(world, port, ncol, nlayer, nsteps, str(tmp_path)), nprocs=world,
                       join=True, start_method="spawn")
    wl = w.coupled_workload(ncol=ncol, nlayer=nlayer, seed=77, top=(w.F, 0.0, w.F, 0.0), bottom=(w.F, 0.0, w.F, 0.0))
    ctx = lh.SoilContext(cuda, wl.config())
    wl.upload(ctx)
    W0 = ctx.budgets()
    ctx.step(0.0, wl.dt, nsteps)
    parts = [np.load(tmp_path / f"state_{r}.npy") for r in range(world)]
    assert np.array_equal(np.concatenate(parts), ctx.get_state(0))      # no halo: bit-identical shards
    b = [np.load(tmp_path / f"budget_{r}.npy") for r in range(world)]
    assert np.array_equal(b[0][2:], b[1][2:])                           # same global budget on every rank
    assert np.allclose(b[0][2:], b[0][:2] + b[1][:2], rtol=1e-15)
    assert np.allclose(b[0][2:], ctx.budgets(), rtol=1e-13)             # sharding-independent to ~1e-13
    assert np.allclose(b[0][2:], W0, rtol=1e-12)                        # zero-flux BCs: conserved

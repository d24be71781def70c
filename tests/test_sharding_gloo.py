"""The N > 1 path on CPU: world_size-2 ``gloo`` processes.

Checks the host-side multi-GPU logic without a GPU: contiguous column shards, the
unique-id hand-off plumbing (rank 0 -> all, through torch.distributed), shard-local stepping and the
global budget = sum of shard budgets.  The ranks drive the CPU checker library through the same
ctypes harness; on the B200 the same code path runs with the CUDA library and NCCL
(tests/test_multi_gpu.py, bench.py --gpus N)."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

import workloads as w

lh = w.lh


def test_shard_ranges_partition_the_columns():
    for ncol, nranks in ((1 << 20, 8), (1000, 3), (7, 7), (1025, 2)):
        s = lh.ColumnShards(ncol, nranks)
        r = s.ranges()
        assert r[0][0] == 0 and r[-1][1] == ncol
        assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert max(s.counts()) - min(s.counts()) <= 1
        assert s.owner(0) == 0 and s.owner(ncol - 1) == nranks - 1
    assert lh.shard_range(1 << 20, 8, 3) == (3 << 17, 4 << 17)
    with pytest.raises(ValueError):
        lh.shard_range(4, 8, 0)


def _worker(rank, world, port, ncol, nlayer, nsteps, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from landhydrology_b200 import sharding

        lib = w.oracle_library()
        lo, hi = lh.shard_range(ncol, world, rank)
        wl = w.coupled_workload(ncol=ncol, nlayer=nlayer, seed=77, col_range=(lo, hi),
                                top=(w.F, 0.0, w.F, 0.0), bottom=(w.F, 0.0, w.F, 0.0))
        ctx = lh.SoilContext(lib, wl.config())
        wl.upload(ctx)
        # plumbing used for the NCCL communicator on the GPU path
        uid = sharding.exchange_unique_id(lib, dist)
        assert len(uid) == 128
        ctx.comm_init(world, rank, uid)
        ctx.step(0.0, wl.dt, nsteps)
        local = ctx.budgets()
        total = sharding.allreduce_budgets_host(local, dist)
        np.save(os.path.join(out_dir, f"state_{rank}.npy"), ctx.get_state(0))
        np.save(os.path.join(out_dir, f"budget_{rank}.npy"), np.concatenate([local, total]))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_run_matches_single(tmp_path):
    ncol, nlayer, nsteps, world = 600, 24, 5, 2
    port = 29500 + (os.getpid() % 2000)
    mp.start_processes(_worker, args=(world, port, ncol, nlayer, nsteps, str(tmp_path)), nprocs=world,
                       join=True, start_method="spawn")
    # single-process run of the whole column set
    lib = w.oracle_library()
    wl = w.coupled_workload(ncol=ncol, nlayer=nlayer, seed=77, top=(w.F, 0.0, w.F, 0.0), bottom=(w.F, 0.0, w.F, 0.0))
    ctx = lh.SoilContext(lib, wl.config())
    wl.upload(ctx)
    W0 = ctx.budgets()
    ctx.step(0.0, wl.dt, nsteps)
    full = ctx.get_state(0)
    parts = [np.load(tmp_path / f"state_{r}.npy") for r in range(world)]
    assert np.array_equal(np.concatenate(parts), full)          # no halo: shards evolve exactly like the whole
    b = [np.load(tmp_path / f"budget_{r}.npy") for r in range(world)]
    assert np.array_equal(b[0][2:], b[1][2:])                   # every rank sees the same global budget
    assert np.allclose(b[0][2:], b[0][:2] + b[1][:2], rtol=1e-15)
    assert np.allclose(b[0][2:], ctx.budgets(), rtol=1e-13)
    assert np.allclose(b[0][2:], W0, rtol=1e-12)                # zero-flux BCs: conserved

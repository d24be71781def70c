"""bench.py's N > 1 control flow on CPU: world_size-2 and -3 ``gloo`` processes run bench._main to the end.

Round 2's first 8-GPU run hung at N = 2 and N = 4 (rank-asymmetric teardown of two NCCL communicators).  bench.py's multi-rank
path cannot run without GPUs, but every collective it issues, in the order it issues them, can: this test swaps the device-only
pieces in the TEST process (torch.cuda calls -> no-ops, the CUDA library -> the CPU checker library, the clock sampler -> a
stub) and lets the unmodified control flow of bench._main run under gloo.  Any rank-asymmetric collective, early return or
exception shows up as a hang (the test has a timeout) or a missing result line.  The numbers it prints are meaningless and are
not looked at."""
import json
import os
import sys

import pytest
import torch.multiprocessing as mp

import workloads as w


def _worker(rank, world, port, out_dir, extra):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import types

    import torch
    import torch.distributed as dist

    sys.path.insert(0, w.ROOT if hasattr(w, "ROOT") else os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench

    lh = w.lh
    oracle = w.oracle_library()
    # device-only pieces -> CPU stand-ins (TEST process only; nothing in the product or in bench.py routes here)
    bench.DEVICE = "cpu"
    bench.EXIT_HARD = False
    torch.cuda.is_available = lambda: True
    torch.cuda.set_device = lambda *_a, **_k: None
    torch.cuda.synchronize = lambda *_a, **_k: None
    torch.cuda.get_device_properties = lambda *_a, **_k: types.SimpleNamespace(multi_processor_count=148)
    lh.cuda_library = lambda: oracle

    class FakeSampler:
        def __init__(self, device):
            pass

        def start(self):
            pass

        def stop(self, t0, t1):
            return {"sm_mhz": 1500.0, "sm_min_mhz": 1500.0, "sm_max_mhz": 1965.0, "power_w_median": 0.0, "reasons": [], "samples": 3}

    bench.ClockSampler = FakeSampler
    if "--hang-e2e-on-rank-1" in extra:                     # the e2e leg of one rank never returns: the deadline path must still print
        extra = [x for x in extra if x != "--hang-e2e-on-rank-1"]
        bench.E2E_DEADLINE_S = 3.0
        if rank == 1:
            import time

            lh.SoilContext.run = lambda *_a, **_k: time.sleep(3600)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    argv = ["bench.py", "--gpus", str(world), "--steps", "2", "--warmup", "1", "--ncol", "768", "--nlayer", "12", "--min-seconds", "0.01"] + extra
    sys.argv = argv
    args = bench.parse_args()
    with open(os.path.join(out_dir, f"out_{rank}.txt"), "w") as stream:
        rc = bench._main(args, stream)
    assert rc == 0
    dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world,extra", [(2, []), (3, ["--model", "richards"]), (2, ["--no-e2e", "--general-vg"]),
                                         (1, ["--no-variants", "--no-cpu-baseline", "--ice"]), (1, ["--no-variants", "--no-e2e"])])
def test_bench_multi_rank_control_flow_terminates(tmp_path, world, extra):
    port = 29500 + (os.getpid() % 2000) + world
    mp.start_processes(_worker, args=(world, port, str(tmp_path), extra), nprocs=world, join=True, start_method="spawn")
    lines = [l for l in open(tmp_path / "out_0.txt").read().splitlines() if l.strip()]
    assert len(lines) == 1                                   # rank 0 prints exactly one line
    d = json.loads(lines[0])
    assert d["n_gpus"] == world and d["steps"] == 2 and d["scaling"] == "strong" and d["sustained"]["blocks"] >= 3
    assert ("e2e" in d) == ("--no-e2e" not in extra)
    assert ("cpu_baseline" in d) == (world == 1 and "--no-cpu-baseline" not in extra) and "extra" not in d      # CPU leg at N = 1 only
    for r in range(1, world):
        assert open(tmp_path / f"out_{r}.txt").read().strip() == ""    # the other ranks print nothing


@pytest.mark.timeout(300)
def test_bench_e2e_deadline_still_prints_the_line(tmp_path):
    """One rank's e2e leg never returns (a stuck transfer, a shard thread that deadlocks): every rank gives up at the deadline,
    rank 0 prints the complete line with the failure recorded under `e2e`, and all processes exit 0 (a non-zero rank would make
    torchrun kill rank 0)."""
    world = 2
    port = 29500 + (os.getpid() % 2000) + 7
    mp.start_processes(_worker, args=(world, port, str(tmp_path), ["--hang-e2e-on-rank-1"]), nprocs=world, join=True, start_method="spawn")
    lines = [l for l in open(tmp_path / "out_0.txt").read().splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["n_gpus"] == 2 and d["value"] > 0 and d["sustained"]["blocks"] >= 3 and "roofline" in d
    assert d["e2e"]["value"] is None and "did not finish" in d["e2e"]["error"]

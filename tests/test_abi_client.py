"""A non-Python client of the C ABI (tests/support/abi_client.c): plain C, includes only include/lh_soil.h, dlopens the
library and drives create -> set_state (reference layout, batched columns) -> rhs -> step -> run (snapshots, budgets,
checkpoint / restart) -> get_state -> budgets.  CPU: built and run against the oracle (the header alone is sufficient to
write a client).  GPU: run against the CUDA library and compared with the oracle's output file."""
import os
import subprocess

import numpy as np
import pytest

import __graft_entry__ as graft

ROOT = graft.ROOT
SRC = os.path.join(ROOT, "tests", "support", "abi_client.c")
NCOL, NLAYER = 48, 20


@pytest.fixture(scope="module")
def client(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("abi_client") / "abi_client")
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    subprocess.run(["gcc", "-std=c11", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
                    "-ldl", "-lm"], check=True, env=env)
    return exe


def _run(client, lib, prefix, out, ncol=NCOL, nlayer=NLAYER):
    r = subprocess.run([client, lib, prefix, out, str(ncol), str(nlayer)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "abi_client ok" in r.stdout
    return np.fromfile(out, dtype=np.float64)


def _split(a, ncol=NCOL, nlayer=NLAYER):
    cells = ncol * nlayer
    k = 0
    out = {}
    for name, n in (("tendency", 2 * cells), ("state5", 2 * cells), ("snaps", 4 * 2 * cells), ("budgets", 12), ("final_budget", 2)):
        out[name] = a[k:k + n]
        k += n
    assert k == a.size
    return out


def test_client_includes_only_the_public_header():
    src = open(SRC).read()
    incs = [l.split()[1] for l in src.splitlines() if l.startswith("#include")]
    assert [i for i in incs if i.startswith('"')] == ['"lh_soil.h"']


def test_c_client_against_oracle(client, tmp_path):
    o = _split(_run(client, graft.build_oracle(), "lho_", str(tmp_path / "oracle.bin")))
    assert np.all(np.isfinite(o["tendency"])) and np.all(np.isfinite(o["snaps"]))
    assert np.any(o["tendency"] != 0.0)
    assert np.array_equal(o["budgets"][-2:], o["final_budget"])          # the last per-step budget is the final one
    cells = NCOL * NLAYER
    assert np.array_equal(o["snaps"][:2 * cells], o["state5"])           # snapshot 0 (save_first) is the state after the 5 steps


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(NCOL, NLAYER), (4099, 64)])
def test_c_client_cuda_matches_oracle(client, tmp_path, shape):
    ncol, nlayer = shape
    o = _split(_run(client, graft.build_oracle(), "lho_", str(tmp_path / "oracle.bin"), ncol, nlayer), ncol, nlayer)
    g = _split(_run(client, graft.load_package().cuda_library().path, "lh_", str(tmp_path / "cuda.bin"), ncol, nlayer), ncol, nlayer)
    cells = ncol * nlayer
    for f in range(2):       # tendencies: 1e-11 of the field's max norm (the strict, cancellation-aware gate is test_gpu_parity.py)
        r, a = o["tendency"][f * cells:(f + 1) * cells], g["tendency"][f * cells:(f + 1) * cells]
        assert np.max(np.abs(a - r)) <= 1e-11 * np.max(np.abs(r))
    for key in ("state5", "snaps"):
        r, a = o[key].reshape(-1, cells), g[key].reshape(-1, cells)
        for k in range(r.shape[0]):
            assert np.max(np.abs(a[k] - r[k])) <= 1e-10 * np.max(np.abs(r[k])), (key, k)
    assert np.max(np.abs(g["budgets"] - o["budgets"]) / np.abs(o["budgets"])) <= 1e-12

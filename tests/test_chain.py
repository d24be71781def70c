"""Block-to-block chaining of consecutive stage launches (LhKernelArgs::chain_flags; include/lh_soil.h
LH_FLAG_NO_CHAIN): block j of a stage waits for block j of the previous stage launch instead of the whole previous
grid.  Results must be bit-identical to the whole-grid dependency, for grids of many waves (where stages really
overlap), with interleaved uploads / downloads / budget reads / shape changes, and for the generic steppers."""
import numpy as np
import pytest

import workloads as w

pytestmark = pytest.mark.gpu
lh, abi = w.lh, w.abi

STAGE = abi.LH_FLAG_STAGE_LAUNCHES


def _pair(cuda, wl, extra=0):
    a = lh.SoilContext(cuda, wl.config(flags=STAGE | extra))
    b = lh.SoilContext(cuda, wl.config(flags=STAGE | abi.LH_FLAG_NO_CHAIN | extra))
    for ctx in (a, b):
        wl.upload(ctx)
    return a, b


CASES = {
    # ~12 waves of 4-warp blocks: the next stage's blocks start while the previous stage's last wave is still running
    "coupled_many_waves": lambda: w.coupled_workload(ncol=1 << 18, nlayer=64, seed=81),
    "coupled_shard_2p8_waves": lambda: w.coupled_workload(ncol=131072, nlayer=64, seed=82),
    "richards_many_waves": lambda: w.richards_workload(ncol=200000, nlayer=100, seed=83),
    "heat_ragged": lambda: w.heat_workload(ncol=70001, nlayer=37, seed=84),
    "coupled_ice": lambda: w.coupled_workload(ncol=50000, nlayer=33, seed=85, ice=True),
    "tiny": lambda: w.coupled_workload(ncol=3, nlayer=5, seed=86, zlim=(-0.5, 0.0)),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_chained_launches_are_bit_identical(cuda, name):
    wl = CASES[name]()
    a, b = _pair(cuda, wl)
    for ctx in (a, b):
        ctx.step(0.0, wl.dt, 7)
        bud = ctx.budgets()                         # a budget kernel between two chained calls
        ctx.step(7 * wl.dt, wl.dt, 3)
        assert np.all(np.isfinite(bud))
    for f in wl.fields:
        ra, rb = a.get_state(f), b.get_state(f)
        assert np.array_equal(ra, rb), (name, f, np.max(np.abs(ra - rb)))
    assert np.array_equal(a.budgets(), b.budgets())


def test_chain_survives_uploads_rhs_and_shape_changes(cuda):
    """Uploads (writers on the same stream), tendency launches, and a switch of kernel variant / launch shape
    (per-column parameters) between chained calls."""
    wl = w.coupled_workload(ncol=40000, nlayer=64, seed=87)
    a, b = _pair(cuda, wl)
    rng = np.random.default_rng(5)
    nu = wl.params.nu * rng.uniform(1.0, 1.1, wl.ncol)
    th2 = wl.fields[0] * 0.97
    for ctx in (a, b):
        ctx.step(0.0, wl.dt, 2)
        ctx.rhs(0.0)
        ctx.step(0.0, wl.dt, 1)
        ctx.set_state(0, th2)                       # overwrite the state between two chained launches
        ctx.step(0.0, wl.dt, 2)
        ctx.set_column_params(nu=nu)                # HET variant: another launch shape
        ctx.step(0.0, wl.dt, 2)
        ctx.set_column_params()                     # back
        ctx.step(0.0, wl.dt, 2)
    for f in wl.fields:
        assert np.array_equal(a.get_state(f), b.get_state(f)), f
    for f in (0, 2):
        assert np.array_equal(a.get_tendency(f), b.get_tendency(f))


@pytest.mark.parametrize("method", ["LH_METHOD_SSPRK43", "LH_METHOD_CK2N54"])
def test_generic_steppers_chain(cuda, method):
    wl = w.coupled_workload(ncol=60000, nlayer=40, seed=88)
    a, b = _pair(cuda, wl)
    st = abi.lh_soil_stepper()
    assert cuda.soil_stepper_named(getattr(abi, method), st) == abi.LH_OK
    for ctx in (a, b):
        ctx.step_with(st, 0.0, wl.dt, 4)
    for f in wl.fields:
        assert np.array_equal(a.get_state(f), b.get_state(f)), f


def test_single_stage_calls_chain(cuda, oracle):
    """lh_soil_stage_ssprk33 called stage by stage (the host path for time-dependent prescribed profiles) is chained
    too, and still matches the oracle."""
    wl = w.coupled_workload(ncol=9000, nlayer=64, seed=89)
    g = lh.SoilContext(cuda, wl.config(flags=STAGE))
    o = lh.SoilContext(oracle, wl.config())
    for ctx in (g, o):
        wl.upload(ctx)
        for _ in range(3):
            for s in (1, 2, 3):
                ctx.stage(s, wl.dt)
    for f in (0, 2):
        r = o.get_state(f)
        assert np.max(np.abs(g.get_state(f) - r)) <= 1e-10 * np.max(np.abs(r))

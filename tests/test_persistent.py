"""The persistent SSPRK33 launch (LH_FLAG_PERSISTENT: all 3 nsteps stages in one kernel, every block keeping
its columns) against the per-stage launches (LH_FLAG_STAGE_LAUNCHES): bit-identical states, with and without a
per-stage boundary-value table, for every model, tall and short columns, and the automatic choice."""
import numpy as np
import pytest

import workloads as w

pytestmark = pytest.mark.gpu
lh, abi = w.lh, w.abi


def _pair(cuda, wl):
    a = lh.SoilContext(cuda, wl.config(flags=abi.LH_FLAG_STAGE_LAUNCHES))
    b = lh.SoilContext(cuda, wl.config(flags=abi.LH_FLAG_PERSISTENT))
    for ctx in (a, b):
        wl.upload(ctx)
    return a, b


CASES = {
    "coupled": lambda: w.coupled_workload(ncol=300, nlayer=64, seed=61),
    "coupled_general_vg_ice": lambda: w.coupled_workload(ncol=96, nlayer=33, seed=62, ice=True,
                                                         viscosity=lh.TemperatureDependentViscosity(), impedance=lh.IceImpedance()),
    "coupled_tall": lambda: w.coupled_workload(ncol=40, nlayer=700, seed=63, zlim=(-20.0, 0.0)),
    "coupled_one_layer": lambda: w.coupled_workload(ncol=64, nlayer=1, seed=64, zlim=(-0.05, 0.0)),
    "richards": lambda: w.richards_workload(ncol=257, nlayer=100, seed=65),
    "richards_viscosity": lambda: w.richards_workload(ncol=64, nlayer=40, seed=66, viscosity=lh.TemperatureDependentViscosity()),
    "heat": lambda: w.heat_workload(ncol=100, nlayer=37, seed=67),
}


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("with_table", [False, True])
def test_persistent_is_bit_identical(cuda, name, with_table):
    wl = CASES[name]()
    a, b = _pair(cuda, wl)
    nsteps = 5
    table = None
    if with_table:
        rng = np.random.default_rng(3)
        base = np.array([wl.top[1], wl.top[3], wl.bottom[1], wl.bottom[3]])
        table = base * (1.0 + 1e-3 * rng.standard_normal((nsteps, 3, 4)))
    for ctx in (a, b):
        ctx.step(0.0, wl.dt, nsteps, table)
        ctx.step(nsteps * wl.dt, wl.dt, 2)          # a second call continues from the state (and bc values) left behind
    assert a.last_step_timing()[1] == 6 and b.last_step_timing()[1] == 1
    for f in wl.fields:
        ra, rb = a.get_state(f), b.get_state(f)
        assert np.array_equal(ra, rb), (name, f, np.max(np.abs(ra - rb)))
    assert np.array_equal(a.budgets(), b.budgets())


def test_automatic_choice_and_oracle(cuda, oracle):
    """Small grids take the persistent launch on their own; large ones keep one launch per stage."""
    small = w.coupled_workload(ncol=2048, nlayer=64, seed=71)
    g, o = lh.SoilContext(cuda, small.config()), lh.SoilContext(oracle, small.config())
    for ctx in (g, o):
        small.upload(ctx)
        ctx.step(0.0, small.dt, 4)
    assert g.last_step_timing()[1] == 1
    for f in (0, 2):
        r = o.get_state(f)
        assert np.max(np.abs(g.get_state(f) - r)) <= 1e-10 * np.max(np.abs(r))
    big = w.coupled_workload(ncol=1 << 19, nlayer=64, seed=72)
    gb = lh.SoilContext(cuda, big.config())
    big.upload(gb)
    gb.step(0.0, big.dt, 2)
    assert gb.last_step_timing()[1] == 6


def test_persistent_check_finite(cuda):
    wl = w.richards_workload(ncol=64, nlayer=20, seed=73)
    wl.fields[0][3, 5] = np.nan
    ctx = lh.SoilContext(cuda, wl.config(flags=abi.LH_FLAG_PERSISTENT | abi.LH_FLAG_CHECK_FINITE))
    wl.upload(ctx)
    with pytest.raises(lh.NonFiniteStateError):
        ctx.step(0.0, wl.dt, 3)


@pytest.mark.parametrize("flag", [abi.LH_FLAG_STAGE_LAUNCHES, abi.LH_FLAG_PERSISTENT])
def test_zero_steps_is_a_no_op(cuda, flag):
    wl = w.coupled_workload(ncol=40, nlayer=12, seed=74)
    ctx = lh.SoilContext(cuda, wl.config(flags=flag))
    wl.upload(ctx)
    before = {f: ctx.get_state(f) for f in wl.fields}
    b0 = ctx.budgets()
    ctx.step(0.0, wl.dt, 0)
    ctx.step(0.0, wl.dt, 0, np.zeros((0, 3, 4)))
    assert ctx.last_step_timing()[1] == 0
    for f, a in before.items():
        assert np.array_equal(ctx.get_state(f), a)
    assert np.array_equal(ctx.budgets(), b0)
    with pytest.raises(lh._abi.SoilError):
        ctx.lib.soil_step_ssprk33  # symbol exists
        ctx._check(ctx.lib.soil_step_ssprk33(ctx._h, 0.0, wl.dt, -1, None))

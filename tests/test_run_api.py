"""lh_soil_run (saveat snapshots + per-step budgets in one call), checkpoint / restart, prescribed-profile tables and the
non-blocking budget read (include/lh_soil.h), on the oracle (CPU: the semantics) and on the CUDA library (GPU: the same
results from the overlapped implementation, and the whole saved history against the oracle).

Reference seams: Simulation(...; saveat, callback) simulation.jl:64-70, used by richards_equation.jl:66-78 and
coupled.jl:94-98; make_update_aux right_hand_side.jl:54-81."""
import numpy as np
import pytest

import workloads as w

lh, abi = w.lh, w.abi


@pytest.fixture(params=["oracle", pytest.param("cuda", marks=pytest.mark.gpu)])
def lib(request, oracle):
    return oracle if request.param == "oracle" else lh.cuda_library()


def _ctx(lib, wl, flags=0):
    ctx = lh.SoilContext(lib, wl.config(flags=flags))
    wl.upload(ctx)
    return ctx


def _table(wl, nsteps, seed=3):
    rng = np.random.default_rng(seed)
    base = np.array([wl.top[1], wl.top[3], wl.bottom[1], wl.bottom[3]])
    return base * (1.0 + 1e-3 * rng.standard_normal((nsteps, 3, 4)))


@pytest.mark.parametrize("case", ["coupled", "richards", "coupled_persistent"])
def test_run_equals_the_step_loop(lib, case):
    """One lh_soil_run call == the host loop step / budgets / get_state, bit for bit."""
    if case == "richards":
        wl, fields = w.richards_workload(ncol=70, nlayer=50, seed=101), (0,)
    else:
        wl, fields = w.coupled_workload(ncol=133, nlayer=24, seed=102), (0, 2)
    flags = abi.LH_FLAG_PERSISTENT if case == "coupled_persistent" else abi.LH_FLAG_STAGE_LAUNCHES
    nsteps, every_b, every_s = 13, 2, 3
    table = _table(wl, nsteps)
    a, b = _ctx(lib, wl, flags), _ctx(lib, wl, flags)
    budgets, snaps = a.run(0.0, wl.dt, nsteps, bc_table=table, budget_every=every_b, save_every=every_s, save_first=True,
                           save_fields=fields)
    assert budgets.shape == (nsteps // every_b, 2) and snaps.shape == (1 + nsteps // every_s, len(fields), wl.ncol, wl.nlayer)
    ref_b, ref_s = [], [[b.get_state(f) for f in fields]]
    for s in range(nsteps):
        b.step(s * wl.dt, wl.dt, 1, table[s:s + 1])
        if (s + 1) % every_b == 0:
            ref_b.append(b.budgets())
        if (s + 1) % every_s == 0:
            ref_s.append([b.get_state(f) for f in fields])
    assert np.array_equal(budgets, np.array(ref_b))
    assert np.array_equal(snaps, np.array(ref_s))
    for f in fields:
        assert np.array_equal(a.get_state(f), b.get_state(f))


def test_run_snapshots_of_a_one_layer_domain(lib):
    """nlayer == 1 with several columns: in the reference layout (col_stride = nlayer = 1, layer_stride = 1) a snapshot is one
    contiguous row of columns.  (Found by the random API sequences of tests/test_hostemu_fuzz.py: the overlapped snapshot path
    handed cudaMemcpy2DAsync a destination pitch of 8 bytes for a 79-column row and the run failed with `invalid argument`;
    the blocking lh_soil_get_state path had the single-layer case, the snapshot path did not.)"""
    wl = w.coupled_workload(ncol=79, nlayer=1, seed=104, zlim=(-0.1, 0.0))
    a, b = _ctx(lib, wl), _ctx(lib, wl)
    _, snaps = a.run(0.0, wl.dt, 4, save_every=2, save_first=True, save_fields=(0, 2))
    assert snaps.shape == (3, 2, 79, 1)
    ref = [[b.get_state(f) for f in (0, 2)]]
    for s in range(2):
        b.step(2 * s * wl.dt, wl.dt, 2)
        ref.append([b.get_state(f) for f in (0, 2)])
    assert np.array_equal(snaps, np.array(ref))


def test_run_snapshots_in_the_column_fastest_layout(lib):
    """Snapshots may also be written column-fastest, (col_stride, layer_stride) = (1, >= ncol) — the layout of the device
    arrays, a strided 2-D copy with no transpose.  Same numbers as the reference layout, transposed; the gap columns of a
    padded layer stride are left untouched."""
    import ctypes as C

    wl = w.coupled_workload(ncol=70, nlayer=9, seed=105, zlim=(-0.9, 0.0))
    a, b = _ctx(lib, wl), _ctx(lib, wl)
    _, ref = a.run(0.0, wl.dt, 4, save_every=2, save_first=True, save_fields=(0, 2))          # (3, 2, ncol, nlayer)
    pad = 75                                                                                   # layer stride > ncol
    out = np.full((3, 2, wl.nlayer, pad), -7.0)
    o = abi.lh_soil_run_opts()
    o.save_every, o.save_first, o.nsave_fields = 2, 1, 2
    o.save_fields[0], o.save_fields[1] = 0, 2
    o.save_out = out.ctypes.data_as(C.POINTER(C.c_double))
    o.snapshot_stride, o.field_stride, o.col_stride, o.layer_stride = 2 * wl.nlayer * pad, wl.nlayer * pad, 1, pad
    assert lib.soil_run(b._h, 0.0, wl.dt, 4, C.byref(o)) == abi.LH_OK
    assert np.array_equal(out[..., :70].transpose(0, 1, 3, 2), ref)
    assert np.all(out[..., 70:] == -7.0)


def test_host_buffers_allocated_by_the_library(lib):
    """lh_soil_alloc_host / lh_soil_free_host: page-locked host memory for hosts without an allocator of their own (a Julia
    caller); used here as upload source, download target and snapshot destination."""
    import ctypes as C

    wl = w.coupled_workload(ncol=50, nlayer=10, seed=106, zlim=(-1.0, 0.0))
    nbytes = 3 * 2 * 50 * 10 * 8
    p = C.c_void_p()
    assert lib.soil_alloc_host(nbytes, C.byref(p)) == abi.LH_OK and p.value
    try:
        buf = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(3, 2, 50, 10))
        a, b = _ctx(lib, wl), _ctx(lib, wl)
        buf[0, 0] = wl.fields[0] * 0.99
        for c in (a, b):
            c.set_state(0, buf[0, 0])                                # upload straight from the library's buffer
        _, ref = b.run(0.0, wl.dt, 4, save_every=2, save_first=True, save_fields=(0, 2))
        _, snaps = a.run(0.0, wl.dt, 4, save_every=2, save_first=True, save_fields=(0, 2), save_out=buf)
        assert snaps is buf and np.array_equal(buf, ref)
        a.get_state(2, out=buf[1, 1])
        assert np.array_equal(buf[1, 1], b.get_state(2))
    finally:
        assert lib.soil_free_host(p) == abi.LH_OK
    assert lib.soil_alloc_host(-1, C.byref(p)) == abi.LH_ERR_INVALID_ARG
    assert lib.soil_free_host(None) == abi.LH_OK


def test_run_argument_checks(lib):
    wl = w.coupled_workload(ncol=8, nlayer=6, seed=103, zlim=(-0.6, 0.0))
    ctx = _ctx(lib, wl)
    o = abi.lh_soil_run_opts()
    o.budget_every = 1                                   # no budgets_out
    assert lib.soil_run(ctx._h, 0.0, wl.dt, 2, o) == abi.LH_ERR_INVALID_ARG
    o = abi.lh_soil_run_opts()
    o.struct_size = 8
    assert lib.soil_run(ctx._h, 0.0, wl.dt, 2, o) == abi.LH_ERR_INVALID_ARG
    o = abi.lh_soil_run_opts()
    assert lib.soil_run(ctx._h, 0.0, wl.dt, -1, o) == abi.LH_ERR_INVALID_ARG
    assert lib.soil_run(ctx._h, 0.0, wl.dt, 0, o) == abi.LH_OK      # nothing to do
    b, s = ctx.run(0.0, wl.dt, 0, save_first=True, save_fields=(0,))
    assert b is None and s.shape[0] == 1 and np.array_equal(s[0, 0], wl.fields[0])


def test_checkpoint_restart_is_bit_identical(lib):
    """Run 9 steps; or run 4, checkpoint, destroy, restore into a NEW ctx, run 5: the same bits."""
    wl = w.coupled_workload(ncol=97, nlayer=31, seed=104, ice=True)
    table = _table(wl, 9)
    a = _ctx(lib, wl)
    a.step(0.0, wl.dt, 9, table)
    b = _ctx(lib, wl)
    b.step(0.0, wl.dt, 4, table[:4])
    blob = b.checkpoint()
    b.close()
    c = lh.SoilContext(lib, wl.config())                 # fresh: nothing uploaded
    c.restore(blob)
    c.step(4 * wl.dt, wl.dt, 5, table[4:])
    for f in wl.fields:
        assert np.array_equal(a.get_state(f), c.get_state(f)), f
    assert np.array_equal(a.budgets(), c.budgets())
    # a checkpoint of another problem is refused
    other = lh.SoilContext(lib, w.coupled_workload(ncol=96, nlayer=31, seed=1).config())
    with pytest.raises(lh._abi.SoilError):
        other.restore(blob)
    with pytest.raises(lh._abi.SoilError):
        c.restore(blob[: blob.size // 2])


def test_prescribed_profile_table_equals_stagewise_uploads(lib):
    """Time-dependent T(z, t) (Richards + viscosity): rows uploaded ahead == set_aux before every stage."""
    wl = w.richards_workload(ncol=40, nlayer=30, seed=105, viscosity=lh.TemperatureDependentViscosity())
    nsteps, dt = 6, wl.dt
    zc = w.zc_of(wl.zmin, wl.zmax, wl.nlayer)
    Tp = lambda t: 288.0 + 5.0 * zc + 0.5 * t
    rows = np.array([Tp(t0 + c * dt) for t0 in (s * dt for s in range(nsteps)) for c in (0.0, 1.0, 0.5)])
    a, b = _ctx(lib, wl, abi.LH_FLAG_STAGE_LAUNCHES), _ctx(lib, wl, abi.LH_FLAG_STAGE_LAUNCHES)
    a.set_aux_table(abi.LH_FIELD_T, rows)
    a.step(0.0, dt, nsteps)
    for s in range(nsteps):
        for stage, c in ((1, 0.0), (2, 1.0), (3, 0.5)):
            b.set_aux(abi.LH_FIELD_T, Tp(s * dt + c * dt), per_layer=True)
            b.stage(stage, dt)
    assert np.array_equal(a.get_state(0), b.get_state(0))
    with pytest.raises(lh._abi.SoilError):               # table exhausted
        a.step(nsteps * dt, dt, 1)
    a.set_aux_table(abi.LH_FIELD_T, None)                # back to the static field (last row stays)
    a.step(nsteps * dt, dt, 1)
    with pytest.raises(lh._abi.SoilError):               # theta_l is prognostic in the Richards model
        a.set_aux_table(abi.LH_FIELD_THETA_L, rows)


def test_prescribed_hydrology_table(lib):
    """Heat model with time-dependent prescribed ϑ_l(z, t) and θ_i(z, t) (PrescribedHydrologyModel)."""
    wl = w.heat_workload(ncol=24, nlayer=20, seed=106)
    nsteps, dt = 4, wl.dt
    zc = w.zc_of(wl.zmin, wl.zmax, wl.nlayer)
    th = lambda t: 0.2 + 0.1 * zc + 1e-4 * t
    ti = lambda t: 0.01 + 0.0 * zc + 1e-6 * t
    times = [s * dt + c * dt for s in range(nsteps) for c in (0.0, 1.0, 0.5)]
    a, b = _ctx(lib, wl, abi.LH_FLAG_STAGE_LAUNCHES), _ctx(lib, wl, abi.LH_FLAG_STAGE_LAUNCHES)
    a.set_aux_table(abi.LH_FIELD_THETA_L, np.array([th(t) for t in times]))
    a.set_aux_table(abi.LH_FIELD_THETA_I, np.array([ti(t) for t in times]))
    a.step(0.0, dt, nsteps)
    k = 0
    for s in range(nsteps):
        for stage in (1, 2, 3):
            b.set_aux(abi.LH_FIELD_THETA_L, th(times[k]), per_layer=True)
            b.set_aux(abi.LH_FIELD_THETA_I, ti(times[k]), per_layer=True)
            b.stage(stage, dt)
            k += 1
    assert np.array_equal(a.get_state(2), b.get_state(2))


def test_async_budgets(lib):
    wl = w.coupled_workload(ncol=64, nlayer=16, seed=107)
    a, b = _ctx(lib, wl), _ctx(lib, wl)
    tickets, ref = [], []
    for s in range(6):
        a.step(s * wl.dt, wl.dt, 1)
        tickets.append(a.budgets_async())                # collected later: the stream is never drained in between
        b.step(s * wl.dt, wl.dt, 1)
        ref.append(b.budgets())
    got = [a.budgets_wait(t) for t in tickets]
    assert np.array_equal(np.array(got), np.array(ref))
    with pytest.raises(lh._abi.SoilError):
        a.budgets_wait(tickets[0])                       # already collected
    for _ in range(8):
        a.budgets_async()
    with pytest.raises(lh._abi.SoilError):               # ring full
        a.budgets_async()


@pytest.mark.gpu
def test_saved_history_matches_oracle(cuda, oracle):
    """The whole sol.u history of a run (snapshots every 4 of 24 steps, budgets every step), CUDA vs oracle: 1e-10 on
    every snapshot (north_star: state after the configured number of steps), budgets to 1e-12 relative."""
    wl = w.coupled_workload(ncol=2500, nlayer=64, seed=108)
    nsteps = 24
    table = _table(wl, nsteps)
    out = {}
    for name, L in (("cuda", cuda), ("oracle", oracle)):
        ctx = _ctx(L, wl)
        out[name] = ctx.run(0.0, wl.dt, nsteps, bc_table=table, budget_every=1, save_every=4, save_first=True, save_fields=(0, 2))
    (bg, sg), (bo, so) = out["cuda"], out["oracle"]
    assert sg.shape == so.shape == (7, 2, wl.ncol, wl.nlayer)
    for k in range(sg.shape[0]):
        for f in range(2):
            assert np.max(np.abs(sg[k, f] - so[k, f])) <= 1e-10 * np.max(np.abs(so[k, f])), (k, f)
    assert np.max(np.abs(bg - bo) / np.abs(bo)) <= 1e-12


@pytest.mark.gpu
def test_run_overlaps_snapshots_with_steps(cuda):
    """Large enough that a blocking download per snapshot would dominate: the overlapped run must not be slower than
    stepping plus ONE snapshot's transfer time per snapshot (sanity bound, generous), and results equal the loop."""
    wl = w.coupled_workload(ncol=1 << 16, nlayer=64, seed=109)
    a, b = _ctx(cuda, wl), _ctx(cuda, wl)
    nsteps = 12
    _, snaps = a.run(0.0, wl.dt, nsteps, save_every=3, save_fields=(0, 2))
    ref = []
    for s in range(nsteps):
        b.step(s * wl.dt, wl.dt, 1)
        if (s + 1) % 3 == 0:
            ref.append([b.get_state(0), b.get_state(2)])
    assert np.array_equal(snaps, np.array(ref))

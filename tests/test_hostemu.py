"""The PRODUCT's own kernels and host logic, executed on the CPU (no GPU in this container).

tests/hostemu.py compiles landhydrology.jl_b200/csrc/*.cu — the sources nvcc compiles for sm_100a, unchanged — with g++ against
an emulated CUDA execution model (tests/support/hostemu/: one fiber per CUDA thread, real __syncthreads / __syncwarp / shuffle
rendezvous, cp.async commit groups, guard-paged and poisoned "device" memory).  This file

* runs the stage kernels of every model under three fiber schedules (ascending, descending, seeded random) and both cp.async
  completion models (at issue / as late as wait_group allows): a kernel whose barriers and wait_groups are sufficient computes
  the same BITS every time, and they must equal the oracle to the parity bar — the racecheck this pool's closed
  compute-sanitizer could not give;
* re-runs the `-m gpu` parity suite (the tests the driver runs on the B200) against the emulated build in a child pytest
  (LH_TEST_HOSTEMU=1, tests/conftest.py), minus the cases sized for a real GPU.

The emulated library is test infrastructure: nothing in the package, bench.py or __graft_entry__ can load it, and a GPU run
never uses it.  What it cannot show: timing, stream / event ordering (everything completes inside the enqueuing call), the
sm_100a code generation itself (SASS identity of refactors is checked separately, tools/sass_identity.py)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import hostemu
import workloads as w

lh, abi = w.lh, w.abi
STAGE, PERSIST = abi.LH_FLAG_STAGE_LAUNCHES, abi.LH_FLAG_PERSISTENT


@pytest.fixture(scope="module")
def emu():
    return hostemu.library(lh)


@pytest.fixture(scope="module")
def knobs():
    h = hostemu.controls()
    yield h
    h.lh_emu_set_schedule(0, 1)
    h.lh_emu_set_cp_async_lazy(0)
    h.lh_emu_set_sm_count(148)
    h.lh_emu_set_device_count(1)


F_ = abi.LH_BC_FLUX


def _atmos():
    ep = lh.EarthParameterSet()
    a = abi.lh_soil_atmos()
    a.u_atm, a.theta_atm, a.z_atm, a.theta_scale, a.rho_a_sfc, a.q_atm = 2.0, 288.0, 2.0, 290.0, 1.17, 0.006
    a.R_v, a.R_d, a.grav, a.cp_d, a.cp_v, a.LH_v0 = ep.R_v, ep.R_d, ep.grav, ep.cp_d, ep.cp_v, ep.LH_v0
    a.press_triple, a.T_triple, a.von_karman = ep.press_triple, ep.T_triple, ep.von_karman_const
    a.Pr_0, a.a_m, a.a_h = ep.Pr_0, ep.a_m, ep.a_h
    return a


def _visc():
    return lh.TemperatureDependentViscosity()


def _imp():
    return lh.IceImpedance()


CASES = {
    # (workload, flags): shapes chosen so that columns are cut into several chunks per column (the shared-memory face
    # exchange and its one __syncthreads), ragged column counts (padding lanes), one-layer and two-layer chunks
    "coupled_n2_chunks": lambda: (w.coupled_workload(ncol=70, nlayer=64, seed=11), STAGE),
    "coupled_general_ice": lambda: (w.coupled_workload(ncol=45, nlayer=33, seed=12, ice=True), STAGE | abi.LH_FLAG_GENERAL_VG),
    "coupled_ice_factors": lambda: (w.coupled_workload(ncol=33, nlayer=24, seed=13, ice=True, viscosity=_visc(), impedance=_imp()), STAGE),
    "richards_sand": lambda: (w.richards_workload(ncol=97, nlayer=100, seed=14), STAGE),
    "richards_viscosity": lambda: (w.richards_workload(ncol=40, nlayer=30, seed=15, viscosity=_visc()), STAGE),
    "heat": lambda: (w.heat_workload(ncol=64, nlayer=37, seed=16, ice=True), STAGE),
    "coupled_persistent": lambda: (w.coupled_workload(ncol=50, nlayer=40, seed=17), PERSIST),
    "coupled_one_layer": lambda: (w.coupled_workload(ncol=40, nlayer=1, seed=18, zlim=(-0.1, 0.0)), STAGE),
    "coupled_tall": lambda: (w.coupled_workload(ncol=32, nlayer=300, seed=19, zlim=(-6.0, 0.0)), STAGE),
}


def _run(lib, wl, flags, nsteps=2):
    ctx = lh.SoilContext(lib, wl.config(flags=flags))
    wl.upload(ctx)
    ctx.rhs(0.0)
    tend = {f: ctx.get_tendency(f) for f in (0, 2) if (f == 0 and wl.model != abi.LH_MODEL_HEAT) or (f == 2 and wl.model != abi.LH_MODEL_RICHARDS)}
    ctx.step(0.0, wl.dt, nsteps)
    state = {f: ctx.get_state(f) for f in tend}
    bud = ctx.budgets()
    info = ctx.kernel_info()
    ctx.close()
    return tend, state, bud, info


@pytest.mark.parametrize("name", sorted(CASES))
def test_kernels_are_schedule_and_copy_timing_independent_and_match_oracle(emu, knobs, oracle, name):
    wl, flags = CASES[name]()
    knobs.lh_emu_set_schedule(0, 1)
    knobs.lh_emu_set_cp_async_lazy(0)
    ref = _run(emu, wl, flags)
    for policy, seed, lazy in [(0, 1, 1), (1, 1, 0), (1, 1, 1), (2, 7, 0), (2, 8, 1), (2, 9, 1)]:
        knobs.lh_emu_set_schedule(policy, seed)
        knobs.lh_emu_set_cp_async_lazy(lazy)
        got = _run(emu, wl, flags)
        for a, b in zip(ref[:2], got[:2]):
            for f in a:
                assert np.array_equal(a[f], b[f]), (name, "schedule", policy, seed, "lazy cp.async", lazy, "field", f)
        assert np.array_equal(ref[2], got[2])
    knobs.lh_emu_set_schedule(0, 1)
    knobs.lh_emu_set_cp_async_lazy(0)
    # and the bits are the right ones: the parity bar of tests/test_gpu_parity.py against the oracle
    o = lh.SoilContext(oracle, wl.config())
    wl.upload(o)
    o.rhs(0.0)
    for f, a in ref[0].items():
        scale = w.tendency_scale(o, f)
        assert np.max(np.abs(a - o.get_tendency(f)) / scale[:, None]) <= 1e-12, (name, f, ref[3])
    o.step(0.0, wl.dt, 2)
    for f, a in ref[1].items():
        r = o.get_state(f)
        assert np.max(np.abs(a - r)) <= 1e-10 * np.max(np.abs(r)), (name, f)
    assert np.allclose(ref[2], o.budgets(), rtol=1e-12, atol=0.0)


def test_schedule_and_copy_timing_fuzz_detects_a_missing_syncwarp():
    """The test of the test for the kernel side: a MUTANT stage kernel without the __syncwarp() between cp.async.wait_group and
    the reads of the ring (a lane reads values another lane's copy delivered).  One thread order happens to work — the mutant
    would pass a plain parity run — but not all of them."""
    mlib = hostemu.build_mutant(
        "ring_read_without_syncwarp", "lh_stage_kernel.cuh",
        "        lh_cp_wait<RING_DEPTH - 2>();\n        __syncwarp();\n",
        "        lh_cp_wait<RING_DEPTH - 2>();\n")
    lib, ctl = lh.SoilLibrary(mlib, "lh_"), hostemu.controls(mlib)
    wl = w.coupled_workload(ncol=70, nlayer=64, seed=11)

    def run(policy, seed, lazy):
        ctl.lh_emu_set_schedule(policy, seed)
        ctl.lh_emu_set_cp_async_lazy(lazy)
        try:
            g = lh.SoilContext(lib, wl.config(flags=STAGE))
            wl.upload(g)
            g.step(0.0, wl.dt, 2)
            out = g.get_state(0)
            g.close()
            return out
        finally:
            ctl.lh_emu_set_schedule(0, 1)
            ctl.lh_emu_set_cp_async_lazy(0)

    runs = [run(p, s, z) for p, s, z in [(0, 1, 0), (0, 1, 1), (1, 1, 0), (1, 1, 1), (2, 7, 0), (2, 8, 1)]]
    assert any(not np.array_equal(runs[0], r, equal_nan=True) for r in runs[1:])


def test_launch_shapes_of_a_148_sm_device_are_the_ones_emulated(emu, knobs):
    """lh_choose_shape sees the 148 SMs of a B200 (so the chunking, block shapes and the persistent / per-stage choice are the
    GPU's), and a different SM count really changes the shape (the knob works)."""
    wl = w.coupled_workload(ncol=4096, nlayer=64, seed=3)
    knobs.lh_emu_set_sm_count(148)
    a = lh.SoilContext(emu, wl.config())
    knobs.lh_emu_set_sm_count(4)
    b = lh.SoilContext(emu, wl.config())
    knobs.lh_emu_set_sm_count(148)
    ia, ib = a.kernel_info(), b.kernel_info()
    assert "persistent" in ia and "W=10" in ia, ia               # 128 column groups on 148 SMs: cut into chunks, one launch per call
    assert "lh_soil_stage_kernel" in ib and "W=2," in ib, ib     # the same columns are several waves on a 4-SM device: long chunks, per-stage launches


def test_no_device_no_fallback(emu, knobs):
    """Without a device lh_soil_create fails with LH_ERR_NO_DEVICE — the emulated library reproduces the product's behaviour,
    it does not add a CPU path to it."""
    wl = w.coupled_workload(ncol=8, nlayer=8, seed=1)
    knobs.lh_emu_set_device_count(0)
    try:
        with pytest.raises(lh._abi.NoDeviceError):
            lh.SoilContext(emu, wl.config())
    finally:
        knobs.lh_emu_set_device_count(1)


def test_fresh_device_memory_is_poisoned_and_never_read(emu):
    """cudaMalloc'ed memory is NaN-filled here.  The stage buffer V, the tendency buffers and every scratch array are written
    before they are read, so a run that starts from freshly allocated buffers is finite everywhere — including the padding
    columns (70 -> 96), which LH_FLAG_CHECK_FINITE counts too."""
    wl = w.coupled_workload(ncol=70, nlayer=20, seed=5)
    ctx = lh.SoilContext(emu, wl.config(flags=STAGE | abi.LH_FLAG_CHECK_FINITE))
    wl.upload(ctx)
    ctx.rhs(0.0)
    ctx.step(0.0, wl.dt, 2)
    assert np.all(np.isfinite(ctx.get_state(0))) and np.all(np.isfinite(ctx.get_state(2)))
    assert np.all(np.isfinite(ctx.budgets()))


def _async_scenario(emu, knobs, mode, seed):
    """Uploads in three host layouts (dense, strided gather, column-fastest), a run with per-step budgets and overlapped
    snapshots, a ticketed budget read with steps enqueued behind it, per-column parameters, a checkpoint round trip, downloads."""
    knobs.lh_emu_set_async(mode, seed)
    try:
        wl = w.coupled_workload(ncol=300, nlayer=24, seed=7)
        g = lh.SoilContext(emu, wl.config(flags=STAGE))
        wl.upload(g)
        wide = np.zeros((300, 48))
        wide[:, ::2] = wl.fields[0]
        g.set_state(0, wide[:, ::2])                                  # layer stride 2: gathered through the pinned staging blocks
        soa = np.ascontiguousarray(wl.fields[2].T)
        g.set_state(2, soa.T)                                         # column-fastest host block: the 2-D copy path
        # snapshots into PINNED host memory (lh_soil_alloc_host): their D2H copies are truly asynchronous — a pageable numpy
        # destination would make every copy a synchronisation point and hide a missing dependency
        import ctypes as C

        pin = C.c_void_p()
        assert emu.soil_alloc_host(3 * 2 * 300 * 24 * 8, C.byref(pin)) == abi.LH_OK
        pinned = np.ctypeslib.as_array(C.cast(pin, C.POINTER(C.c_double)), shape=(3, 2, 300, 24))
        bud, _ = g.run(0.0, wl.dt, 5, budget_every=1, save_every=2, save_first=True, save_fields=(0, 2), save_out=pinned)
        snaps = pinned.copy()
        pinned[0, 0] = snaps[2, 0]
        g.set_state(0, pinned[0, 0])                                  # and an upload from pinned memory: read when the copy runs
        ticket = g.budgets_async()
        g.step(5 * wl.dt, wl.dt, 2)
        b2 = g.budgets_wait(ticket)
        g.set_column_params(Ksat=wl.params.Ksat * np.linspace(0.5, 2.0, 300))
        g.step(7 * wl.dt, wl.dt, 1)
        ck = g.checkpoint()
        g.step(8 * wl.dt, wl.dt, 2)
        after = g.get_state(0)
        g.restore(ck)
        g.step(8 * wl.dt, wl.dt, 2)
        out = (g.get_state(0), g.get_state(2), bud, snaps, b2, after, g.budgets())
        g.close()
        assert emu.soil_free_host(pin) == abi.LH_OK
        return out
    finally:
        knobs.lh_emu_set_async(0, 1)


def test_stream_and_event_dependencies_are_complete(emu, knobs, monkeypatch):
    """The host layer overlaps PCIe copies, layout transposes, snapshots and steps on two streams per context.  Under the
    emulator's asynchronous modes nothing runs until a synchronisation forces it, and then either only what that synchronisation
    needs (lazy) or a seeded random interleaving of everything runnable: a missing event wait or a staging block reused too early
    gives different bits.  LH_STAGE_BLOCK_BYTES = 4096 pushes ten staging blocks per field through the two-buffer pipelines."""
    monkeypatch.setenv("LH_STAGE_BLOCK_BYTES", "4096")                # read by the library when a ctx first needs its staging blocks
    ref = _async_scenario(emu, knobs, 0, 1)
    assert np.array_equal(ref[0], ref[5])                             # restart == uninterrupted, bit for bit
    for mode, seed in [(1, 1), (2, 1), (2, 2), (2, 3)]:
        got = _async_scenario(emu, knobs, mode, seed)
        for a, b in zip(ref, got):
            assert np.array_equal(a, b), (mode, seed)


def test_asynchronous_modes_detect_a_missing_event_wait(monkeypatch):
    """The test of the test: a MUTANT of the product in which the layout transpose of an upload no longer waits for its H2D copy
    (one cudaStreamWaitEvent removed from upload_field).  Synchronously executed it computes the right answer — exactly why such
    a bug survives ordinary tests; under deferred execution it does not."""
    mlib = hostemu.build_mutant(
        "upload_transpose_without_copy_wait", "lh_soil_api.cu",
        "        LH_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_copy[k], 0));\n        LH_CUDA(c, lh_launch_to_soa(",
        "        LH_CUDA(c, lh_launch_to_soa(")
    monkeypatch.setenv("LH_STAGE_BLOCK_BYTES", "4096")
    lib, ctl = lh.SoilLibrary(mlib, "lh_"), hostemu.controls(mlib)
    wl = w.coupled_workload(ncol=300, nlayer=24, seed=7)

    def run(mode, seed):
        ctl.lh_emu_set_async(mode, seed)
        try:
            g = lh.SoilContext(lib, wl.config())
            wl.upload(g)
            g.step(0.0, wl.dt, 2)
            out = g.get_state(0)
            g.close()
            return out
        finally:
            ctl.lh_emu_set_async(0, 1)

    ref = run(0, 1)
    o = lh.SoilContext(w.oracle_library(), wl.config())
    wl.upload(o)
    o.step(0.0, wl.dt, 2)
    assert np.max(np.abs(ref - o.get_state(0))) <= 1e-10 * np.max(np.abs(ref))       # the mutant passes a synchronous parity test
    assert not np.array_equal(run(1, 1), ref)                                         # lazy execution exposes it
    assert any(not np.array_equal(run(2, s), ref) for s in (1, 2, 3, 4))              # and so does random interleaving


def test_deferred_execution_is_real_and_the_raw_pointer_rule_holds(emu, knobs):
    """Negative control for the test above, and the documented rule of lh_soil_device_ptr (include/lh_soil.h): lh_soil_set_state
    returns when the host buffer is free, the layout transform may still be queued on the ctx stream.  In lazy mode the raw
    pointer really shows the old contents until lh_soil_sync — so the asynchronous modes do defer work."""
    import ctypes as C

    wl = w.coupled_workload(ncol=64, nlayer=8, seed=7, zlim=(-0.8, 0.0))
    knobs.lh_emu_set_async(1, 1)
    try:
        g = lh.SoilContext(emu, wl.config())
        g.set_state(2, wl.fields[2])
        ptr, ncp = g.device_ptr(2)
        view = lambda: np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(8, ncp))[:, :64].T.copy()
        assert not np.array_equal(view(), wl.fields[2])
        g.sync()
        assert np.array_equal(view(), wl.fields[2])
        g.close()
    finally:
        knobs.lh_emu_set_async(0, 1)


@pytest.mark.parametrize("op", ["rhs", "stages", "step", "stepper", "run", "colp_then_step", "checkpoint_then_step", "info"])
def test_first_call_after_an_ice_upload_uses_the_ice_kernels(emu, oracle, op):
    """The scan that answers "any ice?" after a theta_i upload is only enqueued by the upload; whichever call comes next and
    launches kernels (or reports / changes the variant) must wait for the answer first.  A ctx that has already stepped with the
    ice-free kernels gets ice uploaded, then `op` — the result must be the oracle's."""
    wl = w.coupled_workload(ncol=40, nlayer=12, seed=23, zlim=(-1.2, 0.0))
    g, o = lh.SoilContext(emu, wl.config(flags=STAGE)), lh.SoilContext(oracle, wl.config())
    ice = np.random.default_rng(3).uniform(0.0, 0.03, wl.fields[1].shape)
    for c in (g, o):
        wl.upload(c)
        c.step(0.0, wl.dt, 1)
    assert "!ICE" in g.kernel_info()
    th = np.minimum(o.get_state(0), 0.95 * (wl.params.nu - ice))
    for c in (g, o):
        c.set_state(1, ice)
        c.set_state(0, th)
    if op == "info":
        assert ":ICE" in g.kernel_info()
        return
    for c in (g, o):
        if op == "rhs":
            c.rhs(0.0)
        elif op == "stages":
            for k in (1, 2, 3):
                c.stage(k, wl.dt)
        elif op == "step":
            c.step(wl.dt, wl.dt, 2)
        elif op == "stepper":
            t = abi.lh_soil_stepper()
            assert c.lib.soil_stepper_named(abi.LH_METHOD_SSPRK43, t) == abi.LH_OK
            c.step_with(t, wl.dt, wl.dt, 2)
        elif op == "run":
            c.run(wl.dt, wl.dt, 3, budget_every=1)
        elif op == "colp_then_step":
            c.set_column_params(Ksat=wl.params.Ksat * np.linspace(0.5, 2.0, wl.ncol))
            c.step(wl.dt, wl.dt, 2)
        elif op == "checkpoint_then_step":
            c.restore(c.checkpoint())
            c.step(wl.dt, wl.dt, 2)
    if op == "rhs":
        for f in (0, 2):
            scale = w.tendency_scale(o, f)
            assert np.max(np.abs(g.get_tendency(f) - o.get_tendency(f)) / scale[:, None]) <= 1e-12, (op, f)
    for f in (0, 2):
        a, r = g.get_state(f), o.get_state(f)
        assert np.max(np.abs(a - r)) <= 1e-10 * np.max(np.abs(r)), (op, f)
    assert ":ICE" in g.kernel_info()


def test_destroy_frees_every_device_and_pinned_allocation(emu, knobs):
    """A ctx that has used every lazily allocated resource — staging blocks, tendency buffers, diagnostic scratch, per-column and
    per-cell parameter tables, flux fields, atmospheric flux arrays, the budget ring, the run history and snapshot buffers, a
    boundary table on the device, a profile table — returns every cudaMalloc / cudaMallocHost block, stream and event on lh_soil_destroy."""
    before, handles = knobs.lh_emu_live_allocations(), knobs.lh_emu_live_handles()
    for model in ("coupled", "richards", "heat"):
        if model == "coupled":
            wl = w.coupled_workload(ncol=40, nlayer=12, seed=41, zlim=(-1.2, 0.0), ice=True, top=(F_, -1.0, F_, -1e-8), bottom=(F_, 0.0, F_, 0.0))
        elif model == "richards":
            wl = w.richards_workload(ncol=40, nlayer=12, seed=42, zlim=(-1.2, 0.0), viscosity=_visc())
        else:
            wl = w.heat_workload(ncol=40, nlayer=12, seed=43, zlim=(0.0, 1.0))
        g = lh.SoilContext(emu, wl.config(flags=PERSIST))
        wl.upload(g)
        g.rhs(0.0)
        g.diagnostic(abi.LH_DIAG_K)
        fields = {"coupled": (0, 2), "richards": (0,), "heat": (2,)}[model]
        table = np.broadcast_to(np.array([wl.top[1], wl.top[3], wl.bottom[1], wl.bottom[3]]), (3, 3, 4)).copy()
        g.run(0.0, wl.dt, 3, bc_table=table, budget_every=1, save_every=1, save_first=True, save_fields=fields)
        g.budgets_wait(g.budgets_async())
        if model != "heat":
            g.set_column_params(Ksat=wl.params.Ksat * np.linspace(0.5, 2.0, 40))
            g.set_cell_params(vg_n=np.full((40, 12), 2.5))
        if model == "coupled":
            g.set_column_heat_params(rho_c_ds=np.full(40, wl.params.rho_c_ds))
            g.set_column_fluxes(top_energy=np.full(40, -1.0))
            g.set_atmos_forcing(_atmos())
        if model == "richards":
            g.set_aux_table(abi.LH_FIELD_T, np.full((3, 12), 290.0))
        g.step(0.0, wl.dt, 1)
        g.restore(g.checkpoint())
        g.close()
    assert knobs.lh_emu_live_allocations() == before
    assert knobs.lh_emu_live_handles() == handles                    # and every stream and event it created


def test_every_allocation_site_fails_cleanly_and_recovers(emu, knobs, oracle):
    """Out-of-memory injection: the n-th cudaMalloc / cudaMallocHost of a session fails, for every n a session reaches.  The
    failing call must return LH_ERR_CUDA naming the site (never crash, never leak), and the SAME call must succeed when repeated
    with memory available again — no half-allocated staging blocks or budget ring left behind — and give the oracle's numbers."""
    wl = w.coupled_workload(ncol=40, nlayer=12, seed=41, zlim=(-1.2, 0.0), ice=True)
    o = lh.SoilContext(oracle, wl.config())
    wl.upload(o)
    o.run(0.0, wl.dt, 2)
    o.set_column_params(Ksat=wl.params.Ksat * np.linspace(0.5, 2.0, 40))
    o.set_cell_params(vg_n=np.full((40, 12), 2.5))
    o.step(0.0, wl.dt, 1)
    ref = o.get_state(0)
    base, handles = knobs.lh_emu_live_allocations(), knobs.lh_emu_live_handles()

    def session(g):
        yield "upload", lambda: wl.upload(g)
        yield "rhs", lambda: g.rhs(0.0)
        yield "diag", lambda: g.diagnostic(abi.LH_DIAG_K)
        yield "run", lambda: g.run(0.0, wl.dt, 2, budget_every=1, save_every=1, save_fields=(0, 2))
        yield "async", lambda: g.budgets_wait(g.budgets_async())
        yield "params", lambda: (g.set_column_params(Ksat=wl.params.Ksat * np.linspace(0.5, 2.0, 40)), g.set_cell_params(vg_n=np.full((40, 12), 2.5)))
        yield "step", lambda: g.step(0.0, wl.dt, 1)

    failures = 0
    for n in range(80):
        knobs.lh_emu_fail_allocation_in(n)
        g = None
        try:
            try:
                g = lh.SoilContext(emu, wl.config())
            except lh._abi.SoilError as e:
                failures += 1
                assert "out of memory" in str(e)
                continue
            for name, call in session(g):
                try:
                    call()
                except lh._abi.SoilError as e:
                    failures += 1
                    assert "out of memory" in str(e), (n, name, e)
                    knobs.lh_emu_fail_allocation_in(-1)
                    if name == "params":                     # the column parameters were set before the per-cell ones failed
                        g.set_column_params()
                        g.set_cell_params()
                    if name in ("upload", "params"):
                        wl.upload(g)                         # a half-finished upload leaves the state to the caller: redo it
                        if name == "params":
                            g.run(0.0, wl.dt, 2)
                    call()                                   # the same call, with memory available again
            a = g.get_state(0)
            assert np.max(np.abs(a - ref)) <= 1e-10 * np.max(np.abs(ref)), n
        finally:
            knobs.lh_emu_fail_allocation_in(-1)
            if g is not None:
                g.close()
        assert knobs.lh_emu_live_allocations() == base and knobs.lh_emu_live_handles() == handles, n
    assert failures >= 25                                    # create: 12 sites, the session: 16 more


def test_contexts_on_concurrent_host_threads(emu):
    """bench.py's e2e leg and any multi-threaded host drive several contexts from several host threads at once (one ctx, one
    stream and one thread per column shard).  The library keeps per-process state (the per-variant shared-memory configuration
    cache, the NCCL loader) and per-thread state (the create-time error string): every shard must come out bit-identical to the
    same shard run alone."""
    from concurrent.futures import ThreadPoolExecutor

    wl = w.coupled_workload(ncol=256, nlayer=24, seed=31)
    cuts = [0, 64, 96, 192, 256]
    K = 3

    def shard(k):
        c0, c1 = cuts[k], cuts[k + 1]
        ctx = lh.SoilContext(emu, wl.config(ncol=c1 - c0, flags=STAGE if k % 2 else 0))
        for f, a in wl.fields.items():
            ctx.set_state(f, np.ascontiguousarray(a[c0:c1]))
        bud, _ = ctx.run(0.0, wl.dt, K, budget_every=1)
        out = {f: ctx.get_state(f) for f in (0, 2)}
        ctx.close()
        return out, bud

    alone = [shard(k) for k in range(4)]
    with ThreadPoolExecutor(4) as pool:
        together = list(pool.map(shard, range(4)))
    for (sa, ba), (st, bt) in zip(alone, together):
        assert np.array_equal(ba, bt)
        for f in sa:
            assert np.array_equal(sa[f], st[f])
    # and the shards are the columns of the unsharded run
    whole = lh.SoilContext(emu, wl.config())
    wl.upload(whole)
    whole.step(0.0, wl.dt, K)
    for f in (0, 2):
        assert np.array_equal(np.concatenate([t_[0][f] for t_ in together]), whole.get_state(f))


@pytest.mark.parametrize("world,ncol", [(2, 4096), (3, 1000)])
def test_column_shards_with_the_budget_allreduce(world, ncol):
    """tests/test_multi_gpu.py::test_two_gpu_sharded_run_matches_single without GPUs: host threads are the ranks, each with a ctx
    over its contiguous column range on the emulated build, and tests/support/hostemu/fake_nccl.cpp (soname libnccl.so.2, loaded
    first so that the product's own dlopen finds it) is the communicator — lh_soil_comm_unique_id / lh_soil_comm_init /
    lh_soil_budgets_allreduce of the unchanged product code.  In a child process: this one may have the real libnccl mapped."""
    import json

    script = os.path.join(w.ROOT, "tests", "support", "hostemu", "multirank_session.py")
    r = subprocess.run([sys.executable, script, str(world), str(ncol), "24", "4"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["shards_bit_identical"] and d["same_total_on_every_rank"] and d["total_is_sum_of_locals"], d
    assert d["total_vs_unsharded_rel"] <= 1e-13 and d["conserved_rel"] <= 1e-12, d        # zero-flux faces: budgets conserved


@pytest.mark.parametrize("model", ["coupled", "richards", "heat"])
def test_misuse_returns_a_status_and_leaves_the_ctx_usable(emu, model):
    """Out-of-range field / diagnostic / stage / method / function ids, NULL pointers, a negative step count, wrong struct sizes,
    a junk checkpoint, an unknown budget ticket, a bad rank: every call returns a non-zero status (never a crash: device memory is
    guard-paged here), and the ctx still steps afterwards."""
    import ctypes as C

    wl = {"coupled": lambda: w.coupled_workload(ncol=10, nlayer=6, seed=1, zlim=(-0.6, 0.0)),
          "richards": lambda: w.richards_workload(ncol=10, nlayer=6, seed=1, zlim=(-0.6, 0.0)),
          "heat": lambda: w.heat_workload(ncol=10, nlayer=6, seed=1, zlim=(0.0, 0.6))}[model]()
    g = lh.SoilContext(emu, wl.config())
    wl.upload(g)
    h, buf = g._h, np.zeros((10, 6))
    dp = C.POINTER(C.c_double)
    p = buf.ctypes.data_as(dp)
    vp = buf.ctypes.data_as(C.c_void_p)
    calls = []
    for f in (-5, -1, 4, 7, 100, 2 ** 31 - 1):
        calls += [lambda f=f: emu.soil_set_state(h, f, p, 6, 1), lambda f=f: emu.soil_get_state(h, f, p, 6, 1),
                  lambda f=f: emu.soil_get_tendency(h, f, p, 6, 1), lambda f=f: emu.soil_diagnostic(h, f, p, 6, 1),
                  lambda f=f: emu.soil_stage_ssprk33(h, f + 10, 1.0), lambda f=f: emu.soil_set_aux_table(h, f, p, 2),
                  lambda f=f: emu.soil_eval_math(h, f + 20, p, p, 4), lambda f=f: emu.soil_device_ptr(h, f, C.byref(C.c_void_p()), None),
                  lambda f=f: emu.soil_stepper_named(f + 10, C.byref(abi.lh_soil_stepper()))]
    bad_kind, bad_stages = abi.lh_soil_stepper(), abi.lh_soil_stepper()
    bad_kind.kind, bad_kind.nstages = 7, 2
    bad_stages.kind, bad_stages.nstages = abi.LH_STEPPER_SHU_OSHER, 99
    too_many, wrong_size = abi.lh_soil_run_opts(), abi.lh_soil_run_opts()
    too_many.save_every, too_many.nsave_fields = 1, 9
    wrong_size.struct_size = 3
    calls += [lambda: emu.soil_set_state(h, 0, None, 6, 1), lambda: emu.soil_get_state(h, 0, None, 6, 1),
              lambda: emu.soil_step_ssprk33(h, 0.0, 1.0, -3, None), lambda: emu.soil_budgets(h, None), lambda: emu.soil_run(h, 0.0, 1.0, 2, None),
              lambda: emu.soil_budgets_wait(h, 12345, p), lambda: emu.soil_checkpoint_save(h, vp, 8),
              lambda: emu.soil_checkpoint_load(h, vp, buf.nbytes), lambda: emu.soil_comm_init(h, 2, 5, (C.c_uint8 * 128)()),
              lambda: emu.soil_step(h, C.byref(bad_kind), 0.0, 1.0, 1, None), lambda: emu.soil_step(h, C.byref(bad_stages), 0.0, 1.0, 1, None),
              lambda: emu.soil_run(h, 0.0, 1.0, 2, C.byref(too_many)), lambda: emu.soil_run(h, 0.0, 1.0, 2, C.byref(wrong_size))]
    for k, call in enumerate(calls):
        assert call() != abi.LH_OK, k
    g.step(0.0, wl.dt, 1)
    f = 2 if model == "heat" else 0
    assert np.all(np.isfinite(g.get_state(f)))
    g.close()


# The driver's `-m gpu` suite against the emulated build.  Left out (by wall time on 8 host cores, not by outcome — every one of
# them passes on the emulated build when given the minutes): cases sized for a real GPU (>= 5e4 columns, the full C4 shape),
# the reference's long integrations (20 000 to 138 240 steps), and what needs torch.cuda, NCCL or nvidia-smi.
EMU_DESELECT = [
    "tests/test_gpu_parity.py::test_full_size_properties",
    "tests/test_persistent.py::test_automatic_choice_and_oracle",
    "tests/test_chain.py::test_chained_launches_are_bit_identical[coupled_many_waves]",
    "tests/test_chain.py::test_chained_launches_are_bit_identical[coupled_shard_2p8_waves]",
    "tests/test_chain.py::test_chained_launches_are_bit_identical[richards_many_waves]",
    "tests/test_chain.py::test_chained_launches_are_bit_identical[heat_ragged]",
    "tests/test_chain.py::test_chained_launches_are_bit_identical[coupled_ice]",
    "tests/test_chain.py::test_generic_steppers_chain[LH_METHOD_CK2N54]",
    "tests/test_run_api.py::test_run_overlaps_snapshots_with_steps",
    "tests/test_host_api.py::test_heat_simulation_time_dependent_dirichlet[cuda]",
    "tests/test_host_api.py::test_richards_simulation_step_and_run[cuda]",
    "tests/test_host_api.py::test_bench_b200_arm_json_line",
]
EMU_FILES = ["tests/test_gpu_parity.py", "tests/test_golden.py", "tests/test_steppers.py", "tests/test_column_params.py",
             "tests/test_persistent.py", "tests/test_chain.py", "tests/test_run_api.py", "tests/test_abi_client.py",
             "tests/test_host_api.py", "tests/test_prescribed_atmos_bc.py", "tests/test_closure_truth.py"]


@pytest.mark.timeout(1500)
def test_gpu_parity_suite_passes_on_the_emulated_build():
    hostemu.build()
    env = dict(os.environ, LH_TEST_HOSTEMU="1")
    cmd = [sys.executable, "-m", "pytest", *EMU_FILES, "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider", "--timeout", "300"]
    for d in EMU_DESELECT:
        cmd += ["--deselect", d]
    r = subprocess.run(cmd, cwd=w.ROOT, env=env, capture_output=True, text=True)
    tail = "\n".join((r.stdout + r.stderr).splitlines()[-40:])
    assert r.returncode == 0, tail
    last = [l for l in r.stdout.splitlines() if " passed" in l][-1]
    assert int(last.split(" passed")[0].split()[-1]) >= 250, last

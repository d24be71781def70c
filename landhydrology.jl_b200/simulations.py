"""``Simulation``, ``step!`` and ``run!``: mirror of reference src/Simulations/simulation.jl:11-87.

The reference wraps ``DiffEqBase.init(ODEProblem(make_rhs(model), Y, tspan, Ya), method; dt,
callback, kwargs...)`` and every caller uses ``SSPRK33()`` with a fixed ``dt``.
Here the integrator keeps the state resident on the GPU and each step is three launches of the
fused RHS+stage kernel (``lh_soil_step_ssprk33``); the host only evaluates the Dirichlet
closures for the stage times ``t, t+dt, t+dt/2`` ahead of the launch.

Python has no ``!`` in identifiers: ``step!`` is ``step_`` and ``run!`` is ``run_``.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import numpy as np

from . import _abi
from .engine import SoilEngine, current_library
from .models import AbstractModel, SoilModel
from .engine import _FIELD_ID
from .states import FieldVector, copy as copy_state, nf


class _Stepper:
    """An explicit low-storage Runge-Kutta method that runs as one fused RHS+stage launch per stage.

    ``table(lib)`` returns the ``lh_soil_stepper`` coefficient table (``None`` for SSPRK33, which has its own
    specialised stage kernels); ``c`` are the stage time offsets the host needs to evaluate the Dirichlet
    closures and prescribed profiles at ``t + c[i] dt``."""

    method_id = None
    c: Sequence[float] = ()

    def table(self, lib):
        tab = _abi.lh_soil_stepper()
        st = lib.soil_stepper_named(self.method_id, tab)
        if st != _abi.LH_OK:
            raise ValueError(f"lh_soil_stepper_named({self.method_id}) failed with status {st}")
        return tab

    @property
    def stages(self):
        return len(self.c)

    def __repr__(self):
        return f"{type(self).__name__}()"


class SSPRK33(_Stepper):
    """OrdinaryDiffEq's 3-stage, third-order SSP Runge-Kutta method (Shu-Osher form): the only stepper the
    reference's tests and experiment use."""

    method_id = _abi.LH_METHOD_SSPRK33
    c = (0.0, 1.0, 0.5)

    def table(self, lib):
        return None          # lh_soil_step_ssprk33: the specialised stage kernels


class Euler(_Stepper):
    method_id = _abi.LH_METHOD_EULER
    c = (0.0,)


class SSPRK22(_Stepper):
    method_id = _abi.LH_METHOD_SSPRK22
    c = (0.0, 1.0)


class SSPRK43(_Stepper):
    method_id = _abi.LH_METHOD_SSPRK43
    c = (0.0, 0.5, 1.0, 0.5)


class CarpenterKennedy2N54(_Stepper):
    """Imported (and marked "does not work") by the reference's test driver, test/runtests.jl:5-10."""

    method_id = _abi.LH_METHOD_CK2N54
    c = (0.0, 1432997174477.0 / 9575080441755.0, 2526269341429.0 / 6820363962896.0,
         2006345519317.0 / 3224310063776.0, 2802321613138.0 / 2924317926251.0)


class ShuOsherRK(_Stepper):
    """Any two-register Shu-Osher method: ``u_i = a[i] u^n + b[i] u_{i-1} + g[i] dt f(u_{i-1}, t + c[i] dt)``
    (e.g. OrdinaryDiffEq's SSPRK73 with its own coefficient table)."""

    kind = _abi.LH_STEPPER_SHU_OSHER

    def __init__(self, a, b, g, c):
        if not (len(a) == len(b) == len(g) == len(c)) or not 1 <= len(a) <= _abi.LH_MAX_STAGES:
            raise ValueError("a, b, g, c must have the same length, 1..LH_MAX_STAGES")
        self.a, self.b, self.g, self.c = map(lambda v: tuple(float(x) for x in v), (a, b, g, c))

    def table(self, lib):
        tab = _abi.lh_soil_stepper()
        tab.kind, tab.nstages = self.kind, len(self.c)
        for i in range(len(self.c)):
            tab.a[i], tab.b[i], tab.g[i], tab.c[i] = self.a[i], self.b[i], self.g[i], self.c[i]
        return tab


class LowStorageRK2N(ShuOsherRK):
    """Any Williamson 2N method: ``r = A[i] r + dt f(u, t + c[i] dt); u = u + B[i] r`` (``A[0] == 0``)."""

    kind = _abi.LH_STEPPER_2N

    def __init__(self, A, B, c):
        super().__init__(A, B, [0.0] * len(A), c)


class Solution:
    """``integrator.sol``: saved times ``t`` and states ``u``."""

    def __init__(self):
        self.t: List[float] = []
        self.u: List[FieldVector] = []


class Integrator:
    """Minimal DEIntegrator: ``u``, ``p`` (= Ya), ``t``, ``dt``, ``sol``."""

    def __init__(self, engine: SoilEngine, Y: FieldVector, Ya: FieldVector, tspan, dt, saveat, callback,
                 max_chunk: int, method: Optional[_Stepper] = None, callback_reupload: str = "if-modified"):
        self.engine = engine
        if callback_reupload not in ("always", "if-modified", "never"):
            raise ValueError("callback_reupload must be 'always', 'if-modified' or 'never'")
        self.reupload = callback_reupload
        self.method = method if method is not None else SSPRK33()
        self._table = self.method.table(engine.lib)
        self.u = copy_state(Y)   # DiffEqBase.init does not alias u0
        self.p = Ya
        self.t0, self.tf = float(tspan[0]), float(tspan[1])
        self.t = self.t0
        self.dt = float(dt)
        self.callback = callback
        self.iter = 0
        self.sol = Solution()
        self.max_chunk = int(max_chunk)
        if saveat is None:
            self._saveat = None  # DiffEq default: save every step
        elif np.isscalar(saveat):
            n = int(np.floor((self.tf - self.t0) / float(saveat) + 1e-9))
            self._saveat = [self.t0 + k * float(saveat) for k in range(n + 1)]
            if self._saveat[-1] < self.tf:
                self._saveat.append(self.tf)
        else:
            self._saveat = sorted(float(s) for s in saveat)
        self._next_save = 0
        self._device_fresh = True
        self._dynamic_aux = engine.has_time_dependent_aux(self.t0, self.dt)
        engine.upload(self.u)
        self._save_if_due(force_first=True)

    # -- saving ------------------------------------------------------------------------------------
    def _save_now(self):
        self.sync_host()
        self.sol.t.append(self.t)
        self.sol.u.append(copy_state(self.u))

    def _save_if_due(self, force_first=False):
        if self._saveat is None:
            self._save_now()
            return
        tol = 1e-9 * max(1.0, abs(self.dt))
        while self._next_save < len(self._saveat) and self._saveat[self._next_save] <= self.t + tol:
            self._save_now()
            self._next_save += 1

    def _steps_to_next_event(self) -> int:
        """Whole steps that can be fused into one device call before the next save point / end."""
        remaining = int(np.floor((self.tf - self.t) / self.dt + 1e-9))
        if self._saveat is None or self.callback is not None:
            return min(1, remaining)
        if self._next_save < len(self._saveat):
            to_save = int(np.ceil((self._saveat[self._next_save] - self.t) / self.dt - 1e-9))
            remaining = min(remaining, max(to_save, 1))
        return min(remaining, self.max_chunk)

    def sync_host(self):
        """Bring ``integrator.u`` up to date with the device state."""
        if not self._device_fresh:
            self.engine.download(self.u)
            self._device_fresh = True

    # -- stepping ----------------------------------------------------------------------------------
    def _bc_table(self, t0: float, nsteps: int, dt: float):
        eng, cs = self.engine, self.method.c
        if not eng.has_dirichlet():
            return None
        table = np.empty((nsteps, len(cs), 4))
        t = t0
        for s in range(nsteps):
            for i, ci in enumerate(cs):
                table[s, i] = eng.bc_values(t + ci * dt)
            t = t + dt
        return table

    def _upload_aux_tables(self, t0: float, nsteps: int, dt: float):
        """Time-dependent prescribed profiles (make_update_aux, right_hand_side.jl:54-81): evaluated AHEAD at every stage
        time of the coming ``nsteps`` steps and uploaded once (``lh_soil_set_aux_table``); the device broadcasts the next
        row before every stage launch, so there is no host round trip per stage."""
        eng, cs = self.engine, self.method.c
        times = []
        t = t0
        for _ in range(nsteps):
            times.extend(t + ci * dt for ci in cs)
            t = t + dt
        rows = {}
        for tk in times:
            for name, prof in eng.prescribed_profiles(tk).items():
                rows.setdefault(name, []).append(prof)
        for name, r in rows.items():
            eng.ctx.set_aux_table(_FIELD_ID[name], np.array(r))
        return rows

    def _clear_aux_tables(self, rows):
        for name, r in rows.items():
            self.engine.ctx.set_aux_table(_FIELD_ID[name], None)
            self.engine._aux_cache[name] = r[-1]            # what the field holds now
            getattr(self.p, self.engine.model.name)[name][...] = r[-1]

    def _advance(self, nsteps: int, dt: float):
        eng = self.engine
        rows = self._upload_aux_tables(self.t, nsteps, dt) if self._dynamic_aux else {}
        table = self._bc_table(self.t, nsteps, dt)
        try:
            if self._table is None:
                eng.ctx.step(self.t, dt, nsteps, table)
            else:
                eng.ctx.step_with(self._table, self.t, dt, nsteps, table)
        finally:
            if rows:
                self._clear_aux_tables(rows)
        t = self.t
        for _ in range(nsteps):
            t = t + dt
        self.t = t
        self.iter += nsteps
        self._device_fresh = False

    def step(self):
        """One time step (DiffEqBase.step!)."""
        if self.t >= self.tf - 1e-12 * max(1.0, abs(self.tf)):
            return
        dt = min(self.dt, self.tf - self.t)
        self._advance(1, dt)
        if abs(self.tf - self.t) < 1e-9 * max(1.0, abs(self.dt)):
            self.t = self.tf
        self._after_step()

    def _after_step(self):
        if self.callback is not None:
            self.sync_host()
            before = copy_state(self.u) if self.reupload == "if-modified" else None
            self.callback(self)
            # DiffEq callbacks commonly modify integrator.u: the device state is authoritative, so send it back
            # (always, or only when it differs from what was downloaded: ``callback_reupload``)
            if self.reupload == "always" or (before is not None and not (before == self.u)):
                self.engine.upload(self.u)
        self._save_if_due()

    # -- run!(sim) as ONE device call --------------------------------------------------------------------------
    def _uniform_save_cadence(self):
        """Steps between save points if ``saveat`` is a whole multiple of dt reaching tf exactly, else ``None``."""
        nsteps = (self.tf - self.t) / self.dt
        if abs(nsteps - round(nsteps)) > 1e-9 * max(1.0, nsteps) or round(nsteps) < 1:
            return None
        nsteps = int(round(nsteps))
        if self._saveat is None:
            return nsteps, 1
        pend = self._saveat[self._next_save:]
        if not pend:
            return None
        every = (pend[0] - self.t) / self.dt
        if abs(every - round(every)) > 1e-9 * max(1.0, every) or round(every) < 1:
            return None
        every = int(round(every))
        expect = [self.t + (k + 1) * every * self.dt for k in range(nsteps // every)]
        if nsteps % every != 0 or len(pend) != len(expect) or any(abs(a - b) > 1e-9 * max(1.0, abs(self.dt)) for a, b in zip(pend, expect)):
            return None
        return nsteps, every

    def _solve_in_one_call(self, nsteps: int, every: int):
        """``lh_soil_run``: all steps, the ``saveat`` snapshots leaving the device on the copy stream while the steps go on
        (include/lh_soil.h); in blocks of at most ``max_chunk`` steps so that the host buffers stay bounded."""
        eng = self.engine
        names = [nf(n) for n in eng.model.prognostic_names]
        fields = [_FIELD_ID[n] for n in names]
        block = max(every, (self.max_chunk // every) * every)
        done = 0
        while done < nsteps:
            n = min(block, nsteps - done)
            rows = self._upload_aux_tables(self.t, n, self.dt) if self._dynamic_aux else {}
            try:
                _, snaps = eng.ctx.run(self.t, self.dt, n, bc_table=self._bc_table(self.t, n, self.dt), save_every=every,
                                       save_fields=fields)
            finally:
                if rows:
                    self._clear_aux_tables(rows)
            t = self.t
            for k in range(n):
                t = t + self.dt
                if (k + 1) % every == 0:
                    u = copy_state(self.u)
                    soil = getattr(u, eng.model.name)
                    for j, name in enumerate(eng.model.prognostic_names):
                        soil[name][...] = snaps[(k + 1) // every - 1, j].reshape(soil[name].shape)
                    self.sol.t.append(t)
                    self.sol.u.append(u)
                    if self._saveat is not None:
                        self._next_save += 1
            self.t = t
            self.iter += n
            self._device_fresh = False
            done += n
        if abs(self.tf - self.t) < 1e-9 * max(1.0, abs(self.dt)):
            self.t = self.tf
            if self.sol.t:
                self.sol.t[-1] = self.t

    def solve(self):
        """Run to ``tspan[2]`` (DiffEqBase.solve!)."""
        tol = 1e-9 * max(1.0, abs(self.dt))
        if self.callback is None and self._table is None and self.engine.column_range == (0, self.engine.model.domain.ncolumns):
            cadence = self._uniform_save_cadence()
            if cadence is not None:
                self._solve_in_one_call(*cadence)
        while self.tf - self.t > tol:
            n = self._steps_to_next_event()
            if n >= 1:
                self._advance(n, self.dt)
            else:  # last, shortened step so the integration ends exactly at tf
                self._advance(1, self.tf - self.t)
            if abs(self.tf - self.t) < tol:
                self.t = self.tf
            self._after_step()
        self.sync_host()
        if not self.sol.t or self.sol.t[-1] != self.t:
            self._save_now()
        return self.sol


class AbstractSimulation:
    pass


class Simulation(AbstractSimulation):
    """``Simulation(model, method; Y_init, dt, tspan, Ya_init, callbacks = nothing, kwargs...)``.

    ``kwargs`` understood: ``saveat`` (scalar spacing or list of times); ``progress`` /
    ``progress_message`` are accepted and ignored; ``max_steps_per_call`` bounds how many steps are
    fused into one ``lh_soil_step_ssprk33`` call; ``column_params`` (new) = per-column ``ν``, ``θr``, ``n``, ``α``,
    ``Ksat`` (``SoilEngine.set_column_params``) and ``ρc_ds``, ``κ_sat_unfrozen``, ``κ_sat_frozen``, ``κ_solid``, ``ν_ss_om``,
    ``ν_ss_quartz``, ``ν_ss_gravel`` (``SoilEngine.set_column_heat_params``) arrays for heterogeneous soils;
    ``callback_reupload`` = "always" | "if-modified" | "never": whether ``integrator.u`` goes back to the device after a callback.
    """

    def __init__(self, model: AbstractModel, method, *, Y_init, dt, tspan, Ya_init, callbacks=None,
                 saveat=None, progress=False, progress_message=None, max_steps_per_call: int = 4096,
                 device: int = 0, check_finite: bool = False, column_params: Optional[dict] = None,
                 callback_reupload: str = "if-modified"):
        if not isinstance(method, _Stepper):
            raise NotImplementedError(
                "the B200 path runs explicit low-storage methods: SSPRK33 (the only stepper the reference uses), Euler, "
                "SSPRK22, SSPRK43, CarpenterKennedy2N54, or a ShuOsherRK / LowStorageRK2N coefficient table")
        if not isinstance(model, SoilModel) or model.kind is None:
            raise TypeError("Simulation needs a SoilModel with at least one dynamic component")
        if Y_init is None:
            # simulation.jl:46-51: the reference's default-state branch references an undefined
            # variable (`soil_model`) and always throws; mirrored as an error here.
            raise NameError("soil_model not defined (reference simulation.jl:50): pass Y_init and Ya_init")
        self.model = model
        self.callbacks = callbacks
        engine = SoilEngine(model, float(tspan[0]), device=device, library=current_library(),
                            check_finite=check_finite)
        if column_params:                                   # heterogeneous soils
            hyd = {k: v for k, v in column_params.items() if nf(k) in {nf(x) for x in ("ν", "θr", "n", "α", "Ksat")}}
            heat = {k: v for k, v in column_params.items() if k not in hyd}
            if hyd:
                engine.set_column_params(**hyd)             # ν, θr, n, α, Ksat per column
            if heat:
                engine.set_column_heat_params(**heat)       # ρc_ds, κ_sat_unfrozen, κ_sat_frozen, κ_solid, ν_ss_* per column
        self.integrator = Integrator(engine, Y_init, Ya_init, tspan, dt, saveat, callbacks, max_steps_per_call, method,
                                     callback_reupload=callback_reupload)


def step_(simulation: AbstractSimulation) -> None:
    """``step!(simulation)`` (simulation.jl:79-80); returns ``None`` like the reference."""
    simulation.integrator.step()
    return None


def run_(simulation: AbstractSimulation):
    """``run!(simulation)`` (simulation.jl:86-87)."""
    return simulation.integrator.solve()

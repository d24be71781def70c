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
from .states import FieldVector, copy as copy_state


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
                 max_chunk: int, method: Optional[_Stepper] = None):
        self.engine = engine
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
        self._dynamic_aux = engine.has_time_dependent_aux(self.t0, self.dt) and self._table is None
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
    def _advance(self, nsteps: int, dt: float):
        eng = self.engine
        cs = self.method.c
        if self._dynamic_aux:
            # time-dependent prescribed profiles: stage by stage so update_aux! sees every stage time
            t = self.t
            for _ in range(nsteps):
                for stage, ts in ((1, t), (2, t + dt), (3, t + 0.5 * dt)):
                    eng.update_aux(ts, self.p)
                    eng.ctx.set_bc_values(eng.bc_values(ts))
                    eng.ctx.stage(stage, dt)
                t = t + dt
            self.t = t
        else:
            table = None
            t = self.t
            if eng.has_dirichlet():
                table = np.empty((nsteps, len(cs), 4))
                for s in range(nsteps):
                    for i, ci in enumerate(cs):
                        table[s, i] = eng.bc_values(t + ci * dt)
                    t = t + dt
            else:
                for _ in range(nsteps):
                    t = t + dt
            if self._table is None:
                eng.ctx.step(self.t, dt, nsteps, table)
            else:
                eng.ctx.step_with(self._table, self.t, dt, nsteps, table)
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
            self.callback(self)
        self._save_if_due()

    def solve(self):
        """Run to ``tspan[2]`` (DiffEqBase.solve!)."""
        tol = 1e-9 * max(1.0, abs(self.dt))
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
    ``Ksat`` arrays for heterogeneous soils (``SoilEngine.set_column_params``).
    """

    def __init__(self, model: AbstractModel, method, *, Y_init, dt, tspan, Ya_init, callbacks=None,
                 saveat=None, progress=False, progress_message=None, max_steps_per_call: int = 4096,
                 device: int = 0, check_finite: bool = False, column_params: Optional[dict] = None):
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
        if column_params:
            engine.set_column_params(**column_params)       # heterogeneous soils: ν, θr, n, α, Ksat per column
        if engine.has_time_dependent_aux(float(tspan[0]), float(dt)) and not isinstance(method, SSPRK33):
            raise NotImplementedError("time-dependent prescribed profiles are streamed stage by stage for SSPRK33 only")
        self.integrator = Integrator(engine, Y_init, Ya_init, tspan, dt, saveat, callbacks, max_steps_per_call, method)


def step_(simulation: AbstractSimulation) -> None:
    """``step!(simulation)`` (simulation.jl:79-80); returns ``None`` like the reference."""
    simulation.integrator.step()
    return None


def run_(simulation: AbstractSimulation):
    """``run!(simulation)`` (simulation.jl:86-87)."""
    return simulation.integrator.solve()

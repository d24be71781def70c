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

from .engine import SoilEngine, current_library
from .models import AbstractModel, SoilModel
from .states import FieldVector, copy as copy_state


class SSPRK33:
    """OrdinaryDiffEq's 3-stage, third-order SSP Runge-Kutta method (Shu-Osher form)."""

    stages = 3

    def __repr__(self):
        return "SSPRK33()"


class Solution:
    """``integrator.sol``: saved times ``t`` and states ``u``."""

    def __init__(self):
        self.t: List[float] = []
        self.u: List[FieldVector] = []


class Integrator:
    """Minimal DEIntegrator: ``u``, ``p`` (= Ya), ``t``, ``dt``, ``sol``."""

    def __init__(self, engine: SoilEngine, Y: FieldVector, Ya: FieldVector, tspan, dt, saveat, callback,
                 max_chunk: int):
        self.engine = engine
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
    def _advance(self, nsteps: int, dt: float):
        eng = self.engine
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
                table = np.empty((nsteps, 3, 4))
                for s in range(nsteps):
                    table[s, 0] = eng.bc_values(t)
                    table[s, 1] = eng.bc_values(t + dt)
                    table[s, 2] = eng.bc_values(t + 0.5 * dt)
                    t = t + dt
            else:
                for _ in range(nsteps):
                    t = t + dt
            eng.ctx.step(self.t, dt, nsteps, table)
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
    fused into one ``lh_soil_step_ssprk33`` call.
    """

    def __init__(self, model: AbstractModel, method, *, Y_init, dt, tspan, Ya_init, callbacks=None,
                 saveat=None, progress=False, progress_message=None, max_steps_per_call: int = 4096,
                 device: int = 0, check_finite: bool = False):
        if not isinstance(method, SSPRK33):
            raise NotImplementedError("only SSPRK33() is on the B200 path (the only stepper the reference uses)")
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
        self.integrator = Integrator(engine, Y_init, Ya_init, tspan, dt, saveat, callbacks, max_steps_per_call)


def step_(simulation: AbstractSimulation) -> None:
    """``step!(simulation)`` (simulation.jl:79-80); returns ``None`` like the reference."""
    simulation.integrator.step()
    return None


def run_(simulation: AbstractSimulation):
    """``run!(simulation)`` (simulation.jl:86-87)."""
    return simulation.integrator.solve()

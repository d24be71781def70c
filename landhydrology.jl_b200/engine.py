"""SoilEngine: the device-resident twin of one ``SoilModel`` (one column shard on one GPU).

It translates the reference's model/BC/parameter objects into an ``lh_soil_config``, owns the
``lh_soil_ctx`` and moves ``FieldVector`` states across the C ABI.  ``make_rhs`` and
``Simulation`` sit on top of it.  The library it talks to is the CUDA library; tests may inject
another :class:`SoilLibrary` (the CPU oracle) through ``use_library`` to check the host logic
without a GPU — the package itself never does.
"""
from __future__ import annotations

import contextlib
from typing import Optional, Sequence

import numpy as np

from . import _abi
from ._abi import SoilContext, SoilLibrary, lh_soil_config
from .models import (
    Dirichlet,
    FreeDrainage,
    NoBC,
    PrescribedAtmosForcing,
    PrescribedHydrologyModel,
    PrescribedTemperatureModel,
    SoilComponentBC,
    SoilHydrologyModel,
    SoilModel,
    VerticalFlux,
)
from .parameterizations import IceImpedance, NoEffect, TemperatureDependentViscosity
from .states import FieldVector, nf

_injected_library: Optional[SoilLibrary] = None


def current_library() -> SoilLibrary:
    return _injected_library if _injected_library is not None else _abi.cuda_library()


@contextlib.contextmanager
def use_library(lib: SoilLibrary):
    """Route new engines to ``lib`` inside the block (test hook)."""
    global _injected_library
    prev, _injected_library = _injected_library, lib
    try:
        yield lib
    finally:
        _injected_library = prev


_FIELD_ID = {nf("ϑ_l"): _abi.LH_FIELD_THETA_L, nf("θ_i"): _abi.LH_FIELD_THETA_I,
             nf("ρe_int"): _abi.LH_FIELD_RHO_E_INT, "T": _abi.LH_FIELD_T}


def _bc_kind_value(bc, t0: float):
    if isinstance(bc, NoBC):
        return _abi.LH_BC_NONE, 0.0
    if isinstance(bc, VerticalFlux):
        return _abi.LH_BC_FLUX, float(bc.flux)
    if isinstance(bc, Dirichlet):
        return _abi.LH_BC_DIRICHLET, float(bc.state_value(t0))
    if isinstance(bc, FreeDrainage):
        return _abi.LH_BC_FREE_DRAINAGE, 0.0
    raise TypeError(f"not an AbstractBC: {bc!r}")


def build_params(model: SoilModel) -> _abi.lh_soil_params:
    sp, ep = model.soil_param_set, model.earth_param_set
    p = _abi.lh_soil_params()
    p.nu, p.S_s = sp.ν, sp.S_s
    p.nu_ss_gravel, p.nu_ss_om, p.nu_ss_quartz = sp.ν_ss_gravel, sp.ν_ss_om, sp.ν_ss_quartz
    p.rho_c_ds, p.kappa_solid, p.rho_p = sp.ρc_ds, sp.κ_solid, sp.ρp
    p.kappa_sat_unfrozen, p.kappa_sat_frozen = sp.κ_sat_unfrozen, sp.κ_sat_frozen
    p.a, p.b, p.kappa_dry_parameter = sp.a, sp.b, sp.κ_dry_parameter
    p.z_0m, p.z_0s = sp.z_0m, sp.z_0s
    hyd = model.hydrology_model
    if isinstance(hyd, SoilHydrologyModel):
        hm = hyd.hydraulic_model
        visc, imp = hyd.viscosity_factor, hyd.impedance_factor
    else:
        from .parameterizations import vanGenuchten
        hm, visc, imp = vanGenuchten(), NoEffect(), NoEffect()
    p.vg_n, p.vg_alpha, p.vg_m, p.theta_r, p.Ksat = hm.n, hm.α, hm.m, hm.θr, hm.Ksat
    p.viscosity_factor = 1 if isinstance(visc, TemperatureDependentViscosity) else 0
    p.impedance_factor = 1 if isinstance(imp, IceImpedance) else 0
    p.visc_gamma = visc.γ if isinstance(visc, TemperatureDependentViscosity) else 2.64e-2
    p.visc_T_ref = visc.T_ref if isinstance(visc, TemperatureDependentViscosity) else 288.0
    p.imp_Omega = imp.Ω if isinstance(imp, IceImpedance) else 7.0
    p.rho_cloud_liq, p.rho_cloud_ice = ep.ρ_cloud_liq, ep.ρ_cloud_ice
    p.cp_l, p.cp_i, p.T_0, p.LH_f0, p.K_therm = ep.cp_l, ep.cp_i, ep.T_0, ep.LH_f0, ep.K_therm
    return p


def build_atmos(model: SoilModel) -> _abi.lh_soil_atmos:
    """``lh_soil_atmos`` from the model's ``PrescribedAtmosForcing`` top BC and its earth parameter set."""
    bc, ep = model.boundary_conditions.top, model.earth_param_set
    a = _abi.lh_soil_atmos()
    a.u_atm, a.theta_atm, a.z_atm, a.theta_scale, a.rho_a_sfc, a.q_atm = bc.u_atm, bc.θ_atm, bc.z_atm, bc.θ_scale, bc.ρ_a_sfc, bc.q_atm
    a.R_v, a.R_d, a.grav, a.cp_d, a.cp_v, a.LH_v0 = ep.R_v, ep.R_d, ep.grav, ep.cp_d, ep.cp_v, ep.LH_v0
    a.press_triple, a.T_triple, a.von_karman = ep.press_triple, ep.T_triple, ep.von_karman_const
    a.Pr_0, a.a_m, a.a_h = ep.Pr_0, ep.a_m, ep.a_h
    return a


def build_config(model: SoilModel, t0: float = 0.0, *, device: int = 0, ncol: Optional[int] = None,
                 check_finite: bool = False) -> lh_soil_config:
    kind = model.kind
    if kind is None:
        raise ValueError("prescribed/prescribed model has no device right-hand side")
    bcs = model.boundary_conditions
    atmos_top = isinstance(bcs.top, PrescribedAtmosForcing)
    if atmos_top and kind != _abi.LH_MODEL_COUPLED:
        # boundary_conditions.jl:103-112 / test_prescribed_atmos_bc.jl:161-185: only with both components prognostic
        raise _abi.SoilError(_abi.LH_ERR_UNSUPPORTED_BC,
                             "PrescribedAtmosForcing needs SoilEnergyModel + SoilHydrologyModel (no method for prescribed components)")
    cfg = lh_soil_config()
    cfg.device = device
    cfg.ncol = int(ncol if ncol is not None else model.domain.ncolumns)
    cfg.nlayer = int(model.domain.nelements)
    cfg.model = kind
    cfg.zmin, cfg.zmax = float(model.domain.zlim[0]), float(model.domain.zlim[1])
    cfg.params = build_params(model)
    for face_cfg, face_bc in ((cfg.top, bcs.top), (cfg.bottom, bcs.bottom)):
        if isinstance(face_bc, PrescribedAtmosForcing):
            # the fluxes come per column from lh_soil_set_atmos_forcing (SoilEngine); as far as the config goes the face is a flux
            face_cfg.energy_kind, face_cfg.energy_value = _abi.LH_BC_FLUX, 0.0
            face_cfg.hydrology_kind, face_cfg.hydrology_value = _abi.LH_BC_FLUX, 0.0
            continue
        face_cfg.energy_kind, face_cfg.energy_value = _bc_kind_value(face_bc.energy, t0)
        face_cfg.hydrology_kind, face_cfg.hydrology_value = _bc_kind_value(face_bc.hydrology, t0)
    cfg.flags = _abi.LH_FLAG_CHECK_FINITE if check_finite else 0
    return cfg


class SoilEngine:
    """Device context + the host-side closures (Dirichlet values, prescribed profiles)."""

    def __init__(self, model: SoilModel, t0: float = 0.0, *, device: int = 0,
                 column_range: Optional[Sequence[int]] = None, library: Optional[SoilLibrary] = None,
                 check_finite: bool = False):
        self.model = model
        self.lib = library if library is not None else current_library()
        ntot = model.domain.ncolumns
        self.column_range = (0, ntot) if column_range is None else (int(column_range[0]), int(column_range[1]))
        lo, hi = self.column_range
        if not (0 <= lo < hi <= ntot):
            raise ValueError(f"bad column range {self.column_range} for {ntot} columns")
        self.cfg = build_config(model, t0, device=device, ncol=hi - lo, check_finite=check_finite)
        self.ctx = SoilContext(self.lib, self.cfg)
        self.zc = self.ctx.zc()
        self._aux_cache = {}
        if isinstance(model.boundary_conditions.top, PrescribedAtmosForcing):
            self.ctx.set_atmos_forcing(build_atmos(model))
        self.update_aux(t0)

    def close(self):
        self.ctx.close()

    # -- boundary values: evaluate Dirichlet closures on the host (boundary_conditions.jl:247,267)
    def _component_bcs(self):
        bcs = self.model.boundary_conditions
        top = (NoBC(), NoBC()) if isinstance(bcs.top, PrescribedAtmosForcing) else (bcs.top.energy, bcs.top.hydrology)
        return top + (bcs.bottom.energy, bcs.bottom.hydrology)

    def bc_values(self, t: float) -> np.ndarray:
        out = np.zeros(4)
        for i, bc in enumerate(self._component_bcs()):
            if isinstance(bc, Dirichlet):
                out[i] = float(bc.state_value(t))
            elif isinstance(bc, VerticalFlux):
                out[i] = float(bc.flux)
        return out

    def has_dirichlet(self) -> bool:
        return any(isinstance(b, Dirichlet) for b in self._component_bcs())

    # -- prescribed profiles: make_update_aux (right_hand_side.jl:54-96) -----------------------
    def prescribed_profiles(self, t: float):
        m = self.model
        out = {}
        if isinstance(m.energy_model, PrescribedTemperatureModel):
            out["T"] = np.array([m.energy_model.T_profile(float(z), t) for z in self.zc], dtype=np.float64)
        if isinstance(m.hydrology_model, PrescribedHydrologyModel):
            hm = m.hydrology_model
            out[nf("ϑ_l")] = np.array([hm.ϑ_l_profile(float(z), t) for z in self.zc], dtype=np.float64)
            out[nf("θ_i")] = np.array([hm.θ_i_profile(float(z), t) for z in self.zc], dtype=np.float64)
        return out

    def update_aux(self, t: float, Ya: Optional[FieldVector] = None):
        """Evaluate the prescribed profiles at ``t``, upload what changed, mirror into ``Ya``."""
        for name, prof in self.prescribed_profiles(t).items():
            cached = self._aux_cache.get(name)
            if cached is None or not np.array_equal(cached, prof):
                self.ctx.set_aux(_FIELD_ID[name], prof, per_layer=True)
                self._aux_cache[name] = prof
            if Ya is not None:
                getattr(Ya, self.model.name)[name][...] = prof

    def has_time_dependent_aux(self, t: float, dt: float) -> bool:
        a, b, c = self.prescribed_profiles(t), self.prescribed_profiles(t + dt), self.prescribed_profiles(t + 0.5 * dt)
        return any(not (np.array_equal(a[k], b[k]) and np.array_equal(a[k], c[k])) for k in a)

    # -- heterogeneous soils (new): per-column hydraulic parameters ---------------------------------
    def set_column_params(self, *, ν=None, θr=None, n=None, α=None, Ksat=None):
        """Per-column ``ν``, ``θr``, van Genuchten ``n`` and ``α``, ``Ksat`` (arrays over ALL columns of the domain; this
        engine takes its own column range) instead of the scalars of ``SoilParams`` / ``vanGenuchten``
        (reference parameters.jl:11-43, SoilWaterParameterizations.jl:151-170).  ``None`` keeps the model's scalar."""
        def shard(a):
            if a is None:
                return None
            a = np.asarray(a, dtype=np.float64)
            if a.shape != (self.model.domain.ncolumns,):
                raise ValueError(f"per-column parameter must have shape ({self.model.domain.ncolumns},)")
            lo, hi = self.column_range
            return np.ascontiguousarray(a[lo:hi])
        self.ctx.set_column_params(nu=shard(ν), theta_r=shard(θr), vg_n=shard(n), vg_alpha=shard(α), Ksat=shard(Ksat))

    def set_cell_params(self, *, ν=None, θr=None, n=None, α=None, Ksat=None):
        """Layered soils: per-CELL ``ν``, ``θr``, van Genuchten ``n`` and ``α``, ``Ksat`` — arrays of shape (ncolumns, nelements)
        over ALL columns of the domain (this engine takes its own column range)."""
        def shard(a):
            if a is None:
                return None
            a = np.asarray(a, dtype=np.float64)
            if a.shape != (self.model.domain.ncolumns, self.model.domain.nelements):
                raise ValueError(f"per-cell parameter must have shape ({self.model.domain.ncolumns}, {self.model.domain.nelements})")
            lo, hi = self.column_range
            return np.ascontiguousarray(a[lo:hi])
        self.ctx.set_cell_params(nu=shard(ν), theta_r=shard(θr), vg_n=shard(n), vg_alpha=shard(α), Ksat=shard(Ksat))

    def set_column_heat_params(self, *, ρc_ds=None, κ_sat_unfrozen=None, κ_sat_frozen=None, κ_solid=None, ν_ss_om=None,
                               ν_ss_quartz=None, ν_ss_gravel=None):
        """Per-column heat parameters of ``SoilParams`` (reference parameters.jl:11-43); arrays over ALL columns of the domain."""
        def shard(a):
            if a is None:
                return None
            a = np.asarray(a, dtype=np.float64)
            if a.shape != (self.model.domain.ncolumns,):
                raise ValueError(f"per-column parameter must have shape ({self.model.domain.ncolumns},)")
            lo, hi = self.column_range
            return np.ascontiguousarray(a[lo:hi])
        self.ctx.set_column_heat_params(rho_c_ds=shard(ρc_ds), kappa_sat_unfrozen=shard(κ_sat_unfrozen),
                                        kappa_sat_frozen=shard(κ_sat_frozen), kappa_solid=shard(κ_solid), nu_ss_om=shard(ν_ss_om),
                                        nu_ss_quartz=shard(ν_ss_quartz), nu_ss_gravel=shard(ν_ss_gravel))

    # -- state transfer ----------------------------------------------------------------------------
    def _shard(self, a: np.ndarray) -> np.ndarray:
        if a.ndim == 1:
            return a
        lo, hi = self.column_range
        return a[lo:hi]

    def upload(self, Y: FieldVector):
        soil = getattr(Y, self.model.name)
        for name in self.model.prognostic_names:
            self.ctx.set_state(_FIELD_ID[nf(name)], np.ascontiguousarray(self._shard(soil[name]), dtype=np.float64))

    def download(self, Y: FieldVector):
        soil = getattr(Y, self.model.name)
        for name in self.model.prognostic_names:
            self.ctx.get_state(_FIELD_ID[nf(name)], self._shard(soil[name]))

    def download_tendency(self, dY: FieldVector):
        soil = getattr(dY, self.model.name)
        for name in self.model.prognostic_names:
            self.ctx.get_tendency(_FIELD_ID[nf(name)], self._shard(soil[name]))


def engine_for(model: SoilModel, t0: float = 0.0) -> SoilEngine:
    """One cached engine per model (and per library, so a test-injected library is honoured)."""
    lib = current_library()
    eng = model._engine
    if eng is None or eng.lib is not lib:
        eng = SoilEngine(model, t0, library=lib)
        model._engine = eng
    return eng

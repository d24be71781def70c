"""Seeded synthetic workloads shared by tests/, __graft_entry__.smoke() and bench.py.

Definitions follow SURVEY.md §8(d): parameters of the reference's own tests, smooth-plus-noise
saturation profiles, ``numpy.random.default_rng(seed)``.  Works at the ctypes-context level
(``SoilContext``) so the same inputs go to the CUDA library and to the oracle.
"""
from __future__ import annotations

import os
import sys
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

lh = graft.load_package()
abi = lh._abi

BASE_SEED = 20211001


def earth_defaults():
    return lh.EarthParameterSet()


def coupled_soil_params():
    """test/SoilModel/coupled.jl:3-32"""
    ν = 0.5
    κ_solid = lh.k_solid(0.0, 0.92, 7.7, 2.5, 0.25)
    return lh.SoilParams(
        ν=ν, S_s=1e-3, ν_ss_gravel=0.0, ν_ss_om=0.0, ν_ss_quartz=0.92,
        ρc_ds=(1 - ν) * 1.926e06, κ_solid=κ_solid,
        κ_sat_unfrozen=lh.ksat_unfrozen(κ_solid, ν, 0.57),
        κ_sat_frozen=lh.ksat_frozen(κ_solid, ν, 2.29),
    )


def coupled_vg():
    return lh.vanGenuchten(n=2.0, α=2.6, Ksat=0.0443 / 3600 / 100, θr=0.0)


def sand_soil_params():
    """test/SoilModel/richards_equation.jl:102-109"""
    return lh.SoilParams(ν=0.287, S_s=1e-3)


def sand_vg():
    return lh.vanGenuchten(n=3.96, α=2.7, Ksat=34 / 3600 / 100, θr=0.075)


def make_params(sp, vg, ep=None, viscosity=None, impedance=None):
    """lh_soil_params from the host-side structs (through the package's own translation)."""
    ep = ep or earth_defaults()
    hyd = lh.SoilHydrologyModel(
        hydraulic_model=vg,
        viscosity_factor=viscosity or lh.NoEffect(),
        impedance_factor=impedance or lh.NoEffect(),
    )
    dom = lh.Column(zlim=(-1.0, 0.0), nelements=2)
    bc = lh.SoilColumnBC(
        top=lh.SoilComponentBC(energy=lh.VerticalFlux(0.0), hydrology=lh.VerticalFlux(0.0)),
        bottom=lh.SoilComponentBC(energy=lh.VerticalFlux(0.0), hydrology=lh.VerticalFlux(0.0)),
    )
    m = lh.SoilModel(domain=dom, energy_model=lh.SoilEnergyModel(), hydrology_model=hyd,
                     boundary_conditions=bc, soil_param_set=sp, earth_param_set=ep)
    return lh.build_params(m)


@dataclass
class Workload:
    model: int
    ncol: int
    nlayer: int
    zmin: float
    zmax: float
    params: object
    top: tuple          # (energy_kind, energy_value, hydrology_kind, hydrology_value)
    bottom: tuple
    dt: float
    fields: dict = field(default_factory=dict)   # field id -> (ncol, nlayer) float64
    aux_T: Optional[np.ndarray] = None           # per-layer prescribed T (Richards)
    name: str = ""
    device: int = 0

    def config(self, ncol: Optional[int] = None, flags: int = 0):
        cfg = abi.lh_soil_config()
        cfg.device = self.device
        cfg.ncol = int(ncol if ncol is not None else self.ncol)
        cfg.nlayer = self.nlayer
        cfg.model = self.model
        cfg.zmin, cfg.zmax = self.zmin, self.zmax
        cfg.params = self.params
        for fc, spec in ((cfg.top, self.top), (cfg.bottom, self.bottom)):
            fc.energy_kind, fc.energy_value, fc.hydrology_kind, fc.hydrology_value = spec
        cfg.flags = flags
        return cfg

    def upload(self, ctx, lo: int = 0, hi: Optional[int] = None):
        hi = self.ncol if hi is None else hi
        for fid, arr in self.fields.items():
            ctx.set_state(fid, np.ascontiguousarray(arr[lo:hi]))
        if self.aux_T is not None:
            ctx.set_aux(abi.LH_FIELD_T, self.aux_T, per_layer=True)

    @property
    def cells(self) -> int:
        return self.ncol * self.nlayer


def zc_of(zmin, zmax, n):
    j = np.arange(n + 1, dtype=np.float64)
    zf = zmin + (zmax - zmin) * j / n
    return (zf[:-1] + zf[1:]) / 2.0


CHUNK = 16384  # columns per independently seeded block: any shard can generate just its own columns


def _chunks(col_range):
    lo, hi = col_range
    for k in range(lo // CHUNK, (hi - 1) // CHUNK + 1):
        c0, c1 = k * CHUNK, (k + 1) * CHUNK
        yield k, max(lo, c0) - c0, min(hi, c1) - c0, max(lo, c0) - lo, min(hi, c1) - lo


def saturation_profiles(seed, col_range, nlayer, zmin, zmax, lo=0.05, hi=0.98):
    """S = clip(s0 + a sin(2π (z - zmin)/L + φ) + 0.02 N(0,1), lo, hi)   (SURVEY §8d).
    Block k of CHUNK columns is drawn from default_rng([seed, 1, k]), so the profile of a column
    does not depend on how the column set is sharded."""
    z = zc_of(zmin, zmax, nlayer)
    L = zmax - zmin
    S = np.empty((col_range[1] - col_range[0], nlayer), dtype=np.float64)
    for k, a0, a1, o0, o1 in _chunks(col_range):
        rng = np.random.default_rng([seed, 1, k])
        s0 = rng.uniform(0.3, 0.8, size=(CHUNK, 1))
        a = rng.uniform(0.0, 0.15, size=(CHUNK, 1))
        phi = rng.uniform(0.0, 2 * np.pi, size=(CHUNK, 1))
        noise = rng.standard_normal(size=(CHUNK, nlayer))
        blk = np.clip(s0 + a * np.sin(2 * np.pi * (z[None, :] - zmin) / L + phi) + 0.02 * noise, lo, hi)
        S[o0:o1] = blk[a0:a1]
    return S


def temperature_profiles(seed, col_range, nlayer, zmin, zmax):
    """T = 285 + 10 U(0,1), smooth in z."""
    z = zc_of(zmin, zmax, nlayer)
    L = zmax - zmin
    T = np.empty((col_range[1] - col_range[0], nlayer), dtype=np.float64)
    for k, a0, a1, o0, o1 in _chunks(col_range):
        rng = np.random.default_rng([seed, 2, k])
        base = rng.uniform(0.0, 1.0, size=(CHUNK, 1))
        amp = rng.uniform(0.0, 0.3, size=(CHUNK, 1))
        phi = rng.uniform(0.0, 2 * np.pi, size=(CHUNK, 1))
        blk = 285.0 + 10.0 * np.clip(base + amp * np.sin(2 * np.pi * (z[None, :] - zmin) / L + phi), 0.0, 1.0)
        T[o0:o1] = blk[a0:a1]
    return T


def ice_profiles(seed, col_range, nlayer, hi):
    out = np.empty((col_range[1] - col_range[0], nlayer), dtype=np.float64)
    for k, a0, a1, o0, o1 in _chunks(col_range):
        rng = np.random.default_rng([seed, 3, k])
        out[o0:o1] = rng.uniform(0.0, hi, size=(CHUNK, nlayer))[a0:a1]
    return out


def rho_e_int_from_T(p, theta_l_aug, theta_i, T):
    """volumetric_heat_capacity + volumetric_internal_energy (SoilHeatParameterizations.jl:65-102)."""
    nu_eff = p.nu - theta_i
    theta_l = np.where(theta_l_aug < nu_eff, theta_l_aug, nu_eff)
    rho_c_s = p.rho_c_ds + theta_l * (p.cp_l * p.rho_cloud_liq) + theta_i * (p.cp_i * p.rho_cloud_ice)
    return rho_c_s * (T - p.T_0) - theta_i * p.rho_cloud_ice * p.LH_f0


D, F, FD, N = abi.LH_BC_DIRICHLET, abi.LH_BC_FLUX, abi.LH_BC_FREE_DRAINAGE, abi.LH_BC_NONE


def coupled_workload(ncol=1 << 20, nlayer=64, seed=None, ice=False, zlim=(-2.0, 0.0),
                     viscosity=None, impedance=None, top=None, bottom=None, sat_hi=0.98, col_range=None):
    """BASELINE configs C2/C4/C5-coupled: coupled.jl parameters, Dirichlet top (ϑ_l, T),
    FreeDrainage (water) / zero flux (energy) bottom, dt = 20 s."""
    seed = BASE_SEED + 3 if seed is None else seed
    cr = (0, ncol) if col_range is None else tuple(col_range)
    ncol = cr[1] - cr[0]       # the workload holds only the requested shard of columns
    p = make_params(coupled_soil_params(), coupled_vg(), viscosity=viscosity, impedance=impedance)
    zmin, zmax = zlim
    S = saturation_profiles(seed, cr, nlayer, zmin, zmax, hi=sat_hi)
    theta_i = ice_profiles(seed, cr, nlayer, 0.05) if ice else np.zeros((ncol, nlayer))
    theta = p.theta_r + S * ((p.nu - theta_i) - p.theta_r)
    T = temperature_profiles(seed, cr, nlayer, zmin, zmax)
    rho_e = rho_e_int_from_T(p, theta, theta_i, T)
    return Workload(
        model=abi.LH_MODEL_COUPLED, ncol=ncol, nlayer=nlayer, zmin=zmin, zmax=zmax, params=p,
        top=top or (D, 288.0, D, 0.4), bottom=bottom or (F, 0.0, FD, 0.0), dt=20.0,
        fields={0: theta, 1: theta_i, 2: rho_e}, name=f"coupled_{ncol}x{nlayer}",
    )


def richards_workload(ncol=1024, nlayer=100, seed=None, ice=False, zlim=(-1.5, 0.0),
                      viscosity=None, impedance=None, top=None, bottom=None, sat_hi=0.98, col_range=None):
    """BASELINE configs C1/C3/C5-Richards: Bonan sand, Dirichlet top ϑ_l = 0.267, FreeDrainage
    bottom, dt = 0.25 s (richards_equation.jl:98-167)."""
    seed = BASE_SEED + 2 if seed is None else seed
    cr = (0, ncol) if col_range is None else tuple(col_range)
    ncol = cr[1] - cr[0]
    p = make_params(sand_soil_params(), sand_vg(), viscosity=viscosity, impedance=impedance)
    zmin, zmax = zlim
    S = saturation_profiles(seed, cr, nlayer, zmin, zmax, hi=sat_hi)
    theta_i = ice_profiles(seed, cr, nlayer, 0.03) if ice else np.zeros((ncol, nlayer))
    theta = p.theta_r + S * ((p.nu - theta_i) - p.theta_r)
    aux_T = None
    if viscosity is not None:
        aux_T = 288.0 + 8.0 * np.sin(np.linspace(0, 3, nlayer))
    return Workload(
        model=abi.LH_MODEL_RICHARDS, ncol=ncol, nlayer=nlayer, zmin=zmin, zmax=zmax, params=p,
        top=top or (N, 0.0, D, 0.267), bottom=bottom or (N, 0.0, FD, 0.0), dt=0.25,
        fields={0: theta, 1: theta_i}, aux_T=aux_T, name=f"richards_{ncol}x{nlayer}",
    )


def heat_workload(ncol=256, nlayer=60, seed=None, ice=False, zlim=(0.0, 1.0), top=None, bottom=None):
    """Heat-only model (prescribed hydrology), coupled.jl soil parameters, Dirichlet T both ends."""
    seed = BASE_SEED + 1 if seed is None else seed
    cr = (0, ncol)
    p = make_params(coupled_soil_params(), coupled_vg())
    zmin, zmax = zlim
    S = saturation_profiles(seed, cr, nlayer, zmin, zmax)
    theta_i = ice_profiles(seed, cr, nlayer, 0.05) if ice else np.zeros((ncol, nlayer))
    theta = p.theta_r + S * ((p.nu - theta_i) - p.theta_r)
    T = temperature_profiles(seed, cr, nlayer, zmin, zmax)
    rho_e = rho_e_int_from_T(p, theta, theta_i, T)
    return Workload(
        model=abi.LH_MODEL_HEAT, ncol=ncol, nlayer=nlayer, zmin=zmin, zmax=zmax, params=p,
        top=top or (D, 290.0, N, 0.0), bottom=bottom or (D, 280.0, N, 0.0), dt=50.0,
        fields={0: theta, 1: theta_i, 2: rho_e}, name=f"heat_{ncol}x{nlayer}",
    )


def oracle_library():
    return lh.SoilLibrary(graft.build_oracle(), "lho_")


def tendency_scale(oracle_ctx, field_id):
    """Per-column scale of the cancellation-aware parity norm:
    max(‖r‖∞ over the column, max_j |face flux_j| / Δz), from the oracle's last rhs call."""
    import ctypes as C

    n, ncol = oracle_ctx.nlayer, oracle_ctx.ncol
    r = oracle_ctx.get_tendency(field_id)
    dz = (oracle_ctx.cfg.zmax - oracle_ctx.cfg.zmin) / n
    fn = oracle_ctx.lib.raw("lho_soil_face_fluxes")
    fn.restype = C.c_int32
    fn.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    Fw = np.empty(n + 1)
    Fe = np.empty(n + 1)
    scale = np.empty(ncol)
    dp = C.POINTER(C.c_double)
    for c in range(ncol):
        fn(oracle_ctx._h, c, Fw.ctypes.data_as(dp), Fe.ctypes.data_as(dp))
        F = Fw if field_id == 0 else Fe
        scale[c] = max(np.max(np.abs(r[c])), np.max(np.abs(F)) / dz)
    scale[scale == 0] = 1.0
    return scale

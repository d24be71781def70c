"""Host-side scalar parameterisation functions and their parameter structs.

Mirror of the reference modules ``SoilWaterParameterizations`` and ``SoilHeatParameterizations``
(src/SoilModel/SoilWaterParameterizations.jl, SoilHeatParameterizations.jl): same names, same
argument order, same expression order.  In the reference these are the functions user scripts
call to build initial conditions and to post-process states (e.g. test/SoilModel/coupled.jl:
78-82, 97-100); the hot path evaluates the SAME closures on the device inside the fused
kernels (csrc/lh_closures.cuh), so these host copies are setup helpers, not a fallback.

Identifiers keep the reference's Greek names.  Python NFKC-normalises identifiers, so ``ϑ_l``
(U+03D1) and ``θ_l`` (U+03B8) are the same identifier here; this module therefore spells the
augmented liquid fraction ``ϑ_l`` only in positions where the reference never has a distinct
``θ_l`` next to it.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

EPS = 2.220446049250313e-16  # eps(Float64)


# ---- earth parameter set (CLIMAParameters v0.1 defaults) ----------------------
@dataclass(frozen=True)
class EarthParameterSet:
    """Stand-in for ``struct EarthParameterSet <: AbstractEarthParameterSet`` (test/runtests.jl:14).

    Only the constants read at SoilHeatParameterizations.jl:12-13 are carried.  They cross the
    C ABI as runtime doubles so a Julia host can supply its own CLIMAParameters values.
    """

    ρ_cloud_liq: float = 1000.0
    ρ_cloud_ice: float = 916.7
    cp_l: float = 4181.0
    cp_i: float = 2100.0
    T_0: float = 273.16
    LH_f0: float = 2.8344e6 - 2.5008e6  # LH_s0 - LH_v0
    K_therm: float = 2.4e-2
    # read only by PrescribedAtmosForcing (boundary_conditions.jl:575-617 and, through SurfaceFluxes / Thermodynamics, the
    # saturation vapour pressure and the similarity functions): CLIMAParameters v0.1 values
    R_v: float = 8.3144598 / 18.01528e-3
    R_d: float = 8.3144598 / 28.97e-3
    grav: float = 9.81
    cp_d: float = (8.3144598 / 28.97e-3) / (2.0 / 7.0)
    cp_v: float = 1859.0
    LH_v0: float = 2.5008e6
    press_triple: float = 611.657
    T_triple: float = 273.16
    von_karman_const: float = 0.4
    # SurfaceFluxes.UniversalFunctions.Businger
    Pr_0: float = 0.74
    a_m: float = 4.7
    a_h: float = 4.7


def q_vap_saturation_liquid(ps, T, ρ):
    """``Thermodynamics.q_vap_saturation_generic(param_set, T, ρ, Liquid())`` restated from the published closed form
    (Clausius-Clapeyron with constant Δcp = cp_v - cp_l; PARITY UNPINNED, include/lh_soil.h): host-side helper for building a
    ``PrescribedAtmosForcing`` the way test_prescribed_atmos_bc.jl:28 does."""
    import math

    dcp = ps.cp_v - ps.cp_l
    p_vs = ps.press_triple * (T / ps.T_triple) ** (dcp / ps.R_v) * math.exp((ps.LH_v0 - dcp * ps.T_0) / ps.R_v * (1.0 / ps.T_triple - 1.0 / T))
    return p_vs / (ρ * ps.R_v * T)


def ρ_cloud_liq(ps): return ps.ρ_cloud_liq
def ρ_cloud_ice(ps): return ps.ρ_cloud_ice
def cp_l(ps): return ps.cp_l
def cp_i(ps): return ps.cp_i
def T_0(ps): return ps.T_0
def LH_f0(ps): return ps.LH_f0
def K_therm(ps): return ps.K_therm


# ---- conductivity factors (SoilWaterParameterizations.jl:30-126) -----------------------------
class AbstractConductivityFactor:
    pass


@dataclass(frozen=True)
class NoEffect(AbstractConductivityFactor):
    """:38"""


@dataclass(frozen=True)
class TemperatureDependentViscosity(AbstractConductivityFactor):
    """:46-52"""
    γ: float = 2.64e-2
    T_ref: float = 288.0


@dataclass(frozen=True)
class IceImpedance(AbstractConductivityFactor):
    """:62-65"""
    Ω: float = 7.0


def impedance_factor(imp: AbstractConductivityFactor, *args) -> float:
    """:76-93"""
    if isinstance(imp, NoEffect):
        return 1.0
    if isinstance(imp, IceImpedance):
        (f_i,) = args
        return 10.0 ** (-imp.Ω * f_i)
    raise TypeError(f"no method impedance_factor(::{type(imp).__name__})")


def viscosity_factor(vm: AbstractConductivityFactor, *args) -> float:
    """:104-126"""
    if isinstance(vm, NoEffect):
        return 1.0
    if isinstance(vm, TemperatureDependentViscosity):
        (T,) = args
        factor = vm.γ * (T - vm.T_ref)
        return math.exp(factor)
    raise TypeError(f"no method viscosity_factor(::{type(vm).__name__})")


# ---- van Genuchten (SoilWaterParameterizations.jl:139-170) ------------------------------------
class AbstractHydraulicsModel:
    pass


class vanGenuchten(AbstractHydraulicsModel):
    """``vanGenuchten{FT}(; n = 1.56, α = 3.6, Ksat = 2.9e-7, θr = 0)``; m = 1 - 1/n (:162-169)."""

    __slots__ = ("n", "α", "m", "θr", "Ksat")

    def __init__(self, *, n: float = 1.56, α: float = 3.6, Ksat: float = 2.9e-7, θr: float = 0.0):
        self.n = float(n)
        self.α = float(α)
        self.m = 1.0 - 1.0 / float(n)
        self.θr = float(θr)
        self.Ksat = float(Ksat)

    def __repr__(self):
        return f"vanGenuchten(n={self.n}, α={self.α}, m={self.m}, θr={self.θr}, Ksat={self.Ksat})"


def volumetric_liquid_fraction(ϑ_l: float, ν_eff: float) -> float:
    """:181-188"""
    return ϑ_l if ϑ_l < ν_eff else ν_eff


def matric_potential(hm: vanGenuchten, S: float) -> float:
    """:196-200"""
    n, α, m = hm.n, hm.α, hm.m
    return -(((S ** (-1.0 / m) - 1.0) * α ** (-n)) ** (1.0 / n))


def effective_saturation(porosity: float, ϑ_l: float, θr: float) -> float:
    """:213-217"""
    ϑ_l_safe = max(ϑ_l, θr + EPS)
    return (ϑ_l_safe - θr) / (porosity - θr)


def pressure_head(hm: vanGenuchten, ϑ_l: float, ν_eff: float, S_s: float) -> float:
    """:229-242"""
    S_l_eff = effective_saturation(ν_eff, ϑ_l, hm.θr)
    if S_l_eff <= 1.0:
        return matric_potential(hm, S_l_eff)
    return (ϑ_l - ν_eff) / S_s


def inverse_matric_potential(hm: vanGenuchten, ψ: float) -> float:
    """:253-258"""
    if ψ > 0:
        raise ValueError("Matric potential is positive")
    return (1.0 + (hm.α * abs(ψ)) ** hm.n) ** (-hm.m)


def hydraulic_conductivity(hm: vanGenuchten, S: float, viscosity_f: float, impedance_f: float) -> float:
    """:269-282"""
    if S < 1.0:
        K = math.sqrt(S) * (1.0 - (1.0 - S ** (1.0 / hm.m)) ** hm.m) ** 2.0
    else:
        K = 1.0
    return K * hm.Ksat * viscosity_f * impedance_f


def hydrostatic_profile(hm: vanGenuchten, z: float, z_interface: float, ν: float, S_s: float) -> float:
    """:290-306"""
    if z > z_interface:
        S = (1.0 + (hm.α * (z - z_interface)) ** hm.n) ** (-hm.m)
        return S * (ν - hm.θr) + hm.θr
    return -S_s * (z - z_interface) + ν


# ---- heat closures (SoilHeatParameterizations.jl) ----------------------------------------------
def temperature_from_ρe_int(ρe_int: float, θ_i: float, ρc_s: float, param_set) -> float:
    """:42-53"""
    return param_set.T_0 + (ρe_int + θ_i * param_set.ρ_cloud_ice * param_set.LH_f0) / ρc_s


def volumetric_heat_capacity(θ_l: float, θ_i: float, ρc_ds: float, param_set) -> float:
    """:65-79"""
    ρcp_i = param_set.cp_i * param_set.ρ_cloud_ice
    ρcp_l = param_set.cp_l * param_set.ρ_cloud_liq
    return ρc_ds + θ_l * ρcp_l + θ_i * ρcp_i


def volumetric_internal_energy(θ_i: float, ρc_s: float, T: float, param_set) -> float:
    """:91-102"""
    return ρc_s * (T - param_set.T_0) - θ_i * param_set.ρ_cloud_ice * param_set.LH_f0


def saturated_thermal_conductivity(θ_l: float, θ_i: float, κ_sat_unfrozen: float, κ_sat_frozen: float) -> float:
    """:114-128"""
    θ_w = θ_l + θ_i
    if θ_w < EPS:
        return 0.0
    return κ_sat_unfrozen ** (θ_l / θ_w) * κ_sat_frozen ** (θ_i / θ_w)


def relative_saturation(θ_l: float, θ_i: float, porosity: float) -> float:
    """:139-142"""
    return (θ_l + θ_i) / porosity


def kersten_number(θ_i: float, S_r: float, soil_param_functions) -> float:
    """:152-174"""
    sp = soil_param_functions
    if θ_i < EPS:
        return S_r ** ((1.0 + sp.ν_ss_om - sp.a * sp.ν_ss_quartz - sp.ν_ss_gravel) / 2.0) * (
            (1.0 + math.exp(-sp.b * S_r)) ** (-3.0) - ((1.0 - S_r) / 2.0) ** 3.0
        ) ** (1.0 - sp.ν_ss_om)
    return S_r ** (1.0 + sp.ν_ss_om)


def thermal_conductivity(κ_dry: float, K_e: float, κ_sat: float) -> float:
    """:185-188"""
    return K_e * κ_sat + (1.0 - K_e) * κ_dry


def volumetric_internal_energy_liq(T: float, param_set) -> float:
    """:198-207"""
    ρcp_l = param_set.cp_l * param_set.ρ_cloud_liq
    return ρcp_l * (T - param_set.T_0)


def k_solid(ν_ss_om: float, ν_ss_quartz: float, κ_quartz: float, κ_minerals: float, κ_om: float) -> float:
    """:223-233"""
    return κ_om ** ν_ss_om * κ_quartz ** ν_ss_quartz * κ_minerals ** (1.0 - ν_ss_om - ν_ss_quartz)


def ksat_frozen(κ_solid: float, porosity: float, κ_ice: float) -> float:
    """:245-247"""
    return κ_solid ** (1.0 - porosity) * κ_ice ** porosity


def ksat_unfrozen(κ_solid: float, porosity: float, κ_l: float) -> float:
    """:258-260"""
    return κ_solid ** (1.0 - porosity) * κ_l ** porosity


def ρb_ss(porosity: float, ρp: float) -> float:
    """:268-270"""
    return (1.0 - porosity) * ρp


def k_dry(param_set, soil_param_functions) -> float:
    """:280-294"""
    sp = soil_param_functions
    κ_air = param_set.K_therm
    ρb_val = ρb_ss(sp.ν, sp.ρp)
    numerator = (sp.κ_dry_parameter * sp.κ_solid - κ_air) * ρb_val + κ_air * sp.ρp
    denom = sp.ρp - (1.0 - sp.κ_dry_parameter) * ρb_val
    return numerator / denom

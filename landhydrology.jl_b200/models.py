"""Model, parameter and boundary-condition types of the soil path.

Mirror of reference src/SoilModel/models.jl, parameters.jl and the type half of
boundary_conditions.jl (:17-161): same constructor names and keyword names, so a user script of
the reference reads the same here.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Optional, Union

import numpy as np

from . import _abi
from .domains import AbstractVerticalDomain
from .parameterizations import (
    AbstractConductivityFactor,
    IceImpedance,
    NoEffect,
    TemperatureDependentViscosity,
    vanGenuchten,
)


class AbstractModel:
    """reference src/Models.jl:11"""


# ---- SoilParams (parameters.jl:11-43) -----------------------------------------------------------
@dataclass(frozen=True)
class SoilParams:
    """``SoilParams{FT}(; ...)``: defaults correspond to loam soil."""

    ν: float = 0.43
    S_s: float = 1e-3
    ν_ss_gravel: float = 0.0
    ν_ss_om: float = 0.0
    ν_ss_quartz: float = 0.41
    ρc_ds: float = 2700.0
    κ_solid: float = 3.97
    ρp: float = 2700.0
    κ_sat_unfrozen: float = 1.72
    κ_sat_frozen: float = 3.13
    a: float = 0.24
    b: float = 18.1
    κ_dry_parameter: float = 0.053
    z_0m: float = 0.001
    z_0s: float = 0.001


# ---- component models (models.jl:7-78) -----------------------------------------------------------
class AbstractSoilComponentModel:
    pass


@dataclass(frozen=True)
class SoilEnergyModel(AbstractSoilComponentModel):
    """models.jl:17"""


@dataclass(frozen=True)
class SoilHydrologyModel(AbstractSoilComponentModel):
    """models.jl:28-33"""

    hydraulic_model: vanGenuchten = field(default_factory=vanGenuchten)
    viscosity_factor: AbstractConductivityFactor = field(default_factory=NoEffect)
    impedance_factor: AbstractConductivityFactor = field(default_factory=NoEffect)

    def __post_init__(self):
        if not isinstance(self.viscosity_factor, (NoEffect, TemperatureDependentViscosity)):
            raise TypeError("viscosity_factor must be NoEffect or TemperatureDependentViscosity")
        if not isinstance(self.impedance_factor, (NoEffect, IceImpedance)):
            raise TypeError("impedance_factor must be NoEffect or IceImpedance")


def _default_T_profile(z, t):
    return 288.0


def _zero_profile(z, t):
    return 0.0


@dataclass(frozen=True)
class PrescribedTemperatureModel(AbstractSoilComponentModel):
    """models.jl:51-54; default T ≡ 288 K."""

    T_profile: Callable = _default_T_profile


@dataclass(frozen=True)
class PrescribedHydrologyModel(AbstractSoilComponentModel):
    """models.jl:73-78; defaults are totally dry soil."""

    ϑ_l_profile: Callable = _zero_profile
    θ_i_profile: Callable = _zero_profile


# ---- boundary conditions (boundary_conditions.jl:17-161) ------------------------------------------
class AbstractBC:
    pass


@dataclass(frozen=True)
class NoBC(AbstractBC):
    """:27"""


@dataclass(frozen=True)
class VerticalFlux(AbstractBC):
    """:43-46; scalar flux, positive = aligned with ẑ (at both faces)."""

    flux: float


@dataclass(frozen=True)
class Dirichlet(AbstractBC):
    """:61-64; ``state_value`` is a host closure ``t -> value`` (T or ϑ_l)."""

    state_value: Callable


@dataclass(frozen=True)
class FreeDrainage(AbstractBC):
    """:77; ∇h = 1 at the boundary."""


class AbstractFaceBC:
    pass


@dataclass(frozen=True)
class SoilComponentBC(AbstractFaceBC):
    """:95-101"""

    energy: AbstractBC = field(default_factory=NoBC)
    hydrology: AbstractBC = field(default_factory=NoBC)


@dataclass(frozen=True)
class PrescribedAtmosForcing(AbstractFaceBC):
    """:119-132.  Turbulent surface fluxes from Monin-Obukhov similarity, top face only, both soil components
    prognostic.  PARITY UNPINNED: the similarity solve and the saturation vapour pressure live in the un-vendored
    SurfaceFluxes / Thermodynamics packages and follow the published formulations here (include/lh_soil.h
    ``lh_soil_set_atmos_forcing``)."""

    u_atm: float
    θ_atm: float
    z_atm: float
    θ_scale: float
    ρ_a_sfc: float
    q_atm: float


@dataclass(frozen=True)
class SoilColumnBC:
    """:144-161"""

    top: Union[SoilComponentBC, PrescribedAtmosForcing] = field(default_factory=SoilComponentBC)
    bottom: SoilComponentBC = field(default_factory=SoilComponentBC)

    def __post_init__(self):
        if not isinstance(self.top, (SoilComponentBC, PrescribedAtmosForcing)):
            raise TypeError("top must be a SoilComponentBC or PrescribedAtmosForcing")
        if not isinstance(self.bottom, SoilComponentBC):
            raise TypeError("bottom must be a SoilComponentBC")


# ---- SoilModel (models.jl:90-135) ---------------------------------------------------------------
class SoilModel(AbstractModel):
    """``SoilModel(FT; domain, energy_model, hydrology_model, boundary_conditions,
    soil_param_set = SoilParams{FT}(), earth_param_set, name = :soil)``."""

    def __init__(
        self,
        FT=np.float64,
        *,
        domain: AbstractVerticalDomain,
        energy_model: AbstractSoilComponentModel,
        hydrology_model: AbstractSoilComponentModel,
        boundary_conditions,
        soil_param_set: Optional[SoilParams] = None,
        earth_param_set,
        name: str = "soil",
    ):
        if FT not in (np.float64, float):
            raise TypeError("the B200 path is fp64 only (FT = Float64)")
        if not isinstance(domain, AbstractVerticalDomain):
            raise TypeError("domain must be an AbstractVerticalDomain")
        if not isinstance(energy_model, AbstractSoilComponentModel) or not isinstance(
            hydrology_model, AbstractSoilComponentModel
        ):
            raise TypeError("energy_model / hydrology_model must be AbstractSoilComponentModel")
        self.FT = np.float64
        self.domain = domain
        self.energy_model = energy_model
        self.hydrology_model = hydrology_model
        self.boundary_conditions = boundary_conditions
        self.soil_param_set = soil_param_set if soil_param_set is not None else SoilParams()
        self.earth_param_set = earth_param_set
        self.name = name
        self._engine = None  # device context, created lazily by make_rhs / Simulation

    # which right-hand side `make_rhs` dispatches to (right_hand_side.jl:103-369)
    @property
    def kind(self) -> Optional[int]:
        e_dyn = isinstance(self.energy_model, SoilEnergyModel)
        h_dyn = isinstance(self.hydrology_model, SoilHydrologyModel)
        if e_dyn and h_dyn:
            return _abi.LH_MODEL_COUPLED
        if e_dyn:
            return _abi.LH_MODEL_HEAT
        if h_dyn:
            return _abi.LH_MODEL_RICHARDS
        return None  # prescribed/prescribed: rhs! is a no-op (:103-112)

    @property
    def prognostic_names(self):
        k = self.kind
        if k == _abi.LH_MODEL_COUPLED:
            return ("ϑ_l", "θ_i", "ρe_int")
        if k == _abi.LH_MODEL_RICHARDS:
            return ("ϑ_l", "θ_i")
        if k == _abi.LH_MODEL_HEAT:
            return ("ρe_int",)
        return ()

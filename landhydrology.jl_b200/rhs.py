"""``make_rhs`` / ``make_update_aux``: mirror of reference src/SoilModel/right_hand_side.jl:33-96.

``make_rhs(model)`` returns the in-place closure ``rhs!(dY, Y, Ya, t) -> dY`` (spelled ``rhs_``
in Python).  Where the reference evaluates ~16 broadcast temporaries and 2 stencil broadcasts on
the CPU, this closure (1) evaluates the host-only closures — prescribed profiles and Dirichlet
``state_value(t)`` — (2) uploads ``Y``, (3) launches ONE fused CUDA kernel through
``lh_soil_rhs`` and (4) downloads ``dY``.  Time stepping should use ``Simulation`` instead, which
keeps the state resident on the device.
"""
from __future__ import annotations

import numpy as np

from .engine import engine_for
from .models import (
    AbstractSoilComponentModel,
    PrescribedHydrologyModel,
    PrescribedTemperatureModel,
    SoilModel,
)


def make_update_aux(component: AbstractSoilComponentModel):
    """right_hand_side.jl:54-96: returns ``update_aux!(Ya, t)``."""
    if isinstance(component, PrescribedTemperatureModel):
        def update_aux_(Ya, t):
            T = Ya.soil.T
            zc = Ya.zc
            T[...] = np.array([component.T_profile(float(z), t) for z in zc], dtype=np.float64)
            return Ya
        return update_aux_
    if isinstance(component, PrescribedHydrologyModel):
        def update_aux_(Ya, t):
            zc = Ya.zc
            Ya.soil.ϑ_l[...] = np.array([component.ϑ_l_profile(float(z), t) for z in zc], dtype=np.float64)
            Ya.soil.θ_i[...] = np.array([component.θ_i_profile(float(z), t) for z in zc], dtype=np.float64)
            return Ya
        return update_aux_

    def update_aux_(Ya, t):
        return None
    return update_aux_


def make_rhs(model: SoilModel):
    """right_hand_side.jl:33-44."""
    if model.kind is None:
        # prescribed temperature + prescribed hydrology: the RHS does nothing (:103-112)
        update_en = make_update_aux(model.energy_model)
        update_hy = make_update_aux(model.hydrology_model)

        def rhs_(dY, Y, Ya, t):
            update_en(Ya, t)
            update_hy(Ya, t)
            return dY
        return rhs_

    def rhs_(dY, Y, Ya, t):
        eng = engine_for(model, t)
        eng.update_aux(t, Ya)                      # update_aux_en!, update_aux_hydr! (:38-39)
        eng.ctx.set_bc_values(eng.bc_values(t))    # Dirichlet state_value(t)
        eng.upload(Y)
        eng.ctx.rhs(t)                             # rhs_soil!(dY, Y, Ya, t) (:40)
        eng.download_tendency(dY)
        return dY

    return rhs_


def compute_turbulent_surface_fluxes(energy, hydrology, model: SoilModel, ϑ_l, θ_i, T):
    """``compute_turbulent_surface_fluxes(energy::SoilEnergyModel, hydrology::SoilHydrologyModel, model, ϑ_l, θ_i, T)``
    (boundary_conditions.jl:555-620): ``(heat_flux, Ẽ)`` for a surface state, evaluated by the library with the model's
    parameters.  Like the reference, there is no method for prescribed components (test_prescribed_atmos_bc.jl:161-185)."""
    from .models import PrescribedAtmosForcing, SoilEnergyModel, SoilHydrologyModel

    if not isinstance(energy, SoilEnergyModel) or not isinstance(hydrology, SoilHydrologyModel):
        raise TypeError("MethodError: no method matching compute_turbulent_surface_fluxes("
                        f"::{type(energy).__name__}, ::{type(hydrology).__name__}, ...)")
    if not isinstance(model.boundary_conditions.top, PrescribedAtmosForcing):
        raise TypeError("the model's top boundary condition is not a PrescribedAtmosForcing")
    eng = engine_for(model, 0.0)
    heat, water = eng.ctx.atmos_fluxes(ϑ_l, θ_i, T)
    if np.ndim(ϑ_l) == 0:
        return float(heat[0]), float(water[0])
    return heat, water


def boundary_fluxes(X, bc, face: str, model: SoilModel, cs=None, t: float = 0.0):
    """``boundary_fluxes(X, bc::PrescribedAtmosForcing, face, model, cs, t)`` (boundary_conditions.jl:516-536): the surface
    fluxes from the interior values next to the top face; any other face is an error (:525-527)."""
    from .models import PrescribedAtmosForcing

    if not isinstance(bc, PrescribedAtmosForcing):
        raise TypeError("host-side boundary_fluxes is provided for PrescribedAtmosForcing only; component BCs are evaluated on the device")
    if face != "top":
        raise ValueError("Prescribed atmosphere-driven boundary conditions are only valid at the top of the soil column.")
    soil = getattr(X, model.name)
    heat, water = compute_turbulent_surface_fluxes(model.energy_model, model.hydrology_model, model,
                                                   np.atleast_2d(soil["ϑ_l"])[..., -1], np.atleast_2d(soil["θ_i"])[..., -1],
                                                   np.atleast_2d(soil["T"])[..., -1])
    return {"fρe_int": heat, "fϑ_l": water}

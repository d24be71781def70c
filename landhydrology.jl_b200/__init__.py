"""landhydrology.jl_b200 — B200-native soil right-hand side + SSPRK33 stage path.

Host-side mirror (Python) of the part of CliMA/LandHydrology.jl's API that sits on the soil RHS
hot path, over the C ABI of ``include/lh_soil.h`` (hand-written sm_100a CUDA kernels in
``csrc/``).  Same names and keyword arguments as the reference:

    Column, HybridBox, SoilParams, vanGenuchten, SoilHydrologyModel, SoilEnergyModel,
    PrescribedTemperatureModel, PrescribedHydrologyModel, VerticalFlux, Dirichlet, FreeDrainage,
    NoBC, SoilComponentBC, SoilColumnBC, SoilModel, initialize_states, make_rhs,
    Simulation, step_ (step!), run_ (run!)

The directory name contains a dot, so it is imported through ``__graft_entry__.load_package()``
(registered as ``landhydrology_b200`` in ``sys.modules``).  There is no CPU fallback: without
``csrc/liblh_soil.so`` and a CUDA device every compute entry point raises.
"""
from . import _abi
from ._abi import (
    ABI_SYMBOLS,
    CUDA_LIBRARY_PATH,
    DomainAssertionError,
    NoDeviceError,
    NonFiniteStateError,
    SoilContext,
    SoilError,
    SoilLibrary,
    UnsupportedBCError,
    cuda_library,
)
from .domains import (
    AbstractDomain,
    AbstractVerticalDomain,
    Column,
    HybridBox,
    length,
    make_function_space,
    ndims,
    size,
)
from .engine import SoilEngine, build_atmos, build_config, build_params, engine_for, use_library
from .models import (
    AbstractBC,
    AbstractModel,
    AbstractSoilComponentModel,
    Dirichlet,
    FreeDrainage,
    NoBC,
    PrescribedAtmosForcing,
    PrescribedHydrologyModel,
    PrescribedTemperatureModel,
    SoilColumnBC,
    SoilComponentBC,
    SoilEnergyModel,
    SoilHydrologyModel,
    SoilModel,
    SoilParams,
    VerticalFlux,
)
from .parameterizations import *  # noqa: F401,F403  (the reference exports every closure)
from .parameterizations import EarthParameterSet
from .rhs import boundary_fluxes, compute_turbulent_surface_fluxes, make_rhs, make_update_aux
from .sharding import ColumnShards, bind_to_gpu_numa_node, gpu_numa_cpus, init_budget_comm, shard_range
from .simulations import (SSPRK22, SSPRK33, SSPRK43, CarpenterKennedy2N54, Euler, LowStorageRK2N, ShuOsherRK,
                          Simulation, run_, step_)
from .states import (
    FieldVector,
    NamedFields,
    coordinates,
    copy,
    initialize_auxiliary,
    initialize_prognostic,
    initialize_states,
    parent,
    similar,
)
from .parameterizations import volumetric_heat_capacity as _vhc, volumetric_internal_energy as _vie


def default_initial_conditions(model):
    """reference models.jl:147-166: isothermal at T0 = 273.16, no ice, ϑ_l = ν/2 — only for the
    coupled (SoilEnergyModel + SoilHydrologyModel) model; any other model errors."""
    if not isinstance(model, SoilModel) or model.kind != _abi.LH_MODEL_COUPLED:
        raise RuntimeError("No default IC exist for this type of soil model.")

    def ic(z, m):
        param_set = m.earth_param_set
        T = 273.16
        θ_i = 0.0
        θ_l = 0.5 * m.soil_param_set.ν
        ρc_s = _vhc(θ_l, θ_i, m.soil_param_set.ρc_ds, param_set)
        ρe_int = _vie(θ_i, ρc_s, T, param_set)
        return {"ϑ_l": θ_l, "θ_i": θ_i, "ρe_int": ρe_int}

    return initialize_states(model, ic, 0.0)

# LandHydrologyB200.jl — the ccall binding a LandHydrology.jl maintainer would add.
#
# NOT EXECUTED in this repository's CI (no Julia in the build image); the tested twin is the
# Python/ctypes host in landhydrology.jl_b200/.  It is kept thin on purpose: it only translates the
# reference's own model objects into `lh_soil_config` and forwards `rhs!` / `step!` / `run!`.
#
# Usage inside LandHydrology.jl:
#     using LandHydrologyB200
#     rhs! = LandHydrologyB200.make_rhs(soil_model)              # drop-in for SoilInterface.make_rhs
#     sim  = LandHydrologyB200.Simulation(soil_model, SSPRK33(); Y_init = Y, dt = dt, tspan = (t0, tf), Ya_init = Ya)
#     LandHydrologyB200.step!(sim); LandHydrologyB200.run!(sim)
module LandHydrologyB200

using LandHydrology.SoilInterface
using LandHydrology.SoilInterface: SoilModel, SoilEnergyModel, SoilHydrologyModel,
    PrescribedTemperatureModel, PrescribedHydrologyModel, SoilComponentBC, NoBC, VerticalFlux,
    Dirichlet, FreeDrainage
using LandHydrology.SoilInterface.SoilWaterParameterizations: NoEffect, TemperatureDependentViscosity, IceImpedance
using CLIMAParameters.Planet: ρ_cloud_liq, ρ_cloud_ice, cp_l, cp_i, T_0, LH_f0
using CLIMAParameters.Atmos.Microphysics: K_therm
import OrdinaryDiffEq        # only for the method types (SSPRK33, CarpenterKennedy2N54, ...) Simulation dispatches on

const LIB = get(ENV, "LH_SOIL_LIBRARY", "liblh_soil.so")

# ---- mirror of include/lh_soil.h (field order and types must match) ---------------------------
struct LhSoilParams
    nu::Cdouble; S_s::Cdouble; nu_ss_gravel::Cdouble; nu_ss_om::Cdouble; nu_ss_quartz::Cdouble
    rho_c_ds::Cdouble; kappa_solid::Cdouble; rho_p::Cdouble; kappa_sat_unfrozen::Cdouble
    kappa_sat_frozen::Cdouble; a::Cdouble; b::Cdouble; kappa_dry_parameter::Cdouble
    z_0m::Cdouble; z_0s::Cdouble
    vg_n::Cdouble; vg_alpha::Cdouble; vg_m::Cdouble; theta_r::Cdouble; Ksat::Cdouble
    viscosity_factor::Int32; impedance_factor::Int32
    visc_gamma::Cdouble; visc_T_ref::Cdouble; imp_Omega::Cdouble
    rho_cloud_liq::Cdouble; rho_cloud_ice::Cdouble; cp_l::Cdouble; cp_i::Cdouble
    T_0::Cdouble; LH_f0::Cdouble; K_therm::Cdouble
end

struct LhSoilFaceBC
    energy_kind::Int32; hydrology_kind::Int32; energy_value::Cdouble; hydrology_value::Cdouble
end

struct LhSoilConfig
    struct_size::Int32; device::Int32; ncol::Int64; nlayer::Int32; model::Int32
    zmin::Cdouble; zmax::Cdouble
    params::LhSoilParams; top::LhSoilFaceBC; bottom::LhSoilFaceBC
    flags::Int32; reserved::Int32
end

const LH_MODEL_RICHARDS, LH_MODEL_HEAT, LH_MODEL_COUPLED = Int32(0), Int32(1), Int32(2)
const LH_BC_NONE, LH_BC_FLUX, LH_BC_DIRICHLET, LH_BC_FREE_DRAINAGE = Int32(0), Int32(1), Int32(2), Int32(3)
const LH_FIELD = Dict(:ϑ_l => Int32(0), :θ_i => Int32(1), :ρe_int => Int32(2), :T => Int32(3))

function check(ctx, status)
    status == 0 && return nothing
    msg = unsafe_string(ccall((:lh_soil_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx))
    status == -3 && throw(MethodError(SoilInterface.vertical_flux, (msg,)))     # LH_ERR_UNSUPPORTED_BC
    status == -2 && throw(AssertionError(msg))                                   # LH_ERR_DOMAIN
    status == -7 && throw(DomainError(NaN, msg))                                 # LH_ERR_NONFINITE
    error("lh_soil status $status: $msg")
end

model_kind(::PrescribedTemperatureModel, ::SoilHydrologyModel) = LH_MODEL_RICHARDS
model_kind(::SoilEnergyModel, ::PrescribedHydrologyModel) = LH_MODEL_HEAT
model_kind(::SoilEnergyModel, ::SoilHydrologyModel) = LH_MODEL_COUPLED

bc_pair(::NoBC, t) = (LH_BC_NONE, 0.0)
bc_pair(bc::VerticalFlux, t) = (LH_BC_FLUX, Float64(bc.flux))
bc_pair(bc::Dirichlet, t) = (LH_BC_DIRICHLET, Float64(bc.state_value(t)))
bc_pair(::FreeDrainage, t) = (LH_BC_FREE_DRAINAGE, 0.0)
face_bc(bc::SoilComponentBC, t) = LhSoilFaceBC(bc_pair(bc.energy, t)[1], bc_pair(bc.hydrology, t)[1],
                                               bc_pair(bc.energy, t)[2], bc_pair(bc.hydrology, t)[2])
bc_values(model, t) = Float64[bc_pair(model.boundary_conditions.top.energy, t)[2],
                              bc_pair(model.boundary_conditions.top.hydrology, t)[2],
                              bc_pair(model.boundary_conditions.bottom.energy, t)[2],
                              bc_pair(model.boundary_conditions.bottom.hydrology, t)[2]]

function params(model::SoilModel)
    sp, ep = model.soil_param_set, model.earth_param_set
    hyd = model.hydrology_model
    hm, visc, imp = hyd isa SoilHydrologyModel ?
        (hyd.hydraulic_model, hyd.viscosity_factor, hyd.impedance_factor) :
        (SoilInterface.SoilWaterParameterizations.vanGenuchten{Float64}(), NoEffect{Float64}(), NoEffect{Float64}())
    LhSoilParams(sp.ν, sp.S_s, sp.ν_ss_gravel, sp.ν_ss_om, sp.ν_ss_quartz, sp.ρc_ds, sp.κ_solid, sp.ρp,
        sp.κ_sat_unfrozen, sp.κ_sat_frozen, sp.a, sp.b, sp.κ_dry_parameter, sp.z_0m, sp.z_0s,
        hm.n, hm.α, hm.m, hm.θr, hm.Ksat,
        Int32(visc isa TemperatureDependentViscosity), Int32(imp isa IceImpedance),
        visc isa TemperatureDependentViscosity ? visc.γ : 2.64e-2,
        visc isa TemperatureDependentViscosity ? visc.T_ref : 288.0,
        imp isa IceImpedance ? imp.Ω : 7.0,
        ρ_cloud_liq(ep), ρ_cloud_ice(ep), cp_l(ep), cp_i(ep), T_0(ep), LH_f0(ep), K_therm(ep))
end

mutable struct Engine
    ctx::Ptr{Cvoid}
    model::SoilModel
    n::Int
end

function Engine(model::SoilModel, t0; device = 0, ncol = 1)
    dom = model.domain
    cfg = LhSoilConfig(Int32(sizeof(LhSoilConfig)), Int32(device), Int64(ncol), Int32(dom.nelements),
        model_kind(model.energy_model, model.hydrology_model), dom.zlim[1], dom.zlim[2], params(model),
        face_bc(model.boundary_conditions.top, t0), face_bc(model.boundary_conditions.bottom, t0), Int32(0), Int32(0))
    out = Ref{Ptr{Cvoid}}(C_NULL)
    st = ccall((:lh_soil_create, LIB), Int32, (Ref{LhSoilConfig}, Ref{Ptr{Cvoid}}), cfg, out)
    check(C_NULL, st)
    e = Engine(out[], model, Int(dom.nelements))
    finalizer(x -> ccall((:lh_soil_destroy, LIB), Int32, (Ptr{Cvoid},), x.ctx), e)
    return e
end

# parent(field) of a single column is an n-vector: col_stride = 0 (one column), layer_stride = 1
function upload!(e::Engine, Y)
    for name in propertynames(Y.soil)
        a = vec(parent(getproperty(Y.soil, name)))
        GC.@preserve a check(e.ctx, ccall((:lh_soil_set_state, LIB), Int32,
            (Ptr{Cvoid}, Int32, Ptr{Cdouble}, Int64, Int64), e.ctx, LH_FIELD[name], a, 0, 1))
    end
end

function download!(e::Engine, Y, fn::Symbol)
    for name in propertynames(Y.soil)
        a = vec(parent(getproperty(Y.soil, name)))
        GC.@preserve a check(e.ctx, ccall((fn, LIB), Int32,
            (Ptr{Cvoid}, Int32, Ptr{Cdouble}, Int64, Int64), e.ctx, LH_FIELD[name], a, 0, 1))
    end
end

function update_aux!(e::Engine, Ya, t)
    m = e.model
    zc = vec(parent(Ya.zc))
    if m.energy_model isa PrescribedTemperatureModel
        prof = Float64[m.energy_model.T_profile(z, t) for z in zc]
        parent(Ya.soil.T) .= prof
        GC.@preserve prof check(e.ctx, ccall((:lh_soil_set_aux, LIB), Int32,
            (Ptr{Cvoid}, Int32, Ptr{Cdouble}, Int64, Int64), e.ctx, LH_FIELD[:T], prof, 0, 1))
    end
    if m.hydrology_model isa PrescribedHydrologyModel
        for (name, f) in ((:ϑ_l, m.hydrology_model.ϑ_l_profile), (:θ_i, m.hydrology_model.θ_i_profile))
            prof = Float64[f(z, t) for z in zc]
            parent(getproperty(Ya.soil, name)) .= prof
            GC.@preserve prof check(e.ctx, ccall((:lh_soil_set_aux, LIB), Int32,
                (Ptr{Cvoid}, Int32, Ptr{Cdouble}, Int64, Int64), e.ctx, LH_FIELD[name], prof, 0, 1))
        end
    end
end

"""
    make_rhs(model::SoilModel)

Drop-in for `SoilInterface.make_rhs` (src/SoilModel/right_hand_side.jl:33-44): returns `rhs!(dY, Y, Ya, t)`.
"""
function make_rhs(model::SoilModel)
    engine = Ref{Union{Nothing, Engine}}(nothing)
    function rhs!(dY, Y, Ya, t)
        engine[] === nothing && (engine[] = Engine(model, t))
        e = engine[]
        update_aux!(e, Ya, t)
        v = bc_values(model, t)
        GC.@preserve v check(e.ctx, ccall((:lh_soil_set_bc_values, LIB), Int32, (Ptr{Cvoid}, Ptr{Cdouble}), e.ctx, v))
        upload!(e, Y)
        check(e.ctx, ccall((:lh_soil_rhs, LIB), Int32, (Ptr{Cvoid}, Cdouble), e.ctx, t))
        download!(e, dY, :lh_soil_get_tendency)
        return dY
    end
    return rhs!
end

"""
    set_column_params!(engine; ν, θr, n, α, Ksat)

Heterogeneous soils (new): per-column vectors instead of the scalars of `SoilParams` / `vanGenuchten`; `nothing` keeps
the model's value.
"""
function set_column_params!(e::Engine; ν = nothing, θr = nothing, n = nothing, α = nothing, Ksat = nothing)
    ptr(a) = a === nothing ? Ptr{Cdouble}(C_NULL) : pointer(a)
    arrs = map(a -> a === nothing ? nothing : Vector{Float64}(a), (ν, θr, n, α, Ksat))
    GC.@preserve arrs check(e.ctx, ccall((:lh_soil_set_column_params, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}), e.ctx, map(ptr, arrs)...))
    return nothing
end

# lh_soil_stepper (include/lh_soil.h): coefficient table of a two-register Shu-Osher or a Williamson 2N method
const LH_MAX_STAGES = 16
struct LhSoilStepper
    kind::Int32
    nstages::Int32
    a::NTuple{LH_MAX_STAGES, Float64}
    b::NTuple{LH_MAX_STAGES, Float64}
    g::NTuple{LH_MAX_STAGES, Float64}
    c::NTuple{LH_MAX_STAGES, Float64}
end

"OrdinaryDiffEq method -> built-in table id (LH_METHOD_*); `nothing` selects the specialised SSPRK33 kernels."
method_id(::OrdinaryDiffEq.SSPRK33) = nothing
method_id(::OrdinaryDiffEq.Euler) = Int32(0)
method_id(::OrdinaryDiffEq.SSPRK22) = Int32(1)
method_id(::OrdinaryDiffEq.SSPRK43) = Int32(3)
method_id(::OrdinaryDiffEq.CarpenterKennedy2N54) = Int32(4)

function stepper_table(method)
    id = method_id(method)
    id === nothing && return nothing
    tab = Ref{LhSoilStepper}()
    status = ccall((:lh_soil_stepper_named, LIB), Int32, (Int32, Ref{LhSoilStepper}), id, tab)
    status == 0 || error("lh_soil_stepper_named($id) failed with status $status")
    return tab[]
end

mutable struct Simulation
    model::SoilModel
    engine::Engine
    u
    p
    t::Float64
    dt::Float64
    tf::Float64
    table::Union{Nothing, LhSoilStepper}      # nothing: SSPRK33
end

"""
    Simulation(model, method; Y_init, dt, tspan, Ya_init, ...)

Drop-in for src/Simulations/simulation.jl:34-73; the state stays on the GPU.  `method` is `SSPRK33()` (what every
reference test uses) or one of `Euler()`, `SSPRK22()`, `SSPRK43()`, `CarpenterKennedy2N54()`.
"""
function Simulation(model::SoilModel, method; Y_init, dt, tspan, Ya_init, kwargs...)
    e = Engine(model, tspan[1])
    u = deepcopy(Y_init)
    update_aux!(e, Ya_init, tspan[1])
    upload!(e, u)
    return Simulation(model, e, u, Ya_init, tspan[1], dt, tspan[2], stepper_table(method))
end

function advance!(sim::Simulation, nsteps::Integer)
    cs = sim.table === nothing ? (0.0, 1.0, 0.5) : sim.table.c[1:sim.table.nstages]   # stage times t + c dt
    ns = length(cs)
    table = Vector{Float64}(undef, 4 * ns * nsteps)
    t = sim.t
    for s in 0:(nsteps - 1)
        for (k, c) in enumerate(cs)
            table[(4ns * s + 4(k - 1) + 1):(4ns * s + 4k)] .= bc_values(sim.model, t + c * sim.dt)
        end
        t += sim.dt
    end
    if sim.table === nothing
        GC.@preserve table check(sim.engine.ctx, ccall((:lh_soil_step_ssprk33, LIB), Int32,
            (Ptr{Cvoid}, Cdouble, Cdouble, Int64, Ptr{Cdouble}), sim.engine.ctx, sim.t, sim.dt, nsteps, table))
    else
        GC.@preserve table check(sim.engine.ctx, ccall((:lh_soil_step, LIB), Int32,
            (Ptr{Cvoid}, Ref{LhSoilStepper}, Cdouble, Cdouble, Int64, Ptr{Cdouble}),
            sim.engine.ctx, sim.table, sim.t, sim.dt, nsteps, table))
    end
    sim.t = t
    return nothing
end

"step!(simulation): one step = one fused RHS+stage kernel launch per stage (simulation.jl:79-80)."
step!(sim::Simulation) = advance!(sim, 1)

"run!(simulation): integrate to tspan[2] and bring the state back (simulation.jl:86-87)."
function run!(sim::Simulation)
    advance!(sim, round(Int, (sim.tf - sim.t) / sim.dt))
    download!(sim.engine, sim.u, :lh_soil_get_state)
    return sim.u
end

end # module

# LandHydrologyB200.jl — the ccall binding a LandHydrology.jl maintainer would add.
#
# NOT EXECUTED in this repository's CI (there is no Julia in the build image).  What IS checked here:
#   * tests/test_julia_shim.py parses the struct mirrors below and compares them field for field (name, order, C type)
#     with include/lh_soil.h, and checks that every ccall names an exported symbol with the right argument count;
#   * tests/support/abi_client.c is a compiled, non-Python client that makes the same sequence of calls as this file
#     (create -> set_state in the reference layout, batched -> rhs -> step -> run with snapshots -> get_state).
# The tested twin of the host logic is the Python/ctypes package in landhydrology.jl_b200/.
#
# Usage inside LandHydrology.jl:
#     using LandHydrologyB200
#     rhs! = LandHydrologyB200.make_rhs(soil_model)              # drop-in for SoilInterface.make_rhs
#     sim  = LandHydrologyB200.Simulation(soil_model, SSPRK33(); Y_init = Y, dt = dt, tspan = (t0, tf), Ya_init = Ya,
#                                         saveat = 3600.0, callbacks = nothing)
#     LandHydrologyB200.step!(sim); sol = LandHydrologyB200.run!(sim)          # sol.t, sol.u
# Batched columns (BASELINE config 3/4: a HybridBox is nx * ny laterally independent columns):
#     box = LandHydrologyB200.HybridBox(Float64; xlim, ylim, zlim = (-2.0, 0.0), nelements = (1024, 1024, 64))
#     sim = LandHydrologyB200.Simulation(soil_model, SSPRK33(); domain = box, Y_init = Yb, ...)   # Yb: (n, nfields, ncol) array
module LandHydrologyB200

using LandHydrology.SoilInterface
using LandHydrology.SoilInterface: SoilModel, SoilEnergyModel, SoilHydrologyModel,
    PrescribedTemperatureModel, PrescribedHydrologyModel, SoilComponentBC, NoBC, VerticalFlux,
    Dirichlet, FreeDrainage, PrescribedAtmosForcing
using LandHydrology.SoilInterface.SoilWaterParameterizations: NoEffect, TemperatureDependentViscosity, IceImpedance
using CLIMAParameters.Planet: ρ_cloud_liq, ρ_cloud_ice, cp_l, cp_i, T_0, LH_f0, R_v, R_d, grav, cp_d, cp_v, LH_v0, press_triple, T_triple
using CLIMAParameters.SubgridScale: von_karman_const
using CLIMAParameters.Atmos.Microphysics: K_therm
import OrdinaryDiffEq        # only for the method types (SSPRK33, CarpenterKennedy2N54, ...) Simulation dispatches on

const LIB = get(ENV, "LH_SOIL_LIBRARY", "liblh_soil.so")

# ---- mirror of include/lh_soil.h (field order and types must match; checked by tests/test_julia_shim.py) --------
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

const LH_NUM_FIELDS = 4
struct LhSoilRunOpts
    struct_size::Int32; save_first::Int32
    bc_table::Ptr{Cdouble}
    budget_every::Int64
    budgets_out::Ptr{Cdouble}
    save_every::Int64
    nsave_fields::Int32
    save_fields::NTuple{LH_NUM_FIELDS, Int32}
    reserved::Int32
    save_out::Ptr{Cdouble}
    snapshot_stride::Int64; field_stride::Int64; col_stride::Int64; layer_stride::Int64
end

# lh_soil_atmos: PrescribedAtmosForcing + the CLIMAParameters / Businger constants its fluxes need
struct LhSoilAtmos
    struct_size::Int32; reserved::Int32
    u_atm::Cdouble; theta_atm::Cdouble; z_atm::Cdouble; theta_scale::Cdouble; rho_a_sfc::Cdouble; q_atm::Cdouble
    R_v::Cdouble; R_d::Cdouble; grav::Cdouble; cp_d::Cdouble; cp_v::Cdouble; LH_v0::Cdouble
    press_triple::Cdouble; T_triple::Cdouble; von_karman::Cdouble
    Pr_0::Cdouble; a_m::Cdouble; a_h::Cdouble
end

# lh_soil_stepper: coefficient table of a two-register Shu-Osher or a Williamson 2N method
const LH_MAX_STAGES = 16
struct LhSoilStepper
    kind::Int32
    nstages::Int32
    a::NTuple{LH_MAX_STAGES, Cdouble}
    b::NTuple{LH_MAX_STAGES, Cdouble}
    g::NTuple{LH_MAX_STAGES, Cdouble}
    c::NTuple{LH_MAX_STAGES, Cdouble}
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

# ---- domains ------------------------------------------------------------------------------------------------------
"""
    HybridBox(FT; xlim, ylim, zlim, nelements = (nx, ny, nz))

Not in the reference (src/Domains defines only `Column`): nx * ny laterally independent columns on one vertical mesh.
No reference operator couples columns (right_hand_side.jl uses vertical operators only), so every column evolves exactly
like a `Column` of the same `zlim` / `nz`.
"""
struct HybridBox{FT}
    xlim::Tuple{FT, FT}
    ylim::Tuple{FT, FT}
    zlim::Tuple{FT, FT}
    nelements::Tuple{Int, Int, Int}
end
function HybridBox(::Type{FT}; xlim = (FT(0), FT(1)), ylim = (FT(0), FT(1)), zlim, nelements) where {FT}
    @assert zlim[1] < zlim[2]                    # as Column (domain.jl:30)
    return HybridBox{FT}(FT.(xlim), FT.(ylim), FT.(zlim), Tuple(Int.(nelements)))
end
ncolumns(d::HybridBox) = d.nelements[1] * d.nelements[2]
nlayers(d::HybridBox) = d.nelements[3]
ncolumns(d) = 1                                   # LandHydrology.Domains.Column
nlayers(d) = Int(d.nelements)

model_kind(::PrescribedTemperatureModel, ::SoilHydrologyModel) = LH_MODEL_RICHARDS
model_kind(::SoilEnergyModel, ::PrescribedHydrologyModel) = LH_MODEL_HEAT
model_kind(::SoilEnergyModel, ::SoilHydrologyModel) = LH_MODEL_COUPLED

bc_pair(::NoBC, t) = (LH_BC_NONE, 0.0)
bc_pair(bc::VerticalFlux, t) = (LH_BC_FLUX, Float64(bc.flux))
bc_pair(bc::Dirichlet, t) = (LH_BC_DIRICHLET, Float64(bc.state_value(t)))
bc_pair(::FreeDrainage, t) = (LH_BC_FREE_DRAINAGE, 0.0)
# PrescribedAtmosForcing replaces the whole top SoilComponentBC: as far as the config goes the face is a (per-column) flux,
# the fluxes themselves come from lh_soil_set_atmos_forcing (Engine constructor)
face_bc(::PrescribedAtmosForcing, t) = LhSoilFaceBC(LH_BC_FLUX, LH_BC_FLUX, 0.0, 0.0)
function atmos(model::SoilModel)
    bc, ep = model.boundary_conditions.top, model.earth_param_set
    LhSoilAtmos(Int32(sizeof(LhSoilAtmos)), Int32(0), bc.u_atm, bc.θ_atm, bc.z_atm, bc.θ_scale, bc.ρ_a_sfc, bc.q_atm,
                R_v(ep), R_d(ep), grav(ep), cp_d(ep), cp_v(ep), LH_v0(ep), press_triple(ep), T_triple(ep), von_karman_const(ep),
                0.74, 4.7, 4.7)                      # SurfaceFluxes.UniversalFunctions.Businger: Pr_0, a_m, a_h
end
function face_bc(bc::SoilComponentBC, t)
    (ek, ev), (hk, hv) = bc_pair(bc.energy, t), bc_pair(bc.hydrology, t)       # each closure evaluated once
    return LhSoilFaceBC(ek, hk, ev, hv)
end
function bc_values(model, t)
    top, bot = face_bc(model.boundary_conditions.top, t), face_bc(model.boundary_conditions.bottom, t)    # atmos top: zeros (unused)
    return Float64[top.energy_value, top.hydrology_value, bot.energy_value, bot.hydrology_value]     # LH_BCV_* order
end

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

# ---- engine: one ctx (= one GPU, one stream) ---------------------------------------------------------------------
mutable struct Engine
    ctx::Ptr{Cvoid}
    model::SoilModel
    n::Int           # layers
    ncol::Int        # columns of this ctx (a shard of the domain's columns in a multi-GPU run)
    names::Vector{Symbol}     # prognostic field names in state order (initial_conditions.jl:101-107)
end

function Engine(model::SoilModel, t0; domain = model.domain, device = 0, ncol = ncolumns(domain), flags = 0)
    n = nlayers(domain)
    cfg = LhSoilConfig(Int32(sizeof(LhSoilConfig)), Int32(device), Int64(ncol), Int32(n),
        model_kind(model.energy_model, model.hydrology_model), domain.zlim[1], domain.zlim[2], params(model),
        face_bc(model.boundary_conditions.top, t0), face_bc(model.boundary_conditions.bottom, t0), Int32(flags), Int32(0))
    out = Ref{Ptr{Cvoid}}(C_NULL)
    st = ccall((:lh_soil_create, LIB), Int32, (Ref{LhSoilConfig}, Ref{Ptr{Cvoid}}), cfg, out)
    check(C_NULL, st)
    e = Engine(out[], model, n, Int(ncol), Symbol[])
    finalizer(x -> ccall((:lh_soil_destroy, LIB), Int32, (Ptr{Cvoid},), x.ctx), e)
    if model.boundary_conditions.top isa PrescribedAtmosForcing
        check(e.ctx, ccall((:lh_soil_set_atmos_forcing, LIB), Int32, (Ptr{Cvoid}, Ref{LhSoilAtmos}), e.ctx, atmos(model)))
    end
    return e
end

# State layouts.  A reference state Y (ClimaCore FieldVector of one Column): parent(Y.soil.<name>) is an n-vector;
# a batched state is a dense Array{Float64,3} of size (n, nfields, ncol): per column the reference's n x nfields matrix
# (layer fastest, fields in IC order), columns one after the other -> col_stride = n * nfields, layer_stride = 1.
field_names(Y) = collect(propertynames(Y.soil))
is_batched(Y) = Y isa AbstractArray{Float64, 3}

function upload!(e::Engine, Y, names = field_names_for(e, Y))
    e.names = names
    if is_batched(Y)
        n, nf, ncol = size(Y)
        @assert n == e.n && ncol == e.ncol && nf == length(names)
        for (k, name) in enumerate(names)
            p = pointer(Y, (k - 1) * n + 1)
            GC.@preserve Y check(e.ctx, ccall((:lh_soil_set_state, LIB), Int32,
                (Ptr{Cvoid}, Int32, Ptr{Cdouble}, Int64, Int64), e.ctx, LH_FIELD[name], p, n * nf, 1))
        end
    else
        for name in names
            a = vec(parent(getproperty(Y.soil, name)))
            GC.@preserve a check(e.ctx, ccall((:lh_soil_set_state, LIB), Int32,
                (Ptr{Cvoid}, Int32, Ptr{Cdouble}, Int64, Int64), e.ctx, LH_FIELD[name], a, length(a), 1))
        end
    end
end
field_names_for(e::Engine, Y) = is_batched(Y) ? (isempty(e.names) ? default_names(e.model) : e.names) : field_names(Y)
default_names(m::SoilModel) = m.energy_model isa SoilEnergyModel ?
    (m.hydrology_model isa SoilHydrologyModel ? [:ϑ_l, :θ_i, :ρe_int] : [:ρe_int]) : [:ϑ_l, :θ_i]

function download_state!(e::Engine, Y)
    names = field_names_for(e, Y)
    if is_batched(Y)
        n, nf, _ = size(Y)
        for (k, name) in enumerate(names)
            p = pointer(Y, (k - 1) * n + 1)
            GC.@preserve Y check(e.ctx, ccall((:lh_soil_get_state, LIB), Int32,
                (Ptr{Cvoid}, Int32, Ptr{Cdouble}, Int64, Int64), e.ctx, LH_FIELD[name], p, n * nf, 1))
        end
    else
        for name in names
            a = vec(parent(getproperty(Y.soil, name)))
            GC.@preserve a check(e.ctx, ccall((:lh_soil_get_state, LIB), Int32,
                (Ptr{Cvoid}, Int32, Ptr{Cdouble}, Int64, Int64), e.ctx, LH_FIELD[name], a, length(a), 1))
        end
    end
end

function download_tendency!(e::Engine, dY)
    names = field_names_for(e, dY)
    if is_batched(dY)
        n, nf, _ = size(dY)
        for (k, name) in enumerate(names)
            p = pointer(dY, (k - 1) * n + 1)
            GC.@preserve dY check(e.ctx, ccall((:lh_soil_get_tendency, LIB), Int32,
                (Ptr{Cvoid}, Int32, Ptr{Cdouble}, Int64, Int64), e.ctx, LH_FIELD[name], p, n * nf, 1))
        end
    else
        for name in names
            a = vec(parent(getproperty(dY.soil, name)))
            GC.@preserve a check(e.ctx, ccall((:lh_soil_get_tendency, LIB), Int32,
                (Ptr{Cvoid}, Int32, Ptr{Cdouble}, Int64, Int64), e.ctx, LH_FIELD[name], a, length(a), 1))
        end
    end
end

# Prescribed profiles (make_update_aux, right_hand_side.jl:54-81): one nlayer profile broadcast to all columns.
function zcentres(e::Engine)
    zc = Vector{Float64}(undef, e.n)
    check(e.ctx, ccall((:lh_soil_get_zc, LIB), Int32, (Ptr{Cvoid}, Ptr{Cdouble}), e.ctx, zc))
    return zc
end

function prescribed_profiles(e::Engine, t)
    m, zc = e.model, zcentres(e)
    out = Pair{Symbol, Vector{Float64}}[]
    if m.energy_model isa PrescribedTemperatureModel
        push!(out, :T => Float64[m.energy_model.T_profile(z, t) for z in zc])
    end
    if m.hydrology_model isa PrescribedHydrologyModel
        push!(out, :ϑ_l => Float64[m.hydrology_model.ϑ_l_profile(z, t) for z in zc])
        push!(out, :θ_i => Float64[m.hydrology_model.θ_i_profile(z, t) for z in zc])
    end
    return out
end

function update_aux!(e::Engine, Ya, t)
    for (name, prof) in prescribed_profiles(e, t)
        Ya === nothing || is_batched(Ya) || (parent(getproperty(Ya.soil, name)) .= prof)
        GC.@preserve prof check(e.ctx, ccall((:lh_soil_set_aux, LIB), Int32,
            (Ptr{Cvoid}, Int32, Ptr{Cdouble}, Int64, Int64), e.ctx, LH_FIELD[name], prof, 0, 1))
    end
end

"Time-dependent prescribed profiles: evaluate them AHEAD at the stage times of `nsteps` steps and upload the rows once."
function upload_aux_tables!(e::Engine, t, dt, nsteps, cs)
    times = Float64[t + s * dt + c * dt for s in 0:(nsteps - 1) for c in cs]
    names = first.(prescribed_profiles(e, t))
    for name in names
        rows = Matrix{Float64}(undef, e.n, length(times))          # column-major: one row of the C table per column here
        for (k, tk) in enumerate(times)
            rows[:, k] .= Dict(prescribed_profiles(e, tk))[name]
        end
        GC.@preserve rows check(e.ctx, ccall((:lh_soil_set_aux_table, LIB), Int32,
            (Ptr{Cvoid}, Int32, Ptr{Cdouble}, Int64), e.ctx, LH_FIELD[name], rows, length(times)))
    end
end

function aux_is_time_dependent(e::Engine, t, dt)
    a, b = prescribed_profiles(e, t), prescribed_profiles(e, t + dt / 2)
    return any(x[2] != y[2] for (x, y) in zip(a, b))
end

"""
    make_rhs(model::SoilModel)

Drop-in for `SoilInterface.make_rhs` (src/SoilModel/right_hand_side.jl:33-44): returns `rhs!(dY, Y, Ya, t)`.
"""
function make_rhs(model::SoilModel; domain = model.domain)
    engine = Ref{Union{Nothing, Engine}}(nothing)
    function rhs!(dY, Y, Ya, t)
        engine[] === nothing && (engine[] = Engine(model, t; domain = domain))
        e = engine[]
        update_aux!(e, Ya, t)
        v = bc_values(model, t)
        GC.@preserve v check(e.ctx, ccall((:lh_soil_set_bc_values, LIB), Int32, (Ptr{Cvoid}, Ptr{Cdouble}), e.ctx, v))
        upload!(e, Y)
        check(e.ctx, ccall((:lh_soil_rhs, LIB), Int32, (Ptr{Cvoid}, Cdouble), e.ctx, t))
        download_tendency!(e, dY)
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

"Layered soils: per-cell ν, θr, n, α, Ksat as (n, ncol) matrices (layer fastest), `nothing` keeps the per-column value."
function set_cell_params!(e::Engine; ν = nothing, θr = nothing, n = nothing, α = nothing, Ksat = nothing)
    ptr(a) = a === nothing ? Ptr{Cdouble}(C_NULL) : pointer(a)
    arrs = map(a -> a === nothing ? nothing : Matrix{Float64}(a), (ν, θr, n, α, Ksat))
    GC.@preserve arrs check(e.ctx, ccall((:lh_soil_set_cell_params, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Int64, Int64), e.ctx, map(ptr, arrs)..., e.n, 1))
    return nothing
end

"Per-column heat parameters of SoilParams (ρc_ds, κ_sat_unfrozen, κ_sat_frozen, κ_solid, ν_ss_om, ν_ss_quartz, ν_ss_gravel)."
function set_column_heat_params!(e::Engine; ρc_ds = nothing, κ_sat_unfrozen = nothing, κ_sat_frozen = nothing, κ_solid = nothing,
                                 ν_ss_om = nothing, ν_ss_quartz = nothing, ν_ss_gravel = nothing)
    ptr(a) = a === nothing ? Ptr{Cdouble}(C_NULL) : pointer(a)
    arrs = map(a -> a === nothing ? nothing : Vector{Float64}(a), (ρc_ds, κ_sat_unfrozen, κ_sat_frozen, κ_solid, ν_ss_om, ν_ss_quartz, ν_ss_gravel))
    GC.@preserve arrs check(e.ctx, ccall((:lh_soil_set_column_heat_params, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}), e.ctx, map(ptr, arrs)...))
    return nothing
end

"Spatially varying VerticalFlux values: per-column vectors for the faces of kind VerticalFlux (`nothing`: the scalar)."
function set_column_fluxes!(e::Engine; top_energy = nothing, top_hydrology = nothing, bottom_energy = nothing, bottom_hydrology = nothing)
    arrs = map(a -> a === nothing ? nothing : Vector{Float64}(a), (top_energy, top_hydrology, bottom_energy, bottom_hydrology))
    ptrs = Ptr{Cdouble}[a === nothing ? Ptr{Cdouble}(C_NULL) : pointer(a) for a in arrs]
    GC.@preserve arrs ptrs check(e.ctx, ccall((:lh_soil_set_column_fluxes, LIB), Int32, (Ptr{Cvoid}, Ptr{Ptr{Cdouble}}), e.ctx, ptrs))
    return nothing
end

"compute_turbulent_surface_fluxes (boundary_conditions.jl:555-620) for surface states, evaluated on the device."
function surface_fluxes(e::Engine, ϑ_l::Vector{Float64}, θ_i::Vector{Float64}, T::Vector{Float64})
    heat, water = similar(T), similar(T)
    check(e.ctx, ccall((:lh_soil_atmos_fluxes, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Int64, Ptr{Cdouble}, Ptr{Cdouble}), e.ctx, ϑ_l, θ_i, T, length(T), heat, water))
    return heat, water
end

# ---- multi-GPU: one process per GPU, contiguous column shards, NCCL only for the budgets -------------------------
"128-byte NCCL unique id (create on rank 0, broadcast with the host's own plumbing, e.g. MPI.Bcast!)."
function comm_unique_id()
    id = Vector{UInt8}(undef, 128)
    st = ccall((:lh_soil_comm_unique_id, LIB), Int32, (Ptr{UInt8},), id)
    st == 0 || error("lh_soil_comm_unique_id failed with status $st (is libnccl.so.2 loadable?)")
    return id
end
comm_init!(e::Engine, nranks, rank, id::Vector{UInt8}) =
    check(e.ctx, ccall((:lh_soil_comm_init, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{UInt8}), e.ctx, nranks, rank, id))
"Columns [lo, hi) (0-based, half open) owned by `rank` of `nranks`: contiguous ranges, no halo."
shard_range(ncol, nranks, rank) = (ncol * rank ÷ nranks, ncol * (rank + 1) ÷ nranks)

function budgets(e::Engine; global_sum = false)
    out = Vector{Float64}(undef, 2)
    if global_sum
        check(e.ctx, ccall((:lh_soil_budgets_allreduce, LIB), Int32, (Ptr{Cvoid}, Ptr{Cdouble}), e.ctx, out))
    else
        check(e.ctx, ccall((:lh_soil_budgets, LIB), Int32, (Ptr{Cvoid}, Ptr{Cdouble}), e.ctx, out))
    end
    return out
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

# ---- Simulation / step! / run! (src/Simulations/simulation.jl:34-87) ----------------------------------------------
struct Solution
    t::Vector{Float64}
    u::Vector{Any}
end

mutable struct Simulation
    model::SoilModel
    engine::Engine
    u                                          # host copy of the state (reference FieldVector or (n, nfields, ncol) array)
    p                                          # Ya
    t::Float64
    dt::Float64
    tf::Float64
    table::Union{Nothing, LhSoilStepper}      # nothing: SSPRK33
    saveat::Union{Nothing, Float64}
    callbacks
    sol::Solution
    host_fresh::Bool                           # `u` equals the device state
    dynamic_aux::Bool
end

"""
    Simulation(model, method; Y_init, dt, tspan, Ya_init, saveat = nothing, callbacks = nothing, domain = model.domain, ...)

Drop-in for src/Simulations/simulation.jl:34-73; the state stays on the GPU.  `method` is `SSPRK33()` (what every
reference test uses) or one of `Euler()`, `SSPRK22()`, `SSPRK43()`, `CarpenterKennedy2N54()`.  `saveat` is a spacing in
time (a multiple of `dt`); `callbacks` is a function `cb(sim)` called after every step with `sim.u` current (and re-uploaded
afterwards, so a callback may modify `u`).  `device`, `column_range` select the GPU and the shard of columns of this process.
"""
function Simulation(model::SoilModel, method; Y_init, dt, tspan, Ya_init, saveat = nothing, callbacks = nothing,
                    domain = model.domain, device = 0, column_range = (0, ncolumns(domain)), kwargs...)
    e = Engine(model, tspan[1]; domain = domain, device = device, ncol = column_range[2] - column_range[1])
    u = deepcopy(Y_init)
    update_aux!(e, Ya_init, tspan[1])
    upload!(e, u)
    sim = Simulation(model, e, u, Ya_init, tspan[1], dt, tspan[2], stepper_table(method),
                     saveat === nothing ? nothing : Float64(saveat), callbacks, Solution(Float64[], Any[]), true,
                     aux_is_time_dependent(e, tspan[1], dt))
    push!(sim.sol.t, sim.t); push!(sim.sol.u, deepcopy(u))       # DiffEq save_start
    return sim
end

stage_offsets(sim::Simulation) = sim.table === nothing ? (0.0, 1.0, 0.5) : sim.table.c[1:sim.table.nstages]

function bc_table(sim::Simulation, nsteps)
    cs = stage_offsets(sim)
    ns = length(cs)
    table = Vector{Float64}(undef, 4 * ns * nsteps)
    for s in 0:(nsteps - 1), (k, c) in enumerate(cs)
        table[(4ns * s + 4(k - 1) + 1):(4ns * s + 4k)] .= bc_values(sim.model, sim.t + s * sim.dt + c * sim.dt)
    end
    return table
end

function advance!(sim::Simulation, nsteps::Integer)
    nsteps <= 0 && return nothing
    table = bc_table(sim, nsteps)
    sim.dynamic_aux && upload_aux_tables!(sim.engine, sim.t, sim.dt, nsteps, stage_offsets(sim))
    if sim.table === nothing
        GC.@preserve table check(sim.engine.ctx, ccall((:lh_soil_step_ssprk33, LIB), Int32,
            (Ptr{Cvoid}, Cdouble, Cdouble, Int64, Ptr{Cdouble}), sim.engine.ctx, sim.t, sim.dt, nsteps, table))
    else
        GC.@preserve table check(sim.engine.ctx, ccall((:lh_soil_step, LIB), Int32,
            (Ptr{Cvoid}, Ref{LhSoilStepper}, Cdouble, Cdouble, Int64, Ptr{Cdouble}),
            sim.engine.ctx, sim.table, sim.t, sim.dt, nsteps, table))
    end
    sim.t += nsteps * sim.dt
    sim.host_fresh = false
    return nothing
end

"Bring `sim.u` up to date with the device state (lazy: step! does not download)."
function sync!(sim::Simulation)
    sim.host_fresh || download_state!(sim.engine, sim.u)
    sim.host_fresh = true
    return sim.u
end

"step!(simulation): one step = one fused RHS+stage kernel launch per stage (simulation.jl:79-80)."
function step!(sim::Simulation)
    advance!(sim, 1)
    after_step!(sim)
    return nothing
end

function after_step!(sim::Simulation)
    if sim.callbacks !== nothing
        sync!(sim)
        sim.callbacks(sim)
        upload!(sim.engine, sim.u)               # a callback may have modified u (DiffEq callbacks commonly do)
    end
    if sim.saveat === nothing || isapprox(rem(sim.t - sim.sol.t[1], sim.saveat), 0; atol = 1e-9 * sim.dt) ||
       isapprox(rem(sim.t - sim.sol.t[1], sim.saveat), sim.saveat; atol = 1e-9 * sim.dt)
        push!(sim.sol.t, sim.t); push!(sim.sol.u, deepcopy(sync!(sim)))
    end
end

"""
    run!(simulation)

Integrate to tspan[2] (simulation.jl:86-87).  Without callbacks and with SSPRK33 the whole run, `saveat` snapshots
included, is ONE `lh_soil_run` call: snapshots leave the device on the copy stream while the steps go on.
"""
function run!(sim::Simulation)
    nsteps = round(Int, (sim.tf - sim.t) / sim.dt)
    if sim.callbacks === nothing && sim.table === nothing && is_batched(sim.u) && nsteps > 0
        e = sim.engine
        every = sim.saveat === nothing ? 1 : max(1, round(Int, sim.saveat / sim.dt))
        n, nf, ncol = size(sim.u)
        nsnap = nsteps ÷ every
        snaps = Array{Float64, 4}(undef, n, ncol, nf, nsnap)      # [snapshot][field][col][layer] in C order
        table = bc_table(sim, nsteps)
        sim.dynamic_aux && upload_aux_tables!(e, sim.t, sim.dt, nsteps, stage_offsets(sim))
        fields = ntuple(k -> k <= nf ? LH_FIELD[e.names[k]] : Int32(0), LH_NUM_FIELDS)
        GC.@preserve table snaps begin
            opts = LhSoilRunOpts(Int32(sizeof(LhSoilRunOpts)), Int32(0), pointer(table), 0, Ptr{Cdouble}(C_NULL), every,
                                 Int32(nf), fields, Int32(0), pointer(snaps), n * ncol * nf, n * ncol, n, 1)
            check(e.ctx, ccall((:lh_soil_run, LIB), Int32, (Ptr{Cvoid}, Cdouble, Cdouble, Int64, Ref{LhSoilRunOpts}),
                               e.ctx, sim.t, sim.dt, nsteps, opts))
        end
        for k in 1:nsnap
            push!(sim.sol.t, sim.t + k * every * sim.dt)
            push!(sim.sol.u, permutedims(snaps[:, :, :, k], (1, 3, 2)))     # back to (n, nfields, ncol)
        end
        sim.t += nsteps * sim.dt
        sim.host_fresh = false
    else
        for _ in 1:nsteps
            advance!(sim, 1)
            after_step!(sim)
        end
    end
    sync!(sim)
    return sim.sol
end

end # module

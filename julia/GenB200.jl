# GenB200.jl -- Julia glue: Gen.jl's particle-filter / importance-sampling generics over libgenb200.so.
#
# NOT RUN in this repository's CI: the build image has no Julia (SURVEY.md fact 3).  The Python ctypes
# mirror (gen.jl_b200/api.py) drives exactly the same C ABI (include/genb200.h) and is what the tests
# exercise; this file is the binding a Gen.jl maintainer would load:
#
#     using Gen, GenB200
#     model = GenB200.bearings_model()                        # a DeviceModel <: GenerativeFunction
#     state = initialize_particle_filter(model, (0,), choicemap((:z0, zs[1])), 10^8)
#     for t in 1:T
#         maybe_resample!(state, ess_threshold = N / 2)
#         particle_filter_step!(state, (t,), (UnknownChange(),), choicemap((:chain => t => :z, zs[t+1])))
#     end
#     log_ml_estimate(state)
#
# Julia's multiple dispatch is the plugin mechanism: the methods below are added to Gen's OWN generic
# functions (src/inference/particle_filter.jl:277-279, importance.jl:124) for the device types.
module GenB200

using Gen
import Gen: initialize_particle_filter, particle_filter_step!, maybe_resample!, log_ml_estimate,
            get_log_weights, get_traces, importance_sampling, importance_resampling

const libgenb200 = get(ENV, "GENB200_LIB", "libgenb200.so")

const GB_MODEL_LGSSM, GB_MODEL_SV, GB_MODEL_BEARINGS, GB_MODEL_HMM, GB_MODEL_BLR = 1, 2, 3, 4, 5
const GB_PROPOSAL_DEFAULT, GB_PROPOSAL_AFFINE_NORMAL, GB_PROPOSAL_CATEGORICAL_TABLE = 0, 1, 2
const GB_RESAMPLE_MULTINOMIAL, GB_RESAMPLE_SYSTEMATIC, GB_RESAMPLE_RESIDUAL, GB_RESAMPLE_MULTINOMIAL_SORTED = 0, 1, 2, 3
const GB_F64, GB_F32 = 0, 1

# every entry returns a status; the message replaces Julia's error(...) (particle_filter.jl:227-229)
function check(rc::Cint)
    rc == 0 && return nothing
    error(unsafe_string(ccall((:gb_last_error, libgenb200), Cstring, ())))
end

# ---- DeviceModel: a generative function of the supported class ----
mutable struct DeviceModel <: GenerativeFunction{Any,Trace}
    handle::Ptr{Cvoid}
    kind::Int
    nobs::Int
    state_dim::Int
    function DeviceModel(kind::Int, params::Vector{Float64})
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:gb_model_create, libgenb200), Cint, (Cint, Ptr{Cdouble}, Cint, Ref{Ptr{Cvoid}}),
                    kind, params, length(params), h))
        m = new(h[], kind,
                ccall((:gb_model_num_obs, libgenb200), Cint, (Ptr{Cvoid},), h[]),
                ccall((:gb_model_state_dim, libgenb200), Cint, (Ptr{Cvoid},), h[]))
        finalizer(x -> ccall((:gb_model_free, libgenb200), Cvoid, (Ptr{Cvoid},), x.handle), m)
        m
    end
end

lgssm_model(; m0=0.0, s0=1.0, a=0.9, q=0.5, r=1.0) = DeviceModel(GB_MODEL_LGSSM, [m0, s0, a, q, r])
sv_model(; mu=-1.0, phi=0.95, sigma=0.25) = DeviceModel(GB_MODEL_SV, [mu, phi, sigma, sigma / sqrt(1 - phi^2)])
bearings_model(; x0=(0.01, 0.01), y0=(0.95, 0.01), vx0=(0.002, 0.01), vy0=(-0.013, 0.01),
               velocity_var=1e-6, measurement_noise=0.005) =
    DeviceModel(GB_MODEL_BEARINGS, Float64[x0..., y0..., vx0..., vy0..., velocity_var, measurement_noise])
# layout of test/inference/particle_filter.jl:52-78: transition[z, prev_z], emission[x, z] (column-major = Julia's)
hmm_model(prior::Vector{Float64}, transition::Matrix{Float64}, emission::Matrix{Float64}) =
    DeviceModel(GB_MODEL_HMM, vcat(Float64[length(prior), size(emission, 1)], prior, vec(transition), vec(emission)))
blr_model(X::Matrix{Float64}, sigma::Float64) =
    DeviceModel(GB_MODEL_BLR, vcat(Float64[size(X, 2), size(X, 1), sigma], vec(permutedims(X))))

struct AffineNormalProposal end            # proposal_args = (c0, c1, std)
struct CategoricalTableProposal end        # proposal_args = (table,)
proposal_kind(::Nothing) = GB_PROPOSAL_DEFAULT
proposal_kind(::AffineNormalProposal) = GB_PROPOSAL_AFFINE_NORMAL
proposal_kind(::CategoricalTableProposal) = GB_PROPOSAL_CATEGORICAL_TABLE
flatten_pargs(::Nothing, args) = Float64[]
flatten_pargs(::AffineNormalProposal, args) = Float64[args...]
flatten_pargs(::CategoricalTableProposal, args) = Float64[vec(args[1])...]

# observations cross the boundary as a flat Float64 vector in the model's fixed slot order
# (SURVEY.md section 8b): all leaf values of the choice map in address order
flatten_obs(obs::ChoiceMap) = Float64[Float64(v) for (_, v) in sort(collect(get_values_deep(obs)), by=first)]
function get_values_deep(cm::ChoiceMap, prefix=())
    out = Pair{Any,Any}[]
    for (k, v) in get_values_shallow(cm); push!(out, (prefix..., k) => v); end
    for (k, sub) in get_submaps_shallow(cm); append!(out, get_values_deep(sub, (prefix..., k))); end
    out
end

# ---- DeviceParticleFilterState: ParticleFilterState (particle_filter.jl:35-41) held as SoA columns in HBM ----
mutable struct DeviceParticleFilterState
    handle::Ptr{Cvoid}
    model::DeviceModel
    num_particles::Int
    resampler::Int
    last_t::Union{Int,Nothing}
    function DeviceParticleFilterState(h, model, n, last_t)
        s = new(h, model, n, GB_RESAMPLE_MULTINOMIAL, last_t)   # reference semantics by default (:261-262)
        finalizer(x -> ccall((:gb_pf_free, libgenb200), Cvoid, (Ptr{Cvoid},), x.handle), s)
        s
    end
end

# lazily materialised fields (tests poke .log_weights / .parents / .log_ml_est: particle_filter.jl:176-197)
function Base.getproperty(s::DeviceParticleFilterState, f::Symbol)
    f === :log_weights && return get_log_weights(s)
    f === :log_ml_est && (r = Ref{Cdouble}(); check(ccall((:gb_pf_get_log_ml_est, libgenb200), Cint, (Ptr{Cvoid}, Ref{Cdouble}), getfield(s, :handle), r)); return r[])
    if f === :parents
        out = Vector{Int64}(undef, getfield(s, :num_particles))
        check(ccall((:gb_pf_get_parents, libgenb200), Cint, (Ptr{Cvoid}, Ptr{Int64}), getfield(s, :handle), out))
        return out
    end
    f === :traces && return get_traces(s)
    getfield(s, f)
end

function Base.copy(s::DeviceParticleFilterState)       # particle_filter.jl:43-51
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:gb_pf_clone, libgenb200), Cint, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}), s.handle, h))
    c = DeviceParticleFilterState(h[], s.model, s.num_particles, s.last_t)
    c.resampler = s.resampler
    c
end

# Base.:(==) / Base.hash (particle_filter.jl:52-63): field-wise over traces, new_traces (scratch: not compared here),
# log_weights, log_ml_est, parents.  The columns are read back once per comparison.
function _columns(s::DeviceParticleFilterState)
    n, S = getfield(s, :num_particles), getfield(s, :model).state_dim
    cols = Matrix{Float64}(undef, n, S)
    for f in 1:S
        check(ccall((:gb_pf_get_state, libgenb200), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}), getfield(s, :handle), f - 1, view(cols, :, f)))
    end
    cols
end
function Base.:(==)(a::DeviceParticleFilterState, b::DeviceParticleFilterState)
    getfield(a, :num_particles) == getfield(b, :num_particles) && a.log_ml_est == b.log_ml_est &&
        a.parents == b.parents && isequal(a.log_weights, b.log_weights) && isequal(_columns(a), _columns(b))
end
Base.hash(s::DeviceParticleFilterState, h::UInt) =
    hash(_columns(s), hash(s.log_weights, hash(s.log_ml_est, hash(s.parents, h))))

# ---- Gen's generics ----
function initialize_particle_filter(model::DeviceModel, model_args::Tuple, observations::ChoiceMap,
                                    proposal, proposal_args::Tuple, num_particles::Int;
                                    dtype::Int=GB_F64, seed::Integer=rand(UInt64))
    obs = flatten_obs(observations)
    pa = flatten_pargs(proposal, proposal_args)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:gb_pf_initialize, libgenb200), Cint,
                (Ptr{Cvoid}, Int64, Cint, UInt64, Ptr{Cdouble}, Cint, Cint, Ptr{Cdouble}, Cint, Ref{Ptr{Cvoid}}),
                model.handle, num_particles, dtype, seed, obs, length(obs), proposal_kind(proposal), pa, length(pa), h))
    last_t = (length(model_args) >= 1 && model_args[1] isa Integer) ? Int(model_args[1]) : nothing
    DeviceParticleFilterState(h[], model, num_particles, last_t)
end
initialize_particle_filter(model::DeviceModel, model_args::Tuple, observations::ChoiceMap, num_particles::Int; kw...) =
    initialize_particle_filter(model, model_args, observations, nothing, (), num_particles; kw...)

function particle_filter_step!(state::DeviceParticleFilterState, new_args::Tuple, argdiffs::Tuple,
                               observations::ChoiceMap, proposal=nothing, proposal_args::Tuple=())
    # only extension by one time step is supported; anything else would produce a non-empty discard,
    # which the reference rejects (particle_filter.jl:227-229)
    if state.last_t !== nothing && length(new_args) >= 1 && new_args[1] != state.last_t + 1
        error("Choices were updated or deleted inside particle filter step")
    end
    obs = flatten_obs(observations)
    pa = flatten_pargs(proposal, proposal_args)
    check(ccall((:gb_pf_step, libgenb200), Cint,
                (Ptr{Cvoid}, Ptr{Cdouble}, Cint, Cint, Ptr{Cdouble}, Cint, Ptr{Cdouble}),
                state.handle, obs, length(obs), proposal_kind(proposal), pa, length(pa), C_NULL))
    length(new_args) >= 1 && (state.last_t = Int(new_args[1]))
    (LazyIncrements(state, nothing),)
end

# The (log_incremental_weights,) return value of particle_filter_step! as a vector that stays on the device until
# it is indexed: at N = 1e8 an eager copy would move 800 MB per step that most callers never read.
mutable struct LazyIncrements <: AbstractVector{Float64}
    state::DeviceParticleFilterState
    data::Union{Nothing,Vector{Float64}}
end
Base.size(v::LazyIncrements) = (getfield(v.state, :num_particles),)
function Base.getindex(v::LazyIncrements, i::Int)
    if v.data === nothing
        out = Vector{Float64}(undef, getfield(v.state, :num_particles))
        check(ccall((:gb_pf_get_log_increments, libgenb200), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), getfield(v.state, :handle), out))
        v.data = out
    end
    v.data[i]
end

function maybe_resample!(state::DeviceParticleFilterState; ess_threshold::Real=state.num_particles / 2, verbose=false)
    did = Ref{Cint}(0)
    check(ccall((:gb_pf_maybe_resample, libgenb200), Cint, (Ptr{Cvoid}, Cdouble, Cint, Ref{Cint}),
                state.handle, Float64(ess_threshold), state.resampler, did))
    verbose && println("doing resample: $(did[] != 0)")
    did[] != 0
end

# Resampler of the next maybe_resample! calls: :multinomial (Gen's semantics, particle_filter.jl:261-262), :systematic,
# :residual, or :multinomial_sorted (the same draws' distribution, returned sorted by ancestor: one streaming merge on the
# device instead of N searches; traces are exchangeable after a resample, so nothing downstream can tell the difference).
function set_resampler!(state::DeviceParticleFilterState, kind::Symbol)
    state.resampler = Dict(:multinomial => GB_RESAMPLE_MULTINOMIAL, :systematic => GB_RESAMPLE_SYSTEMATIC,
                           :residual => GB_RESAMPLE_RESIDUAL, :multinomial_sorted => GB_RESAMPLE_MULTINOMIAL_SORTED)[kind]
    state
end

function log_ml_estimate(state::DeviceParticleFilterState)
    r = Ref{Cdouble}()
    check(ccall((:gb_pf_log_ml_estimate, libgenb200), Cint, (Ptr{Cvoid}, Ref{Cdouble}), state.handle, r))
    r[]
end

function get_log_weights(state::DeviceParticleFilterState)
    out = Vector{Float64}(undef, getfield(state, :num_particles))
    check(ccall((:gb_pf_get_log_weights, libgenb200), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), getfield(state, :handle), out))
    out
end

# traces of the supported class are rows of the SoA store: a lazy column view
struct DeviceTraces; state::DeviceParticleFilterState; end
get_traces(state::DeviceParticleFilterState) = DeviceTraces(state)
function column(t::DeviceTraces, field::Int)
    out = Vector{Float64}(undef, t.state.num_particles)
    check(ccall((:gb_pf_get_state, libgenb200), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}), t.state.handle, field, out))
    out
end

function importance_sampling(model::DeviceModel, model_args::Tuple, observations::ChoiceMap,
                             proposal, proposal_args::Tuple, num_samples::Int;
                             verbose=false, dtype::Int=GB_F64, seed::Integer=rand(UInt64))
    obs = flatten_obs(observations)
    pa = flatten_pargs(proposal, proposal_args)
    lnw = Vector{Float64}(undef, num_samples)
    lml = Ref{Cdouble}()
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:gb_importance_sampling, libgenb200), Cint,
                (Ptr{Cvoid}, Int64, Cint, UInt64, Ptr{Cdouble}, Cint, Cint, Ptr{Cdouble}, Cint,
                 Ptr{Cdouble}, Cint, Ptr{Cdouble}, Cint, Ptr{Cdouble}, Ref{Cdouble}, Ref{Ptr{Cvoid}}),
                model.handle, num_samples, dtype, seed, obs, length(obs), proposal_kind(proposal), pa, length(pa),
                C_NULL, 0, C_NULL, 0, lnw, lml, h))
    (DeviceTraces(DeviceParticleFilterState(h[], model, num_samples, nothing)), lnw, lml[])
end
importance_sampling(model::DeviceModel, model_args::Tuple, observations::ChoiceMap, num_samples::Int; kw...) =
    importance_sampling(model, model_args, observations, nothing, (), num_samples; kw...)

# importance_resampling (importance.jl:77-122): (trace, lml_est); streaming on the device, O(chunk) memory for any
# num_samples (gb_importance_resampling folds chunks by the reference's logsumexp(x1, x2) + Bernoulli swap)
function importance_resampling(model::DeviceModel, model_args::Tuple, observations::ChoiceMap,
                               proposal, proposal_args::Tuple, num_samples::Int;
                               verbose=false, dtype::Int=GB_F64, seed::Integer=rand(UInt64), chunk::Int=0)
    obs = flatten_obs(observations)
    pa = flatten_pargs(proposal, proposal_args)
    row = Vector{Float64}(undef, model.state_dim)
    lml = Ref{Cdouble}()
    check(ccall((:gb_importance_resampling, libgenb200), Cint,
                (Ptr{Cvoid}, Int64, Cint, UInt64, Ptr{Cdouble}, Cint, Cint, Ptr{Cdouble}, Cint, Int64, Ptr{Cdouble}, Ref{Cdouble}),
                model.handle, num_samples, dtype, seed, obs, length(obs), proposal_kind(proposal), pa, length(pa), chunk, row, lml))
    (row, lml[])
end
importance_resampling(model::DeviceModel, model_args::Tuple, observations::ChoiceMap, num_samples::Int; kw...) =
    importance_resampling(model, model_args, observations, nothing, (), num_samples; kw...)

# tail of enumerative_inference (enumerative.jl:39-46) on the device: (log_normalized_weights, log_total_weight)
function enumerative_normalize(log_weights::AbstractArray{Float64}, log_vols::Union{AbstractArray{Float64},Nothing}=nothing)
    lw = vec(collect(log_weights))
    out = similar(lw)
    tot = Ref{Cdouble}()
    check(ccall((:gb_enumerative_normalize, libgenb200), Cint, (Ptr{Cdouble}, Ptr{Cdouble}, Int64, Cint, Ptr{Cdouble}, Ref{Cdouble}),
                lw, log_vols === nothing ? C_NULL : vec(collect(log_vols)), length(lw), GB_F64, out, tot))
    (reshape(out, size(log_weights)), tot[])
end

# ---- ONE filter over several GPUs from this (single) Julia thread: the gb_spf_* entries (no NCCL / MPI; the devices
#      exchange through peer memory inside the kernels).  Same generics, same results bit for bit as on one GPU. ----
mutable struct MultiGpuParticleFilterState
    handle::Ptr{Cvoid}
    model::DeviceModel
    num_particles::Int
    last_t::Union{Int,Nothing}
    function MultiGpuParticleFilterState(h, model, n, last_t)
        s = new(h, model, n, last_t)
        finalizer(x -> ccall((:gb_spf_free, libgenb200), Cvoid, (Ptr{Cvoid},), x.handle), s)
        s
    end
end
function initialize_particle_filter(model::DeviceModel, model_args::Tuple, observations::ChoiceMap, num_particles::Int,
                                    devices::Vector{<:Integer}; dtype::Int=GB_F64, seed::Integer=rand(UInt64))
    obs = flatten_obs(observations)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    dev = Cint.(devices)
    check(ccall((:gb_spf_initialize, libgenb200), Cint,
                (Ptr{Cvoid}, Int64, Cint, UInt64, Cint, Ptr{Cint}, Ptr{Cdouble}, Cint, Cint, Ptr{Cdouble}, Cint, Ref{Ptr{Cvoid}}),
                model.handle, num_particles, dtype, seed, length(dev), dev, obs, length(obs), GB_PROPOSAL_DEFAULT, C_NULL, 0, h))
    last_t = (length(model_args) >= 1 && model_args[1] isa Integer) ? Int(model_args[1]) : nothing
    MultiGpuParticleFilterState(h[], model, num_particles, last_t)
end
function particle_filter_step!(state::MultiGpuParticleFilterState, new_args::Tuple, argdiffs::Tuple, observations::ChoiceMap)
    if state.last_t !== nothing && length(new_args) >= 1 && new_args[1] != state.last_t + 1
        error("Choices were updated or deleted inside particle filter step")          # particle_filter.jl:227-229
    end
    obs = flatten_obs(observations)
    check(ccall((:gb_spf_step, libgenb200), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Cint, Cint, Ptr{Cdouble}, Cint),
                state.handle, obs, length(obs), GB_PROPOSAL_DEFAULT, C_NULL, 0))
    length(new_args) >= 1 && (state.last_t = Int(new_args[1]))
    lazy = Vector{Float64}(undef, 0)     # log incremental weights stay on the devices (gb_spf_get_log_increments reads them)
    (lazy,)
end
function maybe_resample!(state::MultiGpuParticleFilterState; ess_threshold::Real=state.num_particles / 2, verbose=false)
    did = Ref{Cint}(0)
    check(ccall((:gb_spf_maybe_resample, libgenb200), Cint, (Ptr{Cvoid}, Cdouble, Ref{Cint}), state.handle, Float64(ess_threshold), did))
    did[] != 0
end
function log_ml_estimate(state::MultiGpuParticleFilterState)
    r = Ref{Cdouble}()
    check(ccall((:gb_spf_log_ml_estimate, libgenb200), Cint, (Ptr{Cvoid}, Ref{Cdouble}), state.handle, r))
    r[]
end
function get_log_weights(state::MultiGpuParticleFilterState)
    out = Vector{Float64}(undef, state.num_particles)
    check(ccall((:gb_spf_get_log_weights, libgenb200), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), state.handle, out))
    out
end
function particle_filter_run!(state::MultiGpuParticleFilterState, observations::Vector{<:ChoiceMap};
                              ess_threshold::Real=state.num_particles / 2)
    obs = reduce(vcat, [flatten_obs(o) for o in observations])
    nres = Ref{Cint}(0)
    T = length(observations)
    check(ccall((:gb_spf_run, libgenb200), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Cint, Cdouble, Ref{Cint}),
                state.handle, T, obs, div(length(obs), T), Float64(ess_threshold), nres))
    state.last_t !== nothing && (state.last_t += T)
    nres[]
end

# sample_unweighted_traces (particle_filter.jl:101-109): 1-based indices of num_samples iid draws + their rows
function sample_unweighted_traces(state::DeviceParticleFilterState, num_samples::Int; draw::Integer=0)
    idx = Vector{Int64}(undef, num_samples)
    rows = Matrix{Float64}(undef, num_samples, state.model.state_dim)
    check(ccall((:gb_pf_sample_unweighted, libgenb200), Cint, (Ptr{Cvoid}, Int64, UInt64, Ptr{Int64}, Ptr{Cdouble}),
                state.handle, num_samples, UInt64(draw), idx, rows))
    (idx, rows)
end

# The tutorial's rejuvenation loop (docs/src/tutorials/smc.md:459-463, :518-520)
#     for i in 1:N  state.traces[i], _ = mh(state.traces[i], ...)  end
# for moves on the latest latent choices: move = :resimulate is mh(trace, selection) (mh.jl:16-27),
# move = :drift is mh(trace, gaussian_drift_proposal, (drift_std,)) (mh.jl:37-58).  Returns the number accepted.
function rejuvenate!(state::DeviceParticleFilterState, observations::ChoiceMap; move::Symbol=:resimulate, drift_std::Real=0.0)
    obs = flatten_obs(observations)
    acc = Ref{Int64}(0)
    check(ccall((:gb_pf_rejuvenate, libgenb200), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Cint, Cdouble, Ref{Int64}, Ptr{Cdouble}),
                state.handle, move == :drift ? 1 : 0, obs, length(obs), Float64(drift_std), acc, C_NULL))
    acc[]
end

# T x { maybe_resample!; particle_filter_step! [; sweeps x mh on the newest latents] } without a host round trip per step
# (gb_pf_run / gb_pf_run_mh): the loops of docs/src/tutorials/smc.md:674-689 and :513-531 as one call.
# Returns the number of resamples (and, with rejuvenation, the number of accepted moves).
function particle_filter_run!(state::DeviceParticleFilterState, observations::Vector{<:ChoiceMap};
                              ess_threshold::Real=state.num_particles / 2,
                              move::Union{Symbol,Nothing}=nothing, drift_std::Real=0.0, sweeps::Int=1)
    obs = reduce(vcat, [flatten_obs(o) for o in observations])
    nres = Ref{Cint}(0)
    T = length(observations)
    if move === nothing
        check(ccall((:gb_pf_run, libgenb200), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Cint, Cdouble, Ref{Cint}),
                    state.handle, T, obs, div(length(obs), T), Float64(ess_threshold), nres))
        return nres[]
    end
    nacc = Ref{Int64}(0)
    check(ccall((:gb_pf_run_mh, libgenb200), Cint,
                (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Cint, Cdouble, Cint, Cdouble, Cint, Ref{Cint}, Ref{Int64}),
                state.handle, T, obs, div(length(obs), T), Float64(ess_threshold), move == :drift ? 1 : 0, Float64(drift_std), sweeps,
                nres, nacc))
    (nres[], nacc[])
end

# ---- GB_MODEL_SPEC: a static-IR function of the supported class written out as SSA ops (include/genb200.h) ----
# Gen's own IR cannot be lowered automatically (JuliaNode.fn is an opaque closure, src/static_ir/dag.jl:14-19), so the
# program is built with this small builder; every method returns the register holding the node's value.
mutable struct SpecProgram
    ops::Vector{NTuple{5,Float64}}
    nreg::Int
    SpecProgram() = new(NTuple{5,Float64}[], 0)
end
function _emit!(p::SpecProgram, code, dst, a=0, b=0, imm=0.0)
    push!(p.ops, (Float64(code), Float64(dst), Float64(a), Float64(b), Float64(imm))); dst
end
_new!(p::SpecProgram) = (p.nreg >= 32 && error("a spec program may use at most 32 registers"); p.nreg += 1; p.nreg - 1)
spec_const(p, v) = _emit!(p, 0, _new!(p), 0, 0, v)
spec_prev(p, field) = _emit!(p, 1, _new!(p), field)
spec_param(p, idx) = _emit!(p, 2, _new!(p), idx)
spec_obs(p, idx=0) = _emit!(p, 3, _new!(p), idx)
spec_time(p) = _emit!(p, 4, _new!(p))
for (name, code) in ((:add, 10), (:sub, 11), (:mul, 12), (:div, 13), (:atan2, 18))
    @eval $(Symbol("spec_", name))(p, a, b) = _emit!(p, $code, _new!(p), a, b)
end
for (name, code) in ((:neg, 14), (:sqrt, 15), (:exp, 16), (:log, 17), (:abs, 19))
    @eval $(Symbol("spec_", name))(p, a) = _emit!(p, $code, _new!(p), a)
end
sample_normal!(p, mu, std) = _emit!(p, 20, _new!(p), mu, std)          # {addr} ~ normal(mu, std), unconstrained
observe_normal!(p, x, mu, std) = (_emit!(p, 21, x, mu, std); nothing)  # constrained: weight += logpdf
sample_bernoulli!(p, q) = _emit!(p, 22, _new!(p), q)
observe_bernoulli!(p, x, q) = (_emit!(p, 23, x, q); nothing)
sample_gamma!(p, shape, scale) = _emit!(p, 24, _new!(p), shape, scale)
observe_gamma!(p, x, shape, scale) = (_emit!(p, 25, x, shape, scale); nothing)
sample_categorical!(p, offset_reg, len) = _emit!(p, 26, _new!(p), offset_reg, len)
observe_categorical!(p, x, offset_reg, len) = (_emit!(p, 27, x, offset_reg, len); nothing)
store!(p, field, reg) = (_emit!(p, 30, field, reg); nothing)
add_weight!(p, reg) = (_emit!(p, 31, 0, reg); nothing)
logpdf_normal!(p, x, mu, std) = _emit!(p, 32, _new!(p), mu, std, x)
sample_dist!(p, dist::Int, a, b=a) = _emit!(p, 40, _new!(p), a, b, dist)          # dist = GB_DIST_* with a quantile function
observe_dist!(p, dist::Int, x, a, b=a) = (_emit!(p, 41, x, a, b, dist); nothing)

const GB_MODEL_SPEC = 6
"A DeviceModel from its `generate` program (t = 0) and its Unfold kernel program (extend-by-one `update`)."
function spec_model(state_dim::Int, params::Vector{Float64}, num_obs::Int, generate::SpecProgram, step::Union{SpecProgram,Nothing}=nothing)
    sops = step === nothing ? NTuple{5,Float64}[] : step.ops
    flat = Float64[state_dim, length(params), num_obs, length(generate.ops), length(sops)]
    append!(flat, params)
    for op in vcat(generate.ops, sops); append!(flat, op); end
    DeviceModel(GB_MODEL_SPEC, flat)
end

end # module

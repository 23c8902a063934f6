# BNuts.jl — Julia host side of the B200 batched-chain NUTS engine.
#
# Thin `ccall` bindings over include/bnuts.h behind InplaceDHMC.jl's names
# (export list src/InplaceDHMC.jl:3-11).  UNTESTED — NOT EXECUTED ANYWHERE IN THIS REPOSITORY:
# the build image has no Julia; the Python twin (../api.py) drives the same C ABI
# and is what the tests exercise.  Every call below maps 1:1 to a tested C entry point;
# tests/test_julia_binding_static.py checks this file's ccall signatures against include/bnuts.h.
module BNuts

export GaussianKineticEnergy, TuningNUTS, mcmc_with_warmup, threaded_mcmc, default_warmup_stages,
       NoProgressReport, LogProgressReport, NUTS, DualAveraging, InitialStepsizeSearch, FindLocalOptimum,
       TreeStatisticsNUTS, InvalidTree, IIDNormal, Funnel, GaussianTarget, LogisticTarget

const libbnuts = get(ENV, "BNUTS_LIB", joinpath(@__DIR__, "..", "csrc", "libbnuts.so"))

# ≙ InvalidTree / TreeStatisticsNUTS, src/tree.jl:278-281, src/NUTS.jl:229-242 — identical 32-byte layout,
# so the buffer filled by bnuts_sample is a Matrix{TreeStatisticsNUTS} without conversion.
struct InvalidTree
    left::Int32
    right::Int32
end
struct TreeStatisticsNUTS
    π::Float64
    acceptance_rate::Float64
    termination::InvalidTree
    depth::Int32
    steps::Int32
end

struct bnuts_config
    struct_size::Int32; dtype::Int32; n_chains::Int32; dim::Int32; max_depth::Int32; device::Int32
    min_delta::Float64; seed::UInt64; chain_offset::Int32; gradient_path::Int32
end
struct bnuts_dual_averaging
    delta::Float64; gamma::Float64; kappa::Float64; t0::Int32; _pad::Int32
end
struct bnuts_stepsize_search
    a_min::Float64; a_max::Float64; eps0::Float64; C::Float64; maxiter_crossing::Int32; maxiter_bisect::Int32
end

Base.@kwdef struct NUTS; max_depth::Int = 10; min_Δ::Float64 = -1000.0; end                    # src/NUTS.jl:204-220
Base.@kwdef struct DualAveraging; δ::Float64 = 0.8; γ::Float64 = 0.05; κ::Float64 = 0.75; t₀::Int = 10; end  # src/stepsize.jl:191
Base.@kwdef struct InitialStepsizeSearch                                                          # src/stepsize.jl:29-30
    a_min::Float64 = 0.25; a_max::Float64 = 0.75; ϵ₀::Float64 = 1.0; C::Float64 = 2.0
    maxiter_crossing::Int = 400; maxiter_bisect::Int = 400
end
Base.@kwdef struct FindLocalOptimum; magnitude_penalty::Float64 = 1e-4; iterations::Int = 50; end # src/warmup.jl:137-150
struct FixedStepsize end                                                                           # src/stepsize.jl:251-255
struct TuningNUTS{M}                                                                               # src/warmup.jl:217-234
    N::Int; stepsize_adaptation::Union{DualAveraging,FixedStepsize}; λ::Float64
    # inner constructor (an outer method of the same signature would call itself): λ defaults to 5/N, src/warmup.jl:228-229;
    # the reference records N ≥ 20, λ ≥ 0 as commented-out @argcheck's (src/warmup.jl:230-231); enforced here and in the
    # C ABI: λ ≥ 0, and N ≥ 2 when a metric is adapted (the variance needs two draws)
    function TuningNUTS{M}(N::Integer, da::Union{DualAveraging,FixedStepsize}, λ::Real = 5.0 / N) where {M}
        N ≥ (M === Nothing ? 1 : 2) || throw(ArgumentError("TuningNUTS needs N ≥ 2 to adapt a metric"))
        λ ≥ 0 || throw(ArgumentError("TuningNUTS needs λ ≥ 0"))
        new{M}(Int(N), da, Float64(λ))
    end
end
fixed_stepsize_warmup_stages(; local_optimization = FindLocalOptimum(), M = :Diagonal, middle_steps = 25, doubling_stages = 5) =
    (local_optimization, ntuple(d -> TuningNUTS{M}(middle_steps << (d - 1), FixedStepsize()), doubling_stages)...)   # src/warmup.jl:383-389
Base.length(t::TuningNUTS) = t.N
struct GaussianKineticEnergy; M⁻¹::Matrix{Float64}; end     # [D, nchains] diagonals, src/hamiltonian.jl:33-38
struct NoProgressReport end
struct LogProgressReport; step_interval::Int; end

struct IIDNormal; D::Int; end
struct Funnel; D::Int; end
struct GaussianTarget; P::Matrix{Float64}; end
struct LogisticTarget; X::Matrix{Float32}; y::Vector{Float64}; prior_precision::Float64; end   # X is [D, N] (row-major [N][D] for C)
dimension(ℓ::Union{IIDNormal,Funnel}) = ℓ.D
dimension(ℓ::GaussianTarget) = size(ℓ.P, 1)
dimension(ℓ::LogisticTarget) = size(ℓ.X, 1)

function default_warmup_stages(; local_optimization = FindLocalOptimum(), stepsize_search = InitialStepsizeSearch(),
                               M = :Diagonal, stepsize_adaptation = DualAveraging(), init_steps = 75, middle_steps = 25,
                               doubling_stages = 5, terminating_steps = 50)                       # src/warmup.jl:361-372
    (local_optimization, stepsize_search, TuningNUTS{Nothing}(init_steps, stepsize_adaptation),
     ntuple(d -> TuningNUTS{M}(middle_steps << (d - 1), stepsize_adaptation), doubling_stages)...,
     TuningNUTS{Nothing}(terminating_steps, stepsize_adaptation))
end

check(e, rc) = rc == 0 ? nothing :
    error("bnuts error $rc: " * unsafe_string(ccall((:bnuts_last_error, libbnuts), Cstring, (Ptr{Cvoid},), e)))

attach!(e, ::IIDNormal) = check(e, ccall((:bnuts_model_iid_normal, libbnuts), Int32, (Ptr{Cvoid},), e))
attach!(e, ::Funnel) = check(e, ccall((:bnuts_model_funnel, libbnuts), Int32, (Ptr{Cvoid},), e))
attach!(e, ℓ::GaussianTarget) = check(e, ccall((:bnuts_model_gaussian, libbnuts), Int32, (Ptr{Cvoid}, Ptr{Float64}), e, ℓ.P))
# synthetic rows [row_offset, row_offset + N) of the benchmark design matrix, generated on the device (include/bnuts.h)
struct SyntheticLogisticTarget; D::Int; data_seed::UInt64; row_offset::Int64; N::Int64; prior_precision::Float64; end
dimension(ℓ::SyntheticLogisticTarget) = ℓ.D
attach!(e, ℓ::SyntheticLogisticTarget) = check(e, ccall((:bnuts_model_logistic_synthetic, libbnuts), Int32,
    (Ptr{Cvoid}, UInt64, Int64, Int64, Float64, Int32), e, ℓ.data_seed, ℓ.row_offset, ℓ.N, ℓ.prior_precision, 1))
attach!(e, ℓ::LogisticTarget) = check(e, ccall((:bnuts_model_logistic, libbnuts), Int32,
    (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Float64}, Int64, Float64, Int32), e, ℓ.X, 1, ℓ.y, length(ℓ.y), ℓ.prior_precision, 1))

warmup!(e, ::Nothing) = nothing
warmup!(e, s::FindLocalOptimum) =                                                                 # ≙ src/warmup.jl:152-186
    check(e, ccall((:bnuts_find_local_optimum, libbnuts), Int32, (Ptr{Cvoid}, Float64, Int32), e, s.magnitude_penalty, s.iterations))
function warmup!(e, s::InitialStepsizeSearch)                                                     # ≙ src/warmup.jl:188-200
    p = Ref(bnuts_stepsize_search(s.a_min, s.a_max, s.ϵ₀, s.C, s.maxiter_crossing, s.maxiter_bisect))
    check(e, ccall((:bnuts_find_initial_stepsize, libbnuts), Int32, (Ptr{Cvoid}, Ref{bnuts_stepsize_search}), e, p))
end
function warmup!(e, t::TuningNUTS{M}) where {M}                                                   # ≙ src/warmup.jl:269-314
    da = t.stepsize_adaptation
    if da isa FixedStepsize                                                                      # ≙ src/stepsize.jl:251-255: NULL keeps ϵ
        return check(e, ccall((:bnuts_warmup_stage, libbnuts), Int32,
            (Ptr{Cvoid}, Int32, Int32, Ptr{Cvoid}, Float64, Ptr{Float64}, Int64, Int64, Ptr{Cvoid}, Int64, Ptr{Float64}),
            e, t.N, M === Nothing ? 0 : 1, C_NULL, t.λ, C_NULL, 0, 0, C_NULL, 0, C_NULL))
    end
    p = Ref(bnuts_dual_averaging(da.δ, da.γ, da.κ, da.t₀, 0))
    check(e, ccall((:bnuts_warmup_stage, libbnuts), Int32,
        (Ptr{Cvoid}, Int32, Int32, Ref{bnuts_dual_averaging}, Float64, Ptr{Float64}, Int64, Int64, Ptr{Cvoid}, Int64, Ptr{Float64}),
        e, t.N, M === Nothing ? 0 : 1, p, t.λ, C_NULL, 0, 0, C_NULL, 0, C_NULL))
end

"""≙ threaded_mcmc(ℓ, N; δ, initialization, warmup_stages, algorithm, reporter, nchains), src/mcmc.jl:130-159.
Returns `chains::Array{Float64,3}` (D × N × nchains, the reference's layout) and
`tree_statistics::Matrix{TreeStatisticsNUTS}` (N × nchains)."""
function threaded_mcmc(ℓ, N; δ::Float64 = 0.8, initialization = (),
                       warmup_stages = default_warmup_stages(stepsize_adaptation = DualAveraging(δ = δ)),
                       algorithm = NUTS(), reporter = NoProgressReport(), nchains = 4096, dtype = 0, device = 0,
                       seed = UInt64(20261018), gradient_path = 0)
    D = dimension(ℓ)
    cfg = Ref(bnuts_config(sizeof(bnuts_config), dtype, nchains, D, algorithm.max_depth, device, algorithm.min_Δ, seed, 0, gradient_path))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:bnuts_create, libbnuts), Int32, (Ref{bnuts_config}, Ref{Ptr{Cvoid}}), cfg, h)
    rc == 0 || error("bnuts_create failed: " * unsafe_string(ccall((:bnuts_last_error, libbnuts), Cstring, (Ptr{Cvoid},), C_NULL)))
    e = h[]
    try
        attach!(e, ℓ)
        init = NamedTuple(initialization)
        if haskey(init, :κ)   # ≙ GaussianKineticEnergy: a D × nchains matrix of diagonals, or one dense D × D M⁻¹ (src/hamiltonian.jl:44)
            Mi = init.κ.M⁻¹
            if size(Mi) == (D, D) && nchains != D
                check(e, ccall((:bnuts_set_metric_dense, libbnuts), Int32, (Ptr{Cvoid}, Ptr{Float64}), e, Matrix{Float64}(Mi)))
            else
                check(e, ccall((:bnuts_set_metric_diag, libbnuts), Int32, (Ptr{Cvoid}, Ptr{Float64}), e, Mi))
            end
        end
        q = haskey(init, :q) ? init.q : nothing                      # D × nchains
        check(e, ccall((:bnuts_set_positions, libbnuts), Int32, (Ptr{Cvoid}, Ptr{Float64}), e, q === nothing ? C_NULL : q))
        if haskey(init, :ϵ)   # ≙ initialization = (ϵ = …,), src/warmup.jl:87-92,100: one value or one per chain
            ϵs = init.ϵ isa Number ? fill(Float64(init.ϵ), nchains) : Vector{Float64}(init.ϵ)
            check(e, ccall((:bnuts_set_stepsize, libbnuts), Int32, (Ptr{Cvoid}, Ptr{Float64}), e, ϵs))
            # as api.py: a given ϵ replaces the search stage (the reference would run it and overwrite the value)
            warmup_stages = filter(s -> !(s isa InitialStepsizeSearch), collect(warmup_stages))
        end
        foreach(s -> warmup!(e, s), warmup_stages)
        chains = Array{Float64,3}(undef, D, N, nchains)
        stats = Matrix{TreeStatisticsNUTS}(undef, N, nchains)
        GC.@preserve chains stats begin                              # ≙ mcmc!, src/warmup.jl:316-332
            check(e, ccall((:bnuts_sample, libbnuts), Int32,
                (Ptr{Cvoid}, Int32, Ptr{Float64}, Int64, Int64, Ptr{TreeStatisticsNUTS}, Int64, Ptr{Int32}),
                e, N, chains, D, D * N, stats, N, C_NULL))
        end
        return chains, stats
    finally
        ccall((:bnuts_destroy, libbnuts), Int32, (Ptr{Cvoid},), e)
    end
end

"""Row-sharded data (BASELINE config 5): call on every process after `attach!` with the 128-byte id that rank 0
got from `nccl_unique_id()` and the host broadcast.  No reference counterpart (`src/mcmc.jl:150-157` has threads only)."""
nccl_unique_id() = (id = zeros(UInt8, 128); ccall((:bnuts_nccl_unique_id, libbnuts), Int32, (Ptr{UInt8},), id) == 0 || error("NCCL not available"); id)
join_row_shards!(e, id::Vector{UInt8}, world::Integer, rank::Integer) =
    check(e, ccall((:bnuts_set_nccl, libbnuts), Int32, (Ptr{Cvoid}, Ptr{UInt8}, Int32, Int32), e, id, world, rank))
"Reference point of the tensor-core logistic path (β₀ near the mode, e.g. the result of FindLocalOptimum)."
set_reference!(e, β₀::Vector{Float64}) =
    check(e, ccall((:bnuts_logistic_set_reference, libbnuts), Int32, (Ptr{Cvoid}, Ptr{Float64}), e, β₀))

"""≙ WarmupState (z, κ, ϵ), src/warmup.jl:47-51, plus the position of the counter-based generator: everything a run needs
to continue.  `restore!` in a fresh engine (same model, dtype, chain ids) continues bit-identically (tests/test_checkpoint_resume.py)."""
struct WarmupState; q::Matrix{Float64}; M⁻¹::Matrix{Float64}; W::Matrix{Float64}; ϵ::Vector{Float64}; seed::UInt64; next_transition::UInt32; end
function warmup_state(e, D, nchains)
    q = Matrix{Float64}(undef, D, nchains); Mi = similar(q); W = similar(q); ϵ = Vector{Float64}(undef, nchains)
    seed = Ref{UInt64}(0); t = Ref{UInt32}(0)
    check(e, ccall((:bnuts_get_state, libbnuts), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), e, q, C_NULL, C_NULL))
    check(e, ccall((:bnuts_get_metric_diag, libbnuts), Int32, (Ptr{Cvoid}, Ptr{Float64}), e, Mi))
    check(e, ccall((:bnuts_get_metric_diag_w, libbnuts), Int32, (Ptr{Cvoid}, Ptr{Float64}), e, W))
    check(e, ccall((:bnuts_get_stepsize, libbnuts), Int32, (Ptr{Cvoid}, Ptr{Float64}), e, ϵ))
    check(e, ccall((:bnuts_get_rng, libbnuts), Int32, (Ptr{Cvoid}, Ref{UInt64}, Ref{UInt32}), e, seed, t))
    WarmupState(q, Mi, W, ϵ, seed[], t[])
end
function restore!(e, s::WarmupState)
    check(e, ccall((:bnuts_set_metric_diag_pair, libbnuts), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), e, s.M⁻¹, s.W))
    check(e, ccall((:bnuts_set_stepsize, libbnuts), Int32, (Ptr{Cvoid}, Ptr{Float64}), e, s.ϵ))
    check(e, ccall((:bnuts_seed, libbnuts), Int32, (Ptr{Cvoid}, UInt64, UInt32), e, s.seed, s.next_transition))
    check(e, ccall((:bnuts_set_positions, libbnuts), Int32, (Ptr{Cvoid}, Ptr{Float64}), e, s.q))
end

"≙ mcmc_with_warmup(ℓ, N; ...), src/mcmc.jl:109-128 (one chain)."
function mcmc_with_warmup(ℓ, N; kw...)
    chains, stats = threaded_mcmc(ℓ, N; nchains = 1, kw...)
    chains[:, :, 1], stats[:, 1]
end

end # module

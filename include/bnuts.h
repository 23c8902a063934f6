/* bnuts.h — C ABI of the B200 batched-chain NUTS engine (libbnuts.so).
 *
 * Drop-in boundary for the hot path of chriselrod/InplaceDHMC.jl.  The reference
 * has no FFI: its seam is Julia dispatch.  Each entry point below names the
 * reference call it replaces (paths relative to the reference checkout).  A Julia
 * host binds these with `ccall` (see INTEGRATION.md and julia/BNuts.jl).
 *
 * Conventions
 *  - every function returns 0 on success or a negative bnuts_status; the message
 *    is available from bnuts_last_error(); nothing throws or aborts across the ABI
 *  - host buffers are caller-owned, plain pointers + strides, always Float64 /
 *    Int32 (the reference is Float64-only: src/warmup.jl:108,115,119; src/mcmc.jl:117-122)
 *  - an engine is driven by one host thread at a time; engines are independent
 *  - one engine = the chains resident on one GPU; a multi-GPU job creates one
 *    engine per process/GPU with disjoint [chain_offset, chain_offset+n_chains)
 *  - all calls are synchronous on return
 */
#ifndef BNUTS_H
#define BNUTS_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bnuts_engine bnuts_engine;

typedef enum {
  BNUTS_OK = 0,
  BNUTS_ERR_INVALID_ARGUMENT = -1,
  BNUTS_ERR_NO_MODEL = -2,
  BNUTS_ERR_CUDA = -3,
  BNUTS_ERR_NONFINITE_START = -4,  /* ≙ DomainError, src/stepsize.jl:128,136,152-153 */
  BNUTS_ERR_STEPSIZE_SEARCH = -5,  /* ≙ error(), src/stepsize.jl:71,101 */
  BNUTS_ERR_STEPSIZE_COLLAPSE = -6,/* ≙ AssertionError eps < 1e-10, src/warmup.jl:291-296 */
  BNUTS_ERR_UNSUPPORTED = -7,
  BNUTS_ERR_INTERNAL = -8,
  BNUTS_ERR_OPTIMUM = -9           /* ≙ ThrowOptimizationError, src/warmup.jl:151,172 */
} bnuts_status;

enum { BNUTS_F64 = 0, BNUTS_F32 = 1 };                 /* engine arithmetic type */
enum { BNUTS_X_F64 = 0, BNUTS_X_F32 = 1, BNUTS_X_BF16 = 2 }; /* design-matrix storage */
enum { BNUTS_GRAD_AUTO = 0, BNUTS_GRAD_DETERMINISTIC = 1, BNUTS_GRAD_TENSOR = 2 };
enum { BNUTS_METRIC_NONE = 0, BNUTS_METRIC_DIAG = 1 }; /* ≙ TuningNUTS{Nothing|Diagonal}, src/warmup.jl:217-234
                                                         (TuningNUTS{Symmetric} also adapts a diagonal in the reference:
                                                         src/warmup.jl:309 -> src/hamiltonian.jl:117) */

/* ≙ NUTS(max_depth, min_Δ) src/NUTS.jl:204-220 + run geometry */
typedef struct {
  int32_t struct_size;   /* sizeof(bnuts_config), for ABI evolution */
  int32_t dtype;         /* BNUTS_F64 | BNUTS_F32 */
  int32_t n_chains;      /* chains resident in this engine */
  int32_t dim;           /* D */
  int32_t max_depth;     /* default 10 (src/NUTS.jl:214) */
  int32_t device;        /* CUDA device ordinal */
  double  min_delta;     /* default -1000.0 (src/NUTS.jl:214) */
  uint64_t seed;         /* Philox key */
  int32_t chain_offset;  /* global id of chain 0 (RNG is keyed by global id) */
  int32_t gradient_path; /* BNUTS_GRAD_* */
} bnuts_config;

/* ≙ TreeStatisticsNUTS, src/NUTS.jl:229-242 — same 32-byte layout, so Julia can
 * unsafe_wrap the buffer as Vector{TreeStatisticsNUTS}. */
typedef struct {
  double  pi;              /* log density of the selected draw (with its own momentum) */
  double  acceptance_rate; /* min(1, exp(log_sum_alpha)/steps), src/NUTS.jl:84 */
  int32_t term_left;       /* InvalidTree.left  (src/tree.jl:278-300) */
  int32_t term_right;      /* InvalidTree.right; (1,0) = REACHED_MAX_DEPTH */
  int32_t depth;
  int32_t steps;
} bnuts_tree_stats;

/* ≙ DualAveraging(δ,γ,κ,t₀), src/stepsize.jl:173-193 */
typedef struct { double delta, gamma, kappa; int32_t t0; int32_t _pad; } bnuts_dual_averaging;

/* ≙ InitialStepsizeSearch, src/stepsize.jl:16-38 */
typedef struct {
  double a_min, a_max, eps0, C;
  int32_t maxiter_crossing, maxiter_bisect;
} bnuts_stepsize_search;

typedef struct {
  int64_t leapfrogs;        /* gradient evaluations of finished transitions */
  int64_t transitions;      /* finished transitions */
  int64_t lockstep_steps;   /* kernel-level leapfrog rounds launched */
  int64_t kernel_launches;  /* launches of this library's kernels */
  int64_t divergences;
  int64_t gradient_rows;    /* staging rows evaluated by the batched gradient kernel (active chains summed over launches) */
} bnuts_counter_block;

int32_t bnuts_create(const bnuts_config* cfg, bnuts_engine** out);
int32_t bnuts_destroy(bnuts_engine* e);
/* e == NULL returns the message of the last failed bnuts_create on this thread */
const char* bnuts_last_error(const bnuts_engine* e);

/* Models.  The reference takes a user AbstractProbabilityModel and calls
 * logdensity_and_gradient!(∇ℓq, ℓ, q, sptr) (src/kinetic_energy.jl:73); a device
 * engine cannot call a Julia closure, so the benchmark targets are built in. */
int32_t bnuts_model_iid_normal(bnuts_engine* e);
int32_t bnuts_model_funnel(bnuts_engine* e);
int32_t bnuts_model_gaussian(bnuts_engine* e, const double* precision /* [D][D] */);
/* X is [N][D] row-major in x_dtype; y is [N] of 0/1.  row_blocks fixes the
 * summation order of the deterministic path (ignored by the tensor path). */
int32_t bnuts_model_logistic(bnuts_engine* e, const void* X, int32_t x_dtype, const double* y,
                             int64_t N, double prior_precision, int32_t row_blocks);

/* The same model on SYNTHETIC rows [row_offset, row_offset + N) of one conceptual design matrix; no reference
 * counterpart (the reference ships no models or data; SURVEY.md §8d defines the benchmark inputs, and for the
 * row-sharded configuration asks for rows generated where they are used).  Row i, column d is a pure function of
 * (data_seed, i, d): x[i][0] = 1, x[i][d] = bf16(N(0,1)) from Philox4x32-10 keyed by the seed with counter
 * (row, column quad), beta*[d] ~ N(0, 1/D), y[i] ~ Bernoulli(sigma(x[i]·beta*)) — so every sharding of the rows
 * sees the same matrix.  The tensor path generates its shard on the device (only the seed crosses PCIe); the
 * deterministic path generates it on the host and takes the upload route of bnuts_model_logistic. */
int32_t bnuts_model_logistic_synthetic(bnuts_engine* e, uint64_t data_seed, int64_t row_offset, int64_t N,
                                       double prior_precision, int32_t row_blocks);
/* Host copy of the same rows (no engine needed): X [nrows][D] as bf16 bits, y [nrows] of 0/1, beta_true [D];
 * any output may be NULL. */
int32_t bnuts_synth_logistic_rows(uint64_t data_seed, int64_t row_offset, int64_t nrows, int32_t D, uint16_t* X_bf16,
                                  double* y, double* beta_true);

/* Tensor-core logistic path only; no reference counterpart (the reference's first warmup stage,
 * FindLocalOptimum src/warmup.jl:152-186, is what produces such a point).  beta_ref [D] near the
 * posterior mode lets the kernel evaluate X·beta as X·beta_ref (stored) + X·(beta − beta_ref) with
 * two instead of three bf16 terms for the fp32 position.  The point is checked: if the gradient
 * there exceeds sqrt(N·D)/2 the call fails with BNUTS_ERR_INVALID_ARGUMENT and the exact
 * three-term path stays in force.  beta_ref == NULL returns to the three-term path.  For 128 < D <= 256 the kernel always
 * works about a reference point (zero until one is set; there is no three-term form to fall back to), so any finite
 * point is accepted there: optimise, set the reference, optimise again from the more accurate operand, set it again.
 * For tall problems (N >= 3000 D over the whole row group) the engine switches to the REMAINDER MODE (csrc/logistic_rm.cu): the
 * model is expanded about beta_ref row by row, the part that is linear / quadratic in beta − beta_ref (g0 − H0 (beta − beta_ref), D x D
 * constants formed in Float64 at this call) is exact arithmetic per chain, and only the small Taylor remainder goes through the tensor
 * cores; chains far from the reference fall back to closed forms inside the same kernel.  Gradient within 1e-6, log density within
 * 1e-4 of Float64 near the reference.  BNUTS_TC_RMODE=0/1/2 overrides (0: two bf16 terms of sigma(−eta); 1: one term of
 * sigma(−eta) − sigma(−eta_ref), the default for N >= 3.3e5 D where the remainder mode does not apply).
 * The call re-evaluates (log density, gradient) of every chain at its current position: the stored values came from the previous arithmetic. */
int32_t bnuts_logistic_set_reference(bnuts_engine* e, const double* beta_ref);

/* ≙ initialize_warmup_state(q = …), src/warmup.jl:100-129: sets q and evaluates
 * ℓ, ∇ℓ.  q == NULL draws U[-2,2]^D (src/warmup.jl:73) from Philox. */
int32_t bnuts_set_positions(bnuts_engine* e, const double* q /* [C][D] */);
int32_t bnuts_get_state(bnuts_engine* e, double* q, double* grad, double* logdensity);

/* ≙ GaussianKineticEnergy(M⁻¹) src/hamiltonian.jl:50-74; W = 1/sqrt(M⁻¹).
 * minv == NULL resets to the identity (src/warmup.jl:102). Per-chain metric. */
int32_t bnuts_set_metric_diag(bnuts_engine* e, const double* minv /* [C][D] */);
int32_t bnuts_get_metric_diag(bnuts_engine* e, double* minv /* [C][D] */);

/* ≙ GaussianKineticEnergy(M⁻¹::AbstractMatrix) — the dense constructor the reference keeps only as a comment
 * (src/hamiltonian.jl:44, W = cholesky(inv(M⁻¹)).L); its field types force Diagonal (:35-37).  One dense
 * symmetric positive definite M⁻¹ [D][D] shared by all chains; kinetic energy ½pᵀM⁻¹p, p♯ = M⁻¹p, drift
 * q + εM⁻¹p, momenta p = L⁻ᵀz with M⁻¹ = LLᵀ.  Implemented by whitening (see engine_core.h); Gaussian and iid
 * normal targets.  minv == NULL returns to the per-chain diagonal metric.  Existing positions are kept. */
/* ≙ GaussianKineticEnergy(M⁻¹, W), src/hamiltonian.jl:33-38, holds BOTH fields.  The metric update forms W from the
 * Float64 variance before M⁻¹ is rounded to the engine's type (src/hamiltonian.jl:153-162), so in an fp32 engine W is
 * not a function of the stored M⁻¹: a checkpoint carries the pair. */
int32_t bnuts_get_metric_diag_w(bnuts_engine* e, double* w /* [C][D] */);
int32_t bnuts_set_metric_diag_pair(bnuts_engine* e, const double* minv /* [C][D] */, const double* w /* [C][D] */);
int32_t bnuts_set_metric_dense(bnuts_engine* e, const double* minv /* [D][D] */);
int32_t bnuts_get_metric_dense(bnuts_engine* e, double* minv /* [D][D] */);

int32_t bnuts_set_stepsize(bnuts_engine* e, const double* eps /* [C] */);
int32_t bnuts_get_stepsize(bnuts_engine* e, double* eps /* [C] */);

/* Philox key and the transition counter the next transition will use. */
int32_t bnuts_seed(bnuts_engine* e, uint64_t seed, uint32_t next_transition);

/* ≙ WarmupState (src/warmup.jl:47-51) is (z, κ, ϵ); the engine's counterpart of the rng argument every reference entry
 * point takes is (seed, next transition index).  bnuts_get_state / bnuts_get_metric_diag / bnuts_get_stepsize /
 * bnuts_get_rng read everything a run needs to continue; the matching setters restore it in a fresh engine, and the
 * continued run is bit-identical to an uninterrupted one (checkpoint / resume; tests/test_checkpoint_resume.py). */
int32_t bnuts_get_rng(bnuts_engine* e, uint64_t* seed, uint32_t* next_transition);

/* ≙ sample_tree(rng, ...; p = …, directions = …) src/NUTS.jl:251-258 with a scripted rng: the next T transitions of
 * every chain use these directions, momenta and merge exponentials instead of Philox — the three random streams of a
 * transition.  dirs [T][C]; p [T][C][D]; exps [T][C][n_exps]: the k-th value of a row is the k-th randexp() the
 * reference's rand_bool_logprob (src/NUTS.jl:32-34) would CONSUME in that transition, i.e. merges in the order the
 * recursion performs them (post-order; the biased top-level merge of a doubling after its subtree), counting only
 * merges with logprob2 < 0; consumption beyond n_exps falls back to the engine's own stream.  Any of the three may
 * be NULL to keep that stream random. */
int32_t bnuts_inject(bnuts_engine* e, int32_t T, const uint32_t* dirs, const double* p, const double* exps, int32_t n_exps);

/* ≙ stack leapfrog, src/kinetic_energy.jl:164-195: nsteps leapfrogs of signed
 * step eps[c] from the engine's current (q, ∇ℓ) with momentum p_in; engine state
 * is not modified.  Outputs [C][D] / [C], any may be NULL. */
int32_t bnuts_leapfrog(bnuts_engine* e, const double* p_in, const double* eps, int32_t nsteps,
                       double* q_out, double* p_out, double* grad_out, double* logdensity_out);

/* ≙ warmup!(FindLocalOptimum(magnitude_penalty, iterations)) src/warmup.jl:137-186: every chain climbs
 * ℓ(q) − ½·magnitude_penalty·‖q‖² from its current position for at most `iterations` accepted steps and keeps
 * (q, ∇ℓ, ℓ) of the best point; a chain whose start (or result) is non-finite is re-randomised with a doubled
 * penalty, up to 100 times, then reported through bnuts_chain_status / BNUTS_ERR_OPTIMUM.  The reference's
 * inner solver is the un-vendored QuasiNewtonMethods.proptimize!; here: gradient ascent with a
 * Barzilai-Borwein step and Armijo backtracking, batched over chains.  Defaults 1e-4, 50 (src/warmup.jl:143,148). */
int32_t bnuts_find_local_optimum(bnuts_engine* e, double magnitude_penalty, int32_t iterations);

/* ≙ warmup!(InitialStepsizeSearch) src/warmup.jl:188-200 + src/stepsize.jl:51-126,150-164 */
int32_t bnuts_find_initial_stepsize(bnuts_engine* e, const bnuts_stepsize_search* params);

/* ≙ warmup!(TuningNUTS{M}) src/warmup.jl:269-314: N transitions with dual
 * averaging restarted from the current eps (src/stepsize.jl:208-212), then (metric_kind
 * == DIAG) the regularised variance update (src/hamiltonian.jl:117-189) and
 * eps <- final_ϵ (src/stepsize.jl:241).  lambda < 0 means the default 5/N.
 * da == NULL ≙ FixedStepsize (src/stepsize.jl:251-255; fixed_stepsize_warmup_stages, src/warmup.jl:383-389): ϵ is kept.
 * Optional outputs: chain_out/stats_out as bnuts_sample; eps_out [C][N]. */
int32_t bnuts_warmup_stage(bnuts_engine* e, int32_t N, int32_t metric_kind,
                           const bnuts_dual_averaging* da, double lambda,
                           double* chain_out, int64_t stride_draw, int64_t stride_chain,
                           bnuts_tree_stats* stats_out, int64_t stats_stride_chain,
                           double* eps_out);

/* ≙ mcmc! src/warmup.jl:316-332 with the output layout of threaded_mcmc
 * (src/mcmc.jl:140-152): draw n of chain c is written at
 * chain_out[c*stride_chain + n*stride_draw + d] (strides in elements), its
 * statistics at stats_out[c*stats_stride_chain + n].  selected_index (may be
 * NULL) receives the trajectory position of the selected draw, [C][N].
 * Chains whose bnuts_chain_status is non-zero (non-finite start, failed step size search, collapsed step size) do not
 * move: their output rows are zero-filled and the call still returns 0 unless the failure happened inside it
 * (BNUTS_ERR_STEPSIZE_COLLAPSE); callers mask rows with bnuts_chain_status. */
int32_t bnuts_sample(bnuts_engine* e, int32_t N,
                     double* chain_out, int64_t stride_draw, int64_t stride_chain,
                     bnuts_tree_stats* stats_out, int64_t stats_stride_chain,
                     int32_t* selected_index);

/* Row-sharded data (BASELINE config 5; the reference has no collective of any kind,
 * src/mcmc.jl:150-157).  Every engine of a group holds ALL chains (same seed, chain_offset, positions, step
 * sizes) and one shard of the rows of X (pass the local rows to bnuts_model_logistic).  After one of the calls
 * below, each leapfrog step sums the per-shard gradient / log-density partials over the group, so all
 * replicas see bit-identical gradients and make identical tree decisions; the prior term is added once,
 * after the sum.
 *   bnuts_set_nccl        CUDA engines, one per GPU: ncclAllReduce over NVLink on the engine's stream.
 *                         id: the 128 bytes of bnuts_nccl_unique_id() from rank 0, broadcast by the host.
 *   bnuts_set_allreduce   host-provided collective (tests, other transports): fn sums `count` elements
 *                         (dtype 0 = Float64, 1 = Float32) in place across the group and returns 0; buf is
 *                         device memory for the CUDA engine (its stream is idle during the call). */
typedef int32_t (*bnuts_allreduce_fn)(void* ctx, void* buf, int64_t count, int32_t dtype);
int32_t bnuts_set_allreduce(bnuts_engine* e, bnuts_allreduce_fn fn, void* ctx);
int32_t bnuts_nccl_unique_id(uint8_t* id /* [128] */);
int32_t bnuts_set_nccl(bnuts_engine* e, const uint8_t* id /* [128] */, int32_t world, int32_t rank);

/*   bnuts_p2p_export /   CUDA engines in different processes of one NVLink / NVSwitch node: the exchange is done by the
 *   bnuts_p2p_connect    library's own kernels over peer memory instead of NCCL.  The kernel that folds a shard's
 *                        partials also pushes the folded rows into every peer's receive slot (plain stores to
 *                        peer pointers, i.e. NVLink writes) and raises a flag there; the kernel that feeds the
 *                        chains waits for all flags and adds the slots in rank order (bit-identical on every rank).
 *                        export: allocates the receive buffer and returns its 64-byte cudaIpcMemHandle_t;
 *                        connect: the handles of all ranks in rank order (host-side all-gather), then as set_nccl. */
int32_t bnuts_p2p_export(bnuts_engine* e, uint8_t* handle /* [64] */);
int32_t bnuts_p2p_connect(bnuts_engine* e, const uint8_t* handles /* [world][64] */, int32_t world, int32_t rank);

int32_t bnuts_counters(bnuts_engine* e, bnuts_counter_block* out);
/* Measurement hook (no reference counterpart): when enabled, every launch of the
 * batched gradient kernel inside the run loop is bracketed by CUDA events on the
 * engine's stream.  Returns and resets the accumulated device time and launch count. */
int32_t bnuts_profile(bnuts_engine* e, int32_t enable, double* gradient_ms, int64_t* gradient_launches);
/* per-chain numerical status: 0 ok, else a bnuts_status */
int32_t bnuts_chain_status(bnuts_engine* e, int32_t* status /* [C] */);

#ifdef __cplusplus
}
#endif
#endif /* BNUTS_H */

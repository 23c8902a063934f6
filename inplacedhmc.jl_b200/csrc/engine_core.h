// engine_core.h — host-side engine behind the C ABI of include/bnuts.h.
//
// Written once over an execution policy X:
//   CudaExec  (engine_cuda.cu)        — the product: device memory, kernels, streams
//   HostExec  (tests/hostemu/*.cpp)   — a serial emulation used only by the CPU test
//                                       suite to exercise this file and the state
//                                       machine without a GPU
// X provides raw memory (alloc/free/h2d/d2h/zero), the per-chain launches
// (prepare / advance / metric / finish_da / totals) and the batched gradient
// evaluators of the GEMM-shaped targets.
//
// Call structure mirrors the reference drivers: bnuts_warmup_stage ≙
// warmup!(TuningNUTS) (src/warmup.jl:269-314), bnuts_sample ≙ mcmc!
// (src/warmup.jl:316-332), bnuts_find_initial_stepsize ≙ src/warmup.jl:188-200.
#pragma once
#include <algorithm>
#include <cmath>
#include <string>
#include <vector>
#include <cstring>
#include "../../include/bnuts.h"
#include "backend.h"

namespace bn {

static_assert(sizeof(TreeStats) == sizeof(bnuts_tree_stats), "stats layout");

template <class T> struct ModelCtx {
  int32_t kind = MODEL_NONE;
  // gaussian
  T* P = nullptr;            // [D][D] device
  // logistic
  T* X = nullptr;            // [N][D] device (deterministic path)
  T* y = nullptr;            // [N]
  uint16_t* Xb = nullptr;    // [Npad][Dt] bf16 device (tensor path)
  float* yf = nullptr;       // [Npad]
  int64_t N = 0, Npad = 0;
  int32_t row_blocks = 1;
  bool tensor = false;
  bool batched() const { return kind == MODEL_GAUSSIAN || kind == MODEL_LOGISTIC; }
};

// Shared dense metric M⁻¹ = L·Lᵀ (≙ GaussianKineticEnergy(M⁻¹::AbstractMatrix), the upstream constructor the
// reference keeps only as a comment, src/hamiltonian.jl:44).  HMC with a dense metric is, step for step, HMC
// with the unit metric on the whitened variable q̃ = L⁻¹q: p̃ = Lᵀp ~ N(0, I), ½p̃ᵀp̃ = ½pᵀM⁻¹p,
// q̃′ = q̃ + ε p̃ₘ ⇔ q′ = q + ε M⁻¹pₘ, ∇ℓ̃ = Lᵀ∇ℓ, and the turn test ρ̃·p̃ = ρᵀM⁻¹p.  So the per-chain state
// machine runs unchanged in whitened coordinates, the metric lives in the model (for the Gaussian target
// P̃ = LᵀPL: one GEMM operand instead of three metric applications per leapfrog) and positions, momenta and
// gradients are mapped at the C-ABI boundary by batched matrix products on the device.
struct DenseMetric {
  bool on = false;
  std::vector<double> Minv, L, Linv;                 // host, [D][D] row-major
  double* dL = nullptr; double* dLt = nullptr; double* dLinv = nullptr; double* dLinvt = nullptr;   // device
};

// host form of the synthetic-row definition (bnuts_math.h, synth_*): rows [row0, row0 + N) as bf16 bits + labels
inline void synth_rows_host(uint64_t seed, int64_t row0, int64_t N, int32_t D, uint16_t* X, double* y, double* beta_out) {
  std::vector<double> beta(static_cast<size_t>(D));
  for (int32_t d = 0; d < D; ++d) beta[size_t(d)] = synth_beta(seed, uint32_t(d), D);
  if (beta_out) for (int32_t d = 0; d < D; ++d) beta_out[d] = beta[size_t(d)];
#if defined(_OPENMP)
#pragma omp parallel for schedule(static)
#endif
  for (int64_t i = 0; i < N; ++i) {
    const uint64_t row = uint64_t(row0 + i);
    if (X)
      for (int32_t q = 0; 4 * q < D; ++q) {
        float x[4];
        synth_x4(seed, row, uint32_t(q), x);
        for (int e = 0; e < 4 && 4 * q + e < D; ++e) X[size_t(i) * D + 4 * q + e] = uint16_t(f2u(x[e]) >> 16);
      }
    if (y) y[i] = synth_label(seed, row, D, beta.data());
  }
}

template <class T, class X> struct EngineCore {
  bnuts_config cfg{};
  X x;
  EngineMem<T> M{};
  RunParams<T> rp{};
  ModelCtx<T> model;
  uint32_t next_t = 0;
  std::string err;
  bnuts_counter_block counters{};
  // per-call device output buffers (grown on demand)
  size_t cap_draws = 0, cap_stats = 0;
  TreeStats* d_stats = nullptr; int32_t* d_sel = nullptr; double* d_eps_hist = nullptr;
  uint32_t* d_inj_dirs = nullptr; double* d_inj_p = nullptr; double* d_inj_exps = nullptr;
  double* d_tmp_cd = nullptr;   // [C][D] scratch (positions / momenta in)
  double* d_tmp_c = nullptr;    // [C]
  double* d_bare_p = nullptr; double* d_bare_out = nullptr;   // bnuts_leapfrog: momenta in [C][D], (q, p, ∇ℓ, ℓ) out [3][C][D] + [C]
  std::vector<double> h_draws;  // staging when caller strides are not compact
  DenseMetric dm;
  std::vector<double> h_P;      // Gaussian target precision as given by the caller (user coordinates)
  int32_t user_model_kind = MODEL_NONE;
  double* d_xf_in = nullptr; double* d_xf_out = nullptr; size_t cap_xf = 0;   // transform scratch
  // row-sharded data: per-leapfrog sum of the gradient partials over the group
  bool reduce_on = false;
  bool p2p_on = false; uint64_t p2p_seq = 0;   // peer-memory exchange instead of a collective call
  bnuts_allreduce_fn red_fn = nullptr; void* red_ctx = nullptr;
  T* red_g = nullptr;           // [C][Dp] folded gradient partials, summed across the group in place
  double* red_l = nullptr;      // [C]     folded log-density partials (Float64)
  T* wide_q = nullptr; uint16_t* wide_bh = nullptr; uint16_t* wide_bm = nullptr; uint16_t* wide_bl = nullptr;
  int32_t* d_active = nullptr;  // [C] request flags (deterministic row assignment)

  int32_t fail(int32_t code, const std::string& m) { err = m; return code; }

  int32_t init(const bnuts_config& c) {
    cfg = c;
    if (c.max_depth > MAX_LEVELS) return fail(BNUTS_ERR_UNSUPPORTED, "max_depth > 32 not supported (one UInt32 of directions, src/tree.jl:132)");
    int32_t rc = x.init(c.device, err);
    if (rc) return rc;
    M.C = c.n_chains; M.D = c.dim; M.Dp = (c.dim + 31) / 32 * 32;
    M.S = c.max_depth + 4; M.L = c.max_depth;
    M.model_kind = MODEL_NONE;
    const size_t CD = size_t(M.C) * M.Dp;
    M.zs = x.template alloc<T>(CD * M.S * 3);
    M.zlq = x.template alloc<double>(size_t(M.C) * M.S);
    M.st_rho = x.template alloc<T>(CD * M.L);
    M.st_psf = x.template alloc<T>(CD * M.L);
    M.m_rho = x.template alloc<T>(CD); M.m_psm = x.template alloc<T>(CD); M.m_psp = x.template alloc<T>(CD);
    M.ps_cur = x.template alloc<T>(CD); M.Minv = x.template alloc<T>(CD); M.W = x.template alloc<T>(CD);
    M.cs = x.template alloc<ChainState<T>>(M.C);
    d_tmp_cd = x.template alloc<double>(size_t(M.C) * M.D * 4 + M.C);
    d_tmp_c = x.template alloc<double>(M.C);
    {
      const void* need[] = {M.zs, M.zlq, M.st_rho, M.st_psf, M.m_rho, M.m_psm, M.m_psp, M.ps_cur, M.Minv, M.W, M.cs, d_tmp_cd, d_tmp_c};
      for (const void* p : need) if (!p) return fail(BNUTS_ERR_CUDA, "device allocation failed");
      if (x.check(err)) return BNUTS_ERR_CUDA;
    }
    x.zero(M.zs, CD * M.S * 3 * sizeof(T));
    x.zero(M.zlq, size_t(M.C) * M.S * sizeof(double));
    x.zero(M.cs, size_t(M.C) * sizeof(ChainState<T>));
    rp.max_depth = c.max_depth; rp.min_delta = c.min_delta; rp.seed = c.seed;
    rp.chain_offset = c.chain_offset; rp.n_chains = c.n_chains; rp.n_slots = M.S;
    std::vector<double> ones(size_t(M.C) * M.D, 1.0);
    set_metric(nullptr);
    std::vector<double> e1(M.C, 1.0);
    set_stepsize(e1.data());
    return x.check(err);
  }
  void destroy() {
    x.sync();
    free_reduce();
    free_dense();
    if (d_xf_in) { x.free(d_xf_in); x.free(d_xf_out); d_xf_in = d_xf_out = nullptr; }
    void* ptrs[] = {M.zs, M.zlq, M.st_rho, M.st_psf, M.m_rho, M.m_psm, M.m_psp, M.ps_cur, M.Minv, M.W, M.cs,
                    M.stage_q, M.stage_g, M.stage_l, M.stage_ld, M.stage_bh, M.stage_bm, M.stage_bl, M.stage_row, M.draws, d_stats, d_sel, d_eps_hist,
                    d_inj_dirs, d_inj_p, d_inj_exps, d_tmp_cd, d_tmp_c, d_bare_p, d_bare_out, model.P, model.X, model.y, model.Xb, model.yf};
    for (void* p : ptrs) if (p) x.free(p);
    x.shutdown();
  }

  // ---------------------------------------------------------------- models
  void free_reduce() {
    void** ps[] = {(void**)&red_g, (void**)&red_l, (void**)&wide_q, (void**)&wide_bh, (void**)&wide_bm, (void**)&wide_bl,
                   (void**)&d_active};
    for (void** p : ps) if (*p) { x.free(*p); *p = nullptr; }
    reduce_on = false; p2p_on = false;
  }
  int32_t p2p_export(uint8_t* handle) {
    if (model.kind != MODEL_LOGISTIC) return fail(BNUTS_ERR_NO_MODEL, "row sharding needs the logistic model (set it first)");
    return x.p2p_export(size_t(M.C) * M.Dp * sizeof(T), size_t(M.C) * sizeof(double), handle, err);
  }
  int32_t p2p_connect(const uint8_t* handles, int32_t world, int32_t rank) {
    int32_t rc = x.p2p_connect(handles, world, rank, err);
    if (rc) return rc;
    rc = enable_reduce(nullptr, nullptr);
    if (rc) return rc;
    p2p_on = true; p2p_seq = 0; red_world = world;
    return 0;
  }
  // ≙ no reference counterpart (the reference has no collective).  After this call the engine
  // treats its design matrix as one shard of the rows: every lockstep step the folded partials
  // [rows x Dp] gradient + [rows] log density are summed over the group before the chains consume them.
  // All engines of the group must hold all chains (same seed, chain_offset, positions).
  int32_t enable_reduce(bnuts_allreduce_fn fn, void* ctx) {
    if (model.kind != MODEL_LOGISTIC) return fail(BNUTS_ERR_NO_MODEL, "row sharding needs the logistic model (set it first)");
    free_reduce();
    const size_t CD = size_t(M.C) * M.Dp;
    red_g = x.template alloc<T>(CD); red_l = x.template alloc<double>(M.C);
    wide_q = x.template alloc<T>(CD); d_active = x.template alloc<int32_t>(M.C);
    x.zero(wide_q, CD * sizeof(T)); x.zero(d_active, size_t(M.C) * sizeof(int32_t));
    if (M.stage_bh) {
      const size_t nbt = size_t(M.C) * M.Dt;
      wide_bh = x.template alloc<uint16_t>(nbt); wide_bm = x.template alloc<uint16_t>(nbt); wide_bl = x.template alloc<uint16_t>(nbt);
      x.zero(wide_bh, nbt * 2); x.zero(wide_bm, nbt * 2); x.zero(wide_bl, nbt * 2);
    }
    red_fn = fn; red_ctx = ctx;
    reduce_on = true;
    return x.check(err);
  }
  // the view of the staging buffers the chains see in row-sharded mode: they write requests to the wide
  // buffers (row = chain) and read the reduced result (one partial block)
  EngineMem<T> reduce_view(int64_t rows) const {
    EngineMem<T> V = M;
    V.stage_q = wide_q; V.stage_bh = wide_bh; V.stage_bm = wide_bm; V.stage_bl = wide_bl;
    V.stage_active = d_active;
    V.stage_g = red_g; V.stage_ld = red_l; V.stage_nb = 1; V.stage_rows = (int32_t)rows; V.lin_w = nullptr;
    if (!M.lin_H) V.grad0 = nullptr;   // remainder mode: the consumer adds g0 - H0 (q - beta_ref) (constants summed over the group at set-up)
    return V;
  }

  void free_model() {
    free_reduce();
    void** ps[] = {(void**)&model.P, (void**)&model.X, (void**)&model.y, (void**)&model.Xb, (void**)&model.yf,
                   (void**)&M.stage_q, (void**)&M.stage_g, (void**)&M.stage_l, (void**)&M.stage_ld, (void**)&M.stage_bh,
                   (void**)&M.stage_bm, (void**)&M.stage_bl, (void**)&M.stage_row};
    for (void** p : ps) if (*p) { x.free(*p); *p = nullptr; }
    model = ModelCtx<T>();
    user_model_kind = MODEL_NONE; h_P.clear();
    M.stage_nb = 0;
    M.beta_ref = nullptr; M.lin_w = nullptr; M.grad0 = nullptr; M.lin_H = nullptr; M.ell0 = 0.0;
  }
  int32_t model_simple(int kind) {
    free_model();
    model.kind = kind; M.model_kind = kind;
    user_model_kind = kind; h_P.clear();
    if (dm.on) return apply_dense_to_model();
    return 0;
  }
  int32_t upload_gaussian(const std::vector<double>& Pd) {
    const size_t n = size_t(M.D) * M.D;
    std::vector<T> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = T(Pd[i]);
    if (!model.P) model.P = x.template alloc<T>(n);
    x.h2d(model.P, h.data(), n * sizeof(T));
    if (!M.stage_q) alloc_stage(1);
    model.kind = MODEL_GAUSSIAN; M.model_kind = MODEL_GAUSSIAN;
    // tensor-core gradient (fp32 engine): on request, or by default for large D where the GEMM dominates
    const bool want_tensor = cfg.gradient_path == BNUTS_GRAD_TENSOR ||
                             (cfg.gradient_path == BNUTS_GRAD_AUTO && sizeof(T) == 4 && X::has_tensor_path && M.D >= 512);
    model.tensor = false;
    if (want_tensor) {
      if (sizeof(T) != 4 || !X::has_tensor_path)
        return fail(BNUTS_ERR_UNSUPPORTED, "tensor gradient path needs dtype F32 on the CUDA engine");
      int32_t rc = x.gauss_tensor_setup(*this, Pd, err);
      if (rc) return rc;
      model.tensor = true;
    }
    return x.check(err);
  }
  // whitened model: P̃ = Lᵀ P L (P = I for the iid normal target)
  int32_t apply_dense_to_model() {
    const int D = M.D;
    if (user_model_kind == MODEL_NONE) return 0;
    if (user_model_kind != MODEL_GAUSSIAN && user_model_kind != MODEL_IID_NORMAL)
      return fail(BNUTS_ERR_UNSUPPORTED, "dense metric is implemented for the Gaussian and iid normal targets");
    std::vector<double> T1(size_t(D) * D, 0.0), Pt(size_t(D) * D, 0.0);
    const double* L = dm.L.data();
    if (user_model_kind == MODEL_GAUSSIAN) {   // T1 = P L
      for (int i = 0; i < D; ++i)
        for (int k = 0; k < D; ++k) {
          const double pik = h_P[size_t(i) * D + k];
          for (int j = 0; j <= k; ++j) T1[size_t(i) * D + j] += pik * L[size_t(k) * D + j];
        }
    } else {
      T1 = dm.L;
    }
    for (int k = 0; k < D; ++k)               // P̃ = Lᵀ T1
      for (int i = 0; i <= k; ++i) {
        const double lki = L[size_t(k) * D + i];
        for (int j = 0; j < D; ++j) Pt[size_t(i) * D + j] += lki * T1[size_t(k) * D + j];
      }
    for (int i = 0; i < D; ++i)               // symmetrise away rounding asymmetry
      for (int j = 0; j < i; ++j) {
        const double s = 0.5 * (Pt[size_t(i) * D + j] + Pt[size_t(j) * D + i]);
        Pt[size_t(i) * D + j] = s; Pt[size_t(j) * D + i] = s;
      }
    if (model.kind != MODEL_GAUSSIAN) free_model_keep_user();
    return upload_gaussian(Pt);
  }
  void free_model_keep_user() {
    const int32_t k = user_model_kind; std::vector<double> p = h_P;
    free_model();
    user_model_kind = k; h_P = p;
  }
  // partial_rows = rows of the partial-output buffers (>= nb * C)
  void alloc_stage(int nb, size_t partial_rows = 0) {
    const size_t CD = size_t(M.C) * M.Dp;
    if (partial_rows < size_t(nb) * M.C) partial_rows = size_t(nb) * M.C;
    M.stage_nb = nb;
    M.stage_rows = M.C;
    M.stage_q = x.template alloc<T>(CD);
    M.stage_g = x.template alloc<T>(partial_rows * M.Dp);
    M.stage_l = x.template alloc<T>(partial_rows);
    M.stage_row = x.template alloc<int32_t>(M.C);
    M.stage_count = x.counter();
    x.zero(M.stage_q, CD * sizeof(T));
    x.zero(M.stage_g, partial_rows * M.Dp * sizeof(T));
    x.zero(M.stage_l, partial_rows * sizeof(T));
    x.zero(M.stage_row, size_t(M.C) * sizeof(int32_t));
  }
  int32_t model_gaussian(const double* P) {
    if (!P) return fail(BNUTS_ERR_INVALID_ARGUMENT, "precision is NULL");
    free_model();
    const size_t n = size_t(M.D) * M.D;
    h_P.assign(P, P + n);
    user_model_kind = MODEL_GAUSSIAN;
    if (dm.on) return apply_dense_to_model();
    return upload_gaussian(h_P);
  }
  int32_t model_logistic(const void* Xh, int32_t xd, const double* y, int64_t N, double tau, int32_t rb) {
    if (!Xh || !y || N <= 0 || rb <= 0) return fail(BNUTS_ERR_INVALID_ARGUMENT, "bad logistic arguments");
    if (dm.on) return fail(BNUTS_ERR_UNSUPPORTED, "dense metric is implemented for the Gaussian and iid normal targets");
    if (xd != BNUTS_X_F64 && xd != BNUTS_X_F32 && xd != BNUTS_X_BF16) return fail(BNUTS_ERR_INVALID_ARGUMENT, "bad x_dtype");
    free_model();
    const bool want_tensor = cfg.gradient_path == BNUTS_GRAD_TENSOR ||
                             (cfg.gradient_path == BNUTS_GRAD_AUTO && sizeof(T) == 4 && X::has_tensor_path);
    if (want_tensor) {
      if (sizeof(T) != 4 || !X::has_tensor_path)
        return fail(BNUTS_ERR_UNSUPPORTED, "tensor gradient path needs dtype F32 on the CUDA engine");
      int32_t rc = x.logistic_tensor_setup(*this, Xh, xd, y, N, err);
      if (rc) return rc;
      model.tensor = true;
    } else {
      const size_t n = size_t(N) * M.D;
      std::vector<T> h(n);
      std::vector<T> hy(static_cast<size_t>(N));
      for (size_t i = 0; i < n; ++i) {
        double v;
        if (xd == BNUTS_X_F64) v = static_cast<const double*>(Xh)[i];
        else if (xd == BNUTS_X_F32) v = double(static_cast<const float*>(Xh)[i]);
        else v = double(bf16_val(static_cast<const uint16_t*>(Xh)[i]));
        h[i] = T(v);
      }
      for (int64_t i = 0; i < N; ++i) hy[size_t(i)] = T(y[i]);
      model.X = x.template alloc<T>(n); model.y = x.template alloc<T>(size_t(N));
      if (!model.X || !model.y) return fail(BNUTS_ERR_CUDA, "device allocation failed");
      x.h2d(model.X, h.data(), n * sizeof(T));
      x.h2d(model.y, hy.data(), size_t(N) * sizeof(T));
      model.row_blocks = rb;
      alloc_stage(rb);
    }
    model.N = N;
    M.tau = T(tau);
    user_model_kind = MODEL_LOGISTIC;
    model.kind = MODEL_LOGISTIC; M.model_kind = MODEL_LOGISTIC;
    return x.check(err);
  }

  // ≙ SURVEY.md §8d (config c5): the shard's rows [row0, row0 + N) of the synthetic design matrix, generated from
  // Philox keyed by (data seed, global row index) — on the device for the tensor path, by the same definition on the
  // host (synth_rows_host) for the deterministic path, which then takes the ordinary upload route
  int32_t model_logistic_synth(uint64_t seed, int64_t row0, int64_t N, double tau, int32_t rb) {
    if (N <= 0 || row0 < 0 || rb <= 0) return fail(BNUTS_ERR_INVALID_ARGUMENT, "bad synthetic logistic arguments");
    if (dm.on) return fail(BNUTS_ERR_UNSUPPORTED, "dense metric is implemented for the Gaussian and iid normal targets");
    const bool want_tensor = cfg.gradient_path == BNUTS_GRAD_TENSOR ||
                             (cfg.gradient_path == BNUTS_GRAD_AUTO && sizeof(T) == 4 && X::has_tensor_path);
    if (!want_tensor) {
      std::vector<uint16_t> xb(size_t(N) * M.D);
      std::vector<double> y(static_cast<size_t>(N));
      synth_rows_host(seed, row0, N, M.D, xb.data(), y.data(), nullptr);
      return model_logistic(xb.data(), BNUTS_X_BF16, y.data(), N, tau, rb);
    }
    if (sizeof(T) != 4 || !X::has_tensor_path)
      return fail(BNUTS_ERR_UNSUPPORTED, "tensor gradient path needs dtype F32 on the CUDA engine");
    free_model();
    int32_t rc = x.logistic_tensor_setup_synth(*this, seed, row0, N, err);
    if (rc) return rc;
    model.tensor = true;
    model.N = N;
    M.tau = T(tau);
    user_model_kind = MODEL_LOGISTIC;
    model.kind = MODEL_LOGISTIC; M.model_kind = MODEL_LOGISTIC;
    return x.check(err);
  }

  // gradient of staged row 0 after a one-row gradient launch with `nb` partial blocks, folded on the host;
  // in row-sharded mode the partials are first summed over the group (every rank must make the same call)
  int32_t folded_row0(int nb, std::vector<double>& g) {
    g.assign(size_t(M.Dp), 0.0);
    if (reduce_on) {
      M.stage_nb = nb; M.stage_rows = 1;
      if (p2p_on) {
        p2p_seq += 1;
        x.fold_push(M, 1, p2p_seq);
        x.wait_sum(M, 1, p2p_seq, red_g, red_l);
      } else {
        x.fold_partials(M, 1, red_g, red_l);
        int32_t rc = x.allreduce(red_g, M.Dp, sizeof(T) == 4, red_l, 1, red_fn, red_ctx, err);
        if (rc) return rc;
      }
      std::vector<T> r(size_t(M.Dp));
      x.d2h(r.data(), red_g, r.size() * sizeof(T));
      for (int d = 0; d < M.D; ++d) g[d] = double(r[d]);
    } else {
      std::vector<T> r(size_t(nb) * M.Dp);
      x.d2h(r.data(), M.stage_g, r.size() * sizeof(T));
      for (int d = 0; d < M.D; ++d)
        for (int s = 0; s < nb; ++s) g[d] += double(r[size_t(s) * M.Dp + d]);
    }
    return x.check(err);
  }
  int reduce_world() const { return reduce_on ? red_world : 1; }
  int red_world = 1;

  // reference point of the tensor-core logistic path (numerical device, see include/bnuts.h)
  int32_t logistic_set_reference(const double* beta_ref) {
    if (model.kind != MODEL_LOGISTIC || !model.tensor)
      return fail(BNUTS_ERR_UNSUPPORTED, "reference point applies to the tensor-core logistic path only");
    const int32_t rc = x.logistic_reference(*this, beta_ref, err);
    if (rc) return rc;
    // The reference changes the arithmetic of the model (operand split, residual mode, remainder mode): the (ℓ, ∇ℓ) the
    // chains hold were computed by the previous arithmetic.  At N = 1e6 the two differ by ~3e-2 in ℓ; at N = 1e7 by several
    // units, which no step-size search or Armijo test that starts from the stored ℓ survives.  Re-evaluate where the chains are.
    x.gather_state(M, d_tmp_cd);
    M.pos_in = d_tmp_cd;
    PrepareArgs a{}; a.mode = MODE_EVAL;
    if (reduce_on) x.prepare(reduce_view(0), rp, a); else x.prepare(M, rp, a);
    const int32_t rc2 = run(true);
    M.pos_in = nullptr;
    return rc2;
  }

  // ---------------------------------------------------------------- state setters
  // ≙ GaussianKineticEnergy(M⁻¹): W = 1/sqrt(M⁻¹) (src/hamiltonian.jl:53-55) needs positive, finite entries
  int32_t check_metric_values(const double* v, const char* what) {
    const size_t n = size_t(M.C) * M.D;
    for (size_t i = 0; i < n; ++i)
      if (!(v[i] > 0.0) || !(v[i] < 1.7e308)) return fail(BNUTS_ERR_INVALID_ARGUMENT, std::string(what) + " must be positive and finite");
    return 0;
  }
  int32_t set_metric(const double* minv) {
    if (minv) { const int32_t rc = check_metric_values(minv, "M^-1"); if (rc) return rc; }
    std::vector<T> hm(size_t(M.C) * M.Dp, T(1)), hw(size_t(M.C) * M.Dp, T(1));
    for (int c = 0; c < M.C; ++c)
      for (int d = 0; d < M.D; ++d) {
        const double m = minv ? minv[size_t(c) * M.D + d] : 1.0;
        hm[size_t(c) * M.Dp + d] = T(m);
        hw[size_t(c) * M.Dp + d] = T(1.0 / sqrt_(m));  // ≙ src/hamiltonian.jl:53-55
      }
    x.h2d(M.Minv, hm.data(), hm.size() * sizeof(T));
    x.h2d(M.W, hw.data(), hw.size() * sizeof(T));
    return x.check(err);
  }
  // ≙ GaussianKineticEnergy(M⁻¹, W), src/hamiltonian.jl:33-38: the reference type holds BOTH fields.  The metric update
  // forms W from the Float64 variance before M⁻¹ is rounded to T (k_metric / metric_update), so in an fp32 engine W is not
  // a function of the stored M⁻¹: a checkpoint has to carry it (bnuts_get_metric_diag_w / bnuts_set_metric_diag_pair).
  int32_t set_metric_pair(const double* minv, const double* w) {
    if (!minv || !w) return fail(BNUTS_ERR_INVALID_ARGUMENT, "minv and w required");
    { int32_t rc = check_metric_values(minv, "M^-1"); if (!rc) rc = check_metric_values(w, "W"); if (rc) return rc; }
    std::vector<T> hm(size_t(M.C) * M.Dp, T(1)), hw(size_t(M.C) * M.Dp, T(1));
    for (int c = 0; c < M.C; ++c)
      for (int d = 0; d < M.D; ++d) {
        hm[size_t(c) * M.Dp + d] = T(minv[size_t(c) * M.D + d]);
        hw[size_t(c) * M.Dp + d] = T(w[size_t(c) * M.D + d]);
      }
    x.h2d(M.Minv, hm.data(), hm.size() * sizeof(T));
    x.h2d(M.W, hw.data(), hw.size() * sizeof(T));
    return x.check(err);
  }
  int32_t get_metric_w(double* out) {
    std::vector<T> hw(size_t(M.C) * M.Dp);
    x.d2h(hw.data(), M.W, hw.size() * sizeof(T));
    for (int c = 0; c < M.C; ++c)
      for (int d = 0; d < M.D; ++d) out[size_t(c) * M.D + d] = double(hw[size_t(c) * M.Dp + d]);
    return x.check(err);
  }
  // ---------------------------------------------------------------- dense metric (see DenseMetric above)
  void free_dense() {
    void** ps[] = {(void**)&dm.dL, (void**)&dm.dLt, (void**)&dm.dLinv, (void**)&dm.dLinvt};
    for (void** p : ps) if (*p) { x.free(*p); *p = nullptr; }
    dm.on = false; dm.Minv.clear(); dm.L.clear(); dm.Linv.clear();
  }
  void xf_reserve(size_t n) {
    if (n <= cap_xf) return;
    if (d_xf_in) { x.free(d_xf_in); x.free(d_xf_out); }
    d_xf_in = x.template alloc<double>(n); d_xf_out = x.template alloc<double>(n);
    cap_xf = n;
  }
  // rows [R][D] (device, Float64) times a D x D matrix, in place through the scratch buffer
  void xf_device(double* rows, int64_t R, const double* mat) {
    xf_reserve(size_t(R) * M.D);
    x.rows_times_matrix(rows, d_xf_out, R, M.D, mat);
    x.d2d(rows, d_xf_out, size_t(R) * M.D * sizeof(double));
  }
  // host rows -> device scratch, transformed; returns the device pointer of the result
  double* xf_host(const double* rows, int64_t R, const double* mat) {
    xf_reserve(size_t(R) * M.D);
    x.h2d(d_xf_in, rows, size_t(R) * M.D * sizeof(double));
    x.rows_times_matrix(d_xf_in, d_xf_out, R, M.D, mat);
    return d_xf_out;
  }
  int32_t set_metric_dense(const double* minv) {
    const int D = M.D;
    // current positions in user coordinates, to be re-expressed under the new metric
    std::vector<double> q_user;
    const bool have_pos = model.kind != MODEL_NONE && !x.any_status(M, ST_NONFINITE_START);
    if (have_pos) { q_user.resize(size_t(M.C) * D); int32_t rc = get_state(q_user.data(), nullptr, nullptr); if (rc) return rc; }
    if (!minv) {
      free_dense();
      if (user_model_kind == MODEL_GAUSSIAN) { int32_t rc = upload_gaussian(h_P); if (rc) return rc; }
      else if (user_model_kind == MODEL_IID_NORMAL) { const int32_t k = user_model_kind; free_model(); model.kind = k; M.model_kind = k; user_model_kind = k; }
    } else {
      if (user_model_kind != MODEL_NONE && user_model_kind != MODEL_GAUSSIAN && user_model_kind != MODEL_IID_NORMAL)
        return fail(BNUTS_ERR_UNSUPPORTED, "dense metric is implemented for the Gaussian and iid normal targets");
      const size_t n = size_t(D) * D;
      std::vector<double> A(minv, minv + n), L(n, 0.0), Li(n, 0.0);
      for (int i = 0; i < D; ++i)
        for (int j = 0; j < i; ++j)
          if (std::fabs(A[size_t(i) * D + j] - A[size_t(j) * D + i]) > 1e-10 * (std::fabs(A[size_t(i) * D + i]) + std::fabs(A[size_t(j) * D + j])))
            return fail(BNUTS_ERR_INVALID_ARGUMENT, "dense metric must be symmetric");
      for (int j = 0; j < D; ++j) {            // Cholesky M⁻¹ = L Lᵀ
        double s = A[size_t(j) * D + j];
        for (int k = 0; k < j; ++k) s -= L[size_t(j) * D + k] * L[size_t(j) * D + k];
        if (!(s > 0.0)) return fail(BNUTS_ERR_INVALID_ARGUMENT, "dense metric must be positive definite");
        const double ljj = std::sqrt(s);
        L[size_t(j) * D + j] = ljj;
        for (int i = j + 1; i < D; ++i) {
          double t = A[size_t(i) * D + j];
          for (int k = 0; k < j; ++k) t -= L[size_t(i) * D + k] * L[size_t(j) * D + k];
          L[size_t(i) * D + j] = t / ljj;
        }
      }
      for (int c = 0; c < D; ++c) {            // L⁻¹ by forward substitution, column by column
        Li[size_t(c) * D + c] = 1.0 / L[size_t(c) * D + c];
        for (int i = c + 1; i < D; ++i) {
          double t = 0.0;
          for (int k = c; k < i; ++k) t -= L[size_t(i) * D + k] * Li[size_t(k) * D + c];
          Li[size_t(i) * D + c] = t / L[size_t(i) * D + i];
        }
      }
      free_dense();
      dm.Minv = A; dm.L = L; dm.Linv = Li;
      std::vector<double> Lt(n), Lit(n);
      for (int i = 0; i < D; ++i) for (int j = 0; j < D; ++j) { Lt[size_t(j) * D + i] = L[size_t(i) * D + j]; Lit[size_t(j) * D + i] = Li[size_t(i) * D + j]; }
      dm.dL = x.template alloc<double>(n); dm.dLt = x.template alloc<double>(n);
      dm.dLinv = x.template alloc<double>(n); dm.dLinvt = x.template alloc<double>(n);
      x.h2d(dm.dL, L.data(), n * 8); x.h2d(dm.dLt, Lt.data(), n * 8); x.h2d(dm.dLinv, Li.data(), n * 8); x.h2d(dm.dLinvt, Lit.data(), n * 8);
      dm.on = true;
      int32_t rc = apply_dense_to_model();
      if (rc) return rc;
    }
    int32_t rc = set_metric(nullptr);          // the per-chain diagonal lives in the (new) sampling coordinates
    if (rc) return rc;
    if (have_pos && model.kind != MODEL_NONE) return set_positions(q_user.data());
    return x.check(err);
  }
  int32_t get_metric_dense(double* out) {
    const size_t n = size_t(M.D) * M.D;
    if (dm.on) { std::memcpy(out, dm.Minv.data(), n * 8); return 0; }
    std::fill(out, out + n, 0.0);
    for (int d = 0; d < M.D; ++d) out[size_t(d) * M.D + d] = 1.0;
    return 0;
  }

  int32_t get_metric(double* out) {
    std::vector<T> hm(size_t(M.C) * M.Dp);
    x.d2h(hm.data(), M.Minv, hm.size() * sizeof(T));
    for (int c = 0; c < M.C; ++c)
      for (int d = 0; d < M.D; ++d) out[size_t(c) * M.D + d] = double(hm[size_t(c) * M.Dp + d]);
    return x.check(err);
  }
  int32_t set_stepsize(const double* eps) {
    x.h2d(d_tmp_c, eps, size_t(M.C) * sizeof(double));
    x.set_eps(M, d_tmp_c);
    return x.check(err);
  }
  int32_t get_stepsize(double* eps) {
    x.get_eps(M, d_tmp_c);
    x.d2h(eps, d_tmp_c, size_t(M.C) * sizeof(double));
    return x.check(err);
  }

  // ---------------------------------------------------------------- run loop
  // Lockstep: [batched gradient] -> advance (consume leaf, merges, next leapfrog) -> ...
  // until every chain is idle.  Elementwise targets run inside advance().
  // Only chains with a gradient request occupy staging rows (active-chain
  // compaction), so the batched kernel cost follows the number of active chains.
  // Batched targets, one engine: the host does not wait for each step's request count.  Inside one call the
  // number of requesting chains never grows (a chain requests a gradient every lockstep step until it goes idle
  // for the rest of the call), so the newest count that has ARRIVED is an upper bound for the rows of the next
  // gradient launch: the launch covers all real requests plus, at worst, a few stale rows nobody reads.  Counts
  // come back through a small ring of pinned slots (async copy + event); the host runs at most LAG steps ahead.
  int32_t run_pipelined(bool pending) {
    typename X::Range nvtx_range("lockstep loop (pipelined): gradient kernel + k_advance per step");
    constexpr int LAG = 2;
    x.use();
    int64_t np_known = pending ? x.read_count() : -1;   // -1: nothing known yet (first step has no gradient)
    x.reset_counters();
    int64_t step = 0, completed = -1;
    bool first = !pending;
    for (;;) {
      if (np_known > 0) {
        M.stage_nb = x.gradient(*this, (int)np_known);
        M.stage_rows = (int32_t)np_known;
        counters.kernel_launches += x.launches_per_gradient();
      }
      x.advance_async(M, rp, 1, (int)(step % X::RING), (int)(step & 1));
      counters.kernel_launches += 1;
      counters.lockstep_steps += 1;
      ++step;
      // collect the counts that have arrived; block only if the host is LAG steps ahead or knows nothing yet
      while (completed + 1 < step) {
        const int slot = (int)((completed + 1) % X::RING);
        const bool must = first || (step - 1 - completed) >= LAG || np_known == 0;
        if (!must && !x.count_ready(slot)) break;
        const int64_t c = x.count_wait(slot);
        ++completed;
        counters.gradient_rows += c;       // rows requested by step `completed`, evaluated by the next launch
        np_known = c;
        first = false;
      }
      if (np_known == 0 && completed + 1 == step) break;
      if (x.failed()) break;
    }
    return x.check(err);
  }
  int32_t run(bool pending) {
    const bool batched = model.batched();
    x.use();
    if (batched && !reduce_on && X::RING > 0) return run_pipelined(pending);
    typename X::Range nvtx_range(reduce_on ? "lockstep loop (rows sharded): gradient kernel + exchange + k_advance per step"
                                           : (batched ? "lockstep loop" : "k_advance: every chain to the end of the call"));
    const int iters = batched ? 1 : (1 << 30);
    int64_t np = 0;
    if (pending) np = reduce_on ? x.assign_rows(reduce_view(0), M) : x.read_count();
    for (;;) {
      if (np > 0 && batched) {
        M.stage_nb = x.gradient(*this, (int)np);
        M.stage_rows = (int32_t)np;
        counters.kernel_launches += x.launches_per_gradient();
        counters.gradient_rows += np;
        if (reduce_on && p2p_on) {   // fold + push over NVLink in one kernel, wait + sum in rank order in another
          p2p_seq += 1;
          x.fold_push(M, (int)np, p2p_seq);
          x.wait_sum(M, (int)np, p2p_seq, red_g, red_l);
          counters.kernel_launches += 2;
        } else if (reduce_on) {   // fold the partials of this shard, sum them over the group (one exchange per leapfrog)
          x.fold_partials(M, (int)np, red_g, red_l);
          int32_t rc = x.allreduce(red_g, np * M.Dp, sizeof(T) == 4, red_l, np, red_fn, red_ctx, err);
          if (rc) return rc;
          counters.kernel_launches += 1;
        }
      }
      if (reduce_on) {
        const EngineMem<T> V = reduce_view(np);
        x.advance(V, rp, iters);
        np = x.assign_rows(V, M);
      } else {
        np = x.advance(M, rp, iters);
      }
      counters.kernel_launches += 1;
      counters.lockstep_steps += 1;
      if (np <= 0) break;
    }
    if (p2p_on && x.p2p_failed()) return fail(BNUTS_ERR_CUDA, "peer-memory exchange timed out waiting for a rank");
    return x.check(err);
  }
  void ensure_out(int N, bool want_draws) {
    const size_t nd = size_t(M.C) * N * M.D, ns = size_t(M.C) * N;
    if (want_draws && nd > cap_draws) {
      if (M.draws) x.free(M.draws);
      M.draws = x.template alloc<double>(nd);
      cap_draws = nd;
    }
    // chains with a sticky non-zero status (bnuts_chain_status) write no rows: hand back zeros, never stale memory
    if (want_draws) x.zero(M.draws, nd * sizeof(double));
    if (ns > cap_stats) {
      if (d_stats) { x.free(d_stats); x.free(d_sel); x.free(d_eps_hist); }
      d_stats = x.template alloc<TreeStats>(ns);
      d_sel = x.template alloc<int32_t>(ns);
      d_eps_hist = x.template alloc<double>(ns);
      cap_stats = ns;
    }
    x.zero(d_stats, ns * sizeof(TreeStats)); x.zero(d_sel, ns * sizeof(int32_t)); x.zero(d_eps_hist, ns * sizeof(double));
  }
  void copy_out(int N, double* chain_out, int64_t sd, int64_t sc, bnuts_tree_stats* stats_out, int64_t ssc,
                int32_t* sel, double* eps_out) {
    const int C = M.C, D = M.D;
    if (chain_out) {
      if (sd == D && sc == int64_t(N) * D) x.d2h(chain_out, M.draws, size_t(C) * N * D * sizeof(double));
      else {
        h_draws.resize(size_t(C) * N * D);
        x.d2h(h_draws.data(), M.draws, h_draws.size() * sizeof(double));
        for (int c = 0; c < C; ++c)
          for (int n = 0; n < N; ++n)
            std::memcpy(chain_out + c * sc + n * sd, &h_draws[(size_t(c) * N + n) * D], size_t(D) * sizeof(double));
      }
    }
    if (stats_out) {
      if (ssc == N) x.d2h(stats_out, d_stats, size_t(C) * N * sizeof(TreeStats));
      else for (int c = 0; c < C; ++c) x.d2h(stats_out + c * ssc, d_stats + size_t(c) * N, size_t(N) * sizeof(TreeStats));
    }
    if (sel) x.d2h(sel, d_sel, size_t(C) * N * sizeof(int32_t));
    if (eps_out) x.d2h(eps_out, d_eps_hist, size_t(C) * N * sizeof(double));
  }

  int32_t transitions(int N, const bnuts_dual_averaging* da, int metric_kind, double lambda, double* chain_out,
                      int64_t sd, int64_t sc, bnuts_tree_stats* stats_out, int64_t ssc, int32_t* sel, double* eps_out) {
    typename X::Range nvtx_range(da ? (metric_kind == BNUTS_METRIC_DIAG ? "bnuts_warmup_stage (step size + metric)" : "bnuts_warmup_stage (step size)")
                                    : (metric_kind == BNUTS_METRIC_DIAG ? "bnuts_warmup_stage (fixed step size, metric)" : "bnuts_sample"));
    if (model.kind == MODEL_NONE) return fail(BNUTS_ERR_NO_MODEL, "no model set");
    if (N <= 0) return fail(BNUTS_ERR_INVALID_ARGUMENT, "N must be positive");
    if (metric_kind != BNUTS_METRIC_NONE && metric_kind != BNUTS_METRIC_DIAG)
      return fail(BNUTS_ERR_INVALID_ARGUMENT, "bad metric_kind");
    if (metric_kind == BNUTS_METRIC_DIAG && dm.on)
      return fail(BNUTS_ERR_UNSUPPORTED, "diagonal adaptation on top of a dense metric is not supported");
    // the invariants the reference records as (commented-out) @argcheck's, src/stepsize.jl:183-186
    if (da && !(da->delta > 0.0 && da->delta < 1.0 && da->gamma > 0.0 && da->kappa > 0.5 && da->kappa <= 1.0 && da->t0 >= 0))
      return fail(BNUTS_ERR_INVALID_ARGUMENT, "dual averaging needs 0 < delta < 1, gamma > 0, 0.5 < kappa <= 1, t0 >= 0");
    // ≙ the (commented-out) @argcheck's of TuningNUTS, src/warmup.jl:230-231 (N ≥ 20, λ ≥ 0).  Enforced: the variance of
    // src/hamiltonian.jl:153-162 needs two draws (N = 1 gives mulreg = inf, a NaN metric), and λ must be ≥ 0 or the
    // default sentinel (any negative value used to be read as 5/N; only −1 is, now)
    if (metric_kind == BNUTS_METRIC_DIAG && N < 2)
      return fail(BNUTS_ERR_INVALID_ARGUMENT, "a metric-adapting stage needs N >= 2 (the reference records N >= 20)");
    if (metric_kind == BNUTS_METRIC_DIAG && !(lambda >= 0.0 || lambda == -1.0))
      return fail(BNUTS_ERR_INVALID_ARGUMENT, "lambda must be >= 0 (or -1 for the default 5/N)");
    const bool want_draws = chain_out != nullptr || metric_kind == BNUTS_METRIC_DIAG;
    ensure_out(N, want_draws);
    double* keep_draws = M.draws;
    if (!want_draws) M.draws = nullptr;
    rp.da_on = da ? 1 : 0;
    if (da) { rp.da.delta = da->delta; rp.da.gamma = da->gamma; rp.da.kappa = da->kappa; rp.da.t0 = da->t0; }
    rp.stats = d_stats; rp.sel = d_sel; rp.eps_hist = d_eps_hist; rp.n_total = N;
    PrepareArgs a{}; a.mode = MODE_SAMPLE; a.N = N; a.t0 = next_t;
    x.prepare(M, rp, a);
    int32_t rc = run(false);
    if (rc) { M.draws = keep_draws; return rc; }
    if (metric_kind == BNUTS_METRIC_DIAG) x.metric_update(M, N, lambda == -1.0 ? 5.0 / N : lambda);  // ≙ src/warmup.jl:308-309
    if (da) x.finish_da(M);                                                                     // ≙ final_ϵ, :313
    next_t += uint32_t(N);
    if (dm.on && chain_out) xf_device(M.draws, int64_t(M.C) * N, dm.dLt);   // draws back to user coordinates: q = L q̃
    copy_out(N, chain_out, sd, sc, stats_out, ssc, sel, eps_out);
    M.draws = keep_draws;
    rc = x.check(err);
    if (rc) return rc;
    if (x.any_status(M, ST_STEPSIZE_COLLAPSE))
      return fail(BNUTS_ERR_STEPSIZE_COLLAPSE, "step size fell below 1e-10 during adaptation");
    return 0;
  }

  int32_t set_positions(const double* q) {
    typename X::Range nvtx_range("bnuts_set_positions");
    if (model.kind == MODEL_NONE) return fail(BNUTS_ERR_NO_MODEL, "no model set");
    if (q && dm.on) {                       // q̃ = L⁻¹ q
      double* qt = xf_host(q, M.C, dm.dLinvt);
      x.d2d(d_tmp_cd, qt, size_t(M.C) * M.D * sizeof(double));
      M.pos_in = d_tmp_cd;
    } else if (q) { x.h2d(d_tmp_cd, q, size_t(M.C) * M.D * sizeof(double)); M.pos_in = d_tmp_cd; } else M.pos_in = nullptr;
    PrepareArgs a{}; a.mode = MODE_EVAL;
    if (reduce_on) x.prepare(reduce_view(0), rp, a); else x.prepare(M, rp, a);
    int32_t rc = run(true);
    if (rc) return rc;
    if (x.any_status(M, ST_NONFINITE_START)) return fail(BNUTS_ERR_NONFINITE_START, "starting point has non-finite density");
    return 0;
  }
  int32_t get_state(double* q, double* g, double* l) {
    double* o = d_tmp_cd;
    x.gather_state(M, o);
    const size_t n = size_t(M.C) * M.D;
    if (dm.on) { xf_device(o, M.C, dm.dLt); xf_device(o + n, M.C, dm.dLinv); }   // q = L q̃,  ∇ℓ = L⁻ᵀ ∇ℓ̃
    if (q) x.d2h(q, o, n * sizeof(double));
    if (g) x.d2h(g, o + n, n * sizeof(double));
    if (l) x.d2h(l, o + 2 * n, size_t(M.C) * sizeof(double));
    return x.check(err);
  }
  int32_t inject(int32_t Tn, const uint32_t* dirs, const double* p, const double* exps, int32_t n_exps) {
    if (Tn < 0) return fail(BNUTS_ERR_INVALID_ARGUMENT, "T < 0");
    if (exps && n_exps <= 0) return fail(BNUTS_ERR_INVALID_ARGUMENT, "exps given with n_exps <= 0");
    if (d_inj_dirs) { x.free(d_inj_dirs); d_inj_dirs = nullptr; }
    if (d_inj_p) { x.free(d_inj_p); d_inj_p = nullptr; }
    if (d_inj_exps) { x.free(d_inj_exps); d_inj_exps = nullptr; }
    if (exps && Tn > 0) {
      d_inj_exps = x.template alloc<double>(size_t(Tn) * M.C * n_exps);
      x.h2d(d_inj_exps, exps, size_t(Tn) * M.C * n_exps * sizeof(double));
    }
    if (dirs && Tn > 0) {
      d_inj_dirs = x.template alloc<uint32_t>(size_t(Tn) * M.C);
      x.h2d(d_inj_dirs, dirs, size_t(Tn) * M.C * sizeof(uint32_t));
    }
    if (p && Tn > 0) {
      d_inj_p = x.template alloc<double>(size_t(Tn) * M.C * M.D);
      if (dm.on) x.d2d(d_inj_p, xf_host(p, int64_t(Tn) * M.C, dm.dL), size_t(Tn) * M.C * M.D * sizeof(double));   // p̃ = Lᵀ p
      else x.h2d(d_inj_p, p, size_t(Tn) * M.C * M.D * sizeof(double));
    }
    rp.inj_T = Tn; rp.inj_start = next_t; rp.inj_dirs = d_inj_dirs; rp.inj_p = d_inj_p;
    rp.inj_exps = d_inj_exps; rp.inj_nexp = d_inj_exps ? n_exps : 0;
    return x.check(err);
  }
  int32_t leapfrog(const double* p_in, const double* eps, int nsteps, double* q_out, double* p_out, double* g_out,
                   double* l_out) {
    if (model.kind == MODEL_NONE) return fail(BNUTS_ERR_NO_MODEL, "no model set");
    if (!p_in || !eps || nsteps < 0) return fail(BNUTS_ERR_INVALID_ARGUMENT, "p_in, eps required");
    typename X::Range nvtx_range("bnuts_leapfrog");
    const size_t n = size_t(M.C) * M.D;
    if (!d_bare_p) { d_bare_p = x.template alloc<double>(n); d_bare_out = x.template alloc<double>(3 * n + M.C); }   // once per engine
    double* d_p = d_bare_p;
    double* d_out = d_bare_out;
    if (dm.on) x.d2d(d_p, xf_host(p_in, M.C, dm.dL), n * sizeof(double));   // p̃ = Lᵀ p
    else x.h2d(d_p, p_in, n * sizeof(double));
    x.h2d(d_tmp_c, eps, size_t(M.C) * sizeof(double));
    M.bare_p_in = d_p; M.bare_out = d_out;
    PrepareArgs a{}; a.mode = MODE_BARE; a.N = nsteps; a.bare_eps = d_tmp_c;
    x.prepare(M, rp, a);
    int32_t rc = run(false);
    if (!rc && dm.on) { xf_device(d_out, M.C, dm.dLt); xf_device(d_out + n, M.C, dm.dLinv); xf_device(d_out + 2 * n, M.C, dm.dLinv); }
    if (!rc) {
      if (q_out) x.d2h(q_out, d_out, n * sizeof(double));
      if (p_out) x.d2h(p_out, d_out + n, n * sizeof(double));
      if (g_out) x.d2h(g_out, d_out + 2 * n, n * sizeof(double));
      if (l_out) x.d2h(l_out, d_out + 3 * n, size_t(M.C) * sizeof(double));
      rc = x.check(err);
    }
    M.bare_p_in = nullptr; M.bare_out = nullptr;
    return rc;
  }
  int32_t find_initial_stepsize(const bnuts_stepsize_search& P) {
    typename X::Range nvtx_range("bnuts_find_initial_stepsize");
    if (model.kind == MODEL_NONE) return fail(BNUTS_ERR_NO_MODEL, "no model set");
    // ≙ the (commented-out) @argcheck's of InitialStepsizeSearch, src/stepsize.jl:31-35
    if (!(P.a_min > 0.0 && P.a_min < P.a_max && P.a_max < 1.0 && P.C > 1.0 && P.eps0 > 0.0 && P.maxiter_crossing > 0 && P.maxiter_bisect > 0))
      return fail(BNUTS_ERR_INVALID_ARGUMENT, "step size search needs 0 < a_min < a_max < 1, C > 1, eps0 > 0, positive iteration caps");
    rp.search.a_min = P.a_min; rp.search.a_max = P.a_max; rp.search.eps0 = P.eps0; rp.search.C = P.C;
    rp.search.maxiter_crossing = P.maxiter_crossing; rp.search.maxiter_bisect = P.maxiter_bisect;
    rp.da_on = 0;
    PrepareArgs a{}; a.mode = MODE_SEARCH; a.t0 = next_t;
    x.prepare(M, rp, a);
    int32_t rc = run(false);
    next_t += 1;
    if (rc) return rc;
    if (x.any_status(M, ST_NONFINITE_START) || x.any_status(M, ST_STEPSIZE_SEARCH))
      return fail(BNUTS_ERR_STEPSIZE_SEARCH, "initial step size search failed for some chains");
    return 0;
  }
  // ≙ warmup!(FindLocalOptimum), src/warmup.jl:152-186
  int32_t find_local_optimum(double magnitude_penalty, int32_t iterations) {
    typename X::Range nvtx_range("bnuts_find_local_optimum");
    if (model.kind == MODEL_NONE) return fail(BNUTS_ERR_NO_MODEL, "no model set");
    if (!(magnitude_penalty >= 0.0) || iterations < 0) return fail(BNUTS_ERR_INVALID_ARGUMENT, "bad FindLocalOptimum parameters");
    rp.opt.penalty = magnitude_penalty; rp.opt.iterations = iterations;
    rp.da_on = 0;
    PrepareArgs a{}; a.mode = MODE_OPT;
    x.prepare(M, rp, a);
    int32_t rc = run(false);
    if (rc) return rc;
    if (x.any_status(M, ST_OPTIMUM_FAILED))
      return fail(BNUTS_ERR_OPTIMUM, "optimization failed to converge for some chains (100 restarts)");
    return 0;
  }
  int32_t get_counters(bnuts_counter_block* out) {
    int64_t tot[3];
    x.totals(M, tot);
    counters.leapfrogs = tot[0]; counters.transitions = tot[1]; counters.divergences = tot[2];
    *out = counters;
    return x.check(err);
  }
  int32_t chain_status(int32_t* st) {
    x.get_status(M, st);
    return x.check(err);
  }
};

}  // namespace bn

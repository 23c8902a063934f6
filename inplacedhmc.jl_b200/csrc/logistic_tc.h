// logistic_tc.h — host interface of the tcgen05/TMA logistic-regression gradient
// kernel (logistic_tc.cu).  fp32 engine only; X must lie on the bf16 grid.
#pragma once
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/bnuts.h"
#include "backend.h"

namespace bn {

struct LogisticTC {
  // problem
  int32_t C = 0, D = 0, Dp = 0, Dt = 0, dk = 0;   // Dt: smem tile width (64|128), dk: MMA K / N (multiple of 16)
  int32_t aug = 0;           // 1: columns D..D+2 of X~ are reserved for the reference-point term (dk = round16(D+3))
  int32_t nterms = 3;        // bf16 terms of the position operand: 3 (exact split) or 2 (with a reference point)
  int64_t N = 0, Npad = 0;
  int32_t flush_every = 32;  // row blocks per TMEM-accumulator flush (0 = only at the end)
  int32_t sms = 148, max_splits = 1, force_nsplit = 0, last_nsplit = 1;
  int64_t partial_rows = 0;  // rows needed in the partial-output buffers
  // device buffers owned here
  uint16_t* Xb = nullptr;    // [Npad][Dt] bf16, rows sign-folded: X~_i = (2 y_i - 1) X_i
  double* colsum = nullptr;  // [Dp] column sums of X~ (linear part of the log density)
  float* beta_ref = nullptr; // [Dp] reference point (zeros when none is set)
  float* eta0 = nullptr;     // [Npad] X~ beta_ref per row (k_logistic_tc256 only)
  // single-term residual about the reference (see k_logistic_tc, template parameter RR)
  float* c0 = nullptr;       // [Npad] ½ − r0_i, r0_i = σ(−η̃0_i) (zero in the padding rows)
  double* grad0 = nullptr;   // [Dp] X̃ᵀ·r0 in Float64 from the stored fp32 values: added by the consumer (EngineMem::grad0)
  double* grad0_part = nullptr;  // [G0_BLOCKS][Dp] scratch of its two-pass (deterministic) reduction
  int32_t rmode = 0;         // residual operand: 0 two bf16 terms of r; 1 one term of δ = r − r0; 2 remainder mode (logistic_rm.cu)
  // remainder mode (logistic_rm.cu): per-row records (A2, A3, A4, η̃0), H0 = X̃ᵀ diag(w) X̃ (fp32 [D][Dp]), ℓ0, radius² of the Taylor path
  float* rec = nullptr; float* rm_r0 = nullptr; float* rm_w = nullptr; double* rm_f0 = nullptr;
  float* H0 = nullptr; double* H0_part = nullptr; double* rm_part = nullptr; double* ell0 = nullptr;
  double ell0_host = 0.0;
  float kappa2 = 0.f;
  // borrowed from the engine
  const uint16_t* bh = nullptr; const uint16_t* bm = nullptr; const uint16_t* bl = nullptr;  // [C][Dt]
  float* G = nullptr;        // [nsplit][rows][Dp]
  double* Ld = nullptr;      // [nsplit][C] log-density partials (Float64: ~1e5..1e6 in magnitude)
  // opaque tensor maps (4 x CUtensorMap, 128 B each, 64 B aligned): X, βh, βm, βl
  alignas(64) unsigned char tmaps[7][128];   // X (128-row box), βh, βm, βl, X (64-row box), βh and βm with 64-row boxes
  int32_t variant = 128;     // 128: k_logistic_tc (D <= 128, 128-row blocks); 256: k_logistic_tc256 (128 < D <= 256, 64-row blocks)
  bool ready = false;
  cudaError_t last = cudaSuccess;

  int plan_splits(int nrows, int tile_rows = 128) const;
  void run(cudaStream_t s, int nrows);
  void destroy();
};

// build device copies of X (bf16, padded) and y, the tensor maps and the split plan
int32_t logistic_tc_build(LogisticTC& tc, const uint16_t* Xbf16_host /*[N][D]*/, const double* y, int64_t N,
                          int32_t C, int32_t D, int32_t Dp, std::string& err);

// the same with the rows generated on the device (bnuts_model_logistic_synthetic); `fill` writes the sign-folded
// bf16 rows [row0, row0 + N) of the synthetic matrix into Xb [Npad][Dt]
typedef int (*SynthFillFn)(cudaStream_t s, uint64_t seed, int64_t row0, int64_t N, int32_t D, int32_t Dt, uint16_t* Xb);   // 0 or a cudaError_t
int32_t logistic_tc_build_synth(LogisticTC& tc, uint64_t seed, int64_t row0, int64_t N, int32_t C, int32_t D, int32_t Dp,
                                cudaStream_t s, SynthFillFn fill, std::string& err);

int32_t logistic_tc_maps(LogisticTC& tc, std::string& err);

// write H~0 = X~·beta_ref (three bf16 terms) into the reserved columns (beta_ref == nullptr: clear them)
void logistic_tc_write_reference(LogisticTC& tc, cudaStream_t s, const float* beta_ref_dev);
// c0_i = ½ − σ(−X̃_i·beta_ref) per row and grad0 = X̃ᵀ·(½ − c0) in Float64 (fixed summation order)
void logistic_tc_write_residual_reference(LogisticTC& tc, cudaStream_t s, const float* beta_ref_dev);
// staging rows: 1.0 in the reserved columns of the high term
void logistic_tc_init_stage(LogisticTC& tc, cudaStream_t s, uint16_t* bh);
// remainder mode (logistic_rm.cu): reference constants (records, g0 into tc.grad0, ℓ0, H0, κ²) and the launch
// group_sum (or null): adds an fp32 and a Float64 device buffer in place over the engines that share the rows (either may be empty)
typedef int32_t (*RmGroupSum)(void* ctx, float* f32, int64_t nf, double* f64, int64_t nd);
int32_t logistic_rm_write_reference(LogisticTC& tc, cudaStream_t s, const float* beta_ref_dev, RmGroupSum group_sum, void* group_ctx,
                                    int world, std::string& err);
void logistic_rm_launch(LogisticTC& tc, cudaStream_t s, int nrows, int nsplit, int nc);

template <class E> int32_t logistic_tc_attach(LogisticTC& tc, E& eng, std::string& err);

template <class E>
int32_t logistic_tc_setup(LogisticTC& tc, E& eng, const void* Xh, int32_t xd, const double* y, int64_t N, std::string& err) {
  using T = typename std::remove_reference<decltype(*eng.M.zs)>::type;
  if constexpr (!std::is_same<T, float>::value) {
    err = "tensor gradient path needs dtype F32";
    return BNUTS_ERR_UNSUPPORTED;
  } else {
    auto& M = eng.M;
    if (M.D > 256) { err = "tensor gradient path supports D <= 256 in this build"; return BNUTS_ERR_UNSUPPORTED; }
    // X must be exactly representable in bf16 (the data operand of the MMA is not split)
    const size_t n = size_t(N) * M.D;
    std::vector<uint16_t> xb(n);
    for (size_t i = 0; i < n; ++i) {
      if (xd == BNUTS_X_BF16) { xb[i] = static_cast<const uint16_t*>(Xh)[i]; continue; }
      const double v = xd == BNUTS_X_F64 ? static_cast<const double*>(Xh)[i] : double(static_cast<const float*>(Xh)[i]);
      const uint16_t h = bf16_bits(float(v));
      if (double(bf16_val(h)) != v) {
        err = "tensor gradient path needs X on the bf16 grid (element not representable); use BNUTS_GRAD_DETERMINISTIC";
        return BNUTS_ERR_UNSUPPORTED;
      }
      xb[i] = h;
    }
    int32_t rc = logistic_tc_build(tc, xb.data(), y, N, M.C, M.D, M.Dp, err);
    if (rc) return rc;
    return logistic_tc_attach(tc, eng, err);
  }
}

// engine-owned staging of the tensor path (bf16 terms of q, partial outputs), tensor maps
template <class E>
int32_t logistic_tc_attach(LogisticTC& tc, E& eng, std::string& err) {
  auto& M = eng.M;
  M.Dt = tc.Dt;
  const size_t nbt = size_t(M.C) * tc.Dt;
  M.stage_bh = eng.x.template alloc<uint16_t>(nbt);
  M.stage_bm = eng.x.template alloc<uint16_t>(nbt);
  M.stage_bl = eng.x.template alloc<uint16_t>(nbt);
  eng.x.zero(M.stage_bh, nbt * 2); eng.x.zero(M.stage_bm, nbt * 2); eng.x.zero(M.stage_bl, nbt * 2);
  eng.alloc_stage(1, size_t(tc.partial_rows));
  M.stage_ld = eng.x.template alloc<double>(size_t(tc.partial_rows));
  eng.x.zero(M.stage_ld, size_t(tc.partial_rows) * sizeof(double));
  tc.bh = M.stage_bh; tc.bm = M.stage_bm; tc.bl = M.stage_bl; tc.G = M.stage_g; tc.Ld = M.stage_ld;
  logistic_tc_init_stage(tc, eng.x.stream, M.stage_bh);
  M.lin_w = tc.colsum;
  M.beta_ref = tc.beta_ref;
  const int32_t rc = logistic_tc_maps(tc, err);
  if (rc) return rc;
  eng.model.Npad = tc.Npad;
  return 0;
}

template <class E>
int32_t logistic_tc_setup_synth(LogisticTC& tc, E& eng, uint64_t seed, int64_t row0, int64_t N, SynthFillFn fill, std::string& err) {
  using T = typename std::remove_reference<decltype(*eng.M.zs)>::type;
  if constexpr (!std::is_same<T, float>::value) {
    err = "tensor gradient path needs dtype F32";
    return BNUTS_ERR_UNSUPPORTED;
  } else {
    auto& M = eng.M;
    if (M.D > 256) { err = "tensor gradient path supports D <= 256 in this build"; return BNUTS_ERR_UNSUPPORTED; }
    const int32_t rc = logistic_tc_build_synth(tc, seed, row0, N, M.C, M.D, M.Dp, eng.x.stream, fill, err);
    if (rc) return rc;
    return logistic_tc_attach(tc, eng, err);
  }
}

// ≙ no reference counterpart (numerical device of the tensor path, see bnuts.h).  Evaluates the
// gradient at beta_ref with the exact three-term path first and refuses a point whose gradient
// is larger than a typical posterior-bulk gradient sqrt(tr H) <= sqrt(N D)/2: the two-term
// split of beta - beta_ref is accurate relative to H·(beta - beta_ref), which is the gradient
// only if beta_ref sits at the mode to within the posterior width.
template <class E>
int32_t logistic_tc_set_reference(LogisticTC& tc, E& eng, const double* beta_ref, std::string& err) {
  using T = typename std::remove_reference<decltype(*eng.M.zs)>::type;
  if constexpr (!std::is_same<T, float>::value) {
    err = "reference point needs the tensor gradient path (dtype F32)";
    return BNUTS_ERR_UNSUPPORTED;
  } else {
    auto& M = eng.M;
    auto& x = eng.x;
    if (!tc.ready) { err = "reference point needs the tensor gradient path"; return BNUTS_ERR_UNSUPPORTED; }
    // back to the exact path (D <= 128) / to the zero reference (D > 128: always two terms, see k_logistic_tc256)
    tc.nterms = tc.variant == 256 ? 2 : 3;
    tc.rmode = 0;
    M.grad0 = nullptr; M.lin_H = nullptr; M.ell0 = 0.0; M.lin_w = tc.colsum;
    x.zero(tc.beta_ref, size_t(M.Dp) * sizeof(float));
    logistic_tc_write_reference(tc, x.stream, nullptr);
    if (!beta_ref) return x.check(err);
    if (!tc.aug && tc.variant != 256) { err = "no spare K columns for the reference term (125 < D <= 128)"; return BNUTS_ERR_UNSUPPORTED; }
    std::vector<float> b(size_t(M.Dp), 0.f);
    std::vector<uint16_t> h(size_t(tc.Dt), 0), m(size_t(tc.Dt), 0), l(size_t(tc.Dt), 0);
    for (int d = 0; d < M.D; ++d) {
      b[d] = float(beta_ref[d]);
      if (!(b[d] == b[d]) || b[d] - b[d] != 0.f) { err = "reference point is not finite"; return BNUTS_ERR_INVALID_ARGUMENT; }
      h[d] = bf16_bits(b[d]);
      const float r1 = b[d] - bf16_val(h[d]);
      m[d] = bf16_bits(r1);
      l[d] = bf16_bits(r1 - bf16_val(m[d]));
    }
    if (tc.aug) for (int k = 0; k < 3; ++k) h[M.D + k] = 0x3F80;
    x.h2d(M.stage_bh, h.data(), h.size() * 2); x.h2d(M.stage_bm, m.data(), m.size() * 2); x.h2d(M.stage_bl, l.data(), l.size() * 2);
    tc.run(x.stream, 1);
    std::vector<double> g;
    int32_t rc = eng.folded_row0(tc.last_nsplit, g);   // summed over the group when the rows are sharded
    if (rc) return rc;
    double n2 = 0.0;
    for (int d = 0; d < M.D; ++d) {
      const double acc = g[d] - double(M.tau) * double(b[d]);
      n2 += acc * acc;
    }
    const double bound = 0.5 * std::sqrt(double(tc.N) * double(eng.reduce_world()) * double(M.D));
    // BNUTS_TC_DEBUG_NOCHECK=1: timing experiments with ablated kernel builds (-DBNUTS_TC_DEBUG, wrong results by design)
    const char* nochk = std::getenv("BNUTS_TC_DEBUG_NOCHECK");
    // D > 128 (k_logistic_tc256) ALWAYS works about a reference point (zero until one is set) and has no exact
    // alternative to fall back to: any finite point nearer to the chains than the current one improves the operand
    // β − β₀, so a point that is not yet at the mode is accepted there (callers iterate: optimise, set, optimise, set)
    if (!(n2 <= bound * bound) && tc.variant != 256 && !(nochk && std::atoi(nochk) != 0)) {
      err = "reference point rejected: |grad| = " + std::to_string(std::sqrt(n2)) + " exceeds sqrt(N D)/2 = " +
            std::to_string(bound) + " (not at the mode); exact three-term path kept";
      return BNUTS_ERR_INVALID_ARGUMENT;
    }
    x.h2d(tc.beta_ref, b.data(), size_t(M.Dp) * sizeof(float));
    logistic_tc_write_reference(tc, x.stream, tc.beta_ref);
    tc.nterms = 2;
    // residual about the reference as well (one bf16 term of δ = r − r0 instead of r = rh + rl; see k_logistic_tc):
    // error ≈ 1.7e-3·sqrt(D/N)·|∇ℓ|, below the path's tolerance for tall problems only, so the engine takes it by itself
    // when N >= 3.3e5·D over the whole row group (config 5); BNUTS_TC_RREF=0/1 overrides.
    const char* rr = std::getenv("BNUTS_TC_RREF");
    const bool rr_auto = double(tc.N) * double(eng.reduce_world()) >= 3.3e5 * double(M.D);
    // Remainder mode (logistic_rm.cu; D <= 125, rows not sharded): the model is expanded about the reference row by row, the
    // linear / quadratic part is exact D x D arithmetic in the consumer and only the small remainder goes through the tensor
    // cores, as ONE bf16 term.  Default where it applies; BNUTS_TC_RMODE=0/1/2 overrides (0 / 1: the residual operands above).
    const char* rme = std::getenv("BNUTS_TC_RMODE");
    // (D <= 125: the spare K columns are not needed, but the exact path the check above ran on has them; 128 < D <= 256: 64-chain
    // tiles; rows sharded: the set-up sums go over the row group, which needs a collective — NCCL or the host callback, not the
    // peer-memory exchange)
    const bool rm_ok = ((tc.variant == 128 && tc.aug) || tc.variant == 256) && (!eng.reduce_on || !eng.p2p_on);
    // ... by itself only for tall problems: the posterior's row-wise rms of δ is about sqrt(D / (N/5)); beyond ~0.04 most chains
    // would sit outside the Taylor radius and take the closed forms (correct, but no faster than the modes above)
    const bool rm_auto = rm_ok && double(tc.N) * double(eng.reduce_world()) >= 3000.0 * double(M.D);
    int want = rme ? std::atoi(rme) : (rr ? (std::atoi(rr) != 0 ? 1 : 0) : (rm_auto ? 2 : (rr_auto ? 1 : 0)));
    if (want == 2 && !rm_ok) want = rr_auto ? 1 : 0;
    if (want == 2) {
      logistic_tc_write_reference(tc, x.stream, nullptr);   // S = X̃·(β − β₀) only: the reference columns of X̃ stay clear
      struct Ctx { E* eng; } ctx{&eng};
      RmGroupSum gs = nullptr;
      if (eng.reduce_on)
        gs = [](void* c, float* f32, int64_t nf, double* f64, int64_t nd) -> int32_t {
          E& en = *static_cast<Ctx*>(c)->eng;
          // x.allreduce sums one fp32 and one Float64 buffer; an empty one is replaced by a one-element scratch
          float* fb = nf ? f32 : reinterpret_cast<float*>(en.red_g);
          double* db = nd ? f64 : en.red_l;
          return en.x.allreduce(fb, nf ? nf : 1, true, db, nd ? nd : 1, en.red_fn, en.red_ctx, en.err);
        };
      int32_t rc2 = logistic_rm_write_reference(tc, x.stream, tc.beta_ref, gs, &ctx, eng.reduce_world(), err);
      if (rc2) return rc2;
      tc.rmode = 2;
      M.grad0 = tc.grad0; M.lin_H = tc.H0; M.ell0 = tc.ell0_host; M.lin_w = nullptr;
    } else if (want == 1) {
      logistic_tc_write_residual_reference(tc, x.stream, tc.beta_ref);
      tc.rmode = 1;
      M.grad0 = tc.grad0;
    }
    return x.check(err);
  }
}

}  // namespace bn

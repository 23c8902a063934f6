// logistic_tc.h — host interface of the tcgen05/TMA logistic-regression gradient
// kernel (logistic_tc.cu).  fp32 engine only; X must lie on the bf16 grid.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/bnuts.h"
#include "backend.h"

namespace bn {

struct LogisticTC {
  // problem
  int32_t C = 0, D = 0, Dp = 0, Dt = 0, dk = 0;   // Dt: smem tile width (64|128), dk = round16(D): MMA K / N
  int64_t N = 0, Npad = 0;
  int32_t flush_every = 32;  // row blocks per TMEM-accumulator flush (0 = only at the end)
  int32_t sms = 148, max_splits = 1, force_nsplit = 0, last_nsplit = 1;
  int64_t partial_rows = 0;  // rows needed in the partial-output buffers
  // device buffers owned here
  uint16_t* Xb = nullptr;    // [Npad][Dt] bf16
  float* yf = nullptr;       // [Npad]
  // borrowed from the engine
  const uint16_t* bh = nullptr; const uint16_t* bm = nullptr; const uint16_t* bl = nullptr;  // [C][Dt]
  float* G = nullptr;        // [nsplit][rows][Dp]
  double* Ld = nullptr;      // [nsplit][C] log-density partials (Float64: ~1e5..1e6 in magnitude)
  // opaque tensor maps (4 x CUtensorMap, 128 B each, 64 B aligned): X, βh, βm, βl
  alignas(64) unsigned char tmaps[4][128];
  bool ready = false;
  cudaError_t last = cudaSuccess;

  int plan_splits(int nrows) const;
  void run(cudaStream_t s, int nrows);
  void destroy();
};

// build device copies of X (bf16, padded) and y, the tensor maps and the split plan
int32_t logistic_tc_build(LogisticTC& tc, const uint16_t* Xbf16_host /*[N][D]*/, const double* y, int64_t N,
                          int32_t C, int32_t D, int32_t Dp, std::string& err);

int32_t logistic_tc_maps(LogisticTC& tc, std::string& err);

template <class E>
int32_t logistic_tc_setup(LogisticTC& tc, E& eng, const void* Xh, int32_t xd, const double* y, int64_t N, std::string& err) {
  using T = typename std::remove_reference<decltype(*eng.M.zs)>::type;
  if constexpr (!std::is_same<T, float>::value) {
    err = "tensor gradient path needs dtype F32";
    return BNUTS_ERR_UNSUPPORTED;
  } else {
    auto& M = eng.M;
    if (M.D > 128) { err = "tensor gradient path supports D <= 128 in this build"; return BNUTS_ERR_UNSUPPORTED; }
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
    // staging owned by the engine: bf16 hi/lo of q, partial outputs
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
    rc = logistic_tc_maps(tc, err);
    if (rc) return rc;
    eng.model.Npad = tc.Npad;
    return 0;
  }
}

}  // namespace bn

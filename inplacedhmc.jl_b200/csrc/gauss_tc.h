// gauss_tc.h — host interface of the tcgen05/TMA gradient kernel of the Gaussian target (gauss_tc.cu):
// G = −P·Q for thousands of chains in lockstep (BASELINE config 2; with a shared dense metric P is the
// whitened precision P̃ = LᵀPL, see engine_core.h).  fp32 engine only.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/bnuts.h"
#include "backend.h"

namespace bn {

struct GaussTC {
  int32_t C = 0, D = 0, Dp = 0, Kp = 0, Np = 0;   // Kp = round64(D) (MMA K), Np = round128(D) (rows of P)
  uint16_t* P3 = nullptr;                          // [3][Np][Kp] bf16: exact three-term split of −float(P)
  const uint16_t* qh = nullptr; const uint16_t* qm = nullptr; const uint16_t* ql = nullptr;   // borrowed staging [C][Kp]
  float* G = nullptr;                              // borrowed [C][Dp]
  alignas(64) unsigned char tmaps[6][128];         // Qh, Qm, Ql, Ph, Pm, Pl
  bool ready = false;
  void run(cudaStream_t s, int nrows);
  void destroy();
};

int32_t gauss_tc_build(GaussTC& gt, const std::vector<double>& P, int32_t C, int32_t D, int32_t Dp, std::string& err);
int32_t gauss_tc_maps(GaussTC& gt, std::string& err);

}  // namespace bn

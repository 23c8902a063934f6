// gauss_tc.cu — G = −P·Q on tcgen05 + TMA (sm_100a) for the Gaussian target.
//
// Replaces the model call of the reference's leapfrog (logdensity_and_gradient!, call site
// src/kinetic_energy.jl:73) for ℓ(q) = −½ qᵀPq when thousands of chains advance in lockstep: the gradients of
// all requesting chains are one GEMM  G[chains x D] = Q[chains x D] · (−P)ᵀ.
//
// fp32 accuracy on bf16 tensor cores: both operands are split EXACTLY into three bf16 terms
// (x = xh + xm + xl, 3 x 8 mantissa bits; Q by the chains when they stage their request, −float(P) once at
// set-up) and the six products of total order <= 2 are accumulated in fp32 in TMEM
// (lh, hl, mm, mh, hm, hh — small ones first); the dropped terms are below 2^-24 of the result.
//
// One CTA = 128 chains x 128 output coordinates; K runs over D in 64-column chunks, two stages of
// (3 Q tiles + 3 P tiles) x 16 KB.  Warp 0: TMA producer, warp 1: MMA issuer (batched asm blocks, see
// tc_ptx.h), warps 2-5: epilogue (TMEM -> registers -> global, one chain per thread).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>

#include "gauss_tc.h"

namespace bn {

namespace {

#include "tc_ptx.h"

constexpr int GT_THREADS = 192;
constexpr int GT_NS = 2;                              // stages
constexpr int GT_STAGE_BYTES = 6 * CHUNK_BYTES;       // Qh Qm Ql Ph Pm Pl
constexpr int GT_SMEM = GT_NS * GT_STAGE_BYTES + 64 + 1024;

__global__ void __launch_bounds__(GT_THREADS, 1)
k_gauss_tc(const __grid_constant__ CUtensorMap tmQh, const __grid_constant__ CUtensorMap tmQm,
           const __grid_constant__ CUtensorMap tmQl, const __grid_constant__ CUtensorMap tmPh,
           const __grid_constant__ CUtensorMap tmPm, const __grid_constant__ CUtensorMap tmPl, float* G, int nrows, int Dp,
           int nchunks) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GT_NS * GT_STAGE_BYTES);
  uint64_t* full = bars;              // [NS]
  uint64_t* empty = bars + GT_NS;     // [NS]
  uint64_t* acc_full = bars + 2 * GT_NS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * GT_NS + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_m = blockIdx.x, tile_n = blockIdx.y;

  if (threadIdx.x == 0) {
    for (int i = 0; i < GT_NS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kc = 0; kc < nchunks; ++kc) {
        const int st = kc % GT_NS;
        mbar_wait(&empty[st], ((uint32_t)(kc / GT_NS) & 1u) ^ 1u);
        mbar_expect_tx(&full[st], GT_STAGE_BYTES);
        unsigned char* base = smem + st * GT_STAGE_BYTES;
        tma_load_2d(&tmQh, base + 0 * CHUNK_BYTES, &full[st], kc * 64, tile_m * 128);
        tma_load_2d(&tmQm, base + 1 * CHUNK_BYTES, &full[st], kc * 64, tile_m * 128);
        tma_load_2d(&tmQl, base + 2 * CHUNK_BYTES, &full[st], kc * 64, tile_m * 128);
        tma_load_2d(&tmPh, base + 3 * CHUNK_BYTES, &full[st], kc * 64, tile_n * 128);
        tma_load_2d(&tmPm, base + 4 * CHUNK_BYTES, &full[st], kc * 64, tile_n * 128);
        tma_load_2d(&tmPl, base + 5 * CHUNK_BYTES, &full[st], kc * 64, tile_n * 128);
      }
    }
  } else if (warp == 1) {
    // M = 128 (chains), N = 128 (coordinates), bf16 x bf16 -> f32, both operands K-major
    constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t dKM = desc_kmajor(0, 0);
    const uint32_t hi = (uint32_t)(dKM >> 32), lo0 = (uint32_t)dKM;
    // (Q term, P term) pairs, small products first: l·h, h·l, m·m, m·h, h·m, h·h
    constexpr int QA[6] = {2, 0, 1, 1, 0, 0};
    constexpr int PB[6] = {0, 2, 1, 0, 1, 0};
    for (int kc = 0; kc < nchunks; ++kc) {
      const int st = kc % GT_NS;
      mbar_wait(&full[st], (uint32_t)(kc / GT_NS) & 1u);
      tc_fence_after();
      const uint32_t base = lo0 + ((smem_u32(smem) + (uint32_t)st * GT_STAGE_BYTES) >> 4);
#pragma unroll
      for (int pr = 0; pr < 6; ++pr)
        mma_ss_run<4>(tmem, base + (uint32_t)(QA[pr] * (CHUNK_BYTES >> 4)), hi, base + (uint32_t)((3 + PB[pr]) * (CHUNK_BYTES >> 4)), hi,
                      IDESC, (kc | pr) ? 1u : 0u);
      if (elect_one()) tc_commit(&empty[st]);
      __syncwarp();
    }
    if (elect_one()) tc_commit(acc_full);
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int row = tile_m * 128 + q * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    mbar_wait(acc_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int ch = 0; ch < 4; ++ch) {
      uint32_t v[32];
      tmem_ld32(tmem + lane_sel + (uint32_t)ch * 32u, v);
      tmem_ld_wait();
      if (row < nrows) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const int d = tile_n * 128 + ch * 32 + j;
          if (d < Dp)
            *reinterpret_cast<float4*>(G + (size_t)row * Dp + d) =
                make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
  }
}

}  // namespace

void GaussTC::run(cudaStream_t s, int nrows) {
  if (!ready || nrows <= 0) return;
  static unsigned long long attr_done = 0;   // per device ordinal
  int dev = 0;
  cudaGetDevice(&dev);
  if (!((attr_done >> (dev & 63)) & 1ull)) {
    cudaFuncSetAttribute(k_gauss_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, GT_SMEM);
    attr_done |= 1ull << (dev & 63);
  }
  CUtensorMap m[6];
  for (int i = 0; i < 6; ++i) std::memcpy(&m[i], tmaps[i], sizeof(CUtensorMap));
  dim3 grid((nrows + 127) / 128, Np / 128);
  k_gauss_tc<<<grid, GT_THREADS, GT_SMEM, s>>>(m[0], m[1], m[2], m[3], m[4], m[5], G, nrows, Dp, Kp / 64);
}
void GaussTC::destroy() {
  if (P3) cudaFree(P3);
  P3 = nullptr; ready = false;
}

int32_t gauss_tc_build(GaussTC& gt, const std::vector<double>& P, int32_t C, int32_t D, int32_t Dp, std::string& err) {
  gt.destroy();
  gt.C = C; gt.D = D; gt.Dp = Dp;
  gt.Kp = (D + 63) / 64 * 64;
  gt.Np = (D + 127) / 128 * 128;
  const size_t plane = size_t(gt.Np) * gt.Kp;
  std::vector<uint16_t> h(3 * plane, 0);
  for (int i = 0; i < D; ++i)
    for (int k = 0; k < D; ++k) {
      const float v = -float(P[size_t(i) * D + k]);          // the deterministic path multiplies by float(P) as well
      const uint16_t a = bf16_bits(v);
      const float r1 = v - bf16_val(a);
      const uint16_t b = bf16_bits(r1);
      const uint16_t c = bf16_bits(r1 - bf16_val(b));
      const size_t o = size_t(i) * gt.Kp + k;
      h[o] = a; h[plane + o] = b; h[2 * plane + o] = c;
    }
  if (cudaMalloc(&gt.P3, h.size() * 2) != cudaSuccess) { err = "device allocation failed (tensor path P)"; return BNUTS_ERR_CUDA; }
  cudaMemcpy(gt.P3, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  return 0;
}
int32_t gauss_tc_maps(GaussTC& gt, std::string& err) {
  const size_t plane = size_t(gt.Np) * gt.Kp;
  const uint16_t* q[3] = {gt.qh, gt.qm, gt.ql};
  for (int t = 0; t < 3; ++t) {
    if (!encode_map(gt.tmaps[t], q[t], (uint64_t)gt.C, (uint64_t)gt.Kp) ||
        !encode_map(gt.tmaps[3 + t], gt.P3 + t * plane, (uint64_t)gt.Np, (uint64_t)gt.Kp)) {
      err = "cuTensorMapEncodeTiled failed";
      return BNUTS_ERR_CUDA;
    }
  }
  gt.ready = true;
  return 0;
}

}  // namespace bn

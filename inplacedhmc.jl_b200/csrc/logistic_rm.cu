// logistic_rm.cu — "remainder mode" of the tensor-core logistic-regression gradient (sm_100a, tcgen05 + TMA).
//
// Same boundary as logistic_tc.cu (the model call of the reference's leapfrog, logdensity_and_gradient!, call site
// src/kinetic_energy.jl:73, for thousands of chains in lockstep), taken once a reference point β₀ near the posterior
// mode is known (bnuts_logistic_set_reference).  With δ_ic = x̃_i·(β_c − β₀) and, per data row, η̃0_i = x̃_i·β₀,
// r0 = σ(−η̃0), w = σ'(η̃0), the residual r = σ(−η̃0 − δ) is expanded about the reference:
//     r = r0 − w δ + ρ(δ),      ρ = δ²(A₂ + A₃δ + A₄δ²) + O(δ⁵),   A₂ = −wu/2, A₃ = −w(u² − 2w)/6, A₄ = −wu(u² − 8w)/24, u = 2 r0 − 1
// so that
//     ∇ℓ = g0 − H0 (β − β₀) + X̃ᵀρ,                         g0 = X̃ᵀ r0,  H0 = X̃ᵀ diag(w) X̃   (Float64 at set-up)
//     ℓ  = ℓ0 + g0·(β − β₀) − ½ (β − β₀)ᵀH0(β − β₀) + Σ_i λ_i,    λ = δ³(A₂/3 + A₃δ/4 + A₄δ²/5)   (dλ/dδ = ρ)
// and because Σ_i δ_ic ρ_ic = (β_c − β₀)·(X̃ᵀρ)_c is a D-dot product of numbers the consumer holds anyway, the kernel only sums
//     μ = λ − δρ/3 = −δ⁴(A₃/12 + 2 A₄ δ/15)      (fourth order: ~1e-4 of ℓ's remainder in the bulk; the δ⁵ term is dropped in the Taylor form)
// The D x D linear part is EXACT fp32/Float64 arithmetic done by the consumer (backend.h model_grad); only the remainder
// goes through the tensor cores.  In the posterior bulk |δ| ≈ 0.02 and |X̃ᵀρ| is < 1 % of the gradient, so ρ needs three
// digits, not seven: ONE bf16 term carries it (2⁻⁹ relative, random over the rows: ~1e-7 of |∇ℓ| at N = 1e6), the
// elementwise stage is six packed FMAs and one bf16 pack per pair of chains — no exp, no reciprocal, no hi/lo split, no log
// (ablations of round 1's k_logistic_tc, profiles/r2_k_logistic_tc_ablation.txt: 2.4 ms of its 2.8 ms remain with both GEMMs
// removed: the sigmoid, not the tensor pipe, set its pace).
//
// Layout: DATA ROWS are the MMA M dimension (TMEM lane = row), chains the N dimension:
//     GEMM1  S[128 rows x NC chains]   = X̃blk[128 x K] · ΔBᵀ[K x NC]      (A, B from smem, K-major; ΔB = β − β₀ in two bf16 terms)
//     GEMM2  Gᵀ[128 features x NC]    += X̃blkᵀ[128 x 128 rows] · R[128 rows x NC]   (A = the same X̃ tile read MN-major,
//                                                                          B = R written to shared memory by the elementwise warps)
// so the per-row constants (A₂, A₃, A₄, η̃0) are per-THREAD registers (with chains as M they are uniform across a warp and
// every one costs a broadcast shared-memory load: the quadratic-remainder experiment of round 1 died of exactly that), the
// MMA cost follows the number of chains in the tile (NC = 64 for launches of <= 64 chains: HBM-bound, not tile-bound), and
// the per-chain log-density sum is a per-thread register accumulation reduced once at the end.
//
// Chains far from the reference (‖β − β₀‖² > κ², measured by the kernel itself from the staged operand tile): a tile that
// holds one switches, for the whole launch, to a loop whose flagged chains take the closed forms ρ = σ(−η̃0 − δ) − r0 + wδ and
// λ from log σ, and which hands R to GEMM2 in TWO bf16 halves per block (hi, lo: the remainder is no longer small there); the
// log-density remainder is folded into Float64 per chain every block.  Correct everywhere (60 posterior sd: 7e-6 of |∇ℓ|),
// fast near the reference.
//
// Warp roles (608 threads): warps 0-15 elementwise / epilogue: TMEM lane quarter q = warp % 4 (hardware rule), chain group
// warp / 4 (32 chains each) — four warps per scheduler: packed f32x2 arithmetic issues every 2.0 clk with four warps per
// sub-partition and every 2.6 clk with two (scripts/micro/fma_pipe.cu); warp 16 TMA producer, warp 17 GEMM1 issuer + TMEM
// owner, warp 18 GEMM2 issuer.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <cstdlib>

#include <type_traits>

#include "logistic_tc.h"

namespace bn {

namespace {

constexpr int RM_THREADS = 608;      // warps 0-15 elementwise / epilogue, warp 16 TMA producer, warp 17 GEMM1 issuer + TMEM owner, warp 18 GEMM2 issuer
constexpr int RM_TMA_WARP = 16, RM_G1_WARP = 17, RM_G2_WARP = 18;
constexpr int ROWS = 128;            // data rows per block (GEMM1 M, GEMM2 K)
#ifndef BNUTS_RM_DEBUG
#define BNUTS_RM_DEBUG 0      // timing experiments only (results are wrong): 1 skip the elementwise arithmetic, 2 skip GEMM1, 4 skip GEMM2, 8 skip the stores of R
#endif
constexpr int RMDBG = BNUTS_RM_DEBUG;
#ifndef BNUTS_RM_ARRIVE_ALL
#define BNUTS_RM_ARRIVE_ALL 0   // 1: every elementwise thread arrives on the hand-off barriers; 0: one elected lane per warp after __syncwarp
#endif
constexpr int ARR_ALL = BNUTS_RM_ARRIVE_ALL;
#include "tc_ptx.h"

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void ew_arrive(uint64_t* bar, int lane) {
  if (ARR_ALL) { mbar_arrive(bar); return; }
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}
// v[j] = (chains 2j, 2j+1) of this lane's row slot: returns, in lane l, the sum over the warp's 32 lanes of chain l
// (recursive halving: 31 shuffles instead of 32 x 5)
__device__ __forceinline__ float lane_sums(const float2* v2, int lane) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 16; ++j) { v[2 * j] = v2[j].x; v[2 * j + 1] = v2[j].y; }
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int k = 0; k < s; ++k) {
      const float send = up ? v[k] : v[k + s];
      const float keep = up ? v[k + s] : v[k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}
// GEMM2: N4 consecutive K steps (16 data rows each); both operands are MN-major tiles whose K rows are 128 B apart,
// so both descriptors advance 2048 B = 128 units per step
#define BN_MN_STEP "add.u32 al, al, 128;\n\tadd.u32 bl, bl, 128;\n\t"
template <int N4>
__device__ __forceinline__ void mma_ss_mn_run(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                              uint32_t idesc, uint32_t acc_first) {
  static_assert(N4 == 4, "four K steps per call");
  asm volatile(BN_SS_HEAD BN_MMA_SS("pa") BN_MN_STEP BN_MMA_SS("pt") BN_MN_STEP BN_MMA_SS("pt") BN_MN_STEP BN_MMA_SS("pt") "}\n"
               BN_SS_ARGS);
}


template <int DT, int NC, int NT> struct RmPlan {
  static constexpr int KC = DT / 64;
  static constexpr int BR = NC > 64 ? 128 : 64;              // staged rows per ΔB tile (TMA box height)
  static constexpr int B_CHUNK = BR * 128;                   // one 64-column chunk of a ΔB tile
  static constexpr int B_BYTES = KC * B_CHUNK;               // one ΔB term
  static constexpr int X_BYTES = KC * CHUNK_BYTES;           // one X stage: 128 rows x DT columns
  static constexpr int R_BYTES = (NC > 64 ? 2 : 1) * CHUNK_BYTES;   // one R buffer: 128 rows x NC chains (64-chain chunks)
  static constexpr int NSB = 3;                              // S buffers in TMEM
  static constexpr int NRB = 2;                              // R buffers in shared memory
  static constexpr int NGH = DT > 128 ? 2 : 1;               // 128-feature halves of the GEMM2 accumulator
  static_assert(NSB * 128 + NGH * NC <= 512, "TMEM columns");
  static constexpr int FIXED = NT * B_BYTES + NRB * R_BYTES + 512;
  static constexpr int NS0 = (232448 - FIXED) / X_BYTES;     // X stages: as many as fit (the stage of block i is held from its TMA
  static constexpr int NS = NS0 > 8 ? 8 : NS0;               // load until GEMM2(i) has read it, so depth hides the HBM latency)
  static constexpr int OFF_B = 0;
  static constexpr int OFF_X = NT * B_BYTES;
  static constexpr int OFF_RB = OFF_X + NS * X_BYTES;
  static constexpr int OFF_BAR = OFF_RB + NRB * R_BYTES;
  static constexpr int NBAR = 1 + 2 * NS + 2 * NSB + 2 * NRB + 2;
  static constexpr int TOTAL = OFF_BAR + NBAR * 8 + 16;
};

// NT = bf16 terms of β − β₀ in GEMM1.  Two: with one, the truncation of the operand is a coherent perturbation of δ over the rows and
// the gradient error measured 5e-6 (median) … 1.4e-5 of |∇ℓ| at one posterior sd, for 3 % of time: only NT = 2 is instantiated.
template <int DT, int NK, int NC, int NT>
__global__ void __launch_bounds__(RM_THREADS, 1)
k_logistic_rm(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBm,
              const float4* __restrict__ rec, float* G, double* Ld, int nrows, int D, int Dp, int nblk_total, int nsplit,
              int flush_every, float kappa2) {
  using P = RmPlan<DT, NC, NT>;
  static_assert(P::NS >= 2, "shared memory plan");
  constexpr int NS = P::NS, NSB = P::NSB, NRB = P::NRB, KC = P::KC;
  constexpr int NEW = 4 * (NC / 32);         // active elementwise warps: 4 lane quarters x (NC / 32) chain groups of 32
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sB = smem + P::OFF_B;
  unsigned char* sX = smem + P::OFF_X;
  unsigned char* sRb = smem + P::OFF_RB;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P::OFF_BAR);
  uint64_t* bar_b = bars;               // ΔB tiles landed
  uint64_t* x_full = bars + 1;          // [NS]
  uint64_t* x_empty = x_full + NS;      // [NS]  GEMM2 done with the stage
  uint64_t* s_full = x_empty + NS;      // [NSB] GEMM1 done
  uint64_t* s_empty = s_full + NSB;     // [NSB] S is in registers
  uint64_t* r_full = s_empty + NSB;     // [NRB] R written to shared memory
  uint64_t* r_empty = r_full + NRB;     // [NRB] GEMM2 done with the buffer
  uint64_t* g_full = r_empty + NRB;     // accumulator complete for its flush period
  uint64_t* g_empty = g_full + 1;       // accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + P::NBAR);
  uint32_t* far_flag = tmem_slot + 1;   // some chain of the tile is far from the reference: every hand-off of R has two halves (hi, lo)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, split = blockIdx.y;
  const int b0 = (int)(((long long)nblk_total * split) / nsplit);
  const int b1 = (int)(((long long)nblk_total * (split + 1)) / nsplit);
  const int nb = b1 - b0;
  const int fe = flush_every > 0 ? flush_every : 0x7fffffff;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) asm volatile("trap;");
    mbar_init(bar_b, 1);
    for (int i = 0; i < NS; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); }
    for (int i = 0; i < NSB; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], ARR_ALL ? 32 * NEW : NEW); }
    for (int i = 0; i < NRB; ++i) { mbar_init(&r_full[i], ARR_ALL ? 32 * NEW : NEW); mbar_init(&r_empty[i], 1); }
    mbar_init(g_full, 1);
    mbar_init(g_empty, ARR_ALL ? 32 * NEW : NEW);
    *far_flag = 0u;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == RM_G1_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_S = tmem;              // NSB x 128 columns
  const uint32_t tmem_G = tmem + NSB * 128;  // NC columns, lane = feature

  if (warp == RM_TMA_WARP) {
    // ===================================================== TMA producer
    if (lane == 0 && nb > 0) {
      mbar_expect_tx(bar_b, NT * P::B_BYTES);
      for (int kc = 0; kc < KC; ++kc) {
        tma_load_2d(&tmBh, sB + 0 * P::B_BYTES + kc * P::B_CHUNK, bar_b, kc * 64, tile * NC);
        if (NT > 1) tma_load_2d(&tmBm, sB + 1 * P::B_BYTES + kc * P::B_CHUNK, bar_b, kc * 64, tile * NC);
      }
      // L2 prefetch PF blocks ahead of the loads: a stage is refilled only after GEMM2 of its previous block, so the refill
      // itself must not wait on HBM
      constexpr int PF = NS + 3;
      for (int i = 0; i < PF && i < nb; ++i)
        for (int kc = 0; kc < KC; ++kc) tma_prefetch_2d(&tmX, kc * 64, (b0 + i) * ROWS);
      for (int i = 0; i < nb; ++i) {
        const int st = i % NS;
        const uint32_t ph = (uint32_t)(i / NS) & 1u;
        if (i + PF < nb)
          for (int kc = 0; kc < KC; ++kc) tma_prefetch_2d(&tmX, kc * 64, (b0 + i + PF) * ROWS);
        mbar_wait(&x_empty[st], ph ^ 1u);
        mbar_expect_tx(&x_full[st], P::X_BYTES);
        for (int kc = 0; kc < KC; ++kc)
          tma_load_2d(&tmX, sX + st * P::X_BYTES + kc * CHUNK_BYTES, &x_full[st], kc * 64, (b0 + i) * ROWS);
      }
    }
  } else if (warp == RM_G1_WARP) {
    // ===================================================== GEMM1 issuer: S = X̃blk · ΔBᵀ
    if (nb > 0) {
      constexpr uint32_t IDESC1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NC >> 3) << 17) | ((uint32_t)(ROWS >> 4) << 24);
      const uint32_t aX = smem_u32(sX);
      const uint64_t dKM = desc_kmajor(0, 0);
      const uint32_t km_hi = (uint32_t)(dKM >> 32), km_lo0 = (uint32_t)dKM;
      uint32_t bB[NT];
#pragma unroll
      for (int term = 0; term < NT; ++term) bB[term] = km_lo0 + ((smem_u32(sB) + (uint32_t)term * P::B_BYTES) >> 4);
      mbar_wait(bar_b, 0);
      for (int i = 0; i < nb; ++i) {
        const int st = i % NS, buf = i % NSB;
        mbar_wait(&x_full[st], (uint32_t)(i / NS) & 1u);
        if (i >= NSB) mbar_wait(&s_empty[buf], (uint32_t)(i / NSB - 1) & 1u);
        tc_fence_after();
        const uint32_t xlo = km_lo0 + ((aX + (uint32_t)st * P::X_BYTES) >> 4);
        const uint32_t d = tmem_S + (uint32_t)buf * 128u;
#pragma unroll
        for (int term = 0; term < NT; ++term) {
#pragma unroll
          for (int c = 0; c < (NK + 3) / 4; ++c) {
            constexpr int LAST = NK - ((NK + 3) / 4 - 1) * 4;   // K steps in the last chunk
            const uint32_t off = (uint32_t)(c * (CHUNK_BYTES >> 4)), offb = (uint32_t)(c * (P::B_CHUNK >> 4));
            const uint32_t acc = (term | c) ? 1u : 0u;
            if (RMDBG & 2) continue;
            if (c + 1 < (NK + 3) / 4) mma_ss_run<4>(d, xlo + off, km_hi, bB[term] + offb, km_hi, IDESC1, acc);
            else mma_ss_run<LAST>(d, xlo + off, km_hi, bB[term] + offb, km_hi, IDESC1, acc);
          }
        }
        if (elect_one()) tc_commit(&s_full[buf]);
        __syncwarp();
      }
    }
  } else if (warp == RM_G2_WARP) {
    // ===================================================== GEMM2 issuer: Gᵀ += X̃blkᵀ · R
    if (nb > 0) {
      constexpr uint32_t IDESC2 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(NC >> 3) << 17) |
                                  ((uint32_t)(128 >> 4) << 24);
      const uint32_t aX = smem_u32(sX), aR = smem_u32(sRb);
      const uint64_t dMN = desc_mnmajor(0, 0);
      const uint32_t mn_hi = (uint32_t)(dMN >> 32), mn_lo0 = (uint32_t)dMN;
      asm volatile("bar.sync 2, %0;" ::"r"(32 * (NEW + 1)) : "memory");   // the elementwise warps have measured the tile's chains
      const int nph = *reinterpret_cast<volatile uint32_t*>(far_flag) ? 2 : 1;
      int period = 0, in_period = 0, j = 0;
      for (int i = 0; i < nb; ++i) {
        const int st = i % NS;
        const uint32_t xm = mn_lo0 + ((aX + (uint32_t)st * P::X_BYTES) >> 4);
        for (int ph = 0; ph < nph; ++ph, ++j) {
          const int rb = j % NRB;
          mbar_wait(&r_full[rb], (uint32_t)(j / NRB) & 1u);
          if (ph == 0 && in_period == 0 && period >= 1) mbar_wait(g_empty, (uint32_t)(period - 1) & 1u);
          tc_fence_after();
          const uint32_t rm = mn_lo0 + ((aR + (uint32_t)rb * P::R_BYTES) >> 4);
          const uint32_t acc0 = (in_period > 0 || ph > 0) ? 1u : 0u;
          if (!(RMDBG & 4)) {
#pragma unroll
            for (int gh = 0; gh < P::NGH; ++gh) {   // features 128 gh .. 128 gh + 127: the 64-column chunks 2 gh, 2 gh + 1 of the X tile
              const uint32_t xa = xm + (uint32_t)gh * (2u * CHUNK_BYTES >> 4), dg = tmem_G + (uint32_t)(gh * NC);
              mma_ss_mn_run<4>(dg, xa, mn_hi, rm, mn_hi, IDESC2, acc0);
              mma_ss_mn_run<4>(dg, xa + 4u * 128u, mn_hi, rm + 4u * 128u, mn_hi, IDESC2, 1u);
            }
          }
          if (elect_one()) { if (ph + 1 == nph) tc_commit(&x_empty[st]); tc_commit(&r_empty[rb]); }
          __syncwarp();
        }
        ++in_period;
        if (i + 1 == nb || in_period == fe) {
          if (elect_one()) tc_commit(g_full);
          ++period;
          in_period = 0;
        }
        __syncwarp();
      }
    }
  } else if (warp < NEW) {
    // ===================================================== elementwise + epilogue
    const int h = warp >> 2;                   // chain group: chains 32h .. 32h+31 of the tile
    const int q = warp & 3;                    // TMEM lane quarter: rows 32q .. 32q+31 of the block (features in the epilogue)
    const int r = q * 32 + lane;               // row of the block / feature index
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    // far flags of this warp's 32 chains, from the staged operand itself: ‖β − β₀‖² over the high bf16 term (columns < D;
    // columns D..D+2 of the high term hold the 1.0 of the exact path's reference columns and are skipped)
    uint32_t far32 = 0;
    if (nb > 0) {
      mbar_wait(bar_b, 0);
      const int c = h * 32 + lane;             // chain of the tile (row of the ΔB tile)
      float n2 = 0.f;
      const uint32_t base = smem_u32(sB) + (uint32_t)c * 128u;
      for (int kc = 0; kc < KC; ++kc)
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          uint32_t w0, w1, w2, w3;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3)
                       : "r"(base + (uint32_t)kc * P::B_CHUNK + (uint32_t)((u ^ (c & 7)) << 4)));
          const uint32_t ws[4] = {w0, w1, w2, w3};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int col = kc * 64 + u * 8 + e * 2;
            const float lo = __uint_as_float(ws[e] << 16), hi = __uint_as_float(ws[e] & 0xffff0000u);
            if (col < D) n2 = fmaf(lo, lo, n2);
            if (col + 1 < D) n2 = fmaf(hi, hi, n2);
          }
        }
      const bool far = !(n2 <= kappa2) && (tile * NC + c) < nrows;   // NaN counts as far
      far32 = __ballot_sync(0xffffffffu, far);
      if (far32 != 0u && lane == 0) atomicOr(far_flag, 1u);
      asm volatile("bar.sync 2, %0;" ::"r"(32 * (NEW + 1)) : "memory");   // with the GEMM2 issuer
    }
    const bool tile_far = nb > 0 && *reinterpret_cast<volatile uint32_t*>(far_flag) != 0u;
    const bool live = (tile * NC + h * 32) < nrows;   // a chain group beyond the staged rows only keeps the barrier protocol going
    float2 lacc[16];                                   // Σ_rows μ for this thread's row slot, chains (2j, 2j+1) of the group
#pragma unroll
    for (int j = 0; j < 16; ++j) lacc[j] = make_float2(0.f, 0.f);
    const float4* recp = rec + (size_t)b0 * ROWS + r;
    // the records are read from global memory two blocks ahead (one block ahead, the first use of a record was the largest
    // single stall of the kernel: 14 % of the warp samples, ncu source page)
    float4 rc_next = nb > 0 ? __ldg(recp) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 rc_next2 = nb > 1 ? __ldg(recp + ROWS) : make_float4(0.f, 0.f, 0.f, 0.f);
    // this thread's 64 bytes of a row of R: chunk h / 2, 16-byte units 4 (h % 2) + qt (8 chains each), swizzled with the row
    const uint32_t r_row = smem_u32(sRb) + (uint32_t)(h >> 1) * CHUNK_BYTES + (uint32_t)r * 128u;
    const uint32_t r_sw = (uint32_t)(r & 7), u0 = (uint32_t)(h & 1) * 4u;
    // The loop over the row blocks exists twice: the Taylor form only (no chain of the group is far: the steady state), and
    // a copy whose flagged pairs take the closed forms; a warp picks one for the whole launch (the flags are warp-uniform).
    auto run_blocks = [&](auto FAR_TAG, double& lsum_out) {
      constexpr bool FARP = decltype(FAR_TAG)::value;
      int fpos = 0, fper = 0;
      [[maybe_unused]] double dacc = 0.0;   // far tiles: Σ μ of chain `lane` of the group over this warp's rows, folded every block
      for (int i = 0; i < nb; ++i) {
        const int buf = i % NSB;
        const int j0 = FARP ? 2 * i : i;                               // hand-off of R: one per block, two (hi, lo) in far tiles
        const int rb = j0 % NRB;
        const float4 rc = rc_next;                                   // (A2, A3, A4, eta0) of this thread's row
        rc_next = rc_next2;
        if (i + 2 < nb) rc_next2 = __ldg(recp + (size_t)(i + 2) * ROWS);
        mbar_wait(&s_full[buf], (uint32_t)(i / NSB) & 1u);
        tc_fence_after();
        if (j0 >= NRB) mbar_wait(&r_empty[rb], (uint32_t)(j0 / NRB - 1) & 1u);
        const uint32_t rbase = r_row + (uint32_t)rb * P::R_BYTES;
        [[maybe_unused]] uint32_t lo_pk[16];
        if (live) {
          // four quarters of 8 chains: S -> registers one quarter ahead, remainder, packed bf16 -> shared memory
          const uint32_t ts = tmem_S + (uint32_t)buf * 128u + lane_sel + (uint32_t)h * 32u;
          uint32_t v[2][8];
          tmem_ld8(ts, v[0]);
          const float C3s = rc.y * (-1.f / 12.f);
          [[maybe_unused]] float r0 = 0.f, w0 = 0.f, f0 = 0.f;
          if constexpr (FARP) {
            const float e0 = rc.w;
            const float t0 = ex2_approx(-fabsf(e0) * 1.4426950408889634f), rcp0 = rcp_approx(1.f + t0);
            r0 = e0 >= 0.f ? t0 * rcp0 : rcp0;                                               // σ(−η̃0)
            w0 = t0 * rcp0 * rcp0;                                                            // σ'(η̃0)
            f0 = fminf(e0, 0.f) - 0.6931471805599453f * lg2_approx(1.f + t0);                 // log σ(η̃0)
          }
#pragma unroll
          for (int qt = 0; qt < 4; ++qt) {
            tmem_ld_wait();
            if (qt + 1 < 4) tmem_ld8(ts + (uint32_t)(qt + 1) * 8u, v[(qt + 1) & 1]);
            else { tc_fence_before(); ew_arrive(&s_empty[buf], lane); }   // S is in registers: GEMM1 may overwrite the buffer
                                                                                                   // (one arrival per warp: 512 per-thread arrivals on one word serialise)
            const uint32_t* vv = v[qt & 1];
            uint32_t pk[4];
            if constexpr ((RMDBG & 1) != 0) {
#pragma unroll
              for (int j = 0; j < 4; ++j) pk[j] = vv[2 * j] ^ vv[2 * j + 1];
            } else if constexpr (!FARP) {
              const float2 A2 = make_float2(rc.x, rc.x), A3 = make_float2(rc.y, rc.y), A4 = make_float2(rc.z, rc.z);
              const float2 C3 = make_float2(C3s, C3s);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 d = make_float2(__uint_as_float(vv[2 * j]), __uint_as_float(vv[2 * j + 1]));
                const float2 d2 = __fmul2_rn(d, d);
                const float2 t = __ffma2_rn(__ffma2_rn(A4, d, A3), d, A2);
                const float2 rho = __fmul2_rn(d2, t);
                lacc[4 * qt + j] = __ffma2_rn(__fmul2_rn(d2, d2), C3, lacc[4 * qt + j]);   // μ = −A₃δ⁴/12 (+ O(δ⁵) < 1e-11 per row)
                pk[j] = pack_bf16(rho.x, rho.y);
              }
            } else {
              const uint32_t fm = far32 >> (8 * qt);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float dd[2] = {__uint_as_float(vv[2 * j]), __uint_as_float(vv[2 * j + 1])};
                float rho[2], lam[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  if ((fm >> (2 * j + e)) & 1u) {      // this chain is far (warp-uniform): closed forms
                    const float eta = rc.w + dd[e];
                    const float t = ex2_approx(-fabsf(eta) * 1.4426950408889634f), rcp1 = rcp_approx(1.f + t);
                    const float rr = eta >= 0.f ? t * rcp1 : rcp1;
                    rho[e] = fmaf(w0, dd[e], rr - r0);
                    const float f = fminf(eta, 0.f) - 0.6931471805599453f * lg2_approx(1.f + t);
                    lam[e] = fmaf(0.5f * w0 * dd[e], dd[e], fmaf(-r0, dd[e], f - f0));
                  } else {
                    const float d2 = dd[e] * dd[e];
                    rho[e] = d2 * fmaf(fmaf(rc.z, dd[e], rc.y), dd[e], rc.x);
                    lam[e] = fmaf(C3s * d2, d2, (1.f / 3.f) * dd[e] * rho[e]);   // λ = μ + δρ/3
                  }
                }
                // two bf16 terms of ρ in far tiles; μ = λ − δρ/3 with the ρ the tensor core will see (hi + lo), so that the
                // consumer's ⅓ δ·(X̃ᵀρ) cancels exactly
                pk[j] = pack_bf16(rho[0], rho[1]);
                const float h0 = __uint_as_float(pk[j] << 16), h1 = __uint_as_float(pk[j] & 0xffff0000u);
                lo_pk[4 * qt + j] = pack_bf16(rho[0] - h0, rho[1] - h1);
                const float s0 = h0 + __uint_as_float(lo_pk[4 * qt + j] << 16), s1 = h1 + __uint_as_float(lo_pk[4 * qt + j] & 0xffff0000u);
                lacc[4 * qt + j].x += fmaf(-(1.f / 3.f) * dd[0], s0, lam[0]);
                lacc[4 * qt + j].y += fmaf(-(1.f / 3.f) * dd[1], s1, lam[1]);
              }
            }
            if (!(RMDBG & 8)) sts128(rbase + (((u0 + (uint32_t)qt) ^ r_sw) << 4), pk[0], pk[1], pk[2], pk[3]);
            else if (pk[0] == 0x12345678u) sts128(rbase, pk[0], pk[1], pk[2], pk[3]);
          }
        } else {
          tc_fence_before();
          ew_arrive(&s_empty[buf], lane);
#pragma unroll
          for (int u = 0; u < 4; ++u) sts128(rbase + (((u0 + (uint32_t)u) ^ r_sw) << 4), 0u, 0u, 0u, 0u);
        }
        fence_proxy_async();
        ew_arrive(&r_full[rb], lane);
        if constexpr (FARP) {
          // second hand-off of the block: the low halves; then this block's μ folded into Float64 per chain
          const int j1 = j0 + 1, rb1 = j1 % NRB;
          if (j1 >= NRB) mbar_wait(&r_empty[rb1], (uint32_t)(j1 / NRB - 1) & 1u);
          const uint32_t rbase1 = r_row + (uint32_t)rb1 * P::R_BYTES;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (live) sts128(rbase1 + (((u0 + (uint32_t)u) ^ r_sw) << 4), lo_pk[4 * u], lo_pk[4 * u + 1], lo_pk[4 * u + 2], lo_pk[4 * u + 3]);
            else sts128(rbase1 + (((u0 + (uint32_t)u) ^ r_sw) << 4), 0u, 0u, 0u, 0u);
          }
          fence_proxy_async();
          ew_arrive(&r_full[rb1], lane);
          dacc += (double)lane_sums(lacc, lane);
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) lacc[jj] = make_float2(0.f, 0.f);
        }
        const bool closes = (i + 1 == nb) || (fpos == fe - 1);
        const int period = fper;
        if (++fpos == fe) { fpos = 0; ++fper; }
        if (closes) {
          // drain the GEMM2 accumulator (lane = feature, column = chain) and add it outside the tensor core
          mbar_wait(g_full, (uint32_t)period & 1u);
          tc_fence_after();
          if (live) {
#pragma unroll 1
            for (int gq = 0; gq < 4 * P::NGH; ++gq) {
              const int gh = gq >> 2, qt = gq & 3, feat = gh * 128 + r;
              uint32_t w[8];
              tmem_ld8(tmem_G + lane_sel + (uint32_t)(gh * NC) + (uint32_t)h * 32u + (uint32_t)qt * 8u, w);
              tmem_ld_wait();
              if (feat < Dp) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const int row = tile * NC + h * 32 + qt * 8 + j;
                  if (row < nrows) {
                    float* gp = G + ((size_t)split * nrows + (size_t)row) * Dp + feat;   // lanes = consecutive features: one 128-byte line per warp
                    if (RMDBG) { *gp = 0.f; continue; }
                    if (period > 0) atomicAdd(gp, __uint_as_float(w[j]));                 // only this thread ever touches the address: same bits as load + add
                    else *gp = __uint_as_float(w[j]);
                  }
                }
              }
            }
          }
          tc_fence_before();
          ew_arrive(g_empty, lane);
        }
      }
      if constexpr (FARP) lsum_out = dacc; else lsum_out = (double)lane_sums(lacc, lane);
    };
    double lsum;                                        // Σ μ of chain `lane` of the group over this warp's row slots
    if (!tile_far) { run_blocks(std::false_type{}, lsum); } else { run_blocks(std::true_type{}, lsum); }
    // per-chain log-density remainder: the four lane quarters of a chain group add up in a fixed order (X stages are dead by now)
    double* lp = reinterpret_cast<double*>(sX);        // [4 quarters][NC]
    asm volatile("bar.sync 1, %0;" ::"r"(32 * NEW) : "memory");
    lp[q * NC + h * 32 + lane] = lsum;
    asm volatile("bar.sync 1, %0;" ::"r"(32 * NEW) : "memory");
    if (q == 0) {
      const int c = h * 32 + lane, row = tile * NC + c;
      if (row < nrows) Ld[(size_t)split * nrows + row] = RMDBG ? 0.0 : ((lp[c] + lp[NC + c]) + lp[2 * NC + c]) + lp[3 * NC + c];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == RM_G1_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

template <int DT, int NK, int NC, int NT> void launch_rm_nc(LogisticTC& tc, cudaStream_t s, int nrows, int nsplit) {
  using P = RmPlan<DT, NC, NT>;
  static unsigned long long attr_done = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (!((attr_done >> (dev & 63)) & 1ull)) {
    tc.last = cudaFuncSetAttribute(k_logistic_rm<DT, NK, NC, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, P::TOTAL);
    attr_done |= 1ull << (dev & 63);
  }
  const int tiles = (nrows + NC - 1) / NC;
  dim3 grid(tiles, nsplit);
  CUtensorMap m[3];
  std::memcpy(&m[0], tc.tmaps[0], sizeof(CUtensorMap));
  std::memcpy(&m[1], tc.tmaps[NC > 64 ? 1 : 5], sizeof(CUtensorMap));   // ΔB high / middle terms: 128- or 64-row boxes
  std::memcpy(&m[2], tc.tmaps[NC > 64 ? 2 : 6], sizeof(CUtensorMap));
  k_logistic_rm<DT, NK, NC, NT><<<grid, RM_THREADS, P::TOTAL, s>>>(m[0], m[1], m[2], reinterpret_cast<const float4*>(tc.rec), tc.G, tc.Ld, nrows,
                                                               tc.D, tc.Dp, (int)(tc.Npad / ROWS), nsplit, tc.flush_every, tc.kappa2);
}
template <int DT, int NK> void launch_rm_nk(LogisticTC& tc, cudaStream_t s, int nrows, int nsplit, int nc) {
  if constexpr (DT > 128) {
    launch_rm_nc<DT, NK, 64, 2>(tc, s, nrows, nsplit);   // two ΔB terms of 256 columns leave room for 64-chain tiles only
  } else {
    if (nc == 64) launch_rm_nc<DT, NK, 64, 2>(tc, s, nrows, nsplit);
    else launch_rm_nc<DT, NK, 128, 2>(tc, s, nrows, nsplit);
  }
}

// ------------------------------------------------------------------ set-up of the reference constants
constexpr int RMG_BLOCKS = 1024;
// one thread per data row: η̃0 = X̃_i·β₀ in Float64; record (A₂, A₃, A₄, η̃0) in fp32; r0 (for g0), w (for H0), log σ(η̃0) (for ℓ0)
__global__ void k_rm_records(const uint16_t* __restrict__ Xb, const float* __restrict__ beta_ref, float4* rec, float* r0o, float* wo,
                             double* f0o, long long N, long long Npad, int D, int Dt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Npad) return;
  float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
  float r0 = 0.f, w = 0.f;
  double f0 = 0.0;
  if (i < N) {
    const uint16_t* xr = Xb + i * Dt;
    double eta = 0.0;
    for (int d = 0; d < D; ++d) eta = fma((double)bf16_val(xr[d]), (double)beta_ref[d], eta);
    const double th = tanh(0.5 * eta);                 // u = 2 r0 − 1 = −tanh(η̃0 / 2)
    const double u = -th, wd = 0.25 * (1.0 - th * th);
    r0 = (float)(0.5 * (1.0 + u));
    w = (float)wd;
    f0 = (eta < 0.0 ? eta : 0.0) - log1p(exp(-fabs(eta)));
    o.x = (float)(-0.5 * wd * u);
    o.y = (float)(-wd * (u * u - 2.0 * wd) / 6.0);
    o.z = (float)(-wd * u * (u * u - 8.0 * wd) / 24.0);
    o.w = (float)eta;
  }
  rec[i] = o; r0o[i] = r0; wo[i] = w; f0o[i] = f0;
}
// weighted column sums Σ_i X̃_id v_i (v = r0) and the plain sum Σ_i f0_i, fixed order: block b sums its contiguous range
__global__ void k_rm_g0_partial(const uint16_t* __restrict__ Xb, const float* __restrict__ v, const double* __restrict__ f0, double* part,
                                long long N, int D, int Dt, int Dp) {
  const long long r0 = N * blockIdx.x / gridDim.x, r1 = N * (blockIdx.x + 1) / gridDim.x;
  for (int d = threadIdx.x; d <= Dp; d += blockDim.x) {
    double acc = 0.0;
    if (d < D) for (long long i = r0; i < r1; ++i) acc = fma((double)bf16_val(Xb[i * Dt + d]), (double)v[i], acc);
    else if (d == Dp) for (long long i = r0; i < r1; ++i) acc += f0[i];
    part[(size_t)blockIdx.x * (Dp + 1) + d] = acc;
  }
}
__global__ void k_rm_g0_sum(const double* __restrict__ part, double* grad0, double* ell0, int nb, int Dp) {
  for (int d = threadIdx.x; d <= Dp; d += blockDim.x) {
    double acc = 0.0;
    for (int b = 0; b < nb; ++b) acc += part[(size_t)b * (Dp + 1) + d];
    if (d < Dp) grad0[d] = acc; else *ell0 = acc;
  }
}
constexpr int H0_BLOCKS = 148, H0_THREADS = 512, H0_ACC = 32;   // a 128 x 128 block of the matrix = 512 x 32 accumulators per thread block
// H = X̃ᵀ diag(w) X̃ (w == nullptr: X̃ᵀX̃) in Float64 from the stored fp32 w, one 128 x 128 block (a0, b0) of it per launch:
// each thread block sums a contiguous range of rows into register accumulators (pair p = a 128 + b -> thread p % 512, slot
// p / 512), then one pass adds the blocks in order
__global__ void __launch_bounds__(H0_THREADS) k_rm_hess_partial(const uint16_t* __restrict__ Xb, const float* __restrict__ w, double* part,
                                                                long long N, int D, int Dt, int a0, int b0) {
  __shared__ float xa[8][128];
  __shared__ float xb[8][128];
  __shared__ float ws[8];
  const long long r0 = N * blockIdx.x / gridDim.x, r1 = N * (blockIdx.x + 1) / gridDim.x;
  double acc[H0_ACC];
  int pa[H0_ACC], pb[H0_ACC];
#pragma unroll
  for (int m = 0; m < H0_ACC; ++m) {
    acc[m] = 0.0;
    const int p = threadIdx.x + m * H0_THREADS;
    pa[m] = p >> 7; pb[m] = p & 127;
  }
  for (long long it = r0; it < r1; it += 8) {
    const int nr = (int)((r1 - it < 8) ? (r1 - it) : 8);
    __syncthreads();
    for (int idx = threadIdx.x; idx < nr * 128; idx += H0_THREADS) {
      const int rr = idx >> 7, d = idx & 127;
      xa[rr][d] = a0 + d < D ? bf16_val(Xb[(it + rr) * Dt + a0 + d]) : 0.f;
      xb[rr][d] = b0 + d < D ? bf16_val(Xb[(it + rr) * Dt + b0 + d]) : 0.f;
    }
    if (threadIdx.x < nr) ws[threadIdx.x] = w ? w[it + threadIdx.x] : 1.f;
    __syncthreads();
    for (int rr = 0; rr < nr; ++rr) {
      const double wv = (double)ws[rr];
#pragma unroll
      for (int m = 0; m < H0_ACC; ++m) acc[m] = fma((double)(xa[rr][pa[m]] * xb[rr][pb[m]]), wv, acc[m]);   // bf16 x bf16 is exact in fp32
    }
  }
#pragma unroll
  for (int m = 0; m < H0_ACC; ++m)
    if (a0 + pa[m] < D && b0 + pb[m] < D) part[((size_t)blockIdx.x * D + (a0 + pa[m])) * D + (b0 + pb[m])] = acc[m];
}
__global__ void k_rm_hess_sum(const double* __restrict__ part, float* H, int nb, int D, int Dp) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= D * D) return;
  double acc = 0.0;
  for (int b = 0; b < nb; ++b) acc += part[(size_t)b * D * D + p];
  H[(size_t)(p / D) * Dp + (p % D)] = (float)acc;
}

}  // namespace

void logistic_rm_launch(LogisticTC& tc, cudaStream_t s, int nrows, int nsplit, int nc) {
  switch (tc.dk / 16) {
    case 1: launch_rm_nk<64, 1>(tc, s, nrows, nsplit, nc); break;
    case 2: launch_rm_nk<64, 2>(tc, s, nrows, nsplit, nc); break;
    case 3: launch_rm_nk<64, 3>(tc, s, nrows, nsplit, nc); break;
    case 4: launch_rm_nk<64, 4>(tc, s, nrows, nsplit, nc); break;
    case 5: launch_rm_nk<128, 5>(tc, s, nrows, nsplit, nc); break;
    case 6: launch_rm_nk<128, 6>(tc, s, nrows, nsplit, nc); break;
    case 7: launch_rm_nk<128, 7>(tc, s, nrows, nsplit, nc); break;
    case 8: launch_rm_nk<128, 8>(tc, s, nrows, nsplit, nc); break;
    case 9: launch_rm_nk<192, 9>(tc, s, nrows, nsplit, nc); break;
    case 10: launch_rm_nk<192, 10>(tc, s, nrows, nsplit, nc); break;
    case 11: launch_rm_nk<192, 11>(tc, s, nrows, nsplit, nc); break;
    case 12: launch_rm_nk<192, 12>(tc, s, nrows, nsplit, nc); break;
    case 13: launch_rm_nk<256, 13>(tc, s, nrows, nsplit, nc); break;
    case 14: launch_rm_nk<256, 14>(tc, s, nrows, nsplit, nc); break;
    case 15: launch_rm_nk<256, 15>(tc, s, nrows, nsplit, nc); break;
    default: launch_rm_nk<256, 16>(tc, s, nrows, nsplit, nc); break;
  }
}

// per-row records, g0 = X̃ᵀ r0, ℓ0 = Σ log σ(η̃0), H0 = X̃ᵀ diag(w) X̃ and the radius κ of the Taylor path.
// κ: the row-wise rms of δ = x̃·(β − β₀) is at most sqrt(λ_max(X̃ᵀX̃ / N))·‖β − β₀‖; the degree-4 remainder is good to ~1e-2·15·s⁴
// of the gradient at rms s, so chains with a bound above 0.08 take the closed forms instead.
int32_t logistic_rm_write_reference(LogisticTC& tc, cudaStream_t s, const float* beta_ref_dev, RmGroupSum group_sum, void* group_ctx,
                                    int world, std::string& err) {
  const size_t np = size_t(tc.Npad);
  auto need = [&](void** p, size_t bytes) { return *p || cudaMalloc(p, bytes) == cudaSuccess; };
  if (!need((void**)&tc.rec, np * 16) || !need((void**)&tc.rm_r0, np * 4) || !need((void**)&tc.rm_w, np * 4) || !need((void**)&tc.rm_f0, np * 8) ||
      !need((void**)&tc.H0, size_t(tc.Dp) * tc.Dp * 4) || !need((void**)&tc.H0_part, size_t(H0_BLOCKS) * tc.D * tc.D * 8) ||
      !need((void**)&tc.rm_part, size_t(RMG_BLOCKS) * (tc.Dp + 1) * 8) || !need((void**)&tc.ell0, 8)) {
    err = "device allocation failed (remainder-mode reference)";
    return BNUTS_ERR_CUDA;
  }
  k_rm_records<<<(unsigned)((np + 255) / 256), 256, 0, s>>>(tc.Xb, beta_ref_dev, reinterpret_cast<float4*>(tc.rec), tc.rm_r0, tc.rm_w, tc.rm_f0,
                                                           (long long)tc.N, (long long)tc.Npad, tc.D, tc.Dt);
  k_rm_g0_partial<<<RMG_BLOCKS, 160, 0, s>>>(tc.Xb, tc.rm_r0, tc.rm_f0, tc.rm_part, (long long)tc.N, tc.D, tc.Dt, tc.Dp);
  k_rm_g0_sum<<<1, 160, 0, s>>>(tc.rm_part, tc.grad0, tc.ell0, RMG_BLOCKS, tc.Dp);
  // λ_max(X̃ᵀX̃ / N) by power iteration on the host (D x D), then H0 into the same buffer
  cudaMemsetAsync(tc.H0, 0, size_t(tc.Dp) * tc.Dp * 4, s);
  for (int a0 = 0; a0 < tc.D; a0 += 128)
    for (int b0 = 0; b0 < tc.D; b0 += 128)
      k_rm_hess_partial<<<H0_BLOCKS, H0_THREADS, 0, s>>>(tc.Xb, nullptr, tc.H0_part, (long long)tc.N, tc.D, tc.Dt, a0, b0);
  k_rm_hess_sum<<<(tc.D * tc.D + 255) / 256, 256, 0, s>>>(tc.H0_part, tc.H0, H0_BLOCKS, tc.D, tc.Dp);
  // rows sharded over a group of engines: every constant is a sum over ALL rows (g0 and X̃ᵀX̃ here, ℓ0 and H0 below)
  if (group_sum) {
    int32_t rc = group_sum(group_ctx, tc.H0, (int64_t)tc.Dp * tc.Dp, tc.grad0, tc.Dp);
    if (!rc) rc = group_sum(group_ctx, nullptr, 0, tc.ell0, 1);
    if (rc) { err = "remainder-mode reference: the sum over the row group failed"; return rc; }
  }
  std::vector<float> g2(size_t(tc.Dp) * tc.Dp);
  if (cudaMemcpyAsync(g2.data(), tc.H0, g2.size() * 4, cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess) {
    err = "remainder-mode reference: device failure";
    return BNUTS_ERR_CUDA;
  }
  {
    const int D = tc.D;
    std::vector<double> v(size_t(D), 1.0 / std::sqrt(double(D))), y(static_cast<size_t>(D));
    double lam = 0.0;
    for (int it = 0; it < 60; ++it) {
      double n2 = 0.0;
      for (int a = 0; a < D; ++a) {
        double acc = 0.0;
        for (int b = 0; b < D; ++b) acc += double(g2[size_t(a) * tc.Dp + b]) * v[size_t(b)];
        y[size_t(a)] = acc; n2 += acc * acc;
      }
      lam = std::sqrt(n2);
      if (!(lam > 0.0)) break;
      for (int a = 0; a < D; ++a) v[size_t(a)] = y[size_t(a)] / lam;
    }
    lam = 1.1 * lam / (double(tc.N) * double(world));     // power iteration converges from below; equal shards
    const char* ke = std::getenv("BNUTS_TC_RM_RADIUS");   // rms of δ beyond which a chain takes the closed forms (default 0.08)
    const double s_max = ke ? std::atof(ke) : 0.08;
    tc.kappa2 = lam > 0.0 ? float(s_max * s_max / lam) : 0.f;
  }
  cudaMemsetAsync(tc.H0, 0, size_t(tc.Dp) * tc.Dp * 4, s);
  for (int a0 = 0; a0 < tc.D; a0 += 128)
    for (int b0 = 0; b0 < tc.D; b0 += 128)
      k_rm_hess_partial<<<H0_BLOCKS, H0_THREADS, 0, s>>>(tc.Xb, tc.rm_w, tc.H0_part, (long long)tc.N, tc.D, tc.Dt, a0, b0);
  k_rm_hess_sum<<<(tc.D * tc.D + 255) / 256, 256, 0, s>>>(tc.H0_part, tc.H0, H0_BLOCKS, tc.D, tc.Dp);
  if (group_sum) {
    const int32_t rc = group_sum(group_ctx, tc.H0, (int64_t)tc.Dp * tc.Dp, nullptr, 0);
    if (rc) { err = "remainder-mode reference: the sum over the row group failed"; return rc; }
  }
  if (cudaMemcpyAsync(&tc.ell0_host, tc.ell0, 8, cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess) {
    err = "remainder-mode reference: device failure";
    return BNUTS_ERR_CUDA;
  }
  return 0;
}

}  // namespace bn

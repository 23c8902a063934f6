// logistic_tc.cu — fused two-GEMM logistic-regression gradient on tcgen05 + TMA (sm_100a).
//
// Replaces the model call of the reference's leapfrog (logdensity_and_gradient!,
// call site src/kinetic_energy.jl:73) for the Bayesian logistic-regression target
// when thousands of chains advance in lockstep:
//     H = X·B            (N x C, never materialised)
//     R = y − σ(H),  ℓ_c = Σ_i [y_i H_ic − softplus(H_ic)]
//     G = Xᵀ·R           (D x C)
// evaluated in the sign-folded form: with X̃_i = (2y_i − 1)·X_i (a sign flip of bf16 rows,
// done once at set-up) and H̃ = X̃·B,
//     ℓ_c = Σ_i log σ(H̃_ic),   G = X̃ᵀ·σ(−H̃)
// so the kernel never touches y.  Σ_i min(H̃,0) is taken as ½(Σ H̃ − Σ|H̃|); the linear part
// Σ_i H̃_ic = colsum(X̃)·β_c is added by the consumer (backend.h model_grad) in Float64.
// One CTA owns a tile of 128 chains and a contiguous range of 128-row blocks of X.
// Chains are the MMA M dimension, so TMEM lane = chain: each elementwise thread
// owns one chain, the sum over data rows is a serial per-thread accumulation and
// the residual tile goes back to TMEM as the A operand of the second GEMM
// (FlashAttention-shaped: S in TMEM -> elementwise -> P in TMEM -> second MMA).
//
//   GEMM1  S[128 chains x 128 rows]  = Bt[128 x Dt] · Xblk[128 rows x Dt]ᵀ   (A, B from smem, K-major)
//   GEMM2  Gt[128 chains x Dt]      += R[128 x 128 rows] · Xblk[128 rows x Dt] (A from TMEM, B = same smem
//                                                                               tile read MN-major)
// fp32 accuracy on bf16 tensor cores: X is exact in bf16 (checked at set-up); the
// fp32 position is split exactly in three bf16 terms (β = βh + βm + βl, 3 x 8
// mantissa bits).  With a reference point β₀ near the mode (bnuts_logistic_set_reference)
// the kernel evaluates H̃ = H̃₀ + X̃·(β − β₀): H̃₀ = X̃·β₀ is stored, split in three bf16
// terms, in three spare K columns of X̃ (the matching β columns hold 1), and β − β₀ is small
// enough that two bf16 terms carry it (nterms = 2).  The residual is split in two
// (r = rh + rl), all accumulated in fp32; the
// TMEM accumulator of GEMM2 is drained every `flush_every` row blocks and summed
// outside the tensor core (bounds accumulator rounding drift).  Three S/R buffers in
// TMEM give the elementwise warps two block-times of slack behind the tensor pipe.
//
// Warp roles (608 threads): warp 0 TMA producer, warp 1 GEMM1 issuer + TMEM owner, warp 18 GEMM2 issuer,
// warps 2-17 elementwise/epilogue: TMEM lane group = warp % 4 (hardware rule); two groups of eight warps
// ping-pong over the row blocks, a warp owns two 32-row chunks of its group's blocks.
// A second kernel in this file, k_logistic_tc256, covers 128 < D <= 256 (config 5's shape); then the set-up kernels of the
// reference-point modes (k_write_reference / _eta0 / _c0, g0 reduction in a fixed order).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>

#include "logistic_tc.h"

namespace bn {

namespace {

constexpr int TC_THREADS = 608;       // warp 0 TMA, warp 1 GEMM1 issuer, warps 2-17 elementwise, warp 18 GEMM2 issuer
constexpr int G2_WARP = 18;
constexpr int ROWS = 128;            // data rows per block (GEMM1 N, GEMM2 K)
constexpr int CHAINS = 128;          // chains per CTA (MMA M)
#include "tc_ptx.h"

// 1/d for d in (1,2], two lanes at a time.  SW = 0: MUFU.RCP.  SW = 1: FMA-pipe only (quadratic
// minimax seed 32/99 d^2 - 144/99 d + 210/99, relative error 1.01e-2, then y <- y + y(e + e^2 + e^3),
// e = 1 - d y: 1.3e-7 after rounding),
// used for a fraction of the elements to take load off the MUFU pipe.
__device__ __forceinline__ float2 rcp2(float2 d, bool sw) {
  if (!sw) {
    return make_float2(rcp_approx(d.x), rcp_approx(d.y));
  } else {
    const float2 c2 = make_float2(0.32323232f, 0.32323232f), c1 = make_float2(-1.45454545f, -1.45454545f),
                 c0 = make_float2(2.12121212f, 2.12121212f), one = make_float2(1.f, 1.f);
    const float2 md = make_float2(-d.x, -d.y);
    float2 y = __ffma2_rn(__ffma2_rn(c2, d, c1), d, c0);
    const float2 e = __ffma2_rn(md, y, one);
    const float2 p = __ffma2_rn(__ffma2_rn(e, e, e), e, e);   // e + e^2 + e^3
    y = __ffma2_rn(y, p, y);
    return y;
  }
}
#ifndef BNUTS_TC_RCPSW
#define BNUTS_TC_RCPSW 0x55   // bit k set: pair k of every 8 pairs uses the FMA-pipe reciprocal
#endif
constexpr int RCPSW = BNUTS_TC_RCPSW;
#ifndef BNUTS_TC_PREFETCH
#define BNUTS_TC_PREFETCH 0   // measured: no gain (2.98-3.07 ms without, 3.01 ms with), the 4 X stages already cover the HBM latency
#endif
#ifndef BNUTS_TC_DEBUG
#define BNUTS_TC_DEBUG 0      // timing experiments only (results are wrong): 1 skip elementwise, 2 skip GEMM1, 4 skip GEMM2
#endif
constexpr int TCDBG = BNUTS_TC_DEBUG;
#ifndef BNUTS_TC_GROUPS
#define BNUTS_TC_GROUPS 2     // elementwise warp groups: 1 = all 16 warps share every row block (one 32-column chunk
#endif                        // each, lowest latency per block); 2 = two groups of 8 ping-pong over the blocks
constexpr int NG = BNUTS_TC_GROUPS;
constexpr int EW_PER_GROUP = 16 / NG;     // warps per group
constexpr int CPW = NG;                   // 32-column chunks per warp and block
// timing trace (debug builds only, -DBNUTS_TC_TRACE): CTA (0,0) records clock64() at fixed points of
// the first 256 row blocks: trace[role][block][slot], role 0 = TMA producer, 1 = MMA issuer, 2 = elementwise warp 2
#ifdef BNUTS_TC_TRACE
__device__ long long g_tc_trace[3 * 256 * 16];
#define TC_TRACE(role, blk, slot)                                                                   \
  do {                                                                                              \
    if (blockIdx.x == 0 && blockIdx.y == 0 && (blk) < 256 && (threadIdx.x & 31) == 0)               \
      g_tc_trace[((role) * 256 + (blk)) * 16 + (slot)] = clock64();                                  \
  } while (0)
#else
#define TC_TRACE(role, blk, slot) do {} while (0)
#endif

template <int DT> struct SmemPlan {
  static constexpr int KC = DT / 64;
  static constexpr int B_BYTES = KC * CHUNK_BYTES;   // one β term
  static constexpr int X_BYTES = KC * CHUNK_BYTES;   // one X stage
  static constexpr int NS = (DT == 128) ? 4 : 6;     // X stages
  static constexpr int NSB = 3;                      // S/R buffers in TMEM
  static constexpr int OFF_B = 0;                    // 3 terms
  static constexpr int OFF_X = 3 * B_BYTES;
  static constexpr int OFF_R = OFF_X + NS * X_BYTES;   // per-row offsets of the single-term residual mode: NS x 128 floats
  static constexpr int OFF_BAR = OFF_R + NS * 128 * 4;
  static constexpr int NBAR = 1 + 2 * NS + 3 * NSB + 2;
  static constexpr int TOTAL = OFF_BAR + NBAR * 8 + 16;
};

// NK = round16(D)/16: K steps of GEMM1; dk = 16 NK is also N of GEMM2 (columns >= dk of the
// 64-wide smem chunks are never read).  NK is a template parameter so the single MMA-issuing
// thread runs fully unrolled code with constant descriptor offsets (it is latency-critical).
//
// RR (single-term residual, only with a reference point): the residual is carried about the reference,
// r_i = r0_i + δ_i with r0_i = σ(−η̃0_i); X̃ᵀ·r0 is a constant vector the consumer adds in Float64 (EngineMem::grad0)
// and only δ goes through GEMM2, as ONE bf16 term: its rounding error, 2⁻⁹·|δ|, is relative to δ, not to r, and
// averages over the rows, so the gradient error is ≈ 1.7e-3·√(D/N)·|∇ℓ| at any distance from the reference
// (numerical experiment in DESIGN.md §6).  That is below the fp32 tolerance of the path only for tall problems:
// the host enables the mode when N ≥ 3.3e5·D (config 5: N = 1e8, D = 256), not for config 3 (N/D = 1e4: 1.7e-5).
// An fp16 δ would be 8× finer (enough for config 3), but tcgen05.mma kind::f16 with an fp16 A operand next to the
// bf16 B operand X̃ is an illegal instruction on sm_100a (measured), and X̃ is not exact in fp16.
// The kernel reads c_i = ½ − r0_i (fp32, one
// 512-byte bulk copy per X̃ stage) and forms δ = c_i − copysign(1/d − ½, η̃) in the FMA that produced r before:
// no extra arithmetic, no hi/lo split, half the TMEM stores and half the GEMM2 MMAs.
//
// (A quadratic-remainder variant of this idea — ρ = δ + w (η̃ − η̃0) through GEMM2, the linear part as a D x D mat-vec per
// chain — and two pipeline variants, 64-row blocks with β in TMEM and decoupled S / R buffers, were built, validated and
// measured in round 1: none was faster.  They were removed in round 2; DESIGN.md section 6 and profiles/r1_* keep the numbers.)
template <int DT, int NK, int RR>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_logistic_tc(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmBh,
              const __grid_constant__ CUtensorMap tmBm, const __grid_constant__ CUtensorMap tmBl,
              const float* __restrict__ c0, float* G, double* Ld, int nrows, int Dp, long long N, int nblk_total, int nsplit,
              int flush_every, int nterms) {
  using P = SmemPlan<DT>;
  constexpr int dk = NK * 16;
  constexpr int NS = P::NS;
  constexpr int KC = P::KC;
  constexpr int NSB = P::NSB;
  extern __shared__ __align__(1024) unsigned char smem[];   // SW128 tiles need 1024 B alignment (checked below)
  unsigned char* sB = smem + P::OFF_B;
  unsigned char* sX = smem + P::OFF_X;
  float* sR = reinterpret_cast<float*>(smem + P::OFF_R);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P::OFF_BAR);
  uint64_t* bar_b = bars;               // β tiles landed
  uint64_t* x_full = bars + 1;          // [NS]
  uint64_t* x_empty = x_full + NS;      // [NS]
  uint64_t* s_full = x_empty + NS;      // [NSB] GEMM1 done
  uint64_t* r_full = s_full + NSB;      // [NSB] residual written to TMEM
  uint64_t* sr_empty = r_full + NSB;    // [NSB] GEMM2 done with the buffer (it may be overwritten by GEMM1)
  uint64_t* g_full = sr_empty + NSB;    // accumulator complete for its flush period
  uint64_t* g_empty = g_full + 1;       // accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + P::NBAR);

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
    for (int i = 0; i < NSB; ++i) { mbar_init(&s_full[i], 1); mbar_init(&r_full[i], 32 * EW_PER_GROUP); mbar_init(&sr_empty[i], 1); }
    mbar_init(g_full, 1);
    mbar_init(g_empty, 32 * EW_PER_GROUP);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_S = tmem;              // NSB x 128 columns (S, overwritten in place by R)
  const uint32_t tmem_G = tmem + NSB * 128;  // dk <= 128 columns

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0 && nb > 0) {
      mbar_expect_tx(bar_b, 3 * P::B_BYTES);
      for (int kc = 0; kc < KC; ++kc) {
        tma_load_2d(&tmBh, sB + 0 * P::B_BYTES + kc * CHUNK_BYTES, bar_b, kc * 64, tile * CHAINS);
        tma_load_2d(&tmBm, sB + 1 * P::B_BYTES + kc * CHUNK_BYTES, bar_b, kc * 64, tile * CHAINS);
        tma_load_2d(&tmBl, sB + 2 * P::B_BYTES + kc * CHUNK_BYTES, bar_b, kc * 64, tile * CHAINS);
      }
      constexpr int PF = BNUTS_TC_PREFETCH;   // L2 prefetch distance in row blocks (0: none)
      for (int i = 0; i < PF && i < nb; ++i)
        for (int kc = 0; kc < KC; ++kc) tma_prefetch_2d(&tmX, kc * 64, (b0 + i) * ROWS);
      for (int i = 0; i < nb; ++i) {
        const int st = i % NS;
        const uint32_t ph = (uint32_t)(i / NS) & 1u;
        if (PF > 0 && i + PF < nb)
          for (int kc = 0; kc < KC; ++kc) tma_prefetch_2d(&tmX, kc * 64, (b0 + i + PF) * ROWS);
        TC_TRACE(0, i, 0);
        mbar_wait(&x_empty[st], ph ^ 1u);
        TC_TRACE(0, i, 1);
        mbar_expect_tx(&x_full[st], P::X_BYTES + (RR ? ROWS * 4 : 0));
        for (int kc = 0; kc < KC; ++kc)
          tma_load_2d(&tmX, sX + st * P::X_BYTES + kc * CHUNK_BYTES, &x_full[st], kc * 64, (b0 + i) * ROWS);
        if (RR) bulk_load_1d(sR + st * ROWS, c0 + (size_t)(b0 + i) * ROWS, ROWS * 4, &x_full[st]);
      }
    }
  } else if (warp == 1) {
    // ===================================================== GEMM1 issuer
    // Two warps issue the MMAs: this one GEMM1 (S = B·X̃ᵀ), warp G2_WARP GEMM2 (G += R·X̃).  Measured with one
    // issuer: its own bookkeeping (barrier round trips, commits, descriptor set-up, ~1 100 clk per row block on
    // top of ~1 800 clk of queue-throttled MMA issue) was as long as the elementwise stage, so small launches
    // (one chain tile, little elementwise work) still ran at 1.9 us per block.  With two issuers each one's
    // waits overlap the other's MMAs.  tcgen05.mma of DIFFERENT threads are not ordered by the hardware, so
    // the re-use of an S/R buffer is ordered explicitly: GEMM2(i) commits sr_empty, GEMM1(i + NSB) waits for it.
    // The whole warp runs convergently so descriptors stay in uniform registers; one elected lane issues.
    if (nb > 0) {
      constexpr uint32_t IDESC1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(ROWS >> 3) << 17) | ((uint32_t)(CHAINS >> 4) << 24);
      const uint32_t aX = smem_u32(sX);
      // descriptor halves: the high words are constants, the low words carry the start address
      const uint64_t dKM = desc_kmajor(0, 0);
      const uint32_t km_hi = (uint32_t)(dKM >> 32), km_lo0 = (uint32_t)dKM;
      uint32_t bB[3];
#pragma unroll
      for (int term = 0; term < 3; ++term) bB[term] = km_lo0 + ((smem_u32(sB) + (uint32_t)term * P::B_BYTES) >> 4);
      mbar_wait(bar_b, 0);
      for (int i = 0; i < nb; ++i) {
        const int st = i % NS, buf = i % NSB;
        TC_TRACE(1, i, 0);
        mbar_wait(&x_full[st], (uint32_t)(i / NS) & 1u);
        if (i >= NSB) mbar_wait(&sr_empty[buf], (uint32_t)(i / NSB - 1) & 1u);
        TC_TRACE(1, i, 1);
        tc_fence_after();
        const uint32_t xlo = km_lo0 + ((aX + (uint32_t)st * P::X_BYTES) >> 4);
        const uint32_t d = tmem_S + (uint32_t)buf * 128u;
#pragma unroll
        for (int term = 0; term < 3; ++term)
          if (term < nterms) {
#pragma unroll
            for (int c = 0; c < (NK + 3) / 4; ++c) {
              constexpr int LAST = NK - ((NK + 3) / 4 - 1) * 4;   // K steps in the last chunk
              const uint32_t off = (uint32_t)(c * (CHUNK_BYTES >> 4));
              const uint32_t acc = (term | c) ? 1u : 0u;
              if (TCDBG & 2) continue;
              if (c + 1 < (NK + 3) / 4) mma_ss_run<4>(d, bB[term] + off, km_hi, xlo + off, km_hi, IDESC1, acc);
              else mma_ss_run<LAST>(d, bB[term] + off, km_hi, xlo + off, km_hi, IDESC1, acc);
            }
          }
        TC_TRACE(1, i, 7);
        if (elect_one()) tc_commit(&s_full[buf]);
        TC_TRACE(1, i, 2);
        __syncwarp();
      }
    }
  } else if (warp == G2_WARP) {
    // ===================================================== GEMM2 issuer
    if (nb > 0) {
      constexpr uint32_t IDESC2 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(dk >> 3) << 17) | ((uint32_t)(CHAINS >> 4) << 24);
      const uint32_t aX = smem_u32(sX);
      const uint64_t dMN = desc_mnmajor(0, 0);
      const uint32_t mn_hi = (uint32_t)(dMN >> 32), mn_lo0 = (uint32_t)dMN;
      int period = 0, in_period = 0;
      for (int i = 0; i < nb; ++i) {
        const int st = i % NS, buf = i % NSB, u = i / NSB;
        TC_TRACE(1, i, 3);
        mbar_wait(&r_full[buf], (uint32_t)u & 1u);
        TC_TRACE(1, i, 4);
        if (in_period == 0 && period >= 1) mbar_wait(g_empty, (uint32_t)(period - 1) & 1u);
        TC_TRACE(1, i, 5);
        tc_fence_after();
        const uint32_t xm = mn_lo0 + ((aX + (uint32_t)st * P::X_BYTES) >> 4);
        const uint32_t a = tmem_S + (uint32_t)buf * 128u;
        const uint32_t acc0 = in_period > 0 ? 1u : 0u;
#pragma unroll
        for (int term = 0; term < (RR ? 1 : 2); ++term)
#pragma unroll
          for (int hb = 0; hb < 2; ++hb) {   // 64-row halves of the block
            if (TCDBG & 4) continue;
            mma_ts_run4(tmem_G, a + (uint32_t)(hb * 64 + term * 16), xm + (uint32_t)(hb * 4 * 128), mn_hi, IDESC2,
                        (term | hb) ? 1u : acc0);
          }
        TC_TRACE(1, i, 8);
        if (elect_one()) { tc_commit(&x_empty[st]); tc_commit(&sr_empty[buf]); }
        TC_TRACE(1, i, 6);
        ++in_period;
        if (i + 1 == nb || in_period == fe) {
          if (elect_one()) tc_commit(g_full);
          ++period;
          in_period = 0;
        }
        __syncwarp();
        TC_TRACE(1, i, 9);
      }
    }
  } else {
    // ===================================================== elementwise + epilogue (16 warps)
    // NG groups of 16/NG warps; group g takes blocks i = g mod NG.  Inside a group: TMEM lane group
    // q = warp % 4 (hardware rule), column part h; a warp owns CPW = NG chunks of 32 columns per block.
    // Measured (clock64 trace): the latency of one block around the loop S -> elementwise -> R -> GEMM2 ->
    // buffer free -> GEMM1, not the throughput of any pipe, limits the kernel (three TMEM buffers of slack);
    // NG = 2 (3.02 ms per full launch) beats NG = 1 (3.17 ms) once S is prefetched without blocking.
    const int ew = warp - 2;
    const int grp = ew / EW_PER_GROUP;
    const int h = (ew % EW_PER_GROUP) >> 2;
    const int q = warp & 3;
    const int row = tile * CHAINS + q * 32 + lane;   // staging row = chain slot
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    double lsum = 0.0;
    const float2 L2E2 = make_float2(1.4426950408889634f, 1.4426950408889634f);
    const float2 ONE2 = make_float2(1.0f, 1.0f), MHALF2 = make_float2(-0.5f, -0.5f), HALF2 = make_float2(0.5f, 0.5f);
    const float2 MONE2 = make_float2(-1.0f, -1.0f);
    const float LN2 = 0.6931471805599453f;
    float* gout = G + ((size_t)split * nrows + (size_t)row) * Dp;
    int fpos = grp, fper = 0;               // i % fe and i / fe, kept without divisions
    while (fpos >= fe) { fpos -= fe; ++fper; }
    // Work items of this warp: (block i, chunk cc), i = grp, grp + NG, ..., cc < CPW.  The TMEM load of
    // the next item is issued before the stores / barrier traffic of the current one, so its latency
    // (and the s_full wait of the next block, normally already complete) is off the critical path.
    // A warp whose 32 chains are all beyond the staged rows (last, partial tile) does no elementwise work: it
    // only keeps the barrier protocol going, so the MUFU pipe and the issue slots go to the warps with real rows
    // (garbage in the unused TMEM lanes stays in output rows nobody reads: a lane is a chain in both GEMMs).
    const bool live = (tile * CHAINS + q * 32) < nrows;
    uint32_t v[32];
    // `blocking = false`: only if S of that block is already there (the load is then a pure prefetch).  Waiting
    // here for S(i + NG) BEFORE publishing R(i) would chain every block's hand-off to the GEMM1 of a later block
    // (measured with the clock64 trace: 1 700 clk per block lost in exactly that wait).
    auto load_item = [&](int i, int cc, bool blocking) -> bool {
      const int buf = i % NSB, u = i / NSB;
      if (cc == 0) {
        if (!blocking) {
          if (!mbar_test(&s_full[buf], (uint32_t)u & 1u)) return false;
        } else {
          if (warp == 2) TC_TRACE(2, i, 0);
          mbar_wait(&s_full[buf], (uint32_t)u & 1u);
          if (warp == 2) TC_TRACE(2, i, 1);
        }
        tc_fence_after();
      }
      tmem_ld32(tmem_S + (uint32_t)buf * 128u + lane_sel + (uint32_t)(CPW * h + cc) * 32u, v);
      return true;
    };
    bool have_next = false;
    if (grp < nb && !(TCDBG & 1) && live) have_next = load_item(grp, 0, true);
    // bsum / asum run over LFOLD of this warp's blocks before they are folded into the Float64 sum: a DADD (+ conversion)
    // per block and thread was 5 % of all warp samples (ncu: the 16 elementwise warps hit the narrow FP64 pipe together).
    // Eight blocks keep the fp32 partial sums below ~1e3, i.e. their rounding below 1e-4 per fold (ℓ itself is ~1e5..1e6).
    constexpr int LFOLD = 8;
    float bsum = 0.f;   // sum over this thread's elements of  log2(1 + 2^-|u|)
    float asum = 0.f;   // sum of |eta|
    int nfold = 0;
    for (int i = grp; i < nb; i += NG) {
      const int buf = i % NSB;
      if (!live) mbar_wait(&s_full[buf], (uint32_t)(i / NSB) & 1u);   // stay in phase with the buffer, then just arrive
#pragma unroll 1
      for (int cc = 0; cc < (((TCDBG & 1) || !live) ? 0 : CPW); ++cc) {
        const int ch = CPW * h + cc;
        if (!have_next) have_next = load_item(i, cc, true);
        const uint32_t tS = tmem_S + (uint32_t)buf * 128u + lane_sel + (uint32_t)ch * 32u;
        if (warp == 2) TC_TRACE(2, i, 3);
        tmem_ld_wait();
        if (warp == 2) TC_TRACE(2, i, 4);
        uint32_t hi[16], lo[RR ? 1 : 16];
        float2 prod = ONE2;
        float2 as2 = make_float2(0.f, 0.f);   // Σ|η̃| of the even / odd columns: one FADD2 with |.| operand modifiers per pair
        if (RR && cc == 0) mbar_wait(&x_full[i % NS], (uint32_t)(i / NS) & 1u);   // c of this block is visible (long complete)
        const float2* c2p = reinterpret_cast<const float2*>(sR + (i % NS) * ROWS + ch * 32);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          // eta = H~ (natural units); t = exp(-|eta|) in (0,1]; d = 1 + t in (1,2]
          const float e0 = __uint_as_float(v[2 * j]), e1 = __uint_as_float(v[2 * j + 1]);
          const float2 u2 = __fmul2_rn(make_float2(e0, e1), L2E2);
          const float2 d2 = (TCDBG & 16) ? __fadd2_rn(u2, ONE2) : __fadd2_rn(make_float2(ex2_approx(-fabsf(u2.x)), ex2_approx(-fabsf(u2.y))), ONE2);
          prod = __fmul2_rn(prod, d2);                       // 32 factors in (1,2]: no overflow
          as2 = __fadd2_rn(as2, make_float2(fabsf(e0), fabsf(e1)));
          // rc = 1/d = sigma(|eta|) in [1/2,1);  sigma(-eta) = 1/2 - copysign(rc - 1/2, eta)
          const float2 hm = __fadd2_rn((TCDBG & 8) ? d2 : rcp2(d2, ((RCPSW >> (j & 7)) & 1) != 0), MHALF2);
          const float2 cs = make_float2(__uint_as_float(__float_as_uint(hm.x) | (v[2 * j] & 0x80000000u)),
                                        __uint_as_float(__float_as_uint(hm.y) | (v[2 * j + 1] & 0x80000000u)));
          if constexpr (RR == 1) {
            const float2 dl = __ffma2_rn(cs, MONE2, c2p[j]);   // δ = (½ − r0) − copysign(..) = r − r0, one rounding
            hi[j] = pack_bf16(dl.x, dl.y);
          } else {
            const float2 r2 = __ffma2_rn(cs, MONE2, HALF2);
            const uint32_t hh = pack_bf16(r2.x, r2.y);
            const float2 hv = make_float2(__uint_as_float(hh << 16), __uint_as_float(hh & 0xffff0000u));
            const float2 l2 = __ffma2_rn(hv, MONE2, r2);
            hi[j] = hh;
            lo[j] = pack_bf16(l2.x, l2.y);
          }
        }
        bsum += lg2_approx(prod.x * prod.y);
        asum += as2.x + as2.y;
        if (warp == 2) TC_TRACE(2, i, 5);
        // next item's S -> registers (v is dead now): always within the block, across blocks only if it is ready
        have_next = false;
        if (cc + 1 < CPW) have_next = load_item(i, cc + 1, true);
        else if (i + NG < nb) have_next = load_item(i + NG, 0, false);
        tmem_st16(tS, hi);
        if constexpr (!RR) tmem_st16(tS + 16u, lo);
      }
      if (warp == 2) TC_TRACE(2, i, 6);
      tmem_st_wait();
      if (warp == 2) TC_TRACE(2, i, 7);
      tc_fence_before();
      mbar_arrive(&r_full[buf]);
      if (warp == 2) TC_TRACE(2, i, 2);
      if (++nfold == LFOLD) {   // sum log sigma(eta) minus the linear part (added by the consumer)
        lsum += (double)fmaf(-LN2, bsum, -0.5f * asum);
        bsum = 0.f; asum = 0.f; nfold = 0;
      }
      const bool closes = (i + 1 == nb) || (fpos == fe - 1);
      const int period = fper;
      fpos += NG;
      while (fpos >= fe) { fpos -= fe; ++fper; }
      if (closes) {
        // this block closes flush period i / fe: drain the GEMM2 accumulator and add it outside the tensor core
        mbar_wait(g_full, (uint32_t)period & 1u);
        tc_fence_after();
#pragma unroll 1
        for (int cc = 0; cc < CPW; ++cc) {
          const int ch = CPW * h + cc;
          if (ch * 32 < dk && live) {
            uint32_t w[32];
            tmem_ld32(tmem_G + lane_sel + (uint32_t)ch * 32u, w);
            tmem_ld_wait();
            if (row < nrows) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const int d = ch * 32 + j;
                if (d < dk && d < Dp) {   // dk can exceed the row stride Dp when the reference columns spill over
                  float4 a = make_float4(__uint_as_float(w[j]), __uint_as_float(w[j + 1]), __uint_as_float(w[j + 2]),
                                         __uint_as_float(w[j + 3]));
                  float4* gp = reinterpret_cast<float4*>(gout + d);
                  // later periods add into what the first one stored.  A reduction (no return value) instead of load + add +
                  // store: the thread does not wait for the round trip to L2 (6 % of warp samples sat on that load); the adds of
                  // one address come from this thread only, in program order, so the sum is the same bits as before
                  if (period > 0) red_add_f32x4(gp, a);
                  else *gp = a;
                }
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(g_empty);
      }
    }
    lsum += (double)fmaf(-LN2, bsum, -0.5f * asum);
    // rows >= N of the last block are zero padding: eta = 0, y = 0 -> each contributed -log 2
    if (h == 0 && nb > 0 && grp == ((nb - 1) % NG) && b1 == nblk_total)
      lsum += (double)((long long)nblk_total * ROWS - N) * 0.6931471805599453;
    // combine the four partial sums of each chain (2 groups x 2 column halves) and publish
    double* lp = reinterpret_cast<double*>(sX);   // X stages are dead by now
    const int part = grp * (4 / NG) + h;
    asm volatile("bar.sync 1, 512;" ::: "memory");
    if (part > 0) lp[(part - 1) * 128 + q * 32 + lane] = lsum;
    asm volatile("bar.sync 1, 512;" ::: "memory");
    if (part == 0 && row < nrows) {
      if (nb == 0) for (int d = 0; d < Dp; ++d) gout[d] = 0.f;
      const int k = q * 32 + lane;
      Ld[(size_t)split * nrows + row] = ((lsum + lp[k]) + lp[128 + k]) + lp[256 + k];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

template <int DT, int NK, int RR> void launch_rr(LogisticTC& tc, cudaStream_t s, int nrows, int nsplit) {
  using P = SmemPlan<DT>;
  // the attribute is per device: one bit per device ordinal (engines on several GPUs may live in one process)
  static unsigned long long attr_done = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (!((attr_done >> (dev & 63)) & 1ull)) {
    tc.last = cudaFuncSetAttribute(k_logistic_tc<DT, NK, RR>, cudaFuncAttributeMaxDynamicSharedMemorySize, P::TOTAL);
    attr_done |= 1ull << (dev & 63);
  }
  const int tiles = (nrows + CHAINS - 1) / CHAINS;
  dim3 grid(tiles, nsplit);
  CUtensorMap m[4];
  for (int i = 0; i < 4; ++i) std::memcpy(&m[i], tc.tmaps[i], sizeof(CUtensorMap));
  k_logistic_tc<DT, NK, RR><<<grid, TC_THREADS, P::TOTAL, s>>>(m[0], m[1], m[2], m[3], tc.c0, tc.G, tc.Ld, nrows, tc.Dp,
                                                               (long long)tc.N, (int)(tc.Npad / ROWS), nsplit, tc.flush_every, tc.nterms);
}
template <int DT, int NK> void launch(LogisticTC& tc, cudaStream_t s, int nrows, int nsplit) {
  if (tc.rmode == 1) launch_rr<DT, NK, 1>(tc, s, nrows, nsplit);
  else launch_rr<DT, NK, 0>(tc, s, nrows, nsplit);
}


constexpr int ROWS2 = 64;            // data rows per block of k_logistic_tc256
constexpr int CHUNK2 = ROWS2 * 128;   // 64 rows x 64 bf16, one SW128 box
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

// =====================================================================================================
// Variant for 128 < D <= 256 (k_logistic_tc256; BASELINE config 5 has D = 256).
//
// With exact operand splits a D = 256 problem does not fit the layouts above: three β tiles would be 192 KB and
// G (256 columns) leaves no room for 128-column S/R buffers.  Here: 64-row blocks, K = dk
// exactly (no spare K columns), TWO-TERM mode only — the position operand is always β − β₀ split in two bf16
// terms, and η̃₀ = X̃β₀ is added in the elementwise stage from a per-row fp32 vector that travels with each X̃
// stage.  Without a reference point β₀ = 0 and the operand is a 16-bit β (relative gradient error ~1e-5..1e-4:
// enough for FindLocalOptimum, which then supplies the reference).  Shared memory: 2 β tiles (KC x 16 KB each)
// + 3 stages of 64-row X̃ blocks (KC x 8 KB + 256 B of η̃₀); TMEM: G [0, 256) + four 64-column S/R buffers.
// GEMM1 is an SS MMA with N = 64 (shared-memory bound: 6 KB of operands per 32 clk); GEMM2 one N = dk MMA per K step.
template <int KC> struct SmemPlan3 {
  static constexpr int B_BYTES = KC * CHUNK_BYTES;       // one β term: 128 chains x (KC x 64) columns
  static constexpr int X_BYTES = KC * CHUNK2;            // one X stage: 64 rows x (KC x 64) columns
  static constexpr int NS = 3;
  static constexpr int OFF_B = 0;
  static constexpr int OFF_X = 2 * B_BYTES;
  static constexpr int OFF_E = OFF_X + NS * X_BYTES;     // eta0: NS x 64 floats
  static constexpr int OFF_R = OFF_E + NS * ROWS2 * 4;   // ½ − r0 (single-term residual mode): NS x 64 floats
  static constexpr int OFF_BAR = OFF_R + NS * ROWS2 * 4;
  static constexpr int NBAR = 1 + 2 * NS + 3 * 4 + 2;
  static constexpr int TOTAL = OFF_BAR + NBAR * 8 + 16;
};

template <int NK, bool RR>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_logistic_tc256(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmBh,
                 const __grid_constant__ CUtensorMap tmBm, const float* __restrict__ eta0, const float* __restrict__ c0, float* G,
                 double* Ld, int nrows, int Dp, long long N, int nblk_total, int nsplit, int flush_every) {
  constexpr int KC = (NK + 3) / 4;
  using P = SmemPlan3<KC>;
  constexpr int dk = NK * 16;
  constexpr int NS = P::NS;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sB = smem + P::OFF_B;
  unsigned char* sX = smem + P::OFF_X;
  float* sE = reinterpret_cast<float*>(smem + P::OFF_E);
  float* sR = reinterpret_cast<float*>(smem + P::OFF_R);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P::OFF_BAR);
  uint64_t* bar_b = bars;
  uint64_t* x_full = bars + 1;
  uint64_t* x_empty = x_full + NS;
  uint64_t* s_full = x_empty + NS;
  uint64_t* r_full = s_full + 4;
  uint64_t* sr_empty = r_full + 4;
  uint64_t* g_full = sr_empty + 4;
  uint64_t* g_empty = g_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + P::NBAR);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, split = blockIdx.y;
  const int b0 = (int)(((long long)nblk_total * split) / nsplit);
  const int b1 = (int)(((long long)nblk_total * (split + 1)) / nsplit);
  const int nb = b1 - b0;
  const int fe = flush_every > 0 ? 2 * flush_every : 0x7fffffff;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) asm volatile("trap;");
    mbar_init(bar_b, 1);
    for (int i = 0; i < NS; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(&s_full[i], 1); mbar_init(&r_full[i], 256); mbar_init(&sr_empty[i], 1); }
    mbar_init(g_full, 1);
    mbar_init(g_empty, 256);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_G = tmem;             // [0, 256)
  const uint32_t tmem_S = tmem + 256u;      // 4 x 64 columns

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0 && nb > 0) {
      mbar_expect_tx(bar_b, 2 * P::B_BYTES);
      for (int kc = 0; kc < KC; ++kc) {
        tma_load_2d(&tmBh, sB + 0 * P::B_BYTES + kc * CHUNK_BYTES, bar_b, kc * 64, tile * CHAINS);
        tma_load_2d(&tmBm, sB + 1 * P::B_BYTES + kc * CHUNK_BYTES, bar_b, kc * 64, tile * CHAINS);
      }
      for (int i = 0; i < nb; ++i) {
        const int st = i % NS;
        const uint32_t ph = (uint32_t)(i / NS) & 1u;
        mbar_wait(&x_empty[st], ph ^ 1u);
        mbar_expect_tx(&x_full[st], P::X_BYTES + ROWS2 * 4 * (RR ? 2 : 1));
        for (int kc = 0; kc < KC; ++kc)
          tma_load_2d(&tmX, sX + st * P::X_BYTES + kc * CHUNK2, &x_full[st], kc * 64, (b0 + i) * ROWS2);
        bulk_load_1d(sE + st * ROWS2, eta0 + (size_t)(b0 + i) * ROWS2, ROWS2 * 4, &x_full[st]);
        if (RR) bulk_load_1d(sR + st * ROWS2, c0 + (size_t)(b0 + i) * ROWS2, ROWS2 * 4, &x_full[st]);
      }
    }
  } else if (warp == 1) {
    // ===================================================== GEMM1 issuer: S[buf] = (β − β₀) · X̃_iᵀ, both from smem
    if (nb > 0) {
      constexpr uint32_t IDESC1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(ROWS2 >> 3) << 17) | ((uint32_t)(CHAINS >> 4) << 24);
      const uint32_t aX = smem_u32(sX);
      const uint64_t dKM = desc_kmajor(0, 0);
      const uint32_t km_hi = (uint32_t)(dKM >> 32), km_lo0 = (uint32_t)dKM;
      const uint32_t bB0 = km_lo0 + (smem_u32(sB) >> 4);
      mbar_wait(bar_b, 0);
      for (int i = 0; i < nb; ++i) {
        const int st = i % NS, buf = i & 3, rnd = i >> 2;
        mbar_wait(&x_full[st], (uint32_t)(i / NS) & 1u);
        if (rnd >= 1) mbar_wait(&sr_empty[buf], (uint32_t)(rnd - 1) & 1u);
        tc_fence_after();
        const uint32_t xlo = km_lo0 + ((aX + (uint32_t)st * P::X_BYTES) >> 4);
        const uint32_t d = tmem_S + (uint32_t)buf * 64u;
#pragma unroll
        for (int term = 0; term < 2; ++term)
#pragma unroll
          for (int c = 0; c < KC; ++c) {
            constexpr int LAST = NK - (KC - 1) * 4;
            const uint32_t a = bB0 + (uint32_t)((term * P::B_BYTES + c * CHUNK_BYTES) >> 4);
            const uint32_t b = xlo + (uint32_t)(c * (CHUNK2 >> 4));
            const uint32_t acc = (term | c) ? 1u : 0u;
            if (c + 1 < KC) mma_ss_run<4>(d, a, km_hi, b, km_hi, IDESC1, acc);
            else mma_ss_run<LAST>(d, a, km_hi, b, km_hi, IDESC1, acc);
          }
        if (elect_one()) tc_commit(&s_full[buf]);
        __syncwarp();
      }
    }
  } else if (warp == G2_WARP) {
    // ===================================================== GEMM2 issuer: G += R (TMEM) · X̃_i (smem, MN-major), N = dk
    if (nb > 0) {
      constexpr uint32_t IDESC2 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(dk >> 3) << 17) | ((uint32_t)(CHAINS >> 4) << 24);
      const uint32_t aX = smem_u32(sX);
      const uint64_t dMN = make_desc(0, (uint32_t)CHUNK2, 1024u);
      const uint32_t mn_hi = (uint32_t)(dMN >> 32), mn_lo0 = (uint32_t)dMN;
      int period = 0, in_period = 0;
      for (int i = 0; i < nb; ++i) {
        const int st = i % NS, buf = i & 3, rnd = i >> 2;
        mbar_wait(&r_full[buf], (uint32_t)rnd & 1u);
        if (in_period == 0 && period >= 1) mbar_wait(g_empty, (uint32_t)(period - 1) & 1u);
        tc_fence_after();
        const uint32_t xm = mn_lo0 + ((aX + (uint32_t)st * P::X_BYTES) >> 4);
        const uint32_t a = tmem_S + (uint32_t)buf * 64u;
        const uint32_t acc0 = in_period > 0 ? 1u : 0u;
#pragma unroll
        for (int term = 0; term < (RR ? 1 : 2); ++term)
          mma_ts_run4(tmem_G, a + (uint32_t)(term * 16), xm, mn_hi, IDESC2, term ? 1u : acc0);
        if (elect_one()) { tc_commit(&x_empty[st]); tc_commit(&sr_empty[buf]); }
        ++in_period;
        if (i + 1 == nb || in_period == fe) {
          if (elect_one()) tc_commit(g_full);
          ++period;
          in_period = 0;
        }
        __syncwarp();
      }
    }
  } else {
    // ===================================================== elementwise + epilogue (16 warps, two groups of 8)
    const int ew = warp - 2;
    const int grp = ew >> 3;
    const int h = (ew >> 2) & 1;
    const int q = warp & 3;
    const int row = tile * CHAINS + q * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    const bool live = (tile * CHAINS + q * 32) < nrows;
    double lsum = 0.0;
    const float2 L2E2 = make_float2(1.4426950408889634f, 1.4426950408889634f);
    const float2 ONE2 = make_float2(1.0f, 1.0f), MHALF2 = make_float2(-0.5f, -0.5f), HALF2 = make_float2(0.5f, 0.5f);
    const float2 MONE2 = make_float2(-1.0f, -1.0f);
    const float LN2 = 0.6931471805599453f;
    float* gout = G + ((size_t)split * nrows + (size_t)row) * Dp;
    int fpos = grp, fper = 0;
    while (fpos >= fe) { fpos -= fe; ++fper; }
    uint32_t v[32];
    auto load_item = [&](int i, bool blocking) -> bool {
      const int buf = i & 3, rnd = i >> 2;
      if (!blocking) {
        if (!mbar_test(&s_full[buf], (uint32_t)rnd & 1u)) return false;
      } else {
        mbar_wait(&s_full[buf], (uint32_t)rnd & 1u);
      }
      tc_fence_after();
      tmem_ld32(tmem_S + (uint32_t)buf * 64u + lane_sel + (uint32_t)h * 32u, v);
      return true;
    };
    bool have_next = false;
    if (grp < nb && live) have_next = load_item(grp, true);
    constexpr int LFOLD = 8;    // blocks per fold of the fp32 partial sums into Float64 (see k_logistic_tc)
    float bsum = 0.f, asum = 0.f;
    int nfold = 0;
    for (int i = grp; i < nb; i += 2) {
      const int buf = i & 3, st = i % NS;
      if (!live) {
        mbar_wait(&s_full[buf], (uint32_t)(i >> 2) & 1u);
      } else {
        if (!have_next) have_next = load_item(i, true);
        mbar_wait(&x_full[st], (uint32_t)(i / NS) & 1u);     // eta0 of this block is visible (long complete)
        const float4* e4 = reinterpret_cast<const float4*>(sE + st * ROWS2 + h * 32);
        const float2* c2p = reinterpret_cast<const float2*>(sR + st * ROWS2 + h * 32);
        const uint32_t tS = tmem_S + (uint32_t)buf * 64u + lane_sel + (uint32_t)h * 32u;
        tmem_ld_wait();
        uint32_t hi[16], lo[RR ? 1 : 16];
        float2 prod = ONE2;
        float2 as2 = make_float2(0.f, 0.f);   // Σ|η̃| of the even / odd columns: one FADD2 with |.| operand modifiers per pair
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 eo = e4[j4];
          const float off[4] = {eo.x, eo.y, eo.z, eo.w};
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            const int j = 2 * j4 + jj;
            const float e0 = __uint_as_float(v[2 * j]) + off[2 * jj], e1 = __uint_as_float(v[2 * j + 1]) + off[2 * jj + 1];
            const float2 u2 = __fmul2_rn(make_float2(e0, e1), L2E2);
            const float2 d2 = __fadd2_rn(make_float2(ex2_approx(-fabsf(u2.x)), ex2_approx(-fabsf(u2.y))), ONE2);
            prod = __fmul2_rn(prod, d2);
            as2 = __fadd2_rn(as2, make_float2(fabsf(e0), fabsf(e1)));
            const float2 hm = __fadd2_rn(rcp2(d2, ((RCPSW >> (j & 7)) & 1) != 0), MHALF2);
            const float2 cs = make_float2(__uint_as_float(__float_as_uint(hm.x) | (__float_as_uint(e0) & 0x80000000u)),
                                          __uint_as_float(__float_as_uint(hm.y) | (__float_as_uint(e1) & 0x80000000u)));
            if constexpr (RR) {
              const float2 dl = __ffma2_rn(cs, MONE2, c2p[j]);   // δ = r − r0 (see k_logistic_tc)
              hi[j] = pack_bf16(dl.x, dl.y);
            } else {
              const float2 r2 = __ffma2_rn(cs, MONE2, HALF2);
              const uint32_t hh = pack_bf16(r2.x, r2.y);
              const float2 hv = make_float2(__uint_as_float(hh << 16), __uint_as_float(hh & 0xffff0000u));
              const float2 l2 = __ffma2_rn(hv, MONE2, r2);
              hi[j] = hh;
              lo[j] = pack_bf16(l2.x, l2.y);
            }
          }
        }
        bsum += lg2_approx(prod.x * prod.y);
        asum += as2.x + as2.y;
        have_next = false;
        if (i + 2 < nb) have_next = load_item(i + 2, false);
        tmem_st16(tS, hi);
        if constexpr (!RR) tmem_st16(tS + 16u, lo);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&r_full[buf]);
      if (++nfold == LFOLD) { lsum += (double)fmaf(-LN2, bsum, -0.5f * asum); bsum = 0.f; asum = 0.f; nfold = 0; }
      const bool closes = (i + 1 == nb) || (fpos == fe - 1);
      const int period = fper;
      fpos += 2;
      while (fpos >= fe) { fpos -= fe; ++fper; }
      if (closes) {
        mbar_wait(g_full, (uint32_t)period & 1u);
        tc_fence_after();
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          const int ch = 4 * h + cc;
          if (ch * 32 < dk && live) {
            uint32_t w[32];
            tmem_ld32(tmem_G + lane_sel + (uint32_t)ch * 32u, w);
            tmem_ld_wait();
            if (row < nrows) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const int d = ch * 32 + j;
                if (d < dk && d < Dp) {
                  float4 a = make_float4(__uint_as_float(w[j]), __uint_as_float(w[j + 1]), __uint_as_float(w[j + 2]),
                                         __uint_as_float(w[j + 3]));
                  float4* gp = reinterpret_cast<float4*>(gout + d);
                  // later periods add into what the first one stored.  A reduction (no return value) instead of load + add +
                  // store: the thread does not wait for the round trip to L2 (6 % of warp samples sat on that load); the adds of
                  // one address come from this thread only, in program order, so the sum is the same bits as before
                  if (period > 0) red_add_f32x4(gp, a);
                  else *gp = a;
                }
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(g_empty);
      }
    }
    // rows >= N of the last block are zero padding (X̃ = 0, eta0 = 0): each contributed -log 2
    lsum += (double)fmaf(-LN2, bsum, -0.5f * asum);
    if (h == 0 && nb > 0 && grp == ((nb - 1) & 1) && b1 == nblk_total)
      lsum += (double)((long long)nblk_total * ROWS2 - N) * 0.6931471805599453;
    double* lp = reinterpret_cast<double*>(sX);   // X stages are dead by now
    const int part = grp * 2 + h;
    asm volatile("bar.sync 1, 512;" ::: "memory");
    if (part > 0) lp[(part - 1) * 128 + q * 32 + lane] = lsum;
    asm volatile("bar.sync 1, 512;" ::: "memory");
    if (part == 0 && row < nrows) {
      if (nb == 0) for (int d = 0; d < Dp; ++d) gout[d] = 0.f;
      const int k = q * 32 + lane;
      Ld[(size_t)split * nrows + row] = ((lsum + lp[k]) + lp[128 + k]) + lp[256 + k];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

template <int NK, bool RR> void launch256_rr(LogisticTC& tc, cudaStream_t s, int nrows, int nsplit) {
  using P = SmemPlan3<(NK + 3) / 4>;
  static unsigned long long attr_done = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (!((attr_done >> (dev & 63)) & 1ull)) {
    tc.last = cudaFuncSetAttribute(k_logistic_tc256<NK, RR>, cudaFuncAttributeMaxDynamicSharedMemorySize, P::TOTAL);
    attr_done |= 1ull << (dev & 63);
  }
  const int tiles = (nrows + CHAINS - 1) / CHAINS;
  dim3 grid(tiles, nsplit);
  CUtensorMap m[3];
  std::memcpy(&m[0], tc.tmaps[4], sizeof(CUtensorMap));
  std::memcpy(&m[1], tc.tmaps[1], sizeof(CUtensorMap));
  std::memcpy(&m[2], tc.tmaps[2], sizeof(CUtensorMap));
  k_logistic_tc256<NK, RR><<<grid, TC_THREADS, P::TOTAL, s>>>(m[0], m[1], m[2], tc.eta0, tc.c0, tc.G, tc.Ld, nrows, tc.Dp,
                                                            (long long)tc.N, (int)(tc.Npad / ROWS2), nsplit, tc.flush_every);
}
template <int NK> void launch256(LogisticTC& tc, cudaStream_t s, int nrows, int nsplit) {
  if (tc.rmode == 1) launch256_rr<NK, true>(tc, s, nrows, nsplit);
  else launch256_rr<NK, false>(tc, s, nrows, nsplit);
}

}  // namespace

// number of row splits for `nrows` active rows: fill the SMs in as few full waves as possible
int LogisticTC::plan_splits(int nrows, int tile_rows) const {
  const int tiles = (nrows + tile_rows - 1) / tile_rows;
  const int64_t nblk = Npad / ROWS;
  if (force_nsplit > 0) return (int)std::min<int64_t>(force_nsplit, nblk);
  int best = 1;
  double best_eff = 0.0;
  for (int ns = 1; ns <= max_splits; ++ns) {
    if (ns > nblk) break;
    const int ctas = tiles * ns;
    const int waves = (ctas + sms - 1) / sms;
    const double eff = (double)ctas / ((double)waves * sms);
    const double score = eff * (1.0 - 0.01 * waves);   // prefer few waves (per-CTA set-up, partial traffic)
    if (score > best_eff + 1e-9) { best_eff = score; best = ns; }
  }
  return best;
}
void LogisticTC::run(cudaStream_t s, int nrows) {
  if (!ready || nrows <= 0) return;
  if (rmode == 2) {   // remainder mode: chains are the MMA N dimension, 64-chain tiles for small launches
    const int nc = (nrows <= 64 || variant == 256) ? 64 : 128;
    last_nsplit = plan_splits(nrows, nc);
    logistic_rm_launch(*this, s, nrows, last_nsplit, nc);
    return;
  }
  last_nsplit = plan_splits(nrows);
  if (variant == 256) {
    switch (dk / 16) {
      case 9: launch256<9>(*this, s, nrows, last_nsplit); break;
      case 10: launch256<10>(*this, s, nrows, last_nsplit); break;
      case 11: launch256<11>(*this, s, nrows, last_nsplit); break;
      case 12: launch256<12>(*this, s, nrows, last_nsplit); break;
      case 13: launch256<13>(*this, s, nrows, last_nsplit); break;
      case 14: launch256<14>(*this, s, nrows, last_nsplit); break;
      case 15: launch256<15>(*this, s, nrows, last_nsplit); break;
      default: launch256<16>(*this, s, nrows, last_nsplit); break;
    }
    return;
  }
  switch (dk / 16) {
    case 1: launch<64, 1>(*this, s, nrows, last_nsplit); break;
    case 2: launch<64, 2>(*this, s, nrows, last_nsplit); break;
    case 3: launch<64, 3>(*this, s, nrows, last_nsplit); break;
    case 4: launch<64, 4>(*this, s, nrows, last_nsplit); break;
    case 5: launch<128, 5>(*this, s, nrows, last_nsplit); break;
    case 6: launch<128, 6>(*this, s, nrows, last_nsplit); break;
    case 7: launch<128, 7>(*this, s, nrows, last_nsplit); break;
    default: launch<128, 8>(*this, s, nrows, last_nsplit); break;
  }
}
#ifdef BNUTS_TC_TRACE
extern "C" int bnuts_debug_tc_trace(long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_tc_trace, sizeof(long long) * 3 * 256 * 16);
}
#endif
void LogisticTC::destroy() {
  if (Xb) cudaFree(Xb);
  if (colsum) cudaFree(colsum);
  if (beta_ref) cudaFree(beta_ref);
  if (eta0) cudaFree(eta0);
  if (c0) cudaFree(c0);
  if (grad0) cudaFree(grad0);
  if (grad0_part) cudaFree(grad0_part);
  void** rmp[] = {(void**)&rec, (void**)&rm_r0, (void**)&rm_w, (void**)&rm_f0, (void**)&H0, (void**)&H0_part, (void**)&rm_part, (void**)&ell0};
  for (void** p : rmp) if (*p) { cudaFree(*p); *p = nullptr; }
  Xb = nullptr; colsum = nullptr; beta_ref = nullptr; eta0 = nullptr; c0 = nullptr; grad0 = nullptr; grad0_part = nullptr;
  ready = false; nterms = 3; rmode = 0; variant = 128;
}

namespace {
// one thread per data row: H~0_i = sum_d X~_id beta_ref_d in Float64, rounded once to fp32 and
// split exactly in three bf16 terms stored in columns D..D+2 of the row
__global__ void k_write_reference(uint16_t* Xb, const float* __restrict__ beta_ref, long long N, int D, int Dt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  uint16_t* xr = Xb + i * Dt;
  uint16_t t0 = 0, t1 = 0, t2 = 0;
  if (beta_ref) {
    double acc = 0.0;
    for (int d = 0; d < D; ++d) acc = fma((double)bf16_val(xr[d]), (double)beta_ref[d], acc);
    const float e = (float)acc;
    t0 = bf16_bits(e);
    const float r1 = e - bf16_val(t0);
    t1 = bf16_bits(r1);
    t2 = bf16_bits(r1 - bf16_val(t1));
  }
  xr[D] = t0; xr[D + 1] = t1; xr[D + 2] = t2;
}
__global__ void k_init_stage(uint16_t* bh, int C, int D, int Dt) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  for (int k = 0; k < 3; ++k) bh[(size_t)c * Dt + D + k] = 0x3F80;   // 1.0
}
}  // namespace

namespace {
__global__ void k_write_eta0(const uint16_t* __restrict__ Xb, const float* __restrict__ beta_ref, float* eta0, long long N, int D, int Dt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  double acc = 0.0;
  if (beta_ref) {
    const uint16_t* xr = Xb + i * Dt;
    for (int d = 0; d < D; ++d) acc = fma((double)bf16_val(xr[d]), (double)beta_ref[d], acc);
  }
  eta0[i] = (float)acc;
}
}  // namespace
namespace {
constexpr int G0_BLOCKS = 1024;
// one thread per data row: c_i = ½ − σ(−η̃0_i) = ½·tanh(η̃0_i / 2) with η̃0_i = X̃_i·beta_ref in Float64, rounded once
// to fp32.  Any fp32 value near it would do: what matters is that grad0 below is formed from the STORED values.
__global__ void k_write_c0(const uint16_t* __restrict__ Xb, const float* __restrict__ beta_ref, float* c0, long long N, long long Npad,
                           int D, int Dt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Npad) return;
  float c = 0.f;
  if (i < N) {
    const uint16_t* xr = Xb + i * Dt;
    double acc = 0.0;
    for (int d = 0; d < D; ++d) acc = fma((double)bf16_val(xr[d]), (double)beta_ref[d], acc);
    c = (float)(0.5 * tanh(0.5 * acc));
  }
  c0[i] = c;
}
// grad0_d = Σ_i X̃_id (½ − c_i): block b sums its contiguous range of rows (thread = column), then one block
// adds the G0_BLOCKS partials in order — the same bits on every run and every rank
__global__ void k_grad0_partial(const uint16_t* __restrict__ Xb, const float* __restrict__ c0, double* part, long long N, int D,
                                int Dt, int Dp) {
  const long long r0 = N * blockIdx.x / gridDim.x, r1 = N * (blockIdx.x + 1) / gridDim.x;
  for (int d = threadIdx.x; d < Dp; d += blockDim.x) {
    double acc = 0.0;
    if (d < D)
      for (long long i = r0; i < r1; ++i) acc = fma((double)bf16_val(Xb[i * Dt + d]), c0 ? 0.5 - (double)c0[i] : 1.0, acc);
    part[(size_t)blockIdx.x * Dp + d] = acc;
  }
}
__global__ void k_grad0_sum(const double* __restrict__ part, double* grad0, int nb, int Dp) {
  for (int d = threadIdx.x; d < Dp; d += blockDim.x) {
    double acc = 0.0;
    for (int b = 0; b < nb; ++b) acc += part[(size_t)b * Dp + d];
    grad0[d] = acc;
  }
}
}  // namespace
void logistic_tc_write_residual_reference(LogisticTC& tc, cudaStream_t s, const float* beta_ref_dev) {
  k_write_c0<<<(unsigned)((tc.Npad + 255) / 256), 256, 0, s>>>(tc.Xb, beta_ref_dev, tc.c0, (long long)tc.N, (long long)tc.Npad, tc.D,
                                                              tc.Dt);
  k_grad0_partial<<<G0_BLOCKS, 128, 0, s>>>(tc.Xb, tc.c0, tc.grad0_part, (long long)tc.N, tc.D, tc.Dt, tc.Dp);
  k_grad0_sum<<<1, 128, 0, s>>>(tc.grad0_part, tc.grad0, G0_BLOCKS, tc.Dp);
}
void logistic_tc_write_reference(LogisticTC& tc, cudaStream_t s, const float* beta_ref_dev) {
  if (tc.variant == 256) {
    k_write_eta0<<<(unsigned)((tc.N + 255) / 256), 256, 0, s>>>(tc.Xb, beta_ref_dev, tc.eta0, (long long)tc.N, tc.D, tc.Dt);
    return;
  }
  if (!tc.aug) return;
  k_write_reference<<<(unsigned)((tc.N + 255) / 256), 256, 0, s>>>(tc.Xb, beta_ref_dev, (long long)tc.N, tc.D, tc.Dt);
}
void logistic_tc_init_stage(LogisticTC& tc, cudaStream_t s, uint16_t* bh) {
  if (!tc.aug) return;
  k_init_stage<<<(tc.C + 127) / 128, 128, 0, s>>>(bh, tc.C, tc.D, tc.Dt);
}

namespace {
// shape of the problem: K / N of the MMAs, tile width, kernel variant, padded row count, environment overrides
void tc_shape(LogisticTC& tc, int64_t N, int32_t C, int32_t D, int32_t Dp) {
  tc.destroy();
  tc.C = C; tc.D = D; tc.Dp = Dp; tc.N = N;
  tc.aug = (D + 3 <= 128) ? 1 : 0;
  tc.dk = (D + (tc.aug ? 3 : 0) + 15) / 16 * 16;
  tc.Dt = (tc.dk <= 64) ? 64 : 128;
  if (D > 128) {   // k_logistic_tc256: K = dk exactly, two-term mode with eta0 = X~ beta_ref added elementwise
    tc.aug = 0;
    tc.dk = (D + 15) / 16 * 16;
    tc.Dt = (tc.dk + 63) / 64 * 64;
    tc.variant = 256;
    tc.nterms = 2;
  }
  tc.Npad = (N + ROWS - 1) / ROWS * ROWS;
  const char* fe = std::getenv("BNUTS_TC_FLUSH");
  if (fe) tc.flush_every = std::atoi(fe);
  const char* se = std::getenv("BNUTS_TC_NSPLIT");
  tc.force_nsplit = se ? std::atoi(se) : 0;
}
// device buffers (zeroed) and the split plan
int32_t tc_alloc(LogisticTC& tc, std::string& err) {
  const size_t xbytes = size_t(tc.Npad) * tc.Dt * 2;
  if (cudaMalloc(&tc.Xb, xbytes) != cudaSuccess || cudaMalloc(&tc.colsum, size_t(tc.Dp) * 8) != cudaSuccess ||
      cudaMalloc(&tc.beta_ref, size_t(tc.Dp) * 4) != cudaSuccess) {
    err = "device allocation failed (tensor path X)";
    return BNUTS_ERR_CUDA;
  }
  cudaMemset(tc.Xb, 0, xbytes);
  cudaMemset(tc.colsum, 0, size_t(tc.Dp) * 8);
  cudaMemset(tc.beta_ref, 0, size_t(tc.Dp) * 4);
  if (cudaMalloc(&tc.c0, size_t(tc.Npad) * 4) != cudaSuccess || cudaMalloc(&tc.grad0, size_t(tc.Dp) * 8) != cudaSuccess ||
      cudaMalloc(&tc.grad0_part, size_t(G0_BLOCKS) * tc.Dp * 8) != cudaSuccess) {
    err = "device allocation failed (residual reference)";
    return BNUTS_ERR_CUDA;
  }
  cudaMemset(tc.c0, 0, size_t(tc.Npad) * 4);
  cudaMemset(tc.grad0, 0, size_t(tc.Dp) * 8);
  if (tc.variant == 256) {
    if (cudaMalloc(&tc.eta0, size_t(tc.Npad) * 4) != cudaSuccess) { err = "device allocation failed (eta0)"; return BNUTS_ERR_CUDA; }
    cudaMemset(tc.eta0, 0, size_t(tc.Npad) * 4);
  }
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&tc.sms, cudaDevAttrMultiProcessorCount, dev);
  if (tc.sms <= 0) tc.sms = 148;
  // staging rows needed for the partial outputs: max over nrows of nsplit(nrows) * nrows
  tc.max_splits = 3 * tc.sms;
  const int64_t nblk = tc.Npad / ROWS;
  if (tc.max_splits > nblk) tc.max_splits = (int)nblk;
  int64_t worst = tc.C;
  for (int tiles = 1; tiles <= (tc.C + CHAINS - 1) / CHAINS; ++tiles) {
    const int nr = std::min<int>(tc.C, tiles * CHAINS);
    worst = std::max<int64_t>(worst, (int64_t)tc.plan_splits(nr) * nr);
  }
  for (int tiles = 1; tiles <= (tc.C + 63) / 64; ++tiles) {   // remainder mode: 64-chain tiles (small launches; every launch for D > 128)
    const int nr = std::min<int>(tc.C, tiles * 64);
    worst = std::max<int64_t>(worst, (int64_t)tc.plan_splits(nr, 64) * nr);
  }
  tc.partial_rows = worst;
  return 0;
}
}  // namespace

int32_t logistic_tc_build(LogisticTC& tc, const uint16_t* Xh, const double* y, int64_t N, int32_t C, int32_t D, int32_t Dp,
                          std::string& err) {
  tc_shape(tc, N, C, D, Dp);
  // sign-folded design matrix: row i multiplied by (2 y_i - 1) (flip of the bf16 sign bit)
  std::vector<uint16_t> xp(size_t(tc.Npad) * tc.Dt, 0);
  std::vector<double> cs(size_t(Dp), 0.0);
  for (int64_t i = 0; i < N; ++i) {
    if (y[i] != 0.0 && y[i] != 1.0) {
      err = "tensor gradient path needs y in {0, 1}; use BNUTS_GRAD_DETERMINISTIC";
      return BNUTS_ERR_UNSUPPORTED;
    }
    const uint16_t flip = (y[i] == 0.0) ? 0x8000 : 0;
    uint16_t* dst = &xp[size_t(i) * tc.Dt];
    const uint16_t* src = &Xh[size_t(i) * D];
    for (int d = 0; d < D; ++d) {
      dst[d] = src[d] ^ flip;
      cs[d] += double(bf16_val(dst[d]));
    }
  }
  const int32_t rc = tc_alloc(tc, err);
  if (rc) return rc;
  cudaMemcpy(tc.Xb, xp.data(), xp.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(tc.colsum, cs.data(), cs.size() * 8, cudaMemcpyHostToDevice);
  return 0;
}

// ≙ SURVEY.md §8d, config c5: the rows of this shard are generated ON THE DEVICE from Philox keyed by (data seed, global
// row index) (bnuts_math.h, synth_*), sign-folded on the way; nothing but the seed crosses PCIe.  `fill` is the
// generator kernel pair of engine_cuda.cu (compiled without FMA contraction: the same bits as the host generator).
int32_t logistic_tc_build_synth(LogisticTC& tc, uint64_t seed, int64_t row0, int64_t N, int32_t C, int32_t D, int32_t Dp,
                                cudaStream_t s, SynthFillFn fill, std::string& err) {
  tc_shape(tc, N, C, D, Dp);
  const int32_t rc = tc_alloc(tc, err);
  if (rc) return rc;
  const int frc = fill(s, seed, row0, N, D, tc.Dt, tc.Xb);
  if (frc != 0) { err = std::string("synthetic data generation failed: ") + cudaGetErrorString((cudaError_t)frc); return BNUTS_ERR_CUDA; }
  // column sums of X~ in Float64, fixed order (same two-pass reduction as grad0, weight 1)
  k_grad0_partial<<<G0_BLOCKS, 128, 0, s>>>(tc.Xb, nullptr, tc.grad0_part, (long long)tc.N, tc.D, tc.Dt, tc.Dp);
  k_grad0_sum<<<1, 128, 0, s>>>(tc.grad0_part, tc.colsum, G0_BLOCKS, tc.Dp);
  if (cudaStreamSynchronize(s) != cudaSuccess) { err = "synthetic data generation failed"; return BNUTS_ERR_CUDA; }
  return 0;
}

int32_t logistic_tc_maps(LogisticTC& tc, std::string& err) {
  const uint64_t crow = (uint64_t)tc.C;
  if (!encode_map(tc.tmaps[0], tc.Xb, (uint64_t)tc.Npad, (uint64_t)tc.Dt) ||
      !encode_map(tc.tmaps[1], tc.bh, crow, (uint64_t)tc.Dt) || !encode_map(tc.tmaps[2], tc.bm, crow, (uint64_t)tc.Dt) ||
      !encode_map(tc.tmaps[3], tc.bl, crow, (uint64_t)tc.Dt)) {
    err = "cuTensorMapEncodeTiled failed";
    return BNUTS_ERR_CUDA;
  }
  if (!encode_map(tc.tmaps[4], tc.Xb, (uint64_t)tc.Npad, (uint64_t)tc.Dt, 64) || !encode_map(tc.tmaps[5], tc.bh, crow, (uint64_t)tc.Dt, 64) ||
      !encode_map(tc.tmaps[6], tc.bm, crow, (uint64_t)tc.Dt, 64)) {
    err = "cuTensorMapEncodeTiled failed";
    return BNUTS_ERR_CUDA;
  }
  tc.ready = true;
  return 0;
}

}  // namespace bn

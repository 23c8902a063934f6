// engine_cuda.cu — CUDA execution policy of the engine (sm_100a) and libbnuts.so's
// C ABI.  One warp owns one chain; see backend.h for the HBM layout and
// nuts_machine.h for the per-chain state machine these kernels run.
//
// Kernels:
//   k_prepare        per-call chain set-up                     (≙ loop prologues, src/warmup.jl:283-287)
//   k_advance        consume leaf + merges + next leapfrog     (≙ src/tree.jl:321-444, src/kinetic_energy.jl:126-163)
//   k_grad_gaussian  deterministic G = -P·Q                    (≙ logdensity_and_gradient!, src/kinetic_energy.jl:73)
//   k_grad_logistic  deterministic logistic-regression gradient, fixed row-block order
//   logistic_tc.cu   tcgen05/TMA fused two-GEMM logistic gradient (fp32 variant)
//   k_metric, k_finish_da, small gathers
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>   // header-only; ranges cost nothing unless a tool is attached
#include <dlfcn.h>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "engine_core.h"
#include "logistic_tc.h"
#include "gauss_tc.h"

namespace bn {

constexpr int ADV_THREADS = 128;  // 4 chains per CTA
#ifndef BNUTS_ADV_MIN_BLOCKS
#define BNUTS_ADV_MIN_BLOCKS 4
#endif
constexpr int ADV_MIN_BLOCKS = BNUTS_ADV_MIN_BLOCKS;   // register budget of k_advance = 65536 / (128 x this); measured 2 / 3 / 4: funnel f64 74 / 87 / 89 M leapfrog steps/s, c2 25.4 / 25.4 / 29.3 M

template <class T> __global__ void __launch_bounds__(ADV_THREADS) k_prepare(EngineMem<T> M, RunParams<T> rp, PrepareArgs a) {
  const int c = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  if (c >= M.C) return;
  prepare_chain(M, rp, a, c, WarpLanes{(int)(threadIdx.x & 31)});
}

template <class T, int RM>
__global__ void __launch_bounds__(ADV_THREADS, ADV_MIN_BLOCKS) k_advance(EngineMem<T> M, RunParams<T> rp, int iters, unsigned long long* pending) {
  const int c = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  if (c >= M.C) return;
  const int lane = (int)(threadIdx.x & 31);
  const bool p = advance_chain<T, WarpLanes, RM>(M, rp, c, WarpLanes{lane}, iters);
  if (p && lane == 0 && !M.stage_q) atomicAdd(pending, 1ull);  // batched targets count through take_row()
}

// Same, for the pipelined run loop: no memset / copy / event around it.  The request counter alternates between
// two device words (this launch counts into `cnt` and clears `cnt_next` for the following one); the last CTA to
// finish publishes {sequence number, count} into a pinned, device-mapped host word the host polls.
template <class T, int RM>
__global__ void __launch_bounds__(ADV_THREADS, ADV_MIN_BLOCKS) k_advance_ring(EngineMem<T> M, RunParams<T> rp, int iters, unsigned long long* cnt,
                                                              unsigned long long* cnt_next, unsigned int* done,
                                                              volatile unsigned long long* host_slot, unsigned int seq) {
  const int c = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  if (c < M.C) {
    const int lane = (int)(threadIdx.x & 31);
    const bool p = advance_chain<T, WarpLanes, RM>(M, rp, c, WarpLanes{lane}, iters);
    if (p && lane == 0 && !M.stage_q) atomicAdd(cnt, 1ull);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int ticket = atomicAdd(done, 1u);
    if (ticket == gridDim.x - 1) {
      const unsigned long long n = atomicAdd(cnt, 0ull);
      *cnt_next = 0ull;
      *done = 0u;
      __threadfence_system();
      *host_slot = ((unsigned long long)seq << 32) | (n & 0xffffffffull);
    }
  }
}

template <class T> __global__ void __launch_bounds__(ADV_THREADS) k_metric(EngineMem<T> M, int N, double lambda) {
  const int c = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  if (c >= M.C) return;
  metric_update_chain(M, c, N, lambda, WarpLanes{(int)(threadIdx.x & 31)});
}

template <class T> __global__ void k_finish_da(EngineMem<T> M) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < M.C && M.cs[c].status == 0) M.cs[c].eps = exp_(M.cs[c].da_logepsbar);  // ≙ final_ϵ, src/stepsize.jl:241
}
template <class T> __global__ void k_set_eps(EngineMem<T> M, const double* e) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < M.C) M.cs[c].eps = e[c];
}
template <class T> __global__ void k_get_eps(EngineMem<T> M, double* e) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < M.C) e[c] = M.cs[c].eps;
}
template <class T> __global__ void k_get_status(EngineMem<T> M, int32_t* st) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < M.C) st[c] = M.cs[c].status;
}
template <class T> __global__ void k_totals(EngineMem<T> M, unsigned long long* tot) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= M.C) return;
  atomicAdd(&tot[0], (unsigned long long)M.cs[c].tot_leapfrogs);
  atomicAdd(&tot[1], (unsigned long long)M.cs[c].tot_transitions);
  atomicAdd(&tot[2], (unsigned long long)M.cs[c].tot_divergences);
}
template <class T> __global__ void k_gather_state(EngineMem<T> M, double* o) {
  const int c = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  if (c >= M.C) return;
  const int lane = (int)(threadIdx.x & 31);
  const int64_t n = (int64_t)M.C * M.D;
  const int s = M.cs[c].slot_cur;
  const T* q = M.zs + ((int64_t)c * M.S + s) * 3 * M.Dp;
  for (int d = lane; d < M.D; d += 32) {
    o[(int64_t)c * M.D + d] = (double)q[d];
    o[n + (int64_t)c * M.D + d] = (double)q[2 * M.Dp + d];
  }
  if (lane == 0) o[2 * n + c] = M.zlq[(int64_t)c * M.S + s];
}

// ------------------------------------------------------------------ row-sharded mode
// Deterministic row assignment: exclusive scan of the request flags in chain order (one CTA; C is a few
// thousand), so every engine of the group gives chain c the same staging row.
__global__ void __launch_bounds__(1024) k_scan_rows(const int32_t* __restrict__ active, int32_t* stage_row, int C,
                                                    unsigned long long* count) {
  __shared__ int wsum[32];
  __shared__ int base_s;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) base_s = 0;
  __syncthreads();
  for (int c0 = 0; c0 < C; c0 += 1024) {
    const int c = c0 + threadIdx.x;
    const int f = (c < C && active[c]) ? 1 : 0;
    int incl = f;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    if (w == 0) {
      int v = wsum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
      wsum[lane] = v;   // inclusive over warps
    }
    __syncthreads();
    const int base = base_s;
    const int excl = base + (w ? wsum[w - 1] : 0) + incl - f;
    if (f) stage_row[c] = excl;
    __syncthreads();
    if (threadIdx.x == 0) base_s = base + wsum[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = (unsigned long long)base_s;
}
// move the flagged chains' requests from the wide staging (row = chain) to their compact rows
template <class T>
__global__ void __launch_bounds__(ADV_THREADS) k_gather_rows(EngineMem<T> V, EngineMem<T> Mc) {
  const int c = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  if (c >= V.C) return;
  const int lane = (int)(threadIdx.x & 31);
  if (!V.stage_active[c]) return;
  const int64_t r = Mc.stage_row[c];
  for (int d = lane; d < V.Dp; d += 32) Mc.stage_q[r * V.Dp + d] = V.stage_q[(int64_t)c * V.Dp + d];
  if (Mc.stage_bh)
    for (int d = lane; d < V.Dt; d += 32) {
      Mc.stage_bh[r * V.Dt + d] = V.stage_bh[(int64_t)c * V.Dt + d];
      Mc.stage_bm[r * V.Dt + d] = V.stage_bm[(int64_t)c * V.Dt + d];
      Mc.stage_bl[r * V.Dt + d] = V.stage_bl[(int64_t)c * V.Dt + d];
    }
  __syncwarp();
  if (lane == 0) V.stage_active[c] = 0;
}
// fold this shard's partial blocks (row blocks of the deterministic path / splits of the tensor path) in a
// fixed order: one gradient row + one Float64 log-density per staged row, ready for the sum over the group
template <class T>
__global__ void __launch_bounds__(ADV_THREADS) k_fold_partials(EngineMem<T> M, int rows, T* red_g, double* red_l) {
  const int row = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  if (row >= rows) return;
  const int lane = (int)(threadIdx.x & 31);
  const int64_t bs = (int64_t)rows * M.Dp;
  const T* sg = M.stage_g + (int64_t)row * M.Dp;
  double lin = 0.0;
  for (int d = lane; d < M.Dp; d += 32) {
    T acc = T(0);
    if (d < M.D) {
      for (int b = 0; b < M.stage_nb; ++b) acc = acc + sg[b * bs + d];
      if (M.grad0 && !M.lin_H) acc = T((double)acc + M.grad0[d]);   // remainder mode: the consumer adds g0 (summed over the group) with the linear part
      if (M.lin_w) lin = fma(M.lin_w[d], (double)M.stage_q[(int64_t)row * M.Dp + d], lin);
    }
    red_g[(int64_t)row * M.Dp + d] = acc;
  }
  for (int off = 16; off >= 1; off >>= 1) lin += __shfl_xor_sync(0xffffffffu, lin, off);
  if (lane == 0) {
    double l = 0.0;
    for (int b = 0; b < M.stage_nb; ++b)
      l += M.stage_ld ? M.stage_ld[(int64_t)b * rows + row] : (double)M.stage_l[(int64_t)b * rows + row];
    red_l[row] = fma(0.5, lin, l);
  }
}

// ---- peer-memory exchange (bnuts_p2p_*): every rank owns one receive buffer
//   [ flags: 2 parities x 8 ranks x u64 | pad to 256 B | parity 0: rank 0 {G rows, L rows}, rank 1 {...}, ... | parity 1: ... ]
// and holds the (IPC-mapped) base pointers of all ranks' buffers.
struct P2PView {
  unsigned char* peer[8];
  int world, rank;
  size_t g_bytes, l_bytes;      // per (parity, source rank) block: C*Dp*sizeof(T) + C*8
  __host__ __device__ size_t block_off(int parity, int src) const { return 256 + ((size_t)parity * 8 + src) * (g_bytes + l_bytes); }
  __host__ __device__ size_t flag_off(int parity, int src) const { return ((size_t)parity * 8 + src) * 8; }
};
// fold this shard's partial blocks (as k_fold_partials) and push the folded row into the receive slot
// [parity][my rank] of EVERY rank over NVLink; the last CTA to finish raises the flag on every rank
template <class T>
__global__ void __launch_bounds__(ADV_THREADS) k_fold_push(EngineMem<T> M, int rows, P2PView V, unsigned long long seq, unsigned int* done) {
  const int row = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  const int lane = (int)(threadIdx.x & 31);
  const int parity = (int)(seq & 1ull);
  if (row < rows) {
    const int64_t bs = (int64_t)rows * M.Dp;
    const T* sg = M.stage_g + (int64_t)row * M.Dp;
    double lin = 0.0;
    for (int d = lane; d < M.Dp; d += 32) {
      T acc = T(0);
      if (d < M.D) {
        for (int b = 0; b < M.stage_nb; ++b) acc = acc + sg[b * bs + d];
        if (M.grad0 && !M.lin_H) acc = T((double)acc + M.grad0[d]);
        if (M.lin_w) lin = fma(M.lin_w[d], (double)M.stage_q[(int64_t)row * M.Dp + d], lin);
      }
      for (int r = 0; r < V.world; ++r)
        reinterpret_cast<T*>(V.peer[r] + V.block_off(parity, V.rank))[(int64_t)row * M.Dp + d] = acc;
    }
    for (int off = 16; off >= 1; off >>= 1) lin += __shfl_xor_sync(0xffffffffu, lin, off);
    if (lane == 0) {
      double l = 0.0;
      for (int b = 0; b < M.stage_nb; ++b)
        l += M.stage_ld ? M.stage_ld[(int64_t)b * rows + row] : (double)M.stage_l[(int64_t)b * rows + row];
      l = fma(0.5, lin, l);
      for (int r = 0; r < V.world; ++r)
        reinterpret_cast<double*>(V.peer[r] + V.block_off(parity, V.rank) + V.g_bytes)[row] = l;
    }
  }
  // all stores of this CTA system-visible, then count it; the last CTA publishes the sequence number everywhere
  __threadfence_system();
  __syncthreads();
  __shared__ unsigned int ticket;
  if (threadIdx.x == 0) ticket = atomicAdd(done, 1u);
  __syncthreads();
  if (ticket == gridDim.x - 1 && threadIdx.x == 0) {
    __threadfence_system();
    for (int r = 0; r < V.world; ++r)
      *reinterpret_cast<volatile unsigned long long*>(V.peer[r] + V.flag_off(parity, V.rank)) = seq;
    *done = 0u;
    __threadfence_system();
  }
}
// wait until every rank's rows of this step have arrived, then add the slots in rank order
template <class T>
__global__ void __launch_bounds__(ADV_THREADS) k_wait_sum(int rows, int Dp, P2PView V, unsigned long long seq, T* red_g, double* red_l,
                                                          int* err_flag) {
  const int parity = (int)(seq & 1ull);
  const unsigned char* mine = V.peer[V.rank];
  if (threadIdx.x < V.world) {
    const volatile unsigned long long* f = reinterpret_cast<const volatile unsigned long long*>(mine + V.flag_off(parity, threadIdx.x));
    const long long t0 = clock64();
    while (*f < seq) {
      if (clock64() - t0 > 8000000000ll) { *err_flag = 1; break; }   // ~4 s: a peer died; do not hang the GPU
      __nanosleep(200);
    }
  }
  __syncthreads();
  __threadfence_system();
  const int row = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  if (row >= rows) return;
  const int lane = (int)(threadIdx.x & 31);
  for (int d = lane; d < Dp; d += 32) {
    T acc = T(0);
    for (int r = 0; r < V.world; ++r)
      acc = acc + __ldcv(reinterpret_cast<const T*>(mine + V.block_off(parity, r)) + (int64_t)row * Dp + d);
    red_g[(int64_t)row * Dp + d] = acc;
  }
  if (lane == 0) {
    double l = 0.0;
    for (int r = 0; r < V.world; ++r) l += __ldcv(reinterpret_cast<const double*>(mine + V.block_off(parity, r) + V.g_bytes) + row);
    red_l[row] = l;
  }
}

// NCCL is bound at run time (dlopen) so the library loads on hosts without it; only bnuts_set_nccl needs it
struct NcclApi {
  typedef struct { char internal[128]; } UniqueId;
  void* lib = nullptr;
  int (*GetUniqueId)(UniqueId*) = nullptr;
  int (*CommInitRank)(void**, int, UniqueId, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool load(std::string& err);
};

// ------------------------------------------------------------------ deterministic gradients
// G[c][d] = -sum_k P[d][k] Q[c][k], k strictly sequential per output so the result
// is bit-identical to the oracle's scalar loop (no split-K, explicit fma).  Register-tiled SIMT GEMM:
// TM x TN output tile per CTA (d x chains), RM x RN outputs per thread, operands staged through shared
// memory K-slab by K-slab with 16-byte loads along k.
template <class T, int TM, int TN, int RM, int RN>
__global__ void __launch_bounds__((TM / RM) * (TN / RN)) k_grad_gaussian(const T* __restrict__ P, const T* Q, T* G, int C, int D, int Dp) {
  constexpr int TK = 16;
  constexpr int NT = (TM / RM) * (TN / RN);
  constexpr int VEC = 16 / sizeof(T);                 // elements per 16-byte load
  __shared__ __align__(16) T Ps[TK][TM + 4];
  __shared__ __align__(16) T Qs[TK][TN + 4];
  const int d0 = blockIdx.x * TM, c0 = blockIdx.y * TN;
  const int tx = threadIdx.x % (TN / RN), ty = threadIdx.x / (TN / RN);
  T acc[RM][RN];
#pragma unroll
  for (int i = 0; i < RM; ++i)
#pragma unroll
    for (int j = 0; j < RN; ++j) acc[i][j] = T(0);
  const bool vec_ok = (D % VEC) == 0;                  // rows of P stay 16-byte aligned
  for (int k0 = 0; k0 < D; k0 += TK) {
    // stage P[d0.., k0..] and Q[c0.., k0..] transposed ([k][row]); VEC consecutive k per load
    for (int idx = threadIdx.x; idx < TM * (TK / VEC); idx += NT) {
      const int r = idx / (TK / VEC), kv = (idx % (TK / VEC)) * VEC;
      const int d = d0 + r;
      T v[VEC];
      if (vec_ok && d < D && k0 + kv + VEC <= D) {
        *reinterpret_cast<uint4*>(v) = *reinterpret_cast<const uint4*>(P + (int64_t)d * D + k0 + kv);
      } else {
#pragma unroll
        for (int e = 0; e < VEC; ++e) v[e] = (d < D && k0 + kv + e < D) ? P[(int64_t)d * D + k0 + kv + e] : T(0);
      }
#pragma unroll
      for (int e = 0; e < VEC; ++e) Ps[kv + e][r] = v[e];
    }
    for (int idx = threadIdx.x; idx < TN * (TK / VEC); idx += NT) {
      const int r = idx / (TK / VEC), kv = (idx % (TK / VEC)) * VEC;
      const int c = c0 + r;
      T v[VEC];
      if (c < C && k0 + kv + VEC <= D) {               // Dp is a multiple of 32: rows of Q are aligned
        *reinterpret_cast<uint4*>(v) = *reinterpret_cast<const uint4*>(Q + (int64_t)c * Dp + k0 + kv);
      } else {
#pragma unroll
        for (int e = 0; e < VEC; ++e) v[e] = (c < C && k0 + kv + e < D) ? Q[(int64_t)c * Dp + k0 + kv + e] : T(0);
      }
#pragma unroll
      for (int e = 0; e < VEC; ++e) Qs[kv + e][r] = v[e];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      T a[RM], b[RN];
#pragma unroll
      for (int i = 0; i < RM; ++i) a[i] = Ps[k][ty * RM + i];
#pragma unroll
      for (int j = 0; j < RN; ++j) b[j] = Qs[k][tx * RN + j];
#pragma unroll
      for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < RN; ++j) acc[i][j] = fma_(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < RN; ++j) {
    const int c = c0 + tx * RN + j;
    if (c >= C) continue;
#pragma unroll
    for (int i = 0; i < RM; ++i) {
      const int d = d0 + ty * RM + i;
      if (d < D) G[(int64_t)c * Dp + d] = -acc[i][j];
    }
  }
}

// out[r][j] = sum_k in[r][k] mat[k][j] in Float64, k strictly sequential: the coordinate maps of the dense
// metric at the C-ABI boundary (positions, momenta, gradients, draws).  64x64 output tile, 4x4 per thread.
__global__ void __launch_bounds__(256) k_rows_times_matrix(const double* __restrict__ in, double* out, long long R, int D,
                                                           const double* __restrict__ mat) {
  constexpr int TM = 64, TN = 64, TK = 16;
  __shared__ double As[TK][TM + 1];
  __shared__ double Bs[TK][TN + 1];
  const long long r0 = (long long)blockIdx.y * TM;
  const int j0 = blockIdx.x * TN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  for (int k0 = 0; k0 < D; k0 += TK) {
    for (int idx = threadIdx.x; idx < TM * TK; idx += 256) {
      const int r = idx / TK, k = idx % TK;
      As[k][r] = (r0 + r < R && k0 + k < D) ? in[(r0 + r) * D + k0 + k] : 0.0;
      const int kb = idx / TN, jb = idx % TN;
      Bs[kb][jb] = (k0 + kb < D && j0 + jb < D) ? mat[(long long)(k0 + kb) * D + j0 + jb] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[k][ty * 4 + i]; b[i] = Bs[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long r = r0 + ty * 4 + i;
      const int jj = j0 + tx * 4 + j;
      if (r < R && jj < D) out[r * D + jj] = acc[i][j];
    }
}

// One warp = 32 chains, one CTA per (chain group, row block).  Rows of the block are
// walked sequentially; eta and the gradient partials accumulate in index order.
template <class T>
__global__ void __launch_bounds__(32) k_grad_logistic(const T* __restrict__ X, const T* __restrict__ y, const T* Q, T* G,
                                                       T* Lp, int64_t N, int D, int Dp, int C, int64_t R) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int XR = 8;
  T* qs = reinterpret_cast<T*>(smem_raw);  // [D][32]
  T* ps = qs + (size_t)D * 32;             // [D][32]
  T* xs = ps + (size_t)D * 32;             // [XR][D]
  T* ys = xs + (size_t)XR * D;             // [XR]
  const int lane = threadIdx.x;
  const int c = blockIdx.x * 32 + lane;
  const int b = blockIdx.y;
  const bool valid = c < C;
  for (int d = 0; d < D; ++d) {
    qs[d * 32 + lane] = valid ? Q[(int64_t)c * Dp + d] : T(0);
    ps[d * 32 + lane] = T(0);
  }
  T pl = T(0);
  const int64_t i0 = (int64_t)b * R;
  const int64_t i1 = (i0 + R < N) ? i0 + R : N;
  for (int64_t it = i0; it < i1; it += XR) {
    const int nr = (int)((i1 - it < XR) ? (i1 - it) : XR);
    __syncwarp();
    for (int idx = lane; idx < nr * D; idx += 32) xs[idx] = X[it * D + idx];
    if (lane < nr) ys[lane] = y[it + lane];
    __syncwarp();
    for (int r = 0; r < nr; ++r) {
      const T* xr = xs + r * D;
      T eta = T(0);
      for (int d = 0; d < D; ++d) eta = fma_(xr[d], qs[d * 32 + lane], eta);
      T rr, lt;
      logistic_elem(eta, ys[r], &rr, &lt);
      pl = pl + lt;
      for (int d = 0; d < D; ++d) ps[d * 32 + lane] = fma_(xr[d], rr, ps[d * 32 + lane]);
    }
  }
  if (valid) {
    T* g = G + ((int64_t)b * C + c) * Dp;
    for (int d = 0; d < D; ++d) g[d] = ps[d * 32 + lane];
    Lp[(int64_t)b * C + c] = pl;
  }
}

// ------------------------------------------------------------------ synthetic rows (bnuts_model_logistic_synthetic)
// ≙ SURVEY.md §8d, config c5: rows generated on the device from Philox keyed by (data seed, global row index), the
// definition in bnuts_math.h (synth_*).  This file is compiled without FMA contraction, so the bits are the host
// generator's.  k_synth_labels: one thread per row (eta accumulated in the defined order).  k_synth_rows: one
// thread per four columns, coalesced 8-byte stores of the sign-folded bf16 row (X~_i = (2 y_i - 1) X_i).
__global__ void k_synth_labels(uint64_t seed, long long row0, long long N, int D, const double* __restrict__ beta, uint8_t* y) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  y[i] = synth_label(seed, (uint64_t)(row0 + i), D, beta) != 0.0 ? 1 : 0;
}
__global__ void k_synth_beta(uint64_t seed, int D, double* beta) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d < D) beta[d] = synth_beta(seed, (uint32_t)d, D);
}
__global__ void k_synth_rows(uint64_t seed, long long row0, long long N, int D, int Dt, const uint8_t* __restrict__ y, uint16_t* Xb) {
  const int nq = (D + 3) >> 2;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long i = t / nq;
  const int q = (int)(t - i * nq);
  if (i >= N) return;
  float x[4];
  synth_x4(seed, (uint64_t)(row0 + i), (uint32_t)q, x);
  const uint16_t flip = y[i] ? 0 : 0x8000;
  uint16_t h[4];
  for (int e = 0; e < 4; ++e) h[e] = (4 * q + e < D) ? (uint16_t)((f2u(x[e]) >> 16) ^ flip) : (uint16_t)0;
  // Dt is a multiple of 64: the four columns of a quad never straddle the row end and the store is 8-byte aligned
  *reinterpret_cast<uint2*>(Xb + i * Dt + 4 * q) = make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16));
}
int synth_fill_xb(cudaStream_t s, uint64_t seed, int64_t row0, int64_t N, int32_t D, int32_t Dt, uint16_t* Xb) {
  double* beta = nullptr;
  uint8_t* y = nullptr;
  if (cudaMalloc(&beta, (size_t)D * 8) != cudaSuccess) return (int)cudaErrorMemoryAllocation;
  if (cudaMalloc(&y, (size_t)N) != cudaSuccess) { cudaFree(beta); return (int)cudaErrorMemoryAllocation; }
  k_synth_beta<<<(D + 127) / 128, 128, 0, s>>>(seed, D, beta);
  k_synth_labels<<<(unsigned)((N + 127) / 128), 128, 0, s>>>(seed, (long long)row0, (long long)N, D, beta, y);
  const long long nt = (long long)N * ((D + 3) >> 2);
  k_synth_rows<<<(unsigned)((nt + 255) / 256), 256, 0, s>>>(seed, (long long)row0, (long long)N, D, Dt, y, Xb);
  cudaError_t e = cudaGetLastError();
  const cudaError_t e2 = cudaStreamSynchronize(s);
  if (e == cudaSuccess) e = e2;
  cudaFree(beta); cudaFree(y);
  return (int)e;
}

// ------------------------------------------------------------------ execution policy
struct CudaExec {
  static constexpr bool has_tensor_path = true;
  // NVTX range per phase of the C ABI (tracing); shows up in Nsight Systems / ncu --nvtx
  struct Range {
    explicit Range(const char* name) { nvtxRangePushA(name); }
    ~Range() { nvtxRangePop(); }
    Range(const Range&) = delete; Range& operator=(const Range&) = delete;
  };
  cudaStream_t stream = nullptr;
  cudaError_t first_err = cudaSuccess;
  const char* first_where = "";
  unsigned long long* d_scal = nullptr;  // [4]
  unsigned long long* h_scal = nullptr;  // pinned [4]
  int32_t* d_status = nullptr;
  int device = 0;
  LogisticTC tc;
  GaussTC gt;
  void* nccl_comm = nullptr;
  // peer-memory exchange
  unsigned char* p2p_buf = nullptr; size_t p2p_bytes = 0;
  P2PView p2p{};
  unsigned int* p2p_done = nullptr; int* p2p_err = nullptr;
  // measurement hook: CUDA-event pairs around the batched gradient launches
  bool profiling = false;
  std::vector<cudaEvent_t> ev;   // pairs
  size_t ev_used = 0;
  double prof_ms = 0.0;
  int64_t prof_n = 0;
  void prof_fold() {
    if (ev_used == 0) return;
    note(cudaStreamSynchronize(stream), "profile sync");
    for (size_t i = 0; i + 1 < ev_used; i += 2) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, ev[i], ev[i + 1]) == cudaSuccess) { prof_ms += ms; prof_n += 1; }
    }
    ev_used = 0;
  }
  int32_t profile(int32_t enable, double* ms, int64_t* n) {
    prof_fold();
    if (ms) *ms = prof_ms;
    if (n) *n = prof_n;
    prof_ms = 0.0; prof_n = 0;
    profiling = enable != 0;
    return 0;
  }
  cudaEvent_t next_event() {
    if (ev_used == ev.size()) {
      if (ev.size() >= 16384) prof_fold();
      else { cudaEvent_t e; note(cudaEventCreate(&e), "event create"); ev.push_back(e); }
    }
    return ev[ev_used++];
  }

  void note(cudaError_t e, const char* where) {
    if (e != cudaSuccess && first_err == cudaSuccess) { first_err = e; first_where = where; }
  }
  int32_t init(int dev, std::string& err) {
    device = dev;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) { err = std::string("no CUDA device: ") + cudaGetErrorString(e); return BNUTS_ERR_CUDA; }
    if (dev < 0 || dev >= n) { err = "bad device ordinal"; return BNUTS_ERR_INVALID_ARGUMENT; }
    note(cudaSetDevice(dev), "cudaSetDevice");
    note(cudaMalloc(&d_scal, 4 * sizeof(unsigned long long)), "cudaMalloc");
    note(cudaMallocHost(&h_scal, 4 * sizeof(unsigned long long)), "cudaMallocHost");
    return check(err);
  }
  void shutdown() {
    if (nccl_comm && nccl().CommDestroy) nccl().CommDestroy(nccl_comm);
    nccl_comm = nullptr;
    for (int r = 0; r < p2p.world; ++r) if (r != p2p.rank && p2p.peer[r]) cudaIpcCloseMemHandle(p2p.peer[r]);
    if (p2p_buf) cudaFree(p2p_buf);
    if (p2p_done) cudaFree(p2p_done);
    if (p2p_err) cudaFree(p2p_err);
    tc.destroy();
    gt.destroy();
    if (d_scal) cudaFree(d_scal);
    if (h_scal) cudaFreeHost(h_scal);
    if (d_status) cudaFree(d_status);
    if (h_ring) { cudaFreeHost(const_cast<unsigned long long*>(h_ring)); h_ring = nullptr; }
    if (d_done) cudaFree(d_done);
    for (cudaEvent_t e : ev) cudaEventDestroy(e);
  }
  template <class U> U* alloc(size_t n) {
    void* p = nullptr;
    note(cudaSetDevice(device), "cudaSetDevice");
    note(cudaMalloc(&p, (n ? n : 1) * sizeof(U)), "cudaMalloc");
    return static_cast<U*>(p);
  }
  void free(void* p) { cudaFree(p); }
  // engines on several devices may be driven from one host thread: every entry re-selects the engine's device
  void use() { note(cudaSetDevice(device), "cudaSetDevice"); }
  void h2d(void* d, const void* s, size_t n) { use(); note(cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, stream), "h2d"); note(cudaStreamSynchronize(stream), "h2d sync"); }
  void d2h(void* d, const void* s, size_t n) { use(); note(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, stream), "d2h"); note(cudaStreamSynchronize(stream), "d2h sync"); }
  void d2d(void* d, const void* s, size_t n) { note(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToDevice, stream), "d2d"); }
  void rows_times_matrix(const double* in, double* out, int64_t R, int D, const double* mat) {
    dim3 grid((D + 63) / 64, (unsigned)((R + 63) / 64));
    k_rows_times_matrix<<<grid, 256, 0, stream>>>(in, out, (long long)R, D, mat);
    note(cudaGetLastError(), "rows_times_matrix");
  }
  void zero(void* d, size_t n) { note(cudaMemsetAsync(d, 0, n, stream), "memset"); }
  void sync() { note(cudaStreamSynchronize(stream), "sync"); }
  int32_t check(std::string& err) {
    note(cudaGetLastError(), "kernel launch");
    if (first_err == cudaSuccess) return 0;
    err = std::string("CUDA failure in ") + first_where + ": " + cudaGetErrorString(first_err);
    return BNUTS_ERR_CUDA;
  }
  static int warp_grid(int C) { return (C * 32 + ADV_THREADS - 1) / ADV_THREADS; }

  unsigned long long* counter() { return d_scal; }
  int64_t read_count() {
    note(cudaMemcpyAsync(h_scal, d_scal, sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream), "count d2h");
    note(cudaStreamSynchronize(stream), "count sync");
    return first_err == cudaSuccess ? (int64_t)h_scal[0] : 0;
  }
  template <class T> void prepare(const EngineMem<T>& M, const RunParams<T>& rp, const PrepareArgs& a) {
    note(cudaSetDevice(device), "cudaSetDevice");
    note(cudaMemsetAsync(d_scal, 0, sizeof(unsigned long long), stream), "memset");
    k_prepare<T><<<warp_grid(M.C), ADV_THREADS, 0, stream>>>(M, rp, a);
  }
  // ---- pipelined run loop (engine_core.h run_pipelined): request counts return through a ring of pinned,
  // device-mapped host words written by the last CTA of k_advance_ring; two stream operations per lockstep step
  // (gradient kernel, k_advance_ring), no memset, no copy, no event
  static constexpr int RING = 4;
  volatile unsigned long long* h_ring = nullptr;   // pinned + mapped [RING]
  unsigned long long* d_ring = nullptr;            // device alias of h_ring
  unsigned int* d_done = nullptr;
  unsigned int ring_seq = 0, ring_expect[RING] = {};
  template <class T> void advance_async(const EngineMem<T>& M, const RunParams<T>& rp, int iters, int slot, int parity) {
    if (!h_ring) {
      void* hp = nullptr; void* dp = nullptr;
      note(cudaHostAlloc(&hp, RING * sizeof(unsigned long long), cudaHostAllocMapped), "cudaHostAlloc");
      note(cudaHostGetDevicePointer(&dp, hp, 0), "cudaHostGetDevicePointer");
      h_ring = static_cast<volatile unsigned long long*>(hp); d_ring = static_cast<unsigned long long*>(dp);
      for (int i = 0; i < RING; ++i) h_ring[i] = 0ull;
      note(cudaMalloc(&d_done, sizeof(unsigned int)), "cudaMalloc");
      note(cudaMemsetAsync(d_done, 0, sizeof(unsigned int), stream), "memset");
    }
    ring_seq += 1;
    if (ring_seq == 0) ring_seq = 1;
    ring_expect[slot] = ring_seq;
    EngineMem<T> Ml = M;
    Ml.stage_count = d_scal + parity;
    // the remainder mode of the tensor-core logistic path has its own instance (fp32 engines only): see Backend
    bool rm = false;
    if constexpr (std::is_same<T, float>::value) rm = M.lin_H != nullptr;
    if (rm) {
      if constexpr (std::is_same<T, float>::value)
        k_advance_ring<T, 1><<<warp_grid(M.C), ADV_THREADS, 0, stream>>>(Ml, rp, iters, d_scal + parity, d_scal + (parity ^ 1), d_done, d_ring + slot, ring_seq);
    } else {
      k_advance_ring<T, 0><<<warp_grid(M.C), ADV_THREADS, 0, stream>>>(Ml, rp, iters, d_scal + parity, d_scal + (parity ^ 1), d_done, d_ring + slot, ring_seq);
    }
    note(cudaGetLastError(), "k_advance");
  }
  bool count_ready(int slot) { return (unsigned int)(h_ring[slot] >> 32) == ring_expect[slot]; }
  int64_t count_wait(int slot) {
    for (unsigned long long spins = 0; !count_ready(slot); ++spins) {
      if ((spins & 0xfff) == 0xfff) {                 // the kernel may have faulted: never spin on a dead stream
        const cudaError_t e = cudaStreamQuery(stream);
        if (e != cudaSuccess && e != cudaErrorNotReady) { note(e, "k_advance"); return 0; }
        if (e == cudaSuccess && !count_ready(slot)) { note(cudaErrorUnknown, "request count never arrived"); return 0; }
      }
    }
    return first_err == cudaSuccess ? (int64_t)(h_ring[slot] & 0xffffffffull) : 0;
  }
  // both counter words clear at the start of a run
  void reset_counters() { note(cudaMemsetAsync(d_scal, 0, 2 * sizeof(unsigned long long), stream), "memset"); }
  bool failed() const { return first_err != cudaSuccess; }
  template <class T> int64_t advance(const EngineMem<T>& M, const RunParams<T>& rp, int iters) {
    note(cudaMemsetAsync(d_scal, 0, sizeof(unsigned long long), stream), "memset");
    bool rm = false;
    if constexpr (std::is_same<T, float>::value) rm = M.lin_H != nullptr;
    if (rm) {
      if constexpr (std::is_same<T, float>::value) k_advance<T, 1><<<warp_grid(M.C), ADV_THREADS, 0, stream>>>(M, rp, iters, d_scal);
    } else {
      k_advance<T, 0><<<warp_grid(M.C), ADV_THREADS, 0, stream>>>(M, rp, iters, d_scal);
    }
    note(cudaGetLastError(), "k_advance");
    note(cudaMemcpyAsync(h_scal, d_scal, sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream), "pending d2h");
    note(cudaStreamSynchronize(stream), "k_advance sync");
    if (first_err != cudaSuccess) return 0;
    return (int64_t)h_scal[0];
  }
  // evaluate the staged rows [0, rows); returns the number of partial blocks written per row
  template <class E> int gradient(E& eng, int rows) {
    if (profiling) note(cudaEventRecord(next_event(), stream), "event record");
    const int nb = gradient_launch(eng, rows);
    if (profiling) note(cudaEventRecord(next_event(), stream), "event record");
    return nb;
  }
  template <class E> int gradient_launch(E& eng, int rows) {
    auto& M = eng.M;
    using T = typename std::remove_reference<decltype(*M.zs)>::type;
    int nb = 1;
    if (eng.model.kind == MODEL_GAUSSIAN && eng.model.tensor) {
      gt.run(stream, rows);
    } else if (eng.model.kind == MODEL_GAUSSIAN) {
      // tile size by the number of active rows, so straggler steps (a few chains deep in their trees)
      // still spread over many SMs; every variant keeps k sequential per output (bit-identical results)
      if (sizeof(T) == 4 && rows > 1024) {
        dim3 grid((M.D + 127) / 128, (rows + 127) / 128);
        k_grad_gaussian<T, 128, 128, 8, 8><<<grid, 256, 0, stream>>>(eng.model.P, M.stage_q, M.stage_g, rows, M.D, M.Dp);
      } else if (rows > 128) {
        dim3 grid((M.D + 63) / 64, (rows + 63) / 64);
        k_grad_gaussian<T, 64, 64, 4, 4><<<grid, 256, 0, stream>>>(eng.model.P, M.stage_q, M.stage_g, rows, M.D, M.Dp);
      } else {
        dim3 grid((M.D + 31) / 32, (rows + 15) / 16);
        k_grad_gaussian<T, 32, 16, 2, 2><<<grid, 128, 0, stream>>>(eng.model.P, M.stage_q, M.stage_g, rows, M.D, M.Dp);
      }
    } else if (eng.model.kind == MODEL_LOGISTIC) {
      if (eng.model.tensor) { tc.run(stream, rows); nb = tc.last_nsplit; }
      else {
        nb = eng.model.row_blocks;
        const int64_t R = (eng.model.N + nb - 1) / nb;
        const size_t smem = ((size_t)M.D * 64 + (size_t)8 * M.D + 8) * sizeof(T);
        note(cudaFuncSetAttribute(k_grad_logistic<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "smem attr");
        dim3 grid((rows + 31) / 32, nb);
        k_grad_logistic<T><<<grid, 32, smem, stream>>>(eng.model.X, eng.model.y, M.stage_q, M.stage_g, M.stage_l,
                                                      eng.model.N, M.D, M.Dp, rows, R);
      }
    }
    note(cudaGetLastError(), "gradient kernel");
    return nb;
  }
  // Gaussian target on the tensor cores: staging of the three-term bf16 split of q (written by the chains,
  // backend.h stage_put4) + the split of -P
  template <class E> int32_t gauss_tensor_setup(E& eng, const std::vector<double>& P, std::string& err) {
    using T = typename std::remove_reference<decltype(*eng.M.zs)>::type;
    if constexpr (!std::is_same<T, float>::value) {
      err = "tensor gradient path needs dtype F32";
      return BNUTS_ERR_UNSUPPORTED;
    } else {
      auto& M = eng.M;
      int32_t rc = gauss_tc_build(gt, P, M.C, M.D, M.Dp, err);
      if (rc) return rc;
      if (M.stage_bh && M.Dt != gt.Kp) { free(M.stage_bh); free(M.stage_bm); free(M.stage_bl); M.stage_bh = M.stage_bm = M.stage_bl = nullptr; }
      M.Dt = gt.Kp;
      const size_t nbt = size_t(M.C) * gt.Kp;
      if (!M.stage_bh) {
        M.stage_bh = alloc<uint16_t>(nbt); M.stage_bm = alloc<uint16_t>(nbt); M.stage_bl = alloc<uint16_t>(nbt);
        zero(M.stage_bh, nbt * 2); zero(M.stage_bm, nbt * 2); zero(M.stage_bl, nbt * 2);
      }
      gt.qh = M.stage_bh; gt.qm = M.stage_bm; gt.ql = M.stage_bl; gt.G = M.stage_g;
      return gauss_tc_maps(gt, err);
    }
  }
  template <class E> int32_t logistic_tensor_setup(E& eng, const void* Xh, int32_t xd, const double* y, int64_t N, std::string& err) {
    return logistic_tc_setup(tc, eng, Xh, xd, y, N, err);
  }
  template <class E> int32_t logistic_tensor_setup_synth(E& eng, uint64_t seed, int64_t row0, int64_t N, std::string& err) {
    return logistic_tc_setup_synth(tc, eng, seed, row0, N, &synth_fill_xb, err);
  }
  // ---- row-sharded mode
  static NcclApi& nccl() { static NcclApi api; return api; }
  static int32_t nccl_unique_id(uint8_t* id) {
    std::string err;
    if (!nccl().load(err)) return BNUTS_ERR_UNSUPPORTED;
    NcclApi::UniqueId u;
    if (nccl().GetUniqueId(&u) != 0) return BNUTS_ERR_CUDA;
    std::memcpy(id, u.internal, 128);
    return 0;
  }
  int32_t nccl_init(const uint8_t* id, int world, int rank, std::string& err) {
    if (!nccl().load(err)) return BNUTS_ERR_UNSUPPORTED;
    note(cudaSetDevice(device), "cudaSetDevice");
    if (nccl_comm) { nccl().CommDestroy(nccl_comm); nccl_comm = nullptr; }
    NcclApi::UniqueId u;
    std::memcpy(u.internal, id, 128);
    const int r = nccl().CommInitRank(&nccl_comm, world, u, rank);
    if (r != 0) { err = std::string("ncclCommInitRank: ") + nccl().GetErrorString(r); nccl_comm = nullptr; return BNUTS_ERR_CUDA; }
    return 0;
  }
  int32_t p2p_export(size_t g_bytes, size_t l_bytes, uint8_t* handle, std::string& err) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    note(cudaSetDevice(device), "cudaSetDevice");
    if (p2p_buf) { cudaFree(p2p_buf); p2p_buf = nullptr; }
    p2p = P2PView{};
    p2p.g_bytes = g_bytes; p2p.l_bytes = l_bytes;
    p2p_bytes = 256 + 2 * 8 * (g_bytes + l_bytes);
    if (cudaMalloc(&p2p_buf, p2p_bytes) != cudaSuccess) { err = "device allocation failed (peer exchange buffer)"; return BNUTS_ERR_CUDA; }
    note(cudaMemset(p2p_buf, 0, p2p_bytes), "memset");
    if (!p2p_done) { note(cudaMalloc(&p2p_done, 4), "cudaMalloc"); note(cudaMalloc(&p2p_err, 4), "cudaMalloc"); }
    note(cudaMemset(p2p_done, 0, 4), "memset"); note(cudaMemset(p2p_err, 0, 4), "memset");
    cudaIpcMemHandle_t h;
    const cudaError_t e = cudaIpcGetMemHandle(&h, p2p_buf);
    if (e != cudaSuccess) { err = std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e); return BNUTS_ERR_CUDA; }
    std::memcpy(handle, &h, 64);
    return check(err);
  }
  int32_t p2p_connect(const uint8_t* handles, int world, int rank, std::string& err) {
    if (!p2p_buf) { err = "bnuts_p2p_export first"; return BNUTS_ERR_INVALID_ARGUMENT; }
    note(cudaSetDevice(device), "cudaSetDevice");
    p2p.world = world; p2p.rank = rank;
    for (int r = 0; r < world; ++r) {
      if (r == rank) { p2p.peer[r] = p2p_buf; continue; }
      cudaIpcMemHandle_t h;
      std::memcpy(&h, handles + (size_t)r * 64, 64);
      void* ptr = nullptr;
      const cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) { err = std::string("cudaIpcOpenMemHandle (peer access over NVLink needed): ") + cudaGetErrorString(e); return BNUTS_ERR_CUDA; }
      p2p.peer[r] = static_cast<unsigned char*>(ptr);
    }
    return check(err);
  }
  bool p2p_failed() {
    int v = 0;
    d2h(&v, p2p_err, sizeof(int));
    return v != 0;
  }
  template <class T> void fold_push(const EngineMem<T>& M, int rows, uint64_t seq) {
    k_fold_push<T><<<warp_grid(rows), ADV_THREADS, 0, stream>>>(M, rows, p2p, (unsigned long long)seq, p2p_done);
    note(cudaGetLastError(), "fold_push");
  }
  template <class T> void wait_sum(const EngineMem<T>& M, int rows, uint64_t seq, T* red_g, double* red_l) {
    k_wait_sum<T><<<warp_grid(rows), ADV_THREADS, 0, stream>>>(rows, M.Dp, p2p, (unsigned long long)seq, red_g, red_l, p2p_err);
    note(cudaGetLastError(), "wait_sum");
  }
  template <class T> int64_t assign_rows(const EngineMem<T>& V, const EngineMem<T>& Mc) {
    k_scan_rows<<<1, 1024, 0, stream>>>(V.stage_active, Mc.stage_row, V.C, d_scal);
    k_gather_rows<T><<<warp_grid(V.C), ADV_THREADS, 0, stream>>>(V, Mc);
    note(cudaGetLastError(), "row assignment");
    return read_count();
  }
  template <class T> void fold_partials(const EngineMem<T>& M, int rows, T* red_g, double* red_l) {
    k_fold_partials<T><<<warp_grid(rows), ADV_THREADS, 0, stream>>>(M, rows, red_g, red_l);
    note(cudaGetLastError(), "fold partials");
  }
  int32_t allreduce(void* g, int64_t ng, bool g_is_f32, double* l, int64_t nl, bnuts_allreduce_fn fn, void* ctx, std::string& err) {
    if (nccl_comm && !fn) {
      // one fused exchange per leapfrog step: [rows x Dp] gradient + [rows] log density, summed in place
      nccl().GroupStart();
      int r1 = nccl().AllReduce(g, g, (size_t)ng, g_is_f32 ? 7 : 8, 0, nccl_comm, stream);   // ncclFloat32 = 7, ncclFloat64 = 8, ncclSum = 0
      int r2 = nccl().AllReduce(l, l, (size_t)nl, 8, 0, nccl_comm, stream);
      int r3 = nccl().GroupEnd();
      if (r1 || r2 || r3) { err = std::string("ncclAllReduce: ") + nccl().GetErrorString(r1 ? r1 : (r2 ? r2 : r3)); return BNUTS_ERR_CUDA; }
      return 0;
    }
    if (!fn) { err = "row-sharded mode without a collective"; return BNUTS_ERR_INTERNAL; }
    note(cudaStreamSynchronize(stream), "allreduce sync");
    if (fn(ctx, g, ng, g_is_f32 ? 1 : 0) != 0 || fn(ctx, l, nl, 0) != 0) { err = "host allreduce callback failed"; return BNUTS_ERR_INTERNAL; }
    return 0;
  }
  template <class E> int32_t logistic_reference(E& eng, const double* beta_ref, std::string& err) {
    return logistic_tc_set_reference(tc, eng, beta_ref, err);
  }
  int reference_mode() const { return tc.ready ? tc.rmode : 0; }
  int launches_per_gradient() const { return 1; }
  template <class T> void metric_update(const EngineMem<T>& M, int N, double lambda) {
    k_metric<T><<<warp_grid(M.C), ADV_THREADS, 0, stream>>>(M, N, lambda);
  }
  template <class T> void finish_da(const EngineMem<T>& M) { k_finish_da<T><<<(M.C + 127) / 128, 128, 0, stream>>>(M); }
  template <class T> void set_eps(const EngineMem<T>& M, const double* e) { k_set_eps<T><<<(M.C + 127) / 128, 128, 0, stream>>>(M, e); }
  template <class T> void get_eps(const EngineMem<T>& M, double* e) { k_get_eps<T><<<(M.C + 127) / 128, 128, 0, stream>>>(M, e); }
  template <class T> void get_status(const EngineMem<T>& M, int32_t* st) {
    if (!d_status) note(cudaMalloc(&d_status, (size_t)M.C * sizeof(int32_t)), "cudaMalloc");
    k_get_status<T><<<(M.C + 127) / 128, 128, 0, stream>>>(M, d_status);
    d2h(st, d_status, (size_t)M.C * sizeof(int32_t));
  }
  template <class T> bool any_status(const EngineMem<T>& M, int32_t code) {
    std::vector<int32_t> st(M.C);
    get_status(M, st.data());
    for (int32_t v : st) if (v == code) return true;
    return false;
  }
  template <class T> void gather_state(const EngineMem<T>& M, double* o) {
    k_gather_state<T><<<warp_grid(M.C), ADV_THREADS, 0, stream>>>(M, o);
  }
  template <class T> void totals(const EngineMem<T>& M, int64_t* tot) {
    note(cudaMemsetAsync(d_scal, 0, 3 * sizeof(unsigned long long), stream), "memset");
    k_totals<T><<<(M.C + 127) / 128, 128, 0, stream>>>(M, d_scal);
    d2h(h_scal, d_scal, 3 * sizeof(unsigned long long));
    for (int i = 0; i < 3; ++i) tot[i] = (int64_t)h_scal[i];
  }
};

bool NcclApi::load(std::string& err) {
  if (lib) return true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) { lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
  if (!lib) { err = "NCCL not found (dlopen libnccl.so.2): " + std::string(dlerror() ? dlerror() : ""); return false; }
  GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
  CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
  CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
  AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
  GroupStart = (decltype(GroupStart))dlsym(lib, "ncclGroupStart");
  GroupEnd = (decltype(GroupEnd))dlsym(lib, "ncclGroupEnd");
  GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
  if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllReduce || !GroupStart || !GroupEnd || !GetErrorString) {
    err = "NCCL symbols missing"; lib = nullptr; return false;
  }
  return true;
}

}  // namespace bn

#define BNUTS_EXEC bn::CudaExec
#include "capi_impl.h"

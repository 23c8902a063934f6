// hostemu.cpp — serial host build of the PRODUCT state machine (nuts_machine.h,
// backend.h, engine_core.h, capi_impl.h) behind the same C ABI.
//
// Test harness only: it lets the CPU test suite (`-m "not gpu"`) run the exact
// control code and vector-loop code of the CUDA engine against the oracle without
// a GPU.  It is never shipped or loaded by the product; libbnuts.so contains only
// the CUDA execution policy.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <omp.h>

#include "../../inplacedhmc.jl_b200/csrc/engine_core.h"

namespace bn {

struct HostExec {
  static constexpr bool has_tensor_path = false;
  struct Range { explicit Range(const char*) {} };   // tracing ranges exist on the device engine only
  int32_t init(int, std::string&) { return 0; }
  void shutdown() {}
  template <class U> U* alloc(size_t n) { return static_cast<U*>(std::calloc(n ? n : 1, sizeof(U))); }
  void free(void* p) { std::free(p); }
  void h2d(void* d, const void* s, size_t n) { std::memcpy(d, s, n); }
  void d2h(void* d, const void* s, size_t n) { std::memcpy(d, s, n); }
  void d2d(void* d, const void* s, size_t n) { std::memmove(d, s, n); }
  void rows_times_matrix(const double* in, double* out, int64_t R, int D, const double* mat) {
    for (int64_t r = 0; r < R; ++r)
      for (int j = 0; j < D; ++j) {
        double acc = 0.0;
        for (int k = 0; k < D; ++k) acc = std::fma(in[r * D + k], mat[size_t(k) * D + j], acc);
        out[r * D + j] = acc;
      }
  }
  void zero(void* d, size_t n) { std::memset(d, 0, n); }
  void sync() {}
  void use() {}
  int32_t check(std::string&) { return 0; }
  int32_t profile(int32_t, double* ms, int64_t* n) { if (ms) *ms = 0; if (n) *n = 0; return 0; }

  unsigned long long count = 0;
  unsigned long long* counter() { return &count; }
  int64_t read_count() { return (int64_t)count; }
  template <class T> void prepare(const EngineMem<T>& M, const RunParams<T>& rp, const PrepareArgs& a) {
    count = 0;
    for (int c = 0; c < M.C; ++c) prepare_chain(M, rp, a, c, SerialLanes{});
  }
  // pipelined run loop: on the host every count is available at once (same control flow, no lag)
  static constexpr int RING = 4;
  int64_t ring[RING] = {};
  template <class T> void advance_async(const EngineMem<T>& M, const RunParams<T>& rp, int iters, int slot, int) {
    ring[slot] = advance(M, rp, iters);
  }
  void reset_counters() {}
  bool count_ready(int) { return true; }
  int64_t count_wait(int slot) { return ring[slot]; }
  bool failed() const { return false; }
  template <class T> int64_t advance(const EngineMem<T>& M, const RunParams<T>& rp, int iters) {
    int64_t np = 0;
    count = 0;
#pragma omp parallel for schedule(dynamic) reduction(+ : np)
    for (int c = 0; c < M.C; ++c) np += advance_chain(M, rp, c, SerialLanes{}, iters) ? 1 : 0;
    return np;
  }
  // deterministic batched gradients, same summation order as the CUDA kernels
  // evaluates staged rows [0, rows); returns the number of partial blocks per row
  template <class E> int gradient(E& eng, int rows) {
    auto& M = eng.M;
    using T = typename std::remove_reference<decltype(*M.zs)>::type;
    const int C = rows, D = M.D, Dp = M.Dp;
    int nbr = 1;
    if (eng.model.kind == MODEL_GAUSSIAN) {
#pragma omp parallel for
      for (int c = 0; c < C; ++c)
        for (int d = 0; d < D; ++d) {
          T acc = T(0);
          for (int k = 0; k < D; ++k) acc = fma_(eng.model.P[size_t(d) * D + k], M.stage_q[size_t(c) * Dp + k], acc);
          M.stage_g[size_t(c) * Dp + d] = -acc;
        }
    } else if (eng.model.kind == MODEL_LOGISTIC) {
      const int64_t N = eng.model.N;
      const int nb = eng.model.row_blocks;
      nbr = nb;
      const int64_t R = (N + nb - 1) / nb;
#pragma omp parallel for collapse(2)
      for (int b = 0; b < nb; ++b)
        for (int c = 0; c < C; ++c) {
          const int64_t i0 = b * R, i1 = std::min<int64_t>(N, i0 + R);
          T* part = &M.stage_g[(size_t(b) * C + c) * Dp];
          const T* q = &M.stage_q[size_t(c) * Dp];
          for (int d = 0; d < D; ++d) part[d] = T(0);
          T pl = T(0);
          for (int64_t i = i0; i < i1; ++i) {
            const T* xr = &eng.model.X[size_t(i) * D];
            T eta = T(0);
            for (int d = 0; d < D; ++d) eta = fma_(xr[d], q[d], eta);
            T r, lt;
            logistic_elem(eta, eng.model.y[size_t(i)], &r, &lt);
            pl = pl + lt;
            for (int d = 0; d < D; ++d) part[d] = fma_(xr[d], r, part[d]);
          }
          M.stage_l[size_t(b) * C + c] = pl;
        }
    }
    return nbr;
  }
  template <class E> int32_t gauss_tensor_setup(E&, const std::vector<double>&, std::string& err) {
    err = "tensor path is CUDA-only";
    return BNUTS_ERR_UNSUPPORTED;
  }
  template <class E> int32_t logistic_tensor_setup(E&, const void*, int32_t, const double*, int64_t, std::string& err) {
    err = "tensor path is CUDA-only";
    return BNUTS_ERR_UNSUPPORTED;
  }
  int reference_mode() const { return 0; }
  int launches_per_gradient() const { return 1; }
  template <class E> int32_t logistic_tensor_setup_synth(E&, uint64_t, int64_t, int64_t, std::string& err) {
    err = "tensor path is CUDA-only";
    return BNUTS_ERR_UNSUPPORTED;
  }
  // ---- row-sharded mode (same semantics as the CUDA policy, serial loops)
  static int32_t nccl_unique_id(uint8_t*) { return BNUTS_ERR_UNSUPPORTED; }
  int32_t nccl_init(const uint8_t*, int, int, std::string& err) { err = "NCCL is CUDA-only"; return BNUTS_ERR_UNSUPPORTED; }
  int32_t p2p_export(size_t, size_t, uint8_t*, std::string& err) { err = "peer-memory exchange is CUDA-only"; return BNUTS_ERR_UNSUPPORTED; }
  int32_t p2p_connect(const uint8_t*, int, int, std::string& err) { err = "peer-memory exchange is CUDA-only"; return BNUTS_ERR_UNSUPPORTED; }
  bool p2p_failed() { return false; }
  template <class T> void fold_push(const EngineMem<T>&, int, uint64_t) {}
  template <class T> void wait_sum(const EngineMem<T>&, int, uint64_t, T*, double*) {}
  template <class T> int64_t assign_rows(const EngineMem<T>& V, const EngineMem<T>& Mc) {
    int64_t n = 0;
    for (int c = 0; c < V.C; ++c) {
      if (!V.stage_active[c]) continue;
      const int64_t r = n++;
      Mc.stage_row[c] = (int32_t)r;
      for (int d = 0; d < V.Dp; ++d) Mc.stage_q[r * V.Dp + d] = V.stage_q[size_t(c) * V.Dp + d];
      V.stage_active[c] = 0;
    }
    count = 0;
    return n;
  }
  template <class T> void fold_partials(const EngineMem<T>& M, int rows, T* red_g, double* red_l) {
    const int64_t bs = int64_t(rows) * M.Dp;
    for (int row = 0; row < rows; ++row) {
      for (int d = 0; d < M.Dp; ++d) {
        T acc = T(0);
        if (d < M.D) for (int b = 0; b < M.stage_nb; ++b) acc = acc + M.stage_g[b * bs + int64_t(row) * M.Dp + d];
        red_g[int64_t(row) * M.Dp + d] = acc;
      }
      double l = 0.0;
      for (int b = 0; b < M.stage_nb; ++b) l += double(M.stage_l[int64_t(b) * rows + row]);
      red_l[row] = l;
    }
  }
  int32_t allreduce(void* g, int64_t ng, bool g_is_f32, double* l, int64_t nl, bnuts_allreduce_fn fn, void* ctx, std::string& err) {
    if (!fn) { err = "row-sharded mode without a collective"; return BNUTS_ERR_INTERNAL; }
    if (fn(ctx, g, ng, g_is_f32 ? 1 : 0) != 0 || fn(ctx, l, nl, 0) != 0) { err = "host allreduce callback failed"; return BNUTS_ERR_INTERNAL; }
    return 0;
  }
  template <class E> int32_t logistic_reference(E&, const double*, std::string& err) {
    err = "tensor path is CUDA-only";
    return BNUTS_ERR_UNSUPPORTED;
  }
  template <class T> void metric_update(const EngineMem<T>& M, int N, double lambda) {
    for (int c = 0; c < M.C; ++c) metric_update_chain(M, c, N, lambda, SerialLanes{});
  }
  template <class T> void finish_da(const EngineMem<T>& M) {
    for (int c = 0; c < M.C; ++c)
      if (M.cs[c].status == 0) M.cs[c].eps = exp_(M.cs[c].da_logepsbar);
  }
  template <class T> void set_eps(const EngineMem<T>& M, const double* e) { for (int c = 0; c < M.C; ++c) M.cs[c].eps = e[c]; }
  template <class T> void get_eps(const EngineMem<T>& M, double* e) { for (int c = 0; c < M.C; ++c) e[c] = M.cs[c].eps; }
  template <class T> bool any_status(const EngineMem<T>& M, int32_t code) {
    for (int c = 0; c < M.C; ++c) if (M.cs[c].status == code) return true;
    return false;
  }
  template <class T> void get_status(const EngineMem<T>& M, int32_t* st) { for (int c = 0; c < M.C; ++c) st[c] = M.cs[c].status; }
  template <class T> void gather_state(const EngineMem<T>& M, double* o) {
    const size_t n = size_t(M.C) * M.D;
    for (int c = 0; c < M.C; ++c) {
      const int s = M.cs[c].slot_cur;
      const T* q = M.zs + (size_t(c) * M.S + s) * 3 * M.Dp;
      for (int d = 0; d < M.D; ++d) { o[size_t(c) * M.D + d] = double(q[d]); o[n + size_t(c) * M.D + d] = double(q[2 * M.Dp + d]); }
      o[2 * n + c] = double(M.zlq[size_t(c) * M.S + s]);
    }
  }
  template <class T> void totals(const EngineMem<T>& M, int64_t* tot) {
    tot[0] = tot[1] = tot[2] = 0;
    for (int c = 0; c < M.C; ++c) { tot[0] += M.cs[c].tot_leapfrogs; tot[1] += M.cs[c].tot_transitions; tot[2] += M.cs[c].tot_divergences; }
  }
};

}  // namespace bn

#define BNUTS_EXEC bn::HostExec
#include "../../inplacedhmc.jl_b200/csrc/capi_impl.h"

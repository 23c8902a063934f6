// backend.h — HBM data layout of the chain state and the D-vector primitives the
// state machine (nuts_machine.h) calls.
//
// One backend instance is bound to one chain.  All loops are written once over a
// "lane policy" LP:
//   device  (WarpLanes): lane l of the chain's warp handles d = l, l+32, ...;
//            reductions finish with a xor-butterfly of warp shuffles
//   tests   (SerialLanes): one host thread walks d = 0..D-1 and keeps 32 partial
//            accumulators indexed by d & 31, then emulates the same butterfly
// so both produce bit-identical sums (the "warp order" of bnuts_models.h).
//
// Layout (T = engine arithmetic type, Dp = D rounded up to 32, chain-major so a
// warp streams contiguous memory and GEMM staging is K-major):
//   zs      [C][S][3][Dp]  phase-point slots (q, p, ∇ℓ)      ≙ Tree z-slots, src/tree.jl:69-82
//   zlq     [C][S]         ℓ(q) per slot                      ≙ EvaluatedLogDensity.ℓq
//   st_rho  [C][L][Dp]     Σρ of pending left siblings        ≙ Σρ slots, src/tree.jl:95-105
//   st_psf  [C][L][Dp]     p♯ of their first-built leaf       ≙ ρ♯ slots, src/tree.jl:83-94
//   m_rho, m_psm, m_psp, ps_cur [C][Dp]   main-tree turn statistic, p♯ of the newest leaf
//   Minv, W [C][Dp]        GaussianKineticEnergy diagonals    ≙ src/hamiltonian.jl:33-38
#pragma once
#include "nuts_machine.h"

namespace bn {

template <class T> struct EngineMem {
  int32_t C, D, Dp, S, L;
  int32_t model_kind;
  T* zs; T* zlq; T* st_rho; T* st_psf;
  T* m_rho; T* m_psm; T* m_psp; T* ps_cur; T* Minv; T* W;
  ChainState<T>* cs;
  // gradient staging for models evaluated by a separate batched kernel
  T* stage_q;            // [C][Dp]   position to evaluate
  T* stage_g;            // [NB][C][Dp] gradient (partials over NB row blocks / splits)
  T* stage_l;            // [NB][C]   log-density partials
  double* stage_ld;      // [NB][C]   Float64 log-density partials (tensor path), or null
  int32_t stage_nb;      // NB
  uint16_t* stage_bh;    // [C][Dt]   bf16 high / middle / low 8 mantissa bits of q (tensor path), or null
  uint16_t* stage_bm;
  uint16_t* stage_bl;
  // active-chain compaction: a chain that requests a gradient takes the next free
  // staging row, so the batched kernels only see rows [0, stage_rows)
  int32_t* stage_row;            // [C] row of each chain in the current request
  unsigned long long* stage_count;  // rows handed out in the current lockstep step
  int32_t stage_rows;            // rows of the request being consumed (stride of the partials)
  // row-sharded data: every engine of the group runs the same chains and must hand
  // out the same rows, so requests are only flagged here (row = chain in "wide" staging) and a scan in chain
  // order assigns the compact rows afterwards; null in the ordinary (atomic counter) mode
  int32_t* stage_active;         // [C]
  int32_t Dt;            // padded K of the tensor path
  const T* beta_ref;     // [Dp] reference point of the tensor path (staged operand is q − beta_ref), or null
  const double* lin_w;   // [Dp] tensor path: the kernel's log-density partials omit ½ Σ_d lin_w[d] q[d]
  T tau;                 // logistic prior precision
  // per-call outputs
  double* draws;         // [C][N][D]
  double* bare_p_in;     // [C][D]
  double* bare_out;      // [4][C][D] q, p, g, (lq in first C entries of block 3)
  double* pos_in;        // [C][D] set_positions input (or null: Philox U[-2,2])
};

// bf16 round-to-nearest-even of a finite float, as raw bits
BN_HD uint16_t bf16_bits(float x) {
  const uint32_t u = f2u(x);
  const uint32_t r = u + 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(r >> 16);
}
BN_HD float bf16_val(uint16_t h) { return u2f((uint32_t)h << 16); }

// ------------------------------------------------------------------ lane policies
struct SerialLanes {
  static constexpr int NACC = 32;
  BN_HD int first() const { return 0; }
  BN_HD int stride() const { return 1; }
  BN_HD int acc(int d) const { return d & 31; }
  BN_HD bool lane0() const { return true; }
  BN_HD void sync() const {}
  BN_HD int alloc_row(unsigned long long* counter) const {
#if defined(__CUDA_ARCH__)
    return (int)atomicAdd(counter, 1ull);
#else
    return (int)__atomic_fetch_add(counter, 1ull, __ATOMIC_RELAXED);
#endif
  }
  template <class T> BN_HD T reduce(T* part) const {
    for (int off = 16; off >= 1; off >>= 1) {
      T nw[32];
      for (int l = 0; l < 32; ++l) nw[l] = part[l] + part[l ^ off];
      for (int l = 0; l < 32; ++l) part[l] = nw[l];
    }
    return part[0];
  }
};
#if defined(__CUDACC__)
struct WarpLanes {
  static constexpr int NACC = 1;
  int lane;
  BN_HD int first() const { return lane; }
  BN_HD int stride() const { return 32; }
  BN_HD int acc(int) const { return 0; }
  BN_HD bool lane0() const { return lane == 0; }
  BN_HD void sync() const {
#if defined(__CUDA_ARCH__)
    __syncwarp();
#endif
  }
  BN_HD int alloc_row(unsigned long long* counter) const {
    int r = 0;
#if defined(__CUDA_ARCH__)
    if (lane == 0) r = (int)atomicAdd(counter, 1ull);
    r = __shfl_sync(0xffffffffu, r, 0);
#endif
    return r;
  }
  template <class T> BN_HD T reduce(T* part) const {
    T v = part[0];
#if defined(__CUDA_ARCH__)
    for (int off = 16; off >= 1; off >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, off);
#endif
    return v;
  }
};
#endif

template <class T, class LP> struct Backend {
  const EngineMem<T>& M;
  int32_t c;
  LP lp;
  const RunParams<T>& rp;

  BN_HD Backend(const EngineMem<T>& M_, int32_t c_, LP lp_, const RunParams<T>& rp_) : M(M_), c(c_), lp(lp_), rp(rp_) {}

  BN_HD int dim() const { return M.D; }
  BN_HD bool lane0() const { return lp.lane0(); }
  BN_HD void sync() const { lp.sync(); }
  BN_HD T* zq(int s) const { return M.zs + ((int64_t)c * M.S + s) * 3 * M.Dp; }
  BN_HD T* zp(int s) const { return zq(s) + M.Dp; }
  BN_HD T* zg(int s) const { return zq(s) + 2 * M.Dp; }
  BN_HD T* strho(int k) const { return M.st_rho + ((int64_t)c * M.L + k) * M.Dp; }
  BN_HD T* stpsf(int k) const { return M.st_psf + ((int64_t)c * M.L + k) * M.Dp; }
  BN_HD T* cv(T* base) const { return base + (int64_t)c * M.Dp; }

  BN_HD T get_lq(int s) const { return M.zlq[(int64_t)c * M.S + s]; }
  BN_HD void set_lq(int s, T v) const {
    if (lp.lane0()) M.zlq[(int64_t)c * M.S + s] = v;
    lp.sync();
  }

  // momentum refresh (≙ rand_p!, src/kinetic_energy.jl:63) + p♯, K (≙ :14-24, :39-46)
  // + initial turn statistic (≙ leaf_turn_statistic, src/NUTS.jl:113-116)
  BN_HD void start_tx(int slot, const double* inj_p, uint64_t seed, uint32_t gchain, uint32_t t, T* Ksum) const {
    T* p = zp(slot);
    const T* W = cv(M.W);
    const T* Mi = cv(M.Minv);
    T* ps = cv(M.ps_cur); T* mr = cv(M.m_rho); T* mm = cv(M.m_psm); T* mp = cv(M.m_psp);
    T part[LP::NACC];
    for (int i = 0; i < LP::NACC; ++i) part[i] = T(0);
    for (int d = lp.first(); d < M.D; d += lp.stride()) {
      const T pd = inj_p ? T(inj_p[d]) : W[d] * std_normal(seed, gchain, t, (uint32_t)d, T(0));
      const T psd = Mi[d] * pd;
      p[d] = pd; ps[d] = psd; mr[d] = pd; mm[d] = psd; mp[d] = psd;
      T& a = part[lp.acc(d)];
      a = fma_(psd, pd, a);
    }
    *Ksum = lp.reduce(part);
    lp.sync();
  }

  // take a staging row for this chain's gradient request (batched targets only)
  BN_HD int take_row() const {
    if (!M.stage_q) return -1;
    if (M.stage_active) {
      if (lp.lane0()) M.stage_active[c] = 1;
      return c;
    }
    const int row = lp.alloc_row(M.stage_count);
    if (lp.lane0()) M.stage_row[c] = row;
    return row;
  }
  // publish one coordinate of the position to evaluate; the tensor path also gets the
  // exact three-term bf16 split of the fp32 value (3 x 8 mantissa bits)
  BN_HD void stage_put(int row, int d, T qd) const {
    M.stage_q[(int64_t)row * M.Dp + d] = qd;
    if (M.stage_bh) {
      const float qf = (float)qd - (M.beta_ref ? (float)M.beta_ref[d] : 0.f);
      const uint16_t h = bf16_bits(qf);
      const float r1 = qf - bf16_val(h);
      const uint16_t m = bf16_bits(r1);
      const int64_t o = (int64_t)row * M.Dt + d;
      M.stage_bh[o] = h; M.stage_bm[o] = m; M.stage_bl[o] = bf16_bits(r1 - bf16_val(m));
    }
  }

  // ≙ src/kinetic_energy.jl:144-150: pₘ = p + ½ϵ∇ℓ ; q′ = q + ϵ M⁻¹ pₘ
  BN_HD void pre_kick_drift(int src, int dst, T eh, T eps) const {
    const T* q = zq(src); const T* p = zp(src); const T* g = zg(src);
    T* qn = zq(dst); T* pn = zp(dst);
    const T* Mi = cv(M.Minv);
    const int row = take_row();
    for (int d = lp.first(); d < M.D; d += lp.stride()) {
      const T pm = fma_(eh, g[d], p[d]);
      const T qd = fma_(eps * Mi[d], pm, q[d]);
      pn[d] = pm; qn[d] = qd;
      if (row >= 0) stage_put(row, d, qd);
    }
    lp.sync();
  }

  // ≙ src/kinetic_energy.jl:159-161 (p′ = pₘ + ½ϵ∇ℓ′), p♯ (:39-46), Σ p♯p (:19-23);
  // push_level >= 0 also stores (ρ, p♯first) of a new stack entry
  BN_HD T post_kick(int slot, T eh, int push_level) const {
    T* p = zp(slot); const T* g = zg(slot);
    const T* Mi = cv(M.Minv);
    T* ps = cv(M.ps_cur);
    T* pr = push_level >= 0 ? strho(push_level) : nullptr;
    T* pf = push_level >= 0 ? stpsf(push_level) : nullptr;
    T part[LP::NACC];
    for (int i = 0; i < LP::NACC; ++i) part[i] = T(0);
    for (int d = lp.first(); d < M.D; d += lp.stride()) {
      const T pd = fma_(eh, g[d], p[d]);
      const T psd = Mi[d] * pd;
      p[d] = pd; ps[d] = psd;
      if (pr) { pr[d] = pd; pf[d] = psd; }
      T& a = part[lp.acc(d)];
      a = fma_(psd, pd, a);
    }
    const T r = lp.reduce(part);
    lp.sync();
    return r;
  }

  // ≙ combine_turn_statistics + is_turning dots (src/NUTS.jl:139-158) for stack level L
  BN_HD void merge_sub(int L, int rhoR_is_leaf, int slot_leaf, bool fwd, T* dm, T* dp) const {
    T* rl = strho(L);
    const T* rr = rhoR_is_leaf ? zp(slot_leaf) : strho(L + 1);
    const T* pf = stpsf(L);
    const T* pc = cv(M.ps_cur);
    const T* psm = fwd ? pf : pc;
    const T* psp = fwd ? pc : pf;
    T am[LP::NACC], ap[LP::NACC];
    for (int i = 0; i < LP::NACC; ++i) am[i] = ap[i] = T(0);
    for (int d = lp.first(); d < M.D; d += lp.stride()) {
      const T r = fwd ? rl[d] + rr[d] : rr[d] + rl[d];
      rl[d] = r;
      T& a = am[lp.acc(d)]; a = fma_(r, psm[d], a);
      T& b2 = ap[lp.acc(d)]; b2 = fma_(r, psp[d], b2);
    }
    *dm = lp.reduce(am);
    *dp = lp.reduce(ap);
    lp.sync();
  }

  // same for the main tree (src/tree.jl:437-438); the new edge's p♯ replaces p♯₊ (fwd) or p♯₋
  BN_HD void merge_top(int rhoR_is_leaf, int slot_leaf, bool fwd, T* dm, T* dp) const {
    T* mr = cv(M.m_rho);
    const T* rr = rhoR_is_leaf ? zp(slot_leaf) : strho(0);
    T* mm = cv(M.m_psm); T* mp = cv(M.m_psp);
    const T* pc = cv(M.ps_cur);
    T am[LP::NACC], ap[LP::NACC];
    for (int i = 0; i < LP::NACC; ++i) am[i] = ap[i] = T(0);
    for (int d = lp.first(); d < M.D; d += lp.stride()) {
      const T r = fwd ? mr[d] + rr[d] : rr[d] + mr[d];
      mr[d] = r;
      const T pcd = pc[d];
      const T psm = fwd ? mm[d] : pcd;
      const T psp = fwd ? pcd : mp[d];
      if (fwd) mp[d] = pcd; else mm[d] = pcd;
      T& a = am[lp.acc(d)]; a = fma_(r, psm, a);
      T& b2 = ap[lp.acc(d)]; b2 = fma_(r, psp, b2);
    }
    *dm = lp.reduce(am);
    *dp = lp.reduce(ap);
    lp.sync();
  }

  // Σ a[d] b[d] in warp order
  BN_HD T dot(const T* a, const T* b) const {
    T part[LP::NACC];
    for (int i = 0; i < LP::NACC; ++i) part[i] = T(0);
    for (int d = lp.first(); d < M.D; d += lp.stride()) {
      T& x = part[lp.acc(d)];
      x = fma_(a[d], b[d], x);
    }
    return lp.reduce(part);
  }

  // ≙ logdensity_and_gradient! (call site src/kinetic_energy.jl:73); SURVEY.md §A.4 targets.
  // Elementwise targets are evaluated here; batched targets were evaluated by a
  // separate kernel into the staging buffers and are finalised here.
  BN_HD T model_grad(int slot) const {
    const T* q = zq(slot);
    T* g = zg(slot);
    T l;
    switch (M.model_kind) {
      case MODEL_IID_NORMAL: {
        for (int d = lp.first(); d < M.D; d += lp.stride()) g[d] = iid_grad(q[d]);
        l = iid_value(dot(q, q));
        break;
      }
      case MODEL_FUNNEL: {
        T part[LP::NACC];
        for (int i = 0; i < LP::NACC; ++i) part[i] = T(0);
        for (int d = lp.first(); d < M.D; d += lp.stride()) {
          if (d >= 1) { T& x = part[lp.acc(d)]; x = fma_(q[d], q[d], x); }
        }
        const T S = lp.reduce(part);
        const T v = q[0], e = exp_(-v);
        for (int d = lp.first(); d < M.D; d += lp.stride())
          g[d] = (d == 0) ? funnel_grad_v(v, S, e, M.D) : funnel_grad_x(q[d], e);
        l = funnel_value(v, S, e, M.D);
        break;
      }
      case MODEL_GAUSSIAN: {
        const T* sg = M.stage_g + (int64_t)M.stage_row[c] * M.Dp;
        for (int d = lp.first(); d < M.D; d += lp.stride()) g[d] = sg[d];
        lp.sync();
        l = T(0.5) * dot(q, g);
        break;
      }
      case MODEL_LOGISTIC: {
        const int64_t row = M.stage_row[c], rows = M.stage_rows;
        const int64_t bs = rows * M.Dp;
        const T* sg = M.stage_g + row * M.Dp;
        for (int d = lp.first(); d < M.D; d += lp.stride()) {
          T acc = T(0);
          for (int b = 0; b < M.stage_nb; ++b) acc = acc + sg[b * bs + d];
          g[d] = fma_(-M.tau, q[d], acc);
        }
        T ls = T(0);
        if (M.stage_ld) {  // tensor path: partials are ~1e5 in magnitude, summed in Float64
          double lsd = 0.0;
          for (int b = 0; b < M.stage_nb; ++b) lsd = lsd + M.stage_ld[b * rows + row];
          if (M.lin_w) {  // linear part of Σ log σ(η̃): ½ Σ_i η̃_i = ½ colsum(X̃)·q
            double part[LP::NACC];
            for (int i = 0; i < LP::NACC; ++i) part[i] = 0.0;
            for (int d = lp.first(); d < M.D; d += lp.stride()) {
              double& a = part[lp.acc(d)];
              a = fma_(M.lin_w[d], (double)q[d], a);
            }
            lsd = fma_(0.5, lp.reduce(part), lsd);
          }
          ls = T(lsd);
        } else {
          for (int b = 0; b < M.stage_nb; ++b) ls = ls + M.stage_l[b * rows + row];
        }
        l = fma_(T(-0.5) * M.tau, dot(q, q), ls);
        break;
      }
      default: l = lim<T>::nan();
    }
    lp.sync();
    return l;
  }

  // ---- local optimum search (≙ warmup!(FindLocalOptimum), src/warmup.jl:152-186; see nuts_machine.h)
  BN_HD void opt_trial(int src, int dst, T alpha, T lambda) const {
    const T* q = zq(src); const T* g = zg(src);
    T* qn = zq(dst);
    const int row = take_row();
    for (int d = lp.first(); d < M.D; d += lp.stride()) {
      const T dv = fma_(-lambda, q[d], g[d]);
      const T qd = fma_(alpha, dv, q[d]);
      qn[d] = qd;
      if (row >= 0) stage_put(row, d, qd);
    }
    lp.sync();
  }
  BN_HD void opt_norms(int slot, T lambda, T* dd, T* qq) const {
    const T* q = zq(slot); const T* g = zg(slot);
    T a[LP::NACC], c2[LP::NACC];
    for (int i = 0; i < LP::NACC; ++i) a[i] = c2[i] = T(0);
    for (int d = lp.first(); d < M.D; d += lp.stride()) {
      const T dv = fma_(-lambda, q[d], g[d]);
      T& x = a[lp.acc(d)]; x = fma_(dv, dv, x);
      T& y = c2[lp.acc(d)]; y = fma_(q[d], q[d], y);
    }
    *dd = lp.reduce(a);
    *qq = lp.reduce(c2);
  }
  BN_HD void opt_dots(int src, int dst, T lambda, T* dd_new, T* qq_new, T* dod) const {
    const T* q0 = zq(src); const T* g0 = zg(src);
    const T* q1 = zq(dst); const T* g1 = zg(dst);
    T a[LP::NACC], c2[LP::NACC], e[LP::NACC];
    for (int i = 0; i < LP::NACC; ++i) a[i] = c2[i] = e[i] = T(0);
    for (int d = lp.first(); d < M.D; d += lp.stride()) {
      const T d0 = fma_(-lambda, q0[d], g0[d]);
      const T d1 = fma_(-lambda, q1[d], g1[d]);
      T& x = a[lp.acc(d)]; x = fma_(d1, d1, x);
      T& y = c2[lp.acc(d)]; y = fma_(q1[d], q1[d], y);
      T& z = e[lp.acc(d)]; z = fma_(d0, d1, z);
    }
    *dd_new = lp.reduce(a);
    *qq_new = lp.reduce(c2);
    *dod = lp.reduce(e);
  }
  BN_HD void opt_restart_position(int slot, uint64_t seed, uint32_t gchain, uint32_t attempt) const {
    T* q = zq(slot);
    const int row = take_row();
    for (int d = lp.first(); d < M.D; d += lp.stride()) {
      const T qd = T(restart_position(seed, gchain, attempt, (uint32_t)d));
      q[d] = qd;
      if (row >= 0) stage_put(row, d, qd);
    }
    lp.sync();
  }

  // ≙ copyto!(chain[:, n], z.Q.q), src/warmup.jl:299,326
  BN_HD void emit_draw(int slot, int n) const {
    if (!M.draws) return;
    const T* q = zq(slot);
    double* o = M.draws + ((int64_t)c * rp.n_total + n) * M.D;
    for (int d = lp.first(); d < M.D; d += lp.stride()) o[d] = (double)q[d];
  }
  BN_HD void bare_load_p(int slot) const {
    T* p = zp(slot);
    const double* pi = M.bare_p_in + (int64_t)c * M.D;
    for (int d = lp.first(); d < M.D; d += lp.stride()) p[d] = T(pi[d]);
    lp.sync();
  }
  BN_HD void bare_emit(int slot) const {
    const int64_t blk = (int64_t)M.C * M.D;
    double* o = M.bare_out + (int64_t)c * M.D;
    const T* q = zq(slot); const T* p = zp(slot); const T* g = zg(slot);
    for (int d = lp.first(); d < M.D; d += lp.stride()) {
      o[d] = (double)q[d]; o[blk + d] = (double)p[d]; o[2 * blk + d] = (double)g[d];
    }
    if (lp.lane0()) M.bare_out[3 * blk + c] = (double)get_lq(slot);
  }
  // set_positions: q -> slot (and staging); ≙ src/warmup.jl:119 / random_position! :73
  BN_HD void load_position(int slot, uint64_t seed, uint32_t gchain) const {
    T* q = zq(slot);
    const int row = take_row();
    for (int d = lp.first(); d < M.D; d += lp.stride()) {
      const T qd = M.pos_in ? T(M.pos_in[(int64_t)c * M.D + d]) : T(init_position(seed, gchain, (uint32_t)d));
      q[d] = qd;
      if (row >= 0) stage_put(row, d, qd);
    }
    lp.sync();
  }
};

// ------------------------------------------------------------------ per-chain entry points
enum RunMode : int32_t { MODE_SAMPLE = 0, MODE_SEARCH = 1, MODE_BARE = 2, MODE_EVAL = 3, MODE_OPT = 4 };

struct PrepareArgs {
  int32_t mode;
  int32_t N;          // transitions (SAMPLE) or leapfrogs (BARE)
  uint32_t t0;        // first transition counter of the call
  const double* bare_eps;  // [C] signed step (BARE)
  const double* eps_in;    // [C] or null: overwrite s.eps before the call
};

// set up one chain for a call; ≙ the loop prologues of warmup!/mcmc! (src/warmup.jl:283-287, :321-323)
template <class T, class LP>
BN_HD void prepare_chain(const EngineMem<T>& M, const RunParams<T>& rp, const PrepareArgs& a, int32_t c, LP lp) {
  ChainState<T>& s = M.cs[c];
  Backend<T, LP> b(M, c, lp, rp);
  if (a.mode == MODE_EVAL) {
    b.load_position(0, rp.seed, (uint32_t)(rp.chain_offset + c));
    if (lp.lane0()) {
      s.slot_cur = 0; s.slot_minus = s.slot_plus = s.slot_zeta = 0; s.sp = 0;
      s.status = 0; s.phase = PH_EVAL_LEAF;
    }
    return;
  }
  if (!lp.lane0()) return;
  if (a.eps_in) s.eps = a.eps_in[c];
  s.n_done = 0; s.t = a.t0; s.sp = 0;
  s.slot_minus = s.slot_plus = s.slot_zeta = s.slot_cur;
  if (a.mode == MODE_SAMPLE) {
    s.remaining = a.N;
    s.phase = (s.status == 0 && a.N > 0) ? PH_START : PH_IDLE;
    if (rp.da_on) da_init(s);
  } else if (a.mode == MODE_SEARCH) {
    s.phase = (s.status == 0) ? PH_SEARCH_START : PH_IDLE;
  } else if (a.mode == MODE_OPT) {
    s.phase = (s.status == 0 || s.status == ST_NONFINITE_START) ? PH_OPT_START : PH_IDLE;
  } else if (a.mode == MODE_BARE) {
    s.remaining = a.N;
    s.ss_try = a.bare_eps[c];
    s.phase = PH_BARE_START;
  }
}

// advance one chain: consume a pending gradient, run to the next gradient request.
// `max_iters` > 1 lets elementwise targets run many leapfrogs inside one launch.
// Returns true if a gradient is pending for this chain on exit.
template <class T, class LP>
BN_HD bool advance_chain(const EngineMem<T>& M, const RunParams<T>& rp, int32_t c, LP lp, int max_iters) {
  ChainState<T>* g = &M.cs[c];
  if (g->phase == PH_IDLE) return false;
  ChainState<T> s;
  // copy the scalar head (everything before the stack arrays) into registers
  {
    const int nwords = (int)(offsetof(ChainState<T>, st_omega) / 4);
    const uint32_t* src = reinterpret_cast<const uint32_t*>(g);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&s);
    for (int i = 0; i < nwords; ++i) dst[i] = src[i];
  }
  Backend<T, LP> b(M, c, lp, rp);
  Machine<T, Backend<T, LP>> m(b, s, g, rp, c);
  bool pending = false;
  for (int it = 0; it < max_iters; ++it) {
    pending = m.step();
    if (!pending) break;
  }
  lp.sync();
  if (lp.lane0()) {
    const int nwords = (int)(offsetof(ChainState<T>, st_omega) / 4);
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&s);
    uint32_t* dst = reinterpret_cast<uint32_t*>(g);
    for (int i = 0; i < nwords; ++i) dst[i] = src[i];
  }
  return pending;
}

// ≙ GaussianKineticEnergy!(κ, chain, λ), src/hamiltonian.jl:77-101,153-162, over this
// chain's N draws of the stage, followed by W = 1/sqrt(M⁻¹)
template <class T, class LP>
BN_HD void metric_update_chain(const EngineMem<T>& M, int32_t c, int32_t N, double lambda, LP lp) {
  if (M.cs[c].status != 0) return;
  const double Nf = (double)N;
  const double Ninv = 1.0 / Nf;
  const double mulreg = Nf / ((Nf + lambda) * (Nf - 1.0));
  const double addreg = 1e-3 * lambda / (Nf + lambda);
  const double* dr = M.draws + (int64_t)c * N * M.D;
  for (int d = lp.first(); d < M.D; d += lp.stride()) {
    const double mu = dr[d];
    double sd = 0.0, sd2 = 0.0;
    for (int n = 1; n < N; ++n) {
      const double dl = dr[(int64_t)n * M.D + d] - mu;
      sd = dl + sd;
      sd2 = fma_(dl, dl, sd2);
    }
    const double s2nm1 = fma_(-(sd * sd), Ninv, sd2);
    const double reg = fma_(s2nm1, mulreg, addreg);
    M.Minv[(int64_t)c * M.Dp + d] = T(reg);
    M.W[(int64_t)c * M.Dp + d] = T(1.0 / sqrt_(reg));
  }
}

}  // namespace bn

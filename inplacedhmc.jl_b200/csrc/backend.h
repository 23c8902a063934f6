// backend.h — HBM data layout of the chain state and the D-vector primitives the
// state machine (nuts_machine.h) calls.
//
// One backend instance is bound to one chain.  All loops are written once over a
// "lane policy" LP:
//   device  (WarpLanes): lane l of the chain's warp handles the groups of four consecutive
//            elements d = 4l..4l+3 (+128, +256, ...) with 16/32-byte loads and stores;
//            reductions finish with a xor-butterfly of warp shuffles
//   tests   (SerialLanes): one host thread walks d = 0..D-1 and keeps 32 partial
//            accumulators indexed by (d >> 2) & 31, then emulates the same butterfly
// so both produce bit-identical sums (the "warp order" of bnuts_models.h).
//
// Layout (T = engine arithmetic type, Dp = D rounded up to 32, chain-major so a
// warp streams contiguous memory and GEMM staging is K-major):
//   zs      [C][S][3][Dp]  phase-point slots (q, p, ∇ℓ)      ≙ Tree z-slots, src/tree.jl:69-82
//   zlq     [C][S]         ℓ(q) per slot, Float64 for both T  ≙ EvaluatedLogDensity.ℓq
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
  T* zs; double* zlq; T* st_rho; T* st_psf;
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
  const double* grad0;   // [Dp] tensor path, single-term residual / remainder modes: the kernel's gradient partials omit X̃ᵀ·r0, or null
  const float* lin_H;    // [D][Dp] tensor path, remainder mode (logistic_rm.cu): H0 = X̃ᵀ diag(σ'(η̃0)) X̃; the kernel's partials also omit
                         // −H0 (q − beta_ref) and its log density is the remainder beyond ell0 + g0·δ − ½ δᵀH0 δ; or null
  double ell0;           // Σ_i log σ(η̃0_i) (remainder mode)
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
  BN_HD int first4() const { return 0; }
  BN_HD int stride4() const { return 4; }
  BN_HD int acc(int d) const { return (d >> 2) & 31; }
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
  BN_HD int first4() const { return 4 * lane; }
  BN_HD int stride4() const { return 128; }
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

// Groups of four consecutive elements.  Every D-vector of the engine is padded to Dp (a multiple of 32) and
// starts 128-byte aligned, so a whole group can always be LOADED (elements >= D are padding and are never
// used); stores of a partial last group fall back to scalar stores of the valid elements.
template <class T> BN_HD void ld4(const T* p, T (&v)[4]) {
#if defined(__CUDA_ARCH__)
  if constexpr (sizeof(T) == 4) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
    const double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  }
#else
  for (int e = 0; e < 4; ++e) v[e] = p[e];
#endif
}
template <class T> BN_HD void st4(T* p, const T (&v)[4], int nv) {
#if defined(__CUDA_ARCH__)
  if (nv == 4) {
    if constexpr (sizeof(T) == 4) {
      *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
      *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
      *reinterpret_cast<double2*>(p + 2) = make_double2(v[2], v[3]);
    }
    return;
  }
#endif
  for (int e = 0; e < nv; ++e) p[e] = v[e];
}
#define BN_FOR4(d0, nv) \
  for (int d0 = lp.first4(), nv = (M.D - d0 < 4 ? M.D - d0 : 4); d0 < M.D; d0 += lp.stride4(), nv = (M.D - d0 < 4 ? M.D - d0 : 4))

// RM = 1: the kernel instance that serves the remainder mode of the tensor-core logistic path (logistic_rm.cu): only it carries
// the D x D linear part (its registers and code would otherwise be charged to every target's state machine: measured −20 % on
// the Gaussian and funnel configurations when it was compiled into the one kernel)
template <class T, class LP, int RM = 0> struct Backend {
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

  BN_HD double get_lq(int s) const { return M.zlq[(int64_t)c * M.S + s]; }
  BN_HD void set_lq(int s, double v) const {
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
    BN_FOR4(d0, nv) {
      T w[4], mi[4], pv[4], sv[4];
      ld4(W + d0, w); ld4(Mi + d0, mi);
      for (int e = 0; e < 4; ++e) {
        const int d = d0 + e;
        pv[e] = T(0); sv[e] = T(0);
        if (e < nv) {
          pv[e] = inj_p ? T(inj_p[d]) : w[e] * std_normal(seed, gchain, t, (uint32_t)d, T(0));
          sv[e] = mi[e] * pv[e];
          T& a = part[lp.acc(d)];
          a = fma_(sv[e], pv[e], a);
        }
      }
      st4(p + d0, pv, nv); st4(ps + d0, sv, nv); st4(mr + d0, pv, nv); st4(mm + d0, sv, nv); st4(mp + d0, sv, nv);
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
  BN_HD void stage_put4(int row, int d0, const T (&qv)[4], int nv) const {
    st4(M.stage_q + (int64_t)row * M.Dp + d0, qv, nv);
    if (M.stage_bh) {
      T br[4] = {T(0), T(0), T(0), T(0)};
      if (M.beta_ref) ld4(M.beta_ref + d0, br);
      uint16_t h[4], m[4], l[4];
      for (int e = 0; e < 4; ++e) {
        const float qf = (float)qv[e] - (float)br[e];
        h[e] = bf16_bits(qf);
        const float r1 = qf - bf16_val(h[e]);
        m[e] = bf16_bits(r1);
        l[e] = bf16_bits(r1 - bf16_val(m[e]));
      }
      const int64_t o = (int64_t)row * M.Dt + d0;
      if (nv == 4) {   // 8-byte stores (Dt is a multiple of 64, d0 of 4)
        *reinterpret_cast<uint64_t*>(M.stage_bh + o) = (uint64_t)h[0] | ((uint64_t)h[1] << 16) | ((uint64_t)h[2] << 32) | ((uint64_t)h[3] << 48);
        *reinterpret_cast<uint64_t*>(M.stage_bm + o) = (uint64_t)m[0] | ((uint64_t)m[1] << 16) | ((uint64_t)m[2] << 32) | ((uint64_t)m[3] << 48);
        *reinterpret_cast<uint64_t*>(M.stage_bl + o) = (uint64_t)l[0] | ((uint64_t)l[1] << 16) | ((uint64_t)l[2] << 32) | ((uint64_t)l[3] << 48);
      } else {
        for (int e = 0; e < nv; ++e) { M.stage_bh[o + e] = h[e]; M.stage_bm[o + e] = m[e]; M.stage_bl[o + e] = l[e]; }
      }
    }
  }

  // ≙ src/kinetic_energy.jl:144-150: pₘ = p + ½ϵ∇ℓ ; q′ = q + ϵ M⁻¹ pₘ
  BN_HD void pre_kick_drift(int src, int dst, T eh, T eps) const {
    const T* q = zq(src); const T* p = zp(src); const T* g = zg(src);
    T* qn = zq(dst); T* pn = zp(dst);
    const T* Mi = cv(M.Minv);
    const int row = take_row();
    BN_FOR4(d0, nv) {
      T qv[4], pv[4], gv[4], mi[4], pm[4], qd[4];
      ld4(q + d0, qv); ld4(p + d0, pv); ld4(g + d0, gv); ld4(Mi + d0, mi);
      for (int e = 0; e < 4; ++e) {
        pm[e] = fma_(eh, gv[e], pv[e]);
        qd[e] = fma_(eps * mi[e], pm[e], qv[e]);
      }
      st4(pn + d0, pm, nv); st4(qn + d0, qd, nv);
      if (row >= 0) stage_put4(row, d0, qd, nv);
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
    BN_FOR4(d0, nv) {
      T pv[4], gv[4], mi[4], pd[4], sd[4];
      ld4(p + d0, pv); ld4(g + d0, gv); ld4(Mi + d0, mi);
      for (int e = 0; e < 4; ++e) {
        pd[e] = fma_(eh, gv[e], pv[e]);
        sd[e] = mi[e] * pd[e];
        if (e < nv) { T& a = part[lp.acc(d0 + e)]; a = fma_(sd[e], pd[e], a); }
      }
      st4(p + d0, pd, nv); st4(ps + d0, sd, nv);
      if (pr) { st4(pr + d0, pd, nv); st4(pf + d0, sd, nv); }
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
    BN_FOR4(d0, nv) {
      T a4[4], b4[4], m4[4], p4[4], r4[4];
      ld4(rl + d0, a4); ld4(rr + d0, b4); ld4(psm + d0, m4); ld4(psp + d0, p4);
      for (int e = 0; e < 4; ++e) {
        r4[e] = fwd ? a4[e] + b4[e] : b4[e] + a4[e];
        if (e < nv) {
          T& a = am[lp.acc(d0 + e)]; a = fma_(r4[e], m4[e], a);
          T& b2 = ap[lp.acc(d0 + e)]; b2 = fma_(r4[e], p4[e], b2);
        }
      }
      st4(rl + d0, r4, nv);
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
    BN_FOR4(d0, nv) {
      T a4[4], b4[4], c4[4], o4[4], r4[4];
      ld4(mr + d0, a4); ld4(rr + d0, b4); ld4(pc + d0, c4); ld4((fwd ? mm : mp) + d0, o4);
      for (int e = 0; e < 4; ++e) {
        r4[e] = fwd ? a4[e] + b4[e] : b4[e] + a4[e];
        const T psm = fwd ? o4[e] : c4[e];
        const T psp = fwd ? c4[e] : o4[e];
        if (e < nv) {
          T& a = am[lp.acc(d0 + e)]; a = fma_(r4[e], psm, a);
          T& b2 = ap[lp.acc(d0 + e)]; b2 = fma_(r4[e], psp, b2);
        }
      }
      st4(mr + d0, r4, nv);
      st4((fwd ? mp : mm) + d0, c4, nv);
    }
    *dm = lp.reduce(am);
    *dp = lp.reduce(ap);
    lp.sync();
  }

  // Σ a[d] b[d] in warp order
  BN_HD T dot(const T* a, const T* b) const {
    T part[LP::NACC];
    for (int i = 0; i < LP::NACC; ++i) part[i] = T(0);
    BN_FOR4(d0, nv) {
      T a4[4], b4[4];
      ld4(a + d0, a4); ld4(b + d0, b4);
      for (int e = 0; e < nv; ++e) { T& x = part[lp.acc(d0 + e)]; x = fma_(a4[e], b4[e], x); }
    }
    return lp.reduce(part);
  }

  // ≙ logdensity_and_gradient! of the logistic target, second half: the batched kernel has written partial gradients / log
  // densities for this chain's staging row; fold them, add the prior (and, in the remainder mode of the tensor path, the
  // exact linear / quadratic part).  A real function: its registers must not be charged to the state machine of the other targets.
  template <bool WITH_LIN> BN_HD double logistic_finalize_t(const T* q, T* g) const {
    double l;
    const int64_t row = M.stage_row[c], rows = M.stage_rows;
    const int64_t bs = rows * M.Dp;
    const T* sg = M.stage_g + row * M.Dp;
    T rem4[2][4] = {{T(0), T(0), T(0), T(0)}, {T(0), T(0), T(0), T(0)}};   // remainder mode (D <= 256: at most two groups per lane): the folded X̃ᵀρ of this lane's coordinates
    BN_FOR4(d0, nv) {
      T acc[4] = {T(0), T(0), T(0), T(0)}, qv[4], gv[4];
      // the partial blocks are added in the order b = 0, 1, ... (fixed: the deterministic path is bit-identical to the
      // oracle), but loaded four at a time: a small launch of the tensor path has up to 148 of them (one per SM) and one
      // L2 round trip per partial made this fold the longest part of a lockstep step with few active chains
      int b = 0;
      for (; b + 4 <= M.stage_nb; b += 4) {
        T p0[4], p1[4], p2[4], p3[4];
        ld4(sg + (b + 0) * bs + d0, p0); ld4(sg + (b + 1) * bs + d0, p1); ld4(sg + (b + 2) * bs + d0, p2); ld4(sg + (b + 3) * bs + d0, p3);
        for (int e = 0; e < 4; ++e) acc[e] = (((acc[e] + p0[e]) + p1[e]) + p2[e]) + p3[e];
      }
      for (; b < M.stage_nb; ++b) {
        T pv[4];
        ld4(sg + b * bs + d0, pv);
        for (int e = 0; e < 4; ++e) acc[e] = acc[e] + pv[e];
      }
      if constexpr (WITH_LIN) { for (int e = 0; e < 4; ++e) rem4[(d0 >> 7) & 1][e] = acc[e]; }
      if (M.grad0) {  // constant part of the gradient about the reference point (see k_logistic_tc)
        double g0[4];
        ld4(M.grad0 + d0, g0);
        for (int e = 0; e < 4; ++e) acc[e] = T((double)acc[e] + g0[e]);
      }
      ld4(q + d0, qv);
      for (int e = 0; e < 4; ++e) gv[e] = fma_(-M.tau, qv[e], acc[e]);
      st4(g + d0, gv, nv);
    }
    double lin_l = 0.0;
    if constexpr (WITH_LIN) { if (M.lin_H) lin_l = linear_part(q, g, rem4, sg, bs); }   // remainder mode: g −= H0 δ, ℓ += ell0 + δ·(g0 − ½ H0 δ + ⅓ X̃ᵀρ)
    if (M.stage_ld) {  // tensor path: partials are ~1e5..1e7 in magnitude, summed and kept in Float64
      // Float64 partials of the splits: lane l adds the splits l, l + 32, ..., then the butterfly (a fixed order; one round
      // trip to L2 instead of one per split)
      double pl[LP::NACC];
      for (int i = 0; i < LP::NACC; ++i) pl[i] = 0.0;
      for (int b = lp.first(); b < M.stage_nb; b += lp.stride()) pl[LP::NACC == 1 ? 0 : (b & 31)] += M.stage_ld[b * rows + row];
      double lsd = lin_l + lp.reduce(pl);
      if (M.lin_w) {  // linear part of Σ log σ(η̃): ½ Σ_i η̃_i = ½ colsum(X̃)·q
        double part[LP::NACC];
        for (int i = 0; i < LP::NACC; ++i) part[i] = 0.0;
        BN_FOR4(d0, nv) {
          T qv[4]; double wv[4];
          ld4(q + d0, qv); ld4(M.lin_w + d0, wv);
          for (int e = 0; e < nv; ++e) { double& a = part[lp.acc(d0 + e)]; a = fma_(wv[e], (double)qv[e], a); }
        }
        lsd = fma_(0.5, lp.reduce(part), lsd);
      }
      l = fma_(-0.5 * (double)M.tau, (double)dot(q, q), lsd);
    } else {
      T ls = T(0);
      for (int b = 0; b < M.stage_nb; ++b) ls = ls + M.stage_l[b * rows + row];
      l = (double)fma_(T(-0.5) * M.tau, dot(q, q), ls);
    }

    return l;
  }
  BN_HDN double logistic_finalize_rm(const T* q, T* g) const { return logistic_finalize_t<true>(q, g); }

  // ≙ logdensity_and_gradient! (call site src/kinetic_energy.jl:73); SURVEY.md §A.4 targets.
  // Elementwise targets are evaluated here; batched targets were evaluated by a
  // separate kernel into the staging buffers and are finalised here.
  // Returns ℓ as Float64: what a target computes in T is widened once; the tensor path's Float64 partial sums stay Float64.
  BN_HD double model_grad(int slot) const {
    const T* q = zq(slot);
    T* g = zg(slot);
    double l;
    switch (M.model_kind) {
      case MODEL_IID_NORMAL: {
        BN_FOR4(d0, nv) {
          T qv[4], gv[4];
          ld4(q + d0, qv);
          for (int e = 0; e < 4; ++e) gv[e] = iid_grad(qv[e]);
          st4(g + d0, gv, nv);
        }
        l = (double)iid_value(dot(q, q));
        break;
      }
      case MODEL_FUNNEL: {
        T part[LP::NACC];
        for (int i = 0; i < LP::NACC; ++i) part[i] = T(0);
        BN_FOR4(d0, nv) {
          T qv[4];
          ld4(q + d0, qv);
          for (int e = 0; e < nv; ++e)
            if (d0 + e >= 1) { T& x = part[lp.acc(d0 + e)]; x = fma_(qv[e], qv[e], x); }
        }
        const T S = lp.reduce(part);
        const T v = q[0], ex = exp_(-v);
        BN_FOR4(d0, nv) {
          T qv[4], gv[4];
          ld4(q + d0, qv);
          for (int e = 0; e < 4; ++e) gv[e] = (d0 + e == 0) ? funnel_grad_v(v, S, ex, M.D) : funnel_grad_x(qv[e], ex);
          st4(g + d0, gv, nv);
        }
        l = (double)funnel_value(v, S, ex, M.D);
        break;
      }
      case MODEL_GAUSSIAN: {
        const T* sg = M.stage_g + (int64_t)M.stage_row[c] * M.Dp;
        BN_FOR4(d0, nv) {
          T gv[4];
          ld4(sg + d0, gv);
          st4(g + d0, gv, nv);
        }
        lp.sync();
        l = (double)(T(0.5) * dot(q, g));
        break;
      }
      case MODEL_LOGISTIC:
        if constexpr (RM != 0) l = logistic_finalize_rm(q, g); else l = logistic_finalize_t<false>(q, g);
        break;
      default: l = lim<double>::nan();
    }
    lp.sync();
    return l;
  }

  // Remainder mode of the tensor path (logistic_rm.cu): the part of the model that is linear / quadratic in
  // δ = q − beta_ref is exact arithmetic here, one D x D mat-vec per chain: y = H0 δ (fp32 FMAs, k sequential), then
  // g −= y in place and the return value is ell0 + Σ_d δ_d (g0_d − ½ y_d + ⅓ (X̃ᵀρ)_d) in Float64; the last term is the part
  // Σ_i δ_i ρ_i / 3 of the log density's remainder, which the kernel therefore does not sum.  Lane l owns the coordinates
  // 4l..4l+3 (D <= 128, the domain of that kernel); δ_k is broadcast from its owner by a warp shuffle, row k of the
  // symmetric H0 is one coalesced 16-byte load per lane.
  BN_HD double linear_part(const T* q, T* g, const T (&rem)[2][4], const T* sg, int64_t bs) const {
    double part[LP::NACC];
    for (int i = 0; i < LP::NACC; ++i) part[i] = 0.0;
#if defined(__CUDA_ARCH__)
    // lane l owns the groups d = 4l .. 4l+3 and (D > 128) 128 + 4l .. 128 + 4l+3
    const int l4 = lp.first4();
    const int ng = M.D > 128 ? 2 : 1;
    float dv[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, y[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    for (int gI = 0; gI < ng; ++gI) {
      const int d0 = 128 * gI + l4;
      if (d0 < M.D) {
        T qv[4], br[4];
        ld4(q + d0, qv); ld4(M.beta_ref + d0, br);
        for (int e = 0; e < 4; ++e) dv[gI][e] = (d0 + e < M.D) ? (float)qv[e] - (float)br[e] : 0.f;
      }
    }
    // rows k >= D of H0 are zero padding up to Dp (a multiple of 32) and δ_k = 0 there: no guards, so the row loads of two
    // steps are in flight together (guarded, the compiler serialised them: one L2 round trip per row, 35 us per chain)
    const float* Hl = M.lin_H + (l4 < M.Dp ? l4 : 0);   // lanes beyond the row read a valid (unused) address
    if (ng == 1) {
      for (int kk = 0; 4 * kk < M.D; kk += 2) {
        float4 hr[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) hr[u] = *reinterpret_cast<const float4*>(Hl + (int64_t)(4 * kk + u) * M.Dp);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float b = __shfl_sync(0xffffffffu, dv[0][u & 3], kk + (u >> 2));
          y[0][0] = fmaf(hr[u].x, b, y[0][0]); y[0][1] = fmaf(hr[u].y, b, y[0][1]); y[0][2] = fmaf(hr[u].z, b, y[0][2]); y[0][3] = fmaf(hr[u].w, b, y[0][3]);
        }
      }
    } else {
      const bool two = 128 + l4 < M.Dp;
      for (int gk = 0; gk < 2; ++gk) {                     // δ_k is broadcast from lane (k % 128) / 4, group k / 128
        const int kend = (M.D - 128 * gk + 3) / 4 < 32 ? (M.D - 128 * gk + 3) / 4 : 32;
        for (int kk = 0; kk < kend; ++kk) {
          float4 hr[4], hs[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float* rowp = Hl + (int64_t)(128 * gk + 4 * kk + u) * M.Dp;
            hr[u] = *reinterpret_cast<const float4*>(rowp);
            hs[u] = two ? *reinterpret_cast<const float4*>(rowp + 128) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float b = __shfl_sync(0xffffffffu, gk ? dv[1][u] : dv[0][u], kk);
            y[0][0] = fmaf(hr[u].x, b, y[0][0]); y[0][1] = fmaf(hr[u].y, b, y[0][1]); y[0][2] = fmaf(hr[u].z, b, y[0][2]); y[0][3] = fmaf(hr[u].w, b, y[0][3]);
            y[1][0] = fmaf(hs[u].x, b, y[1][0]); y[1][1] = fmaf(hs[u].y, b, y[1][1]); y[1][2] = fmaf(hs[u].z, b, y[1][2]); y[1][3] = fmaf(hs[u].w, b, y[1][3]);
          }
        }
      }
    }
    (void)sg; (void)bs;
    for (int gI = 0; gI < ng; ++gI) {
      const int d0 = 128 * gI + l4;
      if (d0 < M.D) {
        T gv[4]; double g0[4];
        ld4(g + d0, gv); ld4(M.grad0 + d0, g0);
        const int nv = M.D - d0 < 4 ? M.D - d0 : 4;
        for (int e = 0; e < 4; ++e) {
          gv[e] = gv[e] - T(y[gI][e]);
          if (e < nv) part[0] = fma_((double)dv[gI][e], g0[e] - 0.5 * (double)y[gI][e] + (1.0 / 3.0) * (double)rem[gI][e], part[0]);
        }
        st4(g + d0, gv, nv);
      }
    }
#else
    for (int d = 0; d < M.D; ++d) {   // host build: same definition, plain loops (the tensor path itself is CUDA-only)
      float yd = 0.f;
      for (int k = 0; k < M.D; ++k) yd = fmaf(M.lin_H[(int64_t)k * M.Dp + d], (float)q[k] - (float)M.beta_ref[k], yd);
      g[d] = g[d] - T(yd);
      T remd = T(0);
      for (int b = 0; b < M.stage_nb; ++b) remd = remd + sg[b * bs + d];
      double& a = part[lp.acc(d)];
      a = fma_((double)((float)q[d] - (float)M.beta_ref[d]), M.grad0[d] - 0.5 * (double)yd + (1.0 / 3.0) * (double)remd, a);
    }
#endif
    const double r = lp.reduce(part);
    lp.sync();
    return M.ell0 + r;
  }

  // ---- local optimum search (≙ warmup!(FindLocalOptimum), src/warmup.jl:152-186; see nuts_machine.h)
  BN_HD void opt_trial(int src, int dst, T alpha, T lambda) const {
    const T* q = zq(src); const T* g = zg(src);
    T* qn = zq(dst);
    const int row = take_row();
    BN_FOR4(d0, nv) {
      T qv[4], gv[4], qd[4];
      ld4(q + d0, qv); ld4(g + d0, gv);
      for (int e = 0; e < 4; ++e) qd[e] = fma_(alpha, fma_(-lambda, qv[e], gv[e]), qv[e]);
      st4(qn + d0, qd, nv);
      if (row >= 0) stage_put4(row, d0, qd, nv);
    }
    lp.sync();
  }
  BN_HD void opt_norms(int slot, T lambda, T* dd, T* qq) const {
    const T* q = zq(slot); const T* g = zg(slot);
    T a[LP::NACC], c2[LP::NACC];
    for (int i = 0; i < LP::NACC; ++i) a[i] = c2[i] = T(0);
    BN_FOR4(d0, nv) {
      T qv[4], gv[4];
      ld4(q + d0, qv); ld4(g + d0, gv);
      for (int e = 0; e < nv; ++e) {
        const T dv = fma_(-lambda, qv[e], gv[e]);
        T& x = a[lp.acc(d0 + e)]; x = fma_(dv, dv, x);
        T& y = c2[lp.acc(d0 + e)]; y = fma_(qv[e], qv[e], y);
      }
    }
    *dd = lp.reduce(a);
    *qq = lp.reduce(c2);
  }
  BN_HD void opt_dots(int src, int dst, T lambda, T* dd_new, T* qq_new, T* dod) const {
    const T* q0 = zq(src); const T* g0 = zg(src);
    const T* q1 = zq(dst); const T* g1 = zg(dst);
    T a[LP::NACC], c2[LP::NACC], e[LP::NACC];
    for (int i = 0; i < LP::NACC; ++i) a[i] = c2[i] = e[i] = T(0);
    BN_FOR4(d0, nv) {
      T qa[4], ga[4], qb[4], gb[4];
      ld4(q0 + d0, qa); ld4(g0 + d0, ga); ld4(q1 + d0, qb); ld4(g1 + d0, gb);
      for (int k = 0; k < nv; ++k) {
        const T da = fma_(-lambda, qa[k], ga[k]);
        const T db = fma_(-lambda, qb[k], gb[k]);
        T& x = a[lp.acc(d0 + k)]; x = fma_(db, db, x);
        T& y = c2[lp.acc(d0 + k)]; y = fma_(qb[k], qb[k], y);
        T& z = e[lp.acc(d0 + k)]; z = fma_(da, db, z);
      }
    }
    *dd_new = lp.reduce(a);
    *qq_new = lp.reduce(c2);
    *dod = lp.reduce(e);
  }
  BN_HD void opt_restart_position(int slot, uint64_t seed, uint32_t gchain, uint32_t attempt) const {
    T* q = zq(slot);
    const int row = take_row();
    BN_FOR4(d0, nv) {
      T qd[4] = {T(0), T(0), T(0), T(0)};
      for (int e = 0; e < nv; ++e) qd[e] = T(restart_position(seed, gchain, attempt, (uint32_t)(d0 + e)));
      st4(q + d0, qd, nv);
      if (row >= 0) stage_put4(row, d0, qd, nv);
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
    if (lp.lane0()) M.bare_out[3 * blk + c] = get_lq(slot);
  }
  // set_positions: q -> slot (and staging); ≙ src/warmup.jl:119 / random_position! :73
  BN_HD void load_position(int slot, uint64_t seed, uint32_t gchain) const {
    T* q = zq(slot);
    const int row = take_row();
    BN_FOR4(d0, nv) {
      T qd[4] = {T(0), T(0), T(0), T(0)};
      for (int e = 0; e < nv; ++e)
        qd[e] = M.pos_in ? T(M.pos_in[(int64_t)c * M.D + d0 + e]) : T(init_position(seed, gchain, (uint32_t)(d0 + e)));
      st4(q + d0, qd, nv);
      if (row >= 0) stage_put4(row, d0, qd, nv);
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
template <class T, class LP, int RM = 0>
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
  Backend<T, LP, RM> b(M, c, lp, rp);
  Machine<T, Backend<T, LP, RM>> m(b, s, g, rp, c);
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

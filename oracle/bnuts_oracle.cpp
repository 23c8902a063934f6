// bnuts_oracle.cpp — CPU oracle: a restatement of the InplaceDHMC.jl hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product (inplacedhmc.jl_b200/, libbnuts.so)
// may include, link, load or execute this file; only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs use it, as the checker and
// as the timed CPU baseline.
//
// PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures
// (test/runtests.jl:4-6 is an empty testset) and cannot be executed here (no
// Julia; ten un-vendored, un-pinned dependencies; load-time defects, SURVEY.md
// §A.3).  This file follows the cited reference lines statement by statement and
// is validated by analytic known answers (tests/test_oracle_*.py).  Arithmetic the
// reference leaves to un-pinned packages (SIMD reduction order, RNG, exp/log) is
// fixed by the conventions in inplacedhmc.jl_b200/csrc/bnuts_math.h.
//
// It exports the same C ABI as include/bnuts.h so one ctypes binding drives both.
// Structure deliberately mirrors the reference (recursive adjacent_tree), unlike
// the CUDA engine (iterative per-chain state machine).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>
#include <omp.h>

#include "../include/bnuts.h"
#include "../inplacedhmc.jl_b200/csrc/bnuts_math.h"
#include "../inplacedhmc.jl_b200/csrc/bnuts_models.h"

namespace {

using bn::exp_;
using bn::fma_;
using bn::isfinite_;
using bn::log_;
using bn::logaddexp_;

// ------------------------------------------------------------------ reductions
// "warp order": lane l accumulates the groups of four consecutive elements 4l..4l+3 (+128, ...) in index order
// (accumulator (d >> 2) & 31), then xor-butterfly.
template <class T> T butterfly(T* part) {
  for (int off = 16; off >= 1; off >>= 1) {
    T nw[32];
    for (int l = 0; l < 32; ++l) nw[l] = part[l] + part[l ^ off];
    for (int l = 0; l < 32; ++l) part[l] = nw[l];
  }
  return part[0];
}
template <class T> T dot_warp(const T* a, const T* b, int D) {
  T part[32];
  for (int l = 0; l < 32; ++l) part[l] = T(0);
  for (int d = 0; d < D; ++d) part[(d >> 2) & 31] = fma_(a[d], b[d], part[(d >> 2) & 31]);
  return butterfly(part);
}

// ------------------------------------------------------------------ arena
// ≙ the per-thread bump allocator (src/InplaceDHMC.jl:51-78); reset per transition.
template <class T> struct Arena {
  std::vector<std::unique_ptr<T[]>> blocks;
  size_t block_elems = size_t(1) << 20, cur = 0, top = 0;
  T* alloc(size_t n) {
    if (n > block_elems) block_elems = n;
    if (blocks.empty() || top + n > block_elems) {
      if (!blocks.empty()) ++cur;
      if (cur >= blocks.size()) blocks.emplace_back(new T[block_elems]);
      top = 0;
    }
    T* p = blocks[cur].get() + top;
    top += n;
    return p;
  }
  void reset() { cur = 0; top = 0; }
};

// ------------------------------------------------------------------ model
template <class T> struct Model {
  int kind = bn::MODEL_NONE, D = 0;
  std::vector<T> P;            // gaussian precision [D][D]
  std::vector<T> X, y;         // logistic [N][D], [N]
  int64_t N = 0;
  T tau = T(0);
  int row_blocks = 1;

  // ≙ logdensity_and_gradient!(∇ℓq, ℓ, q, sptr) -> ℓq   (call site src/kinetic_energy.jl:73)
  T eval(const T* q, T* g, std::vector<T>& scratch) const {
    switch (kind) {
      case bn::MODEL_IID_NORMAL: {
        for (int d = 0; d < D; ++d) g[d] = bn::iid_grad(q[d]);
        return bn::iid_value(dot_warp(q, q, D));
      }
      case bn::MODEL_FUNNEL: {
        T part[32];
        for (int l = 0; l < 32; ++l) part[l] = T(0);
        for (int d = 1; d < D; ++d) part[(d >> 2) & 31] = fma_(q[d], q[d], part[(d >> 2) & 31]);
        const T S = butterfly(part);
        const T v = q[0], e = exp_(-v);
        g[0] = bn::funnel_grad_v(v, S, e, D);
        for (int d = 1; d < D; ++d) g[d] = bn::funnel_grad_x(q[d], e);
        return bn::funnel_value(v, S, e, D);
      }
      case bn::MODEL_GAUSSIAN: {
        for (int d = 0; d < D; ++d) {
          T acc = T(0);
          const T* row = &P[size_t(d) * D];
          for (int k = 0; k < D; ++k) acc = fma_(row[k], q[k], acc);
          g[d] = -acc;
        }
        return T(0.5) * dot_warp(q, g, D);
      }
      case bn::MODEL_LOGISTIC: {
        // rows are summed sequentially inside each of row_blocks blocks, block
        // partials are summed sequentially; prior added once at the end.
        const int64_t R = (N + row_blocks - 1) / row_blocks;
        scratch.assign(size_t(D), T(0));
        T* part = scratch.data();
        for (int d = 0; d < D; ++d) g[d] = T(0);
        T l = T(0);
        for (int64_t b = 0; b < row_blocks; ++b) {
          const int64_t i0 = b * R, i1 = std::min<int64_t>(N, i0 + R);
          for (int d = 0; d < D; ++d) part[d] = T(0);
          T pl = T(0);
          for (int64_t i = i0; i < i1; ++i) {
            const T* x = &X[size_t(i) * D];
            T eta = T(0);
            for (int d = 0; d < D; ++d) eta = fma_(x[d], q[d], eta);
            T r, lt;
            bn::logistic_elem(eta, y[i], &r, &lt);
            pl = pl + lt;
            for (int d = 0; d < D; ++d) part[d] = fma_(x[d], r, part[d]);
          }
          for (int d = 0; d < D; ++d) g[d] = g[d] + part[d];
          l = l + pl;
        }
        for (int d = 0; d < D; ++d) g[d] = fma_(-tau, q[d], g[d]);
        return fma_(T(-0.5) * tau, dot_warp(q, q, D), l);
      }
    }
    return bn::lim<T>::nan();
  }
};

// ------------------------------------------------------------------ types
// ≙ EvaluatedLogDensity + PhasePoint, src/hamiltonian.jl:237-276
// Energy scalars (ℓ, H, Δ, log-weights) are Float64 for BOTH engine types: the reference is Float64-only, and an fp32 ℓ of
// magnitude N·log 2 has an ulp of 0.03 at N = 1e6 and of 8 at N = 1e8 — no acceptance test survives that.  The fp32 variant
// is fp32 STATE VECTORS (q, p, ∇ℓ, metric) and vector arithmetic; what a target's evaluation returns in T is widened once.
template <class T> struct PhasePoint { T* q; T* p; T* g; double lq; };
// ≙ GeneralizedTurnStatistic, src/NUTS.jl:93-97
template <class T> struct TurnStat { const T* psm; const T* psp; const T* rho; };
// ≙ AcceptanceStatistic, src/NUTS.jl:58-66
template <class T> struct Visited { double lsa; int32_t steps; };
// proposal ζ plus what TreeStatisticsNUTS needs from it (π = logdensity(H, ζ), src/NUTS.jl:262)
template <class T> struct Proposal { PhasePoint<T> z; double pi; int32_t idx; };
template <class T> struct SubTree { Proposal<T> zeta; double omega; TurnStat<T> tau; PhasePoint<T> zlast; int32_t ilast; };
// ≙ InvalidTree, src/tree.jl:278-300
struct Invalid { bool flag; int32_t left, right; };

// ≙ DualAveragingState, src/stepsize.jl:196-202 (kept in Float64 like the reference)
struct DAState { double mu; int64_t m; double Hbar, logeps, logepsbar; };

template <class T> struct ChainCtx;

// Decision margins of one transition (test instrumentation, bnuts_oracle_trace): how far every data-dependent
// decision of the tree was from its threshold, so a test can prove that a tolerance-parity engine (the tensor-core
// gradient path) may only differ where a compared quantity sits within its error bound of the threshold.
//   min_div    min over leaves of |Δ − min_Δ|                                  (divergence test, src/NUTS.jl:180)
//   min_turn   min over merges of min(|ρ·p♯₋| / Σ|ρ_d p♯₋_d|, same for p♯₊)      (is_turning, src/NUTS.jl:148-170)
//   min_sel    min over consuming merges of |e + logprob2|                      (rand_bool_logprob, src/NUTS.jl:32-34)
//   scale_H    max over leaves of |ℓq| + |K|: the magnitude whose rounding bounds the error of every Δ and ω
//   grad_sq    max over leaves of |∇ℓ|²: how strongly an error of the position moves the energy
struct DecisionTrace { double min_div, min_turn, min_sel, scale_H, grad_sq; };
thread_local DecisionTrace* g_trace = nullptr;
inline void trace_min(double& slot, double v) { if (v < slot) slot = v; }

template <class T> struct Engine {
  bnuts_config cfg{};
  int C = 0, D = 0;
  Model<T> model;
  std::vector<T> q, g;           // [C][D], [C][D]
  std::vector<double> lq;        // [C]
  std::vector<T> Minv, W;        // [C][D]   (GaussianKineticEnergy, src/hamiltonian.jl:33-38)
  // shared dense metric (≙ the dense GaussianKineticEnergy constructor kept as a comment at
  // src/hamiltonian.jl:44): M⁻¹ [D][D] and the momentum factor W = L⁻ᵀ with M⁻¹ = L Lᵀ, so W Wᵀ = M.
  // Applied DIRECTLY here (mat-vecs in the kinetic energy, p♯, the drift and the momentum draw); the device
  // engine whitens instead, so agreement of the two is a check of that equivalence.
  bool dense = false;
  std::vector<T> MinvD, WD;      // [D][D] row-major
  std::vector<double> MinvD64;
  std::vector<double> eps;       // [C]
  std::vector<int32_t> status;   // [C]
  uint64_t seed = 0;
  uint32_t next_t = 0;
  // injection (≙ the p = / directions = hooks of sample_tree, src/NUTS.jl:251-258)
  int inj_T = 0;
  uint32_t inj_start = 0;
  std::vector<uint32_t> inj_dirs;
  std::vector<double> inj_p;
  bool has_inj_dirs = false, has_inj_p = false;
  std::vector<DecisionTrace> trace;   // [C] margins of each chain's LAST transition (empty: tracing off)
  std::vector<double> inj_exps;   // [T][C][inj_nexp]: scripted randexp stream of rand_bool_logprob (src/NUTS.jl:32-34)
  int inj_nexp = 0;
  bnuts_counter_block counters{};
  std::string err;
  std::vector<ChainCtx<T>> ctx;  // per OpenMP thread
};

template <class T> struct ChainCtx {
  Arena<T> arena;
  std::vector<T> scratch;
};

// ------------------------------------------------------------------ Hamiltonian pieces
template <class T> struct Ham {
  const Engine<T>* E;
  int c;  // local chain
  const T* Minv() const { return &E->Minv[size_t(c) * E->D]; }
  const T* W() const { return &E->W[size_t(c) * E->D]; }
  bool dense() const { return E->dense; }
  // out = M⁻¹ v, k sequential
  void apply_minv(const T* v, T* out) const {
    const int D = E->D;
    for (int i = 0; i < D; ++i) {
      T acc = T(0);
      const T* row = &E->MinvD[size_t(i) * D];
      for (int k = 0; k < D; ++k) acc = fma_(row[k], v[k], acc);
      out[i] = acc;
    }
  }
  // p = W z  (≙ rand_p!, src/kinetic_energy.jl:63, with a dense W)
  void draw_momentum(T* p, uint64_t seed, uint32_t gchain, uint32_t t) const {
    const int D = E->D;
    if (!E->dense) {
      const T* Wd = W();
      for (int d = 0; d < D; ++d) p[d] = Wd[d] * bn::std_normal(seed, gchain, t, uint32_t(d), T(0));
      return;
    }
    std::vector<T> z(D);
    for (int d = 0; d < D; ++d) z[d] = bn::std_normal(seed, gchain, t, uint32_t(d), T(0));
    for (int i = 0; i < D; ++i) {
      T acc = T(0);
      const T* row = &E->WD[size_t(i) * D];
      for (int k = 0; k < D; ++k) acc = fma_(row[k], z[k], acc);
      p[i] = acc;
    }
  }
};

// ≙ kinetic_energy, src/kinetic_energy.jl:14-24:  ke += p*M⁻¹*p ; 0.5*ke
template <class T> T kinetic_energy(const Ham<T>& H, const T* p) {
  const int D = H.E->D;
  if (H.dense()) {
    std::vector<T> ps(D);
    H.apply_minv(p, ps.data());
    return T(0.5) * dot_warp(ps.data(), p, D);
  }
  const T* Mi = H.Minv();
  T part[32];
  for (int l = 0; l < 32; ++l) part[l] = T(0);
  for (int d = 0; d < D; ++d) part[(d >> 2) & 31] = fma_(p[d] * Mi[d], p[d], part[(d >> 2) & 31]);
  return T(0.5) * butterfly(part);
}
// ≙ logdensity(H, z), src/kinetic_energy.jl:107-112
template <class T> double logdensity(const Ham<T>& H, const PhasePoint<T>& z) {
  if (!isfinite_(z.lq)) return -bn::lim<double>::inf();
  const T K = kinetic_energy(H, z.p);
  return z.lq - (isfinite_(K) ? double(K) : bn::lim<double>::inf());
}
// ≙ calculate_p♯, src/kinetic_energy.jl:39-46
template <class T> T* calculate_psharp(Arena<T>& A, const Ham<T>& H, const T* p) {
  const int D = H.E->D;
  const T* Mi = H.Minv();
  T* ps = A.alloc(D);
  if (H.dense()) { H.apply_minv(p, ps); return ps; }
  for (int d = 0; d < D; ++d) ps[d] = Mi[d] * p[d];
  return ps;
}
// ≙ evaluate_ℓ!, src/kinetic_energy.jl:72-85 (non-finite ℓ -> -Inf, ∇ aliased to q)
template <class T> double evaluate_l(ChainCtx<T>& X, const Ham<T>& H, const T* q, T* g) {
  const T lq = H.E->model.eval(q, g, X.scratch);
  if (isfinite_(lq)) return double(lq);
  for (int d = 0; d < H.E->D; ++d) g[d] = q[d];
  return -bn::lim<double>::inf();
}
// ≙ leapfrog, src/kinetic_energy.jl:126-163 (and the stack twin :164-195)
template <class T> PhasePoint<T> leapfrog(ChainCtx<T>& X, const Ham<T>& H, const PhasePoint<T>& z, T eps) {
  const int D = H.E->D;
  const T* Mi = H.Minv();
  PhasePoint<T> n;
  n.p = X.arena.alloc(D);
  n.q = X.arena.alloc(D);
  n.g = X.arena.alloc(D);
  const T eh = T(0.5) * eps;
  if (H.dense()) {
    for (int d = 0; d < D; ++d) n.p[d] = fma_(eh, z.g[d], z.p[d]);
    H.apply_minv(n.p, n.g);                       // n.g as scratch: M⁻¹ pₘ
    for (int d = 0; d < D; ++d) n.q[d] = fma_(eps, n.g[d], z.q[d]);
  } else
  for (int d = 0; d < D; ++d) {
    const T pm = fma_(eh, z.g[d], z.p[d]);
    n.p[d] = pm;
    n.q[d] = fma_(eps * Mi[d], pm, z.q[d]);
  }
  n.lq = evaluate_l(X, H, n.q, n.g);
  for (int d = 0; d < D; ++d) n.p[d] = fma_(eh, n.g[d], n.p[d]);
  return n;
}

// ------------------------------------------------------------------ NUTS (src/NUTS.jl, src/tree.jl)
template <class T> struct Trajectory {  // ≙ TrajectoryNUTS, src/NUTS.jl:5-16
  double pi0; T eps; double min_delta;
  uint64_t seed;
  uint32_t chain, t;  // RNG position
  const double* exps = nullptr;  // scripted exponentials of this transition (bnuts_inject), or null
  int n_exps = 0;
  mutable int n_exp = 0;         // consumed so far
};

// ≙ leaf, src/NUTS.jl:176-191 (+ leaf_acceptance_statistic :76-78, leaf_turn_statistic :113-116)
template <class T>
void leaf(ChainCtx<T>& X, const Ham<T>& H, const Trajectory<T>& tr, const PhasePoint<T>& z, bool is_initial,
          int32_t idx, Proposal<T>* zeta, double* omega, TurnStat<T>* tau, Visited<T>* v, bool* isdiv) {
  const double Hz = is_initial ? tr.pi0 : logdensity(H, z);
  const double delta = is_initial ? 0.0 : Hz - tr.pi0;
  *isdiv = delta < tr.min_delta;
  if (g_trace) {
    if (!is_initial) trace_min(g_trace->min_div, std::fabs(double(delta) - double(tr.min_delta)));
    const double sc = std::fabs(double(z.lq)) + std::fabs(double(z.lq) - double(Hz));
    if (sc > g_trace->scale_H && std::isfinite(sc)) g_trace->scale_H = sc;
    double g2 = 0.0;
    for (int d = 0; d < H.E->D; ++d) g2 += double(z.g[d]) * double(z.g[d]);
    if (g2 > g_trace->grad_sq && std::isfinite(g2)) g_trace->grad_sq = g2;
  }
  if (is_initial) { v->lsa = -bn::lim<double>::inf(); v->steps = 0; }
  else { v->lsa = (delta < 0.0) ? delta : 0.0; v->steps = 1; }
  zeta->z = z; zeta->pi = Hz; zeta->idx = idx;
  *omega = delta;
  if (*isdiv) { tau->psm = tau->psp = tau->rho = nullptr; return; }
  const T* ps = calculate_psharp(X.arena, H, z.p);
  tau->psm = ps; tau->psp = ps; tau->rho = z.p;
}
// ≙ combine_acceptance_statistics, src/NUTS.jl:68-70
template <class T> Visited<T> combine_visited(const Visited<T>& a, const Visited<T>& b) {
  return Visited<T>{logaddexp_(a.lsa, b.lsa), a.steps + b.steps};
}
// ≙ combine_turn_statistics, src/NUTS.jl:118-145 (x is earlier in time)
template <class T> TurnStat<T> combine_turn(ChainCtx<T>& X, int D, const TurnStat<T>& x, const TurnStat<T>& y) {
  T* rho = X.arena.alloc(D);
  for (int d = 0; d < D; ++d) rho[d] = x.rho[d] + y.rho[d];
  return TurnStat<T>{x.psm, y.psp, rho};
}
// ≙ combine_turn_statistics_in_direction, src/tree.jl:230-236
template <class T>
TurnStat<T> combine_turn_dir(ChainCtx<T>& X, int D, const TurnStat<T>& t1, const TurnStat<T>& t2, bool fwd) {
  return fwd ? combine_turn(X, D, t1, t2) : combine_turn(X, D, t2, t1);
}
// ≙ is_turning, src/NUTS.jl:148-170 (both dots always computed; strict <; NaN => false)
template <class T> bool is_turning(int D, const TurnStat<T>& tau) {
  T pm[32], pp[32];
  for (int l = 0; l < 32; ++l) pm[l] = pp[l] = T(0);
  for (int d = 0; d < D; ++d) {
    const T r = tau.rho[d];
    pm[(d >> 2) & 31] = fma_(r, tau.psm[d], pm[(d >> 2) & 31]);
    pp[(d >> 2) & 31] = fma_(r, tau.psp[d], pp[(d >> 2) & 31]);
  }
  const T dm = butterfly(pm), dp = butterfly(pp);
  if (g_trace) {
    double am = 0.0, ap = 0.0;
    for (int d = 0; d < D; ++d) { am += std::fabs(double(tau.rho[d]) * double(tau.psm[d])); ap += std::fabs(double(tau.rho[d]) * double(tau.psp[d])); }
    trace_min(g_trace->min_turn, std::fabs(double(dm)) / (am > 0 ? am : 1.0));
    trace_min(g_trace->min_turn, std::fabs(double(dp)) / (ap > 0 ? ap : 1.0));
  }
  return (dm < T(0)) | (dp < T(0));
}
// ≙ rand_bool_logprob, src/NUTS.jl:32-34 — the draw is consumed only if logprob < 0
template <class T> bool rand_bool_logprob(const Trajectory<T>& tr, double logprob, uint32_t j, uint32_t k, uint32_t n) {
  // "logprob >= 0" is a decision too: it settles whether a draw is consumed, which shifts a scripted stream
  if (g_trace && std::isfinite(double(logprob))) trace_min(g_trace->min_sel, std::fabs(double(logprob)));
  if (logprob >= 0.0) return true;
  T e;   // the variate itself is a T (the engine type's stream); the comparison is Float64
  if (tr.exps && tr.n_exp < tr.n_exps) e = T(tr.exps[tr.n_exp]);   // scripted stream: consumed in call order
  else e = bn::std_exponential(tr.seed, tr.chain, tr.t, j, k, n, T(0));
  tr.n_exp += 1;
  if (g_trace) trace_min(g_trace->min_sel, std::fabs(double(e) + double(logprob)));
  return double(e) > -logprob;
}
// ≙ combine_proposals_and_logweights, src/tree.jl:238-245 with
//   biased_progressive_logprob2 (src/tree.jl:261-263) and combine_proposals (src/NUTS.jl:40-45)
template <class T>
void combine_proposals_and_logweights(const Trajectory<T>& tr, const Proposal<T>& z1, const Proposal<T>& z2, double w1,
                                      double w2, bool is_doubling, uint32_t j, uint32_t k, uint32_t n,
                                      Proposal<T>* z, double* w) {
  *w = logaddexp_(w1, w2);
  const double logprob2 = w2 - (is_doubling ? w1 : *w);
  *z = rand_bool_logprob(tr, logprob2, j, k, n) ? z2 : z1;
}

// ≙ adjacent_tree, src/tree.jl:321-366.  j = doubling number, base = leaves of this
// doubling built before this subtree (only used to name the merge draws).
template <class T>
void adjacent_tree(ChainCtx<T>& X, const Ham<T>& H, const Trajectory<T>& tr, const PhasePoint<T>& z, int32_t i,
                   int32_t depth, bool fwd, uint32_t j, uint32_t base, SubTree<T>* out, Visited<T>* v,
                   Invalid* inv) {
  const int32_t ip = i + (fwd ? 1 : -1);
  if (depth == 0) {
    const PhasePoint<T> zn = leapfrog(X, H, z, fwd ? tr.eps : -tr.eps);  // ≙ move, src/NUTS.jl:18-21
    bool isdiv;
    leaf(X, H, tr, zn, false, ip, &out->zeta, &out->omega, &out->tau, v, &isdiv);
    out->zlast = zn; out->ilast = ip;
    *inv = Invalid{isdiv, ip, ip};
    return;
  }
  SubTree<T> tm;
  Visited<T> vm;
  adjacent_tree(X, H, tr, z, i, depth - 1, fwd, j, base, &tm, &vm, inv);
  if (inv->flag) { *out = tm; *v = vm; return; }
  SubTree<T> tp;
  Visited<T> vp;
  adjacent_tree(X, H, tr, tm.zlast, tm.ilast, depth - 1, fwd, j, base + (1u << (depth - 1)), &tp, &vp, inv);
  *v = combine_visited(vm, vp);
  if (inv->flag) { *out = tp; return; }
  const TurnStat<T> tau = combine_turn_dir(X, H.E->D, tm.tau, tp.tau, fwd);
  if (is_turning(H.E->D, tau)) { *out = tp; *inv = Invalid{true, ip, tp.ilast}; return; }
  out->tau = tau; out->zlast = tp.zlast; out->ilast = tp.ilast;
  combine_proposals_and_logweights(tr, tm.zeta, tp.zeta, tm.omega, tp.omega, false, j, uint32_t(depth),
                                   base + (1u << depth), &out->zeta, &out->omega);
  *inv = Invalid{false, 1, 0};
}

// ≙ sample_trajectory, src/tree.jl:382-444
template <class T>
void sample_trajectory(ChainCtx<T>& X, const Ham<T>& H, const Trajectory<T>& tr, const PhasePoint<T>& z,
                       int max_depth, uint32_t directions, Proposal<T>* zeta_out, Visited<T>* v_out,
                       Invalid* term_out, int32_t* depth_out) {
  Proposal<T> zeta;
  double omega;
  TurnStat<T> tau;
  Visited<T> v;
  bool isdiv;
  leaf(X, H, tr, z, true, 0, &zeta, &omega, &tau, &v, &isdiv);
  PhasePoint<T> zm = z, zp = z;
  int32_t depth = 0, im = 0, ipl = 0;
  Invalid term{false, 1, 0};  // REACHED_MAX_DEPTH
  while (depth < max_depth) {
    const bool fwd = directions & 1u;  // ≙ next_direction, src/tree.jl:152-155
    directions >>= 1;
    SubTree<T> tn;
    Visited<T> vn;
    Invalid inv;
    adjacent_tree(X, H, tr, fwd ? zp : zm, fwd ? ipl : im, depth, fwd, uint32_t(depth), 0u, &tn, &vn, &inv);
    v = combine_visited(v, vn);
    if (inv.flag) { term = inv; break; }
    if (fwd) { zp = tn.zlast; ipl = tn.ilast; } else { zm = tn.zlast; im = tn.ilast; }
    combine_proposals_and_logweights(tr, zeta, tn.zeta, omega, tn.omega, true, uint32_t(depth), 0u, 0u, &zeta,
                                     &omega);
    depth += 1;
    tau = combine_turn_dir(X, H.E->D, tau, tn.tau, fwd);
    if (is_turning(H.E->D, tau)) { term = Invalid{true, im, ipl}; break; }
  }
  *zeta_out = zeta; *v_out = v; *term_out = term; *depth_out = depth;
}

// ≙ sample_tree, src/NUTS.jl:251-264.  Writes the new position into the engine's
// (q, g, lq) for chain c and returns the statistics.
template <class T>
bnuts_tree_stats sample_tree(Engine<T>& E, ChainCtx<T>& X, int c, double eps, uint32_t t, int32_t* sel_idx) {
  const int D = E.D;
  X.arena.reset();
  Ham<T> H{&E, c};
  const uint32_t gchain = uint32_t(E.cfg.chain_offset + c);
  PhasePoint<T> z;
  z.q = X.arena.alloc(D); z.p = X.arena.alloc(D); z.g = X.arena.alloc(D);
  for (int d = 0; d < D; ++d) { z.q[d] = E.q[size_t(c) * D + d]; z.g[d] = E.g[size_t(c) * D + d]; }
  z.lq = E.lq[c];
  const bool injected = E.inj_T > 0 && t >= E.inj_start && t < E.inj_start + uint32_t(E.inj_T);
  const size_t it = injected ? size_t(t - E.inj_start) : 0;
  uint32_t directions = bn::draw_directions(E.seed, gchain, t);  // drawn first, src/NUTS.jl:252
  if (injected && E.has_inj_dirs) directions = E.inj_dirs[it * E.C + c];
  if (injected && E.has_inj_p) {
    for (int d = 0; d < D; ++d) z.p[d] = T(E.inj_p[(it * E.C + c) * D + d]);
  } else {  // ≙ rand_p!, src/kinetic_energy.jl:63: p = W .* randn
    H.draw_momentum(z.p, E.seed, gchain, t);
  }
  Trajectory<T> tr{logdensity(H, z), T(eps), double(E.cfg.min_delta), E.seed, gchain, t};
  if (injected && E.inj_nexp > 0) { tr.exps = E.inj_exps.data() + (it * E.C + c) * size_t(E.inj_nexp); tr.n_exps = E.inj_nexp; }
  Proposal<T> zeta;
  Visited<T> v;
  Invalid term;
  int32_t depth;
  if (!E.trace.empty()) { E.trace[size_t(c)] = DecisionTrace{1e300, 1e300, 1e300, 0.0, 0.0}; g_trace = &E.trace[size_t(c)]; }
  sample_trajectory(X, H, tr, z, E.cfg.max_depth, directions, &zeta, &v, &term, &depth);
  g_trace = nullptr;
  bnuts_tree_stats st;
  st.pi = double(zeta.pi);
  const double a = exp_(v.lsa) / double(v.steps);  // ≙ acceptance_rate, src/NUTS.jl:84
  st.acceptance_rate = a < 1.0 ? a : 1.0;
  st.term_left = term.left; st.term_right = term.right;
  st.depth = depth; st.steps = v.steps;
  for (int d = 0; d < D; ++d) { E.q[size_t(c) * D + d] = zeta.z.q[d]; E.g[size_t(c) * D + d] = zeta.z.g[d]; }
  E.lq[c] = zeta.z.lq;
  if (sel_idx) *sel_idx = zeta.idx;
  return st;
}

// ------------------------------------------------------------------ step size (src/stepsize.jl)
// ≙ initial_adaptation_state :208-212
DAState da_init(double eps) {
  const double le = log_(eps);
  return DAState{log_(10.0) + le, 0, 0.0, le, 0.0};
}
// ≙ adapt_stepsize :220-229
DAState da_adapt(const bnuts_dual_averaging& P, DAState A, double a) {
  A.m += 1;
  const double m = double(A.m);
  A.Hbar += (P.delta - a - A.Hbar) / (m + double(P.t0));
  A.logeps = A.mu - bn::sqrt_(m) / P.gamma * A.Hbar;
  A.logepsbar += exp_(-P.kappa * log_(m)) * (A.logeps - A.logepsbar);
  return A;
}

// ≙ local_acceptance_ratio :150-160 — A(eps) = exp(H(leapfrog(z, eps)) - H(z))
template <class T> struct LocalAcceptance {
  ChainCtx<T>* X; Ham<T> H; PhasePoint<T> z; double target;
  double operator()(double eps) const {
    X->arena.reset();
    const PhasePoint<T> zn = leapfrog(*X, H, z, T(eps));
    return exp_(logdensity(H, zn) - target);
  }
};
// ≙ find_crossing_stepsize :51-72, bisect_stepsize :83-102, find_initial_stepsize :111-126
template <class F> int find_initial_stepsize(const bnuts_stepsize_search& P, F& A, double* out) {
  double e0 = P.eps0, A0 = A(e0);
  if (P.a_min <= A0 && A0 <= P.a_max) { *out = e0; return 0; }
  const double s = A0 > P.a_max ? 1.0 : -1.0, a = A0 > P.a_max ? P.a_max : P.a_min;
  const double Cf = s < 0 ? 1.0 / P.C : P.C;
  double e1 = 0, A1 = 0;
  bool crossed = false;
  for (int it = 0; it < P.maxiter_crossing; ++it) {
    e1 = e0 * Cf; A1 = A(e1);
    if (s * (A1 - a) <= 0) { crossed = true; break; }
    e0 = e1; A0 = A1;
  }
  if (!crossed) return BNUTS_ERR_STEPSIZE_SEARCH;
  if (P.a_min <= A1 && A1 <= P.a_max) { *out = e1; return 0; }
  double lo = e0, hi = e1;
  if (!(e0 < e1)) { lo = e1; hi = e0; }
  for (int it = 0; it < P.maxiter_bisect; ++it) {
    const double em = 0.5 * (lo + hi);  // ≙ middle()
    const double Am = A(em);
    if (P.a_min <= Am && Am <= P.a_max) { *out = em; return 0; }
    if (Am < P.a_min) hi = em; else lo = em;
  }
  return BNUTS_ERR_STEPSIZE_SEARCH;
}

// ------------------------------------------------------------------ metric (src/hamiltonian.jl:77-101,153-162)
// draws: N vectors of length D with stride `stride` (Float64); writes M⁻¹, W.
template <class T> void metric_update(const double* draws, int64_t stride, int N, int D, double lambda, T* Minv, T* W) {
  const double Nf = double(N);
  const double Ninv = 1.0 / Nf;
  const double mulreg = Nf / ((Nf + lambda) * (Nf - 1.0));
  const double addreg = 1e-3 * lambda / (Nf + lambda);
  for (int d = 0; d < D; ++d) {
    const double mu = draws[d];
    double sd = 0.0, sd2 = 0.0;
    for (int n = 1; n < N; ++n) {
      const double dl = draws[int64_t(n) * stride + d] - mu;
      sd = dl + sd;
      sd2 = fma_(dl, dl, sd2);
    }
    const double s2nm1 = fma_(-(sd * sd), Ninv, sd2);
    const double reg = fma_(s2nm1, mulreg, addreg);
    Minv[d] = T(reg);
    W[d] = T(1.0 / bn::sqrt_(reg));
  }
}

// ------------------------------------------------------------------ engine plumbing
struct AnyEngine {
  int dtype;
  Engine<double>* e64 = nullptr;
  Engine<float>* e32 = nullptr;
};
thread_local std::string g_create_error;

template <class T> int32_t fail(Engine<T>& E, int32_t code, const std::string& msg) { E.err = msg; return code; }

template <class T> void ensure_ctx(Engine<T>& E) {
  const int nt = omp_get_max_threads();
  if (int(E.ctx.size()) < nt) E.ctx.resize(nt);
}

template <class T> int32_t set_positions(Engine<T>& E, const double* q) {
  if (E.model.kind == bn::MODEL_NONE) return fail(E, BNUTS_ERR_NO_MODEL, "no model set");
  ensure_ctx(E);
  const int C = E.C, D = E.D;
  int bad = 0;
#pragma omp parallel for schedule(dynamic) reduction(+ : bad)
  for (int c = 0; c < C; ++c) {
    ChainCtx<T>& X = E.ctx[omp_get_thread_num()];
    for (int d = 0; d < D; ++d)
      E.q[size_t(c) * D + d] = q ? T(q[size_t(c) * D + d])
                                 : T(bn::init_position(E.seed, uint32_t(E.cfg.chain_offset + c), uint32_t(d)));
    Ham<T> H{&E, c};
    E.lq[c] = evaluate_l(X, H, &E.q[size_t(c) * D], &E.g[size_t(c) * D]);
    E.status[c] = isfinite_(E.lq[c]) ? 0 : BNUTS_ERR_NONFINITE_START;
    bad += E.status[c] != 0;
  }
  if (bad) return fail(E, BNUTS_ERR_NONFINITE_START, "starting point has non-finite density");
  return 0;
}

template <class T>
int32_t run_transitions(Engine<T>& E, int N, const bnuts_dual_averaging* da, int metric_kind, double lambda,
                        double* chain_out, int64_t sd, int64_t sc, bnuts_tree_stats* stats_out, int64_t ssc,
                        int32_t* sel, double* eps_out) {
  if (E.model.kind == bn::MODEL_NONE) return fail(E, BNUTS_ERR_NO_MODEL, "no model set");
  if (N <= 0) return fail(E, BNUTS_ERR_INVALID_ARGUMENT, "N must be positive");
  if (metric_kind == BNUTS_METRIC_DIAG && E.dense)
    return fail(E, BNUTS_ERR_UNSUPPORTED, "diagonal adaptation on top of a dense metric is not supported");
  // the invariants the reference records as (commented-out) @argcheck's, src/stepsize.jl:183-186
  if (da && !(da->delta > 0.0 && da->delta < 1.0 && da->gamma > 0.0 && da->kappa > 0.5 && da->kappa <= 1.0 && da->t0 >= 0))
    return fail(E, BNUTS_ERR_INVALID_ARGUMENT, "dual averaging needs 0 < delta < 1, gamma > 0, 0.5 < kappa <= 1, t0 >= 0");
  ensure_ctx(E);
  const int C = E.C, D = E.D;
  const uint32_t t0 = E.next_t;
  int collapsed = 0;
  int64_t leap = 0, divs = 0;
#pragma omp parallel for schedule(dynamic) reduction(+ : collapsed, leap, divs)
  for (int c = 0; c < C; ++c) {  // ≙ Threads.@threads over chains, src/mcmc.jl:150-157
    ChainCtx<T>& X = E.ctx[omp_get_thread_num()];
    std::vector<double> stage;  // this chain's draws of the stage, for the metric update
    if (metric_kind == BNUTS_METRIC_DIAG) stage.resize(size_t(N) * D);
    DAState A{};
    if (da) A = da_init(E.eps[c]);  // ≙ src/warmup.jl:284
    bool dead = E.status[c] != 0;
    for (int n = 0; n < N; ++n) {   // ≙ HOT LOOP src/warmup.jl:288 / :324
      double eps = da ? exp_(A.logeps) : E.eps[c];  // ≙ current_ϵ, src/stepsize.jl:235
      if (da && !dead && eps < 1e-10) { dead = true; E.status[c] = BNUTS_ERR_STEPSIZE_COLLAPSE; collapsed += 1; }
      bnuts_tree_stats st{};
      int32_t si = 0;
      if (!dead) {
        st = sample_tree(E, X, c, eps, t0 + uint32_t(n), &si);
        leap += st.steps;
        divs += (st.term_left == st.term_right);
        if (da) A = da_adapt(*da, A, st.acceptance_rate);  // ≙ src/warmup.jl:303
      }
      if (eps_out) eps_out[size_t(c) * N + n] = eps;
      if (chain_out) for (int d = 0; d < D; ++d) chain_out[c * sc + n * sd + d] = double(E.q[size_t(c) * D + d]);
      if (!stage.empty()) for (int d = 0; d < D; ++d) stage[size_t(n) * D + d] = double(E.q[size_t(c) * D + d]);
      if (stats_out) stats_out[c * ssc + n] = st;
      if (sel) sel[size_t(c) * N + n] = si;
    }
    if (metric_kind == BNUTS_METRIC_DIAG && !dead)  // ≙ src/warmup.jl:308-309
      metric_update(stage.data(), D, N, D, lambda < 0 ? 5.0 / N : lambda, &E.Minv[size_t(c) * D], &E.W[size_t(c) * D]);
    if (da && !dead) E.eps[c] = exp_(A.logepsbar);  // ≙ final_ϵ, src/warmup.jl:313
  }
  E.next_t += uint32_t(N);
  E.counters.transitions += int64_t(N) * C;
  E.counters.leapfrogs += leap;
  E.counters.divergences += divs;
  if (collapsed) return fail(E, BNUTS_ERR_STEPSIZE_COLLAPSE, "step size fell below 1e-10 during adaptation");
  return 0;
}

template <class T> int32_t initial_stepsize(Engine<T>& E, const bnuts_stepsize_search& P) {
  if (E.model.kind == bn::MODEL_NONE) return fail(E, BNUTS_ERR_NO_MODEL, "no model set");
  // ≙ the (commented-out) @argcheck's of InitialStepsizeSearch, src/stepsize.jl:31-35
  if (!(P.a_min > 0.0 && P.a_min < P.a_max && P.a_max < 1.0 && P.C > 1.0 && P.eps0 > 0.0 && P.maxiter_crossing > 0 && P.maxiter_bisect > 0))
    return fail(E, BNUTS_ERR_INVALID_ARGUMENT, "step size search needs 0 < a_min < a_max < 1, C > 1, eps0 > 0, positive iteration caps");
  ensure_ctx(E);
  const int C = E.C, D = E.D;
  const uint32_t t = E.next_t;
  int bad = 0;
#pragma omp parallel for schedule(dynamic) reduction(+ : bad)
  for (int c = 0; c < C; ++c) {
    ChainCtx<T>& X = E.ctx[omp_get_thread_num()];
    Ham<T> H{&E, c};
    std::vector<T> p(D);
    const uint32_t gchain = uint32_t(E.cfg.chain_offset + c);
    H.draw_momentum(p.data(), E.seed, gchain, t);  // ≙ src/warmup.jl:195
    PhasePoint<T> z{&E.q[size_t(c) * D], p.data(), &E.g[size_t(c) * D], E.lq[c]};
    const double target = logdensity(H, z);
    if (!isfinite_(target)) { E.status[c] = BNUTS_ERR_NONFINITE_START; bad += 1; continue; }
    LocalAcceptance<T> A{&X, H, z, target};
    double eps = 0;
    const int rc = find_initial_stepsize(P, A, &eps);
    if (rc != 0) { E.status[c] = rc; bad += 1; continue; }
    E.eps[c] = eps;
  }
  E.next_t += 1;
  if (bad) return fail(E, BNUTS_ERR_STEPSIZE_SEARCH, "initial step size search failed for some chains");
  return 0;
}

// ≙ warmup!(FindLocalOptimum), src/warmup.jl:152-186.  The inner solver of the reference
// (QuasiNewtonMethods.proptimize!) is un-vendored and unpinned; the ascent below (Barzilai-Borwein step,
// Armijo backtracking, restart with a doubled penalty on a non-finite start) is this project's choice and is
// restated independently of the device state machine (nuts_machine.h) so the two can be compared bit for bit.
template <class T> int32_t find_local_optimum(Engine<T>& E, double penalty, int32_t iterations) {
  if (E.model.kind == bn::MODEL_NONE) return fail(E, BNUTS_ERR_NO_MODEL, "no model set");
  if (!(penalty >= 0.0) || iterations < 0) return fail(E, BNUTS_ERR_INVALID_ARGUMENT, "bad FindLocalOptimum parameters");
  ensure_ctx(E);
  const int C = E.C, D = E.D;
  int bad = 0;
#pragma omp parallel for schedule(dynamic) reduction(+ : bad)
  for (int c = 0; c < C; ++c) {
    ChainCtx<T>& X = E.ctx[omp_get_thread_num()];
    Ham<T> H{&E, c};
    const uint32_t gchain = uint32_t(E.cfg.chain_offset + c);
    std::vector<T> q(&E.q[size_t(c) * D], &E.q[size_t(c) * D] + D), g(&E.g[size_t(c) * D], &E.g[size_t(c) * D] + D);
    std::vector<T> qn(D), gn(D), dv(D), dn(D);
    double lq = E.lq[c];
    double lambda = penalty;
    int tries = 0, iter = 0;
    bool failed = false;
    E.status[c] = 0;
    for (;;) {   // (re)start
      if (!isfinite_(lq)) {
        tries += 1;
        if (tries > 100) { failed = true; break; }
        lambda += lambda;
        for (int d = 0; d < D; ++d) q[d] = T(bn::restart_position(E.seed, gchain, uint32_t(tries), uint32_t(d)));
        lq = evaluate_l(X, H, q.data(), g.data());
        iter = 0;
        continue;
      }
      const T lam = T(lambda);
      for (int d = 0; d < D; ++d) dv[d] = fma_(-lam, q[d], g[d]);
      double dd = double(dot_warp(dv.data(), dv.data(), D));
      double f = lq - 0.5 * lambda * double(dot_warp(q.data(), q.data(), D));
      double alpha = 1.0 / (1.0 + bn::sqrt_(dd));
      int bt = 0;
      while (iter < iterations && dd > 0.0) {
        const T al = T(alpha);
        for (int d = 0; d < D; ++d) qn[d] = fma_(al, dv[d], q[d]);
        const double lqn = evaluate_l(X, H, qn.data(), gn.data());
        for (int d = 0; d < D; ++d) dn[d] = fma_(-lam, qn[d], gn[d]);
        const T ddn = dot_warp(dn.data(), dn.data(), D);
        const T qqn = dot_warp(qn.data(), qn.data(), D);
        const T dod = dot_warp(dv.data(), dn.data(), D);
        const double fn = isfinite_(lqn) ? lqn - 0.5 * lambda * double(qqn) : -bn::lim<double>::inf();
        if (fn >= f + 1e-4 * alpha * dd) {
          const double curv = dd - double(dod);
          double an = curv > 0.0 ? alpha * dd / curv : 2.0 * alpha;
          an = an < 1e-12 ? 1e-12 : (an > 1e12 ? 1e12 : an);
          q.swap(qn); g.swap(gn); dv.swap(dn); lq = lqn;
          f = fn; dd = double(ddn); alpha = an; iter += 1; bt = 0;
          if (!(dd > 1e-20 * (1.0 + fn * fn))) break;
        } else {
          alpha *= 0.25; bt += 1;
          if (bt > 40) break;
        }
      }
      break;
    }
    if (failed) { E.status[c] = BNUTS_ERR_OPTIMUM; bad += 1; continue; }
    for (int d = 0; d < D; ++d) { E.q[size_t(c) * D + d] = q[d]; E.g[size_t(c) * D + d] = g[d]; }
    E.lq[c] = lq;
  }
  if (bad) return fail(E, BNUTS_ERR_OPTIMUM, "optimization failed to converge for some chains (100 restarts)");
  return 0;
}

template <class T>
int32_t bare_leapfrog(Engine<T>& E, const double* p_in, const double* eps, int nsteps, double* q_out, double* p_out,
                      double* g_out, double* l_out) {
  if (E.model.kind == bn::MODEL_NONE) return fail(E, BNUTS_ERR_NO_MODEL, "no model set");
  if (!p_in || !eps || nsteps < 0) return fail(E, BNUTS_ERR_INVALID_ARGUMENT, "p_in, eps required");
  ensure_ctx(E);
  const int C = E.C, D = E.D;
#pragma omp parallel for schedule(dynamic)
  for (int c = 0; c < C; ++c) {
    ChainCtx<T>& X = E.ctx[omp_get_thread_num()];
    X.arena.reset();
    Ham<T> H{&E, c};
    PhasePoint<T> z;
    z.q = X.arena.alloc(D); z.p = X.arena.alloc(D); z.g = X.arena.alloc(D);
    for (int d = 0; d < D; ++d) {
      z.q[d] = E.q[size_t(c) * D + d]; z.g[d] = E.g[size_t(c) * D + d]; z.p[d] = T(p_in[size_t(c) * D + d]);
    }
    z.lq = E.lq[c];
    for (int s = 0; s < nsteps; ++s) z = leapfrog(X, H, z, T(eps[c]));
    for (int d = 0; d < D; ++d) {
      if (q_out) q_out[size_t(c) * D + d] = double(z.q[d]);
      if (p_out) p_out[size_t(c) * D + d] = double(z.p[d]);
      if (g_out) g_out[size_t(c) * D + d] = double(z.g[d]);
    }
    if (l_out) l_out[c] = double(z.lq);
  }
  return 0;
}

template <class T> Engine<T>* make_engine(const bnuts_config& cfg) {
  auto* E = new Engine<T>();
  E->cfg = cfg; E->C = cfg.n_chains; E->D = cfg.dim; E->seed = cfg.seed;
  const size_t n = size_t(E->C) * E->D;
  E->q.assign(n, T(0)); E->g.assign(n, T(0)); E->lq.assign(E->C, 0.0);
  E->Minv.assign(n, T(1)); E->W.assign(n, T(1));  // ≙ κ = I, src/warmup.jl:102
  E->eps.assign(E->C, 1.0); E->status.assign(E->C, 0);
  E->model.D = E->D;
  return E;
}

#define DISPATCH(e, expr64, expr32) \
  do { if (!(e)) return BNUTS_ERR_INVALID_ARGUMENT; AnyEngine* ae = reinterpret_cast<AnyEngine*>(e); \
       if (ae->dtype == BNUTS_F64) { auto& E = *ae->e64; (void)E; return (expr64); } \
       else { auto& E = *ae->e32; (void)E; return (expr32); } } while (0)

template <class T> int32_t model_simple(Engine<T>& E, int kind) { E.model.kind = kind; return 0; }
template <class T> int32_t model_gaussian(Engine<T>& E, const double* P) {
  if (!P) return fail(E, BNUTS_ERR_INVALID_ARGUMENT, "precision is NULL");
  E.model.kind = bn::MODEL_GAUSSIAN;
  E.model.P.resize(size_t(E.D) * E.D);
  for (size_t i = 0; i < E.model.P.size(); ++i) E.model.P[i] = T(P[i]);
  return 0;
}
inline double bf16_to_double(uint16_t h) { return double(bn::u2f(uint32_t(h) << 16)); }
template <class T>
int32_t model_logistic(Engine<T>& E, const void* X, int32_t xd, const double* y, int64_t N, double tau, int32_t rb) {
  if (!X || !y || N <= 0 || rb <= 0) return fail(E, BNUTS_ERR_INVALID_ARGUMENT, "bad logistic arguments");
  E.model.kind = bn::MODEL_LOGISTIC;
  E.model.N = N; E.model.tau = T(tau); E.model.row_blocks = rb;
  const size_t n = size_t(N) * E.D;
  E.model.X.resize(n); E.model.y.resize(size_t(N));
  for (size_t i = 0; i < n; ++i) {
    double v;
    if (xd == BNUTS_X_F64) v = static_cast<const double*>(X)[i];
    else if (xd == BNUTS_X_F32) v = double(static_cast<const float*>(X)[i]);
    else if (xd == BNUTS_X_BF16) v = bf16_to_double(static_cast<const uint16_t*>(X)[i]);
    else return fail(E, BNUTS_ERR_INVALID_ARGUMENT, "bad x_dtype");
    E.model.X[i] = T(v);
  }
  for (int64_t i = 0; i < N; ++i) E.model.y[size_t(i)] = T(y[i]);
  return 0;
}
template <class T> int32_t set_metric_dense(Engine<T>& E, const double* minv) {
  const int D = E.D;
  const size_t n = size_t(D) * D;
  if (!minv) { E.dense = false; E.MinvD.clear(); E.WD.clear(); E.MinvD64.clear(); return 0; }
  std::vector<double> L(n, 0.0), Li(n, 0.0);
  for (int j = 0; j < D; ++j) {
    double s = minv[size_t(j) * D + j];
    for (int k = 0; k < j; ++k) s -= L[size_t(j) * D + k] * L[size_t(j) * D + k];
    if (!(s > 0.0)) return fail(E, BNUTS_ERR_INVALID_ARGUMENT, "dense metric must be positive definite");
    const double ljj = std::sqrt(s);
    L[size_t(j) * D + j] = ljj;
    for (int i = j + 1; i < D; ++i) {
      double t = minv[size_t(i) * D + j];
      for (int k = 0; k < j; ++k) t -= L[size_t(i) * D + k] * L[size_t(j) * D + k];
      L[size_t(i) * D + j] = t / ljj;
    }
  }
  for (int c = 0; c < D; ++c) {
    Li[size_t(c) * D + c] = 1.0 / L[size_t(c) * D + c];
    for (int i = c + 1; i < D; ++i) {
      double t = 0.0;
      for (int k = c; k < i; ++k) t -= L[size_t(i) * D + k] * Li[size_t(k) * D + c];
      Li[size_t(i) * D + c] = t / L[size_t(i) * D + i];
    }
  }
  E.MinvD.resize(n); E.WD.resize(n); E.MinvD64.assign(minv, minv + n);
  for (int i = 0; i < D; ++i)
    for (int j = 0; j < D; ++j) {
      E.MinvD[size_t(i) * D + j] = T(minv[size_t(i) * D + j]);
      E.WD[size_t(i) * D + j] = T(Li[size_t(j) * D + i]);   // W = L⁻ᵀ
    }
  E.dense = true;
  return 0;
}
template <class T> int32_t set_metric(Engine<T>& E, const double* minv) {
  const size_t n = size_t(E.C) * E.D;
  for (size_t i = 0; i < n; ++i) {
    const double m = minv ? minv[i] : 1.0;
    E.Minv[i] = T(m);
    E.W[i] = T(1.0 / bn::sqrt_(m));  // ≙ src/hamiltonian.jl:53-55
  }
  return 0;
}
template <class T> int32_t get_state(Engine<T>& E, double* q, double* g, double* l) {
  const size_t n = size_t(E.C) * E.D;
  if (q) for (size_t i = 0; i < n; ++i) q[i] = double(E.q[i]);
  if (g) for (size_t i = 0; i < n; ++i) g[i] = double(E.g[i]);
  if (l) for (int c = 0; c < E.C; ++c) l[c] = double(E.lq[c]);
  return 0;
}
template <class T> int32_t inject(Engine<T>& E, int32_t Tn, const uint32_t* dirs, const double* p, const double* exps, int32_t n_exps) {
  if (Tn < 0) return fail(E, BNUTS_ERR_INVALID_ARGUMENT, "T < 0");
  if (exps && n_exps <= 0) return fail(E, BNUTS_ERR_INVALID_ARGUMENT, "exps given with n_exps <= 0");
  E.inj_T = Tn; E.inj_start = E.next_t;
  E.has_inj_dirs = dirs != nullptr; E.has_inj_p = p != nullptr;
  if (dirs) E.inj_dirs.assign(dirs, dirs + size_t(Tn) * E.C);
  if (p) E.inj_p.assign(p, p + size_t(Tn) * E.C * E.D);
  E.inj_exps.clear(); E.inj_nexp = 0;
  if (exps && Tn > 0) { E.inj_exps.assign(exps, exps + size_t(Tn) * E.C * size_t(n_exps)); E.inj_nexp = n_exps; }
  return 0;
}

}  // namespace

extern "C" {

int32_t bnuts_create(const bnuts_config* cfg, bnuts_engine** out) {
  if (!cfg || !out || cfg->n_chains <= 0 || cfg->dim <= 0 || cfg->max_depth <= 0 || cfg->max_depth > 32 ||
      !(cfg->min_delta < 0) || (cfg->dtype != BNUTS_F64 && cfg->dtype != BNUTS_F32)) {
    g_create_error = "invalid bnuts_config";
    return BNUTS_ERR_INVALID_ARGUMENT;
  }
  auto* ae = new AnyEngine();
  ae->dtype = cfg->dtype;
  if (cfg->dtype == BNUTS_F64) ae->e64 = make_engine<double>(*cfg); else ae->e32 = make_engine<float>(*cfg);
  *out = reinterpret_cast<bnuts_engine*>(ae);
  return 0;
}
int32_t bnuts_destroy(bnuts_engine* e) {
  if (!e) return 0;
  auto* ae = reinterpret_cast<AnyEngine*>(e);
  delete ae->e64; delete ae->e32; delete ae;
  return 0;
}
const char* bnuts_last_error(const bnuts_engine* e) {
  if (!e) return g_create_error.c_str();
  auto* ae = reinterpret_cast<const AnyEngine*>(e);
  return ae->dtype == BNUTS_F64 ? ae->e64->err.c_str() : ae->e32->err.c_str();
}
int32_t bnuts_model_iid_normal(bnuts_engine* e) { DISPATCH(e, model_simple(E, bn::MODEL_IID_NORMAL), model_simple(E, bn::MODEL_IID_NORMAL)); }
int32_t bnuts_model_funnel(bnuts_engine* e) { DISPATCH(e, model_simple(E, bn::MODEL_FUNNEL), model_simple(E, bn::MODEL_FUNNEL)); }
int32_t bnuts_model_gaussian(bnuts_engine* e, const double* P) { DISPATCH(e, model_gaussian(E, P), model_gaussian(E, P)); }
int32_t bnuts_model_logistic(bnuts_engine* e, const void* X, int32_t xd, const double* y, int64_t N, double tau, int32_t rb) {
  DISPATCH(e, model_logistic(E, X, xd, y, N, tau, rb), model_logistic(E, X, xd, y, N, tau, rb));
}
// synthetic rows (include/bnuts.h; definition in bnuts_math.h, synth_*): generated here row by row, then the ordinary model
int32_t bnuts_synth_logistic_rows(uint64_t seed, int64_t row0, int64_t nrows, int32_t D, uint16_t* X, double* y, double* beta_true) {
  if (nrows < 0 || row0 < 0 || D <= 0) return BNUTS_ERR_INVALID_ARGUMENT;
  std::vector<double> beta(static_cast<size_t>(D));
  for (int32_t d = 0; d < D; ++d) beta[size_t(d)] = bn::synth_beta(seed, uint32_t(d), D);
  if (beta_true) std::copy(beta.begin(), beta.end(), beta_true);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nrows; ++i) {
    const uint64_t row = uint64_t(row0 + i);
    float x[4];
    if (X)
      for (int32_t q = 0; 4 * q < D; ++q) {
        bn::synth_x4(seed, row, uint32_t(q), x);
        for (int e = 0; e < 4 && 4 * q + e < D; ++e) X[size_t(i) * D + 4 * q + e] = uint16_t(bn::f2u(x[e]) >> 16);
      }
    if (y) y[i] = bn::synth_label(seed, row, D, beta.data());
  }
  return 0;
}
int32_t bnuts_model_logistic_synthetic(bnuts_engine* e, uint64_t seed, int64_t row0, int64_t N, double tau, int32_t rb) {
  if (!e || N <= 0 || row0 < 0) return BNUTS_ERR_INVALID_ARGUMENT;
  AnyEngine* ae = reinterpret_cast<AnyEngine*>(e);
  const int32_t D = ae->dtype == BNUTS_F64 ? ae->e64->D : ae->e32->D;
  std::vector<uint16_t> X(size_t(N) * D);
  std::vector<double> y(static_cast<size_t>(N));
  bnuts_synth_logistic_rows(seed, row0, N, D, X.data(), y.data(), nullptr);
  return bnuts_model_logistic(e, X.data(), BNUTS_X_BF16, y.data(), N, tau, rb);
}
// numerical device of the CUDA tensor path; the oracle computes in plain arithmetic: accepted, no effect
int32_t bnuts_logistic_set_reference(bnuts_engine* e, const double*) { return e ? 0 : BNUTS_ERR_INVALID_ARGUMENT; }
// row sharding is a property of the device engine's data layout; the oracle always sees all rows
int32_t bnuts_set_allreduce(bnuts_engine*, bnuts_allreduce_fn, void*) { return BNUTS_ERR_UNSUPPORTED; }
int32_t bnuts_nccl_unique_id(uint8_t*) { return BNUTS_ERR_UNSUPPORTED; }
int32_t bnuts_p2p_export(bnuts_engine*, uint8_t*) { return BNUTS_ERR_UNSUPPORTED; }
int32_t bnuts_p2p_connect(bnuts_engine*, const uint8_t*, int32_t, int32_t) { return BNUTS_ERR_UNSUPPORTED; }
int32_t bnuts_set_nccl(bnuts_engine*, const uint8_t*, int32_t, int32_t) { return BNUTS_ERR_UNSUPPORTED; }
int32_t bnuts_set_positions(bnuts_engine* e, const double* q) { DISPATCH(e, set_positions(E, q), set_positions(E, q)); }
int32_t bnuts_get_state(bnuts_engine* e, double* q, double* g, double* l) { DISPATCH(e, get_state(E, q, g, l), get_state(E, q, g, l)); }
int32_t bnuts_set_metric_diag(bnuts_engine* e, const double* m) { DISPATCH(e, set_metric(E, m), set_metric(E, m)); }
int32_t bnuts_get_metric_diag_w(bnuts_engine* e, double* w) {
  if (!w) return BNUTS_ERR_INVALID_ARGUMENT;
  DISPATCH(e, ([&] { for (size_t i = 0; i < E.W.size(); ++i) w[i] = double(E.W[i]); return 0; })(),
              ([&] { for (size_t i = 0; i < E.W.size(); ++i) w[i] = double(E.W[i]); return 0; })());
}
int32_t bnuts_set_metric_diag_pair(bnuts_engine* e, const double* m, const double* w) {
  if (!m || !w) return BNUTS_ERR_INVALID_ARGUMENT;
  auto set = [&](auto& E) {
    using T = typename std::remove_reference<decltype(E.W[0])>::type;
    for (size_t i = 0; i < E.W.size(); ++i) { E.Minv[i] = T(m[i]); E.W[i] = T(w[i]); }
    return 0;
  };
  DISPATCH(e, set(E), set(E));
}
int32_t bnuts_set_metric_dense(bnuts_engine* e, const double* m) { DISPATCH(e, set_metric_dense(E, m), set_metric_dense(E, m)); }
int32_t bnuts_get_metric_dense(bnuts_engine* e, double* m) {
  if (!m) return BNUTS_ERR_INVALID_ARGUMENT;
  auto get = [&](auto& E) {
    const size_t D = size_t(E.D);
    for (size_t i = 0; i < D * D; ++i) m[i] = E.dense ? E.MinvD64[i] : ((i / D == i % D) ? 1.0 : 0.0);
    return 0;
  };
  DISPATCH(e, get(E), get(E));
}
int32_t bnuts_get_metric_diag(bnuts_engine* e, double* m) {
  if (!m) return BNUTS_ERR_INVALID_ARGUMENT;
  DISPATCH(e, ([&] { for (size_t i = 0; i < E.Minv.size(); ++i) m[i] = double(E.Minv[i]); return 0; })(),
           ([&] { for (size_t i = 0; i < E.Minv.size(); ++i) m[i] = double(E.Minv[i]); return 0; })());
}
int32_t bnuts_set_stepsize(bnuts_engine* e, const double* eps) {
  if (!eps) return BNUTS_ERR_INVALID_ARGUMENT;
  DISPATCH(e, ([&] { E.eps.assign(eps, eps + E.C); return 0; })(), ([&] { E.eps.assign(eps, eps + E.C); return 0; })());
}
int32_t bnuts_get_stepsize(bnuts_engine* e, double* eps) {
  if (!eps) return BNUTS_ERR_INVALID_ARGUMENT;
  DISPATCH(e, ([&] { std::copy(E.eps.begin(), E.eps.end(), eps); return 0; })(),
           ([&] { std::copy(E.eps.begin(), E.eps.end(), eps); return 0; })());
}
int32_t bnuts_seed(bnuts_engine* e, uint64_t seed, uint32_t next_t) {
  DISPATCH(e, ([&] { E.seed = seed; E.next_t = next_t; return 0; })(), ([&] { E.seed = seed; E.next_t = next_t; return 0; })());
}
int32_t bnuts_get_rng(bnuts_engine* e, uint64_t* seed, uint32_t* next_t) {
  if (!seed || !next_t) return BNUTS_ERR_INVALID_ARGUMENT;
  DISPATCH(e, ([&] { *seed = E.seed; *next_t = E.next_t; return 0; })(), ([&] { *seed = E.seed; *next_t = E.next_t; return 0; })());
}
int32_t bnuts_inject(bnuts_engine* e, int32_t T, const uint32_t* dirs, const double* p, const double* exps, int32_t n_exps) {
  DISPATCH(e, inject(E, T, dirs, p, exps, n_exps), inject(E, T, dirs, p, exps, n_exps));
}
int32_t bnuts_leapfrog(bnuts_engine* e, const double* p_in, const double* eps, int32_t nsteps, double* q_out,
                       double* p_out, double* g_out, double* l_out) {
  DISPATCH(e, bare_leapfrog(E, p_in, eps, nsteps, q_out, p_out, g_out, l_out),
           bare_leapfrog(E, p_in, eps, nsteps, q_out, p_out, g_out, l_out));
}
int32_t bnuts_find_local_optimum(bnuts_engine* e, double penalty, int32_t iterations) {
  DISPATCH(e, find_local_optimum(E, penalty, iterations), find_local_optimum(E, penalty, iterations));
}
int32_t bnuts_find_initial_stepsize(bnuts_engine* e, const bnuts_stepsize_search* P) {
  if (!P) return BNUTS_ERR_INVALID_ARGUMENT;
  DISPATCH(e, initial_stepsize(E, *P), initial_stepsize(E, *P));
}
int32_t bnuts_warmup_stage(bnuts_engine* e, int32_t N, int32_t metric_kind, const bnuts_dual_averaging* da,
                           double lambda, double* chain_out, int64_t sd, int64_t sc, bnuts_tree_stats* stats_out,
                           int64_t ssc, double* eps_out) {
  // da == NULL ≙ FixedStepsize (src/stepsize.jl:251-255): the stage keeps ϵ and only tunes the metric
  DISPATCH(e, run_transitions(E, N, da, metric_kind, lambda, chain_out, sd, sc, stats_out, ssc, nullptr, eps_out),
           run_transitions(E, N, da, metric_kind, lambda, chain_out, sd, sc, stats_out, ssc, nullptr, eps_out));
}
int32_t bnuts_sample(bnuts_engine* e, int32_t N, double* chain_out, int64_t sd, int64_t sc,
                     bnuts_tree_stats* stats_out, int64_t ssc, int32_t* sel) {
  DISPATCH(e, run_transitions(E, N, nullptr, BNUTS_METRIC_NONE, 0.0, chain_out, sd, sc, stats_out, ssc, sel, nullptr),
           run_transitions(E, N, nullptr, BNUTS_METRIC_NONE, 0.0, chain_out, sd, sc, stats_out, ssc, sel, nullptr));
}
int32_t bnuts_counters(bnuts_engine* e, bnuts_counter_block* out) {
  if (!out) return BNUTS_ERR_INVALID_ARGUMENT;
  DISPATCH(e, ([&] { *out = E.counters; return 0; })(), ([&] { *out = E.counters; return 0; })());
}
int32_t bnuts_profile(bnuts_engine* e, int32_t, double* ms, int64_t* n) {
  if (!e) return BNUTS_ERR_INVALID_ARGUMENT;
  if (ms) *ms = 0.0;
  if (n) *n = 0;
  return 0;
}
int32_t bnuts_chain_status(bnuts_engine* e, int32_t* st) {
  if (!st) return BNUTS_ERR_INVALID_ARGUMENT;
  DISPATCH(e, ([&] { std::copy(E.status.begin(), E.status.end(), st); return 0; })(),
           ([&] { std::copy(E.status.begin(), E.status.end(), st); return 0; })());
}

// ---- probes of the shared scalar math, for the known-answer tests (oracle only)
double bnuts_oracle_exp(double x) { return exp_(x); }
double bnuts_oracle_log(double x) { return log_(x); }
double bnuts_oracle_log1p(double x) { return bn::log1p_(x); }
double bnuts_oracle_logaddexp(double x, double y) { return logaddexp_(x, y); }
float bnuts_oracle_expf(float x) { return exp_(x); }
float bnuts_oracle_logf(float x) { return log_(x); }
void bnuts_oracle_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
  bn::u32x4 c{ctr[0], ctr[1], ctr[2], ctr[3]};
  const bn::u32x4 r = bn::philox4x32_10(c, key[0], key[1]);
  out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}
double bnuts_oracle_normal(uint64_t seed, uint32_t chain, uint32_t t, uint32_t d) { return bn::std_normal(seed, chain, t, d, 0.0); }
float bnuts_oracle_normalf(uint64_t seed, uint32_t chain, uint32_t t, uint32_t d) { return bn::std_normal(seed, chain, t, d, 0.0f); }
double bnuts_oracle_exponential(uint64_t seed, uint32_t chain, uint32_t t, uint32_t j, uint32_t k, uint32_t n) {
  return bn::std_exponential(seed, chain, t, j, k, n, 0.0);
}
// decision margins of every chain's last transition (see DecisionTrace); enable = 1 turns tracing on for later calls.
// out [C][5] = (min_div, min_turn, min_sel, scale_H, grad_sq), may be NULL.
int32_t bnuts_oracle_trace(bnuts_engine* e, int32_t enable, double* out) {
  auto f = [&](auto& E) -> int32_t {
    if (out)
      for (size_t c = 0; c < E.trace.size(); ++c) {
        out[5 * c] = E.trace[c].min_div; out[5 * c + 1] = E.trace[c].min_turn; out[5 * c + 2] = E.trace[c].min_sel;
        out[5 * c + 3] = E.trace[c].scale_H; out[5 * c + 4] = E.trace[c].grad_sq;
      }
    if (enable) E.trace.assign(size_t(E.C), DecisionTrace{1e300, 1e300, 1e300, 0.0, 0.0}); else E.trace.clear();
    return 0;
  };
  DISPATCH(e, f(E), f(E));
}
uint32_t bnuts_oracle_directions(uint64_t seed, uint32_t chain, uint32_t t) { return bn::draw_directions(seed, chain, t); }
// ≙ adapt_stepsize table (src/stepsize.jl:220-229): state = {mu, m, Hbar, logeps, logepsbar}
void bnuts_oracle_da_init(double eps, double* state) {
  const DAState A = da_init(eps);
  state[0] = A.mu; state[1] = double(A.m); state[2] = A.Hbar; state[3] = A.logeps; state[4] = A.logepsbar;
}
void bnuts_oracle_da_adapt(const bnuts_dual_averaging* P, double* state, double a) {
  DAState A{state[0], int64_t(state[1]), state[2], state[3], state[4]};
  A = da_adapt(*P, A, a);
  state[0] = A.mu; state[1] = double(A.m); state[2] = A.Hbar; state[3] = A.logeps; state[4] = A.logepsbar;
}
void bnuts_oracle_metric_update(const double* draws, int64_t stride, int32_t N, int32_t D, double lambda,
                                double* minv, double* w) {
  metric_update<double>(draws, stride, N, D, lambda, minv, w);
}
// gradient probe: model of engine e evaluated at q[D] (single vector)
int32_t bnuts_oracle_eval(bnuts_engine* e, const double* q, double* g, double* l) {
  DISPATCH(e, ([&] { std::vector<double> qq(q, q + E.D), gg(E.D), sc; *l = E.model.eval(qq.data(), gg.data(), sc);
                     std::copy(gg.begin(), gg.end(), g); return 0; })(),
           ([&] { std::vector<float> qq(E.D), gg(E.D), sc; for (int d = 0; d < E.D; ++d) qq[d] = float(q[d]);
                  *l = double(E.model.eval(qq.data(), gg.data(), sc)); for (int d = 0; d < E.D; ++d) g[d] = double(gg[d]);
                  return 0; })());
}

}  // extern "C"

// nuts_machine.h — per-chain NUTS state machine (iterative form of the
// reference's recursive tree builder), host/device scalar control code.
//
// The reference builds each trajectory by recursion (adjacent_tree,
// src/tree.jl:321-366, driven by sample_trajectory, src/tree.jl:382-444).  On the
// GPU every chain is a small state machine that advances by exactly one leapfrog
// per "lockstep step": a leaf is consumed, every merge it completes is applied
// (trailing_zeros(leaf number) levels), and the next leapfrog is issued.  Chains
// are independent (src/mcmc.jl:150-157), so a chain that finishes a transition
// starts its next one immediately; nothing forces a barrier between transitions.
//
// This header contains only scalar control.  All D-vector work goes through a
// backend `B` (device: one warp per chain, lane-strided loops + xor-butterfly
// reductions; tests: a host emulation of the same order), so the same control
// code can be exercised on a CPU without a GPU.
//
// Reference lines restated here:
//   leaf / acceptance / divergence      src/NUTS.jl:176-191, :76-78
//   turn statistic merge + test         src/NUTS.jl:118-170, src/tree.jl:230-236
//   proposal merge, biased / unbiased   src/tree.jl:238-263, src/NUTS.jl:32-45
//   invalid subtree bookkeeping         src/tree.jl:340,347-348,358,414-417,438
//   transition entry / statistics       src/NUTS.jl:251-264
//   dual averaging                      src/stepsize.jl:208-241
//   initial step size search            src/stepsize.jl:51-126,150-164
#pragma once
#include "bnuts_math.h"
#include "bnuts_models.h"

namespace bn {

constexpr int MAX_LEVELS = 32;  // max_depth supported by the engine ≙ MAX_DIRECTIONS_DEPTH, src/tree.jl:132 (one UInt32 of directions)

enum Phase : int32_t {
  PH_IDLE = 0,
  PH_START = 1,         // begin a NUTS transition
  PH_NEXT = 2,          // issue the next leapfrog of the current subtree
  PH_LEAF = 3,          // gradient of slot_new pending -> consume leaf
  PH_SEARCH_START = 4,  // initial step size search: draw momentum, first probe
  PH_SEARCH_LEAF = 5,
  PH_BARE_START = 6,    // bare integrator (bnuts_leapfrog)
  PH_BARE_LEAF = 7,
  PH_EVAL_LEAF = 8,     // set_positions: gradient of slot_cur pending
  PH_OPT_START = 9,     // FindLocalOptimum: begin the ascent from slot_cur
  PH_OPT_LEAF = 10,     // gradient of the trial point pending
  PH_OPT_REEVAL = 11    // gradient of a re-randomised start pending
};

// error codes mirrored from include/bnuts.h (kept numerically identical)
constexpr int32_t ST_NONFINITE_START = -4, ST_STEPSIZE_SEARCH = -5, ST_STEPSIZE_COLLAPSE = -6, ST_OPTIMUM_FAILED = -9;

struct TreeStats {  // ≙ TreeStatisticsNUTS, src/NUTS.jl:229-242 (32 bytes)
  double pi, acceptance_rate;
  int32_t term_left, term_right, depth, steps;
};

template <class T> struct ChainState {
  // ---- persistent across transitions
  double eps;                                   // step size (Float64 like the reference)
  double da_mu, da_Hbar, da_logeps, da_logepsbar;  // ≙ DualAveragingState, src/stepsize.jl:196-202
  int32_t da_m;
  int32_t status;
  int32_t phase;
  int32_t remaining;   // transitions (or bare steps) left in this call
  int32_t n_done;      // transitions finished in this call (output index)
  uint32_t t;          // transition counter = RNG position
  int32_t slot_cur;    // slot holding the chain's current (q, ∇ℓ, ℓ)
  int32_t slot_new;    // slot receiving the leapfrog in flight
  // ---- transition-local (≙ locals of sample_trajectory, src/tree.jl:388-393)
  // Energy scalars (ℓ, H, Δ, log-weights) are Float64 for both engine types: the reference is Float64-only and an fp32 ℓ of
  // magnitude N·log 2 has an ulp of 0.03 at N = 1e6, 8 at N = 1e8.  The fp32 variant is fp32 state VECTORS and vector arithmetic.
  double pi0; T teps;
  uint32_t dirs;
  int32_t depth, fwd; uint32_t n; int32_t sp;   // n: leaves of the current doubling built so far (up to 2^31)
  int32_t i_cur, i_minus, i_plus;
  int32_t slot_minus, slot_plus, slot_zeta, i_zeta;
  double omega, pi_zeta, v_lsa;
  int32_t v_steps;
  int32_t n_exp;       // merge exponentials consumed so far in this transition (index into the injected stream)
  // ---- running totals
  int64_t tot_leapfrogs, tot_transitions, tot_divergences;
  // ---- step size search (≙ locals of find_initial_stepsize, src/stepsize.jl:111-126)
  double ss_e0, ss_A0, ss_lo, ss_hi, ss_try, ss_s, ss_a, ss_Cf;
  double ss_target;
  int32_t ss_stage, ss_iter;
  int32_t bare_src;  // slot to restore after the bare integrator
  // ---- local optimum search (≙ warmup!(FindLocalOptimum), src/warmup.jl:152-186)
  double opt_alpha, opt_f, opt_dd, opt_lambda;
  int32_t opt_iter, opt_bt, opt_tries, opt_pad;
  // ---- subtree stack (one entry per pending left sibling)
  double st_omega[MAX_LEVELS], st_pi[MAX_LEVELS], st_vlsa[MAX_LEVELS];
  int32_t st_first_i[MAX_LEVELS], st_slot[MAX_LEVELS], st_izeta[MAX_LEVELS];
};

struct DualAveragingP { double delta, gamma, kappa; int32_t t0; };
struct SearchP { double a_min, a_max, eps0, C; int32_t maxiter_crossing, maxiter_bisect; };
struct OptP { double penalty; int32_t iterations; int32_t pad; };   // ≙ FindLocalOptimum, src/warmup.jl:137-150

template <class T> struct RunParams {
  int32_t max_depth;
  double min_delta;
  uint64_t seed;
  int32_t chain_offset;
  int32_t n_chains;
  int32_t n_slots;
  int32_t da_on;
  DualAveragingP da;
  SearchP search;
  OptP opt;
  // injection (≙ p = / directions = of sample_tree, src/NUTS.jl:251-258)
  int32_t inj_T;
  uint32_t inj_start;
  const uint32_t* inj_dirs;  // [T][C] or null
  const double* inj_p;       // [T][C][D] or null
  const double* inj_exps;    // [T][C][inj_nexp] or null: the k-th exponential CONSUMED by rand_bool_logprob (src/NUTS.jl:32-34)
  int32_t inj_nexp;
  // outputs of the current call (device / host-emulation buffers), any may be null
  TreeStats* stats;          // [C][N]
  int32_t* sel;              // [C][N]
  double* eps_hist;          // [C][N]
  int32_t n_total;           // N of this call
};

// ≙ initial_adaptation_state, src/stepsize.jl:208-212
template <class T> BN_HD void da_init(ChainState<T>& s) {
  const double le = log_(s.eps);
  s.da_mu = log_(10.0) + le;
  s.da_m = 0;
  s.da_Hbar = 0.0;
  s.da_logeps = le;
  s.da_logepsbar = 0.0;
}
// ≙ adapt_stepsize, src/stepsize.jl:220-229
template <class T> BN_HD void da_adapt(ChainState<T>& s, const DualAveragingP& P, double a) {
  s.da_m += 1;
  const double m = (double)s.da_m;
  s.da_Hbar += (P.delta - a - s.da_Hbar) / (m + (double)P.t0);
  s.da_logeps = s.da_mu - sqrt_(m) / P.gamma * s.da_Hbar;
  s.da_logepsbar += exp_(-P.kappa * log_(m)) * (s.da_logeps - s.da_logepsbar);
}

// The backend B provides, for the chain it is bound to:
//   bool lane0(); void sync();
//   double get_lq(slot); void set_lq(slot, double);      (energy scalars are Float64 for both engine types)
//   void start_tx(slot, const double* inj_p_or_null, seed, gchain, t, T* Ksum)
//            p <- W .* N(0,1) (or injected); p# <- M^-1 p; main rho <- p; main p#- = p#+ <- p#; Ksum = sum p# p
//   void pre_kick_drift(src_slot, dst_slot, T eh, T eps)       // ≙ src/kinetic_energy.jl:146-150
//   T    post_kick(slot, T eh, int push_level)                 // ≙ :159-161 + kinetic energy + p#; returns sum p# p
//   void merge_sub(int L, int rhoR_is_leaf, int slot_leaf, bool fwd, T* dm, T* dp)   // stack level merge, in place
//   void merge_top(int rhoR_is_leaf, int slot_leaf, bool fwd, T* dm, T* dp)          // main tree merge, in place
//   double model_grad(slot)          // elementwise models: evaluate now; GEMM models: finalize staged result
//   void emit_draw(slot, n_done)     // copy q out as Float64
//   void bare_load_p(slot); void bare_emit(slot);
//   void opt_trial(src, dst, T alpha, T lambda)   q_dst = q + alpha (∇ℓ − λ q), staged for its gradient
//   void opt_norms(slot, T lambda, T* dd, T* qq)   |∇ℓ − λq|², |q|²
//   void opt_dots(src, dst, T lambda, T* dd_new, T* qq_new, T* dod)   also d_old·d_new
//   void opt_restart_position(slot, seed, gchain, attempt)
template <class T, class B> struct Machine {
  B& b;
  ChainState<T>& s;        // register/working copy of the scalars
  ChainState<T>* g;        // backing store (stack arrays are accessed here)
  const RunParams<T>& rp;
  int32_t c;               // local chain index

  BN_HD Machine(B& b_, ChainState<T>& s_, ChainState<T>* g_, const RunParams<T>& rp_, int32_t c_)
      : b(b_), s(s_), g(g_), rp(rp_), c(c_) {}

  BN_HD uint32_t gchain() const { return (uint32_t)(rp.chain_offset + c); }

  // ≙ logdensity(H, z), src/kinetic_energy.jl:107-112
  BN_HD static double hamiltonian(double lq, T Ksum) {
    if (!isfinite_(lq)) return -lim<double>::inf();
    const T K = T(0.5) * Ksum;
    return lq - (isfinite_(K) ? (double)K : lim<double>::inf());
  }

  BN_HD int32_t alloc_slot() const {
    uint64_t used = (1ull << s.slot_cur) | (1ull << s.slot_minus) | (1ull << s.slot_plus) | (1ull << s.slot_zeta);
    for (int k = 0; k < s.sp; ++k) used |= 1ull << g->st_slot[k];
    int32_t f = 0;
    while (used & (1ull << f)) ++f;
    return f;  // n_slots = max_depth + 4 <= 36 guarantees f < n_slots
  }

  // ≙ rand_bool_logprob, src/NUTS.jl:32-34
  // The reference consumes one randexp(rng) per call with logprob2 < 0, in call order (post-order over the merges of the
  // recursion, which is the order this machine performs them in); an injected stream is consumed the same way, the
  // engine's own draws are a pure function of the merge's position in the tree.
  BN_HD bool select_second(double logprob2, uint32_t j, uint32_t k, uint32_t n) {
    if (logprob2 >= 0.0) return true;
    T e;   // the variate is a T (the engine type's stream); the comparison is Float64
    const bool injected = rp.inj_exps && rp.inj_T > 0 && s.t >= rp.inj_start && s.t < rp.inj_start + (uint32_t)rp.inj_T;
    if (injected && s.n_exp < rp.inj_nexp)
      e = T(rp.inj_exps[((int64_t)(s.t - rp.inj_start) * rp.n_chains + c) * rp.inj_nexp + s.n_exp]);
    else
      e = std_exponential(rp.seed, gchain(), s.t, j, k, n, T(0));
    s.n_exp += 1;
    return (double)e > -logprob2;
  }

  // ---------------------------------------------------------------- transition entry
  // ≙ sample_tree src/NUTS.jl:251-260 + initial leaf src/tree.jl:388-393
  BN_HD void start_transition() {
    if (rp.da_on) {
      s.eps = exp_(s.da_logeps);  // ≙ current_ϵ, src/stepsize.jl:235
      if (s.eps < 1e-10) {        // ≙ src/warmup.jl:291-296
        s.status = ST_STEPSIZE_COLLAPSE;
        s.phase = PH_IDLE;
        return;
      }
    }
    if (rp.eps_hist && b.lane0()) rp.eps_hist[(int64_t)c * rp.n_total + s.n_done] = s.eps;
    s.teps = T(s.eps);
    const bool injected = rp.inj_T > 0 && s.t >= rp.inj_start && s.t < rp.inj_start + (uint32_t)rp.inj_T;
    const int64_t it = injected ? (int64_t)(s.t - rp.inj_start) : 0;
    s.dirs = draw_directions(rp.seed, gchain(), s.t);
    if (injected && rp.inj_dirs) s.dirs = rp.inj_dirs[it * rp.n_chains + c];
    const double* ip = (injected && rp.inj_p) ? rp.inj_p + (it * rp.n_chains + c) * (int64_t)b.dim() : nullptr;
    T Ksum;
    b.start_tx(s.slot_cur, ip, rp.seed, gchain(), s.t, &Ksum);
    s.pi0 = hamiltonian(b.get_lq(s.slot_cur), Ksum);
    s.slot_zeta = s.slot_cur; s.omega = 0.0; s.pi_zeta = s.pi0; s.i_zeta = 0;
    s.v_lsa = -lim<double>::inf(); s.v_steps = 0; s.n_exp = 0;
    s.slot_minus = s.slot_plus = s.slot_cur;
    s.i_minus = s.i_plus = 0;
    s.depth = 0;
    begin_doubling();
  }

  // ≙ loop head of sample_trajectory, src/tree.jl:395-404
  BN_HD void begin_doubling() {
    if (s.depth >= rp.max_depth) { finish(1, 0); return; }  // REACHED_MAX_DEPTH
    s.fwd = (int32_t)(s.dirs & 1u);
    s.dirs >>= 1;
    s.slot_cur = s.fwd ? s.slot_plus : s.slot_minus;
    s.i_cur = s.fwd ? s.i_plus : s.i_minus;
    s.n = 0; s.sp = 0;
    s.phase = PH_NEXT;
  }

  // ≙ move + first half of leapfrog, src/NUTS.jl:18-21, src/kinetic_energy.jl:144-150
  BN_HD void pre() {
    s.slot_new = alloc_slot();
    const T e = s.fwd ? s.teps : -s.teps;
    b.pre_kick_drift(s.slot_cur, s.slot_new, T(0.5) * e, e);
    s.phase = PH_LEAF;
  }

  // fold the visited statistic of an invalid subtree up through its pending left
  // siblings and into the tree (src/tree.jl:347-348 on the way out, :414 at the top)
  BN_HD void fold_invalid(double acc_v) {
    for (int k = s.sp - 1; k >= 0; --k) acc_v = logaddexp_(g->st_vlsa[k], acc_v);
    s.v_lsa = logaddexp_(s.v_lsa, acc_v);
  }

  // ≙ second half of leapfrog + leaf + every merge this leaf completes
  BN_HD void post_leaf() {
    double lq = b.model_grad(s.slot_new);
    if (!isfinite_(lq)) lq = -lim<double>::inf();  // ≙ evaluate_ℓ!, src/kinetic_energy.jl:80-84
    b.set_lq(s.slot_new, lq);
    const T e = s.fwd ? s.teps : -s.teps;
    const uint32_t n1 = s.n + 1u;
    const bool push_leaf = (n1 & 1u) && s.depth > 0;
    const T Ksum = b.post_kick(s.slot_new, T(0.5) * e, push_leaf ? s.sp : -1);
    s.slot_cur = s.slot_new;
    s.i_cur += s.fwd ? 1 : -1;
    s.n = n1;
    // ≙ leaf, src/NUTS.jl:176-191
    const double Hz = hamiltonian(lq, Ksum);
    const double delta = Hz - s.pi0;
    const bool isdiv = delta < rp.min_delta;
    double acc_v = (delta < 0.0) ? delta : 0.0;  // ≙ leaf_acceptance_statistic, src/NUTS.jl:76-78
    s.v_steps += 1;
    if (isdiv) {  // ≙ InvalidTree(i′), src/tree.jl:332,340,348,417
      fold_invalid(acc_v);
      finish(s.i_cur, s.i_cur);
      return;
    }
    double acc_omega = delta, acc_pi = Hz;
    int32_t acc_slot = s.slot_cur, acc_iz = s.i_cur, acc_first = s.i_cur;
    int rhoR_is_leaf = 1;
    int m = 0;
    while (!((n1 >> m) & 1u)) ++m;  // trailing zeros; n1 <= 2^depth so m <= depth
    for (int k = 1; k <= m; ++k) {  // ≙ adjacent_tree depth-k body, src/tree.jl:347-364
      const int L = s.sp - 1;
      acc_v = logaddexp_(g->st_vlsa[L], acc_v);
      T dm, dp;
      b.merge_sub(L, rhoR_is_leaf, s.slot_cur, s.fwd != 0, &dm, &dp);
      rhoR_is_leaf = 0;
      if ((dm < T(0)) | (dp < T(0))) {  // ≙ is_turning -> InvalidTree(i′, i₊), src/tree.jl:358
        const int32_t left = g->st_first_i[L];
        s.sp = L;
        fold_invalid(acc_v);
        finish(left, s.i_cur);
        return;
      }
      const double wl = g->st_omega[L];
      const double w = logaddexp_(wl, acc_omega);
      const double logprob2 = acc_omega - w;  // unbiased, src/tree.jl:261-263 with bias = false
      if (!select_second(logprob2, (uint32_t)s.depth, (uint32_t)k, n1)) {
        acc_slot = g->st_slot[L]; acc_pi = g->st_pi[L]; acc_iz = g->st_izeta[L];
      }
      acc_omega = w;
      acc_first = g->st_first_i[L];
      s.sp = L;
    }
    if (n1 == (1u << s.depth)) {
      top_merge(acc_v, acc_omega, acc_pi, acc_slot, acc_iz, rhoR_is_leaf);
      return;
    }
    const int P = s.sp;
    if (b.lane0()) {
      g->st_omega[P] = acc_omega; g->st_pi[P] = acc_pi; g->st_vlsa[P] = acc_v;
      g->st_first_i[P] = acc_first; g->st_slot[P] = acc_slot; g->st_izeta[P] = acc_iz;
    }
    b.sync();
    s.sp = P + 1;
    s.phase = PH_NEXT;
  }

  // ≙ sample_trajectory after adjacent_tree returned valid, src/tree.jl:414-438
  BN_HD void top_merge(double acc_v, double acc_omega, double acc_pi, int32_t acc_slot, int32_t acc_iz, int rhoR_is_leaf) {
    s.v_lsa = logaddexp_(s.v_lsa, acc_v);
    if (s.fwd) { s.slot_plus = s.slot_cur; s.i_plus = s.i_cur; }
    else       { s.slot_minus = s.slot_cur; s.i_minus = s.i_cur; }
    const double w = logaddexp_(s.omega, acc_omega);
    const double logprob2 = acc_omega - s.omega;  // biased progressive, bias = true
    if (select_second(logprob2, (uint32_t)s.depth, 0u, 0u)) {
      s.slot_zeta = acc_slot; s.pi_zeta = acc_pi; s.i_zeta = acc_iz;
    }
    s.omega = w;
    s.depth += 1;
    T dm, dp;
    b.merge_top(rhoR_is_leaf, s.slot_cur, s.fwd != 0, &dm, &dp);
    if ((dm < T(0)) | (dp < T(0))) { finish(s.i_minus, s.i_plus); return; }  // src/tree.jl:438
    begin_doubling();
  }

  // ≙ tail of sample_tree (src/NUTS.jl:262) + per-transition work of warmup!/mcmc!
  //   (src/warmup.jl:297-303, :325-326)
  BN_HD void finish(int32_t term_left, int32_t term_right) {
    const double a0 = exp_(s.v_lsa) / (double)s.v_steps;  // ≙ acceptance_rate, src/NUTS.jl:84
    const double a = a0 < 1.0 ? a0 : 1.0;
    if (b.lane0()) {
      if (rp.stats) {
        TreeStats st;
        st.pi = s.pi_zeta; st.acceptance_rate = a;
        st.term_left = term_left; st.term_right = term_right;
        st.depth = s.depth; st.steps = s.v_steps;
        rp.stats[(int64_t)c * rp.n_total + s.n_done] = st;
      }
      if (rp.sel) rp.sel[(int64_t)c * rp.n_total + s.n_done] = s.i_zeta;
    }
    s.slot_cur = s.slot_zeta;
    b.emit_draw(s.slot_cur, s.n_done);
    if (rp.da_on) da_adapt(s, rp.da, a);
    s.tot_leapfrogs += s.v_steps;
    s.tot_transitions += 1;
    s.tot_divergences += (term_left == term_right);
    s.n_done += 1;
    s.t += 1;
    s.remaining -= 1;
    s.sp = 0;
    s.slot_minus = s.slot_plus = s.slot_zeta;
    s.phase = s.remaining > 0 ? PH_START : PH_IDLE;
  }

  // ---------------------------------------------------------------- step size search
  // ≙ warmup!(InitialStepsizeSearch), src/warmup.jl:188-200
  BN_HD void search_start() {
    T Ksum;
    b.start_tx(s.slot_cur, nullptr, rp.seed, gchain(), s.t, &Ksum);
    s.ss_target = hamiltonian(b.get_lq(s.slot_cur), Ksum);
    s.slot_minus = s.slot_plus = s.slot_zeta = s.slot_cur; s.sp = 0;
    if (!isfinite_(s.ss_target)) { s.status = ST_NONFINITE_START; s.phase = PH_IDLE; return; }
    s.ss_stage = 0; s.ss_iter = 0; s.ss_try = rp.search.eps0;
    search_pre();
  }
  BN_HD void search_pre() {
    s.slot_new = alloc_slot();
    const T e = T(s.ss_try);
    b.pre_kick_drift(s.slot_cur, s.slot_new, T(0.5) * e, e);
    s.phase = PH_SEARCH_LEAF;
  }
  BN_HD void search_done(double eps) { s.eps = eps; s.phase = PH_IDLE; }
  BN_HD void search_fail() { s.status = ST_STEPSIZE_SEARCH; s.phase = PH_IDLE; }
  // ≙ find_initial_stepsize / find_crossing_stepsize / bisect_stepsize, src/stepsize.jl:51-126
  BN_HD void search_post() {
    double lq = b.model_grad(s.slot_new);
    if (!isfinite_(lq)) lq = -lim<double>::inf();
    b.set_lq(s.slot_new, lq);
    const T e = T(s.ss_try);
    const T Ksum = b.post_kick(s.slot_new, T(0.5) * e, -1);
    const double A = exp_(hamiltonian(lq, Ksum) - s.ss_target);  // ≙ local_acceptance_ratio
    const SearchP& P = rp.search;
    const bool inside = (P.a_min <= A) && (A <= P.a_max);
    if (s.ss_stage == 0) {
      if (inside) { search_done(s.ss_try); return; }
      s.ss_e0 = s.ss_try; s.ss_A0 = A;
      const bool above = A > P.a_max;
      s.ss_s = above ? 1.0 : -1.0;
      s.ss_a = above ? P.a_max : P.a_min;
      s.ss_Cf = above ? P.C : 1.0 / P.C;
      s.ss_stage = 1; s.ss_iter = 0;
      if (P.maxiter_crossing <= 0) { search_fail(); return; }
      s.ss_try = s.ss_e0 * s.ss_Cf;
    } else if (s.ss_stage == 1) {
      if (s.ss_s * (A - s.ss_a) <= 0) {
        if (inside) { search_done(s.ss_try); return; }
        if (s.ss_e0 < s.ss_try) { s.ss_lo = s.ss_e0; s.ss_hi = s.ss_try; }
        else                    { s.ss_lo = s.ss_try; s.ss_hi = s.ss_e0; }
        s.ss_stage = 2; s.ss_iter = 0;
        if (P.maxiter_bisect <= 0) { search_fail(); return; }
        s.ss_try = 0.5 * (s.ss_lo + s.ss_hi);
      } else {
        s.ss_e0 = s.ss_try; s.ss_A0 = A;
        s.ss_iter += 1;
        if (s.ss_iter >= P.maxiter_crossing) { search_fail(); return; }
        s.ss_try = s.ss_e0 * s.ss_Cf;
      }
    } else {
      if (inside) { search_done(s.ss_try); return; }
      if (A < P.a_min) s.ss_hi = s.ss_try; else s.ss_lo = s.ss_try;
      s.ss_iter += 1;
      if (s.ss_iter >= P.maxiter_bisect) { search_fail(); return; }
      s.ss_try = 0.5 * (s.ss_lo + s.ss_hi);
    }
    search_pre();
  }

  // ---------------------------------------------------------------- local optimum
  // ≙ warmup!(FindLocalOptimum), src/warmup.jl:152-186: maximise ℓ(q) − ½λ‖q‖² for at most `iterations`
  // steps ("we don't need to find the mode, just be in a reasonable region", :146-147); a non-finite result
  // re-randomises q and doubles λ, up to 100 times (:162-172).  The reference delegates the inner solver to
  // QuasiNewtonMethods.proptimize! (un-vendored, unpinned); here it is gradient ascent with a
  // Barzilai-Borwein step and Armijo backtracking, one gradient request per trial point, so thousands of
  // chains search in lockstep through the same batched gradient kernels as the sampler.
  BN_HD void opt_start() {
    s.opt_lambda = rp.opt.penalty; s.opt_iter = 0; s.opt_tries = 0;
    s.status = 0;
    opt_begin();
  }
  BN_HD void opt_begin() {
    s.slot_minus = s.slot_plus = s.slot_zeta = s.slot_cur; s.sp = 0;
    const double lq = b.get_lq(s.slot_cur);
    if (!isfinite_(lq)) { opt_restart(); return; }
    T dd, qq;
    b.opt_norms(s.slot_cur, T(s.opt_lambda), &dd, &qq);
    s.opt_f = lq - 0.5 * s.opt_lambda * (double)qq;
    s.opt_dd = (double)dd;
    s.opt_alpha = 1.0 / (1.0 + sqrt_(s.opt_dd));       // first step no longer than 1
    s.opt_bt = 0;
    if (s.opt_iter >= rp.opt.iterations || !(s.opt_dd > 0.0)) { opt_done(); return; }
    opt_pre();
  }
  BN_HD void opt_restart() {   // ≙ src/warmup.jl:168-170
    s.opt_tries += 1;
    if (s.opt_tries > 100) { s.status = ST_OPTIMUM_FAILED; s.phase = PH_IDLE; return; }
    s.opt_lambda += s.opt_lambda;
    b.opt_restart_position(s.slot_cur, rp.seed, gchain(), (uint32_t)s.opt_tries);
    s.phase = PH_OPT_REEVAL;
  }
  BN_HD void opt_reeval_post() {
    double lq = b.model_grad(s.slot_cur);
    if (!isfinite_(lq)) lq = -lim<double>::inf();
    b.set_lq(s.slot_cur, lq);
    s.opt_iter = 0;
    opt_begin();
  }
  BN_HD void opt_pre() {
    s.slot_new = alloc_slot();
    b.opt_trial(s.slot_cur, s.slot_new, T(s.opt_alpha), T(s.opt_lambda));
    s.phase = PH_OPT_LEAF;
  }
  BN_HD void opt_done() {
    s.slot_minus = s.slot_plus = s.slot_zeta = s.slot_cur; s.sp = 0;
    s.phase = PH_IDLE;
  }
  BN_HD void opt_post() {
    double lq = b.model_grad(s.slot_new);
    if (!isfinite_(lq)) lq = -lim<double>::inf();
    b.set_lq(s.slot_new, lq);
    T ddn, qqn, dod;
    b.opt_dots(s.slot_cur, s.slot_new, T(s.opt_lambda), &ddn, &qqn, &dod);
    const double fn = isfinite_(lq) ? lq - 0.5 * s.opt_lambda * (double)qqn : -lim<double>::inf();
    const bool accept = fn >= s.opt_f + 1e-4 * s.opt_alpha * s.opt_dd;   // Armijo; false for -Inf / NaN
    if (accept) {
      const double curv = s.opt_dd - (double)dod;      // -(s·y)/alpha with s = alpha d, y = d' - d
      double an = curv > 0.0 ? s.opt_alpha * s.opt_dd / curv : 2.0 * s.opt_alpha;   // Barzilai-Borwein
      an = an < 1e-12 ? 1e-12 : (an > 1e12 ? 1e12 : an);
      s.slot_cur = s.slot_new; s.opt_f = fn; s.opt_dd = (double)ddn; s.opt_alpha = an;
      s.opt_iter += 1; s.opt_bt = 0;
      if (s.opt_iter >= rp.opt.iterations || !((double)ddn > 1e-20 * (1.0 + fn * fn))) { opt_done(); return; }
    } else {
      s.opt_alpha *= 0.25; s.opt_bt += 1;
      if (s.opt_bt > 40) { opt_done(); return; }
    }
    opt_pre();
  }

  // ---------------------------------------------------------------- bare integrator
  // ≙ stack leapfrog, src/kinetic_energy.jl:164-195 (engine state is restored)
  BN_HD void bare_start() {
    s.bare_src = s.slot_cur;
    s.slot_minus = s.slot_plus = s.slot_zeta = s.slot_cur; s.sp = 0;
    b.bare_load_p(s.slot_cur);
    s.teps = T(s.ss_try);  // signed step supplied by the host in ss_try
    if (s.remaining <= 0) { bare_finish(); return; }
    bare_pre();
  }
  BN_HD void bare_pre() {
    s.slot_new = alloc_slot();
    b.pre_kick_drift(s.slot_cur, s.slot_new, T(0.5) * s.teps, s.teps);
    s.phase = PH_BARE_LEAF;
  }
  BN_HD void bare_post() {
    double lq = b.model_grad(s.slot_new);
    if (!isfinite_(lq)) lq = -lim<double>::inf();
    b.set_lq(s.slot_new, lq);
    (void)b.post_kick(s.slot_new, T(0.5) * s.teps, -1);
    s.slot_cur = s.slot_new;
    s.remaining -= 1;
    if (s.remaining <= 0) { bare_finish(); return; }
    bare_pre();
  }
  BN_HD void bare_finish() {
    b.bare_emit(s.slot_cur);
    s.slot_cur = s.bare_src;
    s.slot_minus = s.slot_plus = s.slot_zeta = s.slot_cur;
    s.phase = PH_IDLE;
  }

  // ---------------------------------------------------------------- set_positions
  // ≙ initialize_warmup_state, src/warmup.jl:119-124 (q already written to slot_cur)
  BN_HD void eval_post() {
    double lq = b.model_grad(s.slot_cur);
    if (!isfinite_(lq)) lq = -lim<double>::inf();
    b.set_lq(s.slot_cur, lq);
    s.status = isfinite_(lq) ? 0 : ST_NONFINITE_START;
    s.phase = PH_IDLE;
  }

  // Consume a pending gradient (if any), then run until the next gradient is
  // requested or the chain goes idle.  Returns true if a gradient is pending.
  BN_HD bool step() {
    switch (s.phase) {
      case PH_LEAF: post_leaf(); break;
      case PH_SEARCH_LEAF: search_post(); break;
      case PH_BARE_LEAF: bare_post(); break;
      case PH_EVAL_LEAF: eval_post(); break;
      case PH_OPT_LEAF: opt_post(); break;
      case PH_OPT_REEVAL: opt_reeval_post(); break;
      default: break;
    }
    if (s.phase == PH_START) start_transition();
    else if (s.phase == PH_SEARCH_START) search_start();
    else if (s.phase == PH_BARE_START) bare_start();
    else if (s.phase == PH_OPT_START) opt_start();
    if (s.phase == PH_NEXT) pre();
    return s.phase == PH_LEAF || s.phase == PH_SEARCH_LEAF || s.phase == PH_BARE_LEAF || s.phase == PH_EVAL_LEAF ||
           s.phase == PH_OPT_LEAF || s.phase == PH_OPT_REEVAL;
  }
};

}  // namespace bn

// bnuts_models.h — scalar formulas of the benchmark targets.
//
// The reference has no models: the log density is a user-supplied
// AbstractProbabilityModel reached through logdensity_and_gradient!
// (src/kinetic_energy.jl:73,89).  The four targets BASELINE.json names are
// defined here once, as host/device scalar code, so the CUDA kernels and the CPU
// oracle evaluate exactly the same expression tree.  Vector reductions follow
// the "warp order" convention: lane l accumulates the groups of four consecutive elements 4l..4l+3 (+128, ...) with fma
// in index order, then a xor-butterfly (16,8,4,2,1).
#pragma once
#include "bnuts_math.h"

namespace bn {

enum ModelKind : int32_t {
  MODEL_NONE = 0,
  MODEL_IID_NORMAL = 1,  // l = -1/2 q'q
  MODEL_GAUSSIAN = 2,    // l = -1/2 q'Pq, P dense precision
  MODEL_LOGISTIC = 3,    // l = sum_i [y_i eta_i - softplus(eta_i)] - tau/2 b'b
  MODEL_FUNNEL = 4       // Neal's funnel, v = q[0]
};

// iid normal: gradient element and value from S = sum q^2
template <class T> BN_HD T iid_grad(T q) { return -q; }
template <class T> BN_HD T iid_value(T S) { return T(-0.5) * S; }

// funnel: v = q[0], S = sum_{d>=1} q_d^2, e = exp(-v)
//   l   = -v^2/18 - (D-1)/2 v - 1/2 e S
//   dv  = -v/9 - (D-1)/2 + 1/2 e S
//   dqd = -e q_d
template <class T> BN_HD T funnel_value(T v, T S, T e, int D) {
  const T hd = T(0.5) * T(D - 1);
  T l = -(v * v) / T(18);
  l = fma_(-hd, v, l);
  l = fma_(T(-0.5) * e, S, l);
  return l;
}
template <class T> BN_HD T funnel_grad_v(T v, T S, T e, int D) {
  const T hd = T(0.5) * T(D - 1);
  T g = -v / T(9) - hd;
  g = fma_(T(0.5) * e, S, g);
  return g;
}
template <class T> BN_HD T funnel_grad_x(T q, T e) { return -(e * q); }

// logistic: per (row, chain) element from eta = x_i . beta
//   t = exp(-|eta|); sigma = eta >= 0 ? 1/(1+t) : t/(1+t); softplus = max(eta,0) + log1p(t)
template <class T> BN_HD void logistic_elem(T eta, T y, T* resid, T* lterm) {
  const T a = eta < T(0) ? -eta : eta;
  const T t = exp_(-a);
  const T d = T(1) + t;
  const T sig = (eta >= T(0)) ? T(1) / d : t / d;
  const T sp = (eta > T(0) ? eta : T(0)) + log1p_(t);
  *resid = y - sig;
  *lterm = fma_(y, eta, -sp);
}

}  // namespace bn

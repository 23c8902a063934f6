// cpu_fast.cpp — NON-PARITY CPU leg of the benchmark (test / measurement infrastructure, like the rest of oracle/).
//
// The bit-exact oracle (bnuts_oracle.cpp) evaluates the logistic-regression gradient with strictly sequential fma
// chains and hand-built transcendentals so that CPU and GPU agree bit for bit; that makes it latency-bound and
// about an order of magnitude slower than what the reference's own stack would do on the same cores
// (LoopVectorization `@avx` over the rows, SLEEFPirates exp, one chain per Julia thread: src/mcmc.jl:150-157,
// src/kinetic_energy.jl:144-161 around the model call :73).  This file is that "vectorised" CPU number:
// the same leapfrog + gradient in Float64 arithmetic, SIMD over the coordinates of a row and over row blocks,
// one chain per OpenMP thread, no fixed summation order.  bench.py reports it as cpu_baseline.vectorised next
// to the bit-exact port; tests/test_cpu_fast.py checks it against the oracle to 1e-9 (it is never a parity
// reference itself, and the product never links it).
//
// X~ is the sign-folded design matrix ((2y-1)·x, Float32 storage: the bf16-grid values are exact in it), so
//   l(q) = sum_i log sigma(eta_i) - tau/2 |q|^2,   grad = X~' sigma(-eta) - tau q,   eta = X~ q.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <omp.h>

namespace {

// exp(x) for x <= 0 (we only need exp(-|eta|)): Cody-Waite reduction + degree-11 Taylor on |r| <= ln2/2; branch-free
// so the row-block loop vectorises.  Relative error < 3e-16 on [-700, 0].
inline double exp_neg(double x) {
  x = x < -700.0 ? -700.0 : x;
  const double kf = std::nearbyint(x * 1.4426950408889634);
  const double r = std::fma(kf, -1.9082149292705877e-10, std::fma(kf, -0.6931471803691238, x));
  double p = 1.0 / 39916800.0;
  p = std::fma(p, r, 1.0 / 3628800.0); p = std::fma(p, r, 1.0 / 362880.0); p = std::fma(p, r, 1.0 / 40320.0);
  p = std::fma(p, r, 1.0 / 5040.0); p = std::fma(p, r, 1.0 / 720.0); p = std::fma(p, r, 1.0 / 120.0);
  p = std::fma(p, r, 1.0 / 24.0); p = std::fma(p, r, 1.0 / 6.0); p = std::fma(p, r, 0.5);
  p = std::fma(p, r, 1.0); p = std::fma(p, r, 1.0);
  int64_t bits; std::memcpy(&bits, &p, 8);
  bits += static_cast<int64_t>(kf) << 52;
  double out; std::memcpy(&out, &bits, 8);
  return out;
}

constexpr int RB = 16;   // rows per block

__attribute__((target_clones("avx512f", "avx2", "default")))
double grad_one(const float* __restrict__ X, int64_t N, int D, double tau, const double* __restrict__ q, double* __restrict__ g) {
  std::vector<double> acc(static_cast<size_t>(D), 0.0);
  double* a = acc.data();
  double lsum = 0.0;
  for (int64_t i0 = 0; i0 < N; i0 += RB) {
    const int nr = static_cast<int>(N - i0 < RB ? N - i0 : RB);
    double eta[RB], r[RB];
    for (int k = 0; k < nr; ++k) {
      const float* x = X + (i0 + k) * D;
      double s = 0.0;
#pragma omp simd reduction(+ : s)
      for (int d = 0; d < D; ++d) s += static_cast<double>(x[d]) * q[d];
      eta[k] = s;
    }
    for (int k = nr; k < RB; ++k) eta[k] = 0.0;
    double lb = 0.0, prod = 1.0;
#pragma omp simd reduction(+ : lb) reduction(* : prod)
    for (int k = 0; k < RB; ++k) {
      const double e = eta[k];
      const double t = exp_neg(-std::fabs(e));          // exp(-|eta|)
      const double d1 = 1.0 + t;
      const double inv = 1.0 / d1;
      r[k] = e >= 0.0 ? t * inv : inv;                   // sigma(-eta)
      lb += e < 0.0 ? e : 0.0;                           // log sigma(eta) = min(eta, 0) - log(1 + t)
      prod *= d1;                                        // 16 factors in (1, 2]: one log per row block
    }
    lb -= std::log(prod);
    lb += (RB - nr) * 0.6931471805599453;               // padding rows (eta = 0) contributed -log 2 each
    lsum += lb;
    for (int k = 0; k < nr; ++k) {
      const float* x = X + (i0 + k) * D;
      const double rk = r[k];
#pragma omp simd
      for (int d = 0; d < D; ++d) a[d] += rk * static_cast<double>(x[d]);
    }
  }
  double qq = 0.0;
  for (int d = 0; d < D; ++d) { g[d] = a[d] - tau * q[d]; qq += q[d] * q[d]; }
  return lsum - 0.5 * tau * qq;
}

}  // namespace

extern "C" {

int cpufast_max_threads() { return omp_get_max_threads(); }
void cpufast_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }

// value and gradient for C chains, one chain per thread; returns the number of threads that took part
int cpufast_logistic_grad(const float* X, int64_t N, int D, double tau, const double* q /*[C][D]*/, int C, double* g /*[C][D]*/,
                          double* l /*[C]*/) {
  int used = 0;
#pragma omp parallel
  {
#pragma omp single
    used = omp_get_num_threads();
#pragma omp for schedule(static)
    for (int c = 0; c < C; ++c) l[c] = grad_one(X, N, D, tau, q + static_cast<size_t>(c) * D, g + static_cast<size_t>(c) * D);
  }
  return used;
}

// nsteps leapfrogs (unit metric; ≙ src/kinetic_energy.jl:144-161) of every chain from (q, p), in place; g holds the
// gradient at q on entry and on exit.  Returns the number of threads that took part.
int cpufast_logistic_leapfrog(const float* X, int64_t N, int D, double tau, double eps, int nsteps, int C, double* q, double* p,
                              double* g, double* l) {
  int used = 0;
#pragma omp parallel
  {
#pragma omp single
    used = omp_get_num_threads();
#pragma omp for schedule(static)
    for (int c = 0; c < C; ++c) {
      double* qc = q + static_cast<size_t>(c) * D; double* pc = p + static_cast<size_t>(c) * D; double* gc = g + static_cast<size_t>(c) * D;
      for (int s = 0; s < nsteps; ++s) {
        for (int d = 0; d < D; ++d) { pc[d] = std::fma(0.5 * eps, gc[d], pc[d]); qc[d] = std::fma(eps, pc[d], qc[d]); }
        l[c] = grad_one(X, N, D, tau, qc, gc);
        for (int d = 0; d < D; ++d) pc[d] = std::fma(0.5 * eps, gc[d], pc[d]);
      }
    }
  }
  return used;
}

}  // extern "C"

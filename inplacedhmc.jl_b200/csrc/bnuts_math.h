// bnuts_math.h — scalar math shared by host and device builds.
//
// Every function here is built only from IEEE-754 correctly rounded operations
// (+, -, *, /, sqrt, fma) and integer bit manipulation, so a g++ build
// (-ffp-contract=off) and an nvcc build (-fmad=false) return bit-identical
// results.  That is what makes tree decisions comparable bit-for-bit between the
// CUDA engine and the CPU oracle.  No libm / libdevice transcendental is used.
//
// Covers what the reference leaves to un-pinned dependencies:
//   logaddexp            src/InplaceDHMC.jl:27-30   (libm exp + log1p)
//   randexp              src/NUTS.jl:33             (Random stdlib)
//   randn!               src/kinetic_energy.jl:63   (VectorizedRNG)
//   rand(UInt32)         src/tree.jl:144-145        (VectorizedRNG PCG)
// The RNG is replaced by counter-based Philox4x32-10 (no stream parity with
// PCG is possible or required).
#pragma once
#include <stdint.h>
#include <string.h>
#include <math.h>

#if defined(__CUDACC__)
#define BN_HD __host__ __device__ __forceinline__
// large scalar routines (transcendentals, Philox) are real functions on the device: inlined at every call site they
// made k_advance 460 KB of code and the kernel instruction-fetch bound (ncu: 45 % of warp samples "no instruction")
#ifdef BNUTS_INLINE_MATH
#define BN_HDN __host__ __device__ __forceinline__
#else
#define BN_HDN __host__ __device__ inline __noinline__
#endif
#else
#define BN_HD inline
#define BN_HDN inline
#endif

namespace bn {

// ---------------------------------------------------------------- bit casts
BN_HD uint64_t d2u(double x) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(x);
#else
  uint64_t u; memcpy(&u, &x, 8); return u;
#endif
}
BN_HD double u2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  double x; memcpy(&x, &u, 8); return x;
#endif
}
BN_HD uint32_t f2u(float x) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(x);
#else
  uint32_t u; memcpy(&u, &x, 4); return u;
#endif
}
BN_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float x; memcpy(&x, &u, 4); return x;
#endif
}

// explicit fused multiply-add (the only place contraction happens)
BN_HD double fma_(double a, double b, double c) { return ::fma(a, b, c); }
BN_HD float fma_(float a, float b, float c) { return ::fmaf(a, b, c); }
BN_HD double sqrt_(double a) { return ::sqrt(a); }
BN_HD float sqrt_(float a) { return ::sqrtf(a); }

BN_HD bool isfinite_(double x) { return ((d2u(x) >> 52) & 0x7ffu) != 0x7ffu; }
BN_HD bool isfinite_(float x) { return ((f2u(x) >> 23) & 0xffu) != 0xffu; }

template <class T> struct lim;
template <> struct lim<double> {
  static BN_HD double inf() { return u2d(0x7ff0000000000000ull); }
  static BN_HD double nan() { return u2d(0x7ff8000000000000ull); }
};
template <> struct lim<float> {
  static BN_HD float inf() { return u2f(0x7f800000u); }
  static BN_HD float nan() { return u2f(0x7fc00000u); }
};

// ---------------------------------------------------------------- exp
// k = round(x/ln2), r = x - k ln2 (two-term Cody-Waite), Taylor polynomial of
// e^r on |r| <= ln2/2, scaled by 2^k through the exponent field.
BN_HDN double exp_(double x) {
  if (x != x) return x;
  if (x > 709.0) return lim<double>::inf();
  if (x < -708.0) return 0.0;
  const double kf = ::floor(fma_(x, 1.4426950408889634074, 0.5));
  double r = fma_(-kf, 6.93147180369123816490e-01, x);
  r = fma_(-kf, 1.90821492927058770002e-10, r);
  // sum_{n=0}^{13} r^n / n!   (remainder < 4e-18)
  double p = 1.0 / 6227020800.0;
  p = fma_(p, r, 1.0 / 479001600.0);
  p = fma_(p, r, 1.0 / 39916800.0);
  p = fma_(p, r, 1.0 / 3628800.0);
  p = fma_(p, r, 1.0 / 362880.0);
  p = fma_(p, r, 1.0 / 40320.0);
  p = fma_(p, r, 1.0 / 5040.0);
  p = fma_(p, r, 1.0 / 720.0);
  p = fma_(p, r, 1.0 / 120.0);
  p = fma_(p, r, 1.0 / 24.0);
  p = fma_(p, r, 1.0 / 6.0);
  p = fma_(p, r, 0.5);
  p = fma_(p, r, 1.0);
  p = fma_(p, r, 1.0);
  const int64_t k = (int64_t)kf;  // |k| <= 1023 here
  return p * u2d((uint64_t)(k + 1023) << 52);
}
BN_HDN float exp_(float x) {
  if (x != x) return x;
  if (x > 88.0f) return lim<float>::inf();
  if (x < -87.0f) return 0.0f;
  const float kf = ::floorf(fma_(x, 1.44269504088896341f, 0.5f));
  float r = fma_(-kf, 0.693145751953125f, x);
  r = fma_(-kf, 1.42860676533018704e-06f, r);
  float p = 1.0f / 5040.0f;
  p = fma_(p, r, 1.0f / 720.0f);
  p = fma_(p, r, 1.0f / 120.0f);
  p = fma_(p, r, 1.0f / 24.0f);
  p = fma_(p, r, 1.0f / 6.0f);
  p = fma_(p, r, 0.5f);
  p = fma_(p, r, 1.0f);
  p = fma_(p, r, 1.0f);
  const int32_t k = (int32_t)kf;  // |k| <= 127 here
  return p * u2f((uint32_t)(k + 127) << 23);
}

// ---------------------------------------------------------------- log
// x = 2^e * m, m in [sqrt(1/2), sqrt(2)); log m = 2 atanh(s), s = (m-1)/(m+1),
// odd series in s with the exact rational coefficients 1/(2j+1).
BN_HDN double log_(double x) {
  if (x != x) return x;
  if (x < 0.0) return lim<double>::nan();
  if (x == 0.0) return -lim<double>::inf();
  if (!isfinite_(x)) return x;
  int64_t e = 0;
  uint64_t u = d2u(x);
  if ((u >> 52) == 0) {  // subnormal: scale by 2^54
    x = x * 18014398509481984.0;
    u = d2u(x);
    e = -54;
  }
  e += (int64_t)((u >> 52) & 0x7ffu) - 1023;
  double m = u2d((u & 0x000fffffffffffffull) | 0x3ff0000000000000ull);
  if (m > 1.4142135623730951) { m = m * 0.5; e += 1; }
  const double s = (m - 1.0) / (m + 1.0);
  const double z = s * s;
  double q = 1.0 / 23.0;
  q = fma_(q, z, 1.0 / 21.0);
  q = fma_(q, z, 1.0 / 19.0);
  q = fma_(q, z, 1.0 / 17.0);
  q = fma_(q, z, 1.0 / 15.0);
  q = fma_(q, z, 1.0 / 13.0);
  q = fma_(q, z, 1.0 / 11.0);
  q = fma_(q, z, 1.0 / 9.0);
  q = fma_(q, z, 1.0 / 7.0);
  q = fma_(q, z, 1.0 / 5.0);
  q = fma_(q, z, 1.0 / 3.0);
  const double two_s = s + s;
  const double lm = fma_(two_s * z, q, two_s);
  const double ef = (double)e;
  return fma_(ef, 6.93147180369123816490e-01, fma_(ef, 1.90821492927058770002e-10, lm));
}
BN_HDN float log_(float x) {
  if (x != x) return x;
  if (x < 0.0f) return lim<float>::nan();
  if (x == 0.0f) return -lim<float>::inf();
  if (!isfinite_(x)) return x;
  int32_t e = 0;
  uint32_t u = f2u(x);
  if ((u >> 23) == 0) {  // subnormal: scale by 2^25
    x = x * 33554432.0f;
    u = f2u(x);
    e = -25;
  }
  e += (int32_t)((u >> 23) & 0xffu) - 127;
  float m = u2f((u & 0x007fffffu) | 0x3f800000u);
  if (m > 1.41421356f) { m = m * 0.5f; e += 1; }
  const float s = (m - 1.0f) / (m + 1.0f);
  const float z = s * s;
  float q = 1.0f / 11.0f;
  q = fma_(q, z, 1.0f / 9.0f);
  q = fma_(q, z, 1.0f / 7.0f);
  q = fma_(q, z, 1.0f / 5.0f);
  q = fma_(q, z, 1.0f / 3.0f);
  const float two_s = s + s;
  const float lm = fma_(two_s * z, q, two_s);
  const float ef = (float)e;
  return fma_(ef, 0.693145751953125f, fma_(ef, 1.42860676533018704e-06f, lm));
}

// log(1+u) with the rounding error of 1+u compensated
template <class T> BN_HD T log1p_(T u) {
  const T w = T(1) + u;
  if (w == T(1)) return u;
  if (!isfinite_(w)) return log_(w);
  return log_(w) + (u - (w - T(1))) / w;
}

// logaddexp — restates src/InplaceDHMC.jl:27-30 (argument order matters for NaN)
template <class T> BN_HD T logaddexp_(T x, T y) {
  if (!(isfinite_(x) && isfinite_(y))) return x > y ? x : y;
  return x > y ? x + log1p_(exp_(y - x)) : y + log1p_(exp_(x - y));
}

// ---------------------------------------------------------------- sin/cos(2 pi u)
// u in [0,1).  j = nearest quarter turn, f = u - j/4 in [-1/8,1/8] (exact),
// Taylor series of sin/cos on |phi| <= pi/4, then a quadrant rotation.
BN_HDN void sincos2pi_(double u, double* s_out, double* c_out) {
  const double jf = ::floor(fma_(u, 4.0, 0.5));
  const double f = u - jf * 0.25;
  const double x = f * 6.283185307179586476925;
  const double z = x * x;
  // sin x = x (1 - z/3! + z^2/5! - ... + z^8/17!)
  double sp = 1.0 / 355687428096000.0;
  sp = fma_(sp, z, -1.0 / 1307674368000.0);
  sp = fma_(sp, z, 1.0 / 6227020800.0);
  sp = fma_(sp, z, -1.0 / 39916800.0);
  sp = fma_(sp, z, 1.0 / 362880.0);
  sp = fma_(sp, z, -1.0 / 5040.0);
  sp = fma_(sp, z, 1.0 / 120.0);
  sp = fma_(sp, z, -1.0 / 6.0);
  const double s = fma_(x * z, sp, x);
  // cos x = 1 - z/2! + z^2/4! - ... + z^9/18!
  double cp = -1.0 / 6402373705728000.0;
  cp = fma_(cp, z, 1.0 / 20922789888000.0);
  cp = fma_(cp, z, -1.0 / 87178291200.0);
  cp = fma_(cp, z, 1.0 / 479001600.0);
  cp = fma_(cp, z, -1.0 / 3628800.0);
  cp = fma_(cp, z, 1.0 / 40320.0);
  cp = fma_(cp, z, -1.0 / 720.0);
  cp = fma_(cp, z, 1.0 / 24.0);
  cp = fma_(cp, z, -0.5);
  const double c = fma_(cp, z, 1.0);
  const int j = ((int)jf) & 3;
  *s_out = (j == 0) ? s : (j == 1) ? c : (j == 2) ? -s : -c;
  *c_out = (j == 0) ? c : (j == 1) ? -s : (j == 2) ? -c : s;
}
BN_HDN void sincos2pi_(float u, float* s_out, float* c_out) {
  const float jf = ::floorf(fma_(u, 4.0f, 0.5f));
  const float f = u - jf * 0.25f;
  const float x = f * 6.28318530717958648f;
  const float z = x * x;
  float sp = 1.0f / 362880.0f;
  sp = fma_(sp, z, -1.0f / 5040.0f);
  sp = fma_(sp, z, 1.0f / 120.0f);
  sp = fma_(sp, z, -1.0f / 6.0f);
  const float s = fma_(x * z, sp, x);
  float cp = 1.0f / 40320.0f;
  cp = fma_(cp, z, -1.0f / 720.0f);
  cp = fma_(cp, z, 1.0f / 24.0f);
  cp = fma_(cp, z, -0.5f);
  const float c = fma_(cp, z, 1.0f);
  const int j = ((int)jf) & 3;
  *s_out = (j == 0) ? s : (j == 1) ? c : (j == 2) ? -s : -c;
  *c_out = (j == 0) ? c : (j == 1) ? -s : (j == 2) ? -c : s;
}

// ---------------------------------------------------------------- Philox4x32-10
struct u32x4 { uint32_t x, y, z, w; };

BN_HD void mulhilo32(uint32_t a, uint32_t b, uint32_t* hi, uint32_t* lo) {
  const uint64_t p = (uint64_t)a * (uint64_t)b;
  *hi = (uint32_t)(p >> 32);
  *lo = (uint32_t)p;
}
BN_HDN u32x4 philox4x32_10(u32x4 c, uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0, lo0, hi1, lo1;
    mulhilo32(0xD2511F53u, c.x, &hi0, &lo0);
    mulhilo32(0xCD9E8D57u, c.z, &hi1, &lo1);
    u32x4 n;
    n.x = hi1 ^ c.y ^ k0;
    n.y = lo1;
    n.z = hi0 ^ c.w ^ k1;
    n.w = lo0;
    c = n;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c;
}

// Counter layout: (global chain id, transition, purpose|tags, index)
enum : uint32_t { PURPOSE_DIRS = 0, PURPOSE_MOMENTUM = 1, PURPOSE_MERGE = 2, PURPOSE_INIT = 3 };

BN_HD u32x4 draw4(uint64_t seed, uint32_t chain, uint32_t t, uint32_t purpose_tag, uint32_t index) {
  u32x4 c;
  c.x = chain; c.y = t; c.z = purpose_tag; c.w = index;
  return philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
}

// uniforms: open0 in (0,1], half-open in [0,1)
BN_HD double u01_open0(uint32_t hi, uint32_t lo, double) {
  const uint64_t k = ((uint64_t)hi << 21) | ((uint64_t)lo >> 11);
  return (double)(k + 1) * 1.1102230246251565404e-16;  // 2^-53
}
BN_HD double u01_half(uint32_t hi, uint32_t lo, double) {
  const uint64_t k = ((uint64_t)hi << 21) | ((uint64_t)lo >> 11);
  return (double)k * 1.1102230246251565404e-16;
}
BN_HD float u01_open0(uint32_t w, float) { return (float)((w >> 8) + 1u) * 5.9604644775390625e-08f; }
BN_HD float u01_half(uint32_t w, float) { return (float)(w >> 8) * 5.9604644775390625e-08f; }

// Box-Muller pair
template <class T> BN_HD void box_muller(T u1_open0, T u2_half, T* z0, T* z1) {
  const T r = sqrt_(T(-2) * log_(u1_open0));
  T s, c;
  sincos2pi_(u2_half, &s, &c);
  *z0 = r * c;
  *z1 = r * s;
}

// Standard normal for coordinate d of (chain, transition t).  fp64: one Philox
// call gives coordinates (2k, 2k+1); fp32: one call gives (4k .. 4k+3).
BN_HD double std_normal(uint64_t seed, uint32_t chain, uint32_t t, uint32_t d, double tag) {
  const u32x4 w = draw4(seed, chain, t, PURPOSE_MOMENTUM, d >> 1);
  double z0, z1;
  box_muller(u01_open0(w.x, w.y, tag), u01_half(w.z, w.w, tag), &z0, &z1);
  return (d & 1u) ? z1 : z0;
}
BN_HD float std_normal(uint64_t seed, uint32_t chain, uint32_t t, uint32_t d, float tag) {
  const u32x4 w = draw4(seed, chain, t, PURPOSE_MOMENTUM, d >> 2);
  float z0, z1;
  if (d & 2u) box_muller(u01_open0(w.z, tag), u01_half(w.w, tag), &z0, &z1);
  else        box_muller(u01_open0(w.x, tag), u01_half(w.y, tag), &z0, &z1);
  return (d & 1u) ? z1 : z0;
}

// Exponential(1) draw for the merge identified by (doubling j, level k, leaf n)
// of transition t; k = 0, n = 0 is the top-level merge of doubling j.  A pure
// function of the position in the tree, consumed only if logprob2 < 0
// (same laziness as src/NUTS.jl:33).
BN_HD uint32_t merge_tag(uint32_t j, uint32_t k) { return PURPOSE_MERGE | (j << 8) | (k << 16); }
BN_HD double std_exponential(uint64_t seed, uint32_t chain, uint32_t t, uint32_t j, uint32_t k,
                             uint32_t n, double tag) {
  const u32x4 w = draw4(seed, chain, t, merge_tag(j, k), n);
  return -log_(u01_open0(w.x, w.y, tag));
}
BN_HD float std_exponential(uint64_t seed, uint32_t chain, uint32_t t, uint32_t j, uint32_t k,
                            uint32_t n, float tag) {
  const u32x4 w = draw4(seed, chain, t, merge_tag(j, k), n);
  return -log_(u01_open0(w.x, tag));
}
BN_HD uint32_t draw_directions(uint64_t seed, uint32_t chain, uint32_t t) {
  return draw4(seed, chain, t, PURPOSE_DIRS, 0).x;
}
// initial position coordinate ~ U[-2,2]  (src/warmup.jl:73)
// ≙ random_position! on a restart of FindLocalOptimum (src/warmup.jl:169): attempt >= 1
BN_HD double restart_position(uint64_t seed, uint32_t chain, uint32_t attempt, uint32_t d) {
  const u32x4 w = draw4(seed, chain, 0xffffffffu - attempt, PURPOSE_INIT, d);
  return fma_(u01_half(w.x, w.y, 0.0), 4.0, -2.0);
}
BN_HD double init_position(uint64_t seed, uint32_t chain, uint32_t d) {
  const u32x4 w = draw4(seed, chain, 0xffffffffu, PURPOSE_INIT, d);
  return fma_(u01_half(w.x, w.y, 0.0), 4.0, -2.0);
}


// ---- synthetic logistic-regression data (config c5 generates its rows from Philox keyed by
// (data seed, GLOBAL row index), so every sharding of the rows sees the same matrix and nothing crosses PCIe).
// Pure functions of (seed, row, column), the same bits on host and device:
//   x[row][0] = 1 (intercept); x[row][d] = bf16(z), z ~ N(0,1) from fp32 Box-Muller (one Philox call = 4 columns)
//   beta*[d] = z / sqrt(D), z ~ N(0,1) in Float64
//   y[row] = u < sigma(eta), eta = sum_d x[row][d] beta*[d] accumulated in Float64 with fma, d = 0, 1, ..., D-1
enum : uint32_t { PURPOSE_DATA = 4 };
BN_HD void synth_x4(uint64_t seed, uint64_t row, uint32_t quad, float* x) {
  const u32x4 w = draw4(seed, (uint32_t)row, (uint32_t)(row >> 32), PURPOSE_DATA, quad);
  box_muller(u01_open0(w.x, 0.f), u01_half(w.y, 0.f), &x[0], &x[1]);
  box_muller(u01_open0(w.z, 0.f), u01_half(w.w, 0.f), &x[2], &x[3]);
  for (int e = 0; e < 4; ++e) {   // round to nearest even on the bf16 grid
    const uint32_t u = f2u(x[e]);
    x[e] = u2f((u + 0x7fffu + ((u >> 16) & 1u)) & 0xffff0000u);
  }
  if (quad == 0) x[0] = 1.f;
}
BN_HD double synth_beta(uint64_t seed, uint32_t d, int32_t D) {
  const u32x4 w = draw4(seed, d, 0u, PURPOSE_DATA | (1u << 8), 0u);
  double z0, z1;
  box_muller(u01_open0(w.x, w.y, 0.0), u01_half(w.z, w.w, 0.0), &z0, &z1);
  return z0 / sqrt_((double)D);
}
BN_HD double synth_uniform(uint64_t seed, uint64_t row) {
  const u32x4 w = draw4(seed, (uint32_t)row, (uint32_t)(row >> 32), PURPOSE_DATA | (2u << 8), 0u);
  return u01_half(w.x, w.y, 0.0);
}
// label of one row given beta* (length D): the reference implementation of the definition above
BN_HDN double synth_label(uint64_t seed, uint64_t row, int32_t D, const double* beta) {
  double eta = 0.0;
  for (int32_t q = 0; 4 * q < D; ++q) {
    float x[4];
    synth_x4(seed, row, (uint32_t)q, x);
    for (int e = 0; e < 4 && 4 * q + e < D; ++e) eta = fma_((double)x[e], beta[4 * q + e], eta);
  }
  const double p = 1.0 / (1.0 + exp_(-eta));
  return synth_uniform(seed, row) < p ? 1.0 : 0.0;
}

}  // namespace bn

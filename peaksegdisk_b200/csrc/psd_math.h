// psd_math.h -- deterministic fp64 exp/log shared by the CUDA kernels and the host driver.
//
// Why this exists: the reference solver (src/funPieceListLog.cpp:192-234, :29-190) calls libm
// exp()/log() ~200 times per bedGraph row, and its integer outputs (segment ends, peak counts,
// equality constraints, interval counts) and the 15-digit penalty chain of sequentialSearch_dir
// depend on the last bit of those results.  CUDA's exp/log are different
// 1-ulp functions, so the device would drift from the reference.  These two functions perform
// exactly the operation sequence of glibc 2.39's x86-64 FMA variants (__ieee754_exp_fma /
// __ieee754_log_fma, the ones ifunc selects on every AVX2+FMA host), including which
// multiply-adds are fused, so host, device and the reference binary agree bit for bit.
// tests/test_math.py checks this against the system libm on hundreds of millions of inputs.
//
// Build rules: compile with -fmad=false (nvcc) / -ffp-contract=off (gcc); every fused
// multiply-add below is written explicitly as PSD_FMA.
#pragma once
#include <stdint.h>
#include "psd_math_tables.h"

#if defined(__CUDACC__)
#define PSD_HD __host__ __device__ __forceinline__
#else
#define PSD_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define PSD_FMA(a, b, c) __fma_rn((a), (b), (c))
#define PSD_D2U(x) ((uint64_t)__double_as_longlong(x))
#define PSD_U2D(u) __longlong_as_double((long long)(u))
#else
#define PSD_FMA(a, b, c) __builtin_fma((a), (b), (c))
static inline uint64_t psd_d2u_(double x) { uint64_t u; __builtin_memcpy(&u, &x, 8); return u; }
static inline double psd_u2d_(uint64_t u) { double x; __builtin_memcpy(&x, &u, 8); return x; }
#define PSD_D2U(x) psd_d2u_(x)
#define PSD_U2D(u) psd_u2d_(u)
#endif

#define PSD_INF PSD_U2D(0x7ff0000000000000ULL)
#define PSD_NAN PSD_U2D(0x7ff8000000000000ULL)

// Host-side copies of the tables (device code receives pointers to shared-memory copies).
#if !defined(__CUDA_ARCH__)
static const uint64_t psd_exp_tab_host[256] = PSD_EXP_TAB_INIT;
static const uint64_t psd_log_tab_host[256] = PSD_LOG_TAB_INIT;
#endif

// exp(x); T = 256-entry table {tail bits, scale bits} for 2^(i/128).
PSD_HD double psd_exp(double x, const uint64_t* __restrict__ T) {
  const uint64_t ix = PSD_D2U(x);
  uint32_t abstop = (uint32_t)(ix >> 52) & 0x7ffu;
  if (abstop - 0x3c9u >= 0x3fu) {
    if ((int32_t)(abstop - 0x3c9u) < 0) return 1.0 + x;  // |x| < 2^-54
    if (abstop >= 0x409u) {                              // |x| >= 1024
      if (ix == 0xfff0000000000000ULL) return 0.0;
      if (abstop == 0x7ffu) return 1.0 + x;              // +inf, nan
      return (ix >> 63) ? 0.0 : PSD_INF;                 // underflow / overflow
    }
    abstop = 0;                                          // 512 <= |x| < 1024: careful scaling below
  }
  const double shift = PSD_U2D(PSD_EXP_SHIFT);
  double kd = PSD_FMA(x, PSD_U2D(PSD_EXP_INVLN2N), shift);
  const uint64_t ki = PSD_D2U(kd);
  kd = kd - shift;
  double r = PSD_FMA(kd, PSD_U2D(PSD_EXP_NEGLN2HIN), x);
  r = PSD_FMA(kd, PSD_U2D(PSD_EXP_NEGLN2LON), r);
  const uint32_t idx = 2u * ((uint32_t)ki & 127u);
  const uint64_t top = ki << 45;
  const double tail = PSD_U2D(T[idx]);
  uint64_t sbits = T[idx + 1] + top;
  const double r2 = r * r;
  const double p23 = PSD_FMA(PSD_U2D(PSD_EXP_C3), r, PSD_U2D(PSD_EXP_C2));
  const double tr = r + tail;
  const double p45 = PSD_FMA(r, PSD_U2D(PSD_EXP_C5), PSD_U2D(PSD_EXP_C4));
  const double t1 = PSD_FMA(p23, r2, tr);
  const double r4 = r2 * r2;
  const double tmp = PSD_FMA(r4, p45, t1);
  if (abstop == 0) {
    if ((ki & 0x80000000ULL) == 0) {  // k > 0: scale may have overflowed
      sbits -= 1009ULL << 52;
      const double scale = PSD_U2D(sbits);
      return PSD_FMA(scale, tmp, scale) * PSD_U2D(0x7f00000000000000ULL);  // * 2^1009
    }
    sbits += 1022ULL << 52;           // k < 0: result may be subnormal
    const double scale = PSD_U2D(sbits);
    const double p = scale * tmp;
    double y = scale + p;
    if (y < 1.0) {
      double lo = (scale - y) + p;
      const double hi = 1.0 + y;
      lo = ((1.0 - hi) + y) + lo;
      y = (hi + lo) - 1.0;
      if (y == 0.0) y = 0.0;
    }
    return y * PSD_U2D(0x0010000000000000ULL);  // * 2^-1022
  }
  const double scale = PSD_U2D(sbits);
  return PSD_FMA(scale, tmp, scale);
}

// log(x); T = 256-entry table {invc bits, logc bits} for the 128 sub-intervals of [0.6875, 1.375).
PSD_HD double psd_log(double x, const uint64_t* __restrict__ T) {
  uint64_t ix = PSD_D2U(x);
  const uint32_t top = (uint32_t)(ix >> 48);
  if (ix - 0x3fee000000000000ULL < 0x3090000000000ULL) {  // 1-2^-4 <= x < 1+0x1.09p-4
    if (ix == 0x3ff0000000000000ULL) return 0.0;
    const double r = x - 1.0;
    const double B0 = PSD_U2D(PSD_LOG_B0);
    const double q12 = PSD_FMA(PSD_U2D(PSD_LOG_B2), r, PSD_U2D(PSD_LOG_B1));
    const double q45 = PSD_FMA(PSD_U2D(PSD_LOG_B5), r, PSD_U2D(PSD_LOG_B4));
    const double r2 = r * r;
    const double q78 = PSD_FMA(PSD_U2D(PSD_LOG_B8), r, PSD_U2D(PSD_LOG_B7));
    const double q123 = PSD_FMA(r2, PSD_U2D(PSD_LOG_B3), q12);
    const double q456 = PSD_FMA(r2, PSD_U2D(PSD_LOG_B6), q45);
    const double r3 = r * r2;
    double q = PSD_FMA(r2, PSD_U2D(PSD_LOG_B9), q78);
    q = PSD_FMA(r3, PSD_U2D(PSD_LOG_B10), q);
    q = PSD_FMA(q, r3, q456);
    q = PSD_FMA(q, r3, q123);
    const double two27 = PSD_U2D(0x41a0000000000000ULL);
    const double rw = PSD_FMA(r, two27, r);        // r + r*2^27
    const double rhi = PSD_FMA(-two27, r, rw);     // rw - r*2^27
    const double rhi2 = rhi * rhi;
    const double rlo = r - rhi;
    const double hi = PSD_FMA(rhi2, B0, r);
    const double rmh = r - hi;
    const double rs = r + rhi;
    double lo = PSD_FMA(rhi2, B0, rmh);
    const double brlo = B0 * rlo;
    lo = PSD_FMA(brlo, rs, lo);
    const double y = PSD_FMA(q, r3, lo);
    return hi + y;
  }
  if (top - 0x0010u > 0x7fdfu) {  // x < 2^-1022, inf or nan
    if (ix * 2 == 0) return -PSD_INF;
    if (ix == 0x7ff0000000000000ULL) return x;
    if ((top & 0x8000u) || (top & 0x7ff0u) == 0x7ff0u) {
      // glibc: __math_invalid(x) = (x - x) / (x - x).  On x86-64 that is the "real indefinite" NaN
      // (sign bit set, printed as -nan by the reference's iostream) for negative arguments and the
      // quieted argument for a NaN argument.
      if ((ix << 1) > 0xffe0000000000000ULL) return PSD_U2D(ix | 0x0008000000000000ULL);
      return PSD_U2D(0xfff8000000000000ULL);
    }
    ix = PSD_D2U(x * PSD_U2D(0x4330000000000000ULL));  // subnormal: scale by 2^52
    ix -= 52ULL << 52;
  }
  const uint64_t tmp = ix - 0x3fe6000000000000ULL;
  const uint32_t i = (uint32_t)(tmp >> 45) & 127u;
  const int32_t k = (int32_t)((int64_t)tmp >> 52);
  const uint64_t iz = ix - (tmp & 0xfff0000000000000ULL);
  const double invc = PSD_U2D(T[2 * i]);
  const double logc = PSD_U2D(T[2 * i + 1]);
  const double z = PSD_U2D(iz);
  const double kd = (double)k;
  const double w = PSD_FMA(kd, PSD_U2D(PSD_LOG_LN2HI), logc);
  const double r = PSD_FMA(z, invc, -1.0);
  const double p12 = PSD_FMA(PSD_U2D(PSD_LOG_A2), r, PSD_U2D(PSD_LOG_A1));
  const double hi = r + w;
  const double r2 = r * r;
  double lo = (w - hi) + r;
  lo = PSD_FMA(kd, PSD_U2D(PSD_LOG_LN2LO), lo);
  const double r3 = r * r2;
  const double p34 = PSD_FMA(r, PSD_U2D(PSD_LOG_A4), PSD_U2D(PSD_LOG_A3));
  lo = PSD_FMA(r2, PSD_U2D(PSD_LOG_A0), lo);
  const double p = PSD_FMA(p34, r2, p12);
  const double y = PSD_FMA(r3, p, lo);
  return y + hi;
}

// deva_log.cuh - bit-exact restatement of the `log` the reference's PHOLD models call.
//
// The reference computes its event delay as (uint64_t)(-lambda*std::log(1.0 - u))
// (test/phold.cxx:114, test/phold-bcast.cxx:70, bench/phold.cxx:51,110).  std::log is not in
// the reference tree: it is glibc 2.39 libm, sysdeps/ieee754/dbl-64/e_log.c (table-driven,
// N = 128, from ARM optimized-routines), and on every FMA+AVX2 x86-64 host the ifunc picks the
// `__log_fma` build.  Committed results must be bit-exact, so this header restates THAT build
// operation by operation: every add/mul/fma below appears in the same order and with the same
// fusion as in the host's `__log_fma` (checked against its disassembly), written with explicit
// round-to-nearest intrinsics on the device so that nvcc can neither fuse nor split them.
// The table is glibc's `__log_data`, extracted by tools/extract_glibc_log_table.py.
//
// Compiles for the device (nvcc) and for the host (g++ -ffp-contract=off; used by the CPU tests
// that sweep it against the host's std::log, tests/test_log_parity.py).
#pragma once
#include <stdint.h>
#include "glibc_log_data.inc"

#if defined(__CUDACC__)
  #define DEVA_HD __host__ __device__ __forceinline__
  // out-of-line on the device: keeps rarely executed code out of the hot instruction stream
  #define DEVA_HD_NOINLINE __host__ __device__ __noinline__
#else
  #define DEVA_HD inline
  #define DEVA_HD_NOINLINE inline
#endif

#if defined(__CUDA_ARCH__)
  #define DEVA_FMA(a, b, c) __fma_rn((a), (b), (c))
  #define DEVA_ADD(a, b) __dadd_rn((a), (b))
  #define DEVA_SUB(a, b) __dsub_rn((a), (b))
  #define DEVA_MUL(a, b) __dmul_rn((a), (b))
  #define DEVA_DIV(a, b) __ddiv_rn((a), (b))
  #define DEVA_SQRT(a) __dsqrt_rn((a))
  #define DEVA_AS_U64(d) ((uint64_t)__double_as_longlong(d))
  #define DEVA_AS_F64(u) __longlong_as_double((long long)(u))
#else
  #include <math.h>
  #include <string.h>
  #define DEVA_FMA(a, b, c) fma((a), (b), (c))
  #define DEVA_ADD(a, b) ((a) + (b))
  #define DEVA_SUB(a, b) ((a) - (b))
  #define DEVA_MUL(a, b) ((a) * (b))
  #define DEVA_DIV(a, b) ((a) / (b))
  #define DEVA_SQRT(a) sqrt((a))
  static inline uint64_t deva_as_u64_(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
  static inline double deva_as_f64_(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
  #define DEVA_AS_U64(d) deva_as_u64_(d)
  #define DEVA_AS_F64(u) deva_as_f64_(u)
#endif

namespace deva_b200 {

#if defined(__CUDACC__)
// 128 x {invc, logc}.  Indexed by data-dependent i, so plain global memory (L1/L2-resident,
// 2 KB) rather than __constant__, whose divergent reads serialise.
__device__ const uint64_t d_log_tab[256] = DEVA_LOG_TAB;
#endif
static const uint64_t h_log_tab[256] = DEVA_LOG_TAB;

DEVA_HD double log_tab_(int idx) {
#if defined(__CUDA_ARCH__)
  return DEVA_AS_F64(__ldg(&d_log_tab[idx]));
#else
  return DEVA_AS_F64(h_log_tab[idx]);
#endif
}

// glibc 2.39 __log_fma(x) for x in [0, +inf]; x < 0 / NaN are not reachable from the models
// (arguments are 1-u with u in [0,1], or r2 in (0,1]) and return NaN like the host would.
DEVA_HD double glibc_log(double x) {
  const uint64_t A[5] = DEVA_LOG_POLY_A;
  const uint64_t B[11] = DEVA_LOG_POLY_B;
  uint64_t ix = DEVA_AS_U64(x);
  const uint64_t LO = 0x3fee000000000000ull;  // asuint64(1.0 - 0x1p-4)
  const uint64_t HI = 0x3ff1090000000000ull;  // asuint64(1.0 + 0x1.09p-4)
  if (ix - LO < HI - LO) {
    // e_log.c "close to 1.0" branch
    if (ix == 0x3ff0000000000000ull) return 0.0;
    const double r = DEVA_SUB(x, 1.0);
    const double q1a = DEVA_FMA(r, DEVA_AS_F64(B[2]), DEVA_AS_F64(B[1]));
    const double q2a = DEVA_FMA(r, DEVA_AS_F64(B[5]), DEVA_AS_F64(B[4]));
    const double r2 = DEVA_MUL(r, r);
    const double q3a = DEVA_FMA(r, DEVA_AS_F64(B[8]), DEVA_AS_F64(B[7]));
    const double q1 = DEVA_FMA(r2, DEVA_AS_F64(B[3]), q1a);
    const double q2 = DEVA_FMA(r2, DEVA_AS_F64(B[6]), q2a);
    const double r3 = DEVA_MUL(r, r2);
    double q3 = DEVA_FMA(r2, DEVA_AS_F64(B[9]), q3a);
    q3 = DEVA_FMA(r3, DEVA_AS_F64(B[10]), q3);
    double q = DEVA_FMA(q3, r3, q2);
    q = DEVA_FMA(q, r3, q1);
    const double two27 = 134217728.0;  // 0x1p27
    const double t = DEVA_FMA(r, two27, r);        // r + w,  w = r*0x1p27 (fused by gcc)
    const double rhi = DEVA_FMA(-two27, r, t);     // (r + w) - w
    const double b0 = DEVA_AS_F64(B[0]);           // -0.5
    const double rhi2 = DEVA_MUL(rhi, rhi);
    const double rlo = DEVA_SUB(r, rhi);
    const double hi = DEVA_FMA(rhi2, b0, r);       // r + rhi*rhi*B[0]
    const double rmh = DEVA_SUB(r, hi);
    const double rsum = DEVA_ADD(r, rhi);
    double lo = DEVA_FMA(rhi2, b0, rmh);           // r - hi + w
    const double b0rlo = DEVA_MUL(b0, rlo);
    lo = DEVA_FMA(b0rlo, rsum, lo);                // lo += B[0]*rlo*(rhi + r)
    const double y = DEVA_FMA(q, r3, lo);          // y = r3*(...) + lo
    return DEVA_ADD(hi, y);
  }
  uint32_t top = (uint32_t)(ix >> 48);
  if (top - 0x0010u >= 0x7ff0u - 0x0010u) {
    // x < 0x1p-1022, inf or nan
    if ((ix << 1) == 0) return DEVA_AS_F64(0xfff0000000000000ull);  // log(+-0) = -inf
    if (ix == 0x7ff0000000000000ull) return x;                      // log(inf) = inf
    if ((top & 0x8000u) || (top & 0x7ff0u) == 0x7ff0u) return DEVA_AS_F64(0x7ff8000000000000ull);
    // subnormal: normalise
    ix = DEVA_AS_U64(DEVA_MUL(x, 4503599627370496.0));  // x * 0x1p52
    ix -= 52ull << 52;
  }
  const uint64_t OFF = 0x3fe6000000000000ull;
  const uint64_t tmp = ix - OFF;
  const int i = (int)((tmp >> (52 - 7)) & 127);
  const int k = (int)((int64_t)tmp >> 52);
  const uint64_t iz = ix - (tmp & (0xfffull << 52));
  const double invc = log_tab_(2 * i);
  const double logc = log_tab_(2 * i + 1);
  const double z = DEVA_AS_F64(iz);
  const double kd = (double)k;
  const double w = DEVA_FMA(kd, DEVA_AS_F64(DEVA_LOG_LN2HI), logc);   // kd*Ln2hi + logc
  const double r = DEVA_FMA(z, invc, -1.0);
  const double p1 = DEVA_FMA(r, DEVA_AS_F64(A[2]), DEVA_AS_F64(A[1]));
  const double hi = DEVA_ADD(r, w);
  const double r2 = DEVA_MUL(r, r);
  double lo = DEVA_SUB(w, hi);
  lo = DEVA_ADD(lo, r);
  lo = DEVA_FMA(kd, DEVA_AS_F64(DEVA_LOG_LN2LO), lo);                 // w - hi + r + kd*Ln2lo
  const double r3 = DEVA_MUL(r, r2);
  const double p2 = DEVA_FMA(r, DEVA_AS_F64(A[4]), DEVA_AS_F64(A[3]));
  const double s = DEVA_FMA(r2, DEVA_AS_F64(A[0]), lo);               // lo + r2*A[0]
  const double p = DEVA_FMA(p2, r2, p1);
  const double y = DEVA_FMA(r3, p, s);
  return DEVA_ADD(y, hi);
}

// (uint64_t)d as g++ compiles it for x86-64 (cvttsd2si, with the >= 2^63 fix-up); +inf -> 0,
// which is what the reference's binary produces for log(0) (probability 2^-54 per draw).
DEVA_HD uint64_t x86_double_to_u64(double d) {
  const double two63 = 9223372036854775808.0;
  if (d >= two63) {
    const double e = DEVA_SUB(d, two63);
    if (!(e < two63)) return 0ull;  // cvttsd2si "indefinite" 0x8000... ^ 0x8000...
#if defined(__CUDA_ARCH__)
    return ((uint64_t)__double2ll_rz(e)) ^ 0x8000000000000000ull;
#else
    return ((uint64_t)(int64_t)e) ^ 0x8000000000000000ull;
#endif
  }
  if (!(d == d)) return 0x8000000000000000ull;
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double2ll_rz(d);
#else
  return (uint64_t)(int64_t)d;
#endif
}

// (double)u64 with round-to-nearest-even, as x86-64 g++ produces it.
DEVA_HD double u64_to_double(uint64_t u) {
#if defined(__CUDA_ARCH__)
  return __ull2double_rn(u);
#else
  return (double)u;
#endif
}

// test/phold.cxx:114-116 / bench/phold.cxx:110-112:  dt = (uint64_t)(-lambda*log(1 - r/2^64)) + 1
DEVA_HD uint64_t phold_dt(uint64_t r, double neg_lambda) {
  const double u = DEVA_MUL(u64_to_double(r), DEVA_AS_F64(0x3bf0000000000000ull));  // * 2^-64 == / double(-1ull), exact
  const double l = glibc_log(DEVA_SUB(1.0, u));
  return x86_double_to_u64(DEVA_MUL(neg_lambda, l)) + 1;
}

}  // namespace deva_b200

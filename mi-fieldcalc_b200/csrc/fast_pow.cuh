// fast_pow.cuh -- x^y for float x and a call-constant exponent, evaluated in double precision.
//
// Why: the Exner function powf(p/1000, r/cp) is evaluated once PER GRID POINT by every hybrid- and
// atmospheric-level operator (reference FC.cc:308-311) and powf(ff, 0.16f) by windCooling
// (FC.cc:2213).  libdevice's powf costs ~90 issue slots per point, which makes those kernels
// issue-bound at 40-65 % of the HBM roofline (profiles/r01_ncu_full_a_*).  This routine needs
// ~20 FP64-pipe instructions (a separate pipe on sm_100) plus a dozen integer ones, and it is MORE
// accurate: the double result has a relative error < 2^-38, so after rounding to float it is the
// correctly rounded x^y except in ~3 cases per million (tools/gen_pow_tables.py checks this on the
// CPU).  glibc's powf -- what the reference calls -- is itself correctly rounded in 99.94 % of cases,
// so the two agree bit-for-bit almost everywhere and never differ by more than one float ulp.
//
//   x = m * 2^e, m in [1, 2);  i = top 5 mantissa bits;  c_i = 1 + (i + 1/2)/32;  r = m/c_i - 1, |r| <= 2^-6
//   log2 x = e + log2 c_i + r*(A0 + r*(A1 + r*(A2 + r*(A3 + r*A4))))
//   t = y*log2 x = q + j/32 + f, |f| <= 1/64;   x^y = 2^q * 2^(j/32) * (1 + f*(C0 + f*(C1 + f*(C2 + f*C3))))
//
// FMA is used on purpose here: this is our own algorithm, not a restatement of a reference
// expression (the library is otherwise compiled with -fmad=false).
#pragma once

#include <cuda_runtime.h>

#include "pow_tables.inc"

namespace fcb200 {
namespace dev {

static __device__ __constant__ double2 c_pow_log[32] = FCB_POW_LOG_TABLE;
static __device__ __constant__ double c_pow_exp[32] = FCB_POW_EXP_TABLE;

// Tables live in shared memory while a kernel runs: the index is data dependent.
struct PowTable
{
  double2 lg[32]; // {1/c_i, log2 c_i}
  double ex[32];  // 2^(i/32)

  __device__ __forceinline__ void load()
  {
    for (int i = threadIdx.x; i < 32; i += blockDim.x) {
      lg[i] = c_pow_log[i];
      ex[i] = c_pow_exp[i];
    }
  }

  // x^y for finite x > 0 (normal or subnormal)
  __device__ __forceinline__ float pow_pos(float x, double y) const
  {
    constexpr double A[5] = FCB_POW_LOG_COEF;
    constexpr double C[4] = FCB_POW_EXP_COEF;
    int e0 = -127;
    if (x < 1.17549435e-38f) { // subnormal
      x *= 8388608.f;
      e0 -= 23;
    }
    const unsigned ix = __float_as_uint(x);
    const int e = (int)(ix >> 23) + e0;
    const int i = (ix >> 18) & 31;
    // m = 1.mantissa and e as doubles, assembled with integer instructions: the F2F / I2F conversions
    // run on the quarter-rate XU pipe, which these kernels would otherwise saturate
    const unsigned mant = ix & 0x007fffffu;
    const double m = __hiloint2double((int)(0x3ff00000u | (mant >> 3)), (int)(mant << 29));
    const double ed = __hiloint2double(0x43300000, (int)(0x80000000u ^ (unsigned)e)) - 4503601774854144.0; // (2^52 + 2^31 + e) - (2^52 + 2^31)
    const double2 t = lg[i];
    const double r = fma(m, t.x, -1.0);
    double p = fma(A[4], r, A[3]);
    p = fma(p, r, A[2]);
    p = fma(p, r, A[1]);
    p = fma(p, r, A[0]);
    const double lgx = fma(r, p, ed + t.y);
    const double ty = y * lgx;
    // split ty = k/32 + f with the round-to-nearest shift trick
    constexpr double SHIFT = 6755399441055744.0; // 1.5 * 2^52
    double kd = fma(ty, 32.0, SHIFT);
    const int k = __double2loint(kd);
    kd -= SHIFT;
    const double f = fma(kd, -1.0 / 32.0, ty);
    double s = fma(C[3], f, C[2]);
    s = fma(s, f, C[1]);
    s = fma(s, f, C[0]);
    const double tj = ex[k & 31];
    double res = fma(tj * f, s, tj);
    // multiply by 2^(k >> 5): |k >> 5| < 200, res in [1, 2) -> stays a normal double
    const int hi = __double2hiint(res) + ((k >> 5) << 20);
    res = __hiloint2double(hi, __double2loint(res));
    return (float)res;
  }

  // powf(x, y) semantics for y > 0 non-integer: NaN for finite x < 0 or NaN, +0 for +-0, +inf for +-inf
  __device__ __forceinline__ float pow(float x, double y) const
  {
    if (x > 0.f && x < __int_as_float(0x7f800000))
      return pow_pos(x, y);
    if (x == 0.f)
      return 0.f;
    if (x > 0.f || x == __int_as_float(0xff800000))
      return __int_as_float(0x7f800000); // +-inf -> +inf (y is not an odd integer)
    return __int_as_float(0x7fc00000); // x < 0 or NaN
  }
};

} // namespace dev
} // namespace fcb200

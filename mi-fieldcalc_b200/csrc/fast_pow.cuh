// fast_pow.cuh -- x^y for float x and a call-constant exponent, evaluated in double precision.
//
// Why: the Exner function powf(p/1000, r/cp) is evaluated once PER GRID POINT by every hybrid- and
// atmospheric-level operator (reference FC.cc:308-311) and powf(ff, 0.16f) by windCooling
// (FC.cc:2213).  libdevice's powf costs ~90 issue slots per point, which makes those kernels
// issue-bound at 40-65 % of the HBM roofline (profiles/r01_ncu_full_a_*).  This routine needs
// 15 FP64-pipe instructions (a separate pipe on sm_100) plus a dozen integer ones, and it is MORE
// accurate: the double result has a relative error < 2^-38, so after rounding to float it is the
// correctly rounded x^y except in ~3 cases per million (tools/gen_pow_tables.py checks this on the
// CPU).  glibc's powf -- what the reference calls -- is itself correctly rounded in 99.94 % of cases,
// so the two agree bit-for-bit almost everywhere and never differ by more than one float ulp.
//
//   x = m * 2^e, m in [1, 2);  i = top 5 mantissa bits;  c_i = 1 + (i + 1/2)/32;  r = m/c_i - 1, |r| <= 2^-6
//   log2 x = e + log2 c_i + r*(A0 + r*(A1 + r*(A2 + r*(A3 + r*A4))))
//   T = 32*y*log2 x = 32*q + j + g, |g| <= 1/2;   x^y = 2^q * 2^(j/32) * (1 + g*(D0 + g*(D1 + g*(D2 + g*D3))))
//   with D_k = C_k / 32^(k+1) (exact scalings of the Taylor coefficients C_k of 2^f - 1, f = g/32).
//
// Issue-slot economy (ncu source view of the round-1 chain kernel, profiles/r01_ncu_full_chain_summary.csv:
// 78 of 260 issue slots per point went to this routine):
//   * every polynomial coefficient is an operand from the CONSTANT BANK (c_pow_coef): a double literal
//     costs two UMOV per use, because only one DFMA operand may be an immediate and a 64-bit immediate
//     with a non-zero low word is not encodable at all;
//   * one unsigned comparison classifies x as a positive NORMAL float; everything else (zero,
//     subnormal, negative, inf, NaN) goes to a non-inlined slow path;
//   * the 32 of the index split is folded into the exponent constant and the coefficients.
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

namespace powdetail {
constexpr double A_[5] = FCB_POW_LOG_COEF;
constexpr double C_[4] = FCB_POW_EXP_COEF;
} // namespace powdetail

// exponents this library raises to (indices into c_pow_coef's tail)
enum PowExponent { POW_KAPPA = 0, POW_WINDCHILL = 1 };

// [0..4] A0..A4, [5..8] D0..D3, [9] 2^52 + 2^31 (exponent-to-double bias), [10..11] 32*y
static __device__ __constant__ double c_pow_coef[12] = {
    powdetail::A_[0],
    powdetail::A_[1],
    powdetail::A_[2],
    powdetail::A_[3],
    powdetail::A_[4],
    powdetail::C_[0] / 32.0,
    powdetail::C_[1] / 1024.0,
    powdetail::C_[2] / 32768.0,
    powdetail::C_[3] / 1048576.0,
    4503601774854144.0,
    32.0 * (double)(287.f / 1004.f), // kappa = r/cp as the float the reference computes (MC.h:47)
    32.0 * (double)(float)0.16,      // windCooling's exponent 0.16f (FC.cc:2213)
};

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

  // x^y for a positive NORMAL float x (2^-126 <= x < 2^128); `e_adjust` is added to the binary exponent
  // (the slow path rescales subnormals)
  template <int Y>
  __device__ __forceinline__ float pow_normal(float x, int e_adjust = 0) const
  {
    const double* K = c_pow_coef;
    const unsigned ix = __float_as_uint(x);
    const int e = (int)(ix >> 23) - 127 + e_adjust;
    const int i = (ix >> 18) & 31;
    // m = 1.mantissa and e as doubles, assembled with integer instructions: the F2F / I2F conversions
    // run on the quarter-rate XU pipe, which these kernels would otherwise saturate
    const unsigned mant = ix & 0x007fffffu;
    const double m = __hiloint2double((int)(0x3ff00000u | (mant >> 3)), (int)(mant << 29));
    const double ed = __hiloint2double(0x43300000, (int)(0x80000000u ^ (unsigned)e)) - K[9]; // (2^52 + 2^31 + e) - (2^52 + 2^31)
    const double2 t = lg[i];
    const double r = fma(m, t.x, -1.0);
    double p = fma(K[4], r, K[3]);
    p = fma(p, r, K[2]);
    p = fma(p, r, K[1]);
    p = fma(p, r, K[0]);
    const double lgx = fma(r, p, ed + t.y);
    const double ty = K[10 + Y] * lgx; // 32 * y * log2 x
    // split ty = k + g, k integer, |g| <= 1/2, with the round-to-nearest shift trick
    constexpr double SHIFT = 6755399441055744.0; // 1.5 * 2^52
    double kd = ty + SHIFT;
    const int k = __double2loint(kd);
    kd -= SHIFT;
    const double g = ty - kd;
    double s = fma(K[8], g, K[7]);
    s = fma(s, g, K[6]);
    s = fma(s, g, K[5]);
    const double tj = ex[k & 31];
    double res = fma(tj * g, s, tj);
    // multiply by 2^(k >> 5): |k >> 5| < 200, res in [1, 2) -> stays a normal double
    const int hi = __double2hiint(res) + ((k >> 5) << 20);
    res = __hiloint2double(hi, __double2loint(res));
    return (float)res;
  }

  // everything that is not a positive normal float.  powf(x, y) semantics for y > 0 non-integer: NaN for
  // finite x < 0 or NaN, +0 for +-0, +inf for +-inf
  template <int Y>
  __device__ __noinline__ float pow_special(float x) const
  {
    if (x > 0.f && x < 1.17549435e-38f)
      return pow_normal<Y>(x * 8388608.f, -23);
    if (x == 0.f)
      return 0.f;
    if (x > 0.f || x == __int_as_float(0xff800000))
      return __int_as_float(0x7f800000); // +-inf -> +inf (y is not an odd integer)
    return __int_as_float(0x7fc00000); // x < 0 or NaN
  }

  template <int Y>
  __device__ __forceinline__ float pow(float x) const
  {
    // positive normal <=> 0x00800000 <= bits < 0x7f800000 (the sign bit makes negatives huge)
    if (__float_as_uint(x) - 0x00800000u < 0x7f000000u)
      return pow_normal<Y>(x);
    return pow_special<Y>(x);
  }
};

} // namespace dev
} // namespace fcb200

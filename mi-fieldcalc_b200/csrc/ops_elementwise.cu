// ops_elementwise.cu -- pressure-, hybrid- and atmospheric-level thermodynamic conversions,
// windCooling, fieldOPERfield and the momentum coordinates.
//
// Every operator is a functor for the batched engine in elementwise.cuh.  Host code decodes
// `compute` / `unit` and validates arguments with the reference's exact early-return rules, then
// launches one kernel for the whole batch.  Citations: FC.cc = the reference's
// src/mi_fieldcalc/FieldCalculations.cc.
#include "ew_host.cuh"

#include "../../include/fcb200.h"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace fcb200 {
namespace {

using dev::is_def;
using dev::K_CP;
using dev::K_T0;
using dev::K_XLH;

enum Kind { PLEVEL = 0, HLEVEL = 1, ALEVEL = 2 };

inline bool bad_hlevel(float a, float b)
{ // FC.cc:298-301
  return (a < 0.0) || (b < 0.0) || (a == 0.0 && b == 0.0) || (b > 1.0);
}

int temp_compute(int compute, const char* unit)
{ // FC.cc:340-345, 1060-1065, 1322-1327
  if (compute < 3) {
    if (unit_is(unit, "celsius"))
      return 1;
    if (unit_is(unit, "kelvin"))
      return 2;
  }
  return compute;
}

int hum_compute(int compute, const char* unit)
{ // FC.cc:422-425, 1174-1177, 1417-1420
  if (compute > 8 && unit_is(unit, "celsius"))
    return compute - 4;
  if (compute > 4 && compute <= 8 && unit_is(unit, "kelvin"))
    return compute + 4;
  return compute;
}

// ------------------------------------------------------------------------------------ functors

// pleveltemp c1-3 (FC.cc:351-355): pure stream through unaryFunctionField, flag untouched
struct PTempStreamOp
{
  static constexpr int NIN = 1, NOUT = 1, UNROLL = 4;
  static constexpr int NCOUNT = 0;
  static constexpr int MIN_BLOCKS = 5;
  static constexpr bool HEAVY = false;
  static constexpr bool USES_EWT = false, USES_POW = false;
  int compute;
  template <bool ALL>
  __device__ __forceinline__ void point(const float* in, float* out, const PointCtx& c, long long, unsigned*) const
  {
    const float f = in[0], pidcp = c.m.a;
    if (ALL || is_def(f, c.undef))
      out[0] = (compute == 1) ? (f * pidcp - K_T0) : (compute == 2) ? (f * pidcp) : (f / pidcp);
    else
      out[0] = c.undef;
  }
};

// *leveltemp with a table lookup or a per-point Exner function.
// PLEVEL: only c4/c5 come here; meta.a = p, meta.b = pi.   HLEVEL: meta.a/b = alevel/blevel.
template <int KIND>
struct TempOp
{
  static constexpr bool SHARED_INPUT = (KIND == HLEVEL); // ps is one field for the whole batch
  static constexpr int NIN = (KIND == PLEVEL) ? 1 : 2, NOUT = 1, UNROLL = 2;
  static constexpr int NCOUNT = 1;
  static constexpr int MIN_BLOCKS = 4;
  static constexpr bool HEAVY = false;
  static constexpr bool USES_EWT = true, USES_POW = true;
  int compute;
  template <bool ALL>
  __device__ __forceinline__ void point(const float* in, float* out, const PointCtx& c, long long, unsigned* nundef) const
  {
    const float t = in[0];
    bool ok = ALL || is_def(t, c.undef);
    if (KIND != PLEVEL)
      ok = ALL || (ok && is_def(in[NIN - 1], c.undef));
    float r = c.undef;
    if (ok) {
      float p, pi, pidcp = 0.f;
      if (KIND == PLEVEL) {
        p = c.m.a;
        pi = c.m.b;
      } else {
        p = (KIND == HLEVEL) ? dev::p_hlevel(in[1], c.m.a, c.m.b) : in[1];
        pidcp = dev::pidcp_from_p(c.pw, p);
        pi = K_CP * pidcp;
      }
      if (compute == 1)
        r = t * pidcp - K_T0;
      else if (compute == 2)
        r = t * pidcp;
      else if (compute == 3)
        r = t / pidcp;
      else if (compute == 4)
        ok = dev::t_thesat(c.tab, t, p, pi, r);
      else
        ok = dev::th_thesat(c.tab, t, p, pi, r);
    }
    if (!ok) {
      r = c.undef;
      nundef[0] += 1;
    }
    out[0] = r;
  }
};

// h/alevelthe (FC.cc:1128-1139, 1378-1388)
template <int KIND>
struct TheOp
{
  static constexpr bool SHARED_INPUT = (KIND == HLEVEL); // ps is one field for the whole batch
  static constexpr int NIN = 3, NOUT = 1, UNROLL = 2;
  static constexpr int NCOUNT = 1;
  static constexpr int MIN_BLOCKS = 4;
  static constexpr bool HEAVY = false;
  static constexpr bool USES_EWT = false, USES_POW = true;
  int compute;
  template <bool ALL>
  __device__ __forceinline__ void point(const float* in, float* out, const PointCtx& c, long long, unsigned* nundef) const
  {
    const float t = in[0], q = in[1];
    if (ALL || (is_def(t, c.undef) && is_def(q, c.undef) && is_def(in[2], c.undef))) {
      const float p = (KIND == HLEVEL) ? dev::p_hlevel(in[2], c.m.a, c.m.b) : in[2];
      const float pi = K_CP * dev::pidcp_from_p(c.pw, p);
      out[0] = (compute == 1) ? ((t * K_CP + q * K_XLH) / pi) : (t + q * K_XLH / pi);
    } else {
      out[0] = c.undef;
      nundef[0] += 1;
    }
  }
};

// The 12-mode humidity conversion in the a/h-level numbering (FC.cc:1399-1410); plevelhum's own
// numbering is translated by the host.  PLEVEL: meta.a = p, meta.b = tconv (pi/cp for the even
// modes, 1 otherwise -- FC.cc:436), meta.all bit 1 = "p == undef, fill" (FC.cc:429-432).
template <int KIND>
struct HumOp
{
  static constexpr bool SHARED_INPUT = (KIND == HLEVEL); // ps is one field for the whole batch
  static constexpr int NIN = (KIND == PLEVEL) ? 2 : 3, NOUT = 1, UNROLL = 2;
  static constexpr int NCOUNT = 1;
  static constexpr int MIN_BLOCKS = 4;
  static constexpr bool HEAVY = false;
  static constexpr bool USES_EWT = true, USES_POW = true;
  int compute;
  int pcheck; // test `pin != undef` (no NaN test): HLEVEL when p is needed (FC.cc:1187), ALEVEL when it is NOT (FC.cc:1429)
  float tdconv;
  template <bool ALL>
  __device__ __forceinline__ void point(const float* in, float* out, const PointCtx& c, long long, unsigned* nundef) const
  {
    const float t = in[0], h = in[1];
    const bool all = ALL || (c.m.all & 1) != 0;
    bool ok = all || (is_def(t, c.undef) && is_def(h, c.undef));
    if (KIND == PLEVEL) {
      if (c.m.all & 2)
        ok = false;
    } else if (pcheck && !all) {
      ok = ok && (in[NIN - 1] != c.undef);
    }
    float r = c.undef;
    if (ok) {
      float p, tk = t;
      const bool even = (compute & 1) == 0;
      if (KIND == PLEVEL) {
        p = c.m.a;
        tk = t * c.m.b;
      } else if (KIND == HLEVEL) {
        const bool need_p = !(compute == 7 || compute == 11);
        p = need_p ? dev::p_hlevel(in[2], c.m.a, c.m.b) : 0.f;
        if (even)
          tk = t * dev::pidcp_from_p(c.pw, p);
      } else {
        p = in[2];
        if (even)
          tk = t * dev::pidcp_from_p(c.pw, p);
      }
      switch (compute) {
      case 1:
      case 2:
        ok = dev::tk_q_rh(c.tab, tk, h, p, r);
        break;
      case 3:
      case 4:
        ok = dev::tk_rh_q(c.tab, tk, h, p, r);
        break;
      case 5:
      case 6:
      case 9:
      case 10:
        ok = dev::tk_q_td(c.tab, tk, h, p, tdconv, r);
        break;
      default:
        ok = dev::tk_rh_td(c.tab, tk, h, tdconv, r);
        break;
      }
    }
    if (!ok) {
      r = c.undef;
      nundef[0] += 1;
    }
    out[0] = r;
  }
};

// h/alevelducting (FC.cc:1256-1271, 1490-1503)
template <int KIND>
struct DuctOp
{
  static constexpr bool SHARED_INPUT = (KIND == HLEVEL); // ps is one field for the whole batch
  static constexpr int NIN = 3, NOUT = 1, UNROLL = 2;
  static constexpr int NCOUNT = 1;
  static constexpr int MIN_BLOCKS = 4;
  static constexpr bool HEAVY = false;
  static constexpr bool USES_EWT = true, USES_POW = true;
  int compute;
  template <bool ALL>
  __device__ __forceinline__ void point(const float* in, float* out, const PointCtx& c, long long, unsigned* nundef) const
  {
    bool ok = ALL || (is_def(in[0], c.undef) && is_def(in[1], c.undef) && is_def(in[2], c.undef));
    float r = c.undef;
    if (ok) {
      const float p = (KIND == HLEVEL) ? dev::p_hlevel(in[2], c.m.a, c.m.b) : in[2];
      float tk = in[0];
      if ((compute & 1) == 0)
        tk *= dev::pidcp_from_p(c.pw, p);
      if (compute <= 2)
        r = dev::tk_q_duct(tk, in[1], p);
      else
        ok = dev::tk_rh_duct(c.tab, tk, in[1], p, r);
    }
    if (!ok) {
      r = c.undef;
      nundef[0] += 1;
    }
    out[0] = r;
  }
};

// The reference does not validate `compute` in hleveltemp, hlevelthe, hlevelducting and alevelducting:
// with an unknown mode a defined point keeps whatever the output array held and an undefined point
// becomes undef (e.g. FC.cc:1080-1094).  in[NCHK] is the previous content of the output.
template <int NCHK>
struct KeepOrUndefOp
{
  static constexpr bool SHARED_INPUT = true;
  static constexpr int NIN = NCHK + 1, NOUT = 1, UNROLL = 2;
  static constexpr int NCOUNT = 1;
  static constexpr int MIN_BLOCKS = 5;
  static constexpr bool HEAVY = false;
  static constexpr bool USES_EWT = false, USES_POW = false;
  template <bool ALL>
  __device__ __forceinline__ void point(const float* in, float* out, const PointCtx& c, long long, unsigned* nundef) const
  {
    bool ok = true;
#pragma unroll
    for (int k = 0; k < NCHK; ++k)
      ok = ok && is_def(in[k], c.undef);
    if (ALL || ok)
      out[0] = in[NCHK];
    else {
      out[0] = c.undef;
      nundef[0] += 1;
    }
  }
};

// hlevelpressure (FC.cc:1294-1301)
struct HPressureOp
{
  static constexpr int NIN = 1, NOUT = 1, UNROLL = 4;
  static constexpr int NCOUNT = 1;
  static constexpr int MIN_BLOCKS = 5;
  static constexpr bool HEAVY = false;
  static constexpr bool USES_EWT = false, USES_POW = false;
  template <bool ALL>
  __device__ __forceinline__ void point(const float* in, float* out, const PointCtx& c, long long, unsigned* nundef) const
  {
    if (ALL || is_def(in[0], c.undef))
      out[0] = dev::p_hlevel(in[0], c.m.a, c.m.b);
    else {
      out[0] = c.undef;
      nundef[0] += 1;
    }
  }
};

// windCooling (FC.cc:2209-2221); the reference never updates the flag.  Same construction as the fused alevel chain: a
// branch-free path for plausible winds (2^-100 <= u*u + v*v < 2^100: the square root and the power need no range guards, the
// double constants come from the constant bank), the four points of a float4 group in one basic block, and a non-inlined
// redo with the ordinary operators for calm (u = v = 0), NaN, infinite or absurd winds.  The temperature needs no test: it only
// enters double arithmetic that behaves identically for every bit pattern.
#ifndef FCB_PL_U
#define FCB_PL_U 2
#define FCB_PL_MB 3
#define FCB_PL_MB4 4
#endif
#ifndef FCB_WC_U
#define FCB_WC_U 2
#define FCB_WC_MB 3
#endif
struct WindCoolingOp
{
  static constexpr int NIN = 3, NOUT = 1, UNROLL = FCB_WC_U;
  static constexpr int NCOUNT = 0;
  static constexpr int MIN_BLOCKS = FCB_WC_MB;
  static constexpr bool HEAVY = false;
  static constexpr bool USES_EWT = false, USES_POW = true;
  static constexpr bool QUAD = true;
  float tconv;

  __device__ __forceinline__ static float chill(float tc, float ffpow)
  { // FC.cc:2214-2216
    const double* K = dev::c_dconst;
    float d = (float)(K[2] + K[3] * (double)tc - K[4] * (double)ffpow + K[5] * (double)tc * (double)ffpow);
    if (d > 0.f)
      d = 0.f;
    return d;
  }

  __device__ __noinline__ static float ieee(float tc, float u, float v, const dev::PowTable& pw)
  {
    const float ff = (float)((double)dev::absval(u, v) * 3.6);
    return chill(tc, pw.pow<dev::POW_WINDCHILL>(ff));
  }

  __device__ __forceinline__ bool fast(float tc, float u, float v, const dev::PowTable& pw, float& d) const
  {
    const float s = u * u + v * v;
    const bool plausible = __float_as_uint(s) - 0x0d800000u < 0x64000000u; // 2^-100 <= s < 2^100
    const float ff = (float)((double)dev::sqrt_midrange(s) * dev::c_dconst[1]);
    d = chill(tc, pw.pow_normal<dev::POW_WINDCHILL>(ff));
    return plausible;
  }

  template <bool ALL>
  __device__ __forceinline__ void point(const float* in, float* out, const PointCtx& c, long long, unsigned*) const
  {
    const bool def = ALL || (is_def(in[0], c.undef) && is_def(in[1], c.undef) && is_def(in[2], c.undef));
    const float tc = in[0] - tconv;
    float d;
    if (!fast(tc, in[1], in[2], c.pw, d) && def)
      d = ieee(tc, in[1], in[2], c.pw);
    out[0] = def ? d : c.undef;
  }

  template <bool ALL>
  __device__ __forceinline__ void quad(const float (*in)[4], float (*out)[4], const PointCtx& c, unsigned*) const
  {
    float d[4];
    unsigned bad = 0, defined = 0;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const bool def = ALL || (is_def(in[0][w], c.undef) && is_def(in[1][w], c.undef) && is_def(in[2][w], c.undef));
      defined |= def ? (1u << w) : 0u;
      if (!fast(in[0][w] - tconv, in[1][w], in[2][w], c.pw, d[w]) && def)
        bad |= 1u << w;
    }
    if (bad) {
#pragma unroll
      for (int w = 0; w < 4; ++w)
        if (bad & (1u << w))
          d[w] = ieee(in[0][w] - tconv, in[1][w], in[2][w], c.pw);
    }
#pragma unroll
    for (int w = 0; w < 4; ++w)
      out[0][w] = (defined & (1u << w)) ? d[w] : c.undef;
  }
};

// fieldOPERfield (FC.cc:2611-2625)
struct FieldOperOp
{
  static constexpr int NIN = 2, NOUT = 1, UNROLL = 4;
  static constexpr int NCOUNT = 1;
  static constexpr int MIN_BLOCKS = 5;
  static constexpr bool HEAVY = false;
  static constexpr bool USES_EWT = false, USES_POW = false;
  int compute;
  template <bool ALL>
  __device__ __forceinline__ void point(const float* in, float* out, const PointCtx& c, long long, unsigned* nundef) const
  {
    const float a = in[0], b = in[1];
    float r = c.undef;
    bool ok = ALL || (is_def(a, c.undef) && is_def(b, c.undef));
    if (ok) {
      if (compute == 1)
        r = a + b;
      else if (compute == 2)
        r = a - b;
      else if (compute == 3)
        r = a * b;
      else if (b != 0)
        r = a / b;
      else
        ok = false; // divideUndef, FC.cc:84-92
    }
    if (!ok) {
      r = c.undef;
      nundef[0] += 1;
    }
    out[0] = r;
  }
};

// momentumX/Ycoordinate (FC.cc:2371-2383, 2407-2419): elementwise despite the name
template <bool XDIR>
struct MomentumOp
{
  static constexpr int NIN = 3, NOUT = 1, UNROLL = 2;
  static constexpr int NCOUNT = 1;
  static constexpr int MIN_BLOCKS = 5;
  static constexpr bool HEAVY = false;
  static constexpr bool USES_EWT = false, USES_POW = false;
  float fcormin;
  template <bool ALL>
  __device__ __forceinline__ void point(const float* in, float* out, const PointCtx& c, long long idx, unsigned* nundef) const
  {
    if (ALL || is_def(in[0], c.undef)) {
      float fcor = in[2];
      const float fcormax = -fcormin;
      if (fcor >= 0.f && fcor < fcormin)
        fcor = fcormin;
      else if (fcor <= 0.f && fcor > fcormax)
        fcor = fcormax;
      const int i = (int)idx;
      if (XDIR)
        out[0] = (float)(i % c.nx) + in[0] * in[1] / fcor;
      else
        out[0] = (float)(i / c.nx) - in[0] * in[1] / fcor;
    } else {
      out[0] = c.undef;
      nundef[0] += 1;
    }
  }
};


// The MEPS alevel chain (BASELINE.json configs[1]) fused: one read of t, q, p produces what the
// reference computes with four calls that each re-read their inputs (12 + 16 + 16 + 16 = 60 B/point
// -> 28 B/point):
//   out[0] theta   = aleveltemp(c=3)          t / pidcp(p)                         test t, p     FC.cc:1334-1340
//   out[1] RH (%)  = alevelhum(c=1)           tk_q_rh(t, q, p)                     test t, q     FC.cc:1429-1431
//   out[2] Td      = alevelhum(c=5 or 9)      tk_q_td(t, q, p, tdconv)             test t, q     FC.cc:1440-1441
//   out[3] theta_e = alevelthe(c=1)           (t*cp + q*xlh) / (cp*pidcp(p))       test t, q, p  FC.cc:1379-1382
// Each output keeps its own definedness test and its own undefined counter, so values, masks and
// the four flags are exactly those of the four separate calls.  The Exner function and the
// saturation-table lookup are evaluated once per point instead of three / two times.
//
// The same code serves the four operators on their own (aleveltemp c3, alevelhum c1, alevelhum c5/9, alevelthe c1):
// OUTS selects which outputs exist, everything an absent output would need is compiled out.
enum : unsigned {
  O_THETA = 1, O_RH = 2, O_TD = 4, O_THE = 8, O_ALL = 15,
  O_THESAT = 16, // *leveltemp c4: T -> theta_e,sat (FC.cc:196-205)
  O_TDRH = 32    // *levelhum T, RH(%) -> Td (a/h-level c7/11, p-level c5/9; tk_rh_td, FC.cc:254-267): the second input is RH, p is not used
};

// KIND = PLEVEL: the pressure is the field's scalar (FieldMeta::a) instead of a third input field -- plevelhum.
// KIND = HLEVEL: the last input is the surface pressure, p = alevel + blevel * ps with the field's FieldMeta::a, ::b
// (FC.cc:303-306) -- hleveltemp, hlevelhum, whose humidity modes test `ps != undef` (no NaN test, FC.cc:1187) where the
// a-level ones test nothing.
template <int U_, int MB_, int J_ = 1, unsigned OUTS = O_ALL, int KIND = ALEVEL, int PK_ = 1>
struct AlevelChainOpT
{
  static constexpr int PACK = PK_; // 1: two points per packed FP32 instruction (fast2), 0: one point at a time (fast)
  static constexpr bool HAS_Q = (OUTS & (O_RH | O_TD | O_THE | O_TDRH)) != 0; // q (or RH) is an input
  static constexpr bool HAS_TAB = (OUTS & (O_RH | O_TD | O_THESAT | O_TDRH)) != 0; // the saturation table is used
  static constexpr bool USES_P = (OUTS & ~O_TDRH) != 0;                          // the pressure enters the arithmetic
  static constexpr bool Q_IS_RATIO = (OUTS & (O_RH | O_TD | O_THE)) != 0;          // q is divided / divides: its range matters
  static constexpr bool WIDE_P = KIND == ALEVEL && (OUTS & (O_RH | O_TD)) != 0; // an undefined p is an operand of a live output (see `fast`)
  static constexpr bool HAS_POW = (OUTS & (O_THETA | O_THE)) != 0 || ((OUTS & O_THESAT) != 0 && KIND != PLEVEL); // the Exner function is used
  static constexpr bool SHARED_INPUT = (KIND == HLEVEL);
  static constexpr int ITEM_ROUNDS = J_;
  static constexpr int NIN = (HAS_Q ? 2 : 1) + (KIND != PLEVEL ? 1 : 0); // t, [q,] [p or ps]
  static constexpr int NOUT = ((OUTS & 1) ? 1 : 0) + ((OUTS & 2) ? 1 : 0) + ((OUTS & 4) ? 1 : 0) + ((OUTS & 8) ? 1 : 0) + ((OUTS & 16) ? 1 : 0) + ((OUTS & 32) ? 1 : 0);
  static constexpr int UNROLL = U_;
  static constexpr int NCOUNT = NOUT;
  static constexpr int MIN_BLOCKS = MB_;
  static constexpr bool HEAVY = false;
  static constexpr bool USES_EWT = HAS_TAB, USES_POW = HAS_POW;
  static constexpr bool QUAD = true;
  float tdconv;

  struct Raw
  {
    float theta, rh, td, the, thesat, tdrh;
    bool edef; // the saturation-table lookup of t was in range
  };

  // The reference's expressions with ordinary IEEE operators: correct for ANY bit pattern.  Not inlined: it
  // only runs for points whose inputs fail the plausibility test of `fast` (undefined values that flow into
  // the arithmetic, NaN, zero or negative pressure, ...).
  // `pi_field` = the field's pi = cp * (p/p0)^kappa evaluated on the host (KIND == PLEVEL, O_THESAT only)
  __device__ __noinline__ static void ieee_raw(float t, float q, float p, const dev::EwtTable& tab, const dev::PowTable& pw, float tdconv, float pi_field,
                                               Raw& r)
  {
    float pi = pi_field;
    if (HAS_POW) {
      const float pidcp = dev::pidcp_from_p(pw, p);
      pi = K_CP * pidcp;
      if (OUTS & O_THETA)
        r.theta = t / pidcp;
      if (OUTS & O_THE)
        r.the = (t * K_CP + q * K_XLH) / (K_CP * pidcp);
    }
    r.edef = true;
    if (HAS_TAB) {
      const dev::Ewt e(t - K_T0);
      const float et = e.value(tab);
      const float qsat = dev::K_EPS * et / p;
      if (OUTS & O_THESAT)
        r.thesat = (K_CP * t + K_XLH * qsat) / pi;
      if (OUTS & O_RH)
        r.rh = (float)(100. * (double)q / (double)qsat);
      if (OUTS & O_TD) {
        const float rhc = dev::clamp_rh(q / qsat);
        r.td = e.inverse(tab, rhc * et) + tdconv;
      }
      if (OUTS & O_TDRH) {
        const float rhc = dev::clamp_rh((float)(0.01 * (double)q));
        r.tdrh = e.inverse(tab, rhc * et) + tdconv;
      }
      r.edef = e.defined;
    }
  }
  // `out` of the non-inlined call lives in local memory; copying it keeps the caller's own Raw in registers
  __device__ __forceinline__ static void ieee(float t, float q, float p, const dev::EwtTable& tab, const dev::PowTable& pw, float tdconv, float pi_field,
                                              Raw& r)
  {
    Raw tmp;
    ieee_raw(t, q, p, tab, pw, tdconv, pi_field, tmp);
    r.thesat = tmp.thesat;
    r.tdrh = tmp.tdrh;
    r.theta = tmp.theta;
    r.rh = tmp.rh;
    r.td = tmp.td;
    r.the = tmp.the;
    r.edef = tmp.edef;
  }

  // Same expressions, same roundings, for PLAUSIBLE inputs: saturation-table position x in [0, 40)
  // (-100 <= t - 273.15 < 100 degC; without the table: 2^-59 <= |t| < 2^59), p in [2^-7, 2^11) hPa, q = +0 or
  // 2^-90 <= |q| < 2^20.  Then every divisor, quotient and remainder below is a normal number far from the ends of the
  // exponent range, the divisions can use the guard-free sequences of device_common.cuh, the Exner function needs
  // no classification of its argument and the table indices need no clamps -- straight-line code without a single
  // branch, so the compiler interleaves the four points of a thread.  Returns false (after computing harmless
  // garbage) when the inputs are not plausible; the caller then redoes the point with `ieee`.
  //
  // WIDE (a-level humidity outputs of a field that is not ALL_DEFINED): alevelhum tests t and q only, so an UNDEFINED p -- the
  // huge undefined value itself, 1e35 -- flows into RH and Td (FC.cc:1429).  With independent masks that is one point in seven,
  // and the IEEE redo of those points, one lane at a time, halved the masked operators.  The same straight-line code is valid
  // for them: the preconditions of div_midrange (divisor, quotient and dividend * 2^-24 normal) hold for 2^64 <= p < 2^126 as long
  // as qsat = eps * et / p comes out NORMAL and |q| < 2 (q / qsat < 2^128; the double quotient has range to spare); RH then
  // overflows to +inf in the float conversion exactly as in the reference, and Td sees the clamped ratio.  Tested per point.
  template <bool WIDE = false>
  __device__ __forceinline__ bool fast(float t, float q, float p, float pi_field, const dev::EwtTable& tab, const dev::PowTable& pw, Raw& r) const
  {
    float pi = pi_field;
    bool plausible = !USES_P || (__float_as_uint(p) - 0x3c000000u < 0x09000000u); // 2^-7 <= p < 2^11
    const bool wide = WIDE && (__float_as_uint(p) - 0x5f800000u < 0x1f000000u);    // 2^64 <= p < 2^126
    if (WIDE)
      plausible = plausible || (wide && (__float_as_uint(q) & 0x7fffffffu) < 0x40000000u);
    if (Q_IS_RATIO) {
      const unsigned uq = __float_as_uint(q) & 0x7fffffffu;
      plausible = plausible && ((uq - 0x12800000u < 0x37000000u) || __float_as_uint(q) == 0u); // 2^-90 <= |q| < 2^20, or +0
    }
    if (!HAS_TAB)
      plausible = plausible && ((__float_as_uint(t) & 0x7fffffffu) - 0x22000000u < 0x3b000000u); // 2^-59 <= |t| < 2^59

    float rpi = 0.f; // 1 / pi
    if (HAS_POW) {
      // FC.cc:308-316: pidcp = powf(p * p0inv, kappa), pi = cp * pidcp; theta = t / pidcp, theta_e = (...) / pi.  Here the
      // RECIPROCAL Exner factor from the two special-function instructions (dev::exner_recip: relative error < 1e-6, see there)
      // and a multiplication instead of powf + division: ~45 issue slots per point less.  theta, theta_e and theta_e,sat are
      // the outputs north_star gives a tolerance for (1e-5, "transcendental differences documented"; tests/cases.py
      // TRANSCENDENTAL); the Exner factor never reaches a definedness test or a table index, so masks, flags, RH and Td
      // stay bit for bit.  Implausible pressures take `ieee` with the exact powf as before.
      const float rpid = dev::exner_recip(p);
      rpi = rpid * dev::K_CPINV;
      if (OUTS & O_THETA)
        r.theta = t * rpid;
      if (OUTS & O_THE)
        r.the = (t * K_CP + q * K_XLH) * rpi;
    }
    if (HAS_TAB) {
      const float x = (float)(((double)(t - K_T0) + 100.) * dev::c_dconst[0]); // Ewt::Ewt, MC.h:66
      plausible = plausible && (__float_as_uint(x) < 0x42200000u);              // +0 <= x < 40
      const int l = plausible ? (int)x : 0;
      const float2 e = tab.e[l];
      const float et = e.x + e.y * (x - (float)l); // MC.h:78
      const float qsat = dev::div_midrange(dev::K_EPS * et, p);
      if (WIDE)
        plausible = plausible && (!wide || __float_as_uint(qsat) - 0x00800000u < 0x7f000000u); // a positive normal qsat
      if (OUTS & O_THESAT) // t_thesat, FC.cc:196-205 (pi in [36, 1240] for plausible p; the field's own pi comes from the host's powf)
        r.thesat = (HAS_POW && KIND != PLEVEL) ? (K_CP * t + K_XLH * qsat) * rpi : dev::div_midrange(K_CP * t + K_XLH * qsat, pi);
      // (Evaluating this double quotient in float-float arithmetic with a midpoint test -- no conversions, no FP64 -- was
      // measured 8 % SLOWER: the test's dependent chain costs more than the nine DFMA it replaces.)
      if (OUTS & O_RH)
        r.rh = (float)dev::div_midrange(100. * (double)q, (double)qsat); // FC.cc:229
      // dew point from a relative humidity (a NaN stays a NaN through the selects, like clamp_rh, FC.cc:186-194)
      auto dewpoint = [&](float rh) {
        const float rhc = rh < dev::K_RHMIN ? dev::K_RHMIN : (rh > dev::K_RHMAX ? dev::K_RHMAX : rh);
        const float etd = rhc * et;
        // Ewt::inverse (MC.cc:37-45) with the bucket table; 0.02 * ewt[0] < 2^-15 lands below the first bucket -> index 0
        int b = (int)(__float_as_uint(etd) >> 21) - dev::EWT_LUT0;
        b = min(max(b, 0), dev::EWT_NLUT - 1);
        int ll = min((int)tab.lut[b], l);
        const int k = min(ll + 1, l); // the bucket index is the answer or one below it; k == ll when the walk may not go up
        ll = (tab.e[k].x > etd) ? ll : k;
        const float2 e2 = tab.e[ll];
        const float y = (float)ll + dev::div_midrange(etd - e2.x, e2.y);
        // (float)(-100. + (double)y * 5.): for |y| < 64 the double expression is exact (5y has at most 27
        // significant bits and -100 + 5y at most 7 + 46) or, for |y| < 2^-23, rounds to -100 either way -> one fmaf
        return fmaf(5.f, y, -100.f) + tdconv;
      };
      if (OUTS & O_TD)
        r.td = dewpoint(dev::div_midrange(q, qsat));
      if (OUTS & O_TDRH)
        r.tdrh = dewpoint((float)(0.01 * (double)q));
    }
    r.edef = true;
    return plausible;
  }

  // `fast` for TWO points at once: the float arithmetic on Blackwell's packed FP32 instructions (FMUL2 / FFMA2 / FADD2: one issue
  // slot for both points; these kernels are issue-bound, the FMA pipe is not).  Same operations, same roundings -- the packed
  // instructions are IEEE round-to-nearest like the scalar ones; every reference expression that adds a rounded product is
  // written with a scalar add (ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with -fmad=false).  The table
  // lookups, the Exner function and the double-precision RH quotient stay per point.  Returns the implausible points as bits 0, 1.
  template <bool WIDE = false>
  __device__ __forceinline__ unsigned fast2(float2 t, float2 q, float2 p, float pi_field, const dev::EwtTable& tab, const dev::PowTable& pw, Raw& r0, Raw& r1) const
  {
    using namespace dev;
    bool pl0 = !USES_P || (__float_as_uint(p.x) - 0x3c000000u < 0x09000000u); // 2^-7 <= p < 2^11
    bool pl1 = !USES_P || (__float_as_uint(p.y) - 0x3c000000u < 0x09000000u);
    const bool w0 = WIDE && (__float_as_uint(p.x) - 0x5f800000u < 0x1f000000u), w1 = WIDE && (__float_as_uint(p.y) - 0x5f800000u < 0x1f000000u); // see `fast`
    if (WIDE) {
      pl0 = pl0 || (w0 && (__float_as_uint(q.x) & 0x7fffffffu) < 0x40000000u);
      pl1 = pl1 || (w1 && (__float_as_uint(q.y) & 0x7fffffffu) < 0x40000000u);
    }
    if (Q_IS_RATIO) {
      const unsigned u0 = __float_as_uint(q.x) & 0x7fffffffu, u1 = __float_as_uint(q.y) & 0x7fffffffu;
      pl0 = pl0 && ((u0 - 0x12800000u < 0x37000000u) || __float_as_uint(q.x) == 0u); // 2^-90 <= |q| < 2^20, or +0
      pl1 = pl1 && ((u1 - 0x12800000u < 0x37000000u) || __float_as_uint(q.y) == 0u);
    }
    if (!HAS_TAB) {
      pl0 = pl0 && ((__float_as_uint(t.x) & 0x7fffffffu) - 0x22000000u < 0x3b000000u); // 2^-59 <= |t| < 2^59
      pl1 = pl1 && ((__float_as_uint(t.y) & 0x7fffffffu) - 0x22000000u < 0x3b000000u);
    }
    const float2 pi = make_float2(pi_field, pi_field);
    float2 rpi = make_float2(0.f, 0.f);
    if (HAS_POW) { // the reciprocal Exner factor (see `fast`)
      const float2 rpid = make_float2(exner_recip(p.x), exner_recip(p.y));
      rpi = pk_mul(rpid, K_CPINV);
      if (OUTS & O_THETA) {
        const float2 th = pk_mul(t, rpid);
        r0.theta = th.x, r1.theta = th.y;
      }
      if (OUTS & O_THE) {
        const float2 the = pk_mul(pk_add_products(pk_mul(t, K_CP), pk_mul(q, K_XLH)), rpi);
        r0.the = the.x, r1.the = the.y;
      }
    }
    if (HAS_TAB) {
      const float2 tc = pk_add(t, -K_T0);
      const float x0 = (float)(((double)tc.x + 100.) * c_dconst[0]), x1 = (float)(((double)tc.y + 100.) * c_dconst[0]); // Ewt::Ewt, MC.h:66
      pl0 = pl0 && (__float_as_uint(x0) < 0x42200000u); // +0 <= x < 40
      pl1 = pl1 && (__float_as_uint(x1) < 0x42200000u);
      const int l0 = pl0 ? (int)x0 : 0, l1 = pl1 ? (int)x1 : 0;
      const float2 e0 = tab.e[l0], e1 = tab.e[l1];
      const float2 et = make_float2(e0.x + e0.y * (x0 - (float)l0), e1.x + e1.y * (x1 - (float)l1)); // MC.h:78
      const float2 qsat = pk_div_midrange(pk_mul(et, K_EPS), p);
      if (WIDE) {
        pl0 = pl0 && (!w0 || __float_as_uint(qsat.x) - 0x00800000u < 0x7f000000u);
        pl1 = pl1 && (!w1 || __float_as_uint(qsat.y) - 0x00800000u < 0x7f000000u);
      }
      if (OUTS & O_THESAT) {
        const float2 num = pk_add_products(pk_mul(t, K_CP), pk_mul(qsat, K_XLH)); // t_thesat, FC.cc:196-205
        const float2 ts = (HAS_POW && KIND != PLEVEL) ? pk_mul(num, rpi) : pk_div_midrange(num, pi);
        r0.thesat = ts.x, r1.thesat = ts.y;
      }
      if (OUTS & O_RH) {
        r0.rh = (float)div_midrange(100. * (double)q.x, (double)qsat.x); // FC.cc:229
        r1.rh = (float)div_midrange(100. * (double)q.y, (double)qsat.y);
      }
      // dew point from a relative humidity (see `fast`)
      auto dewpoint2 = [&](float2 rh) {
        const float2 rhc = make_float2(rh.x < K_RHMIN ? K_RHMIN : (rh.x > K_RHMAX ? K_RHMAX : rh.x), rh.y < K_RHMIN ? K_RHMIN : (rh.y > K_RHMAX ? K_RHMAX : rh.y));
        const float2 etd = pk_mul(rhc, et);
        int b0 = (int)(__float_as_uint(etd.x) >> 21) - EWT_LUT0, b1 = (int)(__float_as_uint(etd.y) >> 21) - EWT_LUT0;
        b0 = min(max(b0, 0), EWT_NLUT - 1), b1 = min(max(b1, 0), EWT_NLUT - 1);
        int ll0 = min((int)tab.lut[b0], l0), ll1 = min((int)tab.lut[b1], l1);
        const int k0 = min(ll0 + 1, l0), k1 = min(ll1 + 1, l1);
        ll0 = (tab.e[k0].x > etd.x) ? ll0 : k0;
        ll1 = (tab.e[k1].x > etd.y) ? ll1 : k1;
        const float2 f0 = tab.e[ll0], f1 = tab.e[ll1];
        const float2 quo = pk_div_midrange(make_float2(etd.x - f0.x, etd.y - f1.x), make_float2(f0.y, f1.y));
        const float2 y = pk_add(make_float2((float)ll0, (float)ll1), quo);
        return pk_add(pk_fma(make_float2(5.f, 5.f), y, make_float2(-100.f, -100.f)), tdconv);
      };
      if (OUTS & O_TD) {
        const float2 td = dewpoint2(pk_div_midrange(q, qsat));
        r0.td = td.x, r1.td = td.y;
      }
      if (OUTS & O_TDRH) {
        const float2 td = dewpoint2(make_float2((float)(0.01 * (double)q.x), (float)(0.01 * (double)q.y)));
        r0.tdrh = td.x, r1.tdrh = td.y;
      }
    }
    r0.edef = true, r1.edef = true;
    return (pl0 ? 0u : 1u) | (pl1 ? 0u : 2u);
  }

  // definedness tests of the reference calls + one counter per output
  template <bool ALL>
  __device__ __forceinline__ void finish(float t, float q, float p, const Raw& r, const PointCtx& c, float* out, unsigned* nundef) const
  {
    const float undef = c.undef;
    // PLEVEL: FieldMeta::all bit 0 = the flag is ALL_DEFINED, bit 1 = p == undef -> every point undefined (FC.cc:429-432)
    const bool all = ALL || (KIND == PLEVEL && (c.m.all & 1) != 0);
    const bool dt = (all || is_def(t, undef)) && !(KIND == PLEVEL && (c.m.all & 2) != 0);
    const bool dq = all || !HAS_Q || is_def(q, undef), dp = all || KIND == PLEVEL || is_def(p, undef);
    const bool ok_theta = dt && dp;          // a/hleveltemp test t, p (ps)
    // alevelhum c1/c5 test t, q only -- an undefined p flows into the arithmetic; hlevelhum also tests ps != undef
    const bool ok_hum = dt && dq && r.edef && (KIND != HLEVEL || all || p != undef);
    const bool ok_the = dt && dq && dp;      // alevelthe tests t, q, p
    int o = 0;
    if (OUTS & O_THETA) {
      out[o] = ok_theta ? r.theta : undef;
      nundef[o++] += ok_theta ? 0u : 1u;
    }
    if (OUTS & O_RH) {
      out[o] = ok_hum ? r.rh : undef;
      nundef[o++] += ok_hum ? 0u : 1u;
    }
    if (OUTS & O_TD) {
      out[o] = ok_hum ? r.td : undef;
      nundef[o++] += ok_hum ? 0u : 1u;
    }
    if (OUTS & O_THE) {
      out[o] = ok_the ? r.the : undef;
      nundef[o++] += ok_the ? 0u : 1u;
    }
    if (OUTS & O_TDRH) { // alevelhum c7/11 test `p != undef` although p is not used (FC.cc:1429); hlevelhum and plevelhum do not
      const bool ok = dt && dq && r.edef && (KIND != ALEVEL || all || p != undef);
      out[o] = ok ? r.tdrh : undef;
      nundef[o++] += ok ? 0u : 1u;
    }
    if (OUTS & O_THESAT) { // *leveltemp c4 tests t and p (ps), then the table range
      const bool ok = ok_theta && r.edef;
      out[o] = ok ? r.thesat : undef;
      nundef[o++] += ok ? 0u : 1u;
    }
  }

  // For a field that is not ALL_DEFINED: an undefined q only reaches outputs that are undefined anyway, so it is
  // replaced by +0 (plausible) for the arithmetic; the IEEE redo is only needed if some output of the point can still
  // be defined -- never with an undefined t; with an undefined p only RH and Td, into which it flows (FC.cc:1429).
  // `praw` = the last input as stored (p, or ps for HLEVEL); the level pressure the arithmetic uses
  __device__ __forceinline__ static float level_p(float praw, const PointCtx& c)
  {
    return KIND == PLEVEL ? c.m.a : (KIND == HLEVEL ? dev::p_hlevel(praw, c.m.a, c.m.b) : praw);
  }

  template <bool ALL>
  __device__ __forceinline__ bool eval(float t, float q, float praw, const PointCtx& c, Raw& r) const
  {
    const bool dq = ALL || !HAS_Q || is_def(q, c.undef);
    const float qe = dq ? q : 0.f;
    const bool plausible = fast<WIDE_P && !ALL>(t, qe, level_p(praw, c), c.m.b, c.tab, c.pw, r);
    if (ALL)
      return plausible;
    const bool dt = is_def(t, c.undef), dp = KIND == PLEVEL || is_def(praw, c.undef);
    const bool hum_live = (OUTS & (O_RH | O_TD)) != 0 && dq && (KIND != HLEVEL || praw != c.undef);
    const bool live = dt && (hum_live || ((OUTS & O_TDRH) && dq && (KIND != ALEVEL || praw != c.undef)) || ((OUTS & O_THESAT) && dp) ||
                             (HAS_POW && dp && ((OUTS & O_THETA) || dq)));
    return plausible || !live;
  }

  // `eval` for two points (bits 0, 1 of the result: the point needs the IEEE redo)
  template <bool ALL>
  __device__ __forceinline__ unsigned eval2(float2 t, float2 q, float2 praw, const PointCtx& c, Raw& r0, Raw& r1) const
  {
    const bool dq0 = ALL || !HAS_Q || is_def(q.x, c.undef), dq1 = ALL || !HAS_Q || is_def(q.y, c.undef);
    const float2 qe = make_float2(dq0 ? q.x : 0.f, dq1 ? q.y : 0.f);
    const float2 pl = make_float2(level_p(praw.x, c), level_p(praw.y, c));
    const unsigned implausible = fast2<WIDE_P && !ALL>(t, qe, pl, c.m.b, c.tab, c.pw, r0, r1);
    if (ALL)
      return implausible;
    auto live = [&](float tt, bool dq, float pr) {
      const bool dt = is_def(tt, c.undef), dp = KIND == PLEVEL || is_def(pr, c.undef);
      const bool hum_live = (OUTS & (O_RH | O_TD)) != 0 && dq && (KIND != HLEVEL || pr != c.undef);
      return dt && (hum_live || ((OUTS & O_TDRH) && dq && (KIND != ALEVEL || pr != c.undef)) || ((OUTS & O_THESAT) && dp) ||
                    (HAS_POW && dp && ((OUTS & O_THETA) || dq)));
    };
    return (((implausible & 1u) && live(t.x, dq0, praw.x)) ? 1u : 0u) | (((implausible & 2u) && live(t.y, dq1, praw.y)) ? 2u : 0u);
  }

  // The IEEE redo of one point.  If the field is not ALL_DEFINED and p (ps) itself is undefined, theta and theta_e are undefined
  // whatever the arithmetic says and only RH / Td -- into which the undefined p flows (FC.cc:1429) -- are needed: that redo is
  // the humidity-only instantiation, without the Exner function.
  template <bool ALL>
  __device__ __forceinline__ void redo(float t, float q, float praw, const PointCtx& c, Raw& r) const
  {
    constexpr unsigned HUM = OUTS & (O_RH | O_TD);
    if constexpr (!ALL && HUM != 0 && HUM != OUTS && KIND != PLEVEL) {
      if (!is_def(praw, c.undef)) {
        typename AlevelChainOpT<U_, MB_, J_, HUM, KIND, PK_>::Raw h;
        AlevelChainOpT<U_, MB_, J_, HUM, KIND, PK_>::ieee_raw(t, q, level_p(praw, c), c.tab, c.pw, tdconv, c.m.b, h);
        r.rh = h.rh;
        r.td = h.td;
        r.edef = h.edef;
        return;
      }
    }
    ieee(t, q, level_p(praw, c), c.tab, c.pw, tdconv, c.m.b, r);
  }

  __device__ __forceinline__ static float Q(const float* in) { return HAS_Q ? in[1] : 0.f; }
  __device__ __forceinline__ static float P(const float* in) { return KIND == PLEVEL ? 0.f : in[NIN - 1]; }

  template <bool ALL>
  __device__ __forceinline__ void point(const float* in, float* out, const PointCtx& c, long long, unsigned* nundef) const
  {
    Raw r;
    if (!eval<ALL>(in[0], Q(in), P(in), c, r))
      redo<ALL>(in[0], Q(in), P(in), c, r);
    finish<ALL>(in[0], Q(in), P(in), r, c, out, nundef);
  }

  // four consecutive points of one thread: all fast evaluations first (one basic block), the rare IEEE redo after
  template <bool ALL>
  __device__ __forceinline__ void quad(const float (*in)[4], float (*out)[4], const PointCtx& c, unsigned* nundef) const
  {
    Raw r[4];
    unsigned bad = 0;
    if constexpr (PACK != 0) {
#pragma unroll
      for (int w = 0; w < 4; w += 2) {
        const float2 t = make_float2(in[0][w], in[0][w + 1]);
        const float2 q = HAS_Q ? make_float2(in[1][w], in[1][w + 1]) : make_float2(0.f, 0.f);
        const float2 pr = KIND == PLEVEL ? make_float2(0.f, 0.f) : make_float2(in[NIN - 1][w], in[NIN - 1][w + 1]);
        bad |= eval2<ALL>(t, q, pr, c, r[w], r[w + 1]) << w;
      }
    } else {
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const float q = HAS_Q ? in[1][w] : 0.f;
        bad |= eval<ALL>(in[0][w], q, KIND == PLEVEL ? 0.f : in[NIN - 1][w], c, r[w]) ? 0u : (1u << w);
      }
    }
    // The IEEE redo, one flagged point of EVERY lane per pass: a pass per slot (w = 0 .. 3) runs four times whenever each slot is
    // flagged in some lane of the warp -- always, when 15 % of the points are flagged (a defined t, q over an undefined p, which the
    // reference lets flow into RH and Td, FC.cc:1429) --, each time with a sixth of the lanes; a pass per "k-th flagged point of the
    // lane" runs max-over-lanes(popc(bad)) times, 2.3 on average for the same field.
    if constexpr (OUTS == O_ALL) {
      // (the four-output chain keeps the pass per slot: the other loop costs its all-defined path 8 % -- 150 -> 138 Gpt/s -- through
      // the register allocation of a kernel that sits at 127 registers)
      if (bad) {
#pragma unroll
        for (int w = 0; w < 4; ++w)
          if (bad & (1u << w))
            redo<ALL>(in[0][w], HAS_Q ? in[1][w] : 0.f, KIND == PLEVEL ? 0.f : in[NIN - 1][w], c, r[w]);
      }
      bad = 0;
    }
    while (bad) {
      const int w = __ffs(bad) - 1;
      bad &= bad - 1;
      const float tw = w == 0 ? in[0][0] : w == 1 ? in[0][1] : w == 2 ? in[0][2] : in[0][3];
      const float qw = !HAS_Q ? 0.f : w == 0 ? in[HAS_Q ? 1 : 0][0] : w == 1 ? in[HAS_Q ? 1 : 0][1] : w == 2 ? in[HAS_Q ? 1 : 0][2] : in[HAS_Q ? 1 : 0][3];
      const float pw = KIND == PLEVEL ? 0.f : w == 0 ? in[NIN - 1][0] : w == 1 ? in[NIN - 1][1] : w == 2 ? in[NIN - 1][2] : in[NIN - 1][3];
      Raw rr = w == 0 ? r[0] : w == 1 ? r[1] : w == 2 ? r[2] : r[3]; // (a humidity-only redo leaves theta and theta_e as they are)
      redo<ALL>(tw, qw, pw, c, rr);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (k == w)
          r[k] = rr;
    }
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      float o[NOUT];
      finish<ALL>(in[0][w], HAS_Q ? in[1][w] : 0.f, KIND == PLEVEL ? 0.f : in[NIN - 1][w], r[w], c, o, nundef);
#pragma unroll
      for (int k = 0; k < NOUT; ++k)
        out[k][w] = o[k];
    }
  }
};

// ------------------------------------------------------------------------------------ host drivers


// ---- pressure levels --------------------------------------------------------------------------------

int impl_pleveltemp(const Batch& b, const float* tinp, const float* p, const char* unit, int compute, float* tout, int* fDefined, float undef)
{ // FC.cc:328-367
  for (int k = 0; k < b.nfields; ++k)
    if (p[k] <= 0)
      return 0;
  compute = temp_compute(compute, unit);
  if (compute < 1 || compute > 5)
    return 0;
  const float* in[1] = {tinp};
  const int pf[1] = {1};
  if (compute <= 3) {
    PTempStreamOp op{compute};
    return run_elementwise(b, op, in, pf, tout, fDefined, undef, FLAG_UNCHANGED, [&](int k, FieldMeta& m) { m.a = host_pidcp(p[k]); });
  }
  auto fill = [&](int k, FieldMeta& m) {
    m.a = p[k];
    m.b = host_pidcp(p[k]) * H_CP;
  };
  if (compute == 4) // T -> theta_e,sat: one-output form of the fused chain's branch-free code (0.49 -> 0.61; 4 CTAs/SM beats 3 and 5)
    return run_elementwise(b, AlevelChainOpT<FCB_PL_U, FCB_PL_MB4, 2, O_THESAT, PLEVEL>{0.f}, in, pf, tout, fDefined, undef, FLAG_FROM_COUNT, fill);
  TempOp<PLEVEL> op{compute};
  return run_elementwise(b, op, in, pf, tout, fDefined, undef, FLAG_FROM_COUNT, fill);
}

int impl_plevelhum(const Batch& b, const float* t, const float* huminp, const float* p, const char* unit, int compute, float* humout, int* fDefined,
                   float undef)
{ // FC.cc:400-464
  if (compute <= 0 || compute >= 13)
    return 0;
  for (int k = 0; k < b.nfields; ++k)
    if (p[k] <= 0)
      return 0;
  compute = hum_compute(compute, unit);
  const bool rh_td = (compute == 5 || compute == 6 || compute == 9 || compute == 10);
  // p-level numbering -> a/h-level numbering (5,6,9,10 <-> 7,8,11,12; FCT.cc:73 documents the permutation)
  static const int to_ah[13] = {0, 1, 2, 3, 4, 7, 8, 5, 6, 11, 12, 9, 10};
  HumOp<PLEVEL> op{to_ah[compute], 0, (compute >= 9) ? H_T0 : 0.f};
  const bool even = (compute % 2 == 0);
  const float* in[2] = {t, huminp};
  const int pf[2] = {1, 1};
  auto fill = [&](int k, FieldMeta& m) {
    if (p[k] == undef && !rh_td)
      m.all |= 2; // fillUndef -> every point undefined -> NONE_DEFINED (FC.cc:429-432)
    m.a = p[k];
    m.b = 1.f;
  };
  // T, q -> RH and T, q -> Td: the fused chain's branch-free code with one output and the field's scalar pressure
  if (to_ah[compute] == 1)
    return run_elementwise(b, AlevelChainOpT<FCB_PL_U, FCB_PL_MB, 2, O_RH, PLEVEL>{0.f}, in, pf, humout, fDefined, undef, FLAG_FROM_COUNT, fill);
  if (to_ah[compute] == 5 || to_ah[compute] == 9)
    return run_elementwise(b, AlevelChainOpT<FCB_PL_U, FCB_PL_MB, 2, O_TD, PLEVEL>{(compute >= 9) ? H_T0 : 0.f}, in, pf, humout, fDefined, undef, FLAG_FROM_COUNT, fill);
  if (to_ah[compute] == 7 || to_ah[compute] == 11) // T, RH -> Td (p-level numbering 5 / 9): p is not used, not even when it is undefined
    return run_elementwise(b, AlevelChainOpT<FCB_PL_U, FCB_PL_MB, 2, O_TDRH, PLEVEL>{(compute >= 9) ? H_T0 : 0.f}, in, pf, humout, fDefined, undef, FLAG_FROM_COUNT, fill);
  return run_elementwise(b, op, in, pf, humout, fDefined, undef, FLAG_FROM_COUNT, [&](int k, FieldMeta& m) {
    if (p[k] == undef && !rh_td)
      m.all |= 2; // fillUndef -> every point undefined -> NONE_DEFINED (FC.cc:429-432)
    const float pi = H_CP * host_pidcp(p[k]);
    m.a = p[k];
    m.b = even ? (pi / H_CP) : 1.f;
  });
}

// ---- hybrid / atmospheric levels --------------------------------------------------------------------


template <int KIND>
int impl_xleveltemp(const Batch& b, const float* tinp, const float* pin, const float* alevel, const float* blevel, int compute, float* tout,
                    int* fDefined, float undef)
{ // FC.cc:1074-1097 / 1329-1352
  const int pf[3] = {1, KIND == ALEVEL ? 1 : 0, 1};
  auto levels = [&](int k, FieldMeta& m) {
    if (KIND == HLEVEL) {
      m.a = alevel[k];
      m.b = blevel[k];
    }
  };
  if (compute < 1 || compute > 5) { // only reachable for hleveltemp
    const float* in[3] = {tinp, pin, tout};
    KeepOrUndefOp<2> op;
    return run_elementwise(b, op, in, pf, tout, fDefined, undef, FLAG_FROM_COUNT, levels);
  }
  const float* in[2] = {tinp, pin};
  // T -> theta: the fused chain's branch-free code with one output.  Shapes measured on B200 (MEPS x 96, fraction of the HBM
  // roofline): 3 CTAs/SM 0.65, 4 CTAs/SM 0.63, 2 CTAs/SM 0.56; the generic TempOp 0.55.
  // (round 2, after the Exner factor moved to the special-function unit: a-level 0.89 with 3 CTAs/SM, 0.84 with 4; hybrid level -- one
  // input field less to wait for -- 0.65 with 3, 0.72 with 4; one float4 group per thread or 5 CTAs/SM change nothing)
  if (compute == 3 && KIND == HLEVEL)
    return run_elementwise(b, AlevelChainOpT<2, 4, 2, O_THETA, KIND>{0.f}, in, pf, tout, fDefined, undef, FLAG_FROM_COUNT, levels);
  if (compute == 3)
    return run_elementwise(b, AlevelChainOpT<2, 3, 2, O_THETA, KIND>{0.f}, in, pf, tout, fDefined, undef, FLAG_FROM_COUNT, levels);
  if (compute == 4)
    return run_elementwise(b, AlevelChainOpT<2, 3, 2, O_THESAT, KIND>{0.f}, in, pf, tout, fDefined, undef, FLAG_FROM_COUNT, levels);
  TempOp<KIND> op{compute};
  return run_elementwise(b, op, in, pf, tout, fDefined, undef, FLAG_FROM_COUNT, levels);
}

template <int KIND>
int impl_xlevelthe(const Batch& b, const float* t, const float* q, const float* pin, const float* alevel, const float* blevel, int compute,
                   float* the, int* fDefined, float undef)
{ // FC.cc:1125-1142 / 1375-1391
  const int pf[4] = {1, 1, KIND == ALEVEL ? 1 : 0, 1};
  auto levels = [&](int k, FieldMeta& m) {
    if (KIND == HLEVEL) {
      m.a = alevel[k];
      m.b = blevel[k];
    }
  };
  if (compute != 1 && compute != 2) { // only reachable for hlevelthe
    const float* in[4] = {t, q, pin, the};
    KeepOrUndefOp<3> op;
    return run_elementwise(b, op, in, pf, the, fDefined, undef, FLAG_FROM_COUNT, levels);
  }
  const float* in[3] = {t, q, pin};
  // compute 1 (T, q -> theta_e): the one-output form of the fused chain's branch-free code (reciprocal Exner factor from the
  // special-function unit, two multiplications); compute 2 (theta, q -> theta_e) keeps the generic operator
  if (compute == 1)
    return run_elementwise(b, AlevelChainOpT<2, 3, 2, O_THE, KIND>{0.f}, in, pf, the, fDefined, undef, FLAG_FROM_COUNT, levels);
  TheOp<KIND> op{compute};
  return run_elementwise(b, op, in, pf, the, fDefined, undef, FLAG_FROM_COUNT, levels);
}

template <int KIND>
int impl_xlevelhum(const Batch& b, const float* t, const float* huminp, const float* pin, const float* alevel, const float* blevel,
                   const char* unit, int compute, float* humout, int* fDefined, float undef)
{ // FC.cc:1145-1217 / 1394-1458
  compute = hum_compute(compute, unit);
  const bool no_p = (compute == 7 || compute == 11);
  HumOp<KIND> op{compute, (KIND == HLEVEL) ? !no_p : no_p, (compute >= 9) ? H_T0 : 0.f};
  const float* in[3] = {t, huminp, pin};
  const int pf[3] = {1, 1, KIND == ALEVEL ? 1 : 0};
  // T, q -> RH and T, q -> Td: one-output forms of the fused chain (0.77 / 0.69 of the roofline at 3 CTAs/SM against 0.62 / 0.47 for
  // the generic HumOp; 2 CTAs/SM 0.65 / 0.62, 4 CTAs/SM 0.73 / 0.66).  Price: a field whose p is undefined where t and q are
  // defined sends those points through the IEEE redo (p flows into the arithmetic, FC.cc:1429): 0.31 / 0.24 instead of 0.50 / 0.40
  // with 30 % of t, q, p independently undefined.
  auto levels = [&](int k, FieldMeta& m) {
    if (KIND == HLEVEL) {
      m.a = alevel[k];
      m.b = blevel[k];
    }
  };
  if (compute == 1)
    return run_elementwise(b, AlevelChainOpT<2, 3, 2, O_RH, KIND>{0.f}, in, pf, humout, fDefined, undef, FLAG_FROM_COUNT, levels);
  if (compute == 5 || compute == 9)
    return run_elementwise(b, AlevelChainOpT<2, 3, 2, O_TD, KIND>{(compute >= 9) ? H_T0 : 0.f}, in, pf, humout, fDefined, undef, FLAG_FROM_COUNT, levels);
  if (compute == 7 || compute == 11) // T, RH -> Td
    return run_elementwise(b, AlevelChainOpT<2, 3, 2, O_TDRH, KIND>{(compute >= 9) ? H_T0 : 0.f}, in, pf, humout, fDefined, undef, FLAG_FROM_COUNT, levels);
  return run_elementwise(b, op, in, pf, humout, fDefined, undef, FLAG_FROM_COUNT, levels);
}

template <int KIND>
int impl_xlevelducting(const Batch& b, const float* t, const float* h, const float* pin, const float* alevel, const float* blevel, int compute,
                       float* duct, int* fDefined, float undef)
{ // FC.cc:1252-1273 / 1485-1504 (the a-level variant never updates the flag)
  const FlagRule rule = (KIND == HLEVEL) ? FLAG_FROM_COUNT : FLAG_UNCHANGED;
  const int pf[4] = {1, 1, KIND == ALEVEL ? 1 : 0, 1};
  auto levels = [&](int k, FieldMeta& m) {
    if (KIND == HLEVEL) {
      m.a = alevel[k];
      m.b = blevel[k];
    }
  };
  if (compute < 1 || compute > 4) {
    const float* in[4] = {t, h, pin, duct};
    KeepOrUndefOp<3> op;
    return run_elementwise(b, op, in, pf, duct, fDefined, undef, rule, levels);
  }
  const float* in[3] = {t, h, pin};
  DuctOp<KIND> op{compute};
  return run_elementwise(b, op, in, pf, duct, fDefined, undef, rule, levels);
}

bool any_bad_hlevel(const Batch& b, const float* alevel, const float* blevel)
{
  for (int k = 0; k < b.nfields; ++k)
    if (bad_hlevel(alevel[k], blevel[k]))
      return true;
  return false;
}

} // namespace
} // namespace fcb200

// =========================================================================================== C-ABI
using namespace fcb200;

extern "C" {

// ---- pressure levels
int fcb200_pleveltemp(int nx, int ny, const float* tinp, float p, const char* unit, int compute, float* tout, int* fDefined, float undef)
{
  return impl_pleveltemp(make_batch(nx, ny, 1), tinp, &p, unit, compute, tout, fDefined, undef);
}
int fcb200_pleveltemp_batched(int nx, int ny, int nfields, const float* tinp, const float* p, const char* unit, int compute, float* tout,
                              int* fDefined, float undef)
{
  return impl_pleveltemp(make_batch(nx, ny, nfields), tinp, p, unit, compute, tout, fDefined, undef);
}

int fcb200_plevelhum(int nx, int ny, const float* t, const float* huminp, float p, const char* unit, int compute, float* humout, int* fDefined,
                     float undef)
{
  return impl_plevelhum(make_batch(nx, ny, 1), t, huminp, &p, unit, compute, humout, fDefined, undef);
}
int fcb200_plevelhum_batched(int nx, int ny, int nfields, const float* t, const float* huminp, const float* p, const char* unit, int compute,
                             float* humout, int* fDefined, float undef)
{
  return impl_plevelhum(make_batch(nx, ny, nfields), t, huminp, p, unit, compute, humout, fDefined, undef);
}

// ---- hybrid levels
int fcb200_hleveltemp_batched(int nx, int ny, int nfields, const float* tinp, const float* ps, const float* alevel, const float* blevel,
                              const char* unit, int compute, float* tout, int* fDefined, float undef)
{ // FC.cc:1046-1098: unit remap first, then bad_hlevel
  const Batch b = make_batch(nx, ny, nfields);
  compute = temp_compute(compute, unit);
  if (any_bad_hlevel(b, alevel, blevel))
    return 0;
  return impl_xleveltemp<HLEVEL>(b, tinp, ps, alevel, blevel, compute, tout, fDefined, undef);
}
int fcb200_hleveltemp(int nx, int ny, const float* tinp, const float* ps, float alevel, float blevel, const char* unit, int compute, float* tout,
                      int* fDefined, float undef)
{
  return fcb200_hleveltemp_batched(nx, ny, 1, tinp, ps, &alevel, &blevel, unit, compute, tout, fDefined, undef);
}

int fcb200_hlevelthe_batched(int nx, int ny, int nfields, const float* t, const float* q, const float* ps, const float* alevel,
                             const float* blevel, int compute, float* the, int* fDefined, float undef)
{ // FC.cc:1100-1143
  const Batch b = make_batch(nx, ny, nfields);
  if (any_bad_hlevel(b, alevel, blevel))
    return 0;
  return impl_xlevelthe<HLEVEL>(b, t, q, ps, alevel, blevel, compute, the, fDefined, undef);
}
int fcb200_hlevelthe(int nx, int ny, const float* t, const float* q, const float* ps, float alevel, float blevel, int compute, float* the,
                     int* fDefined, float undef)
{
  return fcb200_hlevelthe_batched(nx, ny, 1, t, q, ps, &alevel, &blevel, compute, the, fDefined, undef);
}

int fcb200_hlevelhum_batched(int nx, int ny, int nfields, const float* t, const float* huminp, const float* ps, const float* alevel,
                             const float* blevel, const char* unit, int compute, float* humout, int* fDefined, float undef)
{ // FC.cc:1145-1217
  if (compute <= 0 || compute >= 13)
    return 0;
  const Batch b = make_batch(nx, ny, nfields);
  if (any_bad_hlevel(b, alevel, blevel))
    return 0;
  return impl_xlevelhum<HLEVEL>(b, t, huminp, ps, alevel, blevel, unit, compute, humout, fDefined, undef);
}
int fcb200_hlevelhum(int nx, int ny, const float* t, const float* huminp, const float* ps, float alevel, float blevel, const char* unit,
                     int compute, float* humout, int* fDefined, float undef)
{
  return fcb200_hlevelhum_batched(nx, ny, 1, t, huminp, ps, &alevel, &blevel, unit, compute, humout, fDefined, undef);
}

int fcb200_hlevelducting_batched(int nx, int ny, int nfields, const float* t, const float* h, const float* ps, const float* alevel,
                                 const float* blevel, int compute, float* duct, int* fDefined, float undef)
{ // FC.cc:1219-1274
  const Batch b = make_batch(nx, ny, nfields);
  if (any_bad_hlevel(b, alevel, blevel))
    return 0;
  return impl_xlevelducting<HLEVEL>(b, t, h, ps, alevel, blevel, compute, duct, fDefined, undef);
}
int fcb200_hlevelducting(int nx, int ny, const float* t, const float* h, const float* ps, float alevel, float blevel, int compute, float* duct,
                         int* fDefined, float undef)
{
  return fcb200_hlevelducting_batched(nx, ny, 1, t, h, ps, &alevel, &blevel, compute, duct, fDefined, undef);
}

int fcb200_hlevelpressure_batched(int nx, int ny, int nfields, const float* ps, const float* alevel, const float* blevel, float* p, int* fDefined,
                                  float undef)
{ // FC.cc:1276-1304
  const Batch b = make_batch(nx, ny, nfields);
  if (any_bad_hlevel(b, alevel, blevel))
    return 0;
  const float* in[1] = {ps};
  const int pf[1] = {0};
  HPressureOp op;
  return run_elementwise(b, op, in, pf, p, fDefined, undef, FLAG_FROM_COUNT, [&](int k, FieldMeta& m) {
    m.a = alevel[k];
    m.b = blevel[k];
  });
}
int fcb200_hlevelpressure(int nx, int ny, const float* ps, float alevel, float blevel, float* p, int* fDefined, float undef)
{
  return fcb200_hlevelpressure_batched(nx, ny, 1, ps, &alevel, &blevel, p, fDefined, undef);
}

// ---- atmospheric levels
int fcb200_aleveltemp_batched(int nx, int ny, int nfields, const float* tinp, const float* p, const char* unit, int compute, float* tout,
                              int* fDefined, float undef)
{ // FC.cc:1310-1353: range check first, then unit remap
  if (compute <= 0 || compute >= 6)
    return 0;
  compute = temp_compute(compute, unit);
  return impl_xleveltemp<ALEVEL>(make_batch(nx, ny, nfields), tinp, p, nullptr, nullptr, compute, tout, fDefined, undef);
}
int fcb200_aleveltemp(int nx, int ny, const float* tinp, const float* p, const char* unit, int compute, float* tout, int* fDefined, float undef)
{
  return fcb200_aleveltemp_batched(nx, ny, 1, tinp, p, unit, compute, tout, fDefined, undef);
}

int fcb200_alevelthe_batched(int nx, int ny, int nfields, const float* t, const float* q, const float* p, int compute, float* the, int* fDefined,
                             float undef)
{ // FC.cc:1355-1392
  if (compute != 1 && compute != 2)
    return 0;
  return impl_xlevelthe<ALEVEL>(make_batch(nx, ny, nfields), t, q, p, nullptr, nullptr, compute, the, fDefined, undef);
}
int fcb200_alevelthe(int nx, int ny, const float* t, const float* q, const float* p, int compute, float* the, int* fDefined, float undef)
{
  return fcb200_alevelthe_batched(nx, ny, 1, t, q, p, compute, the, fDefined, undef);
}

int fcb200_alevelhum_batched(int nx, int ny, int nfields, const float* t, const float* huminp, const float* p, const char* unit, int compute,
                             float* humout, int* fDefined, float undef)
{ // FC.cc:1394-1458
  if (compute <= 0 || compute >= 13)
    return 0;
  return impl_xlevelhum<ALEVEL>(make_batch(nx, ny, nfields), t, huminp, p, nullptr, nullptr, unit, compute, humout, fDefined, undef);
}
int fcb200_alevelhum(int nx, int ny, const float* t, const float* huminp, const float* p, const char* unit, int compute, float* humout,
                     int* fDefined, float undef)
{
  return fcb200_alevelhum_batched(nx, ny, 1, t, huminp, p, unit, compute, humout, fDefined, undef);
}

int fcb200_alevelducting_batched(int nx, int ny, int nfields, const float* t, const float* h, const float* p, int compute, float* duct,
                                 int* fDefined, float undef)
{ // FC.cc:1460-1505
  return impl_xlevelducting<ALEVEL>(make_batch(nx, ny, nfields), t, h, p, nullptr, nullptr, compute, duct, fDefined, undef);
}
int fcb200_alevelducting(int nx, int ny, const float* t, const float* h, const float* p, int compute, float* duct, int* fDefined, float undef)
{
  return fcb200_alevelducting_batched(nx, ny, 1, t, h, p, compute, duct, fDefined, undef);
}

// ---- level independent
int fcb200_windCooling_batched(int nx, int ny, int nfields, const float* t, const float* u, const float* v, int compute, float* dtcool,
                               int* fDefined, float undef)
{ // FC.cc:2181-2229
  if (compute != 1 && compute != 2)
    return 0;
  WindCoolingOp op{(compute == 1) ? H_T0 : 0.f};
  const float* in[3] = {t, u, v};
  const int pf[3] = {1, 1, 1};
  return run_elementwise(make_batch(nx, ny, nfields), op, in, pf, dtcool, fDefined, undef, FLAG_UNCHANGED, NoMeta());
}
int fcb200_windCooling(int nx, int ny, const float* t, const float* u, const float* v, int compute, float* dtcool, int* fDefined, float undef)
{
  return fcb200_windCooling_batched(nx, ny, 1, t, u, v, compute, dtcool, fDefined, undef);
}

int fcb200_fieldOPERfield_batched(int compute, int nx, int ny, int nfields, const float* field1, const float* field2, float* fres, int* fDefined,
                                  float undef)
{ // FC.cc:2611-2625: only the division recomputes the flag
  if (compute < 1 || compute > 4)
    return 0;
  FieldOperOp op{compute};
  const float* in[2] = {field1, field2};
  const int pf[2] = {1, 1};
  return run_elementwise(make_batch(nx, ny, nfields), op, in, pf, fres, fDefined, undef, compute == 4 ? FLAG_FROM_COUNT : FLAG_UNCHANGED, NoMeta());
}
int fcb200_fieldOPERfield(int compute, int nx, int ny, const float* field1, const float* field2, float* fres, int* fDefined, float undef)
{
  return fcb200_fieldOPERfield_batched(compute, nx, ny, 1, field1, field2, fres, fDefined, undef);
}

int fcb200_momentumXcoordinate_batched(int nx, int ny, int nfields, const float* v, const float* xmapr, const float* fcoriolis, float fcoriolisMin,
                                       float* mxy, int* fDefined, float undef)
{ // FC.cc:2351-2386
  if (nx < 3 || ny < 3)
    return 0;
  MomentumOp<true> op{fabsf(fcoriolisMin)};
  const float* in[3] = {v, xmapr, fcoriolis};
  const int pf[3] = {1, 0, 0};
  return run_elementwise(make_batch(nx, ny, nfields), op, in, pf, mxy, fDefined, undef, FLAG_FROM_COUNT, NoMeta());
}
int fcb200_momentumXcoordinate(int nx, int ny, const float* v, const float* xmapr, const float* fcoriolis, float fcoriolisMin, float* mxy,
                               int* fDefined, float undef)
{
  return fcb200_momentumXcoordinate_batched(nx, ny, 1, v, xmapr, fcoriolis, fcoriolisMin, mxy, fDefined, undef);
}

int fcb200_momentumYcoordinate_batched(int nx, int ny, int nfields, const float* u, const float* ymapr, const float* fcoriolis, float fcoriolisMin,
                                       float* nxy, int* fDefined, float undef)
{ // FC.cc:2388-2422
  if (nx < 3 || ny < 3)
    return 0;
  MomentumOp<false> op{fabsf(fcoriolisMin)};
  const float* in[3] = {u, ymapr, fcoriolis};
  const int pf[3] = {1, 0, 0};
  return run_elementwise(make_batch(nx, ny, nfields), op, in, pf, nxy, fDefined, undef, FLAG_FROM_COUNT, NoMeta());
}
int fcb200_momentumYcoordinate(int nx, int ny, const float* u, const float* ymapr, const float* fcoriolis, float fcoriolisMin, float* nxy,
                               int* fDefined, float undef)
{
  return fcb200_momentumYcoordinate_batched(nx, ny, 1, u, ymapr, fcoriolis, fcoriolisMin, nxy, fDefined, undef);
}

int fcb200_alevel_chain_batched(int nx, int ny, int nfields, const float* t, const float* q, const float* p, const char* td_unit, float* theta,
                                float* rh, float* td, float* thetae, const int* fDefinedIn, int* fDefinedOut, float undef)
{ // = aleveltemp(c=3) + alevelhum(c=1) + alevelhum(c=5, unit) + alevelthe(c=1) on the same inputs
  const int td_compute = hum_compute(5, td_unit); // 5 (Celsius) or 9 (Kelvin), FC.cc:1417-1420
  const float tdconv = (td_compute >= 9) ? H_T0 : 0.f;
  auto run = [&](auto op) {
    EwJob<decltype(op)> job;
    job.nx = nx;
    job.ny = ny;
    job.nfields = nfields;
    job.in[0] = t;
    job.in[1] = q;
    job.in[2] = p;
    for (int k = 0; k < 3; ++k)
      job.per_field[k] = true;
    job.out[0] = theta;
    job.out[1] = rh;
    job.out[2] = td;
    job.out[3] = thetae;
    job.flags_in = fDefinedIn;
    for (int o = 0; o < 4; ++o)
      job.flags_out[o] = fDefinedOut + (size_t)o * (nfields > 0 ? nfields : 0);
    job.undef = undef;
    return run_ew_job(op, job);
  };
  // Kernel shape (measured on B200): one float4 group per thread and round, 4 rounds per item, 3 CTAs of 256 threads per SM
  // (80 registers, 84 bytes of spill).  With the FP64 Exner power of round 1 the straight-line code of 8 interleaved points
  // needed 127 registers (2 CTAs/SM: profiles/r01_chain_tuning.txt); with the reciprocal Exner factor from the special-function
  // unit the kernel is 113 instructions per point and latency-bound at 2 CTAs/SM (issue 58 %, DRAM 59 %): 168 Gpt/s there,
  // 179 Gpt/s with 3 (profiles/r02aq_chain_variants.txt; 4 CTAs at 64 registers 166, the packed form 177).
  // Fields with undefined points take another shape: the definedness tests, the `live` logic and four counters cost the masked
  // instantiation the registers the 3-CTA shape does not have -- 2 CTAs/SM at 126 registers, two groups per thread, packed FP32:
  // 0.63 of the roofline with 30 % undefined against 0.45 for the all-defined shape (which gets 0.72 / 0.77 from the two shapes).
  bool all = true;
  for (int k = 0; k < nfields; ++k)
    all = all && fDefinedIn[k] == ALL_DEFINED;
  if (!all)
    return run(AlevelChainOpT<2, 2, 4, O_ALL, ALEVEL, 1>{tdconv});
  return run(AlevelChainOpT<1, 3, 4, O_ALL, ALEVEL, 0>{tdconv});
}

int fcb200_hlevel_chain_batched(int nx, int ny, int nfields, const float* t, const float* q, const float* ps, const float* alevel, const float* blevel,
                                const char* td_unit, float* theta, float* rh, float* td, float* thetae, const int* fDefinedIn, int* fDefinedOut,
                                float undef)
{ // = hleveltemp(c=3) + hlevelhum(c=1) + hlevelhum(c=5, unit) + hlevelthe(c=1) on the same inputs; p = alevel + blevel * ps with ONE surface
  // pressure field for the batch (the MEPS layout: 65 hybrid levels above one ps): 24 bytes per point instead of the 28 of the a-level
  // chain, which needs a pressure field per level
  if (nfields > 0 && any_bad_hlevel(make_batch(nx, ny, nfields), alevel, blevel))
    return 0; // FC.cc:1070, 1121, 1170
  const int td_compute = hum_compute(5, td_unit);
  auto run = [&](auto op) {
    EwJob<decltype(op)> job;
    job.nx = nx;
    job.ny = ny;
    job.nfields = nfields;
    job.in[0] = t;
    job.in[1] = q;
    job.in[2] = ps;
    job.per_field[0] = job.per_field[1] = true;
    job.per_field[2] = false;
    job.out[0] = theta;
    job.out[1] = rh;
    job.out[2] = td;
    job.out[3] = thetae;
    job.flags_in = fDefinedIn;
    for (int o = 0; o < 4; ++o)
      job.flags_out[o] = fDefinedOut + (size_t)o * (nfields > 0 ? nfields : 0);
    job.undef = undef;
    job.fill_meta = [&](int k, FieldMeta& m) {
      m.a = alevel[k];
      m.b = blevel[k];
    };
    return run_ew_job(op, job);
  };
  const float tdconv = (td_compute >= 9) ? H_T0 : 0.f;
  bool all = true; // (the shapes: see fcb200_alevel_chain_batched)
  for (int k = 0; k < nfields; ++k)
    all = all && fDefinedIn[k] == ALL_DEFINED;
  if (!all)
    return run(AlevelChainOpT<2, 2, 4, O_ALL, HLEVEL, 1>{tdconv});
  return run(AlevelChainOpT<1, 3, 4, O_ALL, HLEVEL, 0>{tdconv});
}

} // extern "C"

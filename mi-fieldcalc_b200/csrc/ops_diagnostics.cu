// ops_diagnostics.cu -- the fixed-level stability indices and the level-independent conversions that
// complete the reference's Python subset (python/py_mi_fieldcalc.cc:189-207):
// kIndex, ductingIndex, showalterIndex, boydenIndex, sweatIndex, seaSoundSpeed, cvtemp, cvhum, abshum,
// underCooledRain.  All of them are point-wise: functors for the batched engine in elementwise.cuh, same
// rules as ops_elementwise.cu (reference expression types, no FMA contraction, IEEE division and sqrt).
// Scalar level pressures are turned into Exner factors on the HOST with glibc powf, exactly like the
// reference does before its loop, so none of these operators carries a device-powf ulp.
// Citations: FC.cc = the reference's src/mi_fieldcalc/FieldCalculations.cc, MC.h = MetConstants.h.
#include "ew_host.cuh"

#include "../../include/fcb200.h"

namespace fcb200 {
namespace {

using dev::is_def;
using dev::K_CP;
using dev::K_T0;
using dev::K_XLH;

// MC.h:42: rcp = r / cp, cplr = xlh / rcp, exl = eps * xlh (float constant expressions)
constexpr float K_RCP = dev::K_R / dev::K_CP;
constexpr float K_CPLR = dev::K_XLH / K_RCP;
constexpr float K_EXL = dev::K_EPS * dev::K_XLH;

__device__ __forceinline__ float rh_fraction(float rh100)
{ // clamp_rh(0.01 * rh): the product is double, the clamp's argument float (FC.cc:186, 788)
  return dev::clamp_rh((float)(0.01 * (double)rh100));
}

// the same without a branch (selects): a NaN stays a NaN, exactly like clamp_rh
__device__ __forceinline__ float rh_fraction_select(float rh100)
{
  const float rh = (float)(0.01 * (double)rh100);
  return rh < dev::K_RHMIN ? dev::K_RHMIN : (rh > dev::K_RHMAX ? dev::K_RHMAX : rh);
}

__device__ __forceinline__ float ms2knots(float ff)
{ // MC.h:53, 132-135
  return (float)((double)ff * (3600.0 / 1852.0));
}

// kIndex, FC.cc:745-814.  Same construction as CvHumOp: branch-free inside the saturation table (both levels), redo outside.
#ifndef FCB_KI_U
#define FCB_KI_U 2
#define FCB_KI_MB 2
#endif
struct KIndexOp
{
  static constexpr int NIN = 5, NOUT = 1, UNROLL = FCB_KI_U;
  static constexpr int NCOUNT = 1;
  static constexpr int MIN_BLOCKS = FCB_KI_MB;
  static constexpr bool HEAVY = false;
  static constexpr bool USES_EWT = true, USES_POW = false;
  static constexpr bool QUAD = true;
  float cvt500, cvt700, cvt850;

  __device__ __noinline__ static bool exact(float t500, float t700, float rh700, float t850, float rh850, float cvt500, float cvt700, float cvt850,
                                            const dev::EwtTable& tab, float& r)
  {
    const float rh8 = rh_fraction(rh850);
    const float tc850 = cvt850 * t850 - K_T0;
    const float tc700 = cvt700 * t700 - K_T0;
    const dev::Ewt e850(tc850), e700(tc700);
    if (!(e850.defined && e700.defined))
      return false;
    const float tdc850 = e850.inverse(tab, e850.value(tab) * rh8);
    const float rh7 = rh_fraction(rh700);
    const float tdc700 = e700.inverse(tab, e700.value(tab) * rh7);
    const float tc500 = cvt500 * t500 - K_T0;
    r = (tc850 + tdc850) - (tc700 - tdc700) - tc500;
    return true;
  }

  __device__ __forceinline__ bool fast(float t500, float t700, float rh700, float t850, float rh850, const dev::EwtTable& tab, float& r) const
  {
    const float tc850 = cvt850 * t850 - K_T0;
    const float tc700 = cvt700 * t700 - K_T0;
    const dev::EwtFast e850(tab, tc850), e700(tab, tc700);
    const float tdc850 = e850.dewpoint(tab, e850.et * rh_fraction_select(rh850));
    const float tdc700 = e700.dewpoint(tab, e700.et * rh_fraction_select(rh700));
    const float tc500 = cvt500 * t500 - K_T0;
    r = (tc850 + tdc850) - (tc700 - tdc700) - tc500;
    return e850.plausible && e700.plausible;
  }

  template <bool ALL>
  __device__ __forceinline__ bool defined(float a, float b, float c2, float d, float e, float undef) const
  {
    return ALL || (is_def(a, undef) && is_def(b, undef) && is_def(c2, undef) && is_def(d, undef) && is_def(e, undef));
  }

  template <bool ALL>
  __device__ __forceinline__ void point(const float* in, float* out, const PointCtx& c, long long, unsigned* nundef) const
  {
    const bool def = defined<ALL>(in[0], in[1], in[2], in[3], in[4], c.undef);
    float r;
    bool ok = fast(in[0], in[1], in[2], in[3], in[4], c.tab, r);
    if (!ok && def)
      ok = exact(in[0], in[1], in[2], in[3], in[4], cvt500, cvt700, cvt850, c.tab, r);
    const bool good = def && ok;
    out[0] = good ? r : c.undef;
    nundef[0] += good ? 0u : 1u;
  }

  template <bool ALL>
  __device__ __forceinline__ void quad(const float (*in)[4], float (*out)[4], const PointCtx& c, unsigned* nundef) const
  {
    float r[4];
    unsigned okm = 0, defm = 0;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      defm |= defined<ALL>(in[0][w], in[1][w], in[2][w], in[3][w], in[4][w], c.undef) ? (1u << w) : 0u;
      okm |= fast(in[0][w], in[1][w], in[2][w], in[3][w], in[4][w], c.tab, r[w]) ? (1u << w) : 0u;
    }
    const unsigned redo = defm & ~okm;
    if (redo) {
#pragma unroll
      for (int w = 0; w < 4; ++w)
        if (redo & (1u << w)) {
          float v = 0.f;
          if (exact(in[0][w], in[1][w], in[2][w], in[3][w], in[4][w], cvt500, cvt700, cvt850, c.tab, v)) {
            okm |= 1u << w;
            r[w] = v;
          }
        }
    }
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const bool good = ((defm & okm) >> w) & 1u;
      out[0][w] = good ? r[w] : c.undef;
      nundef[0] += good ? 0u : 1u;
    }
  }
};

// ductingIndex, FC.cc:816-870.  Same construction as CvHumOp: branch-free inside the saturation table, redo outside.
#ifndef FCB_DI_U
#define FCB_DI_U 2
#define FCB_DI_MB 4
#endif
struct DuctingIndexOp
{
  static constexpr int NIN = 2, NOUT = 1, UNROLL = FCB_DI_U;
  static constexpr int NCOUNT = 1;
  static constexpr int MIN_BLOCKS = FCB_DI_MB;
  static constexpr bool HEAVY = false;
  static constexpr bool USES_EWT = true, USES_POW = false;
  static constexpr bool QUAD = true;
  float tconvert;

  __device__ __noinline__ static bool exact(float t, float h, float tconvert, const dev::EwtTable& tab, float& r)
  {
    const float bduct = (float)3.8e+5;
    const float rh = rh_fraction(h);
    const float tk = t * tconvert;
    const dev::Ewt e(tk - K_T0);
    if (!e.defined)
      return false;
    const float et = e.value(tab);
    const float etd = et * rh;
    const float tdk = e.inverse(tab, etd) + K_T0;
    r = bduct * (et / (tk * tk) - etd / (tdk * tdk));
    return true;
  }

  __device__ __forceinline__ bool fast(float t, float h, const dev::EwtTable& tab, float& r) const
  {
    const float bduct = (float)3.8e+5;
    const float rh = rh_fraction_select(h);
    const float tk = t * tconvert;
    const dev::EwtFast e(tab, tk - K_T0);
    const float etd = e.et * rh;
    const float tdk = e.dewpoint(tab, etd) + K_T0; // in [168, 373] for a plausible position: both squares are mid-range divisors
    r = bduct * (dev::div_midrange(e.et, tk * tk) - dev::div_midrange(etd, tdk * tdk));
    return e.plausible;
  }

  template <bool ALL>
  __device__ __forceinline__ void point(const float* in, float* out, const PointCtx& c, long long, unsigned* nundef) const
  {
    const bool def = ALL || (is_def(in[0], c.undef) && is_def(in[1], c.undef));
    float r;
    bool ok = fast(in[0], in[1], c.tab, r);
    if (!ok && def)
      ok = exact(in[0], in[1], tconvert, c.tab, r);
    const bool good = def && ok;
    out[0] = good ? r : c.undef;
    nundef[0] += good ? 0u : 1u;
  }

  template <bool ALL>
  __device__ __forceinline__ void quad(const float (*in)[4], float (*out)[4], const PointCtx& c, unsigned* nundef) const
  {
    float r[4];
    unsigned okm = 0, defm = 0;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const bool def = ALL || (is_def(in[0][w], c.undef) && is_def(in[1][w], c.undef));
      defm |= def ? (1u << w) : 0u;
      okm |= fast(in[0][w], in[1][w], c.tab, r[w]) ? (1u << w) : 0u;
    }
    const unsigned redo = defm & ~okm;
    if (redo) {
#pragma unroll
      for (int w = 0; w < 4; ++w)
        if (redo & (1u << w)) {
          float v = 0.f;
          if (exact(in[0][w], in[1][w], tconvert, c.tab, v)) {
            okm |= 1u << w;
            r[w] = v;
          }
        }
    }
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const bool good = ((defm & okm) >> w) & 1u;
      out[0][w] = good ? r[w] : c.undef;
      nundef[0] += good ? 0u : 1u;
    }
  }
};

// showalterIndex, FC.cc:872-971.  in[3] is the OUTPUT field itself: an undefined input point is counted
// but its output point is not written (:966-968), so the old value is carried through.
struct ShowalterOp
{
  static constexpr int NIN = 4, NOUT = 1, UNROLL = 1;
  static constexpr int NCOUNT = 1;
  static constexpr int MIN_BLOCKS = 3;
  static constexpr bool HEAVY = false;
  static constexpr bool USES_EWT = true, USES_POW = false;
  float cvt500, cvt850, dryadiabat, p500, p850;
  template <bool ALL>
  __device__ __forceinline__ void point(const float* in, float* out, const PointCtx& c, long long, unsigned* nundef) const
  {
    const float t500 = in[0], t850 = in[1], rh850 = in[2];
    if (!(ALL || (is_def(t500, c.undef) && is_def(t850, c.undef) && is_def(rh850, c.undef)))) {
      out[0] = in[3];
      nundef[0] += 1;
      return;
    }
    const float tk500 = cvt500 * t500;
    const float tk850 = cvt850 * t850;
    const float rh = rh_fraction(rh850);
    const dev::Ewt e(tk850 - K_T0);
    if (!e.defined) {
      out[0] = c.undef;
      nundef[0] += 1;
      return;
    }
    const float etd = e.value(c.tab) * rh;
    // moist adiabat: lift along the dry adiabat, then adjust humidity and heat in 7 iterations (:938-960)
    float tcl = dryadiabat * t850;
    float qcl = dev::K_EPS * etd / p850;
    bool live = true;
#pragma unroll 1
    for (int it = 0; it < 7; ++it) {
      const dev::Ewt e2(tcl / K_CP - K_T0);
      live = live && e2.defined; // `break` of the reference: nothing changes once a lookup was out of range
      if (live) {
        const float esat = e2.value(c.tab);
        const float qsat = dev::K_EPS * esat / p500;
        float dq = qcl - qsat;
        const float a1 = K_CPLR * qcl / tcl;
        const float a2 = K_EXL / tcl;
        dq = (float)((double)dq / (1. + (double)(a1 * a2)));
        qcl = qcl - dq;
        tcl = tcl + dq * K_XLH;
      }
    }
    out[0] = tk500 - tcl / K_CP;
  }
};

// boydenIndex, FC.cc:973-1014
struct BoydenOp
{
  static constexpr int NIN = 3, NOUT = 1, UNROLL = 2;
  static constexpr int NCOUNT = 1;
  static constexpr int MIN_BLOCKS = 4;
  static constexpr bool HEAVY = false;
  static constexpr bool USES_EWT = false, USES_POW = false;
  float tconv;
  template <bool ALL>
  __device__ __forceinline__ void point(const float* in, float* out, const PointCtx& c, long long, unsigned* nundef) const
  {
    if (ALL || (is_def(in[0], c.undef) && is_def(in[1], c.undef) && is_def(in[2], c.undef))) {
      const float tc700 = in[0] * tconv - K_T0;
      out[0] = (float)((double)(in[1] - in[2]) / 10. - (double)tc700 - 200.);
    } else {
      out[0] = c.undef;
      nundef[0] += 1;
    }
  }
};

// sweatIndex, FC.cc:1016-1040: the float terms are summed left to right in float, the last term is double
struct SweatOp
{
  static constexpr int NIN = 8, NOUT = 1, UNROLL = 1;
  static constexpr int NCOUNT = 1;
  static constexpr int MIN_BLOCKS = 4;
  static constexpr bool HEAVY = false;
  static constexpr bool USES_EWT = false, USES_POW = false;
  template <bool ALL>
  __device__ __forceinline__ void point(const float* in, float* out, const PointCtx& c, long long, unsigned* nundef) const
  {
    bool ok = true;
    if (!ALL) {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        ok = ok && is_def(in[k], c.undef);
    }
    if (ok) {
      const float t850 = in[0], t500 = in[1], td850 = in[2], u850 = in[4], v850 = in[5], u500 = in[6], v500 = in[7];
      const float ff850 = dev::absval(u850, v850);
      const float ff500 = dev::absval(u500, v500);
      const float sind = (u500 * v850 - v500 * u850) / (ff850 * ff500);
      const float lhs = 32.f * td850 + 20.f * t850 - 40.f * t500 - 980.f + 2.f * ms2knots(ff850) + ms2knots(ff500);
      out[0] = (float)((double)lhs + 125. * ((double)sind + 0.2));
    } else {
      out[0] = c.undef;
      nundef[0] += 1;
    }
  }
};

// seaSoundSpeed, FC.cc:1555-1602 (Ross 1978): double polynomials of the float temperature and salinity
#ifndef FCB_SS_U
#define FCB_SS_U 2
#define FCB_SS_MB 5 // (0.74 -> 0.80 of the roofline against 4; profiles/r02ay_shape_variants.txt)
#endif
struct SeaSoundOp
{
  static constexpr int NIN = 2, NOUT = 1, UNROLL = FCB_SS_U;
  static constexpr int NCOUNT = 1;
  static constexpr int MIN_BLOCKS = FCB_SS_MB;
  static constexpr bool HEAVY = false;
  static constexpr bool USES_EWT = false, USES_POW = false;
  float tconv;
  double Cz;
  template <bool ALL>
  __device__ __forceinline__ void point(const float* in, float* out, const PointCtx& c, long long, unsigned* nundef) const
  {
    if (ALL || (is_def(in[0], c.undef) && is_def(in[1], c.undef))) {
      const double T = (double)(in[0] - tconv);
      const double S = (double)in[1];
      const double Ct = 4.565 * T - 0.0517 * T * T + 0.000221 * T * T * T;
      const double Cs = (1.338 - 0.013 * T + 0.0001 * T * T) * (S - 35.0);
      out[0] = (float)(1449.1 + Ct + Cs + Cz);
    } else {
      out[0] = c.undef;
      nundef[0] += 1;
    }
  }
};

// cvtemp's conversion loop, FC.cc:1662-1671
struct CvTempOp
{
  static constexpr int NIN = 1, NOUT = 1, UNROLL = 4;
  static constexpr int NCOUNT = 1;
  static constexpr int MIN_BLOCKS = 5;
  static constexpr bool HEAVY = false;
  static constexpr bool USES_EWT = false, USES_POW = false;
  float tconvert;
  template <bool ALL>
  __device__ __forceinline__ void point(const float* in, float* out, const PointCtx& c, long long, unsigned* nundef) const
  {
    if (ALL || is_def(in[0], c.undef))
      out[0] = in[0] + tconvert;
    else {
      out[0] = c.undef;
      nundef[0] += 1;
    }
  }
};

// cvtemp's "input does not seem to need converting" branch: a bit copy, flag untouched (FC.cc:1651-1657)
struct CopyOp
{
  static constexpr int NIN = 1, NOUT = 1, UNROLL = 4;
  static constexpr int NCOUNT = 0;
  static constexpr int MIN_BLOCKS = 5;
  static constexpr bool HEAVY = false;
  static constexpr bool USES_EWT = false, USES_POW = false;
  template <bool ALL>
  __device__ __forceinline__ void point(const float* in, float* out, const PointCtx&, long long, unsigned*) const
  {
    out[0] = in[0];
  }
};

// cvhum, FC.cc:1738-1817.  Branch-free path for temperatures inside the saturation table (dev::EwtFast), the four points of a
// float4 group in one basic block; points outside the table (their result is undefined or, just below -100 degC, extrapolated)
// are redone by `exact`, the reference's expressions with the ordinary operators.
#ifndef FCB_CV_U
#define FCB_CV_U 2
#define FCB_CV_MB 4
#endif
struct CvHumOp
{
  static constexpr int NIN = 2, NOUT = 1, UNROLL = FCB_CV_U;
  static constexpr int NCOUNT = 1;
  static constexpr int MIN_BLOCKS = FCB_CV_MB;
  static constexpr bool HEAVY = false;
  static constexpr bool USES_EWT = true, USES_POW = false;
  static constexpr bool QUAD = true;
  int compute; // 1..5 after the unit remap
  float tconv, tdconv, unit_scale;

  // returns "the table lookups were in range"; r = the value
  __device__ __noinline__ static bool exact(float t, float h, int compute, float tconv, float tdconv, float unit_scale, const dev::EwtTable& tab, float& r)
  {
    const dev::Ewt e(t - tconv);
    if (compute <= 3) {
      if (!e.defined)
        return false;
      const float et = e.value(tab);
      const float rh = rh_fraction(h);
      r = e.inverse(tab, rh * et) + tdconv;
      return true;
    }
    const dev::Ewt e2(h - tconv);
    if (!(e.defined && e2.defined))
      return false;
    r = (e2.value(tab) / e.value(tab)) * unit_scale;
    return true;
  }

  __device__ __forceinline__ bool fast(float t, float h, const dev::EwtTable& tab, float& r) const
  {
    const dev::EwtFast e(tab, t - tconv);
    if (compute <= 3) {
      r = e.dewpoint(tab, rh_fraction_select(h) * e.et) + tdconv;
      return e.plausible;
    }
    const dev::EwtFast e2(tab, h - tconv);
    r = dev::div_midrange(e2.et, e.et) * unit_scale; // both in [3.4e-5, 1013.25]
    return e.plausible && e2.plausible;
  }

  template <bool ALL>
  __device__ __forceinline__ void finish(bool def, bool ok, float r, float undef, float& out, unsigned* nundef) const
  {
    const bool good = def && ok;
    out = good ? r : undef;
    nundef[0] += good ? 0u : 1u;
  }

  template <bool ALL>
  __device__ __forceinline__ void point(const float* in, float* out, const PointCtx& c, long long, unsigned* nundef) const
  {
    const bool def = ALL || (is_def(in[0], c.undef) && is_def(in[1], c.undef));
    float r;
    bool ok = fast(in[0], in[1], c.tab, r);
    if (!ok && def)
      ok = exact(in[0], in[1], compute, tconv, tdconv, unit_scale, c.tab, r);
    finish<ALL>(def, ok, r, c.undef, out[0], nundef);
  }

  template <bool ALL>
  __device__ __forceinline__ void quad(const float (*in)[4], float (*out)[4], const PointCtx& c, unsigned* nundef) const
  {
    float r[4];
    unsigned okm = 0, defm = 0;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const bool def = ALL || (is_def(in[0][w], c.undef) && is_def(in[1][w], c.undef));
      defm |= def ? (1u << w) : 0u;
      okm |= fast(in[0][w], in[1][w], c.tab, r[w]) ? (1u << w) : 0u;
    }
    const unsigned redo = defm & ~okm;
    if (redo) {
#pragma unroll
      for (int w = 0; w < 4; ++w)
        if (redo & (1u << w)) {
          float v = 0.f;
          if (exact(in[0][w], in[1][w], compute, tconv, tdconv, unit_scale, c.tab, v)) {
            okm |= 1u << w;
            r[w] = v;
          }
        }
    }
#pragma unroll
    for (int w = 0; w < 4; ++w)
      finish<ALL>((defm >> w) & 1u, (okm >> w) & 1u, r[w], c.undef, out[0][w], nundef);
  }
};

// abshum, FC.cc:1676-1736 (Vaisala).  The reference includes <cmath> without `using namespace std`, so its
// unqualified sqrt(v) and exp(x) on float arguments are the C library's DOUBLE functions: v*sqrt(v) and
// Pc*exp(..) are double expressions rounded to float on assignment.  exp() here is CUDA's (<= 1 ulp of a
// double): the float result differs from glibc's in about one case in 2^29.
#ifndef FCB_AH_U
#define FCB_AH_U 1
#define FCB_AH_MB 4
#endif
struct AbsHumOp
{
  static constexpr int NIN = 2, NOUT = 1, UNROLL = FCB_AH_U;
  static constexpr int NCOUNT = 1;
  static constexpr int MIN_BLOCKS = FCB_AH_MB;
  static constexpr bool HEAVY = false;
  static constexpr bool USES_EWT = false, USES_POW = false;
  template <bool ALL>
  __device__ __forceinline__ void point(const float* in, float* out, const PointCtx& c, long long, unsigned* nundef) const
  {
    const float C = (float)2.16679, C1 = (float)-7.85951783, C2 = (float)1.84408259, C3 = (float)-11.7866497, C4 = (float)22.6807411,
                C5 = (float)-15.9618719, C6 = (float)1.80122502, Tc = (float)647.096, Pc = 220640.f;
    if (ALL || (is_def(in[0], c.undef) && is_def(in[1], c.undef))) {
      const float t = in[0];
      const float v = 1.f - t / Tc, tii = 1.f / t;
      const float v2 = v * v, v3 = v * v2, v4 = v2 * v2;
      const float v1_5 = (float)((double)v * sqrt((double)v));
      const float v3_5 = v2 * v1_5, v7_5 = v4 * v3_5;
      const float arg = Tc * tii * (C1 * v + C2 * v1_5 + C3 * v3 + C4 * v3_5 + C5 * v4 + C6 * v7_5);
      const float Pws = (float)((double)Pc * exp((double)arg));
      const float Pw = Pws * in[1];
      out[0] = C * Pw * 100.f * tii;
    } else {
      out[0] = c.undef;
      nundef[0] += 1;
    }
  }
};

// underCooledRain, FC.cc:2231-2264
struct UnderCooledOp
{
  static constexpr int NIN = 3, NOUT = 1, UNROLL = 2;
  static constexpr int NCOUNT = 1;
  static constexpr int MIN_BLOCKS = 5;
  static constexpr bool HEAVY = false;
  static constexpr bool USES_EWT = false, USES_POW = false;
  float precipMin, tkMax, snowRateMax;
  template <bool ALL>
  __device__ __forceinline__ void point(const float* in, float* out, const PointCtx& c, long long, unsigned* nundef) const
  {
    if (ALL || (is_def(in[0], c.undef) && is_def(in[1], c.undef) && is_def(in[2], c.undef))) {
      out[0] = (in[0] >= precipMin && in[2] <= tkMax && in[1] <= in[0] * snowRateMax) ? 1.f : 0.f;
    } else {
      out[0] = c.undef;
      nundef[0] += 1;
    }
  }
};

// ------------------------------------------------------------------------------------ cvtemp's pre-pass
// compute 3 / 4 average the defined input first: a FLOAT accumulation in index order in the reference's
// (serial) build, compared with t0/2.  A parallel sum has a different rounding, so the decision is taken from
// a double-precision sum together with a rigorous bound on how far the sequential float sum can be from it
// (|fl(sum) - sum| <= (n-1) u sum|x| to first order, u = 2^-24); only if 136.575 lies inside that band -- never
// for a temperature field -- the sequential sum itself is evaluated by one thread.
struct MeanStats
{
  double sum, abs_sum;
  unsigned long long count;
  float seq_avg_sum;
  int pad;
};

template <bool ALL>
__global__ void __launch_bounds__(256) mean_stats_kernel(const float* __restrict__ x, long long n, float undef, MeanStats* st)
{
  double s = 0., a = 0.;
  unsigned long long cnt = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    if (ALL || is_def(v, undef)) {
      s += (double)v;
      a += fabs((double)v);
      cnt += 1;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_down_sync(0xffffffffu, s, o);
    a += __shfl_down_sync(0xffffffffu, a, o);
    cnt += __shfl_down_sync(0xffffffffu, cnt, o);
  }
  if ((threadIdx.x & 31) == 0 && cnt) {
    atomicAdd(&st->sum, s);
    atomicAdd(&st->abs_sum, a);
    atomicAdd(&st->count, cnt);
  }
}

template <bool ALL>
__global__ void sequential_sum_kernel(const float* __restrict__ x, long long n, float undef, MeanStats* st)
{ // FC.cc:1640-1645 as written: one thread, index order, float accumulator
  float tavg = 0.f;
  for (long long i = 0; i < n; ++i) {
    const float v = x[i];
    if (ALL || is_def(v, undef))
      tavg += v;
  }
  st->seq_avg_sum = tavg;
}

// 1 = the reference takes the copy branch, 0 = it converts, < 0 runtime error
int cvtemp_seems_converted(Call& call, const float* dx, long long n, bool all, float undef, int compute)
{
  MeanStats* st = static_cast<MeanStats*>(call.scratch(sizeof(MeanStats)));
  if (!call.ok())
    return -1;
  cudaStream_t s = call.stream();
  if (!cuda_ok(cudaMemsetAsync(st, 0, sizeof(MeanStats), s), "cvtemp: clear statistics"))
    return -1;
  const int grid = sm_count() * 8;
  if (all)
    mean_stats_kernel<true><<<grid, 256, 0, s>>>(dx, n, undef, st);
  else
    mean_stats_kernel<false><<<grid, 256, 0, s>>>(dx, n, undef, st);
  count_launch();
  MeanStats h;
  if (!cuda_ok(cudaMemcpyAsync(&h, st, sizeof(h), cudaMemcpyDeviceToHost, s), "cvtemp: fetch statistics") ||
      !cuda_ok(cudaStreamSynchronize(s), "cvtemp: statistics"))
    return -1;
  const double half_t0 = (double)H_T0 / 2.;
  float tavg = 0.f;
  bool decided = false;
  if (h.count == 0) {
    decided = true; // tavg stays 0 (FC.cc:1646)
  } else {
    const double mean = h.sum / (double)h.count;
    const double band = ((double)n * 5.97e-8 + 1e-6) * (h.abs_sum / (double)h.count) * 1.01 + 1e-30;
    if (mean - band > half_t0 || mean + band < half_t0) {
      tavg = (float)mean;
      decided = std::isfinite(mean);
    }
  }
  if (!decided) {
    if (all)
      sequential_sum_kernel<true><<<1, 1, 0, s>>>(dx, n, undef, st);
    else
      sequential_sum_kernel<false><<<1, 1, 0, s>>>(dx, n, undef, st);
    count_launch();
    if (!cuda_ok(cudaMemcpyAsync(&h.seq_avg_sum, &st->seq_avg_sum, sizeof(float), cudaMemcpyDeviceToHost, s), "cvtemp: fetch sum") ||
        !cuda_ok(cudaStreamSynchronize(s), "cvtemp: sequential sum"))
      return -1;
    tavg = h.seq_avg_sum;
    if (h.count > 0)
      tavg /= (float)(int)h.count;
  }
  return ((compute == 3 && (double)tavg < half_t0) || (compute == 4 && (double)tavg > half_t0)) ? 1 : 0;
}

} // namespace
} // namespace fcb200

// =========================================================================================== C-ABI
using namespace fcb200;

extern "C" {

int fcb200_kIndex_batched(int nx, int ny, int nfields, const float* t500, const float* t700, const float* rh700, const float* t850, const float* rh850,
                          float p500, float p700, float p850, int compute, float* kfield, int* fDefined, float undef)
{ // FC.cc:745-814
  if (p500 <= 0.0 || p500 >= p700 || p700 >= p850)
    return 0;
  KIndexOp op;
  if (compute == 1)
    op = KIndexOp{1.f, 1.f, 1.f};
  else if (compute == 2)
    op = KIndexOp{host_pidcp(p500), host_pidcp(p700), host_pidcp(p850)};
  else
    return 0;
  const float* in[5] = {t500, t700, rh700, t850, rh850};
  const int pf[5] = {1, 1, 1, 1, 1};
  return run_elementwise(make_batch(nx, ny, nfields), op, in, pf, kfield, fDefined, undef, FLAG_FROM_COUNT, NoMeta());
}
int fcb200_kIndex(int nx, int ny, const float* t500, const float* t700, const float* rh700, const float* t850, const float* rh850, float p500, float p700,
                  float p850, int compute, float* kfield, int* fDefined, float undef)
{
  return fcb200_kIndex_batched(nx, ny, 1, t500, t700, rh700, t850, rh850, p500, p700, p850, compute, kfield, fDefined, undef);
}

int fcb200_ductingIndex_batched(int nx, int ny, int nfields, const float* t850, const float* rh850, float p850, int compute, float* duct,
                                int* fDefined, float undef)
{ // FC.cc:816-870
  if (p850 <= 0.0)
    return 0;
  DuctingIndexOp op;
  if (compute == 1)
    op.tconvert = 1.f;
  else if (compute == 2)
    op.tconvert = host_pidcp(p850);
  else
    return 0;
  const float* in[2] = {t850, rh850};
  const int pf[2] = {1, 1};
  return run_elementwise(make_batch(nx, ny, nfields), op, in, pf, duct, fDefined, undef, FLAG_FROM_COUNT, NoMeta());
}
int fcb200_ductingIndex(int nx, int ny, const float* t850, const float* rh850, float p850, int compute, float* duct, int* fDefined, float undef)
{
  return fcb200_ductingIndex_batched(nx, ny, 1, t850, rh850, p850, compute, duct, fDefined, undef);
}

int fcb200_showalterIndex_batched(int nx, int ny, int nfields, const float* t500, const float* t850, const float* rh850, float p500, float p850,
                                  int compute, float* sfield, int* fDefined, float undef)
{ // FC.cc:872-971
  if (p500 <= 0.0 || p500 >= p850)
    return 0;
  const float pi500 = H_CP * host_pidcp(p500), pi850 = H_CP * host_pidcp(p850);
  ShowalterOp op;
  if (compute == 1) {
    op.cvt500 = 1.f;
    op.cvt850 = 1.f;
    op.dryadiabat = H_CP * (H_CP / pi850) * (pi500 / H_CP);
  } else if (compute == 2) {
    op.cvt500 = pi500 / H_CP;
    op.cvt850 = pi850 / H_CP;
    op.dryadiabat = H_CP * (pi500 / H_CP);
  } else
    return 0;
  op.p500 = p500;
  op.p850 = p850;
  const float* in[4] = {t500, t850, rh850, sfield}; // the output doubles as an input: untouched where the input is undefined
  const int pf[4] = {1, 1, 1, 1};
  return run_elementwise(make_batch(nx, ny, nfields), op, in, pf, sfield, fDefined, undef, FLAG_FROM_COUNT, NoMeta());
}
int fcb200_showalterIndex(int nx, int ny, const float* t500, const float* t850, const float* rh850, float p500, float p850, int compute, float* sfield,
                          int* fDefined, float undef)
{
  return fcb200_showalterIndex_batched(nx, ny, 1, t500, t850, rh850, p500, p850, compute, sfield, fDefined, undef);
}

int fcb200_boydenIndex_batched(int nx, int ny, int nfields, const float* t700, const float* z700, const float* z1000, float p700, float p1000,
                               int compute, float* bfield, int* fDefined, float undef)
{ // FC.cc:973-1014
  if (compute <= 0 || compute >= 3)
    return 0;
  if (p700 <= 0.0 || p700 >= p1000)
    return 0;
  const float pi700 = H_CP * powf(p700 / 1000.f, 287.f / 1004.f); // :999 spells the Exner function out with p/p0 instead of p*p0inv
  BoydenOp op{(compute == 2) ? pi700 / H_CP : 1.f};
  const float* in[3] = {t700, z700, z1000};
  const int pf[3] = {1, 1, 1};
  return run_elementwise(make_batch(nx, ny, nfields), op, in, pf, bfield, fDefined, undef, FLAG_FROM_COUNT, NoMeta());
}
int fcb200_boydenIndex(int nx, int ny, const float* t700, const float* z700, const float* z1000, float p700, float p1000, int compute, float* bfield,
                       int* fDefined, float undef)
{
  return fcb200_boydenIndex_batched(nx, ny, 1, t700, z700, z1000, p700, p1000, compute, bfield, fDefined, undef);
}

int fcb200_sweatIndex_batched(int nx, int ny, int nfields, const float* t850, const float* t500, const float* td850, const float* td500,
                              const float* u850, const float* v850, const float* u500, const float* v500, float* sindex, int* fDefined, float undef)
{ // FC.cc:1016-1040
  SweatOp op;
  const float* in[8] = {t850, t500, td850, td500, u850, v850, u500, v500};
  const int pf[8] = {1, 1, 1, 1, 1, 1, 1, 1};
  return run_elementwise(make_batch(nx, ny, nfields), op, in, pf, sindex, fDefined, undef, FLAG_FROM_COUNT, NoMeta());
}
int fcb200_sweatIndex(int nx, int ny, const float* t850, const float* t500, const float* td850, const float* td500, const float* u850, const float* v850,
                      const float* u500, const float* v500, float* sindex, int* fDefined, float undef)
{
  return fcb200_sweatIndex_batched(nx, ny, 1, t850, t500, td850, td500, u850, v850, u500, v500, sindex, fDefined, undef);
}

int fcb200_seaSoundSpeed_batched(int nx, int ny, int nfields, const float* t, const float* s, float z, int compute, float* soundspeed, int* fDefined,
                                 float undef)
{ // FC.cc:1555-1602
  if (compute != 1 && compute != 2)
    return 0;
  const double Z = fabsf(z);
  SeaSoundOp op{(compute == 1) ? 0.f : H_T0, 0.01635 * Z + 0.000000175 * Z * Z};
  const float* in[2] = {t, s};
  const int pf[2] = {1, 1};
  return run_elementwise(make_batch(nx, ny, nfields), op, in, pf, soundspeed, fDefined, undef, FLAG_FROM_COUNT, NoMeta());
}
int fcb200_seaSoundSpeed(int nx, int ny, const float* t, const float* s, float z, int compute, float* soundspeed, int* fDefined, float undef)
{
  return fcb200_seaSoundSpeed_batched(nx, ny, 1, t, s, z, compute, soundspeed, fDefined, undef);
}

int fcb200_cvtemp(int nx, int ny, const float* tinp, int compute, float* tout, int* fDefined, float undef)
{ // FC.cc:1608-1674
  if (compute < 1 || compute > 4)
    return 0;
  const float tconvert = (compute == 1 || compute == 3) ? -H_T0 : +H_T0;
  const Batch b = make_batch(nx, ny, 1);
  const float* in[1] = {tinp};
  const int pf[1] = {1};
  if (compute >= 3) {
    if (!b.valid()) {
      set_error("fcb200: invalid grid (nx=%d ny=%d)", nx, ny);
      return -1;
    }
    int copy;
    {
      Call call;
      if (!call.ok())
        return -1;
      const float* dx = call.in(tinp, (size_t)b.n);
      if (!call.ok())
        return -1;
      copy = cvtemp_seems_converted(call, dx, b.n, *fDefined == ALL_DEFINED, undef, compute);
      if (call.finish(Finalizer()) < 0 || copy < 0)
        return -1;
    }
    if (copy) {
      if (tout == tinp)
        return 1;
      return run_elementwise(b, CopyOp(), in, pf, tout, fDefined, undef, FLAG_UNCHANGED, NoMeta());
    }
  }
  return run_elementwise(b, CvTempOp{tconvert}, in, pf, tout, fDefined, undef, FLAG_FROM_COUNT, NoMeta());
}

int fcb200_cvtemp_batched(int nx, int ny, int nfields, const float* tinp, int compute, float* tout, int* fDefined, float undef)
{ // compute 1 / 2: one launch; compute 3 / 4 decide per field whether to convert at all -> field by field
  if (compute < 1 || compute > 4)
    return 0;
  if (compute <= 2) {
    const float* in[1] = {tinp};
    const int pf[1] = {1};
    return run_elementwise(make_batch(nx, ny, nfields), CvTempOp{(compute == 1) ? -H_T0 : +H_T0}, in, pf, tout, fDefined, undef, FLAG_FROM_COUNT,
                           NoMeta());
  }
  const size_t n = (size_t)nx * (size_t)ny;
  for (int k = 0; k < nfields; ++k) {
    const int rc = fcb200_cvtemp(nx, ny, tinp + k * n, compute, tout + k * n, fDefined + k, undef);
    if (rc != 1)
      return rc;
  }
  return 1;
}

int fcb200_cvhum_batched(int nx, int ny, int nfields, const float* t, const float* huminp, const char* unit, int compute, float* humout, int* fDefined,
                         float undef)
{ // FC.cc:1738-1817
  CvHumOp op;
  op.unit_scale = 100.f;
  if (compute == 1 && unit_is(unit, "celsius"))
    compute = 2;
  if ((compute == 4 || compute == 5) && unit_is(unit, "1"))
    op.unit_scale = 1.f;
  if (compute < 1 || compute > 5)
    return 0;
  op.compute = compute;
  op.tconv = (compute == 1 || compute == 2 || compute == 4) ? H_T0 : 0.f;
  op.tdconv = (compute == 1) ? H_T0 : 0.f;
  const float* in[2] = {t, huminp};
  const int pf[2] = {1, 1};
  return run_elementwise(make_batch(nx, ny, nfields), op, in, pf, humout, fDefined, undef, FLAG_FROM_COUNT, NoMeta());
}
int fcb200_cvhum(int nx, int ny, const float* t, const float* huminp, const char* unit, int compute, float* humout, int* fDefined, float undef)
{
  return fcb200_cvhum_batched(nx, ny, 1, t, huminp, unit, compute, humout, fDefined, undef);
}

int fcb200_abshum_batched(int nx, int ny, int nfields, const float* t, const float* rhum, float* abshumout, int* fDefined, float undef)
{ // FC.cc:1676-1736
  const float* in[2] = {t, rhum};
  const int pf[2] = {1, 1};
  return run_elementwise(make_batch(nx, ny, nfields), AbsHumOp(), in, pf, abshumout, fDefined, undef, FLAG_FROM_COUNT, NoMeta());
}
int fcb200_abshum(int nx, int ny, const float* t, const float* rhum, float* abshumout, int* fDefined, float undef)
{
  return fcb200_abshum_batched(nx, ny, 1, t, rhum, abshumout, fDefined, undef);
}

int fcb200_underCooledRain_batched(int nx, int ny, int nfields, const float* precip, const float* snow, const float* tk, float precipMin,
                                   float snowRateMax, float tcMax, float* undercooled, int* fDefined, float undef)
{ // FC.cc:2231-2264
  UnderCooledOp op{precipMin, tcMax + H_T0, snowRateMax};
  const float* in[3] = {precip, snow, tk};
  const int pf[3] = {1, 1, 1};
  return run_elementwise(make_batch(nx, ny, nfields), op, in, pf, undercooled, fDefined, undef, FLAG_FROM_COUNT, NoMeta());
}
int fcb200_underCooledRain(int nx, int ny, const float* precip, const float* snow, const float* tk, float precipMin, float snowRateMax, float tcMax,
                           float* undercooled, int* fDefined, float undef)
{
  return fcb200_underCooledRain_batched(nx, ny, 1, precip, snow, tk, precipMin, snowRateMax, tcMax, undercooled, fDefined, undef);
}

} // extern "C"

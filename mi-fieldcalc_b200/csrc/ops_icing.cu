// ops_icing.cu -- the four sea-spray vessel icing models.
// Reference: src/mi_fieldcalc/FieldCalculationsVesselIcing.cc (VI.cc).
//
// Overland and Mertins are cheap and HBM-bound (28 B/point).  ModStall (all double: RK4 x 50 with four
// exp each, two fixed-point loops) and MINCOG (float RK4 x 50 + 17 bisection steps per height) need
// several hundred transcendentals per point against 48 B/point of traffic: they are instruction-bound
// and are reported against an instruction bound, not the HBM roofline.
//
// Everything that depends only on the call's scalars (vs, alpha, zmin, zmax) is evaluated ONCE on the
// host with the same libm the reference uses and passed to the kernel, so those terms are bit-exact.
#include "ew_driver.cuh"

#include "../../include/fcb200.h"

#include <cmath>
#include <cstdlib>

namespace fcb200 {
namespace {

using dev::is_def;
using dev::K_T0;

__device__ __forceinline__ bool def6(bool all, const float* in, float undef)
{
  bool ok = true;
#pragma unroll
  for (int k = 0; k < 6; ++k)
    ok = ok && is_def(in[k], undef);
  return all || ok;
}

// freezing point of sea water, Stallabrass (1980): double, but sal*sal is a float product (VI.cc:95, 127, 245)
__device__ __forceinline__ double freezing_point(float sal)
{
  return (-0.002 - 0.0524 * (double)sal) - 6.0E-5 * (double)(sal * sal);
}

// in: airtemp, seatemp, u, v, sal, aice
// Shape (round 2): one float4 group per thread, 6 CTAs/SM -- six input fields are 48 registers of loads in flight with two groups;
// with one group the kernel fits 40 registers and six CTAs hide the latency: 0.70 -> 0.80 (30 % masked 0.72 -> 0.84)
template <int U_ = 1, int MB_ = 6>
struct OverlandOpT
{ // VI.cc:77-112
  static constexpr int NIN = 6, NOUT = 1, UNROLL = U_;
  static constexpr int NCOUNT = 1;
  static constexpr int MIN_BLOCKS = MB_;
  static constexpr bool HEAVY = false;
  static constexpr bool USES_EWT = false, USES_POW = false;
  template <bool ALL>
  __device__ __forceinline__ void point(const float* in, float* out, const PointCtx& c, long long, unsigned* nundef) const
  {
    const float airtemp = in[0], seatemp = in[1], sal = in[4], aice = in[5];
    bool ok = def6(ALL, in, c.undef) && ((double)aice < 0.4);
    float r = c.undef;
    if (ok) {
      const double Tf = freezing_point(sal);
      if ((double)seatemp < Tf)
        ok = false;
      else {
        const double ff = (double)dev::absval(in[2], in[3]);
        const double ppr = ff * (Tf - (double)airtemp) / (1 + 0.3 * ((double)seatemp - Tf));
        r = (float)(2.73e-2 * ppr + 2.91e-4 * (ppr * ppr) + 1.84e-6 * ppr * ppr * ppr);
      }
    }
    if (!ok) {
      r = c.undef;
      nundef[0] += 1;
    }
    out[0] = r;
  }
};

// (shape: see OverlandOpT; 0.61 -> 0.85, 30 % masked 0.71 -> 0.88)
template <int U_ = 1, int MB_ = 6>
struct MertinsOpT
{ // VI.cc:114-180
  static constexpr int NIN = 6, NOUT = 1, UNROLL = U_;
  static constexpr int NCOUNT = 1;
  static constexpr int MIN_BLOCKS = MB_;
  static constexpr bool HEAVY = false;
  static constexpr bool USES_EWT = false, USES_POW = false;
  template <bool ALL>
  __device__ __forceinline__ void point(const float* in, float* out, const PointCtx& c, long long, unsigned* nundef) const
  {
    const float sal = in[4], aice = in[5];
    bool ok = def6(ALL, in, c.undef) && ((double)aice < 0.4);
    float r = c.undef;
    if (ok) {
      const double Tf = freezing_point(sal);
      if ((double)in[1] < Tf)
        ok = false;
      else {
        const double ff = (double)dev::absval(in[2], in[3]);
        const double ta = in[0], sst = in[1];
        if (!(ff >= 10.8)) {
          r = 0.f;
        } else {
          double t1, t2, t3;
          if (ff < 17.2) {
            t1 = -1.15 * sst - 4.3;
            t2 = -1.5 * sst - 10;
            t3 = -10000;
          } else if (ff < 20.8) {
            t1 = -0.6 * sst - 3.2;
            t2 = -1.05 * sst - 5.6;
            t3 = -1.75 * sst - 12.5;
          } else if (ff < 28.5) {
            t1 = -0.3 * sst - 2.6;
            t2 = -0.66 * sst - 3.32;
            t3 = -1.325 * sst - 7.651;
          } else {
            t1 = -0.14 * sst - 2.28;
            t2 = -0.3 * sst - 2.6;
            t3 = -1.16 * sst - 5.22;
          }
          if (ta > -2)
            r = 0.f;
          else if (ta > t1)
            r = (float)0.8333;
          else if (ta > t2)
            r = (float)2.0833;
          else if (ta <= t3 || ff < 17.2)
            r = (float)4.375;
          else
            r = (float)6.25;
        }
      }
    }
    if (!ok) {
      r = c.undef;
      nundef[0] += 1;
    }
    out[0] = r;
  }
};

__device__ __forceinline__ double icing_f1_d(double t)
{ // VI.cc:53-57, T = double
  return 0.6112 * exp(17.67 * t / (t + 243.5));
}

__device__ __forceinline__ float icing_f1_f(float t)
{ // VI.cc:53-57, T = float
  return (float)0.6112 * expf((float)17.67 * t / (t + (float)243.5));
}

__device__ __forceinline__ float kT4_f(float tc)
{ // VI.cc:65-70, T = float
  const float sigma = (float)5.67e-8;
  const float a = tc + K_T0;
  const float a2 = a * a;
  return sigma * (a2 * a2);
}

// in: sal, wave, x_wind, y_wind, airtemp, rh, sst, p, Pw, aice, depth ; Pw (in[8]) is not tested (VI.cc:208, 696)
__device__ __forceinline__ bool def10_no_pw(bool all, const float* in, float undef)
{
  bool ok = true;
#pragma unroll
  for (int k = 0; k < 11; ++k)
    if (k != 8)
      ok = ok && is_def(in[k], undef);
  return all || ok;
}

struct ModStallOp
{ // VI.cc:182-337 -- everything in double
  static constexpr int NIN = 11, NOUT = 1, UNROLL = 1;
  static constexpr int NCOUNT = 1;
  static constexpr int MIN_BLOCKS = 2;
  static constexpr bool HEAVY = true;
  static constexpr bool USES_EWT = false, USES_POW = false;
  double vs_cos_alpha; // vs * cos(alpha), host libm
  float zmin;
  int number;
  template <bool ALL>
  __device__ __forceinline__ void point(const float* in, float* out, const PointCtx& c, long long, unsigned* nundef) const
  {
    const float sal = in[0], wave = in[1], airtemp = in[4], rh = in[5], sst = in[6], p = in[7], Pw = in[8], aice = in[9], depth = in[10];
    if (!(def10_no_pw(ALL, in, c.undef) && ((double)aice < 0.4))) {
      out[0] = c.undef;
      nundef[0] += 1;
      return;
    }
    // deep-water wave speed, then the shallow-water fixed point (:218-237)
    double cw = (9.81 / (2 * M_PI)) * (double)Pw;
    if ((double)depth <= cw * (double)Pw && cw != 0) {
      cw = 1.0;
      double err = 1.0;
      int j = 0;
      // The reference iterates until |c_new - c| <= 1e-5 or 10001 times, then gives up with c = 0.  An iterate that EQUALS the one
      // two steps back closes a cycle: the same two values and the same error repeat for ever, so if that error is above the
      // tolerance the loop can only end at its cap, with c = 0 -- decided here at once, same result.  (An undefined period,
      // which the reference does not test, FC VI.cc:208, makes c jump between 1 and g*depth: 10001 tanh per point otherwise.)
      double c_before = __longlong_as_double(0x7ff8000000000000LL); // NaN: equals nothing
      while (err > 1e-5) {
        const double c_new = (9.81 * (double)Pw / (2 * M_PI)) * tanh(2 * M_PI * (double)depth / ((double)Pw * cw));
        err = fabs(c_new - cw);
        const bool cycle = c_new == c_before && err > 1e-5;
        c_before = cw;
        cw = c_new;
        j = j + 1;
        if (j > 10000 || cycle) {
          cw = 0.0;
          break;
        }
      }
    }
    const double Vr = cw - vs_cos_alpha;
    const double v = (double)dev::absval(in[2], in[3]);
    const double Tf = freezing_point(sal);
    const double ha_ = 5.17, ha = ha_ * pow(v, 0.8);
    const double ratio = 89.5 / ha_;
    const double tau = 11.25 - v / 4.0;
    double td = sst;
    if (tau > 0.0) { // droplet cooling: RK4, 50 steps (:262-281)
      const double K = 311000.0 / (((double)p / 10.0) * 1005.0);
      const double M = 0.2 * (double)airtemp + K * (double)rh * (double)icing_f1_f(airtemp); // float argument -> icing_f1<float>
      const double h = tau / 50.0;
      double y = sst;
#pragma unroll 1
      for (int s = 0; s < 50; ++s) {
        const double k1 = (M - 0.2 * y) - K * icing_f1_d(y);
        const double y2 = y + 0.5 * h * k1;
        const double k2 = (M - 0.2 * y2) - K * icing_f1_d(y2);
        const double y3 = y + 0.5 * h * k2;
        const double k3 = (M - 0.2 * y3) - K * icing_f1_d(y3);
        const double y4 = y + h * k3;
        y += h * ((1.0 / 6.0) * (((k1 + 2.0 * k2) + 2.0 * k3) + ((M - 0.2 * y4) - K * icing_f1_d(y4))));
      }
      td = y;
    }
    const double rh_f1_air = (double)(rh * icing_f1_f(airtemp)); // float * icing_f1<float>: a float product (VI.cc:306)
    double ice = 0;
#pragma unroll 1
    for (int k = 0; k < number; ++k) { // freezing fraction per height (:288-326)
      const double rw = 6.46E-5 * (double)wave * (Vr * Vr) * exp(-0.55 * ((double)zmin + 0.5 * k)) * v;
      double N = 0.0, err = 1.0;
      int j = 0;
      while (err >= 1.0E-5 && N >= 0 && N <= 1) {
        const double Ts = (1.0 + N) * Tf;
        const double ri = (0.012012012 * rw * (Ts - td) + (ha / 333000.0) * ((Ts - (double)airtemp) + ratio * (icing_f1_d(Ts) - rh_f1_air)));
        const double N1 = ri / rw;
        err = fabs(N1 - N);
        N = N1;
        j = j + 1;
        if (j > 1000) {
          N = 0.0;
          break;
        }
      }
      if (N < 0.0)
        N = 0.0;
      else if (N > 1.0)
        N = 1.0;
      ice += N * (rw / 890.0) * 3600.0 * 100.0;
    }
    out[0] = (float)fabs(ice / number);
  }
};

// ---- MINCOG, V = float (VI.cc:339-675) -----------------------------------------------------------------

struct Ffz
{
  float Sw, Ta, ha, he, ea, RH, rw, Tsp, Lwdown, Swdown;
};

__device__ __forceinline__ float freeze_frac_zero(const Ffz& z, float N)
{ // VI.cc:345-361
  const float cw = 4000;
  const float lfs = (float)(3.33e5 * 0.7);
  const float Sb = (float)((double)z.Sw / (1 - (double)N * (1 - 0.3)));
  const float Ts = (float)-54.1126 * (Sb / (1000 - Sb));
  const float es = 10 * icing_f1_f(Ts);
  const float Qc = z.ha * (Ts - z.Ta);
  const float Qe = z.he * (es - z.RH * z.ea);
  const float Qd = z.rw * cw * (Ts - z.Tsp);
  const float Lwup = kT4_f(Ts);
  const float Qr = (float)((double)(Lwup - z.Lwdown) - 0.44 * (double)z.Swdown);
  const float ri = (1 / lfs) * (Qc + Qe + Qd + Qr);
  const float N1 = ri / z.rw;
  return N1 - N;
}

__device__ __forceinline__ float mincog_bisection(const Ffz& z, float a, float b, int iterations)
{ // VI.cc:381-415; `iterations` = min(int(log2f((b-a)/eps)), 100) is a call constant (17)
  float ffa = freeze_frac_zero(z, a);
  const float ffb = freeze_frac_zero(z, b);
  if ((ffa > 0) == (ffb > 0))
    return 0;
  float cc = 0;
  int j = 0;
#pragma unroll 1
  for (; j < iterations; ++j) {
    cc = (a + b) / 2;
    const float ffc = freeze_frac_zero(z, cc);
    if (ffc == 0)
      return cc;
    if ((ffc > 0) != (ffa > 0)) {
      b = cc;
    } else {
      a = cc;
      ffa = ffc;
    }
  }
  if (j >= 100)
    cc = 0;
  return cc;
}

__device__ __forceinline__ float f10mk(float t, float M, float K)
{ // VI.cc:59-63
  return (M - (float)0.2 * t) - K * 10 * icing_f1_f(t);
}

struct MincogOp
{ // VI.cc:465-705
  static constexpr int NIN = 11, NOUT = 1, UNROLL = 1;
  static constexpr int NCOUNT = 1;
  static constexpr int MIN_BLOCKS = 2;
  static constexpr bool HEAVY = true;
  static constexpr bool USES_EWT = false, USES_POW = false;
  // call constants, evaluated on the host exactly as the reference's expressions
  float vs, cos_alpha, sin_beta, drag, Vf, zmin;
  double cos_beta_d; // cos((double)beta)
  int number, alt, iterations;

  __device__ __forceinline__ float model(float sal, float wave, float x_wind, float y_wind, float airtemp, float rh, float sst, float p, float Pw,
                                         float depth) const
  {
    const float v = dev::absval(x_wind, y_wind);
    if (v < 1 || (double)wave < 0.1)
      return 0;

    const float c_0 = (float)(9.81 / (2 * M_PI) * (double)Pw);
    float c = c_0;
    if (depth <= c * Pw && c_0 != 0) {
      c = 1;
      int j = 0;
      const float a = (float)(2 * M_PI * (double)depth / (double)Pw);
#pragma unroll 1
      float c_before = __int_as_float(0x7fc00000); // NaN: equals nothing
      for (; j < 1000; ++j) {
        const float c_new = (float)((double)c_0 * tanh((double)(a / c)));
        const float err = fabsf(c_new - c);
        const bool cycle = c_new == c_before; // (see ModStallOp: the iterates repeat with this error for ever)
        c_before = c;
        c = c_new;
        if ((double)err <= 1e-5)
          break;
        if (cycle) {
          j = 1000;
          break;
        }
      }
      if (j >= 1000)
        c = 0;
    }

    const float Vr = c - vs * cos_alpha;
    const float tper = fabsf(c * Pw / Vr);
    if (tper <= 0)
      return 0;

    const float Wrx = (float)fabs((double)v * cos_beta_d - (double)vs);
    const float Wry = fabsf(v * sin_beta);
    const float Wr_inv = 1 / dev::absval(Wrx, Wry);

    const float hax = (float)(6.0617 * pow((double)Wrx, 1.82));
    const float hay = (float)(4.8496 * pow((double)Wry, 1.8));
    const float ha = (hax + hay) / (Wrx + Wry);

    const float vmax5 = (v < 5.f) ? 5.f : v; // std::max<V>(v, 5)
    const float tdur = (float)(0.1230 + 0.7008 * (double)fabsf(Vr * wave) / (double)vmax5);
    const float Nf = 1 / (4 * tper);

    const float beta_r = (float)(M_PI - (double)asinf(v * sin_beta * Wr_inv));
    float br;
    if ((double)beta_r <= (M_PI / 2))
      br = (float)(91 * M_PI / 180);
    else if ((double)beta_r > (M_PI))
      br = (float)M_PI;
    else
      br = beta_r;
    const float sin_br = sinf(br);
    const float sin_beta_r_2 = sin_br * sin_br;
    const float cos_beta_r = cosf(br);
    const float cos_2_beta_r = cosf(2 * br);

    const float r0 = (float)13.18, a0 = (float)32.88, b0 = (float)6.605;
    const float a0_2 = a0 * a0, b0_2 = b0 * b0, r0_2 = r0 * r0;
    const float c0 = (float)(1.4142135623730951 * (double)a0 * (double)b0 *
                             (double)sqrtf((b0_2 - a0_2) * cos_2_beta_r + a0_2 + b0_2 - 2 * r0_2 * sin_beta_r_2));
    const float r = (r0 * 2 * b0_2 * cos_beta_r + c0) / ((b0_2 - a0_2) * cos_2_beta_r + a0_2 + b0_2);

    const float tau_const = r * Wr_inv;
    const float tau = tau_const * drag;

    const float ea = 10 * icing_f1_f(airtemp);
    const float K = (float)(0.2 * 0.622 * 2.5E6 / ((double)p * 1005.0));
    const float M = (float)(0.2 * (double)airtemp + (double)(K * rh * ea));

    float y = sst;
    {
      const float h = tau / 50, h2 = h / 2;
#pragma unroll 1
      for (int s = 0; s < 50; ++s) { // VI.cc:450-463
        const float k1 = h2 * f10mk(y, M, K);
        const float k2 = h * f10mk(y + k1, M, K);
        const float k3 = h * f10mk(y + k2 / 2, M, K);
        const float k4 = h2 * f10mk(y + k3, M, K);
        y += (k1 + k2 + k3 + k4) / 3;
      }
    }
    const float Td = y;
    const float Tsp = (float)(0.5 * (double)(Td + sst));

    const float Vdz = (float)6.67;
    const float Vdcomp = (float)((double)Wrx * 0.9962 + (double)Vdz * 0.0872);

    float lwc0;
    if (alt == 1) {
      lwc0 = (float)(6.36E-5 * (double)wave * (double)(Vr * Vr));
    } else {
      const float lambda = c * Pw, dl = (float)(4 * M_PI * (double)depth / (double)lambda);
      const float cg = (c / 2) * (1 + dl / sinhf(dl));
      const float Vgr = cg - vs * cos_alpha;
      lwc0 = (float)(9.5205E-4 * (double)(wave * wave) * (double)sqrtf(wave / lambda) * (double)Vgr);
    }
    lwc0 = fabsf(lwc0);

    const float he = (float)((double)ha * 1738.6 / (double)p);
    const float Swdown_model = 0;
    const float eps_atm = (float)0.7;
    const float Lwdown = eps_atm * kT4_f(airtemp);
    const float Swdown = Swdown_model * Vf;

    float icing = 0;
#pragma unroll 1
    for (int k = 0; k < number; ++k) {
      const float lwc = (float)((double)lwc0 * exp(-0.55 * ((double)zmin + 0.5 * k)));
      const float rw = lwc * Vdcomp * Nf * tdur;
      const Ffz z = {sal, airtemp, ha, he, ea, rh, rw, Tsp, Lwdown, Swdown};
      float N = mincog_bisection(z, (float)-0.5, (float)1.3, iterations);
      if (N < 0)
        N = 0;
      else if (1 < N)
        N = 1;
      icing += rw * N;
    }
    return fabsf(icing / number) * (float)(3600.0 * 100.0 / 890.0);
  }

  template <bool ALL>
  __device__ __forceinline__ void point(const float* in, float* out, const PointCtx& c, long long, unsigned* nundef) const
  {
    const float sal = in[0], sst = in[6], aice = in[9];
    if (def10_no_pw(ALL, in, c.undef) && ((double)aice < 0.4) && ((double)sst > (-54.1126 * (double)sal / (double)(1000 - sal)))) {
      out[0] = model(sal, in[1], in[2], in[3], in[4], in[5], sst, in[7], in[8], in[10]);
    } else {
      out[0] = c.undef;
      nundef[0] += 1;
    }
  }
};

template <class Op>
int run_icing(const Op& op, int nx, int ny, int nfields, const float* const* host_in, float* host_out, int* fDefined, float undef)
{
  EwJob<Op> job;
  job.nx = nx;
  job.ny = ny;
  job.nfields = nfields;
  for (int k = 0; k < Op::NIN; ++k) {
    job.in[k] = host_in[k];
    job.per_field[k] = true;
  }
  job.out[0] = host_out;
  job.flags_in = fDefined;
  job.flags_out[0] = fDefined;
  job.undef = undef;
  return run_ew_job(op, job);
}

} // namespace
} // namespace fcb200

// =========================================================================================== C-ABI
using namespace fcb200;

extern "C" {

int fcb200_vesselIcingOverland_batched(int nx, int ny, int nfields, const float* airtemp, const float* seatemp, const float* u, const float* v,
                                       const float* sal, const float* aice, float* icing, int* fDefined, float undef)
{
  const float* in[6] = {airtemp, seatemp, u, v, sal, aice};
  return run_icing(OverlandOpT<>(), nx, ny, nfields, in, icing, fDefined, undef);
}
int fcb200_vesselIcingOverland(int nx, int ny, const float* airtemp, const float* seatemp, const float* u, const float* v, const float* sal,
                               const float* aice, float* icing, int* fDefined, float undef)
{
  return fcb200_vesselIcingOverland_batched(nx, ny, 1, airtemp, seatemp, u, v, sal, aice, icing, fDefined, undef);
}

int fcb200_vesselIcingMertins_batched(int nx, int ny, int nfields, const float* airtemp, const float* seatemp, const float* u, const float* v,
                                      const float* sal, const float* aice, float* icing, int* fDefined, float undef)
{
  const float* in[6] = {airtemp, seatemp, u, v, sal, aice};
  return run_icing(MertinsOpT<>(), nx, ny, nfields, in, icing, fDefined, undef);
}
int fcb200_vesselIcingMertins(int nx, int ny, const float* airtemp, const float* seatemp, const float* u, const float* v, const float* sal,
                              const float* aice, float* icing, int* fDefined, float undef)
{
  return fcb200_vesselIcingMertins_batched(nx, ny, 1, airtemp, seatemp, u, v, sal, aice, icing, fDefined, undef);
}

int fcb200_vesselIcingModStall_batched(int nx, int ny, int nfields, const float* sal, const float* wave, const float* x_wind, const float* y_wind,
                                       const float* airtemp, const float* rh, const float* sst, const float* p, const float* Pw, const float* aice,
                                       const float* depth, float vs, float alpha, float zmin, float zmax, float* icing, int* fDefined, float undef)
{ // VI.cc:192-201
  const double num = zmax - zmin;
  const int number = (int)(num * 2 + 1);
  if (zmax < zmin || fmod(num, 1) != 0)
    return 0;
  if (vs < 0 || alpha < 0 || zmin < 0 || zmax < 0)
    return 0;
  ModStallOp op;
  op.vs_cos_alpha = vs * cos((double)alpha);
  op.zmin = zmin;
  op.number = number;
  const float* in[11] = {sal, wave, x_wind, y_wind, airtemp, rh, sst, p, Pw, aice, depth};
  return run_icing(op, nx, ny, nfields, in, icing, fDefined, undef);
}
int fcb200_vesselIcingModStall(int nx, int ny, const float* sal, const float* wave, const float* x_wind, const float* y_wind, const float* airtemp,
                               const float* rh, const float* sst, const float* p, const float* Pw, const float* aice, const float* depth, float vs,
                               float alpha, float zmin, float zmax, float* icing, int* fDefined, float undef)
{
  return fcb200_vesselIcingModStall_batched(nx, ny, 1, sal, wave, x_wind, y_wind, airtemp, rh, sst, p, Pw, aice, depth, vs, alpha, zmin, zmax, icing,
                                            fDefined, undef);
}

int fcb200_vesselIcingMincog_batched(int nx, int ny, int nfields, const float* sal, const float* wave, const float* x_wind, const float* y_wind,
                                     const float* airtemp, const float* rh, const float* sst, const float* p, const float* Pw, const float* aice,
                                     const float* depth, float vs, float alpha, float zmin, float zmax, int alt, float* icing, int* fDefined,
                                     float undef)
{ // VI.cc:688-689
  if (vs < 0 || alpha < 0 || zmin < 0 || zmax < 0 || zmax < zmin || fmod(zmax - zmin, 1) != 0)
    return 0;
  MincogOp op;
  const float beta = alpha;
  op.vs = vs;
  op.cos_alpha = (float)cos((double)alpha);                      // VI.cc:510
  op.sin_beta = (float)sin((double)beta);                        // VI.cc:518
  op.cos_beta_d = cos((double)beta);                             // VI.cc:521
  const float beta_deg = (float)(beta * (180 / M_PI));           // VI.cc:574
  op.drag = (float)(-0.0046 * beta_deg + 2.1912);                // VI.cc:575
  op.Vf = (float)((1 + cos(85 * M_PI / 180)) / 2);               // VI.cc:608
  op.zmin = zmin;
  const float num = zmax - zmin;
  op.number = (int)(num * 2 + 1);                                // VI.cc:618-619
  op.alt = alt;
  int it = (int)log2f(((float)1.3 - (float)-0.5) / (float)1e-5); // VI.cc:391
  op.iterations = it < 100 ? it : 100;
  const float* in[11] = {sal, wave, x_wind, y_wind, airtemp, rh, sst, p, Pw, aice, depth};
  return run_icing(op, nx, ny, nfields, in, icing, fDefined, undef);
}
int fcb200_vesselIcingMincog(int nx, int ny, const float* sal, const float* wave, const float* x_wind, const float* y_wind, const float* airtemp,
                             const float* rh, const float* sst, const float* p, const float* Pw, const float* aice, const float* depth, float vs,
                             float alpha, float zmin, float zmax, int alt, float* icing, int* fDefined, float undef)
{
  return fcb200_vesselIcingMincog_batched(nx, ny, 1, sal, wave, x_wind, y_wind, airtemp, rh, sst, p, Pw, aice, depth, vs, alpha, zmin, zmax, alt,
                                          icing, fDefined, undef);
}

} // extern "C"

// ops_arith.cu -- field arithmetic, element functions and the remaining pressure-level siblings
// (SURVEY.md 8f rank 1): fieldOPERconstant, constantOPERfield, sumFields, min / max / abs / log / exp / pow,
// replaceUndefined / replaceDefined, values2classes, vectorabs, pressure2FlightLevel, snow_in_cm, plevelthe,
// pleveldz2tmean, plevelducting.  All point-wise: functors for the batched engine in elementwise.cuh, same rules
// as ops_elementwise.cu (reference expression types, no FMA contraction, IEEE division and sqrt).
//
// Flag rules follow the reference's four generic loops (FC.cc:94-179): unaryFunctionField and
// binaryFunctionFieldField write undef for undefined input and leave the flag alone; their ...Undef variants
// count and recompute it.  logf / log10f / expf / powf and the double pow / exp are CUDA's (1 - 4 ulp).
// Citations: FC.cc = the reference's src/mi_fieldcalc/FieldCalculations.cc, MC.h = MetConstants.h.
#include "ew_host.cuh"

#include "../../include/fcb200.h"

#include <algorithm>
#include <vector>

namespace fcb200 {
namespace {

using dev::is_def;
using dev::K_CP;
using dev::K_T0;
using dev::K_XLH;

__device__ __forceinline__ float std_min(float a, float b) { return (b < a) ? b : a; } // std::min(a, b)
__device__ __forceinline__ float std_max(float a, float b) { return (a < b) ? b : a; } // std::max(a, b)

enum UnaryCode {
  U_ADDC, U_SUBC, U_MULC, U_DIVC,  // field (op) constant, FC.cc:2627-2645
  U_CADD, U_CSUB, U_CMUL,          // constant (op) field, FC.cc:2647-2660
  U_MINC, U_MAXC,                  // FC.cc:2507-2529
  U_ABS, U_LOG10, U_POW10, U_LOG, U_EXP, U_POWC, // FC.cc:2531-2563
  U_FILL                           // fillUndef / std::fill: the output is `value` everywhere
};

// unaryFunctionField (FC.cc:94-122): undefined -> undef, flag untouched
template <int CODE>
struct UnaryOp
{
  static constexpr int NIN = 1, NOUT = 1, UNROLL = 4;
  static constexpr int NCOUNT = 0;
  static constexpr int MIN_BLOCKS = 5;
  static constexpr bool HEAVY = false;
  static constexpr bool USES_EWT = false, USES_POW = false;
  float value;
  template <bool ALL>
  __device__ __forceinline__ void point(const float* in, float* out, const PointCtx& c, long long, unsigned*) const
  {
    const float a = in[0];
    if (CODE == U_FILL) {
      out[0] = value;
      return;
    }
    if (!(ALL || is_def(a, c.undef))) {
      out[0] = c.undef;
      return;
    }
    float r;
    switch (CODE) {
    case U_ADDC: r = a + value; break;
    case U_SUBC: r = a - value; break;
    case U_MULC: r = a * value; break;
    case U_DIVC: r = a / value; break;
    case U_CADD: r = value + a; break;
    case U_CSUB: r = value - a; break;
    case U_CMUL: r = value * a; break;
    case U_MINC: r = std_min(a, value); break;
    case U_MAXC: r = std_max(a, value); break;
    case U_ABS: r = fabsf(a); break;
    case U_LOG10: r = log10f(a); break;
    case U_POW10: r = (float)pow(10.0, (double)a); break; // math_util.h:121-125: std::pow(10, float) is the double pow
    case U_LOG: r = logf(a); break;
    case U_EXP: r = expf(a); break;
    default: r = powf(a, value); break;
    }
    out[0] = r;
  }
};

enum BinaryCode { B_MIN, B_MAX, B_DZ2TMEAN, B_DUCT_Q };

// binaryFunctionFieldField (FC.cc:126-140): undefined -> undef, flag untouched
template <int CODE>
struct BinaryOp
{
  static constexpr int NIN = 2, NOUT = 1, UNROLL = 4;
  static constexpr int NCOUNT = 0;
  static constexpr int MIN_BLOCKS = 5;
  static constexpr bool HEAVY = false;
  static constexpr bool USES_EWT = false, USES_POW = false;
  float s0, s1; // B_DZ2TMEAN: convert, tconvert;  B_DUCT_Q: tconv, p
  template <bool ALL>
  __device__ __forceinline__ void point(const float* in, float* out, const PointCtx& c, long long, unsigned*) const
  {
    const float a = in[0], b = in[1];
    if (!(ALL || (is_def(a, c.undef) && is_def(b, c.undef)))) {
      out[0] = c.undef;
      return;
    }
    switch (CODE) {
    case B_MIN: out[0] = std_min(a, b); break;
    case B_MAX: out[0] = std_max(a, b); break;
    case B_DZ2TMEAN: out[0] = (a - b) * s0 + s1; break; // FC.cc:501
    default: out[0] = dev::tk_q_duct(a * s0, b, s1); break; // FC.cc:626
    }
  }
};

// constant / field: divideUndef (FC.cc:84-92, 2661-2664)
struct ConstDivFieldOp
{
  static constexpr int NIN = 1, NOUT = 1, UNROLL = 4;
  static constexpr int NCOUNT = 1;
  static constexpr int MIN_BLOCKS = 5;
  static constexpr bool HEAVY = false;
  static constexpr bool USES_EWT = false, USES_POW = false;
  float value;
  template <bool ALL>
  __device__ __forceinline__ void point(const float* in, float* out, const PointCtx& c, long long, unsigned* nundef) const
  {
    if ((ALL || is_def(in[0], c.undef)) && in[0] != 0)
      out[0] = value / in[0];
    else {
      out[0] = c.undef;
      nundef[0] += 1;
    }
  }
};

// replaceUndefined / replaceDefined on a SOME_DEFINED field (FC.cc:2581, 2604): the test is `== undef` only
template <bool REPLACE_DEFINED>
struct ReplaceOp
{
  static constexpr int NIN = 1, NOUT = 1, UNROLL = 4;
  static constexpr int NCOUNT = 0;
  static constexpr int MIN_BLOCKS = 5;
  static constexpr bool HEAVY = false;
  static constexpr bool USES_EWT = false, USES_POW = false;
  float value;
  template <bool ALL>
  __device__ __forceinline__ void point(const float* in, float* out, const PointCtx& c, long long, unsigned*) const
  {
    const bool is_undef = in[0] == c.undef;
    out[0] = (is_undef != REPLACE_DEFINED) ? value : in[0];
  }
};

// plevelthe, FC.cc:369-398 with tk_rh_the (FC.cc:269-278)
struct PlevelTheOp
{
  static constexpr int NIN = 2, NOUT = 1, UNROLL = 2;
  static constexpr int NCOUNT = 1;
  static constexpr int MIN_BLOCKS = 4;
  static constexpr bool HEAVY = false;
  static constexpr bool USES_EWT = true, USES_POW = false;
  float tconv, cvrh, thconv;
  template <bool ALL>
  __device__ __forceinline__ void point(const float* in, float* out, const PointCtx& c, long long, unsigned* nundef) const
  {
    bool ok = ALL || (is_def(in[0], c.undef) && is_def(in[1], c.undef));
    float r = c.undef;
    if (ok) {
      const float tk = in[0] * tconv, rh = in[1] * cvrh;
      const dev::Ewt e(tk - K_T0);
      ok = e.defined;
      if (ok)
        r = tk * thconv + e.value(c.tab) * rh;
    }
    out[0] = r;
    nundef[0] += ok ? 0u : 1u;
  }
};

// plevelducting compute 3 / 4, FC.cc:627-630 with tk_rh_duct (FC.cc:285-296)
struct PlevelDuctRhOp
{
  static constexpr int NIN = 2, NOUT = 1, UNROLL = 2;
  static constexpr int NCOUNT = 1;
  static constexpr int MIN_BLOCKS = 4;
  static constexpr bool HEAVY = false;
  static constexpr bool USES_EWT = true, USES_POW = false;
  float tconv, p;
  template <bool ALL>
  __device__ __forceinline__ void point(const float* in, float* out, const PointCtx& c, long long, unsigned* nundef) const
  {
    bool ok = ALL || (is_def(in[0], c.undef) && is_def(in[1], c.undef));
    float r = c.undef;
    if (ok)
      ok = dev::tk_rh_duct(c.tab, in[0] * tconv, in[1], p, r);
    out[0] = ok ? r : c.undef;
    nundef[0] += ok ? 0u : 1u;
  }
};

// vectorabs, FC.cc:1819-1841
struct VectorAbsOp
{
  static constexpr int NIN = 2, NOUT = 1, UNROLL = 4;
  static constexpr int NCOUNT = 1;
  static constexpr int MIN_BLOCKS = 5;
  static constexpr bool HEAVY = false;
  static constexpr bool USES_EWT = false, USES_POW = false;
  template <bool ALL>
  __device__ __forceinline__ void point(const float* in, float* out, const PointCtx& c, long long, unsigned* nundef) const
  {
    if (ALL || (is_def(in[0], c.undef) && is_def(in[1], c.undef)))
      out[0] = dev::absval(in[0], in[1]);
    else {
      out[0] = c.undef;
      nundef[0] += 1;
    }
  }
};

// pressure2FlightLevel, FC.cc:2311-2349; tables MC.h:87-89.  The level tables are read with a data-dependent index: from
// constant memory that serialises per distinct index in a warp (0.22 of the roofline on white-noise pressures), so they are
// staged in shared memory with the saturation table (USES_EWT = the engine's "stage the tables" switch).
struct FlightLevelOp
{
  static constexpr int NIN = 1, NOUT = 1, UNROLL = 2;
  static constexpr int NCOUNT = 1;
  static constexpr int MIN_BLOCKS = 4;
  static constexpr bool HEAVY = false;
  static constexpr bool USES_EWT = true, USES_POW = false;
  template <bool ALL>
  __device__ __forceinline__ void point(const float* in, float* out, const PointCtx& c, long long, unsigned* nundef) const
  {
    if (ALL || is_def(in[0], c.undef)) {
      constexpr int nTab = 15;
      float p = in[0];
      if (p > 1000.f)
        p = 1000.f;
      if (p < 10.f)
        p = 10.f;
      // k = 1 + the number of levels 1 .. 14 above p: the table decreases, so the reference's walk `while (k < nTab &&
      // pLevelTable[k] > p) k++` stops exactly there (a NaN under ALL_DEFINED compares false everywhere: k = 1, as on the CPU)
      int k = 1;
#pragma unroll
      for (int m = 1; m < nTab; ++m)
        k += (c.tab.level[m].x > p) ? 1 : 0;
      const float2 lo = c.tab.level[k - 1], hi = c.tab.level[k];
      const float ratio = (p - lo.x) / (hi.x - lo.x);
      out[0] = lo.y + (hi.y - lo.y) * ratio;
    } else {
      out[0] = c.undef;
      nundef[0] += 1;
    }
  }
};

// values2classes, FC.cc:2462-2499: at most MAX_CLASS_VALUES limits, passed by value in the functor
constexpr int MAX_CLASS_VALUES = 64;
struct ClassesOp
{
  static constexpr int NIN = 1, NOUT = 1, UNROLL = 2;
  static constexpr int NCOUNT = 1;
  static constexpr int MIN_BLOCKS = 4;
  static constexpr bool HEAVY = false;
  static constexpr bool USES_EWT = false, USES_POW = false;
  int nvalues; // size - 2
  float fmin, fmax;
  float values[MAX_CLASS_VALUES]; // the limits, by value (kernel parameter space: an indexed constant-bank load)
  template <bool ALL>
  __device__ __forceinline__ void point(const float* in, float* out, const PointCtx& c, long long, unsigned* nundef) const
  {
    const float v = in[0];
    if ((ALL || is_def(v, c.undef)) && v >= fmin && v < fmax) {
      int j = 1;
      while (j < nvalues && values[j] < v)
        j++;
      out[0] = (float)(j - 1);
    } else {
      out[0] = c.undef;
      nundef[0] += 1;
    }
  }
};

// snow_in_cm, FC.cc:3063-3118: every line is a double expression rounded to float on assignment
struct SnowCmOp
{
  static constexpr int NIN = 3, NOUT = 1, UNROLL = 1;
  static constexpr int NCOUNT = 1;
  static constexpr int MIN_BLOCKS = 4;
  static constexpr bool HEAVY = false;
  static constexpr bool USES_EWT = false, USES_POW = false;
  template <bool ALL>
  __device__ __forceinline__ void point(const float* in, float* out, const PointCtx& c, long long, unsigned* nundef) const
  {
    if (ALL || (is_def(in[0], c.undef) && is_def(in[1], c.undef) && is_def(in[2], c.undef))) {
      const float sw = in[0];
      if ((double)sw <= 0.) {
        out[0] = 0.f;
        return;
      }
      const float t = (float)((double)(in[1] + in[2]) / 2.);
      const double ex = exp(((double)t - 274.3) * 3.5);
      const float logit_t = (float)((1. - ex) / (1. + ex));
      const double d = ((double)t - 252.0) / 20.0;
      const float mm2cm_t = (float)(0.13 / (0.02 + 0.1 * d * d));
      const float fac = logit_t * mm2cm_t;
      out[0] = ((double)fac <= 1.) ? sw : sw * fac;
    } else {
      out[0] = c.undef;
      nundef[0] += 1;
    }
  }
};

// sumFields, FC.cc:2671-2694: members added in order into a float accumulator; the first undefined member ends the point
constexpr int MAX_SUM_FIELDS = 64;
struct SumArgs
{
  const float* f[MAX_SUM_FIELDS];
  int nfields;
  long long n;
  float undef;
  int all;
  float* out;
  unsigned long long* counter;
};

// `table` = nullptr: the member pointers travel by value in the kernel parameters (up to MAX_SUM_FIELDS);
// otherwise they are read from a device table (any number of fields)
__global__ void __launch_bounds__(256) sum_fields_kernel(const SumArgs a, const float* const* __restrict__ table)
{
  unsigned nundef = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < a.n; i += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    bool ok = true;
    if (table == nullptr) {
      for (int j = 0; j < a.nfields; ++j) {
        const float v = a.f[j][i];
        if (ok && (a.all || is_def(v, a.undef)))
          s += v;
        else
          ok = false;
      }
    } else {
      for (int j = 0; j < a.nfields && ok; ++j) {
        const float v = table[j][i];
        if (a.all || is_def(v, a.undef))
          s += v;
        else
          ok = false;
      }
    }
    a.out[i] = ok ? s : a.undef;
    nundef += ok ? 0u : 1u;
  }
  const unsigned total = __reduce_add_sync(0xffffffffu, nundef);
  if ((threadIdx.x & 31) == 0 && total)
    atomicAdd(a.counter, (unsigned long long)total);
}

// values2classes with more than MAX_CLASS_VALUES limits: the same point function, limits in device memory
__global__ void __launch_bounds__(256) classes_table_kernel(const float* __restrict__ in, float* __restrict__ out, long long n, const float* __restrict__ values,
                                                            int nvalues, float fmin, float fmax, float undef, const FieldMeta* __restrict__ meta,
                                                            unsigned long long* counters)
{
  const int field = blockIdx.y;
  const bool all = meta[field].all != 0;
  const float* src = in + (long long)field * n;
  float* dst = out + (long long)field * n;
  unsigned nundef = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = src[i];
    if ((all || is_def(v, undef)) && v >= fmin && v < fmax) {
      int j = 1;
      while (j < nvalues && values[j] < v)
        j++;
      dst[i] = (float)(j - 1);
    } else {
      dst[i] = undef;
      nundef += 1;
    }
  }
  const unsigned total = __reduce_add_sync(0xffffffffu, nundef);
  if ((threadIdx.x & 31) == 0 && total)
    atomicAdd(counters + field, (unsigned long long)total);
}

template <class Op>
int run_unary(int nx, int ny, int nfields, const Op& op, const float* field, float* fres, int* fDefined, float undef, FlagRule rule)
{
  const float* in[1] = {field};
  const int pf[1] = {1};
  return run_elementwise(make_batch(nx, ny, nfields), op, in, pf, fres, fDefined, undef, rule, NoMeta());
}

template <class Op>
int run_binary(int nx, int ny, int nfields, const Op& op, const float* f1, const float* f2, float* fres, int* fDefined, float undef, FlagRule rule)
{
  const float* in[2] = {f1, f2};
  const int pf[2] = {1, 1};
  return run_elementwise(make_batch(nx, ny, nfields), op, in, pf, fres, fDefined, undef, rule, NoMeta());
}

// fillUndef (FC.cc:76-82) / std::fill: the input is not read by the arithmetic but keeps the engine's shape
int fill_field(int nx, int ny, int nfields, const float* any_field, float value, float* fres, int* fDefined, float undef, int new_flag)
{
  const int rc = run_unary(nx, ny, nfields, UnaryOp<U_FILL>{value}, any_field, fres, fDefined, undef, FLAG_UNCHANGED);
  if (rc == 1)
    for (int k = 0; k < nfields; ++k)
      fDefined[k] = new_flag;
  return rc;
}

template <int CODE>
int unary_entry(int nx, int ny, int nfields, const float* field, float value, float* fres, int* fDefined, float undef)
{
  return run_unary(nx, ny, nfields, UnaryOp<CODE>{value}, field, fres, fDefined, undef, FLAG_UNCHANGED);
}

} // namespace
} // namespace fcb200

// =========================================================================================== C-ABI
using namespace fcb200;

extern "C" {

// ---- pressure-level siblings
int fcb200_plevelthe_batched(int nx, int ny, int nfields, const float* t, const float* rh, float p, int compute, float* the, int* fDefined, float undef)
{ // FC.cc:369-398
  if (compute != 1 && compute != 2)
    return 0;
  if (p <= 0.0)
    return 0;
  const float pidcp = host_pidcp(p), pi = pidcp * H_CP;
  PlevelTheOp op;
  op.cvrh = (float)(0.01 * (double)((float)2.501e+6 / pi) * (double)(float)0.622 / (double)p);
  op.tconv = (compute == 2) ? pidcp : 1.f;
  op.thconv = 1 / pidcp;
  return run_binary(nx, ny, nfields, op, t, rh, the, fDefined, undef, FLAG_FROM_COUNT);
}
int fcb200_plevelthe(int nx, int ny, const float* t, const float* rh, float p, int compute, float* the, int* fDefined, float undef)
{
  return fcb200_plevelthe_batched(nx, ny, 1, t, rh, p, compute, the, fDefined, undef);
}

int fcb200_pleveldz2tmean_batched(int nx, int ny, int nfields, const float* z1, const float* z2, float p1, float p2, int compute, float* tmean,
                                  int* fDefined, float undef)
{ // FC.cc:466-503
  if (p1 <= 0 || p2 <= 0 || p1 == p2)
    return 0;
  const float g = (float)9.8;
  const float pi1 = H_CP * host_pidcp(p1), pi2 = H_CP * host_pidcp(p2);
  float convert, tconvert;
  switch (compute) {
  case 1:
    convert = (float)((double)g * 0.5 * (double)(pi1 + pi2) / (double)((pi2 - pi1) * H_CP));
    tconvert = -H_T0;
    break;
  case 2:
    convert = (float)((double)g * 0.5 * (double)(pi1 + pi2) / (double)((pi2 - pi1) * H_CP));
    tconvert = 0.f;
    break;
  case 3:
    convert = g / (pi2 - pi1);
    tconvert = 0.f;
    break;
  default:
    return 0;
  }
  return run_binary(nx, ny, nfields, BinaryOp<B_DZ2TMEAN>{convert, tconvert}, z1, z2, tmean, fDefined, undef, FLAG_UNCHANGED);
}
int fcb200_pleveldz2tmean(int nx, int ny, const float* z1, const float* z2, float p1, float p2, int compute, float* tmean, int* fDefined, float undef)
{
  return fcb200_pleveldz2tmean_batched(nx, ny, 1, z1, z2, p1, p2, compute, tmean, fDefined, undef);
}

int fcb200_plevelducting_batched(int nx, int ny, int nfields, const float* t, const float* h, float p, int compute, float* duct, int* fDefined,
                                 float undef)
{ // FC.cc:597-636
  if (p <= 0)
    return 0;
  const float tconv = (compute % 2 == 0) ? host_pidcp(p) : 1.f;
  if (compute == 1 || compute == 2)
    return run_binary(nx, ny, nfields, BinaryOp<B_DUCT_Q>{tconv, p}, t, h, duct, fDefined, undef, FLAG_UNCHANGED);
  if (compute == 3 || compute == 4)
    return run_binary(nx, ny, nfields, PlevelDuctRhOp{tconv, p}, t, h, duct, fDefined, undef, FLAG_FROM_COUNT);
  return 0;
}
int fcb200_plevelducting(int nx, int ny, const float* t, const float* h, float p, int compute, float* duct, int* fDefined, float undef)
{
  return fcb200_plevelducting_batched(nx, ny, 1, t, h, p, compute, duct, fDefined, undef);
}

// ---- level independent
int fcb200_vectorabs_batched(int nx, int ny, int nfields, const float* u, const float* v, float* ff, int* fDefined, float undef)
{ // FC.cc:1819-1841
  return run_binary(nx, ny, nfields, VectorAbsOp(), u, v, ff, fDefined, undef, FLAG_FROM_COUNT);
}
int fcb200_vectorabs(int nx, int ny, const float* u, const float* v, float* ff, int* fDefined, float undef)
{
  return fcb200_vectorabs_batched(nx, ny, 1, u, v, ff, fDefined, undef);
}

int fcb200_pressure2FlightLevel_batched(int nx, int ny, int nfields, const float* pressure, float* flightlevel, int* fDefined, float undef)
{ // FC.cc:2311-2349
  return run_unary(nx, ny, nfields, FlightLevelOp(), pressure, flightlevel, fDefined, undef, FLAG_FROM_COUNT);
}
int fcb200_pressure2FlightLevel(int nx, int ny, const float* pressure, float* flightlevel, int* fDefined, float undef)
{
  return fcb200_pressure2FlightLevel_batched(nx, ny, 1, pressure, flightlevel, fDefined, undef);
}

int fcb200_values2classes_batched(int nx, int ny, int nfields, const float* fvalue, float* fclass, const float* values, int nvalues, int* fDefined,
                                  float undef)
{ // FC.cc:2462-2499
  if (nvalues < 2)
    return 0;
  if (nvalues > MAX_CLASS_VALUES) { // more limits than fit the functor: the table kernel
    const long long n = (long long)nx * ny;
    if (nx <= 0 || ny <= 0 || nfields <= 0 || nfields > 65535 || n >= 0x7fffffffLL) {
      set_error("fcb200: invalid grid or batch size (nx=%d ny=%d nfields=%d)", nx, ny, nfields);
      return -1;
    }
    Call call;
    const float* d_in = call.in(fvalue, (size_t)n * nfields);
    float* d_out = call.out(fclass, (size_t)n * nfields);
    const float* d_values = static_cast<const float*>(call.upload_small(values, sizeof(float) * (size_t)nvalues));
    FieldMeta* meta = call.meta_host(nfields);
    if (!call.ok())
      return -1;
    for (int k = 0; k < nfields; ++k) {
      meta[k].all = (fDefined[k] == ALL_DEFINED) ? 1 : 0;
      meta[k].a = meta[k].b = meta[k].c = 0.f;
    }
    const FieldMeta* d_meta = call.upload_meta();
    unsigned long long* counters = call.counters(nfields);
    if (!call.ok())
      return -1;
    long long blocks = (n + 255) / 256;
    const long long cap = std::max(1LL, (long long)sm_count() * 16 / nfields);
    if (blocks > cap)
      blocks = cap;
    classes_table_kernel<<<dim3((unsigned)blocks, (unsigned)nfields), 256, 0, call.stream()>>>(d_in, d_out, n, d_values, nvalues - 2, values[0], values[nvalues - 1],
                                                                                                undef, d_meta, counters);
    count_launch();
    const unsigned long long un = (unsigned long long)n;
    return call.finish([=](const unsigned long long* cnt) {
      for (int k = 0; k < nfields; ++k)
        fDefined[k] = check_defined(cnt[k], un);
    });
  }
  ClassesOp op;
  op.nvalues = nvalues - 2;
  op.fmin = values[0];
  op.fmax = values[nvalues - 1];
  for (int j = 0; j < MAX_CLASS_VALUES; ++j)
    op.values[j] = (j < nvalues) ? values[j] : 0.f;
  return run_unary(nx, ny, nfields, op, fvalue, fclass, fDefined, undef, FLAG_FROM_COUNT);
}
int fcb200_values2classes(int nx, int ny, const float* fvalue, float* fclass, const float* values, int nvalues, int* fDefined, float undef)
{
  return fcb200_values2classes_batched(nx, ny, 1, fvalue, fclass, values, nvalues, fDefined, undef);
}

// ---- min / max / abs / log / exp / pow: `void` in the reference (FC.h:254-272), 1 here
int fcb200_minvalueFields_batched(int nx, int ny, int nfields, const float* field1, const float* field2, float* fres, int* fDefined, float undef)
{
  return run_binary(nx, ny, nfields, BinaryOp<B_MIN>{0.f, 0.f}, field1, field2, fres, fDefined, undef, FLAG_UNCHANGED);
}
int fcb200_minvalueFields(int nx, int ny, const float* field1, const float* field2, float* fres, int* fDefined, float undef)
{
  return fcb200_minvalueFields_batched(nx, ny, 1, field1, field2, fres, fDefined, undef);
}
int fcb200_maxvalueFields_batched(int nx, int ny, int nfields, const float* field1, const float* field2, float* fres, int* fDefined, float undef)
{
  return run_binary(nx, ny, nfields, BinaryOp<B_MAX>{0.f, 0.f}, field1, field2, fres, fDefined, undef, FLAG_UNCHANGED);
}
int fcb200_maxvalueFields(int nx, int ny, const float* field1, const float* field2, float* fres, int* fDefined, float undef)
{
  return fcb200_maxvalueFields_batched(nx, ny, 1, field1, field2, fres, fDefined, undef);
}

#define FCB_UNARY_CONST(name, CODE)                                                                                                                  \
  int fcb200_##name##_batched(int nx, int ny, int nfields, const float* field, float value, float* fres, int* fDefined, float undef)                \
  {                                                                                                                                                  \
    if (value == undef) /* fillUndef: every point undefined, NONE_DEFINED */                                                                        \
      return fill_field(nx, ny, nfields, field, undef, fres, fDefined, undef, NONE_DEFINED);                                                         \
    return unary_entry<CODE>(nx, ny, nfields, field, value, fres, fDefined, undef);                                                                  \
  }                                                                                                                                                  \
  int fcb200_##name(int nx, int ny, const float* field, float value, float* fres, int* fDefined, float undef)                                       \
  {                                                                                                                                                  \
    return fcb200_##name##_batched(nx, ny, 1, field, value, fres, fDefined, undef);                                                                  \
  }
FCB_UNARY_CONST(minvalueFieldConst, U_MINC) // FC.cc:2507-2514
FCB_UNARY_CONST(maxvalueFieldConst, U_MAXC) // FC.cc:2522-2529
FCB_UNARY_CONST(powerField, U_POWC)         // FC.cc:2556-2563

#define FCB_UNARY(name, CODE)                                                                                                                        \
  int fcb200_##name##_batched(int nx, int ny, int nfields, const float* field, float* fres, int* fDefined, float undef)                             \
  {                                                                                                                                                  \
    return unary_entry<CODE>(nx, ny, nfields, field, 0.f, fres, fDefined, undef);                                                                    \
  }                                                                                                                                                  \
  int fcb200_##name(int nx, int ny, const float* field, float* fres, int* fDefined, float undef)                                                    \
  {                                                                                                                                                  \
    return fcb200_##name##_batched(nx, ny, 1, field, fres, fDefined, undef);                                                                         \
  }
FCB_UNARY(absvalueField, U_ABS) // FC.cc:2531-2534
FCB_UNARY(log10Field, U_LOG10)  // FC.cc:2536-2539
FCB_UNARY(pow10Field, U_POW10)  // FC.cc:2541-2544
FCB_UNARY(logField, U_LOG)      // FC.cc:2546-2549
FCB_UNARY(expField, U_EXP)      // FC.cc:2551-2554

int fcb200_replaceUndefined_batched(int nx, int ny, int nfields, const float* field, float value, float* fres, int* fDefined, float undef)
{ // FC.cc:2565-2587, field by field: the branch depends on each field's flag
  const size_t n = (size_t)nx * (size_t)ny;
  for (int k = 0; k < nfields; ++k) {
    const float* f = field + k * n;
    float* o = fres + k * n;
    int rc;
    if (value == undef || fDefined[k] == ALL_DEFINED) {
      if (o == f)
        continue;
      rc = run_unary(nx, ny, 1, ReplaceOp<false>{undef}, f, o, fDefined + k, undef, FLAG_UNCHANGED); // undef -> undef: a bit copy
      if (rc != 1)
        return rc;
      continue;
    }
    if (fDefined[k] == NONE_DEFINED)
      rc = fill_field(nx, ny, 1, f, value, o, fDefined + k, undef, ALL_DEFINED);
    else
      rc = run_unary(nx, ny, 1, ReplaceOp<false>{value}, f, o, fDefined + k, undef, FLAG_UNCHANGED);
    if (rc != 1)
      return rc;
    fDefined[k] = ALL_DEFINED;
  }
  return 1;
}
int fcb200_replaceUndefined(int nx, int ny, const float* field, float value, float* fres, int* fDefined, float undef)
{
  return fcb200_replaceUndefined_batched(nx, ny, 1, field, value, fres, fDefined, undef);
}

int fcb200_replaceDefined_batched(int nx, int ny, int nfields, const float* field, float value, float* fres, int* fDefined, float undef)
{ // FC.cc:2589-2609
  const size_t n = (size_t)nx * (size_t)ny;
  for (int k = 0; k < nfields; ++k) {
    const float* f = field + k * n;
    float* o = fres + k * n;
    int rc;
    if (value == undef || fDefined[k] == NONE_DEFINED) {
      rc = fill_field(nx, ny, 1, f, undef, o, fDefined + k, undef, NONE_DEFINED);
      if (rc != 1)
        return rc;
      continue;
    }
    if (fDefined[k] == ALL_DEFINED)
      rc = fill_field(nx, ny, 1, f, value, o, fDefined + k, undef, ALL_DEFINED);
    else
      rc = run_unary(nx, ny, 1, ReplaceOp<true>{value}, f, o, fDefined + k, undef, FLAG_UNCHANGED);
    if (rc != 1)
      return rc;
    fDefined[k] = ALL_DEFINED;
  }
  return 1;
}
int fcb200_replaceDefined(int nx, int ny, const float* field, float value, float* fres, int* fDefined, float undef)
{
  return fcb200_replaceDefined_batched(nx, ny, 1, field, value, fres, fDefined, undef);
}

// ---- field (op) constant, constant (op) field, sum of fields
int fcb200_fieldOPERconstant_batched(int compute, int nx, int ny, int nfields, const float* field, float value, float* fres, int* fDefined,
                                     float undef)
{ // FC.cc:2627-2645
  if ((value == undef) || (compute == 4 && value == 0))
    return fill_field(nx, ny, nfields, field, undef, fres, fDefined, undef, NONE_DEFINED);
  switch (compute) {
  case 1:
    return unary_entry<U_ADDC>(nx, ny, nfields, field, value, fres, fDefined, undef);
  case 2:
    return unary_entry<U_SUBC>(nx, ny, nfields, field, value, fres, fDefined, undef);
  case 3:
    return unary_entry<U_MULC>(nx, ny, nfields, field, value, fres, fDefined, undef);
  case 4:
    return unary_entry<U_DIVC>(nx, ny, nfields, field, value, fres, fDefined, undef);
  default:
    return 0;
  }
}
int fcb200_fieldOPERconstant(int compute, int nx, int ny, const float* field, float value, float* fres, int* fDefined, float undef)
{
  return fcb200_fieldOPERconstant_batched(compute, nx, ny, 1, field, value, fres, fDefined, undef);
}

int fcb200_constantOPERfield_batched(int compute, int nx, int ny, int nfields, float value, const float* field, float* fres, int* fDefined,
                                     float undef)
{ // FC.cc:2647-2669
  if (value == undef)
    return fill_field(nx, ny, nfields, field, undef, fres, fDefined, undef, NONE_DEFINED);
  switch (compute) {
  case 1:
    return unary_entry<U_CADD>(nx, ny, nfields, field, value, fres, fDefined, undef);
  case 2:
    return unary_entry<U_CSUB>(nx, ny, nfields, field, value, fres, fDefined, undef);
  case 3:
    return unary_entry<U_CMUL>(nx, ny, nfields, field, value, fres, fDefined, undef);
  case 4:
    return run_unary(nx, ny, nfields, ConstDivFieldOp{value}, field, fres, fDefined, undef, FLAG_FROM_COUNT);
  default:
    return 0;
  }
}
int fcb200_constantOPERfield(int compute, int nx, int ny, float value, const float* field, float* fres, int* fDefined, float undef)
{
  return fcb200_constantOPERfield_batched(compute, nx, ny, 1, value, field, fres, fDefined, undef);
}

int fcb200_sumFields(int nx, int ny, const float* const* fields, int nfields, float* fres, int* fDefined, float undef)
{ // FC.cc:2671-2694
  const long long n = (long long)nx * ny;
  if (nx <= 0 || ny <= 0 || n >= 0x7fffffffLL || nfields < 0) {
    set_error("fcb200: invalid grid (nx=%d ny=%d nfields=%d)", nx, ny, nfields);
    return -1;
  }
  Call call;
  if (!call.ok())
    return -1;
  SumArgs a;
  const float* const* d_table = nullptr;
  if (nfields <= MAX_SUM_FIELDS) {
    for (int j = 0; j < nfields; ++j)
      a.f[j] = call.in(fields[j], (size_t)n);
  } else { // any number of fields: the pointer table goes to device memory
    std::vector<const float*> ptrs((size_t)nfields);
    for (int j = 0; j < nfields; ++j)
      ptrs[j] = call.in(fields[j], (size_t)n);
    d_table = static_cast<const float* const*>(call.upload_small(ptrs.data(), sizeof(float*) * (size_t)nfields));
  }
  a.nfields = nfields;
  a.n = n;
  a.undef = undef;
  a.all = (*fDefined == ALL_DEFINED) ? 1 : 0;
  a.out = call.out(fres, (size_t)n);
  a.counter = call.counters(1);
  if (!call.ok())
    return -1;
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap)
    blocks = cap;
  sum_fields_kernel<<<(unsigned)blocks, 256, 0, call.stream()>>>(a, d_table);
  count_launch();
  const unsigned long long un = (unsigned long long)n;
  return call.finish([=](const unsigned long long* cnt) { *fDefined = check_defined(cnt[0], un); });
}
int fcb200_sumFields_batched(int nx, int ny, int ntimes, const float* const* fields, int nfields, float* fres, int* fDefined, float undef)
{ // member j is a dense [ntimes][ny][nx] array (the ensemble operators' batched layout)
  const size_t n = (size_t)nx * (size_t)ny;
  std::vector<const float*> f((size_t)(nfields > 0 ? nfields : 0));
  for (int t = 0; t < ntimes; ++t) {
    for (int j = 0; j < nfields; ++j)
      f[j] = fields[j] + t * n;
    const int rc = fcb200_sumFields(nx, ny, f.data(), nfields, fres + t * n, fDefined + t, undef);
    if (rc != 1)
      return rc;
  }
  return 1;
}

int fcb200_snow_in_cm_batched(int nx, int ny, int nfields, const float* snow_water, const float* tk2m, const float* td2m, float* snow_cm, int* fDefined,
                              float undef)
{ // FC.cc:3063-3118
  const float* in[3] = {snow_water, tk2m, td2m};
  const int pf[3] = {1, 1, 1};
  return run_elementwise(make_batch(nx, ny, nfields), SnowCmOp(), in, pf, snow_cm, fDefined, undef, FLAG_FROM_COUNT, NoMeta());
}
int fcb200_snow_in_cm(int nx, int ny, const float* snow_water, const float* tk2m, const float* td2m, float* snow_cm, int* fDefined, float undef)
{
  return fcb200_snow_in_cm_batched(nx, ny, 1, snow_water, tk2m, td2m, snow_cm, fDefined, undef);
}

} // extern "C"

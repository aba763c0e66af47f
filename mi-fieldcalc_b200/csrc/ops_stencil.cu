// ops_stencil.cu -- the five-point map-ratio stencil family, thermalFrontParameter and the
// shapiro2_filter smoother.
//
// The reference writes every stencil as ONE flat loop over i in [nx, N-nx) -- edge columns included,
// with their wrapped i-1 / i+1 neighbours -- counts undefined points there, derives the flag from
// N - 2*nx, and then overwrites the border ring with fillEdges (FC.cc:59-74), which amounts to
//     out(x, y) = interior(clamp(x, 1, nx-2), clamp(y, 1, ny-2)).
// Here one kernel does all of that: interior points compute and store their value AND the border
// cells that clamp onto them; edge-column points only evaluate the definedness test (they count
// towards n_undefined although their values never survive -- SURVEY.md 7, hard part 1).
//
// Citations: FC.cc = the reference's src/mi_fieldcalc/FieldCalculations.cc.
#include "device_common.cuh"
#include "stencil_tile.cuh"
#include "tfp_tile.cuh"

#include "../../include/fcb200.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace fcb200 {
namespace {

using dev::is_def;

constexpr int ST_THREADS = 256;
constexpr int ST_UNROLL = 4;

struct StencilGeom
{
  int nx, ny;
  int n; // nx*ny < 2^31 (the reference's `int fsize`)
  int nfields;
  int chunks;
  int group;      // fields per grid group (gridDim.x)
  int chunk_base; // first chunk of this launch (a range of more than 65535 chunks takes several launches)
  int lo, hi; // the reference's flat loop range [lo, hi): [nx, N-nx), or [1, N-1) for gradient c=1
  float undef;
  const FieldMeta* meta;
  unsigned long long* counters;
};

// Op interface (every member is evaluated per thread; pointers are per-field after at()):
//   static constexpr int NOUT;
//   static constexpr bool TESTS_WHEN_ALL;        // eval() can fail even when allDefined (TFP's absdelt != 0)
//   __device__ bool all_defined(int field, bool in_all) const;   // the allDefined of this pass
//   __device__ Op at(int field, int n) const;                    // the same operator with its field pointers advanced
//   template <bool ALL> struct In;                               // the operands of ONE point, in registers
//   template <bool ALL> __device__ In<ALL> load(int i, int nx) const;
//   template <bool ALL> __device__ bool eval(const In<ALL>&, float undef, float* val) const;
//   __device__ float* out(int k) const;
//
// The main kernel is the reference's flat loop and nothing else: a thread owns ST_UNROLL points one
// CTA-width apart (every warp access is a contiguous 128 B run), issues ALL its loads, then computes
// and stores.  No (x, y) is ever computed; edge-column points are evaluated with their wrapped
// neighbours exactly like the reference does, and fill_edges_kernel overwrites the border ring
// afterwards (stream order), exactly like fillEdges.  Undefined points are counted in a register and
// flushed with one warp reduction + one atomic per warp.
template <class Op, bool ALL>
__device__ __forceinline__ unsigned stencil_point(const Op& op, const typename Op::template In<ALL>& in, int i, float undef)
{
  float val[Op::NOUT];
  const bool ok = op.template eval<ALL>(in, undef, val);
#pragma unroll
  for (int k = 0; k < Op::NOUT; ++k)
    op.out(k)[i] = ok ? val[k] : undef;
  return ok ? 0u : 1u;
}

template <class Op, bool ALL>
__device__ __forceinline__ void stencil_body(const Op& op, const StencilGeom& g, int chunk, unsigned long long* counter)
{
  typedef typename Op::template In<ALL> In;
  const int nx = g.nx, hi = g.hi;
  const int c0 = g.lo + chunk * (ST_THREADS * ST_UNROLL);
  const int i0 = c0 + (int)threadIdx.x;
  unsigned nundef = 0;
  if (c0 + ST_THREADS * ST_UNROLL <= hi) {
    // full chunk (all but the last of a field): no guards, every address is a constant offset from i0
    In in[ST_UNROLL];
#pragma unroll
    for (int u = 0; u < ST_UNROLL; ++u)
      in[u] = op.template load<ALL>(i0 + u * ST_THREADS, nx);
#pragma unroll
    for (int u = 0; u < ST_UNROLL; ++u)
      nundef += stencil_point<Op, ALL>(op, in[u], i0 + u * ST_THREADS, g.undef);
  } else {
#pragma unroll 1
    for (int i = i0; i < hi; i += ST_THREADS)
      nundef += stencil_point<Op, ALL>(op, op.template load<ALL>(i, nx), i, g.undef);
  }
  // one global atomic per CTA: all CTAs in flight belong to one or two fields, and per-warp atomics on
  // a single address serialise in L2 (measured: 2.3x slower with 30 % undefined points)
  if (!ALL || Op::TESTS_WHEN_ALL) {
    __shared__ unsigned s_count;
    if (threadIdx.x == 0)
      s_count = 0;
    __syncthreads();
    nundef = __reduce_add_sync(0xffffffffu, nundef);
    if ((threadIdx.x & 31) == 0 && nundef)
      atomicAdd(&s_count, nundef);
    __syncthreads();
    if (threadIdx.x == 0 && s_count)
      atomicAdd(counter, (unsigned long long)s_count);
  }
}

// Grid = (ST_GROUP fields, chunks, field groups), x fastest: the CTAs in flight work on the SAME few
// chunks of up to ST_GROUP different fields, so a chunk of the grid-constant arrays (xmapr, ymapr,
// fcoriolis) is fetched from HBM once per group and then hit in L2.  Field-major order thrashes: one
// ECMWF field streams 26 MB in + 26 MB out + 52 MB of map ratios, more than the L2 keeps (measured
// 2.9x DRAM read amplification).
constexpr int ST_GROUP = 16;

template <class Op>
__global__ void __launch_bounds__(ST_THREADS) stencil_kernel(const Op op0, const StencilGeom g)
{
  const int field = blockIdx.z * g.group + blockIdx.x, chunk = blockIdx.y + g.chunk_base;
  if (field >= g.nfields)
    return;
  const bool all = op0.all_defined(field, g.meta[field].all != 0);
  const Op op = op0.at(field, g.n);
  if (all)
    stencil_body<Op, true>(op, g, chunk, g.counters + field);
  else
    stencil_body<Op, false>(op, g, chunk, g.counters + field);
}

__device__ __forceinline__ bool def2(float a, float b, float undef)
{
  return is_def(a, undef) & is_def(b, undef);
}
__device__ __forceinline__ bool def4(float a, float b, float c, float d, float undef)
{
  return is_def(a, undef) & is_def(b, undef) & is_def(c, undef) & is_def(d, undef); // (bitwise: no short-circuit branches)
}

// centred difference times map ratio, evaluated like `0.5 * mapr[i] * (f[i+d] - f[i-d])`:
// float difference, double products
__device__ __forceinline__ double half_map_diff(float mapr, float hi, float lo)
{
  return 0.5 * (double)mapr * (double)(hi - lo);
}

// (float)half_map_diff(...), i.e. a derivative that the reference rounds to float on its own, without double
// arithmetic: 0.5*mapr and the product of two floats are exact in double, so the reference's value is the
// exact product rounded ONCE to float -- which is what a float multiplication returns (overflow, underflow
// to subnormals, infinities and NaN included), provided 0.5f*mapr is exact in float, i.e. mapr is not
// tiny (|mapr| >= 2^-100 or 0; tile::map_is_regular).  The float<->double conversions run on the quarter-rate XU pipe (16/clk/SM on B200, measured):
// with 4..9 of them per point these operators were XU-bound well below the HBM roofline.
template <bool FAST>
__device__ __forceinline__ float half_map_diff_f(float mapr, float hi, float lo)
{
  if (FAST) // the tile kernel checked that none of the CTA's map ratios is tiny, subnormal or NaN
    return (0.5f * mapr) * (hi - lo);
  return (float)(0.5 * (double)mapr * (double)(hi - lo));
}

// relvort / absvort / divergence (FC.cc:1843-1940).  MODE: 0 relvort, 1 absvort, 2 divergence
template <int MODE>
struct VortDivOp
{
  static constexpr bool ASM_STORE = true; // (stencil_tile.cuh, tile_compute)
  static constexpr int NOUT = 1;
  static constexpr bool TESTS_WHEN_ALL = false;
  const float *u, *v, *xm, *ym, *fc;
  float* o;
  // relvort/absvort: a = v[i-1], b = v[i+1], c = u[i-nx], d = u[i+nx] are both the tested and the used
  // operands.  divergence TESTS the same four (FC.cc:1927) but COMPUTES from u[i+-1], v[i+-nx]
  // (FC.cc:1928): the masked path loads both sets, the all-defined path only the second.
  template <bool ALL>
  struct In
  {
    float a, b, c, d, xm, ym, fc;
    float ta, tb, tc, td;
  };
  __device__ __forceinline__ bool all_defined(int, bool in_all) const { return in_all; }
  __device__ __forceinline__ VortDivOp at(int field, int n) const
  {
    VortDivOp r = *this;
    const long long off = (long long)field * n;
    r.u += off;
    r.v += off;
    r.o += off;
    return r;
  }
  __host__ __device__ __forceinline__ float* out(int) const { return o; }
  template <bool ALL>
  __device__ __forceinline__ In<ALL> load(int i, int nx) const
  {
    In<ALL> r;
    if (MODE == 2) {
      r.a = u[i - 1];
      r.b = u[i + 1];
      r.c = v[i - nx];
      r.d = v[i + nx];
      if (!ALL) {
        r.ta = v[i - 1];
        r.tb = v[i + 1];
        r.tc = u[i - nx];
        r.td = u[i + nx];
      }
    } else {
      r.a = v[i - 1];
      r.b = v[i + 1];
      r.c = u[i - nx];
      r.d = u[i + nx];
    }
    r.xm = xm[i];
    r.ym = ym[i];
    if (MODE == 1)
      r.fc = fc[i];
    return r;
  }
  // tile engine: array 0 is read at x +- 1, array 1 at y +- 1 (divergence's masked path reads both at both)
  static constexpr int HX = 1, EXTRA_FLOATS = 0;
  static constexpr bool CUSTOM_TILE = false;
  template <class V, class M>
  __device__ unsigned tile_custom(const V&, const M&, float*, bool, bool, int, int, int, int, int, int, int, float) const { return 0; }
  static constexpr int TY = 8, NARR = 2, NMAPS = (MODE == 1) ? 3 : 2;
  __host__ __device__ __forceinline__ const float* arr(int k) const { return (MODE == 2) ? (k == 0 ? u : v) : (k == 0 ? v : u); }
  __host__ __device__ static constexpr int halo(int k) { return (MODE == 2) ? 1 : k; }
  __host__ __device__ __forceinline__ const float* map(int k) const { return k == 0 ? xm : k == 1 ? ym : fc; }
  template <bool ALL, class View>
  __device__ __forceinline__ In<ALL> fetch(const View& t, int r, const float* m) const
  {
    In<ALL> in;
    in.a = t.template at<0>(r, -1);
    in.b = t.template at<0>(r, 1);
    in.c = t.template at<1>(r - 1, 0);
    in.d = t.template at<1>(r + 1, 0);
    if (MODE == 2 && !ALL) {
      in.ta = t.template at<1>(r, -1);
      in.tb = t.template at<1>(r, 1);
      in.tc = t.template at<0>(r - 1, 0);
      in.td = t.template at<0>(r + 1, 0);
    }
    in.xm = m[0];
    in.ym = m[1];
    if (MODE == 1)
      in.fc = m[2];
    return in;
  }
  template <bool ALL, bool FAST = false>
  __device__ __forceinline__ bool eval(const In<ALL>& r, float undef, float* val) const
  {
    bool ok = true;
    if (!ALL)
      ok = (MODE == 2) ? def4(r.ta, r.tb, r.tc, r.td, undef) : def4(r.a, r.b, r.c, r.d, undef);
    // (no early return: a warp computes as soon as ONE of its points is defined, so the undefined ones cost nothing extra and
    // the branch around the arithmetic only adds instructions and divergence; the caller selects by `ok`)
    // 0.5*xm*dx -+ 0.5*ym*dy: the products of two floats are exact in double and so is the scaling by 0.5,
    // hence RN(0.5*px -+ 0.5*py) = 0.5 * RN(px -+ py) = 0.5 * fma(-+ym, dy, px): three double instructions
    const double px = (double)r.xm * (double)(r.b - r.a);
    const double ym = (MODE == 2) ? (double)r.ym : -(double)r.ym;
    const double h = 0.5 * fma(ym, (double)(r.d - r.c), px);
    val[0] = (MODE == 1) ? (float)(h + (double)r.fc) : (float)h;
    return ok;
  }
};

// advection (FC.cc:1942-1983)
struct AdvectionOp
{
  static constexpr int NOUT = 1;
  static constexpr bool TESTS_WHEN_ALL = false;
  const float *f, *u, *v, *xm, *ym;
  float scale;
  float* o;
  template <bool ALL>
  struct In
  {
    float ui, vi, fd, fl, fr, fu, xm, ym;
  };
  __device__ __forceinline__ bool all_defined(int, bool in_all) const { return in_all; }
  __device__ __forceinline__ AdvectionOp at(int field, int n) const
  {
    AdvectionOp r = *this;
    const long long off = (long long)field * n;
    r.f += off;
    r.u += off;
    r.v += off;
    r.o += off;
    return r;
  }
  __host__ __device__ __forceinline__ float* out(int) const { return o; }
  template <bool ALL>
  __device__ __forceinline__ In<ALL> load(int i, int nx) const
  {
    In<ALL> r;
    r.ui = u[i];
    r.vi = v[i];
    r.fd = f[i - nx];
    r.fl = f[i - 1];
    r.fr = f[i + 1];
    r.fu = f[i + nx];
    r.xm = xm[i];
    r.ym = ym[i];
    return r;
  }
  static constexpr int HX = 1, EXTRA_FLOATS = 0;
  static constexpr bool CUSTOM_TILE = false;
  template <class V, class M>
  __device__ unsigned tile_custom(const V&, const M&, float*, bool, bool, int, int, int, int, int, int, int, float) const { return 0; }
  static constexpr int TY = 8, NARR = 3, NMAPS = 2;
  __host__ __device__ __forceinline__ const float* arr(int k) const { return k == 0 ? f : k == 1 ? u : v; }
  __host__ __device__ static constexpr int halo(int k) { return k == 0 ? 1 : 0; }
  __host__ __device__ __forceinline__ const float* map(int k) const { return k == 0 ? xm : ym; }
  template <bool ALL, class View>
  __device__ __forceinline__ In<ALL> fetch(const View& t, int r, const float* m) const
  {
    In<ALL> in;
    in.ui = t.template at<1>(r, 0);
    in.vi = t.template at<2>(r, 0);
    in.fd = t.template at<0>(r - 1, 0);
    in.fl = t.template at<0>(r, -1);
    in.fr = t.template at<0>(r, 1);
    in.fu = t.template at<0>(r + 1, 0);
    in.xm = m[0];
    in.ym = m[1];
    return in;
  }
  template <bool ALL, bool FAST = false>
  __device__ __forceinline__ bool eval(const In<ALL>& r, float undef, float* val) const
  {
    bool ok = true;
    if (!ALL)
      ok = def2(r.ui, r.vi, undef) & def4(r.fd, r.fl, r.fr, r.fu, undef);
    const double ax = (double)r.ui * 0.5 * (double)r.xm * (double)(r.fr - r.fl);
    const double ay = (double)r.vi * 0.5 * (double)r.ym * (double)(r.fu - r.fd);
    val[0] = (float)((ax + ay) * (double)scale);
    return ok;
  }
};

// gradient (FC.cc:1985-2074), COMPUTE = 1..4
template <int COMPUTE>
struct GradientOp
{
  static constexpr int NOUT = 1;
  static constexpr bool TESTS_WHEN_ALL = false;
  static constexpr int TILE_CTAS = 4; // one staged array: small stages, 56 registers are enough (0.66 / 0.63 / 0.68 with 2 / 3 / 4 CTAs per SM)
  const float *f, *xm, *ym;
  float* o;
  template <bool ALL>
  struct In
  {
    float fd, fl, fc, fr, fu, xm, ym;
  };
  __device__ __forceinline__ bool all_defined(int, bool in_all) const { return in_all; }
  __device__ __forceinline__ GradientOp at(int field, int n) const
  {
    GradientOp r = *this;
    const long long off = (long long)field * n;
    r.f += off;
    r.o += off;
    return r;
  }
  __host__ __device__ __forceinline__ float* out(int) const { return o; }
  template <bool ALL>
  __device__ __forceinline__ In<ALL> load(int i, int nx) const
  {
    In<ALL> r;
    if (COMPUTE != 2) {
      r.fl = f[i - 1];
      r.fr = f[i + 1];
      r.xm = xm[i];
    }
    if (COMPUTE != 1) {
      r.fd = f[i - nx];
      r.fu = f[i + nx];
      r.ym = ym[i];
    }
    if (COMPUTE == 4)
      r.fc = f[i];
    return r;
  }
  static constexpr int HX = 1, EXTRA_FLOATS = 0;
  static constexpr bool CUSTOM_TILE = false;
  template <class V, class M>
  __device__ unsigned tile_custom(const V&, const M&, float*, bool, bool, int, int, int, int, int, int, int, float) const { return 0; }
  static constexpr int TY = 8, NARR = 1, NMAPS = (COMPUTE <= 2) ? 1 : 2;
  __host__ __device__ __forceinline__ const float* arr(int) const { return f; }
  __host__ __device__ static constexpr int halo(int) { return (COMPUTE == 1) ? 0 : 1; }
  __host__ __device__ __forceinline__ const float* map(int k) const { return (COMPUTE == 2 || k == 1) ? ym : xm; }
  template <bool ALL, class View>
  __device__ __forceinline__ In<ALL> fetch(const View& t, int r, const float* m) const
  {
    In<ALL> in;
    if (COMPUTE != 2) {
      in.fl = t.template at<0>(r, -1);
      in.fr = t.template at<0>(r, 1);
      in.xm = m[0];
    }
    if (COMPUTE != 1) {
      in.fd = t.template at<0>(r - 1, 0);
      in.fu = t.template at<0>(r + 1, 0);
      in.ym = (COMPUTE == 2) ? m[0] : m[1];
    }
    if (COMPUTE == 4)
      in.fc = t.template at<0>(r, 0);
    return in;
  }
  template <bool ALL, bool FAST = false>
  __device__ __forceinline__ bool eval(const In<ALL>& r, float undef, float* val) const
  {
    bool ok = true;
    if (!ALL) {
      if (COMPUTE == 1)
        ok = def2(r.fl, r.fr, undef);
      else if (COMPUTE == 2)
        ok = def2(r.fd, r.fu, undef);
      else if (COMPUTE == 3)
        ok = def4(r.fd, r.fl, r.fr, r.fu, undef);
      else
        ok = def4(r.fd, r.fl, r.fr, r.fu, undef) & is_def(r.fc, undef);
    }
    if (COMPUTE == 1) {
      val[0] = half_map_diff_f<FAST>(r.xm, r.fr, r.fl);
    } else if (COMPUTE == 2) {
      val[0] = half_map_diff_f<FAST>(r.ym, r.fu, r.fd);
    } else if (COMPUTE == 3) {
      const float dfdx = half_map_diff_f<FAST>(r.xm, r.fr, r.fl);
      const float dfdy = half_map_diff_f<FAST>(r.ym, r.fu, r.fd);
      // dev::absval; an undefined point (its value is dropped by the caller) takes sqrtf(1): the huge sums of squares of `undef`
      // operands would send its lane, and with it the warp, through sqrtf's out-of-range path
      const float s2 = dfdx * dfdx + dfdy * dfdy;
      val[0] = sqrtf(ALL ? s2 : (ok ? s2 : 1.f));
    } else {
      const float d2fdx = (float)((double)r.fl - 2.0 * (double)r.fc + (double)r.fr);
      const float d2fdy = (float)((double)r.fd - 2.0 * (double)r.fc + (double)r.fu);
      const double mx = (double)r.xm, my = (double)r.ym;
      val[0] = (float)(4.0 * (0.25 * mx * mx * (double)d2fdx + 0.25 * my * my * (double)d2fdy));
    }
    return ok;
  }
};

// jacobian (FC.cc:2424-2460): four derivatives rounded to float, a*b - c*d in float without FMA
struct JacobianOp
{
  static constexpr bool ASM_STORE = true; // (stencil_tile.cuh, tile_compute)
  static constexpr int NOUT = 1;
  static constexpr bool TESTS_WHEN_ALL = false;
  const float *f1, *f2, *xm, *ym;
  float* o;
  template <bool ALL>
  struct In
  {
    float ad, al, ar, au, bd, bl, br, bu, xm, ym;
  };
  __device__ __forceinline__ bool all_defined(int, bool in_all) const { return in_all; }
  __device__ __forceinline__ JacobianOp at(int field, int n) const
  {
    JacobianOp r = *this;
    const long long off = (long long)field * n;
    r.f1 += off;
    r.f2 += off;
    r.o += off;
    return r;
  }
  __host__ __device__ __forceinline__ float* out(int) const { return o; }
  template <bool ALL>
  __device__ __forceinline__ In<ALL> load(int i, int nx) const
  {
    In<ALL> r;
    r.ad = f1[i - nx];
    r.al = f1[i - 1];
    r.ar = f1[i + 1];
    r.au = f1[i + nx];
    r.bd = f2[i - nx];
    r.bl = f2[i - 1];
    r.br = f2[i + 1];
    r.bu = f2[i + nx];
    r.xm = xm[i];
    r.ym = ym[i];
    return r;
  }
  static constexpr int HX = 1, EXTRA_FLOATS = 0;
  static constexpr bool CUSTOM_TILE = false;
  template <class V, class M>
  __device__ unsigned tile_custom(const V&, const M&, float*, bool, bool, int, int, int, int, int, int, int, float) const { return 0; }
  static constexpr int TY = 8, NARR = 2, NMAPS = 2;
  __host__ __device__ __forceinline__ const float* arr(int k) const { return k == 0 ? f1 : f2; }
  __host__ __device__ static constexpr int halo(int) { return 1; }
  __host__ __device__ __forceinline__ const float* map(int k) const { return k == 0 ? xm : ym; }
  template <bool ALL, class View>
  __device__ __forceinline__ In<ALL> fetch(const View& t, int r, const float* m) const
  {
    In<ALL> in;
    in.ad = t.template at<0>(r - 1, 0);
    in.al = t.template at<0>(r, -1);
    in.ar = t.template at<0>(r, 1);
    in.au = t.template at<0>(r + 1, 0);
    in.bd = t.template at<1>(r - 1, 0);
    in.bl = t.template at<1>(r, -1);
    in.br = t.template at<1>(r, 1);
    in.bu = t.template at<1>(r + 1, 0);
    in.xm = m[0];
    in.ym = m[1];
    return in;
  }
  template <bool ALL, bool FAST = false>
  __device__ __forceinline__ bool eval(const In<ALL>& r, float undef, float* val) const
  {
    bool ok = true;
    if (!ALL)
      ok = def4(r.ad, r.al, r.ar, r.au, undef) & def4(r.bd, r.bl, r.br, r.bu, undef);
    const float df1dx = half_map_diff_f<FAST>(r.xm, r.ar, r.al);
    const float df1dy = half_map_diff_f<FAST>(r.ym, r.au, r.ad);
    const float df2dx = half_map_diff_f<FAST>(r.xm, r.br, r.bl);
    const float df2dy = half_map_diff_f<FAST>(r.ym, r.bu, r.bd);
    val[0] = df1dx * df2dy - df1dy * df2dx;
    return ok;
  }
};

// ilevelgwind (FC.cc:1511-1549): two outputs
struct GwindOp
{
  static constexpr int NOUT = 2;
  static constexpr bool TESTS_WHEN_ALL = false;
  const float *m, *xm, *ym, *fc;
  float *ug, *vg;
  template <bool ALL>
  struct In
  {
    float md, ml, mr, mu, xm, ym, fc;
  };
  __device__ __forceinline__ bool all_defined(int, bool in_all) const { return in_all; }
  __device__ __forceinline__ GwindOp at(int field, int n) const
  {
    GwindOp r = *this;
    const long long off = (long long)field * n;
    r.m += off;
    r.ug += off;
    r.vg += off;
    return r;
  }
  __host__ __device__ __forceinline__ float* out(int k) const { return k == 0 ? ug : vg; }
  template <bool ALL>
  __device__ __forceinline__ In<ALL> load(int i, int nx) const
  {
    In<ALL> r;
    r.md = m[i - nx];
    r.ml = m[i - 1];
    r.mr = m[i + 1];
    r.mu = m[i + nx];
    r.xm = xm[i];
    r.ym = ym[i];
    r.fc = fc[i];
    return r;
  }
  static constexpr int HX = 1, EXTRA_FLOATS = 0;
  static constexpr bool CUSTOM_TILE = false;
  template <class V, class M>
  __device__ unsigned tile_custom(const V&, const M&, float*, bool, bool, int, int, int, int, int, int, int, float) const { return 0; }
  static constexpr int TY = 8, NARR = 1, NMAPS = 3;
  __host__ __device__ __forceinline__ const float* arr(int) const { return m; }
  __host__ __device__ static constexpr int halo(int) { return 1; }
  __host__ __device__ __forceinline__ const float* map(int k) const { return k == 0 ? xm : k == 1 ? ym : fc; }
  template <bool ALL, class View>
  __device__ __forceinline__ In<ALL> fetch(const View& t, int r, const float* mp) const
  {
    In<ALL> in;
    in.md = t.template at<0>(r - 1, 0);
    in.ml = t.template at<0>(r, -1);
    in.mr = t.template at<0>(r, 1);
    in.mu = t.template at<0>(r + 1, 0);
    in.xm = mp[0];
    in.ym = mp[1];
    in.fc = mp[2];
    return in;
  }
  template <bool ALL, bool FAST = false>
  __device__ __forceinline__ bool eval(const In<ALL>& r, float undef, float* val) const
  {
    bool ok = true;
    if (!ALL)
      ok = def4(r.md, r.ml, r.mr, r.mu, undef);
    if (!ok)
      return false; // also keeps undefined operands out of the double division
    const double f = (double)r.fc;
    val[0] = (float)(-0.5 * (double)r.ym * (double)(r.mu - r.md) / f);
    val[1] = (float)(0.5 * (double)r.xm * (double)(r.mr - r.ml) / f);
    return ok;
  }
};

// thermalFrontParameter, second pass of the UNFUSED form (FC.cc:2286-2303): |grad T| comes from a scratch field written by
// GradientOp<3> (with its fillEdges).  all2 -- this pass's allDefined -- is pass 1's OUTPUT flag, i.e. "pass 1 counted
// nothing": read from pass 1's device counter, no host round trip.
struct TfpPass2Op
{
  static constexpr int NOUT = 1;
  static constexpr bool TESTS_WHEN_ALL = true;
  const float *tx, *ad, *xm, *ym;
  const unsigned long long* pass1_undef; // per field
  float* o;
  template <bool ALL>
  struct In
  {
    float td, tl, tr, tu, adn, al, ac, ar, au, xm, ym;
  };
  __device__ __forceinline__ bool all_defined(int field, bool) const { return pass1_undef[field] == 0; }
  __device__ __forceinline__ TfpPass2Op at(int field, int n) const
  {
    TfpPass2Op r = *this;
    const long long off = (long long)field * n;
    r.tx += off;
    r.ad += off;
    r.o += off;
    return r;
  }
  __host__ __device__ __forceinline__ float* out(int) const { return o; }
  template <bool ALL>
  __device__ __forceinline__ In<ALL> load(int i, int nx) const
  {
    In<ALL> r;
    r.td = tx[i - nx], r.tl = tx[i - 1], r.tr = tx[i + 1], r.tu = tx[i + nx];
    r.adn = ad[i - nx], r.al = ad[i - 1], r.ac = ad[i], r.ar = ad[i + 1], r.au = ad[i + nx];
    r.xm = xm[i];
    r.ym = ym[i];
    return r;
  }
  static constexpr int HX = 1, EXTRA_FLOATS = 0;
  static constexpr bool CUSTOM_TILE = false;
  template <class V, class M>
  __device__ unsigned tile_custom(const V&, const M&, float*, bool, bool, int, int, int, int, int, int, int, float) const { return 0; }
  static constexpr int TY = 8, NARR = 2, NMAPS = 2;
  __host__ __device__ __forceinline__ const float* arr(int k) const { return k == 0 ? tx : ad; }
  __host__ __device__ static constexpr int halo(int) { return 1; }
  __host__ __device__ __forceinline__ const float* map(int k) const { return k == 0 ? xm : ym; }
  template <bool ALL, class View>
  __device__ __forceinline__ In<ALL> fetch(const View& v, int r, const float* m) const
  {
    In<ALL> in;
    in.td = v.template at<0>(r - 1, 0), in.tl = v.template at<0>(r, -1), in.tr = v.template at<0>(r, 1), in.tu = v.template at<0>(r + 1, 0);
    in.adn = v.template at<1>(r - 1, 0), in.al = v.template at<1>(r, -1), in.ac = v.template at<1>(r, 0), in.ar = v.template at<1>(r, 1),
    in.au = v.template at<1>(r + 1, 0);
    in.xm = m[0];
    in.ym = m[1];
    return in;
  }
  template <bool ALL, bool FAST = false>
  __device__ __forceinline__ bool eval(const In<ALL>& r, float undef, float* val) const
  {
    bool ok = true;
    if (!ALL)
      ok = def4(r.td, r.tl, r.tr, r.tu, undef) && def4(r.adn, r.al, r.ar, r.au, undef) && is_def(r.ac, undef);
    ok = ok && (r.ac != 0); // tested even when allDefined (FC.cc:2292)
    if (!ok)
      return false;
    const float dadx = half_map_diff_f<FAST>(r.xm, r.ar, r.al);
    const float dady = half_map_diff_f<FAST>(r.ym, r.au, r.adn);
    const float dtdxa = (float)(half_map_diff(r.xm, r.tr, r.tl) / (double)r.ac);
    const float dtdya = (float)(half_map_diff(r.ym, r.tu, r.td) / (double)r.ac);
    val[0] = -(dadx * dtdxa + dady * dtdya);
    return true;
  }
};

// plevelgwind_xcomp / plevelgwind_ycomp / plevelgvort (FC.cc:638-743): geostrophic wind components
// and geostrophic vorticity from the height of a pressure surface.  Same tile shape as GwindOp (one staged array, three
// grid-constant maps); MODE 0 = xcomp, 1 = ycomp, 2 = gvort (which also reads the centre point).
template <int MODE>
struct GeoOp
{
  static constexpr int NOUT = 1;
  static constexpr bool TESTS_WHEN_ALL = false;
  const float *z, *xm, *ym, *fc;
  float* o;
  template <bool ALL>
  struct In
  {
    float zd, zl, zc, zr, zu, xm, ym, fc;
  };
  __device__ __forceinline__ bool all_defined(int, bool in_all) const { return in_all; }
  __device__ __forceinline__ GeoOp at(int field, int n) const
  {
    GeoOp r = *this;
    const long long off = (long long)field * n;
    r.z += off;
    r.o += off;
    return r;
  }
  __host__ __device__ __forceinline__ float* out(int) const { return o; }
  template <bool ALL>
  __device__ __forceinline__ In<ALL> load(int i, int nx) const
  {
    In<ALL> r;
    r.zd = z[i - nx];
    r.zl = z[i - 1];
    r.zc = z[i];
    r.zr = z[i + 1];
    r.zu = z[i + nx];
    r.xm = xm[i];
    r.ym = ym[i];
    r.fc = fc[i];
    return r;
  }
  static constexpr int HX = 1, EXTRA_FLOATS = 0;
  static constexpr bool CUSTOM_TILE = false;
  template <class V, class M>
  __device__ unsigned tile_custom(const V&, const M&, float*, bool, bool, int, int, int, int, int, int, int, float) const { return 0; }
  static constexpr int TY = 8, NARR = 1, NMAPS = 3;
  __host__ __device__ __forceinline__ const float* arr(int) const { return z; }
  __host__ __device__ static constexpr int halo(int) { return 1; }
  __host__ __device__ __forceinline__ const float* map(int k) const { return k == 0 ? xm : k == 1 ? ym : fc; }
  template <bool ALL, class View>
  __device__ __forceinline__ In<ALL> fetch(const View& t, int r, const float* mp) const
  {
    In<ALL> in;
    in.zd = t.template at<0>(r - 1, 0);
    in.zl = t.template at<0>(r, -1);
    in.zc = t.template at<0>(r, 0);
    in.zr = t.template at<0>(r, 1);
    in.zu = t.template at<0>(r + 1, 0);
    in.xm = mp[0];
    in.ym = mp[1];
    in.fc = mp[2];
    return in;
  }
  template <bool ALL, bool FAST = false>
  __device__ __forceinline__ bool eval(const In<ALL>& r, float undef, float* val) const
  {
    bool ok = true;
    if (!ALL)
      ok = def4(r.zd, r.zl, r.zr, r.zu, undef) && (MODE != 2 || is_def(r.zc, undef));
    if (!ok)
      return false;
    constexpr float g = (float)9.8; // MC.h:46
    const double f = (double)r.fc;
    if (MODE == 0)
      val[0] = (float)(-0.5 * (double)r.ym * (double)(r.zu - r.zd) * (double)g / f); // :662
    else if (MODE == 1)
      val[0] = (float)(0.5 * (double)r.xm * (double)(r.zr - r.zl) * (double)g / f); // :694
    else {
      constexpr float g4 = (float)((double)g * 4.); // :717
      const double zc2 = 2. * (double)r.zc;
      const double d2x = (double)r.zl - zc2 + (double)r.zr, d2y = (double)r.zd - zc2 + (double)r.zu;
      val[0] = (float)((0.25 * (double)r.xm * (double)r.xm * d2x + 0.25 * (double)r.ym * (double)r.ym * d2y) * (double)g4 / f); // :728-729
    }
    return true;
  }
};

// plevelqvector, second pass (FC.cc:566-589): five-point differences of the geostrophic wind components (two scratch
// fields produced by GeoOp<0> / GeoOp<1>, including their fillEdges) and of the temperature.  The tests are plain
// `!= undef` comparisons (a NaN passes) and apply whatever the flag says.  COMP 0 = x component, 1 = y component.
template <int COMP>
struct QvecOp
{
  static constexpr int NOUT = 1;
  static constexpr bool TESTS_WHEN_ALL = true;
  const float *ug, *vg, *t, *xm, *ym;
  float tscale, c;
  float* o;
  template <bool ALL>
  struct In
  {
    float ud, ul, ur, uu, vd, vl, vr, vu, td, tl, tr, tu, xm, ym;
  };
  __device__ __forceinline__ bool all_defined(int, bool in_all) const { return in_all; }
  __device__ __forceinline__ QvecOp at(int field, int n) const
  {
    QvecOp r = *this;
    const long long off = (long long)field * n;
    r.ug += off;
    r.vg += off;
    r.t += off;
    r.o += off;
    return r;
  }
  __host__ __device__ __forceinline__ float* out(int) const { return o; }
  template <bool ALL>
  __device__ __forceinline__ In<ALL> load(int i, int nx) const
  {
    In<ALL> r;
    r.ud = ug[i - nx], r.ul = ug[i - 1], r.ur = ug[i + 1], r.uu = ug[i + nx];
    r.vd = vg[i - nx], r.vl = vg[i - 1], r.vr = vg[i + 1], r.vu = vg[i + nx];
    r.td = t[i - nx], r.tl = t[i - 1], r.tr = t[i + 1], r.tu = t[i + nx];
    r.xm = xm[i];
    r.ym = ym[i];
    return r;
  }
  static constexpr int HX = 1, EXTRA_FLOATS = 0;
  static constexpr bool CUSTOM_TILE = false;
  template <class V, class M>
  __device__ unsigned tile_custom(const V&, const M&, float*, bool, bool, int, int, int, int, int, int, int, float) const { return 0; }
  static constexpr int TY = 8, NARR = 3, NMAPS = 2;
  __host__ __device__ __forceinline__ const float* arr(int k) const { return k == 0 ? ug : k == 1 ? vg : t; }
  __host__ __device__ static constexpr int halo(int) { return 1; }
  __host__ __device__ __forceinline__ const float* map(int k) const { return k == 0 ? xm : ym; }
  template <bool ALL, class View>
  __device__ __forceinline__ In<ALL> fetch(const View& v, int r, const float* m) const
  {
    In<ALL> in;
    in.ud = v.template at<0>(r - 1, 0), in.ul = v.template at<0>(r, -1), in.ur = v.template at<0>(r, 1), in.uu = v.template at<0>(r + 1, 0);
    in.vd = v.template at<1>(r - 1, 0), in.vl = v.template at<1>(r, -1), in.vr = v.template at<1>(r, 1), in.vu = v.template at<1>(r + 1, 0);
    in.td = v.template at<2>(r - 1, 0), in.tl = v.template at<2>(r, -1), in.tr = v.template at<2>(r, 1), in.tu = v.template at<2>(r + 1, 0);
    in.xm = m[0];
    in.ym = m[1];
    return in;
  }
  template <bool ALL, bool FAST = false>
  __device__ __forceinline__ bool eval(const In<ALL>& r, float undef, float* val) const
  {
    const bool ok = r.ud != undef && r.ul != undef && r.ur != undef && r.uu != undef && r.vd != undef && r.vl != undef && r.vr != undef &&
                    r.vu != undef && r.td != undef && r.tl != undef && r.tr != undef && r.tu != undef;
    if (!ok)
      return false;
    const float dtdx = (float)(0.5 * (double)r.xm * (double)tscale * (double)(r.tr - r.tl));
    const float dtdy = (float)(0.5 * (double)r.ym * (double)tscale * (double)(r.tu - r.td));
    float a, b;
    if (COMP == 0) {
      a = (float)(0.5 * (double)r.xm * (double)(r.ur - r.ul)); // dug/dx
      b = (float)(0.5 * (double)r.xm * (double)(r.vr - r.vl)); // dvg/dx
    } else {
      a = (float)(0.5 * (double)r.ym * (double)(r.uu - r.ud)); // dug/dy
      b = (float)(0.5 * (double)r.ym * (double)(r.vu - r.vd)); // dvg/dy
    }
    val[0] = c * (a * dtdx + b * dtdy);
    return true;
  }
};

// thermalFrontParameter, FUSED (FC.cc:2266-2309).  The reference runs gradient(c=3) into a scratch field
// `absdelt` (with its fillEdges, so absdelt(x, y) = G(clamp(x, 1, nx-2), clamp(y, 1, ny-2)), G = |grad T| at an
// interior point) and then a second five-point pass over T and absdelt.  Here a tile of T with a halo of
// two is staged once, |grad T| is evaluated for the tile plus a halo of one into shared memory and the
// second pass reads it from there: 8 B/point of HBM traffic instead of 20, no scratch field.
//   all1 (pass 1) = the input flag; all2 (pass 2) = pass 1's output flag = "pass 1 counted nothing" =
//   "no element of T is undefined" (the flat loop [nx, N-nx) tests every element of the field): for fields
//   whose input flag is not ALL_DEFINED a pre-pass counts the undefined elements of T (t_undef).
// The scalar load() evaluates G on the fly for the five absdelt operands: it serves the border ring
// (stencil_edge_kernel) and the small grids of the flat kernel.
constexpr int TFP_APITCH = tile::TX + 4; // |grad T| tile: (TY + 2) rows of TX + 2 columns

struct TfpFusedOp
{
  static constexpr int NOUT = 1;
  static constexpr bool TESTS_WHEN_ALL = true;
  const float *tx, *xm, *ym;
  const unsigned long long* t_undef; // per field: undefined elements of T (only counted when the flag is not ALL)
  const FieldMeta* meta;
  int gnx, gny;
  float undef_;
  bool all1;
  float* o;
  template <bool ALL>
  struct In
  {
    float td, tl, tr, tu, adn, al, ac, ar, au, xm, ym;
  };
  __device__ __forceinline__ bool all_defined(int field, bool in_all) const { return in_all || t_undef[field] == 0; }
  __device__ __forceinline__ TfpFusedOp at(int field, int n) const
  {
    TfpFusedOp r = *this;
    const long long off = (long long)field * n;
    r.tx += off;
    r.o += off;
    r.all1 = meta[field].all != 0;
    return r;
  }
  __host__ __device__ __forceinline__ float* out(int) const { return o; }
  // pass 1 at an interior point (FC.cc:2037-2042)
  template <bool FAST = false>
  __device__ __forceinline__ float grad(float fd, float fl, float fr, float fu, float xmi, float ymi) const
  {
    if (!all1 && !def4(fd, fl, fr, fu, undef_))
      return undef_;
    const float dfdx = half_map_diff_f<FAST>(xmi, fr, fl);
    const float dfdy = half_map_diff_f<FAST>(ymi, fu, fd);
    return dev::absval(dfdx, dfdy);
  }
  // absdelt at flat index i: pass 1's value at the clamped interior point
  __device__ float absdelt(int i) const
  {
    const int y = i / gnx, x = i - y * gnx;
    const int cx = min(max(x, 1), gnx - 2), cy = min(max(y, 1), gny - 2);
    const int j = cy * gnx + cx;
    return grad(tx[j - gnx], tx[j - 1], tx[j + 1], tx[j + gnx], xm[j], ym[j]);
  }
  template <bool ALL>
  __device__ In<ALL> load(int i, int nx) const
  {
    In<ALL> r;
    r.td = tx[i - nx];
    r.tl = tx[i - 1];
    r.tr = tx[i + 1];
    r.tu = tx[i + nx];
    r.adn = absdelt(i - nx);
    r.al = absdelt(i - 1);
    r.ac = absdelt(i);
    r.ar = absdelt(i + 1);
    r.au = absdelt(i + nx);
    r.xm = xm[i];
    r.ym = ym[i];
    return r;
  }
  // the second pass, FC.cc:2286-2303
  template <bool ALL, bool FAST = false>
  __device__ __forceinline__ bool eval(const In<ALL>& r, float undef, float* val) const
  {
    bool ok = true;
    if (!ALL)
      ok = def4(r.td, r.tl, r.tr, r.tu, undef) && def4(r.adn, r.al, r.ar, r.au, undef) && is_def(r.ac, undef);
    ok = ok && (r.ac != 0); // tested even when allDefined (FC.cc:2292)
    if (!ok)
      return false;
    const float dadx = half_map_diff_f<FAST>(r.xm, r.ar, r.al);
    const float dady = half_map_diff_f<FAST>(r.ym, r.au, r.adn);
    // (a shared-reciprocal division with a Markstein correction was measured SLOWER than the two IEEE divisions)
    const float dtdxa = (float)(half_map_diff(r.xm, r.tr, r.tl) / (double)r.ac);
    const float dtdya = (float)(half_map_diff(r.ym, r.tu, r.td) / (double)r.ac);
    val[0] = -(dadx * dtdxa + dady * dtdya);
    return true;
  }

  // ---- tile engine: T staged with a halo of two rows and two columns
  static constexpr int HX = 2, TY = 8, NARR = 1, NMAPS = 2;
  static constexpr int EXTRA_FLOATS = (TY + 2) * TFP_APITCH;
  static constexpr bool CUSTOM_TILE = true;
  __host__ __device__ __forceinline__ const float* arr(int) const { return tx; }
  __host__ __device__ static constexpr int halo(int) { return 2; }
  __host__ __device__ __forceinline__ const float* map(int k) const { return k == 0 ? xm : ym; }
  template <bool ALL, class View>
  __device__ __forceinline__ In<ALL> fetch(const View&, int, const float*) const
  {
    return In<ALL>();
  }

  template <bool ALL2, bool FAST, class View>
  __device__ __forceinline__ unsigned pass2(const View& tv, const tile::MapSlots& maps, const float* ad, int nrows, bool col_ok, int i0, int nx, float undef) const
  {
    unsigned nundef = 0;
#pragma unroll
    for (int r = 0; r < TY; ++r) {
      if (r < nrows) { // warp-uniform
        In<ALL2> in;
        in.td = tv.template at<0>(r - 1, 0);
        in.tl = tv.template at<0>(r, -1);
        in.tr = tv.template at<0>(r, 1);
        in.tu = tv.template at<0>(r + 1, 0);
        const float* a = ad + (r + 1) * TFP_APITCH; // row r of the |grad T| tile, this thread's column
        in.adn = a[-TFP_APITCH];
        in.al = a[-1];
        in.ac = a[0];
        in.ar = a[1];
        in.au = a[TFP_APITCH];
        in.xm = maps.get(0, r, TY);
        in.ym = maps.get(1, r, TY);
        float val[1];
        const bool ok = eval<ALL2, FAST>(in, undef, val);
        if (col_ok) {
          if (!ok)
            nundef += 1;
          o[i0 + r * nx] = ok ? val[0] : undef;
        }
      }
    }
    return nundef;
  }

  // consumers only (TX threads); `scratch` = EXTRA_FLOATS floats of shared memory
  template <class View>
  __device__ unsigned tile_custom(const View& tv, const tile::MapSlots& maps, float* scratch, bool all2, bool fast, int x0, int y0, int xlast, int ylast, int c, int nx,
                                  int ny, float undef) const
  {
    if (fast)
      return tile_two_phase<true>(tv, maps, scratch, all2, x0, y0, xlast, ylast, c, nx, ny, undef);
    return tile_two_phase<false>(tv, maps, scratch, all2, x0, y0, xlast, ylast, c, nx, ny, undef);
  }
  template <bool FAST, class View>
  __device__ __forceinline__ unsigned tile_two_phase(const View& tv, const tile::MapSlots& maps, float* scratch, bool all2, int x0, int y0, int xlast, int ylast, int c,
                                                     int nx, int ny, float undef) const
  {
    const int nrows = ylast - y0 + 1, ncols = xlast - x0 + 1;
    const bool col_ok = c < ncols;
    float* ad = scratch + c + 1; // column c of the |grad T| tile; tile row rr lives at (rr + 1) * TFP_APITCH
    // phase 1a: G at the interior points of tile rows -1 .. TY, this thread's column.  The map ratios of rows
    // -1 and TY are not in registers: two extra (L2-resident) loads per thread.
    asm volatile("bar.sync 1, %0;" ::"n"(tile::TX)); // the previous field's pass 2 no longer reads the scratch tile
    const int x = x0 + c;
#pragma unroll
    for (int rr = -1; rr <= TY; ++rr) {
      const int y = y0 + rr;
      if (rr <= nrows && y >= 1 && y <= ny - 2 && x <= nx - 2) {
        float xmi, ymi;
        if (rr >= 0 && rr < TY && rr < nrows && col_ok) {
          xmi = maps.get(0, rr < 0 ? 0 : (rr >= TY ? TY - 1 : rr), TY);
          ymi = maps.get(1, rr < 0 ? 0 : (rr >= TY ? TY - 1 : rr), TY);
        } else {
          xmi = xm[y * nx + x];
          ymi = ym[y * nx + x];
        }
        // (the halo rows' map ratios are not covered by the CTA's regularity check: reference arithmetic there)
        ad[(rr + 1) * TFP_APITCH] =
            (rr >= 0 && rr < TY) ? grad<FAST>(tv.template at<0>(rr - 1, 0), tv.template at<0>(rr, -1), tv.template at<0>(rr, 1), tv.template at<0>(rr + 1, 0), xmi, ymi)
                                 : grad<false>(tv.template at<0>(rr - 1, 0), tv.template at<0>(rr, -1), tv.template at<0>(rr, 1), tv.template at<0>(rr + 1, 0), xmi, ymi);
      }
    }
    // the two halo columns -1 and ncols (when they are interior columns of the grid): threads 0 .. TY+1 and 32 .. 32+TY+1
    {
      const int which = c >> 5, rr = (c & 31) - 1; // which: 0 = column -1, 1 = column ncols
      if (which < 2 && rr <= TY && rr <= nrows) {
        const int cc = which == 0 ? -1 : ncols;
        const int xx = x0 + cc, y = y0 + rr;
        if (xx >= 1 && xx <= nx - 2 && y >= 1 && y <= ny - 2) {
          const float* t = tx + (long long)y * nx + xx; // 10 + 10 points per tile: straight from global memory
          scratch[(rr + 1) * TFP_APITCH + cc + 1] = grad(t[-nx], t[-1], t[1], t[nx], xm[y * nx + xx], ym[y * nx + xx]);
        }
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(tile::TX));
    // phase 1b: fillEdges of pass 1 inside the tile's halo -- border columns first, then border rows (FC.cc:59-74)
    if (x0 == 1 || xlast == nx - 2) {
      const int which = c >> 5, rr = (c & 31) - 1;
      if (which < 2 && rr <= TY && rr <= nrows) {
        const int y = y0 + rr;
        if (y >= 1 && y <= ny - 2) {
          if (which == 0 && x0 == 1)
            scratch[(rr + 1) * TFP_APITCH + 0] = scratch[(rr + 1) * TFP_APITCH + 1];
          if (which == 1 && xlast == nx - 2)
            scratch[(rr + 1) * TFP_APITCH + ncols + 1] = scratch[(rr + 1) * TFP_APITCH + ncols];
        }
      }
    }
    if (y0 == 1 || ylast == ny - 2) {
      asm volatile("bar.sync 1, %0;" ::"n"(tile::TX));
      // columns -1 .. ncols: thread c takes column c - 1, threads 0 and 1 also take ncols - 1 + 1 .. (TX + 2 columns, TX threads)
      for (int cc = c - 1; cc <= ncols; cc += tile::TX) {
        if (y0 == 1)
          scratch[0 * TFP_APITCH + cc + 1] = scratch[1 * TFP_APITCH + cc + 1];
        if (ylast == ny - 2)
          scratch[(nrows + 1) * TFP_APITCH + cc + 1] = scratch[nrows * TFP_APITCH + cc + 1];
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(tile::TX));
    // phase 2
    const int i0 = y0 * nx + x;
    if (all2)
      return pass2<true, FAST>(tv, maps, ad, nrows, col_ok, i0, nx, undef);
    return pass2<false, FAST>(tv, maps, ad, nrows, col_ok, i0, nx, undef);
  }
};

// undefined elements of each field whose flag is not ALL_DEFINED (the fused TFP's all2, see above)
__global__ void __launch_bounds__(256) count_undefined_kernel(const float* __restrict__ f, int n, int nfields, float undef, const FieldMeta* meta,
                                                              unsigned long long* counters)
{
  const int field = blockIdx.y;
  if (meta[field].all != 0)
    return;
  const float* p = f + (long long)field * n;
  unsigned bad = 0;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256)
    bad += is_def(p[i], undef) ? 0u : 1u;
  bad = __reduce_add_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0 && bad)
    atomicAdd(counters + field, (unsigned long long)bad);
}

// main kernel over the reference's flat range, then the border ring
template <class Op>
bool launch_stencil(Call& call, const Op& op, int nx, int ny, int nfields, float undef, const FieldMeta* meta, unsigned long long* counters,
                    bool flat_range = false)
{
  StencilGeom g;
  g.nx = nx;
  g.ny = ny;
  g.n = nx * ny;
  g.nfields = nfields;
  g.lo = flat_range ? 1 : nx;
  g.hi = flat_range ? g.n - 1 : g.n - nx;
  g.undef = undef;
  g.meta = meta;
  g.counters = counters;
  const int ring = 2 * nx + 2 * (ny - 2);
  const dim3 edge_grid((ring + tile::EDGE_THREADS - 1) / tile::EDGE_THREADS, nfields < 4096 ? nfields : 4096);
  // the tile engine stages every per-field input with one shift pattern: they must share their 16-byte alignment
  bool same_align = true;
  for (int k = 1; k < Op::NARR; ++k)
    same_align = same_align && ((reinterpret_cast<uintptr_t>(op.arr(k)) ^ reinterpret_cast<uintptr_t>(op.arr(0))) & 15) == 0;
  if (nx >= 34 && g.n >= 8192 && same_align) {
    // tiles staged in shared memory by the TMA unit (stencil_tile.cuh), then the border ring
    typedef tile::TileLayout<Op> L;
    tile::TileGeom t;
    t.nx = nx;
    t.ny = ny;
    t.n = g.n;
    t.nfields = nfields;
    t.tiles_x = (nx - 2 + tile::TX - 1) / tile::TX;
    t.tiles_y = (ny - 2 + Op::TY - 1) / Op::TY;
    t.period = (g.n % 4 == 0) ? 1 : (g.n % 2 == 0) ? 2 : 4;
    t.undef = undef;
    t.meta = meta;
    t.counters = counters;
    const long long tiles = (long long)t.tiles_x * t.tiles_y;
    // as many fields per CTA as possible (the maps are fetched once per CTA, the pipeline fills once)
    // while the grid still gives every SM a few CTAs
    t.fb = 16;
    while (t.fb > 1 && tiles * ((nfields + t.fb - 1) / t.fb) < 6LL * sm_count())
      t.fb /= 2;
    if (t.fb > nfields)
      t.fb = nfields;
    // pipeline depth: up to 3 fields in flight, two CTAs per SM
    // CTAs per SM (tile::tile_ctas<Op>: 3, gradient 4): measured per operator on B200 (profiles/r02at_tile_ctas.txt) -- with 2
    // CTAs of 9 warps the consumers mostly wait on each other's dependent instructions (ncu: `wait` 2.9 warps per issue,
    // occupancy 27 %); the third CTA costs a pipeline stage (shared memory) and 24 registers per thread and wins 3 - 19 %
    constexpr int CTAS = tile::tile_ctas<Op>::value;
    t.stages = t.fb < tile::MAX_STAGES ? t.fb : tile::MAX_STAGES;
    while (t.stages > 1 && L::smem_bytes(t.stages) > (CTAS == 4 ? 54 : CTAS == 3 ? 73 : 110) * 1024)
      t.stages -= 1;
    const size_t smem = L::smem_bytes(t.stages);
    // the opt-in to more than 48 KB of dynamic shared memory is a PER-DEVICE attribute of the kernel: one bit per
    // device and instantiation (a host thread per GPU is the recommended multi-GPU set-up)
    static std::atomic<unsigned long long> attr_set{0};
    int device = 0;
    if (!cuda_ok(cudaGetDevice(&device), "cudaGetDevice"))
      return false;
    const unsigned long long dev_bit = (device >= 0 && device < 64) ? (1ull << device) : 0ull;
    if (!(attr_set.load(std::memory_order_acquire) & dev_bit) || dev_bit == 0) {
      if (!cuda_ok(cudaFuncSetAttribute(tile::stencil_tile_kernel<Op, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::smem_bytes(tile::MAX_STAGES)),
                   "cudaFuncSetAttribute(stencil_tile_kernel)"))
        return false;
      attr_set.fetch_or(dev_bit, std::memory_order_release);
    }
    // one-dimensional grid, field blocks fastest (CTAs in flight share their tile's map ratios in L2): gridDim.x
    // holds 2^31-1 CTAs, so any grid with nx*ny < 2^31 fits
    t.field_blocks = (nfields + t.fb - 1) / t.fb;
    const long long ctas = tiles * t.field_blocks;
    if (ctas > 0x7fffffffLL) {
      set_error("fcb200: batch too large for one launch (%lld CTAs)", ctas);
      return false;
    }
    tile::stencil_tile_kernel<Op, CTAS><<<(unsigned)ctas, tile::TILE_THREADS, smem, call.stream()>>>(op, t);
    tile::stencil_edge_kernel<Op><<<edge_grid, tile::EDGE_THREADS, 0, call.stream()>>>(op, nx, ny, nfields, g.lo, g.hi, undef, meta, counters, true);
    count_launch(2);
    return true;
  }
  // small grids: the flat scalar kernel (it counts the ring cells of its range itself), then the border ring
  g.group = nfields < ST_GROUP ? nfields : ST_GROUP;
  const int groups = (nfields + g.group - 1) / g.group;
  if (groups > 65535) {
    set_error("fcb200: batch too large for one launch (%d fields)", nfields);
    return false;
  }
  const int per_chunk = ST_THREADS * ST_UNROLL;
  const int total_chunks = (g.hi - g.lo + per_chunk - 1) / per_chunk;
  // gridDim.y <= 65535: a longer range takes several launches
  for (int c0 = 0; c0 < total_chunks; c0 += 65535) {
    g.chunk_base = c0;
    g.chunks = total_chunks - c0 < 65535 ? total_chunks - c0 : 65535;
    const dim3 grid((unsigned)g.group, (unsigned)g.chunks, (unsigned)groups);
    stencil_kernel<Op><<<grid, ST_THREADS, 0, call.stream()>>>(op, g);
    count_launch();
  }
  tile::stencil_edge_kernel<Op><<<edge_grid, tile::EDGE_THREADS, 0, call.stream()>>>(op, nx, ny, nfields, g.lo, g.hi, undef, meta, counters, false);
  count_launch();
  return true;
}

// thermalFrontParameter in one pass (tfp_tile.cuh): interior points by the marching tile kernel, then the border ring
template <class Op>
bool launch_tfp_tile(Call& call, const Op& op, int nx, int ny, int nfields, float undef, const FieldMeta* meta, unsigned long long* counters)
{
  tfp2::Geom t;
  t.nx = nx;
  t.ny = ny;
  t.n = nx * ny;
  t.nfields = nfields;
  t.tiles_x = (nx - 2 + tfp2::TXO - 1) / tfp2::TXO;
  t.tiles_y = (ny - 2 + tfp2::TY - 1) / tfp2::TY;
  t.period = (t.n % 4 == 0) ? 1 : (t.n % 2 == 0) ? 2 : 4;
  t.undef = undef;
  t.meta = meta;
  t.counters = counters;
  const long long tiles = (long long)t.tiles_x * t.tiles_y;
  t.fb = 16;
  while (t.fb > 1 && tiles * ((nfields + t.fb - 1) / t.fb) < 6LL * sm_count())
    t.fb /= 2;
  if (t.fb > nfields)
    t.fb = nfields;
  t.stages = tfp2::STAGES;
  const size_t smem = tfp2::smem_bytes();
  static std::atomic<unsigned long long> attr_set{0}; // per device (see launch_stencil)
  int device = 0;
  if (!cuda_ok(cudaGetDevice(&device), "cudaGetDevice"))
    return false;
  const unsigned long long dev_bit = (device >= 0 && device < 64) ? (1ull << device) : 0ull;
  if (!(attr_set.load(std::memory_order_acquire) & dev_bit) || dev_bit == 0) {
    if (!cuda_ok(cudaFuncSetAttribute(tfp2::tfp_tile_kernel<Op>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tfp2::smem_bytes()),
                 "cudaFuncSetAttribute(tfp_tile_kernel)"))
      return false;
    attr_set.fetch_or(dev_bit, std::memory_order_release);
  }
  t.field_blocks = (nfields + t.fb - 1) / t.fb;
  const long long ctas = tiles * t.field_blocks;
  if (ctas > 0x7fffffffLL) {
    set_error("fcb200: batch too large for one launch (%lld CTAs)", ctas);
    return false;
  }
  tfp2::tfp_tile_kernel<Op><<<(unsigned)ctas, tfp2::THREADS, smem, call.stream()>>>(op, t);
  const int ring = 2 * nx + 2 * (ny - 2);
  const dim3 edge_grid((ring + tile::EDGE_THREADS - 1) / tile::EDGE_THREADS, nfields < 4096 ? nfields : 4096);
  tile::stencil_edge_kernel<Op><<<edge_grid, tile::EDGE_THREADS, 0, call.stream()>>>(op, nx, ny, nfields, nx, t.n - nx, undef, meta, counters, true);
  count_launch(2);
  return true;
}

bool grid_ok(int nx, int ny, int nfields)
{
  if (nfields <= 0 || (long long)nx * ny >= 0x7fffffffLL) {
    set_error("fcb200: invalid grid or batch size (nx=%d ny=%d nfields=%d)", nx, ny, nfields);
    return false;
  }
  return true;
}

const FieldMeta* flags_to_meta(Call& call, const int* fDefined, int nfields)
{
  FieldMeta* meta = call.meta_host(nfields);
  if (!call.ok())
    return nullptr;
  for (int k = 0; k < nfields; ++k) {
    meta[k].all = (fDefined[k] == ALL_DEFINED) ? 1 : 0;
    meta[k].a = meta[k].b = meta[k].c = 0.f;
  }
  return call.upload_meta();
}

// ------------------------------------------------------------------------------------ shapiro2_filter
// FC.cc:2076-2179.  Two iterations of (x pass, y pass); each pass stores float.  Mathematically a
// separable 2-D filter with these boundary rules (the flat-loop wrap never survives, FC.cc:2117-2120):
//   x pass: out(x,y) = filtered for 1 <= x <= nx-2 (EVERY row), copy for x = 0, nx-1
//   y pass: out(x,y) = filtered for 1 <= y <= ny-2 (EVERY column), copy for y = 0, ny-1
//   all-defined branch: s = +0.25 then -0.25; f + s*(fl + fr - 2.*f): float sum, the rest in double.
//   masked branch: weights 0.25/0 from 3-point definedness of the ORIGINAL field, BOTH iterations use
//   +0.25 (the `s = -0.25` update is dead, FC.cc:2136-2168), all-float arithmetic.
//
// All four passes run in REGISTERS: a thread owns W adjacent columns (W = 4: one float4 per row when
// every row is 16-byte aligned, four 4-byte accesses otherwise; W = 1 is kept for comparison) and marches down a band of rows.  x passes take their
// neighbour columns from the adjacent lanes by warp shuffle; y passes use a three-row window of the
// previous pass kept in registers (the pipeline lags two rows behind the load).  A warp strip of 32*W
// loaded columns yields (32 - 2*HL)*W final columns: the HL outermost lanes on each side are halo
// (the result at column x depends on the input columns x-2 .. x+2); a band of RB rows loads RB+4.
// The field is read once and written once: 8 B/point.
constexpr int SH_WARPS = 4;
#ifndef FCB_SH_VEC_ALL_CTAS
#define FCB_SH_VEC_ALL_CTAS 8
#endif
#ifndef FCB_SH_ANYCOL_CTAS
#define FCB_SH_ANYCOL_CTAS 5
#endif
#ifndef FCB_SH_ANYCOL_ALL_CTAS
#define FCB_SH_ANYCOL_ALL_CTAS 6
#endif
#ifndef FCB_SH_PF
#define FCB_SH_PF 3
#endif
constexpr int SH_PF = FCB_SH_PF; // rows in flight per thread (tools/shape_variants.sh)

template <int W>
struct ShapiroGeom
{
  static constexpr int HL = (W == 4) ? 1 : 2;           // halo lanes per side
  static constexpr int USEFUL = (32 - 2 * HL) * W;      // output columns per warp strip
};

// One point of one pass of the all-defined branch: (float)(f + s*((lo + hi) - 2.*f)), float sum, the rest in double
// (FC.cc:2113, 2121).  The three float<->double conversions per point and pass run on the quarter-rate XU
// pipe (16/clk/SM on B200): with all four passes in double the kernel is XU-bound at a third of the HBM
// roofline.  FLOATPATH evaluates the same value in float: T = S - 2f (S = lo + hi) is checked for
// exactness with Knuth's TwoSum; if it is exact (smooth data: always), f + s*T is a sum of two floats,
// for which rounding through double and rounding once agree, and that single rounding is fmaf(s, T, f).
// Anything else -- inexact T, NaN, infinities -- takes the double expression.  All four passes take the float form; the
// TwoSum test itself is the second level behind the one-comparison test of shapiro_pair_sterbenz below.
// (every form takes S = RN(lo + hi), the float sum the reference forms first)
__device__ __forceinline__ float shapiro_point_double(float S, float f, double sd)
{
  return (float)fma(sd, fma(-2.0, (double)f, (double)S), (double)f);
}
// float evaluation; `inexact` is raised when the result must not be used
__device__ __forceinline__ float shapiro_point_float(float S, float f, float sf, bool& inexact)
{
  const float b = f + f;
  const float T = S - b;
  const float a1 = T + b;    // TwoSum(S, -b): a1 ~ S
  const float b1 = T - a1;   //                b1 ~ -b
  const float da = S - a1;
  const float db = (-b) - b1;
  const float err = da + db; // the rounding error of T, exactly (NaN if anything overflowed)
  inexact = inexact || !(err == 0.f);
  return __fmaf_rn(sf, T, f);
}
// The same value with a ONE-comparison exactness test: |T| < |f| (T = RN(S - b), b = 2f) means that the real difference is
// smaller than |b| / 2 too (rounding is monotone and |f| is a float), so S lies strictly between b/2 and 3b/2 and S - b is
// exact by Sterbenz' lemma.  True wherever the second difference of the field is smaller than the field itself -- everywhere
// but next to zeros of the field; a warp with a point that fails (or with a NaN / infinity: every comparison false) repeats the
// pass with the TwoSum test above, which decides between the float and the double expression.  On the packed FP32 instructions
// (FADD2 / FFMA2: one issue slot per pair of columns; the same IEEE operations -- there is no product followed by an add here
// for ptxas to contract): 3 packed + 2 scalar instructions per pair of points after S, against 8 + 2 with the TwoSum test.
__device__ __forceinline__ float2 shapiro_pair_sterbenz(float2 S, float2 f, float sf, bool& unsure)
{
  const float2 b = __fadd2_rn(f, f);
  const float2 T = __fadd2_rn(S, make_float2(-b.x, -b.y));
  unsure = unsure || !(fabsf(T.x) < fabsf(f.x)) || !(fabsf(T.y) < fabsf(f.y));
  return __ffma2_rn(make_float2(sf, sf), T, f);
}
__device__ __forceinline__ float shapiro_point_sterbenz(float S, float f, float sf, bool& unsure)
{
  const float T = S - (f + f);
  unsure = unsure || !(fabsf(T) < fabsf(f));
  return __fmaf_rn(sf, T, f);
}

template <int W, bool ALL, bool FLOATPATH>
__device__ __forceinline__ void shapiro_pass(const float (&S)[W], const float (&f)[W], float (&r)[W], float s, unsigned wbits)
{
  if (!ALL) {
    // f + w * (S - 2.f * f) in float, w = 0.25 or 0 (FC.cc:2136-2168).  2.f * f = f + f, and w * T is exact (a power of two or
    // zero; NaN for 0 * inf either way), so the sum is the single rounding fmaf(w, T, f) -- three packed instructions per pair
    // of points instead of two multiplications and three additions per point.  (Only a |T| below 2^-124 next to an f of that
    // size could tell the difference: 0.25 * T then rounds in the subnormal range before the addition.)
    if constexpr (W == 4) {
#pragma unroll
      for (int j = 0; j < W; j += 2) {
        const float2 ff = make_float2(f[j], f[j + 1]);
        const float2 w = make_float2(((wbits >> j) & 1u) ? 0.25f : 0.f, ((wbits >> (j + 1)) & 1u) ? 0.25f : 0.f);
        const float2 b = __fadd2_rn(ff, ff);
        const float2 T = __fadd2_rn(make_float2(S[j], S[j + 1]), make_float2(-b.x, -b.y));
        const float2 v = __ffma2_rn(w, T, ff);
        r[j] = v.x;
        r[j + 1] = v.y;
      }
    } else {
#pragma unroll
      for (int j = 0; j < W; ++j) {
        const float w = ((wbits >> j) & 1u) ? 0.25f : 0.f;
        r[j] = __fmaf_rn(w, S[j] - (f[j] + f[j]), f[j]);
      }
    }
  } else if (FLOATPATH) {
    bool unsure = false;
    if constexpr (W == 4) {
#pragma unroll
      for (int j = 0; j < W; j += 2) {
        const float2 v = shapiro_pair_sterbenz(make_float2(S[j], S[j + 1]), make_float2(f[j], f[j + 1]), s, unsure);
        r[j] = v.x;
        r[j + 1] = v.y;
      }
    } else {
#pragma unroll
      for (int j = 0; j < W; ++j)
        r[j] = shapiro_point_sterbenz(S[j], f[j], s, unsure);
    }
    if (__any_sync(0xffffffffu, unsure)) { // next to a zero of the field (warp-uniform: real branches, not selects)
      bool inexact = false;
#pragma unroll
      for (int j = 0; j < W; ++j)
        r[j] = shapiro_point_float(S[j], f[j], s, inexact);
      if (__any_sync(0xffffffffu, inexact)) { // rare
#pragma unroll
        for (int j = 0; j < W; ++j)
          r[j] = shapiro_point_double(S[j], f[j], (double)s);
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < W; ++j)
      r[j] = shapiro_point_double(S[j], f[j], (double)s);
  }
}

template <int W, bool ALL, bool FLOATPATH, bool ANYCOL = false>
__device__ __forceinline__ void shapiro_xpass(const float (&f)[W], float (&out)[W], float s, unsigned wbits, unsigned copybits)
{
  // scalar sums: the neighbour pairs (left, f0), (f1, f2), (f3, right) are not register pairs of the row -- as packed operands
  // they cost six moves per pass
  float S[W], r[W];
  const float left = __shfl_up_sync(0xffffffffu, f[W - 1], 1);
  const float right = __shfl_down_sync(0xffffffffu, f[0], 1);
#pragma unroll
  for (int j = 0; j < W; ++j)
    S[j] = (j == 0 ? left : f[j > 0 ? j - 1 : 0]) + (j == W - 1 ? right : f[j < W - 1 ? j + 1 : 0]);
  shapiro_pass<W, ALL, FLOATPATH>(S, f, r, s, wbits);
  // columns 0 and nx-1 are copied.  W = 4 with float4 rows: x0 and nx are multiples of 4, so they can only be the lane's
  // first resp. last column (ANYCOL: rows of any alignment, any of the lane's columns)
#pragma unroll
  for (int j = 0; j < W; ++j)
    out[j] = ((W == 1 || ANYCOL || j == 0 || j == W - 1) && ((copybits >> j) & 1u)) ? f[j] : r[j];
}

template <int W, bool ALL, bool FLOATPATH>
__device__ __forceinline__ void shapiro_ypass(const float (&lo)[W], const float (&f)[W], const float (&hi)[W], float (&out)[W], float s, unsigned wbits,
                                              bool copy)
{
  // rows 0 and ny-1 (and rows outside the grid) are copied: selected afterwards -- as a branch the copy is hoisted above it and
  // costs every row eight moves
  float S[W], r[W];
  if constexpr (W == 4) {
#pragma unroll
    for (int j = 0; j < W; j += 2) {
      const float2 v = __fadd2_rn(make_float2(lo[j], lo[j + 1]), make_float2(hi[j], hi[j + 1]));
      S[j] = v.x;
      S[j + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int j = 0; j < W; ++j)
      S[j] = lo[j] + hi[j];
  }
  shapiro_pass<W, ALL, FLOATPATH>(S, f, r, s, wbits);
#pragma unroll
  for (int j = 0; j < W; ++j)
    out[j] = copy ? f[j] : r[j];
}

// VEC = the lane's W = 4 columns are one float4 (every row on a 16-byte boundary).  W = 4 without VEC: rows of ANY alignment (an odd
// row length: MEPS) -- the same register arithmetic on four adjacent columns, but four 4-byte loads and stores per row, each
// column with its own bounds test (the lane's columns may straddle the grid's last column).  The accesses of a warp then have a
// 16-byte stride: four times the L1 wavefronts of the float4 form, which this kernel has room for (it is issue-bound), against
// the one-column-per-lane form (W = 1), which needs twice the instructions per point.
template <int W, bool ALL, bool VEC = (W == 4)>
__device__ __forceinline__ void shapiro_band(const float* __restrict__ src, float* __restrict__ dst, int nx, int ny, int x0, int r0, int r1, float undef)
{
  constexpr int HL = ShapiroGeom<W>::HL;
  constexpr bool ANYCOL = (W == 4) && !VEC;
  const int lane = threadIdx.x & 31;
  const bool col_ok = ANYCOL ? (x0 + W > 0 && x0 < nx) : (x0 >= 0 && x0 + W <= nx); // VEC: nx % 4 == 0 and x0 % 4 == 0, a float4 is entirely in or out
  const bool store_lane = col_ok && lane >= HL && lane < 32 - HL;
  unsigned okbits = 0; // ANYCOL: the lane's columns inside the grid
#pragma unroll
  for (int j = 0; j < W; ++j)
    if (x0 + j >= 0 && x0 + j < nx)
      okbits |= 1u << j;
  unsigned copybits = 0; // columns 0 and nx-1 are copied by the x passes
#pragma unroll
  for (int j = 0; j < W; ++j)
    if (x0 + j <= 0 || x0 + j >= nx - 1)
      copybits |= 1u << j;

  const int rbeg = r0 - 2, rend = r1 + 2; // rows marched: [rbeg, rend), stores lag two rows
  // Loads are unconditional: a row outside the grid (or past the band) is read as the nearest row inside, a column outside as
  // the nearest column inside.  What is computed from such values never reaches a stored point: rows -1 and ny only feed the
  // y passes of rows 0 and ny-1, columns -1 and nx only the x passes of columns 0 and nx-1, and those are copies; the masked
  // branch takes its definedness bits from the row / column numbers as before.  (Zero fill under a predicate had cost every
  // row a register copy of the loaded values.)
  const int rlast = min(ny, rend) - 1;
  int xc[ANYCOL ? W : 1];
  if (ANYCOL) {
#pragma unroll
    for (int j = 0; j < W; ++j)
      xc[ANYCOL ? j : 0] = min(max(x0 + j, 0), nx - 1);
  } else {
    xc[0] = min(max(x0, 0), nx - W);
  }
  auto load_row = [&](int r, float (&q)[W]) {
    const float* p = src + (long long)min(max(r, 0), rlast) * nx;
    if (ANYCOL) {
#pragma unroll
      for (int j = 0; j < W; ++j)
        q[j] = p[xc[ANYCOL ? j : 0]];
    } else if (W == 4) {
      const float4 t = *reinterpret_cast<const float4*>(p + xc[0]);
      q[0] = t.x;
      q[W > 1 ? 1 : 0] = t.y;
      q[W > 2 ? 2 : 0] = t.z;
      q[W > 3 ? 3 : 0] = t.w;
    } else {
      q[0] = p[xc[0]];
    }
  };

  // The march is unrolled over 2 * SH_PF = 6 rows with everything that rotates in statically indexed slots: two row buffers
  // that alternate between "being loaded" and "being consumed", and the three-row windows of the two iterations (a, c), the
  // definedness words and the weights of the previous rows in slots numbered by the row's position in the body (6 is a
  // multiple of every period), so that no value is ever moved from one register to another to change its role.
  static_assert((2 * SH_PF) % 3 == 0 && (2 * SH_PF) % 2 == 0, "the unrolled body must close the three-row rotation");
  float qa[SH_PF][W], qb[SH_PF][W];
#pragma unroll
  for (int k = 0; k < SH_PF; ++k)
    load_row(rbeg + k, qa[k]);

  float a[3][W] = {}, c[3][W] = {};
  unsigned dd[3] = {0, 0, 0};  // definedness bits of the original rows (slot = position % 3)
  unsigned mxs[2] = {0, 0};    // x weights (slot = position % 2): this row's and the previous row's
  unsigned mys[2] = {0, 0};    // y weights

  // position k of the body (a compile-time number after unrolling), row r, loaded values f
  auto do_row = [&](const int k, const int r, const float (&f)[W]) {
    const int s2 = k % 3, s1 = (k + 2) % 3, s0 = (k + 1) % 3; // slots of rows r, r-1, r-2
    const int t1 = k % 2, t0 = (k + 1) % 2;
    if (!ALL) {
      unsigned d2 = 0;
      if (r >= 0 && r < ny) {
#pragma unroll
        for (int j = 0; j < W; ++j)
          if (is_def(f[j], undef))
            d2 |= 1u << j;
      }
      const unsigned dl = (__shfl_up_sync(0xffffffffu, d2, 1) >> (W - 1)) & 1u;
      const unsigned dr = __shfl_down_sync(0xffffffffu, d2, 1) & 1u;
      const unsigned ext = dl | (d2 << 1) | (dr << (W + 1));
      dd[s2] = d2;
      mxs[t1] = ext & (ext >> 1) & (ext >> 2);
      mys[t1] = dd[s0] & dd[s1] & d2;
    }
    // iteration 1: x pass on row r, y pass on row r-1
    float b[W], d[W];
    shapiro_xpass<W, ALL, true, ANYCOL>(f, a[s2], 0.25f, mxs[t1], copybits);
    shapiro_ypass<W, ALL, true>(a[s0], a[s1], a[s2], b, 0.25f, mys[t1], r - 1 <= 0 || r - 1 >= ny - 1);
    // iteration 2: x pass on row r-1, y pass on row r-2
    shapiro_xpass<W, ALL, true, ANYCOL>(b, c[s2], -0.25f, mxs[t0], copybits);
    shapiro_ypass<W, ALL, W == 4>(c[s0], c[s1], c[s2], d, -0.25f, mys[t0], r - 2 <= 0 || r - 2 >= ny - 1); // (W = 4: the packed float form for all four passes)
    const int ro = r - 2;
    if (store_lane && ro >= r0 && ro < r1) {
      float* p = dst + (long long)ro * nx + x0;
      if (ANYCOL) {
#pragma unroll
        for (int j = 0; j < W; ++j)
          if ((okbits >> j) & 1u)
            p[j] = d[j];
      } else if (W == 4)
        *reinterpret_cast<float4*>(p) = make_float4(d[0], d[W > 1 ? 1 : 0], d[W > 2 ? 2 : 0], d[W > 3 ? 3 : 0]);
      else
        *p = d[0];
    }
  };

  for (int rb = rbeg; rb < rend; rb += 2 * SH_PF) {
#pragma unroll
    for (int k = 0; k < SH_PF; ++k)
      load_row(rb + SH_PF + k, qb[k]);
#pragma unroll
    for (int k = 0; k < SH_PF; ++k) {
      if (rb + k >= rend)
        return;
      do_row(k, rb + k, qa[k]);
    }
#pragma unroll
    for (int k = 0; k < SH_PF; ++k)
      load_row(rb + 2 * SH_PF + k, qa[k]);
#pragma unroll
    for (int k = 0; k < SH_PF; ++k) {
      if (rb + SH_PF + k >= rend)
        return;
      do_row(SH_PF + k, rb + SH_PF + k, qb[k]);
    }
  }
}

// ONLY_ALL: every field of the batch is flagged all-defined (known on the host) -- the kernel then holds the all-defined march
// alone and gets its register budget
template <int W, bool VEC = (W == 4), bool ONLY_ALL = false>
__device__ __forceinline__ void shapiro2_body(const float* __restrict__ fin, float* __restrict__ fout, int nx, int ny, int strips,
                                                                int bands, int rows_per_band, const FieldMeta* meta, float undef)
{
  constexpr int HL = ShapiroGeom<W>::HL;
  constexpr int USEFUL = ShapiroGeom<W>::USEFUL;
  const int strip_groups = (strips + SH_WARPS - 1) / SH_WARPS;
  const int per_field = strip_groups * bands;
  const int field = blockIdx.x / per_field;
  const int t = blockIdx.x - field * per_field;
  const int band = t / strip_groups;
  const int strip = (t - band * strip_groups) * SH_WARPS + (threadIdx.x >> 5);
  if (strip >= strips)
    return; // whole warp
  const int lane = threadIdx.x & 31;
  const int x0 = strip * USEFUL + (lane - HL) * W;
  const int r0 = band * rows_per_band;
  const int r1 = min(r0 + rows_per_band, ny);
  const float* src = fin + (long long)field * nx * ny;
  float* dst = fout + (long long)field * nx * ny;
  if (ONLY_ALL || meta[field].all != 0)
    shapiro_band<W, true, VEC>(src, dst, nx, ny, x0, r0, r1, undef);
  else
    shapiro_band<W, false, VEC>(src, dst, nx, ny, x0, r0, r1, undef);
}

template <int W, bool ONLY_ALL>
__global__ void __launch_bounds__(SH_WARPS * 32) shapiro2_kernel(const float* __restrict__ fin, float* __restrict__ fout, int nx, int ny, int strips, int bands,
                                                                int rows_per_band, const FieldMeta* meta, float undef)
{
  shapiro2_body<W, W == 4, ONLY_ALL>(fin, fout, nx, ny, strips, bands, rows_per_band, meta, undef);
}
// float4 rows, every field all-defined: 8 CTAs per SM at 64 registers (ECMWF x 32: 0.78 -> 0.82 of the roofline against the
// 66 registers / 7 CTAs the compiler picks on its own)
__global__ void __launch_bounds__(SH_WARPS * 32, FCB_SH_VEC_ALL_CTAS) shapiro2_kernel_vec_all(const float* __restrict__ fin, float* __restrict__ fout, int nx, int ny,
                                                                                             int strips, int bands, int rows_per_band, const FieldMeta* meta,
                                                                                             float undef)
{
  shapiro2_body<4, true, true>(fin, fout, nx, ny, strips, bands, rows_per_band, meta, undef);
}
// rows of any alignment: their own register budgets (four separate loads and stores per row keep more addresses alive)
__global__ void __launch_bounds__(SH_WARPS * 32, FCB_SH_ANYCOL_CTAS) shapiro2_kernel_anycol(const float* __restrict__ fin, float* __restrict__ fout, int nx, int ny,
                                                                                           int strips, int bands, int rows_per_band, const FieldMeta* meta,
                                                                                           float undef)
{
  shapiro2_body<4, false, false>(fin, fout, nx, ny, strips, bands, rows_per_band, meta, undef);
}
__global__ void __launch_bounds__(SH_WARPS * 32, FCB_SH_ANYCOL_ALL_CTAS) shapiro2_kernel_anycol_all(const float* __restrict__ fin, float* __restrict__ fout, int nx,
                                                                                                   int ny, int strips, int bands, int rows_per_band,
                                                                                                   const FieldMeta* meta, float undef)
{
  shapiro2_body<4, false, true>(fin, fout, nx, ny, strips, bands, rows_per_band, meta, undef);
}

} // namespace
} // namespace fcb200

// =========================================================================================== C-ABI
using namespace fcb200;

namespace {

template <class Op>
int run_stencil(Call& call, const Op& op, int nx, int ny, int nfields, int* fDefined, float undef, unsigned long long denom, bool flat_range = false)
{
  const FieldMeta* meta = flags_to_meta(call, fDefined, nfields);
  unsigned long long* counters = call.counters(nfields);
  if (!call.ok())
    return -1;
  if (!launch_stencil(call, op, nx, ny, nfields, undef, meta, counters, flat_range))
    return -1;
  return call.finish_counted(fDefined, nfields, denom);
}

template <int MODE>
int vortdiv(int nx, int ny, int nfields, const float* u, const float* v, const float* xmapr, const float* ymapr, const float* fcoriolis, float* out,
            int* fDefined, float undef)
{
  if (nx < 3 || ny < 3)
    return 0;
  if (!grid_ok(nx, ny, nfields))
    return -1;
  const size_t n = (size_t)nx * ny;
  Call call;
  VortDivOp<MODE> op;
  op.u = call.in(u, n * nfields);
  op.v = call.in(v, n * nfields);
  op.xm = call.in(xmapr, n);
  op.ym = call.in(ymapr, n);
  op.fc = (MODE == 1) ? call.in(fcoriolis, n) : nullptr;
  op.o = call.out(out, n * nfields);
  return run_stencil(call, op, nx, ny, nfields, fDefined, undef, n - 2 * (size_t)nx);
}

template <int MODE>
int geostrophic(int nx, int ny, int nfields, const float* z, const float* xmapr, const float* ymapr, const float* fcoriolis, float* out, int* fDefined,
                float undef)
{ // FC.cc:638-743.  plevelgwind_ycomp has no `nx < 3 || ny < 3` test in the reference and then reads and writes out of
  // bounds in fillEdges (undefined behaviour); it is rejected here like its two siblings.
  if (nx < 3 || ny < 3)
    return 0;
  if (!grid_ok(nx, ny, nfields))
    return -1;
  const size_t n = (size_t)nx * ny;
  Call call;
  GeoOp<MODE> op;
  op.z = call.in(z, n * nfields);
  op.xm = call.in(xmapr, n);
  op.ym = call.in(ymapr, n);
  op.fc = call.in(fcoriolis, n);
  op.o = call.out(out, n * nfields);
  if (MODE != 0)
    return run_stencil(call, op, nx, ny, nfields, fDefined, undef, n - 2 * (size_t)nx);
  // plevelgwind_xcomp increments n_undefined for EVERY point of the loop (:664 is outside the else): the flag is
  // checkDefined(N - 2nx, N - 2nx) = NONE_DEFINED whatever the data
  const FieldMeta* meta = flags_to_meta(call, fDefined, nfields);
  unsigned long long* counters = call.counters(nfields);
  if (!call.ok())
    return -1;
  if (!launch_stencil(call, op, nx, ny, nfields, undef, meta, counters, false))
    return -1;
  return call.finish([=](const unsigned long long*) {
    for (int k = 0; k < nfields; ++k)
      fDefined[k] = NONE_DEFINED;
  });
}

template <int COMP>
int qvector(int nx, int ny, int nfields, const float* z, const float* t, const float* xmapr, const float* ymapr, const float* fcoriolis, float tscale,
            float c, float* qcomp, int* fDefined, float undef)
{ // FC.cc:555-595: plevelgwind_xcomp and _ycomp into two scratch fields, then the second pass.  xcomp is given the caller's
  // flag and leaves NONE_DEFINED behind (its quirk), so ycomp always runs with every test on.
  const size_t n = (size_t)nx * ny;
  Call call;
  const float* d_z = call.in(z, n * nfields);
  const float* d_t = call.in(t, n * nfields);
  const float* d_xm = call.in(xmapr, n);
  const float* d_ym = call.in(ymapr, n);
  const float* d_fc = call.in(fcoriolis, n);
  float* d_out = call.out(qcomp, n * nfields);
  float* d_ug = static_cast<float*>(call.scratch(sizeof(float) * n * nfields));
  float* d_vg = static_cast<float*>(call.scratch(sizeof(float) * n * nfields));
  unsigned long long* counters = call.counters(3 * nfields);
  if (!call.ok())
    return -1;
  const FieldMeta* meta_in = flags_to_meta(call, fDefined, nfields);
  std::vector<int> none((size_t)nfields, (int)NONE_DEFINED);
  const FieldMeta* meta_none = flags_to_meta(call, none.data(), nfields);
  if (!call.ok())
    return -1;
  GeoOp<0> gx;
  gx.z = d_z, gx.xm = d_xm, gx.ym = d_ym, gx.fc = d_fc, gx.o = d_ug;
  GeoOp<1> gy;
  gy.z = d_z, gy.xm = d_xm, gy.ym = d_ym, gy.fc = d_fc, gy.o = d_vg;
  QvecOp<COMP> q;
  q.ug = d_ug, q.vg = d_vg, q.t = d_t, q.xm = d_xm, q.ym = d_ym, q.tscale = tscale, q.c = c, q.o = d_out;
  if (!launch_stencil(call, gx, nx, ny, nfields, undef, meta_in, counters) || !launch_stencil(call, gy, nx, ny, nfields, undef, meta_none, counters + nfields) ||
      !launch_stencil(call, q, nx, ny, nfields, undef, meta_none, counters + 2 * nfields))
    return -1;
  return call.finish_counted(fDefined, nfields, n - 2 * (size_t)nx, 2 * nfields);
}

} // namespace

extern "C" {

int fcb200_relvort_batched(int nx, int ny, int nfields, const float* u, const float* v, const float* xmapr, const float* ymapr, float* rvort,
                           int* fDefined, float undef)
{
  return vortdiv<0>(nx, ny, nfields, u, v, xmapr, ymapr, nullptr, rvort, fDefined, undef);
}
int fcb200_relvort(int nx, int ny, const float* u, const float* v, const float* xmapr, const float* ymapr, float* rvort, int* fDefined, float undef)
{
  return vortdiv<0>(nx, ny, 1, u, v, xmapr, ymapr, nullptr, rvort, fDefined, undef);
}

int fcb200_absvort_batched(int nx, int ny, int nfields, const float* u, const float* v, const float* xmapr, const float* ymapr,
                           const float* fcoriolis, float* avort, int* fDefined, float undef)
{
  return vortdiv<1>(nx, ny, nfields, u, v, xmapr, ymapr, fcoriolis, avort, fDefined, undef);
}
int fcb200_absvort(int nx, int ny, const float* u, const float* v, const float* xmapr, const float* ymapr, const float* fcoriolis, float* avort,
                   int* fDefined, float undef)
{
  return vortdiv<1>(nx, ny, 1, u, v, xmapr, ymapr, fcoriolis, avort, fDefined, undef);
}

int fcb200_divergence_batched(int nx, int ny, int nfields, const float* u, const float* v, const float* xmapr, const float* ymapr, float* diverg,
                              int* fDefined, float undef)
{
  return vortdiv<2>(nx, ny, nfields, u, v, xmapr, ymapr, nullptr, diverg, fDefined, undef);
}
int fcb200_divergence(int nx, int ny, const float* u, const float* v, const float* xmapr, const float* ymapr, float* diverg, int* fDefined,
                      float undef)
{
  return vortdiv<2>(nx, ny, 1, u, v, xmapr, ymapr, nullptr, diverg, fDefined, undef);
}

int fcb200_advection_batched(int nx, int ny, int nfields, const float* f, const float* u, const float* v, const float* xmapr, const float* ymapr,
                             float hours, float* advec, int* fDefined, float undef)
{ // FC.cc:1942-1983
  if (nx < 3 || ny < 3)
    return 0;
  if (!grid_ok(nx, ny, nfields))
    return -1;
  const size_t n = (size_t)nx * ny;
  Call call;
  AdvectionOp op;
  op.f = call.in(f, n * nfields);
  op.u = call.in(u, n * nfields);
  op.v = call.in(v, n * nfields);
  op.xm = call.in(xmapr, n);
  op.ym = call.in(ymapr, n);
  op.scale = (float)(-3600. * hours);
  op.o = call.out(advec, n * nfields);
  return run_stencil(call, op, nx, ny, nfields, fDefined, undef, n - 2 * (size_t)nx);
}
int fcb200_advection(int nx, int ny, const float* f, const float* u, const float* v, const float* xmapr, const float* ymapr, float hours,
                     float* advec, int* fDefined, float undef)
{
  return fcb200_advection_batched(nx, ny, 1, f, u, v, xmapr, ymapr, hours, advec, fDefined, undef);
}

int fcb200_gradient_batched(int nx, int ny, int nfields, const float* field, const float* xmapr, const float* ymapr, int compute, float* fgrad,
                            int* fDefined, float undef)
{ // FC.cc:1985-2074
  if (nx < 3 || ny < 3)
    return 0;
  if (compute < 1 || compute > 4)
    return 0;
  if (!grid_ok(nx, ny, nfields))
    return -1;
  const size_t n = (size_t)nx * ny;
  Call call;
  const float* d_f = call.in(field, n * nfields);
  // the reference only dereferences the map ratio(s) the mode needs; keep that so callers may pass the other as null
  const float* d_xm = (compute != 2) ? call.in(xmapr, n) : nullptr;
  const float* d_ym = (compute != 1) ? call.in(ymapr, n) : nullptr;
  float* d_o = call.out(fgrad, n * nfields);
  const unsigned long long denom = n - 2 * (size_t)nx; // for every mode, also c=1 whose loop is [1, N-1) (FC.cc:2014, 2068)
  switch (compute) {
  case 1:
    return run_stencil(call, GradientOp<1>{d_f, d_xm, d_ym, d_o}, nx, ny, nfields, fDefined, undef, denom, true);
  case 2:
    return run_stencil(call, GradientOp<2>{d_f, d_xm, d_ym, d_o}, nx, ny, nfields, fDefined, undef, denom);
  case 3:
    return run_stencil(call, GradientOp<3>{d_f, d_xm, d_ym, d_o}, nx, ny, nfields, fDefined, undef, denom);
  default:
    return run_stencil(call, GradientOp<4>{d_f, d_xm, d_ym, d_o}, nx, ny, nfields, fDefined, undef, denom);
  }
}
int fcb200_gradient(int nx, int ny, const float* field, const float* xmapr, const float* ymapr, int compute, float* fgrad, int* fDefined,
                    float undef)
{
  return fcb200_gradient_batched(nx, ny, 1, field, xmapr, ymapr, compute, fgrad, fDefined, undef);
}

int fcb200_jacobian_batched(int nx, int ny, int nfields, const float* field1, const float* field2, const float* xmapr, const float* ymapr,
                            float* fjacobian, int* fDefined, float undef)
{ // FC.cc:2424-2460
  if (nx < 3 || ny < 3)
    return 0;
  if (!grid_ok(nx, ny, nfields))
    return -1;
  const size_t n = (size_t)nx * ny;
  Call call;
  JacobianOp op;
  op.f1 = call.in(field1, n * nfields);
  op.f2 = call.in(field2, n * nfields);
  op.xm = call.in(xmapr, n);
  op.ym = call.in(ymapr, n);
  op.o = call.out(fjacobian, n * nfields);
  return run_stencil(call, op, nx, ny, nfields, fDefined, undef, n - 2 * (size_t)nx);
}
int fcb200_jacobian(int nx, int ny, const float* field1, const float* field2, const float* xmapr, const float* ymapr, float* fjacobian,
                    int* fDefined, float undef)
{
  return fcb200_jacobian_batched(nx, ny, 1, field1, field2, xmapr, ymapr, fjacobian, fDefined, undef);
}

int fcb200_ilevelgwind_batched(int nx, int ny, int nfields, const float* mpot, const float* xmapr, const float* ymapr, const float* fcoriolis,
                               float* ug, float* vg, int* fDefined, float undef)
{ // FC.cc:1511-1549; flag denominator is N (:1543)
  if (nx < 3 || ny < 3)
    return 0;
  if (!grid_ok(nx, ny, nfields))
    return -1;
  const size_t n = (size_t)nx * ny;
  Call call;
  GwindOp op;
  op.m = call.in(mpot, n * nfields);
  op.xm = call.in(xmapr, n);
  op.ym = call.in(ymapr, n);
  op.fc = call.in(fcoriolis, n);
  op.ug = call.out(ug, n * nfields);
  op.vg = call.out(vg, n * nfields);
  return run_stencil(call, op, nx, ny, nfields, fDefined, undef, n);
}
int fcb200_ilevelgwind(int nx, int ny, const float* mpot, const float* xmapr, const float* ymapr, const float* fcoriolis, float* ug, float* vg,
                       int* fDefined, float undef)
{
  return fcb200_ilevelgwind_batched(nx, ny, 1, mpot, xmapr, ymapr, fcoriolis, ug, vg, fDefined, undef);
}

int fcb200_plevelgwind_xcomp_batched(int nx, int ny, int nfields, const float* z, const float* xmapr, const float* ymapr, const float* fcoriolis,
                                     float* ug, int* fDefined, float undef)
{
  return geostrophic<0>(nx, ny, nfields, z, xmapr, ymapr, fcoriolis, ug, fDefined, undef);
}
int fcb200_plevelgwind_xcomp(int nx, int ny, const float* z, const float* xmapr, const float* ymapr, const float* fcoriolis, float* ug, int* fDefined,
                             float undef)
{
  return geostrophic<0>(nx, ny, 1, z, xmapr, ymapr, fcoriolis, ug, fDefined, undef);
}
int fcb200_plevelgwind_ycomp_batched(int nx, int ny, int nfields, const float* z, const float* xmapr, const float* ymapr, const float* fcoriolis,
                                     float* vg, int* fDefined, float undef)
{
  return geostrophic<1>(nx, ny, nfields, z, xmapr, ymapr, fcoriolis, vg, fDefined, undef);
}
int fcb200_plevelgwind_ycomp(int nx, int ny, const float* z, const float* xmapr, const float* ymapr, const float* fcoriolis, float* vg, int* fDefined,
                             float undef)
{
  return geostrophic<1>(nx, ny, 1, z, xmapr, ymapr, fcoriolis, vg, fDefined, undef);
}
int fcb200_plevelgvort_batched(int nx, int ny, int nfields, const float* z, const float* xmapr, const float* ymapr, const float* fcoriolis,
                               float* gvort, int* fDefined, float undef)
{
  return geostrophic<2>(nx, ny, nfields, z, xmapr, ymapr, fcoriolis, gvort, fDefined, undef);
}
int fcb200_plevelgvort(int nx, int ny, const float* z, const float* xmapr, const float* ymapr, const float* fcoriolis, float* gvort, int* fDefined,
                       float undef)
{
  return geostrophic<2>(nx, ny, 1, z, xmapr, ymapr, fcoriolis, gvort, fDefined, undef);
}

int fcb200_plevelqvector_batched(int nx, int ny, int nfields, const float* z, const float* t, const float* xmapr, const float* ymapr,
                                 const float* fcoriolis, float p, int compute, float* qcomp, int* fDefined, float undef)
{ // FC.cc:505-595
  if (p <= 0.0)
    return 0;
  if (nx < 3 || ny < 3)
    return 0;
  if (compute < 1 || compute > 4)
    return 0;
  if (!grid_ok(nx, ny, nfields))
    return -1;
  float tscale = 1.0f;
  if (compute == 2 || compute == 4) {
    const float pi = 1004.f * powf(p / 1000.f, 287.f / 1004.f); // :537 spells the Exner function out with p / p0
    tscale = pi / 1004.f;
  }
  const float c = (float)((double)(-287.f) / ((double)p * 100.));
  if (compute < 3)
    return qvector<0>(nx, ny, nfields, z, t, xmapr, ymapr, fcoriolis, tscale, c, qcomp, fDefined, undef);
  return qvector<1>(nx, ny, nfields, z, t, xmapr, ymapr, fcoriolis, tscale, c, qcomp, fDefined, undef);
}
int fcb200_plevelqvector(int nx, int ny, const float* z, const float* t, const float* xmapr, const float* ymapr, const float* fcoriolis, float p,
                         int compute, float* qcomp, int* fDefined, float undef)
{
  return fcb200_plevelqvector_batched(nx, ny, 1, z, t, xmapr, ymapr, fcoriolis, p, compute, qcomp, fDefined, undef);
}

int fcb200_thermalFrontParameter_batched(int nx, int ny, int nfields, const float* t, const float* xmapr, const float* ymapr, float* tfp,
                                         int* fDefined, float undef)
{ // FC.cc:2266-2309
  if (nx < 3 || ny < 3)
    return 0;
  if (!grid_ok(nx, ny, nfields))
    return -1;
  const size_t n = (size_t)nx * ny;
  Call call;
  const float* d_t = call.in(t, n * nfields);
  const float* d_xm = call.in(xmapr, n);
  const float* d_ym = call.in(ymapr, n);
  float* d_out = call.out(tfp, n * nfields);
  const FieldMeta* meta = flags_to_meta(call, fDefined, nfields);
  unsigned long long* counters = call.counters(2 * nfields); // [0, nfields): undefined elements of T, [nfields, 2 nfields): pass 2
  if (!call.ok())
    return -1;
  // Default: ONE pass over T (tfp_tile.cuh: 8 B/point, |grad T| in registers, float-float quotients).  FCB200_TFP=unfused selects the
  // reference's own two passes through a scratch field (round 1's default: gradient(c=3) on the tile engine, then TfpPass2Op,
  // 20 B/point); FCB200_TFP=fused_v1 round 1's fused kernel (|grad T| through shared memory, four block barriers per field).
  static const char* mode_env = getenv("FCB200_TFP");
  static const int mode = (mode_env && !strcmp(mode_env, "unfused")) ? 1 : (mode_env && !strcmp(mode_env, "fused_v1")) ? 2 : 0;
  if (mode == 1) {
    float* d_ad = static_cast<float*>(call.scratch(sizeof(float) * n * nfields));
    if (!call.ok())
      return -1;
    if (!launch_stencil(call, GradientOp<3>{d_t, d_xm, d_ym, d_ad}, nx, ny, nfields, undef, meta, counters))
      return -1;
    TfpPass2Op p2;
    p2.tx = d_t, p2.ad = d_ad, p2.xm = d_xm, p2.ym = d_ym, p2.pass1_undef = counters, p2.o = d_out;
    if (!launch_stencil(call, p2, nx, ny, nfields, undef, meta, counters + nfields))
      return -1;
    return call.finish_counted(fDefined, nfields, n - 2 * (size_t)nx, nfields);
  }
  bool any_masked = false;
  for (int k = 0; k < nfields; ++k)
    any_masked = any_masked || fDefined[k] != ALL_DEFINED;
  if (any_masked) {
    if (nfields > 65535) {
      set_error("fcb200: batch too large for one launch (%d fields)", nfields);
      return -1;
    }
    count_undefined_kernel<<<dim3((unsigned)std::min<size_t>((n + 4095) / 4096, 1024), (unsigned)nfields), 256, 0, call.stream()>>>(d_t, (int)n, nfields, undef, meta,
                                                                                                                            counters);
    count_launch();
  }
  TfpFusedOp op;
  op.tx = d_t;
  op.xm = d_xm;
  op.ym = d_ym;
  op.t_undef = counters;
  op.meta = meta;
  op.gnx = nx;
  op.gny = ny;
  op.undef_ = undef;
  op.all1 = false;
  op.o = d_out;
  const bool tiled = nx >= 34 && n >= 8192 && !std::isnan(undef); // the tile engine's threshold (smaller grids: the flat scalar kernel); the march tests definedness with one ordered comparison
  if (mode == 0 && tiled) {
    if (!launch_tfp_tile(call, op, nx, ny, nfields, undef, meta, counters + nfields))
      return -1;
  } else if (!launch_stencil(call, op, nx, ny, nfields, undef, meta, counters + nfields))
    return -1;
  return call.finish_counted(fDefined, nfields, n - 2 * (size_t)nx, nfields);
}
int fcb200_thermalFrontParameter(int nx, int ny, const float* t, const float* xmapr, const float* ymapr, float* tfp, int* fDefined, float undef)
{
  return fcb200_thermalFrontParameter_batched(nx, ny, 1, t, xmapr, ymapr, tfp, fDefined, undef);
}

int fcb200_shapiro2_filter_batched(int nx, int ny, int nfields, float* field, float* fsmooth, int* fDefined, float undef)
{ // FC.cc:2076-2179
  if (nx < 3 || ny < 3)
    return 0;
  if (!grid_ok(nx, ny, nfields))
    return -1;
  const size_t n = (size_t)nx * ny;
  Call call;
  const float* d_in = call.in(field, n * nfields);
  float* d_out = call.out(fsmooth, n * nfields);
  if (!call.ok())
    return -1;
  // the tile kernel is out of place; `field == fsmooth` is allowed by the reference (FC.cc:2088)
  float* d_tmp = d_out;
  const bool aliased = (static_cast<const float*>(d_out) == d_in);
  if (aliased)
    d_tmp = static_cast<float*>(call.scratch(sizeof(float) * n * nfields));
  const FieldMeta* meta = flags_to_meta(call, fDefined, nfields);
  if (!call.ok())
    return -1;
  // W = 4 needs every row of every field on a 16-byte boundary
  const bool vec = (nx % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_in) | reinterpret_cast<uintptr_t>(d_tmp)) & 15) == 0;
  // anything else: four adjacent columns per lane with 4-byte accesses (the one-column form, W = 1, is kept for comparison:
  // FCB200_SHAPIRO_W1; MEPS x 96, 30 % masked: 0.29 of the roofline with W = 1)
  static const bool w1 = getenv("FCB200_SHAPIRO_W1") != nullptr; // (development switch)
  const int useful = (vec || !w1) ? ShapiroGeom<4>::USEFUL : ShapiroGeom<1>::USEFUL;
  const int strips = (nx + useful - 1) / useful;
  const int strip_groups = (strips + SH_WARPS - 1) / SH_WARPS;
  // rows per band: as tall as possible (a band re-reads 4 halo rows) while the launch still fills the GPU
  int rows_per_band = 64;
  while (rows_per_band > 8 && (long long)strips * ((ny + rows_per_band - 1) / rows_per_band) * nfields < 24LL * sm_count())
    rows_per_band /= 2;
  const int bands = (ny + rows_per_band - 1) / rows_per_band;
  const long long grid = (long long)strip_groups * bands * nfields;
  if (grid > 0x7fffffffLL) {
    set_error("fcb200: batch too large for one launch (%lld CTAs)", grid);
    return -1;
  }
  bool only_all = true; // every field flagged all-defined: the kernels that hold the all-defined march alone
  for (int k = 0; k < nfields; ++k)
    only_all = only_all && fDefined[k] == ALL_DEFINED;
#define FCB_SHAPIRO_LAUNCH(K) K<<<(unsigned)grid, SH_WARPS * 32, 0, call.stream()>>>(d_in, d_tmp, nx, ny, strips, bands, rows_per_band, meta, undef)
  if (vec && only_all)
    FCB_SHAPIRO_LAUNCH(shapiro2_kernel_vec_all);
  else if (vec)
    FCB_SHAPIRO_LAUNCH((shapiro2_kernel<4, false>));
  else if (!w1 && only_all)
    FCB_SHAPIRO_LAUNCH(shapiro2_kernel_anycol_all);
  else if (!w1)
    FCB_SHAPIRO_LAUNCH(shapiro2_kernel_anycol);
  else
    FCB_SHAPIRO_LAUNCH((shapiro2_kernel<1, false>));
#undef FCB_SHAPIRO_LAUNCH
  count_launch();
  if (aliased) {
    if (!cuda_ok(cudaMemcpyAsync(d_out, d_tmp, sizeof(float) * n * nfields, cudaMemcpyDeviceToDevice, call.stream()), "cudaMemcpyAsync(D2D)"))
      return -1;
  }
  return call.finish([=](const unsigned long long*) {
    for (int k = 0; k < nfields; ++k)
      fDefined[k] = ALL_DEFINED; // FC.cc:2176
  });
}
int fcb200_shapiro2_filter(int nx, int ny, float* field, float* fsmooth, int* fDefined, float undef)
{
  return fcb200_shapiro2_filter_batched(nx, ny, 1, field, fsmooth, fDefined, undef);
}

} // extern "C"

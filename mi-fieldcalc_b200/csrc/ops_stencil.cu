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

#include "../../include/fcb200.h"

#include <cmath>

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
  float undef;
  const FieldMeta* meta;
  unsigned long long* counters;
};

__device__ __forceinline__ void store_with_edges(float* out, int i, int x, int y, int nx, int ny, float v)
{
  out[i] = v;
  const bool l = (x == 1), r = (x == nx - 2), t = (y == 1), b = (y == ny - 2);
  if (l)
    out[i - 1] = v;
  if (r)
    out[i + 1] = v;
  if (t) {
    out[i - nx] = v;
    if (l)
      out[i - nx - 1] = v;
    if (r)
      out[i - nx + 1] = v;
  }
  if (b) {
    out[i + nx] = v;
    if (l)
      out[i + nx - 1] = v;
    if (r)
      out[i + nx + 1] = v;
  }
}

// Op interface:
//   static constexpr int NOUT;
//   static constexpr bool TESTS_WHEN_ALL;  // eval() can fail even when allDefined (TFP's absdelt != 0)
//   bool count_flat;                      // count over [1, N-1) instead of rows 1..ny-2 (gradient c=1)
//   __device__ bool all_defined(int field, bool in_all) const;       // the allDefined of this pass
//   __device__ bool eval(int field, int i, int nx, int n, bool all, float undef, bool want, float* val) const;
//   __device__ float* out(int k, int field, int n) const;
template <class Op>
__global__ void __launch_bounds__(ST_THREADS) stencil_kernel(const Op op, const StencilGeom g)
{
  const int field = blockIdx.x / g.chunks;
  const int chunk = blockIdx.x - field * g.chunks;
  const bool all = op.all_defined(field, g.meta[field].all != 0);
  const int nx = g.nx, ny = g.ny, n = g.n;
  unsigned nundef = 0;

#pragma unroll
  for (int u = 0; u < ST_UNROLL; ++u) {
    const int i = chunk * (ST_THREADS * ST_UNROLL) + u * ST_THREADS + threadIdx.x;
    if (i >= n)
      continue;
    const int y = i / nx, x = i - y * nx;
    const bool row_in = (y >= 1 && y <= ny - 2);
    const bool interior = row_in && x >= 1 && x <= nx - 2;
    const bool count_here = op.count_flat ? (i >= 1 && i <= n - 2) : row_in;
    if (!(interior || (count_here && (!all || Op::TESTS_WHEN_ALL))))
      continue;
    float val[Op::NOUT];
    const bool ok = op.eval(field, i, nx, n, all, g.undef, interior, val);
    if (!ok)
      nundef += 1;
    if (interior) {
#pragma unroll
      for (int k = 0; k < Op::NOUT; ++k)
        store_with_edges(op.out(k, field, n), i, x, y, nx, ny, ok ? val[k] : g.undef);
    }
  }
  dev::block_add_counter(nundef, g.counters + field);
}

__device__ __forceinline__ bool def4(bool all, float a, float b, float c, float d, float undef)
{
  return all || (is_def(a, undef) && is_def(b, undef) && is_def(c, undef) && is_def(d, undef));
}

// centred difference times map ratio, evaluated like `0.5 * mapr[i] * (f[i+d] - f[i-d])`:
// float difference, double products
__device__ __forceinline__ double half_map_diff(float mapr, float hi, float lo)
{
  return 0.5 * (double)mapr * (double)(hi - lo);
}

// relvort / absvort / divergence (FC.cc:1843-1940)
struct VortDivOp
{
  static constexpr int NOUT = 1;
  static constexpr bool TESTS_WHEN_ALL = false;
  bool count_flat;
  int mode; // 0 relvort, 1 absvort, 2 divergence
  const float *u, *v, *xm, *ym, *fc;
  float* o;
  __device__ __forceinline__ bool all_defined(int, bool in_all) const { return in_all; }
  __device__ __forceinline__ float* out(int, int field, int n) const { return o + (long long)field * n; }
  __device__ __forceinline__ bool eval(int field, int i, int nx, int n, bool all, float undef, bool want, float* val) const
  {
    const float* uu = u + (long long)field * n;
    const float* vv = v + (long long)field * n;
    const float vl = vv[i - 1], vr = vv[i + 1], ud = uu[i - nx], uup = uu[i + nx];
    // the test of all three operators reads v[i+-1] and u[i+-nx] -- divergence included (FC.cc:1927)
    if (!def4(all, vl, vr, ud, uup, undef))
      return false;
    if (want) {
      if (mode == 2)
        val[0] = (float)(half_map_diff(xm[i], uu[i + 1], uu[i - 1]) + half_map_diff(ym[i], vv[i + nx], vv[i - nx]));
      else if (mode == 1)
        val[0] = (float)(half_map_diff(xm[i], vr, vl) - half_map_diff(ym[i], uup, ud) + (double)fc[i]);
      else
        val[0] = (float)(half_map_diff(xm[i], vr, vl) - half_map_diff(ym[i], uup, ud));
    }
    return true;
  }
};

// advection (FC.cc:1942-1983)
struct AdvectionOp
{
  static constexpr int NOUT = 1;
  static constexpr bool TESTS_WHEN_ALL = false;
  bool count_flat;
  const float *f, *u, *v, *xm, *ym;
  float scale;
  float* o;
  __device__ __forceinline__ bool all_defined(int, bool in_all) const { return in_all; }
  __device__ __forceinline__ float* out(int, int field, int n) const { return o + (long long)field * n; }
  __device__ __forceinline__ bool eval(int field, int i, int nx, int n, bool all, float undef, bool want, float* val) const
  {
    const long long off = (long long)field * n;
    const float* ff = f + off;
    const float ui = u[off + i], vi = v[off + i];
    const float fd = ff[i - nx], fl = ff[i - 1], fr = ff[i + 1], fu = ff[i + nx];
    if (!(all || (is_def(ui, undef) && is_def(vi, undef) && is_def(fd, undef) && is_def(fl, undef) && is_def(fr, undef) && is_def(fu, undef))))
      return false;
    if (want) {
      const double ax = (double)ui * 0.5 * (double)xm[i] * (double)(fr - fl);
      const double ay = (double)vi * 0.5 * (double)ym[i] * (double)(fu - fd);
      val[0] = (float)((ax + ay) * (double)scale);
    }
    return true;
  }
};

// gradient (FC.cc:1985-2074)
struct GradientOp
{
  static constexpr int NOUT = 1;
  static constexpr bool TESTS_WHEN_ALL = false;
  bool count_flat; // compute == 1 loops over [1, N-1)
  int compute;
  const float *f, *xm, *ym;
  float* o;
  __device__ __forceinline__ bool all_defined(int, bool in_all) const { return in_all; }
  __device__ __forceinline__ float* out(int, int field, int n) const { return o + (long long)field * n; }
  __device__ __forceinline__ bool eval(int field, int i, int nx, int n, bool all, float undef, bool want, float* val) const
  {
    const float* ff = f + (long long)field * n;
    if (compute == 1) {
      const float fl = ff[i - 1], fr = ff[i + 1];
      if (!(all || (is_def(fl, undef) && is_def(fr, undef))))
        return false;
      if (want)
        val[0] = (float)half_map_diff(xm[i], fr, fl);
      return true;
    }
    const float fd = ff[i - nx], fu = ff[i + nx];
    if (compute == 2) {
      if (!(all || (is_def(fd, undef) && is_def(fu, undef))))
        return false;
      if (want)
        val[0] = (float)half_map_diff(ym[i], fu, fd);
      return true;
    }
    const float fl = ff[i - 1], fr = ff[i + 1];
    if (!def4(all, fd, fl, fr, fu, undef))
      return false;
    if (compute == 3) {
      if (want) {
        const float dfdx = (float)half_map_diff(xm[i], fr, fl);
        const float dfdy = (float)half_map_diff(ym[i], fu, fd);
        val[0] = dev::absval(dfdx, dfdy);
      }
      return true;
    }
    const float fc = ff[i];
    if (!(all || is_def(fc, undef)))
      return false;
    if (want) {
      const float d2fdx = (float)((double)fl - 2.0 * (double)fc + (double)fr);
      const float d2fdy = (float)((double)fd - 2.0 * (double)fc + (double)fu);
      const double mx = (double)xm[i], my = (double)ym[i];
      val[0] = (float)(4.0 * (0.25 * mx * mx * (double)d2fdx + 0.25 * my * my * (double)d2fdy));
    }
    return true;
  }
};

// jacobian (FC.cc:2424-2460): four derivatives rounded to float, a*b - c*d in float without FMA
struct JacobianOp
{
  static constexpr int NOUT = 1;
  static constexpr bool TESTS_WHEN_ALL = false;
  bool count_flat;
  const float *f1, *f2, *xm, *ym;
  float* o;
  __device__ __forceinline__ bool all_defined(int, bool in_all) const { return in_all; }
  __device__ __forceinline__ float* out(int, int field, int n) const { return o + (long long)field * n; }
  __device__ __forceinline__ bool eval(int field, int i, int nx, int n, bool all, float undef, bool want, float* val) const
  {
    const float* a = f1 + (long long)field * n;
    const float* b = f2 + (long long)field * n;
    const float ad = a[i - nx], al = a[i - 1], ar = a[i + 1], au = a[i + nx];
    const float bd = b[i - nx], bl = b[i - 1], br = b[i + 1], bu = b[i + nx];
    if (!(def4(all, ad, al, ar, au, undef) && def4(all, bd, bl, br, bu, undef)))
      return false;
    if (want) {
      const float xmi = xm[i], ymi = ym[i];
      const float df1dx = (float)half_map_diff(xmi, ar, al);
      const float df1dy = (float)half_map_diff(ymi, au, ad);
      const float df2dx = (float)half_map_diff(xmi, br, bl);
      const float df2dy = (float)half_map_diff(ymi, bu, bd);
      val[0] = df1dx * df2dy - df1dy * df2dx;
    }
    return true;
  }
};

// ilevelgwind (FC.cc:1511-1549): two outputs
struct GwindOp
{
  static constexpr int NOUT = 2;
  static constexpr bool TESTS_WHEN_ALL = false;
  bool count_flat;
  const float *m, *xm, *ym, *fc;
  float *ug, *vg;
  __device__ __forceinline__ bool all_defined(int, bool in_all) const { return in_all; }
  __device__ __forceinline__ float* out(int k, int field, int n) const { return (k == 0 ? ug : vg) + (long long)field * n; }
  __device__ __forceinline__ bool eval(int field, int i, int nx, int n, bool all, float undef, bool want, float* val) const
  {
    const float* mm = m + (long long)field * n;
    const float md = mm[i - nx], ml = mm[i - 1], mr = mm[i + 1], mu = mm[i + nx];
    if (!def4(all, md, ml, mr, mu, undef))
      return false;
    if (want) {
      const double f = (double)fc[i];
      val[0] = (float)(-0.5 * (double)ym[i] * (double)(mu - md) / f);
      val[1] = (float)(0.5 * (double)xm[i] * (double)(mr - ml) / f);
    }
    return true;
  }
};

// second pass of thermalFrontParameter (FC.cc:2286-2303).  `ad` = abs(grad T) from gradient(c=3) INCLUDING
// its fillEdges; the pass's allDefined is the first pass's OUTPUT flag, i.e. "the first pass counted
// nothing" (FC.cc:2281-2286) -- read from the first pass's device counter, no host round trip.
struct TfpOp
{
  static constexpr int NOUT = 1;
  static constexpr bool TESTS_WHEN_ALL = true;
  bool count_flat;
  const float *tx, *ad, *xm, *ym;
  const unsigned long long* pass1_counters;
  float* o;
  __device__ __forceinline__ bool all_defined(int field, bool) const { return pass1_counters[field] == 0; }
  __device__ __forceinline__ float* out(int, int field, int n) const { return o + (long long)field * n; }
  __device__ __forceinline__ bool eval(int field, int i, int nx, int n, bool all, float undef, bool want, float* val) const
  {
    const float* t = tx + (long long)field * n;
    const float* a = ad + (long long)field * n;
    const float td = t[i - nx], tl = t[i - 1], tr = t[i + 1], tu = t[i + nx];
    const float adn = a[i - nx], al = a[i - 1], ac = a[i], ar = a[i + 1], au = a[i + nx];
    if (!(def4(all, td, tl, tr, tu, undef) && def4(all, adn, al, ar, au, undef) && (all || is_def(ac, undef))))
      return false;
    if (!(ac != 0)) // tested even when allDefined (FC.cc:2292)
      return false;
    if (want) {
      const float xmi = xm[i], ymi = ym[i];
      const float dadx = (float)half_map_diff(xmi, ar, al);
      const float dady = (float)half_map_diff(ymi, au, adn);
      const float dtdxa = (float)(half_map_diff(xmi, tr, tl) / (double)ac);
      const float dtdya = (float)(half_map_diff(ymi, tu, td) / (double)ac);
      val[0] = -(dadx * dtdxa + dady * dtdya);
    }
    return true;
  }
};

template <class Op>
bool launch_stencil(Call& call, const Op& op, int nx, int ny, int nfields, float undef, const FieldMeta* meta, unsigned long long* counters)
{
  StencilGeom g;
  g.nx = nx;
  g.ny = ny;
  g.n = nx * ny;
  g.nfields = nfields;
  g.chunks = (g.n + ST_THREADS * ST_UNROLL - 1) / (ST_THREADS * ST_UNROLL);
  g.undef = undef;
  g.meta = meta;
  g.counters = counters;
  const long long grid = (long long)g.chunks * nfields;
  if (grid > 0x7fffffffLL) {
    set_error("fcb200: batch too large for one launch (%lld CTAs)", grid);
    return false;
  }
  stencil_kernel<Op><<<(unsigned)grid, ST_THREADS, 0, call.stream()>>>(op, g);
  count_launch();
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

Finalizer flags_from_counters(int* fDefined, int nfields, unsigned long long denom, int counter_offset = 0)
{
  return [=](const unsigned long long* cnt) {
    for (int k = 0; k < nfields; ++k)
      fDefined[k] = check_defined(cnt[counter_offset + k], denom);
  };
}

// ------------------------------------------------------------------------------------ shapiro2_filter
// FC.cc:2076-2179.  Two iterations of (x pass, y pass); each pass stores float.  Mathematically a
// separable 2-D filter with these boundary rules (the flat-loop wrap never survives, FC.cc:2117-2120):
//   x pass: out(x,y) = filtered for 1 <= x <= nx-2 (EVERY row), copy for x = 0, nx-1
//   y pass: out(x,y) = filtered for 1 <= y <= ny-2 (EVERY column), copy for y = 0, ny-1
// All four passes run on one shared-memory tile with a halo of 2: each pass invalidates one more ring
// of the halo, and after four passes exactly the tile interior is final.  8 B/point of HBM traffic
// instead of the reference's nine array sweeps.
//   all-defined branch: s = +0.25 then -0.25; f + s*(fl + fr - 2.*f): float sum, the rest in double.
//   masked branch: weights 0.25/0 from 3-point definedness of the ORIGINAL field, BOTH iterations use
//   +0.25 (the `s = -0.25` update is dead, FC.cc:2136-2168), all-float arithmetic.
constexpr int SH_TX = 64, SH_TY = 32, SH_H = 2;
constexpr int SH_EX = SH_TX + 2 * SH_H, SH_EY = SH_TY + 2 * SH_H;
constexpr int SH_THREADS = 256;

__global__ void __launch_bounds__(SH_THREADS) shapiro2_kernel(const float* __restrict__ fin, float* __restrict__ fout, int nx, int ny, int tiles_x,
                                                              int tiles_y, const FieldMeta* meta, float undef)
{
  __shared__ float bufA[SH_EY][SH_EX + 1];
  __shared__ float bufB[SH_EY][SH_EX + 1];
  __shared__ unsigned char wmask[SH_EY][SH_EX + 1]; // bit 0: x weight is 0.25, bit 1: y weight is 0.25

  const int tiles = tiles_x * tiles_y;
  const int field = blockIdx.x / tiles;
  const int t = blockIdx.x - field * tiles;
  const int ty = t / tiles_x, tx = t - ty * tiles_x;
  const int x0 = tx * SH_TX - SH_H, y0 = ty * SH_TY - SH_H; // global coordinates of cell (0,0)
  const bool all = meta[field].all != 0;
  const float* src = fin + (long long)field * nx * ny;
  float* dst = fout + (long long)field * nx * ny;

  // load: out-of-grid cells get 0 and are never consumed by an in-grid filter point
  for (int c = threadIdx.x; c < SH_EX * SH_EY; c += SH_THREADS) {
    const int ly = c / SH_EX, lx = c - ly * SH_EX;
    const int gx = x0 + lx, gy = y0 + ly;
    float v = 0.f;
    if (gx >= 0 && gx < nx && gy >= 0 && gy < ny)
      v = src[(long long)gy * nx + gx];
    bufA[ly][lx] = v;
  }
  __syncthreads();

  if (!all) {
    for (int c = threadIdx.x; c < SH_EX * SH_EY; c += SH_THREADS) {
      const int ly = c / SH_EX, lx = c - ly * SH_EX;
      unsigned char m = 0;
      const bool dc = is_def(bufA[ly][lx], undef);
      if (lx >= 1 && lx < SH_EX - 1 && dc && is_def(bufA[ly][lx - 1], undef) && is_def(bufA[ly][lx + 1], undef))
        m |= 1;
      if (ly >= 1 && ly < SH_EY - 1 && dc && is_def(bufA[ly - 1][lx], undef) && is_def(bufA[ly + 1][lx], undef))
        m |= 2;
      wmask[ly][lx] = m;
    }
    __syncthreads();
  }

  float(*cur)[SH_EX + 1] = bufA;
  float(*nxt)[SH_EX + 1] = bufB;
#pragma unroll 1
  for (int pass = 0; pass < 4; ++pass) {
    const bool xpass = (pass & 1) == 0;
    const float s = (pass < 2) ? 0.25f : -0.25f;
    for (int c = threadIdx.x; c < SH_EX * SH_EY; c += SH_THREADS) {
      const int ly = c / SH_EX, lx = c - ly * SH_EX;
      const int gx = x0 + lx, gy = y0 + ly;
      const float f = cur[ly][lx];
      float r = f;
      const bool in_grid = gx >= 0 && gx < nx && gy >= 0 && gy < ny;
      bool filt;
      float lo = 0.f, hi = 0.f;
      if (xpass) {
        filt = in_grid && gx >= 1 && gx <= nx - 2 && lx >= 1 && lx < SH_EX - 1;
        if (filt) {
          lo = cur[ly][lx - 1];
          hi = cur[ly][lx + 1];
        }
      } else {
        filt = in_grid && gy >= 1 && gy <= ny - 2 && ly >= 1 && ly < SH_EY - 1;
        if (filt) {
          lo = cur[ly - 1][lx];
          hi = cur[ly + 1][lx];
        }
      }
      if (filt) {
        if (all) {
          r = (float)((double)f + (double)s * ((double)(lo + hi) - 2. * (double)f));
        } else {
          const float w = (wmask[ly][lx] & (xpass ? 1 : 2)) ? 0.25f : 0.f;
          r = f + w * (lo + hi - 2.f * f);
        }
      }
      nxt[ly][lx] = r;
    }
    __syncthreads();
    float(*tmp)[SH_EX + 1] = cur;
    cur = nxt;
    nxt = tmp;
  }

  for (int c = threadIdx.x; c < SH_TX * SH_TY; c += SH_THREADS) {
    const int ly = c / SH_TX, lx = c - ly * SH_TX;
    const int gx = tx * SH_TX + lx, gy = ty * SH_TY + ly;
    if (gx < nx && gy < ny)
      dst[(long long)gy * nx + gx] = cur[ly + SH_H][lx + SH_H];
  }
}

} // namespace
} // namespace fcb200

// =========================================================================================== C-ABI
using namespace fcb200;

namespace {

template <class Op>
int run_stencil(Call& call, const Op& op, int nx, int ny, int nfields, int* fDefined, float undef, unsigned long long denom)
{
  const FieldMeta* meta = flags_to_meta(call, fDefined, nfields);
  unsigned long long* counters = call.counters(nfields);
  if (!call.ok())
    return -1;
  if (!launch_stencil(call, op, nx, ny, nfields, undef, meta, counters))
    return -1;
  return call.finish(flags_from_counters(fDefined, nfields, denom));
}

int vortdiv(int mode, int nx, int ny, int nfields, const float* u, const float* v, const float* xmapr, const float* ymapr, const float* fcoriolis,
            float* out, int* fDefined, float undef)
{
  if (nx < 3 || ny < 3)
    return 0;
  if (!grid_ok(nx, ny, nfields))
    return -1;
  const size_t n = (size_t)nx * ny;
  Call call;
  VortDivOp op;
  op.count_flat = false;
  op.mode = mode;
  op.u = call.in(u, n * nfields);
  op.v = call.in(v, n * nfields);
  op.xm = call.in(xmapr, n);
  op.ym = call.in(ymapr, n);
  op.fc = (mode == 1) ? call.in(fcoriolis, n) : nullptr;
  op.o = call.out(out, n * nfields);
  return run_stencil(call, op, nx, ny, nfields, fDefined, undef, n - 2 * (size_t)nx);
}

} // namespace

extern "C" {

int fcb200_relvort_batched(int nx, int ny, int nfields, const float* u, const float* v, const float* xmapr, const float* ymapr, float* rvort,
                           int* fDefined, float undef)
{
  return vortdiv(0, nx, ny, nfields, u, v, xmapr, ymapr, nullptr, rvort, fDefined, undef);
}
int fcb200_relvort(int nx, int ny, const float* u, const float* v, const float* xmapr, const float* ymapr, float* rvort, int* fDefined, float undef)
{
  return vortdiv(0, nx, ny, 1, u, v, xmapr, ymapr, nullptr, rvort, fDefined, undef);
}

int fcb200_absvort_batched(int nx, int ny, int nfields, const float* u, const float* v, const float* xmapr, const float* ymapr,
                           const float* fcoriolis, float* avort, int* fDefined, float undef)
{
  return vortdiv(1, nx, ny, nfields, u, v, xmapr, ymapr, fcoriolis, avort, fDefined, undef);
}
int fcb200_absvort(int nx, int ny, const float* u, const float* v, const float* xmapr, const float* ymapr, const float* fcoriolis, float* avort,
                   int* fDefined, float undef)
{
  return vortdiv(1, nx, ny, 1, u, v, xmapr, ymapr, fcoriolis, avort, fDefined, undef);
}

int fcb200_divergence_batched(int nx, int ny, int nfields, const float* u, const float* v, const float* xmapr, const float* ymapr, float* diverg,
                              int* fDefined, float undef)
{
  return vortdiv(2, nx, ny, nfields, u, v, xmapr, ymapr, nullptr, diverg, fDefined, undef);
}
int fcb200_divergence(int nx, int ny, const float* u, const float* v, const float* xmapr, const float* ymapr, float* diverg, int* fDefined,
                      float undef)
{
  return vortdiv(2, nx, ny, 1, u, v, xmapr, ymapr, nullptr, diverg, fDefined, undef);
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
  op.count_flat = false;
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
  GradientOp op;
  op.count_flat = (compute == 1);
  op.compute = compute;
  op.f = call.in(field, n * nfields);
  // the reference only dereferences the map ratio(s) the mode needs; keep that so callers may pass the other as null
  op.xm = (compute != 2) ? call.in(xmapr, n) : nullptr;
  op.ym = (compute != 1) ? call.in(ymapr, n) : nullptr;
  op.o = call.out(fgrad, n * nfields);
  return run_stencil(call, op, nx, ny, nfields, fDefined, undef, n - 2 * (size_t)nx);
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
  op.count_flat = false;
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
  op.count_flat = false;
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

int fcb200_thermalFrontParameter_batched(int nx, int ny, int nfields, const float* t, const float* xmapr, const float* ymapr, float* tfp,
                                         int* fDefined, float undef)
{ // FC.cc:2266-2309: gradient(c=3) into scratch (with its fillEdges and flag), then the second five-point pass
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
  float* d_ad = static_cast<float*>(call.scratch(sizeof(float) * n * nfields));
  const FieldMeta* meta = flags_to_meta(call, fDefined, nfields);
  unsigned long long* counters = call.counters(2 * nfields); // [0, nfields): pass 1, [nfields, 2 nfields): pass 2
  if (!call.ok())
    return -1;
  GradientOp g;
  g.count_flat = false;
  g.compute = 3;
  g.f = d_t;
  g.xm = d_xm;
  g.ym = d_ym;
  g.o = d_ad;
  if (!launch_stencil(call, g, nx, ny, nfields, undef, meta, counters))
    return -1;
  TfpOp op;
  op.count_flat = false;
  op.tx = d_t;
  op.ad = d_ad;
  op.xm = d_xm;
  op.ym = d_ym;
  op.pass1_counters = counters;
  op.o = d_out;
  if (!launch_stencil(call, op, nx, ny, nfields, undef, meta, counters + nfields))
    return -1;
  return call.finish(flags_from_counters(fDefined, nfields, n - 2 * (size_t)nx, nfields));
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
  const int tiles_x = (nx + SH_TX - 1) / SH_TX, tiles_y = (ny + SH_TY - 1) / SH_TY;
  const long long grid = (long long)tiles_x * tiles_y * nfields;
  if (grid > 0x7fffffffLL) {
    set_error("fcb200: batch too large for one launch (%lld CTAs)", grid);
    return -1;
  }
  shapiro2_kernel<<<(unsigned)grid, SH_THREADS, 0, call.stream()>>>(d_in, d_tmp, nx, ny, tiles_x, tiles_y, meta, undef);
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

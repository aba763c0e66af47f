// ops_ensemble.cu -- ensemble reductions meanValue / stddevValue / extremeValue / probability
// (reference FC.cc:2696-2860).
//
// One thread owns W (4 or 1) grid points and walks the M member fields IN MEMBER ORDER with float
// accumulators, exactly like the reference's inner loops: a cross-lane tree reduction would change
// the rounding order (51 % bit-equal) while thread-per-point is bit-exact and
// just as coalesced -- M independent 16-byte streams per thread.  Warp/block reductions are only
// used for the undefined-point counter.
//
// Batched over `ntimes`: member j is a dense [ntimes][n] array, one CTA works on one chunk of one
// time, the output is [ntimes][n].
#include "device_common.cuh"

#include "../../include/fcb200.h"

#include <vector>

namespace fcb200 {
namespace {

using dev::is_def;

constexpr int EN_THREADS = 256;
#ifndef FCB_EN_AHEAD
#define FCB_EN_AHEAD 3
#endif
enum { EN_MEAN = 0, EN_STDDEV = 1, EN_EXTREME = 2, EN_PROB = 3 };
enum { MF_ALL = 1, MF_NOT_NONE = 2 }; // per (time, member) flag bits
// members in flight per thread in the load-ahead kernels (measured with 2, 3 and 4: 5 %-masked probability 0.72 / 0.76 / 0.84 of
// the roofline, extremeValue 0.82 / 0.80 / 0.73, meanValue 0.87 / 0.91 / 0.85)
template <int MODE>
constexpr int en_ahead()
{
  return MODE == EN_PROB ? FCB_EN_AHEAD + 1 : FCB_EN_AHEAD;
}

struct EnsArgs
{
  const float* const* members; // device table of M base pointers
  const int* member_flags;     // device [ntimes][M]
  const FieldMeta* meta;       // per time: all = input flag ALL_DEFINED (extremeValue), a = probability divisor lo, b = hi (double split)
  const double* prob_div;      // per time: nfields_defined / 100.0 (probability c1-3), or nullptr
  const float* recip;          // [M]: {j+1 as float, RN(1 / (j+1))} pairs for the all-defined Welford update
  float* out;
  unsigned long long* counters;
  long long n;
  int nmembers, ntimes, chunks, align0;
  int compute;
  int tables_global; // member tables are read from device memory instead of shared memory (very large ensembles)
  int check_above, check_below;
  float v_above, v_below;
  float undef;
};

// W points starting at `base` of one time step: walk the members in order, then store.
// FAST = every member field of this time step is flagged ALL_DEFINED (and, for extremeValue, the in/out flag
// too): no per-point definedness test at all, like the reference's allDefined short-circuit -- and the
// Welford divisor is the same j+1 for every point, so `delta / n` becomes a multiplication by the
// correctly rounded reciprocal RN(1/n) plus one residual correction (Markstein):
//   q0 = RN(delta*y), r = delta - n*q0 (exact, one fma), q = RN(q0 + r*y)
// q is the correctly rounded quotient for normal operands: |q0 + r*y - delta/n| <= 2^-47 |delta/n| while a
// quotient of a 24-bit float by an integer n <= 4096 that is not itself a float is at least 2^-36 away
// (relative) from every rounding boundary, and cannot be an exact tie (n*midpoint has >= 25 significant
// bits).  Zero, tiny, huge and non-finite deltas take the IEEE division.
// AHEAD > 0 (kernels launched for batches with undefined points, tables in shared memory): the member loads run AHEAD
// members in front of their use as volatile loads.  The per-point tests split the plain loop into basic blocks and the compiler
// then keeps ONE load in flight per warp (ncu: 8 warps per issue waiting on it).
template <int MODE, int W, bool FAST, int AHEAD = 0>
__device__ __forceinline__ void ensemble_points(const EnsArgs& a, const float* const* mptr, const int* mflag, const float2* recip, int time,
                                                long long base, float* out, bool in_all, unsigned& nundef, long long moff = 0)
{
  const float undef = a.undef;
  const int M = a.nmembers;
  float acc0[W], acc1[W]; // mean: sum,-   stddev: m, m2   extreme: cur, idx   probability: count,-
  int cnt[W];
#pragma unroll
  for (int w = 0; w < W; ++w) {
    acc0[w] = (MODE == EN_EXTREME) ? undef : 0.f;
    acc1[w] = (MODE == EN_EXTREME) ? undef : 0.f;
    cnt[w] = 0;
  }
  const bool want_max = (a.compute == 1 || a.compute == 3);
  float dmax = 1.f, dmin = 1.f; // range of |delta| seen by the reciprocal-based Welford update

  struct Pts
  {
    float v[W];
  };
  auto load = [&](int j) {
    Pts p;
    if constexpr (AHEAD > 0) {
      if constexpr (W == 4)
        asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(p.v[0]), "=f"(p.v[1]), "=f"(p.v[2]), "=f"(p.v[3]) : "l"(mptr[j] + moff + base));
      else
        asm volatile("ld.global.f32 %0, [%1];" : "=f"(p.v[0]) : "l"(mptr[j] + moff + base));
    } else if constexpr (W == 4) {
      const float4 q = *reinterpret_cast<const float4*>(mptr[j] + moff + base);
      p.v[0] = q.x;
      p.v[1] = q.y;
      p.v[2] = q.z;
      p.v[3] = q.w;
    } else {
      p.v[0] = mptr[j][moff + base];
    }
    return p;
  };
  auto member = [&](const int j, const Pts& x) {
    const int fl = FAST ? (MF_ALL | MF_NOT_NONE) : mflag[j];
    if (MODE == EN_PROB && !(fl & MF_NOT_NONE))
      return; // a member whose FIELD flag is NONE_DEFINED is not counted (FC.cc:2841)
    float2 ny = make_float2(0.f, 0.f);
    if (MODE == EN_STDDEV && FAST)
      ny = recip[j];
#pragma unroll
    for (int w = 0; w < W; ++w) {
      const float xv = x.v[w];
      if (MODE == EN_MEAN) { // FC.cc:2709-2714
        if ((fl & MF_ALL) || is_def(xv, undef)) {
          cnt[w] += 1;
          acc0[w] += xv;
        }
      } else if (MODE == EN_STDDEV && FAST) {
        const float delta = xv - acc0[w];
        const float q0 = delta * ny.y;
        const float rem = __fmaf_rn(-q0, ny.x, delta);
        acc0[w] += __fmaf_rn(rem, ny.y, q0);
        acc1[w] += delta * (xv - acc0[w]);
        dmax = fmaxf(dmax, fabsf(delta)); // (a NaN delta makes the result NaN on either path)
        dmin = fminf(dmin, fabsf(delta));
      } else if (MODE == EN_STDDEV) { // FC.cc:2739-2747, Welford in float, no FMA
        if ((fl & MF_ALL) || is_def(xv, undef)) {
          const float delta = xv - acc0[w];
          cnt[w] += 1;
          acc0[w] += delta / (float)cnt[w];
          acc1[w] += delta * (xv - acc0[w]);
        }
      } else if (MODE == EN_EXTREME) { // FC.cc:2778-2783, 2792-2798
        if (acc0[w] == undef || ((FAST || in_all || is_def(xv, undef)) && (want_max ? (acc0[w] < xv) : (acc0[w] > xv)))) {
          acc0[w] = xv;
          acc1[w] = (float)j;
        }
      } else { // FC.cc:2843-2846
        if ((xv != undef) && (!a.check_above || xv > a.v_above) && (!a.check_below || xv < a.v_below))
          acc0[w] += 1.f;
      }
    }
  };

  if constexpr (AHEAD > 0) {
    Pts buf[AHEAD > 0 ? AHEAD : 1];
#pragma unroll
    for (int d = 0; d < AHEAD; ++d)
      buf[d] = load(min(d, M - 1));
    int j = 0;
#pragma unroll 2
    for (; j + AHEAD <= M; j += AHEAD) {
#pragma unroll
      for (int d = 0; d < AHEAD; ++d) {
        const Pts x = buf[d];
        buf[d] = load(min(j + d + AHEAD, M - 1));
        member(j + d, x);
      }
    }
#pragma unroll
    for (int d = 0; d < AHEAD - 1; ++d)
      if (j + d < M)
        member(j + d, buf[d]);
  } else {
#pragma unroll 6
    for (int j = 0; j < M; ++j)
      member(j, load(j));
  }

  if (MODE == EN_STDDEV && FAST && !(dmin >= 1e-30f && dmax <= 1e30f)) {
    // a zero, tiny, huge or infinite delta: the correction above is only proven for normal operands -- redo
    // these points with the IEEE division
    ensemble_points<MODE, W, false>(a, mptr, mflag, recip, time, base, out, in_all, nundef, moff);
    return;
  }
  float r[W];
#pragma unroll
  for (int w = 0; w < W; ++w) {
    if ((MODE == EN_MEAN || MODE == EN_STDDEV) && FAST)
      cnt[w] = M;
    if (MODE == EN_MEAN) {
      if (cnt[w] > 0)
        r[w] = acc0[w] / (float)cnt[w];
      else {
        r[w] = undef;
        nundef += 1;
      }
    } else if (MODE == EN_STDDEV) {
      if (cnt[w] > 0)
        r[w] = sqrtf(acc1[w] / (float)cnt[w]); // == float(sqrt(double(m2/n))): double rounding is innocuous for sqrt
      else {
        r[w] = undef;
        nundef += 1;
      }
    } else if (MODE == EN_EXTREME) {
      r[w] = (a.compute >= 3) ? acc1[w] : acc0[w];
      if (r[w] == undef)
        nundef += 1;
    } else {
      // nfields_defined is a property of the time step (member FIELD flags), not of the point
      if (a.meta[time].b == 0.f) {
        r[w] = undef;
        nundef += 1;
      } else if (a.compute < 4)
        r[w] = (float)((double)acc0[w] / a.prob_div[time]);
      else
        r[w] = acc0[w];
    }
  }
  if constexpr (W == 4)
    *reinterpret_cast<float4*>(out + base) = make_float4(r[0], r[1], r[2], r[3]);
  else
    out[base] = r[0];
}

// stddevValue on members with undefined points, WITHOUT a branch per point (BF kernel instantiation): the reference's
//   if (defined) { delta = x - m; n += 1; m += delta / n; m2 += delta * (x - m); }          (FC.cc:2739-2747)
// is computed for every point.  An undefined value is replaced by the running mean itself: delta = +0, the quotient is +0
// and both accumulators keep their bits (m is never -0: it starts at +0 and a sum of opposite numbers is +0), so nothing has
// to be selected afterwards.  The divisor is the point's own count of defined members so far: {n, RN(1/n)} comes from the
// shared-memory table indexed by that count (entry 0 = a dummy for "none yet") and the quotient is Markstein's corrected
// product exactly as in the all-defined form above (same proof: an integer divisor <= 4096, a normal delta; a ZERO delta
// gives q0 = rem = q = 0 exactly, in both signs the same sums as the IEEE quotient).
// Definedness is ONE comparison per point, `x != u` (unordered: FSETP.NEU) with u = undef, or u = NaN for a member whose
// field flag is ALL_DEFINED (the reference does not look at its values): true for every x in the second case, and in the
// first for every x but undef -- including a NaN x, which the reference skips.  Such a NaN makes m NaN for good, which is
// tested once at the end: the group is then redone by the branching form (as is a group whose NaN the reference does use).
// The host keeps a NaN `undef` away from this kernel.
// Four points of a thread are independent instruction streams and the loads run D members ahead -- with a branch per
// point (BSSY / BSYNC, an IEEE division with its slow-path call inside) neither happens: 25 instructions per point and
// member against 15, one load in flight per warp.  A non-zero delta below 1e-30 (found as the unsigned minimum of
// 2 * bits - 1, which sends both zeros to the top), or one above 1e30, also sends the group to the branching form.
template <int W>
__device__ __forceinline__ bool stddev_points_bf(const EnsArgs& a, const float* const* mptr, const int* mflag, const float2* recip0, long long base,
                                                 float* out, unsigned& nundef)
{
  constexpr int D = 3; // members in flight
  struct Pts
  {
    float v[W];
  };
  const float undef = a.undef;
  const int M = a.nmembers;
  float m[W], m2[W];
  // the point's count of defined members, kept as the shared-memory address of its {n, RN(1/n)} table entry
  const unsigned tab0 = (unsigned)__cvta_generic_to_shared(recip0);
  unsigned tab[W];
#pragma unroll
  for (int w = 0; w < W; ++w) {
    m[w] = 0.f;
    m2[w] = 0.f;
    tab[w] = tab0;
  }
  float dmax = 1.f;
  unsigned vmin = 0xffffffffu;

  auto load = [&](int j) {
    Pts p;
    // (volatile: the compiler otherwise sinks the loads back to their first use and one is in flight again)
    if constexpr (W == 4)
      asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(p.v[0]), "=f"(p.v[1]), "=f"(p.v[2]), "=f"(p.v[3]) : "l"(mptr[j] + base));
    else
      asm volatile("ld.global.f32 %0, [%1];" : "=f"(p.v[0]) : "l"(mptr[j] + base));
    return p;
  };
  auto step = [&](int j, const Pts& x) {
    const float u = (mflag[j] & MF_ALL) ? __int_as_float(0x7fc00000) : undef;
    float delta[W];
#pragma unroll
    for (int w = 0; w < W; ++w) {
      // def = x != u (unordered);  if (def) count += 1;  xv = def ? x : m;  {n, 1/n} = table[count]
      float xv;
      float2 ny;
      asm("{\n\t.reg .pred p;\n\t"
          "setp.neu.f32 p, %4, %5;\n\t"
          "@p add.u32 %0, %0, 8;\n\t"
          "selp.f32 %1, %4, %6, p;\n\t"
          "ld.shared.v2.f32 {%2, %3}, [%0];\n\t}"
          : "+r"(tab[w]), "=f"(xv), "=f"(ny.x), "=f"(ny.y)
          : "f"(x.v[w]), "f"(u), "f"(m[w]));
      delta[w] = xv - m[w];
      const float q0 = delta[w] * ny.y;
      const float rem = __fmaf_rn(-q0, ny.x, delta[w]);
      m[w] += __fmaf_rn(rem, ny.y, q0);
      m2[w] += delta[w] * (xv - m[w]);
    }
#pragma unroll
    for (int w = 0; w < W; w += 2) {
      if (w + 1 < W) {
        dmax = fmaxf(fmaxf(dmax, fabsf(delta[w])), fabsf(delta[w + 1]));
        vmin = min(min(vmin, 2u * __float_as_uint(delta[w]) - 1u), 2u * __float_as_uint(delta[w + 1]) - 1u);
      } else {
        dmax = fmaxf(dmax, fabsf(delta[w]));
        vmin = min(vmin, 2u * __float_as_uint(delta[w]) - 1u);
      }
    }
  };

  Pts buf[D];
#pragma unroll
  for (int d = 0; d < D; ++d)
    buf[d] = load(min(d, M - 1));
  int j = 0;
#pragma unroll 2
  for (; j + D <= M; j += D) {
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const Pts x = buf[d];
      buf[d] = load(min(j + d + D, M - 1));
      step(j + d, x);
    }
  }
#pragma unroll
  for (int d = 0; d < D - 1; ++d)
    if (j + d < M)
      step(j + d, buf[d]);

  bool ok = dmax <= 1e30f && vmin >= 2u * __float_as_uint(1e-30f) - 1u;
#pragma unroll
  for (int w = 0; w < W; ++w)
    ok = ok && m[w] == m[w];
  if (!ok)
    return false;
  float r[W];
#pragma unroll
  for (int w = 0; w < W; ++w) {
    const int cnt = (int)((tab[w] - tab0) >> 3);
    if (cnt > 0)
      r[w] = sqrtf(m2[w] / (float)cnt);
    else {
      r[w] = undef;
      nundef += 1;
    }
  }
  if constexpr (W == 4)
    *reinterpret_cast<float4*>(out + base) = make_float4(r[0], r[1], r[2], r[3]);
  else
    out[base] = r[0];
  return true;
}

// TG (tables_global) = the member tables are too large for shared memory (more than 2048 members): read from the device tables
// directly, the time offset added at each use.  A separate instantiation: the common kernel keeps its tables in shared memory
// with shared-memory loads and no offset arithmetic (as one kernel with a run-time switch, generic loads and a 64-bit add per
// member load cost stddevValue a fifth of its throughput: 0.82 -> 0.64 of the roofline).
// BF (stddevValue, tables in shared memory) = time steps with undefined points take stddev_points_bf.  Its own instantiation,
// launched only when some time step of the batch is not all-defined: the all-defined kernel keeps its registers.
// AH (the other three reductions, tables in shared memory) = time steps with undefined points load AH members ahead; launched like BF.
template <int MODE, int W, bool TG, bool BF = false, int AH = 0>
__global__ void __launch_bounds__(EN_THREADS) ensemble_kernel(const EnsArgs a)
{
  static_assert(AH == 0 || (!TG && !BF), "load-ahead form: shared-memory tables, not stddevValue");
  static_assert(!BF || (MODE == EN_STDDEV && !TG), "the branch-free masked form exists for stddevValue with shared-memory tables");
  extern __shared__ unsigned char smem_raw[];
  const int time = blockIdx.x / a.chunks;
  const int chunk = blockIdx.x - time * a.chunks;
  const int M = a.nmembers;
  const long long n = a.n;
  const float** s_ptr = reinterpret_cast<const float**>(smem_raw);
  float2* s_recip = reinterpret_cast<float2*>(smem_raw + sizeof(float*) * M);
  // BF: the {n, 1/n} table has a leading dummy entry (index = number of defined members so far, 0 = none yet)
  int* s_flag = reinterpret_cast<int*>(smem_raw + (sizeof(float*) + sizeof(float2)) * M + (BF ? sizeof(float2) : 0));
  if (!TG) {
    for (int j = threadIdx.x; j < M; j += EN_THREADS) {
      s_ptr[j] = a.members[j] + (long long)time * n;
      s_flag[j] = a.member_flags[(long long)time * M + j];
      s_recip[j + (BF ? 1 : 0)] = reinterpret_cast<const float2*>(a.recip)[j];
    }
    if (BF && threadIdx.x == 0)
      s_recip[0] = make_float2(1.f, 1.f);
    __syncthreads();
  }
  const float* const* mptr = TG ? a.members : s_ptr;
  const int* mflag = TG ? a.member_flags + (long long)time * M : s_flag;
  const float2* recip = TG ? reinterpret_cast<const float2*>(a.recip) : s_recip + (BF ? 1 : 0);
  const long long moff = TG ? (long long)time * n : 0;

  // per-time peel so that the float4 groups are 16-byte aligned (see elementwise.cuh)
  const int head = (W == 4) ? ((4 - ((a.align0 + (int)(((long long)time * n) & 3)) & 3)) & 3) : 0;
  const long long groups = (n - head) / W;
  float* out = a.out + (long long)time * n;
  const bool in_all = a.meta[time].all != 0;
  const bool fast = a.meta[time].c != 0.f; // every member ALL_DEFINED (and the in/out flag for extremeValue)
  unsigned nundef = 0;

  const long long g = (long long)chunk * EN_THREADS + threadIdx.x;
  if (g < groups) {
    if (fast)
      ensemble_points<MODE, W, true, AH>(a, mptr, mflag, recip, time, head + g * W, out, in_all, nundef, moff);
    else if (!BF || !stddev_points_bf<W>(a, mptr, mflag, s_recip, head + g * W, out, nundef))
      ensemble_points<MODE, W, false, AH>(a, mptr, mflag, recip, time, head + g * W, out, in_all, nundef, moff);
  }

  if (W == 4 && chunk == 0) {
    const long long tail0 = head + groups * 4;
    const int ntail = (int)(n - tail0);
    long long idx = -1;
    if ((int)threadIdx.x < head)
      idx = threadIdx.x;
    else if (threadIdx.x >= 32 && (int)threadIdx.x - 32 < ntail)
      idx = tail0 + (threadIdx.x - 32);
    if (idx >= 0)
      ensemble_points<MODE, 1, false>(a, mptr, mflag, recip, time, idx, out, in_all, nundef, moff);
  }

  dev::block_add_counter(nundef, a.counters + time);
}

struct EnsHost
{
  int mode, compute;
  int nx, ny, ntimes, nmembers;
  const float* const* fields;
  const int* fDefinedIn; // [ntimes][nmembers] or nullptr (extremeValue)
  const float* limits;
  int nlimits;
  float* fres;
  int* fDefinedOut; // [ntimes]; for extremeValue also the input flag
  float undef;
};

int run_ensemble(const EnsHost& h)
{
  if (h.nx <= 0 || h.ny <= 0 || h.ntimes <= 0 || (long long)h.nx * h.ny >= 0x7fffffffLL) {
    set_error("fcb200: invalid grid or batch size (nx=%d ny=%d ntimes=%d)", h.nx, h.ny, h.ntimes);
    return -1;
  }
  const int M = h.nmembers > 0 ? h.nmembers : 0;
  const long long n = (long long)h.nx * h.ny;
  Call call;

  EnsArgs a;
  a.compute = h.compute;
  a.check_above = a.check_below = 0;
  a.v_above = a.v_below = 0.f;
  a.prob_div = nullptr;
  if (h.mode == EN_PROB) { // FC.cc:2821-2825
    const bool between = (h.nlimits >= 2) && (h.compute == 3 || h.compute == 6);
    a.check_above = (h.nlimits >= 1) && (h.compute == 1 || h.compute == 4 || between);
    a.check_below = (h.nlimits >= 1) && (h.compute == 2 || h.compute == 5 || between);
    if (a.check_above || a.check_below) {
      a.v_above = h.limits[0];
      a.v_below = between ? h.limits[1] : h.limits[0];
    }
  }

  std::vector<const float*> dptr(M > 0 ? M : 1, nullptr);
  for (int j = 0; j < M; ++j)
    dptr[j] = call.in(h.fields[j], (size_t)(n * h.ntimes));
  float* d_out = call.out(h.fres, (size_t)(n * h.ntimes));
  if (!call.ok())
    return -1;

  // every small table of the call in ONE host-to-device copy: member pointers, {n, 1/n} pairs, member flags,
  // per-time metadata, probability divisors
  const size_t Mp = (size_t)(M > 0 ? M : 1);
  const size_t off_ptr = 0;
  const size_t off_rcp = off_ptr + sizeof(float*) * Mp;
  const size_t off_meta = (off_rcp + sizeof(float) * 2 * Mp + 15) & ~size_t(15);
  const size_t off_div = off_meta + sizeof(FieldMeta) * (size_t)h.ntimes;
  const size_t off_flag = off_div + sizeof(double) * (size_t)h.ntimes;
  const size_t blob_bytes = off_flag + sizeof(int) * (size_t)h.ntimes * Mp;
  std::vector<unsigned char> blob(blob_bytes, 0);
  const float** b_ptr = reinterpret_cast<const float**>(blob.data() + off_ptr);
  float* b_rcp = reinterpret_cast<float*>(blob.data() + off_rcp);
  FieldMeta* meta = reinterpret_cast<FieldMeta*>(blob.data() + off_meta);
  double* pdiv = reinterpret_cast<double*>(blob.data() + off_div);
  int* mflags = reinterpret_cast<int*>(blob.data() + off_flag);
  for (int j = 0; j < M; ++j) {
    b_ptr[j] = dptr[j];
    b_rcp[2 * j] = (float)(j + 1);
    b_rcp[2 * j + 1] = 1.0f / (float)(j + 1); // IEEE: the correctly rounded reciprocal
  }
  for (int t = 0; t < h.ntimes; ++t) {
    int counted = 0;
    bool all_members = h.fDefinedIn != nullptr || h.mode == EN_EXTREME;
    for (int j = 0; j < M; ++j) {
      int f = 0;
      if (h.fDefinedIn) {
        const int fd = h.fDefinedIn[(size_t)t * M + j];
        if (fd == ALL_DEFINED)
          f |= MF_ALL;
        else
          all_members = false;
        if (fd != NONE_DEFINED) {
          f |= MF_NOT_NONE;
          counted += 1;
        }
      }
      mflags[(size_t)t * Mp + j] = f;
    }
    meta[t].all = (h.mode == EN_EXTREME && h.fDefinedOut[t] == ALL_DEFINED) ? 1 : 0;
    if (h.mode == EN_EXTREME)
      all_members = meta[t].all != 0;
    meta[t].a = 0.f;
    meta[t].b = (float)counted;
    // (the reciprocal form of the Welford update is proven for divisors up to 4096 only)
    meta[t].c = (all_members && M > 0 && !(h.mode == EN_STDDEV && M > 4096)) ? 1.f : 0.f;
    pdiv[t] = counted / 100.0; // FC.cc:2855
  }
  const unsigned char* d_blob = static_cast<const unsigned char*>(call.upload_small(blob.data(), blob_bytes));
  a.counters = call.counters(h.ntimes);
  if (!call.ok() || !d_blob)
    return -1;
  a.members = reinterpret_cast<const float* const*>(d_blob + off_ptr);
  a.recip = reinterpret_cast<const float*>(d_blob + off_rcp);
  a.meta = reinterpret_cast<const FieldMeta*>(d_blob + off_meta);
  a.prob_div = reinterpret_cast<const double*>(d_blob + off_div);
  a.member_flags = reinterpret_cast<const int*>(d_blob + off_flag);

  a.out = d_out;
  a.n = n;
  a.nmembers = M;
  a.ntimes = h.ntimes;
  a.undef = h.undef;

  bool vec = n >= 16;
  const uintptr_t a0 = reinterpret_cast<uintptr_t>(d_out) & 15;
  if (a0 & 3)
    vec = false;
  for (int j = 0; j < M; ++j)
    if ((reinterpret_cast<uintptr_t>(dptr[j]) & 15) != a0)
      vec = false;
  a.align0 = (int)(a0 >> 2);
  const int width = vec ? 4 : 1;
  const long long per_cta = (long long)EN_THREADS * width;
  a.chunks = (int)((n + per_cta - 1) / per_cta);
  const long long grid = (long long)a.chunks * h.ntimes;
  if (grid > 0x7fffffffLL) {
    set_error("fcb200: batch too large for one launch (%lld CTAs)", grid);
    return -1;
  }
  // 20 bytes of shared memory per member: up to 2048 members fit the 48 KB every device grants without an opt-in
  a.tables_global = M > 2048 ? 1 : 0;
  // stddevValue with some time step that is not all-defined: the kernel with the branch-free masked form (divisors <= 4096:
  // implied by the shared-memory tables)
  bool masked_steps = false;
  for (int t = 0; t < h.ntimes; ++t)
    masked_steps = masked_steps || meta[t].c == 0.f;
  const bool bf = h.mode == EN_STDDEV && !a.tables_global && M > 0 && masked_steps && h.undef == h.undef;
  const size_t smem = a.tables_global ? 0 : (sizeof(float*) + sizeof(float2) + sizeof(int)) * Mp + (bf ? sizeof(float2) : 0);

  // the kernels whose member loads run ahead of their use: batches with undefined points (some time step not all-defined), and
  // two of the reductions always
  // (all-defined batches, measured: extremeValue 0.86 -> 0.91 and meanValue 0.93 -> 0.94 of the roofline with the loads ahead,
  // stddevValue unchanged, probability 0.90 -> 0.88)
  const bool ahead = !a.tables_global && M > 0 && ((h.mode != EN_STDDEV && masked_steps) || h.mode == EN_EXTREME || h.mode == EN_MEAN);
#define FCB_LAUNCH_ENS(MODE)                                                                                                                         \
  do {                                                                                                                                               \
    if (a.tables_global) {                                                                                                                           \
      if (vec)                                                                                                                                       \
        ensemble_kernel<MODE, 4, true><<<(unsigned)grid, EN_THREADS, smem, call.stream()>>>(a);                                                     \
      else                                                                                                                                           \
        ensemble_kernel<MODE, 1, true><<<(unsigned)grid, EN_THREADS, smem, call.stream()>>>(a);                                                     \
    } else if (ahead) {                                                                                                                              \
      if (vec)                                                                                                                                       \
        ensemble_kernel<MODE, 4, false, false, en_ahead<MODE>()><<<(unsigned)grid, EN_THREADS, smem, call.stream()>>>(a);         \
      else                                                                                                                                           \
        ensemble_kernel<MODE, 1, false, false, en_ahead<MODE>()><<<(unsigned)grid, EN_THREADS, smem, call.stream()>>>(a);         \
    } else if (vec)                                                                                                                                  \
      ensemble_kernel<MODE, 4, false><<<(unsigned)grid, EN_THREADS, smem, call.stream()>>>(a);                                                      \
    else                                                                                                                                             \
      ensemble_kernel<MODE, 1, false><<<(unsigned)grid, EN_THREADS, smem, call.stream()>>>(a);                                                      \
  } while (0)
  switch (h.mode) {
  case EN_MEAN:
    FCB_LAUNCH_ENS(EN_MEAN);
    break;
  case EN_STDDEV:
    if (bf && vec)
      ensemble_kernel<EN_STDDEV, 4, false, true><<<(unsigned)grid, EN_THREADS, smem, call.stream()>>>(a);
    else if (bf)
      ensemble_kernel<EN_STDDEV, 1, false, true><<<(unsigned)grid, EN_THREADS, smem, call.stream()>>>(a);
    else
      FCB_LAUNCH_ENS(EN_STDDEV);
    break;
  case EN_EXTREME:
    FCB_LAUNCH_ENS(EN_EXTREME);
    break;
  default:
    FCB_LAUNCH_ENS(EN_PROB);
    break;
  }
#undef FCB_LAUNCH_ENS
  count_launch();

  int* flags = h.fDefinedOut;
  const int ntimes = h.ntimes;
  return call.finish([=](const unsigned long long* cnt) {
    for (int t = 0; t < ntimes; ++t)
      flags[t] = check_defined(cnt[t], (unsigned long long)n);
  });
}

} // namespace
} // namespace fcb200

// =========================================================================================== C-ABI
using namespace fcb200;

extern "C" {

int fcb200_meanValue_batched(int nx, int ny, int ntimes, const float* const* fields, int nmembers, const int* fDefinedIn, float* fres,
                             int* fDefinedOut, float undef)
{ // FC.cc:2696-2724
  EnsHost h{EN_MEAN, 0, nx, ny, ntimes, nmembers, fields, fDefinedIn, nullptr, 0, fres, fDefinedOut, undef};
  return run_ensemble(h);
}
int fcb200_meanValue(int nx, int ny, const float* const* fields, int nfields, const int* fDefinedIn, float* fres, int* fDefinedOut, float undef)
{
  return fcb200_meanValue_batched(nx, ny, 1, fields, nfields, fDefinedIn, fres, fDefinedOut, undef);
}

int fcb200_stddevValue_batched(int nx, int ny, int ntimes, const float* const* fields, int nmembers, const int* fDefinedIn, float* fres,
                               int* fDefinedOut, float undef)
{ // FC.cc:2726-2757
  EnsHost h{EN_STDDEV, 0, nx, ny, ntimes, nmembers, fields, fDefinedIn, nullptr, 0, fres, fDefinedOut, undef};
  return run_ensemble(h);
}
int fcb200_stddevValue(int nx, int ny, const float* const* fields, int nfields, const int* fDefinedIn, float* fres, int* fDefinedOut, float undef)
{
  return fcb200_stddevValue_batched(nx, ny, 1, fields, nfields, fDefinedIn, fres, fDefinedOut, undef);
}

int fcb200_extremeValue_batched(int compute, int nx, int ny, int ntimes, const float* const* fields, int nmembers, float* fres, int* fDefined,
                                float undef)
{ // FC.cc:2759-2805
  if (nmembers <= 0)
    return 0;
  if (compute < 1 || compute > 4) {
    // the reference computes nothing, leaves fres untouched, sets ALL_DEFINED and returns true (:2774-2804)
    for (int t = 0; t < ntimes; ++t)
      fDefined[t] = ALL_DEFINED;
    return 1;
  }
  EnsHost h{EN_EXTREME, compute, nx, ny, ntimes, nmembers, fields, nullptr, nullptr, 0, fres, fDefined, undef};
  return run_ensemble(h);
}
int fcb200_extremeValue(int compute, int nx, int ny, const float* const* fields, int nfields, float* fres, int* fDefined, float undef)
{
  return fcb200_extremeValue_batched(compute, nx, ny, 1, fields, nfields, fres, fDefined, undef);
}

int fcb200_probability_batched(int compute, int nx, int ny, int ntimes, const float* const* fields, int nmembers, const int* fDefinedIn,
                               const float* limits, int nlimits, float* fres, int* fDefinedOut, float undef)
{ // FC.cc:2807-2860
  const bool between = (nlimits >= 2) && (compute == 3 || compute == 6);
  const bool above = (nlimits >= 1) && (compute == 1 || compute == 4 || between);
  const bool below = (nlimits >= 1) && (compute == 2 || compute == 5 || between);
  EnsHost h{EN_PROB, compute, nx, ny, ntimes, nmembers, fields, fDefinedIn, limits, nlimits, fres, fDefinedOut, undef};
  if (!(above || below)) {
    // fill undef, NONE_DEFINED, return false (:2827-2833): run the kernel with no member counted
    EnsHost f = h;
    f.nmembers = 0;
    f.compute = 4;
    const int r = run_ensemble(f);
    if (r < 0)
      return r;
    return 0;
  }
  return run_ensemble(h);
}
int fcb200_probability(int compute, int nx, int ny, const float* const* fields, int nfields, const int* fDefinedIn, const float* limits, int nlimits,
                       float* fres, int* fDefinedOut, float undef)
{
  return fcb200_probability_batched(compute, nx, ny, 1, fields, nfields, fDefinedIn, limits, nlimits, fres, fDefinedOut, undef);
}

} // extern "C"

// elementwise.cuh -- the batched per-point engine behind every operator without neighbours
// (p/h/a-level conversions, windCooling, fieldOPERfield, momentum
// coordinates, vessel icing).  It replaces the reference's generic loops unaryFunctionField /
// binaryFunctionFieldField and their ...Undef variants (FC.cc:94-179) and the hand-written loops
// of the same shape.
//
// Layout: a batch is `nfields` dense fields of `n = nx*ny` floats, field k at base + k*stride
// (stride = n, or 0 for an array shared by the whole batch).  One CTA works on one chunk of one
// field, so the per-field metadata is CTA-uniform and the undefined counter needs one atomic per
// CTA.  Every thread first issues ALL its loads (UNROLL x NIN vectors of 16 B), then computes, then
// stores: the kernels are HBM-bound, bytes in flight are what matters.
//
// Vector path (W = 4): fields of a batch are only 4-byte aligned in general (n = 949*1069 is odd),
// so every field is peeled: `head` scalar points up to the first 16-byte boundary, float4 groups,
// then < 4 tail points.  The host only selects W = 4 when all arrays share the same misalignment.
#pragma once

#include "device_common.cuh"

namespace fcb200 {

constexpr int EW_THREADS = 256;

template <int NIN, int NOUT>
struct EwArgs
{
  const float* in[NIN];
  long long in_stride[NIN];
  float* out[NOUT];
  long long n;  // points per field
  int nfields;
  int chunks;   // CTAs per field
  int align0;   // (address of field 0, element 0) / 4 mod 4 -- identical for every array (W = 4 only)
  int nx;
  float undef;
  const FieldMeta* meta;
  unsigned long long* counters; // one per field, or nullptr when the operator never counts
};

struct PointCtx
{
  const dev::EwtTable& tab;
  FieldMeta m;
  float undef;
  int nx;
};

// An operator type provides:
//   static constexpr int NIN, NOUT, UNROLL;  static constexpr bool USES_EWT, COUNTS;
//   __device__ void point(const float* in, float* out, const PointCtx& c, long long idx, unsigned& nundef) const;
// `in`/`out` hold the NIN inputs / NOUT outputs of ONE grid point, `idx` is the point's flat index
// inside its field.

template <class Op, int W>
__global__ void __launch_bounds__(EW_THREADS) ew_kernel(const Op op, const EwArgs<Op::NIN, Op::NOUT> a)
{
  constexpr int NIN = Op::NIN, NOUT = Op::NOUT;
  constexpr int U = (W == 4) ? Op::UNROLL : Op::UNROLL * 2;

  __shared__ dev::EwtTable tab;
  if (Op::USES_EWT) {
    tab.load();
    __syncthreads();
  }

  const int field = blockIdx.x / a.chunks;
  const int chunk = blockIdx.x - field * a.chunks;
  const PointCtx c{tab, a.meta[field], a.undef, a.nx};

  const float* in[NIN];
  float* out[NOUT];
#pragma unroll
  for (int k = 0; k < NIN; ++k)
    in[k] = a.in[k] + (long long)field * a.in_stride[k];
#pragma unroll
  for (int k = 0; k < NOUT; ++k)
    out[k] = a.out[k] + (long long)field * a.n;

  const long long n = a.n;
  const int head = (W == 4) ? ((4 - ((a.align0 + (int)(((long long)field * n) & 3)) & 3)) & 3) : 0;
  const long long groups = (n - head) / W;
  unsigned nundef = 0;

  // ---- body: all loads first, then compute + store
  float v[U][NIN][W];
  const long long g0 = (long long)chunk * (EW_THREADS * U) + threadIdx.x;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const long long g = g0 + (long long)u * EW_THREADS;
    if (g < groups) {
#pragma unroll
      for (int k = 0; k < NIN; ++k) {
        if constexpr (W == 4) {
          const float4 q = *reinterpret_cast<const float4*>(in[k] + head + g * 4);
          v[u][k][0] = q.x;
          v[u][k][1] = q.y;
          v[u][k][2] = q.z;
          v[u][k][3] = q.w;
        } else {
          v[u][k][0] = in[k][g];
        }
      }
    }
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const long long g = g0 + (long long)u * EW_THREADS;
    if (g < groups) {
      float r[NOUT][W];
#pragma unroll
      for (int w = 0; w < W; ++w) {
        float pin[NIN], pout[NOUT];
#pragma unroll
        for (int k = 0; k < NIN; ++k)
          pin[k] = v[u][k][w];
        op.point(pin, pout, c, head + g * W + w, nundef);
#pragma unroll
        for (int k = 0; k < NOUT; ++k)
          r[k][w] = pout[k];
      }
#pragma unroll
      for (int k = 0; k < NOUT; ++k) {
        if constexpr (W == 4)
          *reinterpret_cast<float4*>(out[k] + head + g * 4) = make_float4(r[k][0], r[k][1], r[k][2], r[k][3]);
        else
          out[k][g] = r[k][0];
      }
    }
  }

  // ---- peel: < 4 head points and < 4 tail points of the field, done by its first CTA
  if (W == 4 && chunk == 0) {
    const long long tail0 = head + groups * 4;
    const int ntail = (int)(n - tail0);
    long long idx = -1;
    if ((int)threadIdx.x < head)
      idx = threadIdx.x;
    else if (threadIdx.x >= 32 && (int)threadIdx.x - 32 < ntail)
      idx = tail0 + (threadIdx.x - 32);
    if (idx >= 0) {
      float pin[NIN], pout[NOUT];
#pragma unroll
      for (int k = 0; k < NIN; ++k)
        pin[k] = in[k][idx];
      op.point(pin, pout, c, idx, nundef);
#pragma unroll
      for (int k = 0; k < NOUT; ++k)
        out[k][idx] = pout[k];
    }
  }

  if (Op::COUNTS)
    dev::block_add_counter(nundef, a.counters + field);
}

// Host side: pick the vector width, size the grid, launch.  `in`/`out` are DEVICE pointers.
template <class Op>
bool launch_elementwise(Call& call, const Op& op, const float* const* in, const long long* in_stride, float* const* out, long long n, int nfields,
                        int nx, float undef, const FieldMeta* meta, unsigned long long* counters)
{
  constexpr int NIN = Op::NIN, NOUT = Op::NOUT;
  EwArgs<NIN, NOUT> a;
  bool vec = n >= 16;
  const uintptr_t a0 = reinterpret_cast<uintptr_t>(in[0]) & 15;
  for (int k = 0; k < NIN; ++k) {
    a.in[k] = in[k];
    a.in_stride[k] = in_stride[k];
    if ((reinterpret_cast<uintptr_t>(in[k]) & 15) != a0 || (a0 & 3))
      vec = false;
    if (in_stride[k] != n && !(nfields == 1 || (n & 3) == 0))
      vec = false; // a shared array cannot follow the per-field misalignment of the strided ones
  }
  for (int k = 0; k < NOUT; ++k) {
    a.out[k] = out[k];
    if ((reinterpret_cast<uintptr_t>(out[k]) & 15) != a0)
      vec = false;
  }
  a.n = n;
  a.nfields = nfields;
  a.align0 = (int)(a0 >> 2);
  a.nx = nx;
  a.undef = undef;
  a.meta = meta;
  a.counters = counters;
  const int width = vec ? 4 : 1;
  const int unroll = vec ? Op::UNROLL : Op::UNROLL * 2;
  const long long per_cta = (long long)EW_THREADS * unroll * width;
  a.chunks = (int)((n + per_cta - 1) / per_cta);
  const long long grid = (long long)a.chunks * nfields;
  if (grid <= 0 || grid > 0x7fffffffLL) {
    set_error("fcb200: batch too large for one launch (%lld CTAs)", grid);
    return false;
  }
  if (vec)
    ew_kernel<Op, 4><<<(unsigned)grid, EW_THREADS, 0, call.stream()>>>(op, a);
  else
    ew_kernel<Op, 1><<<(unsigned)grid, EW_THREADS, 0, call.stream()>>>(op, a);
  count_launch();
  return true;
}

} // namespace fcb200

// elementwise.cuh -- the batched per-point engine behind every operator without neighbours
// (p/h/a-level conversions, windCooling, fieldOPERfield, momentum
// coordinates, vessel icing).  It replaces the reference's generic loops unaryFunctionField /
// binaryFunctionFieldField and their ...Undef variants (FC.cc:94-179) and the hand-written loops
// of the same shape.
//
// Layout: a batch is `nfields` dense fields of `n = nx*ny` floats, field k at base + k*stride
// (stride = n, or 0 for an array shared by the whole batch).  The work is cut into items of
// EW_THREADS * U * W consecutive points of one field; the grid is PERSISTENT (a multiple of the SM
// count) and the items are dealt round-robin to the CTAs, so that
//   * the saturation / pow tables are staged in shared memory once per CTA, not once per item;
//   * the per-field metadata is CTA-uniform per item and the ALL_DEFINED fast path (no undefined
//     tests at all, exactly like the reference's `allDefined` short-circuit, FC.h:47-98) is chosen by
//     a uniform branch into code specialised at compile time;
//   * undefined points are counted in registers and flushed with one warp reduction + one atomic
//     per warp only when the CTA moves on to another field -- no block barriers in the loop.
// Every thread first issues ALL its loads (U x NIN vectors of 16 B), then computes, then stores:
// the kernels are HBM- or issue-bound, never latency-bound.
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
  long long n;      // points per field
  long long items;  // nfields * chunks
  int nfields;
  int chunks;       // items per field
  int align0;       // (address of field 0, element 0) / 4 mod 4 -- identical for every array (W = 4 only)
  int nx;
  float undef;
  const FieldMeta* meta;
  unsigned long long* counters; // Op::NCOUNT per field, or nullptr when the operator never counts
};

struct PointCtx
{
  const dev::EwtTable& tab;
  const dev::PowTable& pw;
  FieldMeta m;
  float undef;
  int nx;
};

// An operator type provides:
//   static constexpr int NIN, NOUT, UNROLL, NCOUNT, MIN_BLOCKS;  static constexpr bool USES_EWT, USES_POW;
//   static constexpr bool HEAVY;   // hundreds of instructions per point: one CTA per item, hardware load balancing
//   (MIN_BLOCKS = CTAs per SM the register allocation must allow)
//   template <bool ALL>
//   __device__ void point(const float* in, float* out, const PointCtx& c, long long idx, unsigned* nundef) const;
// `in`/`out` hold the NIN inputs / NOUT outputs of ONE grid point, `idx` is the point's flat index
// inside its field, `nundef` are the thread's NCOUNT private undefined-point counts (field k's
// counters live at counters[k*NCOUNT ...]).  ALL = the field's input flag is ALL_DEFINED: the
// operator must then skip every is_defined test (NaN / undef flow through the arithmetic).

template <class Op, int W, bool ALL>
__device__ __forceinline__ void ew_item(const Op& op, const EwArgs<Op::NIN, Op::NOUT>& a, const PointCtx& c, int field, int chunk, unsigned* nundef)
{
  constexpr int NIN = Op::NIN, NOUT = Op::NOUT;
  constexpr int U = (W == 4) ? Op::UNROLL : Op::UNROLL * 2;

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
        op.template point<ALL>(pin, pout, c, head + g * W + w, nundef);
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

  // ---- peel: < 4 head points and < 4 tail points of the field, done with its first item
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
      op.template point<ALL>(pin, pout, c, idx, nundef);
#pragma unroll
      for (int k = 0; k < NOUT; ++k)
        out[k][idx] = pout[k];
    }
  }
}

template <int NCOUNT>
__device__ __forceinline__ void ew_flush(unsigned* nundef, unsigned long long* counters, int field)
{
#pragma unroll
  for (int k = 0; k < NCOUNT; ++k) {
    const unsigned total = __reduce_add_sync(0xffffffffu, nundef[k]);
    if ((threadIdx.x & 31) == 0 && total)
      atomicAdd(counters + (long long)field * NCOUNT + k, (unsigned long long)total);
    nundef[k] = 0;
  }
}

// CTA-level flush for the one-item-per-CTA kernel: every CTA in flight works on the same one or two
// fields, and one atomic per WARP on a single address serialises in L2 when many points are undefined.
template <int NCOUNT>
__device__ __forceinline__ void ew_flush_block(unsigned* nundef, unsigned long long* counters, int field)
{
  __shared__ unsigned s_count[NCOUNT > 0 ? NCOUNT : 1];
  if (threadIdx.x < NCOUNT)
    s_count[threadIdx.x] = 0;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NCOUNT; ++k) {
    const unsigned total = __reduce_add_sync(0xffffffffu, nundef[k]);
    if ((threadIdx.x & 31) == 0 && total)
      atomicAdd(&s_count[k], total);
  }
  __syncthreads();
  if (threadIdx.x < NCOUNT && s_count[threadIdx.x])
    atomicAdd(counters + (long long)field * NCOUNT + threadIdx.x, (unsigned long long)s_count[threadIdx.x]);
}

template <class Op, int W>
__global__ void __launch_bounds__(EW_THREADS, Op::MIN_BLOCKS) ew_kernel(const Op op, const EwArgs<Op::NIN, Op::NOUT> a)
{
  __shared__ dev::EwtTable tab;
  __shared__ dev::PowTable pw;
  if (Op::USES_EWT)
    tab.load();
  if (Op::USES_POW)
    pw.load();
  if (Op::USES_EWT || Op::USES_POW)
    __syncthreads();

  // Items are dealt round-robin: at any moment the resident CTAs stream one compact window of the
  // batch (good DRAM page locality), exactly like a non-persistent grid would.  (field, chunk) are
  // advanced incrementally -- no division in the loop -- and the next item's metadata is fetched
  // before the current item is processed, so its latency is hidden.
  constexpr int NC = Op::NCOUNT > 0 ? Op::NCOUNT : 1;
  unsigned nundef[NC] = {};
  const unsigned items = (unsigned)a.items, grid = gridDim.x, chunks = (unsigned)a.chunks;
  const unsigned step_f = grid / chunks, step_c = grid - step_f * chunks;
  unsigned item = blockIdx.x;
  unsigned field = item / chunks, chunk = item - field * chunks;
  FieldMeta m = a.meta[field < (unsigned)a.nfields ? field : 0];
  while (item < items) {
    const unsigned nitem = item + grid;
    unsigned nfield = field + step_f, nchunk = chunk + step_c;
    if (nchunk >= chunks) {
      nchunk -= chunks;
      nfield += 1;
    }
    FieldMeta mn = m;
    if (nitem < items && nfield != field)
      mn = a.meta[nfield];
    const PointCtx c{tab, pw, m, a.undef, a.nx};
    if (m.all == 1)
      ew_item<Op, W, true>(op, a, c, (int)field, (int)chunk, nundef);
    else
      ew_item<Op, W, false>(op, a, c, (int)field, (int)chunk, nundef);
    if (Op::NCOUNT > 0 && (nfield != field || nitem >= items))
      ew_flush<Op::NCOUNT>(nundef, a.counters, (int)field);
    item = nitem;
    field = nfield;
    chunk = nchunk;
    m = mn;
  }
}

// One CTA per item: used for operators without shared-memory tables (pure streaming: every register
// goes to loads in flight) and for compute-heavy ones (the hardware scheduler balances the load).
template <class Op, int W>
__global__ void __launch_bounds__(EW_THREADS, Op::MIN_BLOCKS) ew_kernel_once(const Op op, const EwArgs<Op::NIN, Op::NOUT> a)
{
  __shared__ dev::EwtTable tab;
  __shared__ dev::PowTable pw;
  if (Op::USES_EWT)
    tab.load();
  if (Op::USES_POW)
    pw.load();
  if (Op::USES_EWT || Op::USES_POW)
    __syncthreads();
  constexpr int NC = Op::NCOUNT > 0 ? Op::NCOUNT : 1;
  unsigned nundef[NC] = {};
  const unsigned chunks = (unsigned)a.chunks;
  const unsigned field = blockIdx.x / chunks, chunk = blockIdx.x - field * chunks;
  const PointCtx c{tab, pw, a.meta[field], a.undef, a.nx};
  if (c.m.all == 1)
    ew_item<Op, W, true>(op, a, c, (int)field, (int)chunk, nundef);
  else
    ew_item<Op, W, false>(op, a, c, (int)field, (int)chunk, nundef);
  if (Op::NCOUNT > 0)
    ew_flush_block<Op::NCOUNT>(nundef, a.counters, (int)field);
}

// Host side: pick the vector width, size the persistent grid, launch.  `in`/`out` are DEVICE pointers.
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
  const long long per_item = (long long)EW_THREADS * unroll * width;
  a.chunks = (int)((n + per_item - 1) / per_item);
  a.items = (long long)a.chunks * nfields;

  if (a.items > 0x7fffffffLL) {
    set_error("fcb200: batch too large for one launch (%lld work items)", a.items);
    return false;
  }
  // Operators that stage tables in shared memory run a persistent grid (as many CTAs as stay
  // resident: occupancy x SM count) so that the staging is paid once per CTA; pure streaming and
  // compute-heavy operators get one CTA per item and leave load balancing to the hardware scheduler.
  const bool persistent = (Op::USES_EWT || Op::USES_POW) && !Op::HEAVY;
  if (!persistent) {
    if (vec)
      ew_kernel_once<Op, 4><<<(unsigned)a.items, EW_THREADS, 0, call.stream()>>>(op, a);
    else
      ew_kernel_once<Op, 1><<<(unsigned)a.items, EW_THREADS, 0, call.stream()>>>(op, a);
    count_launch();
    return true;
  }
  int occ = 0;
  cudaError_t e = vec ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ew_kernel<Op, 4>, EW_THREADS, 0)
                      : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ew_kernel<Op, 1>, EW_THREADS, 0);
  if (e != cudaSuccess || occ < 1) {
    cudaGetLastError();
    occ = 2;
  }
  long long grid = (long long)occ * sm_count();
  if (grid > a.items)
    grid = a.items;
  if (grid < 1)
    grid = 1;
  if (vec)
    ew_kernel<Op, 4><<<(unsigned)grid, EW_THREADS, 0, call.stream()>>>(op, a);
  else
    ew_kernel<Op, 1><<<(unsigned)grid, EW_THREADS, 0, call.stream()>>>(op, a);
  count_launch();
  return true;
}

} // namespace fcb200

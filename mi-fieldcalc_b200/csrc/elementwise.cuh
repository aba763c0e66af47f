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

#include <cstdlib>
#include <type_traits>

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
  int align0;       // (address of field 0, element 0) / 4 mod 4 -- identical for every per-field array (W = 4 only)
  int scalar_mask;  // W = 4: inputs (bit k) shared by the batch whose alignment cannot follow the per-field arrays -> four 4-byte loads
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

// Optional: `static constexpr bool QUAD = true` + quad<ALL>(const float (*in)[4], float (*out)[4], ctx, nundef)
// evaluates the four points of one float4 group together (in[k][w] = input k of point w), so that an operator
// can keep its common path free of branches across all four and handle rare cases afterwards.
template <class Op, class = void>
struct op_has_quad : std::false_type
{
};
template <class Op>
struct op_has_quad<Op, std::void_t<decltype(Op::QUAD)>> : std::integral_constant<bool, Op::QUAD>
{
};

// Optional: `static constexpr int ITEM_ROUNDS = J` makes an item of the PERSISTENT kernel J times larger (J rounds
// of U groups per thread): the per-item bookkeeping (next item, field metadata, counters, index arithmetic) is
// ~100 issue slots and only pays for itself when spread over enough points.
template <class Op, class = void>
struct op_item_rounds : std::integral_constant<int, 1>
{
};
template <class Op>
struct op_item_rounds<Op, std::void_t<decltype(Op::ITEM_ROUNDS)>> : std::integral_constant<int, Op::ITEM_ROUNDS>
{
};

// Optional: `static constexpr bool SHARED_INPUT = true` -- some input may be ONE field shared by the whole batch (hybrid-level
// ps).  Only such operators carry the per-input choice between a float4 load and four 4-byte loads (EwArgs::scalar_mask):
// predicated-off loads still cost issue slots, so everybody else compiles the plain float4 load.
template <class Op, class = void>
struct op_shared_input : std::false_type
{
};
template <class Op>
struct op_shared_input<Op, std::void_t<decltype(Op::SHARED_INPUT)>> : std::integral_constant<bool, Op::SHARED_INPUT>
{
};

template <class Op, int W>
struct EwShape
{
  static constexpr int U = (W == 4) ? Op::UNROLL : Op::UNROLL * 2;
  static constexpr int J = (U >= 2) ? op_item_rounds<Op>::value : 1; // rounds per item in the persistent kernel
};

// Index arithmetic of one item.  n < 2^31 (the reference's `int fsize`), so everything inside a field is
// 32-bit; only field * stride needs 64 bits (one IMAD.WIDE).
template <int W>
struct EwPos
{
  unsigned head;   // first float4-aligned point of the field (W = 4): fields of a batch are only 4-byte aligned in general
  unsigned groups; // number of full W-point groups after the head
  unsigned g0;     // this thread's first group

  template <int NIN, int NOUT>
  __device__ __forceinline__ EwPos(const EwArgs<NIN, NOUT>& a, unsigned field, unsigned chunk, int U)
  {
    const unsigned n = (unsigned)a.n;
    head = (W == 4) ? ((4u - (((unsigned)a.align0 + field * n) & 3u)) & 3u) : 0u;
    groups = (n - head) / W;
    g0 = chunk * (unsigned)(EW_THREADS * U) + threadIdx.x;
  }
  __device__ __forceinline__ unsigned group(int u) const { return g0 + (unsigned)u * EW_THREADS; }
  __device__ __forceinline__ unsigned first_point(int u) const { return head + group(u) * W; }
};

// ---- the loads of group u of one item
template <class Op, int W>
__device__ __forceinline__ void ew_load_group(const EwArgs<Op::NIN, Op::NOUT>& a, unsigned field, const EwPos<W>& pos, int u, float (&v)[Op::NIN][W])
{
  if (pos.group(u) < pos.groups) {
    const unsigned e = pos.first_point(u);
#pragma unroll
    for (int k = 0; k < Op::NIN; ++k) {
      const float* in = a.in[k] + ((unsigned long long)field * (unsigned)a.in_stride[k] + e);
      if constexpr (W == 4) {
        if (op_shared_input<Op>::value && ((a.scalar_mask >> k) & 1)) { // warp-uniform: a grid-constant array next to odd-sized fields
          v[k][0] = in[0];
          v[k][1] = in[1];
          v[k][2] = in[2];
          v[k][3] = in[3];
        } else {
          const float4 q = *reinterpret_cast<const float4*>(in);
          v[k][0] = q.x;
          v[k][1] = q.y;
          v[k][2] = q.z;
          v[k][3] = q.w;
        }
      } else {
        v[k][0] = in[0];
      }
    }
  }
}

template <class Op, int W>
__device__ __forceinline__ void ew_load(const EwArgs<Op::NIN, Op::NOUT>& a, unsigned field, unsigned chunk, float (&v)[EwShape<Op, W>::U][Op::NIN][W],
                                        int groups_per_item = EwShape<Op, W>::U)
{
  constexpr int U = EwShape<Op, W>::U;
  const EwPos<W> pos(a, field, chunk, groups_per_item);
#pragma unroll
  for (int u = 0; u < U; ++u)
    ew_load_group<Op, W>(a, field, pos, u, v[u]);
}

// ---- compute + store of group u of one item whose inputs are in registers
template <class Op, int W, bool ALL>
__device__ __forceinline__ void ew_compute_group(const Op& op, const EwArgs<Op::NIN, Op::NOUT>& a, const PointCtx& c, unsigned field, const EwPos<W>& pos, int u,
                                                 const float (&v)[Op::NIN][W], unsigned* nundef)
{
  constexpr int NIN = Op::NIN, NOUT = Op::NOUT;
  if (pos.group(u) < pos.groups) {
    const unsigned e = pos.first_point(u);
    float r[NOUT][W];
    if constexpr (W == 4 && op_has_quad<Op>::value) {
      op.template quad<ALL>(v, r, c, nundef);
    } else {
#pragma unroll
      for (int w = 0; w < W; ++w) {
        float pin[NIN], pout[NOUT];
#pragma unroll
        for (int k = 0; k < NIN; ++k)
          pin[k] = v[k][w];
        op.template point<ALL>(pin, pout, c, (long long)(e + w), nundef);
#pragma unroll
        for (int k = 0; k < NOUT; ++k)
          r[k][w] = pout[k];
      }
    }
    const unsigned long long off = (unsigned long long)field * (unsigned)a.n + e;
#pragma unroll
    for (int k = 0; k < NOUT; ++k) {
      if constexpr (W == 4)
        *reinterpret_cast<float4*>(a.out[k] + off) = make_float4(r[k][0], r[k][1], r[k][2], r[k][3]);
      else
        a.out[k][off] = r[k][0];
    }
  }
}

// ---- the < 4 head and < 4 tail points of a field, done (with their own scalar loads) together with its first item
template <class Op, int W, bool ALL>
__device__ __forceinline__ void ew_peel(const Op& op, const EwArgs<Op::NIN, Op::NOUT>& a, const PointCtx& c, unsigned field, unsigned chunk, const EwPos<W>& pos,
                                        unsigned* nundef)
{
  constexpr int NIN = Op::NIN, NOUT = Op::NOUT;
  if (W == 4 && chunk == 0) {
    const unsigned n = (unsigned)a.n;
    const unsigned tail0 = pos.head + pos.groups * 4;
    const unsigned ntail = n - tail0;
    long long idx = -1;
    if (threadIdx.x < pos.head)
      idx = threadIdx.x;
    else if (threadIdx.x >= 32 && threadIdx.x - 32 < ntail)
      idx = tail0 + (threadIdx.x - 32);
    if (idx >= 0) {
      float pin[NIN], pout[NOUT];
#pragma unroll
      for (int k = 0; k < NIN; ++k)
        pin[k] = (a.in[k] + (unsigned long long)field * (unsigned)a.in_stride[k])[idx];
      op.template point<ALL>(pin, pout, c, idx, nundef);
#pragma unroll
      for (int k = 0; k < NOUT; ++k)
        (a.out[k] + (unsigned long long)field * n)[idx] = pout[k];
    }
  }
}

template <class Op, int W, bool ALL>
__device__ __forceinline__ void ew_compute(const Op& op, const EwArgs<Op::NIN, Op::NOUT>& a, const PointCtx& c, unsigned field, unsigned chunk,
                                           const float (&v)[EwShape<Op, W>::U][Op::NIN][W], unsigned* nundef)
{
  constexpr int U = EwShape<Op, W>::U;
  const EwPos<W> pos(a, field, chunk, U);
#pragma unroll
  for (int u = 0; u < U; ++u)
    ew_compute_group<Op, W, ALL>(op, a, c, field, pos, u, v[u], nundef);
  ew_peel<Op, W, ALL>(op, a, c, field, chunk, pos, nundef);
}

// Persistent kernel, U >= 2: an item is J rounds of U groups per thread.  The group that will reuse a register
// buffer (same u, next round -- or round 0 of the NEXT item) is requested right after the buffer has been
// consumed: every load has the compute time of the other U - 1 groups plus the loop turn-around to arrive, and
// no register is copied.
template <class Op, int W, bool ALL>
__device__ __forceinline__ void ew_compute_and_refill(const Op& op, const EwArgs<Op::NIN, Op::NOUT>& a, const PointCtx& c, unsigned field, unsigned chunk,
                                                      bool has_next, unsigned nfield, unsigned nchunk, float (&v)[EwShape<Op, W>::U][Op::NIN][W],
                                                      unsigned* nundef)
{
  constexpr int U = EwShape<Op, W>::U, J = EwShape<Op, W>::J;
  const EwPos<W> pos(a, field, chunk, U * J);
  const EwPos<W> npos(a, nfield, nchunk, U * J);
#pragma unroll 1
  for (int j = 0; j < J; ++j) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      ew_compute_group<Op, W, ALL>(op, a, c, field, pos, j * U + u, v[u], nundef);
      if (j + 1 < J)
        ew_load_group<Op, W>(a, field, pos, (j + 1) * U + u, v[u]);
      else if (has_next)
        ew_load_group<Op, W>(a, nfield, npos, u, v[u]);
    }
  }
  ew_peel<Op, W, ALL>(op, a, c, field, chunk, pos, nundef);
}

template <class Op, int W, bool ALL>
__device__ __forceinline__ void ew_item(const Op& op, const EwArgs<Op::NIN, Op::NOUT>& a, const PointCtx& c, int field, int chunk, unsigned* nundef)
{
  float v[EwShape<Op, W>::U][Op::NIN][W];
  ew_load<Op, W>(a, (unsigned)field, (unsigned)chunk, v);
  ew_compute<Op, W, ALL>(op, a, c, (unsigned)field, (unsigned)chunk, v, nundef);
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
  // The inputs of item i+1 are requested before item i is computed (one register buffer ahead): the
  // kernels on this path spend ~150 issue slots per point, and without the prefetch a warp sits on its
  // loads while too few other warps are resident to fill the issue slots (ncu: long_scoreboard).
  constexpr int U = EwShape<Op, W>::U;
  float v[U][Op::NIN][W];
  float vn[U >= 2 ? 1 : U][Op::NIN][W]; // only U == 1 needs a second buffer (and a register copy per item)
  if (item < items)
    ew_load<Op, W>(a, field, chunk, v, U * EwShape<Op, W>::J); // round 0 of the first item
  while (item < items) {
    const unsigned nitem = item + grid;
    unsigned nfield = field + step_f, nchunk = chunk + step_c;
    if (nchunk >= chunks) {
      nchunk -= chunks;
      nfield += 1;
    }
    const bool has_next = nitem < items;
    FieldMeta mn = m;
    if (has_next && nfield != field)
      mn = a.meta[nfield];
    const PointCtx c{tab, pw, m, a.undef, a.nx};
    if constexpr (U >= 2) {
      if (m.all == 1)
        ew_compute_and_refill<Op, W, true>(op, a, c, field, chunk, has_next, nfield, nchunk, v, nundef);
      else
        ew_compute_and_refill<Op, W, false>(op, a, c, field, chunk, has_next, nfield, nchunk, v, nundef);
    } else {
      if (has_next)
        ew_load<Op, W>(a, nfield, nchunk, vn);
      if (m.all == 1)
        ew_compute<Op, W, true>(op, a, c, field, chunk, v, nundef);
      else
        ew_compute<Op, W, false>(op, a, c, field, chunk, v, nundef);
#pragma unroll
      for (int k = 0; k < Op::NIN; ++k)
#pragma unroll
        for (int w = 0; w < W; ++w)
          v[0][k][w] = vn[0][k][w];
    }
    if (Op::NCOUNT > 0 && (nfield != field || !has_next))
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
  // the vector path groups points by the alignment of the PER-FIELD arrays (taken from the first of them); an array shared
  // by the batch cannot follow the per-field misalignment of odd-sized fields and is then read with 4-byte loads
  uintptr_t a0 = reinterpret_cast<uintptr_t>(in[0]) & 15;
  for (int k = NIN - 1; k >= 0; --k)
    if (in_stride[k] == n)
      a0 = reinterpret_cast<uintptr_t>(in[k]) & 15;
  a.scalar_mask = 0;
  for (int k = 0; k < NIN; ++k) {
    a.in[k] = in[k];
    a.in_stride[k] = in_stride[k];
    const uintptr_t ak = reinterpret_cast<uintptr_t>(in[k]) & 15;
    if ((ak & 3) || (a0 & 3))
      vec = false;
    if (in_stride[k] != n) {
      if (ak != a0 || !(nfields == 1 || (n & 3) == 0)) {
        a.scalar_mask |= 1 << k;
        if (!op_shared_input<Op>::value)
          vec = false;
      }
    } else if (ak != a0) {
      vec = false;
    }
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
  // Operators that stage tables in shared memory run a persistent grid (as many CTAs as stay
  // resident: occupancy x SM count) so that the staging is paid once per CTA; pure streaming and
  // compute-heavy operators get one CTA per item and leave load balancing to the hardware scheduler.
  constexpr bool persistent = (Op::USES_EWT || Op::USES_POW) && !Op::HEAVY;
  const int width = vec ? 4 : 1;
  const int unroll = vec ? Op::UNROLL : Op::UNROLL * 2;
  const int rounds = !persistent ? 1 : (vec ? EwShape<Op, 4>::J : EwShape<Op, 1>::J);
  const long long per_item = (long long)EW_THREADS * unroll * width * rounds;
  a.chunks = (int)((n + per_item - 1) / per_item);
  a.items = (long long)a.chunks * nfields;

  if (a.items > 0x7fffffffLL) {
    set_error("fcb200: batch too large for one launch (%lld work items)", a.items);
    return false;
  }
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

// stencil_tile.cuh -- the tiled engine behind the five-point map-ratio stencils.
//
// Why tiles: on B200 an L2 hit costs about as much L2 sector throughput as a DRAM fill (measured with
// tools/probes/stencil_probe.cu: a kernel that streams 12 B/point from HBM runs at 6.8 TB/s, the same
// kernel with 8 B/point of extra L2-resident reads at 5.5 TB/s, relvort's access pattern with 20 B/point
// of L2 reads at 3.9 TB/s).  A five-point stencil that reads every neighbour through the cache hierarchy
// is therefore L2-bound at ~60 % of the HBM roofline whatever its instruction mix.  Here every input
// element crosses L2 once:
//   * a CTA owns a tile of TX x TY output points and walks over a block of fields;
//   * the rows of each per-field input (plus the halo rows above/below and one halo column left/right)
//     are brought to shared memory by the TMA unit: one bulk copy (cp.async.bulk, completion on an
//     mbarrier) per row segment, issued by the lanes of warp 0 -- no registers, no LSU instructions,
//     and up to NSTAGE fields in flight per CTA whatever the occupancy;
//   * the grid-constant arrays (xmapr, ymapr, fcoriolis) of the tile stay on chip for all the fields of
//     the block (a thread owns one column of the tile and parks its values in shared memory);
//   * neighbours are read from shared memory, conflict free (consecutive lanes, consecutive words).
// Fields are dense and unpadded, so a row segment starts anywhere relative to a 16-byte boundary
// (nx = 949 is odd; fields of a batch are 4-byte aligned only).  A bulk copy needs 16-byte aligned
// source, destination and size: each segment is copied from its aligned-down start and the consumer
// adds the segment's shift (0..3 elements) when it indexes the row.
//
// The tile kernel computes the interior points (1 <= x <= nx-2, 1 <= y <= ny-2).  The reference's
// flat loop also evaluates the edge-column points (with wrapped neighbours) for the undefined COUNT and
// then overwrites the border ring (fillEdges): stencil_edge_kernel does both for the ring.
#pragma once

#include "device_common.cuh"

#include <type_traits>

namespace fcb200 {
namespace tile {

constexpr int TX = 256;          // tile width = threads per CTA: a thread owns one column
constexpr int PITCH = TX + 8;    // floats per shared-memory row: halo columns + alignment slack, multiple of 4
constexpr int MAX_STAGES = 4;
constexpr int MAX_ARR = 3;
constexpr int MAX_FB = 32;       // fields per CTA, at most

__device__ __forceinline__ unsigned smem_u32(const void* p)
{
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
  asm volatile("{\n"
               ".reg .pred p;\n"
               "WAIT_%=:\n"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
               "@p bra DONE_%=;\n"
               "bra WAIT_%=;\n"
               "DONE_%=:\n"
               "}" ::"r"(smem_u32(bar)),
               "r"(parity)
               : "memory");
}
// TMA bulk copy global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
               "r"(smem_u32(bar))
               : "memory");
}

struct TileGeom
{
  int nx, ny, n;
  int nfields;
  int tiles_x, tiles_y;
  int field_blocks; // CTAs per tile: the grid is one-dimensional, blockIdx.x = tile * field_blocks + field block
  int fb;      // fields per CTA
  int stages;  // pipeline depth: fields in flight per CTA
  int period;  // fields k and k + period share their 16-byte alignment: 1, 2 or 4
  float undef;
  const FieldMeta* meta;
  unsigned long long* counters;
};

// Op interface on top of the scalar one (ops_stencil.cu):
//   static constexpr int TY;                       // tile height
//   static constexpr int NARR, NMAPS;              // per-field input arrays, grid-constant arrays
//   __host__ __device__ const float* arr(int k) const;   static constexpr int halo(int k);  // 1 = rows above/below are read
//   __host__ __device__ const float* map(int k) const;
//   template <bool ALL, class View> __device__ In<ALL> fetch(const View&, int r, const float* m) const;  // m[NMAPS]
//     with View::at<K>(rr, dc): array K at tile row rr (-1 .. TY where the halo is staged) and column delta dc (-1, 0, +1)
template <class Op>
struct TileLayout
{
  static constexpr int TY = Op::TY;
  static constexpr int HX = Op::HX;     // halo columns staged left and right of the tile (1, or 2 for the fused TFP)
  __host__ __device__ static constexpr int hmax(int k) { return k < 0 ? 0 : (Op::halo(k) > hmax(k - 1) ? Op::halo(k) : hmax(k - 1)); }
  static constexpr int HMAX = hmax(Op::NARR - 1); // largest row halo of the operator's arrays
  __host__ __device__ static constexpr int rows(int k) { return TY + 2 * Op::halo(k); }
  __host__ __device__ static constexpr int row0(int k) { return k == 0 ? 0 : row0(k - 1) + rows(k - 1); } // first row slot of array k
  static constexpr int ROWS = row0(Op::NARR);                                                             // row slots per stage
  static constexpr int STAGE_FLOATS = ROWS * PITCH;
  // 128 bytes of mbarriers, the stages, the operator's own scratch (the fused TFP's |grad T| tile), the maps
  static constexpr int MAP_FLOATS = Op::NMAPS * TY * TX; // the tile's grid-constant values, one private slot per consumer thread and row
  static constexpr size_t smem_bytes(int stages) { return 128 + ((size_t)stages * STAGE_FLOATS + Op::EXTRA_FLOATS + MAP_FLOATS) * sizeof(float); }
};

// What a consumer thread needs to index the staged rows of one field.  off[rr + HMAX] is the float offset,
// from the stage's base, of "array 0, tile row rr, this thread's column": the row's shift (0..3
// elements, it grows by nx & 3 per row) is folded in, so a neighbour is ONE LDS with an immediate
// offset.  Every per-field array of an operator has the same shift pattern (the host checks that
// they share their 16-byte alignment); array K's rows follow array 0's at a compile-time distance.
template <class Op>
struct TileView
{
  const float* stage;
  int off[Op::TY + 2 * TileLayout<Op>::HMAX];
  // array K at tile row rr (-halo(K) .. TY-1+halo(K)) and column delta dc (|dc| <= HX) from the thread's column
  template <int K>
  __device__ __forceinline__ float at(int rr, int dc) const
  {
    typedef TileLayout<Op> L;
    constexpr int KOFF = (L::row0(K) + Op::halo(K) - Op::halo(0)) * PITCH;
    return stage[off[rr + L::HMAX] + KOFF + dc];
  }
  // sh0 = shift of tile row 0, c = the thread's column counted from the first staged column
  __device__ __forceinline__ void set_rows(int sh0, int snx, int c)
  {
    typedef TileLayout<Op> L;
    int sh = (sh0 - L::HMAX * snx) & 3; // row -HMAX
#pragma unroll
    for (int q = 0; q < Op::TY + 2 * L::HMAX; ++q) {
      off[q] = (Op::halo(0) + q - L::HMAX) * PITCH + sh + c;
      sh = (sh + snx) & 3;
    }
  }
};

// The tile's grid-constant values (xmapr, ymapr, fcoriolis) stay on chip for every field of the block.  Each
// consumer thread parks the values of ITS column in shared memory (slot [k][r][c]: private to the thread, so
// no barrier; consecutive lanes, consecutive words) instead of 16..24 registers.
struct MapSlots
{
  const float* p; // this thread's column of map 0, row 0
  __device__ __forceinline__ float get(int k, int r, int ty) const { return p[(k * ty + r) * TX]; }
};

// 0.5f * m is exact in float (see half_map_diff_f in ops_stencil.cu)
__device__ __forceinline__ bool map_is_regular(float m)
{
  return fabsf(m) >= 7.8886e-31f /* 2^-100 */ || m == 0.f; // false for NaN
}

// the fields of a block in an order that keeps fields of equal alignment together (fields k and k + period
// start at the same offset from a 16-byte boundary): the consumers rebuild their row offsets only when
// the alignment changes
struct FieldOrder
{
  int period, nf, cls, f;
  __device__ __forceinline__ FieldOrder(int period_, int nf_) : period(period_), nf(nf_), cls(0), f(0) {}
  __device__ __forceinline__ void next()
  {
    f += period;
    if (f >= nf) {
      cls += 1;
      f = cls;
    }
  }
};

// Op::ASM_STORE = the operator's outputs are stored by a predicated instruction of its own through pointers that walk down the
// column (below).  Per operator, measured: relvort 0.80 -> 0.83, divergence 0.75 -> 0.78, 30 %-masked jacobian 0.57 -> 0.60 of the
// roofline, gradient unchanged, but the geostrophic operators lose a fifth (ilevelgwind 0.46 -> 0.36, plevelgvort 0.37 -> 0.30).
#ifdef FCB_TILE_ASM_STORE // (tools/shape_variants.sh: every operator)
template <class Op, class = void>
struct tile_asm_store : std::true_type
{
};
#else
template <class Op, class = void>
struct tile_asm_store : std::false_type
{
};
template <class Op>
struct tile_asm_store<Op, std::void_t<decltype(Op::ASM_STORE)>> : std::integral_constant<bool, Op::ASM_STORE>
{
};
#endif

template <class Op, bool ALL, bool FULL, bool FAST>
__device__ __forceinline__ unsigned tile_compute(const Op& op, const TileView<Op>& tv, const MapSlots& maps, int i0, bool col_ok, int nrows, int nx, float undef)
{
  unsigned nundef = 0;
  float* o[Op::NOUT];
#pragma unroll
  for (int k = 0; k < Op::NOUT; ++k)
    o[k] = op.out(k) + i0;
  constexpr bool ASM_STORE = tile_asm_store<Op>::value;
  const long long row_bytes = (long long)nx * (long long)sizeof(float);
#pragma unroll
  for (int r = 0; r < Op::TY; ++r) {
    if (FULL || r < nrows) { // warp-uniform
      float m[Op::NMAPS > 0 ? Op::NMAPS : 1];
#pragma unroll
      for (int k = 0; k < Op::NMAPS; ++k)
        m[k] = maps.get(k, r, Op::TY);
      // columns past the tile's last one compute on whatever the stage holds there; nothing is stored or counted
      const typename Op::template In<ALL> in = op.template fetch<ALL>(tv, r, m);
      float val[Op::NOUT];
      const bool ok = op.template eval<ALL, FAST>(in, undef, val);
      if ((!ALL || Op::TESTS_WHEN_ALL) && col_ok && !ok)
        nundef += 1;
#pragma unroll
      for (int k = 0; k < Op::NOUT; ++k) {
        const float v = ok ? val[k] : undef;
        if constexpr (ASM_STORE) {
          // the store as a predicated instruction of its own (the compiler otherwise wraps the whole row into a branch on col_ok:
          // BSSY / BRA / BSYNC per row); no memory clobber: the outputs are never read back, the staged inputs are in shared memory
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p st.global.f32 [%0], %1;\n\t}" ::"l"(o[k]), "f"(v), "r"((int)col_ok));
        } else if (col_ok) {
          o[k][r * nx] = v;
        }
      }
    }
    if constexpr (ASM_STORE) {
#pragma unroll
      for (int k = 0; k < Op::NOUT; ++k)
        asm("add.s64 %0, %0, %1;" : "+l"(o[k]) : "l"(row_bytes)); // the output pointers walk down the column (opaque to the optimiser,
                                                                  // which otherwise rebuilds base + r * nx every row: five instructions)
    }
  }
  return nundef;
}

constexpr int TILE_THREADS = TX + 32; // TX consumers (one column each) + one producer warp

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Warp-specialised: the last warp is the PRODUCER -- for each field it waits until the consumers have
// released the next stage ("empty" barrier), then its lanes issue the bulk copies of the field's row
// segments, which complete on the stage's "full" barrier.  The other warps are CONSUMERS: wait for
// "full", compute their column of the tile, store, release the stage.  Nothing else synchronises the
// CTA, so the producer runs up to `stages` fields ahead.
// resident CTAs per SM the kernel is compiled for (and the host sizes its pipeline for): Op::TILE_CTAS if the operator says so
template <class Op, class = void>
struct tile_ctas : std::integral_constant<int, 3>
{
};
template <class Op>
struct tile_ctas<Op, std::void_t<decltype(Op::TILE_CTAS)>> : std::integral_constant<int, Op::TILE_CTAS>
{
};

template <class Op, int MINB = tile_ctas<Op>::value>
__global__ void __launch_bounds__(TILE_THREADS, MINB) stencil_tile_kernel(const Op op0, const TileGeom g)
{
  typedef TileLayout<Op> L;
  constexpr int TY = Op::TY, NARR = Op::NARR, NMAPS = Op::NMAPS;
  constexpr int CONSUMER_WARPS = TX / 32;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned long long* full = reinterpret_cast<unsigned long long*>(smem_raw); // [MAX_STAGES]
  unsigned long long* empty = full + MAX_STAGES;                               // [MAX_STAGES]
  float* stage0 = reinterpret_cast<float*>(smem_raw + 128);
  __shared__ unsigned s_count[MAX_FB];

  const int nx = g.nx, ny = g.ny;
  const int tile = blockIdx.x / g.field_blocks;
  const int ty = tile / g.tiles_x, tx = tile - ty * g.tiles_x;
  const int x0 = 1 + tx * TX, y0 = 1 + ty * TY; // first output column / row of the tile
  const int xlast = min(x0 + TX - 1, nx - 2), ylast = min(y0 + TY - 1, ny - 2);
  const int f0 = (blockIdx.x - tile * g.field_blocks) * g.fb;
  const int nf = min(g.fb, g.nfields - f0);
  const int S = g.stages;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  if (threadIdx.x < MAX_FB)
    s_count[threadIdx.x] = 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 32);
      mbar_init(&empty[s], CONSUMER_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // consumers: park the grid-constant values of their column (rows of the tile) in shared memory and check
  // that the float shortcut for 0.5 * map * difference is exact for all of them
  float* extra = stage0 + (size_t)S * L::STAGE_FLOATS; // the operator's scratch, then the map slots
  MapSlots maps;
  maps.p = extra + Op::EXTRA_FLOATS + threadIdx.x;
  int irregular = 0;
  if (warp < CONSUMER_WARPS) {
    const int x = x0 + (int)threadIdx.x;
#pragma unroll
    for (int k = 0; k < NMAPS; ++k) {
      const float* mp = op0.map(k);
#pragma unroll
      for (int r = 0; r < TY; ++r) {
        const float m = (x <= xlast && y0 + r <= ylast) ? mp[(y0 + r) * nx + x] : 0.f;
        extra[Op::EXTRA_FLOATS + (k * TY + r) * TX + threadIdx.x] = m;
        irregular |= map_is_regular(m) ? 0 : 1;
      }
    }
  }
  const bool fast = __syncthreads_or(irregular) == 0;

  FieldOrder order(g.period, nf);
  int s = 0;
  unsigned phase = 0; // parity of the current round over the stages
  if (warp == CONSUMER_WARPS) {
    // ---------------------------------------------------------------- producer
    const int xs = max(x0 - L::HX, 0);                      // first staged column
    const int ncols = min(xlast + L::HX, nx - 1) - xs + 1; // .. last staged column
#pragma unroll 1
    for (int j = 0; j < nf; ++j) {
      mbar_wait(&empty[s], phase ^ 1u); // passes at once the first time a stage is used
      const Op op = op0.at(f0 + order.f, g.n);
      float* sbase = stage0 + (size_t)s * L::STAGE_FLOATS;
      // a lane copies row slots lane, lane+32, ...: first the byte count it will transfer, then the copies
      const float* src[(L::ROWS + 31) / 32];
      unsigned len[(L::ROWS + 31) / 32];
      unsigned bytes = 0;
#pragma unroll
      for (int q = 0; q < (L::ROWS + 31) / 32; ++q) {
        const int slot = lane + 32 * q;
        len[q] = 0;
        src[q] = nullptr;
        if (slot < L::ROWS) {
          int k = 0;
#pragma unroll
          for (int a = 1; a < NARR; ++a)
            if (slot >= L::row0(a))
              k = a;
          const int y = y0 + slot - L::row0(k) - Op::halo(k); // tile rows -halo .. TY-1+halo
          if (y >= 0 && y <= min(ylast + Op::halo(k), ny - 1)) {
            const float* p = op.arr(k) + (long long)y * nx + xs;
            const int sh = (int)((reinterpret_cast<uintptr_t>(p) >> 2) & 3);
            const int cols = ncols;
            src[q] = p - sh;
            len[q] = (unsigned)((sh + cols + 3) & ~3) * 4u;
            bytes += len[q];
          }
        }
      }
      mbar_arrive_expect_tx(&full[s], bytes);
#pragma unroll
      for (int q = 0; q < (L::ROWS + 31) / 32; ++q)
        if (len[q])
          bulk_g2s(sbase + (size_t)(lane + 32 * q) * PITCH, src[q], len[q], &full[s]);
      order.next();
      if (++s == S) {
        s = 0;
        phase ^= 1u;
      }
    }
  } else {
    // ---------------------------------------------------------------- consumers
    const int c = threadIdx.x, x = x0 + c;
    const bool col_ok = x <= xlast;
    const int nrows = ylast - y0 + 1;
    const int xs = max(x0 - L::HX, 0); // first staged column
    const int i0 = y0 * nx + x;
    TileView<Op> tv;
    int cur_sh0 = -1;
#pragma unroll 1
    for (int j = 0; j < nf; ++j) {
      const int field = f0 + order.f;
      const Op op = op0.at(field, g.n);
      const bool all = op0.all_defined(field, g.meta[field].all != 0);
      const int sh0 = (int)(((reinterpret_cast<uintptr_t>(op.arr(0)) >> 2) + (long long)y0 * nx + xs) & 3);
      if (sh0 != cur_sh0) { // warp-uniform
        tv.set_rows(sh0, nx & 3, c + (x0 - xs));
        cur_sh0 = sh0;
      }
      tv.stage = stage0 + (size_t)s * L::STAGE_FLOATS;
      mbar_wait(&full[s], phase);
      unsigned nundef;
      if (Op::CUSTOM_TILE) {
        nundef = op.tile_custom(tv, maps, extra, all, fast, x0, y0, xlast, ylast, c, nx, ny, g.undef);
      } else if (nrows == TY && fast) {
        if (all)
          nundef = tile_compute<Op, true, true, true>(op, tv, maps, i0, col_ok, nrows, nx, g.undef);
        else
          nundef = tile_compute<Op, false, true, true>(op, tv, maps, i0, col_ok, nrows, nx, g.undef);
      } else { // the grid's last tile row, or a tile with a tiny / NaN map ratio: the reference's double arithmetic
        if (all)
          nundef = tile_compute<Op, true, false, false>(op, tv, maps, i0, col_ok, nrows, nx, g.undef);
        else
          nundef = tile_compute<Op, false, false, false>(op, tv, maps, i0, col_ok, nrows, nx, g.undef);
      }
      __syncwarp();
      if (lane == 0)
        mbar_arrive(&empty[s]); // this warp no longer reads stage s
      if (!all || Op::TESTS_WHEN_ALL) {
        nundef = __reduce_add_sync(0xffffffffu, nundef);
        if (lane == 0 && nundef)
          atomicAdd(&s_count[order.f], nundef);
      }
      order.next();
      if (++s == S) {
        s = 0;
        phase ^= 1u;
      }
    }
  }
  __syncthreads();
  if ((int)threadIdx.x < nf && s_count[threadIdx.x])
    atomicAdd(g.counters + f0 + threadIdx.x, (unsigned long long)s_count[threadIdx.x]);
}

// The border ring of every field: (1) the reference's flat loop [lo, hi) also evaluates the ring cells
// it contains (edge columns; rows 0 and ny-1 for gradient c=1) with their wrapped neighbours and COUNTS
// them; (2) fillEdges (FC.cc:59-74) then sets out(x, y) = interior(clamp(x, 1, nx-2), clamp(y, 1, ny-2)).
// One thread per ring cell; the sources of (2) are interior cells, which this kernel never writes.
constexpr int EDGE_THREADS = 128;

template <class Op>
__global__ void __launch_bounds__(EDGE_THREADS) stencil_edge_kernel(const Op op0, int nx, int ny, int nfields, int lo, int hi, float undef,
                                                                     const FieldMeta* meta, unsigned long long* counters, bool count)
{
  const int ring = 2 * nx + 2 * (ny - 2);
  const int b = blockIdx.x * EDGE_THREADS + threadIdx.x;
  int x = 0, y = 0;
  const bool cell = b < ring;
  if (b < nx) {
    x = b;
    y = 0;
  } else if (b < 2 * nx) {
    x = b - nx;
    y = ny - 1;
  } else if (cell) {
    const int q = b - 2 * nx;
    y = 1 + (q >> 1);
    x = (q & 1) ? nx - 1 : 0;
  }
  const int sx = min(max(x, 1), nx - 2), sy = min(max(y, 1), ny - 2);
  const int dst = y * nx + x, src = sy * nx + sx;
  const int n = nx * ny;
  for (int field = blockIdx.y; field < nfields; field += gridDim.y) {
    const Op op = op0.at(field, n);
    if (count) {
      const bool all = op0.all_defined(field, meta[field].all != 0);
      unsigned bad = 0;
      if (cell && dst >= lo && dst < hi && (!all || Op::TESTS_WHEN_ALL)) {
        float val[Op::NOUT];
        if (all)
          bad = op.template eval<true>(op.template load<true>(dst, nx), undef, val) ? 0u : 1u;
        else
          bad = op.template eval<false>(op.template load<false>(dst, nx), undef, val) ? 0u : 1u;
      }
      bad = __reduce_add_sync(0xffffffffu, bad);
      if ((threadIdx.x & 31) == 0 && bad)
        atomicAdd(counters + field, (unsigned long long)bad);
    }
    if (cell) {
#pragma unroll
      for (int k = 0; k < Op::NOUT; ++k)
        op.out(k)[dst] = op.out(k)[src];
    }
  }
}

} // namespace tile
} // namespace fcb200

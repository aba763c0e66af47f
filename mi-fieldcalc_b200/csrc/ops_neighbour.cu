// ops_neighbour.cu -- neighbourhood statistics: neighbourProbFunctions (box probability through
// a summed-area table, FC.cc:2862-2953) and neighbourFunctions (window mean / max / min / percentile / probability on
// a coarse grid of window centres, painted as step x step blocks, FC.cc:2955-3061).
//
// Bit-exactness rules:
//   * the summed-area table is built with the reference's two sweeps and the reference's summation ORDER -- down
//     every column (one thread per column, coalesced), then along every row (one thread per row): float prefix sums
//     are order dependent for arbitrary data (compute other than 5 / 6 integrates whatever the output buffer held),
//     and exact in any order for the 0 / 1 indicator of compute 5 / 6;
//   * a window statistic is evaluated by ONE thread in the reference's row-major window order (float accumulation);
//   * the percentile is the ii-th smallest value of the window: a 32-step bitwise selection on order-preserving
//     integer keys instead of std::sort -- same element.
// Parameter combinations for which the reference indexes out of bounds (range < 0 or > min(nx, ny) for the SAT;
// step / 2 > range or a percentile index beyond the window) are rejected with 0: see oracle/fc_oracle.c.
#include "runtime.h"

#include "../../include/fcb200.h"

#include <cstddef>

namespace fcb200 {
namespace {

__global__ void __launch_bounds__(256) indicator_kernel(const float* __restrict__ f, float* __restrict__ o, long long n, float limit, int above)
{
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    o[i] = (above ? f[i] > limit : f[i] < limit) ? 1.f : 0.f;
}

// FC.cc:2899-2904: tmp(i, j) = fres(i, j) + tmp(i, j-1), sequential in j for every column i
__global__ void __launch_bounds__(128) sat_columns_kernel(const float* __restrict__ fres, float* __restrict__ tmp, int nx, int ny)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nx)
    return;
  float acc = fres[i];
  tmp[i] = acc;
  for (int j = 1; j < ny; ++j) {
    acc = fres[i + (size_t)j * nx] + acc;
    tmp[i + (size_t)j * nx] = acc;
  }
}

// FC.cc:2905-2908: tmp(i, j) += tmp(i-1, j), sequential in i for every row j.  A warp owns 32 rows and moves 32 columns at a
// time through a padded shared-memory tile, so that global accesses are coalesced while each thread still adds in order.
__global__ void __launch_bounds__(32) sat_rows_kernel(float* __restrict__ tmp, int nx, int ny)
{
  __shared__ float tile[32][33];
  const int lane = threadIdx.x;
  const int row0 = blockIdx.x * 32;
  float acc = 0.f;
  bool first = true;
  for (int x0 = 0; x0 < nx; x0 += 32) {
    // load: lane = column within the chunk, loop over the 32 rows
    for (int r = 0; r < 32; ++r) {
      const int y = row0 + r, x = x0 + lane;
      tile[r][lane] = (y < ny && x < nx) ? tmp[(size_t)y * nx + x] : 0.f;
    }
    __syncwarp();
    // scan: lane = row
    const int ncol = min(32, nx - x0);
    for (int c = 0; c < ncol; ++c) {
      if (first) {
        acc = tile[lane][c]; // column 0 is kept as it is
        first = false;
      } else {
        acc = tile[lane][c] + acc;
        tile[lane][c] = acc;
      }
    }
    __syncwarp();
    for (int r = 0; r < 32; ++r) {
      const int y = row0 + r, x = x0 + lane;
      if (y < ny && x < nx)
        tmp[(size_t)y * nx + x] = tile[r][lane];
    }
    __syncwarp();
  }
}

// FC.cc:2910-2949: box sum from four table entries (in the reference's order of operations), divided by N; border undefined
__global__ void __launch_bounds__(256) sat_window_kernel(const float* __restrict__ tmp, float* __restrict__ fres, int nx, int ny, int range, float undef)
{
  const long long n = (long long)nx * ny;
  const int N = (2 * range + 1) * (2 * range + 1);
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx / nx), i = (int)(idx - (long long)j * nx);
    if (i < range || i >= nx - range || j < range || j >= ny - range) {
      fres[idx] = undef;
      continue;
    }
    const int imax = i + range, jmax = j + range;
    float v = tmp[imax + (size_t)jmax * nx];
    if (i > range) {
      v -= tmp[i - range - 1 + (size_t)jmax * nx];
      if (j > range)
        v += tmp[i - range - 1 + (size_t)(j - range - 1) * nx] - tmp[imax + (size_t)(j - range - 1) * nx];
    } else if (j > range) {
      v -= tmp[imax + (size_t)(j - range - 1) * nx];
    }
    fres[idx] = v / (float)N;
  }
}

// compute 5 / 6 with a small window: the summed-area table of a 0 / 1 indicator holds exact integers (< 2^24), so the
// reference's four-entry difference IS the number of points of the window beyond the limit -- counted here directly
// (the window's rows come from L1), one thread per output point.  Same quotient count / N, same bits.
__global__ void __launch_bounds__(256) box_probability_kernel(const float* __restrict__ f, float* __restrict__ fres, int nx, int ny, int range, float limit,
                                                              int above, float undef)
{
  const long long n = (long long)nx * ny;
  const int N = (2 * range + 1) * (2 * range + 1);
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx / nx), i = (int)(idx - (long long)j * nx);
    if (i < range || i >= nx - range || j < range || j >= ny - range) {
      fres[idx] = undef;
      continue;
    }
    int count = 0;
    for (int y = j - range; y <= j + range; ++y) {
      const float* row = f + (size_t)y * nx;
      for (int x = i - range; x <= i + range; ++x)
        count += (above ? row[x] > limit : row[x] < limit) ? 1 : 0;
    }
    fres[idx] = (float)count / (float)N;
  }
}

__global__ void __launch_bounds__(256) border_undef_kernel(float* __restrict__ fres, int nx, int ny, int range, float undef)
{
  const long long n = (long long)nx * ny;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx / nx), i = (int)(idx - (long long)j * nx);
    if (i < range || i >= nx - range || j < range || j >= ny - range)
      fres[idx] = undef;
  }
}

__device__ __forceinline__ unsigned order_key(float v)
{ // monotone map float -> unsigned (-0 and +0 get the same key, like operator<)
  if (v == 0.f)
    v = 0.f;
  const unsigned b = __float_as_uint(v);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// one thread per window centre (FC.cc:3017-3058)
__global__ void __launch_bounds__(128) neighbour_windows_kernel(const float* __restrict__ field, float* __restrict__ fres, int nx, int ny, int range, int step,
                                                                int compute, int limit, int ii, int nwx, int nwy)
{
  const long long nwin = (long long)nwx * nwy;
  const float ngridp = (float)((2 * range + 1) * (2 * range + 1));
  const float flimit = (float)limit;
  for (long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x; w < nwin; w += (long long)gridDim.x * blockDim.x) {
    const int wj = (int)(w / nwx), wi = (int)(w - (long long)wj * nwx);
    const int i = range + wi * step, j = range + wj * step;
    float value = 0.f;
    if (compute == 2 || compute == 3)
      value = field[(i - range) + (size_t)(j - range) * nx];
    if (compute == 4) {
      // ii-th smallest (0-based) by bitwise selection: at each bit, count the candidates (same prefix) with a 0 bit.  Windows of
      // up to 11 x 11 points are first copied (as keys) into a per-thread array: 32 passes over L1-resident local memory
      // instead of 32 passes over the field.
      constexpr int MAXW = 121;
      unsigned keys[MAXW];
      const int side = 2 * range + 1, count = side * side;
      const bool local = count <= MAXW;
      if (local) {
        int n = 0;
        for (int y = j - range; y <= j + range; ++y)
          for (int x = i - range; x <= i + range; ++x)
            keys[n++] = order_key(field[x + (size_t)y * nx]);
      }
      unsigned prefix = 0, mask = 0;
      int k = ii;
      for (int bit = 31; bit >= 0; --bit) {
        const unsigned b = 1u << bit;
        int zeros = 0;
        if (local) {
          for (int n = 0; n < count; ++n)
            zeros += ((keys[n] & mask) == prefix && !(keys[n] & b)) ? 1 : 0;
        } else {
          for (int y = j - range; y <= j + range; ++y)
            for (int x = i - range; x <= i + range; ++x) {
              const unsigned key = order_key(field[x + (size_t)y * nx]);
              zeros += ((key & mask) == prefix && !(key & b)) ? 1 : 0;
            }
        }
        if (k >= zeros) {
          k -= zeros;
          prefix |= b;
        }
        mask |= b;
      }
      // the element itself (the key is a bijection apart from the sign of zero: return the first window value with this key)
      for (int y = j - range; y <= j + range; ++y)
        for (int x = i - range; x <= i + range; ++x) {
          const float v = field[x + (size_t)y * nx];
          if (order_key(v) == prefix) {
            value = v;
            y = j + range + 1;
            break;
          }
        }
    } else {
      for (int y = j - range; y <= j + range; ++y)
        for (int x = i - range; x <= i + range; ++x) {
          const float v = field[x + (size_t)y * nx];
          if (compute == 1)
            value += v;
          if ((compute == 2 && v > value) || (compute == 3 && v < value))
            value = v;
          if ((compute == 5 && v > flimit) || (compute == 6 && v < flimit))
            value += 1.f;
        }
    }
    if (compute == 1 || compute > 4)
      value /= ngridp;
    for (int l = j - (step - 1) / 2; l < j + step / 2 + 1; ++l)
      for (int k2 = i - (step - 1) / 2; k2 < i + step / 2 + 1; ++k2)
        fres[k2 + (size_t)l * nx] = value;
  }
}

unsigned grid_for(long long n, int threads)
{
  long long b = (n + threads - 1) / threads;
  const long long cap = (long long)sm_count() * 32;
  if (b > cap)
    b = cap;
  return (unsigned)(b < 1 ? 1 : b);
}

bool grid_valid(int nx, int ny)
{
  if (nx <= 0 || ny <= 0 || (long long)nx * ny >= 0x7fffffffLL) {
    set_error("fcb200: invalid grid (nx=%d ny=%d)", nx, ny);
    return false;
  }
  return true;
}

} // namespace
} // namespace fcb200

using namespace fcb200;

extern "C" {

int fcb200_neighbourProbFunctions(int nx, int ny, const float* field, const float* constants, int nconstants, int compute, float* fres, int* fDefined,
                                  float undef)
{ // FC.cc:2862-2953
  if (*fDefined != ALL_DEFINED)
    return 0;
  if (nconstants < 2)
    return 0;
  const int limit = (int)constants[0];
  const int range = (int)constants[1];
  if (range < 0 || range > nx || range > ny)
    return 0;
  const bool indicator = (compute == 5 || compute == 6);
  if (!indicator && range == 0)
    return 1; // nothing is written at all (FC.cc:2894): the caller's buffer and flag stay as they are
  if (!grid_valid(nx, ny))
    return -1;
  const long long n = (long long)nx * ny;
  Call call;
  if (!call.ok())
    return -1;
  const float* d_f = indicator ? call.in(field, (size_t)n) : nullptr;
  // compute other than 5 / 6 integrates what the output buffer holds (the reference never writes it before the sweep)
  float* d_o = indicator ? call.out(fres, (size_t)n) : call.inout(fres, (size_t)n);
  if (!call.ok())
    return -1;
  cudaStream_t s = call.stream();
  if (indicator && range >= 1 && range <= 8 && n < (1LL << 24) && static_cast<const void*>(d_f) != static_cast<const void*>(d_o)) {
    box_probability_kernel<<<grid_for(n, 256), 256, 0, s>>>(d_f, d_o, nx, ny, range, (float)limit, compute == 5 ? 1 : 0, undef);
    count_launch();
    return call.finish([=](const unsigned long long*) { *fDefined = SOME_DEFINED; });
  }
  if (indicator) {
    indicator_kernel<<<grid_for(n, 256), 256, 0, s>>>(d_f, d_o, n, (float)limit, compute == 5 ? 1 : 0);
    count_launch();
  }
  if (range == 0)
    return call.finish(Finalizer());
  float* d_tmp = static_cast<float*>(call.scratch(sizeof(float) * (size_t)n));
  if (!call.ok())
    return -1;
  sat_columns_kernel<<<(unsigned)((nx + 127) / 128), 128, 0, s>>>(d_o, d_tmp, nx, ny);
  sat_rows_kernel<<<(unsigned)((ny + 31) / 32), 32, 0, s>>>(d_tmp, nx, ny);
  sat_window_kernel<<<grid_for(n, 256), 256, 0, s>>>(d_tmp, d_o, nx, ny, range, undef);
  count_launch(3);
  return call.finish([=](const unsigned long long*) { *fDefined = SOME_DEFINED; });
}

int fcb200_neighbourFunctions(int nx, int ny, const float* field, const float* constants, int nconstants, int compute, float* fres, int* fDefined,
                              float undef)
{ // FC.cc:2955-3061
  if (*fDefined != ALL_DEFINED)
    return 0;
  if (nconstants < 1 || (nconstants < 2 && compute > 3))
    return 0;
  int range = 3, step = 3, limit = 0;
  if (compute < 4) {
    range = (int)constants[0];
    if (nconstants == 2)
      step = (int)constants[1];
  } else {
    limit = (int)constants[0];
    range = (int)constants[1];
    if (nconstants == 3)
      step = (int)constants[2];
  }
  if (range > nx || range > ny || range < 1)
    return 0;
  if (step < 1)
    return 0;
  const int nwin1 = (2 * range + 1) * (2 * range + 1);
  const float ngridp = (float)nwin1;
  const int ii = (int)(ngridp * limit / 100);
  if (step / 2 > range)
    return 0;
  if (compute == 4 && (ii < 0 || ii >= nwin1))
    return 0;
  if (!grid_valid(nx, ny))
    return -1;
  const long long n = (long long)nx * ny;
  Call call;
  if (!call.ok())
    return -1;
  const float* d_f = call.in(field, (size_t)n);
  float* d_o = call.inout(fres, (size_t)n); // points outside the border and the painted blocks keep their old value
  if (!call.ok())
    return -1;
  cudaStream_t s = call.stream();
  border_undef_kernel<<<grid_for(n, 256), 256, 0, s>>>(d_o, nx, ny, range, undef);
  count_launch();
  const int nwx = (nx - 2 * range > 0) ? (nx - 2 * range + step - 1) / step : 0;
  const int nwy = (ny - 2 * range > 0) ? (ny - 2 * range + step - 1) / step : 0;
  if (nwx > 0 && nwy > 0) {
    neighbour_windows_kernel<<<grid_for((long long)nwx * nwy, 128), 128, 0, s>>>(d_f, d_o, nx, ny, range, step, compute, limit, ii, nwx, nwy);
    count_launch();
  }
  return call.finish([=](const unsigned long long*) { *fDefined = SOME_DEFINED; });
}

// batched twins: field by field (every field has its own flag-dependent early return)
int fcb200_neighbourProbFunctions_batched(int nx, int ny, int nfields, const float* field, const float* constants, int nconstants, int compute, float* fres,
                                          int* fDefined, float undef)
{
  const size_t n = (size_t)nx * (size_t)ny;
  for (int k = 0; k < nfields; ++k)
    if (fDefined[k] != ALL_DEFINED)
      return 0;
  for (int k = 0; k < nfields; ++k) {
    const int rc = fcb200_neighbourProbFunctions(nx, ny, field + k * n, constants, nconstants, compute, fres + k * n, fDefined + k, undef);
    if (rc != 1)
      return rc;
  }
  return 1;
}
int fcb200_neighbourFunctions_batched(int nx, int ny, int nfields, const float* field, const float* constants, int nconstants, int compute, float* fres,
                                      int* fDefined, float undef)
{
  const size_t n = (size_t)nx * (size_t)ny;
  for (int k = 0; k < nfields; ++k)
    if (fDefined[k] != ALL_DEFINED)
      return 0;
  for (int k = 0; k < nfields; ++k) {
    const int rc = fcb200_neighbourFunctions(nx, ny, field + k * n, constants, nconstants, compute, fres + k * n, fDefined + k, undef);
    if (rc != 1)
      return rc;
  }
  return 1;
}

} // extern "C"

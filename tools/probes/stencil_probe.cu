// stencil_probe.cu -- throwaway microbenchmark: which ingredient of relvort keeps it below the copy rate?
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false -o stencil_probe stencil_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int T = 256, U = 4;

template <int VAR>
__global__ void __launch_bounds__(T) k(const float* __restrict__ u, const float* __restrict__ v, const float* __restrict__ xm, const float* __restrict__ ym,
                                        float* __restrict__ o, int nx, int n, int chunks)
{
  const int field = blockIdx.y;
  const long long off = (long long)field * n;
  u += off; v += off; o += off;
  const int i0 = nx + blockIdx.x * (T * U) + threadIdx.x;
  float a[U], b[U], c[U], d[U], m1[U], m2[U];
#pragma unroll
  for (int q = 0; q < U; ++q) {
    const int i = i0 + q * T;
    if (i < n - nx) {
      if (VAR == 0) { a[q] = u[i]; b[q] = v[i]; }
      else if (VAR == 6) { a[q] = u[i]; b[q] = v[i]; m1[q] = xm[i]; m2[q] = ym[i]; }
      else if (VAR == 7) { a[q] = u[i]; b[q] = v[i]; c[q] = u[i + nx]; }
      else if (VAR == 8) { a[q] = u[i]; b[q] = v[i]; c[q] = u[i + nx]; m1[q] = xm[i]; m2[q] = ym[i]; }
      else {
        a[q] = v[i - 1]; b[q] = v[i + 1];
        if (VAR == 5) { c[q] = u[i]; d[q] = u[i]; } else { c[q] = u[i - nx]; d[q] = u[i + nx]; }
        if (VAR == 1 || VAR == 2 || VAR == 5) { m1[q] = xm[i]; m2[q] = ym[i]; } else { m1[q] = 2e-4f; m2[q] = 2.1e-4f; }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < U; ++q) {
    const int i = i0 + q * T;
    if (i < n - nx) {
      float r;
      if (VAR == 0) r = a[q] + b[q];
      else if (VAR == 6) r = a[q] + b[q] + m1[q] + m2[q];
      else if (VAR == 7) r = a[q] + b[q] + c[q];
      else if (VAR == 8) r = a[q] + b[q] + c[q] + m1[q] + m2[q];
      else if (VAR == 1 || VAR == 4) r = 0.5f * m1[q] * (b[q] - a[q]) - 0.5f * m2[q] * (d[q] - c[q]);
      else r = (float)(0.5 * (double)m1[q] * (double)(b[q] - a[q]) - 0.5 * (double)m2[q] * (double)(d[q] - c[q]));
      o[i] = r;
    }
  }
}

template <int FPC, bool PREFETCH>
__global__ void __launch_bounds__(T) kloop(const float* __restrict__ u, const float* __restrict__ v, const float* __restrict__ xm, const float* __restrict__ ym,
                                           float* __restrict__ o, int nx, int n, int nf)
{
  const int c = blockIdx.x & 3, s = blockIdx.x >> 2;
  const int i0 = nx + blockIdx.y * (T * U) + threadIdx.x;
  double m1[U], m2[U];
#pragma unroll
  for (int q = 0; q < U; ++q) {
    const int i = min(i0 + q * T, n - nx - 1);
    m1[q] = 0.5 * (double)xm[i];
    m2[q] = 0.5 * (double)ym[i];
  }
  float a[U], b[U], cc[U], d[U];
  int f = c + 4 * (s * FPC);
  if (PREFETCH && f < nf) {
    const float* uu = u + (long long)f * n; const float* vv = v + (long long)f * n;
#pragma unroll
    for (int q = 0; q < U; ++q) {
      const int i = min(i0 + q * T, n - nx - 1);
      a[q] = vv[i - 1]; b[q] = vv[i + 1]; cc[q] = uu[i - nx]; d[q] = uu[i + nx];
    }
  }
#pragma unroll 1
  for (int j = 0; j < FPC; ++j, f += 4) {
    if (f >= nf) break;
    float* oo = o + (long long)f * n;
    float a2[U], b2[U], c2[U], d2[U];
    if (!PREFETCH) {
      const float* uu = u + (long long)f * n; const float* vv = v + (long long)f * n;
#pragma unroll
      for (int q = 0; q < U; ++q) {
        const int i = min(i0 + q * T, n - nx - 1);
        a[q] = vv[i - 1]; b[q] = vv[i + 1]; cc[q] = uu[i - nx]; d[q] = uu[i + nx];
      }
    } else if (j + 1 < FPC && f + 4 < nf) {
      const float* uu = u + (long long)(f + 4) * n; const float* vv = v + (long long)(f + 4) * n;
#pragma unroll
      for (int q = 0; q < U; ++q) {
        const int i = min(i0 + q * T, n - nx - 1);
        a2[q] = vv[i - 1]; b2[q] = vv[i + 1]; c2[q] = uu[i - nx]; d2[q] = uu[i + nx];
      }
    }
#pragma unroll
    for (int q = 0; q < U; ++q) {
      const int i = i0 + q * T;
      if (i < n - nx)
        oo[i] = (float)(m1[q] * (double)(b[q] - a[q]) - m2[q] * (double)(d[q] - cc[q]));
    }
    if (PREFETCH) {
#pragma unroll
      for (int q = 0; q < U; ++q) { a[q] = a2[q]; b[q] = b2[q]; cc[q] = c2[q]; d[q] = d2[q]; }
    }
  }
}

template <int FPC, bool PF>
void runloop(const float* u, const float* v, const float* xm, const float* ym, float* o, int nx, int n, int nf, cudaEvent_t e0, cudaEvent_t e1)
{
  const int chunks = (n - 2 * nx + T * U - 1) / (T * U);
  const int per_class = (nf + 3) / 4;
  dim3 grid(4 * ((per_class + FPC - 1) / FPC), chunks);
  float best = 1e9f;
  for (int rep = 0; rep < 12; ++rep) {
    const size_t s = (size_t)(rep & 1) * n * nf;
    cudaEventRecord(e0);
    kloop<FPC, PF><<<grid, T>>>(u + s, v + s, xm, ym, o + s, nx, n, nf);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep >= 2 && ms < best) best = ms;
  }
  printf("field loop FPC=%2d prefetch=%d                      %8.4f ms  %7.1f GB/s (12 B/pt)\n", FPC, (int)PF, best, 12.0 * n * nf / best * 1e-6);
}

int main(int argc, char** argv)
{
  const int nx = 949, ny = 1069, n = nx * ny, nf = 64;
  const int reps = argc > 1 ? atoi(argv[1]) : 12;
  float *u, *v, *o, *xm, *ym;
  cudaMalloc(&u, sizeof(float) * (size_t)n * nf * 2);
  cudaMalloc(&v, sizeof(float) * (size_t)n * nf * 2);
  cudaMalloc(&o, sizeof(float) * (size_t)n * nf * 2);
  cudaMalloc(&xm, sizeof(float) * n);
  cudaMalloc(&ym, sizeof(float) * n);
  cudaMemset(u, 0, sizeof(float) * (size_t)n * nf * 2);
  cudaMemset(v, 0, sizeof(float) * (size_t)n * nf * 2);
  cudaMemset(xm, 0, sizeof(float) * n);
  cudaMemset(ym, 0, sizeof(float) * n);
  const int chunks = (n - 2 * nx + T * U - 1) / (T * U);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const char* names[] = {"0 copy-like u[i]+v[i]", "1 relvort loads, float math, maps", "2 relvort loads, double math, maps", "3 relvort loads, double math, no maps",
                         "4 relvort loads, float math, no maps", "5 u[i] twice instead of u[i+-nx], double, maps", "6 copy-like + maps", "7 copy-like + u[i+nx]",
                         "8 copy-like + u[i+nx] + maps"};
  for (int ai = 2; ai < (argc > 2 ? argc : 11); ++ai) {
    const int var = argc > 2 ? atoi(argv[ai]) : ai - 2;
    float best = 1e9f;
    for (int rep = 0; rep < reps; ++rep) {
      const size_t s = (size_t)(rep & 1) * n * nf;
      cudaEventRecord(e0);
      dim3 grid(chunks, nf);
      switch (var) {
      case 0: k<0><<<grid, T>>>(u + s, v + s, xm, ym, o + s, nx, n, chunks); break;
      case 1: k<1><<<grid, T>>>(u + s, v + s, xm, ym, o + s, nx, n, chunks); break;
      case 2: k<2><<<grid, T>>>(u + s, v + s, xm, ym, o + s, nx, n, chunks); break;
      case 3: k<3><<<grid, T>>>(u + s, v + s, xm, ym, o + s, nx, n, chunks); break;
      case 4: k<4><<<grid, T>>>(u + s, v + s, xm, ym, o + s, nx, n, chunks); break;
      case 5: k<5><<<grid, T>>>(u + s, v + s, xm, ym, o + s, nx, n, chunks); break;
      case 6: k<6><<<grid, T>>>(u + s, v + s, xm, ym, o + s, nx, n, chunks); break;
      case 7: k<7><<<grid, T>>>(u + s, v + s, xm, ym, o + s, nx, n, chunks); break;
      case 8: k<8><<<grid, T>>>(u + s, v + s, xm, ym, o + s, nx, n, chunks); break;
      }
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep >= 2 && ms < best) best = ms;
    }
    printf("%-50s %8.4f ms  %7.1f GB/s (12 B/pt)\n", names[var], best, 12.0 * n * nf / best * 1e-6);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

// Issue rate of the packed FP32 instructions (FFMA2 / FMUL2 / FADD2) against the scalar ones on sm_100a:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/ffma2_probe tools/probes/ffma2_probe.cu && tools/probes/ffma2_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b)
{
  float2 x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
    x[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f - i);
  const float2 A = make_float2(a, a), B = make_float2(b, b);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { // 8 scalar FFMA
        x[i].x = fmaf(x[i].x, a, b);
      } else if (MODE == 1) { // 8 packed FFMA2 (16 flops pairs)
        x[i] = __ffma2_rn(x[i], A, B);
      } else if (MODE == 2) { // 16 scalar FFMA
        x[i].x = fmaf(x[i].x, a, b);
        x[i].y = fmaf(x[i].y, a, b);
      } else if (MODE == 3) { // 4 FFMA2 + 4 FFMA interleaved
        if (i & 1)
          x[i] = __ffma2_rn(x[i], A, B);
        else
          x[i].x = fmaf(x[i].x, a, b);
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    s += x[i].x + x[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int instr_per_iter)
{
  float* out;
  cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<MODE><<<148 * 8, 256>>>(out, 100, 1.0001f, 0.5f);
  cudaEventRecord(e0);
  k<MODE><<<148 * 8, 256>>>(out, iters, 1.0001f, 0.5f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const double warp_instr = (double)148 * 8 * 8 * iters * instr_per_iter; // warps * iters * instr
  const double cycles = ms * 1e-3 * clk_khz * 1e3;
  printf("%-28s %8.3f ms  %.2f warp-instr/clk/SM (at %d MHz nominal)\n", name, ms, warp_instr / cycles / 148, clk_khz / 1000);
  cudaFree(out);
}

int main()
{
  run<0>("8 FFMA", 8);
  run<1>("8 FFMA2", 8);
  run<2>("16 FFMA", 16);
  run<3>("4 FFMA2 + 4 FFMA", 8);
  return 0;
}

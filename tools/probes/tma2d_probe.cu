// Stand-alone check of the 2-D tensor-map copy the TFP tile kernel uses (box 68 x 12 floats, no swizzle, negative start column):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/tma2d_probe tools/probes/tma2d_probe.cu && tools/probes/tma2d_probe
#include <cstdio>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>
#include <vector>

constexpr int BOXH = 12;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void k(const __grid_constant__ CUtensorMap tmap, int cx, int cy, float* out, int BOXW)
{
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem);
  float* stage = reinterpret_cast<float*>(smem + 128);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(32) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const unsigned bytes = threadIdx.x == 0 ? BOXW * BOXH * 4 : 0;
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
  if (threadIdx.x == 0) {
#ifdef USE_HINT
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(stage)),
                 "l"(reinterpret_cast<unsigned long long>(&tmap)), "r"(smem_u32(bar)), "r"(cx), "r"(cy), "l"(0x1000000000000000ull)
                 : "memory");
#else
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(stage)),
                 "l"(reinterpret_cast<unsigned long long>(&tmap)), "r"(smem_u32(bar)), "r"(cx), "r"(cy)
                 : "memory");
#endif
  }
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(smem_u32(bar)), "r"(0) : "memory");
  for (int i = threadIdx.x; i < BOXW * BOXH; i += 32)
    out[i] = stage[i];
}

int main()
{
  const int nx = 300, rows = 80;
  std::vector<float> h((size_t)nx * rows);
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < nx; ++x)
      h[(size_t)y * nx + x] = y * 1000.f + x;
  float *d, *o;
  cudaMalloc(&d, h.size() * 4);
  cudaMalloc(&o, 128 * BOXH * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  typedef CUresult (*Encode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  printf("entry point: %d %d %p\n", (int)e, (int)q, fn);
  for (int BOXW : {BOXW0}) {
  CUtensorMap tmap;
  const cuuint64_t dims[2] = {(cuuint64_t)nx, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)nx * 4};
  const cuuint32_t box[2] = {(cuuint32_t)BOXW, BOXH};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = ((Encode)fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("box %d: encode: %d  map bytes:", BOXW, (int)r);
  for (int i = 0; i < 16; ++i) printf(" %016llx", ((unsigned long long*)&tmap)[i]);
  printf("\n");
  for (int cx : {5, -3, 250}) {
    k<<<1, 32, 128 + BOXW * BOXH * 4>>>(tmap, cx, 7, o, BOXW);
    e = cudaDeviceSynchronize();
    std::vector<float> got(BOXW * BOXH);
    cudaMemcpy(got.data(), o, got.size() * 4, cudaMemcpyDeviceToHost);
    printf("cx=%d: %s  row0: %g %g %g %g ... %g  row11: %g\n", cx, cudaGetErrorString(e), got[0], got[1], got[3], got[4], got[BOXW - 1], got[11 * BOXW + 4]);
  }
  }
  return 0;
}

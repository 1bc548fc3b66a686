#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tmap, float* out, int x, int y, int c, int n, int bytes, int cnt) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint64_t* bar = (uint64_t*)(sm + 65536);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(s32(sm)), "l"(&tmap), "r"(x), "r"(y), "r"(c), "r"(n), "r"(s32(bar)) : "memory");
  }
  uint32_t ok = 0;
  while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(s32(bar)) : "memory");
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) out[i] = ((float*)sm)[i];
}
int main(int argc, char** argv) {
  int W = atoi(argv[1]), H = atoi(argv[2]), C = atoi(argv[3]), N = atoi(argv[4]);
  int bw = atoi(argv[5]), bh = atoi(argv[6]), bc = atoi(argv[7]);
  int x = atoi(argv[8]), y = atoi(argv[9]);
  int promo = atoi(argv[10]);
  float* d; size_t n = (size_t)W * H * C * N;
  cudaMalloc(&d, n * 4);
  float* h = (float*)malloc(n * 4);
  for (size_t i = 0; i < n; ++i) h[i] = (float)(i % 1000);
  cudaMemcpy(d, h, n * 4, cudaMemcpyHostToDevice);
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  auto enc = (PFN_cuTensorMapEncodeTiled_v12000)fp;
  CUtensorMap tm;
  cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)N};
  cuuint64_t str[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * C * 4};
  cuuint32_t box[4] = {(cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bc, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, promo ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d ", (int)r);
  int cnt = bw * bh * bc;
  float* o; cudaMalloc(&o, cnt * 4);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 66000);
  k<<<1, 128, 65536 + 64>>>(tm, o, x, y, 0, 0, cnt * 4, cnt);
  cudaError_t e = cudaDeviceSynchronize();
  float* ho = (float*)malloc(cnt * 4);
  cudaMemcpy(ho, o, cnt * 4, cudaMemcpyDeviceToHost);
  // check first row
  int bad = 0;
  for (int cc = 0; cc < bc; ++cc) for (int yy = 0; yy < bh; ++yy) for (int xx = 0; xx < bw; ++xx) {
    int gx = x + xx, gy = y + yy; float want = 0;
    if (gx >= 0 && gx < W && gy >= 0 && gy < H) want = h[((size_t)cc * H + gy) * W + gx];
    if (ho[(cc * bh + yy) * bw + xx] != want) ++bad;
  }
  printf("sync=%s bad=%d\n", cudaGetErrorString(e), bad);
  return 0;
}

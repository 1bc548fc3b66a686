// tcgen05 tensor-core GEMM path (placeholder until the UMMA kernel lands):
// reports "shape not handled" so a2c_gemm falls through to the fp32 FFMA kernel.
#include "common.cuh"
namespace a2c {
int gemm_tc_try(const float*, long long, long long, const float*, long long, long long, float*,
                long long, const float*, int, int, int, int, int, int, void*, long long,
                cudaStream_t) {
  return 1;
}
}  // namespace a2c

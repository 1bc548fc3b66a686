// Shared helpers for the a2c_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/a2c_b200.h"

namespace a2c {

// Last error text, per host thread. Entry points never throw; they return a
// cudaError_t-compatible int (0 = ok, negative = argument error).
void set_error(const char* fmt, ...);

#define A2C_ARG_CHECK(cond, ...)                 \
  do {                                           \
    if (!(cond)) {                               \
      ::a2c::set_error(__VA_ARGS__);             \
      return A2C_ERR_ARG;                        \
    }                                            \
  } while (0)

#define A2C_CUDA_CHECK(expr)                                                   \
  do {                                                                         \
    cudaError_t e__ = (expr);                                                  \
    if (e__ != cudaSuccess) {                                                  \
      ::a2c::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                       __FILE__, __LINE__);                                    \
      return (int)e__;                                                         \
    }                                                                          \
  } while (0)

#define A2C_LAUNCH_CHECK() A2C_CUDA_CHECK(cudaGetLastError())

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device): the attribute is
// per device, so a process-wide "done" flag would leave a second GPU without the opt-in.
cudaError_t set_max_dynamic_smem(const void* kernel, int bytes);

// Profiling knobs (stage isolation etc.) exist only in builds made with -DA2C_DEBUG_KNOBS;
// release builds never read the environment on a launch path.
inline int debug_knob(const char* name) {
#ifdef A2C_DEBUG_KNOBS
  const char* e = getenv(name);
  return e ? atoi(e) : 0;
#else
  (void)name;
  return 0;
#endif
}

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

__host__ __device__ inline int cdiv(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline long long cdivll(long long a, long long b) { return (a + b - 1) / b; }

// ---- device helpers -------------------------------------------------------

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum; result valid in thread 0. `scratch` holds >= 32 elements.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* scratch) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? scratch[threadIdx.x] : T(0);
  if (w == 0) v = warp_sum(v);
  return v;
}

// Asks L2 for the 128-byte line holding p (no register result, no scoreboard): the producers of the
// convolution kernels call it one work item ahead, so that the loads of the item they convert next
// find their data in L2 instead of paying HBM latency inside the wait -> load -> store -> MMA chain.
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier / TMA (sm_90+ PTX, used as on sm_100a) ------------------------

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// For waits that are expected to be long (a role that is ahead of the pipeline): back off
// so the spinning warp does not take issue slots from the warps doing the work.  The wait is
// given a suspend-time hint (the hardware parks the warp until the phase completes or the hint
// expires) and sleeps between polls.
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity, unsigned ns = 64) {
  if (mbar_try_wait(bar, parity)) return;
  while (!mbar_try_wait_hint(bar, parity, ns)) __nanosleep(ns);
}

// A group of `nwarps` warps waits for the same mbarrier phase: only the group's first warp (`leader`)
// watches the mbarrier, the others park at named hardware barrier `bar_id` (1..15, one id per group).
// A warp waiting on an mbarrier is woken by every arrival on any mbarrier of the CTA and re-polls;
// with 8-16 waiters that polling took 25-60 % of the instructions these kernels issued (ncu, round 2).
// All warps of the group must call this the same number of times.
__device__ __forceinline__ void mbar_wait_group(uint64_t* bar, uint32_t parity, bool leader, int bar_id, int nwarps,
                                                unsigned ns = 64) {
  if (leader) mbar_wait_relaxed(bar, parity, ns);
  asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(32 * nwarps) : "memory");
}

// Stage index + phase parity of a circular mbarrier pipeline, advanced without div/mod.
struct StageRing {
  int s, n;
  uint32_t par;
  __device__ __forceinline__ explicit StageRing(int n_) : s(0), n(n_), par(0) {}
  __device__ __forceinline__ void next() {
    if (++s == n) { s = 0; par ^= 1u; }
  }
};

}  // namespace a2c

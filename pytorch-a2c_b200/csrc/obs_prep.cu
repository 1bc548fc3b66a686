// Stage 1: fused Pong observation prep + frame-stack ring.
//
// Replaces a2c/preprocessing.py:11-17 (crop rows 35:195, every 2nd row/col,
// channel 0, 144/109 -> 0, else non-zero -> 1) and the plane store behind
// a2c/utils.py:26-43 / a2c/runner.py:199.
//
// Per frame the kernel touches only the 80 selected rows (80 x 480 B = 38,400 B,
// each row 32-B-sector aligned) and writes one 6,400 B u8 plane into the ring:
// 44,800 algorithmic bytes per transition, HBM-bound.
//
// Variant 0 (default): persistent CTAs, one 3-D TMA box (120 u32 x 80 rows x 1
// frame; row pitch 960 B skips the odd rows) per frame into a shared memory ring
// guarded by mbarriers (two CTAs per SM with a 2-deep ring each: one CTA's loads
// and classification fill the other's barrier / refill bubbles); threads then
// pick byte 6k of every row with 128-bit shared loads and classify 4 pixels per
// SIMD compare.  Variant 2: one CTA per SM with a 4-deep ring (the round-1 shape).
// Variant 1: same arithmetic straight from global memory with 128-bit loads.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdarg.h>

#include <atomic>
#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"

namespace a2c {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
std::atomic<long long> g_launches{0};

cudaError_t set_max_dynamic_smem(const void* kernel, int bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, int> done;      // (kernel, device) -> bytes granted
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(mu);
  auto it = done.find({kernel, dev});
  if (it != done.end() && it->second >= bytes) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) done[{kernel, dev}] = bytes;
  return e;
}

constexpr int FRAME_H = 210, FRAME_W = 160, FRAME_C = 3;
constexpr int ROW_BYTES = FRAME_W * FRAME_C;         // 480
constexpr int FRAME_BYTES = FRAME_H * ROW_BYTES;     // 100800
constexpr int CROP_TOP = 35;
constexpr int OUT = 80;                              // 80 x 80 plane
constexpr int PLANE_BYTES = OUT * OUT;               // 6400
constexpr int SEL_BYTES = OUT * ROW_BYTES;           // 38400 selected-row bytes per frame
constexpr int BAND_ROWS = 2 * OUT - 1;               // source rows 35 .. 193
constexpr int BAND_BYTES = BAND_ROWS * ROW_BYTES;    // 76320
constexpr int ITEMS = OUT * 10;                      // 800 work items of 48 B -> 8 px
constexpr int PREP_THREADS = 256;

// 48 input bytes (12 words, w[2], w[5], w[8], w[11] unused) -> 8 output bytes.
__device__ __forceinline__ uint2 classify48(const uint4 a, const uint4 b, const uint4 c) {
  // pixels at byte offsets 0,6,12,18 | 24,30,36,42
  // words:            a.x b0, a.y b2, a.w b0, b.x b2 | b.z b0, b.w b2, c.y b0, c.z b2
  uint32_t lo = __byte_perm(__byte_perm(a.x, a.y, 0x0060), __byte_perm(a.w, b.x, 0x0060), 0x5410);
  uint32_t hi = __byte_perm(__byte_perm(b.z, b.w, 0x0060), __byte_perm(c.y, c.z, 0x0060), 0x5410);
  uint32_t bg_lo = __vcmpeq4(lo, 0u) | __vcmpeq4(lo, 0x90909090u) | __vcmpeq4(lo, 0x6d6d6d6du);
  uint32_t bg_hi = __vcmpeq4(hi, 0u) | __vcmpeq4(hi, 0x90909090u) | __vcmpeq4(hi, 0x6d6d6d6du);
  return make_uint2(~bg_lo & 0x01010101u, ~bg_hi & 0x01010101u);
}

// STAGES x CTAS: 4-deep ring with one CTA per SM, or 2-deep rings with two CTAs per SM (the
// second CTA's loads and classification fill the first one's barrier / refill bubbles).
template <int STAGES, int CTAS>
__global__ void __launch_bounds__(PREP_THREADS, CTAS)
obs_prep_tma_kernel(const __grid_constant__ CUtensorMap tmap, const uint8_t* __restrict__ reset,
                    uint8_t* __restrict__ planes, uint8_t* __restrict__ pflag, int NF, int F,
                    int P, int j0) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * SEL_BYTES);
  const int tid = threadIdx.x;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));

  auto issue = [&](int frame, int stage) {
    mbar_expect_tx(&full[stage], SEL_BYTES);
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
        ".L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;" ::"r"(
            smem_u32(smem + stage * SEL_BYTES)),
        "l"(&tmap), "r"(0), "r"(0), "r"(frame), "r"(smem_u32(&full[stage])), "l"(policy)
        : "memory");
  };

  const int first = blockIdx.x, step = gridDim.x;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      int fr = first + s * step;
      if (fr < NF) issue(fr, s);
    }
  }
  int k = 0;
  for (int fr = first; fr < NF; fr += step, ++k) {
    const int stage = k % STAGES;
    const uint32_t parity = (k / STAGES) & 1;
    mbar_wait(&full[stage], parity);
    const uint8_t* src = smem + stage * SEL_BYTES;
    const int r = fr / F, f = fr - r * F;
    uint8_t* dst = planes + ((long long)r * P + j0 + f) * PLANE_BYTES;
#pragma unroll 2
    for (int it = tid; it < ITEMS; it += PREP_THREADS) {
      const uint4* p = reinterpret_cast<const uint4*>(src + it * 48);
      uint2 o = classify48(p[0], p[1], p[2]);
      *reinterpret_cast<uint2*>(dst + it * 8) = o;
    }
    if (tid == 0) pflag[(long long)r * P + j0 + f] = reset[fr] ? 1 : 0;
    __syncthreads();   // every thread is done with this stage
    if (tid == 0) {
      int nx = fr + STAGES * step;
      if (nx < NF) issue(nx, stage);
    }
  }
}

__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

__global__ void __launch_bounds__(256)
obs_prep_ldg_kernel(const uint8_t* __restrict__ frames, const uint8_t* __restrict__ reset,
                    uint8_t* __restrict__ planes, uint8_t* __restrict__ pflag, int NF, int F,
                    int P, int j0) {
  const long long total = (long long)NF * ITEMS;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    const int fr = (int)(g / ITEMS), it = (int)(g - (long long)fr * ITEMS);
    const int row = it / 10, seg = it - row * 10;
    const uint4* p = reinterpret_cast<const uint4*>(
        frames + (long long)fr * FRAME_BYTES + (CROP_TOP + 2 * row) * ROW_BYTES + seg * 48);
    uint4 a = ldg_stream(p), b = ldg_stream(p + 1), c = ldg_stream(p + 2);
    const int r = fr / F, f = fr - r * F;
    uint8_t* dst = planes + ((long long)r * P + j0 + f) * PLANE_BYTES;
    *reinterpret_cast<uint2*>(dst + it * 8) = classify48(a, b, c);
    if (it == 0) pflag[(long long)r * P + j0 + f] = reset[fr] ? 1 : 0;
  }
}

__global__ void ring_carry_kernel(uint8_t* __restrict__ planes, uint8_t* __restrict__ pflag, int R,
                                  int P, int fs) {
  // one block per (slot, carried plane); source index P-fs+c -> destination c
  const int r = blockIdx.x / fs, c = blockIdx.x % fs;
  const uint4* src = reinterpret_cast<const uint4*>(planes + ((long long)r * P + (P - fs + c)) * PLANE_BYTES);
  uint4* dst = reinterpret_cast<uint4*>(planes + ((long long)r * P + c) * PLANE_BYTES);
  for (int i = threadIdx.x; i < PLANE_BYTES / 16; i += blockDim.x) dst[i] = src[i];
  if (threadIdx.x == 0) pflag[(long long)r * P + c] = pflag[(long long)r * P + (P - fs + c)];
}

// One thread per 4 output pixels of one (n, c) plane.
template <typename OutT>
__global__ void __launch_bounds__(256)
materialize_kernel(a2c_state_view v, OutT* __restrict__ out, long long N) {
  const int fs = v.fs;
  const long long total = N * fs * (PLANE_BYTES / 4);
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(g % (PLANE_BYTES / 4));
    const long long nc = g / (PLANE_BYTES / 4);
    const int c = (int)(nc % fs);
    const long long n = nc / fs;
    const long long r = n / v.T;
    const int t = (int)(n - r * v.T);
    bool valid = true;
    if (v.pflag) {
      const uint8_t* pf = v.pflag + r * v.pflag_stride_r + t;
      for (int j = c + 1; j < fs; ++j) valid = valid && (pf[j] == 0);
    }
    const long long off = r * v.stride_r + (long long)t * v.stride_t + (long long)c * v.stride_c + q * 4;
    float x0 = 0, x1 = 0, x2 = 0, x3 = 0;
    if (valid) {
      if (v.dtype == 0) {
        const uint32_t w = *reinterpret_cast<const uint32_t*>(static_cast<const uint8_t*>(v.base) + off);
        x0 = w & 0xff; x1 = (w >> 8) & 0xff; x2 = (w >> 16) & 0xff; x3 = w >> 24;
      } else {
        const float4 w = *reinterpret_cast<const float4*>(static_cast<const float*>(v.base) + off);
        x0 = w.x; x1 = w.y; x2 = w.z; x3 = w.w;
      }
    }
    if constexpr (sizeof(OutT) == 1) {
      uint32_t w = (uint32_t)x0 | ((uint32_t)x1 << 8) | ((uint32_t)x2 << 16) | ((uint32_t)x3 << 24);
      *reinterpret_cast<uint32_t*>(out + g * 4) = w;
    } else {
      *reinterpret_cast<float4*>(out + g * 4) = make_float4(x0, x1, x2, x3);
    }
  }
}

void* tensormap_encoder();   // shared with the conv patch kernels
static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

void* tensormap_encoder() { return reinterpret_cast<void*>(get_encode()); }

template <int STAGES, int CTAS>
static int launch_prep_tma(const CUtensorMap& tmap, const uint8_t* reset, uint8_t* planes, uint8_t* pflag,
                           int NF, int F, int P, int j0, cudaStream_t st) {
  const int smem = STAGES * SEL_BYTES + STAGES * 8;
  auto kern = obs_prep_tma_kernel<STAGES, CTAS>;
  A2C_CUDA_CHECK(set_max_dynamic_smem(reinterpret_cast<const void*>(kern), smem));
  const int cap = num_sms() * CTAS;
  const int grid = NF < cap ? NF : cap;
  kern<<<grid, PREP_THREADS, smem, st>>>(tmap, reset, planes, pflag, NF, F, P, j0);
  return 0;
}

// Source layouts: packed == 0: raw frames u8 [NF, 210, 160, 3] (the TMA row pitch of 960 B skips
// the even rows); packed == 1: the 80 selected rows only, u8 [NF, 80, 480]; packed == 2: the
// cropped band, source rows 35..193, u8 [NF, 159, 480] (the staging layouts of
// a2c_obs_ingest_rows_async).
static int launch_prep(const uint8_t* frames, const uint8_t* reset, uint8_t* planes, uint8_t* pflag,
                       int R, int F, int P, int j0, int variant, int packed, cudaStream_t st) {
  A2C_ARG_CHECK(frames && reset && planes && pflag, "obs_prep: null pointer");
  A2C_ARG_CHECK(R >= 0 && F >= 0 && P > 0 && j0 >= 0 && j0 + F <= P,
                "obs_prep: bad sizes R=%d F=%d P=%d j0=%d", R, F, P, j0);
  A2C_ARG_CHECK((reinterpret_cast<uintptr_t>(frames) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(planes) & 15) == 0,
                "obs_prep: frames/planes must be 16-byte aligned");
  const long long nf_ll = (long long)R * F;
  A2C_ARG_CHECK(nf_ll < (1ll << 31), "obs_prep: too many frames");
  const int NF = (int)nf_ll;
  if (NF == 0) return 0;
  if (variant == 0 || variant == 2) {
    auto encode = get_encode();
    if (!encode) {
      set_error("obs_prep: cuTensorMapEncodeTiled unavailable");
      return (int)cudaErrorNotSupported;
    }
    CUtensorMap tmap;
    cuuint64_t dims[3] = {ROW_BYTES / 4, OUT, (cuuint64_t)NF};
    cuuint64_t strides[2] = {2 * ROW_BYTES, FRAME_BYTES};   // bytes, dims 1..2
    if (packed == 1) { strides[0] = ROW_BYTES; strides[1] = SEL_BYTES; }
    if (packed == 2) strides[1] = BAND_BYTES;
    cuuint32_t box[3] = {ROW_BYTES / 4, OUT, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3,
                         const_cast<uint8_t*>(frames) + (packed ? 0 : CROP_TOP * ROW_BYTES), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
      set_error("obs_prep: cuTensorMapEncodeTiled failed (%d)", (int)cr);
      return (int)cudaErrorInvalidValue;
    }
    // measured on B200 (5,120 frames): 2 CTAs/SM x 2 stages 39.6 us (5.79 TB/s of algorithmic bytes),
    // 1 CTA/SM x 4 stages 44.1 us (5.20 TB/s)
    int rc = variant == 2 ? launch_prep_tma<4, 1>(tmap, reset, planes, pflag, NF, F, P, j0, st)
                          : launch_prep_tma<2, 2>(tmap, reset, planes, pflag, NF, F, P, j0, st);
    if (rc) return rc;
  } else {
    A2C_ARG_CHECK(!packed, "obs_prep: the direct-load variant reads raw frames only");
    long long blocks = cdivll((long long)NF * ITEMS, 256);
    const long long cap = (long long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    obs_prep_ldg_kernel<<<(int)blocks, 256, 0, st>>>(frames, reset, planes, pflag, NF, F, P, j0);
  }
  g_launches++;
  A2C_LAUNCH_CHECK();
  return 0;
}

}  // namespace a2c

using namespace a2c;

extern "C" {

int a2c_abi_version(void) { return A2C_ABI_VERSION; }
const char* a2c_last_error(void) { return a2c::g_err; }
long long a2c_launch_count(void) { return a2c::g_launches.load(); }

int a2c_obs_prep_stack(const uint8_t* frames, const uint8_t* reset, uint8_t* planes, uint8_t* pflag,
                       int R, int F, int P, int j0, int variant, void* stream) {
  return launch_prep(frames, reset, planes, pflag, R, F, P, j0, variant, 0, as_stream(stream));
}

int a2c_obs_prep_stack_rows(const uint8_t* rows, const uint8_t* reset, uint8_t* planes, uint8_t* pflag,
                            int R, int F, int P, int j0, int layout, void* stream) {
  A2C_ARG_CHECK(layout >= 0 && layout <= 2, "obs_prep_rows: layout must be 0, 1 or 2");
  return launch_prep(rows, reset, planes, pflag, R, F, P, j0, 0, layout == 0 ? 1 : (layout == 1 ? 2 : 0),
                     as_stream(stream));
}

long long a2c_obs_ingest_bytes_per_frame(int layout) {
  return layout == 0 ? SEL_BYTES : (layout == 1 ? BAND_BYTES : FRAME_BYTES);
}

int a2c_obs_ingest_rows_async(const uint8_t* frames_host, uint8_t* rows, long long n_frames, int layout,
                              void* stream) {
  A2C_ARG_CHECK(frames_host && rows && n_frames >= 0 && layout >= 0 && layout <= 2, "obs_ingest: bad args");
  if (n_frames == 0) return 0;
  if (layout == 2) {                       // whole frames, one contiguous copy
    A2C_CUDA_CHECK(cudaMemcpyAsync(rows, frames_host, (size_t)n_frames * FRAME_BYTES, cudaMemcpyHostToDevice,
                                   as_stream(stream)));
    return 0;
  }
  if (layout == 1) {                       // the cropped band of every frame: 76,320 contiguous bytes each
    A2C_CUDA_CHECK(cudaMemcpy2DAsync(rows, BAND_BYTES, frames_host + CROP_TOP * ROW_BYTES, FRAME_BYTES, BAND_BYTES,
                                     (size_t)n_frames, cudaMemcpyHostToDevice, as_stream(stream)));
    return 0;
  }
  // layout 0: source seen as [n_frames][105 double rows][960 B], of which double rows 17..96,
  // bytes 480..959 (= source rows 35, 37, ..., 193) are copied: 38,400 B per frame.
  cudaMemcpy3DParms p;
  memset(&p, 0, sizeof(p));
  p.srcPtr = make_cudaPitchedPtr(const_cast<uint8_t*>(frames_host), 2 * ROW_BYTES, 2 * ROW_BYTES, FRAME_H / 2);
  p.srcPos = make_cudaPos(ROW_BYTES, CROP_TOP / 2, 0);
  p.dstPtr = make_cudaPitchedPtr(rows, ROW_BYTES, ROW_BYTES, OUT);
  p.dstPos = make_cudaPos(0, 0, 0);
  p.extent = make_cudaExtent(ROW_BYTES, OUT, (size_t)n_frames);
  p.kind = cudaMemcpyHostToDevice;
  A2C_CUDA_CHECK(cudaMemcpy3DAsync(&p, as_stream(stream)));
  return 0;
}

int a2c_obs_prep_stack_host(const uint8_t* frames_host, const uint8_t* reset_host, uint8_t* planes,
                            uint8_t* pflag, int R, int F, int P, int j0, void* stream) {
  // Zero-copy: the TMA boxes are fetched straight from pinned, mapped host
  // memory, so only the 80 selected rows of each frame (38 % of its bytes)
  // cross PCIe.  The buffers must be page-locked (cudaHostAlloc/HostRegister).
  cudaPointerAttributes at;
  A2C_CUDA_CHECK(cudaPointerGetAttributes(&at, frames_host));
  A2C_ARG_CHECK(at.type == cudaMemoryTypeHost && at.devicePointer != nullptr,
                "obs_prep_host: frames must be pinned (page-locked, mapped) host memory");
  cudaPointerAttributes at2;
  A2C_CUDA_CHECK(cudaPointerGetAttributes(&at2, reset_host));
  A2C_ARG_CHECK(at2.type == cudaMemoryTypeHost && at2.devicePointer != nullptr,
                "obs_prep_host: reset must be pinned (page-locked, mapped) host memory");
  int rc = launch_prep(static_cast<const uint8_t*>(at.devicePointer),
                       static_cast<const uint8_t*>(at2.devicePointer), planes, pflag, R, F, P, j0, 0, 0,
                       as_stream(stream));
  if (rc) return rc;
  A2C_CUDA_CHECK(cudaStreamSynchronize(as_stream(stream)));
  return 0;
}

int a2c_ring_carry(uint8_t* planes, uint8_t* pflag, int R, int P, int fs, void* stream) {
  A2C_ARG_CHECK(planes && pflag && R >= 0 && fs > 0 && P >= 2 * fs,
                "ring_carry: bad args R=%d P=%d fs=%d (need P >= 2*fs)", R, P, fs);
  if (R == 0) return 0;
  ring_carry_kernel<<<R * fs, 128, 0, as_stream(stream)>>>(planes, pflag, R, P, fs);
  g_launches++;
  A2C_LAUNCH_CHECK();
  return 0;
}

int a2c_states_materialize(const a2c_state_view* view, void* out, int out_dtype, int R, void* stream) {
  A2C_ARG_CHECK(view && view->base && out, "materialize: null pointer");
  A2C_ARG_CHECK(view->fs > 0 && view->T > 0 && R >= 0, "materialize: bad sizes");
  A2C_ARG_CHECK((view->stride_r % 4) == 0 && (view->stride_t % 4) == 0 && (view->stride_c % 4) == 0,
                "materialize: strides must be multiples of 4 elements");
  const long long N = (long long)R * view->T;
  if (N == 0) return 0;
  const long long total = N * view->fs * (PLANE_BYTES / 4);
  long long blocks = cdivll(total, 256);
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (out_dtype == 0)
    materialize_kernel<uint8_t><<<(int)blocks, 256, 0, as_stream(stream)>>>(*view, static_cast<uint8_t*>(out), N);
  else
    materialize_kernel<float><<<(int)blocks, 256, 0, as_stream(stream)>>>(*view, static_cast<float*>(out), N);
  g_launches++;
  A2C_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"

// tcgen05 (5th-gen tensor core) GEMM for the dense contractions of stage 4:
// C[M,N] (+)= A[M,K] * B[K,N] + bias, fp32 in / fp32 out.
//
//   precision 1: 3xTF32 -- every operand is split hi = rna_tf32(x), lo = x - hi in
//                the producer warps; D += A_lo*B_hi + A_hi*B_lo + A_hi*B_hi with fp32
//                accumulation in TMEM (the fp32-parity mode on tensor cores);
//   precision 2: single-pass TF32 (10-bit mantissa operands; stated tolerance 2e-3).
//
// One CTA per 128 x BN output tile.  Warp roles:
//   warps 0-3   producers of the A tile
//   warps 4-7   producers of the B tile
//   warps 8-15  accumulator drain + epilogue (TMEM -> registers -> global)
//   warp  16    MMA issuer (one elected lane issues tcgen05.mma, commits to mbarriers)
//   warp  17    TMEM allocator
// The TMEM accumulator adds with truncation (measured: error grows linearly with
// the number of chained MMAs, 2e-4 at K = 25600), so K is processed in chunks of
// KC k-blocks into two ping-pong TMEM accumulators; the drain warps add every
// finished chunk into fp32 registers with round-to-nearest while the tensor core
// works on the next chunk.  That keeps 3xTF32 within 1e-5 of fp64 at any K.
// Operand tiles live in shared memory in the canonical K-major SWIZZLE_128B UMMA
// layout (rows of 32 fp32 = 128 B, 16-byte chunk index XOR (row & 7), 1024-B
// aligned 8-row groups).  Producers load global memory with 128-bit accesses
// along whichever dimension is contiguous (transposing in registers when the
// operand is MN-contiguous), so no operand ever needs a transposed copy in HBM.
// A 2..4-stage ring of {full, empty} mbarriers pipelines producers and MMA.
#include <atomic>

#include "common.cuh"

namespace a2c {
extern std::atomic<long long> g_launches;

namespace tc {

constexpr int TM = 128;            // UMMA M (cta_group::1)
constexpr int KB = 32;             // k-block: 32 fp32 = one 128-byte swizzle row
constexpr int PROD_THREADS = 128;  // per operand
constexpr int THREADS = 576;       // 18 warps
constexpr int NUM_PROD_WARPS = 8;
constexpr int NUM_EPI_WARPS = 8;
constexpr int KC = 8;              // k-blocks per TMEM accumulation chunk (256 k)

struct Operand {
  const float* p;
  long long s_mn, s_k;
  int mode;   // 0: k contiguous, 1: mn contiguous
};

struct Params {
  Operand A, B;
  float* C;
  long long ldc;
  const float* bias;
  int M, N, K;
  int relu, beta;
  int stages;
};

__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);        // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                             // leading byte offset (unused for SW128 K-major) = 1
  d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                             // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
  return d;
}

// kind::tf32, fp32 accumulate, both operands K-major (cute::UMMA::InstrDescriptor)
__host__ __device__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// Stores one 16-byte chunk (4 consecutive k of one row) into the swizzled tile(s).
template <bool SPLIT3>
__device__ __forceinline__ void store_chunk(uint8_t* hi_tile, uint8_t* lo_tile, int row, int chunk,
                                            float4 v) {
  const int off = row * 128 + ((chunk ^ (row & 7)) << 4);
  if constexpr (SPLIT3) {
    float4 h = make_float4(tf32_rna(v.x), tf32_rna(v.y), tf32_rna(v.z), tf32_rna(v.w));
    *reinterpret_cast<float4*>(hi_tile + off) = h;
    *reinterpret_cast<float4*>(lo_tile + off) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
  } else {
    *reinterpret_cast<float4*>(hi_tile + off) = v;
  }
}

// Fills a ROWS x 32 tile from global memory.  t = thread index within the operand's producers.
template <int ROWS, bool SPLIT3>
__device__ __forceinline__ void produce_tile(const Operand& op, int MN, int K, int mn0, int k0,
                                             uint8_t* hi_tile, uint8_t* lo_tile, int t) {
  if (op.mode == 0) {
    // k contiguous: unit = (row, 16-byte chunk); a quarter-warp covers one row's 128 bytes
    constexpr int ITERS = ROWS * 8 / PROD_THREADS;
    float4 v[ITERS];
#pragma unroll
    for (int i = 0; i < ITERS; ++i) {
      const int u = i * PROD_THREADS + t, row = u >> 3, chunk = u & 7;
      const int mn = mn0 + row, k = k0 + chunk * 4;
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (mn < MN && k < K) v[i] = __ldg(reinterpret_cast<const float4*>(op.p + (long long)mn * op.s_mn + k));
    }
#pragma unroll
    for (int i = 0; i < ITERS; ++i) {
      const int u = i * PROD_THREADS + t;
      store_chunk<SPLIT3>(hi_tile, lo_tile, u >> 3, u & 7, v[i]);
    }
  } else {
    // mn contiguous: unit = (group of 4 rows, k-chunk); 4 loads along mn, 4x4 register transpose
    constexpr int UNITS = (ROWS / 4) * 8;
    constexpr int ITERS = (UNITS + PROD_THREADS - 1) / PROD_THREADS;
#pragma unroll
    for (int i = 0; i < ITERS; ++i) {
      const int g = i * (PROD_THREADS / 8) + (t >> 3), chunk = t & 7;
      if (g * 4 >= ROWS) break;
      const int mn = mn0 + g * 4;
      float4 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = k0 + chunk * 4 + j;
        v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (mn < MN && k < K) v[j] = __ldg(reinterpret_cast<const float4*>(op.p + (long long)k * op.s_k + mn));
      }
      store_chunk<SPLIT3>(hi_tile, lo_tile, g * 4 + 0, chunk, make_float4(v[0].x, v[1].x, v[2].x, v[3].x));
      store_chunk<SPLIT3>(hi_tile, lo_tile, g * 4 + 1, chunk, make_float4(v[0].y, v[1].y, v[2].y, v[3].y));
      store_chunk<SPLIT3>(hi_tile, lo_tile, g * 4 + 2, chunk, make_float4(v[0].z, v[1].z, v[2].z, v[3].z));
      store_chunk<SPLIT3>(hi_tile, lo_tile, g * 4 + 3, chunk, make_float4(v[0].w, v[1].w, v[2].w, v[3].w));
    }
  }
}

template <int BN, bool SPLIT3>
__global__ void __launch_bounds__(THREADS, 1) gemm_tc_kernel(const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int A_BYTES = TM * 128, B_BYTES = BN * 128;
  constexpr int STAGE_BYTES = (SPLIT3 ? 2 : 1) * (A_BYTES + B_BYTES);
  constexpr int HALF = BN / 2;      // columns per drain warp
  // 1024-byte alignment of every tile is required by SWIZZLE_128B
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + p.stages * STAGE_BYTES);
  uint64_t* empty = full + p.stages;
  uint64_t* acc_full = empty + p.stages;     // [2]
  uint64_t* acc_empty = acc_full + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * BN;
  const int nkb = (p.K + KB - 1) / KB;
  const int nchunks = (nkb + KC - 1) / KC;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], NUM_PROD_WARPS);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], NUM_EPI_WARPS);
    }
    mbar_fence_init();
  }
  if (warp == 17) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"((uint32_t)(2 * BN)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto stage_ptr = [&](int s) { return smem + s * STAGE_BYTES; };
  // stage layout: [A_hi][A_lo?][B_hi][B_lo?]
  if (warp < NUM_PROD_WARPS) {
    const bool isA = warp < 4;
    const int t = threadIdx.x & (PROD_THREADS - 1);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % p.stages;
      const uint32_t parity = (kb / p.stages) & 1;
      mbar_wait(&empty[s], parity ^ 1);
      uint8_t* st = stage_ptr(s);
      if (isA) {
        produce_tile<TM, SPLIT3>(p.A, p.M, p.K, m0, kb * KB, st, st + A_BYTES, t);
      } else {
        uint8_t* b = st + (SPLIT3 ? 2 : 1) * A_BYTES;
        produce_tile<BN, SPLIT3>(p.B, p.N, p.K, n0, kb * KB, b, b + B_BYTES, t);
      }
      fence_proxy_async();      // generic-proxy stores -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[s]);
    }
  } else if (warp == 16) {
    constexpr uint32_t idesc = make_idesc(BN);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % p.stages;
      const uint32_t parity = (kb / p.stages) & 1;
      const int chunk = kb / KC, buf = chunk & 1;
      if (kb % KC == 0) {                       // new chunk: its TMEM buffer must be drained
        mbar_wait(&acc_empty[buf], ((chunk >> 1) & 1) ^ 1);
        tc_fence_after();
      }
      mbar_wait(&full[s], parity);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t tmem_d = tmem_base + buf * BN;
        const uint32_t a_hi = smem_u32(stage_ptr(s));
        const uint32_t a_lo = a_hi + A_BYTES;
        const uint32_t b_hi = a_hi + (SPLIT3 ? 2 : 1) * A_BYTES;
        const uint32_t b_lo = b_hi + B_BYTES;
#pragma unroll
        for (int kk = 0; kk < KB / 8; ++kk) {
          const uint32_t koff = kk * 32;     // 8 tf32 = 32 bytes inside the 128-byte swizzle row
          const uint32_t first = ((kb % KC) | kk) ? 1u : 0u;
          if constexpr (SPLIT3) {
            mma_tf32(tmem_d, make_desc(a_lo + koff), make_desc(b_hi + koff), idesc, first);
            mma_tf32(tmem_d, make_desc(a_hi + koff), make_desc(b_lo + koff), idesc, 1u);
            mma_tf32(tmem_d, make_desc(a_hi + koff), make_desc(b_hi + koff), idesc, 1u);
          } else {
            mma_tf32(tmem_d, make_desc(a_hi + koff), make_desc(b_hi + koff), idesc, first);
          }
        }
        mma_commit(&empty[s]);                                   // frees the stage when these MMAs retire
        if (kb % KC == KC - 1 || kb == nkb - 1) mma_commit(&acc_full[buf]);   // chunk complete
      }
      __syncwarp();
    }
  } else if (warp >= 8 && warp < 8 + NUM_EPI_WARPS) {
    // ---- drain + epilogue: thread <-> accumulator row (TMEM lane), half of the columns ----
    const int e = warp - 8, q = e & 3, hsel = e >> 2;
    const int row = m0 + q * 32 + lane;
    float acc[HALF];
#pragma unroll
    for (int j = 0; j < HALF; ++j) acc[j] = 0.0f;
    for (int chunk = 0; chunk < nchunks; ++chunk) {
      const int buf = chunk & 1;
      mbar_wait(&acc_full[buf], (chunk >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + hsel * HALF;
#pragma unroll
      for (int c0 = 0; c0 < HALF; c0 += 16) {
        uint32_t r[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
              "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
              "=r"(r[14]), "=r"(r[15])
            : "r"(taddr + (uint32_t)c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[c0 + j] += __uint_as_float(r[j]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
    if (row < p.M) {
      const int nb = n0 + hsel * HALF;
      float* dst = p.C + (long long)row * p.ldc + nb;
      const bool vec = ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) && (nb + HALF <= p.N);
#pragma unroll
      for (int j = 0; j < HALF; j += 4) {
        float x[4];
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          const int n = nb + j + qq;
          float val = acc[j + qq];
          if (n < p.N) {
            if (p.bias) val += p.bias[n];
            if (p.beta) val += dst[j + qq];
            if (p.relu) val = fmaxf(val, 0.0f);
          }
          x[qq] = val;
        }
        if (vec) {
          *reinterpret_cast<float4*>(dst + j) = make_float4(x[0], x[1], x[2], x[3]);
        } else {
#pragma unroll
          for (int qq = 0; qq < 4; ++qq)
            if (nb + j + qq < p.N) dst[j + qq] = x[qq];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 17) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)(2 * BN)));
  }
}

template <int BN, bool SPLIT3>
static int launch(Params p, cudaStream_t st) {
  constexpr int STAGE_BYTES = (SPLIT3 ? 2 : 1) * (TM * 128 + BN * 128);
  int stages = (200 * 1024) / STAGE_BYTES;
  if (stages > 4) stages = 4;
  if (stages < 2) return 1;
  p.stages = stages;
  const size_t smem = (size_t)stages * STAGE_BYTES + 1024 /*align slack*/ + (2 * stages + 4) * 8 + 16;
  auto kern = gemm_tc_kernel<BN, SPLIT3>;
  static bool configured = false;
  if (!configured) {
    A2C_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  dim3 grid(cdiv(p.N, BN), cdiv(p.M, TM));
  kern<<<grid, THREADS, smem, st>>>(p);
  g_launches++;
  A2C_LAUNCH_CHECK();
  return 0;
}

static bool classify(const float* ptr, long long s_mn, long long s_k, int MN, int K, Operand* op) {
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return false;
  op->p = ptr; op->s_mn = s_mn; op->s_k = s_k;
  if (s_k == 1 && (s_mn % 4) == 0 && (K % 4) == 0) { op->mode = 0; return true; }
  if (s_mn == 1 && (s_k % 4) == 0 && (MN % 4) == 0) { op->mode = 1; return true; }
  return false;
}

}  // namespace tc

// Returns 0 when the GEMM ran on the tensor cores, 1 when the shape does not
// qualify (caller falls back to the fp32 FFMA kernel), anything else = error.
int gemm_tc_try(const float* A, long long sam, long long sak, const float* B, long long sbk,
                long long sbn, float* C, long long ldc, const float* bias, int M, int N, int K,
                int relu, int beta, int precision, void* workspace, long long workspace_bytes,
                cudaStream_t st) {
  (void)workspace; (void)workspace_bytes;
  // tiny problems are launch-latency bound and gain nothing from a 128-row MMA tile
  if (M < 64 || K < 64 || (long long)M * N * K < (1ll << 22)) return 1;
  tc::Params p{};
  if (!tc::classify(A, sam, sak, M, K, &p.A)) return 1;
  if (!tc::classify(B, sbn, sbk, N, K, &p.B)) return 1;
  p.C = C; p.ldc = ldc; p.bias = bias; p.M = M; p.N = N; p.K = K; p.relu = relu; p.beta = beta;
  const bool split3 = (precision == 1);
  // widest tile that does not leave most of the MMA idle
  int bn = N > 64 ? 128 : (N > 32 ? 64 : 32);
  // prefer more CTAs when the grid would not fill the machine
  const int sms = num_sms();
  while (bn > 64 && (long long)cdiv(N, bn) * cdiv(M, tc::TM) < sms) bn >>= 1;
  if (split3) {
    switch (bn) {
      case 128: return tc::launch<128, true>(p, st);
      case 64: return tc::launch<64, true>(p, st);
      default: return tc::launch<32, true>(p, st);
    }
  }
  switch (bn) {
    case 128: return tc::launch<128, false>(p, st);
    case 64: return tc::launch<64, false>(p, st);
    default: return tc::launch<32, false>(p, st);
  }
}

}  // namespace a2c

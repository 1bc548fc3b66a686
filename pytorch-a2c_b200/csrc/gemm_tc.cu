// tcgen05 GEMM for the dense contractions of stage 4 (nn.Linear forward /
// dgrad / wgrad and the GRU gate products): C[M,N] (+)= A[M,K] * B[K,N] + bias,
// fp32 in / fp32 out, on the tile engine of tc_core.cuh.
//
//   precision 1: 3xTF32 (fp32-parity mode: 1e-5 of fp64 at any K);
//   precision 2: single-pass TF32 (10-bit mantissa operands; stated tolerance 2e-3).
//
// Producers load global memory with 128-bit accesses along whichever dimension
// of the operand is contiguous and transpose in registers when that is the MN
// dimension, so dgrad / wgrad never need a transposed copy of a matrix in HBM.
#include <atomic>

#include "tc_core.cuh"

namespace a2c {
extern std::atomic<long long> g_launches;

namespace tc {

struct GemmOperand {
  const float* p;
  long long s_mn, s_k;
  int mode;   // 0: k contiguous, 1: mn contiguous
  int dbg;
};

__device__ __forceinline__ float4 gemm_ld4(const GemmOperand& op, const float* a) {
#ifdef A2C_DEBUG_KNOBS
  float4 v;
  if ((op.dbg & 3) == 1) {
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(a));
    return v;
  } else if ((op.dbg & 3) == 2) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(a));
    return v;
  } else if ((op.dbg & 3) == 3) {
    asm volatile("ld.global.cs.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(a));
    return v;
  }
  if (op.dbg & 8) return make_float4(1.f, 2.f, 3.f, 4.f);
#endif
  return __ldg(reinterpret_cast<const float4*>(a));
}

struct GemmParams {
  GemmOperand A, B;
  float* C;
  long long ldc;
  const float* bias;
  int M, N, K;
  int relu, beta;
  int nkb, kb_per_split;
  float* partial;     // split-K: [splits][M][N] partial products (bias / beta / relu applied by the reduction)
  const float* cmask; // optional, laid out like C: C = cmask > 0 ? value : 0 (ReLU of the producer of C's forward twin)
};

constexpr int PROD_THREADS = 256;   // 8 producer warps per operand

template <int ROWS>
struct GemmFrag {
  float4 v[4];             // ROWS (<= 128) x 32 fp32 spread over 256 producer threads
};

// Transposing path: thread t <-> (group of 4 mn-rows, 16-byte k-chunk).  Across a warp the 8 groups
// are consecutive (a warp load covers 4 k-rows x 128 contiguous bytes); inside every quarter-warp
// (the unit in which a 128-bit shared store is checked for bank conflicts) the lanes take 2 group
// parities x 4 chunks, which the 128B swizzle (chunk ^ row & 7, row = 4*group + q) maps to 8
// different 16-byte bank groups.  (ncu: 74 % of the shared wavefronts of the wgrad GEMM were
// conflicts with 8 groups x 1 chunk per quarter-warp.)
__device__ __forceinline__ int mn_group(int t) { return (((t >> 3) & 3) << 1) | (t & 1) | (((t >> 5) & 3) << 3); }
__device__ __forceinline__ int k_chunk(int t) { return ((t >> 1) & 3) | ((t >> 7) << 2); }

// Per-thread producer state of one operand for the current tile: everything that does not depend on
// the k-block is computed once per tile (the producers are issue-bound: with the row decode, bounds
// checks and 64-bit address arithmetic inside the k loop every 128-bit load cost ~40 instructions).
struct GemmRowState {
  const char* base;    // first byte this thread loads at k-block 0
  long long step;      // BYTES between the thread's four loads: mode 0: 32 rows; mode 1: one k-row
  int kfirst;          // k of the thread's first element inside a k-block
  int valid;           // mode 0: bit i = unit i is a row of the matrix; mode 1: the 4-row group is (0 / 15)
  int soff;            // byte offset of the thread's first store inside the swizzled tile (tile-independent)
};

template <int ROWS>
__device__ __forceinline__ void gemm_setup(const GemmOperand& op, int MN, int mn0, int t, GemmRowState& st) {
  if (op.mode == 0) {
    // k contiguous: unit i = (row (t >> 3) + 32 i, 16-byte chunk t & 7); a quarter-warp covers one row's 128 bytes
    const int row = t >> 3, chunk = t & 7;
    st.kfirst = chunk * 4;
    st.base = reinterpret_cast<const char*>(op.p + (long long)(mn0 + row) * op.s_mn + st.kfirst);
    st.step = 32 * op.s_mn * (long long)sizeof(float);
    st.valid = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (row + 32 * i < ROWS && mn0 + row + 32 * i < MN) st.valid |= 1 << i;
    st.soff = swz(row, chunk);           // unit i: + 4096 i (32 rows; row & 7 does not change)
  } else {
    // mn contiguous: unit = (group of 4 rows, k-chunk): 4 loads along mn (transposed at store time).
    // 8 consecutive lanes take 8 consecutive groups, so a warp load covers 4 k-rows x 128 contiguous bytes
    const int g = mn_group(t), chunk = k_chunk(t);
    st.kfirst = chunk * 4;
    st.base = reinterpret_cast<const char*>(op.p + (long long)st.kfirst * op.s_k + mn0 + g * 4);
    st.step = op.s_k * (long long)sizeof(float);
    st.valid = (g * 4 < ROWS && mn0 + g * 4 < MN) ? 15 : 0;
    // row g*4 + q, chunk ^ (row & 7):  (row & 7) = ((g & 1) << 2) | q, so store q sits at
    // ((soff ^ (q << 4)) + 128 q)
    st.soff = g * 512 + ((chunk ^ ((g & 1) << 2)) << 4);
  }
}

// Load phase: global -> registers.  Byte pointers, one 64-bit add per load, and a fast path without
// predicates for the interior of the matrix: about 15 instructions for the four 128-bit loads.
template <int ROWS>
__device__ __forceinline__ void gemm_load(const GemmOperand& op, const GemmRowState& st, int K, int k0,
                                          GemmFrag<ROWS>& f) {
  // rows of k this thread may touch in this k-block (mode 0 needs one: its chunk; mode 1 needs four)
  const int krem = K - k0 - st.kfirst;
  const char* src = st.base + (op.mode == 0 ? (long long)k0 * (long long)sizeof(float) : (long long)k0 * st.step);
  if (st.valid == 15 && krem >= 4) {
    const char* s1 = src + st.step;
    const char* s2 = s1 + st.step;
    const char* s3 = s2 + st.step;
    f.v[0] = gemm_ld4(op, reinterpret_cast<const float*>(src));
    f.v[1] = gemm_ld4(op, reinterpret_cast<const float*>(s1));
    f.v[2] = gemm_ld4(op, reinterpret_cast<const float*>(s2));
    f.v[3] = gemm_ld4(op, reinterpret_cast<const float*>(s3));
    return;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f.v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool ok = op.mode == 0 ? (krem > 0 && ((st.valid >> i) & 1)) : (st.valid != 0 && i < krem);
    if (ok) f.v[i] = gemm_ld4(op, reinterpret_cast<const float*>(src + i * st.step));
  }
}

// Store phase: registers -> swizzled tile(s), with the hi/lo split.
template <int ROWS, bool SPLIT3>
__device__ __forceinline__ void gemm_store(const GemmOperand& op, const GemmRowState& st, const GemmFrag<ROWS>& f,
                                           uint8_t* hi, uint8_t* lo, int t) {
#ifdef A2C_DEBUG_KNOBS
  if (op.dbg & 16) {
    if (f.v[0].x == 1.2345f && f.v[1].y == 3.21f && f.v[2].x == 7.7f && f.v[3].x == 1.1f) *hi = 1;
    return;
  }
#endif
  if (op.mode == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if ((t >> 3) + 32 * i < ROWS) store_chunk_at<SPLIT3>(hi, lo, st.soff + 4096 * i, f.v[i]);
  } else {
    if (mn_group(t) * 4 < ROWS) {
      const float4* v = f.v;
      store_chunk_at<SPLIT3>(hi, lo, st.soff, make_float4(v[0].x, v[1].x, v[2].x, v[3].x));
      store_chunk_at<SPLIT3>(hi, lo, (st.soff ^ 16) + 128, make_float4(v[0].y, v[1].y, v[2].y, v[3].y));
      store_chunk_at<SPLIT3>(hi, lo, (st.soff ^ 32) + 256, make_float4(v[0].z, v[1].z, v[2].z, v[3].z));
      store_chunk_at<SPLIT3>(hi, lo, (st.soff ^ 48) + 384, make_float4(v[0].w, v[1].w, v[2].w, v[3].w));
    }
  }
}

// KCHAIN: k-blocks chained in one TMEM accumulator before the drain adds the partial result in fp32
// registers (round to nearest); TMEM adds truncate.  (Chains of 2 were tried for the weight-gradient
// GEMMs over millions of rows: 2x slower -- the drain becomes the bottleneck -- and no more accurate;
// the residual error of those gradients comes from flipped ReLU mask bits, not from accumulation.)
template <int BN, bool SPLIT3, int KCHAIN = 8>
struct GemmProblem {
  using Params = GemmParams;
  // 16 producer warps, loads two k-blocks ahead: 64 KB of operand loads in flight per SM
  static constexpr int NA = 8, NB = 8, KC = KCHAIN, DEPTH = 2;
  using AState = GemmRowState;
  using BState = GemmRowState;
  using AFrag = GemmFrag<TM>;
  using BFrag = GemmFrag<BN>;
  static __device__ __forceinline__ void init(const Params&, uint8_t*, int, int) {}
  static __device__ __forceinline__ void a_setup(const Params& p, const uint8_t*, int m0, int t, AState& st) {
    gemm_setup<TM>(p.A, p.M, m0, t, st);
  }
  static __device__ __forceinline__ void b_setup(const Params& p, const uint8_t*, int n0, int t, BState& st) {
    gemm_setup<BN>(p.B, p.N, n0, t, st);
  }
  static __device__ __forceinline__ void load_a(const Params& p, const AState& st, const uint8_t*, int, int kb,
                                                int, AFrag& f) {
    gemm_load<TM>(p.A, st, p.K, kb * KB, f);
  }
  static __device__ __forceinline__ void store_a(const Params& p, const AState& st, const AFrag& f, uint8_t* hi,
                                                 uint8_t* lo, int t) {
    gemm_store<TM, SPLIT3>(p.A, st, f, hi, lo, t);
  }
  static __device__ __forceinline__ void load_b(const Params& p, const BState& st, const uint8_t*, int, int kb,
                                                int, BFrag& f) {
    gemm_load<BN>(p.B, st, p.K, kb * KB, f);
  }
  static __device__ __forceinline__ void store_b(const Params& p, const BState& st, const BFrag& f, uint8_t* hi,
                                                 uint8_t* lo, int t) {
    gemm_store<BN, SPLIT3>(p.B, st, f, hi, lo, t);
  }
  template <int HALF>
  static __device__ __forceinline__ void epilogue(const Params& p, int row, int nb, const float* acc, int split) {
    if (row >= p.M) return;
    if (p.partial) {
      float* dst = p.partial + ((long long)split * p.M + row) * p.N + nb;
#pragma unroll
      for (int j = 0; j < HALF; ++j)
        if (nb + j < p.N) dst[j] = acc[j];
      return;
    }
    float* dst = p.C + (long long)row * p.ldc + nb;
    const bool vec = ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) && (nb + HALF <= p.N);
    const float* msk = p.cmask ? p.cmask + (long long)row * p.ldc + nb : nullptr;
    const bool mvec = msk && vec && ((reinterpret_cast<uintptr_t>(msk) & 15) == 0);
#pragma unroll
    for (int j = 0; j < HALF; j += 4) {
      float x[4];
      float mk[4] = {1.f, 1.f, 1.f, 1.f};
      if (mvec) {
        const float4 m4 = __ldg(reinterpret_cast<const float4*>(msk + j));
        mk[0] = m4.x; mk[1] = m4.y; mk[2] = m4.z; mk[3] = m4.w;
      } else if (msk) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (nb + j + q < p.N) mk[q] = __ldg(msk + j + q);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int n = nb + j + q;
        float val = acc[j + q];
        if (n < p.N) {
          if (p.bias) val += p.bias[n];
          if (p.beta) val += dst[j + q];
          if (p.relu) val = fmaxf(val, 0.0f);
          if (!(mk[q] > 0.0f)) val = 0.0f;
        }
        x[q] = val;
      }
      if (vec) {
        *reinterpret_cast<float4*>(dst + j) = make_float4(x[0], x[1], x[2], x[3]);
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (nb + j + q < p.N) dst[j + q] = x[q];
      }
    }
  }
};

static bool classify(const float* ptr, long long s_mn, long long s_k, int MN, int K, GemmOperand* op) {
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return false;
  op->p = ptr; op->s_mn = s_mn; op->s_k = s_k;
  if (s_k == 1 && (s_mn % 4) == 0 && (K % 4) == 0) { op->mode = 0; return true; }
  if (s_mn == 1 && (s_k % 4) == 0 && (MN % 4) == 0) { op->mode = 1; return true; }
  return false;
}

template <int BN, bool SPLIT3>
static int launch_gemm(const GemmParams& p, int splits, cudaStream_t st) {
  int rc = tc_launch<BN, SPLIT3, GemmProblem<BN, SPLIT3>>(p, cdiv(p.M, TM), cdiv(p.N, BN), splits, st);
  if (rc) return rc;
  g_launches++;
  A2C_LAUNCH_CHECK();
  return 0;
}

}  // namespace tc

// Returns 0 when the GEMM ran on the tensor cores, 1 when the shape does not
// qualify (caller falls back to the fp32 FFMA kernel), anything else = error.
int splitk_reduce(const float* partial, float* C, long long ldc, const float* bias, int M, int N, int splits,
                  int relu, int beta, cudaStream_t st);   // gemm_simt.cu

int gemm_tc_try(const float* A, long long sam, long long sak, const float* B, long long sbk,
                long long sbn, float* C, long long ldc, const float* bias, int M, int N, int K,
                int relu, int beta, int precision, void* workspace, long long workspace_bytes,
                const float* cmask, cudaStream_t st) {
  // tiny problems are launch-latency bound and gain nothing from a 128-row MMA tile
  if (M < 64 || K < 64 || (long long)M * N * K < (1ll << 22)) return 1;
  tc::GemmParams p{};
  if (!tc::classify(A, sam, sak, M, K, &p.A)) return 1;
  if (!tc::classify(B, sbn, sbk, N, K, &p.B)) return 1;
  p.C = C; p.ldc = ldc; p.bias = bias; p.M = M; p.N = N; p.K = K; p.relu = relu; p.beta = beta;
  p.cmask = cmask;
  p.A.dbg = p.B.dbg = debug_knob("A2C_GEMM_DBG");
  p.nkb = cdiv(K, tc::KB); p.kb_per_split = p.nkb;
  const bool split3 = (precision == 1);
  int bn = N > 64 ? 128 : (N > 32 ? 64 : 32);
  // prefer more CTAs when the grid would not fill the machine
  const int sms = num_sms();
  while (bn > 64 && (long long)cdiv(N, bn) * cdiv(M, tc::TM) < sms) bn >>= 1;
  // Split-K: an output with few tiles and a long reduction (weight gradients: dW[256, 256] over
  // K = 16,384 samples is 4 tiles) would keep a handful of SMs busy; split K over the others and
  // reduce the partial products (fp32, fixed order) in a second small kernel.
  int splits = 1;
  const long long tiles = (long long)cdiv(N, bn) * cdiv(M, tc::TM);
  if (workspace && !cmask && tiles * 2 <= sms && p.nkb >= 64) {        // K >= 2048
    long long want = sms / tiles;
    if (want > p.nkb / 8) want = p.nkb / 8;                 // at least 8 k-blocks (256 k) per split
    const long long per = (long long)M * N * (long long)sizeof(float);
    if (want * per > workspace_bytes) want = workspace_bytes / per;
    if (want > 1) splits = (int)want;
  }
  if (splits > 1) {
    p.kb_per_split = cdiv(p.nkb, splits);
    splits = cdiv(p.nkb, p.kb_per_split);
    p.partial = static_cast<float*>(workspace);
  }
  int rc;
  if (split3) {
    switch (bn) {
      case 128: rc = tc::launch_gemm<128, true>(p, splits, st); break;
      case 64: rc = tc::launch_gemm<64, true>(p, splits, st); break;
      default: rc = tc::launch_gemm<32, true>(p, splits, st); break;
    }
  } else {
    switch (bn) {
      case 128: rc = tc::launch_gemm<128, false>(p, splits, st); break;
      case 64: rc = tc::launch_gemm<64, false>(p, splits, st); break;
      default: rc = tc::launch_gemm<32, false>(p, splits, st); break;
    }
  }
  if (rc || splits == 1) return rc;
  return splitk_reduce(p.partial, C, ldc, bias, M, N, splits, relu, beta, st);
}

}  // namespace a2c

#ifdef A2C_TC_TRACE
extern "C" int a2c_debug_tc_trace(long long* host) {
  return (int)cudaMemcpyFromSymbol(host, a2c::tc::g_tc_trace, sizeof(a2c::tc::g_tc_trace));
}
#endif

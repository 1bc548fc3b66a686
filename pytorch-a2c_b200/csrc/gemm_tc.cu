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
};

struct GemmParams {
  GemmOperand A, B;
  float* C;
  long long ldc;
  const float* bias;
  int M, N, K;
  int relu, beta;
  int nkb, kb_per_split;
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

// Load phase: global -> registers.  t = thread index within the operand's producers.
template <int ROWS>
__device__ __forceinline__ void gemm_load(const GemmOperand& op, int MN, int K, int mn0, int k0, int t,
                                          GemmFrag<ROWS>& f) {
  if (op.mode == 0) {
    // k contiguous: unit = (row, 16-byte chunk); a quarter-warp covers one row's 128 bytes
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int u = i * PROD_THREADS + t, row = u >> 3, chunk = u & 7;
      const int mn = mn0 + row, k = k0 + chunk * 4;
      f.v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < ROWS && mn < MN && k < K)
        f.v[i] = __ldg(reinterpret_cast<const float4*>(op.p + (long long)mn * op.s_mn + k));
    }
  } else {
    // mn contiguous: unit = (group of 4 rows, k-chunk): 4 loads along mn (transposed at store time).
    // 8 consecutive lanes take 8 consecutive groups, so a warp load covers 4 k-rows x 128 contiguous bytes
    const int g = mn_group(t), chunk = k_chunk(t);
    const int mn = mn0 + g * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + chunk * 4 + j;
      f.v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (g * 4 < ROWS && mn < MN && k < K)
        f.v[j] = __ldg(reinterpret_cast<const float4*>(op.p + (long long)k * op.s_k + mn));
    }
  }
}

// Store phase: registers -> swizzled tile(s), with the hi/lo split.
template <int ROWS, bool SPLIT3>
__device__ __forceinline__ void gemm_store(const GemmOperand& op, const GemmFrag<ROWS>& f, uint8_t* hi,
                                           uint8_t* lo, int t) {
  if (op.mode == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int u = i * PROD_THREADS + t;
      if ((u >> 3) < ROWS) store_chunk<SPLIT3>(hi, lo, u >> 3, u & 7, f.v[i]);
    }
  } else {
    const int g = mn_group(t), chunk = k_chunk(t);
    if (g * 4 < ROWS) {
      const float4* v = f.v;
      store_chunk<SPLIT3>(hi, lo, g * 4 + 0, chunk, make_float4(v[0].x, v[1].x, v[2].x, v[3].x));
      store_chunk<SPLIT3>(hi, lo, g * 4 + 1, chunk, make_float4(v[0].y, v[1].y, v[2].y, v[3].y));
      store_chunk<SPLIT3>(hi, lo, g * 4 + 2, chunk, make_float4(v[0].z, v[1].z, v[2].z, v[3].z));
      store_chunk<SPLIT3>(hi, lo, g * 4 + 3, chunk, make_float4(v[0].w, v[1].w, v[2].w, v[3].w));
    }
  }
}

template <int BN, bool SPLIT3>
struct GemmProblem {
  using Params = GemmParams;
  // 16 producer warps, loads two k-blocks ahead: 64 KB of operand loads in flight per SM
  static constexpr int NA = 8, NB = 8, KC = 8, DEPTH = 2;
  struct AState {};
  using AFrag = GemmFrag<TM>;
  using BFrag = GemmFrag<BN>;
  static __device__ __forceinline__ void init(const Params&, uint8_t*, int, int) {}
  static __device__ __forceinline__ void a_setup(const Params&, const uint8_t*, int, int, AState&) {}
  static __device__ __forceinline__ void load_a(const Params& p, const AState&, const uint8_t*, int m0, int kb,
                                                int t, AFrag& f) {
    gemm_load<TM>(p.A, p.M, p.K, m0, kb * KB, t, f);
  }
  static __device__ __forceinline__ void store_a(const Params& p, const AFrag& f, uint8_t* hi, uint8_t* lo, int t) {
    gemm_store<TM, SPLIT3>(p.A, f, hi, lo, t);
  }
  static __device__ __forceinline__ void load_b(const Params& p, const uint8_t*, int n0, int kb, int t, BFrag& f) {
    gemm_load<BN>(p.B, p.N, p.K, n0, kb * KB, t, f);
  }
  static __device__ __forceinline__ void store_b(const Params& p, const BFrag& f, uint8_t* hi, uint8_t* lo, int t) {
    gemm_store<BN, SPLIT3>(p.B, f, hi, lo, t);
  }
  template <int HALF>
  static __device__ __forceinline__ void epilogue(const Params& p, int row, int nb, const float* acc, int) {
    if (row >= p.M) return;
    float* dst = p.C + (long long)row * p.ldc + nb;
    const bool vec = ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) && (nb + HALF <= p.N);
#pragma unroll
    for (int j = 0; j < HALF; j += 4) {
      float x[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int n = nb + j + q;
        float val = acc[j + q];
        if (n < p.N) {
          if (p.bias) val += p.bias[n];
          if (p.beta) val += dst[j + q];
          if (p.relu) val = fmaxf(val, 0.0f);
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
static int launch_gemm(const GemmParams& p, cudaStream_t st) {
  int rc = tc_launch<BN, SPLIT3, GemmProblem<BN, SPLIT3>>(p, cdiv(p.M, TM), cdiv(p.N, BN), 1, st);
  if (rc) return rc;
  g_launches++;
  A2C_LAUNCH_CHECK();
  return 0;
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
  tc::GemmParams p{};
  if (!tc::classify(A, sam, sak, M, K, &p.A)) return 1;
  if (!tc::classify(B, sbn, sbk, N, K, &p.B)) return 1;
  p.C = C; p.ldc = ldc; p.bias = bias; p.M = M; p.N = N; p.K = K; p.relu = relu; p.beta = beta;
  p.nkb = cdiv(K, tc::KB); p.kb_per_split = p.nkb;
  const bool split3 = (precision == 1);
  int bn = N > 64 ? 128 : (N > 32 ? 64 : 32);
  // prefer more CTAs when the grid would not fill the machine
  const int sms = num_sms();
  while (bn > 64 && (long long)cdiv(N, bn) * cdiv(M, tc::TM) < sms) bn >>= 1;
  if (split3) {
    switch (bn) {
      case 128: return tc::launch_gemm<128, true>(p, st);
      case 64: return tc::launch_gemm<64, true>(p, st);
      default: return tc::launch_gemm<32, true>(p, st);
    }
  }
  switch (bn) {
    case 128: return tc::launch_gemm<128, false>(p, st);
    case 64: return tc::launch_gemm<64, false>(p, st);
    default: return tc::launch_gemm<32, false>(p, st);
  }
}

}  // namespace a2c

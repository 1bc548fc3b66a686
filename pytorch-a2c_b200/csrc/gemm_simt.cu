// fp32 FFMA GEMM with arbitrary element strides (the exact-parity path of
// stage 4).  Replaces torch's Linear / mm forward and their autograd matmuls
// (a2c/models.py:40,46-47,247-263,380-394,472-475,626-636).
//
// 128x128x16 CTA tile, 256 threads, 8x8 register tile per thread, shared
// memory double-buffered with register-staged global prefetch.  Each operand is
// described as an (MN, K) matrix with two element strides and is loaded with
// 128-bit accesses along whichever dimension is contiguous.  A "state view"
// operand reads the u8 plane ring (or f32 states) of stage 1 directly, so the
// first layer never needs a materialised fp32 state tensor.
#include <atomic>

#include "common.cuh"

namespace a2c {
extern std::atomic<long long> g_launches;

constexpr int BM = 128, BN = 128, BK = 16, GTHREADS = 256, SPAD = 4;
enum { MODE_K = 0, MODE_MN = 1, MODE_GEN = 2, MODE_VIEW_K = 3, MODE_VIEW_MN = 4 };

struct Operand {
  const float* p;
  long long s_mn, s_k;   // element (i, k) at p[i*s_mn + k*s_k]
  a2c_state_view view;   // MODE_VIEW_*
};

struct GemmParams {
  Operand A, B;
  float* C;
  long long ldc;
  const float* bias;
  int M, N, K;
  int relu, beta;
  int splits, k_per_split;   // k_per_split multiple of BK
  float* partial;            // [splits, M, N] when splits > 1
};

// plane validity of channel c for sample (r, t): no reset flag in planes t+c+1 .. t+fs-1
__device__ __forceinline__ bool view_valid(const a2c_state_view& v, long long r, int t, int c) {
  if (!v.pflag) return true;
  const uint8_t* pf = v.pflag + r * v.pflag_stride_r + t;
  bool ok = true;
  for (int j = c + 1; j < v.fs; ++j) ok = ok && (pf[j] == 0);
  return ok;
}

__device__ __forceinline__ float4 view_load4(const a2c_state_view& v, long long off) {
  if (v.dtype == 0) {
    const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(static_cast<const uint8_t*>(v.base) + off));
    return make_float4((float)(w & 0xff), (float)((w >> 8) & 0xff), (float)((w >> 16) & 0xff),
                       (float)(w >> 24));
  }
  return __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(v.base) + off));
}

// Loads one 128 x 16 operand tile into registers (8 floats per thread).
template <int MODE>
__device__ __forceinline__ void tile_load(const Operand& op, int MN, int mn0, int k0, int kend,
                                          float (&reg)[8]) {
  const int tid = threadIdx.x;
  if constexpr (MODE == MODE_K) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int v = tid + i * GTHREADS, mn = v >> 2, kq = v & 3;
      float4 x = make_float4(0, 0, 0, 0);
      if (mn0 + mn < MN && k0 + kq * 4 < kend)
        x = __ldg(reinterpret_cast<const float4*>(op.p + (long long)(mn0 + mn) * op.s_mn + k0 + kq * 4));
      reg[i * 4 + 0] = x.x; reg[i * 4 + 1] = x.y; reg[i * 4 + 2] = x.z; reg[i * 4 + 3] = x.w;
    }
  } else if constexpr (MODE == MODE_MN) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int v = tid + i * GTHREADS, k = v >> 5, mq = v & 31;
      float4 x = make_float4(0, 0, 0, 0);
      if (mn0 + mq * 4 < MN && k0 + k < kend)
        x = __ldg(reinterpret_cast<const float4*>(op.p + (long long)(k0 + k) * op.s_k + mn0 + mq * 4));
      reg[i * 4 + 0] = x.x; reg[i * 4 + 1] = x.y; reg[i * 4 + 2] = x.z; reg[i * 4 + 3] = x.w;
    }
  } else if constexpr (MODE == MODE_GEN) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int e = tid + i * GTHREADS, mn = e & 127, k = e >> 7;
      float x = 0.0f;
      if (mn0 + mn < MN && k0 + k < kend)
        x = __ldg(op.p + (long long)(mn0 + mn) * op.s_mn + (long long)(k0 + k) * op.s_k);
      reg[i] = x;
    }
  } else if constexpr (MODE == MODE_VIEW_K) {
    // rows = samples, K = (channel, pixel); a 16-wide K tile lies in one channel
    const a2c_state_view& vw = op.view;
    const int c = k0 / 6400, pix = k0 - c * 6400;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int v = tid + i * GTHREADS, mn = v >> 2, kq = v & 3;
      float4 x = make_float4(0, 0, 0, 0);
      const int n = mn0 + mn;
      if (n < MN && k0 + kq * 4 < kend) {
        const long long r = n / vw.T;
        const int t = n - (int)r * vw.T;
        if (view_valid(vw, r, t, c))
          x = view_load4(vw, r * vw.stride_r + (long long)t * vw.stride_t + (long long)c * vw.stride_c + pix + kq * 4);
      }
      reg[i * 4 + 0] = x.x; reg[i * 4 + 1] = x.y; reg[i * 4 + 2] = x.z; reg[i * 4 + 3] = x.w;
    }
  } else {  // MODE_VIEW_MN: MN = (channel, pixel) features, K = samples
    const a2c_state_view& vw = op.view;
    const int c = mn0 / 6400, pix = mn0 - c * 6400;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int v = tid + i * GTHREADS, k = v >> 5, mq = v & 31;
      float4 x = make_float4(0, 0, 0, 0);
      const int n = k0 + k;
      if (mn0 + mq * 4 < MN && n < kend) {
        const long long r = n / vw.T;
        const int t = n - (int)r * vw.T;
        if (view_valid(vw, r, t, c))
          x = view_load4(vw, r * vw.stride_r + (long long)t * vw.stride_t + (long long)c * vw.stride_c + pix + mq * 4);
      }
      reg[i * 4 + 0] = x.x; reg[i * 4 + 1] = x.y; reg[i * 4 + 2] = x.z; reg[i * 4 + 3] = x.w;
    }
  }
}

template <int MODE>
__device__ __forceinline__ void tile_store(float (*S)[BM + SPAD], const float (&reg)[8]) {
  const int tid = threadIdx.x;
  if constexpr (MODE == MODE_K || MODE == MODE_VIEW_K) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int v = tid + i * GTHREADS, mn = v >> 2, kq = v & 3;
#pragma unroll
      for (int j = 0; j < 4; ++j) S[kq * 4 + j][mn] = reg[i * 4 + j];
    }
  } else if constexpr (MODE == MODE_MN || MODE == MODE_VIEW_MN) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int v = tid + i * GTHREADS, k = v >> 5, mq = v & 31;
      *reinterpret_cast<float4*>(&S[k][mq * 4]) =
          make_float4(reg[i * 4 + 0], reg[i * 4 + 1], reg[i * 4 + 2], reg[i * 4 + 3]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int e = tid + i * GTHREADS, mn = e & 127, k = e >> 7;
      S[k][mn] = reg[i];
    }
  }
}

template <int AMODE, int BMODE>
__global__ void __launch_bounds__(GTHREADS, 2) gemm_kernel(const GemmParams p) {
  __shared__ __align__(16) float As[2][BK][BM + SPAD];
  __shared__ __align__(16) float Bs[2][BK][BN + SPAD];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * p.k_per_split;
  const int kend = min(p.K, kbeg + p.k_per_split);

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

  float ra[8], rb[8];
  tile_load<AMODE>(p.A, p.M, m0, kbeg, kend, ra);
  tile_load<BMODE>(p.B, p.N, n0, kbeg, kend, rb);
  tile_store<AMODE>(As[0], ra);
  tile_store<BMODE>(Bs[0], rb);
  __syncthreads();

  int buf = 0;
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    const bool more = k0 + BK < kend;
    if (more) {
      tile_load<AMODE>(p.A, p.M, m0, k0 + BK, kend, ra);
      tile_load<BMODE>(p.B, p.N, n0, k0 + BK, kend, rb);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) {
      tile_store<AMODE>(As[buf ^ 1], ra);
      tile_store<BMODE>(Bs[buf ^ 1], rb);
      __syncthreads();
      buf ^= 1;
    }
  }

  // epilogue
  const bool split = p.splits > 1;
  float* out = split ? p.partial + (long long)blockIdx.z * p.M * p.N : p.C;
  const long long ld = split ? p.N : p.ldc;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= p.M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int n = n0 + jh * 64 + tx * 4;
      float v[4] = {acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]};
      float* dst = out + (long long)m * ld + n;
      if (!split) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (n + j < p.N) {
            float x = v[j];
            if (p.bias) x += p.bias[n + j];
            if (p.beta) x += dst[j];
            if (p.relu) x = fmaxf(x, 0.0f);
            v[j] = x;
          }
        }
      }
      if (n + 3 < p.N && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
        *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n + j < p.N) dst[j] = v[j];
      }
    }
  }
}

__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ partial, float* __restrict__ C, long long ldc,
                     const float* __restrict__ bias, int M, int N, int splits, int relu, int beta) {
  const long long total = (long long)M * N;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(e / N), n = (int)(e - (long long)m * N);
    double sd = 0.0;                                   // partial products of a long reduction: sum in double
    for (int z = 0; z < splits; ++z) sd += (double)partial[(long long)z * total + e];
    float s = (float)sd;
    if (bias) s += bias[n];
    float* dst = C + (long long)m * ldc + n;
    if (beta) s += *dst;
    if (relu) s = fmaxf(s, 0.0f);
    *dst = s;
  }
}

__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ X, float* __restrict__ out, long long M, int N,
              long long rows_per_block, int accumulate) {
  // blockDim = (32, 8): x -> column, y -> row lane.  grid.x over column tiles,
  // grid.y over row chunks (atomics combine chunks).
  __shared__ float sm[8][33];
  const int n = blockIdx.x * 32 + threadIdx.x;
  const long long r0 = blockIdx.y * rows_per_block;
  const long long r1 = min(M, r0 + rows_per_block);
  float s = 0.0f;
  if (n < N)
    for (long long r = r0 + threadIdx.y; r < r1; r += 8) s += X[r * N + n];
  sm[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    float t = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sm[i][threadIdx.x];
    if (gridDim.y == 1 && !accumulate) out[n] = t;
    else atomicAdd(&out[n], t);
  }
}

__global__ void __launch_bounds__(256)
relu_bwd_kernel(const float* __restrict__ dY, const float* __restrict__ Y, float* __restrict__ dX,
                long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    dX[i] = Y[i] > 0.0f ? dY[i] : 0.0f;
}

static int pick_mode(const float* p, long long s_mn, long long s_k, int MN, int K, int kgran) {
  const bool al = (reinterpret_cast<uintptr_t>(p) & 15) == 0;
  if (s_k == 1 && al && (s_mn % 4) == 0 && (K % 4) == 0 && (kgran % 4) == 0) return MODE_K;
  if (s_mn == 1 && al && (s_k % 4) == 0 && (MN % 4) == 0) return MODE_MN;
  return MODE_GEN;
}

template <int AMODE>
static void launch_b(int bmode, const GemmParams& p, dim3 grid, cudaStream_t st) {
  switch (bmode) {
    case MODE_K: gemm_kernel<AMODE, MODE_K><<<grid, GTHREADS, 0, st>>>(p); break;
    case MODE_MN: gemm_kernel<AMODE, MODE_MN><<<grid, GTHREADS, 0, st>>>(p); break;
    case MODE_VIEW_MN: gemm_kernel<AMODE, MODE_VIEW_MN><<<grid, GTHREADS, 0, st>>>(p); break;
    default: gemm_kernel<AMODE, MODE_GEN><<<grid, GTHREADS, 0, st>>>(p); break;
  }
}

int gemm_dispatch(GemmParams p, int amode, int bmode, void* workspace, long long ws_bytes,
                  cudaStream_t st) {
  const int tm = cdiv(p.M, BM), tn = cdiv(p.N, BN);
  const long long tiles = (long long)tm * tn;
  int splits = 1;
  const int sms = num_sms();
  if (workspace && tiles < sms && p.K >= 8 * BK) {
    long long want = (2ll * sms) / tiles;
    long long maxk = p.K / (4 * BK);
    if (want > maxk) want = maxk;
    const long long per = (long long)p.M * p.N * sizeof(float);
    if (per > 0 && want * per > ws_bytes) want = ws_bytes / per;
    if (want > 1) splits = (int)want;
  }
  int kps = cdiv(cdiv(p.K, splits), BK) * BK;
  splits = cdiv(p.K, kps);
  p.splits = splits;
  p.k_per_split = kps;
  p.partial = static_cast<float*>(workspace);
  dim3 grid(tn, tm, splits);
  switch (amode) {
    case MODE_K: launch_b<MODE_K>(bmode, p, grid, st); break;
    case MODE_MN: launch_b<MODE_MN>(bmode, p, grid, st); break;
    case MODE_VIEW_K: launch_b<MODE_VIEW_K>(bmode, p, grid, st); break;
    default: launch_b<MODE_GEN>(bmode, p, grid, st); break;
  }
  g_launches++;
  A2C_LAUNCH_CHECK();
  if (splits > 1) {
    long long total = (long long)p.M * p.N;
    long long blocks = cdivll(total, 256);
    if (blocks > sms * 8) blocks = sms * 8;
    splitk_reduce_kernel<<<(int)blocks, 256, 0, st>>>(p.partial, p.C, p.ldc, p.bias, p.M, p.N, splits,
                                                    p.relu, p.beta);
    g_launches++;
    A2C_LAUNCH_CHECK();
  }
  return 0;
}

int splitk_reduce(const float* partial, float* C, long long ldc, const float* bias, int M, int N, int splits,
                  int relu, int beta, cudaStream_t st) {
  long long blocks = cdivll((long long)M * N, 256);
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  splitk_reduce_kernel<<<(int)blocks, 256, 0, st>>>(partial, C, ldc, bias, M, N, splits, relu, beta);
  g_launches++;
  A2C_LAUNCH_CHECK();
  return 0;
}

int gemm_tc_try(const float* A, long long sam, long long sak, const float* B, long long sbk,
                long long sbn, float* C, long long ldc, const float* bias, int M, int N, int K,
                int relu, int beta, int precision, void* workspace, long long workspace_bytes,
                const float* cmask, cudaStream_t st);   // gemm_tc.cu; returns 1 when it did not handle the shape

static int check_view(const a2c_state_view* v) {
  A2C_ARG_CHECK(v && v->base, "state view: null");
  A2C_ARG_CHECK(v->fs > 0 && v->T > 0, "state view: bad fs/T");
  A2C_ARG_CHECK(v->dtype == 0 || v->dtype == 1, "state view: dtype must be 0 (u8) or 1 (f32)");
  A2C_ARG_CHECK((v->stride_r % 4) == 0 && (v->stride_t % 4) == 0 && (v->stride_c % 4) == 0 &&
                    (reinterpret_cast<uintptr_t>(v->base) & 15) == 0,
                "state view: strides must be multiples of 4 elements, base 16-byte aligned");
  return 0;
}

}  // namespace a2c

using namespace a2c;

extern "C" {

int a2c_gemm(const float* A, long long sam, long long sak, const float* B, long long sbk,
             long long sbn, float* C, long long ldc, const float* bias, int M, int N, int K,
             int relu, int beta, int precision, void* workspace, long long workspace_bytes,
             void* stream) {
  return a2c_gemm_masked(A, sam, sak, B, sbk, sbn, C, ldc, bias, nullptr, M, N, K, relu, beta, precision,
                         workspace, workspace_bytes, stream);
}

int a2c_gemm_masked(const float* A, long long sam, long long sak, const float* B, long long sbk,
                    long long sbn, float* C, long long ldc, const float* bias, const float* Cmask, int M,
                    int N, int K, int relu, int beta, int precision, void* workspace,
                    long long workspace_bytes, void* stream) {
  A2C_ARG_CHECK(A && B && C, "gemm: null pointer");
  A2C_ARG_CHECK(M >= 0 && N >= 0 && K >= 0 && ldc >= N, "gemm: bad sizes M=%d N=%d K=%d ldc=%lld", M, N, K, ldc);
  A2C_ARG_CHECK(beta == 0 || beta == 1, "gemm: beta must be 0 or 1");
  if (M == 0 || N == 0) return 0;
  cudaStream_t st = as_stream(stream);
  if (precision != 0) {
    int rc = gemm_tc_try(A, sam, sak, B, sbk, sbn, C, ldc, bias, M, N, K, relu, beta, precision,
                         workspace, workspace_bytes, Cmask, st);
    if (rc != 1) return rc;   // 0 = done on tensor cores, other = error; 1 = shape not handled
  }
  A2C_ARG_CHECK(!Cmask || ldc == N, "gemm_masked: the FFMA path needs a contiguous C (ldc == N)");
  GemmParams p{};
  p.A.p = A; p.A.s_mn = sam; p.A.s_k = sak;
  p.B.p = B; p.B.s_mn = sbn; p.B.s_k = sbk;
  p.C = C; p.ldc = ldc; p.bias = bias; p.M = M; p.N = N; p.K = K; p.relu = relu; p.beta = beta;
  const int amode = pick_mode(A, sam, sak, M, K, BK);
  const int bmode = pick_mode(B, sbn, sbk, N, K, BK);
  if (int rc = gemm_dispatch(p, amode, bmode, workspace, workspace_bytes, st)) return rc;
  if (Cmask) {
    const long long n = (long long)M * N;
    long long blocks = cdivll(n, 256);
    if (blocks > num_sms() * 16) blocks = num_sms() * 16;
    relu_bwd_kernel<<<(int)blocks, 256, 0, st>>>(C, Cmask, C, n);
    g_launches++;
    A2C_LAUNCH_CHECK();
  }
  return 0;
}

int a2c_linear_state_fwd(const a2c_state_view* x, const float* W, const float* bias, float* Y, int R,
                         int Nout, int relu, void* workspace, long long workspace_bytes, void* stream) {
  if (int rc = check_view(x)) return rc;
  A2C_ARG_CHECK(W && Y && R >= 0 && Nout > 0, "linear_state_fwd: bad args");
  A2C_ARG_CHECK((reinterpret_cast<uintptr_t>(W) & 15) == 0, "linear_state_fwd: W must be 16-byte aligned");
  const int M = R * x->T, K = x->fs * 6400;
  if (M == 0) return 0;
  GemmParams p{};
  p.A.view = *x;
  p.B.p = W; p.B.s_mn = K; p.B.s_k = 1;
  p.C = Y; p.ldc = Nout; p.bias = bias; p.M = M; p.N = Nout; p.K = K; p.relu = relu; p.beta = 0;
  // few output tiles (C1: 60 x 256 = 2 tiles over K = 12,800): split K over the SMs
  return gemm_dispatch(p, MODE_VIEW_K, MODE_K, workspace, workspace_bytes, as_stream(stream));
}

int a2c_linear_state_wgrad(const a2c_state_view* x, const float* dY, float* dW, int R, int Nout,
                           void* workspace, long long workspace_bytes, void* stream) {
  if (int rc = check_view(x)) return rc;
  A2C_ARG_CHECK(dY && dW && R >= 0 && Nout > 0, "linear_state_wgrad: bad args");
  const int Ms = R * x->T, Kf = x->fs * 6400;
  // dW[Nout, Kf] = dY^T[Nout, Ms] * X[Ms, Kf]
  GemmParams p{};
  p.A.p = dY; p.A.s_mn = 1; p.A.s_k = Nout;           // A(m = out feature, k = sample)
  p.B.view = *x;
  p.C = dW; p.ldc = Kf; p.M = Nout; p.N = Kf; p.K = Ms;
  if (Ms == 0) {
    A2C_CUDA_CHECK(cudaMemsetAsync(dW, 0, sizeof(float) * Nout * Kf, as_stream(stream)));
    return 0;
  }
  const int amode = pick_mode(dY, 1, Nout, Nout, Ms, BK);
  return gemm_dispatch(p, amode, MODE_VIEW_MN, workspace, workspace_bytes, as_stream(stream));
}

int a2c_colsum(const float* X, float* out, long long M, int N, void* stream) {
  A2C_ARG_CHECK(X && out && M >= 0 && N > 0, "colsum: bad args");
  cudaStream_t st = as_stream(stream);
  int chunks = (int)cdivll(M, 2048);
  if (chunks < 1) chunks = 1;
  if (chunks > 1024) chunks = 1024;
  const long long rpb = cdivll(M > 0 ? M : 1, chunks);
  if (chunks > 1) A2C_CUDA_CHECK(cudaMemsetAsync(out, 0, sizeof(float) * N, st));
  colsum_kernel<<<dim3(cdiv(N, 32), chunks), dim3(32, 8), 0, st>>>(X, out, M, N, rpb, 0);
  g_launches++;
  A2C_LAUNCH_CHECK();
  return 0;
}

int a2c_relu_bwd(const float* dY, const float* Y, float* dX, long long n, void* stream) {
  A2C_ARG_CHECK(dY && Y && dX && n >= 0, "relu_bwd: bad args");
  if (n == 0) return 0;
  long long blocks = cdivll(n, 256);
  if (blocks > num_sms() * 16) blocks = num_sms() * 16;
  relu_bwd_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(dY, Y, dX, n);
  g_launches++;
  A2C_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"

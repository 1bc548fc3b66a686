// LayerNorm (FCModel value head, a2c/models.py:392), GRU gate math
// (a2c/models.py:465-476) and the fused clip + RMSprop/Adam step
// (a2c/updater.py:129-132).  All HBM/latency bound elementwise or row kernels.
#include <atomic>

#include "common.cuh"

namespace a2c {
extern std::atomic<long long> g_launches;

// ---- LayerNorm: one warp per row ---------------------------------------------
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const float* __restrict__ X, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float* __restrict__ Y, float* __restrict__ mean,
                     float* __restrict__ rstd, long long rows, int H) {
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* x = X + row * H;
  float s = 0.0f;
  for (int i = lane; i < H; i += 32) s += x[i];
  const float mu = warp_sum(s) / H;
  float v = 0.0f;
  for (int i = lane; i < H; i += 32) { const float d = x[i] - mu; v += d * d; }
  const float rs = rsqrtf(warp_sum(v) / H + 1e-5f);
  for (int i = lane; i < H; i += 32) Y[row * H + i] = (x[i] - mu) * rs * gamma[i] + beta[i];
  if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
}

// dX per row; dgamma/dbeta accumulated per block in shared memory then atomically.
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float* __restrict__ dY, const float* __restrict__ X,
                     const float* __restrict__ gamma, const float* __restrict__ mean,
                     const float* __restrict__ rstd, float* __restrict__ dX,
                     float* __restrict__ dgamma, float* __restrict__ dbeta, long long rows, int H) {
  extern __shared__ float sm[];   // [2][H]
  float* sg = sm;
  float* sb = sm + H;
  for (int i = threadIdx.x; i < 2 * H; i += blockDim.x) sm[i] = 0.0f;
  __syncthreads();
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (long long row = blockIdx.x * (long long)wpb + (threadIdx.x >> 5); row < rows;
       row += (long long)gridDim.x * wpb) {
    const float mu = mean[row], rs = rstd[row];
    const float* x = X + row * H;
    const float* dy = dY + row * H;
    float s1 = 0.0f, s2 = 0.0f;
    for (int i = lane; i < H; i += 32) {
      const float xh = (x[i] - mu) * rs, g = dy[i] * gamma[i];
      s1 += g; s2 += g * xh;
      atomicAdd(&sg[i], dy[i] * xh);
      atomicAdd(&sb[i], dy[i]);
    }
    s1 = warp_sum(s1) / H; s2 = warp_sum(s2) / H;
    for (int i = lane; i < H; i += 32) {
      const float xh = (x[i] - mu) * rs, g = dy[i] * gamma[i];
      dX[row * H + i] = rs * (g - s1 - xh * s2);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    atomicAdd(&dgamma[i], sg[i]);
    atomicAdd(&dbeta[i], sb[i]);
  }
}

// ---- GRU gates -----------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(256)
gru_gates1_kernel(const float* __restrict__ gx, const float* __restrict__ gh,
                  const float* __restrict__ h, float* __restrict__ z, float* __restrict__ r,
                  float* __restrict__ rh, long long B, int H) {
  const long long total = B * H;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long b = e / H;
    const int j = (int)(e - b * H);
    const float zz = sigmoidf_(gx[b * 3 * H + j] + gh[b * 2 * H + j]);
    const float rr = sigmoidf_(gx[b * 3 * H + H + j] + gh[b * 2 * H + H + j]);
    z[e] = zz; r[e] = rr; rh[e] = rr * h[e];
  }
}

__global__ void __launch_bounds__(256)
gru_gates2_kernel(const float* __restrict__ gx, const float* __restrict__ gc,
                  const float* __restrict__ h, const float* __restrict__ z,
                  const float* __restrict__ keep, float* __restrict__ cand,
                  float* __restrict__ hnew, float* __restrict__ hnext, long long B, int H) {
  const long long total = B * H;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long b = e / H;
    const int j = (int)(e - b * H);
    const float c = tanhf(gx[b * 3 * H + 2 * H + j] + gc[e]);
    const float zz = z[e];
    const float hn = zz * h[e] + (1.0f - zz) * c;
    cand[e] = c; hnew[e] = hn;
    if (hnext) hnext[e] = keep ? hn * keep[b] : hn;
  }
}

__global__ void __launch_bounds__(256)
gru_gates2_bwd_kernel(const float* __restrict__ dhnew, const float* __restrict__ dhnext,
                      const float* __restrict__ keep, const float* __restrict__ h,
                      const float* __restrict__ z, const float* __restrict__ cand,
                      float* __restrict__ dgc, float* __restrict__ dz_out, float* __restrict__ dh,
                      long long total, int H) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    float g = dhnew ? dhnew[e] : 0.0f;
    if (dhnext) g += dhnext[e] * (keep ? keep[e / H] : 1.0f);
    const float zz = z[e], c = cand[e];
    dgc[e] = g * (1.0f - zz) * (1.0f - c * c);     // grad wrt candidate pre-activation
    dz_out[e] = g * (h[e] - c);                   // grad wrt z (post-sigmoid)
    dh[e] = g * zz;                               // direct path z*h
  }
}

__global__ void __launch_bounds__(256)
gru_gates1_bwd_kernel(const float* __restrict__ drh, const float* __restrict__ dz_out,
                      const float* __restrict__ h, const float* __restrict__ z,
                      const float* __restrict__ r, float* __restrict__ dgzr,
                      float* __restrict__ dh, long long B, int H) {
  const long long total = B * H;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long b = e / H;
    const int j = (int)(e - b * H);
    const float zz = z[e], rr = r[e], d = drh[e];
    dgzr[b * 2 * H + j] = dz_out[e] * zz * (1.0f - zz);
    dgzr[b * 2 * H + H + j] = d * h[e] * rr * (1.0f - rr);
    dh[e] += d * rr;
  }
}

// ---- clip + optimizer ------------------------------------------------------------
__global__ void __launch_bounds__(256)
grad_sumsq_kernel(const float* __restrict__ g, double* __restrict__ out, long long n) {
  __shared__ double scratch[32];
  double s = 0;
  const long long n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    const float4 v = g4[i];
    s += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long i = n4 << 2; i < n; ++i) s += (double)g[i] * g[i];
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) atomicAdd(out, s);
}

__global__ void __launch_bounds__(256)
clip_optim_kernel(float* __restrict__ param, float* __restrict__ grad, float* __restrict__ s1,
                  float* __restrict__ s2, const double* __restrict__ sumsq, a2c_optim_cfg cfg,
                  float* __restrict__ norm_out, long long n) {
  const float norm = (float)sqrt(*sumsq);
  float coef = cfg.max_norm / (norm + 1e-6f);          // torch clip_grad_norm_
  coef = coef > 1.0f ? 1.0f : coef;
  if (norm_out && blockIdx.x == 0 && threadIdx.x == 0) *norm_out = norm;
  float bc1 = 1.0f, bc2s = 1.0f;
  if (cfg.kind == 1) {
    bc1 = 1.0f - powf(cfg.beta1, (float)cfg.step);
    bc2s = sqrtf(1.0f - powf(cfg.beta2, (float)cfg.step));
  }
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float g = grad[i] * coef;
    float p = param[i];
    if (cfg.kind == 0) {                                // torch.optim.RMSprop defaults
      const float sq = cfg.alpha * s1[i] + (1.0f - cfg.alpha) * g * g;
      s1[i] = sq;
      p -= cfg.lr * g / (sqrtf(sq) + cfg.eps);
    } else {                                            // torch.optim.Adam defaults
      const float m = cfg.beta1 * s1[i] + (1.0f - cfg.beta1) * g;
      const float v = cfg.beta2 * s2[i] + (1.0f - cfg.beta2) * g * g;
      s1[i] = m; s2[i] = v;
      const float denom = sqrtf(v) / bc2s + cfg.eps;
      p -= (cfg.lr / bc1) * m / denom;
    }
    param[i] = p;
    if (cfg.zero_grad) grad[i] = 0.0f;
  }
}

static inline int ew_blocks(long long total) {
  long long b = cdivll(total, 256);
  const long long cap = (long long)num_sms() * 16;
  return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace a2c

using namespace a2c;

extern "C" {

int a2c_layernorm_fwd(const float* X, const float* gamma, const float* beta, float* Y, float* mean,
                      float* rstd, long long rows, int Hdim, void* stream) {
  A2C_ARG_CHECK(X && gamma && beta && Y && mean && rstd && rows >= 0 && Hdim > 0, "layernorm_fwd: bad args");
  if (rows == 0) return 0;
  layernorm_fwd_kernel<<<(unsigned)cdivll(rows, 8), 256, 0, as_stream(stream)>>>(X, gamma, beta, Y, mean,
                                                                               rstd, rows, Hdim);
  g_launches++;
  A2C_LAUNCH_CHECK();
  return 0;
}

int a2c_layernorm_bwd(const float* dY, const float* X, const float* gamma, const float* mean,
                      const float* rstd, float* dX, float* dgamma, float* dbeta, long long rows,
                      int Hdim, void* stream) {
  A2C_ARG_CHECK(dY && X && gamma && mean && rstd && dX && dgamma && dbeta && rows >= 0 && Hdim > 0,
                "layernorm_bwd: bad args");
  A2C_ARG_CHECK(Hdim <= 4096, "layernorm_bwd: H too large");
  cudaStream_t st = as_stream(stream);
  A2C_CUDA_CHECK(cudaMemsetAsync(dgamma, 0, sizeof(float) * Hdim, st));
  A2C_CUDA_CHECK(cudaMemsetAsync(dbeta, 0, sizeof(float) * Hdim, st));
  if (rows == 0) return 0;
  long long blocks = cdivll(rows, 8 * 16);
  if (blocks > num_sms() * 2) blocks = num_sms() * 2;
  if (blocks < 1) blocks = 1;
  layernorm_bwd_kernel<<<(unsigned)blocks, 256, 2 * Hdim * sizeof(float), st>>>(
      dY, X, gamma, mean, rstd, dX, dgamma, dbeta, rows, Hdim);
  g_launches++;
  A2C_LAUNCH_CHECK();
  return 0;
}

int a2c_gru_gates1(const float* gx, const float* gh, const float* h, float* z, float* r, float* rh,
                   long long B, int Hdim, void* stream) {
  A2C_ARG_CHECK(gx && gh && h && z && r && rh && B >= 0 && Hdim > 0, "gru_gates1: bad args");
  if (B == 0) return 0;
  gru_gates1_kernel<<<ew_blocks(B * Hdim), 256, 0, as_stream(stream)>>>(gx, gh, h, z, r, rh, B, Hdim);
  g_launches++;
  A2C_LAUNCH_CHECK();
  return 0;
}

int a2c_gru_gates2(const float* gx, const float* gc, const float* h, const float* z,
                   const float* keep, float* cand, float* hnew, float* hnext, long long B, int Hdim,
                   void* stream) {
  A2C_ARG_CHECK(gx && gc && h && z && cand && hnew && B >= 0 && Hdim > 0, "gru_gates2: bad args");
  if (B == 0) return 0;
  gru_gates2_kernel<<<ew_blocks(B * Hdim), 256, 0, as_stream(stream)>>>(gx, gc, h, z, keep, cand, hnew,
                                                                       hnext, B, Hdim);
  g_launches++;
  A2C_LAUNCH_CHECK();
  return 0;
}

int a2c_gru_gates2_bwd(const float* dhnew, const float* dhnext, const float* keep, const float* h,
                       const float* z, const float* cand, float* dgc, float* dz_out, float* dh,
                       long long B, int Hdim, void* stream) {
  A2C_ARG_CHECK((dhnew || dhnext) && h && z && cand && dgc && dz_out && dh && B >= 0 && Hdim > 0,
                "gru_gates2_bwd: bad args");
  if (B == 0) return 0;
  gru_gates2_bwd_kernel<<<ew_blocks(B * Hdim), 256, 0, as_stream(stream)>>>(
      dhnew, dhnext, keep, h, z, cand, dgc, dz_out, dh, B * Hdim, Hdim);
  g_launches++;
  A2C_LAUNCH_CHECK();
  return 0;
}

int a2c_gru_gates1_bwd(const float* drh, const float* dz_out, const float* h, const float* z,
                       const float* r, float* dgzr, float* dh, long long B, int Hdim, void* stream) {
  A2C_ARG_CHECK(drh && dz_out && h && z && r && dgzr && dh && B >= 0 && Hdim > 0,
                "gru_gates1_bwd: bad args");
  if (B == 0) return 0;
  gru_gates1_bwd_kernel<<<ew_blocks(B * Hdim), 256, 0, as_stream(stream)>>>(drh, dz_out, h, z, r, dgzr,
                                                                           dh, B, Hdim);
  g_launches++;
  A2C_LAUNCH_CHECK();
  return 0;
}

int a2c_grad_sumsq(const float* grad, double* sumsq, long long n, void* stream) {
  A2C_ARG_CHECK(grad && sumsq && n >= 0, "grad_sumsq: bad args");
  A2C_ARG_CHECK((reinterpret_cast<uintptr_t>(grad) & 15) == 0, "grad_sumsq: grad must be 16-byte aligned");
  if (n == 0) return 0;
  long long blocks = cdivll(n, 1024);
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  grad_sumsq_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(grad, sumsq, n);
  g_launches++;
  A2C_LAUNCH_CHECK();
  return 0;
}

int a2c_clip_optim_step(float* param, float* grad, float* state1, float* state2, const double* sumsq,
                        const a2c_optim_cfg* cfg, float* norm_out, long long n, void* stream) {
  A2C_ARG_CHECK(param && grad && state1 && sumsq && cfg && n >= 0, "clip_optim_step: bad args");
  A2C_ARG_CHECK(cfg->kind == 0 || (cfg->kind == 1 && state2 && cfg->step >= 1),
                "clip_optim_step: kind must be 0 (RMSprop) or 1 (Adam with state2, step >= 1)");
  if (n == 0) return 0;
  clip_optim_kernel<<<ew_blocks(n), 256, 0, as_stream(stream)>>>(param, grad, state1, state2, sumsq, *cfg,
                                                               norm_out, n);
  g_launches++;
  A2C_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"

// Stage 2 (rollout bookkeeping + reverse segmented scan), stage 3 (fused A2C
// loss forward + backward) and the acting-time action sampler.
//
// All three are latency/launch bound (28 B and 48 B per transition); the point
// of the kernels is to remove the reference's per-element Python loops and the
// autograd graph, not to chase bandwidth.  fp32 arithmetic follows the
// reference's association order with explicit round-to-nearest mul/add (no FMA
// contraction) so that deltas, advantages and returns are bit-identical to the
// CPU path.
#include <atomic>

#include "common.cuh"

namespace a2c {
extern std::atomic<long long> g_launches;

// ---------------------------------------------------------------------------
// a2c/runner.py:228-245 + a2c/utils.py:63-79, one thread per rollout slot.
//
// Why one thread per slot and not a warp-shuffle scan: the reference evaluates
//   rs = a[i] + g * rs   sequentially in fp32 (utils.py:76-77) and parity for the
// scans is BIT-exact.  A parallel (shuffle) scan composes the affine maps
// (a, g) o (a', g') = (a + g a', g g'), which rounds differently from the
// sequential recurrence, so the recurrence stays sequential inside a slot and
// the parallelism is across the R slots (a warp = 32 slots).  What a warp
// should not do is read global memory with a stride of T floats: the staged
// kernel below moves whole [slots x T] tiles through shared memory with
// coalesced accesses (odd row pitch -> conflict-free column walks) and each
// thread walks its slot there.  The direct kernel is the fallback for T too
// large for a shared-memory tile.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
delta_scan_direct_kernel(float* __restrict__ rewards, float* __restrict__ dones,
                  const float* __restrict__ values, float* __restrict__ deltas,
                  float* __restrict__ advs, float* __restrict__ returns, double* __restrict__ stats,
                  int R, int T, float gamma, float gl) {
  __shared__ double scratch[32];
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  double s_a = 0, s_a2 = 0, s_r = 0, s_r2 = 0;
  if (r < R) {
    const long long base = (long long)r * T;
    const float* v = values + (long long)r * (T + 1);
    float rew = rewards[base + T - 1], d = dones[base + T - 1];
    if (d == 0.0f) {                                   // runner.py:236-243
      rew = __fadd_rn(rew, __fmul_rn(gamma, v[T]));
      d = 1.0f;
      rewards[base + T - 1] = rew;
      dones[base + T - 1] = d;
    }
    float adv_rs = 0.0f, ret_rs = 0.0f;
    float v_next = 0.0f;
    for (int t = T - 1; t >= 0; --t) {
      if (t != T - 1) {
        rew = rewards[base + t];
        d = dones[base + t];
      }
      const float vt = v[t];
      float delta;
      if (t == T - 1)
        delta = __fsub_rn(rew, vt);                    // runner.py:244
      else                                             // runner.py:231
        delta = __fsub_rn(__fadd_rn(rew, __fmul_rn(__fmul_rn(gamma, v_next), __fsub_rn(1.0f, d))), vt);
      if (d == 1.0f) {                                 // utils.py:76
        adv_rs = 0.0f;
        ret_rs = 0.0f;
      }
      adv_rs = __fadd_rn(delta, __fmul_rn(gl, adv_rs));   // utils.py:77
      ret_rs = __fadd_rn(rew, __fmul_rn(gamma, ret_rs));
      deltas[base + t] = delta;
      if (advs) advs[base + t] = adv_rs;
      if (returns) returns[base + t] = ret_rs;
      s_a += adv_rs; s_a2 += (double)adv_rs * adv_rs;
      s_r += ret_rs; s_r2 += (double)ret_rs * ret_rs;
      v_next = vt;
    }
  }
  if (stats) {
    s_a = block_sum(s_a, scratch);
    s_a2 = block_sum(s_a2, scratch);
    s_r = block_sum(s_r, scratch);
    s_r2 = block_sum(s_r2, scratch);
    if (threadIdx.x == 0) {
      atomicAdd(&stats[0], s_a); atomicAdd(&stats[1], s_a2);
      atomicAdd(&stats[2], s_r); atomicAdd(&stats[3], s_r2);
    }
  }
}

// Staged variant: a block owns SB consecutive slots.  Tiles: rew, done (in place), val -> adv
// (in place: v[t] is consumed before adv[t] is written), delta, ret.
__global__ void __launch_bounds__(64)
delta_scan_staged_kernel(float* __restrict__ rewards, float* __restrict__ dones,
                         const float* __restrict__ values, float* __restrict__ deltas,
                         float* __restrict__ advs, float* __restrict__ returns, double* __restrict__ stats,
                         int R, int T, int SB, int TP, float gamma, float gl) {
  extern __shared__ float tile[];
  __shared__ double scratch[32];
  float* s_rew = tile;                       // [SB][TP]
  float* s_done = s_rew + SB * TP;
  float* s_val = s_done + SB * TP;           // [SB][TP], T + 1 <= TP; becomes adv
  float* s_delta = s_val + SB * TP;
  float* s_ret = s_delta + SB * TP;
  const int r0 = blockIdx.x * SB;
  const int nslot = min(SB, R - r0);
  const long long g0 = (long long)r0 * T;
  const int nel = nslot * T;
  for (int i = threadIdx.x; i < nel; i += blockDim.x) {
    const int sl = i / T, t = i - sl * T;
    s_rew[sl * TP + t] = rewards[g0 + i];
    s_done[sl * TP + t] = dones[g0 + i];
  }
  const long long v0 = (long long)r0 * (T + 1);
  for (int i = threadIdx.x; i < nslot * (T + 1); i += blockDim.x) {
    const int sl = i / (T + 1), t = i - sl * (T + 1);
    s_val[sl * TP + t] = values[v0 + i];
  }
  __syncthreads();
  double s_a = 0, s_a2 = 0, s_r = 0, s_r2 = 0;
  for (int sl = threadIdx.x; sl < nslot; sl += blockDim.x) {
    float* rw = s_rew + sl * TP;
    float* dn = s_done + sl * TP;
    float* v = s_val + sl * TP;
    float rew = rw[T - 1], d = dn[T - 1];
    if (d == 0.0f) {                                   // runner.py:236-243
      rew = __fadd_rn(rew, __fmul_rn(gamma, v[T]));
      d = 1.0f;
      rw[T - 1] = rew;
      dn[T - 1] = d;
    }
    float adv_rs = 0.0f, ret_rs = 0.0f, v_next = 0.0f;
    for (int t = T - 1; t >= 0; --t) {
      if (t != T - 1) { rew = rw[t]; d = dn[t]; }
      const float vt = v[t];
      float delta;
      if (t == T - 1)
        delta = __fsub_rn(rew, vt);                    // runner.py:244
      else                                             // runner.py:231
        delta = __fsub_rn(__fadd_rn(rew, __fmul_rn(__fmul_rn(gamma, v_next), __fsub_rn(1.0f, d))), vt);
      if (d == 1.0f) { adv_rs = 0.0f; ret_rs = 0.0f; }       // utils.py:76
      adv_rs = __fadd_rn(delta, __fmul_rn(gl, adv_rs));      // utils.py:77
      ret_rs = __fadd_rn(rew, __fmul_rn(gamma, ret_rs));
      s_delta[sl * TP + t] = delta;
      v[t] = adv_rs;
      s_ret[sl * TP + t] = ret_rs;
      s_a += adv_rs; s_a2 += (double)adv_rs * adv_rs;
      s_r += ret_rs; s_r2 += (double)ret_rs * ret_rs;
      v_next = vt;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nel; i += blockDim.x) {
    const int sl = i / T, t = i - sl * T, o = sl * TP + t;
    rewards[g0 + i] = s_rew[o];
    dones[g0 + i] = s_done[o];
    deltas[g0 + i] = s_delta[o];
    if (advs) advs[g0 + i] = s_val[o];
    if (returns) returns[g0 + i] = s_ret[o];
  }
  if (stats) {
    s_a = block_sum(s_a, scratch);
    s_a2 = block_sum(s_a2, scratch);
    s_r = block_sum(s_r, scratch);
    s_r2 = block_sum(s_r2, scratch);
    if (threadIdx.x == 0) {
      atomicAdd(&stats[0], s_a); atomicAdd(&stats[1], s_a2);
      atomicAdd(&stats[2], s_r); atomicAdd(&stats[3], s_r2);
    }
  }
}

// utils.discount on a flat buffer: the thread sitting on a segment end walks
// its segment backwards, so every segment is summed sequentially (bit-exact).
__global__ void __launch_bounds__(256)
discount_kernel(const float* __restrict__ a, const float* __restrict__ dones,
                float* __restrict__ out, double* __restrict__ stats, long long N, float g) {
  __shared__ double scratch[32];
  double s1 = 0, s2 = 0;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < N && (i == N - 1 || dones[i] == 1.0f)) {
    float rs = 0.0f;
    long long j = i;
    do {
      rs = __fadd_rn(a[j], __fmul_rn(g, rs));
      out[j] = rs;
      s1 += rs; s2 += (double)rs * rs;
      --j;
    } while (j >= 0 && dones[j] != 1.0f);
  }
  if (stats) {
    s1 = block_sum(s1, scratch);
    s2 = block_sum(s2, scratch);
    if (threadIdx.x == 0) { atomicAdd(&stats[0], s1); atomicAdd(&stats[1], s2); }
  }
}

// Both scans of an update (a2c/updater.py:70-71 and :86-88) in one launch: advs from deltas with
// factor ga, returns from rewards with factor gb (b may be NULL: use_nstep_rets).  Same
// segment-end walk as above, so both are bit-exact.  stats f64 [4]: sum / sumsq of each.
__global__ void __launch_bounds__(256)
discount2_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ dones,
                 float* __restrict__ outa, float* __restrict__ outb, double* __restrict__ stats, long long N,
                 float ga, float gb) {
  __shared__ double scratch[32];
  double a1 = 0, a2 = 0, b1 = 0, b2 = 0;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < N && (i == N - 1 || dones[i] == 1.0f)) {
    float ra = 0.0f, rb = 0.0f;
    long long j = i;
    do {
      ra = __fadd_rn(a[j], __fmul_rn(ga, ra));
      outa[j] = ra;
      a1 += ra; a2 += (double)ra * ra;
      if (b) {
        rb = __fadd_rn(b[j], __fmul_rn(gb, rb));
        outb[j] = rb;
        b1 += rb; b2 += (double)rb * rb;
      }
      --j;
    } while (j >= 0 && dones[j] != 1.0f);
  }
  if (stats) {
    a1 = block_sum(a1, scratch);
    a2 = block_sum(a2, scratch);
    b1 = block_sum(b1, scratch);
    b2 = block_sum(b2, scratch);
    if (threadIdx.x == 0) {
      atomicAdd(&stats[0], a1); atomicAdd(&stats[1], a2);
      if (b) { atomicAdd(&stats[2], b1); atomicAdd(&stats[3], b2); }
    }
  }
}

// ---------------------------------------------------------------------------
// a2c/runner.py:94-97 + a2c/utils.py:45-60
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sample_action_kernel(const float* __restrict__ x, const float* __restrict__ u,
                     long long* __restrict__ actions, int B, int A, int is_prob) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* row = x + (long long)b * A;
  float m = -INFINITY, s = 0.0f;
  if (!is_prob) {
    for (int i = 0; i < A; ++i) m = fmaxf(m, row[i]);
    for (int i = 0; i < A; ++i) s += expf(row[i] - m);
  }
  const float ub = u[b];
  float cum = 0.0f;
  long long act = -1;
  for (int i = 0; i < A; ++i) {
    const float p = is_prob ? row[i] : __fdiv_rn(expf(row[i] - m), s);
    cum = __fadd_rn(cum, p);
    if (act < 0 && cum >= ub) act = i;
  }
  actions[b] = act;
}

// ---------------------------------------------------------------------------
// a2c/updater.py:97-106,119-127 and its backward, one thread per transition.
// ---------------------------------------------------------------------------
template <int A_MAX>
__global__ void __launch_bounds__(256)
loss_kernel(const float* __restrict__ logits, const float* __restrict__ vals,
            const long long* __restrict__ actions, const float* __restrict__ advs,
            const float* __restrict__ returns, const double* __restrict__ adv_stats,
            const float* __restrict__ ret_norm, a2c_loss_cfg cfg, float* __restrict__ dlogits,
            float* __restrict__ dvals, float* __restrict__ out, long long N, int A) {
  __shared__ double scratch[32];
  const double ng = (double)cfg.n_global;
  float a_mean = 0.0f, a_den = 1.0f;
  if (cfg.norm_advs) {
    const double mean = adv_stats[0] / ng;
    double var = (adv_stats[1] - ng * mean * mean) / (ng - 1.0);
    if (var < 0) var = 0;
    a_mean = (float)mean;
    a_den = (float)sqrt(var) + 1e-6f;
  }
  float r_mean = 0.0f, r_den = 1.0f;
  if (ret_norm) { r_mean = ret_norm[0]; r_den = ret_norm[1] + 1e-6f; }
  const float inv_n = (float)(1.0 / ng);

  double s_pi = 0, s_val = 0, s_ent = 0;
  for (long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x; n < N;
       n += (long long)gridDim.x * blockDim.x) {
    float z[A_MAX];
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < A_MAX; ++j)
      if (j < A) { z[j] = logits[n * A + j]; m = fmaxf(m, z[j]); }
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < A_MAX; ++j)
      if (j < A) { z[j] = z[j] - m; s += expf(z[j]); }
    const float ls = logf(s);
    float S = 0.0f;
    float p[A_MAX];
#pragma unroll
    for (int j = 0; j < A_MAX; ++j)
      if (j < A) { z[j] = z[j] - ls; p[j] = expf(z[j]); S += p[j] * z[j]; }   // z is now log-softmax
    long long a = actions[n];
    if (a < 0) a += A;                                  // torch negative indexing
    const float adv_raw = advs[n];
    const float v = vals[n];
    float ret = cfg.use_nstep ? adv_raw + v : returns[n];
    ret = (ret - r_mean) / r_den;
    const float adv = cfg.norm_advs ? (adv_raw - a_mean) / a_den : adv_raw;
    float lp_a = 0.0f;
#pragma unroll
    for (int j = 0; j < A_MAX; ++j)
      if (j < A) {
        const float onehot = (j == a) ? 1.0f : 0.0f;
        if (j == a) lp_a = z[j];
        dlogits[n * A + j] = cfg.pi_coef * (-adv * inv_n) * (onehot - p[j]) +
                             cfg.entr_coef * inv_n * p[j] * (z[j] - S);
      }
    const float dv = v - ret;
    dvals[n] = cfg.val_coef * 2.0f * dv * inv_n;
    s_pi += (double)(-lp_a * adv);
    s_val += (double)dv * dv;
    s_ent += (double)S;
  }
  s_pi = block_sum(s_pi, scratch);
  s_val = block_sum(s_val, scratch);
  s_ent = block_sum(s_ent, scratch);
  if (threadIdx.x == 0) {
    const float pi_loss = (float)(cfg.pi_coef * s_pi / ng);
    const float val_loss = (float)(cfg.val_coef * s_val / ng);
    const float entr_loss = (float)(-cfg.entr_coef * s_ent / ng);
    atomicAdd(&out[0], pi_loss + val_loss - entr_loss);
    atomicAdd(&out[1], pi_loss);
    atomicAdd(&out[2], val_loss);
    atomicAdd(&out[3], entr_loss);
  }
}

}  // namespace a2c

using namespace a2c;

extern "C" {

int a2c_delta_scan(float* rewards, float* dones, const float* values, float* deltas, float* advs,
                   float* returns, double* stats, int R, int T, float gamma, float gamma_lambda,
                   void* stream) {
  A2C_ARG_CHECK(rewards && dones && values && deltas, "delta_scan: null pointer");
  A2C_ARG_CHECK(R >= 0 && T >= 1, "delta_scan: bad sizes R=%d T=%d", R, T);
  if (R == 0) return 0;
  // staged through shared memory when a [slots x T] tile set fits in 48 KB
  const int TP = (T + 1) | 1;                       // odd pitch >= T + 1
  int SB = 0;
  for (int c = 64; c >= 8; c >>= 1)
    if ((size_t)5 * c * TP * sizeof(float) <= 48 * 1024) { SB = c; break; }
  if (SB) {
    const size_t smem = (size_t)5 * SB * TP * sizeof(float);
    delta_scan_staged_kernel<<<cdiv(R, SB), 64, smem, as_stream(stream)>>>(
        rewards, dones, values, deltas, advs, returns, stats, R, T, SB, TP, gamma, gamma_lambda);
  } else {
    delta_scan_direct_kernel<<<cdiv(R, 128), 128, 0, as_stream(stream)>>>(rewards, dones, values, deltas, advs,
                                                                        returns, stats, R, T, gamma,
                                                                        gamma_lambda);
  }
  g_launches++;
  A2C_LAUNCH_CHECK();
  return 0;
}

int a2c_discount2(const float* deltas, const float* rewards, const float* dones, float* advs,
                  float* returns, double* stats, long long N, float gamma_lambda, float gamma,
                  void* stream) {
  A2C_ARG_CHECK(deltas && dones && advs && N >= 0, "discount2: bad args");
  A2C_ARG_CHECK(!rewards || returns, "discount2: returns required when rewards are given");
  if (N == 0) return 0;
  discount2_kernel<<<(int)cdivll(N, 256), 256, 0, as_stream(stream)>>>(deltas, rewards, dones, advs, returns,
                                                                    stats, N, gamma_lambda, gamma);
  g_launches++;
  A2C_LAUNCH_CHECK();
  return 0;
}

int a2c_discount(const float* array, const float* dones, float* out, double* stats, long long N,
                 float factor, void* stream) {
  A2C_ARG_CHECK(array && dones && out && N >= 0, "discount: bad args");
  if (N == 0) return 0;
  discount_kernel<<<(int)cdivll(N, 256), 256, 0, as_stream(stream)>>>(array, dones, out, stats, N, factor);
  g_launches++;
  A2C_LAUNCH_CHECK();
  return 0;
}

int a2c_sample_action(const float* logits, const float* u, long long* actions, int B, int A,
                      int is_prob, void* stream) {
  A2C_ARG_CHECK(logits && u && actions && B >= 0 && A >= 1, "sample_action: bad args");
  if (B == 0) return 0;
  sample_action_kernel<<<cdiv(B, 256), 256, 0, as_stream(stream)>>>(logits, u, actions, B, A, is_prob);
  g_launches++;
  A2C_LAUNCH_CHECK();
  return 0;
}

int a2c_loss_fwd_bwd(const float* logits, const float* vals, const long long* actions,
                     const float* advs, const float* returns, const double* adv_stats,
                     const float* ret_norm, const a2c_loss_cfg* cfg, float* dlogits, float* dvals,
                     float* out, long long N, int A, void* stream) {
  A2C_ARG_CHECK(logits && vals && actions && advs && cfg && dlogits && dvals && out,
                "loss: null pointer");
  A2C_ARG_CHECK(cfg->use_nstep || returns, "loss: returns required unless use_nstep");
  A2C_ARG_CHECK(!cfg->norm_advs || adv_stats, "loss: adv_stats required for norm_advs");
  A2C_ARG_CHECK(A >= 1 && A <= 16, "loss: A=%d unsupported (1..16)", A);
  A2C_ARG_CHECK(cfg->n_global >= (cfg->norm_advs ? 2 : 1) && N >= 0, "loss: bad n_global/N");
  cudaStream_t st = as_stream(stream);
  A2C_CUDA_CHECK(cudaMemsetAsync(out, 0, 4 * sizeof(float), st));
  if (N == 0) return 0;
  long long blocks = cdivll(N, 256);
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  if (A <= 4)
    loss_kernel<4><<<(int)blocks, 256, 0, st>>>(logits, vals, actions, advs, returns, adv_stats,
                                               ret_norm, *cfg, dlogits, dvals, out, N, A);
  else
    loss_kernel<16><<<(int)blocks, 256, 0, st>>>(logits, vals, actions, advs, returns, adv_stats,
                                                ret_norm, *cfg, dlogits, dvals, out, N, A);
  g_launches++;
  A2C_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"

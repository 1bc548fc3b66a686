// fp32 direct convolution: forward, backward-data and backward-weight.
// Replaces nn.Conv2d (+ReLU) and its autograd in a2c/models.py:98,312,685.
//
// Forward: one CTA per (sample, band of output rows).  The input patch of a
// chunk of input channels is staged in shared memory (zero-filled halo), the
// chunk's weights are staged as [ci][ky][kx][co]; each thread owns PX
// consecutive output pixels x CTT output channels in registers and streams
// 128-bit shared loads.  The staging loader reads f32 NCHW, the u8 plane ring
// of stage 1 (state view), or -- for backward-data -- the ReLU-masked output
// gradient dilated by the stride, so backward-data is the same compute kernel
// with stride 1 and flipped/transposed weights.
//
// Backward-weight: persistent CTAs over (sample, band) work items; a thread
// owns COT output channels x 1 input channel x KxK taps in registers, pixels
// are split over thread groups, partials are reduced deterministically.
#include <atomic>

#include "common.cuh"

namespace a2c {
extern std::atomic<long long> g_launches;

struct ConvSrc {
  const float* X;          // f32 NCHW source (fwd input, or dY for bwd-data)
  const float* mask;       // same shape as X; value passes where mask > 0 (NULL = none)
  a2c_state_view view;     // used when use_view
  int use_view;
  int SC, SH, SW;          // source tensor dims (channels, height, width)
  int dil;                 // logical coordinate = source coordinate * dil
};

__device__ __forceinline__ float src_value(const ConvSrc& s, int n, int c, int iy, int ix) {
  // (iy, ix) are logical coordinates; returns 0 outside / between dilated taps
  if (iy < 0 || ix < 0) return 0.0f;
  int sy = iy, sx = ix;
  if (s.dil > 1) {
    if ((iy % s.dil) | (ix % s.dil)) return 0.0f;
    sy = iy / s.dil; sx = ix / s.dil;
  }
  if (sy >= s.SH || sx >= s.SW) return 0.0f;
  if (s.use_view) {
    const a2c_state_view& v = s.view;
    const long long r = n / v.T;
    const int t = n - (int)r * v.T;
    if (v.pflag) {
      const uint8_t* pf = v.pflag + r * v.pflag_stride_r + t;
      for (int j = c + 1; j < v.fs; ++j)
        if (pf[j]) return 0.0f;
    }
    const long long off = r * v.stride_r + (long long)t * v.stride_t + (long long)c * v.stride_c +
                          (long long)sy * s.SW + sx;
    if (v.dtype == 0) return (float)static_cast<const uint8_t*>(v.base)[off];
    return static_cast<const float*>(v.base)[off];
  }
  const long long off = (((long long)n * s.SC + c) * s.SH + sy) * s.SW + sx;
  float x = __ldg(s.X + off);
  if (s.mask && !(__ldg(s.mask + off) > 0.0f)) x = 0.0f;
  return x;
}

struct ConvFwdParams {
  ConvSrc src;
  const float* Wt;         // [Cout][Cin][K][K]; flipT: [Cin_w][Cout_w][K][K] read transposed+flipped
  int flipT;
  const float* bias;
  float* Y;                // [N][Cout][OH][OW]
  int N, Cin, Cout, OH, OW, pad, relu;
  int rows, WP, PR, CC;    // band rows, patch pitch, patch rows, cin chunk
  int CoutP;               // Cout rounded up to 4: pitch of the staged weights
};

template <int K, int S, int PX, int CTT>
__global__ void __launch_bounds__(512, 1) conv_fwd_kernel(const ConvFwdParams p) {
  extern __shared__ __align__(16) float sm[];
  float* patch = sm;                                   // [CC][PR][WP]
  float* Ws = sm + (size_t)p.CC * p.PR * p.WP;         // [CC][K][K][Cout]
  const int bands = cdiv(p.OH, p.rows);
  const int n = blockIdx.x / bands, band = blockIdx.x - n * bands;
  const int oy0 = band * p.rows;
  const int npg = p.OW / PX;
  const int pixthreads = p.rows * npg;
  const int tid = threadIdx.x;
  const int g = tid / pixthreads, pt = tid - g * pixthreads;
  const int row = pt / npg, xg = pt - row * npg;
  const bool active = (g * CTT < p.CoutP) && (oy0 + row < p.OH);
  constexpr int NIN = (PX - 1) * S + K;
  constexpr int NL = (PX == 4) ? (NIN + 3) / 4 : 0;

  float acc[PX][CTT];
#pragma unroll
  for (int i = 0; i < PX; ++i)
#pragma unroll
    for (int j = 0; j < CTT; ++j) acc[i][j] = 0.0f;

  const int iy_base = oy0 * S - p.pad;
  for (int c0 = 0; c0 < p.Cin; c0 += p.CC) {
    const int cc_n = min(p.CC, p.Cin - c0);
    __syncthreads();
    // stage the input patch: smem column = ix + pad
    const int patch_elems = cc_n * p.PR * p.WP;
    for (int e = tid; e < patch_elems; e += blockDim.x) {
      const int col = e % p.WP;
      const int rr = (e / p.WP) % p.PR;
      const int cc = e / (p.WP * p.PR);
      patch[e] = src_value(p.src, n, c0 + cc, iy_base + rr, col - p.pad);
    }
    const int w_elems = cc_n * K * K * p.CoutP;
    for (int e = tid; e < w_elems; e += blockDim.x) {
      const int co = e % p.CoutP;
      int rest = e / p.CoutP;
      const int kx = rest % K; rest /= K;
      const int ky = rest % K;
      const int cc = rest / K;
      long long gi;
      if (!p.flipT) gi = (((long long)co * p.Cin + c0 + cc) * K + ky) * K + kx;
      else gi = (((long long)(c0 + cc) * p.Cout + co) * K + (K - 1 - ky)) * K + (K - 1 - kx);
      Ws[e] = co < p.Cout ? __ldg(p.Wt + gi) : 0.0f;
    }
    __syncthreads();
    if (active) {
      for (int cc = 0; cc < cc_n; ++cc) {
#pragma unroll
        for (int ky = 0; ky < K; ++ky) {
          const float* prow = patch + ((size_t)cc * p.PR + row * S + ky) * p.WP + xg * PX * S;
          float xin[(PX == 4) ? NL * 4 : NIN];
          if constexpr (PX == 4) {
#pragma unroll
            for (int q = 0; q < NL; ++q) {
              const float4 v = *reinterpret_cast<const float4*>(prow + q * 4);
              xin[q * 4 + 0] = v.x; xin[q * 4 + 1] = v.y; xin[q * 4 + 2] = v.z; xin[q * 4 + 3] = v.w;
            }
          } else {
#pragma unroll
            for (int q = 0; q < NIN; ++q) xin[q] = prow[q];
          }
#pragma unroll
          for (int kx = 0; kx < K; ++kx) {
            const float* wp = Ws + ((size_t)(cc * K + ky) * K + kx) * p.CoutP + g * CTT;
#pragma unroll
            for (int q = 0; q < CTT / 4; ++q) {
              const float4 w = *reinterpret_cast<const float4*>(wp + q * 4);
#pragma unroll
              for (int px = 0; px < PX; ++px) {
                const float x = xin[px * S + kx];
                acc[px][q * 4 + 0] = fmaf(x, w.x, acc[px][q * 4 + 0]);
                acc[px][q * 4 + 1] = fmaf(x, w.y, acc[px][q * 4 + 1]);
                acc[px][q * 4 + 2] = fmaf(x, w.z, acc[px][q * 4 + 2]);
                acc[px][q * 4 + 3] = fmaf(x, w.w, acc[px][q * 4 + 3]);
              }
            }
          }
        }
      }
    }
  }
  if (!active) return;
  const int oy = oy0 + row;
#pragma unroll
  for (int j = 0; j < CTT; ++j) {
    const int co = g * CTT + j;
    if (co >= p.Cout) break;
    const float b = p.bias ? p.bias[co] : 0.0f;
    float v[PX];
#pragma unroll
    for (int px = 0; px < PX; ++px) {
      v[px] = acc[px][j] + b;
      if (p.relu) v[px] = fmaxf(v[px], 0.0f);
    }
    float* dst = p.Y + (((long long)n * p.Cout + co) * p.OH + oy) * p.OW + xg * PX;
    if constexpr (PX == 4) {
      *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int px = 0; px < PX; ++px) dst[px] = v[px];
    }
  }
}

struct FwdPlan {
  int rows, WP, PR, CC, threads, G;
  size_t smem;
};

template <int K, int S, int PX, int CTT>
static int launch_fwd_t(ConvFwdParams p, int logicalW, cudaStream_t st) {
  constexpr int NIN = (PX - 1) * S + K;
  constexpr int NL = (PX == 4) ? (NIN + 3) / 4 : 0;
  const int npg = p.OW / PX;
  p.CoutP = (p.Cout + 3) & ~3;
  const int G = cdiv(p.CoutP, CTT);
  int rows = 512 / (npg * G);
  if (rows < 1) rows = 1;
  if (rows > p.OH) rows = p.OH;
  // balance bands
  const int bands = cdiv(p.OH, rows);
  rows = cdiv(p.OH, bands);
  int WP = (npg - 1) * PX * S + (PX == 4 ? NL * 4 : NIN);
  const int need = logicalW + 2 * p.pad;
  if (WP < need) WP = need;
  WP = (WP + 3) & ~3;
  const int PR = (rows - 1) * S + K;
  int CC = p.Cin;
  auto smem_for = [&](int cc) { return ((size_t)cc * PR * WP + (size_t)cc * K * K * p.CoutP) * sizeof(float); };
  while (CC > 1 && smem_for(CC) > 96 * 1024) CC = (CC + 1) / 2;
  const size_t smem = smem_for(CC);
  A2C_ARG_CHECK(smem <= 200 * 1024, "conv: shared memory plan too large (%zu)", smem);
  const int threads = rows * npg * G;
  A2C_ARG_CHECK(threads <= 512, "conv: thread plan too large (%d)", threads);
  p.rows = rows; p.WP = WP; p.PR = PR; p.CC = CC;
  auto kern = conv_fwd_kernel<K, S, PX, CTT>;
  static size_t configured = 0;
  if (smem > configured) {
    A2C_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = 200 * 1024;
  }
  const long long blocks = (long long)p.N * cdiv(p.OH, rows);
  kern<<<(unsigned)blocks, ((threads + 31) / 32) * 32, smem, st>>>(p);
  g_launches++;
  A2C_LAUNCH_CHECK();
  return 0;
}

template <int K, int S>
static int launch_fwd_ks(const ConvFwdParams& p, int logicalW, cudaStream_t st) {
  const int c = p.Cout;
  if (p.OW % 4 == 0) {
    if (c % 16 == 0) return launch_fwd_t<K, S, 4, 16>(p, logicalW, st);
    if (c % 12 == 0) return launch_fwd_t<K, S, 4, 12>(p, logicalW, st);
    if (c % 8 == 0) return launch_fwd_t<K, S, 4, 8>(p, logicalW, st);
    return launch_fwd_t<K, S, 4, 4>(p, logicalW, st);   // also covers Cout not a multiple of 4
  }
  if (c % 16 == 0) return launch_fwd_t<K, S, 1, 16>(p, logicalW, st);
  if (c % 12 == 0) return launch_fwd_t<K, S, 1, 12>(p, logicalW, st);
  if (c % 8 == 0) return launch_fwd_t<K, S, 1, 8>(p, logicalW, st);
  return launch_fwd_t<K, S, 1, 4>(p, logicalW, st);
}

static int launch_fwd(const ConvFwdParams& p, int K, int S, int logicalW, cudaStream_t st) {
  if (K == 3 && S == 1) return launch_fwd_ks<3, 1>(p, logicalW, st);
  if (K == 3 && S == 2) return launch_fwd_ks<3, 2>(p, logicalW, st);
  if (K == 4 && S == 1) return launch_fwd_ks<4, 1>(p, logicalW, st);
  if (K == 4 && S == 2) return launch_fwd_ks<4, 2>(p, logicalW, st);
  if (K == 8 && S == 4) return launch_fwd_ks<8, 4>(p, logicalW, st);
  if (K == 8 && S == 1) return launch_fwd_ks<8, 1>(p, logicalW, st);
  set_error("conv: unsupported kernel/stride K=%d S=%d", K, S);
  return A2C_ERR_ARG;
}

// ---------------------------------------------------------------------------
// backward weight
// ---------------------------------------------------------------------------
struct ConvWgParams {
  ConvSrc src;             // forward input (dil = 1)
  const float* dY;         // [N][Cout][OH][OW]
  const float* Ymask;
  float* partial;          // [grid][Cout*Cin*K*K + Cout]
  int N, Cin, Cout, OH, OW, pad;
  int rows, bands, WP, PR, CoutP, PG, pairs;
};

template <int K, int S, int COT>
__global__ void __launch_bounds__(512, 1) conv_wgrad_kernel(const ConvWgParams p) {
  extern __shared__ __align__(16) float sm[];
  float* gs = sm;                                            // [rows][OW][CoutP]
  float* xs = sm + (size_t)p.rows * p.OW * p.CoutP;          // [Cin][PR][WP]
  const int tid = threadIdx.x;
  const int ncog = p.Cout / COT;
  const int pr = tid % p.pairs, pg = tid / p.pairs;
  const int cog = pr % ncog, ci = pr / ncog;
  const bool active = pg < p.PG;

  float acc[COT][K * K];
  float bacc[COT];
  const int wsize = p.Cout * p.Cin * K * K;
  // this thread's slots in the partial of (CTA, pixel group): flushed after every
  // work item so no fp32 chain is longer than one item's pixels
  float* outp = p.partial + ((size_t)blockIdx.x * p.PG + (active ? pg : 0)) * (wsize + p.Cout);
  bool first = true;

  const long long items = (long long)p.N * p.bands;
  for (long long it = blockIdx.x; it < items; it += gridDim.x) {
    const int n = (int)(it / p.bands), band = (int)(it - (long long)n * p.bands);
    const int oy0 = band * p.rows;
    const int nrows = min(p.rows, p.OH - oy0);
#pragma unroll
    for (int i = 0; i < COT; ++i) {
      bacc[i] = 0.0f;
#pragma unroll
      for (int j = 0; j < K * K; ++j) acc[i][j] = 0.0f;
    }
    __syncthreads();
    // stage masked dY as [row][ox][co]
    const int g_elems = p.Cout * nrows * p.OW;
    for (int e = tid; e < g_elems; e += blockDim.x) {
      const int ox = e % p.OW;
      const int rr = (e / p.OW) % nrows;
      const int co = e / (p.OW * nrows);
      const long long gi = (((long long)n * p.Cout + co) * p.OH + oy0 + rr) * p.OW + ox;
      float v = __ldg(p.dY + gi);
      if (p.Ymask && !(__ldg(p.Ymask + gi) > 0.0f)) v = 0.0f;
      gs[((size_t)rr * p.OW + ox) * p.CoutP + co] = v;
    }
    const int iy_base = oy0 * S - p.pad;
    const int x_elems = p.Cin * p.PR * p.WP;
    for (int e = tid; e < x_elems; e += blockDim.x) {
      const int col = e % p.WP;
      const int rr = (e / p.WP) % p.PR;
      const int c = e / (p.WP * p.PR);
      xs[e] = src_value(p.src, n, c, iy_base + rr, col - p.pad);
    }
    __syncthreads();
    if (active) {
      const int npix = nrows * p.OW;
      const int chunk = cdiv(npix, p.PG);
      const int pbeg = pg * chunk, pend = min(npix, pbeg + chunk);
      int rr = pbeg / p.OW, ox = pbeg - rr * p.OW;
      for (int pix = pbeg; pix < pend; ++pix) {
        const float* gp = gs + ((size_t)rr * p.OW + ox) * p.CoutP + cog * COT;
        const float* xp = xs + ((size_t)ci * p.PR + rr * S) * p.WP + ox * S;
        float gv[COT];
        if constexpr (COT % 4 == 0) {
#pragma unroll
          for (int q = 0; q < COT / 4; ++q) {
            const float4 t = *reinterpret_cast<const float4*>(gp + q * 4);
            gv[q * 4 + 0] = t.x; gv[q * 4 + 1] = t.y; gv[q * 4 + 2] = t.z; gv[q * 4 + 3] = t.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < COT; ++i) gv[i] = gp[i];
        }
        if (ci == 0) {
#pragma unroll
          for (int i = 0; i < COT; ++i) bacc[i] += gv[i];
        }
#pragma unroll
        for (int ky = 0; ky < K; ++ky)
#pragma unroll
          for (int kx = 0; kx < K; ++kx) {
            const float xv = xp[ky * p.WP + kx];
#pragma unroll
            for (int i = 0; i < COT; ++i) acc[i][ky * K + kx] = fmaf(gv[i], xv, acc[i][ky * K + kx]);
          }
        if (++ox == p.OW) { ox = 0; ++rr; }
      }
#pragma unroll
      for (int i = 0; i < COT; ++i) {
        const int co = cog * COT + i;
        float* o = outp + ((size_t)co * p.Cin + ci) * K * K;
#pragma unroll
        for (int j = 0; j < K * K; ++j) o[j] = first ? acc[i][j] : o[j] + acc[i][j];
        if (ci == 0) outp[wsize + co] = first ? bacc[i] : outp[wsize + co] + bacc[i];
      }
      first = false;
    }
  }
}

// Sums the per-(CTA, pixel group) partials in double.  block = (32 elements, 8 part stripes).
__global__ void __launch_bounds__(256)
partial_reduce_kernel(const float* __restrict__ partial, int nparts, int wsize, int cout,
                      float* __restrict__ dW, float* __restrict__ dbias) {
  __shared__ double sm[8][33];
  const int total = wsize + cout;
  const int e = blockIdx.x * 32 + threadIdx.x;
  double s = 0.0;
  if (e < total)
    for (int k = threadIdx.y; k < nparts; k += 8) s += (double)partial[(size_t)k * total + e];
  sm[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && e < total) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sm[i][threadIdx.x];
    if (e < wsize) dW[e] = (float)t;
    else if (dbias) dbias[e - wsize] = (float)t;
  }
}

struct WgPlan {
  int rows, bands, WP, PR, CoutP, PG, pairs, threads, grid, COT;
  size_t smem;
};

static int plan_wgrad(const a2c_conv_shape* s, WgPlan* pl) {
  const int K = s->K, S = s->stride;
  pl->COT = (K == 8) ? 1 : 4;
  if (pl->COT == 4 && (s->Cout / 4) * s->Cin > 512) pl->COT = 8;
  if (s->Cout % pl->COT) return -1;
  pl->pairs = (s->Cout / pl->COT) * s->Cin;
  if (pl->pairs > 512) return -1;
  pl->PG = 256 / pl->pairs;
  if (pl->PG < 1) pl->PG = 1;
  pl->threads = ((pl->pairs * pl->PG + 31) / 32) * 32;
  pl->CoutP = s->Cout + 4;
  pl->WP = (((s->OW - 1) * S + K) > (s->W + 2 * s->pad) ? ((s->OW - 1) * S + K) : (s->W + 2 * s->pad));
  pl->WP = (pl->WP + 3) & ~3;
  int rows = s->OH;
  auto smem_for = [&](int r) {
    const int PR = (r - 1) * S + K;
    return ((size_t)r * s->OW * pl->CoutP + (size_t)s->Cin * PR * pl->WP) * sizeof(float);
  };
  while (rows > 1 && smem_for(rows) > 100 * 1024) rows = (rows + 1) / 2;
  pl->bands = cdiv(s->OH, rows);
  pl->rows = cdiv(s->OH, pl->bands);
  pl->PR = (pl->rows - 1) * S + K;
  pl->smem = smem_for(pl->rows);
  if (pl->smem > 200 * 1024) return -1;
  pl->grid = num_sms() * 2;
  return 0;
}

template <int K, int S, int COT>
static int launch_wgrad_t(ConvWgParams p, const WgPlan& pl, int wsize, float* dW, float* dbias,
                          cudaStream_t st) {
  auto kern = conv_wgrad_kernel<K, S, COT>;
  A2C_CUDA_CHECK(set_max_dynamic_smem(reinterpret_cast<const void*>(kern), 200 * 1024));
  const long long items = (long long)p.N * p.bands;
  int grid = pl.grid;
  if (items < grid) grid = (int)items;
  kern<<<grid, pl.threads, pl.smem, st>>>(p);
  g_launches++;
  A2C_LAUNCH_CHECK();
  const int total = wsize + p.Cout;
  partial_reduce_kernel<<<cdiv(total, 32), dim3(32, 8), 0, st>>>(p.partial, grid * pl.PG, wsize, p.Cout,
                                                             dW, dbias);
  g_launches++;
  A2C_LAUNCH_CHECK();
  return 0;
}

// Which kernel family the last a2c_conv2d_* call of this host thread used:
// 0 = tcgen05 shift-GEMM, 1 = tcgen05 implicit GEMM (im2col tiles), 2 = fp32 FFMA.  The host layer
// warns once per layer shape when a precision >= 1 request ends on the FFMA kernels (a shape that
// misses every tensor-core path would otherwise be a silent 5-10x slowdown).
static thread_local int g_last_path = -1;
#define A2C_PATH_RETURN(path, rc_expr) \
  do {                                   \
    int rc__ = (rc_expr);                \
    if (rc__ != 1) {                     \
      g_last_path = (path);              \
      return rc__;                       \
    }                                    \
  } while (0)

static int check_shape(const a2c_conv_shape* s) {
  A2C_ARG_CHECK(s, "conv: null shape");
  A2C_ARG_CHECK(s->Cin > 0 && s->Cout > 0 && s->H > 0 && s->W > 0 && s->K > 0 && s->stride > 0 && s->pad >= 0,
                "conv: bad shape");
  A2C_ARG_CHECK(s->OH == (s->H - s->K + 2 * s->pad) / s->stride + 1 &&
                    s->OW == (s->W - s->K + 2 * s->pad) / s->stride + 1,
                "conv: OH/OW inconsistent with H/W/K/stride/pad");
  return 0;
}

static int fill_src(ConvSrc* src, const float* X, const a2c_state_view* xview, const a2c_conv_shape* s) {
  memset(src, 0, sizeof(*src));
  src->X = X; src->SC = s->Cin; src->SH = s->H; src->SW = s->W; src->dil = 1;
  if (xview) {
    A2C_ARG_CHECK(xview->base && xview->fs == s->Cin && xview->T > 0 && s->H == 80 && s->W == 80,
                  "conv: state view must have fs == Cin and 80x80 planes");
    A2C_ARG_CHECK(xview->dtype == 0 || xview->dtype == 1, "conv: state view dtype");
    src->view = *xview; src->use_view = 1;
  } else {
    A2C_ARG_CHECK(X, "conv: null input");
  }
  return 0;
}

// conv_tc.cu: 0 = ran on the tensor cores, 1 = shape not handled, else error
int conv_fwd_tc_try(const float* X, const a2c_state_view* xview, const float* Wt, const float* bias,
                    float* Y, int N, const a2c_conv_shape* s, int relu, int precision, cudaStream_t st);
int conv_dgrad_tc_try(const float* dY, const float* Wt, float* dX, int N, const a2c_conv_shape* s,
                      int precision, cudaStream_t st);
long long conv_wgrad_tc_workspace(const a2c_conv_shape* s);
int conv_wgrad_tc_try(const float* X, const a2c_state_view* xview, const float* dY, float* dW,
                      float* dbias, int N, const a2c_conv_shape* s, void* workspace,
                      long long workspace_bytes, int precision, cudaStream_t st);

// conv_shift_tc.cu (3xTF32 only): 3x3 pad-1 layers as shifted-descriptor GEMMs, no im2col
int conv_fwd_shift_try(const float* X, const a2c_state_view* xview, const float* Wt, const float* bias, float* Y,
                       int N, const a2c_conv_shape* s, int relu, cudaStream_t st, uint8_t* relu_bits = nullptr);
int conv_dgrad_shift_try(const float* dY, const float* Ymask, const float* Wt, float* dX, const float* Xmask,
                         double* dbias_prev, int N, const a2c_conv_shape* s, cudaStream_t st,
                         const uint8_t* Xbits = nullptr);

// bits u8 [N, C / 8, HW]: bit (c & 7) of byte [n, c >> 3, pos] = Y[n, c, pos] > 0  (fallback producer of the
// 1-bit ReLU masks for forward paths that do not write them in their epilogue)
__global__ void __launch_bounds__(256)
relu_bits_kernel(const float* __restrict__ Y, uint8_t* __restrict__ bits, long long nbytes, int C, int HW) {
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < nbytes; i += (long long)gridDim.x * 256) {
    const int pos = (int)(i % HW);
    const long long nc8 = i / HW;
    const int c8 = (int)(nc8 % (C >> 3));
    const long long n = nc8 / (C >> 3);
    const float* y = Y + (n * C + c8 * 8) * HW + pos;
    uint32_t b = 0;
#pragma unroll
    for (int u = 0; u < 8; ++u) b |= (__ldg(y + (long long)u * HW) > 0.0f ? 1u : 0u) << u;
    bits[i] = (uint8_t)b;
  }
}

// acc f64 [C] += per-channel sums of X f32 [N, C, HW] (where mask > 0, mask optional)
__global__ void __launch_bounds__(256)
channel_sum_kernel(const float* __restrict__ X, const float* __restrict__ mask, double* __restrict__ acc, int N,
                   int C, int HW) {
  __shared__ double scratch[32];
  const int c = blockIdx.y;
  double s = 0.0;
  const long long per = (long long)N * HW;
  const bool vec = (HW & 3) == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0 &&
                   (!mask || (reinterpret_cast<uintptr_t>(mask) & 15) == 0);
  if (vec) {
    const int HW4 = HW >> 2;
    const long long per4 = per >> 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per4; i += (long long)gridDim.x * blockDim.x) {
      const int n = (int)(i / HW4);
      const long long a = ((long long)n * C + c) * HW4 + (i - (long long)n * HW4);
      const float4 v = __ldg(reinterpret_cast<const float4*>(X) + a);
      float t;
      if (mask) {
        const float4 m = __ldg(reinterpret_cast<const float4*>(mask) + a);
        t = ((m.x > 0.f ? v.x : 0.f) + (m.y > 0.f ? v.y : 0.f)) + ((m.z > 0.f ? v.z : 0.f) + (m.w > 0.f ? v.w : 0.f));
      } else {
        t = (v.x + v.y) + (v.z + v.w);
      }
      s += (double)t;
    }
  } else {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per; i += (long long)gridDim.x * blockDim.x) {
      const int n = (int)(i / HW);
      const long long a = ((long long)n * C + c) * HW + (i - (long long)n * HW);
      const float v = __ldg(X + a);
      s += (double)((!mask || __ldg(mask + a) > 0.f) ? v : 0.f);
    }
  }
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) atomicAdd(&acc[c], s);
}
__global__ void f64_to_f32_kernel(const double* __restrict__ in, float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)in[i];
}
__global__ void __launch_bounds__(256)
relu_mask_inplace_kernel(float* __restrict__ X, const float* __restrict__ M, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (!(M[i] > 0.0f)) X[i] = 0.0f;
}

static int channel_sum_launch(const float* X, const float* mask, double* acc, int N, int C, int HW, cudaStream_t st) {
  int bx = (int)cdivll((long long)N * HW, 256 * 16);
  if (bx > 128) bx = 128;
  if (bx < 1) bx = 1;
  channel_sum_kernel<<<dim3(bx, C), 256, 0, st>>>(X, mask, acc, N, C, HW);
  g_launches++;
  A2C_LAUNCH_CHECK();
  return 0;
}

bool conv_dgrad_shift_ok(const a2c_conv_shape* s);
bool conv_wgrad_shift_ok(const a2c_conv_shape* s);
bool conv_wgrad_shift_bias_free(const a2c_conv_shape* s);
long long conv_wgrad_shift_workspace(const a2c_conv_shape* s);
int conv_wgrad_shift_try(const float* X, const a2c_state_view* xview, const float* dY, const float* Ymask,
                         float* dW, float* dbias, int N, const a2c_conv_shape* s, void* workspace,
                         long long workspace_bytes, cudaStream_t st);

int conv_fwd_patch_try(const float* X, const a2c_state_view* xview, const float* Wt, const float* bias,
                       float* Y, int N, const a2c_conv_shape* s, int relu, int precision, cudaStream_t st);
int conv_dgrad_patch_try(const float* dY, const float* Wt, float* dX, int N, const a2c_conv_shape* s,
                         int precision, cudaStream_t st);
long long conv_wgrad_patch_workspace(const a2c_conv_shape* s);
int conv_wgrad_patch_try(const float* X, const a2c_state_view* xview, const float* dY, float* dW,
                         float* dbias, int N, const a2c_conv_shape* s, void* workspace,
                         long long workspace_bytes, int precision, cudaStream_t st);

}  // namespace a2c

using namespace a2c;

extern "C" {

int a2c_conv2d_fwd(const float* X, const a2c_state_view* xview, const float* Wt, const float* bias,
                   float* Y, int N, const a2c_conv_shape* s, int relu, int precision, void* stream) {
  if (int rc = check_shape(s)) return rc;
  A2C_ARG_CHECK(Wt && Y && N >= 0, "conv_fwd: bad args");
  if (N == 0) return 0;
  // 3x3 pad-1 layers (first layers included: they read the uint8 ring directly) run as shift-GEMMs;
  // the other tensor-core paths take layers with more than 4 input channels, the rest is FFMA.
  if (precision == 1) {
    A2C_ARG_CHECK(xview || X, "conv_fwd: null input");
    A2C_PATH_RETURN(0, conv_fwd_shift_try(X, xview, Wt, bias, Y, N, s, relu, as_stream(stream)));
  }
  if (precision != 0 && s->Cin > 4) {
    A2C_ARG_CHECK(xview || X, "conv_fwd: null input");
    A2C_PATH_RETURN(1, conv_fwd_patch_try(X, xview, Wt, bias, Y, N, s, relu, precision, as_stream(stream)));
    A2C_PATH_RETURN(1, conv_fwd_tc_try(X, xview, Wt, bias, Y, N, s, relu, precision, as_stream(stream)));
  }
  g_last_path = 2;
  ConvFwdParams p{};
  if (int rc = fill_src(&p.src, X, xview, s)) return rc;
  p.Wt = Wt; p.flipT = 0; p.bias = bias; p.Y = Y; p.N = N; p.Cin = s->Cin; p.Cout = s->Cout;
  p.OH = s->OH; p.OW = s->OW; p.pad = s->pad; p.relu = relu;
  return launch_fwd(p, s->K, s->stride, s->W, as_stream(stream));
}

int a2c_conv2d_fwd_relu_bits(const float* X, const a2c_state_view* xview, const float* Wt, const float* bias,
                             float* Y, unsigned char* relu_bits, int N, const a2c_conv_shape* s, int precision,
                             void* stream) {
  if (int rc = check_shape(s)) return rc;
  A2C_ARG_CHECK(Wt && Y && relu_bits && N >= 0 && (xview || X), "conv_fwd_relu_bits: bad args");
  A2C_ARG_CHECK((s->Cout & 7) == 0, "conv_fwd_relu_bits: Cout must be a multiple of 8");
  if (N == 0) return 0;
  if (precision == 1) {
    // the shift-GEMM epilogue writes the bits next to Y
    A2C_PATH_RETURN(0, conv_fwd_shift_try(X, xview, Wt, bias, Y, N, s, 1, as_stream(stream), relu_bits));
  }
  // any other path: plain forward, then one pass over Y
  if (int rc = a2c_conv2d_fwd(X, xview, Wt, bias, Y, N, s, 1, precision, stream)) return rc;
  const long long nbytes = (long long)N * (s->Cout >> 3) * s->OH * s->OW;
  long long blocks = cdivll(nbytes, 256);
  if (blocks > num_sms() * 16) blocks = num_sms() * 16;
  relu_bits_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(Y, relu_bits, nbytes, s->Cout, s->OH * s->OW);
  g_launches++;
  A2C_LAUNCH_CHECK();
  return 0;
}

int a2c_conv2d_bwd_data(const float* dY, const float* Ymask, const float* Wt, float* dX, int N,
                        const a2c_conv_shape* s, int precision, void* stream) {
  if (int rc = check_shape(s)) return rc;
  A2C_ARG_CHECK(dY && Wt && dX && N >= 0, "conv_bwd_data: bad args");
  if (N == 0) return 0;
  if (precision == 1) {
    A2C_PATH_RETURN(0, conv_dgrad_shift_try(dY, Ymask, Wt, dX, nullptr, nullptr, N, s, as_stream(stream)));
  }
  if (precision != 0 && !Ymask) {
    A2C_PATH_RETURN(1, conv_dgrad_patch_try(dY, Wt, dX, N, s, precision, as_stream(stream)));
    A2C_PATH_RETURN(1, conv_dgrad_tc_try(dY, Wt, dX, N, s, precision, as_stream(stream)));
  }
  g_last_path = 2;
  // dX = stride-1 conv of the dilated, masked dY with flipped/transposed weights
  ConvFwdParams p{};
  p.src.X = dY; p.src.mask = Ymask; p.src.SC = s->Cout; p.src.SH = s->OH; p.src.SW = s->OW;
  p.src.dil = s->stride;
  p.Wt = Wt; p.flipT = 1; p.bias = nullptr; p.Y = dX; p.N = N;
  p.Cin = s->Cout;   // channels being reduced
  p.Cout = s->Cin;   // channels being produced
  p.OH = s->H; p.OW = s->W; p.pad = s->K - 1 - s->pad; p.relu = 0;
  A2C_ARG_CHECK(p.pad >= 0, "conv_bwd_data: pad > K-1 unsupported");
  const int logicalW = (s->OW - 1) * s->stride + 1;
  return launch_fwd(p, s->K, 1, logicalW, as_stream(stream));
}

static int bwd_data_relu_impl(const float* dY, const float* Ymask, const float* Wt, const float* X,
                              const unsigned char* Xbits, float* dX, float* dbias_prev, int N,
                              const a2c_conv_shape* s, int precision, void* workspace, long long workspace_bytes,
                              void* stream);

int a2c_conv2d_bwd_data_relu(const float* dY, const float* Ymask, const float* Wt, const float* X, float* dX,
                             float* dbias_prev, int N, const a2c_conv_shape* s, int precision, void* workspace,
                             long long workspace_bytes, void* stream) {
  return bwd_data_relu_impl(dY, Ymask, Wt, X, nullptr, dX, dbias_prev, N, s, precision, workspace, workspace_bytes,
                            stream);
}

int a2c_conv2d_bwd_data_relu_bits(const float* dY, const float* Ymask, const float* Wt, const float* X,
                                  const unsigned char* Xbits, float* dX, float* dbias_prev, int N,
                                  const a2c_conv_shape* s, int precision, void* workspace,
                                  long long workspace_bytes, void* stream) {
  A2C_ARG_CHECK(Xbits, "conv_bwd_data_relu_bits: null Xbits");
  return bwd_data_relu_impl(dY, Ymask, Wt, X, Xbits, dX, dbias_prev, N, s, precision, workspace, workspace_bytes,
                            stream);
}

static int bwd_data_relu_impl(const float* dY, const float* Ymask, const float* Wt, const float* X,
                              const unsigned char* Xbits, float* dX, float* dbias_prev, int N,
                              const a2c_conv_shape* s, int precision, void* workspace, long long workspace_bytes,
                              void* stream) {
  if (int rc = check_shape(s)) return rc;
  A2C_ARG_CHECK(dY && Wt && X && dX && N >= 0, "conv_bwd_data_relu: bad args");
  A2C_ARG_CHECK(!dbias_prev || (workspace && workspace_bytes >= (long long)sizeof(double) * s->Cin),
                "conv_bwd_data_relu: workspace of Cin doubles needed for dbias_prev");
  cudaStream_t st = as_stream(stream);
  double* acc = dbias_prev ? static_cast<double*>(workspace) : nullptr;
  if (acc) A2C_CUDA_CHECK(cudaMemsetAsync(acc, 0, sizeof(double) * s->Cin, st));
  if (N > 0) {
    int rc = 1;
    if (precision == 1) rc = conv_dgrad_shift_try(dY, Ymask, Wt, dX, X, acc, N, s, st, Xbits);
    if (rc != 0 && rc != 1) return rc;
    if (rc == 0) g_last_path = 0;
    if (rc == 1) {
      // other paths: plain input gradient, then the mask and the channel sums as separate passes
      if (int r2 = a2c_conv2d_bwd_data(dY, Ymask, Wt, dX, N, s, precision, stream)) return r2;
      const long long n = (long long)N * s->Cin * s->H * s->W;
      long long blocks = cdivll(n, 256);
      if (blocks > num_sms() * 16) blocks = num_sms() * 16;
      relu_mask_inplace_kernel<<<(int)blocks, 256, 0, st>>>(dX, X, n);
      g_launches++;
      A2C_LAUNCH_CHECK();
      if (acc)
        if (int r3 = channel_sum_launch(dX, nullptr, acc, N, s->Cin, s->H * s->W, st)) return r3;
    }
  }
  if (acc) {
    f64_to_f32_kernel<<<cdiv(s->Cin, 64), 64, 0, st>>>(acc, dbias_prev, s->Cin);
    g_launches++;
    A2C_LAUNCH_CHECK();
  }
  return 0;
}

int a2c_channel_sum(const float* X, const float* mask, float* out, int N, int C, int HW, void* workspace,
                    long long workspace_bytes, void* stream) {
  A2C_ARG_CHECK(X && out && N >= 0 && C > 0 && HW > 0, "channel_sum: bad args");
  A2C_ARG_CHECK(workspace && workspace_bytes >= (long long)sizeof(double) * C, "channel_sum: workspace of C doubles");
  cudaStream_t st = as_stream(stream);
  double* acc = static_cast<double*>(workspace);
  A2C_CUDA_CHECK(cudaMemsetAsync(acc, 0, sizeof(double) * C, st));
  if (N > 0)
    if (int rc = channel_sum_launch(X, mask, acc, N, C, HW, st)) return rc;
  f64_to_f32_kernel<<<cdiv(C, 64), 64, 0, st>>>(acc, out, C);
  g_launches++;
  A2C_LAUNCH_CHECK();
  return 0;
}

int a2c_conv2d_last_path(void) { return g_last_path; }

int a2c_conv2d_fused_mask(const a2c_conv_shape* s, int precision, int input_is_view) {
  if (!s || check_shape(s)) return 0;
  if (precision == 0) return 3;                       // the FFMA kernels always take Ymask
  int bits = 0;
  if (precision == 1 && conv_dgrad_shift_ok(s)) bits |= 1;
  const bool wg_shift = precision == 1 && (!input_is_view || s->Cin <= 4) && conv_wgrad_shift_ok(s);
  if (wg_shift || s->Cin <= 4) bits |= 2;
  if (wg_shift && conv_wgrad_shift_bias_free(s)) bits |= 4;
  return bits;
}

long long a2c_conv2d_bwd_weight_workspace(const a2c_conv_shape* s) {
  if (!s) return 0;
  WgPlan pl;
  if (plan_wgrad(s, &pl)) return 0;
  const long long simt = (long long)pl.grid * pl.PG *
                         ((long long)s->Cout * s->Cin * s->K * s->K + s->Cout) * (long long)sizeof(float);
  long long tcw = conv_wgrad_tc_workspace(s);
  const long long pw = conv_wgrad_patch_workspace(s);
  if (pw > tcw) tcw = pw;
  const long long sw = conv_wgrad_shift_workspace(s);
  if (sw > tcw) tcw = sw;
  return simt > tcw ? simt : tcw;
}

int a2c_conv2d_bwd_weight(const float* X, const a2c_state_view* xview, const float* dY,
                          const float* Ymask, float* dW, float* dbias, int N,
                          const a2c_conv_shape* s, void* workspace, long long workspace_bytes,
                          int precision, void* stream) {
  if (int rc = check_shape(s)) return rc;
  A2C_ARG_CHECK(dY && dW && N >= 0, "conv_bwd_weight: bad args");
  A2C_ARG_CHECK(workspace && workspace_bytes >= a2c_conv2d_bwd_weight_workspace(s),
                "conv_bwd_weight: workspace too small");
  const int wsize = s->Cout * s->Cin * s->K * s->K;
  cudaStream_t st = as_stream(stream);
  if (N == 0) {
    A2C_CUDA_CHECK(cudaMemsetAsync(dW, 0, sizeof(float) * wsize, st));
    if (dbias) A2C_CUDA_CHECK(cudaMemsetAsync(dbias, 0, sizeof(float) * s->Cout, st));
    return 0;
  }
  // 3x3 pad-1 layers run as shift-GEMMs (short TMEM chains, fp32 running sums, double reduction);
  // the older tensor-core paths take the remaining layers with more than 4 input channels.
  if (precision == 1) {
    A2C_PATH_RETURN(0, conv_wgrad_shift_try(X, xview, dY, Ymask, dW, dbias, N, s, workspace, workspace_bytes, st));
  }
  if (precision != 0 && !Ymask && s->Cin > 4) {
    A2C_ARG_CHECK(xview || X, "conv_bwd_weight: null input");
    A2C_PATH_RETURN(1, conv_wgrad_patch_try(X, xview, dY, dW, dbias, N, s, workspace, workspace_bytes, precision, st));
    A2C_PATH_RETURN(1, conv_wgrad_tc_try(X, xview, dY, dW, dbias, N, s, workspace, workspace_bytes, precision, st));
  }
  g_last_path = 2;
  WgPlan pl;
  if (plan_wgrad(s, &pl)) {
    set_error("conv_bwd_weight: unsupported shape");
    return A2C_ERR_ARG;
  }
  ConvWgParams p{};
  if (int rc = fill_src(&p.src, X, xview, s)) return rc;
  p.dY = dY; p.Ymask = Ymask; p.partial = static_cast<float*>(workspace);
  p.N = N; p.Cin = s->Cin; p.Cout = s->Cout; p.OH = s->OH; p.OW = s->OW; p.pad = s->pad;
  p.rows = pl.rows; p.bands = pl.bands; p.WP = pl.WP; p.PR = pl.PR; p.CoutP = pl.CoutP;
  p.PG = pl.PG; p.pairs = pl.pairs;
  const int K = s->K, S = s->stride;
  if (K == 3 && S == 1 && pl.COT == 4) return launch_wgrad_t<3, 1, 4>(p, pl, wsize, dW, dbias, st);
  if (K == 3 && S == 2 && pl.COT == 4) return launch_wgrad_t<3, 2, 4>(p, pl, wsize, dW, dbias, st);
  if (K == 3 && S == 1 && pl.COT == 8) return launch_wgrad_t<3, 1, 8>(p, pl, wsize, dW, dbias, st);
  if (K == 3 && S == 2 && pl.COT == 8) return launch_wgrad_t<3, 2, 8>(p, pl, wsize, dW, dbias, st);
  if (K == 4 && S == 2) return launch_wgrad_t<4, 2, 4>(p, pl, wsize, dW, dbias, st);
  if (K == 8 && S == 4) return launch_wgrad_t<8, 4, 1>(p, pl, wsize, dW, dbias, st);
  set_error("conv_bwd_weight: unsupported kernel/stride K=%d S=%d", K, S);
  return A2C_ERR_ARG;
}

}  // extern "C"

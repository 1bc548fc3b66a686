// Implicit-GEMM convolutions on the tcgen05 tile engine (tc_core.cuh): forward,
// backward-data and backward-weight of nn.Conv2d (a2c/models.py:98,312,685),
// NCHW fp32 (first layer: also the u8 plane ring of stage 1).  Nothing is
// im2col'ed in HBM: the producer warps gather the operand tile straight into
// the swizzled UMMA layout, lanes running along the contiguous pixel dimension
// so every global load instruction is coalesced, and all loads of a k-block are
// issued before the first conversion / shared store (latency is overlapped
// inside a thread, not only across warps).
//
//   forward        D[pixel, co]   = sum_{ci,ky,kx} X[n,ci,oy*s+ky-p,ox*s+kx-p] * W[co,ci,ky,kx]
//   backward-data  D[inpixel, ci] = sum_{co,ky,kx} dY[n,co,(iy+p-ky)/s,(ix+p-kx)/s] * W[co,ci,ky,kx]
//   backward-weight D[(ci,ky,kx), co] = sum_{pixel} X[...] * dY[n,co,oy,ox]   (split over pixels)
// dY is expected ReLU-masked already (a2c_relu_bwd).
#include <atomic>

#include "tc_core.cuh"

namespace a2c {
extern std::atomic<long long> g_launches;

namespace tc {

struct ConvParams {
  const float* src;        // fwd / wgrad: X [N,Cin,H,W];  dgrad: dY [N,Cout,OH,OW]
  a2c_state_view view;     // first layer input (fwd / wgrad) when use_view
  int use_view;
  const float* wt;         // [Cout][Cin][K][K]
  const float* dy;         // wgrad: dY
  const float* bias;
  float* out;              // fwd: Y, dgrad: dX, wgrad: partial [splits][mrows][ncols]
  int N, Cin, Cout, H, W, OH, OW, K, S, P, logS, relu;
  int red;                 // reduction length (fwd Cin*K*K, dgrad Cout*K*K, wgrad N*OH*OW)
  long long rows;          // D rows (fwd N*OH*OW, dgrad N*H*W, wgrad Cin*K*K)
  int mrows, ncols;        // wgrad partial pitch
  int nkb, kb_per_split;
};

// packed tap table entry: delta[0,20) ky[20,24) kx[24,28) valid[28] c[29,32)
__device__ __forceinline__ uint32_t pack_tap(int delta, int ky, int kx, int valid, int c) {
  return (uint32_t)delta | ((uint32_t)ky << 20) | ((uint32_t)kx << 24) | ((uint32_t)valid << 28) |
         ((uint32_t)(c & 7) << 29);
}

__device__ __forceinline__ uint32_t view_chanmask(const a2c_state_view& v, long long r, int t) {
  uint32_t m = 0xffffffffu;
  if (v.pflag) {
    const uint8_t* pf = v.pflag + r * v.pflag_stride_r + t;
    m = 0;
    bool ok = true;
    for (int c = v.fs - 1; c >= 0; --c) {   // channel c valid iff no reset flag in planes c+1 .. fs-1
      if (ok) m |= 1u << c;
      ok = ok && (pf[c] == 0);
    }
  }
  return m;
}

// Source of the gathered X operand: fp32 NCHW, or the stage-1 state view (u8 / f32).
struct XRow {
  const uint8_t* p8;
  const float* p32;
  int iy0, ix0;
  uint32_t cmask;
  bool ok;
};

__device__ __forceinline__ void xrow_setup(const ConvParams& p, bool ok, int n, int iy0, int ix0, XRow& r) {
  r.ok = ok; r.iy0 = iy0; r.ix0 = ix0; r.cmask = 0xffffffffu; r.p8 = nullptr; r.p32 = nullptr;
  const long long sp = (long long)iy0 * p.W + ix0;
  if (p.use_view) {
    const long long rr = n / p.view.T;
    const int tt = n - (int)rr * p.view.T;
    const long long base = rr * p.view.stride_r + (long long)tt * p.view.stride_t + sp;
    r.cmask = view_chanmask(p.view, rr, tt);
    if (p.view.dtype == 0) r.p8 = static_cast<const uint8_t*>(p.view.base) + base;
    else r.p32 = static_cast<const float*>(p.view.base) + base;
  } else {
    r.p32 = p.src + (long long)n * p.Cin * p.H * p.W + sp;
  }
}

__device__ __forceinline__ float xrow_load(const ConvParams& p, const XRow& r, uint32_t e) {
  const int ky = (e >> 20) & 15, kx = (e >> 24) & 15;
  const bool ok = r.ok && ((e >> 28) & 1) && (unsigned)(r.iy0 + ky) < (unsigned)p.H &&
                  (unsigned)(r.ix0 + kx) < (unsigned)p.W && ((r.cmask >> (e >> 29)) & 1);
  if (!ok) return 0.0f;
  const int d = (int)(e & 0xFFFFF);
  return r.p8 ? (float)r.p8[d] : __ldg(r.p32 + d);
}

template <int NV>
struct Frag {
  float v[NV];
};

// weights tile (B of forward / backward-data): BN x 32; unit u -> row = u % BN, chunk = u / BN so
// that a quarter-warp stores 8 consecutive rows of one chunk (conflict-free); addr(row, k) by functor
template <int BN, int NT, class F>
__device__ __forceinline__ void load_weights(Frag<BN / 4>& f, int t, F addr) {
#pragma unroll
  for (int i = 0; i < BN * 8 / NT; ++i) {
    const int u = i * NT + t, row = u % BN, c = u / BN;
#pragma unroll
    for (int j = 0; j < 4; ++j) f.v[i * 4 + j] = addr(row, c * 4 + j);
  }
}
template <int BN, bool SPLIT3, int NT>
__device__ __forceinline__ void store_weights(const Frag<BN / 4>& f, uint8_t* hi, uint8_t* lo, int t) {
#pragma unroll
  for (int i = 0; i < BN * 8 / NT; ++i) {
    const int u = i * NT + t, row = u % BN, c = u / BN;
    store_chunk<SPLIT3>(hi, lo, row, c, make_float4(f.v[i * 4], f.v[i * 4 + 1], f.v[i * 4 + 2], f.v[i * 4 + 3]));
  }
}

// ------------------------------------------------------------------ forward --------
template <int BN, bool SPLIT3>
struct ConvFwdProblem {
  using Params = ConvParams;
  static constexpr int NA = 8, NB = 4, KC = 8, DEPTH = 1;
  using AState = XRow;
  static __device__ __forceinline__ void init(const Params& p, uint8_t* scratch, int tid, int nt) {
    uint32_t* tab = reinterpret_cast<uint32_t*>(scratch);
    const int KK = p.K * p.K;
    const int plane = p.use_view ? (int)p.view.stride_c : p.H * p.W;
    for (int k = tid; k < p.nkb * KB; k += nt) {
      const int ci = k / KK, r = k - ci * KK, ky = r / p.K, kx = r - ky * p.K;
      tab[k] = pack_tap(ci * plane + ky * p.W + kx, ky, kx, k < p.red, ci);
    }
  }
  // thread t: row = t & 127 (a pixel), chunks 4*(t>>7) .. +3 of every k-block
  static __device__ __forceinline__ void a_setup(const Params& p, const uint8_t*, int m0, int t, AState& st) {
    const long long m = (long long)m0 + (t & 127);
    const bool ok = m < p.rows;
    const int ohw = p.OH * p.OW;
    const int n = ok ? (int)(m / ohw) : 0;
    const int pix = ok ? (int)(m - (long long)n * ohw) : 0;
    const int oy = pix / p.OW, ox = pix - oy * p.OW;
    xrow_setup(p, ok, n, oy * p.S - p.P, ox * p.S - p.P, st);
  }
  using AFrag = Frag<16>;
  using BFrag = Frag<BN / 4>;
  struct BState {};
  static __device__ __forceinline__ void b_setup(const Params&, const uint8_t*, int, int, BState&) {}
  static __device__ __forceinline__ void load_a(const Params& p, const AState& st, const uint8_t* scratch,
                                                int, int kb, int t, AFrag& f) {
    const uint32_t* tab = reinterpret_cast<const uint32_t*>(scratch) + kb * KB + (t >> 7) * 16;
#pragma unroll
    for (int j = 0; j < 16; ++j) f.v[j] = xrow_load(p, st, tab[j]);
  }
  static __device__ __forceinline__ void store_a(const Params&, const AState&, const AFrag& f, uint8_t* hi, uint8_t* lo, int t) {
    const int row = t & 127, c0 = (t >> 7) * 4;
#pragma unroll
    for (int c = 0; c < 4; ++c)
      store_chunk<SPLIT3>(hi, lo, row, c0 + c,
                          make_float4(f.v[c * 4], f.v[c * 4 + 1], f.v[c * 4 + 2], f.v[c * 4 + 3]));
  }
  static __device__ __forceinline__ void load_b(const Params& p, const BState&, const uint8_t*, int n0, int kb, int t, BFrag& f) {
    load_weights<BN, 32 * NB>(f, t, [&](int row, int kk) {
      const int co = n0 + row, k = kb * KB + kk;
      return (co < p.Cout && k < p.red) ? __ldg(p.wt + (long long)co * p.red + k) : 0.0f;
    });
  }
  static __device__ __forceinline__ void store_b(const Params&, const BState&, const BFrag& f, uint8_t* hi, uint8_t* lo, int t) {
    store_weights<BN, SPLIT3, 32 * NB>(f, hi, lo, t);
  }
  template <int HALF>
  static __device__ __forceinline__ void epilogue(const Params& p, int mrow, int nb, const float* acc, int) {
    const long long m = mrow;
    if (m >= p.rows) return;
    const int ohw = p.OH * p.OW;
    const int n = (int)(m / ohw);
    const int pix = (int)(m - (long long)n * ohw);
    float* dst = p.out + (long long)n * p.Cout * ohw + pix;
#pragma unroll
    for (int j = 0; j < HALF; ++j) {
      const int co = nb + j;
      if (co < p.Cout) {
        float x = acc[j] + (p.bias ? p.bias[co] : 0.0f);
        if (p.relu) x = fmaxf(x, 0.0f);
        dst[(long long)co * ohw] = x;     // lanes = consecutive pixels: coalesced per channel
      }
    }
  }
};

// ------------------------------------------------------------ backward data --------
struct DgradRow {
  const float* ptr;   // dY + n*Cout*OH*OW
  int ty0, tx0;
  bool ok;
};

template <int BN, bool SPLIT3>
struct ConvDgradProblem {
  using Params = ConvParams;
  static constexpr int NA = 8, NB = 4, KC = 8, DEPTH = 1;
  using AState = DgradRow;
  static __device__ __forceinline__ void init(const Params& p, uint8_t* scratch, int tid, int nt) {
    uint32_t* tab = reinterpret_cast<uint32_t*>(scratch);
    const int KK = p.K * p.K;
    for (int k = tid; k < p.nkb * KB; k += nt) {
      const int co = k / KK, r = k - co * KK, ky = r / p.K, kx = r - ky * p.K;
      tab[k] = pack_tap(co * p.OH * p.OW, ky, kx, k < p.red, 0);
    }
  }
  static __device__ __forceinline__ void a_setup(const Params& p, const uint8_t*, int m0, int t, AState& st) {
    const long long m = (long long)m0 + (t & 127);
    st.ok = m < p.rows;
    const int hw = p.H * p.W;
    const int n = st.ok ? (int)(m / hw) : 0;
    const int pix = st.ok ? (int)(m - (long long)n * hw) : 0;
    const int iy = pix / p.W, ix = pix - iy * p.W;
    st.ty0 = iy + p.P; st.tx0 = ix + p.P;
    st.ptr = p.src + (long long)n * p.Cout * p.OH * p.OW;
  }
  using AFrag = Frag<16>;
  using BFrag = Frag<BN / 4>;
  struct BState {};
  static __device__ __forceinline__ void b_setup(const Params&, const uint8_t*, int, int, BState&) {}
  static __device__ __forceinline__ void load_a(const Params& p, const AState& st, const uint8_t* scratch,
                                                int, int kb, int t, AFrag& f) {
    const uint32_t* tab = reinterpret_cast<const uint32_t*>(scratch) + kb * KB + (t >> 7) * 16;
    const int smask = p.S - 1;
    float* v = f.v;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const uint32_t e = tab[j];
      const int ty = st.ty0 - (int)((e >> 20) & 15), tx = st.tx0 - (int)((e >> 24) & 15);
      const int oy = ty >> p.logS, ox = tx >> p.logS;
      const bool ok = st.ok && ((e >> 28) & 1) && ty >= 0 && tx >= 0 && !((ty | tx) & smask) &&
                      oy < p.OH && ox < p.OW;
      v[j] = ok ? __ldg(st.ptr + (int)(e & 0xFFFFF) + oy * p.OW + ox) : 0.0f;
    }
  }
  static __device__ __forceinline__ void store_a(const Params&, const AState&, const AFrag& f, uint8_t* hi, uint8_t* lo, int t) {
    const int row = t & 127, c0 = (t >> 7) * 4;
#pragma unroll
    for (int c = 0; c < 4; ++c)
      store_chunk<SPLIT3>(hi, lo, row, c0 + c,
                          make_float4(f.v[c * 4], f.v[c * 4 + 1], f.v[c * 4 + 2], f.v[c * 4 + 3]));
  }
  static __device__ __forceinline__ void load_b(const Params& p, const BState&, const uint8_t*, int n0, int kb, int t, BFrag& f) {
    // B[ci][k = (co,ky,kx)] = wt[((co*Cin + ci)*K + ky)*K + kx]
    const int KK = p.K * p.K;
    load_weights<BN, 32 * NB>(f, t, [&](int row, int kk) {
      const int ci = n0 + row, k = kb * KB + kk;
      const int co = k / KK, r = k - co * KK;
      return (ci < p.Cin && k < p.red) ? __ldg(p.wt + ((long long)co * p.Cin + ci) * KK + r) : 0.0f;
    });
  }
  static __device__ __forceinline__ void store_b(const Params&, const BState&, const BFrag& f, uint8_t* hi, uint8_t* lo, int t) {
    store_weights<BN, SPLIT3, 32 * NB>(f, hi, lo, t);
  }
  template <int HALF>
  static __device__ __forceinline__ void epilogue(const Params& p, int mrow, int nb, const float* acc, int) {
    const long long m = mrow;
    if (m >= p.rows) return;
    const int hw = p.H * p.W;
    const int n = (int)(m / hw);
    const int pix = (int)(m - (long long)n * hw);
    float* dst = p.out + (long long)n * p.Cin * hw + pix;
#pragma unroll
    for (int j = 0; j < HALF; ++j) {
      const int ci = nb + j;
      if (ci < p.Cin) dst[(long long)ci * hw] = acc[j];
    }
  }
};

// ---------------------------------------------------------- backward weight --------
template <int BN, bool SPLIT3>
struct ConvWgradProblem {
  using Params = ConvParams;
  // short TMEM chains: the reduction runs over millions of cancelling terms, and the
  // accumulator's truncation is a bias, not noise (see tc_core.cuh)
  static constexpr int NA = 8, NB = 4, KC = 2, DEPTH = 1;
  struct AState {};
  static __device__ __forceinline__ void init(const Params& p, uint8_t* scratch, int tid, int nt) {
    uint32_t* tab = reinterpret_cast<uint32_t*>(scratch);   // one entry per tap row of the layer
    const int KK = p.K * p.K;
    const int plane = p.use_view ? (int)p.view.stride_c : p.H * p.W;
    for (int m = tid; m < p.mrows; m += nt) {
      const int ci = m / KK, r = m - ci * KK, ky = r / p.K, kx = r - ky * p.K;
      tab[m] = pack_tap(ci * plane + ky * p.W + kx, ky, kx, m < (int)p.rows, ci);
    }
  }
  static __device__ __forceinline__ void a_setup(const Params&, const uint8_t*, int, int, AState&) {}
  // lane <-> pixel of the k-block; warp w walks over its 16 tap rows
  using AFrag = Frag<16>;
  using BFrag = Frag<BN / 4>;
  struct BState {};
  static __device__ __forceinline__ void b_setup(const Params&, const uint8_t*, int, int, BState&) {}
  static __device__ __forceinline__ void load_a(const Params& p, const AState&, const uint8_t* scratch,
                                                int m0, int kb, int t, AFrag& f) {
    const int lane = t & 31, w = t >> 5;
    const uint32_t* tab = reinterpret_cast<const uint32_t*>(scratch) + m0 + w * 16;
    const long long px = (long long)kb * KB + lane;
    const bool pok = px < p.red;
    const int ohw = p.OH * p.OW;
    const int n = pok ? (int)(px / ohw) : 0;
    const int pix = pok ? (int)(px - (long long)n * ohw) : 0;
    const int oy = pix / p.OW, ox = pix - oy * p.OW;
    XRow xr;
    xrow_setup(p, pok, n, oy * p.S - p.P, ox * p.S - p.P, xr);
#pragma unroll
    for (int rr = 0; rr < 16; ++rr) f.v[rr] = xrow_load(p, xr, tab[rr]);
  }
  static __device__ __forceinline__ void store_a(const Params&, const AState&, const AFrag& f, uint8_t* hi, uint8_t* lo, int t) {
    const int lane = t & 31, w = t >> 5;
#pragma unroll
    for (int rr = 0; rr < 16; ++rr) store_elem<SPLIT3>(hi, lo, w * 16 + rr, lane, f.v[rr]);
  }
  static __device__ __forceinline__ void load_b(const Params& p, const BState&, const uint8_t*, int n0, int kb, int t, BFrag& f) {
    const int lane = t & 31, w = t >> 5;
    const long long px = (long long)kb * KB + lane;
    const bool pok = px < p.red;
    const int ohw = p.OH * p.OW;
    const int n = pok ? (int)(px / ohw) : 0;
    const int pix = pok ? (int)(px - (long long)n * ohw) : 0;
    const float* src = p.dy + (long long)n * p.Cout * ohw + pix;
    constexpr int RW = BN / NB;
#pragma unroll
    for (int rr = 0; rr < RW; ++rr) {
      const int co = n0 + w * RW + rr;
      f.v[rr] = (pok && co < p.Cout) ? __ldg(src + (long long)co * ohw) : 0.0f;
    }
  }
  static __device__ __forceinline__ void store_b(const Params&, const BState&, const BFrag& f, uint8_t* hi, uint8_t* lo, int t) {
    const int lane = t & 31, w = t >> 5;
    constexpr int RW = BN / NB;
#pragma unroll
    for (int rr = 0; rr < RW; ++rr) store_elem<SPLIT3>(hi, lo, w * RW + rr, lane, f.v[rr]);
  }
  template <int HALF>
  static __device__ __forceinline__ void epilogue(const Params& p, int mrow, int nb, const float* acc,
                                                  int split) {
    float* dst = p.out + ((long long)split * p.mrows + mrow) * p.ncols + nb;
#pragma unroll
    for (int j = 0; j < HALF; j += 4)
      *reinterpret_cast<float4*>(dst + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
  }
};

// dW[co][tap] = sum_split partial[split][tap][co]   (double accumulation)
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int mrows, int ncols, int taps,
                    int cout, float* __restrict__ dW) {
  const int total = taps * cout;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int tap = e / cout, co = e - tap * cout;
    double s = 0.0;
    for (int z = 0; z < splits; ++z) s += (double)partial[((long long)z * mrows + tap) * ncols + co];
    dW[(long long)co * taps + tap] = (float)s;
  }
}

// per-channel sum over N and pixels (bias gradient): X [N][C][HW]
__global__ void __launch_bounds__(256)
channel_sum_kernel(const float* __restrict__ X, double* __restrict__ acc, int N, int C, int HW) {
  __shared__ double scratch[32];
  const int c = blockIdx.y;
  double s = 0.0;
  const long long per = (long long)N * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / HW);
    const int px = (int)(i - (long long)n * HW);
    s += (double)X[((long long)n * C + c) * HW + px];
  }
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) atomicAdd(&acc[c], s);
}
__global__ void channel_sum_finish_kernel(const double* __restrict__ acc, float* __restrict__ out, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) out[c] = (float)acc[c];
}

static int ilog2(int s) { int l = 0; while ((1 << l) < s) ++l; return l; }

static bool conv_tc_ok(const a2c_conv_shape* s) {
  if (s->K > 15 || (s->stride & (s->stride - 1))) return false;
  const long long plane = (long long)s->H * s->W;
  if ((long long)s->Cin * plane + (long long)s->K * s->W + s->K >= (1 << 20)) return false;
  if ((long long)s->Cout * s->OH * s->OW >= (1 << 20)) return false;
  return true;
}

template <template <int, bool> class Prob>
static int launch_conv(const ConvParams& p, int ncols_needed, int mtiles, int splits, bool split3,
                       cudaStream_t st) {
  const int bn = ncols_needed > 32 ? 64 : 32;
  const int ntiles = cdiv(ncols_needed, bn);
  int rc;
  if (split3) {
    rc = bn == 64 ? tc_launch<64, true, Prob<64, true>>(p, mtiles, ntiles, splits, st)
                  : tc_launch<32, true, Prob<32, true>>(p, mtiles, ntiles, splits, st);
  } else {
    rc = bn == 64 ? tc_launch<64, false, Prob<64, false>>(p, mtiles, ntiles, splits, st)
                  : tc_launch<32, false, Prob<32, false>>(p, mtiles, ntiles, splits, st);
  }
  if (rc) return rc;
  g_launches++;
  A2C_LAUNCH_CHECK();
  return 0;
}

static void fill_common(ConvParams* p, const a2c_conv_shape* s, int N) {
  memset(p, 0, sizeof(*p));
  p->N = N; p->Cin = s->Cin; p->Cout = s->Cout; p->H = s->H; p->W = s->W; p->OH = s->OH; p->OW = s->OW;
  p->K = s->K; p->S = s->stride; p->P = s->pad; p->logS = ilog2(s->stride);
}

}  // namespace tc

// Each returns 0 = ran on tensor cores, 1 = shape not handled (caller uses the FFMA kernel).
int conv_fwd_tc_try(const float* X, const a2c_state_view* xview, const float* Wt, const float* bias,
                    float* Y, int N, const a2c_conv_shape* s, int relu, int precision, cudaStream_t st) {
  if (!tc::conv_tc_ok(s)) return 1;
  const int red = s->Cin * s->K * s->K;
  if (cdiv(red, tc::KB) * tc::KB * 4 > tc::SCRATCH_BYTES) return 1;
  tc::ConvParams p;
  tc::fill_common(&p, s, N);
  p.src = X; p.wt = Wt; p.bias = bias; p.out = Y; p.relu = relu;
  if (xview) {
    if (xview->stride_c >= (1 << 17)) return 1;
    p.view = *xview; p.use_view = 1;
  }
  p.red = red;
  p.rows = (long long)N * s->OH * s->OW;
  p.nkb = cdiv(red, tc::KB); p.kb_per_split = p.nkb;
  const long long mt = cdivll(p.rows, tc::TM);
  if (mt > 0x7fffffff / tc::TM) return 1;
  return tc::launch_conv<tc::ConvFwdProblem>(p, s->Cout, (int)mt, 1, precision == 1, st);
}

int conv_dgrad_tc_try(const float* dY, const float* Wt, float* dX, int N, const a2c_conv_shape* s,
                      int precision, cudaStream_t st) {
  if (!tc::conv_tc_ok(s)) return 1;
  const int red = s->Cout * s->K * s->K;
  if (cdiv(red, tc::KB) * tc::KB * 4 > tc::SCRATCH_BYTES) return 1;
  tc::ConvParams p;
  tc::fill_common(&p, s, N);
  p.src = dY; p.wt = Wt; p.out = dX;
  p.red = red;
  p.rows = (long long)N * s->H * s->W;
  p.nkb = cdiv(red, tc::KB); p.kb_per_split = p.nkb;
  const long long mt = cdivll(p.rows, tc::TM);
  if (mt > 0x7fffffff / tc::TM) return 1;
  return tc::launch_conv<tc::ConvDgradProblem>(p, s->Cin, (int)mt, 1, precision == 1, st);
}

static void wgrad_plan(const a2c_conv_shape* s, int* mtiles, int* bn, int* ntiles, int* splits) {
  const int taps = s->Cin * s->K * s->K;
  *mtiles = cdiv(taps, tc::TM);
  *bn = s->Cout > 32 ? 64 : 32;
  *ntiles = cdiv(s->Cout, *bn);
  int sp = num_sms() / (*mtiles * *ntiles);
  *splits = sp < 1 ? 1 : sp;
}

long long conv_wgrad_tc_workspace(const a2c_conv_shape* s) {
  int mtiles, bn, ntiles, splits;
  wgrad_plan(s, &mtiles, &bn, &ntiles, &splits);
  return (long long)splits * mtiles * tc::TM * ntiles * bn * 4 + (long long)s->Cout * 8;
}

int conv_wgrad_tc_try(const float* X, const a2c_state_view* xview, const float* dY, float* dW,
                      float* dbias, int N, const a2c_conv_shape* s, void* workspace,
                      long long workspace_bytes, int precision, cudaStream_t st) {
  if (!tc::conv_tc_ok(s)) return 1;
  const int taps = s->Cin * s->K * s->K;
  int mtiles, bn, ntiles, splits;
  wgrad_plan(s, &mtiles, &bn, &ntiles, &splits);
  if (mtiles * tc::TM * 4 > tc::SCRATCH_BYTES) return 1;
  const long long npix = (long long)N * s->OH * s->OW;
  if (npix >= (1ll << 31) - 64) return 1;
  if (workspace_bytes < conv_wgrad_tc_workspace(s)) return 1;
  tc::ConvParams p;
  tc::fill_common(&p, s, N);
  p.src = X; p.dy = dY;
  if (xview) {
    if (xview->stride_c >= (1 << 17)) return 1;
    p.view = *xview; p.use_view = 1;
  }
  p.red = (int)npix;
  p.rows = taps;
  p.nkb = (int)cdivll(npix, tc::KB);
  int kps = cdiv(cdiv(p.nkb, splits), 8) * 8;     // whole accumulation chunks per split
  splits = cdiv(p.nkb, kps);
  p.kb_per_split = kps;
  p.mrows = mtiles * tc::TM;
  p.ncols = ntiles * bn;
  double* bias_acc = reinterpret_cast<double*>(workspace);
  p.out = reinterpret_cast<float*>(bias_acc + s->Cout);
  int rc;
  const bool split3 = precision == 1;
  if (split3) {
    rc = bn == 64 ? tc::tc_launch<64, true, tc::ConvWgradProblem<64, true>>(p, mtiles, ntiles, splits, st)
                  : tc::tc_launch<32, true, tc::ConvWgradProblem<32, true>>(p, mtiles, ntiles, splits, st);
  } else {
    rc = bn == 64 ? tc::tc_launch<64, false, tc::ConvWgradProblem<64, false>>(p, mtiles, ntiles, splits, st)
                  : tc::tc_launch<32, false, tc::ConvWgradProblem<32, false>>(p, mtiles, ntiles, splits, st);
  }
  if (rc) return rc;
  g_launches++;
  A2C_LAUNCH_CHECK();
  const int total = taps * s->Cout;
  tc::wgrad_reduce_kernel<<<cdiv(total, 256), 256, 0, st>>>(p.out, splits, p.mrows, p.ncols, taps, s->Cout, dW);
  g_launches++;
  A2C_LAUNCH_CHECK();
  if (dbias) {
    A2C_CUDA_CHECK(cudaMemsetAsync(bias_acc, 0, sizeof(double) * s->Cout, st));
    const int hw = s->OH * s->OW;
    int bx = (int)cdivll((long long)N * hw, 256 * 8);
    if (bx > 256) bx = 256;
    if (bx < 1) bx = 1;
    tc::channel_sum_kernel<<<dim3(bx, s->Cout), 256, 0, st>>>(dY, bias_acc, N, s->Cout, hw);
    tc::channel_sum_finish_kernel<<<cdiv(s->Cout, 64), 64, 0, st>>>(bias_acc, dbias, s->Cout);
    g_launches += 2;
    A2C_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace a2c

// Explicit im2col / col2im for the convolutions that have no shift-GEMM form (A3CModel's 8x8
// stride-4 and 4x4 stride-2 layers, a2c/models.py:22-38, 98): the patch matrix is written to HBM
// once per layer and every contraction of the layer (forward, weight gradient, input gradient)
// then runs on the tcgen05 GEMM (a2c_gemm):
//
//     col [N*OH*OW, Kc]      Y  = relu(col W^T + b)          [N*OH*OW, Cout]   (channels-last)
//     dW^T [Kc, Cout]        = col^T dY                      (split-K over the N*OH*OW rows)
//     dcol [N*OH*OW, Kc]     = dY W        ->  dX = col2im(dcol) * (X > 0)
//
// These layers are HBM-bound (10.7 MFLOP per state against ~0.26 MB of patch matrix), so the
// cost is the patch-matrix traffic: 3.8 GB written once and read twice for the first A3C layer at
// C3.  Activations between such layers are kept channels-last ([N, H, W, C]) so that the GEMM
// output is the next layer's input without a transpose.
//
// Column orders (they must match the flattening of the weight the GEMM is given):
//   ORDER_CKK  k = (ci*K + ky)*K + kx : the reference's  W.view(Cout, Cin*K*K)  as is -- used for
//              planar sources (the uint8 plane ring / fp32 states of the first layer);
//   ORDER_KKC  k = (ky*K + kx)*Cin + ci : channels fastest, so channels-last sources are read and
//              gradients scattered in 16-byte pieces; the host permutes the (tiny) weight.
#include <atomic>

#include "common.cuh"

namespace a2c {
extern std::atomic<long long> g_launches;

struct Im2colParams {
  a2c_state_view view;      // planar source (use_view)
  int use_view;
  const float* x;           // channels-last source f32 [N, H, W, C]
  float* col;
  const float* dcol;        // col2im: f32 [N*OH*OW, Kc] in ORDER_KKC
  const float* xmask;       // col2im: optional, channels-last like dx: dx = xmask > 0 ? dx : 0
  float* dx;                // col2im: f32 [N, H, W, C]
  int N, C, H, W, OH, OW, K, S, P;
};

// planar source -> ORDER_CKK, 4 consecutive kx per thread (K % 4 == 0, S % 4 == 0 or P-aligned not
// required: loads are per element unless the uint8 fast path applies)
__global__ void __launch_bounds__(256)
im2col_planar_kernel(const Im2colParams p) {
  const int K4 = p.K >> 2;
  const int Kc = p.C * p.K * p.K;
  const long long rows = (long long)p.N * p.OH * p.OW;
  const long long total = rows * (Kc >> 2);
  const a2c_state_view& v = p.view;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    const long long row = g / (Kc >> 2);
    int q = (int)(g - row * (Kc >> 2));              // float4 index inside the row
    const int kx0 = (q % K4) << 2;
    q /= K4;
    const int ky = q % p.K, ci = q / p.K;
    const long long n = row / (p.OH * p.OW);
    const int r = (int)(row - n * (p.OH * p.OW));
    const int oy = r / p.OW, ox = r - oy * p.OW;
    const int y = oy * p.S + ky - p.P, x0 = ox * p.S + kx0 - p.P;
    // state n = (slot rr, step tt); planes older than a reset read as zero (a2c/utils.py:37-40)
    const long long rr = n / v.T;
    const int tt = (int)(n - rr * v.T);
    bool live = true;
    if (v.pflag) {
      const uint8_t* pf = v.pflag + rr * v.pflag_stride_r + tt;
      for (int j = ci + 1; j < v.fs; ++j) live = live && (pf[j] == 0);
    }
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live && (unsigned)y < (unsigned)p.H) {
      const long long base = rr * v.stride_r + (long long)tt * v.stride_t + (long long)ci * v.stride_c + (long long)y * p.W;
      float e[4];
      if (v.dtype == 0 && x0 >= 0 && x0 + 3 < p.W && ((base + x0) & 3) == 0) {
        // uint8 ring, 4 aligned pixels: one 32-bit load
        const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(static_cast<const uint8_t*>(v.base) + base + x0));
        e[0] = (float)(w & 0xff); e[1] = (float)((w >> 8) & 0xff); e[2] = (float)((w >> 16) & 0xff); e[3] = (float)(w >> 24);
      } else
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int x = x0 + j;
        float val = 0.0f;
        if ((unsigned)x < (unsigned)p.W)
          val = v.dtype == 0 ? (float)static_cast<const uint8_t*>(v.base)[base + x]
                             : static_cast<const float*>(v.base)[base + x];
        e[j] = val;
      }
      o = make_float4(e[0], e[1], e[2], e[3]);
    }
    *reinterpret_cast<float4*>(p.col + (row * Kc + ((long long)(ci * p.K + ky) * p.K + kx0))) = o;
  }
}

// Same result, one block per (state n, output row oy): the C x K source rows the output row needs
// are staged in shared memory as fp32 once (5 KB for the first A3C layer), then the OW x Kc patch
// rows are written with coalesced 128-bit stores and no per-element index arithmetic beyond one
// small division.  Used when W % 4 == 0, S % 4 == 0, P % 4 == 0 and the tile fits 48 KB; the
// generic kernel above covers everything else.  (Measured, first A3C layer at C3, 3.8 GB of patch
// matrix: generic 2.6 ms, this one 2.1 ms.  A uint8 patch matrix -- 4x fewer bytes -- was tried and
// dropped: neither this kernel nor the GEMM that consumes it is bandwidth-bound.)
__global__ void __launch_bounds__(256)
im2col_planar_rows_kernel(const Im2colParams p) {
  extern __shared__ float tile[];                    // [C][K][W]
  const a2c_state_view& v = p.view;
  const int Kc = p.C * p.K * p.K, K4 = p.K >> 2, W4 = p.W >> 2;
  const long long nrow_blocks = (long long)p.N * p.OH;
  for (long long blk = blockIdx.x; blk < nrow_blocks; blk += gridDim.x) {
    const long long n = blk / p.OH;
    const int oy = (int)(blk - n * p.OH);
    const long long rr = n / v.T;
    const int tt = (int)(n - rr * v.T);
    __syncthreads();                                 // previous tile fully consumed
    for (int u = threadIdx.x; u < p.C * p.K * W4; u += blockDim.x) {
      const int x4 = u % W4;
      const int ck = u / W4;
      const int ky = ck % p.K, ci = ck / p.K;
      const int y = oy * p.S + ky - p.P;
      bool live = (unsigned)y < (unsigned)p.H;
      if (live && v.pflag) {
        const uint8_t* pf = v.pflag + rr * v.pflag_stride_r + tt;
        for (int j = ci + 1; j < v.fs; ++j) live = live && (pf[j] == 0);
      }
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live) {
        const long long a = rr * v.stride_r + (long long)tt * v.stride_t + (long long)ci * v.stride_c +
                            (long long)y * p.W + 4 * x4;
        if (v.dtype == 0) {
          const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(static_cast<const uint8_t*>(v.base) + a));
          o = make_float4((float)(w & 0xff), (float)((w >> 8) & 0xff), (float)((w >> 16) & 0xff), (float)(w >> 24));
        } else {
          o = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(v.base) + a));
        }
      }
      reinterpret_cast<float4*>(tile)[u] = o;
    }
    __syncthreads();
    const long long out_off = ((n * p.OH + oy) * (long long)p.OW) * Kc;
    const int per_row = Kc >> 2;                     // 4-element groups per patch row
    for (int u = threadIdx.x; u < p.OW * per_row; u += blockDim.x) {
      const int ox = u / per_row, q = u - ox * per_row;
      const int kx0 = (q % K4) << 2, ck = q / K4;    // ck = ci*K + ky
      const int x0 = ox * p.S + kx0 - p.P;           // multiple of 4
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      if ((unsigned)x0 < (unsigned)p.W) o = reinterpret_cast<const float4*>(tile)[ck * W4 + (x0 >> 2)];
      reinterpret_cast<float4*>(p.col + out_off)[u] = o;
    }
  }
}

// channels-last source -> ORDER_KKC, 4 consecutive channels per thread (C % 4 == 0)
__global__ void __launch_bounds__(256)
im2col_nhwc_kernel(const Im2colParams p) {
  const int C4 = p.C >> 2;
  const int Kc = p.C * p.K * p.K;
  const long long rows = (long long)p.N * p.OH * p.OW;
  const long long total = rows * (Kc >> 2);
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    const long long row = g / (Kc >> 2);
    int q = (int)(g - row * (Kc >> 2));
    const int c0 = (q % C4) << 2;
    q /= C4;
    const int kx = q % p.K, ky = q / p.K;
    const long long n = row / (p.OH * p.OW);
    const int r = (int)(row - n * (p.OH * p.OW));
    const int oy = r / p.OW, ox = r - oy * p.OW;
    const int y = oy * p.S + ky - p.P, x = ox * p.S + kx - p.P;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if ((unsigned)y < (unsigned)p.H && (unsigned)x < (unsigned)p.W)
      o = __ldg(reinterpret_cast<const float4*>(p.x + (((n * p.H + y) * p.W + x) * p.C + c0)));
    *reinterpret_cast<float4*>(p.col + (row * Kc + ((long long)(ky * p.K + kx) * p.C + c0))) = o;
  }
}

// dX[n, y, x, c] = sum over taps (ky, kx) that map an output (oy, ox) onto (y, x) of
// dcol[(n, oy, ox), (ky, kx, c)], gated by xmask > 0.  4 channels per thread (gather, no atomics).
__global__ void __launch_bounds__(256)
col2im_nhwc_kernel(const Im2colParams p) {
  const int C4 = p.C >> 2;
  const int Kc = p.C * p.K * p.K;
  const long long total = (long long)p.N * p.H * p.W * C4;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(g % C4) << 2;
    long long t = g / C4;
    const int x = (int)(t % p.W);
    t /= p.W;
    const int y = (int)(t % p.H);
    const long long n = t / p.H;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int ky = (y + p.P) % p.S; ky < p.K; ky += p.S) {
      const int oy = (y + p.P - ky) / p.S;
      if (y + p.P - ky < 0 || oy >= p.OH) continue;
      for (int kx = (x + p.P) % p.S; kx < p.K; kx += p.S) {
        const int ox = (x + p.P - kx) / p.S;
        if (x + p.P - kx < 0 || ox >= p.OW) continue;
        const float4 d = __ldg(reinterpret_cast<const float4*>(
            p.dcol + (((n * p.OH + oy) * p.OW + ox) * Kc + ((long long)(ky * p.K + kx) * p.C + c0))));
        acc.x += d.x; acc.y += d.y; acc.z += d.z; acc.w += d.w;
      }
    }
    const long long off = ((n * p.H + y) * p.W + x) * p.C + c0;
    if (p.xmask) {
      const float4 m = __ldg(reinterpret_cast<const float4*>(p.xmask + off));
      acc.x = m.x > 0.f ? acc.x : 0.f; acc.y = m.y > 0.f ? acc.y : 0.f;
      acc.z = m.z > 0.f ? acc.z : 0.f; acc.w = m.w > 0.f ? acc.w : 0.f;
    }
    *reinterpret_cast<float4*>(p.dx + off) = acc;
  }
}

static int ew_grid(long long n) {
  long long b = cdivll(n, 256);
  const long long cap = (long long)num_sms() * 32;
  return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace a2c

using namespace a2c;

extern "C" {

int a2c_im2col(const a2c_state_view* xview, const float* x_nhwc, float* col, int N, const a2c_conv_shape* s,
               void* stream) {
  A2C_ARG_CHECK(s && col && N >= 0 && (xview != nullptr) != (x_nhwc != nullptr),
                "im2col: give exactly one of xview (planar source) / x_nhwc (channels-last source)");
  A2C_ARG_CHECK(s->OH == (s->H - s->K + 2 * s->pad) / s->stride + 1 && s->OW == (s->W - s->K + 2 * s->pad) / s->stride + 1,
                "im2col: OH/OW inconsistent with H/W/K/stride/pad");
  if (N == 0) return 0;
  Im2colParams p{};
  p.col = col; p.N = N; p.C = s->Cin; p.H = s->H; p.W = s->W; p.OH = s->OH; p.OW = s->OW;
  p.K = s->K; p.S = s->stride; p.P = s->pad;
  const long long total = (long long)N * s->OH * s->OW * ((long long)s->Cin * s->K * s->K / 4);
  if (xview) {
    A2C_ARG_CHECK(xview->base && xview->fs == s->Cin && xview->T > 0 && (xview->dtype == 0 || xview->dtype == 1),
                  "im2col: bad state view");
    A2C_ARG_CHECK((s->K & 3) == 0, "im2col: planar sources need K %% 4 == 0");
    p.view = *xview; p.use_view = 1;
    const size_t tile_bytes = (size_t)s->Cin * s->K * s->W * sizeof(float);
    const bool aligned = (s->W & 3) == 0 && (s->stride & 3) == 0 && (s->pad & 3) == 0 &&
                         ((xview->stride_r | xview->stride_t | xview->stride_c) & 3) == 0 &&
                         (reinterpret_cast<uintptr_t>(xview->base) & (xview->dtype == 0 ? 3 : 15)) == 0;
    if (aligned && tile_bytes <= 48 * 1024) {
      long long blocks = (long long)N * s->OH;
      const long long cap = (long long)num_sms() * 8;
      if (blocks > cap) blocks = cap;
      im2col_planar_rows_kernel<<<(int)blocks, 256, tile_bytes, as_stream(stream)>>>(p);
    } else {
      im2col_planar_kernel<<<ew_grid(total), 256, 0, as_stream(stream)>>>(p);
    }
  } else {
    A2C_ARG_CHECK((s->Cin & 3) == 0 && (reinterpret_cast<uintptr_t>(x_nhwc) & 15) == 0,
                  "im2col: channels-last sources need Cin %% 4 == 0 and a 16-byte aligned pointer");
    p.x = x_nhwc;
    im2col_nhwc_kernel<<<ew_grid(total), 256, 0, as_stream(stream)>>>(p);
  }
  g_launches++;
  A2C_LAUNCH_CHECK();
  return 0;
}

int a2c_col2im_nhwc(const float* dcol, const float* xmask, float* dx, int N, const a2c_conv_shape* s,
                    void* stream) {
  A2C_ARG_CHECK(s && dcol && dx && N >= 0, "col2im: bad args");
  A2C_ARG_CHECK((s->Cin & 3) == 0 && (reinterpret_cast<uintptr_t>(dcol) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(dx) & 15) == 0 &&
                    (!xmask || (reinterpret_cast<uintptr_t>(xmask) & 15) == 0),
                "col2im: Cin %% 4 == 0 and 16-byte aligned pointers required");
  if (N == 0) return 0;
  Im2colParams p{};
  p.dcol = dcol; p.xmask = xmask; p.dx = dx; p.N = N; p.C = s->Cin; p.H = s->H; p.W = s->W;
  p.OH = s->OH; p.OW = s->OW; p.K = s->K; p.S = s->stride; p.P = s->pad;
  col2im_nhwc_kernel<<<ew_grid((long long)N * s->H * s->W * (s->Cin / 4)), 256, 0, as_stream(stream)>>>(p);
  g_launches++;
  A2C_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"

// The reference's other prep_fxn callables (a2c/preprocessing.py:4-9,19-32) as batched device
// kernels: variants of the stage-1 observation kernel for the non-Pong environments.  They are
// elementwise and HBM-bound; outputs are bit-identical to the numpy arithmetic of the reference:
//   normalize_prep : 3 * (pic - 255/2) / (255/2) in float64 (numpy promotes uint8 - float to
//                    float64), shape kept, leading axis added by the host wrapper   (:4-6)
//   snake_prep     : colour classes of an RGB frame -> float32 plane, later rules override
//                    earlier ones (ch0 == 1 -> 1, ch0 == 255 -> 1.5, ch1 == 255 -> 0,
//                    ch2 == 255 -> .33)                                              (:25-32)
//   breakout_prep  : rows 35:195, columns 8:-8, every 2nd row / column, channel 0; the
//                    reference then calls skimage's rgb2grey on that 2-D array, which the
//                    skimage releases that still export `rgb2grey` return unchanged    (:19-23)
#include <atomic>

#include "common.cuh"

namespace a2c {
extern std::atomic<long long> g_launches;

__global__ void __launch_bounds__(256)
normalize_prep_kernel(const uint8_t* __restrict__ in, double* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __ddiv_rn(__dmul_rn(3.0, __dsub_rn((double)in[i], 127.5)), 127.5);
}

__global__ void __launch_bounds__(256)
snake_prep_kernel(const uint8_t* __restrict__ in, float* __restrict__ out, long long npix, int C) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
    const uint8_t* px = in + i * C;
    float v = 0.0f;
    if (px[0] == 1) v = 1.0f;
    if (px[0] == 255) v = 1.5f;
    if (px[1] == 255) v = 0.0f;
    if (px[2] == 255) v = 0.33f;
    out[i] = v;
  }
}

__global__ void __launch_bounds__(256)
breakout_prep_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, long long nframes, int H, int W,
                     int C, int OH, int OW) {
  const long long total = nframes * OH * OW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long f = i / (OH * OW);
    const int r = (int)(i - f * OH * OW);
    const int oy = r / OW, ox = r - oy * OW;
    out[i] = in[((f * H + 35 + 2 * oy) * W + 8 + 2 * ox) * C];
  }
}

static int ew_grid(long long n) {
  long long b = cdivll(n, 256);
  const long long cap = (long long)num_sms() * 16;
  return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace a2c

using namespace a2c;

extern "C" {

int a2c_prep_other(const uint8_t* frames, void* out, int kind, long long n_frames, int H, int W, int C,
                   void* stream) {
  A2C_ARG_CHECK(frames && out && n_frames >= 0 && H > 0 && W > 0 && C > 0, "prep_other: bad args");
  if (n_frames == 0) return 0;
  cudaStream_t st = as_stream(stream);
  const long long npix = n_frames * H * W;
  switch (kind) {
    case 1:
      normalize_prep_kernel<<<ew_grid(npix * C), 256, 0, st>>>(frames, static_cast<double*>(out), npix * C);
      break;
    case 2:
      A2C_ARG_CHECK(C >= 3, "snake_prep needs an RGB frame");
      snake_prep_kernel<<<ew_grid(npix), 256, 0, st>>>(frames, static_cast<float*>(out), npix, C);
      break;
    case 3: {
      A2C_ARG_CHECK(H >= 195 && W > 16, "breakout_prep needs a frame of at least 195 x 17");
      const int OH = 80, OW = (W - 16 + 1) / 2;
      breakout_prep_kernel<<<ew_grid(n_frames * OH * OW), 256, 0, st>>>(frames, static_cast<uint8_t*>(out), n_frames,
                                                                      H, W, C, OH, OW);
      break;
    }
    default:
      set_error("prep_other: kind must be 1 (normalize), 2 (snake) or 3 (breakout)");
      return A2C_ERR_ARG;
  }
  g_launches++;
  A2C_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"

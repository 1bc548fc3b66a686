// Shift-GEMM convolutions on tcgen05: no im2col at all.
//
// A band of input rows is converted once (NCHW fp32 -> hi/lo TF32 split) into shared
// memory in the "interleaved pixel layout"
//      element (channel group g, position q, channel c&3)  at  g*PLANE + q*16 + (c&3)*4  bytes
// where q = row*PW + col runs over the zero-padded band.  In that layout the 128 x 8 operand
// of  D[position, co] += X[position + shift(tap), ci] * W[tap][co][ci]  for ANY tap is just the
// band itself read through a no-swizzle K-major UMMA descriptor (8 positions x 16 B core
// matrices, SBO = 128 B, LBO = PLANE) whose start address is advanced by shift(tap)*16 bytes.
// So every input element is loaded, split and stored once and each tap is a descriptor add --
// the MMA warp issues  taps x (C/8) x 3  instructions per 128 output positions, nothing else
// touches the operand.  Output positions are enumerated in the PADDED row pitch PW (the few
// columns >= the real width are computed and dropped), which is what makes a tap a pure shift.
//
//   stride-2 forward : the band is stored as 4 parity planes (y&1, x&1); tap (ky,kx) reads
//                      plane (ky&1, kx&1) at shift (ky>>1, kx>>1)
//   backward-data    : one pass per output parity class over the dY band (stride-1
//                      correlation with the taps of that class); the ReLU mask of the layer
//                      (Ymask > 0) is applied while the band is converted, so the masked
//                      gradient never exists in HBM
//
// Weights for all taps stay resident in shared memory, hi and lo copies side by side along N:
// the 3xTF32 product  x_hi*w_hi + x_hi*w_lo + x_lo*w_hi  is issued as TWO instructions,
//     D[:, 0:2N] (+)= X_hi * [W_hi | W_lo]   (one MMA of width 2N)    and    D[:, 0:N] += X_lo * W_hi,
// and the drain adds the two column halves.  A small-N MMA is bound by the 128 x 32 B read of
// its A operand from shared memory, whatever N is, so two MMAs instead of
// three cut the tensor core's shared-memory traffic per product from 3 x (4 KB + B) to
// 2 x 4 KB + 1.5 B.  TMEM accumulators add with truncation (tc_core.cuh), so the taps of a tile
// are split into `nch` short chains that the drain warps sum in fp32 registers.
//
// Replaces nn.Conv2d forward / input-gradient of a2c/models.py:98,312,685 for 3x3 pad-1 layers.
#include <stdlib.h>

#include <atomic>

#include <type_traits>

#include "tc_core.cuh"

namespace a2c {
extern std::atomic<long long> g_launches;

namespace tc {

constexpr int SH_MAXTAPS = 16;
constexpr int SH_PROD = 12;                                   // producer warps
constexpr int SH_NISS = 4;                                    // MMA issuer warps (tiles round-robin)
constexpr int SH_DRAIN0 = SH_PROD, SH_MMA = SH_DRAIN0 + NUM_EPI_WARPS, SH_ALLOC = SH_MMA + SH_NISS;
constexpr int SH_THREADS = 32 * (SH_ALLOC + 1);               // 25 warps
constexpr int SH_MAXSLOT = 4;                                 // TMEM accumulator slots
static_assert((SH_NISS & (SH_NISS - 1)) == 0, "issuer count must be a power of two");
static_assert(SH_DRAIN0 % 4 == 0, "drain warps must start at a multiple of 4");

struct ShiftTap {
  uint32_t a_off16;       // class offset + shift, in positions (= 16-byte units)
  uint32_t woff;          // offset of the tap inside one (n, k) weight slice
};
struct ShiftClass {       // one output parity class
  int tap0, ntaps, a, b, qh, qw;
};

struct ShiftParams {
  const float* src;       // f32 [N][C][SH][SW], SW % 4 == 0
  const float* mask;      // optional, same shape: source element counts only where mask > 0
  a2c_state_view view;    // first-layer source: uint8 plane ring (use_view)
  int use_view;
  int cvalid;             // real source channels (first layers: C is padded to 8, channels >= cvalid are zero)
  int nsplit;             // 2: the source has a lo part (3 MMAs per product); 1: exact in TF32 ({0,1} ring) -> 2 MMAs
  int N, C, SH, SW, ncg, ksteps;
  int scalar_w;           // SW % 4 != 0 (10x10 / 5x5 planes): one pixel per lane, scalar loads
  int cs;                 // input class step (1, or 2 for the stride-2 forward)
  int p0y, p0x;           // plane (row r, col c, class cy,cx) <-> source (r*cs + cy - p0y, c*cs + cx - p0x)
  int PW;                 // class plane width
  int rext_y[2];          // halo rows of the planes of class row cy
  int cls_pos[4];         // first position of class (cy, cx) inside a channel-group plane
  int plane_pos;          // positions per channel-group plane over all classes (multiple of 8)
  int QH, Rb, nbands;     // band grid
  int noc, os;            // output classes, output step
  ShiftClass oc[4];
  ShiftTap taps[SH_MAXTAPS];
  int ntaps_total, nch;
  const float* wt;        // W(n, k, tap) = wt[n*wsn + k*wsk + woff(tap)]
  int wsn, wsk, J, n0;    // J output channels in total, this launch produces [n0, n0 + NPAD)
  const float* bias;
  int relu;
  float* out;
  const float* omask;     // optional, same shape as out: out = omask > 0 ? value : 0  (the ReLU of the layer
                          // that produced this layer's input, applied where its gradient is born)
  // 1-bit ReLU masks, u8 [N, J / 8, OutH * OutW]: bit (c & 7) of byte [n, c >> 3, pos] = out[n, c, pos] > 0.
  // obits (forward, relu): written next to `out`; omaskbits (MASKOUT): read instead of the fp32 omask
  // (one byte per 8 channels and position instead of eight strided floats: the masked epilogue of a
  // stride-2 input gradient spent 2 ms of its 4.2 ms on those loads).
  uint8_t* obits;
  const uint8_t* omaskbits;
  double* obias;          // optional f64 [J]: per-channel sums of the (masked) output = bias gradient of that layer
  int OutH, OutW;
  int half_bytes, buf_bytes, w_half_bytes, nbuf;
  int dbg;                // A2C_SHIFT_DBG, stage isolation for profiling: 2 = no MMAs, 4 = no loads, 8 = no stores
};

// no-swizzle K-major descriptor: 8-row groups SBO bytes apart, the two 16-byte k chunks LBO apart
__device__ __forceinline__ uint64_t make_desc_nosw(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

template <int NC>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t* r);
template <>
__device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
template <>
__device__ __forceinline__ void tmem_ld<8>(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr));
}

__host__ __device__ constexpr int tmem_cols_pow2(int n) { return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512; }

struct BandInfo {
  int img, r0, rows, ntile;
};
__device__ __forceinline__ BandInfo band_info(const ShiftParams& p, int item) {
  BandInfo b;
  b.img = item / p.nbands;
  const int bi = item - b.img * p.nbands;
  b.r0 = bi * p.Rb;
  b.rows = min(p.Rb, p.QH - b.r0);
  b.ntile = (b.rows * p.PW + TM - 1) / TM;
  return b;
}

// (v[r], v[r+1], v[r+2], v[r+3]) with indices mod 4
__device__ __forceinline__ float4 rot4(float4 v, int r) {
  const float4 a = (r & 1) ? make_float4(v.y, v.z, v.w, v.x) : v;
  return (r & 2) ? make_float4(a.z, a.w, a.x, a.y) : a;
}

__device__ __forceinline__ void store_split(uint8_t* dst, int half_bytes, float a, float b, float c, float d) {
  const float4 h = make_float4(tf32_rna(a), tf32_rna(b), tf32_rna(c), tf32_rna(d));
  *reinterpret_cast<float4*>(dst) = h;
  // (lo rounded to TF32 as well: see tc_core.cuh store_chunk)
  *reinterpret_cast<float4*>(dst + half_bytes) =
      make_float4(tf32_rna(a - h.x), tf32_rna(b - h.y), tf32_rna(c - h.z), tf32_rna(d - h.w));
}

// FIRST: first-layer variant (<= 4 real channels padded to one k-step, optional uint8 ring source,
// optional 2-product split); a template flag so the general kernel's hot loops stay branch-free.
// MASKOUT: backward-data variant whose epilogue applies the ReLU mask of the layer below and accumulates
// obias: 1 = fp32 mask (omask), 2 = 1-bit mask (omaskbits; one register instead of NCOL -- with both forms
// in one kernel the running bias sums spilled to local memory and the kernel ran 3x slower).
template <int NPAD, bool FIRST, int MASKOUT = 0>
__global__ void __launch_bounds__(SH_THREADS, 1) conv_shift_kernel(const __grid_constant__ ShiftParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int NCOL = NPAD / 2;                      // accumulator columns per drain warp
  constexpr int LDN = NCOL >= 16 ? 16 : 8;
  constexpr int WBLK = 2 * NPAD * 32;                 // bytes of one (tap, kstep) weight block: [kchunk][hi n | lo n][4]
  constexpr int NACC = 2 * NPAD;                      // TMEM columns of one accumulator chain (hi | lo halves)
  uint8_t* base = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  uint8_t* wsm = base;                                // [tap][kstep][kchunk][hi n | lo n][4]
  uint8_t* bufs = wsm + 2 * p.w_half_bytes;           // nbuf band buffers, each [hi | lo][cg][class][pos][4]
  uint64_t* bars = reinterpret_cast<uint64_t*>(bufs + (size_t)p.nbuf * p.buf_bytes + 4096);  // 4 KB overrun slack
  uint64_t* patch_full = bars;                        // [2]
  uint64_t* patch_empty = bars + 2;                   // [2]
  uint64_t* acc_full = bars + 4;                      // [SH_MAXSLOT]
  uint64_t* acc_empty = bars + 4 + SH_MAXSLOT;        // [SH_MAXSLOT]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 + 2 * SH_MAXSLOT);
  float* bias_s = reinterpret_cast<float*>(tmem_slot + 4);   // [NPAD]

  // shuffle-broadcast: tells the compiler the warp index (and everything derived from it) is
  // warp-uniform, so role branches are uniform and the MMA loop can live in uniform registers
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int nitems = p.N * p.nbands;
  // accumulator slots: a tile (nch chains of NPAD columns) per slot, tile i uses slot i % nslot
  const uint32_t nslot = (SH_MAXSLOT * p.nch * NACC <= 512) ? SH_MAXSLOT : 2;
  const uint32_t tcols = (uint32_t)tmem_cols_pow2(nslot * p.nch * NACC);

  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(&patch_full[b], SH_PROD);
      mbar_init(&patch_empty[b], SH_NISS);
    }
    for (int b = 0; b < SH_MAXSLOT; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], NUM_EPI_WARPS);
    }
    mbar_fence_init();
  }
  if (warp == SH_ALLOC) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(tcols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // resident weights: block (tap g, kstep ks) = [kchunk][n][4], k = ks*8 + kchunk*4 + e
  {
    const int units = p.ntaps_total * p.ksteps * 2 * NPAD;
    for (int u = threadIdx.x; u < units; u += SH_THREADS) {
      const int n = u % NPAD;
      int r = u / NPAD;
      const int kc = r & 1;
      r >>= 1;
      const int ks = r % p.ksteps, g = r / p.ksteps;
      float v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = ks * 8 + kc * 4 + e;
        v[e] = (p.n0 + n < p.J && k < p.cvalid)
                   ? __ldg(p.wt + (long long)(p.n0 + n) * p.wsn + (long long)k * p.wsk + p.taps[g].woff)
                   : 0.0f;
      }
      store_split(wsm + (size_t)(g * p.ksteps + ks) * WBLK + kc * (NACC * 16) + n * 16, NPAD * 16, v[0], v[1],
                  v[2], v[3]);
    }
    for (int n = threadIdx.x; n < NPAD; n += SH_THREADS)
      bias_s[n] = (p.bias && p.n0 + n < p.J) ? __ldg(p.bias + p.n0 + n) : 0.0f;
    // the pad columns of the planes are never written by the producers: clear the buffers once
    const int nz = p.nbuf * p.buf_bytes / 16;
    for (int i = threadIdx.x; i < nz; i += SH_THREADS) reinterpret_cast<uint4*>(bufs)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp < SH_PROD) {
    // ---------------- producers: source band -> interleaved hi/lo planes ----------------
    // A lane converts 4 consecutive pixels of 4 channels: 4 LDG.128 in, the 4x4 block is
    // transposed in registers (free) and leaves as 4 + 4 STS.128 (hi, lo).  A source row is
    // NX4 float4 wide, so 32 / NX4 (row, channel group) tasks share a warp.
    const int NX4 = p.scalar_w ? p.SW : p.SW >> 2;             // lanes per source row
    const int tpw = 32 / NX4;                                  // tasks per warp pass
    const int sub = lane / NX4, m = lane - sub * NX4;
    const bool lane_on = sub < tpw;
    const int plane = p.SH * p.SW;                             // elements per channel
    StageRing ring(p.nbuf);
    // L2 prefetch of a band's source rows (and mask rows), one item ahead, spread over the producer lanes
    auto prefetch_band = [&](int nitem) {
      if (FIRST || p.use_view) return;                      // the uint8 ring is 1/6400th of the traffic
      const BandInfo nb = band_info(p, nitem);
      const int sy0 = nb.r0 * p.cs - p.p0y;
      const int nrow = (nb.rows + p.rext_y[0]) * p.cs;
      const int lines = (p.SW * 4 + 127) >> 7;               // 128-byte lines per source row (+1 when unaligned)
      const int nseg = nrow * p.C * (lines + 1);
      for (int sg = warp * 32 + lane; sg < nseg; sg += SH_PROD * 32) {
        const int ln = sg % (lines + 1);
        const int rc = sg / (lines + 1);
        const int r = rc / p.C, c = rc - r * p.C;
        const int sy = sy0 + r;
        if ((unsigned)sy >= (unsigned)p.SH) continue;
        const long long off = ((long long)nb.img * p.C + c) * plane + sy * p.SW + min(ln * 32, p.SW - 1);
        prefetch_l2(p.src + off);
        if (p.mask) prefetch_l2(p.mask + off);
      }
    };
    if ((int)blockIdx.x < nitems) prefetch_band(blockIdx.x);
    for (int item = blockIdx.x; item < nitems; item += gridDim.x, ring.next()) {
      const BandInfo bd = band_info(p, item);
      uint8_t* buf = bufs + (size_t)ring.s * p.buf_bytes;
      if (item + (int)gridDim.x < nitems) prefetch_band(item + gridDim.x);
      mbar_wait_relaxed(&patch_empty[ring.s], ring.par ^ 1);   // long wait: do not burn issue slots
      const float4* simg = reinterpret_cast<const float4*>(p.src + (long long)bd.img * p.C * plane);
      // first layers: element offset of channel 0, channel stride and live-plane mask of this image
      long long vbase = (long long)bd.img * p.cvalid * plane, vcstride = plane;
      uint32_t vmask = 0xfu;
      if (p.use_view) {
        const long long rr = bd.img / p.view.T;
        const int tt = bd.img - (int)rr * p.view.T;
        vbase = rr * p.view.stride_r + (long long)tt * p.view.stride_t;
        vcstride = p.view.stride_c;
        if (p.view.pflag) {                         // planes older than a reset read as zero (a2c/utils.py:37-40)
          const uint8_t* pf = p.view.pflag + rr * p.view.pflag_stride_r + tt;
          vmask = 0;
          bool live = true;
          for (int c = p.view.fs - 1; c >= 0; --c) {
            if (live) vmask |= 1u << c;
            live = live && (pf[c] == 0);
          }
        }
      }
      const float4* mimg = p.mask ? reinterpret_cast<const float4*>(p.mask + (long long)bd.img * p.C * plane) : nullptr;
      for (int cy = 0; cy < p.cs && !(p.dbg & 4); ++cy) {
        const int ntask = (bd.rows + p.rext_y[cy]) * p.ncg;    // (r, cg)
        for (int task = warp * tpw + sub; task < ntask; task += SH_PROD * tpw) {
          if (!lane_on) continue;
          const int r = task / p.ncg, cg = task - r * p.ncg;
          const int sy = (bd.r0 + r) * p.cs + cy - p.p0y;
          const bool rowok = (unsigned)sy < (unsigned)p.SH;
          if (!FIRST && p.scalar_w) {
            // narrow planes whose rows are not 16-byte multiples (GRUModel's 10x10 and 5x5 layers):
            // lane m handles pixel x = m of 4 channels with scalar loads; consecutive lanes store
            // consecutive 16-byte units (conflict-free)
            float c4[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const long long a = (long long)bd.img * p.C * plane + (long long)(cg * 4 + c) * plane + sy * p.SW + m;
              float v = rowok ? __ldg(p.src + a) : 0.0f;
              if (rowok && p.mask) v = __ldg(p.mask + a) > 0.f ? v : 0.f;
              c4[c] = v;
            }
            const int xq = m + p.p0x;
            uint8_t* d = buf + ((size_t)cg * p.plane_pos + r * p.PW) * 16;
            if (p.cs == 1) d += (size_t)(p.cls_pos[0] + xq) * 16;
            else d += (size_t)(p.cls_pos[cy * 2 + (xq & 1)] + (xq >> 1)) * 16;
            store_split(d, p.half_bytes, c4[0], c4[1], c4[2], c4[3]);
            continue;
          }
          float4 v0, v1, v2, v3;
          int e0 = 0, pl4 = 0;
          if constexpr (FIRST) {
            // first layer: cvalid (<= 4) channels, f32 tensor or uint8 ring view; the second channel group is
            // padding that stays zero from the initial clear
            if (cg != 0) continue;
            float4 vv[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              vv[c] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (rowok && c < p.cvalid && ((vmask >> c) & 1u)) {
                const long long a = vbase + c * vcstride + (long long)sy * p.SW + 4 * m;
                if (p.use_view) {
                  const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(static_cast<const uint8_t*>(p.view.base) + a));
                  vv[c] = make_float4((float)(w & 0xff), (float)((w >> 8) & 0xff), (float)((w >> 16) & 0xff), (float)(w >> 24));
                } else {
                  vv[c] = __ldg(reinterpret_cast<const float4*>(p.src + a));
                }
              }
            }
            v0 = vv[0]; v1 = vv[1]; v2 = vv[2]; v3 = vv[3];
          } else {
          e0 = (rowok ? (cg * 4 * plane + sy * p.SW) >> 2 : 0) + m;   // float4 index, channel 0
          pl4 = rowok ? plane >> 2 : 0;
          v0 = __ldg(simg + e0); v1 = __ldg(simg + e0 + pl4); v2 = __ldg(simg + e0 + 2 * pl4);
          v3 = __ldg(simg + e0 + 3 * pl4);
          }
          if (mimg) {
            const float4 m0 = __ldg(mimg + e0), m1 = __ldg(mimg + e0 + pl4), m2 = __ldg(mimg + e0 + 2 * pl4),
                         m3 = __ldg(mimg + e0 + 3 * pl4);
            v0.x = m0.x > 0.f ? v0.x : 0.f; v0.y = m0.y > 0.f ? v0.y : 0.f; v0.z = m0.z > 0.f ? v0.z : 0.f; v0.w = m0.w > 0.f ? v0.w : 0.f;
            v1.x = m1.x > 0.f ? v1.x : 0.f; v1.y = m1.y > 0.f ? v1.y : 0.f; v1.z = m1.z > 0.f ? v1.z : 0.f; v1.w = m1.w > 0.f ? v1.w : 0.f;
            v2.x = m2.x > 0.f ? v2.x : 0.f; v2.y = m2.y > 0.f ? v2.y : 0.f; v2.z = m2.z > 0.f ? v2.z : 0.f; v2.w = m2.w > 0.f ? v2.w : 0.f;
            v3.x = m3.x > 0.f ? v3.x : 0.f; v3.y = m3.y > 0.f ? v3.y : 0.f; v3.z = m3.z > 0.f ? v3.z : 0.f; v3.w = m3.w > 0.f ? v3.w : 0.f;
          }
          if (!rowok) v0 = v1 = v2 = v3 = make_float4(0.f, 0.f, 0.f, 0.f);
          uint8_t* dst = buf + ((size_t)cg * p.plane_pos + r * p.PW) * 16;
          const int xq = 4 * m + p.p0x;                        // plane x' of pixel 0
          if (p.cs == 1) {
            // In step j the lane stores pixel (j + rot) & 3 of its 4-pixel block: with every lane on the
            // same pixel index the 8 lanes of a quarter-warp hit only 2 of the 8 16-byte bank groups
            // (ncu: ~70 % of the shared wavefronts were conflicts); rotated, their 16-byte units are 8
            // consecutive residues mod 8.
            const int rot = (m >> 1) & 3;
            v0 = rot4(v0, rot); v1 = rot4(v1, rot); v2 = rot4(v2, rot); v3 = rot4(v3, rot);
            dst += (size_t)(p.cls_pos[0] + xq) * 16;
            store_split(dst + ((0 + rot) & 3) * 16, p.half_bytes, v0.x, v1.x, v2.x, v3.x);
            store_split(dst + ((1 + rot) & 3) * 16, p.half_bytes, v0.y, v1.y, v2.y, v3.y);
            store_split(dst + ((2 + rot) & 3) * 16, p.half_bytes, v0.z, v1.z, v2.z, v3.z);
            store_split(dst + ((3 + rot) & 3) * 16, p.half_bytes, v0.w, v1.w, v2.w, v3.w);
          } else {
            const int* cp = p.cls_pos + cy * 2;
            store_split(dst + (size_t)(cp[xq & 1] + (xq >> 1)) * 16, p.half_bytes, v0.x, v1.x, v2.x, v3.x);
            store_split(dst + (size_t)(cp[(xq + 1) & 1] + ((xq + 1) >> 1)) * 16, p.half_bytes, v0.y, v1.y, v2.y, v3.y);
            store_split(dst + (size_t)(cp[xq & 1] + ((xq + 2) >> 1)) * 16, p.half_bytes, v0.z, v1.z, v2.z, v3.z);
            store_split(dst + (size_t)(cp[(xq + 1) & 1] + ((xq + 3) >> 1)) * 16, p.half_bytes, v0.w, v1.w, v2.w, v3.w);
          }
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&patch_full[ring.s]);
    }
  } else if (warp >= SH_MMA && warp < SH_MMA + SH_NISS) {
    // Several issuer warps: one warp sustains roughly one small tcgen05.mma per ~100 cycles when
    // its operands go through vector registers; the loop below lives entirely in uniform
    // registers (about 3 instructions per MMA) and tiles are dealt round-robin to SH_NISS warps.
    const uint32_t me = (uint32_t)(warp - SH_MMA);
    const uint32_t niss = nslot < (uint32_t)SH_NISS ? nslot : (uint32_t)SH_NISS;   // active issuer warps
    constexpr uint32_t idesc1 = make_idesc(NPAD);             // X_lo * W_hi
    constexpr uint32_t idesc2 = make_idesc(NACC);             // X_hi * [W_hi | W_lo]
    constexpr uint32_t DESC_HI = (128u >> 4) | (1u << 14);    // SBO = 128 B, descriptor version 1
    const uint32_t a_w0 = (uint32_t)make_desc_nosw(smem_u32(bufs), (uint32_t)p.plane_pos * 16u, 128);
    const uint32_t b_w0 = (uint32_t)make_desc_nosw(smem_u32(wsm), NACC * 16, 128);
    // 32-bit arithmetic on the low descriptor word only (address >> 4 in bits 0-13, no carry out)
    const uint32_t buf16 = (uint32_t)p.buf_bytes >> 4, half16 = (uint32_t)p.half_bytes >> 4;
    const uint32_t kstep16 = 2u * (uint32_t)p.plane_pos;      // 8 channels = 2 channel-group planes
    StageRing ring(p.nbuf);
    uint32_t tcount = 0, slot = 0, sphase = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x, ring.next()) {
      const BandInfo bd = band_info(p, item);
      mbar_wait(&patch_full[ring.s], ring.par);
      tc_fence_after();
      const uint32_t a_buf = a_w0 + ring.s * buf16;
      for (int oc = 0; oc < p.noc; ++oc) {
        const int tpc = p.oc[oc].ntaps / p.nch;
        for (int t = 0; t < bd.ntile; ++t, ++tcount) {
          const uint32_t ab = slot, aphase = sphase;
          if (++slot == nslot) { slot = 0; sphase ^= 1u; }
          // An accumulator slot must always be (re)used by the SAME issuer warp: its wait on
          // acc_empty is a parity wait, which cannot tell "two phases behind" from "done", so a
          // warp may only ever be one use ahead of the drain -- true when tile i goes to warp
          // i % niss with niss dividing into the slot rotation (niss = min(SH_NISS, nslot)).
          if ((tcount & (niss - 1)) != me) continue;
          mbar_wait(&acc_empty[ab], aphase ^ 1u);
          tc_fence_after();
          if (!(p.dbg & 2) && elect_one()) {
            int g = p.oc[oc].tap0;
            for (int ch = 0; ch < p.nch; ++ch) {
              const uint32_t tmem_d = tmem_base + (ab * p.nch + ch) * NACC;
              uint32_t acc = 0;
              for (int tp = 0; tp < tpc; ++tp, ++g) {
                uint32_t a_hi = a_buf + p.taps[g].a_off16 + (uint32_t)(t * TM);
                uint32_t b_hi = b_w0 + (uint32_t)(g * p.ksteps) * (WBLK >> 4);
                for (int ks = 0; ks < p.ksteps; ++ks) {
                  mma_tf32_imm<idesc2, DESC_HI>(tmem_d, a_hi, b_hi, acc);
                  // ({0,1} ring sources are exact in TF32: no X_lo term)
                  if (!FIRST || p.nsplit == 2) mma_tf32_imm<idesc1, DESC_HI>(tmem_d, a_hi + half16, b_hi, 1u);
                  acc = 1u;
                  a_hi += kstep16;
                  b_hi += (WBLK >> 4);
                }
              }
            }
          }
          __syncwarp();
          if (elect_one()) mma_commit(&acc_full[ab]);
        }
      }
      if (elect_one()) mma_commit(&patch_empty[ring.s]);
    }
  } else if (warp >= SH_DRAIN0 && warp < SH_DRAIN0 + NUM_EPI_WARPS) {
    // 8 warps per tile: TMEM lane quarter q, column half hsel.  The SM is issue-bound, so the
    // loop is kept lean: incremental (row, col) of the lane's position, 32-bit store offsets.
    const int e = warp - SH_DRAIN0, q = e & 3, hsel = e >> 2;
    const int row = q * 32 + lane;
    const int col0 = hsel * NCOL;
    const int plane = p.OutH * p.OutW;
    const int nvalid = min(NCOL, p.J - p.n0 - col0);
    const float floor_v = p.relu ? 0.0f : -INFINITY;
    const float4* bias4 = reinterpret_cast<const float4*>(bias_s + col0);
    const int q128 = TM / p.PW, r128 = TM - q128 * p.PW;     // position step of one tile in (row, col)
    const int jr0 = row / p.PW, i0 = row - jr0 * p.PW;
    uint32_t slot = 0, sphase = 0;
    float bsum[MASKOUT ? NCOL : 1];
#pragma unroll
    for (int jj = 0; jj < (MASKOUT ? NCOL : 1); ++jj) bsum[jj] = 0.0f;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
      const BandInfo bd = band_info(p, item);
      float* oimg = p.out + ((long long)bd.img * p.J + p.n0 + col0) * plane;
      const float* mimg = MASKOUT == 1 ? p.omask + ((long long)bd.img * p.J + p.n0 + col0) * plane : nullptr;
      for (int oc = 0; oc < p.noc; ++oc) {
        const ShiftClass c = p.oc[oc];
        int jr = jr0, i = i0;
        for (int t = 0; t < bd.ntile; ++t) {
          const uint32_t ab = slot;
          // position of this lane's accumulator row; the mask values of its NCOL channels are
          // requested BEFORE the wait on the accumulator so their latency hides behind the MMAs
          const int j = bd.r0 + jr;
          const bool pvalid = jr < bd.rows && j < c.qh && i < c.qw && !(p.dbg & 8);
          const int idx0 = (j * p.os + c.a) * p.OutW + i * p.os + c.b;
          float mk[MASKOUT == 1 ? NCOL : 1];
          uint32_t mbits = 0;
          if constexpr (MASKOUT == 2) {
            const uint8_t* bimg = p.omaskbits + ((long long)bd.img * (p.J >> 3) + ((p.n0 + col0) >> 3)) * plane;
#pragma unroll
            for (int b8 = 0; b8 < NCOL / 8; ++b8)
              if (pvalid && b8 * 8 < nvalid) mbits |= (uint32_t)__ldg(bimg + (long long)b8 * plane + idx0) << (8 * b8);
          }
          if constexpr (MASKOUT == 1) {
            // the NEXT tile's mask lines go to L2 now (no registers held), so that its loads below hit L2:
            // the drain handles tiles one after the other and would otherwise pay HBM latency per tile
            {
              int jn = jr + q128, in_ = i + r128;
              if (in_ >= p.PW) { in_ -= p.PW; ++jn; }
              const int j2 = bd.r0 + jn;
              if (t + 1 < bd.ntile && jn < bd.rows && j2 < c.qh && in_ < c.qw) {
                const float* a = mimg + (j2 * p.os + c.a) * p.OutW + in_ * p.os + c.b;
#pragma unroll
                for (int jj = 0; jj < NCOL; ++jj)
                  if (jj < nvalid) prefetch_l2(a + jj * plane);
              }
            }
#pragma unroll
            for (int jj = 0; jj < NCOL; ++jj)
              mk[jj] = (pvalid && jj < nvalid) ? __ldg(mimg + idx0 + jj * plane) : 0.0f;
          }
          mbar_wait(&acc_full[ab], sphase);
          if (++slot == nslot) { slot = 0; sphase ^= 1u; }
          tc_fence_after();
          // chain ch: columns [ch*NACC, +NPAD) hold x*w_hi (both x terms), [ch*NACC + NPAD, +NPAD) x_hi*w_lo
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + ab * p.nch * NACC + col0;
          float acc[NCOL];
          if (NCOL <= 16 && p.nch == 3) {
            // three chains in flight per wait: first the main halves, then the w_lo halves
            uint32_t r0[NCOL], r1[NCOL], r2[NCOL];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              const uint32_t ta = taddr + half * NPAD;
              tmem_ld<LDN>(ta, r0);
              tmem_ld<LDN>(ta + NACC, r1);
              tmem_ld<LDN>(ta + 2 * NACC, r2);
              if (NCOL > LDN) {
                tmem_ld<LDN>(ta + LDN, r0 + NCOL - LDN);
                tmem_ld<LDN>(ta + NACC + LDN, r1 + NCOL - LDN);
                tmem_ld<LDN>(ta + 2 * NACC + LDN, r2 + NCOL - LDN);
              }
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
              for (int jj = 0; jj < NCOL; ++jj) {
                const float t = (__uint_as_float(r0[jj]) + __uint_as_float(r1[jj])) + __uint_as_float(r2[jj]);
                acc[jj] = half ? acc[jj] + t : t;
              }
            }
          } else {
#pragma unroll
            for (int jj = 0; jj < NCOL; ++jj) acc[jj] = 0.0f;
            for (int ch = 0; ch < 2 * p.nch; ++ch) {          // (chain, half) pairs: NPAD columns apart
              uint32_t r[NCOL];
#pragma unroll
              for (int c0 = 0; c0 < NCOL; c0 += LDN) tmem_ld<LDN>(taddr + ch * NPAD + (uint32_t)c0, r + c0);
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
              for (int jj = 0; jj < NCOL; ++jj) acc[jj] += __uint_as_float(r[jj]);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[ab]);
          if (pvalid) {
            // WB: also pack the ReLU bits of the stored values (forward layers that leave a 1-bit mask);
            // a separate instantiation so that the other layers do not pay two instructions per element
            // FULL: all NCOL channels of this warp exist (no per-element bound check); the output pointer
            // advances by one plane per channel (the drain is a critical role of these kernels: 64-bit
            // index arithmetic and a compare + branch per element were 2/3 of its instructions)
            auto emit = [&](auto wb_tag, auto full_tag) {
              constexpr bool WB = decltype(wb_tag)::value;
              constexpr bool FULL = decltype(full_tag)::value;
              float* optr = oimg + idx0;           // MASKOUT == 0
              int idx = idx0;                      // MASKOUT != 0 (one register less: those variants are at the limit)
              uint32_t rbits = 0;
#pragma unroll
              for (int j4 = 0; j4 < NCOL / 4; ++j4) {
                const float4 b4 = bias4[j4];
                const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  if (FULL || j4 * 4 + u < nvalid) {
                    float v = fmaxf(acc[j4 * 4 + u] + bb[u], floor_v);
                    if constexpr (MASKOUT != 0) {
                      const bool on = MASKOUT == 1 ? mk[MASKOUT == 1 ? j4 * 4 + u : 0] > 0.0f : ((mbits >> (j4 * 4 + u)) & 1u) != 0;
                      v = on ? v : 0.0f;
                      bsum[j4 * 4 + u] += v;
                    }
                    if constexpr (WB) rbits |= (v > 0.0f ? 1u : 0u) << (j4 * 4 + u);
                    if constexpr (MASKOUT == 0) *optr = v; else oimg[idx] = v;
                  }
                  if constexpr (MASKOUT == 0) optr += plane; else idx += plane;
                }
              }
              if constexpr (WB) {
                uint8_t* bimg = p.obits + ((long long)bd.img * (p.J >> 3) + ((p.n0 + col0) >> 3)) * plane + idx0;
#pragma unroll
                for (int b8 = 0; b8 < NCOL / 8; ++b8)
                  if (b8 * 8 < nvalid) bimg[(long long)b8 * plane] = (uint8_t)(rbits >> (8 * b8));
              }
            };
            if constexpr (MASKOUT != 0) {
              emit(std::false_type{}, std::false_type{});      // (more variants spill the running bias sums)
            } else if (nvalid == NCOL) {
              if (p.obits) emit(std::true_type{}, std::true_type{});
              else emit(std::false_type{}, std::true_type{});
            } else {
              if (p.obits) emit(std::true_type{}, std::false_type{});
              else emit(std::false_type{}, std::false_type{});
            }
          }
          jr += q128; i += r128;
          if (i >= p.PW) { i -= p.PW; ++jr; }
        }
      }
    }
    if constexpr (MASKOUT) {
      // per-thread fp32 running sums (a few thousand terms each) -> warp sum -> one f64 atomic per channel
      if (p.obias) {
#pragma unroll
        for (int jj = 0; jj < NCOL; ++jj) {
          const float t = warp_sum(bsum[jj]);
          if (lane == 0 && jj < nvalid) atomicAdd(p.obias + p.n0 + col0 + jj, (double)t);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == SH_ALLOC) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tcols));
  }
}

// ---------------------------------------------------------------------------- host side ----------
static constexpr int SH_SMEM_LIMIT = 227 * 1024;

static size_t shift_smem_bytes(const ShiftParams& p) {
  return 2 * (size_t)p.w_half_bytes + (size_t)p.nbuf * p.buf_bytes + 4096 + 128 + 512;
}

// positions of one channel-group plane for bands of rb rows; fills cls_pos
static int shift_plane_pos(ShiftParams* p, int rb) {
  int pos = 0;
  for (int cy = 0; cy < p->cs; ++cy)
    for (int cx = 0; cx < p->cs; ++cx) {
      p->cls_pos[cy * p->cs + cx] = pos;
      pos += (rb + p->rext_y[cy]) * p->PW;
    }
  return (pos + 7) & ~7;
}

// Chooses band height and buffer count; false when the layer does not fit.
static bool shift_plan(ShiftParams* p, int npad) {
  p->w_half_bytes = p->ntaps_total * p->ksteps * npad * 32;
  for (int nbuf = 2; nbuf >= 1; --nbuf) {
    const long long avail = (long long)SH_SMEM_LIMIT - 2LL * p->w_half_bytes - 4096 - 640;
    if (avail <= 0) return false;
    int rb = 0;
    for (int cand = p->QH; cand >= 1; --cand)
      if (2LL * p->ncg * shift_plane_pos(p, cand) * 16 * nbuf <= avail) { rb = cand; break; }
    if (rb < 1) continue;
    if (nbuf == 2 && rb < 2 && p->QH > 1) continue;     // a one-row band is mostly halo: single buffer instead
    const int nb = cdiv(p->QH, rb);
    rb = cdiv(p->QH, nb);
    p->Rb = rb; p->nbands = nb; p->nbuf = nbuf;
    p->plane_pos = shift_plane_pos(p, rb);
    p->half_bytes = p->ncg * p->plane_pos * 16;
    p->buf_bytes = 2 * p->half_bytes;
    if (shift_smem_bytes(*p) <= (size_t)SH_SMEM_LIMIT && p->plane_pos * 16 < (1 << 18)) return true;
  }
  return false;
}

template <int NPAD, bool FIRST = false, int MASKOUT = 0>
static int shift_launch(ShiftParams& p, cudaStream_t st) {
  auto kern = conv_shift_kernel<NPAD, FIRST, MASKOUT>;
  A2C_CUDA_CHECK(set_max_dynamic_smem(reinterpret_cast<const void*>(kern), SH_SMEM_LIMIT));
  const int nitems = p.N * p.nbands;
  int grid = num_sms();
  if (grid > nitems) grid = nitems;
  kern<<<grid, SH_THREADS, shift_smem_bytes(p), st>>>(p);
  g_launches++;
  A2C_LAUNCH_CHECK();
  return 0;
}

static int floordiv_i(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }

}  // namespace tc

using tc::ShiftParams;

static bool shift_enabled() {
  static const bool off = debug_knob("A2C_NO_SHIFT") != 0;
  return !off;
}

// Tries output-channel chunks of 64 / 32 / 16 (the resident weights must leave room for the
// bands), fills the tap table for the resulting plan and launches one kernel per chunk.
template <class FillTaps>
static int shift_run(ShiftParams& p, FillTaps fill, cudaStream_t st) {
  const int cands[3] = {64, 32, 16};
  for (int ci = 0; ci < 3; ++ci) {
    const int npad = cands[ci];
    if (npad >= 2 * p.J && npad > 16) continue;                 // do not pad a narrow layer to a wide tile
    if ((p.omask || p.omaskbits) && npad == 64) continue;       // masked epilogue: 16 / 32 wide only
    if (2 * p.nch * 2 * npad > 512) continue;                   // TMEM columns: 2 slots x nch chains x (hi | lo) halves
    if ((long long)p.ntaps_total * p.ksteps * npad * 64 > 80 * 1024 && npad > 16) continue;
    if (!tc::shift_plan(&p, npad)) continue;
    if (!fill(p)) return 1;
    p.dbg = debug_knob("A2C_SHIFT_DBG");
    for (int n0 = 0; n0 < p.J; n0 += npad) {
      p.n0 = n0;
      int rc;
      if (p.cvalid < p.C) {
        if (npad == 64) return 1;                                 // first layers: 16 / 32 wide only
        rc = npad == 32 ? tc::shift_launch<32, true>(p, st) : tc::shift_launch<16, true>(p, st);
      } else if (p.omask || p.omaskbits) {
        if (npad == 64) return 1;                                 // masked epilogue: up to 32 channels per pass
        if (p.omaskbits) rc = npad == 32 ? tc::shift_launch<32, false, 2>(p, st) : tc::shift_launch<16, false, 2>(p, st);
        else rc = npad == 32 ? tc::shift_launch<32, false, 1>(p, st) : tc::shift_launch<16, false, 1>(p, st);
      } else {
        rc = npad == 64 ? tc::shift_launch<64>(p, st) : npad == 32 ? tc::shift_launch<32>(p, st)
                                                                   : tc::shift_launch<16>(p, st);
      }
      if (rc) return rc;
    }
    return 0;
  }
  return 1;
}

// 0 = ran, 1 = not supported (caller falls back), else error
int conv_fwd_shift_try(const float* X, const a2c_state_view* xview, const float* Wt, const float* bias, float* Y,
                       int N, const a2c_conv_shape* s, int relu, cudaStream_t st, uint8_t* relu_bits) {
  const int K = s->K, S = s->stride, P = s->pad;
  // first layers (<= 4 channels, stride 1): channels padded to one 8-deep k-step; the source may be
  // the uint8 plane ring, whose {0,1} values are exact in TF32 (no lo operand)
  const bool first = s->Cin <= 4 && S == 1;
  if (!shift_enabled() || (!X && !xview) || K != 3 || P != 1 || (S != 1 && S != 2)) return 1;
  if (!first && ((s->Cin % 8) != 0 || xview)) return 1;
  if (xview && (xview->dtype != 0 || xview->fs != s->Cin || (reinterpret_cast<uintptr_t>(xview->base) & 3) ||
                ((xview->stride_r | xview->stride_t | xview->stride_c) & 3)))
    return 1;
  if ((S == 2 && ((s->H | s->W) & 1)) || s->Cout > 128) return 1;
  static thread_local ShiftParams p;
  memset(&p, 0, sizeof(p));
  const int Cp = first ? 8 : s->Cin;
  p.src = xview ? nullptr : X; p.N = N; p.C = Cp; p.SH = s->H; p.SW = s->W; p.ncg = Cp / 4; p.ksteps = Cp / 8;
  p.cvalid = s->Cin; p.nsplit = xview ? 1 : 2;
  if (xview) { p.view = *xview; p.use_view = 1; }
  // source (y, x) lives in parity plane ((y+P) % S, (x+P) % S) at ((y+P) / S, (x+P) / S); tap
  // (ky, kx) of output (oy, ox) reads plane (ky % S, kx % S) at (oy + ky / S, ox + kx / S)
  p.cs = S; p.p0y = P; p.p0x = P;
  p.PW = s->OW + (K - 1) / S;
  for (int cy = 0; cy < S; ++cy) p.rext_y[cy] = (K - 1 - cy) / S;
  if (s->W & 3) {
    if (first || s->W > 32) return 1;
    p.scalar_w = 1;
  } else if (s->W > 128 || (!xview && (reinterpret_cast<uintptr_t>(X) & 15))) {
    return 1;
  }
  p.QH = s->OH; p.noc = 1; p.os = 1;
  p.oc[0] = {0, K * K, 0, 0, s->OH, s->OW};
  p.ntaps_total = K * K; p.nch = K;                 // one accumulator chain per kernel row
  p.wt = Wt; p.wsn = s->Cin * K * K; p.wsk = K * K; p.J = s->Cout; p.bias = bias; p.relu = relu;
  p.out = Y; p.OutH = s->OH; p.OutW = s->OW;
  if (relu_bits && (!relu || (s->Cout & 7))) return 1;
  p.obits = relu_bits;
  return shift_run(p, [K, S](ShiftParams& q) {
    for (int ky = 0; ky < K; ++ky)
      for (int kx = 0; kx < K; ++kx) {
        const int g = ky * K + kx, cls = (ky % S) * S + (kx % S);
        q.taps[g].a_off16 = (uint32_t)(q.cls_pos[cls] + (ky / S) * q.PW + (kx / S));
        q.taps[g].woff = (uint32_t)g;
      }
    return true;
  }, st);
}

bool conv_dgrad_shift_ok(const a2c_conv_shape* s) {
  const int S = s->stride;
  if (!shift_enabled() || s->K != 3 || s->pad != 1 || (S != 1 && S != 2) || (s->Cout % 8) != 0) return false;
  if ((S == 2 && ((s->H | s->W) & 1)) || s->Cin > 128) return false;
  return (s->OW & 3) == 0 ? s->OW <= 128 : s->OW <= 32;     // narrow odd planes: scalar producers
}

// Xmask / dbias_prev (both optional): dX = Xmask > 0 ? dX : 0 and dbias_prev f64 [Cin] += per-channel sums
// of the masked dX (the caller zeroes it) -- the ReLU + bias gradient of the layer that produced X.
int conv_dgrad_shift_try(const float* dY, const float* Ymask, const float* Wt, float* dX, const float* Xmask,
                         double* dbias_prev, int N, const a2c_conv_shape* s, cudaStream_t st,
                         const uint8_t* Xbits) {
  const int K = s->K, S = s->stride, P = s->pad;
  if (!conv_dgrad_shift_ok(s)) return 1;
  if (dbias_prev && !Xmask && !Xbits) return 1;
  if (Xbits && (s->Cin & 7)) Xbits = nullptr;          // (the fp32 mask is always passed as well)
  static thread_local ShiftParams p;
  memset(&p, 0, sizeof(p));
  p.src = dY; p.mask = Ymask; p.N = N; p.C = s->Cout; p.SH = s->OH; p.SW = s->OW;
  p.ncg = s->Cout / 4; p.ksteps = s->Cout / 8;
  p.cvalid = s->Cout; p.nsplit = 2;
  p.cs = 1;
  // tap k contributes to dX row y = S*j + a from dY row j + d, d = (a + P - k) / S when divisible
  int dmin = 0, dmax = 0;
  for (int a = 0; a < S; ++a)
    for (int k = 0; k < K; ++k)
      if (((a + P - k) % S + S) % S == 0) {
        const int d = tc::floordiv_i(a + P - k, S);
        dmin = d < dmin ? d : dmin;
        dmax = d > dmax ? d : dmax;
      }
  p.p0y = -dmin; p.p0x = -dmin;
  p.rext_y[0] = dmax - dmin;
  p.PW = cdiv(s->W, S) + p.rext_y[0];
  if (s->OW & 3) {
    p.scalar_w = 1;
  } else if ((reinterpret_cast<uintptr_t>(dY) & 15) || (Ymask && (reinterpret_cast<uintptr_t>(Ymask) & 15))) {
    return 1;
  }
  p.QH = cdiv(s->H, S); p.noc = S * S; p.os = S;
  p.nch = 1;
  p.wt = Wt; p.wsn = K * K; p.wsk = s->Cin * K * K; p.J = s->Cin;
  p.out = dX; p.OutH = s->H; p.OutW = s->W;
  p.omask = Xbits ? nullptr : Xmask; p.omaskbits = Xbits; p.obias = dbias_prev;
  p.ntaps_total = K * K;
  const int H = s->H, W = s->W;
  return shift_run(p, [K, S, P, H, W, dmin](ShiftParams& q) {
    int g = 0;
    for (int a = 0; a < S; ++a)
      for (int b = 0; b < S; ++b) {
        tc::ShiftClass& c = q.oc[a * S + b];
        c.tap0 = g; c.a = a; c.b = b;
        c.qh = (H - a + S - 1) / S; c.qw = (W - b + S - 1) / S;
        for (int ky = 0; ky < K; ++ky) {
          if (((a + P - ky) % S + S) % S != 0) continue;
          for (int kx = 0; kx < K; ++kx) {
            if (((b + P - kx) % S + S) % S != 0) continue;
            const int dy = tc::floordiv_i(a + P - ky, S) - dmin, dx = tc::floordiv_i(b + P - kx, S) - dmin;
            q.taps[g].a_off16 = (uint32_t)(dy * q.PW + dx);
            q.taps[g].woff = (uint32_t)(ky * K + kx);
            ++g;
          }
        }
        c.ntaps = g - c.tap0;
        if (c.ntaps == 0) return false;
      }
    return true;
  }, st);
}

}  // namespace a2c

// Shift-GEMM weight gradient of 3x3 pad-1 convolutions on tcgen05.
//
//   dW[co][ci][ky][kx] = sum over images and output positions of  X[ci][pos + shift(ky,kx)] * dY[co][pos]
//
// is a GEMM whose reduction (K) dimension is the output position.  Both operands are staged in
// shared memory "position-major": one 128-byte row per position holding 32 channel slots, hi / lo
// TF32 copies, 32-byte chunks XOR-swizzled by (row & 3) -- the SWIZZLE_128B_BASE32B layout, the
// only one tcgen05 defines for MN-major 32-bit operands (the no-swizzle form returns zeros).
// A k-step is 8 consecutive positions (= 8 rows, 1 KB); channel slots are the M / N extent.  A
// tap is a shift of the X start address by whole rows, so there is no im2col and no transpose.
//
//   Cin <= 16           : an X row holds [hi (16 slots) | lo (16 slots)] of its channels in ONE 32-slot atom, so
//                         the atoms of rows ky = 0, 1, 2 stack along M (M = 128, full tensor rate) and a single MMA
//                         per kx against [dY_hi | dY_lo] yields all hi / lo cross products: no second MMA, no
//                         separate lo copy of the band (half the X footprint -> taller bands).
//   stride 1, Cin <= 4  : (first layers) rows y, y+1, y+2 share one atom, 4 slots each: M = 64.
//                         The input may be the uint8 plane ring (state view): {0,1} values are
//                         exact in TF32, so the x_lo * dy_hi product is skipped.
//   stride 1, Cin  > 16 : atoms of rows y, y+1, y+2 stack along M (M = 128).
//   stride 2            : source rows are stored one after the other, each split into its two
//                         x-parity planes; the atoms of rows 2oy-1, 2oy, 2oy+1 stack along M.
// Image rows are processed in column chunks of <= 40 (stride 2: 20) positions so that two band
// buffers fit.
// dY is masked by the layer's ReLU (Ymask > 0) while it is converted.  The 3 kx accumulators
// (3 x 32 TMEM columns) are chained over a few rows only (TMEM adds truncate, tc_core.cuh); the
// drain warps keep running sums in fp32 registers and write one partial per CTA, reduced in
// double afterwards.  The bias gradient is a masked per-channel sum in a separate small kernel.
//
// Replaces the weight / bias gradient of nn.Conv2d in a2c/models.py:98,312,685 (autograd).
#include <stdlib.h>

#include <atomic>
#include <type_traits>

#include "tc_core.cuh"

namespace a2c {
extern std::atomic<long long> g_launches;

namespace tc {

constexpr int WG_PROD = 12;
constexpr int WG_NISS = 4;
constexpr int WG_DRAIN0 = WG_PROD, WG_MMA = WG_DRAIN0 + NUM_EPI_WARPS, WG_ALLOC = WG_MMA + WG_NISS;
constexpr int WG_THREADS = 32 * (WG_ALLOC + 1);
constexpr int WG_NSLOT = 2;
constexpr int WG_NACT = WG_NSLOT < 4 ? WG_NSLOT : 4;   // issuer warps that take chunks (<= WG_NISS)
constexpr int WG_N = 32;                     // output channels per launch
constexpr int WG_NACC = 3 * WG_N;            // accumulator values per row (kx = 0, 1, 2) in the per-CTA partials
// TMEM columns of one accumulator set: per kx [x * dy_hi (32) | x_hi * dy_lo (32)].  The 3xTF32 product is
// issued as two MMAs,  D[:, 0:64] (+)= X_hi * [dY_hi | dY_lo]  and  D[:, 0:32] += X_lo * dY_hi  (the dY lo
// copy sits one LBO = o_half bytes after the hi copy, i.e. it is the second 32-slot atom along N), which
// cuts the tensor core's shared-memory operand reads per product from 3 (A + B) to 2 A + 1.5 B.
constexpr int WG_TCOLS = 3 * 2 * WG_N;
constexpr int WG_ROWB = 128;                 // bytes of one position row (32 slots)
// descriptor high word: SBO = 512 B (4 rows), version 1, base offset 0, SWIZZLE_128B_BASE32B (= 1)
constexpr uint32_t WG_DESC_HI = (512u >> 4) | (1u << 14) | (1u << 29);
static_assert(WG_DRAIN0 % 4 == 0, "drain warps must start at a multiple of 4");

struct WgShiftParams {
  const float* x;        // f32 [N][Cin][H][W] (NULL when use_view)
  a2c_state_view view;   // first-layer source: uint8 plane ring
  int use_view;
  int pack3;             // Cin <= 4: slots [4ky, 4ky+4) of a row = channels of image row y + ky
  int nsplit;            // 2: X has a lo part (3 MMAs per product), 1: X exact in TF32 (2 MMAs)
  const float* dy;       // f32 [N][Cout][OH][OW]
  const float* mask;     // optional (layer output Y): dY counts only where mask > 0
  int N, Cin, H, W, Cout, OH, OW;   // Cin: input channels of THIS launch (a chunk of <= 32 of CinTot)
  int CinTot, ci0;       // channels per image of x, first channel of this launch
  int scalar_w;          // W or OW not a multiple of 4 (10x10 / 5x5 planes): one pixel per lane, scalar loads
  int opack;             // Cout <= 16: a dY row holds [hi (16 slots) | lo (16 slots)] in ONE 32-slot atom, no second copy
  int S;                 // stride
  int xpack;             // Cin <= 16 (not a first layer): X rows hold [hi 16 | lo 16] in one atom, no lo copy
  int ncx;               // x-parity planes per X row (1, or 2 for stride 2)
  int OWc, ncol;         // output columns per chunk, chunks per row
  int PWx, PWo;          // positions per X plane row / per dY row (PWo multiple of 8)
  int xrows;             // source rows of a full band (Rb + 2 for stride 1, 2 Rb + 1 for stride 2)
  int Rb, nbands, rpc;   // band height, bands per image, rows per accumulator chunk
  int ones_chunk;        // 16-byte slot chunk of every X row holding the constant 1 (bias gradient row), or -1
  int pair_o;            // dY lane mapping: 1 = channel-group pairs on even / odd lanes, 0 = one channel group per task
  int pgroups;           // producer groups: 2 = each half of the producer warps fills its own band buffer
  int n0;                // first output channel of this launch
  int ncgo_valid;        // channel groups of dY present in this launch (<= 8)
  int inv_ncgx, inv_ncgo; // ceil(65536 / n): task / n == (task * inv) >> 16 for the small task counts here
  int mrows;             // MMA M (64 or 128)
  int a_lbo;             // bytes between the M atoms of A
  int x_half, x_buf;     // bytes of one hi copy / one buffer of the X band
  int o_half, o_buf;     // same for the dY band
  float* partial;        // [grid][128][WG_NACC]
  int dbg;               // A2C_WG_DBG, stage isolation for profiling: 2 = no MMAs, 4 = no producer work
};

// byte offset of 16-byte chunk j of the position row at byte address row_addr (bands 1024-byte aligned)
__device__ __forceinline__ uint32_t wg_swz(uint32_t row_addr, int j) {
  return ((((uint32_t)j >> 1) ^ ((row_addr >> 7) & 3u)) << 5) | (((uint32_t)j & 1u) << 4);
}

// (v[r], v[r+1], v[r+2], v[r+3]) with indices mod 4: the pixel a lane handles in step j is (j + r) & 3
__device__ __forceinline__ float4 rot4(float4 v, int r) {
  const float4 a = (r & 1) ? make_float4(v.y, v.z, v.w, v.x) : v;
  return (r & 2) ? make_float4(a.z, a.w, a.x, a.y) : a;
}

__device__ __forceinline__ void wg_store_split(uint8_t* dst, int half_bytes, float a, float b, float c, float d) {
  const float4 h = make_float4(tf32_rna(a), tf32_rna(b), tf32_rna(c), tf32_rna(d));
  *reinterpret_cast<float4*>(dst) = h;
  *reinterpret_cast<float4*>(dst + half_bytes) =
      make_float4(tf32_rna(a - h.x), tf32_rna(b - h.y), tf32_rna(c - h.z), tf32_rna(d - h.w));
}

// Stores the hi / lo split of 4 channels (16-byte chunk `cg` of a dY position row).  Unpacked: the lo copy
// lives o_half bytes after the hi copy.  Packed (Cout <= 16): it is chunk cg + 4 of the SAME row.
__device__ __forceinline__ void wg_store_dy(uint8_t* buf, uint32_t pos, uint32_t cg, int o_half, int opack,
                                            float a, float b, float c, float d) {
  const uint32_t c16 = (cg & 1u) << 4, c32 = cg >> 1;
  uint8_t* row = buf + pos * 128u;
  uint8_t* hi = row + (((c32 ^ (pos & 3u)) << 5) | c16);
  uint8_t* lo = opack ? row + ((((c32 + 2u) ^ (pos & 3u)) << 5) | c16) : hi + o_half;
  const float4 h = make_float4(tf32_rna(a), tf32_rna(b), tf32_rna(c), tf32_rna(d));
  *reinterpret_cast<float4*>(hi) = h;
  *reinterpret_cast<float4*>(lo) = make_float4(tf32_rna(a - h.x), tf32_rna(b - h.y), tf32_rna(c - h.z), tf32_rna(d - h.w));
}

__device__ __forceinline__ void wg_tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

__device__ __forceinline__ void wg_tmem_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr));
}

// Per-role clock stamps of CTA 0 (build flag -DA2C_TC_TRACE; read back with a2c_debug_wg_trace,
// scratch/wg_trace.py): the tool that found the GEMM's stalls (DESIGN.md lesson 19).
#ifdef A2C_TC_TRACE
static __device__ long long g_wg_trace[3][64][4];
#define WGT(role, it, ev)                                                                    \
  do {                                                                                       \
    if (blockIdx.x == 0 && lane == 0 && (it) >= 20 && (it) < 84) g_wg_trace[role][(it)-20][ev] = clock64(); \
  } while (0)
#else
#define WGT(role, it, ev) do { } while (0)
#endif

struct WgItem {
  int img, r0, rows, c0;          // image, first output row, rows, first output column of the chunk
};

// kind::tf32, fp32 accumulate, both operands MN-major
__host__ __device__ constexpr uint32_t make_idesc_mn(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}

__global__ void __launch_bounds__(WG_THREADS, 1) conv_shift_wgrad_kernel(const __grid_constant__ WgShiftParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int NC = WG_NACC / 2;                      // accumulator columns per drain warp
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* xb = base;                                  // 2 X buffers: [hi | lo][row][cx][c][32 slots]
  uint8_t* ob = xb + 2 * p.x_buf;                      // 2 dY buffers: [hi | lo][row][c][32 slots]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ob + 2 * p.o_buf + 8192);   // 8 KB: over-read of junk M atoms
  uint64_t* band_full = bars;                          // [2]
  uint64_t* band_empty = bars + 2;                     // [2]
  uint64_t* acc_full = bars + 4;                       // [WG_NSLOT]
  uint64_t* acc_empty = bars + 4 + WG_NSLOT;           // [WG_NSLOT]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 + 2 * WG_NSLOT);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int first_item = (int)blockIdx.x, item_step = (int)gridDim.x;
  const int per_img = p.nbands * p.ncol;
  const int nitems = p.N * per_img;

  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(&band_full[b], WG_PROD / p.pgroups);
      mbar_init(&band_empty[b], WG_NISS);
    }
    for (int b = 0; b < WG_NSLOT; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], NUM_EPI_WARPS);
    }
    mbar_fence_init();
  }
  if (warp == WG_ALLOC) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  {
    // pad positions / pad slots are never written: clear both bands (and the over-read slack) once
    const int nz = (2 * p.x_buf + 2 * p.o_buf + 8192) / 16;
    for (int i = threadIdx.x; i < nz; i += WG_THREADS) reinterpret_cast<uint4*>(xb)[i] = make_uint4(0, 0, 0, 0);
    if (p.ones_chunk >= 0) {
      // a free slot chunk of every X row holds 1.0 (hi copy; lo stays 0): that accumulator row becomes
      // sum(dY) = the bias gradient, for free
      __syncthreads();
      const int npos = p.xrows * p.ncx * p.PWx;
      for (int i = threadIdx.x; i < 2 * npos; i += WG_THREADS) {
        const int b = i / npos, pos = i - b * npos;
        uint8_t* d = xb + (size_t)b * p.x_buf + (size_t)pos * WG_ROWB +
                     (((((uint32_t)p.ones_chunk >> 1) ^ ((uint32_t)pos & 3u)) << 5) | (((uint32_t)p.ones_chunk & 1u) << 4));
        *reinterpret_cast<float4*>(d) = make_float4(1.f, 1.f, 1.f, 1.f);
      }
    }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  auto item_of = [&](int item) {
    WgItem it;
    it.img = item / per_img;
    int r = item - it.img * per_img;
    const int bi = r / p.ncol;
    it.c0 = (r - bi * p.ncol) * p.OWc;
    it.r0 = bi * p.Rb;
    it.rows = min(p.Rb, p.OH - it.r0);
    return it;
  };

  if (warp < WG_PROD) {
    // ---------------- producers ----------------
    const int xplane = p.H * p.W, oplane = p.OH * p.OW;
    const int ncgx = p.Cin >> 2;
    // Lane mapping of the X / dY units.  A task is (row, PAIR of channel groups); inside a task
    // slot even / odd lanes take the two channel groups, lane >> 1 the float4 group gi along the
    // row.  A lane holds a 4 pixel x 4 channel block and stores pixel (j + gi) & 3 in step j: the 8
    // lanes of a quarter-warp (4 consecutive gi x 2 channel groups) then hit 4 swizzle phases x 2
    // chunk parities = 8 different 16-byte bank groups.  (With all lanes on the same pixel index
    // and channel group every 128-bit store was an 8-way bank conflict: ncu showed 82 % of the
    // shared wavefronts as conflicts, 3 ms of shared-memory time per launch of conv2's layer.)
    const int ngx = p.S == 1 ? (p.OWc >> 2) + 2 : (p.OWc >> 1) + 1;   // float4 groups covering a source row chunk
    const int tpwx = 32 / (2 * ngx) > 0 ? 32 / (2 * ngx) : 1;
    const int subx = lane / (2 * ngx), lx = lane - subx * 2 * ngx, cgpx = lx & 1, gix = lx >> 1;
    const int XN = p.S == 1 ? p.OWc + 2 : 2 * p.OWc + 1;              // source columns needed
    const int npairx = (ncgx + 1) >> 1;
    // (narrow dY, <= 4 channel groups: one group per task keeps more warps busy per band than pairs do)
    const int ngo = p.OWc >> 2, lpo = p.pair_o ? 2 * ngo : ngo;        // lanes per dY task
    const int tpwo = 32 / lpo > 0 ? 32 / lpo : 1;
    const int subo = lane / lpo, lo = lane - subo * lpo;
    const int cgpo = p.pair_o ? lo & 1 : 0, gio = p.pair_o ? lo >> 1 : lo;
    // Layers whose band is a single pass of work for half of the producer warps (few channels)
    // run two producer groups, group g only ever filling buffer g: the per-band latency chain
    // (wait, load, convert, store) of consecutive bands then overlaps.
    const int pgw = WG_PROD / p.pgroups;                    // warps per group
    const int grp = warp / pgw, pw = warp - grp * pgw;
    struct { int s; uint32_t par; } ring = {p.pgroups == 2 ? grp : 0, 0u};
    int tit = warp == 0 ? 0 : -1000000;
    for (int item = first_item + grp * item_step; item < nitems; item += p.pgroups * item_step, ++tit) {
      const WgItem it = item_of(item);
      WGT(0, tit, 0);
      mbar_wait_relaxed(&band_empty[ring.s], ring.par ^ 1);
      WGT(0, tit, 1);
      if (!(p.dbg & 4)) {
      if (p.pack3) {
        // --- first layer: <= 4 channels, f32 tensor or uint8 ring view; a row's pixel goes to
        //     chunk ky of row r - ky (ky = 0, 1, 2) ---
        uint8_t* buf = xb + (size_t)ring.s * p.x_buf;
        const uint32_t buf_a = smem_u32(buf);
        const int xs = it.c0 - 1, xe = it.c0 + p.OWc + 1;
        const int g0 = (xs < 0 ? xs - 3 : xs) / 4;
        const int ng = (xe + 3) / 4 - g0;
        const int tpw = 32 / ng > 0 ? 32 / ng : 1;
        const int sub = lane / ng, gi = lane - sub * ng;
        long long ibase = (long long)it.img * p.Cin * xplane, cstride = xplane;
        uint32_t cmask = 0xfu;
        if (p.use_view) {
          const long long rr = it.img / p.view.T;
          const int tt = it.img - (int)rr * p.view.T;
          ibase = rr * p.view.stride_r + (long long)tt * p.view.stride_t;
          cstride = p.view.stride_c;
          if (p.view.pflag) {                       // planes older than a reset read as zero (a2c/utils.py:37-40)
            const uint8_t* pf = p.view.pflag + rr * p.view.pflag_stride_r + tt;
            cmask = 0;
            bool live = true;
            for (int c = p.view.fs - 1; c >= 0; --c) {
              if (live) cmask |= 1u << c;
              live = live && (pf[c] == 0);
            }
          }
        }
        const int nrow = it.rows + 2;
        for (int r = pw * tpw + sub; r < nrow; r += pgw * tpw) {
          if (sub >= tpw) continue;
          const int sy = it.r0 + r - 1;
          const int x4 = 4 * (g0 + gi);
          const bool ok = (unsigned)sy < (unsigned)p.H && (unsigned)x4 < (unsigned)p.W;
          // Lane gi holds a 4 pixel x 4 channel block.  In step (jj, kk) it stores pixel
          // (jj + gi) & 3 into chunk k = (kk + (gi >> 2)) % 3 of row r - k: the 8 lanes of a
          // quarter-warp (4 consecutive gi at one k, the next 4 at k + 1) then hit 4 swizzle
          // phases x 2 chunk parities = 8 different 16-byte bank groups (2-way only when the two
          // k's are 2 and 0).  With every lane on the same (pixel, k) all 8 stores of a
          // quarter-warp landed in one bank group (ncu r1: 70 % of this kernel's shared
          // wavefronts were conflicts).
          float4 vc[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok && c < p.Cin && ((cmask >> c) & 1u)) {
              const long long a = ibase + c * cstride + (long long)sy * p.W + x4;
              if (p.use_view) {
                const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(static_cast<const uint8_t*>(p.view.base) + a));
                v = make_float4((float)(w & 0xff), (float)((w >> 8) & 0xff), (float)((w >> 16) & 0xff), (float)(w >> 24));
              } else {
                v = __ldg(reinterpret_cast<const float4*>(p.x + a));
              }
            }
            vc[c] = rot4(v, gi);
          }
          const float px[4][4] = {{vc[0].x, vc[1].x, vc[2].x, vc[3].x}, {vc[0].y, vc[1].y, vc[2].y, vc[3].y},
                                  {vc[0].z, vc[1].z, vc[2].z, vc[3].z}, {vc[0].w, vc[1].w, vc[2].w, vc[3].w}};
          const int krot = gi >> 2;
#pragma unroll
          for (int kk = 0; kk < 3; ++kk) {
            int k = kk + krot;
            k -= 3 * (k / 3);
            if (r - k < 0) continue;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const int j = (jj + gi) & 3;
              const int xp = x4 + j - xs;
              if (xp < 0 || x4 + j >= xe) continue;
              const uint32_t off = (uint32_t)((r - k) * p.PWx + xp) * WG_ROWB;
              uint8_t* d = buf + off + wg_swz(buf_a + off, k);
              if (p.nsplit == 2) {
                wg_store_split(d, p.x_half, px[jj][0], px[jj][1], px[jj][2], px[jj][3]);
              } else {
                // {0,1} ring values are exact in TF32: the lo copy stays zero and is never read
                *reinterpret_cast<float4*>(d) = make_float4(px[jj][0], px[jj][1], px[jj][2], px[jj][3]);
              }
            }
          }
        }
      } else {
      // --- X band.  A lane converts 4 consecutive pixels (one float4 per channel) of 4 channels. ---
        uint8_t* buf = xb + (size_t)ring.s * p.x_buf;
        const float* ximg = p.x + ((long long)it.img * p.CinTot + p.ci0) * xplane;
        const int nrow_s = p.S == 1 ? it.rows + 2 : 2 * it.rows + 1;
        if (p.scalar_w) {
          // narrow planes (rows not 16-byte multiples): lane <-> one needed source column, 4 channels,
          // scalar loads; a task is (source row, channel group)
          const int tpws = 32 / XN > 0 ? 32 / XN : 1;
          const int subs = lane / XN, xp = lane - subs * XN;
          const int col = p.S * it.c0 - 1 + xp;
          const int ntask = nrow_s * ncgx;
          for (int task = pw * tpws + subs; task < ntask; task += pgw * tpws) {
            if (subs >= tpws) continue;
            const int r = task / ncgx, cg = task - r * ncgx;
            const int sy = p.S * it.r0 + r - 1;
            const bool ok = (unsigned)col < (unsigned)p.W && (unsigned)sy < (unsigned)p.H;
            float c4[4];
#pragma unroll
            for (int c = 0; c < 4; ++c)
              c4[c] = ok ? __ldg(ximg + (long long)(cg * 4 + c) * xplane + sy * p.W + col) : 0.0f;
            const uint32_t pos = (uint32_t)(p.ncx * r * p.PWx + (p.S == 1 ? xp : (xp & 1) * p.PWx + (xp >> 1)));
            const uint32_t c16 = ((uint32_t)cg & 1u) << 4, c32 = (uint32_t)cg >> 1;
            const float4 h = make_float4(tf32_rna(c4[0]), tf32_rna(c4[1]), tf32_rna(c4[2]), tf32_rna(c4[3]));
            const float4 l = make_float4(tf32_rna(c4[0] - h.x), tf32_rna(c4[1] - h.y), tf32_rna(c4[2] - h.z), tf32_rna(c4[3] - h.w));
            uint8_t* d = buf + pos * WG_ROWB + (((c32 ^ (pos & 3u)) << 5) | c16);
            *reinterpret_cast<float4*>(d) = h;
            *reinterpret_cast<float4*>(p.xpack ? buf + pos * WG_ROWB + ((((c32 + 2u) ^ (pos & 3u)) << 5) | c16)
                                               : d + p.x_half) = l;
          }
        } else {
        const float4* simg = reinterpret_cast<const float4*>(ximg);
        const int x4 = 4 * (((p.S * it.c0) >> 2) - 1 + gix);           // first source column of this lane's group
        const bool okx = (unsigned)x4 < (unsigned)p.W;
        const int nrow = nrow_s;                                       // source rows of the band
        const int ntask = nrow * npairx;
        for (int task = pw * tpwx + subx; task < ntask; task += pgw * tpwx) {
          const int r = (task * p.inv_ncgx) >> 16, cg = 2 * (task - r * npairx) + cgpx;
          if (subx >= tpwx || cg >= ncgx) continue;
          const int sy = p.S * it.r0 + r - 1;
          const bool ok = okx && (unsigned)sy < (unsigned)p.H;
          const int e0 = ok ? ((cg * 4 * xplane + sy * p.W + x4) >> 2) : 0;
          const int pl4 = ok ? xplane >> 2 : 0;
          float4 v0 = __ldg(simg + e0), v1 = __ldg(simg + e0 + pl4), v2 = __ldg(simg + e0 + 2 * pl4),
                 v3 = __ldg(simg + e0 + 3 * pl4);
          if (!ok) v0 = v1 = v2 = v3 = make_float4(0.f, 0.f, 0.f, 0.f);
          v0 = rot4(v0, gix); v1 = rot4(v1, gix); v2 = rot4(v2, gix); v3 = rot4(v3, gix);
          const float px[4][4] = {{v0.x, v1.x, v2.x, v3.x}, {v0.y, v1.y, v2.y, v3.y},
                                  {v0.z, v1.z, v2.z, v3.z}, {v0.w, v1.w, v2.w, v3.w}};
          const int rowpos = p.ncx * r * p.PWx;                        // stride 2: two parity planes per source row
          const uint32_t c16 = ((uint32_t)cg & 1u) << 4, c32 = (uint32_t)cg >> 1;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int xp = 4 * gix + ((j + gix) & 3) - 3;              // source column (0-based in the needed range) of this step's pixel
            if (xp < 0 || xp >= XN) continue;
            const uint32_t pos = (uint32_t)(rowpos + (p.S == 1 ? xp : (xp & 1) * p.PWx + (xp >> 1)));
            const float4 h = make_float4(tf32_rna(px[j][0]), tf32_rna(px[j][1]), tf32_rna(px[j][2]), tf32_rna(px[j][3]));
            const float4 l = make_float4(tf32_rna(px[j][0] - h.x), tf32_rna(px[j][1] - h.y), tf32_rna(px[j][2] - h.z), tf32_rna(px[j][3] - h.w));
            // buffers are 1024-byte aligned: the swizzle phase of a row is its position & 3
            uint8_t* d = buf + pos * WG_ROWB + (((c32 ^ (pos & 3u)) << 5) | c16);
            *reinterpret_cast<float4*>(d) = h;
            // (packed: the lo part is chunk cg + 4 of the same row, i.e. slots [16, 32))
            *reinterpret_cast<float4*>(p.xpack ? buf + pos * WG_ROWB + ((((c32 + 2u) ^ (pos & 3u)) << 5) | c16)
                                               : d + p.x_half) = l;
          }
        }
        }
      }
      WGT(0, tit, 2);
      // --- dY band (masked) ---
      if (p.scalar_w) {
        uint8_t* buf = ob + (size_t)ring.s * p.o_buf;
        const float* dimg = p.dy + ((long long)it.img * p.Cout + p.n0) * oplane;
        const float* mimg = p.mask ? p.mask + ((long long)it.img * p.Cout + p.n0) * oplane : nullptr;
        const int tpws = 32 / p.OWc > 0 ? 32 / p.OWc : 1;
        const int subs = lane / p.OWc, xo = lane - subs * p.OWc;
        const int ntask = it.rows * p.ncgo_valid;
        for (int task = pw * tpws + subs; task < ntask; task += pgw * tpws) {
          if (subs >= tpws) continue;
          const int jj = task / p.ncgo_valid, cg = task - jj * p.ncgo_valid;
          float c4[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const long long a = (long long)(cg * 4 + c) * oplane + (it.r0 + jj) * p.OW + it.c0 + xo;
            float v = __ldg(dimg + a);
            if (mimg) v = __ldg(mimg + a) > 0.f ? v : 0.f;
            c4[c] = v;
          }
          wg_store_dy(buf, (uint32_t)(jj * p.PWo + xo), (uint32_t)cg, p.o_half, p.opack, c4[0], c4[1], c4[2], c4[3]);
        }
      } else {
        uint8_t* buf = ob + (size_t)ring.s * p.o_buf;
        const float4* simg = reinterpret_cast<const float4*>(p.dy + ((long long)it.img * p.Cout + p.n0) * oplane);
        const float4* mimg =
            p.mask ? reinterpret_cast<const float4*>(p.mask + ((long long)it.img * p.Cout + p.n0) * oplane) : nullptr;
        const int npairo = p.pair_o ? (p.ncgo_valid + 1) >> 1 : p.ncgo_valid;   // tasks per row
        const int ntask = it.rows * npairo;
        const int pl4 = oplane >> 2;
        for (int task = pw * tpwo + subo; task < ntask; task += pgw * tpwo) {
          const int jj = (task * p.inv_ncgo) >> 16;
          const int cg = p.pair_o ? 2 * (task - jj * npairo) + cgpo : task - jj * npairo;
          if (subo >= tpwo || cg >= p.ncgo_valid) continue;
          const int e0 = ((cg * 4 * oplane + (it.r0 + jj) * p.OW + it.c0) >> 2) + gio;
          float4 v0 = __ldg(simg + e0), v1 = __ldg(simg + e0 + pl4), v2 = __ldg(simg + e0 + 2 * pl4),
                 v3 = __ldg(simg + e0 + 3 * pl4);
          if (mimg) {
            const float4 m0 = __ldg(mimg + e0), m1 = __ldg(mimg + e0 + pl4), m2 = __ldg(mimg + e0 + 2 * pl4),
                         m3 = __ldg(mimg + e0 + 3 * pl4);
            v0.x = m0.x > 0.f ? v0.x : 0.f; v0.y = m0.y > 0.f ? v0.y : 0.f; v0.z = m0.z > 0.f ? v0.z : 0.f; v0.w = m0.w > 0.f ? v0.w : 0.f;
            v1.x = m1.x > 0.f ? v1.x : 0.f; v1.y = m1.y > 0.f ? v1.y : 0.f; v1.z = m1.z > 0.f ? v1.z : 0.f; v1.w = m1.w > 0.f ? v1.w : 0.f;
            v2.x = m2.x > 0.f ? v2.x : 0.f; v2.y = m2.y > 0.f ? v2.y : 0.f; v2.z = m2.z > 0.f ? v2.z : 0.f; v2.w = m2.w > 0.f ? v2.w : 0.f;
            v3.x = m3.x > 0.f ? v3.x : 0.f; v3.y = m3.y > 0.f ? v3.y : 0.f; v3.z = m3.z > 0.f ? v3.z : 0.f; v3.w = m3.w > 0.f ? v3.w : 0.f;
          }
          v0 = rot4(v0, gio); v1 = rot4(v1, gio); v2 = rot4(v2, gio); v3 = rot4(v3, gio);
          const float px[4][4] = {{v0.x, v1.x, v2.x, v3.x}, {v0.y, v1.y, v2.y, v3.y},
                                  {v0.z, v1.z, v2.z, v3.z}, {v0.w, v1.w, v2.w, v3.w}};
          const uint32_t pos0 = (uint32_t)(jj * p.PWo + 4 * gio);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            wg_store_dy(buf, pos0 + (uint32_t)((j + gio) & 3), (uint32_t)cg, p.o_half, p.opack, px[j][0], px[j][1],
                        px[j][2], px[j][3]);
        }
      }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&band_full[ring.s]);
      WGT(0, tit, 3);
      if (p.pgroups == 2) ring.par ^= 1u;                    // same buffer again, next phase
      else if ((ring.s ^= 1) == 0) ring.par ^= 1u;           // alternate buffers
    }
  } else if (warp >= WG_MMA && warp < WG_MMA + WG_NISS) {
    // ---------------- MMA issuers (chunks dealt round-robin) ----------------
    const uint32_t me = (uint32_t)(warp - WG_MMA);
    // descriptor low word: start address >> 4 | LBO >> 4 << 16; high word: SBO = 512 B (4 rows),
    // version 1, base offset 0, layout SWIZZLE_128B_BASE32B (= 1)
    const uint32_t a_w0 = ((smem_u32(xb) & 0x3FFFFu) >> 4) | (((uint32_t)p.a_lbo >> 4) << 16);
    const uint32_t b_w0 = ((smem_u32(ob) & 0x3FFFFu) >> 4) | (((uint32_t)p.o_half >> 4) << 16);   // LBO: hi -> lo atom
    const uint32_t xbuf16 = (uint32_t)p.x_buf >> 4, xhalf16 = (uint32_t)p.x_half >> 4;
    const uint32_t obuf16 = (uint32_t)p.o_buf >> 4;
    const uint32_t xrow16 = (uint32_t)(p.S * p.ncx * p.PWx) * (WG_ROWB >> 4);   // 16-byte units per output row of X
    const uint32_t orow16 = (uint32_t)p.PWo * (WG_ROWB >> 4);
    const int ksteps = p.PWo >> 3;
    StageRing ring(2);
    uint32_t ccount = 0, slot = 0, sphase = 0;
    int tit = me == 0 ? 0 : -1000000;
    for (int item = first_item; item < nitems; item += item_step, ring.next(), ++tit) {
      const WgItem it = item_of(item);
      WGT(1, tit, 0);
      mbar_wait(&band_full[ring.s], ring.par);
      WGT(1, tit, 1);
      tc_fence_after();
      const uint32_t a_buf = a_w0 + ring.s * xbuf16, b_buf = b_w0 + ring.s * obuf16;
      for (int j0 = 0; j0 < it.rows; j0 += p.rpc, ++ccount) {
        const uint32_t ab = slot, aphase = sphase;
        if (++slot == WG_NSLOT) { slot = 0; sphase ^= 1u; }
        // a slot is always reused by the same issuer warp (parity waits cannot see two phases back):
        // chunks go to warp c % min(WG_NISS, WG_NSLOT)
        if ((ccount & (WG_NACT - 1)) != me) continue;
        mbar_wait(&acc_empty[ab], aphase ^ 1u);
        tc_fence_after();
        if (!(p.dbg & 2) && elect_one()) {
          const uint32_t tmem_d = tmem_base + ab * WG_TCOLS;
          const int j1 = min(it.rows, j0 + p.rpc);
          // idesc / descriptor high word as immediates (uniform-register issue loop, see mma_tf32_imm)
          auto issue = [&](auto mtag, auto ntag) {
            constexpr int NH = decltype(ntag)::value;                                   // dY_hi columns: 32, or 16 when packed
            constexpr uint32_t ID1 = make_idesc_mn(decltype(mtag)::value, NH);          // X_lo * dY_hi
            constexpr uint32_t ID2 = make_idesc_mn(decltype(mtag)::value, 2 * NH);      // X_hi * [dY_hi | dY_lo]
            uint32_t acc = 0;
            for (int jj = j0; jj < j1; ++jj) {
              uint32_t a_row = a_buf + (uint32_t)jj * xrow16;
              uint32_t b_row = b_buf + (uint32_t)jj * orow16;
              for (int kk = 0; kk < ksteps; ++kk) {
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                  // stride 1: row x0 + kx; stride 2: parity plane kx & 1, row x0 + (kx >> 1)
                  const uint32_t a_hi = a_row + (p.ncx == 1 ? (uint32_t)kx : (uint32_t)((kx & 1) * p.PWx + (kx >> 1))) * 8u;
                  // (measured: the swizzle phase comes from the absolute address bits, so a start that is
                  //  not aligned to the 512-byte pattern needs NO base offset in the descriptor)
                  mma_tf32_imm<ID2, WG_DESC_HI>(tmem_d + kx * 2 * NH, a_hi, b_row, acc);
                  // ({0,1} ring sources are exact in TF32: no X_lo term)
                  // (packed X rows carry their lo part as extra M rows of the same atom: nothing more to issue)
                  if (p.nsplit == 2 && !p.xpack) mma_tf32_imm<ID1, WG_DESC_HI>(tmem_d + kx * 2 * NH, a_hi + xhalf16, b_row, 1u);
                }
                acc = 1u;
                a_row += 8 * 8;
                b_row += 8 * 8;
              }
            }
          };
          if (p.opack) {
            if (p.mrows == 64) issue(std::integral_constant<int, 64>{}, std::integral_constant<int, 16>{});
            else issue(std::integral_constant<int, 128>{}, std::integral_constant<int, 16>{});
          } else {
            if (p.mrows == 64) issue(std::integral_constant<int, 64>{}, std::integral_constant<int, 32>{});
            else issue(std::integral_constant<int, 128>{}, std::integral_constant<int, 32>{});
          }
        }
        __syncwarp();
        if (elect_one()) mma_commit(&acc_full[ab]);
      }
      if (elect_one()) mma_commit(&band_empty[ring.s]);
      WGT(1, tit, 2);
    }
  } else if (warp >= WG_DRAIN0 && warp < WG_DRAIN0 + NUM_EPI_WARPS) {
    // ---------------- drain: running sums in registers ----------------
    const int e = warp - WG_DRAIN0, q = e & 3, hsel = e >> 2;
    float acc[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) acc[j] = 0.0f;
    uint32_t slot = 0, sphase = 0;
    int tit = e == 0 ? 0 : -1000000;
    // The drain warps are idle ~70 % of the time (clock trace) while the producers -- the critical role
    // -- spend ~1,500 of the ~2,000 clk of every task waiting for HBM.  The drain therefore asks L2 for
    // the source rows of the band three items ahead (no registers, no scoreboard); the producers'
    // loads then hit L2.  (Prefetching from the producer warps themselves was slower: the prefetches
    // queued in front of their own demand loads.)
    auto prefetch_item = [&](int nitem) {
      if (nitem >= nitems || p.scalar_w) return;
      const WgItem n = item_of(nitem);
      const int tid = e * 32 + lane;
      const int xplane = p.H * p.W, oplane = p.OH * p.OW;
      if (!p.pack3) {                                        // (first layer: the uint8 ring is 1/16 of the bytes)
        const float* ximg = p.x + ((long long)n.img * p.CinTot + p.ci0) * xplane;
        const int nrow_s = p.S == 1 ? n.rows + 2 : 2 * n.rows + 1;
        const int x0 = max(0, p.S * n.c0 - 4), x1 = min(p.W, p.S * (n.c0 + p.OWc) + 4);
        const int nl = (x1 - x0 + 31) / 32 + 1;              // 128-byte lines per (row, channel), rows are not line-aligned
        const int total = nrow_s * p.Cin * nl;
        for (int i = tid; i < total; i += 32 * NUM_EPI_WARPS) {
          const int l = i % nl, rc = i / nl, c = rc % p.Cin, r = rc / p.Cin;
          const int sy = p.S * n.r0 + r - 1;
          if ((unsigned)sy < (unsigned)p.H)
            prefetch_l2(ximg + (long long)c * xplane + sy * p.W + min(x0 + l * 32, x1 - 1));
        }
      }
      {
        const long long obase = ((long long)n.img * p.Cout + p.n0) * oplane;
        const int nch = p.ncgo_valid * 4;
        const int nl = (p.OWc + 31) / 32 + 1;
        const int total = n.rows * nch * nl;
        for (int i = tid; i < total; i += 32 * NUM_EPI_WARPS) {
          const int l = i % nl, rc = i / nl, c = rc % nch, r = rc / nch;
          const long long a = obase + (long long)c * oplane + (n.r0 + r) * p.OW + min(n.c0 + l * 32, n.c0 + p.OWc - 1);
          prefetch_l2(p.dy + a);
          if (p.mask) prefetch_l2(p.mask + a);
        }
      }
    };
    prefetch_item(first_item + item_step);
    prefetch_item(first_item + 2 * item_step);
    for (int item = first_item; item < nitems; item += item_step, ++tit) {
      const WgItem it = item_of(item);
      prefetch_item(item + 3 * item_step);
      WGT(2, tit, 0);
      for (int j0 = 0; j0 < it.rows; j0 += p.rpc) {
        const uint32_t ab = slot;
        mbar_wait(&acc_full[ab], sphase);
        if (j0 == 0) { WGT(2, tit, 1); }
        if (++slot == WG_NSLOT) { slot = 0; sphase ^= 1u; }
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + ab * WG_TCOLS;
#pragma unroll
        for (int c0 = 0; c0 < NC; c0 += 8) {
          // output (kx, co) = hsel * NC + c0 + j  lives in TMEM columns kx * 64 + co (and + 32 for the dy_lo
          // half); 8 columns at a time keeps the 48 running sums + both halves inside the register budget
          // (packed dY, Cout <= 16: per kx [hi 16 | lo 16]; outputs co >= 16 are never read by the reduction)
          const int o = hsel * NC + c0;
          const uint32_t nh = p.opack ? 16u : (uint32_t)WG_N;
          const uint32_t col = (uint32_t)(o / WG_N) * 2u * nh + (uint32_t)(o % WG_N);
          uint32_t r[8], r2[8];
          wg_tmem_ld8(taddr + col, r);
          wg_tmem_ld8(taddr + col + nh, r2);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[c0 + j] += __uint_as_float(r[j]) + __uint_as_float(r2[j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[ab]);
      }
      WGT(2, tit, 2);
    }
    float* dst = p.partial + ((size_t)blockIdx.x * 128 + q * 32 + lane) * WG_NACC + hsel * NC;
#pragma unroll
    for (int j = 0; j < NC; ++j) dst[j] = acc[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == WG_ALLOC) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

// dW from the per-CTA partials (double accumulation).  Accumulator row of (ky, ci):
//   first layer (Cin <= 4) : ky * 4 + ci     otherwise : ky * 32 + ci  (packed X, Cin <= 16: plus row + 16 for the lo part)
// M = 64 accumulators live in TMEM lanes (m % 16) + 32 * (m / 16).
__global__ void __launch_bounds__(256)
wg_shift_reduce_kernel(const float* __restrict__ partial, int grid, int pack, int mrows, int Cin, int CinTot,
                       int ci0, int Cout, int n0, float* __restrict__ dW, int bias_row, float* __restrict__ dbias) {
  const int nco = min(WG_N, Cout - n0);
  const int nw = nco * Cin * 9;
  const int total = nw + (dbias && bias_row >= 0 ? nco : 0);
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const bool is_bias = e >= nw;
    const int kx = is_bias ? 0 : e % 3;
    int r = e / 3;
    const int ky = r % 3;
    r /= 3;
    const int ci = r % Cin, co = is_bias ? e - nw : r / Cin;
    const int m = is_bias ? bias_row : (pack == 3 ? ky * 4 + ci : ky * 32 + ci);
    const int lane_row = mrows == 128 ? m : (m & 15) + 32 * (m >> 4);
    double s = 0.0;
    for (int b = 0; b < grid; ++b)
      s += (double)partial[((size_t)b * 128 + lane_row) * WG_NACC + kx * WG_N + co];
    if (pack == 4 && !is_bias)                 // packed X: the lo channels' products sit 16 accumulator rows further
      for (int b = 0; b < grid; ++b)
        s += (double)partial[((size_t)b * 128 + lane_row + 16) * WG_NACC + kx * WG_N + co];
    if (is_bias) dbias[n0 + co] = (float)s;
    else dW[(((size_t)(n0 + co) * CinTot + ci0 + ci) * 3 + ky) * 3 + kx] = (float)s;
  }
}

// dbias[c] = sum over images and positions of dY (where Ymask > 0)
__global__ void __launch_bounds__(256)
wg_bias_sum_kernel(const float* __restrict__ dY, const float* __restrict__ mask, double* __restrict__ acc, int N,
                   int C, int HW4) {
  __shared__ double scratch[32];
  const int c = blockIdx.y;
  double s = 0.0;
  const long long per = (long long)N * HW4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / HW4);
    const long long a = ((long long)n * C + c) * HW4 + (i - (long long)n * HW4);
    const float4 v = __ldg(reinterpret_cast<const float4*>(dY) + a);
    float t;
    if (mask) {
      const float4 m = __ldg(reinterpret_cast<const float4*>(mask) + a);
      t = ((m.x > 0.f ? v.x : 0.f) + (m.y > 0.f ? v.y : 0.f)) + ((m.z > 0.f ? v.z : 0.f) + (m.w > 0.f ? v.w : 0.f));
    } else {
      t = (v.x + v.y) + (v.z + v.w);
    }
    s += (double)t;
  }
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) atomicAdd(&acc[c], s);
}
// same for planes whose size is not a multiple of 4 (5x5)
__global__ void __launch_bounds__(256)
wg_bias_sum_scalar_kernel(const float* __restrict__ dY, const float* __restrict__ mask, double* __restrict__ acc,
                          int N, int C, int HW) {
  __shared__ double scratch[32];
  const int c = blockIdx.y;
  double s = 0.0;
  const long long per = (long long)N * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / HW);
    const long long a = ((long long)n * C + c) * HW + (i - (long long)n * HW);
    const float v = __ldg(dY + a);
    s += (double)((!mask || __ldg(mask + a) > 0.f) ? v : 0.f);
  }
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) atomicAdd(&acc[c], s);
}
__global__ void wg_bias_finish_kernel(const double* __restrict__ acc, float* __restrict__ out, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) out[c] = (float)acc[c];
}

static constexpr int WG_SMEM_LIMIT = 227 * 1024;

static size_t wg_smem_bytes(const WgShiftParams& p) {
  return 2 * (size_t)p.x_buf + 2 * (size_t)p.o_buf + 8192 + 1024 + 256;
}

static bool wg_plan(WgShiftParams* p) {
  for (int rb = p->OH; rb >= 1; --rb) {
    const int nb = cdiv(p->OH, rb);
    p->Rb = cdiv(p->OH, nb);
    p->nbands = nb;
    p->xrows = p->S == 1 ? p->Rb + 2 : 2 * p->Rb + 1;
    // every hi / lo copy and buffer stays 1024-byte aligned (the swizzle phase is taken from address bits)
    const long long xh = ((long long)p->xrows * p->ncx * p->PWx * WG_ROWB + 1023) & ~1023LL;
    const long long oh = ((long long)p->Rb * p->PWo * WG_ROWB + 1023) & ~1023LL;
    // (sources that are exact in TF32 have no lo copy; packed dY rows hold hi and lo side by side)
    p->x_half = (int)xh; p->x_buf = (int)(((p->nsplit == 2 && !p->xpack) ? 2 : 1) * xh);
    p->o_half = (int)oh; p->o_buf = (int)((p->opack ? 1 : 2) * oh);
    if (wg_smem_bytes(*p) <= (size_t)WG_SMEM_LIMIT) return true;
  }
  return false;
}

}  // namespace tc

static bool wg_shift_enabled() {
  static const bool off = debug_knob("A2C_NO_SHIFT") != 0 || debug_knob("A2C_NO_SHIFT_WGRAD") != 0;
  return !off;
}

long long conv_wgrad_shift_workspace(const a2c_conv_shape* s) {
  return (long long)num_sms() * 128 * tc::WG_NACC * 4 + (long long)(s ? s->Cout + 2 : 2) * 8 + 256;
}

// Output columns per chunk; 0 when the row cannot be tiled.  Planes whose rows are 16-byte multiples
// use the float4 producers (chunks of <= 40 / 20 positions, multiple of 4); narrow odd planes (10x10,
// 5x5: GRUModel's tail) use the scalar producers, whose chunk and halo must fit one warp.
static int wg_chunk_cols(const a2c_conv_shape* s, bool* scalar) {
  const int S = s->stride;
  *scalar = ((s->W | s->OW) & 3) != 0;
  if (*scalar) {
    const int owc = s->OW;                                   // one chunk
    const int xn = S == 1 ? owc + 2 : 2 * owc + 1;
    return (owc <= 32 && xn <= 32) ? owc : 0;
  }
  const int ncol = cdiv(s->OW, S == 1 ? 40 : 20);
  if (s->OW % ncol || ((s->OW / ncol) & 3)) return 0;
  return s->OW / ncol;
}

bool conv_wgrad_shift_ok(const a2c_conv_shape* s) {
  const int S = s->stride;
  if (!wg_shift_enabled() || s->K != 3 || s->pad != 1 || (S != 1 && S != 2)) return false;
  const bool first = s->Cin <= 4 && S == 1;
  if ((!first && (s->Cin & 3)) || (s->Cout & 3)) return false;
  if (S == 2 && ((s->H | s->W) & 1)) return false;
  bool scalar;
  if (wg_chunk_cols(s, &scalar) == 0) return false;
  return !(first && scalar);
}

// true when the shift-GEMM weight gradient of this layer also yields the bias gradient at no extra cost
// (a free 4-slot chunk of the X rows carries the constant 1, see ones_chunk below)
bool conv_wgrad_shift_bias_free(const a2c_conv_shape* s) {
  if (!conv_wgrad_shift_ok(s)) return false;
  const bool first = s->Cin <= 4 && s->stride == 1;
  if (first) return true;
  if (s->Cin <= 16) return false;                              // packed X rows: slots [16, 32) carry the lo parts
  for (int ci0 = 0; ci0 < s->Cin; ci0 += 32)
    if ((s->Cin - ci0 < 32 ? s->Cin - ci0 : 32) <= 28) return true;
  return false;
}

// 0 = ran, 1 = not supported (caller falls back), else error
int conv_wgrad_shift_try(const float* X, const a2c_state_view* xview, const float* dY, const float* Ymask,
                         float* dW, float* dbias, int N, const a2c_conv_shape* s, void* workspace,
                         long long workspace_bytes, cudaStream_t st) {
  const int S = s->stride;
  if ((!X && !xview) || !conv_wgrad_shift_ok(s)) return 1;
  const bool first = s->Cin <= 4 && S == 1;
  if (xview) {
    // uint8 plane ring only: 4-byte aligned rows, one plane per channel
    if (!first || xview->dtype != 0 || xview->fs != s->Cin || (reinterpret_cast<uintptr_t>(xview->base) & 3) ||
        ((xview->stride_r | xview->stride_t | xview->stride_c) & 3))
      return 1;
  }
  bool scalar;
  const int owc = wg_chunk_cols(s, &scalar);
  if (!scalar && ((X && (reinterpret_cast<uintptr_t>(X) & 15)) || (reinterpret_cast<uintptr_t>(dY) & 15) ||
                  (Ymask && (reinterpret_cast<uintptr_t>(Ymask) & 15))))
    return 1;
  if (workspace_bytes < conv_wgrad_shift_workspace(s)) return 1;
  double* bias_acc = reinterpret_cast<double*>(workspace);
  bool bias_done = dbias == nullptr;
  int grid = num_sms();
  // input channels in chunks of 32 (the channel slots of one 128-byte position row)
  for (int ci0 = 0; ci0 < s->Cin; ci0 += 32) {
  const int cin = s->Cin - ci0 < 32 ? s->Cin - ci0 : 32;
  static thread_local tc::WgShiftParams p;
  memset(&p, 0, sizeof(p));
  p.x = xview ? nullptr : X; p.dy = dY; p.mask = Ymask; p.N = N; p.Cin = cin; p.H = s->H; p.W = s->W;
  p.CinTot = s->Cin; p.ci0 = ci0; p.scalar_w = scalar ? 1 : 0;
  if (xview) { p.view = *xview; p.use_view = 1; }
  p.pack3 = first ? 1 : 0;
  { const int np = first ? 1 : (cin / 4 + 1) / 2; p.inv_ncgx = (65536 + np - 1) / np; }
  p.nsplit = xview ? 1 : 2;
  p.opack = s->Cout <= 16 ? 1 : 0;
  p.Cout = s->Cout; p.OH = s->OH; p.OW = s->OW; p.S = S; p.ncx = S;
  p.OWc = owc;
  p.ncol = s->OW / owc;
  p.PWx = S == 1 ? p.OWc + 2 : p.OWc + 1;
  p.PWo = (p.OWc + 7) & ~7;
  p.xpack = (s->Cin <= 16 && !first) ? 1 : 0;
  p.mrows = p.pack3 ? 64 : 128;
  p.a_lbo = p.ncx * p.PWx * tc::WG_ROWB;                         // one source row between the ky atoms
  if (p.a_lbo >= (1 << 18)) return 1;
  if (!tc::wg_plan(&p)) return 1;
  // rows per accumulator chunk: a TMEM chain of at most ~64 MMAs (2-3 per k-step)
  const int ksteps = p.PWo / 8;
  p.rpc = 20 / ksteps;
  if (p.rpc > p.Rb / tc::WG_NACT) p.rpc = p.Rb / tc::WG_NACT;   // and every active issuer warp gets a share of each band
  if (p.rpc < 1) p.rpc = 1;
  {
    // two producer groups when half of the warps finish a band's X and dY tasks in one pass
    const int half = tc::WG_PROD / 2;
    const int ngx = S == 1 ? p.OWc / 4 + 2 : p.OWc / 2 + 1, ngo = p.OWc / 4 > 0 ? p.OWc / 4 : 1;
    const int tpwx = first ? (32 / ngx > 0 ? 32 / ngx : 1) : (32 / (2 * ngx) > 0 ? 32 / (2 * ngx) : 1);
    const int ncgo = (s->Cout < tc::WG_N ? s->Cout : tc::WG_N) / 4;
    p.pair_o = ncgo > 4 ? 1 : 0;
    const int lpo = p.pair_o ? 2 * ngo : ngo;
    const int tpwo = 32 / lpo > 0 ? 32 / lpo : 1;
    const int xtasks = first ? p.Rb + 2 : p.xrows * ((cin / 4 + 1) / 2);
    const int otasks = p.Rb * (p.pair_o ? (ncgo + 1) / 2 : ncgo);
    p.pgroups = (!scalar && xtasks <= half * tpwx && otasks <= half * tpwo) ? 2 : 1;
    if (const int k = debug_knob("A2C_WG_PGROUPS")) p.pgroups = k == 2 ? 2 : 1;
  }
  // bias gradient from the MMA when the X rows have a free 4-slot chunk (first layers: chunk 3 of the
  // packed atom; chunks of 20..28 channels: chunk cin / 4); otherwise the separate masked sum below
  p.ones_chunk = bias_done ? -1 : (first ? 3 : ((!p.xpack && cin <= 28) ? cin / 4 : -1));
  const int bias_row = p.ones_chunk >= 0 ? 4 * p.ones_chunk : -1;     // accumulator row (ky = 0 atom)
  p.partial = reinterpret_cast<float*>(bias_acc + ((s->Cout + 1) & ~1));
  p.dbg = debug_knob("A2C_WG_DBG");
  auto kern = tc::conv_shift_wgrad_kernel;
  A2C_CUDA_CHECK(set_max_dynamic_smem(reinterpret_cast<const void*>(kern), tc::WG_SMEM_LIMIT));
  const int items = N * p.nbands * p.ncol;
  grid = num_sms();
  if (grid > items) grid = items;
  for (int n0 = 0; n0 < s->Cout; n0 += tc::WG_N) {
    p.n0 = n0;
    p.ncgo_valid = (s->Cout - n0 < tc::WG_N ? s->Cout - n0 : tc::WG_N) / 4;
    { const int np = p.pair_o ? (p.ncgo_valid + 1) / 2 : p.ncgo_valid; p.inv_ncgo = (65536 + np - 1) / np; }
    kern<<<grid, tc::WG_THREADS, tc::wg_smem_bytes(p), st>>>(p);
    g_launches++;
    A2C_LAUNCH_CHECK();
    const int nco = s->Cout - n0 < tc::WG_N ? s->Cout - n0 : tc::WG_N;
    tc::wg_shift_reduce_kernel<<<cdiv(nco * cin * 9 + nco, 256), 256, 0, st>>>(
        p.partial, grid, p.pack3 ? 3 : (p.xpack ? 4 : 0), p.mrows, cin, s->Cin, ci0, s->Cout, n0, dW, bias_row,
        bias_row >= 0 ? dbias : nullptr);
    g_launches++;
    A2C_LAUNCH_CHECK();
  }
  if (bias_row >= 0) bias_done = true;
  }
  if (!bias_done) {
    A2C_CUDA_CHECK(cudaMemsetAsync(bias_acc, 0, sizeof(double) * s->Cout, st));
    if (((s->OH * s->OW) & 3) == 0 && (reinterpret_cast<uintptr_t>(dY) & 15) == 0 &&
        (!Ymask || (reinterpret_cast<uintptr_t>(Ymask) & 15) == 0)) {
      const int hw4 = s->OH * s->OW / 4;
      int bx = (int)cdivll((long long)N * hw4, 256 * 4);
      if (bx > 128) bx = 128;
      if (bx < 1) bx = 1;
      tc::wg_bias_sum_kernel<<<dim3(bx, s->Cout), 256, 0, st>>>(dY, Ymask, bias_acc, N, s->Cout, hw4);
    } else {
      int bx = (int)cdivll((long long)N * s->OH * s->OW, 256 * 8);
      if (bx > 128) bx = 128;
      if (bx < 1) bx = 1;
      tc::wg_bias_sum_scalar_kernel<<<dim3(bx, s->Cout), 256, 0, st>>>(dY, Ymask, bias_acc, N, s->Cout,
                                                                       s->OH * s->OW);
    }
    tc::wg_bias_finish_kernel<<<cdiv(s->Cout, 64), 64, 0, st>>>(bias_acc, dbias, s->Cout);
    g_launches += 2;
    A2C_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace a2c

#ifdef A2C_TC_TRACE
extern "C" int a2c_debug_wg_trace(long long* host) {
  return (int)cudaMemcpyFromSymbol(host, a2c::tc::g_wg_trace, sizeof(a2c::tc::g_wg_trace));
}
#endif

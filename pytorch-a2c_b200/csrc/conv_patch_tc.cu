// Patch-staged implicit-GEMM convolutions on tcgen05 (the fast conv path).
//
// The tile of work is an 8 x 16 block of output positions (= 128 GEMM rows) of
// one image.  "Stager" warps copy the input patch that block needs (all channels,
// zero-filled halo) from HBM / L2 into shared memory once -- coalesced, bounds
// checked once per element instead of once per tap.  "Builder" warps then expand
// the patch into K-major SWIZZLE_128B operand tiles with one table-driven shared
// load per element (no address arithmetic, no predicates beyond a select), split
// every value hi/lo for 3xTF32 and hand the stage to the MMA warp through
// mbarriers.  Patches are double-buffered so staging, building, MMA and the
// TMEM drain of consecutive tiles all overlap inside one persistent CTA.
//
//   forward / backward-data ("gather conv"):
//       D[pos, j] = sum_k Patch[c_k][pos_y*SS + dy_k][pos_x*SS + dx_k] * W(j, k)
//     forward:        taps (ci,ky,kx), SS = stride, W(j,k) = wt[j][k]
//     backward-data:  one launch per output parity class (iy % S, ix % S); the taps that
//                     contribute to that class form a stride-1 correlation over dY, so no
//                     multiply-by-zero work is done for strided layers
//   backward-weight:
//       D[co, tap] = sum_pos dY[co][pos] * Patch[c_tap][pos_y*S + ky][pos_x*S + kx]
//     (M side = dY rows, N side = taps built from the patch, K = positions; persistent CTAs
//      accumulate their share of the positions, partials are reduced in double)
//
// Replaces nn.Conv2d forward + autograd of a2c/models.py:98,312,685.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>

#include <atomic>

#include "tc_core.cuh"

namespace a2c {
extern std::atomic<long long> g_launches;
void* tensormap_encoder();   // obs_prep.cu

namespace tc {

constexpr int PT_TH = 8, PT_TW = 16;     // spatial tile
constexpr int MAXK = 448;
constexpr uint32_t NOTAP = 0xFFFFFFFFu;
constexpr uint16_t NOTAP16 = 0xFFFFu;

struct PatchParams {
  const float* src;                      // f32 [N][SC][SH][SW]
  a2c_state_view view;                   // first-layer source when use_view
  int use_view;
  int N, SC, SH, SW;
  int QH, QW;                            // positions per image
  int SS;                                // source step per position
  int y0, x0;                            // source coordinate of patch element (0,0) for position (0,0)
  int PH, PW;
  int tiles_y, tiles_x;
  int nk, nkb;
  const float* wt;                       // fwd / dgrad: W(j,k) = wt[j*wstride_j + wofs[k]]
  long long wstride_j;
  int J;
  const float* dy;                       // wgrad: dY [N][Cout][QH][QW]
  int Cout;
  const float* bias;
  int relu;
  float* out;
  int OutH, OutW, OS, oa, ob;            // fwd / dgrad output placement
  int ncols;                             // wgrad partial pitch
  int stages, patch_bufs, patch_floats;  // patch_floats: SC*PH*PW rounded up to 32
  int use_tma;                           // patch staged by one TMA box per tile (f32 NCHW source)
  int kc;                                // fwd / dgrad: k-blocks chained in TMEM before a drain
  int b_resident;                        // fwd / dgrad: all weight k-blocks live in smem for the whole kernel
  int stage_bytes;                       // bytes of one pipeline stage (A [+ B when streamed])
  // 16-bit tables keep the whole parameter block under the classic 4 KB limit (the TMA
  // descriptor of a kernel using the large-parameter path was rejected as illegal instruction)
  uint16_t delta[MAXK];                  // 0xFFFF = no tap
  uint16_t wofs[MAXK];
};
static_assert(sizeof(PatchParams) + 128 < 4096, "kernel parameters must stay below 4 KB");

// TMEM allocations are powers of two >= 32 columns
__host__ __device__ constexpr int tmem_cols(int n) { return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512; }

constexpr int W_STAGE = 4, W_BUILD = 8, W_DIRECT = 4;            // warps per producer role
constexpr int PW0 = 0, PW1 = W_STAGE, PW2 = W_STAGE + W_BUILD, PW3 = W_STAGE + W_BUILD + W_DIRECT;
constexpr int P_DRAIN0 = PW3, P_MMA = PW3 + NUM_EPI_WARPS, P_ALLOC = P_MMA + 1;
constexpr int P_THREADS = 32 * (P_ALLOC + 1);                    // 26 warps = 832 threads
static_assert(P_DRAIN0 % 4 == 0, "drain warps must start at a multiple of 4");

struct PatchSmem {
  uint8_t* stages;
  uint8_t* bres;
  float* patch;
  uint32_t* tab;
  uint32_t* wtab;
  uint64_t *full, *empty, *acc_full, *acc_empty, *patch_full, *patch_empty;
  uint32_t* tmem_slot;
};

__device__ __forceinline__ PatchSmem carve(uint8_t* smem_raw, const PatchParams& p, int bres_bytes) {
  PatchSmem s;
  // pointer arithmetic on the shared array keeps the shared address space (LDS/STS, not generic)
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  s.stages = base;
  uint8_t* q = base + (size_t)p.stages * p.stage_bytes;
  s.bres = q;
  q += bres_bytes;
  s.patch = reinterpret_cast<float*>(q);
  q += (size_t)p.patch_bufs * p.patch_floats * 4;
  s.tab = reinterpret_cast<uint32_t*>(q); q += MAXK * 4;
  s.wtab = reinterpret_cast<uint32_t*>(q); q += MAXK * 4;
  s.full = reinterpret_cast<uint64_t*>(q); q += 8 * p.stages;
  s.empty = reinterpret_cast<uint64_t*>(q); q += 8 * p.stages;
  s.acc_full = reinterpret_cast<uint64_t*>(q); q += 16;
  s.acc_empty = reinterpret_cast<uint64_t*>(q); q += 16;
  s.patch_full = reinterpret_cast<uint64_t*>(q); q += 32;     // up to 4 patch buffers
  s.patch_empty = reinterpret_cast<uint64_t*>(q); q += 32;
  s.tmem_slot = reinterpret_cast<uint32_t*>(q);
  return s;
}

static size_t patch_smem_bytes(const PatchParams& p, int bres_bytes) {
  return (size_t)p.stages * p.stage_bytes + bres_bytes + (size_t)p.patch_bufs * p.patch_floats * 4 +
         2 * MAXK * 4 + 16 * p.stages + 96 + 16 + 1024;
}

__device__ __forceinline__ void decode_tile(const PatchParams& p, int tile, int& img, int& qy0, int& qx0) {
  const int per = p.tiles_y * p.tiles_x;
  img = tile / per;
  const int r = tile - img * per;
  const int ty = r / p.tiles_x;
  qy0 = ty * PT_TH;
  qx0 = (r - ty * p.tiles_x) * PT_TW;
}

// One 4-D TMA box (PW x PH x SC x 1) per tile; out-of-bounds (halo) elements arrive as zeros.
__device__ __forceinline__ void tma_patch(const CUtensorMap* tmap, float* dst, uint64_t* bar, int x, int y,
                                          int img, uint32_t bytes) {
  mbar_expect_tx(bar, bytes);
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(smem_u32(dst)), "l"(tmap), "r"(x), "r"(y), "r"(0), "r"(img), "r"(smem_u32(bar))
      : "memory");
}

// Stagers: patch[c][py][px] = Src[img][c][qy0*SS + y0 + py][qx0*SS + x0 + px] (0 outside)
__device__ __forceinline__ void stage_patch(const PatchParams& p, float* patch, int img, int qy0, int qx0,
                                            int w, int lane) {
  const int sy0 = qy0 * p.SS + p.y0, sx0 = qx0 * p.SS + p.x0;
  long long ibase = 0;
  uint32_t cmask = 0xffffffffu;
  const uint8_t* p8 = nullptr;
  const float* p32 = p.src;
  long long cstride = (long long)p.SH * p.SW;
  if (p.use_view) {
    const long long rr = img / p.view.T;
    const int tt = img - (int)rr * p.view.T;
    ibase = rr * p.view.stride_r + (long long)tt * p.view.stride_t;
    cstride = p.view.stride_c;
    if (p.view.pflag) {
      const uint8_t* pf = p.view.pflag + rr * p.view.pflag_stride_r + tt;
      cmask = 0;
      bool ok = true;
      for (int c = p.view.fs - 1; c >= 0; --c) {
        if (ok) cmask |= 1u << c;
        ok = ok && (pf[c] == 0);
      }
    }
    if (p.view.dtype == 0) { p8 = static_cast<const uint8_t*>(p.view.base); p32 = nullptr; }
    else p32 = static_cast<const float*>(p.view.base);
  } else {
    ibase = (long long)img * p.SC * cstride;
  }
  // flat element index e = (c*PH + py)*PW + px; 12 independent loads in flight per thread
  const int total = p.SC * p.PH * p.PW;
  const int ts = w * 32 + lane;
  const float inv_pw = 1.0f / (float)p.PW, inv_ph = 1.0f / (float)p.PH;
  constexpr int NB = 12;
  for (int eb = 0; eb < total; eb += NB * 32 * W_STAGE) {
    float v[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      const int e = eb + i * (32 * W_STAGE) + ts;
      v[i] = 0.0f;
      if (e < total) {
        int row = (int)(((float)e + 0.5f) * inv_pw);
        if (row * p.PW > e) --row;
        if ((row + 1) * p.PW <= e) ++row;
        const int px = e - row * p.PW;
        int c = (int)(((float)row + 0.5f) * inv_ph);
        if (c * p.PH > row) --c;
        if ((c + 1) * p.PH <= row) ++c;
        const int py = row - c * p.PH;
        const int sy = sy0 + py, sx = sx0 + px;
        if ((unsigned)sy < (unsigned)p.SH && (unsigned)sx < (unsigned)p.SW && ((cmask >> (c & 31)) & 1)) {
          const long long a = ibase + (long long)c * cstride + (long long)sy * p.SW + sx;
          v[i] = p8 ? (float)p8[a] : __ldg(p32 + a);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      const int e = eb + i * (32 * W_STAGE) + ts;
      if (e < total) patch[e] = v[i];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// forward / backward-data
// ------------------------------------------------------------------------------------------------
template <int BN, bool SPLIT3>
__global__ void __launch_bounds__(P_THREADS, 1) conv_patch_fwd_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ PatchParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int A_BYTES = TM * 128, B_BYTES = BN * 128;
  constexpr int APAIR = (SPLIT3 ? 2 : 1) * A_BYTES, BPAIR = (SPLIT3 ? 2 : 1) * B_BYTES;
  constexpr int HALF = BN / 2;
  const int KC = p.kc;   // forward: 1 (activations feed ReLU masks: keep the truncating TMEM chains short)
  PatchSmem sm = carve(smem_raw, p, p.b_resident ? p.nkb * BPAIR : 0);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int ntiles = p.N * p.tiles_y * p.tiles_x;
  const int nchunks = (p.nkb + KC - 1) / KC;

  for (int k = threadIdx.x; k < MAXK; k += P_THREADS) {
    sm.tab[k] = (k < p.nkb * KB && p.delta[k] != NOTAP16) ? (uint32_t)p.delta[k] : NOTAP;
    sm.wtab[k] = k < p.nkb * KB ? (uint32_t)p.wofs[k] : 0;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&sm.full[s], p.b_resident ? W_BUILD : W_BUILD + W_DIRECT);
      mbar_init(&sm.empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&sm.acc_full[b], 1);
      mbar_init(&sm.acc_empty[b], NUM_EPI_WARPS);
    }
    for (int b = 0; b < p.patch_bufs; ++b) {
      mbar_init(&sm.patch_full[b], p.use_tma ? 1 : W_STAGE);
      mbar_init(&sm.patch_empty[b], W_BUILD);
    }
    mbar_fence_init();
  }
  if (warp == P_ALLOC) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(sm.tmem_slot)),
                 "r"((uint32_t)tmem_cols(2 * BN)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  __syncthreads();      // tables visible
  if (p.b_resident) {
    // the weights are the same for every tile: build all their k-blocks once
    const int units = p.nkb * BN * 8;
    for (int u = threadIdx.x; u < units; u += P_THREADS) {
      const int kb = u / (BN * 8), r = u - kb * (BN * 8), jrow = r % BN, c = r / BN;
      float v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = kb * KB + c * 4 + e;
        v[e] = (jrow < p.J && k < p.nk) ? __ldg(p.wt + (long long)jrow * p.wstride_j + sm.wtab[k]) : 0.0f;
      }
      uint8_t* b = sm.bres + (size_t)kb * BPAIR;
      store_chunk<SPLIT3>(b, b + B_BYTES, jrow, c, make_float4(v[0], v[1], v[2], v[3]));
    }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *sm.tmem_slot, 0);
  auto stage_ptr = [&](int s) { return sm.stages + (size_t)s * p.stage_bytes; };

  if (warp < PW1) {
    // ---------------- stagers ----------------
    int j = 0;
    StageRing pring(p.patch_bufs);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++j, pring.next()) {
      const int pb = pring.s;
      const int use = pring.par;
      int img, qy0, qx0;
      decode_tile(p, tile, img, qy0, qx0);
      if (p.use_tma) {
        if (warp == 0 && lane == 0) {
          mbar_wait_relaxed(&sm.patch_empty[pb], (use & 1) ^ 1);
          tma_patch(&tmap, sm.patch + (size_t)pb * p.patch_floats, &sm.patch_full[pb],
                    qx0 * p.SS + p.x0, qy0 * p.SS + p.y0, img, (uint32_t)(p.SC * p.PH * p.PW * 4));
        }
      } else {
        mbar_wait_relaxed(&sm.patch_empty[pb], (use & 1) ^ 1);
        stage_patch(p, sm.patch + (size_t)pb * p.patch_floats, img, qy0, qx0, warp, lane);
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.patch_full[pb]);
      }
    }
  } else if (warp < PW2) {
    // ---------------- builders: A tile from the patch ----------------
    const int t = threadIdx.x - 32 * PW1;          // 0..255
    const int row = t & 127, half = t >> 7;
    const int ty = row >> 4, tx = row & 15;
    const int base = ty * p.SS * p.PW + tx * p.SS;
    int j = 0, it = 0;
    StageRing ring(p.stages);
    StageRing pring(p.patch_bufs);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++j, pring.next()) {
      const int pb = pring.s;
      const int use = pring.par;
      int img, qy0, qx0;
      decode_tile(p, tile, img, qy0, qx0);
      const bool posok = (qy0 + ty < p.QH) && (qx0 + tx < p.QW);
      mbar_wait(&sm.patch_full[pb], use & 1);
      const float* patch = sm.patch + (size_t)pb * p.patch_floats + base;
      for (int kb = 0; kb < p.nkb; ++kb, ++it, ring.next()) {
        const uint32_t* tab = sm.tab + kb * KB + half * 16;
        float v[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const uint32_t d = tab[e];
          const float x = patch[d == NOTAP ? 0 : d];
          v[e] = (d != NOTAP && posok) ? x : 0.0f;
        }
        const int s = ring.s;
        mbar_wait(&sm.empty[s], ring.par ^ 1);
        uint8_t* st = stage_ptr(s);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          store_chunk<SPLIT3>(st, st + A_BYTES, row, half * 4 + c,
                              make_float4(v[c * 4], v[c * 4 + 1], v[c * 4 + 2], v[c * 4 + 3]));
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.full[s]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.patch_empty[pb]);
    }
  } else if (warp < PW3) {
    // ---------------- weights: B tile (only when they do not fit in shared memory) ----------------
    const int t = threadIdx.x - 32 * PW2;          // 0..127
    int it = 0;
    StageRing ring(p.stages);
    for (int tile = blockIdx.x; tile < (p.b_resident ? 0 : ntiles); tile += gridDim.x) {
      for (int kb = 0; kb < p.nkb; ++kb, ++it, ring.next()) {
        float v[BN / 4];
#pragma unroll
        for (int i = 0; i < BN / 16; ++i) {
          const int u = i * 128 + t, jrow = u % BN, c = u / BN;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int k = kb * KB + c * 4 + e;
            v[i * 4 + e] = (jrow < p.J && k < p.nk) ? __ldg(p.wt + (long long)jrow * p.wstride_j + sm.wtab[k]) : 0.0f;
          }
        }
        const int s = ring.s;
        mbar_wait(&sm.empty[s], ring.par ^ 1);
        uint8_t* b = stage_ptr(s) + APAIR;
#pragma unroll
        for (int i = 0; i < BN / 16; ++i) {
          const int u = i * 128 + t, jrow = u % BN, c = u / BN;
          store_chunk<SPLIT3>(b, b + B_BYTES, jrow, c,
                              make_float4(v[i * 4], v[i * 4 + 1], v[i * 4 + 2], v[i * 4 + 3]));
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.full[s]);
      }
    }
  } else if (warp == P_MMA) {
    // convergent warp, one elected lane, descriptor low words in uniform registers (mma_tf32_imm)
    constexpr uint32_t idesc = make_idesc(BN);
    constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO, version, SWIZZLE_128B
    const uint32_t a_w0 = ((smem_u32(sm.stages) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t b_res0 = ((smem_u32(sm.bres) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t stage_step = (uint32_t)p.stage_bytes >> 4;
    StageRing ring(p.stages);
    int chunk = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      int inchunk = 0;
      for (int i = 0; i < p.nkb; ++i, ring.next()) {
        const int buf = chunk & 1;
        if (inchunk == 0) {
          mbar_wait(&sm.acc_empty[buf], ((chunk >> 1) & 1) ^ 1);
          tc_fence_after();
        }
        mbar_wait(&sm.full[ring.s], ring.par);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * BN;
        const uint32_t a_hi = a_w0 + ring.s * stage_step;
        const uint32_t a_lo = a_hi + (A_BYTES >> 4);
        const uint32_t b_hi = p.b_resident ? b_res0 + (uint32_t)i * (BPAIR >> 4) : a_hi + (APAIR >> 4);
        const uint32_t b_lo = b_hi + (B_BYTES >> 4);
        const bool last = (inchunk + 1 == KC) || (i == p.nkb - 1);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < KB / 8; ++kk) {
            const uint32_t ko = kk * 2;      // 32 bytes per UMMA_K step
            const uint32_t first = (inchunk | kk) ? 1u : 0u;
            if constexpr (SPLIT3) {
              mma_tf32_imm<idesc, DESC_HI>(tmem_d, a_lo + ko, b_hi + ko, first);
              mma_tf32_imm<idesc, DESC_HI>(tmem_d, a_hi + ko, b_lo + ko, 1u);
              mma_tf32_imm<idesc, DESC_HI>(tmem_d, a_hi + ko, b_hi + ko, 1u);
            } else {
              mma_tf32_imm<idesc, DESC_HI>(tmem_d, a_hi + ko, b_hi + ko, first);
            }
          }
          mma_commit(&sm.empty[ring.s]);
          if (last) mma_commit(&sm.acc_full[buf]);
        }
        __syncwarp();
        if (last) { inchunk = 0; ++chunk; } else ++inchunk;
      }
    }
  } else if (warp >= P_DRAIN0 && warp < P_DRAIN0 + NUM_EPI_WARPS) {
    const int e = warp - P_DRAIN0, q = e & 3, hsel = e >> 2;
    const int row = q * 32 + lane, ty = row >> 4, tx = row & 15;
    int gchunk = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      float acc[HALF];
#pragma unroll
      for (int jj = 0; jj < HALF; ++jj) acc[jj] = 0.0f;
      for (int chunk = 0; chunk < nchunks; ++chunk, ++gchunk) {
        const int buf = gchunk & 1;
        mbar_wait_relaxed(&sm.acc_full[buf], (gchunk >> 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + hsel * HALF;
#pragma unroll
        for (int c0 = 0; c0 < HALF; c0 += 16) {
          uint32_t r[16];
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
              : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
                "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]),
                "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
              : "r"(taddr + (uint32_t)c0));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) acc[c0 + jj] += __uint_as_float(r[jj]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.acc_empty[buf]);
      }
      int img, qy0, qx0;
      decode_tile(p, tile, img, qy0, qx0);
      const int qy = qy0 + ty, qx = qx0 + tx;
      if (qy < p.QH && qx < p.QW) {
        const long long plane = (long long)p.OutH * p.OutW;
        float* dst = p.out + (long long)img * p.J * plane + (long long)(qy * p.OS + p.oa) * p.OutW +
                     qx * p.OS + p.ob;
#pragma unroll
        for (int jj = 0; jj < HALF; ++jj) {
          const int jcol = hsel * HALF + jj;
          if (jcol < p.J) {
            float x = acc[jj] + (p.bias ? p.bias[jcol] : 0.0f);
            if (p.relu) x = fmaxf(x, 0.0f);
            dst[(long long)jcol * plane] = x;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == P_ALLOC) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)tmem_cols(2 * BN)));
  }
}

// ------------------------------------------------------------------------------------------------
// backward weight: D[tap, co] = sum_pos Patch(tap, pos) * dY[co][pos]
//   M side (128 rows) = taps built from the patch, N side (BN = 32 / 64) = dY rows, K = positions.
//   (taps on M keeps the 128-row MMA mostly full; with Cout on M 80 % of every MMA was padding)
// ------------------------------------------------------------------------------------------------
template <int BN, bool SPLIT3>
__global__ void __launch_bounds__(P_THREADS, 1) conv_patch_wgrad_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ PatchParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int A_BYTES = TM * 128, B_BYTES = BN * 128;
  constexpr int STAGE_BYTES = (SPLIT3 ? 2 : 1) * (A_BYTES + B_BYTES);
  constexpr int HALF = BN / 2;
  constexpr int KC = 1;                       // shortest TMEM chains: truncation is a bias (tc_core.cuh)
  constexpr int KBT = (PT_TH * PT_TW) / KB;   // 4 k-blocks (of 32 positions) per spatial tile
  PatchSmem sm = carve(smem_raw, p, 0);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int ntiles = p.N * p.tiles_y * p.tiles_x;
  const int m0 = blockIdx.x * TM;             // first tap row of this CTA
  int my_tiles = 0;
  for (int tile = blockIdx.y; tile < ntiles; tile += gridDim.y) ++my_tiles;
  const int nkb = my_tiles * KBT;
  const int nchunks = (nkb + KC - 1) / KC;

  for (int k = threadIdx.x; k < MAXK; k += P_THREADS)
    sm.tab[k] = (k < p.nk && p.delta[k] != NOTAP16) ? (uint32_t)p.delta[k] : NOTAP;
  // tap rows >= nk and dY rows >= Cout are never written: clear every stage once
  for (int i = threadIdx.x; i < p.stages * STAGE_BYTES / 16; i += P_THREADS)
    reinterpret_cast<uint4*>(sm.stages)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&sm.full[s], W_BUILD + W_DIRECT);
      mbar_init(&sm.empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&sm.acc_full[b], 1);
      mbar_init(&sm.acc_empty[b], NUM_EPI_WARPS);
    }
    for (int b = 0; b < p.patch_bufs; ++b) {
      mbar_init(&sm.patch_full[b], p.use_tma ? 1 : W_STAGE);
      mbar_init(&sm.patch_empty[b], W_BUILD);
    }
    mbar_fence_init();
  }
  if (warp == P_ALLOC) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(sm.tmem_slot)),
                 "r"((uint32_t)tmem_cols(2 * BN)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *sm.tmem_slot, 0);
  auto stage_ptr = [&](int s) { return sm.stages + (size_t)s * STAGE_BYTES; };

  if (warp < PW1) {
    int j = 0;
    StageRing pring(p.patch_bufs);
    for (int tile = blockIdx.y; tile < ntiles; tile += gridDim.y, ++j, pring.next()) {
      const int pb = pring.s;
      int img, qy0, qx0;
      decode_tile(p, tile, img, qy0, qx0);
      if (p.use_tma) {
        if (warp == 0 && lane == 0) {
          mbar_wait_relaxed(&sm.patch_empty[pb], pring.par ^ 1);
          tma_patch(&tmap, sm.patch + (size_t)pb * p.patch_floats, &sm.patch_full[pb],
                    qx0 * p.SS + p.x0, qy0 * p.SS + p.y0, img, (uint32_t)(p.SC * p.PH * p.PW * 4));
        }
      } else {
        mbar_wait_relaxed(&sm.patch_empty[pb], pring.par ^ 1);
        stage_patch(p, sm.patch + (size_t)pb * p.patch_floats, img, qy0, qx0, warp, lane);
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.patch_full[pb]);
      }
    }
  } else if (warp < PW2) {
    // ---------------- builders: A tile (128 taps x 32 positions) from the patch ----------------
    const int w = warp - PW1;                   // 0..7, owns tap rows w*16 .. +15
    const bool any = m0 + w * 16 < p.nk;        // rows beyond the last tap stay zero
    uint32_t d[16];
#pragma unroll
    for (int rr = 0; rr < 16; ++rr) {
      const int tap = m0 + w * 16 + rr;
      d[rr] = tap < MAXK ? sm.tab[tap] : NOTAP;
    }
    int j = 0;
    StageRing ring(p.stages), pring(p.patch_bufs);
    for (int tile = blockIdx.y; tile < ntiles; tile += gridDim.y, ++j, pring.next()) {
      const int pb = pring.s;
      int img, qy0, qx0;
      decode_tile(p, tile, img, qy0, qx0);
      mbar_wait(&sm.patch_full[pb], pring.par);
      const float* patch = sm.patch + (size_t)pb * p.patch_floats;
      for (int q = 0; q < KBT; ++q, ring.next()) {
        const int ty = 2 * q + (lane >> 4), tx = lane & 15;
        const bool posok = (qy0 + ty < p.QH) && (qx0 + tx < p.QW);
        const float* pp = patch + ty * p.SS * p.PW + tx * p.SS;
        float v[16];
        if (any) {
#pragma unroll
          for (int rr = 0; rr < 16; ++rr) {
            const float x = pp[d[rr] == NOTAP ? 0 : d[rr]];
            v[rr] = (d[rr] != NOTAP && posok) ? x : 0.0f;
          }
        }
        mbar_wait(&sm.empty[ring.s], ring.par ^ 1);
        if (any) {
          uint8_t* a = stage_ptr(ring.s);
#pragma unroll
          for (int rr = 0; rr < 16; ++rr) store_elem<SPLIT3>(a, a + A_BYTES, w * 16 + rr, lane, v[rr]);
          fence_proxy_async();
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.full[ring.s]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.patch_empty[pb]);
    }
  } else if (warp < PW3) {
    // ---------------- dY: B tile rows co < Cout, k = position ----------------
    // (global loads of k-block i+1 are issued before k-block i is converted and stored)
    const int w = warp - PW2;                   // 0..3, rows w*(BN/4) .. +BN/4
    constexpr int RW = BN / W_DIRECT;
    const long long plane = (long long)p.QH * p.QW;
    auto load = [&](int tile, int q, float (&v)[RW]) {
      int img, qy0, qx0;
      decode_tile(p, tile, img, qy0, qx0);
      const int qy = qy0 + 2 * q + (lane >> 4), qx = qx0 + (lane & 15);
      const bool posok = qy < p.QH && qx < p.QW;
      const float* src = p.dy + ((long long)img * p.Cout * p.QH + qy) * p.QW + qx;
#pragma unroll
      for (int rr = 0; rr < RW; ++rr) {
        const int co = w * RW + rr;
        v[rr] = (posok && co < p.Cout) ? __ldg(src + co * plane) : 0.0f;
      }
    };
    int tile = blockIdx.y, q = 0;
    StageRing ring(p.stages);
    float cur[RW], nxt[RW];
    if (tile < ntiles) load(tile, 0, cur);
    while (tile < ntiles) {
      int nq = q + 1, ntile = tile;
      if (nq == KBT) { nq = 0; ntile = tile + gridDim.y; }
      if (ntile < ntiles) load(ntile, nq, nxt);
      mbar_wait(&sm.empty[ring.s], ring.par ^ 1);
      uint8_t* b = stage_ptr(ring.s) + (SPLIT3 ? 2 : 1) * A_BYTES;
#pragma unroll
      for (int rr = 0; rr < RW; ++rr)
        if (w * RW + rr < p.Cout) store_elem<SPLIT3>(b, b + B_BYTES, w * RW + rr, lane, cur[rr]);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.full[ring.s]);
#pragma unroll
      for (int rr = 0; rr < RW; ++rr) cur[rr] = nxt[rr];
      tile = ntile; q = nq; ring.next();
    }
  } else if (warp == P_MMA) {
    // convergent warp, one elected lane, descriptor low words in uniform registers (mma_tf32_imm)
    constexpr uint32_t idesc = make_idesc(BN);
    constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO, version, SWIZZLE_128B
    const uint32_t a_w0 = ((smem_u32(sm.stages) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t stage_step = (uint32_t)STAGE_BYTES >> 4;
    StageRing ring(p.stages);
    int chunk = 0, inchunk = 0;
    for (int i = 0; i < nkb; ++i, ring.next()) {
      const int buf = chunk & 1;
      if (inchunk == 0) {
        mbar_wait(&sm.acc_empty[buf], ((chunk >> 1) & 1) ^ 1);
        tc_fence_after();
      }
      mbar_wait(&sm.full[ring.s], ring.par);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + buf * BN;
      const uint32_t a_hi = a_w0 + ring.s * stage_step;
      const uint32_t a_lo = a_hi + (A_BYTES >> 4);
      const uint32_t b_hi = a_hi + (((SPLIT3 ? 2 : 1) * A_BYTES) >> 4);
      const uint32_t b_lo = b_hi + (B_BYTES >> 4);
      const bool last = (inchunk + 1 == KC) || (i == nkb - 1);
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < KB / 8; ++kk) {
          const uint32_t ko = kk * 2;
          const uint32_t first = (inchunk | kk) ? 1u : 0u;
          if constexpr (SPLIT3) {
            mma_tf32_imm<idesc, DESC_HI>(tmem_d, a_lo + ko, b_hi + ko, first);
            mma_tf32_imm<idesc, DESC_HI>(tmem_d, a_hi + ko, b_lo + ko, 1u);
            mma_tf32_imm<idesc, DESC_HI>(tmem_d, a_hi + ko, b_hi + ko, 1u);
          } else {
            mma_tf32_imm<idesc, DESC_HI>(tmem_d, a_hi + ko, b_hi + ko, first);
          }
        }
        mma_commit(&sm.empty[ring.s]);
        if (last) mma_commit(&sm.acc_full[buf]);
      }
      __syncwarp();
      if (last) { inchunk = 0; ++chunk; } else ++inchunk;
    }
  } else if (warp >= P_DRAIN0 && warp < P_DRAIN0 + NUM_EPI_WARPS) {
    const int e = warp - P_DRAIN0, q = e & 3, hsel = e >> 2;
    const int tap = m0 + q * 32 + lane;
    float acc[HALF];
#pragma unroll
    for (int jj = 0; jj < HALF; ++jj) acc[jj] = 0.0f;
    for (int chunk = 0; chunk < nchunks; ++chunk) {
      const int buf = chunk & 1;
      mbar_wait_relaxed(&sm.acc_full[buf], (chunk >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + hsel * HALF;
#pragma unroll
      for (int c0 = 0; c0 < HALF; c0 += 16) {
        uint32_t r[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
              "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
              "=r"(r[14]), "=r"(r[15])
            : "r"(taddr + (uint32_t)c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) acc[c0 + jj] += __uint_as_float(r[jj]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.acc_empty[buf]);
    }
    if (tap < p.nk) {
      // partial[g][tap][co] ; ncols = BN
      float* dst = p.out + ((long long)blockIdx.y * p.nk + tap) * p.ncols + hsel * HALF;
#pragma unroll
      for (int jj = 0; jj < HALF; jj += 4)
        *reinterpret_cast<float4*>(dst + jj) = make_float4(acc[jj], acc[jj + 1], acc[jj + 2], acc[jj + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == P_ALLOC) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)tmem_cols(2 * BN)));
  }
}

// dW[co][tap] = sum_g partial[g][tap][co]   (double accumulation)
__global__ void __launch_bounds__(256)
patch_wgrad_reduce_kernel(const float* __restrict__ partial, int groups, int cout, int ncols, int taps,
                          float* __restrict__ dW) {
  const int total = cout * taps;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int tap = e / cout, co = e - tap * cout;
    double s = 0.0;
    for (int g = 0; g < groups; ++g) s += (double)partial[((long long)g * taps + tap) * ncols + co];
    dW[(long long)co * taps + tap] = (float)s;
  }
}

// ---------------------------------------------------------------------------- host side ----------
static int smem_limit() { return 227 * 1024; }

// picks stages / patch buffers; returns false when the layer does not fit
// apair / bpair: bytes of the A / B part of a stage; nkb_b > 0 lets the B k-blocks be resident
static bool plan_smem(PatchParams* p, int apair, int bpair, int nkb_b, int* bres_bytes) {
  p->patch_floats = (p->SC * p->PH * p->PW + 31) & ~31;
  for (int resident = (nkb_b > 0 && nkb_b * bpair <= 72 * 1024) ? 1 : 0; resident >= 0; --resident) {
    p->b_resident = resident;
    p->stage_bytes = apair + (resident ? 0 : bpair);
    *bres_bytes = resident ? nkb_b * bpair : 0;
    for (int bufs = 2; bufs >= 1; --bufs)
      for (int st = 4; st >= 2; --st) {
        p->stages = st; p->patch_bufs = bufs;
        if (patch_smem_bytes(*p, *bres_bytes) <= (size_t)smem_limit()) return true;
      }
  }
  return false;
}

// Tensor map for the patch source (f32 NCHW).  Needs 16-byte multiples for the row pitch of the
// source and of the box; otherwise the register staging path is used.
static bool make_patch_tmap(PatchParams& p, CUtensorMap* tmap) {
  memset(tmap, 0, sizeof(*tmap));
  p.use_tma = 0;
  if (p.use_view || !p.src) return true;
  if (debug_knob("A2C_NO_TMA")) return true;
  if ((p.SW % 4) != 0 || (p.PW % 4) != 0 || (reinterpret_cast<uintptr_t>(p.src) & 15)) return true;
  if (p.PW > 256 || p.PH > 256 || p.SC > 256) return true;
  auto encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(tensormap_encoder());
  if (!encode) return true;
  cuuint64_t dims[4] = {(cuuint64_t)p.SW, (cuuint64_t)p.SH, (cuuint64_t)p.SC, (cuuint64_t)p.N};
  cuuint64_t strides[3] = {(cuuint64_t)p.SW * 4, (cuuint64_t)p.SW * p.SH * 4, (cuuint64_t)p.SW * p.SH * p.SC * 4};
  cuuint32_t box[4] = {(cuuint32_t)p.PW, (cuuint32_t)p.PH, (cuuint32_t)p.SC, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult cr = encode(tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(p.src), dims, strides, box,
                       estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr == CUDA_SUCCESS) p.use_tma = 1;
  return true;
}

template <int BN, bool SPLIT3>
static int launch_fwd_like(PatchParams& p, cudaStream_t st) {
  constexpr int AP = (SPLIT3 ? 2 : 1) * TM * 128, BP = (SPLIT3 ? 2 : 1) * BN * 128;
  int bres = 0;
  if (!plan_smem(&p, AP, BP, p.nkb, &bres)) return 1;
  auto kern = conv_patch_fwd_kernel<BN, SPLIT3>;
  A2C_CUDA_CHECK(set_max_dynamic_smem(reinterpret_cast<const void*>(kern), smem_limit()));
  const int ntiles = p.N * p.tiles_y * p.tiles_x;
  int grid = num_sms();
  if (grid > ntiles) grid = ntiles;
  CUtensorMap tmap;
  make_patch_tmap(p, &tmap);
  kern<<<grid, P_THREADS, patch_smem_bytes(p, bres), st>>>(tmap, p);
  g_launches++;
  A2C_LAUNCH_CHECK();
  return 0;
}

static int dispatch_fwd_like(PatchParams& p, bool split3, cudaStream_t st) {
  const int bn = p.J > 32 ? 64 : 32;
  if (p.J > 64) return 1;
  if (split3) return bn == 64 ? launch_fwd_like<64, true>(p, st) : launch_fwd_like<32, true>(p, st);
  return bn == 64 ? launch_fwd_like<64, false>(p, st) : launch_fwd_like<32, false>(p, st);
}

static void set_source(PatchParams* p, const float* src, const a2c_state_view* view, int N, int C, int H, int W) {
  p->src = src; p->N = N; p->SC = C; p->SH = H; p->SW = W;
  if (view) { p->view = *view; p->use_view = 1; }
}

static int floordiv(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }

}  // namespace tc

using tc::PatchParams;

// 0 = ran, 1 = not supported (caller falls back), else error
int conv_fwd_patch_try(const float* X, const a2c_state_view* xview, const float* Wt, const float* bias,
                       float* Y, int N, const a2c_conv_shape* s, int relu, int precision, cudaStream_t st) {
  const int KK = s->K * s->K, nk = s->Cin * KK;
  if (nk > tc::MAXK || s->Cout > 64) return 1;
  static thread_local PatchParams p;
  memset(&p, 0, sizeof(p));
  tc::set_source(&p, X, xview, N, s->Cin, s->H, s->W);
  // TMA needs a 16-byte aligned box start: shift the patch origin left to a multiple of 4 floats
  const int xoff = ((-s->pad % 4) + 4) % 4;
  p.QH = s->OH; p.QW = s->OW; p.SS = s->stride; p.y0 = -s->pad; p.x0 = -s->pad - xoff;
  p.PH = (tc::PT_TH - 1) * s->stride + s->K; p.PW = ((tc::PT_TW - 1) * s->stride + s->K + xoff + 3) & ~3;
  p.tiles_y = cdiv(s->OH, tc::PT_TH); p.tiles_x = cdiv(s->OW, tc::PT_TW);
  p.nk = nk; p.nkb = cdiv(nk, tc::KB);
  for (int k = 0; k < nk; ++k) {
    const int ci = k / KK, r = k % KK, ky = r / s->K, kx = r % s->K;
    p.delta[k] = (uint16_t)(ci * p.PH * p.PW + ky * p.PW + kx + xoff);
    p.wofs[k] = (uint16_t)k;
  }
  for (int k = nk; k < p.nkb * tc::KB; ++k) p.delta[k] = tc::NOTAP16;
  p.wt = Wt; p.wstride_j = nk; p.J = s->Cout; p.bias = bias; p.relu = relu;
  p.out = Y; p.OutH = s->OH; p.OutW = s->OW; p.OS = 1; p.oa = 0; p.ob = 0;
  p.kc = 1;
  return tc::dispatch_fwd_like(p, precision == 1, st);
}

int conv_dgrad_patch_try(const float* dY, const float* Wt, float* dX, int N, const a2c_conv_shape* s,
                         int precision, cudaStream_t st) {
  const int K = s->K, S = s->stride, P = s->pad, KK = K * K;
  if (s->Cin > 64) return 1;
  // feasibility over all parity classes first (so we never leave dX half written)
  for (int a = 0; a < S; ++a) {
    int ny = 0;
    for (int ky = 0; ky < K; ++ky) if (((a + P - ky) % S + S) % S == 0) ++ny;
    if (ny == 0) return 1;
    for (int b = 0; b < S; ++b) {
      int nx = 0;
      for (int kx = 0; kx < K; ++kx) if (((b + P - kx) % S + S) % S == 0) ++nx;
      if (nx == 0 || s->Cout * ny * nx > tc::MAXK) return 1;
    }
  }
  static thread_local PatchParams p;
  for (int a = 0; a < S; ++a)
    for (int b = 0; b < S; ++b) {
      if (a >= s->H || b >= s->W) continue;
      memset(&p, 0, sizeof(p));
      tc::set_source(&p, dY, nullptr, N, s->Cout, s->OH, s->OW);
      int dys[16], kys[16], dxs[16], kxs[16], ny = 0, nx = 0;
      for (int ky = 0; ky < K; ++ky)
        if (((a + P - ky) % S + S) % S == 0) { dys[ny] = tc::floordiv(a + P - ky, S); kys[ny++] = ky; }
      for (int kx = 0; kx < K; ++kx)
        if (((b + P - kx) % S + S) % S == 0) { dxs[nx] = tc::floordiv(b + P - kx, S); kxs[nx++] = kx; }
      int dymin = dys[0], dymax = dys[0], dxmin = dxs[0], dxmax = dxs[0];
      for (int i = 0; i < ny; ++i) { dymin = dys[i] < dymin ? dys[i] : dymin; dymax = dys[i] > dymax ? dys[i] : dymax; }
      for (int i = 0; i < nx; ++i) { dxmin = dxs[i] < dxmin ? dxs[i] : dxmin; dxmax = dxs[i] > dxmax ? dxs[i] : dxmax; }
      p.QH = cdiv(s->H - a, S); p.QW = cdiv(s->W - b, S);
      const int xoff = ((dxmin % 4) + 4) % 4;      // 16-byte aligned TMA box start
      p.SS = 1; p.y0 = dymin; p.x0 = dxmin - xoff;
      p.PH = tc::PT_TH + (dymax - dymin); p.PW = (tc::PT_TW + (dxmax - dxmin) + xoff + 3) & ~3;
      p.tiles_y = cdiv(p.QH, tc::PT_TH); p.tiles_x = cdiv(p.QW, tc::PT_TW);
      int k = 0;
      for (int co = 0; co < s->Cout; ++co)
        for (int iy = 0; iy < ny; ++iy)
          for (int ix = 0; ix < nx; ++ix, ++k) {
            p.delta[k] = (uint16_t)(co * p.PH * p.PW + (dys[iy] - dymin) * p.PW + (dxs[ix] - dxmin + xoff));
            p.wofs[k] = (uint16_t)(co * s->Cin * KK + kys[iy] * K + kxs[ix]);
          }
      p.nk = k; p.nkb = cdiv(k, tc::KB);
      for (int kk = k; kk < p.nkb * tc::KB; ++kk) p.delta[kk] = tc::NOTAP16;
      p.wt = Wt; p.wstride_j = KK; p.J = s->Cin; p.bias = nullptr; p.relu = 0;
      p.out = dX; p.OutH = s->H; p.OutW = s->W; p.OS = S; p.oa = a; p.ob = b;
      p.kc = 8;
      int rc = tc::dispatch_fwd_like(p, precision == 1, st);
      if (rc) return rc == 1 && (a | b) ? A2C_ERR_ARG : rc;
    }
  return 0;
}

long long conv_wgrad_patch_workspace(const a2c_conv_shape* s) {
  const int taps = s->Cin * s->K * s->K;
  const int mtiles = cdiv(taps, tc::TM);
  const int bn = s->Cout > 32 ? 64 : 32;
  int groups = num_sms() / mtiles;
  if (groups < 1) groups = 1;
  return (long long)groups * taps * bn * 4 + (long long)s->Cout * 8 + 64;
}

namespace tc {
template <int BN, bool SPLIT3>
static int launch_wgrad(PatchParams& p, int ntiles_n, int groups, cudaStream_t st) {
  constexpr int AP = (SPLIT3 ? 2 : 1) * TM * 128, BP = (SPLIT3 ? 2 : 1) * BN * 128;
  int bres = 0;
  if (!plan_smem(&p, AP, BP, 0, &bres)) return 1;
  auto kern = conv_patch_wgrad_kernel<BN, SPLIT3>;
  A2C_CUDA_CHECK(set_max_dynamic_smem(reinterpret_cast<const void*>(kern), smem_limit()));
  CUtensorMap tmap;
  make_patch_tmap(p, &tmap);
  kern<<<dim3(ntiles_n, groups), P_THREADS, patch_smem_bytes(p, 0), st>>>(tmap, p);
  g_launches++;
  A2C_LAUNCH_CHECK();
  return 0;
}
// per-channel sum over N and positions (bias gradient): X [N][C][HW]
__global__ void __launch_bounds__(256)
patch_channel_sum_kernel(const float* __restrict__ X, double* __restrict__ acc, int N, int C, int HW) {
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
__global__ void patch_channel_sum_finish_kernel(const double* __restrict__ acc, float* __restrict__ out, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) out[c] = (float)acc[c];
}
}  // namespace tc

int conv_wgrad_patch_try(const float* X, const a2c_state_view* xview, const float* dY, float* dW,
                         float* dbias, int N, const a2c_conv_shape* s, void* workspace,
                         long long workspace_bytes, int precision, cudaStream_t st) {
  const int KK = s->K * s->K, taps = s->Cin * KK;
  if (taps > tc::MAXK || s->Cout > 64) return 1;
  if (workspace_bytes < conv_wgrad_patch_workspace(s)) return 1;
  const int mtiles = cdiv(taps, tc::TM);
  const int bn = s->Cout > 32 ? 64 : 32;
  static thread_local PatchParams p;
  memset(&p, 0, sizeof(p));
  tc::set_source(&p, X, xview, N, s->Cin, s->H, s->W);
  // TMA needs a 16-byte aligned box start: shift the patch origin left to a multiple of 4 floats
  const int xoff = ((-s->pad % 4) + 4) % 4;
  p.QH = s->OH; p.QW = s->OW; p.SS = s->stride; p.y0 = -s->pad; p.x0 = -s->pad - xoff;
  p.PH = (tc::PT_TH - 1) * s->stride + s->K; p.PW = ((tc::PT_TW - 1) * s->stride + s->K + xoff + 3) & ~3;
  p.tiles_y = cdiv(s->OH, tc::PT_TH); p.tiles_x = cdiv(s->OW, tc::PT_TW);
  p.nk = taps;
  for (int k = 0; k < taps; ++k) {
    const int ci = k / KK, r = k % KK, ky = r / s->K, kx = r % s->K;
    p.delta[k] = (uint16_t)(ci * p.PH * p.PW + ky * p.PW + kx + xoff);
  }
  p.dy = dY; p.Cout = s->Cout;
  const int total_tiles = N * p.tiles_y * p.tiles_x;
  int groups = num_sms() / mtiles;
  if (groups < 1) groups = 1;
  if (groups > total_tiles) groups = total_tiles;
  p.ncols = bn;
  double* bias_acc = reinterpret_cast<double*>(workspace);
  p.out = reinterpret_cast<float*>(bias_acc + ((s->Cout + 1) & ~1));
  const bool split3 = precision == 1;
  int rc;
  if (bn == 32) rc = split3 ? tc::launch_wgrad<32, true>(p, mtiles, groups, st) : tc::launch_wgrad<32, false>(p, mtiles, groups, st);
  else rc = split3 ? tc::launch_wgrad<64, true>(p, mtiles, groups, st) : tc::launch_wgrad<64, false>(p, mtiles, groups, st);
  if (rc) return rc;
  const int total = s->Cout * taps;
  tc::patch_wgrad_reduce_kernel<<<cdiv(total, 256), 256, 0, st>>>(p.out, groups, s->Cout, p.ncols, taps, dW);
  g_launches++;
  A2C_LAUNCH_CHECK();
  if (dbias) {
    A2C_CUDA_CHECK(cudaMemsetAsync(bias_acc, 0, sizeof(double) * s->Cout, st));
    const int hw = s->OH * s->OW;
    int bx = (int)cdivll((long long)N * hw, 256 * 8);
    if (bx > 256) bx = 256;
    if (bx < 1) bx = 1;
    tc::patch_channel_sum_kernel<<<dim3(bx, s->Cout), 256, 0, st>>>(dY, bias_acc, N, s->Cout, hw);
    tc::patch_channel_sum_finish_kernel<<<cdiv(s->Cout, 64), 64, 0, st>>>(bias_acc, dbias, s->Cout);
    g_launches += 2;
    A2C_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace a2c

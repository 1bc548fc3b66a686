// tcgen05 tile engine shared by the dense GEMM and the implicit-GEMM convolutions.
//
// One CTA computes a 128 x BN fp32 tile  D = sum_k A[m,k] * B[n,k]  on the 5th-gen
// tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM):
//   NA warps    producers of the A tile   (Problem::load_a / store_a)
//   NB warps    producers of the B tile   (Problem::load_b / store_b)
//   8 warps     accumulator drain + epilogue (Problem::epilogue)
//   1 warp      MMA issuer (one elected lane), which also owns the TMEM allocation
// Operand tiles live in shared memory in the canonical K-major SWIZZLE_128B UMMA
// layout: rows of 32 fp32 (128 B), 16-byte chunk index XOR (row & 7), 8-row
// groups 1024 B apart.  Producers are ordinary threads, so a Problem can gather
// (im2col, transposed, masked, u8 -> f32) while it fills the tile -- no operand
// is ever materialised in HBM in a tensor-core friendly layout.
//
// SPLIT3 (3xTF32): hi = tf32(x), lo = tf32(x - hi) (round to nearest, one F2FP each);  D += A_lo*B_hi + A_hi*B_lo + A_hi*B_hi, issued as
// TWO instructions per k-step:  D[:, 0:2BN] (+)= A_hi * [B_hi | B_lo]  (the lo tile follows the hi tile in
// shared memory, so one descriptor of width 2 BN covers both) and  D[:, 0:BN] += A_lo * B_hi;  the drain
// adds the two column halves.  Same tensor-pipe work as three MMAs, but the A tile is fetched from
// shared memory twice instead of three times.  (What bounds the kernel is the serial loop of the producer
// warps and the L2 -> SM latency their two fragments in flight have to cover: DESIGN.md lessons 18-19.)
// The TMEM accumulator adds with truncation (measured on B200: error grows
// linearly with the number of chained MMAs, 2e-4 at K = 25600), so K is consumed
// in chunks of KC k-blocks into two ping-pong TMEM accumulators and the drain
// warps add every finished chunk into fp32 registers with round-to-nearest
// while the tensor core works on the next chunk: 1e-5 of fp64 at any K.
#pragma once
#include "common.cuh"

namespace a2c {
namespace tc {

constexpr int TM = 128;            // UMMA M (cta_group::1)
constexpr int KB = 32;             // k-block: 32 fp32 = one 128-byte swizzle row
constexpr int NUM_EPI_WARPS = 8;
constexpr int TILE_GROUP_M = 8;     // m-tiles per band of the grouped tile order
constexpr int SCRATCH_BYTES = 4096;  // per-CTA scratch handed to Problem::init
constexpr int MAX_THREADS = 800;     // up to 16 producer warps + 8 drain + MMA

#ifdef A2C_TC_TRACE
static __device__ long long g_tc_trace[4][64][6];
#define TCT(role, it, ev)                                                                   \
  do {                                                                                      \
    if (blockIdx.x == 0 && lane == 0 && (it) >= 200 && (it) < 264) g_tc_trace[role][(it)-200][ev] = clock64(); \
  } while (0)
#else
#define TCT(role, it, ev) do { } while (0)
#endif

__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);        // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                             // leading byte offset (unused for SW128 K-major) = 1
  d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                             // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
  return d;
}

// kind::tf32, fp32 accumulate, both operands K-major (cute::UMMA::InstrDescriptor)
__host__ __device__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Variant for issue-bound loops: the instruction descriptor and the constant high descriptor
// word are immediates (ptxas keeps them in uniform registers instead of re-materialising them
// through R2UR for every instruction); only the low descriptor words are register operands.
template <uint32_t IDESC, uint32_t DESC_HI>
__device__ __forceinline__ void mma_tf32_imm(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t.reg .b32 hi, id;\n\t"
      "setp.ne.b32 p, %3, 0;\n\t"
      "mov.b32 hi, %4;\n\t"
      "mov.b32 id, %5;\n\t"
      "mov.b64 da, {%1, hi};\n\t"
      "mov.b64 db, {%2, hi};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, id, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(accumulate), "n"(DESC_HI), "n"(IDESC)
      : "memory");
}

__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}

// One lane of a converged warp, chosen by the hardware (elect.sync): unlike `lane == 0` the
// compiler knows exactly one thread is active in the guarded region, so a tcgen05.mma inside it
// takes its (warp-uniform) operands straight from uniform registers.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// Round to nearest (even) TF32: ONE instruction (F2FP.TF32.F32, full rate).  cvt.rna (ties away) has
// no hardware form on sm_100a -- ptxas emulates it with FSETP + IADD + SEL + LOP3 per element, which
// made the operand producers issue-bound (ncu: 29 % of the GEMM's instructions were that emulation).
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rn.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

__device__ __forceinline__ int swz(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

// Stores one 16-byte chunk (4 consecutive k of one row) at byte offset `off` = swz(row, chunk) of the tile(s).
template <bool SPLIT3>
__device__ __forceinline__ void store_chunk_at(uint8_t* hi_tile, uint8_t* lo_tile, int off, float4 v) {
  if constexpr (SPLIT3) {
    float4 h = make_float4(tf32_rna(v.x), tf32_rna(v.y), tf32_rna(v.z), tf32_rna(v.w));
    *reinterpret_cast<float4*>(hi_tile + off) = h;
    // the lo term is rounded to TF32 too: the tensor core would otherwise TRUNCATE its low 13 bits
    // (a one-sided error of up to 2^-10 of lo); rounded, the 3xTF32 product error is unbiased and ~3x smaller
    *reinterpret_cast<float4*>(lo_tile + off) =
        make_float4(tf32_rna(v.x - h.x), tf32_rna(v.y - h.y), tf32_rna(v.z - h.z), tf32_rna(v.w - h.w));
  } else {
    *reinterpret_cast<float4*>(hi_tile + off) = v;
  }
}
template <bool SPLIT3>
__device__ __forceinline__ void store_chunk(uint8_t* hi_tile, uint8_t* lo_tile, int row, int chunk,
                                            float4 v) {
  store_chunk_at<SPLIT3>(hi_tile, lo_tile, swz(row, chunk), v);
}

// Stores one element (row, k) into the swizzled tile(s).
template <bool SPLIT3>
__device__ __forceinline__ void store_elem(uint8_t* hi_tile, uint8_t* lo_tile, int row, int k, float v) {
  const int off = swz(row, k >> 2) + ((k & 3) << 2);
  if constexpr (SPLIT3) {
    const float h = tf32_rna(v);
    *reinterpret_cast<float*>(hi_tile + off) = h;
    *reinterpret_cast<float*>(lo_tile + off) = tf32_rna(v - h);
  } else {
    *reinterpret_cast<float*>(hi_tile + off) = v;
  }
}

// Problem concept:
//   static constexpr int NA, NB;      producer warps for A / B (NA + NB a multiple of 4)
//   static constexpr int KC;          k-blocks chained in one TMEM accumulator before it is drained
//   static constexpr int DEPTH;       k-blocks the producers' global loads run ahead of their stores (1 or 2)
//   struct Params { ...; int nkb; int kb_per_split; };        (nkb = total k-blocks)
//   struct AState, BState;            per-thread producer state of the current m / n tile (row decode etc.);
//                                     the stores may lag one tile behind the loads, so whatever store_a /
//                                     store_b read from the state must not depend on the tile
//   static __device__ void init(const Params&, uint8_t* scratch, int tid, int nthreads);
//   static __device__ void a_setup(const Params&, const uint8_t* scratch, int m0, int t, AState&);
//   static __device__ void b_setup(const Params&, const uint8_t* scratch, int n0, int t, BState&);
//   static __device__ void load_a / load_b(const Params&, const State&, scratch, int mn0, int kb, int t, Frag&);
//   static __device__ void store_a / store_b(const Params&, const State&, const Frag&, hi, lo, int t);
//   static __device__ void epilogue<HALF>(const Params&, int m, int n_begin, const float* acc, int split);
// t = thread index within the operand's producer warps.
template <int BN, bool SPLIT3, class P>
__global__ void __launch_bounds__(32 * (P::NA + P::NB + NUM_EPI_WARPS + 1), 1)
tc_kernel(const typename P::Params p, int stages, int mtiles, int ntiles, int group_m) {
  constexpr int NUM_PROD_WARPS = P::NA + P::NB;
  constexpr int KC = P::KC;
  constexpr int EPI0 = NUM_PROD_WARPS, MMA_WARP = EPI0 + NUM_EPI_WARPS, ALLOC_WARP = MMA_WARP;   // the MMA warp owns TMEM
  constexpr int THREADS = 32 * (MMA_WARP + 1);
  static_assert(NUM_PROD_WARPS % 4 == 0, "drain warps must start at a multiple of 4 (TMEM lane quarters)");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int A_BYTES = TM * 128, B_BYTES = BN * 128;
  constexpr int STAGE_BYTES = (SPLIT3 ? 2 : 1) * (A_BYTES + B_BYTES);
  constexpr int HALF = BN / 2;      // columns per drain warp
  // 1024-byte alignment of every tile is required by SWIZZLE_128B
  // (pointer arithmetic on the shared array keeps the shared address space: LDS/STS, not generic)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* scratch = smem + stages * STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(scratch + SCRATCH_BYTES);
  uint64_t* empty = full + stages;
  uint64_t* acc_full = empty + stages;       // [2]
  uint64_t* acc_empty = acc_full + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  // shuffle-broadcast marks the warp index as warp-uniform for the compiler (uniform role branches)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  // Persistent CTAs walk the output tiles in GROUPED order: tile t -> a band of TILE_GROUP_M m-tiles
  // swept along n.  The ~148 tiles in flight then form a compact 2-D block, so A and B tiles are
  // shared through L2 (with one n-tile per CTA and a loop over m, the B tiles of 148 CTAs did not
  // fit in L2 and the dgrad / wgrad GEMMs re-read 8-11x their operands from HBM).
  const int total_tiles = mtiles * ntiles;
  const int tile0 = blockIdx.x, tile_step = gridDim.x;
  auto tile_of = [&](int t, int& mt, int& nt) {
    const int per_group = group_m * ntiles;
    const int g = t / per_group, r = t - g * per_group;
    const int first_m = g * group_m;
    const int gsize = min(group_m, mtiles - first_m);
    nt = r / gsize;
    mt = first_m + (r - nt * gsize);
  };
  const int kb_begin = blockIdx.z * p.kb_per_split;
  const int kb_end = min(p.nkb, kb_begin + p.kb_per_split);
  const int nkb = kb_end - kb_begin;
  const int nchunks = (nkb + KC - 1) / KC;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full[s], NUM_PROD_WARPS);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], NUM_EPI_WARPS);
    }
    mbar_fence_init();
  }
  if (warp == ALLOC_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"((uint32_t)(4 * BN)));      // two ping-pong accumulators of [hi | lo] halves
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  P::init(p, scratch, threadIdx.x, THREADS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  auto stage_ptr = [&](int s) { return smem + s * STAGE_BYTES; };
  // stage layout: [A_hi][A_lo?][B_hi][B_lo?]
  if (warp < P::NA) {
    // ---- A producers.  P::DEPTH + 1 fragments per thread: while one is converted and stored, the
    // loads of the next P::DEPTH k-blocks are in flight (the L2 -> SM latency is ~1,500 clk under this
    // load, about one and a half k-block periods); the GEMM uses 16 producer warps.
    const int t = threadIdx.x;
    StageRing ring(stages);
    if (tile0 < total_tiles && nkb > 0) {
      typename P::AState ast;
      typename P::AFrag f0, f1, f2;
      int s_tile = tile0, s_i = 0;                       // store cursor (tile, k-block)
      int l_tile = tile0, l_i = 0, l_mt, l_nt;           // load cursor
      tile_of(l_tile, l_mt, l_nt);
      P::a_setup(p, scratch, l_mt * TM, t, ast);
      auto load_next = [&](typename P::AFrag& f) {
        if (l_tile >= total_tiles) return;
        P::load_a(p, ast, scratch, l_mt * TM, kb_begin + l_i, t, f);
        if (++l_i == nkb) {
          l_i = 0;
          l_tile += tile_step;
          if (l_tile < total_tiles) {
            int mt2;
            tile_of(l_tile, mt2, l_nt);
            if (mt2 != l_mt) { l_mt = mt2; P::a_setup(p, scratch, l_mt * TM, t, ast); }
          }
        }
      };
      // one k-block: start the loads of the k-block DEPTH ahead into `fl`, convert and store `fs`.
      // The loop below is unrolled over the DEPTH + 1 fragments so that they rotate by NAME (a
      // `f0 = f1; f1 = f2` rotation cost 32 register moves per k-block in an issue-bound loop).
      int kit = 0;
      // one k-block: convert and store fragment `f`, then refill the same registers with the k-block
      // DEPTH + 1 ahead.  The loads are issued AFTER the stores: issued before them, they shared a
      // hardware scoreboard with the fragment being consumed in one of the three unrolled copies and
      // every third k-block waited a full memory latency (clock trace: 1,400-1,900 clk instead of 300).
      auto step = [&](typename P::AFrag& f) -> bool {
        TCT(0, kit, 0);
        const int s = ring.s;
        mbar_wait(&empty[s], ring.par ^ 1);
        TCT(0, kit, 1);
        uint8_t* st = stage_ptr(s);
        P::store_a(p, ast, f, st, st + A_BYTES, t);
        TCT(0, kit, 2);
        fence_proxy_async();      // generic-proxy stores -> visible to the tensor core (async proxy)
        TCT(0, kit, 3);
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[s]);
        TCT(0, kit, 4);
        load_next(f);
        TCT(0, kit, 5);
        if (warp == 0) ++kit;
        if (++s_i == nkb) { s_i = 0; s_tile += tile_step; }
        ring.next();
        return s_tile < total_tiles;
      };
      load_next(f0);
      load_next(f1);
      if constexpr (P::DEPTH == 2) {
        load_next(f2);
        for (;;) {
          if (!step(f0)) break;
          if (!step(f1)) break;
          if (!step(f2)) break;
        }
      } else {
        for (;;) {
          if (!step(f0)) break;
          if (!step(f1)) break;
        }
      }
    }
  } else if (warp < NUM_PROD_WARPS) {
    // ---- B producers (same pipelining; the B tile depends on the k-block and the n tile)
    const int t = (int)threadIdx.x - 32 * P::NA;
    StageRing ring(stages);
    if (tile0 < total_tiles && nkb > 0) {
      typename P::BState bst;
      typename P::BFrag f0, f1, f2;
      int s_tile = tile0, s_i = 0;
      int l_tile = tile0, l_i = 0, l_mt, l_nt;
      tile_of(l_tile, l_mt, l_nt);
      P::b_setup(p, scratch, l_nt * BN, t, bst);
      auto load_next = [&](typename P::BFrag& f) {
        if (l_tile >= total_tiles) return;
        P::load_b(p, bst, scratch, l_nt * BN, kb_begin + l_i, t, f);
        if (++l_i == nkb) {
          l_i = 0;
          l_tile += tile_step;
          if (l_tile < total_tiles) {
            int nt2;
            tile_of(l_tile, l_mt, nt2);
            if (nt2 != l_nt) { l_nt = nt2; P::b_setup(p, scratch, l_nt * BN, t, bst); }
          }
        }
      };
      int kit = 0;
      auto step = [&](typename P::BFrag& f) -> bool {
        TCT(1, kit, 0);
        const int s = ring.s;
        mbar_wait(&empty[s], ring.par ^ 1);
        TCT(1, kit, 1);
        uint8_t* b = stage_ptr(s) + (SPLIT3 ? 2 : 1) * A_BYTES;
        P::store_b(p, bst, f, b, b + B_BYTES, t);
        TCT(1, kit, 2);
        fence_proxy_async();
        TCT(1, kit, 3);
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[s]);
        TCT(1, kit, 4);
        load_next(f);
        TCT(1, kit, 5);
        if (warp == P::NA) ++kit; else kit = -1;
        if (++s_i == nkb) { s_i = 0; s_tile += tile_step; }
        ring.next();
        return s_tile < total_tiles;
      };
      load_next(f0);
      load_next(f1);
      if constexpr (P::DEPTH == 2) {
        load_next(f2);
        for (;;) {
          if (!step(f0)) break;
          if (!step(f1)) break;
          if (!step(f2)) break;
        }
      } else {
        for (;;) {
          if (!step(f0)) break;
          if (!step(f1)) break;
        }
      }
    }
  } else if (warp == MMA_WARP) {
    constexpr uint32_t idesc = make_idesc(BN);
    constexpr uint32_t idesc2 = make_idesc(2 * BN);
    int it = 0, gchunk0 = 0;
    StageRing ring(stages);
    for (int tile = tile0; tile < total_tiles; tile += tile_step, gchunk0 += nchunks)
    for (int i = 0; i < nkb; ++i, ++it, ring.next()) {
      const int s = ring.s;
      const uint32_t parity = ring.par;
      const int chunk = gchunk0 + i / KC, buf = chunk & 1;
      TCT(2, it, 0);
      if (i % KC == 0) {                        // new chunk: its TMEM buffer must have been drained
        mbar_wait(&acc_empty[buf], ((chunk >> 1) & 1) ^ 1);
        tc_fence_after();
      }
      TCT(2, it, 1);
      mbar_wait(&full[s], parity);
      TCT(2, it, 2);
      tc_fence_after();
      // convergent warp, one elected lane, descriptor low words in uniform registers (see
      // mma_tf32_imm): about 3 issue instructions per MMA instead of ~25
      constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO, version, SWIZZLE_128B
      constexpr uint32_t LBO1 = 1u << 16;
      const uint32_t tmem_d = tmem_base + buf * 2 * BN;
      const uint32_t a_hi = ((smem_u32(stage_ptr(s)) & 0x3FFFFu) >> 4) | LBO1;
      const uint32_t a_lo = a_hi + (A_BYTES >> 4);
      const uint32_t b_hi = a_hi + (((SPLIT3 ? 2 : 1) * A_BYTES) >> 4);
      (void)B_BYTES;
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < KB / 8; ++kk) {
          const uint32_t koff = kk * 2;      // 8 tf32 = 32 bytes inside the 128-byte swizzle row
          const uint32_t first = ((i % KC) | kk) ? 1u : 0u;
          if constexpr (SPLIT3) {
            mma_tf32_imm<idesc2, DESC_HI>(tmem_d, a_hi + koff, b_hi + koff, first);
            mma_tf32_imm<idesc, DESC_HI>(tmem_d, a_lo + koff, b_hi + koff, 1u);
          } else {
            mma_tf32_imm<idesc, DESC_HI>(tmem_d, a_hi + koff, b_hi + koff, first);
          }
        }
        mma_commit(&empty[s]);                                   // frees the stage when these MMAs retire
        if (i % KC == KC - 1 || i == nkb - 1) mma_commit(&acc_full[buf]);   // chunk complete
      }
      __syncwarp();
      TCT(2, it, 3);
    }
  } else if (warp >= EPI0 && warp < EPI0 + NUM_EPI_WARPS) {
    // ---- drain + epilogue: thread <-> accumulator row (TMEM lane), half of the columns ----
    const int e = warp - EPI0, q = e & 3, hsel = e >> 2;
    int gchunk = 0;
    for (int tile = tile0; tile < total_tiles; tile += tile_step) {
    int mt, nt;
    tile_of(tile, mt, nt);
    const int m0 = mt * TM, n0 = nt * BN;
    float acc[HALF];
#pragma unroll
    for (int j = 0; j < HALF; ++j) acc[j] = 0.0f;
    for (int chunk = 0; chunk < nchunks; ++chunk, ++gchunk) {
      const int buf = gchunk & 1;
      // only one drain warp watches the mbarrier (see mbar_wait_group)
      mbar_wait_group(&acc_full[buf], (gchunk >> 1) & 1, e == 0, 1, NUM_EPI_WARPS, KC >= 8 ? 400 : 64);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 2 * BN + hsel * HALF;
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
        for (int j = 0; j < 16; ++j) acc[c0 + j] += __uint_as_float(r[j]);
        if constexpr (SPLIT3) {                       // the A_hi * B_lo half
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
              : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
                "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
                "=r"(r[14]), "=r"(r[15])
              : "r"(taddr + (uint32_t)(BN + c0)));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[c0 + j] += __uint_as_float(r[j]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
      if (e == 0) { TCT(3, 200 + gchunk, 3); }
    }
    P::template epilogue<HALF>(p, m0 + q * 32 + lane, n0 + hsel * HALF, acc, (int)blockIdx.z);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == ALLOC_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)(4 * BN)));
  }
}

template <int BN, bool SPLIT3>
constexpr int stage_bytes() { return (SPLIT3 ? 2 : 1) * (TM * 128 + BN * 128); }

// Launch helper: grid = (persistent CTAs, 1, splits).
template <int BN, bool SPLIT3, class P>
int tc_launch(const typename P::Params& p, int mtiles, int ntiles, int splits, cudaStream_t st) {
  constexpr int SB = stage_bytes<BN, SPLIT3>();
  int stages = (196 * 1024) / SB;
  if (stages > 4) stages = 4;
  if (stages < 2) return 1;
  const size_t smem = (size_t)stages * SB + 1024 /*align slack*/ + SCRATCH_BYTES + (2 * stages + 4) * 8 + 16;
  auto kern = tc_kernel<BN, SPLIT3, P>;
  {
    cudaError_t e = set_max_dynamic_smem(reinterpret_cast<const void*>(kern), (int)(227 * 1024));
    if (e != cudaSuccess) {
      set_error("tc_launch: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return (int)e;
    }
  }
  // one CTA per SM (the tile ring takes most of shared memory): persistent over the output tiles
  int ncta = num_sms() / (splits > 0 ? splits : 1);
  if (ncta < 1) ncta = 1;
  if (ncta > mtiles * ntiles) ncta = mtiles * ntiles;
  dim3 grid(ncta, 1, splits);
  int group_m = TILE_GROUP_M;
  if (const int k = debug_knob("A2C_TC_GROUP")) group_m = k;
  kern<<<grid, 32 * (P::NA + P::NB + NUM_EPI_WARPS + 1), smem, st>>>(p, stages, mtiles, ntiles, group_m);
  return 0;
}

}  // namespace tc
}  // namespace a2c

// ClipLoss on CTA pairs (tcgen05 cta_group::2), bf16 / fp16 features, dim <= 768.
//
// Replaces open_clip/loss.py:109-116 + 126-129 and their autograd with kernels that execute
// 8*n*N*D FLOP for the 6*n*N*D credited to forward + backward (one logit sweep each way, no
// feature-slab recompute as in clip_tc.cu):
//
//   pair_sweep_kernel<Mode, Tail> : persistent CTA pairs sweep tiles of S = X . Y^T (256 x 128
//       per pair).  The X row block lives in TMEM as the A operand (tcgen05.mma "TS" form, so
//       shared memory only streams Y; with dim > 512 the columns beyond 512 stay in shared
//       memory and use the "SS" form), S accumulates in a double-buffered TMEM tile, and two
//       ping-pong groups of eight epilogue warps consume it:
//         FwdRows / FwdBoth : flash-style online (max, sum) per row and, for FwdBoth, the column
//           sums of the SAME exponentials (rows re-weighted to a warp reference, 32 lanes added
//           with a halving butterfly) -- S is computed once for both cross-entropies;
//         Grad : the gradient weights
//           G_ij = 2^gs * ( exp(S_ij - lseA_i) + cb * exp(S_ij - lseB_j) - cd * [j == label_i] )
//           with one ex2 per logit (rank-one column factors), fp16, TMA-stored as 16 KB blocks
//           [128 rows x 64 cols].  The logits are never stored; G is a scratch of the backward.
//   pair_gemm_kernel       : dX += G . Y (A K-major) and dY += G^T . X (A MN-major, same G) as
//       ONE persistent stream-K launch: 256 x 512 tiles per pair fill TMEM, partial tiles are
//       reduced with red.global.add.v4.f32 -- for several ranks straight into the owner rank's
//       peer-mapped accumulator (fused reduce-scatter over NVLink).
//   grad_scale_cast_kernel : out_scale * acc -> gradient dtype.
#include "latte_common.cuh"
#include "tc_ptx.cuh"

#include <stdlib.h>

namespace latte {

using namespace ptx;

namespace {

constexpr int kThreads = 384;           // stream-K GEMM: 4 service warps + 8 epilogue warps
constexpr int kEpiWarp0 = 4;
constexpr int kGemmEpiWarps = 8;
constexpr int kSweepThreads = 640;      // sweep: 4 service warps + 2 groups of 8 epilogue warps
constexpr int kNumEpiWarps = 16;
constexpr int kPM = 128;              // rows per CTA (256 per pair)
constexpr int kTN = 128;              // S tile columns (pair MMA N)
constexpr int kBK = 64;               // feature columns per smem chunk (128 bytes)
constexpr int kYChunkBytes = 64 * kBK * 2;      // this CTA's half of a Y chunk: 64 rows
constexpr int kChunksPerStage = 4;               // one barrier round trip per 4 chunks (K = 256)
constexpr int kSweepStages = 4;
constexpr int kRingStageBytes = kChunksPerStage * kYChunkBytes;   // 32 KB
constexpr int kTmemChunks = 8;                   // X chunks held in TMEM (256 of the 512 columns)
constexpr int kXTailBytes = kPM * kBK * 2;       // 16 KB: one chunk of this CTA's X rows in smem
constexpr int kXTailOffset = 2 * kRingStageBytes;  // dim > 512: the ring keeps 64 KB, the tail gets 64 KB
constexpr int kStageBytes = 32 * 128;           // one epilogue warp's G staging: 32 rows x 128 B
constexpr int kMiscBytes = 1024;
constexpr int kVecBytes = 64 * 4;               // one epilogue warp's column factors of a tile
constexpr int kSweepSmem = kSweepStages * kRingStageBytes + kNumEpiWarps * (kStageBytes + kVecBytes) +
                           kMiscBytes;

constexpr int kGBlockElems = 128 * 64;          // one G block: 128 rows x 64 cols fp16

// G is stored as fp16 scaled by 2^gs; gs (>= 13) is chosen per call by the prep kernels from a
// bound on max |G| so that a nearly converged batch (all weights tiny) keeps fp16 precision.

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// =========================================================================== sweep (G)
struct SweepParams {
  const void* x; int64_t ldx;        // [n_loc, dim] rows of this rank (16-bit features)
  int64_t n_loc, n_all, dim;
  int64_t label_offset;
  const float* logit_scale;
  const float* lse_a2;               // base-2 LSE of the x side, indexed label_offset + i
  const float* lse_b2;               // base-2 LSE of the y side, indexed j (zero padded)
  const float* e_a;                  // 2^(lse_a2 - rho)   (x side, indexed like lse_a2)
  const float* einv_b;               // 2^(rho - lse_b2)   (y side, indexed j)
  const int* fast_flag;              // 1: the LSE range allows the one-ex2 epilogue
  const float* gscale_log2;          // log2 of the fp16 scale of G
  const float* nll_a;                // nullable: lse - label logit of the x / y side
  const float* nll_b;
  float cb, cd;
  float ds_cb, ds_cd;                // weights of the same terms inside d loss / d s
  float* ds_partial;                 // [gridDim.x]
  const float* logit_bias;           // SigLIP modes: nullable device scalar added to every logit
  float* aux_partial;                // [gridDim.x] SigLIP: loss partial (forward) / d bias partial (backward)
  // forward modes: per-(slot, half) online-softmax partials of the rows, raw label dots,
  // and (kModeFwdBoth) per-128-row-block column partial sums with their reference exponents
  float* part_max;                   // [4 * slots, n_loc]  (slot, epilogue group, tile half)
  float* part_sum;
  float* diag;                       // [n_loc]
  float* col_part;                   // [2 * row_blocks, ld_colpart]
  float* col_ref;                    // [2 * row_blocks, 4 * col_tiles]  (one per 32 columns)
  int64_t ld_colpart;
  int* zero2;                        // nullable: two ints cleared by the first CTA (forward modes)
  // Multi-rank: y is a gathered buffer whose shard of rank w (rows [w * rows_per_rank, ...)) is
  // written by rank w's push kernel; `landed[w] >= landed_gen` (a flag in this GPU's memory, set
  // by rank w after its stores) says the shard may be read.  NULL: y is complete at launch.
  const int* landed;
  int landed_gen;
  int64_t rows_per_rank;
  int kch;                           // ceil(dim / 64)
  int cps;                           // feature chunks per ring stage (4, or 2 with an X tail)
  int tail_chunks;                   // chunks of X beyond the 8 held in TMEM (dim > 512): smem
  int col_tiles;                     // tiles of 128 columns (even: columns padded to 256)
  int row_blocks;                    // blocks of 256 rows
  int ncb;                           // G block columns = 2 * col_tiles
  uint32_t idesc;
};

constexpr int kModeGrad = 0;      // backward: gradient weights G
constexpr int kModeFwdRows = 1;   // forward: row log-sum-exp partials
constexpr int kModeFwdBoth = 2;   // forward, world size 1: row partials + column partials of the same tile
constexpr int kModeSigFwd = 3;    // SigLIP forward: sum of softplus(-label * z) over the tile range
constexpr int kModeSigGrad = 4;   // SigLIP backward: G = sigmoid(z) - delta

// MUFU.RCP alone (__fdividef adds range fix-ups that a denominator in [1, 2] never needs)
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

struct TrueTag { static constexpr bool value = true; };
struct FalseTag { static constexpr bool value = false; };

// two fp32 lanes per instruction (FFMA2 on sm_100a)
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t mul_f32x2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// log1p(t0) + log1p(t1) for arguments in [0, 1]: log1p(t) = t * q(t) with q the degree-8 Chebyshev
// fit of log1p(t)/t (relative error 2e-7 in fp32 Horner form; MUFU lg2 near 1 only has an ABSOLUTE
// error of 2^-22, useless for the small terms that make up most of a sigmoid loss), both arguments
// in one chain of FFMA2.  Returns the pair (log1p(t0), log1p(t1)).
__device__ __forceinline__ uint64_t log1p_unit_pair(float t0, float t1) {
  const uint64_t t = pack_f32x2(t0, t1);
  uint64_t q = pack_f32x2(0.00525352f, 0.00525352f);
  q = fma_f32x2(q, t, pack_f32x2(-0.02958887f, -0.02958887f));
  q = fma_f32x2(q, t, pack_f32x2(0.07836246f, 0.07836246f));
  q = fma_f32x2(q, t, pack_f32x2(-0.13674858f, -0.13674858f));
  q = fma_f32x2(q, t, pack_f32x2(0.19111485f, 0.19111485f));
  q = fma_f32x2(q, t, pack_f32x2(-0.24844388f, -0.24844388f));
  q = fma_f32x2(q, t, pack_f32x2(0.33319275f, 0.33319275f));
  q = fma_f32x2(q, t, pack_f32x2(-0.49999502f, -0.49999502f));
  q = fma_f32x2(q, t, pack_f32x2(0.99999997f, 0.99999997f));
  return mul_f32x2(q, t);
}

template <int MODE, bool TAIL>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kSweepThreads, 1)
pair_sweep_kernel(const __grid_constant__ CUtensorMap tmy, const __grid_constant__ CUtensorMap tmg,
                  const __grid_constant__ CUtensorMap tmx, const SweepParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = smem_u32(smem);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();

  const uint32_t ring = smem_base;
  const uint32_t stage_base = ring + kSweepStages * kRingStageBytes;
  const uint32_t vec_base = stage_base + kNumEpiWarps * kStageBytes;
  const uint32_t misc = vec_base + kNumEpiWarps * kVecBytes;
  const uint32_t bar_full = misc;                         // [kSweepStages]
  const uint32_t bar_empty = bar_full + 8 * kSweepStages; // [kSweepStages]
  const uint32_t bar_tfull = bar_empty + 8 * kSweepStages;  // [2]
  const uint32_t bar_tempty = bar_tfull + 16;               // [2]   (leader's are used)
  const uint32_t bar_aready = bar_tempty + 16;              // [1]   (leader's is used)
  const uint32_t bar_xt = bar_aready + 8;                   // [1]   X tail landed (leader's is used)
  const uint32_t tmem_slot = bar_xt + 8;
  const uint32_t red_slot = tmem_slot + 8;
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - smem_base));
  float* red_ptr = reinterpret_cast<float*>(smem + (red_slot - smem_base));

  if (threadIdx.x == 0 && (smem_base & 1023u) != 0) __trap();
  if (blockIdx.x == 0 && threadIdx.x == 0 && p.zero2) { p.zero2[0] = 0; p.zero2[1] = 0; }

  // contiguous range of the flattened (row block, column tile) space for this pair
  const int64_t total = (int64_t)p.row_blocks * p.col_tiles;
  const int64_t ncl = gridDim.x >> 1;
  const int64_t cl = blockIdx.x >> 1;
  const int64_t u0 = cl * total / ncl;
  const int64_t u1 = (cl + 1) * total / ncl;
  // (row block, column tile) of the first tile; the loops below step them without dividing
  const int ntile = (int)(u1 - u0);
  const int rb0 = (int)(u0 / p.col_tiles);
  const int ct0 = (int)(u0 % p.col_tiles);

  // ring stages of cps chunks; with dim > 512 the X columns beyond the TMEM-resident 512 live in
  // the upper half of the ring area and their MMAs take A from shared memory ("SS" form)
  constexpr int kCps = TAIL ? 2 : kChunksPerStage;      // compile-time for the issue loops
  constexpr uint32_t ring_stage_bytes = (uint32_t)kCps * kYChunkBytes;
  const uint32_t xtail = ring + kXTailOffset;

  if (warp == 0 && elect_one()) {
    prefetch_tensormap(&tmy);
    prefetch_tensormap(&tmg);
    prefetch_tensormap(&tmx);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < kSweepStages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_tfull + 8 * b, 1);
      mbar_init(bar_tempty + 8 * b, 16);      // one group (8 warps) of each CTA drains a buffer
    }
    mbar_init(bar_aready, 16);                // group 0 (8 warps) of each CTA loads X
    mbar_init(bar_xt, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(tmem_slot, 512);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t tmem_a = tmem_base;            // X block, packed 16-bit pairs: columns [0, 256)
  const uint32_t tmem_s = tmem_base + 256;      // two S buffers of 128 columns

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    if (elect_one()) {
      const uint32_t lead_full = mapa_rank(bar_full, 0);
      const uint64_t keep = policy_evict_last();      // features are re-read by every pair
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t lead_xt = mapa_rank(bar_xt, 0);
      int ct = ct0, rb_i = rb0;
      // ranks whose shard is known to be here (bits of ranks that do not exist are set)
      uint32_t landed_mask = 0xffffffffu;
      if (p.landed) {
        const int nranks = (int)((p.n_all + p.rows_per_rank - 1) / p.rows_per_rank);
        landed_mask = nranks >= 32 ? 0u : ~((1u << nranks) - 1u);
      }
      for (int it = 0; it < ntile; ++it) {
        if (landed_mask != 0xffffffffu) {
          // owners of this tile's rows of y (a tile can straddle two shards)
          const int64_t r_first = (int64_t)ct * kTN;
          const int64_t r_last = min(r_first + kTN, p.n_all) - 1;
          if (r_first < p.n_all) {
            const int w0 = (int)(r_first / p.rows_per_rank), w1 = (int)(r_last / p.rows_per_rank);
            bool waited = false;
            for (int w = w0; w <= w1; ++w)
              if (!((landed_mask >> w) & 1u)) {
                flag_wait_ge(p.landed + w, p.landed_gen);
                landed_mask |= 1u << w;
                waited = true;
              }
            if (waited) fence_proxy_async_all();
          }
        }
        if (TAIL && (it == 0 || ct == 0)) {
          // new row block: its X tail replaces the previous one once every MMA of the
          // previous block is done (tfull of its last tile)
          if (it > 0) {
            const int last = it - 1;
            mbar_wait(bar_tfull + 8 * (last & 1), (last >> 1) & 1);
          }
          if (rank == 0) mbar_arrive_expect_tx(bar_xt, 2u * (uint32_t)p.tail_chunks * kXTailBytes);
          for (int j = 0; j < p.tail_chunks; ++j)
            tma_load_2d_pair_hint(xtail + j * kXTailBytes, &tmx, lead_xt, (kTmemChunks + j) * kBK,
                                  rb_i * 256 + (int)rank * kPM, keep);
        }
        for (int c0 = 0; c0 < p.kch; c0 += kCps) {
          const int nc = min(kCps, p.kch - c0);
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          if (rank == 0) mbar_arrive_expect_tx(bar_full + 8 * stage, 2 * nc * kYChunkBytes);
          for (int c = 0; c < nc; ++c)
            tma_load_2d_pair_hint(ring + stage * ring_stage_bytes + c * kYChunkBytes, &tmy,
                                  lead_full + 8 * stage, (c0 + c) * kBK, ct * kTN + (int)rank * 64,
                                  keep);
          if (++stage == kSweepStages) { stage = 0; phase ^= 1; }
        }
        if (++ct == p.col_tiles) { ct = 0; ++rb_i; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA)
    if (rank == 0 && elect_one()) {
      int stage = 0;
      uint32_t phase = 0, a_phase = 0;
      int ct = ct0;
      for (int it = 0; it < ntile; ++it) {
        if (it == 0 || ct == 0) {          // first tile of a row block: wait for its X block
          mbar_wait(bar_aready, a_phase);
          if (TAIL) mbar_wait(bar_xt, a_phase);
          a_phase ^= 1;
        }
        ct = (ct + 1 == p.col_tiles) ? 0 : ct + 1;
        const int buf = it & 1;
        mbar_wait(bar_tempty + 8 * buf, ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_s + buf * kTN;
        for (int c0 = 0; c0 < p.kch; c0 += kCps) {
          const int nc = min(kCps, p.kch - c0);
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          // descriptor of the stage base; the start-address field counts 16-byte units
          const uint64_t db0 = make_smem_desc_sw128(ring + stage * ring_stage_bytes, 16, 1024);
#pragma unroll
          for (int c = 0; c < kCps; ++c) {
            if (c < nc) {
              const int gc = c0 + c;
              if (!TAIL || gc < kTmemChunks) {
#pragma unroll
                for (int k = 0; k < kBK / 16; ++k) {
                  const uint64_t db = db0 + (uint64_t)((c * kYChunkBytes + k * 32) >> 4);
                  mma2_ts(tmem_d, tmem_a + gc * 32 + k * 8, db, p.idesc, (gc | k) != 0);
                }
              } else {
                const uint64_t da0 =
                    make_smem_desc_sw128(xtail + (gc - kTmemChunks) * kXTailBytes, 16, 1024);
#pragma unroll
                for (int k = 0; k < kBK / 16; ++k) {
                  const uint64_t db = db0 + (uint64_t)((c * kYChunkBytes + k * 32) >> 4);
                  mma2_ss(tmem_d, da0 + (uint64_t)(k * 2), db, p.idesc, 1u);
                }
              }
            }
          }
          tc_commit_pair(bar_empty + 8 * stage, 3);
          if (++stage == kSweepStages) { stage = 0; phase ^= 1; }
        }
        tc_commit_pair(bar_tfull + 8 * buf, 3);
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ------------------------------------------------------------ epilogue (both CTAs)
    // Two groups of eight warps take alternate tiles (group g owns TMEM buffer g), so each
    // group has two tile times for its tile; inside a group, warp%4 selects the 32-lane TMEM
    // quarter and the next bit the 64-column half of the tile.
    const int e = warp - kEpiWarp0;
    const int group = e >> 3;
    const int half = (e >> 2) & 1;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const float c2 = __ldg(p.logit_scale) * kLog2e;
    const uint32_t lead_tempty = mapa_rank(bar_tempty, 0) + 8 * group;
    const uint32_t lead_aready = mapa_rank(bar_aready, 0);
    const uint32_t my_tfull = bar_tfull + 8 * group;
    const uint32_t tmem_tile = tmem_s + lane_base + group * kTN + half * 64;

    // tile `group` of the range, then every second tile
    int rb_i = rb0, ct = ct0;
    auto step_tile = [&]() {
      if (++ct == p.col_tiles) { ct = 0; ++rb_i; }
    };
    if (group == 1) step_tile();
    int cur_rb = -1;
    int64_t grow = 0, label = 0, warp_label0 = 0;
    bool row_ok = false;

    // New row block: refresh the row constants; group 0 also writes this CTA's 128 rows of X
    // into TMEM (the A operand of the TS MMA) once every MMA of the previous block is done.
    auto enter_row_block = [&](int it) {
      cur_rb = rb_i;
      grow = (int64_t)rb_i * 256 + (int64_t)rank * kPM + row;
      row_ok = grow < p.n_loc;
      label = p.label_offset + grow;
      warp_label0 = p.label_offset + (int64_t)rb_i * 256 + (int64_t)rank * kPM + q * 32;
      if (group != 0) return;
      const int first = it - ct;                 // first tile of this row block (ct is 0 or 1 here)
      if (first > 0) {
        const int last = first - 1;              // last tile of the previous row block
        mbar_wait(bar_tfull + 8 * (last & 1), (last >> 1) & 1);
        tc_fence_after();
      }
      const uint16_t* xrow = reinterpret_cast<const uint16_t*>(p.x) + (row_ok ? grow : 0) * p.ldx;
      const int groups = min(p.kch, kTmemChunks) * 2;   // groups of 32 features = 16 packed columns
      for (int g = half; g < groups; g += 2) {
        uint32_t w[16];
#pragma unroll
        for (int v4 = 0; v4 < 4; ++v4) {
          const int e0 = g * 32 + v4 * 8;
          uint4 val = make_uint4(0u, 0u, 0u, 0u);
          if (row_ok && e0 < p.dim) val = __ldg(reinterpret_cast<const uint4*>(xrow + e0));
          w[v4 * 4 + 0] = val.x; w[v4 * 4 + 1] = val.y;
          w[v4 * 4 + 2] = val.z; w[v4 * 4 + 3] = val.w;
        }
        tmem_st_32x16(tmem_a + lane_base + g * 16, w);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(lead_aready);
    };

    // A row block that starts at the LAST tile of this range when that tile has an odd local index
    // belongs to group 1 alone: group 0 has no later tile at whose top it would load the block's X,
    // and the MMA issuer would wait for it forever.  Called by every mode after its tile loop
    // (group 0's (rb_i, ct) then name tile `ntile`, one past the end).
    auto trailing_row_block = [&]() {
      if (group == 0 && ntile >= 2 && (ntile & 1) == 0 && ct == 1 && rb_i != cur_rb)
        enter_row_block(ntile);
    };

    if constexpr (MODE == kModeGrad) {
      // ======================================================== gradient weights
      const float s = __ldg(p.logit_scale);
      (void)s;
      const float gs = __ldg(p.gscale_log2);
      const float gscale = exp2f(gs);
      const float cd_scaled = p.cd * gscale;
      const float ds_cd_scaled = p.ds_cd * gscale;
      const float c2h = 0.5f * c2;
      const bool fast = __ldg(p.fast_flag) != 0;
      const bool ds_both = p.ds_cb != 0.f;
      const uint32_t my_stage = stage_base + e * kStageBytes;
      const uint64_t stream_pol = policy_evict_first();   // G is written once, read once
      // column factors B_j of this warp's half tile, staged in shared memory one tile ahead
      float* my_vec = reinterpret_cast<float*>(smem + (vec_base - smem_base) + e * kVecBytes);
      if (group < ntile) {
        const int64_t c0 = (int64_t)ct * kTN + half * 64;
        my_vec[lane] = __ldg(p.einv_b + c0 + lane);
        my_vec[32 + lane] = __ldg(p.einv_b + c0 + 32 + lane);
        __syncwarp();
      }
      float ds_acc = 0.f;
      float a2 = 0.f, cb = 0.f, ds_cb = 0.f, off_h = 0.f, cbA = 0.f;
      for (int it = group, k = 0; it < ntile; it += 2, ++k) {
        if (rb_i != cur_rb) {
          enter_row_block(it);
          // rows past the end: huge LSE and no cross term -> G = 0
          a2 = row_ok ? __ldg(p.lse_a2 + label) - gs : 1.0e30f;
          cb = row_ok ? p.cb : 0.f;
          ds_cb = row_ok ? p.ds_cb : 0.f;
          off_h = row_ok ? 0.5f * (gs - __ldg(p.lse_a2 + label)) : -1.0e30f;
          cbA = (row_ok && fast) ? p.cb * __ldg(p.e_a + label) : 0.f;
        }
        const int ct_cur = ct;
        const int rb_cur = rb_i;
        step_tile();
        step_tile();                              // (rb_i, ct) now name this group's next tile
        float bn0 = 0.f, bn1 = 0.f;
        if (it + 2 < ntile) {
          const int64_t cn = (int64_t)ct * kTN + half * 64;
          bn0 = __ldg(p.einv_b + cn + lane);
          bn1 = __ldg(p.einv_b + cn + 32 + lane);
        }
        const int64_t col0 = (int64_t)ct_cur * kTN + half * 64;
        const bool ragged = col0 + 64 > p.n_all;
        const bool diag_tile = warp_label0 < col0 + 64 && warp_label0 + 32 > col0;
        const bool plain = fast && !ragged && !diag_tile;
        const uint32_t row_addr = my_stage + lane * 128;

        mbar_wait(my_tfull, k & 1);
        tc_fence_after();
        if (lane == 0) bulk_wait_group_read<0>();   // previous store has left the staging buffer
        __syncwarp();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_tile + h * 32, r);
          tmem_ld_wait();
          if (h == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(lead_tempty);
          }
          uint32_t packed[16];
          if (plain) {
            // One ex2 per logit: ea = hh*hh with hh = 2^((x - a_i + 13)/2); the y-side softmax
            // term is ea * 2^(a_i - b_j) = ea * A_i * B_j (rank one; the prep kernel checked
            // that the LSE range keeps every factor inside fp32 range).
            const float4* pB = reinterpret_cast<const float4*>(my_vec) + h * 8;
#pragma unroll
            for (int i4 = 0; i4 < 8; ++i4) {
              const float4 b4 = pB[i4];
              const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
              float g[4];
#pragma unroll
              for (int x = 0; x < 4; ++x) {
                const float v = __uint_as_float(r[i4 * 4 + x]);
                const float hh = fast_exp2(fmaf(v, c2h, off_h));
                const float ea = hh * hh;
                g[x] = ea * fmaf(cbA, bb[x], 1.0f);
                ds_acc = fmaf(ds_both ? g[x] : ea, v, ds_acc);
              }
              packed[i4 * 2 + 0] = pack2(g[0], g[1]);
              packed[i4 * 2 + 1] = pack2(g[2], g[3]);
            }
          } else {
            int want = -1;
            if (row_ok && label >= col0 + h * 32 && label < col0 + h * 32 + 32)
              want = (int)(label - col0) - h * 32;
            // label entry: P - 1 = expm1(-nll), no cancellation when the label dominates
            const bool have_nll = p.nll_a != nullptr;
            float g_lab = 0.f, dw_lab = 0.f;
            if (want >= 0 && have_nll) {
              const float pa1 = expm1f(-__ldg(p.nll_a + label)) * gscale;
              const float pb1 = expm1f(-__ldg(p.nll_b + label)) * gscale;
              g_lab = fmaf(cb, pb1, pa1);
              dw_lab = fmaf(ds_cb, pb1, pa1);
            }
            const float4* pb = reinterpret_cast<const float4*>(p.lse_b2 + col0) + h * 8;
#pragma unroll
            for (int i4 = 0; i4 < 8; ++i4) {
              const float4 b4 = __ldg(pb + i4);
              const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
              float g[4];
#pragma unroll
              for (int x = 0; x < 4; ++x) {
                const int i = i4 * 4 + x;
                const float v = __uint_as_float(r[i]);
                float ea = fast_exp2(fmaf(v, c2, -a2));
                float eb = fast_exp2(fmaf(v, c2, gs - bb[x]));
                if (ragged && col0 + h * 32 + i >= p.n_all) { ea = 0.f; eb = 0.f; }
                float dw = fmaf(ds_cb, eb, ea);
                g[x] = fmaf(cb, eb, ea);
                if (i == want) {
                  if (have_nll) {
                    g[x] = g_lab;
                    dw = dw_lab;
                  } else {
                    g[x] -= cd_scaled;
                    dw -= ds_cd_scaled;
                  }
                }
                ds_acc = fmaf(dw, v, ds_acc);
              }
              packed[i4 * 2 + 0] = pack2(g[0], g[1]);
              packed[i4 * 2 + 1] = pack2(g[2], g[3]);
            }
          }
          // 32 columns = four 16-byte chunks of this row of the 128B-swizzled staging piece
#pragma unroll
          for (int j = 0; j < 4; ++j)
            st_shared_v4(row_addr + ((uint32_t)((h * 4 + j) ^ (lane & 7)) << 4), packed[4 * j],
                         packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
        }
        my_vec[lane] = bn0;            // column factors of this group's next tile
        my_vec[32 + lane] = bn1;
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          const int64_t block = ((int64_t)rb_cur * 2 + rank) * (int64_t)p.ncb + (ct_cur * 2 + half);
          tma_store_2d_hint(&tmg, my_stage, 0, (int32_t)(block * 128 + q * 32), stream_pol);
          bulk_commit_group();
        }
      }
      if (lane == 0) bulk_wait_group<0>();
      trailing_row_block();

      // ---- d loss / d s partial of this CTA
      float v = ds_acc * exp2f(-gs);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) red_ptr[e] = v;
      named_bar_sync(1, kNumEpiWarps * 32);
      if (e == 0 && lane == 0) {
        float tot = 0.f;
        for (int w = 0; w < kNumEpiWarps; ++w) tot += red_ptr[w];
        p.ds_partial[blockIdx.x] = tot;
      }
    } else if constexpr (MODE == kModeSigGrad) {
      // ======================================================== SigLIP gradient weights
      // z = s * <x_i, y_j> + b;  d softplus(-label z) / dz = sigmoid(z) - [j == label_i]
      // (loss.py:509-519).  Stored as fp16 scaled by 2^13 like the softmax weights.
      const float b2 = p.logit_bias ? __ldg(p.logit_bias) * kLog2e : 0.f;
      constexpr float kSigScale = 8192.0f;
      const uint32_t my_stage = stage_base + e * kStageBytes;
      const uint64_t stream_pol = policy_evict_first();
      float ds_acc = 0.f, db_acc = 0.f;
      for (int it = group, k = 0; it < ntile; it += 2, ++k) {
        if (rb_i != cur_rb) enter_row_block(it);
        const int ct_cur = ct;
        const int rb_cur = rb_i;
        step_tile();
        step_tile();
        const int64_t col0 = (int64_t)ct_cur * kTN + half * 64;
        const uint32_t row_addr = my_stage + lane * 128;
        mbar_wait(my_tfull, k & 1);
        tc_fence_after();
        if (lane == 0) bulk_wait_group_read<0>();
        __syncwarp();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_tile + h * 32, r);
          tmem_ld_wait();
          if (h == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(lead_tempty);
          }
          const int64_t colh = col0 + h * 32;
          int want = -1;
          if (row_ok && label >= colh && label < colh + 32) want = (int)(label - colh);
          const int valid = !row_ok ? 0 : (colh + 32 <= p.n_all ? 32 : (colh >= p.n_all ? 0 : (int)(p.n_all - colh)));
          // warp-uniform: no label entry, no ragged column, no row past the end in this 32x32 piece
          const bool plain = !(warp_label0 < colh + 32 && warp_label0 + 32 > colh) && colh + 32 <= p.n_all &&
                             __all_sync(0xffffffffu, row_ok);
          uint32_t packed[16];
          float dsh = 0.f, dbh = 0.f;
          const uint64_t c2p = pack_f32x2(c2, c2), b2p = pack_f32x2(b2, b2), onep = pack_f32x2(1.f, 1.f);
          const uint64_t scp = pack_f32x2(kSigScale, kSigScale);
          uint64_t dsp = pack_f32x2(0.f, 0.f), dbp = pack_f32x2(0.f, 0.f);
          auto piece = [&](auto plain_tag) {
            constexpr bool kPlain = decltype(plain_tag)::value;
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              // packed fp32 pairs (FFMA2 / FADD2 / FMUL2) wherever both lanes do the same thing
              const uint64_t vp = pack_f32x2(__uint_as_float(r[i]), __uint_as_float(r[i + 1]));
              float z[2], t[2], den[2], tr[2], sg[2];
              unpack_f32x2(fma_f32x2(vp, c2p, b2p), z[0], z[1]);
              t[0] = fast_exp2(-fabsf(z[0]));
              t[1] = fast_exp2(-fabsf(z[1]));
              const uint64_t tp = pack_f32x2(t[0], t[1]);
              unpack_f32x2(add_f32x2(tp, onep), den[0], den[1]);
              const float rr[2] = {rcp_approx(den[0]), rcp_approx(den[1])};   // den in [1, 2]
              unpack_f32x2(mul_f32x2(tp, pack_f32x2(rr[0], rr[1])), tr[0], tr[1]);
#pragma unroll
              for (int x = 0; x < 2; ++x) {
                const bool pos = z[x] >= 0.f;
                sg[x] = pos ? rr[x] : tr[x];                        // sigmoid(z)
                if constexpr (!kPlain) {
                  if (i + x == want) sg[x] = pos ? -tr[x] : -rr[x]; // sigmoid(z) - 1 without cancellation
                  if (i + x >= valid) sg[x] = 0.f;
                }
              }
              const uint64_t sgp = pack_f32x2(sg[0], sg[1]);
              dsp = fma_f32x2(sgp, vp, dsp);
              dbp = add_f32x2(dbp, sgp);
              float g0, g1;
              unpack_f32x2(mul_f32x2(sgp, scp), g0, g1);
              packed[i >> 1] = pack2(g0, g1);
            }
          };
          if (plain) piece(TrueTag{}); else piece(FalseTag{});
          {
            float a0, a1, b0, b1;
            unpack_f32x2(dsp, a0, a1);
            unpack_f32x2(dbp, b0, b1);
            dsh = a0 + a1;
            dbh = b0 + b1;
          }
          ds_acc += dsh;
          db_acc += dbh;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            st_shared_v4(row_addr + ((uint32_t)((h * 4 + j) ^ (lane & 7)) << 4), packed[4 * j],
                         packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          const int64_t block = ((int64_t)rb_cur * 2 + rank) * (int64_t)p.ncb + (ct_cur * 2 + half);
          tma_store_2d_hint(&tmg, my_stage, 0, (int32_t)(block * 128 + q * 32), stream_pol);
          bulk_commit_group();
        }
      }
      if (lane == 0) bulk_wait_group<0>();
      trailing_row_block();
      float v0 = ds_acc, v1 = db_acc;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        v0 += __shfl_xor_sync(0xffffffffu, v0, o);
        v1 += __shfl_xor_sync(0xffffffffu, v1, o);
      }
      if (lane == 0) { red_ptr[e] = v0; red_ptr[kNumEpiWarps + e] = v1; }
      named_bar_sync(1, kNumEpiWarps * 32);
      if (e == 0 && lane == 0) {
        double t0 = 0.0, t1 = 0.0;
        for (int w = 0; w < kNumEpiWarps; ++w) { t0 += red_ptr[w]; t1 += red_ptr[kNumEpiWarps + w]; }
        p.ds_partial[blockIdx.x] = (float)t0;
        p.aux_partial[blockIdx.x] = (float)t1;
      }
    } else if constexpr (MODE == kModeSigFwd) {
      // ======================================================== SigLIP loss terms
      // -logsigmoid(label * z) = softplus(z) for the negatives and softplus(z) - z for the label
      // entry; softplus(z) = max(z, 0) + log1p(exp(-|z|))  (loss.py:515-519)
      const float b2 = p.logit_bias ? __ldg(p.logit_bias) * kLog2e : 0.f;
      float sum = 0.f, comp = 0.f;                 // compensated running sum of this thread
      for (int it = group, k = 0; it < ntile; it += 2, ++k) {
        if (rb_i != cur_rb) enter_row_block(it);
        const int ct_cur = ct;
        step_tile();
        step_tile();
        const int64_t col0 = (int64_t)ct_cur * kTN + half * 64;
        mbar_wait(my_tfull, k & 1);
        tc_fence_after();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_tile + h * 32, r);
          tmem_ld_wait();
          if (h == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(lead_tempty);
          }
          const int64_t colh = col0 + h * 32;
          int want = -1;
          if (row_ok && label >= colh && label < colh + 32) want = (int)(label - colh);
          const int valid = !row_ok ? 0 : (colh + 32 <= p.n_all ? 32 : (colh >= p.n_all ? 0 : (int)(p.n_all - colh)));
          float big = 0.f;                                   // log2 units
          uint64_t smallp = pack_f32x2(0.f, 0.f), bigp = pack_f32x2(0.f, 0.f);   // ln units / log2 units
          // warp-uniform: no label entry, no ragged column, no row past the end in this 32x32 piece
          const bool plain = !(warp_label0 < colh + 32 && warp_label0 + 32 > colh) && colh + 32 <= p.n_all &&
                             __all_sync(0xffffffffu, row_ok);
          const uint64_t c2p = pack_f32x2(c2, c2), b2p = pack_f32x2(b2, b2);
          auto piece = [&](auto plain_tag) {
            constexpr bool kPlain = decltype(plain_tag)::value;
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float z0, z1;
              unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), c2p, b2p), z0, z1);
              if constexpr (!kPlain) {
                if (i >= valid) z0 = -INFINITY;
                if (i + 1 >= valid) z1 = -INFINITY;
              }
              smallp = add_f32x2(smallp, log1p_unit_pair(fast_exp2(-fabsf(z0)), fast_exp2(-fabsf(z1))));
              bigp = add_f32x2(bigp, pack_f32x2(fmaxf(z0, 0.f), fmaxf(z1, 0.f)));
              if constexpr (!kPlain) {
                if (i == want) big -= z0;
                if (i + 1 == want) big -= z1;
              }
            }
          };
          if (plain) piece(TrueTag{}); else piece(FalseTag{});
          float s0, s1, b0, b1;
          unpack_f32x2(smallp, s0, s1);
          unpack_f32x2(bigp, b0, b1);
          const float term = fmaf(big + (b0 + b1), kLn2, s0 + s1);
          const float y = term - comp;               // Kahan
          const float tsum = sum + y;
          comp = (tsum - sum) - y;
          sum = tsum;
        }
      }
      trailing_row_block();
      float v0 = sum;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v0 += __shfl_xor_sync(0xffffffffu, v0, o);
      if (lane == 0) red_ptr[e] = v0;
      named_bar_sync(1, kNumEpiWarps * 32);
      if (e == 0 && lane == 0) {
        double t0 = 0.0;
        for (int w = 0; w < kNumEpiWarps; ++w) t0 += red_ptr[w];
        p.aux_partial[blockIdx.x] = (float)t0;
      }
    } else {
      // ======================================================== forward
      constexpr bool kBothSides = MODE == kModeFwdBoth;
      // exponent FMA, sums and column weights on packed fp32 pairs (FFMA2 / FADD2 / FMUL2: same
      // roundings, half the issue slots; the epilogue is issue-bound).  false = scalar reference form
      constexpr bool kPackedMath = true;
      // (Column sums through shared memory -- 8 STS.128 + 8 LDS.128 + 2 shuffle steps per half tile
      // instead of the 31-shuffle butterfly below -- were measured 11 % SLOWER on the same box:
      // the MMA's B-operand reads and the TMA ring already take most of the shared-memory bandwidth.)
      // column-partial exchange between the four lane-quarter warps of one half tile
      float* colbuf = reinterpret_cast<float*>(smem + (stage_base - smem_base));   // [2][2][2][4][64]
      float* refbuf = colbuf + 2 * 2 * 2 * 4 * 64;                                  // [2][2][2][4][2]
      float m = -INFINITY, l = 0.f;
      auto flush_rows = [&]() {
        if (cur_rb >= 0 && row_ok) {
          const int64_t slot = cl - cluster_of_tile((int64_t)cur_rb * p.col_tiles, total, ncl);
          const int64_t idx = ((slot * 2 + group) * 2 + half) * p.n_loc + grow;
          p.part_max[idx] = m;
          p.part_sum[idx] = l;
        }
      };
      for (int it = group, k = 0; it < ntile; it += 2, ++k) {
        if (rb_i != cur_rb) {
          flush_rows();
          enter_row_block(it);
          m = -INFINITY;
          l = 0.f;
        }
        const int ct_cur = ct;
        const int rb_cur = rb_i;
        step_tile();
        step_tile();
        const int64_t col0 = (int64_t)ct_cur * kTN + half * 64;
        const bool diag_tile = warp_label0 < col0 + 64 && warp_label0 + 32 > col0;
        const int par = k & 1;
        float* cb_w = colbuf + (((par * 2 + group) * 2 + half) * 4 + q) * 64;
        float* rf_w = refbuf + (((par * 2 + group) * 2 + half) * 4 + q) * 2;

        mbar_wait(my_tfull, k & 1);
        tc_fence_after();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_tile + h * 32, r);
          tmem_ld_wait();
          if (h == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(lead_tempty);
          }
          const int64_t colh = col0 + h * 32;
          if (colh + 32 > p.n_all) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (colh + i >= p.n_all) r[i] = 0xff800000u;  // -inf
          }
          if (diag_tile && row_ok && label >= colh && label < colh + 32) {
            const int want = (int)(label - colh);
            float dv = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i == want) dv = __uint_as_float(r[i]);
            p.diag[grow] = dv;
          }
          float tmax0 = -INFINITY, tmax1 = -INFINITY;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            tmax0 = fmax3(tmax0, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
            tmax1 = fmax3(tmax1, __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
          }
          const float m_new = fmaxf(m, fmaxf(tmax0, tmax1) * c2);
          if (m_new > -INFINITY) {
            if constexpr (kPackedMath) {
              // same roundings as the scalar form, two lanes per FFMA2 / FADD2
              const uint64_t c2p = pack_f32x2(c2, c2), nmp = pack_f32x2(-m_new, -m_new);
              uint64_t accp = pack_f32x2(0.f, 0.f);
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                float x0, x1;
                unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), c2p, nmp),
                             x0, x1);
                const float e0 = fast_exp2(x0), e1 = fast_exp2(x1);
                accp = add_f32x2(accp, pack_f32x2(e0, e1));
                r[i] = __float_as_uint(e0);
                r[i + 1] = __float_as_uint(e1);
              }
              float acc0, acc1;
              unpack_f32x2(accp, acc0, acc1);
              l = l * fast_exp2(m - m_new) + (acc0 + acc1);
            } else {
              float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                const float e0 = fast_exp2(fmaf(__uint_as_float(r[i]), c2, -m_new));
                const float e1 = fast_exp2(fmaf(__uint_as_float(r[i + 1]), c2, -m_new));
                acc0 += e0;
                acc1 += e1;
                if constexpr (kBothSides) {
                  r[i] = __float_as_uint(e0);
                  r[i + 1] = __float_as_uint(e1);
                }
              }
              l = l * fast_exp2(m - m_new) + (acc0 + acc1);
            }
            m = m_new;
          } else if constexpr (kBothSides) {
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = 0u;
          }
          if constexpr (kBothSides) {
            // Column sums of the same exponentials: weight row i by 2^(m_i - M_w) (M_w = the
            // largest running max among the warp's rows), add over the 32 lanes with a halving
            // butterfly (lane c ends with column c).
            float mw = row_ok ? m : -INFINITY;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mw = fmaxf(mw, __shfl_xor_sync(0xffffffffu, mw, o));
            const float w = (row_ok && m > -INFINITY) ? fast_exp2(m - mw) : 0.f;
            float v[32];
            if constexpr (kPackedMath) {
              const uint64_t wp = pack_f32x2(w, w);
#pragma unroll
              for (int i = 0; i < 32; i += 2)
                unpack_f32x2(mul_f32x2(pack_f32x2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), wp),
                             v[i], v[i + 1]);
#pragma unroll
              for (int o = 16; o > 1; o >>= 1) {
                const bool up = (lane & o) != 0;
#pragma unroll
                for (int kk = 0; kk < o; kk += 2) {
                  const float s0 = up ? v[kk] : v[kk + o], s1 = up ? v[kk + 1] : v[kk + 1 + o];
                  const float k0 = up ? v[kk + o] : v[kk], k1 = up ? v[kk + 1 + o] : v[kk + 1];
                  const float g0 = __shfl_xor_sync(0xffffffffu, s0, o);
                  const float g1 = __shfl_xor_sync(0xffffffffu, s1, o);
                  unpack_f32x2(add_f32x2(pack_f32x2(k0, k1), pack_f32x2(g0, g1)), v[kk], v[kk + 1]);
                }
              }
              {
                const bool up = (lane & 1) != 0;
                const float send = up ? v[0] : v[1];
                const float keep = up ? v[1] : v[0];
                v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) * w;
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) {
                const bool up = (lane & o) != 0;
#pragma unroll
                for (int kk = 0; kk < o; ++kk) {
                  const float send = up ? v[kk] : v[kk + o];
                  const float keep = up ? v[kk + o] : v[kk];
                  v[kk] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                }
              }
            }
            cb_w[h * 32 + lane] = v[0];
            if (lane == 0) rf_w[h] = mw;
          }
        }
        if constexpr (kBothSides) {
          // merge the four lane-quarter warps: warp q finishes columns [16q, 16q + 16)
          named_bar_sync(1 + group * 2 + half, 128);
          if (lane < 16) {
            const int piece = q >> 1;
            const float* rf = refbuf + ((par * 2 + group) * 2 + half) * 4 * 2 + piece;
            const float* cbh = colbuf + ((par * 2 + group) * 2 + half) * 4 * 64;
            const float M = fmaxf(fmaxf(rf[0], rf[2]), fmaxf(rf[4], rf[6]));
            float c = 0.f;
#pragma unroll
            for (int qq = 0; qq < 4; ++qq)
              if (rf[qq * 2] > -INFINITY)
                c = fmaf(cbh[qq * 64 + q * 16 + lane], fast_exp2(rf[qq * 2] - M), c);
            const int64_t rblk = (int64_t)rb_cur * 2 + rank;
            p.col_part[rblk * p.ld_colpart + col0 + q * 16 + lane] = c;
            if ((q & 1) == 0 && lane == 0)
              p.col_ref[rblk * (4 * p.col_tiles) + ct_cur * 4 + half * 2 + piece] = M;
          }
        }
      }
      flush_rows();
      trailing_row_block();
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_pair(tmem_base, 512);
}

// =========================================================================== stream-K GEMM
constexpr int kGemmStages = 4;
constexpr int kGemmABytes = 128 * kBK * 2;      // 16 KB: this CTA's 128 rows (or columns) of G
constexpr int kGemmBChunk = 64 * 64 * 2;        // 8 KB: 64 k-rows x 64 feature columns
constexpr int kGemmStageBytes = kGemmABytes + 4 * kGemmBChunk;   // 48 KB
constexpr int kGemmSmem = kGemmStages * kGemmStageBytes + kMiscBytes;
// fused reduce-scatter through TMA: every epilogue warp stages 32 rows x 32 fp32 columns (4 KB,
// 128B-swizzled) and one cp.reduce.async.bulk.tensor adds the box into the owner's accumulator
constexpr int kPeerStageBytes = 32 * 128;
constexpr int kGemmSmemPeer = kGemmSmem + kGemmEpiWarps * kPeerStageBytes;
struct PeerMaps {
  CUtensorMap m[8];
};

struct GemmProblem {
  int mode;            // 0: A = G (K-major), 1: A = G^T (MN-major)
  int m_tiles;         // tiles of 256 output rows
  int k_chunks;        // chunks of 64 along the contraction
  int64_t m_rows;      // valid output rows
  float* out;          // [m_rows, ld_out] fp32, accumulated with red.add
  int64_t ld_out;
  const float* scale;  // nullable device scalar applied to every partial before the red.add
  // Fused reduce-scatter (multi-rank text gradient): output row m belongs to rank m / rows_per_peer
  // and is added straight into that rank's accumulator through its peer mapping (NVLink).
  float* peers[8];
  int npeers;
  int64_t rows_per_peer;
  // feature slab of this problem: TMEM holds 512 accumulator columns, so dim > 512 is split
  int product;         // 0: B operand from tmb0, 1: from tmb1
  int d_off;           // first feature column of the slab (multiple of 128)
  int nhalf;           // 64-column feature chunks per CTA in this slab (1..4)
  uint32_t idesc[2];   // per MMA group of the slab
  // A segment that covers a WHOLE tile (every k chunk) never meets another cluster's partial sum:
  // its epilogue writes direct_scale * acc straight into the gradient (nullable) and the fp32
  // accumulator `out` is only used -- zeroed, added to and cast by gemm_fixup_kernel -- for tiles
  // that the stream-K schedule splits between clusters.
  void* direct;        // [m_rows, ld_direct] in direct_dtype, or NULL: always red.add into `out`
  int direct_dtype;
  int64_t ld_direct;
  const float* direct_scale;
};

struct GemmParams {
  GemmProblem prob[4];
  int nprob;
  int ncb;             // G block columns
  int dim;
  // fused reduce-scatter: when every add of the launch is out, the last CTA publishes
  // done_flags[w][done_slot] = done_gen on every rank w (n_done = 0: no signal)
  int* done_flags[8];
  int n_done;
  unsigned int* done_counter;
  int done_gen;
  int done_slot;
  // Two-part schedule (fused reduce-scatter): the units [0, split_units) -- the problems that add
  // into the peers -- are dealt evenly to ALL clusters and run first, the local problems after
  // them.  Every cluster's remote reductions are then out in the first part of the kernel, the
  // "done" flag is published when the last CTA finishes that part, and the transfer drains behind
  // the local product instead of after the kernel.  0 = one part.
  int64_t split_units;
};

// kPeerTma: problems with peers add their tiles with TMA reduce operations (one 4 KB box per
// 32 x 32 piece, 128-byte bursts on the wire) instead of 16-byte red.global.add per thread.
template <bool kPeerTma>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
pair_gemm_kernel(const __grid_constant__ CUtensorMap tma0, const __grid_constant__ CUtensorMap tma1,
                 const __grid_constant__ CUtensorMap tmb0, const __grid_constant__ CUtensorMap tmb1,
                 const __grid_constant__ PeerMaps peer_maps, const GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = smem_u32(smem);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();

  const uint32_t misc = smem_base + kGemmStages * kGemmStageBytes;
  const uint32_t bar_full = misc;
  const uint32_t bar_empty = bar_full + 8 * kGemmStages;
  const uint32_t bar_tfull = bar_empty + 8 * kGemmStages;
  const uint32_t bar_tempty = bar_tfull + 8;
  const uint32_t tmem_slot = bar_tempty + 8;
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - smem_base));

  if (threadIdx.x == 0 && (smem_base & 1023u) != 0) __trap();

  int64_t ubase[5];
  ubase[0] = 0;
  for (int i = 0; i < 4; ++i)
    ubase[i + 1] = ubase[i] + (i < p.nprob ? (int64_t)p.prob[i].m_tiles * p.prob[i].k_chunks : 0);
  const int64_t total = ubase[4];
  const int64_t ncl = gridDim.x >> 1;
  const int64_t cl = blockIdx.x >> 1;
  const int nparts = p.split_units > 0 ? 2 : 1;
  // this cluster's contiguous unit range inside part `part`
  auto part_range = [&](int part, int64_t& a, int64_t& b) {
    const int64_t lo = part == 0 ? 0 : p.split_units;
    const int64_t hi = (nparts == 2 && part == 0) ? p.split_units : total;
    a = lo + cl * (hi - lo) / ncl;
    b = lo + (cl + 1) * (hi - lo) / ncl;
  };

  if (warp == 0 && elect_one()) {
    prefetch_tensormap(&tma0);
    prefetch_tensormap(&tma1);
    prefetch_tensormap(&tmb0);
    prefetch_tensormap(&tmb1);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < kGemmStages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_tfull, 1);
    mbar_init(bar_tempty, 2 * kGemmEpiWarps);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(tmem_slot, 512);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // A segment = (problem, output tile, contiguous chunk range); all roles walk the same list.
  auto next_segment = [&](int64_t u, int64_t u1, int& pi, int& mt, int& k0, int& k1) {
    pi = 0;
    while (pi + 1 < p.nprob && u >= ubase[pi + 1]) ++pi;
    const int64_t local = u - ubase[pi];
    const int kc = p.prob[pi].k_chunks;
    mt = (int)(local / kc);
    k0 = (int)(local % kc);
    const int64_t left = u1 - u;
    k1 = (int)((int64_t)(kc - k0) < left ? kc : k0 + left);
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    if (elect_one()) {
      const uint32_t lead_full = mapa_rank(bar_full, 0);
      const uint64_t keep = policy_evict_last();        // features: re-read by every tile
      const uint64_t stream_pol = policy_evict_first(); // G: read once per product
      int stage = 0;
      uint32_t phase = 0;
      for (int part = 0; part < nparts; ++part) {
      int64_t u0, u1;
      part_range(part, u0, u1);
      for (int64_t u = u0; u < u1;) {
        int pi, mt, k0, k1;
        next_segment(u, u1, pi, mt, k0, k1);
        const int mode = p.prob[pi].mode;
        const int nhalf = p.prob[pi].nhalf;
        const int dch0 = p.prob[pi].d_off / 64;
        const uint32_t stage_tx = 2u * (uint32_t)(kGemmABytes + nhalf * kGemmBChunk);
        const CUtensorMap* tb = p.prob[pi].product ? &tmb1 : &tmb0;
        for (int kc = k0; kc < k1; ++kc) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t sa = smem_base + stage * kGemmStageBytes;
          const uint32_t sb = sa + kGemmABytes;
          const uint32_t lf = lead_full + 8 * stage;
          if (rank == 0) mbar_arrive_expect_tx(bar_full + 8 * stage, stage_tx);
          if (mode == 0) {
            const int64_t block = (int64_t)(2 * mt + (int)rank) * p.ncb + kc;
            tma_load_2d_pair_hint(sa, &tma0, lf, 0, (int32_t)(block * 128), stream_pol);
          } else {
            const int64_t block = (int64_t)(kc >> 1) * p.ncb + 4 * mt + 2 * (int)rank;
            const int32_t r0 = (int32_t)(block * 128 + (kc & 1) * 64);
            tma_load_2d_pair_hint(sa, &tma1, lf, 0, r0, stream_pol);
            tma_load_2d_pair_hint(sa + 8192, &tma1, lf, 0, r0 + 128, stream_pol);
          }
          for (int lc = 0; lc < nhalf; ++lc) {
            const int g = lc >> 1;
            const int cnt = min(2, nhalf - 2 * g);
            const int dchunk = dch0 + 4 * g + (int)rank * cnt + (lc - 2 * g);
            tma_load_2d_pair_hint(sb + lc * kGemmBChunk, tb, lf, dchunk * 64, kc * 64, keep);
          }
          if (++stage == kGemmStages) { stage = 0; phase ^= 1; }
        }
        u += k1 - k0;
      }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA)
    if (rank == 0 && elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int seg = 0;
      for (int part = 0; part < nparts; ++part) {
      int64_t u0, u1;
      part_range(part, u0, u1);
      for (int64_t u = u0; u < u1; ++seg) {
        int pi, mt, k0, k1;
        next_segment(u, u1, pi, mt, k0, k1);
        const int mode = p.prob[pi].mode;
        const int nmma = (p.prob[pi].nhalf + 1) / 2;
        mbar_wait(bar_tempty, (seg & 1) ^ 1);
        tc_fence_after();
        for (int kc = k0; kc < k1; ++kc) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * kGemmStageBytes;
          const uint32_t sb = sa + kGemmABytes;
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint64_t da = mode == 0 ? make_smem_desc_sw128(sa + k * 32, 16, 1024)
                                          : make_smem_desc_sw128(sa + k * 2048, 8192, 1024);
            for (int g = 0; g < nmma; ++g) {
              const uint64_t db = make_smem_desc_sw128(sb + g * 2 * kGemmBChunk + k * 2048, 8192, 1024);
              mma2_ss(tmem_base + g * 256, da, db, p.prob[pi].idesc[g], (kc > k0 || k > 0) ? 1u : 0u);
            }
          }
          tc_commit_pair(bar_empty + 8 * stage, 3);
          if (++stage == kGemmStages) { stage = 0; phase ^= 1; }
        }
        tc_commit_pair(bar_tfull, 3);
        u += k1 - k0;
      }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ------------------------------------------------------------ epilogue (both CTAs)
    const int q = warp & 3;
    const int half = (warp - kEpiWarp0) >> 2;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const uint32_t lead_tempty = mapa_rank(bar_tempty, 0);
    int seg = 0;
    for (int part = 0; part < nparts; ++part) {
    int64_t u0, u1;
    part_range(part, u0, u1);
    for (int64_t u = u0; u < u1; ++seg) {
      int pi, mt, k0, k1;
      next_segment(u, u1, pi, mt, k0, k1);
      const GemmProblem& pr = p.prob[pi];
      const int cols_half = pr.nhalf * 64;        // accumulator columns per epilogue half
      const int64_t m = (int64_t)mt * 256 + (int64_t)rank * kPM + q * 32 + lane;
      const bool m_ok = m < pr.m_rows;
      float* orow = pr.out + (m_ok ? m : 0) * pr.ld_out;
      if (pr.npeers > 0 && m_ok) {
        const int64_t w = m / pr.rows_per_peer;
        orow = pr.peers[w] + (m - w * pr.rows_per_peer) * pr.ld_out;
      }
      const float osc = pr.scale ? __ldg(pr.scale) : 1.0f;
      const bool whole = pr.direct != nullptr && k0 == 0 && k1 == pr.k_chunks;
      mbar_wait(bar_tfull, seg & 1);
      tc_fence_after();
      if (whole) {
        const float dsc = __ldg(pr.direct_scale);
        const int64_t mrow = m_ok ? m : 0;
        for (int cc = 0; cc < cols_half; cc += 32) {
          const int col = half * cols_half + cc;
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + lane_base + col, v);
          tmem_ld_wait();
          if (!m_ok) continue;
          const int dcol = pr.d_off + col;
          if (pr.direct_dtype == LATTE_F32) {
            float* o = static_cast<float*>(pr.direct) + mrow * pr.ld_direct + dcol;
#pragma unroll
            for (int i = 0; i < 32; i += 4)
              if (dcol + i < p.dim)
                *reinterpret_cast<float4*>(o + i) =
                    make_float4(dsc * __uint_as_float(v[i]), dsc * __uint_as_float(v[i + 1]),
                                dsc * __uint_as_float(v[i + 2]), dsc * __uint_as_float(v[i + 3]));
          } else {
            uint16_t* o = static_cast<uint16_t*>(pr.direct) + mrow * pr.ld_direct + dcol;
            const bool bf = pr.direct_dtype == LATTE_BF16;
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
              uint32_t w[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float lo = dsc * __uint_as_float(v[i + 2 * e]);
                const float hi = dsc * __uint_as_float(v[i + 2 * e + 1]);
                if (bf) {
                  __nv_bfloat162 b = __floats2bfloat162_rn(lo, hi);
                  w[e] = *reinterpret_cast<uint32_t*>(&b);
                } else {
                  w[e] = pack2(lo, hi);
                }
              }
              if (dcol + i < p.dim) *reinterpret_cast<uint4*>(o + i) = make_uint4(w[0], w[1], w[2], w[3]);
            }
          }
        }
      } else if (kPeerTma && pr.npeers > 0) {
        // owner of this warp's 32 rows (rows_per_peer % 32 == 0: a warp never straddles two ranks)
        const int64_t m0 = (int64_t)mt * 256 + (int64_t)rank * kPM + q * 32;
        const int64_t wown = m0 / pr.rows_per_peer;
        const bool any = m0 < pr.m_rows;
        const CUtensorMap* pm = &peer_maps.m[any ? wown : 0];
        const int32_t prow = (int32_t)(m0 - wown * pr.rows_per_peer);
        const uint32_t my_stage = smem_base + kGemmSmem + (uint32_t)(warp - kEpiWarp0) * kPeerStageBytes;
        const uint32_t row_addr = my_stage + lane * 128;
        for (int cc = 0; cc < cols_half; cc += 32) {
          const int col = half * cols_half + cc;
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + lane_base + col, v);
          tmem_ld_wait();
          if (!any || pr.d_off + col >= p.dim) continue;      // warp-uniform
          if (lane == 0) bulk_wait_group_read<0>();           // the previous box has left the stage
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            st_shared_v4(row_addr + ((uint32_t)(j ^ (lane & 7)) << 4),
                         __float_as_uint(osc * __uint_as_float(v[4 * j])),
                         __float_as_uint(osc * __uint_as_float(v[4 * j + 1])),
                         __float_as_uint(osc * __uint_as_float(v[4 * j + 2])),
                         __float_as_uint(osc * __uint_as_float(v[4 * j + 3])));
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            // rows past the owner's last row and columns past dim are clipped by the tensor map
            tma_reduce_add_2d(pm, my_stage, pr.d_off + col, prow);
            bulk_commit_group();
          }
        }
      } else {
        for (int cc = 0; cc < cols_half; cc += 32) {
          const int col = half * cols_half + cc;
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + lane_base + col, v);
          tmem_ld_wait();
          if (m_ok) {
            if (pr.npeers > 0) {           // peer memory: the adds of several GPUs meet here
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                if (pr.d_off + col + i < p.dim)
                  red_add_v4_sys(orow + pr.d_off + col + i, osc * __uint_as_float(v[i]),
                                 osc * __uint_as_float(v[i + 1]), osc * __uint_as_float(v[i + 2]),
                                 osc * __uint_as_float(v[i + 3]));
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                if (pr.d_off + col + i < p.dim)
                  red_add_v4(orow + pr.d_off + col + i, osc * __uint_as_float(v[i]), osc * __uint_as_float(v[i + 1]),
                             osc * __uint_as_float(v[i + 2]), osc * __uint_as_float(v[i + 3]));
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(lead_tempty);
      u += k1 - k0;
    }
    if (part == 0 && nparts == 2 && p.n_done > 0) {
      // End of the peer part of this CTA: wait until its reductions are complete (the MMAs of the
      // local part run meanwhile), then count the CTA; the last one publishes "done" on every rank.
      if (kPeerTma && lane == 0) {
        bulk_wait_group<0>();
        fence_proxy_async_all();
      }
      __threadfence_system();
      named_bar_sync(2, kGemmEpiWarps * 32);
      if (warp == kEpiWarp0 && lane == 0) {
        if (atomicInc(p.done_counter, gridDim.x - 1) == gridDim.x - 1) {
          __threadfence_system();
          for (int w = 0; w < p.n_done; ++w)
            if (p.done_flags[w]) st_release_sys(p.done_flags[w] + p.done_slot, p.done_gen);
        }
      }
    }
    }
  }

  // one-part schedule with peers: the adds are ordered before the cluster barrier and the last CTA
  // signals after it
  const bool tail_signal = p.n_done > 0 && nparts == 1;
  if (kPeerTma && warp >= kEpiWarp0 && lane == 0) {
    bulk_wait_group<0>();              // the TMA reductions of this warp are complete
    fence_proxy_async_all();
  }
  if (tail_signal && warp >= kEpiWarp0) __threadfence_system();
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_pair(tmem_base, 512);
  if (tail_signal && threadIdx.x == 0) {
    if (atomicInc(p.done_counter, gridDim.x - 1) == gridDim.x - 1) {
      __threadfence_system();
      for (int w = 0; w < p.n_done; ++w)
        if (p.done_flags[w]) st_release_sys(p.done_flags[w] + p.done_slot, p.done_gen);
    }
  }
}

// Tiles that the stream-K schedule splits between clusters go through the fp32 accumulators:
// kFixZero clears them before the GEMM, kFixCast turns them into gradients after it (whole tiles
// were written by the GEMM epilogue itself).  One CTA per (problem, tile); a tile is split when its
// first and last unit belong to different clusters of the contiguous schedule.
constexpr int kFixZero = 0;
constexpr int kFixCast = 1;
struct FixupParams {
  int nprob;
  int64_t ubase[5];
  int64_t total, ncl;
  int64_t split_units;               // two-part schedule of the GEMM (0 = one part)
  int m_tiles[4], k_chunks[4], d_off[4], ncols[4];
  int64_t m_rows[4];
  float* acc[4];
  int64_t ld_acc;
  void* out[4];
  int out_dtype;
  int64_t ld_out;
  const float* out_scale;
  int dim;
};

constexpr int kFixSplit = 8;      // CTAs per tile (32 rows each)

template <int MODE>
__global__ void __launch_bounds__(256) gemm_fixup_kernel(const FixupParams p) {
  int pi = 0, t = blockIdx.x;
  while (pi < p.nprob && t >= p.m_tiles[pi]) { t -= p.m_tiles[pi]; ++pi; }
  if (pi >= p.nprob || p.out[pi] == nullptr) return;
  int64_t u_first = p.ubase[pi] + (int64_t)t * p.k_chunks[pi];
  // unit range of the schedule part this problem belongs to
  const bool second = p.split_units > 0 && u_first >= p.split_units;
  const int64_t part_lo = second ? p.split_units : 0;
  const int64_t part_n = (p.split_units > 0 && !second ? p.split_units : p.total) - part_lo;
  u_first -= part_lo;
  const int64_t u_last = u_first + p.k_chunks[pi] - 1;
  if (cluster_of_tile(u_first, part_n, p.ncl) == cluster_of_tile(u_last, part_n, p.ncl)) return;
  const int c0 = p.d_off[pi];
  const int c1 = min(p.dim, c0 + p.ncols[pi]);
  const int per_row = (c1 - c0) / 4;
  constexpr int kRowsPerCta = 256 / kFixSplit;
  const int64_t r0 = (int64_t)t * 256 + (int64_t)blockIdx.y * kRowsPerCta;
  const int rows = (int)min((int64_t)kRowsPerCta, p.m_rows[pi] - r0);
  if (rows <= 0) return;
  const float cs = MODE == kFixCast ? __ldg(p.out_scale) : 0.f;
  for (int idx = threadIdx.x; idx < rows * per_row; idx += 256) {
    const int64_t r = r0 + idx / per_row;
    const int c = c0 + (idx % per_row) * 4;
    float4* a = reinterpret_cast<float4*>(p.acc[pi] + r * p.ld_acc + c);
    if (MODE == kFixZero) {
      *a = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      const float4 v = __ldcs(a);
      const float o[4] = {v.x * cs, v.y * cs, v.z * cs, v.w * cs};
      if (p.out_dtype == LATTE_F32) {
        *reinterpret_cast<float4*>(static_cast<float*>(p.out[pi]) + r * p.ld_out + c) =
            make_float4(o[0], o[1], o[2], o[3]);
      } else if (p.out_dtype == LATTE_BF16) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(o[0], o[1]), hi = __floats2bfloat162_rn(o[2], o[3]);
        *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(p.out[pi]) + r * p.ld_out + c) =
            make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
      } else {
        *reinterpret_cast<uint2*>(static_cast<__half*>(p.out[pi]) + r * p.ld_out + c) =
            make_uint2(pack2(o[0], o[1]), pack2(o[2], o[3]));
      }
    }
  }
}

// out[i, d] = coef * s * 2^-13 * acc[i, d]   (coef = grad_loss * grad_mult / (2 n_loc))
__global__ void grad_scale_cast_kernel(const float* acc0, const float* acc1, int64_t ld_acc,
                                       void* out0, void* out1, int out_dtype,
                                       int64_t ld_out, int64_t rows, int64_t dim,
                                       const float* out_scale) {
  const float* acc = blockIdx.y ? acc1 : acc0;
  void* out = blockIdx.y ? out1 : out0;
  const int64_t per_row = dim / 4;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * per_row) return;
  const int64_t r = idx / per_row, c = (idx % per_row) * 4;
  const float cs = __ldg(out_scale);
  const float4 a = *reinterpret_cast<const float4*>(acc + r * ld_acc + c);
  const float o[4] = {a.x * cs, a.y * cs, a.z * cs, a.w * cs};
  if (out_dtype == LATTE_F32) {
    float* po = reinterpret_cast<float*>(out) + r * ld_out + c;
#pragma unroll
    for (int k = 0; k < 4; ++k) po[k] = o[k];
  } else if (out_dtype == LATTE_BF16) {
    __nv_bfloat16* po = reinterpret_cast<__nv_bfloat16*>(out) + r * ld_out + c;
#pragma unroll
    for (int k = 0; k < 4; ++k) po[k] = __float2bfloat16_rn(o[k]);
  } else {
    __half* po = reinterpret_cast<__half*>(out) + r * ld_out + c;
#pragma unroll
    for (int k = 0; k < 4; ++k) po[k] = __float2half_rn(o[k]);
  }
}

// =========================================================================== host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess)
      return nullptr;
    if (q != cudaDriverEntryPointSuccess) return nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

// 16-bit row-major [rows, cols] (row pitch ld elements) -> boxes [box_rows x 64 cols], 128B swizzle
int make_map16(CUtensorMap* map, const void* base, int dtype, int64_t rows, int64_t cols, int64_t ld,
               int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return LATTE_ERR_CUDA;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, dtype == LATTE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                           : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                  2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? LATTE_OK : LATTE_ERR_CUDA;
}

}  // namespace

namespace {
template <int MODE, bool TAIL>
int launch_sweep_t(int grid, cudaStream_t stream, const CUtensorMap& tmy, const CUtensorMap& tmg,
                   const CUtensorMap& tmx, const SweepParams& p) {
  LATTE_CUDA_OK(cudaFuncSetAttribute(pair_sweep_kernel<MODE, TAIL>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, kSweepSmem));
  pair_sweep_kernel<MODE, TAIL><<<grid, kSweepThreads, kSweepSmem, stream>>>(tmy, tmg, tmx, p);
  return LATTE_OK;
}
int launch_sweep(int mode, bool tail, int grid, cudaStream_t stream, const CUtensorMap& tmy,
                 const CUtensorMap& tmg, const CUtensorMap& tmx, const SweepParams& p) {
  if (mode == kModeGrad)
    return tail ? launch_sweep_t<kModeGrad, true>(grid, stream, tmy, tmg, tmx, p)
                : launch_sweep_t<kModeGrad, false>(grid, stream, tmy, tmg, tmx, p);
  if (mode == kModeFwdBoth)
    return tail ? launch_sweep_t<kModeFwdBoth, true>(grid, stream, tmy, tmg, tmx, p)
                : launch_sweep_t<kModeFwdBoth, false>(grid, stream, tmy, tmg, tmx, p);
  if (mode == kModeSigFwd)
    return tail ? launch_sweep_t<kModeSigFwd, true>(grid, stream, tmy, tmg, tmx, p)
                : launch_sweep_t<kModeSigFwd, false>(grid, stream, tmy, tmg, tmx, p);
  if (mode == kModeSigGrad)
    return tail ? launch_sweep_t<kModeSigGrad, true>(grid, stream, tmy, tmg, tmx, p)
                : launch_sweep_t<kModeSigGrad, false>(grid, stream, tmy, tmg, tmx, p);
  return tail ? launch_sweep_t<kModeFwdRows, true>(grid, stream, tmy, tmg, tmx, p)
              : launch_sweep_t<kModeFwdRows, false>(grid, stream, tmy, tmg, tmx, p);
}
}  // namespace

bool clip_pair_supported(int dtype, int64_t dim, int64_t ldx, int64_t ldy, const void* x,
                         const void* y) {
  if (dtype != LATTE_BF16 && dtype != LATTE_F16) return false;
  if (dim < 8 || dim > 768 || (dim % 8) != 0) return false;
  if ((ldx % 8) != 0 || (ldy % 8) != 0) return false;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 15)) return false;
  return true;
}

PairGeom clip_pair_geom(int64_t n_loc, int64_t n_all) {
  PairGeom g;
  g.row_blocks = (int)((n_loc + 255) / 256);
  g.col_tiles = (int)((n_all + 255) / 256) * 2;
  g.ncb = g.col_tiles * 2;
  g.g_elems = (size_t)g.row_blocks * 2 * (size_t)g.ncb * kGBlockElems;
  return g;
}

int clip_pair_ds_count() { return device_sm_count() / 2 * 2; }

PairFwdGeom clip_pair_fwd_geom(int64_t n_loc, int64_t n_all) {
  const PairGeom geo = clip_pair_geom(n_loc, n_all);
  PairFwdGeom f;
  f.row_blocks = geo.row_blocks;
  f.col_tiles = geo.col_tiles;
  f.total = (int64_t)geo.row_blocks * geo.col_tiles;
  f.ncl = device_sm_count() / 2;
  if (f.total < f.ncl) f.ncl = (int)f.total;
  const int64_t tpc = f.total / f.ncl;              // >= 1 tiles per cluster
  f.slots = (int)((geo.col_tiles + tpc - 1) / tpc + 1);
  f.ld_colpart = (int64_t)geo.col_tiles * kTN;
  return f;
}

int clip_pair_fwd_sweep(const PairFwdArgs& a, cudaStream_t stream) {
  if (!clip_pair_supported(a.dtype, a.dim, a.ldx, a.ldy, a.x, a.y)) return LATTE_ERR_UNSUPPORTED;
  const PairFwdGeom f = clip_pair_fwd_geom(a.n_loc, a.n_all);
  CUtensorMap tmy;
  int rc = make_map16(&tmy, a.y, a.dtype, a.n_all, a.dim, a.ldy, 64);
  if (rc) return rc;
  SweepParams p = {};
  p.x = a.x; p.ldx = a.ldx;
  p.n_loc = a.n_loc; p.n_all = a.n_all; p.dim = a.dim;
  p.label_offset = a.label_offset;
  p.logit_scale = a.logit_scale;
  p.part_max = a.part_max; p.part_sum = a.part_sum; p.diag = a.diag;
  p.col_part = a.col_part; p.col_ref = a.col_ref; p.ld_colpart = f.ld_colpart;
  p.zero2 = a.zero2;
  p.landed = a.landed; p.landed_gen = a.landed_gen;
  p.rows_per_rank = a.rows_per_rank > 0 ? a.rows_per_rank : a.n_all;
  p.kch = (int)((a.dim + kBK - 1) / kBK);
  p.tail_chunks = p.kch > kTmemChunks ? p.kch - kTmemChunks : 0;
  p.cps = p.tail_chunks ? 2 : kChunksPerStage;
  CUtensorMap tmx;
  rc = make_map16(&tmx, a.x, a.dtype, a.n_loc, a.dim, a.ldx, 128);
  if (rc) return rc;
  p.col_tiles = f.col_tiles;
  p.row_blocks = f.row_blocks;
  p.ncb = f.col_tiles * 2;
  p.idesc = make_idesc_f16(256, kTN, a.dtype == LATTE_BF16 ? 1u : 0u, 0, 0);
  rc = launch_sweep(a.col_part ? kModeFwdBoth : kModeFwdRows, p.tail_chunks > 0, 2 * f.ncl, stream,
                    tmy, tmy, tmx, p);
  if (rc) return rc;
  LATTE_LAUNCH_OK();
  return LATTE_OK;
}

int clip_pair_sweep(const PairSweepArgs& a, cudaStream_t stream) {
  if (!clip_pair_supported(a.dtype, a.dim, a.ldx, a.ldy, a.x, a.y)) return LATTE_ERR_UNSUPPORTED;
  const PairGeom geo = clip_pair_geom(a.n_loc, a.n_all);
  CUtensorMap tmy, tmg;
  int rc = make_map16(&tmy, a.y, a.dtype, a.n_all, a.dim, a.ldy, 64);
  if (rc) return rc;
  // blocked view of G: every 16 KB block is 128 consecutive rows of a [blocks*128, 64] matrix
  const int64_t g_rows = (int64_t)geo.row_blocks * 2 * geo.ncb * 128;
  rc = make_map16(&tmg, a.g, LATTE_F16, g_rows, 64, 64, 32);
  if (rc) return rc;
  SweepParams p = {};
  p.x = a.x; p.ldx = a.ldx;
  p.n_loc = a.n_loc; p.n_all = a.n_all; p.dim = a.dim;
  p.label_offset = a.label_offset;
  p.logit_scale = a.logit_scale;
  p.lse_a2 = a.lse_a2; p.lse_b2 = a.lse_b2;
  p.e_a = a.e_a; p.einv_b = a.einv_b; p.fast_flag = a.fast_flag; p.gscale_log2 = a.gscale_log2;
  p.nll_a = a.nll_a; p.nll_b = a.nll_b;
  p.cb = a.cross_terms ? 1.f : 0.f;
  p.cd = a.cross_terms ? 2.f : 1.f;
  if (a.no_label) { p.cd = 0.f; p.nll_a = p.nll_b = nullptr; }
  // one sweep standing for both directions (world size 1) carries both softmax terms in ds
  p.ds_cb = a.ds_both ? p.cb : 0.f;
  p.ds_cd = a.no_label ? 0.f : (a.ds_both ? p.cd : 1.f);
  p.ds_partial = a.ds_partial;
  p.kch = (int)((a.dim + kBK - 1) / kBK);
  p.tail_chunks = p.kch > kTmemChunks ? p.kch - kTmemChunks : 0;
  p.cps = p.tail_chunks ? 2 : kChunksPerStage;
  CUtensorMap tmx;
  rc = make_map16(&tmx, a.x, a.dtype, a.n_loc, a.dim, a.ldx, 128);
  if (rc) return rc;
  p.col_tiles = geo.col_tiles;
  p.row_blocks = geo.row_blocks;
  p.ncb = geo.ncb;
  p.idesc = make_idesc_f16(256, kTN, a.dtype == LATTE_BF16 ? 1u : 0u, 0, 0);
  int ncl = device_sm_count() / 2;
  const int64_t total = (int64_t)geo.row_blocks * geo.col_tiles;
  if (total < ncl) ncl = (int)total;
  // every CTA of the launch writes its ds partial; unused slots are zeroed by the caller
  rc = launch_sweep(kModeGrad, p.tail_chunks > 0, 2 * ncl, stream, tmy, tmg, tmx, p);
  if (rc) return rc;
  LATTE_LAUNCH_OK();
  return LATTE_OK;
}

// SigLIP sweeps (loss.py:509-519): forward = per-CTA partial sums of the loss terms, backward = G
// (fp16, scaled by 2^13) + per-CTA partials of d loss / d logit_scale and d loss / d logit_bias.
int clip_pair_sig_sweep(const PairSigArgs& a, cudaStream_t stream) {
  if (!clip_pair_supported(a.dtype, a.dim, a.ldx, a.ldy, a.x, a.y)) return LATTE_ERR_UNSUPPORTED;
  const PairGeom geo = clip_pair_geom(a.n_loc, a.n_all);
  CUtensorMap tmy, tmg, tmx;
  int rc = make_map16(&tmy, a.y, a.dtype, a.n_all, a.dim, a.ldy, 64);
  if (rc) return rc;
  tmg = tmy;
  if (a.g) {
    const int64_t g_rows = (int64_t)geo.row_blocks * 2 * geo.ncb * 128;
    rc = make_map16(&tmg, a.g, LATTE_F16, g_rows, 64, 64, 32);
    if (rc) return rc;
  }
  rc = make_map16(&tmx, a.x, a.dtype, a.n_loc, a.dim, a.ldx, 128);
  if (rc) return rc;
  SweepParams p = {};
  p.x = a.x; p.ldx = a.ldx;
  p.n_loc = a.n_loc; p.n_all = a.n_all; p.dim = a.dim;
  p.label_offset = a.label_offset;
  p.logit_scale = a.logit_scale;
  p.logit_bias = a.logit_bias;
  p.ds_partial = a.ds_partial;
  p.aux_partial = a.aux_partial;
  p.kch = (int)((a.dim + kBK - 1) / kBK);
  p.tail_chunks = p.kch > kTmemChunks ? p.kch - kTmemChunks : 0;
  p.cps = p.tail_chunks ? 2 : kChunksPerStage;
  p.col_tiles = geo.col_tiles;
  p.row_blocks = geo.row_blocks;
  p.ncb = geo.ncb;
  p.idesc = make_idesc_f16(256, kTN, a.dtype == LATTE_BF16 ? 1u : 0u, 0, 0);
  int ncl = device_sm_count() / 2;
  const int64_t total = (int64_t)geo.row_blocks * geo.col_tiles;
  if (total < ncl) ncl = (int)total;
  rc = launch_sweep(a.g ? kModeSigGrad : kModeSigFwd, p.tail_chunks > 0, 2 * ncl, stream, tmy, tmg, tmx, p);
  if (rc) return rc;
  LATTE_LAUNCH_OK();
  return LATTE_OK;
}

namespace {
// Problem list of one gradient-GEMM launch (shared by the GEMM and the fix-up kernels): one
// problem per (product, feature slab of <= 512 accumulator columns).
int build_gemm_problems(const PairGemmArgs& a, GemmParams& p, int64_t& total, int& ncl) {
  const PairGeom geo = clip_pair_geom(a.n_loc, a.n_all);
  p = GemmParams{};
  p.ncb = geo.ncb;
  p.dim = (int)a.dim;
  const int units128 = (int)((a.dim + 127) / 128);
  const int nslab = (units128 + 3) / 4;
  const int nproducts = a.x16 ? 2 : 1;
  if (a.dy_peers && a.n_peers > 8) return LATTE_ERR_UNSUPPORTED;
  if (a.feat_dtype != LATTE_F16) return LATTE_ERR_UNSUPPORTED;   // A (= G) is fp16: one format per MMA
  // direct stores need 16-byte vectors on the output rows
  const bool direct_ok = a.out_scale != nullptr && (a.ld_out % 8) == 0;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  p.nprob = 0;
  total = 0;
  p.split_units = 0;
  // with peers the text-side product (the one that adds into the other ranks) goes first
  const bool peers_first = nproducts == 2 && a.dy_peers && a.n_peers > 1;
  for (int pidx = 0; pidx < nproducts; ++pidx) {
    const int prod = peers_first ? 1 - pidx : pidx;
    if (peers_first && pidx == 1) p.split_units = total;
    int done = 0;
    for (int sl = 0; sl < nslab; ++sl) {
      const int nh = (units128 - done + (nslab - sl) - 1) / (nslab - sl);   // even split
      GemmProblem& q = p.prob[p.nprob++];
      q.product = prod;
      q.d_off = done * 128;
      q.nhalf = nh;
      done += nh;
      for (int g = 0; g < 2; ++g) {
        const int cnt = nh - 2 * g >= 2 ? 2 : (nh - 2 * g == 1 ? 1 : 0);
        q.idesc[g] = cnt ? make_idesc_f16(256, 2 * cnt * 64, 0u, prod, 1) : 0u;
      }
      q.scale = nullptr;
      q.npeers = 0;
      q.rows_per_peer = 1;
      q.direct = nullptr; q.direct_dtype = a.out_dtype; q.ld_direct = a.ld_out;
      q.direct_scale = a.out_scale;
      for (int w = 0; w < 8; ++w) q.peers[w] = nullptr;
      if (prod == 0) {
        // dX = G . Y : rows of this rank, contraction over all columns
        q.mode = 0;
        q.m_tiles = geo.row_blocks;
        q.k_chunks = (int)((a.n_all + 63) / 64);
        q.m_rows = a.n_loc;
        q.out = a.dx32;
        q.ld_out = a.ld32;
        if (direct_ok && a.dx_out && al16(a.dx_out)) q.direct = a.dx_out;
      } else {
        // dY = G^T . X : rows = columns of G, contraction over the rows of G
        q.mode = 1;
        q.m_tiles = geo.col_tiles / 2;
        q.k_chunks = (int)((a.n_loc + 63) / 64);
        q.m_rows = a.n_all;
        q.out = a.dy32;
        q.ld_out = a.ld_dy32;
        q.scale = a.dy_scale;
        if (a.dy_peers && a.n_peers > 1) {
          q.npeers = a.n_peers;
          q.rows_per_peer = a.n_all / a.n_peers;
          for (int w = 0; w < a.n_peers; ++w) q.peers[w] = a.dy_peers[w];
        } else if (direct_ok && a.dy_out && al16(a.dy_out) && !a.dy_scale) {
          q.direct = a.dy_out;
        }
      }
      total += (int64_t)q.m_tiles * q.k_chunks;
    }
  }
  ncl = device_sm_count() / 2;
  if (total < ncl) ncl = (int)total;
  p.n_done = a.dy_peers ? a.n_done : 0;
  for (int w = 0; w < 8; ++w) p.done_flags[w] = w < p.n_done ? a.done_flags[w] : nullptr;
  p.done_counter = a.done_counter;
  p.done_gen = a.done_gen;
  p.done_slot = a.done_slot;
  if (p.n_done > 0 && !p.done_counter) return LATTE_ERR_BAD_ARG;
  return LATTE_OK;
}
}  // namespace

namespace {
// fp32 row-major [rows, cols] -> boxes [32 rows x 32 cols] (128-byte rows), 128B swizzle
int make_map32(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return LATTE_ERR_CUDA;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32u, 32u};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? LATTE_OK : LATTE_ERR_CUDA;
}
}  // namespace

int clip_pair_gemm(const PairGemmArgs& a, cudaStream_t stream) {
  const PairGeom geo = clip_pair_geom(a.n_loc, a.n_all);
  const int64_t g_rows = (int64_t)geo.row_blocks * 2 * geo.ncb * 128;
  CUtensorMap tma0, tma1, tmb0, tmb1;
  int rc = make_map16(&tma0, a.g, LATTE_F16, g_rows, 64, 64, 128);
  if (rc) return rc;
  rc = make_map16(&tma1, a.g, LATTE_F16, g_rows, 64, 64, 64);
  if (rc) return rc;
  rc = make_map16(&tmb0, a.y16, a.feat_dtype, a.n_all, a.dim, a.ldy16, 64);
  if (rc) return rc;
  rc = make_map16(&tmb1, a.x16 ? a.x16 : a.y16, a.feat_dtype, a.x16 ? a.n_loc : a.n_all, a.dim,
                  a.x16 ? a.ldx16 : a.ldy16, 64);
  if (rc) return rc;
  GemmParams p;
  int64_t total;
  int ncl;
  rc = build_gemm_problems(a, p, total, ncl);
  if (rc) return rc;
  // TMA reductions into the peers' accumulators: shard rows a multiple of the 32-row warp pieces,
  // 16-byte aligned rows (LATTE_B200_PEER_RED=1 keeps the per-thread red.global.add form)
  static const bool no_tma = []() {
    const char* e = getenv("LATTE_B200_PEER_RED");
    return e != nullptr && e[0] == '1';
  }();
  PeerMaps pm;
  bool peer_tma = false;
  if (a.dy_peers && a.n_peers > 1 && !no_tma) {
    const int64_t rpp = a.n_all / a.n_peers;
    peer_tma = (rpp % 32) == 0 && (a.ld_dy32 % 4) == 0;
    for (int w = 0; w < a.n_peers && peer_tma; ++w) {
      if ((reinterpret_cast<uintptr_t>(a.dy_peers[w]) & 15) != 0) peer_tma = false;
      else if (make_map32(&pm.m[w], a.dy_peers[w], rpp, a.dim, a.ld_dy32)) peer_tma = false;
    }
    if (peer_tma)
      for (int w = a.n_peers; w < 8; ++w) pm.m[w] = pm.m[0];
  }
  if (peer_tma) {
    LATTE_CUDA_OK(cudaFuncSetAttribute(pair_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       kGemmSmemPeer));
    pair_gemm_kernel<true><<<2 * ncl, kThreads, kGemmSmemPeer, stream>>>(tma0, tma1, tmb0, tmb1, pm, p);
  } else {
    for (int w = 0; w < 8; ++w) pm.m[w] = tma0;        // unused
    LATTE_CUDA_OK(cudaFuncSetAttribute(pair_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       kGemmSmem));
    pair_gemm_kernel<false><<<2 * ncl, kThreads, kGemmSmem, stream>>>(tma0, tma1, tmb0, tmb1, pm, p);
  }
  LATTE_LAUNCH_OK();
  return LATTE_OK;
}

// cast = 0: clear the fp32 accumulators of the tiles the GEMM's schedule splits (before the GEMM);
// cast = 1: turn those tiles into gradients (after it).  No-op for problems without direct output.
int clip_pair_gemm_fixup(const PairGemmArgs& a, int cast, cudaStream_t stream) {
  GemmParams g;
  int64_t total;
  int ncl;
  int rc = build_gemm_problems(a, g, total, ncl);
  if (rc) return rc;
  FixupParams f = {};
  f.nprob = g.nprob;
  f.total = total; f.ncl = ncl;
  f.split_units = g.split_units;
  f.ld_acc = a.ld32; f.out_dtype = a.out_dtype; f.ld_out = a.ld_out; f.out_scale = a.out_scale;
  f.dim = (int)a.dim;
  int tiles = 0;
  bool any = false;
  f.ubase[0] = 0;
  for (int i = 0; i < g.nprob; ++i) {
    const GemmProblem& q = g.prob[i];
    f.m_tiles[i] = q.m_tiles; f.k_chunks[i] = q.k_chunks; f.d_off[i] = q.d_off;
    f.ncols[i] = q.nhalf * 128; f.m_rows[i] = q.m_rows; f.acc[i] = q.out; f.out[i] = q.direct;
    f.ubase[i + 1] = f.ubase[i] + (int64_t)q.m_tiles * q.k_chunks;
    if (q.direct && q.ld_out != a.ld32) return LATTE_ERR_BAD_ARG;
    any = any || q.direct != nullptr;
    tiles += q.m_tiles;
  }
  if (!any || tiles == 0) return LATTE_OK;
  if (cast) gemm_fixup_kernel<kFixCast><<<dim3(tiles, kFixSplit), 256, 0, stream>>>(f);
  else gemm_fixup_kernel<kFixZero><<<dim3(tiles, kFixSplit), 256, 0, stream>>>(f);
  LATTE_LAUNCH_OK();
  return LATTE_OK;
}

// true when clip_pair_gemm writes this product's whole tiles itself (fix-up kernels handle the rest)
bool clip_pair_gemm_direct(const PairGemmArgs& a, int product) {
  GemmParams g;
  int64_t total;
  int ncl;
  if (build_gemm_problems(a, g, total, ncl)) return false;
  for (int i = 0; i < g.nprob; ++i)
    if (g.prob[i].product == product && g.prob[i].direct) return true;
  return false;
}

int clip_pair_scale_cast(const float* acc0, const float* acc1, int64_t ld_acc, void* out0, void* out1,
                         int out_dtype, int64_t ld_out, int64_t rows, int64_t dim,
                         const float* out_scale, cudaStream_t stream) {
  const int64_t work = rows * (dim / 4);
  grad_scale_cast_kernel<<<dim3((unsigned)((work + 255) / 256), acc1 ? 2 : 1), 256, 0, stream>>>(
      acc0, acc1, ld_acc, out0, out1, out_dtype, ld_out, rows, dim, out_scale);
  LATTE_LAUNCH_OK();
  return LATTE_OK;
}

}  // namespace latte

// ClipLoss row kernels on tcgen05 tensor cores (sm_100a), bf16 / fp16 features.
//
// Replaces, without ever storing the logits, open_clip/loss.py:109-116 (the two logit
// GEMMs), :126-129 (the two cross-entropies) and their autograd.
//
//   forward  : for a block of 128 rows of X, sweep Y in tiles of 128 rows; the tile
//              S = X_blk . Y_tile^T is accumulated by tcgen05.mma into TMEM (fp32),
//              epilogue warps read it back with tcgen05.ld and keep a flash-style online
//              (max, sum) per row in base-2 units.
//   backward : same sweep; the epilogue turns S into G = exp2(S*c - lseA_i) +
//              cb*exp2(S*c - lseB_j), writes G (16-bit) back into the TMEM columns S
//              occupied, and a second tcgen05.mma (A from TMEM, B = the same Y tile read
//              MN-major) accumulates dX_blk += G . Y_tile in TMEM.  A 128 x D fp32
//              accumulator for D = 512 would fill all 512 TMEM columns, so the feature
//              axis of dX is split into slabs of 256 columns across CTAs
//              (blockIdx.y); each slab CTA recomputes S.
//
// Warp roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM
// allocator, warps 4-11 = epilogue (warp%4 selects the 32-lane TMEM quarter, (warp-4)/4
// the 64-column half of the tile).
//
// Shared memory: X block resident as D/64 chunks of [128 rows x 64 cols] (16 KB each,
// 128B-swizzled by TMA), plus a ring of 16 KB chunks for Y.  One CTA per SM.
#include "latte_common.cuh"
#include "tc_ptx.cuh"

namespace latte {

using namespace ptx;

namespace {

constexpr int kBM = 128;            // X rows per CTA
constexpr int kBN = 128;            // Y rows per tile
constexpr int kBK = 64;             // feature columns per smem chunk (128 bytes)
constexpr int kChunkBytes = kBM * kBK * 2;
constexpr int kThreads = 384;
constexpr int kEpiWarp0 = 4;
constexpr int kNumEpiWarps = 8;
constexpr int kMaxChunks = 14;      // (kch + stages) * 16 KB + misc <= 227 KB
constexpr int kMiscBytes = 1024;
constexpr int kMaxStages = 6;

struct FwdParams {
  int64_t n_loc, n_all, dim;
  int64_t label_offset;
  const float* logit_scale;
  float* part_max;  // [2 * splits, n_loc]: one partial per (column split, tile half)
  float* part_sum;
  float* diag;
  int kch;          // ceil(dim / 64)
  int stages;
  int tiles_total;  // ceil(n_all / 128)
  int splits;       // column splits (gridDim.y)
  uint32_t idesc;
  const int* gate;  // nullable: run only when *gate != 0 (exact fallback of the pair forward)
};

struct BwdParams {
  int64_t n_loc, n_all, dim;
  int64_t label_offset;
  const float* logit_scale;
  const float* lse_a2;
  const float* lse_b2;
  const float* grad_loss;
  float grad_mult;
  float cb;          // weight of the y-side softmax term (0 or 1)
  float cd;          // weight of the one-hot term (1 or 2)
  void* dx; int grad_dtype; int64_t ld_dx;
  float* ds_partial;
  int kch;
  int stages;
  int tiles_total;
  int is_bf16;
  uint32_t idesc_g1;
  uint32_t idesc_g2;
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// =========================================================================== forward
__global__ void __launch_bounds__(kThreads, 1)
clip_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmy,
                   const FwdParams p) {
  if (p.gate != nullptr && *p.gate == 0) return;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = smem_u32(smem);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const uint32_t x_smem = smem_base;
  const uint32_t ring = smem_base + p.kch * kChunkBytes;
  const uint32_t misc = ring + p.stages * kChunkBytes;
  // barrier layout in misc: [0] x_full, [1..6] full, [7..12] empty, [13,14] tmem_full,
  // [15,16] tmem_empty, then the TMEM base slot.
  const uint32_t bar_x = misc;
  const uint32_t bar_full = misc + 8;
  const uint32_t bar_empty = misc + 8 + 8 * kMaxStages;
  const uint32_t bar_tfull = misc + 8 + 16 * kMaxStages;
  const uint32_t bar_tempty = bar_tfull + 16;
  const uint32_t tmem_slot = bar_tempty + 16;
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - smem_base));

  if (threadIdx.x == 0 && (smem_base & 1023u) != 0) __trap();

  const int row0 = blockIdx.x * kBM;
  // tiles [t0, t1) of this column split
  const int per = (p.tiles_total + p.splits - 1) / p.splits;
  const int t0 = blockIdx.y * per;
  const int t1 = min(p.tiles_total, t0 + per);
  const int ntiles = max(0, t1 - t0);

  if (warp == 0 && elect_one()) {
    prefetch_tensormap(&tmx);
    prefetch_tensormap(&tmy);
  }
  if (warp == 1 && elect_one()) {
    mbar_init(bar_x, 1);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_tfull + 8 * b, 1);
      mbar_init(bar_tempty + 8 * b, kNumEpiWarps);      // every epilogue warp drains its part
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one() && ntiles > 0) {
      mbar_arrive_expect_tx(bar_x, p.kch * kChunkBytes);
      for (int c = 0; c < p.kch; ++c)
        tma_load_2d(x_smem + c * kChunkBytes, &tmx, bar_x, c * kBK, row0);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = t0; t < t1; ++t) {
        for (int c = 0; c < p.kch; ++c) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          mbar_arrive_expect_tx(bar_full + 8 * stage, kChunkBytes);
          tma_load_2d(ring + stage * kChunkBytes, &tmy, bar_full + 8 * stage, c * kBK, t * kBN);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one() && ntiles > 0) {
      mbar_wait(bar_x, 0);
      tc_fence_after();
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < ntiles; ++it) {
        const int buf = it & 1;
        mbar_wait(bar_tempty + 8 * buf, ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * kBN;
        for (int c = 0; c < p.kch; ++c) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t a_addr = x_smem + c * kChunkBytes;
          const uint32_t b_addr = ring + stage * kChunkBytes;
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint64_t da = make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
            const uint64_t db = make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
            mma_ss(tmem_d, da, db, p.idesc, (c | k) != 0);
          }
          tc_commit(bar_empty + 8 * stage);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        tc_commit(bar_tfull + 8 * buf);
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ------------------------------------------------------------ epilogue
    const int q = warp & 3;
    const int half = (warp - kEpiWarp0) >> 2;
    const int row = q * 32 + lane;
    const int64_t grow = (int64_t)row0 + row;
    const float c2 = __ldg(p.logit_scale) * kLog2e;
    const int64_t label = p.label_offset + grow;
    float m = -INFINITY, l = 0.f;
    for (int it = 0; it < ntiles; ++it) {
      const int t = t0 + it;
      const int buf = it & 1;
      mbar_wait(bar_tfull + 8 * buf, (it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * kBN + half * 64;
      uint32_t r[2][32];
      tmem_ld_32x32(taddr, r[0]);
      tmem_ld_32x32(taddr + 32, r[1]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);

      const int64_t col0 = (int64_t)t * kBN + half * 64;
      if (col0 + 64 > p.n_all) {
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (col0 + h * 32 + i >= p.n_all) r[h][i] = 0xff800000u;  // -inf
      }
      if (label >= col0 && label < col0 + 64 && grow < p.n_loc) {
        const int want = (int)(label - col0);
        float dv = 0.f;
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (h * 32 + i == want) dv = __uint_as_float(r[h][i]);
        p.diag[grow] = dv;
      }
      float tmax = -INFINITY;
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int i = 0; i < 32; ++i) tmax = fmaxf(tmax, __uint_as_float(r[h][i]));
      const float m_new = fmaxf(m, tmax * c2);
      if (m_new > -INFINITY) {
        float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          acc0 += fast_exp2(fmaf(__uint_as_float(r[0][i]), c2, -m_new));
          acc1 += fast_exp2(fmaf(__uint_as_float(r[1][i]), c2, -m_new));
        }
        l = l * fast_exp2(m - m_new) + (acc0 + acc1);
        m = m_new;
      }
    }
    if (grow < p.n_loc) {
      const int64_t slot = ((int64_t)blockIdx.y * 2 + half) * p.n_loc + grow;
      p.part_max[slot] = m;
      p.part_sum[slot] = l;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 256);
}

// =========================================================================== backward
template <typename OutT>
__device__ __forceinline__ void store_out(OutT* p, float v);
template <>
__device__ __forceinline__ void store_out<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void store_out<__nv_bfloat16>(__nv_bfloat16* p, float v) {
  *p = __float2bfloat16_rn(v);
}
template <>
__device__ __forceinline__ void store_out<__half>(__half* p, float v) { *p = __float2half_rn(v); }

// G is handed to the second MMA as fp16 (11-bit significand; bf16's 8 bits cost ~5e-3 of
// gradient accuracy) scaled by 2^13 so that weights down to ~1e-11 stay representable.
// tcgen05.mma kind::f16 needs A and B in the same format (a mixed fp16 x bf16 descriptor
// raises an illegal-instruction fault on sm_100a), so for bf16 features the second GEMM
// reads Y from an fp16 copy (exact for |y| in [6e-5, 65504], which covers CLIP features).
constexpr float kGScaleLog2 = 13.0f;
constexpr float kGScaleInv = 1.0f / 8192.0f;

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__global__ void __launch_bounds__(kThreads, 1)
clip_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmy,
                   const __grid_constant__ CUtensorMap tmy16, const BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = smem_u32(smem);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const uint32_t x_smem = smem_base;
  const uint32_t ring = smem_base + p.kch * kChunkBytes;
  const uint32_t misc = ring + p.stages * kChunkBytes;
  // [0] x_full, [1..6] full, [7..12] empty, [13,14] s_full, [15,16] g_ready, [17] acc_full
  const uint32_t bar_x = misc;
  const uint32_t bar_full = misc + 8;
  const uint32_t bar_empty = misc + 8 + 8 * kMaxStages;
  const uint32_t bar_sfull = misc + 8 + 16 * kMaxStages;
  const uint32_t bar_gready = bar_sfull + 16;
  const uint32_t bar_acc = bar_gready + 16;
  const uint32_t tmem_slot = bar_acc + 8;
  const uint32_t red_slot = tmem_slot + 8;   // 8 floats for the ds reduction
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - smem_base));
  float* red_ptr = reinterpret_cast<float*>(smem + (red_slot - smem_base));

  if (threadIdx.x == 0 && (smem_base & 1023u) != 0) __trap();

  const int row0 = blockIdx.x * kBM;
  const int dsplit = blockIdx.y;
  // feature chunks (of 64 columns) owned by this slab: [ch0, ch0 + nch)
  const int ch0 = dsplit * 4;
  const int nch = min(4, p.kch - ch0);
  const int ntiles = p.tiles_total;

  if (warp == 0 && elect_one()) {
    prefetch_tensormap(&tmx);
    prefetch_tensormap(&tmy);
    prefetch_tensormap(&tmy16);
  }
  if (warp == 1 && elect_one()) {
    mbar_init(bar_x, 1);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_sfull + 8 * b, 1);
      mbar_init(bar_gready + 8 * b, kNumEpiWarps);
    }
    mbar_init(bar_acc, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t tmem_acc = tmem_base;          // columns [0, 256)
  const uint32_t tmem_s = tmem_base + 256;      // two S/G buffers of 128 columns

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      mbar_arrive_expect_tx(bar_x, p.kch * kChunkBytes);
      for (int c = 0; c < p.kch; ++c)
        tma_load_2d(x_smem + c * kChunkBytes, &tmx, bar_x, c * kBK, row0);
      int stage = 0;
      uint32_t phase = 0;
      auto load_chunk = [&](const CUtensorMap* map, int chunk, int tile) {
        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
        mbar_arrive_expect_tx(bar_full + 8 * stage, kChunkBytes);
        tma_load_2d(ring + stage * kChunkBytes, map, bar_full + 8 * stage, chunk * kBK,
                    tile * kBN);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      };
      for (int c = 0; c < p.kch; ++c) load_chunk(&tmy, c, 0);               // GEMM1(0)
      for (int it = 0; it < ntiles; ++it) {
        if (it + 1 < ntiles)
          for (int c = 0; c < p.kch; ++c) load_chunk(&tmy, c, it + 1);      // GEMM1(it+1)
        for (int c = 0; c < nch; ++c) load_chunk(&tmy16, ch0 + c, it);      // GEMM2(it), fp16 Y
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      mbar_wait(bar_x, 0);
      tc_fence_after();
      int stage = 0;
      uint32_t phase = 0;
      auto gemm1 = [&](int it) {   // S[buf] = X_blk . Y_tile^T
        const int buf = it & 1;
        const uint32_t tmem_d = tmem_s + buf * kBN;
        for (int c = 0; c < p.kch; ++c) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t a_addr = x_smem + c * kChunkBytes;
          const uint32_t b_addr = ring + stage * kChunkBytes;
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint64_t da = make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
            const uint64_t db = make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
            mma_ss(tmem_d, da, db, p.idesc_g1, (c | k) != 0);
          }
          tc_commit(bar_empty + 8 * stage);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        tc_commit(bar_sfull + 8 * buf);
      };
      auto gemm2 = [&](int it) {   // acc[:, chunk] += G[buf] . Y_tile[:, chunk]
        const int buf = it & 1;
        mbar_wait(bar_gready + 8 * buf, (it >> 1) & 1);
        tc_fence_after();
        for (int c = 0; c < nch; ++c) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t b_addr = ring + stage * kChunkBytes;
#pragma unroll
          for (int k = 0; k < kBN / 16; ++k) {
            // A: 16 j-columns of G = 8 packed TMEM columns; half h lives at +64*h
            const uint32_t ta = tmem_s + buf * kBN + (k >> 2) * 64 + (k & 3) * 8;
            // B: rows j = 16k .. 16k+15 of the chunk, MN-major (feature axis contiguous)
            const uint64_t db = make_smem_desc_sw128(b_addr + k * 2048, 16384, 1024);
            mma_ts(tmem_acc + c * 64, ta, db, p.idesc_g2, (it | k) != 0);
          }
          tc_commit(bar_empty + 8 * stage);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      };
      gemm1(0);
      for (int it = 0; it < ntiles; ++it) {
        if (it + 1 < ntiles) gemm1(it + 1);
        gemm2(it);
      }
      tc_commit(bar_acc);
    }
  } else if (warp >= kEpiWarp0) {
    // ------------------------------------------------------------ epilogue
    const int q = warp & 3;
    const int half = (warp - kEpiWarp0) >> 2;
    const int row = q * 32 + lane;
    const int64_t grow = (int64_t)row0 + row;
    const bool row_ok = grow < p.n_loc;
    const float s = __ldg(p.logit_scale);
    const float c2 = s * kLog2e;
    const int64_t label = p.label_offset + grow;
    // exponents are biased by +13: every weight below carries the factor 2^13
    const float a2 = (row_ok ? __ldg(p.lse_a2 + label) : 0.f) - kGScaleLog2;
    const float cb = p.cb;
    const float cd_scaled = p.cd * 8192.0f;
    float ds_acc = 0.f;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;

    for (int it = 0; it < ntiles; ++it) {
      const int buf = it & 1;
      const int64_t col0 = (int64_t)it * kBN + half * 64;
      mbar_wait(bar_sfull + 8 * buf, (it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_s + lane_base + buf * kBN + half * 64;
      uint32_t r[2][32];
      tmem_ld_32x32(taddr, r[0]);
      tmem_ld_32x32(taddr + 32, r[1]);
      tmem_ld_wait();

      const bool ragged = col0 + 64 > p.n_all;
      // column of this row's positive pair inside this half tile (-1: not here)
      int want = -1;
      if (label >= col0 && label < col0 + 64 && row_ok) want = (int)(label - col0);
      const float4* pb = reinterpret_cast<const float4*>(p.lse_b2 + col0);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t packed[16];
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          const float4 b4 = __ldg(pb + h * 8 + i4);
          const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
          float g[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = i4 * 4 + e;
            const float v = __uint_as_float(r[h][i]);
            float ea = fast_exp2(fmaf(v, c2, -a2));
            float eb = fast_exp2(fmaf(v, c2, kGScaleLog2 - bb[e]));
            if (ragged && col0 + h * 32 + i >= p.n_all) { ea = 0.f; eb = 0.f; }
            ds_acc = fmaf(ea, v, ds_acc);
            g[e] = fmaf(cb, eb, ea);
            if (h * 32 + i == want) {      // one-hot term, subtracted in fp32 before rounding
              g[e] -= cd_scaled;
              ds_acc = fmaf(-8192.0f, v, ds_acc);
            }
          }
          packed[i4 * 2 + 0] = pack2(g[0], g[1]);
          packed[i4 * 2 + 1] = pack2(g[2], g[3]);
        }
        // G for columns [32h, 32h+32) of this half -> 16 packed TMEM columns
        tmem_st_32x16(tmem_s + lane_base + buf * kBN + half * 64 + h * 16, packed);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_gready + 8 * buf);
    }

    // ---- final: dX slab = coef * s * 2^-13 * acc ----
    mbar_wait(bar_acc, 0);
    tc_fence_after();
    const float coef = __ldg(p.grad_loss) * p.grad_mult / (2.0f * (float)p.n_loc);
    const float cs = coef * s * kGScaleInv;
    const int cols_slab = nch * 64;
    const int cols_half = cols_slab / 2;      // 32, 64, 96 or 128
    for (int cc = 0; cc < cols_half; cc += 32) {
      const int col = half * cols_half + cc;
      uint32_t v[32];
      tmem_ld_32x32(tmem_acc + lane_base + col, v);
      tmem_ld_wait();
      if (row_ok) {
        const int64_t d0 = (int64_t)ch0 * 64 + col;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int64_t d = d0 + i;
          if (d < p.dim) {
            const float o = cs * __uint_as_float(v[i]);
            if (p.grad_dtype == LATTE_F32)
              store_out(reinterpret_cast<float*>(p.dx) + grow * p.ld_dx + d, o);
            else if (p.grad_dtype == LATTE_BF16)
              store_out(reinterpret_cast<__nv_bfloat16*>(p.dx) + grow * p.ld_dx + d, o);
            else
              store_out(reinterpret_cast<__half*>(p.dx) + grow * p.ld_dx + d, o);
          }
        }
      }
    }
    // ---- d loss / d s partial (slab 0 only; every slab computed the same S) ----
    if (dsplit == 0) {
      float v = row_ok ? ds_acc * kGScaleInv : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) red_ptr[warp - kEpiWarp0] = v;
      named_bar_sync(1, kNumEpiWarps * 32);
      if (warp == kEpiWarp0 && lane == 0) {
        float tot = 0.f;
        for (int w = 0; w < kNumEpiWarps; ++w) tot += red_ptr[w];
        p.ds_partial[blockIdx.x] = tot;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// =========================================================================== host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
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

// [rows, dim] 16-bit row-major matrix -> boxes of [128 rows x 64 cols], 128B swizzle,
// out-of-bounds elements read as zero (ragged rows / ragged feature tail).
int make_map(CUtensorMap* map, const void* base, int dtype, int64_t rows, int64_t dim, int64_t ld) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return LATTE_ERR_CUDA;
  cuuint64_t gdim[2] = {(cuuint64_t)dim, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)kBM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, dtype == LATTE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                           : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                  2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? LATTE_OK : LATTE_ERR_CUDA;
}

int smem_bytes_for(int kch, int stages) { return (kch + stages) * kChunkBytes + kMiscBytes; }

}  // namespace

bool clip_tc_supported(int dtype, int64_t dim, int64_t ldx, int64_t ldy, const void* x,
                       const void* y) {
  if (dtype != LATTE_BF16 && dtype != LATTE_F16) return false;
  const int64_t kch = (dim + kBK - 1) / kBK;
  if (kch < 1 || kch > kMaxChunks - 2) return false;      // dim <= 768
  if ((ldx % 8) != 0 || (ldy % 8) != 0) return false;      // TMA: 16-byte row pitch
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 15)) return false;
  return true;
}

int clip_tc_nparts(int64_t n_loc, int64_t n_all, int sm_count) {
  // Column splits so that row_blocks * splits fills whole waves of one CTA per SM while
  // every CTA still sweeps enough tiles to amortise loading its X block.
  const int64_t rb = (n_loc + kBM - 1) / kBM;
  const int64_t tiles = (n_all + kBN - 1) / kBN;
  int best = 1;
  double best_eff = 0.0;
  for (int s = 1; s <= 8; ++s) {
    if (s > 1 && tiles / s < 16) break;
    const int64_t ctas = rb * s;
    const int64_t waves = (ctas + sm_count - 1) / sm_count;
    const double eff = (double)ctas / (double)(waves * sm_count);
    if (eff > best_eff + 1e-9) { best_eff = eff; best = s; }
  }
  return 2 * best;
}

int clip_tc_ds_count(int64_t n_loc, int64_t dim) {
  (void)dim;
  return (int)((n_loc + kBM - 1) / kBM);
}

int clip_fwd_rows_tc(const ClipFwdArgs& a, cudaStream_t stream) {
  if (!clip_tc_supported(a.dtype, a.dim, a.ldx, a.ldy, a.x, a.y)) return LATTE_ERR_UNSUPPORTED;
  CUtensorMap tmx, tmy;
  int rc = make_map(&tmx, a.x, a.dtype, a.n_loc, a.dim, a.ldx);
  if (rc) return rc;
  rc = make_map(&tmy, a.y, a.dtype, a.n_all, a.dim, a.ldy);
  if (rc) return rc;
  FwdParams p;
  p.n_loc = a.n_loc; p.n_all = a.n_all; p.dim = a.dim;
  p.label_offset = a.label_offset;
  p.logit_scale = a.logit_scale;
  p.part_max = a.part_max; p.part_sum = a.part_sum; p.diag = a.diag;
  p.kch = (int)((a.dim + kBK - 1) / kBK);
  p.stages = kMaxChunks - p.kch < kMaxStages ? kMaxChunks - p.kch : kMaxStages;
  p.tiles_total = (int)((a.n_all + kBN - 1) / kBN);
  p.splits = a.nparts / 2;
  p.idesc = make_idesc_f16(kBM, kBN, a.dtype == LATTE_BF16 ? 1u : 0u, 0, 0);
  p.gate = a.gate;
  const int smem = smem_bytes_for(p.kch, p.stages);
  LATTE_CUDA_OK(cudaFuncSetAttribute(clip_fwd_tc_kernel,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  dim3 grid((unsigned)((a.n_loc + kBM - 1) / kBM), (unsigned)p.splits);
  clip_fwd_tc_kernel<<<grid, kThreads, smem, stream>>>(tmx, tmy, p);
  LATTE_LAUNCH_OK();
  return LATTE_OK;
}

int clip_bwd_rows_tc(const ClipBwdArgs& a, cudaStream_t stream) {
  if (!clip_tc_supported(a.dtype, a.dim, a.ldx, a.ldy, a.x, a.y)) return LATTE_ERR_UNSUPPORTED;
  CUtensorMap tmx, tmy;
  int rc = make_map(&tmx, a.x, a.dtype, a.n_loc, a.dim, a.ldx);
  if (rc) return rc;
  rc = make_map(&tmy, a.y, a.dtype, a.n_all, a.dim, a.ldy);
  if (rc) return rc;
  CUtensorMap tmy16;
  rc = make_map(&tmy16, a.y16, LATTE_F16, a.n_all, a.dim, a.ldy16);
  if (rc) return rc;
  BwdParams p;
  p.n_loc = a.n_loc; p.n_all = a.n_all; p.dim = a.dim;
  p.label_offset = a.label_offset;
  p.logit_scale = a.logit_scale;
  p.lse_a2 = a.lse_a2; p.lse_b2 = a.lse_b2;
  p.grad_loss = a.grad_loss; p.grad_mult = a.grad_mult;
  p.cb = a.cross_terms ? 1.f : 0.f;
  p.cd = a.cross_terms ? 2.f : 1.f;
  p.dx = a.dx; p.grad_dtype = a.grad_dtype; p.ld_dx = a.ld_dx;
  p.ds_partial = a.ds_partial;
  p.kch = (int)((a.dim + kBK - 1) / kBK);
  p.stages = kMaxChunks - p.kch < kMaxStages ? kMaxChunks - p.kch : kMaxStages;
  p.tiles_total = (int)((a.n_all + kBN - 1) / kBN);
  p.is_bf16 = a.dtype == LATTE_BF16;
  const uint32_t fmt = p.is_bf16 ? 1u : 0u;
  p.idesc_g1 = make_idesc_f16(kBM, kBN, fmt, 0, 0);
  p.idesc_g2 = make_idesc_f16(kBM, 64, /*fp16*/ 0u, 0, 1);
  const int smem = smem_bytes_for(p.kch, p.stages);
  LATTE_CUDA_OK(cudaFuncSetAttribute(clip_bwd_tc_kernel,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  dim3 grid((unsigned)((a.n_loc + kBM - 1) / kBM), (unsigned)((p.kch + 3) / 4));
  clip_bwd_tc_kernel<<<grid, kThreads, smem, stream>>>(tmx, tmy, tmy16, p);
  LATTE_LAUNCH_OK();
  return LATTE_OK;
}

}  // namespace latte

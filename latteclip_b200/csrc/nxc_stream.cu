// N x C prototype similarity as ONE HBM-bound launch (up to four stacked jobs, C <= 64 per job).
//
// Per step the reference runs five distinct N x C products (train.py:410-411 pseudo-label argmax,
// compute_text_weights train.py:292-303 on the image-description, group-description and class-name
// text features -- SURVEY 2c K8/K9).  latte_nxc_multi runs them as one persistent launch:
//   * prototypes are split ONCE per matrix into 16-bit operand planes (latte_nxc_split_prototypes,
//     optionally fused with the row normalisation of train.py:384-389): three bf16 planes
//     p = p0 + p1 + p2 for bf16 rows, two fp16 planes p = h0 + 2^-11 h1 for fp16 / fp32 rows;
//   * feature rows are streamed exactly once.  16-bit rows are MMA operands as they are; fp32 rows
//     land raw in shared memory ([128 rows x 64 features]) and sixteen converter warps split them
//     into two fp16 planes in place (x = h0 + 2^-11 h1: the residual after two planes is <= 2^-24 |x|,
//     fp32's own rounding; the low plane is stored times 2^11 to stay clear of fp16's subnormals) in
//     the 128B-swizzled operand layout; tcgen05.mma accumulates the plane pairs (0,0), (0,1), (1,0)
//     in fp32 (TMEM): 3 MMAs per 16 features for fp32 and bf16 rows, 2 for fp16 rows.  The first version
//     of this kernel used three bf16 planes for every input (6 / 5 / 3 MMAs) and was bound by the
//     NUMBER of small MMAs (an M128 x N48 x K16 MMA costs ~4x its FLOP time): 0.53 of HBM for every
//     input type; now 0.65 / 0.62 / 0.57 (fp32 / fp16 / bf16 rows);
//   * four accumulators per tile (the large term x0.p0 spread over thirds of the feature axis + one
//     for the small terms, summed in round-to-nearest fp32 by the epilogue) make the result as
//     accurate as an fp32 FMA loop although the tensor core truncates its accumulator per MMA;
//   * TMEM holds two sets of accumulators, so the argmax / top-2 epilogue of tile t overlaps the loads
//     and MMAs of tile t + 1; CTAs are persistent (grid = SM count) and the 32-row groups of all jobs
//     are dealt to them in equal contiguous ranges (one 128-row TMA box per full tile, 32-row boxes
//     for the partial tiles at range ends, one box for the packed prototype planes of a chunk).
// Algorithmic bytes: N * D * sizeof(x) + outputs; the [N, C] logits never exist in memory.
#include "latte_common.cuh"
#include "tc_ptx.cuh"

namespace latte {

using namespace ptx;

namespace {

constexpr int kRows = 128;                     // rows per tile (MMA M)
constexpr int kGrpRows = 32;                   // rows per work unit = rows of one TMA box
constexpr int kBK = 64;                        // features per chunk (128 bytes of bf16)
constexpr int kMaxCls = 64;                    // classes per job (MMA N), padded to 16
constexpr int kPlaneBytes = kRows * kBK * 2;   // 16 KB: one bf16 plane of one chunk
constexpr int kRawBytes = kRows * kBK * 4;     // 32 KB: one raw fp32 chunk (two 128-byte-row boxes)
constexpr int kPBytes = 3 * kMaxCls * kBK * 2; // 24 KB: three prototype planes of one chunk
constexpr int kMaxJobs = 4;
constexpr int kAccCols = 64;                   // TMEM columns per accumulator
constexpr int kBufCols = 4 * kAccCols;         // per TMEM buffer: 4 accumulators

// converting variant (fp32 rows): stage = raw 32 KB + two fp16 planes 32 KB + prototypes 24 KB; two stages
constexpr int kCvtStages = 2;
constexpr int kCvtXPlanes = 2;                 // fp32 = h0 + 2^-11 h1 with two fp16 planes (see below)
constexpr int kCvtStageBytes = kRawBytes + kCvtXPlanes * kPlaneBytes + kPBytes;        // 88 KB
constexpr float kLoScale = 2048.f;             // the low plane is stored times 2^11 (fp16 subnormals)
constexpr int kCvtSmem = kCvtStages * kCvtStageBytes + 1024;
constexpr int kCvtWarps = 16;                  // converter warps: 4 per scheduler hide the cvt -> sub chains
constexpr int kCvtThreads = (6 + kCvtWarps) * 32;   // TMA, MMA, 4 epilogue warps, converters
// direct variant (bf16 rows): stage = one plane 16 KB + prototypes 24 KB; five stages
constexpr int kDirStages = 5;
constexpr int kDirStageBytes = kPlaneBytes + kPBytes;                         // 40 KB
constexpr int kDirSmem = kDirStages * kDirStageBytes + 1024;
constexpr int kDirThreads = 6 * 32;

struct StreamJob {
  int64_t n;                 // rows
  int groups;                // ceil(n / 32): the work unit is one 32-row group (one TMA box)
  int group0;                // first global group index of this job
  int kch;                   // ceil(dim / 64)
  int cp;                    // classes padded to a multiple of 16 (<= 64)
  int num_classes;
  int x_planes;              // operand planes of a feature row: 1 (bf16, fp16: the row itself), 2 (fp32)
  int p_planes;              // prototype planes: 3 bf16 planes (bf16 rows) or 2 fp16 planes (fp16 / fp32 rows)
  int ab_fmt;                // MMA operand format: 1 bf16, 0 fp16
  float scale;
  int64_t* argmax_out; float* margin_out; float* top1_out;
};
struct StreamParams {
  StreamJob job[kMaxJobs];
  int njobs;
  int total_groups;
};
struct StreamMaps {
  CUtensorMap x[kMaxJobs];   // raw rows, 32-row boxes (fp32: 32 features wide; 16-bit: 64 wide): partial tiles
  CUtensorMap xf[kMaxJobs];  // the same rows, 128-row boxes: full tiles (one TMA instead of four -- the
                             // issue rate of the single producer thread bounds the 16-bit variant)
  CUtensorMap p[kMaxJobs];   // prototype planes of one operand format, box 64 x (planes * cp): one TMA per chunk
};

// Row tiles of one CTA.  The 32-row groups of all jobs form one flattened space that is dealt to the
// CTAs in equal contiguous ranges (the persistent grid then finishes together: 128-row blocks dealt
// round-robin left the last wave 73 % full at 32768 rows); a tile is up to four consecutive groups
// of ONE job (MMA M = 128; rows past `ng` groups hold stale data and are masked by the epilogue).
struct TileIter {
  int g, g_end, j;           // next group, end of this CTA's range, job cursor
  int job, ng;               // current tile: job, groups (1..4)
  int32_t row0;              //               first row inside the job
};
__device__ __forceinline__ TileIter tile_iter(const StreamParams& p) {
  TileIter t;
  t.g = (int)((int64_t)blockIdx.x * p.total_groups / gridDim.x);
  t.g_end = (int)((int64_t)(blockIdx.x + 1) * p.total_groups / gridDim.x);
  t.j = 0; t.job = 0; t.ng = 0; t.row0 = 0;
  return t;
}
__device__ __forceinline__ bool next_tile(const StreamParams& p, TileIter& t) {
  if (t.g >= t.g_end) return false;
  while (t.j + 1 < p.njobs && t.g >= p.job[t.j + 1].group0) ++t.j;
  const int job_end = p.job[t.j].group0 + p.job[t.j].groups;
  t.job = t.j;
  t.row0 = (t.g - p.job[t.j].group0) * kGrpRows;
  t.ng = min(kRows / kGrpRows, min(t.g_end, job_end) - t.g);
  t.g += t.ng;
  return true;
}

template <bool kConvert>
__global__ void __launch_bounds__(kConvert ? kCvtThreads : kDirThreads, 1)
nxc_stream_kernel(const __grid_constant__ StreamMaps maps, const StreamParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = smem_u32(smem);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int kStages = kConvert ? kCvtStages : kDirStages;
  constexpr int kStageBytes = kConvert ? kCvtStageBytes : kDirStageBytes;
  constexpr int kOffPlanes = kConvert ? kRawBytes : 0;                     // inside a stage
  constexpr int kOffP = kOffPlanes + (kConvert ? kCvtXPlanes : 1) * kPlaneBytes;
  const uint32_t misc = smem_base + kStages * kStageBytes;
  const uint32_t bar_raw_full = misc;                     // [kStages] raw tile landed (kConvert)
  const uint32_t bar_raw_empty = misc + 8 * kStages;      // [kStages] converters done with the raw tile
  const uint32_t bar_op_full = misc + 16 * kStages;       // [kStages] x planes ready (converters / TMA)
  const uint32_t bar_p_full = misc + 24 * kStages;        // [kStages] prototype planes landed (kConvert)
  const uint32_t bar_op_empty = misc + 32 * kStages;      // [kStages] MMAs of the stage complete
  const uint32_t bar_tfull = misc + 40 * kStages;         // [2]
  const uint32_t bar_tempty = bar_tfull + 16;             // [2]
  const uint32_t tmem_slot = bar_tempty + 16;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - smem_base));
  if (threadIdx.x == 0 && (smem_base & 1023u) != 0) __trap();

  if (warp == 0 && elect_one()) {
    for (int j = 0; j < p.njobs; ++j) {
      prefetch_tensormap(&maps.x[j]);
      prefetch_tensormap(&maps.xf[j]);
      prefetch_tensormap(&maps.p[j]);
    }
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_raw_full + 8 * s, 1);
      mbar_init(bar_raw_empty + 8 * s, kCvtWarps);
      mbar_init(bar_op_full + 8 * s, kConvert ? kCvtWarps : 1);
      mbar_init(bar_p_full + 8 * s, 1);
      mbar_init(bar_op_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_tfull + 8 * b, 1);
      mbar_init(bar_tempty + 8 * b, 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      const uint64_t stream_pol = policy_evict_first();   // feature rows are read once
      const uint64_t keep_pol = policy_evict_last();      // prototype planes: re-read by every block
      int stage = 0;
      uint32_t phase = 0;
      constexpr uint32_t kGrpBytes = kGrpRows * 128;        // one box: 32 rows of 128 bytes
      for (TileIter t = tile_iter(p); next_tile(p, t);) {
        const int j = t.job;
        const StreamJob& jb = p.job[j];
        const uint32_t p_tx = (uint32_t)jb.p_planes * (uint32_t)jb.cp * (kBK * 2);
        for (int c = 0; c < jb.kch; ++c) {
          const uint32_t st = smem_base + stage * kStageBytes;
          if (kConvert) {
            mbar_wait(bar_raw_empty + 8 * stage, phase ^ 1);
            const uint32_t rf = bar_raw_full + 8 * stage;
            mbar_arrive_expect_tx(rf, 2u * (uint32_t)t.ng * kGrpBytes);
            if (t.ng == kRows / kGrpRows) {
              tma_load_2d_hint(st, &maps.xf[j], rf, c * kBK, t.row0, stream_pol);
              tma_load_2d_hint(st + kRawBytes / 2, &maps.xf[j], rf, c * kBK + 32, t.row0, stream_pol);
            } else {
              for (int g = 0; g < t.ng; ++g) {
                tma_load_2d_hint(st + g * kGrpBytes, &maps.x[j], rf, c * kBK, t.row0 + g * kGrpRows, stream_pol);
                tma_load_2d_hint(st + kRawBytes / 2 + g * kGrpBytes, &maps.x[j], rf, c * kBK + 32,
                                 t.row0 + g * kGrpRows, stream_pol);
              }
            }
            mbar_wait(bar_op_empty + 8 * stage, phase ^ 1);
            const uint32_t pf = bar_p_full + 8 * stage;
            mbar_arrive_expect_tx(pf, p_tx);
            tma_load_2d_hint(st + kOffP, &maps.p[j], pf, c * kBK, 0, keep_pol);
          } else {
            mbar_wait(bar_op_empty + 8 * stage, phase ^ 1);
            const uint32_t of = bar_op_full + 8 * stage;
            mbar_arrive_expect_tx(of, (uint32_t)t.ng * kGrpBytes + p_tx);
            if (t.ng == kRows / kGrpRows) {
              tma_load_2d_hint(st, &maps.xf[j], of, c * kBK, t.row0, stream_pol);
            } else {
              for (int g = 0; g < t.ng; ++g)
                tma_load_2d_hint(st + g * kGrpBytes, &maps.x[j], of, c * kBK, t.row0 + g * kGrpRows, stream_pol);
            }
            tma_load_2d_hint(st + kOffP, &maps.p[j], of, c * kBK, 0, keep_pol);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      // plane pairs (a, b), small terms first.  bf16 rows (x = x0, p = p0 + p1 + p2 in bf16):
      // (0,2) (0,1) (0,0).  fp16 planes (16-bit halves of an fp32 value: v = h0 + 2^-11 h1 with
      // |v - h0 - 2^-11 h1| <= 2^-24 |v|; the low plane is stored times 2^11): fp16 rows (0,1) (0,0),
      // fp32 rows (0,1) (1,0) (0,0) -- the (1,1) term is below 2^-22 |x||p| and dropped.  The large
      // term (0,0) goes to accumulator `third`, all others to accumulator 3 (scaled by 2^-11 in the
      // epilogue when the planes are fp16).
      const int pa[4] = {0, 0, 1, 0};
      const int pb[4] = {2, 1, 0, 0};
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (TileIter t = tile_iter(p); next_tile(p, t); ++it) {
        const StreamJob& jb = p.job[t.job];
        const int buf = it & 1;
        const uint32_t idesc = make_idesc_f16(kRows, jb.cp, (uint32_t)jb.ab_fmt, 0, 0);
        mbar_wait(bar_tempty + 8 * buf, ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        uint32_t used = 0;                 // bit a: accumulator a already holds a partial sum
        for (int c = 0; c < jb.kch; ++c) {
          mbar_wait(bar_op_full + 8 * stage, phase);
          if (kConvert) mbar_wait(bar_p_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t st = smem_base + stage * kStageBytes;
          const int third = (c * 3) / jb.kch;            // accumulator of the x0.p0 term
          for (int q = 0; q < 4; ++q) {
            if (pa[q] >= jb.x_planes || pb[q] >= jb.p_planes) continue;
            const int acc_id = (pa[q] | pb[q]) == 0 ? third : 3;
            const uint32_t tmem_d = tmem_base + buf * kBufCols + acc_id * kAccCols;
            const uint64_t da0 = make_smem_desc_sw128(st + kOffPlanes + pa[q] * kPlaneBytes, 16, 1024);
            const uint64_t db0 = make_smem_desc_sw128(st + kOffP + pb[q] * (jb.cp * kBK * 2), 16, 1024);
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              mma_ss(tmem_d, da0 + (uint64_t)(k * 2), db0 + (uint64_t)(k * 2), idesc, (used >> acc_id) & 1u);
              used |= 1u << acc_id;
            }
          }
          tc_commit(bar_op_empty + 8 * stage);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        tc_commit(bar_tfull + 8 * buf);
      }
    }
  } else if (warp < 6) {
    // ------------------------------------------------------------ epilogue (warps 2-5)
    const int q = warp & 3;                               // TMEM lane quarter of this warp
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    int it = 0;
    for (TileIter t = tile_iter(p); next_tile(p, t); ++it) {
      const StreamJob& jb = p.job[t.job];
      const int buf = it & 1;
      const int64_t gr = (int64_t)t.row0 + q * 32 + lane;
      uint32_t used_mask = 8u;
      for (int c = 0; c < jb.kch; ++c) used_mask |= 1u << ((c * 3) / jb.kch);
      float v1 = -INFINITY, v2 = -INFINITY;
      int i1 = 0;
      const float lo_scale = jb.ab_fmt == 0 ? 1.f / kLoScale : 1.f;   // fp16 planes: low plane stored times 2^11
      mbar_wait(bar_tfull + 8 * buf, (it >> 1) & 1);
      tc_fence_after();
      for (int cc = 0; cc < jb.cp; cc += 16) {
        uint32_t r[4][16];
#pragma unroll
        for (int a = 0; a < 4; ++a)
          if ((used_mask >> a) & 1u)
            tmem_ld_32x16(tmem_base + lane_base + buf * kBufCols + a * kAccCols + cc, r[a]);
        tmem_ld_wait();
        const int m = min(16, jb.num_classes - cc);
#pragma unroll
        for (int jx = 0; jx < 16; ++jx) {
          float s = 0.f;                       // small terms first, round-to-nearest adds
          if ((used_mask >> 3) & 1u) s = __uint_as_float(r[3][jx]) * lo_scale;
          if ((used_mask >> 2) & 1u) s += __uint_as_float(r[2][jx]);
          if ((used_mask >> 1) & 1u) s += __uint_as_float(r[1][jx]);
          if (used_mask & 1u) s += __uint_as_float(r[0][jx]);
          if (jx < m) {
            if (s > v1) { v2 = v1; v1 = s; i1 = cc + jx; }
            else if (s > v2) { v2 = s; }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
      if (q < t.ng && gr < jb.n) {
        if (jb.argmax_out) jb.argmax_out[gr] = i1;
        if (jb.margin_out) jb.margin_out[gr] = v1 - v2;
        if (jb.top1_out) jb.top1_out[gr] = jb.scale * v1;
      }
    }
  } else if (kConvert) {
    // ------------------------------------------------------------ converters (warps 6-21)
    // thread -> (row r of the block, quarter h of the chunk): 16 consecutive features
    const int t = threadIdx.x - 6 * 32;
    const int r = t & 127, h = t >> 7;
    const uint32_t sw = (uint32_t)(r & 7);
    int stage = 0;
    uint32_t phase = 0;
    for (TileIter t = tile_iter(p); next_tile(p, t);) {
      const StreamJob& jb = p.job[t.job];
      for (int c = 0; c < jb.kch; ++c) {
        const uint32_t st = smem_base + stage * kStageBytes;
        mbar_wait(bar_raw_full + 8 * stage, phase);
        const bool live = (r >> 5) < t.ng;       // warp-uniform: this warp's 32-row group is part of the tile
        float v[16];
        if (!live) {
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] = 0.f;
        } else {
          // box h / 2 holds 32 features: row r = 128 bytes, 16-byte chunk j at position j ^ (r & 7);
          // this thread's 16 floats are chunks 4 (h & 1) .. 4 (h & 1) + 3
          const uint32_t row_addr = st + (h >> 1) * (kRawBytes / 2) + r * 128;
#pragma unroll
          for (int jx = 0; jx < 4; ++jx) {
            uint32_t a, b, cw, d;
            ld_shared_v4(row_addr + (((uint32_t)(4 * (h & 1) + jx) ^ sw) << 4), a, b, cw, d);
            v[4 * jx] = __uint_as_float(a); v[4 * jx + 1] = __uint_as_float(b);
            v[4 * jx + 2] = __uint_as_float(cw); v[4 * jx + 3] = __uint_as_float(d);
          }
        }
        // the raw tile is in registers: hand the buffer back before converting
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_raw_empty + 8 * stage);
        mbar_wait(bar_op_empty + 8 * stage, phase ^ 1);     // MMAs that read these planes are done
        const uint32_t prow = st + kOffPlanes + r * 128;
        if (live) {
          // h0 = fp16(v), h1 = fp16(2^11 (v - h0)): the residual is exact in fp32 and the scaling keeps
          // the low plane out of fp16's subnormal range (|v| < 65504 is assumed: unit-norm features)
#pragma unroll
          for (int pl = 0; pl < kCvtXPlanes; ++pl) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              uint32_t w4[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const __half2 h2 = __floats2half2_rn(v[8 * g + 2 * e], v[8 * g + 2 * e + 1]);
                w4[e] = *reinterpret_cast<const uint32_t*>(&h2);
                if (pl == 0) {
                  const float2 f = __half22float2(h2);
                  v[8 * g + 2 * e] = (v[8 * g + 2 * e] - f.x) * kLoScale;
                  v[8 * g + 2 * e + 1] = (v[8 * g + 2 * e + 1] - f.y) * kLoScale;
                }
              }
              st_shared_v4(prow + pl * kPlaneBytes + (((uint32_t)(2 * h + g) ^ sw) << 4), w4[0], w4[1], w4[2], w4[3]);
            }
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_op_full + 8 * stage);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// prototypes [C, D] fp32 -> bf16 planes [3][cp][dim_pad] followed by fp16 planes [2][cp][dim_pad]
// (high half, and 2^11 times the residual; rows >= C and columns >= D are zero), optionally
// L2-normalising each row first (train.py:384-389: F.normalize(stack(bank), dim=1)).  bf16 feature
// rows multiply the bf16 planes, fp16 / fp32 rows the fp16 planes.
__global__ void __launch_bounds__(256)
nxc_split_protos_kernel(const float* protos, int64_t ld, int num_classes, int dim, int cp, int dim_pad,
                        int normalize, __nv_bfloat16* out, float* normalized_out, int64_t ld_norm) {
  __shared__ float red[8];
  const int c = blockIdx.x;
  float inv = 1.0f;
  if (c < num_classes && normalize) {
    float ss = 0.f;
    for (int d = threadIdx.x; d < dim; d += 256) { const float v = protos[(int64_t)c * ld + d]; ss = fmaf(v, v, ss); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
    for (int w = 0; w < 8; ++w) tot += red[w];
    inv = 1.0f / fmaxf(sqrtf(tot), 1e-12f);
  }
  for (int d = threadIdx.x; d < dim_pad; d += 256) {
    float v = (c < num_classes && d < dim) ? protos[(int64_t)c * ld + d] : 0.f;
    if (normalize) v *= inv;            // the same x * (1 / max(norm, eps)) as latte_normalize_rows
    if (normalized_out && c < num_classes && d < dim) normalized_out[(int64_t)c * ld_norm + d] = v;
    {
      __half* out16 = reinterpret_cast<__half*>(out + (int64_t)3 * cp * dim_pad);
      const __half h0 = __float2half_rn(v);
      out16[(int64_t)c * dim_pad + d] = h0;
      out16[((int64_t)cp + c) * dim_pad + d] = __float2half_rn((v - __half2float(h0)) * kLoScale);
    }
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) {
      const __nv_bfloat16 b = __float2bfloat16_rn(v);
      out[((int64_t)pl * cp + c) * dim_pad + d] = b;
      v -= __bfloat162float(b);
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn stream_encode_fn() {
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

int stream_make_map(CUtensorMap* map, const void* base, CUtensorMapDataType dt, int esize, int64_t rows,
                    int64_t cols, int64_t ld, int box_cols, int box_rows) {
  EncodeTiledFn fn = stream_encode_fn();
  if (!fn) return LATTE_ERR_CUDA;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * esize};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? LATTE_OK : LATTE_ERR_CUDA;
}

}  // namespace
}  // namespace latte

using namespace latte;

extern "C" int latte_nxc_planes_bytes(int64_t num_classes, int64_t dim, size_t* bytes) {
  LATTE_CHECK_ARG(bytes && num_classes > 0 && dim > 0);
  if (num_classes > kMaxCls) return LATTE_ERR_UNSUPPORTED;
  const int64_t cp = (num_classes + 15) / 16 * 16;
  const int64_t dim_pad = (dim + kBK - 1) / kBK * kBK;
  *bytes = (size_t)(3 + 2) * cp * dim_pad * sizeof(__nv_bfloat16);      // 3 bf16 planes + 2 fp16 planes
  return LATTE_OK;
}

extern "C" int latte_nxc_split_prototypes(const float* protos, int64_t ld, int64_t num_classes, int64_t dim,
                                          int normalize, void* planes, float* normalized_out,
                                          int64_t ld_norm, void* stream) {
  LATTE_CHECK_ARG(protos && planes && num_classes > 0 && dim > 0 && ld >= dim);
  LATTE_CHECK_ARG(!normalized_out || ld_norm >= dim);
  if (num_classes > kMaxCls) return LATTE_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(planes) & 127) != 0) return LATTE_ERR_BAD_ARG;
  const int cp = (int)((num_classes + 15) / 16 * 16);
  const int dim_pad = (int)((dim + kBK - 1) / kBK * kBK);
  nxc_split_protos_kernel<<<cp, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      protos, ld, (int)num_classes, (int)dim, cp, dim_pad, normalize, static_cast<__nv_bfloat16*>(planes),
      normalized_out, ld_norm);
  LATTE_LAUNCH_OK();
  return LATTE_OK;
}

extern "C" int latte_nxc_multi(const latte_nxc_job_t* jobs, int njobs, void* stream) {
  LATTE_CHECK_ARG(jobs && njobs >= 1 && njobs <= kMaxJobs);
  StreamParams p = {};
  StreamMaps maps;
  p.njobs = njobs;
  int groups = 0;
  bool convert = false, direct = false;
  for (int j = 0; j < njobs; ++j) {
    const latte_nxc_job_t& in = jobs[j];
    LATTE_CHECK_ARG(in.x && in.planes && in.n > 0 && in.dim > 0 && in.num_classes > 0 && in.ldx >= in.dim);
    LATTE_CHECK_ARG(in.x_dtype >= LATTE_F32 && in.x_dtype <= LATTE_F16);
    if (in.num_classes > kMaxCls) return LATTE_ERR_UNSUPPORTED;
    const int esize = in.x_dtype == LATTE_F32 ? 4 : 2;
    if (((in.ldx * esize) % 16) != 0 || (reinterpret_cast<uintptr_t>(in.x) & 15) != 0 ||
        (reinterpret_cast<uintptr_t>(in.planes) & 127) != 0 || in.n >= (1ll << 31))
      return LATTE_ERR_UNSUPPORTED;
    StreamJob& jb = p.job[j];
    jb.n = in.n;
    jb.groups = (int)((in.n + kGrpRows - 1) / kGrpRows);
    jb.group0 = groups;
    groups += jb.groups;
    jb.kch = (int)((in.dim + kBK - 1) / kBK);
    jb.cp = (int)((in.num_classes + 15) / 16 * 16);
    jb.num_classes = (int)in.num_classes;
    jb.x_planes = in.x_dtype == LATTE_F32 ? kCvtXPlanes : 1;
    jb.p_planes = in.x_dtype == LATTE_BF16 ? 3 : 2;
    jb.ab_fmt = in.x_dtype == LATTE_BF16 ? 1 : 0;
    jb.scale = in.scale;
    jb.argmax_out = in.argmax_out; jb.margin_out = in.margin_out; jb.top1_out = in.top1_out;
    (in.x_dtype == LATTE_F32 ? convert : direct) = true;
    const int64_t dim_pad = (int64_t)jb.kch * kBK;
    int rc = LATTE_OK;
    for (int full = 0; full < 2 && !rc; ++full) {
      CUtensorMap* m = full ? &maps.xf[j] : &maps.x[j];
      const int box_rows = full ? kRows : kGrpRows;
      if (in.x_dtype == LATTE_F32)
        rc = stream_make_map(m, in.x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, in.n, in.dim, in.ldx, 32, box_rows);
      else
        rc = stream_make_map(m, in.x, in.x_dtype == LATTE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                                : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                             2, in.n, in.dim, in.ldx, kBK, box_rows);
    }
    if (rc) return rc;
    if (in.x_dtype == LATTE_BF16)
      rc = stream_make_map(&maps.p[j], in.planes, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 3 * (int64_t)jb.cp,
                           dim_pad, dim_pad, kBK, 3 * jb.cp);
    else
      rc = stream_make_map(&maps.p[j], static_cast<const char*>(in.planes) + (size_t)3 * jb.cp * dim_pad * 2,
                           CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, 2 * (int64_t)jb.cp, dim_pad, dim_pad, kBK, 2 * jb.cp);
    if (rc) return rc;
  }
  // one launch handles one operand class: all jobs 16-bit (rows are MMA operands as they are) or all
  // fp32 (converting)
  if (convert && direct) return LATTE_ERR_UNSUPPORTED;
  for (int j = njobs; j < kMaxJobs; ++j) { maps.x[j] = maps.x[0]; maps.xf[j] = maps.xf[0]; maps.p[j] = maps.p[0]; }
  p.total_groups = groups;
  int grid = device_sm_count();
  if (grid > groups) grid = groups;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (convert) {
    LATTE_CUDA_OK(cudaFuncSetAttribute(nxc_stream_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       kCvtSmem));
    nxc_stream_kernel<true><<<grid, kCvtThreads, kCvtSmem, st>>>(maps, p);
  } else {
    LATTE_CUDA_OK(cudaFuncSetAttribute(nxc_stream_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       kDirSmem));
    nxc_stream_kernel<false><<<grid, kDirThreads, kDirSmem, st>>>(maps, p);
  }
  LATTE_LAUNCH_OK();
  return LATTE_OK;
}

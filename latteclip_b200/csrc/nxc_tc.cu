// N x C prototype similarity on tcgen05 tensor cores with fp32-grade products.
//
// Replaces the fp32 SIMT tile kernel of nxc.cu for large problems (train.py:410-411,
// compute_text_weights train.py:292-303, zero_shot.py:14-20,40).  Pseudo-labels must be exact
// wherever the reference's fp32 GEMM has no tie and margins are differences of near-equal
// dot products, so a plain 16-bit MMA is not accurate enough.  Operands are split into 16-bit
// planes and the plane products are accumulated in fp32 (TMEM):
//   bf16 rows: x is its own plane, p = b0 + b1 + b2 in bf16 (8 + 8 + 8 significand bits, exact):
//              x.p = x.b0 + x.b1 + x.b2                                         (3 MMAs)
//   fp16 / fp32 rows: fp16 planes v = h0 + 2^-11 h1 (11 + 11 bits + sign trick: the residual after
//              two planes is <= 2^-24 |v|; the low plane is stored times 2^11 to stay clear of
//              fp16's subnormals; |v| < 65504): x.p = x0.p0 + x0.p1 + x1.p0      (3 MMAs for fp32
//              rows, 2 for fp16 rows which are their own plane; the x1.p1 term is < 2^-22 |x||p|).
// Round 1 used three bf16 planes for every input (6 MMAs for fp32 rows, 5 for fp16): this kernel is
// tensor-bound from C ~ 400, so half the MMAs is half the time.
// The tensor core truncates its fp32 accumulator on every MMA (measured: a bias of ~-3e-7 on
// unit-norm dots after 192 accumulations), so the large term x0.p0 is spread over three TMEM
// accumulators (thirds of the feature axis), the small terms go to a fourth, and the epilogue
// adds the four in round-to-nearest fp32: the result is as accurate as an fp32 FMA loop.
//
//   nxc_split_planes_kernel : (optionally gathered) rows -> 16-bit planes [P][rows][dim_pad]
//   nxc_tc_kernel           : 128 rows per CTA, class tiles of <= 128, K chunks of 64 through a
//                             2-stage TMA ring, accumulators double-buffered in TMEM; four
//                             epilogue warps keep the running argmax / top-2 / top-k per row.
// The [N, C] logits are never written to HBM.  TMEM: 4 accumulators x 128 columns.
#include "latte_common.cuh"
#include "tc_ptx.cuh"

namespace latte {

using namespace ptx;

namespace {

constexpr int kRows = 128;                 // rows per CTA (MMA M)
constexpr int kCls = 128;                  // classes per tile (MMA N <= 128)
constexpr int kBK = 64;                    // features per chunk (128 bytes of bf16)
constexpr int kPlaneBytes = kRows * kBK * 2;        // 16 KB: one plane of one chunk
constexpr int kStages = 2;
constexpr int kStageBytes = 6 * kPlaneBytes;        // 3 x planes + 3 p planes
constexpr int kNxcThreads = 192;           // warp 0 TMA, warp 1 MMA + TMEM alloc, warps 2-5 epilogue
constexpr int kNxcSmem = kStages * kStageBytes + 1024;
constexpr int kMaxTopK = 16;

// ---- split into 16-bit planes -------------------------------------------------------------
__device__ __forceinline__ float load_as_float(const void* base, int64_t idx, int dtype) {
  if (dtype == LATTE_F32) return __ldg(reinterpret_cast<const float*>(base) + idx);
  if (dtype == LATTE_BF16)
    return __bfloat162float(__ldg(reinterpret_cast<const __nv_bfloat16*>(base) + idx));
  return __half2float(__ldg(reinterpret_cast<const __half*>(base) + idx));
}

// eight consecutive features per thread: 16-byte stores into every plane
__global__ void __launch_bounds__(256)
nxc_split_planes_kernel(const void* src, int64_t ld, int dtype, const int64_t* row_index,
                        int64_t rows, int64_t dim, int64_t dim_pad, int planes, int fp16_planes,
                        __nv_bfloat16* out, int vec_ok) {
  const int64_t per_row = dim_pad / 8;
  const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= rows * per_row) return;
  const int64_t r = idx / per_row, d0 = (idx % per_row) * 8;
  const int64_t sr = row_index ? row_index[r] : r;
  float v[8];
  if (vec_ok && dtype == LATTE_F32 && d0 + 8 <= dim) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(src) + sr * ld + d0));
    const float4 c = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(src) + sr * ld + d0) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = d0 + e < dim ? load_as_float(src, sr * ld + d0 + e, dtype) : 0.f;
  }
  if (fp16_planes) {
    // h0 = fp16(v), h1 = fp16(2^11 (v - h0)): the residual is exact in fp32
    for (int p = 0; p < planes; ++p) {
      uint4 o;
      __half2* h = reinterpret_cast<__half2*>(&o);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const __half2 b = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
        h[e] = b;
        const float2 f = __half22float2(b);
        v[2 * e] = (v[2 * e] - f.x) * 2048.f;
        v[2 * e + 1] = (v[2 * e + 1] - f.y) * 2048.f;
      }
      *reinterpret_cast<uint4*>(out + ((int64_t)p * rows + r) * dim_pad + d0) = o;
    }
    return;
  }
  for (int p = 0; p < planes; ++p) {
    uint4 o;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const __nv_bfloat162 b = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
      h[e] = b;
      v[2 * e] -= __bfloat162float(b.x);            // exact in fp32
      v[2 * e + 1] -= __bfloat162float(b.y);
    }
    *reinterpret_cast<uint4*>(out + ((int64_t)p * rows + r) * dim_pad + d0) = o;
  }
}

// ---- the GEMM + row-reduction kernel ----------------------------------------------------
struct NxcTcParams {
  int64_t n, num_classes;
  int kch;                 // ceil(dim / 64)
  int x_planes;            // 1 (bf16 / fp16 rows) or 2 (fp32 rows)
  int p_planes;            // 3 bf16 planes (bf16 rows) or 2 fp16 planes
  int ab_fmt;              // MMA operand format: 1 bf16, 0 fp16
  int passes;              // ceil(C / 128)
  float scale;
  int64_t* argmax_out; float* margin_out; float* top1_out;
  int k; int64_t* topk_idx; float* topk_val;
};

template <bool kTopK>
__global__ void __launch_bounds__(kNxcThreads, 1)
nxc_tc_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmp,
              const NxcTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = smem_u32(smem);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t misc = smem_base + kStages * kStageBytes;
  const uint32_t bar_full = misc;                   // [kStages]
  const uint32_t bar_empty = misc + 8 * kStages;    // [kStages]
  const uint32_t bar_tfull = bar_empty + 8 * kStages;   // [2]
  const uint32_t bar_tempty = bar_tfull + 16;           // [2]
  const uint32_t tmem_slot = bar_tempty + 16;
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - smem_base));
  if (threadIdx.x == 0 && (smem_base & 1023u) != 0) __trap();

  const int64_t row0 = (int64_t)blockIdx.x * kRows;
  const int nplanes_x = p.x_planes;
  // accumulators in use: the thirds of the feature axis that own a chunk + the small-term one
  uint32_t used_mask = 8u;
  for (int c = 0; c < p.kch; ++c) used_mask |= 1u << ((c * 3) / p.kch);

  if (warp == 0 && elect_one()) {
    prefetch_tensormap(&tmx);
    prefetch_tensormap(&tmp);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
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
      const uint32_t tx = (uint32_t)(nplanes_x + p.p_planes) * kPlaneBytes;
      int stage = 0;
      uint32_t phase = 0;
      for (int pass = 0; pass < p.passes; ++pass) {
        for (int c = 0; c < p.kch; ++c) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t sa = smem_base + stage * kStageBytes;
          const uint32_t full = bar_full + 8 * stage;
          mbar_arrive_expect_tx(full, tx);
          for (int a = 0; a < nplanes_x; ++a)
            tma_load_2d(sa + a * kPlaneBytes, &tmx, full, c * kBK, (int32_t)(a * p.n + row0));
          for (int b = 0; b < p.p_planes; ++b)
            tma_load_2d(sa + (3 + b) * kPlaneBytes, &tmp, full, c * kBK,
                        (int32_t)(b * p.num_classes + pass * kCls));
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      // plane pairs (a, b), small terms first; (0, 0) goes to the accumulator of the chunk's third,
      // every other pair to accumulator 3 (rescaled by 2^-11 in the epilogue for fp16 planes)
      const int pa[4] = {0, 0, 1, 0};
      const int pb[4] = {2, 1, 0, 0};
      int stage = 0;
      uint32_t phase = 0;
      for (int pass = 0; pass < p.passes; ++pass) {
        const int64_t left = p.num_classes - (int64_t)pass * kCls;
        const int ncols = (int)(left >= kCls ? kCls : (left + 15) / 16 * 16);
        const uint32_t idesc = make_idesc_f16(kRows, ncols, (uint32_t)p.ab_fmt, 0, 0);
        mbar_wait(bar_tempty, (pass & 1) ^ 1);
        tc_fence_after();
        uint32_t used = 0;                 // bit a: accumulator a already holds a partial sum
        for (int c = 0; c < p.kch; ++c) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * kStageBytes;
          const int third = (c * 3) / p.kch;          // accumulator of the x0.p0 term
          for (int q = 0; q < 4; ++q) {
            if (pa[q] >= nplanes_x || pb[q] >= p.p_planes) continue;
            const int acc_id = (pa[q] | pb[q]) == 0 ? third : 3;
            const uint32_t tmem_d = tmem_base + acc_id * kCls;
            const uint64_t da0 = make_smem_desc_sw128(sa + pa[q] * kPlaneBytes, 16, 1024);
            const uint64_t db0 = make_smem_desc_sw128(sa + (3 + pb[q]) * kPlaneBytes, 16, 1024);
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              mma_ss(tmem_d, da0 + (uint64_t)(k * 2), db0 + (uint64_t)(k * 2), idesc,
                     (used >> acc_id) & 1u);
              used |= 1u << acc_id;
            }
          }
          tc_commit(bar_empty + 8 * stage);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        tc_commit(bar_tfull);
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2-5)
    const int q = warp & 3;                               // TMEM lane quarter of this warp
    const int64_t gr = row0 + q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    float v1 = -INFINITY, v2 = -INFINITY;
    int i1 = 0;
    const float lo_scale = p.ab_fmt == 0 ? 1.f / 2048.f : 1.f;   // fp16 planes: the low plane is stored times 2^11
    float tv[kTopK ? kMaxTopK : 1];
    int ti[kTopK ? kMaxTopK : 1];
    if (kTopK) {
#pragma unroll
      for (int j = 0; j < kMaxTopK; ++j) { tv[j] = -INFINITY; ti[j] = 0x7fffffff; }
    }
    for (int pass = 0; pass < p.passes; ++pass) {
      const int64_t c0 = (int64_t)pass * kCls;
      const int lim = (int)min((int64_t)kCls, p.num_classes - c0);
      mbar_wait(bar_tfull, pass & 1);
      tc_fence_after();
      const uint32_t used = used_mask;
      for (int cc = 0; cc < lim; cc += 16) {
        uint32_t r[4][16];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          if ((used >> a) & 1u) tmem_ld_32x16(tmem_base + lane_base + a * kCls + cc, r[a]);
        }
        tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float s = 0.f;                       // small terms first, round-to-nearest adds
          if ((used >> 3) & 1u) s = __uint_as_float(r[3][j]) * lo_scale;
          if ((used >> 2) & 1u) s += __uint_as_float(r[2][j]);
          if ((used >> 1) & 1u) s += __uint_as_float(r[1][j]);
          if (used & 1u) s += __uint_as_float(r[0][j]);
          v[j] = s;
        }
        const int m = min(16, lim - cc);
        if (kTopK) {
          for (int j = 0; j < m; ++j) {
            const float vv = v[j];
            // descending list; equal values keep the lower class index first
            if (vv > tv[p.k - 1]) {
              int pos = p.k - 1;
              while (pos > 0 && vv > tv[pos - 1]) { tv[pos] = tv[pos - 1]; ti[pos] = ti[pos - 1]; --pos; }
              tv[pos] = vv; ti[pos] = (int)(c0 + cc + j);
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (j < m) {
              const float vv = v[j];
              if (vv > v1) { v2 = v1; v1 = vv; i1 = (int)(c0 + cc + j); }
              else if (vv > v2) { v2 = vv; }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty);
    }
    if (gr < p.n) {
      if (kTopK) {
        for (int j = 0; j < p.k; ++j) {
          p.topk_idx[gr * p.k + j] = ti[j];
          p.topk_val[gr * p.k + j] = p.scale * tv[j];
        }
      } else {
        if (p.argmax_out) p.argmax_out[gr] = i1;
        if (p.margin_out) p.margin_out[gr] = v1 - v2;
        if (p.top1_out) p.top1_out[gr] = p.scale * v1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn nxc_encode_fn() {
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

// bf16 [rows, cols] row-major (pitch ld) -> boxes [128 rows x 64 cols], 128B swizzle, OOB = 0
int nxc_make_map(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, bool fp16 = false) {
  EncodeTiledFn fn = nxc_encode_fn();
  if (!fn) return LATTE_ERR_CUDA;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, 128u};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                  const_cast<void*>(base), gdim, gstride,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? LATTE_OK : LATTE_ERR_CUDA;
}

}  // namespace

namespace {
struct NxcWsLayout { size_t x_elems, x_al, p_elems, total; int64_t dim_pad; int x_planes, p_planes; bool direct_x; };
NxcWsLayout nxc_ws_layout(const void* x, int64_t ldx, int x_dtype, bool gathered, int64_t n, int64_t dim,
                          int64_t num_classes) {
  NxcWsLayout w;
  w.dim_pad = (dim + kBK - 1) / kBK * kBK;
  // x == NULL (size query): assume the planes of x have to be materialised unless bf16 rows are used as is
  w.direct_x = x_dtype != LATTE_F32 && !gathered && (ldx % 8) == 0 &&
               (reinterpret_cast<uintptr_t>(x) & 15) == 0;          // 16-bit rows are MMA operands as they are
  w.x_planes = x_dtype == LATTE_F32 ? 2 : 1;
  w.p_planes = x_dtype == LATTE_BF16 ? 3 : 2;
  w.x_elems = w.direct_x ? 0 : (size_t)w.x_planes * (size_t)n * (size_t)w.dim_pad;
  w.p_elems = (size_t)w.p_planes * (size_t)num_classes * (size_t)w.dim_pad;
  w.x_al = (w.x_elems + 127) / 128 * 128;
  w.total = (w.x_al + w.p_elems) * sizeof(__nv_bfloat16);
  return w;
}
}  // namespace

// Scratch of one N x C call: 16-bit planes of the prototypes and (unless 16-bit rows are read in place)
// of x.  Upper bound over pointer alignments of x.
size_t nxc_tc_workspace_bytes(int x_dtype, bool gathered, int64_t n, int64_t dim, int64_t num_classes) {
  // an unaligned bf16 x also needs its plane: query with direct_x = false
  NxcWsLayout w = nxc_ws_layout(reinterpret_cast<const void*>(1), 1, x_dtype, gathered, n, dim, num_classes);
  return w.total + 256;
}

// Returns LATTE_ERR_UNSUPPORTED when the caller should use the SIMT kernel instead.
int nxc_tc_run(const void* x, int64_t ldx, int x_dtype, const int64_t* row_index, int64_t n,
               int64_t dim, const float* protos, int64_t ldp, int64_t num_classes, float scale,
               int64_t* argmax_out, float* margin_out, float* top1_out, int k, int64_t* topk_idx,
               float* topk_val, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  // plane row offsets are 32-bit TMA coordinates
  if (3 * n >= (1ll << 31) || 3 * num_classes >= (1ll << 31)) return LATTE_ERR_UNSUPPORTED;
  const int fp16_planes = x_dtype == LATTE_BF16 ? 0 : 1;
  const NxcWsLayout w = nxc_ws_layout(x, ldx, x_dtype, row_index != nullptr, n, dim, num_classes);
  const int64_t dim_pad = w.dim_pad;
  const bool direct_x = w.direct_x;
  const int x_planes = w.x_planes;
  if (!workspace) return LATTE_ERR_BAD_ARG;
  const uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 255) / 256 * 256;
  if (base - reinterpret_cast<uintptr_t>(workspace) + w.total > workspace_bytes) return LATTE_ERR_WORKSPACE;
  __nv_bfloat16* buf = reinterpret_cast<__nv_bfloat16*>(base);
  __nv_bfloat16* xp = buf;
  __nv_bfloat16* pp = buf + w.x_al;
  int rc = LATTE_OK;
  if (!direct_x) {
    const int64_t work = n * (dim_pad / 8);
    const int vec_ok = (ldx % 4) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
    nxc_split_planes_kernel<<<(unsigned)((work + 255) / 256), 256, 0, st>>>(
        x, ldx, x_dtype, row_index, n, dim, dim_pad, x_planes, fp16_planes, xp, vec_ok);
  }
  {
    const int64_t work = num_classes * (dim_pad / 8);
    const int vec_ok = (ldp % 4) == 0 && (reinterpret_cast<uintptr_t>(protos) & 15) == 0;
    nxc_split_planes_kernel<<<(unsigned)((work + 255) / 256), 256, 0, st>>>(
        protos, ldp, LATTE_F32, nullptr, num_classes, dim, dim_pad, w.p_planes, fp16_planes, pp, vec_ok);
  }
  CUtensorMap tmx, tmpm;
  if (direct_x) rc = nxc_make_map(&tmx, x, n, dim, ldx, fp16_planes != 0);
  else rc = nxc_make_map(&tmx, xp, (int64_t)x_planes * n, dim_pad, dim_pad, fp16_planes != 0);
  if (!rc) rc = nxc_make_map(&tmpm, pp, (int64_t)w.p_planes * num_classes, dim_pad, dim_pad, fp16_planes != 0);
  if (!rc) {
    NxcTcParams p;
    p.n = n; p.num_classes = num_classes;
    p.kch = (int)(dim_pad / kBK);
    p.x_planes = x_planes;
    p.p_planes = w.p_planes;
    p.ab_fmt = fp16_planes ? 0 : 1;
    p.passes = (int)((num_classes + kCls - 1) / kCls);
    p.scale = scale;
    p.argmax_out = argmax_out; p.margin_out = margin_out; p.top1_out = top1_out;
    p.k = k; p.topk_idx = topk_idx; p.topk_val = topk_val;
    const unsigned grid = (unsigned)((n + kRows - 1) / kRows);
    if (topk_idx) {
      if (cudaFuncSetAttribute(nxc_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               kNxcSmem) != cudaSuccess) rc = LATTE_ERR_CUDA;
      else nxc_tc_kernel<true><<<grid, kNxcThreads, kNxcSmem, st>>>(tmx, tmpm, p);
    } else {
      if (cudaFuncSetAttribute(nxc_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               kNxcSmem) != cudaSuccess) rc = LATTE_ERR_CUDA;
      else nxc_tc_kernel<false><<<grid, kNxcThreads, kNxcSmem, st>>>(tmx, tmpm, p);
    }
    if (!rc && cudaGetLastError() != cudaSuccess) rc = LATTE_ERR_CUDA;
  }
  return rc;
}

}  // namespace latte

// HBM-bound prototype kernels: row normalisation, text mixture + EMA (forward and
// backward), memory-bank segment sums and per-class normalise.
// Replaces the ~20 elementwise launches of train.py:460-488, the per-sample Python loops
// of train.py:415-431 (gathers) and :508-530 (bank update), and F.normalize at :388/:530.
// All kernels: 128-bit vectorised, coalesced access where the layout allows it,
// warp-shuffle reductions, no floating-point atomics (deterministic; per-class accumulation
// follows the reference's sample order).
#include <type_traits>

#include "latte_common.cuh"
#include "tc_ptx.cuh"

namespace latte {
namespace {

__device__ __forceinline__ float ld1(const void* base, int64_t idx, int dtype) {
  if (dtype == LATTE_F32) return __ldg(reinterpret_cast<const float*>(base) + idx);
  if (dtype == LATTE_BF16)
    return __bfloat162float(__ldg(reinterpret_cast<const __nv_bfloat16*>(base) + idx));
  return __half2float(__ldg(reinterpret_cast<const __half*>(base) + idx));
}
__device__ __forceinline__ void st1(void* base, int64_t idx, int dtype, float v) {
  if (dtype == LATTE_F32) reinterpret_cast<float*>(base)[idx] = v;
  else if (dtype == LATTE_BF16) reinterpret_cast<__nv_bfloat16*>(base)[idx] = __float2bfloat16_rn(v);
  else reinterpret_cast<__half*>(base)[idx] = __float2half_rn(v);
}

// 4 consecutive elements (16 bytes for fp32, 8 bytes for 16-bit types)
__device__ __forceinline__ float4 ld4(const void* base, int64_t idx, int dtype) {
  if (dtype == LATTE_F32) return __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx));
  const uint2 raw = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(base) + idx));
  float4 o;
  if (dtype == LATTE_BF16) {
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&raw.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&raw.y);
    o.x = __bfloat162float(a.x); o.y = __bfloat162float(a.y);
    o.z = __bfloat162float(b.x); o.w = __bfloat162float(b.y);
  } else {
    const __half2 a = *reinterpret_cast<const __half2*>(&raw.x);
    const __half2 b = *reinterpret_cast<const __half2*>(&raw.y);
    o.x = __half2float(a.x); o.y = __half2float(a.y);
    o.z = __half2float(b.x); o.w = __half2float(b.y);
  }
  return o;
}
__device__ __forceinline__ void st4(void* base, int64_t idx, int dtype, float4 v) {
  if (dtype == LATTE_F32) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + idx) = v;
    return;
  }
  uint2 raw;
  if (dtype == LATTE_BF16) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    raw.x = *reinterpret_cast<uint32_t*>(&a); raw.y = *reinterpret_cast<uint32_t*>(&b);
  } else {
    __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    raw.x = *reinterpret_cast<uint32_t*>(&a); raw.y = *reinterpret_cast<uint32_t*>(&b);
  }
  *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(base) + idx) = raw;
}

// dtype as a template parameter: with a run-time dtype every load sits behind a branch and the
// 16-bit conversions consume each load where it is issued
template <int DT> __device__ __forceinline__ float4 ld4t(const void* base, int64_t idx) { return ld4(base, idx, DT); }
template <int DT> __device__ __forceinline__ void st4t(void* base, int64_t idx, float4 v) { st4(base, idx, DT, v); }

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < nw; ++i) t += red[i];
  return t;
}

// ------------------------------------------------------------------ normalize rows
__global__ void __launch_bounds__(256)
normalize_rows_kernel(const float* in, int64_t ld_in, float* out, int64_t ld_out, int64_t dim) {
  __shared__ float red[8];
  const int64_t r = blockIdx.x;
  const float* src = in + r * ld_in;
  float ss = 0.f;
  for (int64_t d = threadIdx.x; d < dim; d += blockDim.x) { const float v = src[d]; ss = fmaf(v, v, ss); }
  const float tot = block_sum(ss, red);
  const float inv = 1.0f / fmaxf(sqrtf(tot), 1e-12f);
  for (int64_t d = threadIdx.x; d < dim; d += blockDim.x) out[r * ld_out + d] = src[d] * inv;
}

// ------------------------------------------------------------------ mixture + EMA forward
struct MixArgs {
  const void* class_text; int64_t ld_ct;
  const void* per_image; int64_t ld_pi;
  const void* per_group; int64_t ld_pg;
  const float* bank; int64_t ld_bank;
  const int64_t* preds; const int64_t* zs;
  const float* w_lbl; const float* w_lbl_zs; const float* w_img; const float* w_grp;
  float alpha; int label_axis; int dtype;
  int64_t batch, dim;
  void* t_ft; void* t_zs; int64_t ld_out;
  int vec;
  int64_t num_classes;   // 1 when every pointer / stride allows 4-element vector access
};

__device__ __forceinline__ float mix_one(float wl, float l, float p, float wi, float g, float wg,
                                         float tot, float m, float alpha) {
  // same operation order as train.py:476-479 and :487
  float mix = wl * l + p * wi;
  mix = mix + g * wg;
  mix = mix / tot;
  return m + alpha * (mix - m);
}

// One WARP per row, persistent grid (4 CTAs of 8 warps per SM): a row needs two dependent global round
// trips (ids and weights, then the table / feature rows), and one 128-thread CTA per row (32768 CTAs of
// one iteration each) left that latency exposed at every CTA wave -- 16-bit rows took the same time as
// fp32 rows.  A lane issues the loads of all its 4-element pieces of the row before the arithmetic.
// Small batches (the reference's 512) keep one 128-thread CTA per row: a single round of loads.
template <int DT, bool kWarpRow>
__global__ void __launch_bounds__(256, 4) mix_ema_fwd_kernel(MixArgs a) {
  const int lane = kWarpRow ? (threadIdx.x & 31) : threadIdx.x;            // position inside the row team
  const int team = kWarpRow ? 32 : blockDim.x;                               // threads per row
  const int64_t nteams = kWarpRow ? (int64_t)gridDim.x * (blockDim.x >> 5) : gridDim.x;
  const int64_t team0 = kWarpRow ? (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5) : blockIdx.x;
  const bool quirk = a.label_axis == LATTE_LABEL_AXIS_QUIRK;
  for (int64_t i = team0; i < a.batch; i += nteams) {
    // ids index the [C, D] tables: clamp them so that a bad id (a stale pickled feature record) cannot
    // read out of bounds; the class sums drop such ids, the reference would raise KeyError
    const int64_t cp = min(max(a.preds[i], (int64_t)0), a.num_classes - 1);
    const int64_t cz = min(max(a.zs[i], (int64_t)0), a.num_classes - 1);
    const float wl_row = a.w_lbl[i], wi = a.w_img[i], wg = a.w_grp[i];
    const float tot = wl_row + wi + wg;              // train.py:472
    const float tot_zs = a.w_lbl_zs[i] + wi + wg;    // train.py:473
    if (a.vec) {
#pragma unroll 1
      for (int64_t d = (int64_t)lane * 4; d < a.dim; d += (int64_t)team * 4) {
        const float4 lf = ld4t<DT>(a.class_text, cp * a.ld_ct + d);
        const float4 lz = ld4t<DT>(a.class_text, cz * a.ld_ct + d);
        const float4 p = ld4t<DT>(a.per_image, i * a.ld_pi + d);
        const float4 g = ld4t<DT>(a.per_group, i * a.ld_pg + d);
        const float4 mf = __ldg(reinterpret_cast<const float4*>(a.bank + cp * a.ld_bank + d));
        const float4 mz = __ldg(reinterpret_cast<const float4*>(a.bank + cz * a.ld_bank + d));
        float4 wl = make_float4(wl_row, wl_row, wl_row, wl_row);
        if (quirk) wl = __ldg(reinterpret_cast<const float4*>(a.w_lbl + d));
        float4 of, oz;
        of.x = mix_one(wl.x, lf.x, p.x, wi, g.x, wg, tot, mf.x, a.alpha);
        of.y = mix_one(wl.y, lf.y, p.y, wi, g.y, wg, tot, mf.y, a.alpha);
        of.z = mix_one(wl.z, lf.z, p.z, wi, g.z, wg, tot, mf.z, a.alpha);
        of.w = mix_one(wl.w, lf.w, p.w, wi, g.w, wg, tot, mf.w, a.alpha);
        oz.x = mix_one(wl.x, lz.x, p.x, wi, g.x, wg, tot_zs, mz.x, a.alpha);
        oz.y = mix_one(wl.y, lz.y, p.y, wi, g.y, wg, tot_zs, mz.y, a.alpha);
        oz.z = mix_one(wl.z, lz.z, p.z, wi, g.z, wg, tot_zs, mz.z, a.alpha);
        oz.w = mix_one(wl.w, lz.w, p.w, wi, g.w, wg, tot_zs, mz.w, a.alpha);
        st4t<DT>(a.t_ft, i * a.ld_out + d, of);
        st4t<DT>(a.t_zs, i * a.ld_out + d, oz);
      }
    } else {
      for (int64_t d = lane; d < a.dim; d += team) {
        const float wl = quirk ? a.w_lbl[d] : wl_row;
        const float p = ld1(a.per_image, i * a.ld_pi + d, a.dtype);
        const float g = ld1(a.per_group, i * a.ld_pg + d, a.dtype);
        const float of = mix_one(wl, ld1(a.class_text, cp * a.ld_ct + d, a.dtype), p, wi, g, wg, tot,
                                 a.bank[cp * a.ld_bank + d], a.alpha);
        const float oz = mix_one(wl, ld1(a.class_text, cz * a.ld_ct + d, a.dtype), p, wi, g, wg, tot_zs,
                                 a.bank[cz * a.ld_bank + d], a.alpha);
        st1(a.t_ft, i * a.ld_out + d, a.dtype, of);
        st1(a.t_zs, i * a.ld_out + d, a.dtype, oz);
      }
    }
  }
}

// ------------------------------------------------------------------ mixture + EMA backward
// d_per_image[i,d] = w_img[i] * alpha * (dT_ft[i,d]/tot_i + dT_zs[i,d]/totz_i)   (same for group)
struct MixBwdArgs {
  const void* d_t_ft; const void* d_t_zs; int64_t ld_dt;
  const float* w_lbl; const float* w_lbl_zs; const float* w_img; const float* w_grp;
  float alpha; int dtype;
  int64_t batch, dim;
  void* d_per_image; void* d_per_group; int64_t ld_dp;
};

__global__ void __launch_bounds__(128) mix_ema_bwd_rows_kernel(MixBwdArgs a) {
  const int64_t i = blockIdx.x;
  const float wi = a.w_img[i], wg = a.w_grp[i];
  const float inv_ft = a.alpha / (a.w_lbl[i] + wi + wg);
  const float inv_zs = a.alpha / (a.w_lbl_zs[i] + wi + wg);
  for (int64_t d = threadIdx.x; d < a.dim; d += blockDim.x) {
    const float gf = ld1(a.d_t_ft, i * a.ld_dt + d, a.dtype);
    const float gz = ld1(a.d_t_zs, i * a.ld_dt + d, a.dtype);
    const float dm = gf * inv_ft + gz * inv_zs;
    st1(a.d_per_image, i * a.ld_dp + d, a.dtype, wi * dm);
    st1(a.d_per_group, i * a.ld_dp + d, a.dtype, wg * dm);
  }
}

// ------------------------------------------------------------------ per-class segment sums
// sums[c] = sum over the entries of class c, in the order of the reference loop
// (train.py:511-527: for every sample i ascending, row_zs(i) into class zs[i], then row_ft(i)
// into class preds[i]).  Entry e = 2 i + kind (kind 0 = zs list, 1 = ft list).
// Deterministic and parallel over the batch:
//   1. seg_hist    : per-slice class histograms of the 2B entries
//   2. seg_scan    : exclusive scans -> slice bases, class offsets, piece offsets, counts
//   3. seg_scatter : stable counting-sort placement -> order[] (entries grouped by class)
//   4. seg_piece   : one CTA per piece of <= 32 entries of one class: ordered partial row sum
//   5. seg_final   : per class, ordered sum of its pieces (+ column weights, accumulate)
// Every feature row is read once (step 4); no atomics on floating-point data.
constexpr int kSegSlice = 512;       // entries per slice
constexpr int kSegPiece = 32;        // entries per piece

struct SegArgs {
  const void* src_ft; const void* src_zs; int64_t ld_src; int dtype;
  const int64_t* preds; const int64_t* zs;
  int64_t batch, dim;
  int num_classes;
  // Optional factors (class-text gradient of the mixture, train.py:476-488):
  //   row from the ft list is scaled by alpha / (w_lbl[i]    + w_img[i] + w_grp[i]) * wl
  //   row from the zs list is scaled by alpha / (w_lbl_zs[i] + w_img[i] + w_grp[i]) * wl
  // with wl = w_lbl[i] (row axis) or w_lbl[d] (quirk axis).  All NULL => plain sums.
  const float* w_lbl; const float* w_lbl_zs; const float* w_img; const float* w_grp;
  float alpha; int label_axis;
  float* out; int64_t ld_out; int accumulate;
  float* counts;            // nullable
  float post_scale;
  // scratch (one stream-ordered allocation)
  int* hist;                // [slices, C] -> exclusive slice bases
  int* class_off;           // [C + 1]
  int* piece_off;           // [C + 1]
  int* order;               // [2 B]
  float* partial;           // [max_pieces, dim]
  int slices, max_pieces, vec;
};

__device__ __forceinline__ int seg_class_of(const SegArgs& a, int64_t e) {
  const int64_t i = e >> 1;
  const int64_t c = (e & 1) ? a.preds[i] : a.zs[i];
  return (c >= 0 && c < a.num_classes) ? (int)c : -1;     // out-of-range ids are dropped
}

__global__ void __launch_bounds__(256) seg_hist_kernel(SegArgs a) {
  extern __shared__ int sh[];
  for (int c = threadIdx.x; c < a.num_classes; c += 256) sh[c] = 0;
  __syncthreads();
  const int64_t e0 = (int64_t)blockIdx.x * kSegSlice;
  for (int k = threadIdx.x; k < kSegSlice; k += 256) {
    const int64_t e = e0 + k;
    if (e < 2 * a.batch) {
      const int c = seg_class_of(a, e);
      if (c >= 0) atomicAdd(&sh[c], 1);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < a.num_classes; c += 256)
    a.hist[(int64_t)blockIdx.x * a.num_classes + c] = sh[c];
}

__global__ void __launch_bounds__(1024) seg_scan_kernel(SegArgs a) {
  // one warp per class: exclusive scan over the slices (in place), 32 slices per step
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = warp; c < a.num_classes; c += 32) {
    int carry = 0;
    for (int s0 = 0; s0 < a.slices; s0 += 32) {
      const int s = s0 + lane;
      const int64_t idx = (int64_t)s * a.num_classes + c;
      const int v = s < a.slices ? a.hist[idx] : 0;
      int inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      if (s < a.slices) a.hist[idx] = carry + inc - v;
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) a.class_off[c + 1] = carry;          // class totals for now
  }
  __syncthreads();
  // exclusive scans over the classes (entries and pieces): warp 0, 32 classes per step
  if (warp == 0) {
    int eo = 0, po = 0;
    for (int c0 = 0; c0 < a.num_classes; c0 += 32) {
      const int c = c0 + lane;
      const int n = c < a.num_classes ? a.class_off[c + 1] : 0;
      const int pc = (n + kSegPiece - 1) / kSegPiece;
      int ie = n, ip = pc;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int te = __shfl_up_sync(0xffffffffu, ie, o);
        const int tp = __shfl_up_sync(0xffffffffu, ip, o);
        if (lane >= o) { ie += te; ip += tp; }
      }
      __syncwarp();
      if (c < a.num_classes) {
        if (a.counts) a.counts[c] = (float)n;
        a.class_off[c + 1] = eo + ie;
        a.piece_off[c + 1] = po + ip;
      }
      eo += __shfl_sync(0xffffffffu, ie, 31);
      po += __shfl_sync(0xffffffffu, ip, 31);
    }
    if (lane == 0) { a.class_off[0] = 0; a.piece_off[0] = 0; }
  }
}

__global__ void __launch_bounds__(256) seg_scatter_kernel(SegArgs a) {
  __shared__ int cls[kSegSlice];
  const int64_t e0 = (int64_t)blockIdx.x * kSegSlice;
  for (int k = threadIdx.x; k < kSegSlice; k += 256) {
    const int64_t e = e0 + k;
    cls[k] = e < 2 * a.batch ? seg_class_of(a, e) : -1;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < kSegSlice; k += 256) {
    const int c = cls[k];
    if (c < 0) continue;
    int rank = 0;                       // earlier entries of this slice with the same class
    for (int j = 0; j < k; ++j) rank += (cls[j] == c);
    const int pos = a.class_off[c] + a.hist[(int64_t)blockIdx.x * a.num_classes + c] + rank;
    a.order[pos] = (int)(e0 + k);
  }
}

__device__ __forceinline__ float seg_row_scale(const SegArgs& a, int64_t i, int kind, bool quirk) {
  if (!a.w_lbl) return 1.f;
  float sc = a.alpha / ((kind ? a.w_lbl[i] : a.w_lbl_zs[i]) + a.w_img[i] + a.w_grp[i]);
  if (!quirk) sc *= a.w_lbl[i];
  return sc;
}

__global__ void __launch_bounds__(128) seg_piece_kernel(SegArgs a) {
  const int b = blockIdx.x;
  if (b >= a.piece_off[a.num_classes]) return;
  int lo = 0, hi = a.num_classes;       // class c with piece_off[c] <= b < piece_off[c + 1]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (a.piece_off[mid] <= b) lo = mid; else hi = mid;
  }
  const int c = lo;
  const int first = a.class_off[c] + (b - a.piece_off[c]) * kSegPiece;
  const int last = min(first + kSegPiece, a.class_off[c + 1]);
  const bool quirk = a.w_lbl && a.label_axis == LATTE_LABEL_AXIS_QUIRK;
  float* dst = a.partial + (int64_t)b * a.dim;
  if (a.vec) {
    for (int64_t d = (int64_t)threadIdx.x * 4; d < a.dim; d += 512) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int k0 = first; k0 < last; k0 += 4) {
        float4 v[4];
        float sc[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int k = k0 + u;
          if (k < last) {
            const int e = a.order[k];
            const int64_t i = e >> 1;
            const int kind = e & 1;
            v[u] = ld4(kind ? a.src_ft : a.src_zs, i * a.ld_src + d, a.dtype);
            sc[u] = seg_row_scale(a, i, kind, quirk);
          } else {
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            sc[u] = 0.f;
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          acc.x = fmaf(v[u].x, sc[u], acc.x); acc.y = fmaf(v[u].y, sc[u], acc.y);
          acc.z = fmaf(v[u].z, sc[u], acc.z); acc.w = fmaf(v[u].w, sc[u], acc.w);
        }
      }
      *reinterpret_cast<float4*>(dst + d) = acc;
    }
  } else {
    for (int64_t d = threadIdx.x; d < a.dim; d += 128) {
      float acc = 0.f;
      for (int k = first; k < last; ++k) {
        const int e = a.order[k];
        const int64_t i = e >> 1;
        const int kind = e & 1;
        acc = fmaf(ld1(kind ? a.src_ft : a.src_zs, i * a.ld_src + d, a.dtype),
                   seg_row_scale(a, i, kind, quirk), acc);
      }
      dst[d] = acc;
    }
  }
}

__global__ void __launch_bounds__(128) seg_final_kernel(SegArgs a) {
  const int c = blockIdx.x;
  const int p0 = a.piece_off[c], p1 = a.piece_off[c + 1];
  const bool quirk = a.w_lbl && a.label_axis == LATTE_LABEL_AXIS_QUIRK;
  for (int64_t d = threadIdx.x; d < a.dim; d += 128) {
    float acc = 0.f;
    for (int pb = p0; pb < p1; ++pb) acc += a.partial[(int64_t)pb * a.dim + d];
    if (quirk) acc *= a.w_lbl[d];
    acc *= a.post_scale;
    float* o = a.out + (int64_t)c * a.ld_out + d;
    *o = a.accumulate ? *o + acc : acc;
  }
}

// ------------------------------------------------------------------ per-class sums, streaming form
// When the [C, D] fp32 accumulator fits in shared memory (C * D * 4 <= 200 KB: 47 x 512 = 94 KB) the
// per-class sums need no sort: the batch is cut into a FIXED number of row chunks (a function of the
// batch size only, so the result does not depend on the device), one CTA per chunk walks its rows
// in order -- entry order of the reference loop, zs row then ft row of every sample,
// train.py:511-527 -- adding each row into the accumulator row of its class (a thread owns its
// feature columns: no conflicts, no atomics), and writes its [C, D] partial; a second kernel adds
// the chunk partials in chunk order.  Every feature row is read exactly once, staged through a ring of
// 1-D bulk copies (up to 64 row pairs in flight per SM).  With `d_per_image` the same pass is the mixer's backward: it also writes the two
// row-shaped gradients from the rows it holds, so d_t_ft / d_t_zs are read once for all four
// outputs (the five-launch sort path re-read them for the class sums).
constexpr int kClsChunks = 148;          // fixed (one wave of a B200; the same on any device)
constexpr size_t kClsSmemMax = 200 * 1024;

struct ClsArgs {
  const void* src_ft; const void* src_zs; int64_t ld_src; int dtype;
  const int64_t* preds; const int64_t* zs;
  int64_t batch, dim;
  int num_classes;
  const float* w_lbl; const float* w_lbl_zs; const float* w_img; const float* w_grp;   // nullable set
  float alpha; int label_axis;
  void* d_per_image; void* d_per_group; int64_t ld_dp;      // nullable: row-shaped outputs
  float* partial;           // [chunks, C, dim]
  int* cnt_partial;         // [chunks, C]
  int chunks; int64_t rows_per_chunk;
  // final step
  float* out; int64_t ld_out; int accumulate; float* counts; float post_scale;
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// Shared-window accesses by 32-bit address: the accumulator read-modify-writes below are a dependent
// chain per class, and generic pointers to dynamic shared memory cost an S2UR + address rebuild per
// access (measured: 40 % of the consumer loop's stall samples).
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
// element `col` of a row staged in shared memory
template <int DT> __device__ __forceinline__ float lds_elem(uint32_t row, int col) {
  if (DT == LATTE_F32) return lds_f32(row + col * 4);
  unsigned short raw;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(raw) : "r"(row + col * 2) : "memory");
  if (DT == LATTE_BF16) return __uint_as_float((uint32_t)raw << 16);
  return __half2float(__ushort_as_half(raw));
}
template <int DT> __device__ __forceinline__ void st_elem(void* base, int64_t idx, float v) {
  if (DT == LATTE_F32) static_cast<float*>(base)[idx] = v;
  else if (DT == LATTE_BF16) static_cast<__nv_bfloat16*>(base)[idx] = __float2bfloat16_rn(v);
  else static_cast<__half*>(base)[idx] = __float2half_rn(v);
}

// Shared memory of cls_stream_kernel: [C, D] fp32 accumulator | ring of `groups` x 4 row pairs (the zs
// row and the ft row of one sample, filled by 1-D bulk copies: bytes in flight are bounded by the
// ring, not by registers; one mbarrier per group of 4 samples keeps the per-row bookkeeping small) |
// ids and per-sample factors of 256 samples | class counts | mbarriers.
constexpr int kClsIdRows = 256;
constexpr int kClsGrp = 4;
struct __align__(16) ClsIdA { int p, z; float sc_z, sc_f; };          // byte offsets of the class rows in the
                                                                      // accumulator (-1: dropped), scales
struct __align__(16) ClsIdB { float inv_ft, inv_zs, wi, wg; };        // row-output factors (mixer backward)

// DT: feature dtype; BWD: mixer backward (weighted rows, optional row-shaped outputs); VPT: columns per
// consumer thread.
// One column per thread (D <= 992) gives 16-24 consumer warps: each row costs a dependent
// ld.shared -> fma -> st.shared on the accumulator row of its class, and only warp-level parallelism
// hides that chain (8 warps with 2 columns each ran at 0.28 instructions / cycle / scheduler and made the
// kernel compute-bound at 0.3 of the HBM rate for fp32 rows, 0.15 for bf16).
template <int DT, bool BWD, int VPT>
__global__ void __launch_bounds__(1024) cls_stream_kernel(ClsArgs a, int groups, int row_bytes) {
  extern __shared__ __align__(128) uint8_t cls_smem[];
  using namespace ptx;
  const int dim = (int)a.dim;
  const size_t acc_bytes = ((size_t)a.num_classes * a.dim * 4 + 127) / 128 * 128;
  float* acc = reinterpret_cast<float*>(cls_smem);                                          // [C][dim]
  uint8_t* ring = cls_smem + acc_bytes;                                    // [groups][4][2][row_bytes]
  const int grp_bytes = kClsGrp * 2 * row_bytes;
  ClsIdA* ida = reinterpret_cast<ClsIdA*>(ring + (size_t)groups * grp_bytes);               // [256]
  ClsIdB* idb = reinterpret_cast<ClsIdB*>(ida + kClsIdRows);                                // [256]
  int* cnt = reinterpret_cast<int*>(idb + kClsIdRows);                                      // [C]
  const uint32_t bar_full = smem_u32(cnt + ((a.num_classes + 3) / 4 * 4));                  // [groups]
  const uint32_t bar_empty = bar_full + 8 * groups;                                         // [groups]
  const int ncons = (int)blockDim.x - 32;                  // consumer threads (the last warp produces)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool producer = warp == (ncons >> 5);
  const int64_t r0 = (int64_t)blockIdx.x * a.rows_per_chunk;
  const int64_t r1 = min(a.batch, r0 + a.rows_per_chunk);
  const int rows = (int)(r1 - r0);
  const int ngrp = (rows + kClsGrp - 1) / kClsGrp;

  for (int k = threadIdx.x; k < a.num_classes * dim; k += blockDim.x) acc[k] = 0.f;
  for (int k = threadIdx.x; k < a.num_classes; k += blockDim.x) cnt[k] = 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < groups; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, ncons >> 5);
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (producer) {
    if (lane == 0) {
      const uint64_t pol = policy_evict_first();          // every row is read once
      const uint8_t* sf = static_cast<const uint8_t*>(a.src_ft);
      const uint8_t* sz = static_cast<const uint8_t*>(a.src_zs);
      const int64_t pitch = a.ld_src * (int64_t)(row_bytes / a.dim);      // bytes per source row
      int s = 0;
      uint32_t phase = 0;
      for (int gk = 0; gk < ngrp; ++gk) {
        const int n = min(kClsGrp, rows - gk * kClsGrp);
        mbar_wait(bar_empty + 8 * s, phase ^ 1);
        const uint32_t dst = smem_u32(ring + (size_t)s * grp_bytes);
        const uint32_t bar = bar_full + 8 * s;
        mbar_arrive_expect_tx(bar, (uint32_t)(n * 2 * row_bytes));
        const int64_t off = (r0 + gk * kClsGrp) * pitch;
        if (pitch == row_bytes) {
          // contiguous rows: one copy per list for the whole group (the issue rate of bulk copies,
          // not their bytes, bounds a single producer thread)
          bulk_load_1d(dst, sz + off, (uint32_t)(n * row_bytes), bar, pol);
          bulk_load_1d(dst + kClsGrp * row_bytes, sf + off, (uint32_t)(n * row_bytes), bar, pol);
        } else {
          for (int u = 0; u < n; ++u) {
            bulk_load_1d(dst + u * row_bytes, sz + off + u * pitch, (uint32_t)row_bytes, bar, pol);
            bulk_load_1d(dst + (kClsGrp + u) * row_bytes, sf + off + u * pitch, (uint32_t)row_bytes, bar, pol);
          }
        }
        if (++s == groups) { s = 0; phase ^= 1; }
      }
    }
    return;
  }

  // ------------------------------------------------------------ consumers: thread -> VPT columns
  const int d0 = (int)threadIdx.x * VPT;
  const bool col_ok = d0 < dim;
  const int d = col_ok ? d0 : 0;                              // idle lanes re-read column 0
  const bool quirk = BWD && a.label_axis == LATTE_LABEL_AXIS_QUIRK;
  const uint32_t acc_col = smem_u32(acc) + (uint32_t)d * 4;
  const uint32_t ring_base = smem_u32(ring);
  const uint32_t ida_base = smem_u32(ida), idb_base = smem_u32(idb);
  int s = 0;
  uint32_t phase = 0;
  for (int k0 = 0; k0 < rows; k0 += kClsIdRows) {
    // ids and factors of the next 256 samples (coalesced), validated and counted once
    named_bar_sync(1, ncons);
    for (int t = threadIdx.x; t < kClsIdRows && k0 + t < rows; t += ncons) {
      const int64_t i = r0 + k0 + t;
      const int64_t p = a.preds[i], z = a.zs[i];
      ClsIdA q;
      const bool p_ok = p >= 0 && p < a.num_classes, z_ok = z >= 0 && z < a.num_classes;
      q.p = p_ok ? (int)p * dim * 4 : -1;                     // out-of-range ids are dropped
      q.z = z_ok ? (int)z * dim * 4 : -1;
      q.sc_z = q.sc_f = 1.f;
      if (BWD) {
        const float wl = a.w_lbl[i], wlz = a.w_lbl_zs[i], wi = a.w_img[i], wg = a.w_grp[i];
        ClsIdB f;
        f.inv_ft = a.alpha / (wl + wi + wg);
        f.inv_zs = a.alpha / (wlz + wi + wg);
        f.wi = wi; f.wg = wg;
        const float wle = quirk ? 1.f : wl;   // quirk axis: the label weight is a column factor (final step)
        q.sc_z = f.inv_zs * wle;
        q.sc_f = f.inv_ft * wle;
        idb[t] = f;
      }
      ida[t] = q;
      if (z_ok) atomicAdd(cnt + (int)z, 1);                   // integer counts: order does not matter
      if (p_ok) atomicAdd(cnt + (int)p, 1);
    }
    named_bar_sync(1, ncons);
    const int kend = min(rows, k0 + kClsIdRows);
    for (int k = k0; k < kend; k += kClsGrp) {
      mbar_wait(bar_full + 8 * s, phase);
      const uint32_t grp = ring_base + (uint32_t)s * grp_bytes;
      float gz[kClsGrp][VPT], gf[kClsGrp][VPT];
#pragma unroll
      for (int u = 0; u < kClsGrp; ++u)                        // rows past the chunk: stale bytes, unused
#pragma unroll
        for (int v = 0; v < VPT; ++v) {                        // group layout: 4 zs rows, then 4 ft rows
          gz[u][v] = lds_elem<DT>(grp + u * row_bytes, d + v);
          gf[u][v] = lds_elem<DT>(grp + (kClsGrp + u) * row_bytes, d + v);
        }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_empty + 8 * s);            // this warp has its copy of the group
      if (++s == groups) { s = 0; phase ^= 1; }
      if (!col_ok) continue;
      auto one_sample = [&](int u) {
        const uint4 qa = lds_u4(ida_base + (uint32_t)(k + u - k0) * 16);
        const int qp = (int)qa.x, qz = (int)qa.y;
        const float sc_z = __uint_as_float(qa.z), sc_f = __uint_as_float(qa.w);
        if (BWD && a.d_per_image) {
          const uint4 qb = lds_u4(idb_base + (uint32_t)(k + u - k0) * 16);
          const float inv_ft = __uint_as_float(qb.x), inv_zs = __uint_as_float(qb.y);
          const float wi = __uint_as_float(qb.z), wg = __uint_as_float(qb.w);
          const int64_t i = r0 + k + u;
#pragma unroll
          for (int v = 0; v < VPT; ++v) {
            const float dx = gf[u][v] * inv_ft + gz[u][v] * inv_zs;
            st_elem<DT>(a.d_per_image, i * a.ld_dp + d + v, wi * dx);
            st_elem<DT>(a.d_per_group, i * a.ld_dp + d + v, wg * dx);
          }
        }
        if (qz >= 0) {             // entry 2i: the zs-list row first (train.py:524)
          const uint32_t o = acc_col + (uint32_t)qz;
#pragma unroll
          for (int v = 0; v < VPT; ++v) sts_f32(o + 4 * v, fmaf(gz[u][v], sc_z, lds_f32(o + 4 * v)));
        }
        if (qp >= 0) {             // entry 2i + 1: the ft-list row (train.py:525)
          const uint32_t o = acc_col + (uint32_t)qp;
#pragma unroll
          for (int v = 0; v < VPT; ++v) sts_f32(o + 4 * v, fmaf(gf[u][v], sc_f, lds_f32(o + 4 * v)));
        }
      };
      if (k + kClsGrp <= kend) {
#pragma unroll
        for (int u = 0; u < kClsGrp; ++u) one_sample(u);
      } else {
        for (int u = 0; u < kend - k; ++u) {
          // a partial group at the end of the chunk: select the row by index (no dynamic register indexing)
          if (u == 0) one_sample(0);
          else if (u == 1) one_sample(1);
          else one_sample(2);
        }
      }
    }
  }
  named_bar_sync(1, ncons);
  float* dst = a.partial + (int64_t)blockIdx.x * a.num_classes * dim;
  for (int k = threadIdx.x; k < a.num_classes * dim; k += ncons) dst[k] = acc[k];
  for (int k = threadIdx.x; k < a.num_classes; k += ncons)
    a.cnt_partial[(int64_t)blockIdx.x * a.num_classes + k] = cnt[k];
}

// out[c] (+)= post_scale * colweight * sum over the chunks.  One CTA per (class, 32 vector columns):
// warp g adds the chunks of quarter g in chunk order, the four quarter sums are added in order.
__global__ void __launch_bounds__(128) cls_final_kernel(ClsArgs a) {
  __shared__ float4 quarter[4][32];
  __shared__ int cnt_red[4];
  const int nvec = (int)(a.dim / 4);
  const int c = blockIdx.x;
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int v = blockIdx.y * 32 + lane;
  if (blockIdx.y == 0 && a.counts) {
    int n = 0;
    for (int k = threadIdx.x; k < a.chunks; k += 128) n += a.cnt_partial[(int64_t)k * a.num_classes + c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if (lane == 0) cnt_red[grp] = n;
  }
  const int per = (a.chunks + 3) / 4;
  const int k_lo = grp * per, k_hi = min(a.chunks, k_lo + per);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (v < nvec) {
    const float4* src = reinterpret_cast<const float4*>(a.partial) + (int64_t)c * nvec + v;
    const int64_t step = (int64_t)a.num_classes * nvec;
    for (int k0 = k_lo; k0 < k_hi; k0 += 8) {
      float4 t[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) t[u] = __ldcg(src + (int64_t)min(k0 + u, k_hi - 1) * step);
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (k0 + u < k_hi) { s.x += t[u].x; s.y += t[u].y; s.z += t[u].z; s.w += t[u].w; }
    }
  }
  quarter[grp][lane] = s;
  __syncthreads();
  if (blockIdx.y == 0 && threadIdx.x == 0 && a.counts)
    a.counts[c] = (float)(cnt_red[0] + cnt_red[1] + cnt_red[2] + cnt_red[3]);
  if (grp != 0 || v >= nvec) return;
#pragma unroll
  for (int g = 1; g < 4; ++g) {
    const float4 t = quarter[g][lane];
    s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
  }
  if (a.w_lbl && a.label_axis == LATTE_LABEL_AXIS_QUIRK) {
    const float4 w = __ldg(reinterpret_cast<const float4*>(a.w_lbl) + v);
    s.x *= w.x; s.y *= w.y; s.z *= w.z; s.w *= w.w;
  }
  s.x *= a.post_scale; s.y *= a.post_scale; s.z *= a.post_scale; s.w *= a.post_scale;
  float4* o = reinterpret_cast<float4*>(a.out + (int64_t)c * a.ld_out) + v;
  if (a.accumulate) { const float4 p = *o; s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w; }
  *o = s;
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

struct ClsGeom { bool ok; int chunks; int64_t rows_per_chunk; size_t smem, part_bytes, cnt_bytes; int slots; };
ClsGeom cls_geom(int64_t batch, int64_t dim, int64_t num_classes, int esz = 4) {
  ClsGeom g{};
  const size_t acc = ((size_t)num_classes * (size_t)dim * 4 + 127) / 128 * 128;
  const size_t fixed = acc + (size_t)kClsIdRows * (sizeof(ClsIdA) + sizeof(ClsIdB)) +
                       ((size_t)num_classes + 3) / 4 * 16 + 64;
  const size_t row_bytes = (size_t)dim * esz;
  const size_t grp = (size_t)kClsGrp * 2 * row_bytes + 16;          // one group of 4 row pairs + 2 mbarriers
  g.ok = batch > 0 && dim % 8 == 0 && dim / 2 <= 992 && fixed + 2 * grp <= kClsSmemMax;
  if (!g.ok) return g;
  size_t groups = (kClsSmemMax - fixed) / grp;
  if (groups > 16) groups = 16;
  g.slots = (int)groups;
  g.smem = fixed + groups * grp;
  g.rows_per_chunk = (batch + kClsChunks - 1) / kClsChunks;
  if (g.rows_per_chunk < 16) g.rows_per_chunk = 16;
  g.chunks = (int)((batch + g.rows_per_chunk - 1) / g.rows_per_chunk);
  g.part_bytes = ((size_t)g.chunks * (size_t)num_classes * (size_t)dim * 4 + 255) / 256 * 256;
  g.cnt_bytes = ((size_t)g.chunks * (size_t)num_classes * 4 + 255) / 256 * 256;
  return g;
}

// -> LATTE_OK when the streaming form ran, 1 when it does not apply (caller takes the sort path)
int class_sums_stream(ClsArgs a, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  const int64_t esz = (int64_t)dtype_size(a.dtype);
  const ClsGeom g = cls_geom(a.batch, a.dim, a.num_classes, (int)esz);
  // bulk copies want 16-byte aligned rows; the row outputs and the final step use 4-element vectors
  const bool vec = g.ok && ((a.ld_src * esz) % 16 == 0) && (a.ld_out % 4 == 0) && al16(a.out) &&
                   al16(a.src_ft) && al16(a.src_zs) &&
                   (!a.d_per_image || ((a.ld_dp % 4 == 0) &&
                                       (reinterpret_cast<uintptr_t>(a.d_per_image) % (4 * esz) == 0) &&
                                       (reinterpret_cast<uintptr_t>(a.d_per_group) % (4 * esz) == 0))) &&
                   (!a.w_lbl || al16(a.w_lbl));
  if (!vec) return 1;
  if (!workspace) return LATTE_ERR_BAD_ARG;
  const uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 255) / 256 * 256;
  if (base - reinterpret_cast<uintptr_t>(workspace) + g.part_bytes + g.cnt_bytes > workspace_bytes)
    return LATTE_ERR_WORKSPACE;
  a.partial = reinterpret_cast<float*>(base);
  a.cnt_partial = reinterpret_cast<int*>(base + g.part_bytes);
  a.chunks = g.chunks;
  a.rows_per_chunk = g.rows_per_chunk;
  const int nvec = (int)(a.dim / 4);
  const int vpt = a.dim <= 992 ? 1 : 2;                            // columns per consumer thread
  const int threads = (int)((a.dim / vpt + 31) / 32 * 32) + 32;    // consumers + the producer warp
  const int row_bytes = (int)(a.dim * esz);
  const bool bwd = a.w_lbl != nullptr;
  auto launch = [&](auto kernel) -> int {
    if (g.smem > 48 * 1024)
      LATTE_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
    kernel<<<g.chunks, threads, g.smem, st>>>(a, g.slots, row_bytes);
    return LATTE_OK;
  };
  auto pick = [&](auto dt) -> int {
    constexpr int DT = decltype(dt)::value;
    if (bwd) return vpt == 1 ? launch(cls_stream_kernel<DT, true, 1>) : launch(cls_stream_kernel<DT, true, 2>);
    return vpt == 1 ? launch(cls_stream_kernel<DT, false, 1>) : launch(cls_stream_kernel<DT, false, 2>);
  };
  int lrc;
  if (a.dtype == LATTE_F32) lrc = pick(std::integral_constant<int, LATTE_F32>{});
  else if (a.dtype == LATTE_BF16) lrc = pick(std::integral_constant<int, LATTE_BF16>{});
  else lrc = pick(std::integral_constant<int, LATTE_F16>{});
  if (lrc) return lrc;
  cls_final_kernel<<<dim3((unsigned)a.num_classes, (unsigned)((nvec + 31) / 32)), 128, 0, st>>>(a);
  if (cudaGetLastError() != cudaSuccess) return LATTE_ERR_CUDA;
  return LATTE_OK;
}

struct SegWs { size_t n_hist, int_bytes, part_bytes; int slices, max_pieces; };
SegWs seg_ws(int64_t batch, int64_t dim, int64_t num_classes) {
  SegWs w;
  const int64_t entries = 2 * batch;
  w.slices = (int)((entries + kSegSlice - 1) / kSegSlice);
  if (w.slices < 1) w.slices = 1;
  w.max_pieces = (int)((entries + kSegPiece - 1) / kSegPiece) + (int)num_classes;
  w.n_hist = (size_t)w.slices * (size_t)num_classes;
  const size_t n_int = w.n_hist + 2 * ((size_t)num_classes + 1) + (size_t)entries + 8;
  w.int_bytes = (n_int * sizeof(int) + 255) / 256 * 256;
  w.part_bytes = (size_t)w.max_pieces * (size_t)dim * sizeof(float);
  return w;
}

// Runs the five steps on `stream` in the caller's workspace (latte_seg_workspace_bytes).
int segment_sums(SegArgs a, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  const SegWs w = seg_ws(a.batch, a.dim, a.num_classes);
  a.slices = w.slices;
  a.max_pieces = w.max_pieces;
  const size_t n_hist = w.n_hist, int_bytes = w.int_bytes;
  if (!workspace) return LATTE_ERR_BAD_ARG;
  const uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 255) / 256 * 256;
  if (base - reinterpret_cast<uintptr_t>(workspace) + w.int_bytes + w.part_bytes > workspace_bytes)
    return LATTE_ERR_WORKSPACE;
  char* buf = reinterpret_cast<char*>(base);
  int* ip = reinterpret_cast<int*>(buf);
  a.hist = ip; ip += n_hist;
  a.class_off = ip; ip += a.num_classes + 1;
  a.piece_off = ip; ip += a.num_classes + 1;
  a.order = ip;
  a.partial = reinterpret_cast<float*>(buf + int_bytes);
  const int64_t esz = (int64_t)dtype_size(a.dtype);
  a.vec = (a.dim % 4 == 0) && (a.ld_src % 4 == 0) &&
          (reinterpret_cast<uintptr_t>(a.src_ft) % (4 * esz) == 0) &&
          (reinterpret_cast<uintptr_t>(a.src_zs) % (4 * esz) == 0);
  int rc = LATTE_OK;
  seg_hist_kernel<<<a.slices, 256, (size_t)a.num_classes * sizeof(int), st>>>(a);
  seg_scan_kernel<<<1, 1024, 0, st>>>(a);
  seg_scatter_kernel<<<a.slices, 256, 0, st>>>(a);
  seg_piece_kernel<<<a.max_pieces, 128, 0, st>>>(a);
  seg_final_kernel<<<a.num_classes, 128, 0, st>>>(a);
  if (cudaGetLastError() != cudaSuccess) rc = LATTE_ERR_CUDA;
  return rc;
}

// ------------------------------------------------------------------ bank finalize
__global__ void __launch_bounds__(256)
bank_finalize_kernel(const float* sums, int64_t ld_sums, const float* counts, float* bank,
                     int64_t ld_bank, int64_t dim) {
  __shared__ float red[8];
  const int64_t c = blockIdx.x;
  const float cnt = counts[c];
  if (!(cnt > 0.f)) return;                     // untouched class keeps its row (train.py:528)
  const float* s = sums + c * ld_sums;
  float ss = 0.f;
  for (int64_t d = threadIdx.x; d < dim; d += blockDim.x) {
    const float v = s[d] / cnt;                 // train.py:529
    ss = fmaf(v, v, ss);
  }
  const float tot = block_sum(ss, red);
  const float denom = fmaxf(sqrtf(tot), 1e-12f);  // F.normalize eps, train.py:530
  for (int64_t d = threadIdx.x; d < dim; d += blockDim.x) bank[c * ld_bank + d] = (s[d] / cnt) / denom;
}

}  // namespace
}  // namespace latte

using namespace latte;

extern "C" int latte_seg_workspace_bytes(int64_t batch, int64_t dim, int64_t num_classes,
                                         size_t* bytes) {
  LATTE_CHECK_ARG(bytes && batch >= 0 && dim > 0 && num_classes > 0);
  const SegWs w = seg_ws(batch, dim, num_classes);
  *bytes = w.int_bytes + w.part_bytes + 256;
  const ClsGeom g = cls_geom(batch, dim, num_classes, 2);   // streaming form: chunk partials
  if (g.ok && g.part_bytes + g.cnt_bytes + 256 > *bytes) *bytes = g.part_bytes + g.cnt_bytes + 256;
  return LATTE_OK;
}

extern "C" int latte_normalize_rows(const float* in, int64_t ld_in, float* out, int64_t ld_out,
                                    int64_t rows, int64_t dim, void* stream) {
  LATTE_CHECK_ARG(in && out && rows >= 0 && dim > 0 && ld_in >= dim && ld_out >= dim);
  if (rows == 0) return LATTE_OK;
  normalize_rows_kernel<<<(unsigned)rows, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      in, ld_in, out, ld_out, dim);
  LATTE_LAUNCH_OK();
  return LATTE_OK;
}

extern "C" int latte_mix_ema_fwd(const void* class_text, int64_t ld_ct, const void* per_image,
                                 int64_t ld_pi, const void* per_group, int64_t ld_pg,
                                 const float* bank, int64_t ld_bank, const int64_t* preds,
                                 const int64_t* zs, const float* w_lbl, const float* w_lbl_zs,
                                 const float* w_img, const float* w_grp, float alpha,
                                 int label_axis, int dtype, int64_t batch, int64_t dim,
                                 int64_t num_classes, void* t_ft, void* t_zs, int64_t ld_out,
                                 void* stream) {
  LATTE_CHECK_ARG(class_text && per_image && per_group && bank && preds && zs && w_lbl &&
                  w_lbl_zs && w_img && w_grp && t_ft && t_zs);
  LATTE_CHECK_ARG(batch >= 0 && dim > 0 && num_classes > 0);
  LATTE_CHECK_ARG(dtype >= LATTE_F32 && dtype <= LATTE_F16);
  LATTE_CHECK_ARG(label_axis == LATTE_LABEL_AXIS_ROW || label_axis == LATTE_LABEL_AXIS_QUIRK);
  // the reference's literal broadcast only exists for B == D (train.py:476)
  if (label_axis == LATTE_LABEL_AXIS_QUIRK && batch != dim) return LATTE_ERR_UNSUPPORTED;
  if (batch == 0) return LATTE_OK;
  MixArgs a{class_text, ld_ct, per_image, ld_pi, per_group, ld_pg, bank, ld_bank, preds, zs,
            w_lbl, w_lbl_zs, w_img, w_grp, alpha, label_axis, dtype, batch, dim, t_ft, t_zs,
            ld_out, 0, num_classes};
  const int64_t esz = (int64_t)dtype_size(dtype);
  const int64_t vb = 4 * esz;   // bytes per 4-element vector
  a.vec = (dim % 4 == 0) && (ld_ct % 4 == 0) && (ld_pi % 4 == 0) && (ld_pg % 4 == 0) &&
          (ld_bank % 4 == 0) && (ld_out % 4 == 0) && al16(bank) && al16(w_lbl) &&
          (reinterpret_cast<uintptr_t>(class_text) % vb == 0) &&
          (reinterpret_cast<uintptr_t>(per_image) % vb == 0) &&
          (reinterpret_cast<uintptr_t>(per_group) % vb == 0) &&
          (reinterpret_cast<uintptr_t>(t_ft) % vb == 0) && (reinterpret_cast<uintptr_t>(t_zs) % vb == 0);
  cudaStream_t mst = static_cast<cudaStream_t>(stream);
  if (batch >= 8192) {
    int64_t mgrid = (batch + 7) / 8;                                   // 8 warps (rows) per CTA
    if (mgrid > 4 * (int64_t)device_sm_count()) mgrid = 4 * (int64_t)device_sm_count();
    if (dtype == LATTE_F32) mix_ema_fwd_kernel<LATTE_F32, true><<<(unsigned)mgrid, 256, 0, mst>>>(a);
    else if (dtype == LATTE_BF16) mix_ema_fwd_kernel<LATTE_BF16, true><<<(unsigned)mgrid, 256, 0, mst>>>(a);
    else mix_ema_fwd_kernel<LATTE_F16, true><<<(unsigned)mgrid, 256, 0, mst>>>(a);
  } else {
    if (dtype == LATTE_F32) mix_ema_fwd_kernel<LATTE_F32, false><<<(unsigned)batch, 128, 0, mst>>>(a);
    else if (dtype == LATTE_BF16) mix_ema_fwd_kernel<LATTE_BF16, false><<<(unsigned)batch, 128, 0, mst>>>(a);
    else mix_ema_fwd_kernel<LATTE_F16, false><<<(unsigned)batch, 128, 0, mst>>>(a);
  }
  LATTE_LAUNCH_OK();
  return LATTE_OK;
}

extern "C" int latte_mix_ema_bwd(const void* d_t_ft, const void* d_t_zs, int64_t ld_dt,
                                 const int64_t* preds, const int64_t* zs, const float* w_lbl,
                                 const float* w_lbl_zs, const float* w_img, const float* w_grp,
                                 float alpha, int label_axis, int dtype, int64_t batch,
                                 int64_t dim, int64_t num_classes, float* d_class_text,
                                 int64_t ld_dct, void* d_per_image, void* d_per_group,
                                 int64_t ld_dp, float* d_bank, int64_t ld_dbank, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  LATTE_CHECK_ARG(d_t_ft && d_t_zs && preds && zs && w_lbl && w_lbl_zs && w_img && w_grp);
  LATTE_CHECK_ARG(batch >= 0 && dim > 0 && num_classes > 0);
  LATTE_CHECK_ARG(dtype >= LATTE_F32 && dtype <= LATTE_F16);
  if (label_axis == LATTE_LABEL_AXIS_QUIRK && batch != dim) return LATTE_ERR_UNSUPPORTED;
  if (num_classes > 12000) return LATTE_ERR_UNSUPPORTED;     // class histogram lives in smem
  if (batch == 0) return LATTE_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  bool rows_done = false, class_done = false;
  if (d_class_text) {
    // one pass over d_t_ft / d_t_zs for the row gradients and the class-text sums
    ClsArgs c{};
    c.src_ft = d_t_ft; c.src_zs = d_t_zs; c.ld_src = ld_dt; c.dtype = dtype;
    c.preds = preds; c.zs = zs; c.batch = batch; c.dim = dim; c.num_classes = (int)num_classes;
    c.w_lbl = w_lbl; c.w_lbl_zs = w_lbl_zs; c.w_img = w_img; c.w_grp = w_grp;
    c.alpha = alpha; c.label_axis = label_axis;
    const bool rows = d_per_image && d_per_group;
    c.d_per_image = rows ? d_per_image : nullptr; c.d_per_group = rows ? d_per_group : nullptr; c.ld_dp = ld_dp;
    c.out = d_class_text; c.ld_out = ld_dct; c.accumulate = 1; c.counts = nullptr; c.post_scale = 1.f;
    const int rc = class_sums_stream(c, workspace, workspace_bytes, st);
    if (rc < 0) return rc;
    if (rc == LATTE_OK) { class_done = true; rows_done = rows; }
  }
  if (d_per_image && d_per_group && !rows_done) {
    MixBwdArgs a{d_t_ft, d_t_zs, ld_dt, w_lbl, w_lbl_zs, w_img, w_grp, alpha, dtype, batch, dim,
                 d_per_image, d_per_group, ld_dp};
    mix_ema_bwd_rows_kernel<<<(unsigned)batch, 128, 0, st>>>(a);
    LATTE_LAUNCH_OK();
  }
  if (d_class_text && !class_done) {
    // d_class_text[c] += sum_{zs_i=c} wl * alpha/totz_i * dT_zs[i] + sum_{preds_i=c} wl * alpha/tot_i * dT_ft[i]
    SegArgs s{};
    s.src_ft = d_t_ft; s.src_zs = d_t_zs; s.ld_src = ld_dt; s.dtype = dtype;
    s.preds = preds; s.zs = zs; s.batch = batch; s.dim = dim; s.num_classes = (int)num_classes;
    s.w_lbl = w_lbl; s.w_lbl_zs = w_lbl_zs; s.w_img = w_img; s.w_grp = w_grp;
    s.alpha = alpha; s.label_axis = label_axis;
    s.out = d_class_text; s.ld_out = ld_dct; s.accumulate = 1; s.counts = nullptr; s.post_scale = 1.f;
    const int rc = segment_sums(s, workspace, workspace_bytes, st);
    if (rc) return rc;
  }
  if (d_bank) {
    // (1 - alpha) * d_t scattered by class (train.py:487-488 wrt membank_features)
    SegArgs s{};
    s.src_ft = d_t_ft; s.src_zs = d_t_zs; s.ld_src = ld_dt; s.dtype = dtype;
    s.preds = preds; s.zs = zs; s.batch = batch; s.dim = dim; s.num_classes = (int)num_classes;
    s.label_axis = LATTE_LABEL_AXIS_ROW;
    s.out = d_bank; s.ld_out = ld_dbank; s.accumulate = 1; s.counts = nullptr;
    s.post_scale = 1.f - alpha;
    const int rc = segment_sums(s, workspace, workspace_bytes, st);
    if (rc) return rc;
  }
  return LATTE_OK;
}

extern "C" int latte_bank_accumulate(const void* t_ft, const void* t_zs, int64_t ld_t, int dtype,
                                     const int64_t* preds, const int64_t* zs, int64_t batch,
                                     int64_t dim, int64_t num_classes, float* sums,
                                     int64_t ld_sums, float* counts, void* workspace,
                                     size_t workspace_bytes, void* stream) {
  LATTE_CHECK_ARG(t_ft && t_zs && preds && zs && sums && counts);
  LATTE_CHECK_ARG(batch >= 0 && dim > 0 && num_classes > 0 && ld_t >= dim && ld_sums >= dim);
  LATTE_CHECK_ARG(dtype >= LATTE_F32 && dtype <= LATTE_F16);
  if (num_classes > 12000) return LATTE_ERR_UNSUPPORTED;       // class histogram lives in smem
  if (batch > 0) {
    ClsArgs c{};
    c.src_ft = t_ft; c.src_zs = t_zs; c.ld_src = ld_t; c.dtype = dtype;
    c.preds = preds; c.zs = zs; c.batch = batch; c.dim = dim; c.num_classes = (int)num_classes;
    c.label_axis = LATTE_LABEL_AXIS_ROW;
    c.out = sums; c.ld_out = ld_sums; c.accumulate = 0; c.counts = counts; c.post_scale = 1.f;
    const int rc = class_sums_stream(c, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
    if (rc <= 0) return rc;
  }
  SegArgs s{};
  s.src_ft = t_ft; s.src_zs = t_zs; s.ld_src = ld_t; s.dtype = dtype;
  s.preds = preds; s.zs = zs; s.batch = batch; s.dim = dim; s.num_classes = (int)num_classes;
  s.label_axis = LATTE_LABEL_AXIS_ROW;
  s.out = sums; s.ld_out = ld_sums; s.accumulate = 0; s.counts = counts; s.post_scale = 1.f;
  return segment_sums(s, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

extern "C" int latte_bank_finalize(const float* sums, int64_t ld_sums, const float* counts,
                                   float* bank, int64_t ld_bank, int64_t dim,
                                   int64_t num_classes, void* stream) {
  LATTE_CHECK_ARG(sums && counts && bank && dim > 0 && num_classes > 0);
  bank_finalize_kernel<<<(unsigned)num_classes, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      sums, ld_sums, counts, bank, ld_bank, dim);
  LATTE_LAUNCH_OK();
  return LATTE_OK;
}

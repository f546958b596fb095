// HBM-bound prototype kernels: row normalisation, text mixture + EMA (forward and
// backward), memory-bank segment sums and per-class normalise.
// Replaces the ~20 elementwise launches of train.py:460-488, the per-sample Python loops
// of train.py:415-431 (gathers) and :508-530 (bank update), and F.normalize at :388/:530.
// All kernels: 128-bit vectorised, coalesced access where the layout allows it,
// warp-shuffle reductions, no atomics (deterministic; per-class accumulation follows the
// reference's sample order).
#include "latte_common.cuh"

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
  int vec;   // 1 when every pointer / stride allows 4-element vector access
};

__device__ __forceinline__ float mix_one(float wl, float l, float p, float wi, float g, float wg,
                                         float tot, float m, float alpha) {
  // same operation order as train.py:476-479 and :487
  float mix = wl * l + p * wi;
  mix = mix + g * wg;
  mix = mix / tot;
  return m + alpha * (mix - m);
}

__global__ void __launch_bounds__(128) mix_ema_fwd_kernel(MixArgs a) {
  const int64_t i = blockIdx.x;
  const int64_t cp = a.preds[i], cz = a.zs[i];
  const float wl_row = a.w_lbl[i], wi = a.w_img[i], wg = a.w_grp[i];
  const float tot = wl_row + wi + wg;              // train.py:472
  const float tot_zs = a.w_lbl_zs[i] + wi + wg;    // train.py:473
  const bool quirk = a.label_axis == LATTE_LABEL_AXIS_QUIRK;
  if (a.vec) {
    for (int64_t d = (int64_t)threadIdx.x * 4; d < a.dim; d += (int64_t)blockDim.x * 4) {
      const float4 lf = ld4(a.class_text, cp * a.ld_ct + d, a.dtype);
      const float4 lz = ld4(a.class_text, cz * a.ld_ct + d, a.dtype);
      const float4 p = ld4(a.per_image, i * a.ld_pi + d, a.dtype);
      const float4 g = ld4(a.per_group, i * a.ld_pg + d, a.dtype);
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
      st4(a.t_ft, i * a.ld_out + d, a.dtype, of);
      st4(a.t_zs, i * a.ld_out + d, a.dtype, oz);
    }
  } else {
    for (int64_t d = threadIdx.x; d < a.dim; d += blockDim.x) {
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
// One CTA per class c scans the sample indices in order; for every sample i (ascending)
// it adds row_zs(i) if zs[i] == c and then row_ft(i) if preds[i] == c -- the order of the
// reference loop (train.py:511-527).  Each thread owns columns tid, tid + 256, ...
constexpr int kSegThreads = 256;
constexpr int kSegMaxCols = 8;   // dim <= 2048

struct SegArgs {
  const void* src_ft; const void* src_zs; int64_t ld_src; int dtype;
  const int64_t* preds; const int64_t* zs;
  int64_t batch, dim;
  // Optional factors (class-text gradient of the mixture, train.py:476-488):
  //   row from the ft list is scaled by alpha / (w_lbl[i]    + w_img[i] + w_grp[i]) * wl
  //   row from the zs list is scaled by alpha / (w_lbl_zs[i] + w_img[i] + w_grp[i]) * wl
  // with wl = w_lbl[i] (row axis) or w_lbl[d] (quirk axis).  All NULL => plain sums.
  const float* w_lbl; const float* w_lbl_zs; const float* w_img; const float* w_grp;
  float alpha; int label_axis;
  float* out; int64_t ld_out; int accumulate;
  float* counts;            // nullable
  float post_scale;
};

__global__ void __launch_bounds__(kSegThreads) segment_sum_kernel(SegArgs a) {
  __shared__ unsigned mz_s[kSegThreads / 32], mp_s[kSegThreads / 32];
  const int c = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float acc[kSegMaxCols];
#pragma unroll
  for (int j = 0; j < kSegMaxCols; ++j) acc[j] = 0.f;
  float colw[kSegMaxCols];
  const bool quirk = a.w_lbl && a.label_axis == LATTE_LABEL_AXIS_QUIRK;
#pragma unroll
  for (int j = 0; j < kSegMaxCols; ++j) {
    const int64_t d = tid + (int64_t)j * kSegThreads;
    colw[j] = (quirk && d < a.dim) ? a.w_lbl[d] : 1.f;
  }
  int count = 0;
  for (int64_t base = 0; base < a.batch; base += kSegThreads) {
    const int64_t i = base + tid;
    const bool mz = i < a.batch && a.zs[i] == c;
    const bool mp = i < a.batch && a.preds[i] == c;
    const unsigned bz = __ballot_sync(0xffffffffu, mz);
    const unsigned bp = __ballot_sync(0xffffffffu, mp);
    __syncthreads();           // previous chunk fully consumed
    if (lane == 0) { mz_s[warp] = bz; mp_s[warp] = bp; }
    __syncthreads();
    for (int w = 0; w < kSegThreads / 32; ++w) {
      const unsigned z = mz_s[w], pm = mp_s[w];
      unsigned any = z | pm;
      count += __popc(z) + __popc(pm);
      while (any) {
        const int b = __ffs(any) - 1;
        any &= any - 1;
        const int64_t ii = base + w * 32 + b;
        if (z & (1u << b)) {
          float sc = 1.f;
          if (a.w_lbl) {
            sc = a.alpha / (a.w_lbl_zs[ii] + a.w_img[ii] + a.w_grp[ii]);
            if (!quirk) sc *= a.w_lbl[ii];
          }
#pragma unroll
          for (int j = 0; j < kSegMaxCols; ++j) {
            const int64_t d = tid + (int64_t)j * kSegThreads;
            if (d < a.dim) acc[j] += ld1(a.src_zs, ii * a.ld_src + d, a.dtype) * sc * colw[j];
          }
        }
        if (pm & (1u << b)) {
          float sc = 1.f;
          if (a.w_lbl) {
            sc = a.alpha / (a.w_lbl[ii] + a.w_img[ii] + a.w_grp[ii]);
            if (!quirk) sc *= a.w_lbl[ii];
          }
#pragma unroll
          for (int j = 0; j < kSegMaxCols; ++j) {
            const int64_t d = tid + (int64_t)j * kSegThreads;
            if (d < a.dim) acc[j] += ld1(a.src_ft, ii * a.ld_src + d, a.dtype) * sc * colw[j];
          }
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < kSegMaxCols; ++j) {
    const int64_t d = tid + (int64_t)j * kSegThreads;
    if (d < a.dim) {
      float* o = a.out + (int64_t)c * a.ld_out + d;
      const float v = acc[j] * a.post_scale;
      *o = a.accumulate ? *o + v : v;
    }
  }
  if (a.counts && tid == 0) a.counts[c] = (float)count;
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

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace
}  // namespace latte

using namespace latte;

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
            ld_out, 0};
  const int64_t esz = (int64_t)dtype_size(dtype);
  const int64_t vb = 4 * esz;   // bytes per 4-element vector
  a.vec = (dim % 4 == 0) && (ld_ct % 4 == 0) && (ld_pi % 4 == 0) && (ld_pg % 4 == 0) &&
          (ld_bank % 4 == 0) && (ld_out % 4 == 0) && al16(bank) && al16(w_lbl) &&
          (reinterpret_cast<uintptr_t>(class_text) % vb == 0) &&
          (reinterpret_cast<uintptr_t>(per_image) % vb == 0) &&
          (reinterpret_cast<uintptr_t>(per_group) % vb == 0) &&
          (reinterpret_cast<uintptr_t>(t_ft) % vb == 0) && (reinterpret_cast<uintptr_t>(t_zs) % vb == 0);
  mix_ema_fwd_kernel<<<(unsigned)batch, 128, 0, static_cast<cudaStream_t>(stream)>>>(a);
  LATTE_LAUNCH_OK();
  return LATTE_OK;
}

extern "C" int latte_mix_ema_bwd(const void* d_t_ft, const void* d_t_zs, int64_t ld_dt,
                                 const int64_t* preds, const int64_t* zs, const float* w_lbl,
                                 const float* w_lbl_zs, const float* w_img, const float* w_grp,
                                 float alpha, int label_axis, int dtype, int64_t batch,
                                 int64_t dim, int64_t num_classes, float* d_class_text,
                                 int64_t ld_dct, void* d_per_image, void* d_per_group,
                                 int64_t ld_dp, float* d_bank, int64_t ld_dbank, void* stream) {
  LATTE_CHECK_ARG(d_t_ft && d_t_zs && preds && zs && w_lbl && w_lbl_zs && w_img && w_grp);
  LATTE_CHECK_ARG(batch >= 0 && dim > 0 && num_classes > 0);
  LATTE_CHECK_ARG(dtype >= LATTE_F32 && dtype <= LATTE_F16);
  if (label_axis == LATTE_LABEL_AXIS_QUIRK && batch != dim) return LATTE_ERR_UNSUPPORTED;
  if (dim > (int64_t)kSegThreads * kSegMaxCols) return LATTE_ERR_UNSUPPORTED;
  if (batch == 0) return LATTE_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (d_per_image && d_per_group) {
    MixBwdArgs a{d_t_ft, d_t_zs, ld_dt, w_lbl, w_lbl_zs, w_img, w_grp, alpha, dtype, batch, dim,
                 d_per_image, d_per_group, ld_dp};
    mix_ema_bwd_rows_kernel<<<(unsigned)batch, 128, 0, st>>>(a);
    LATTE_LAUNCH_OK();
  }
  if (d_class_text) {
    // d_class_text[c] += sum_{zs_i=c} wl * alpha/totz_i * dT_zs[i] + sum_{preds_i=c} wl * alpha/tot_i * dT_ft[i]
    SegArgs s{d_t_ft, d_t_zs, ld_dt, dtype, preds, zs, batch, dim,
              w_lbl, w_lbl_zs, w_img, w_grp, alpha, label_axis,
              d_class_text, ld_dct, 1, nullptr, 1.f};
    segment_sum_kernel<<<(unsigned)num_classes, kSegThreads, 0, st>>>(s);
    LATTE_LAUNCH_OK();
  }
  if (d_bank) {
    // (1 - alpha) * d_t scattered by class (train.py:487-488 wrt membank_features)
    SegArgs s{d_t_ft, d_t_zs, ld_dt, dtype, preds, zs, batch, dim,
              nullptr, nullptr, nullptr, nullptr, 0.f, LATTE_LABEL_AXIS_ROW,
              d_bank, ld_dbank, 1, nullptr, 1.f - alpha};
    segment_sum_kernel<<<(unsigned)num_classes, kSegThreads, 0, st>>>(s);
    LATTE_LAUNCH_OK();
  }
  return LATTE_OK;
}

extern "C" int latte_bank_accumulate(const void* t_ft, const void* t_zs, int64_t ld_t, int dtype,
                                     const int64_t* preds, const int64_t* zs, int64_t batch,
                                     int64_t dim, int64_t num_classes, float* sums,
                                     int64_t ld_sums, float* counts, void* stream) {
  LATTE_CHECK_ARG(t_ft && t_zs && preds && zs && sums && counts);
  LATTE_CHECK_ARG(batch >= 0 && dim > 0 && num_classes > 0 && ld_t >= dim && ld_sums >= dim);
  LATTE_CHECK_ARG(dtype >= LATTE_F32 && dtype <= LATTE_F16);
  if (dim > (int64_t)kSegThreads * kSegMaxCols) return LATTE_ERR_UNSUPPORTED;
  SegArgs s{t_ft, t_zs, ld_t, dtype, preds, zs, batch, dim,
            nullptr, nullptr, nullptr, nullptr, 0.f, LATTE_LABEL_AXIS_ROW,
            sums, ld_sums, 0, counts, 1.f};
  segment_sum_kernel<<<(unsigned)num_classes, kSegThreads, 0, static_cast<cudaStream_t>(stream)>>>(s);
  LATTE_LAUNCH_OK();
  return LATTE_OK;
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

// C ABI of the SigLIP loss (open_clip loss.py:453-560) on the CTA-pair tile engine of clip_pair.cu:
// one logit sweep of this rank's images against ALL texts per pass (the reference reaches the same
// pairs by passing text shards round the ring, loss.py:521-558), nothing of the [n, N] logit matrix
// stored.  Forward: sum of -logsigmoid(label * z).  Backward: G = sigmoid(z) - delta as fp16 blocks,
// then the same stream-K gradient GEMMs as ClipLoss (d_img = G . T, d_txt (partial) = G^T . I).
#include "latte_common.cuh"

namespace latte {
namespace {

struct SigLayout {
  size_t off_part0, off_part1, off_scale, off_x16, off_y16, off_g, off_acc0, off_acc1;
  size_t ld16, ld32, parts;
  size_t total;
};

SigLayout sig_layout(int64_t n_loc, int64_t n_all, int64_t dim, int dtype, bool bwd, bool own_dtxt) {
  SigLayout w;
  auto up = [](size_t x) { return (x + 63) / 64 * 64; };
  w.parts = up((size_t)clip_pair_ds_count());
  w.ld16 = ((size_t)dim + 7) / 8 * 8;
  w.ld32 = ((size_t)dim + 3) / 4 * 4;
  size_t o = 0;
  w.off_part0 = o; o += w.parts;
  w.off_part1 = o; o += w.parts;
  w.off_scale = o; o += 64;
  w.off_x16 = w.off_y16 = w.off_g = w.off_acc0 = w.off_acc1 = o;
  if (bwd) {
    if (dtype == LATTE_BF16) {
      w.off_x16 = o; o += up(((size_t)n_loc * w.ld16 + 1) / 2);
      w.off_y16 = o; o += up(((size_t)n_all * w.ld16 + 1) / 2);
    }
    const PairGeom geo = clip_pair_geom(n_loc, n_all);
    w.off_g = o; o += up((geo.g_elems + 1) / 2);
    w.off_acc0 = o; o += up((size_t)n_loc * w.ld32);
    if (own_dtxt) { w.off_acc1 = o; o += up((size_t)n_all * w.ld32); }
  }
  w.total = o * sizeof(float);
  return w;
}

__global__ void __launch_bounds__(256)
sig_loss_kernel(const float* partial, int count, int64_t n_loc, float* loss) {
  __shared__ double red[8];
  double acc = 0.0;
  for (int i = threadIdx.x; i < count; i += 256) acc += (double)partial[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < 8; ++w) tot += red[w];
    *loss = (float)(tot / (double)n_loc);            // loss.py:519: sum / image_features.shape[0]
  }
}

// out_scale = grad_loss * s / (n * 2^13): turns the GEMM accumulators (G scaled by 2^13) into
// feature gradients
__global__ void sig_scale_kernel(const float* grad_loss, const float* logit_scale, int64_t n_loc,
                                 float* out_scale) {
  *out_scale = __ldg(grad_loss) * __ldg(logit_scale) / ((float)n_loc * 8192.0f);
}

__global__ void __launch_bounds__(256)
sig_scalar_grads_kernel(const float* ds_partial, const float* db_partial, int count,
                        const float* grad_loss, int64_t n_loc, float* d_scale, float* d_bias) {
  __shared__ double red[2][8];
  double a0 = 0.0, a1 = 0.0;
  for (int i = threadIdx.x; i < count; i += 256) {
    a0 += (double)ds_partial[i];
    a1 += (double)db_partial[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a0; red[1][threadIdx.x >> 5] = a1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t0 = 0.0, t1 = 0.0;
    for (int w = 0; w < 8; ++w) { t0 += red[0][w]; t1 += red[1][w]; }
    const double c = (double)__ldg(grad_loss) / (double)n_loc;
    if (d_scale) *d_scale = (float)(t0 * c);
    if (d_bias) *d_bias = (float)(t1 * c);
  }
}

__global__ void __launch_bounds__(256)
sig_to_fp16_kernel(const __nv_bfloat16* in, __half* out, int64_t ld_in, int64_t ld_out, int64_t rows,
                   int64_t dim) {
  const int64_t per_row = dim / 8;                   // dim % 8 == 0 on this path
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * per_row) return;
  const int64_t r = idx / per_row, c = (idx % per_row) * 8;
  const uint4 raw = *reinterpret_cast<const uint4*>(in + r * ld_in + c);
  const __nv_bfloat162* v = reinterpret_cast<const __nv_bfloat162*>(&raw);
  uint4 o;
  __half2* h = reinterpret_cast<__half2*>(&o);
#pragma unroll
  for (int k = 0; k < 4; ++k) h[k] = __float22half2_rn(__bfloat1622float2(v[k]));
  *reinterpret_cast<uint4*>(out + r * ld_out + c) = o;
}

float* align_ws(void* workspace) {
  return reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) / 256 * 256);
}

}  // namespace
}  // namespace latte

using namespace latte;

extern "C" int latte_siglip_supported(int dtype, int64_t dim) {
  return ((dtype == LATTE_BF16 || dtype == LATTE_F16) && dim >= 8 && dim <= 768 && dim % 8 == 0) ? 1 : 0;
}

extern "C" int latte_siglip_workspace_bytes(int64_t n_loc, int64_t n_all, int64_t dim, int dtype,
                                            int backward, int own_d_txt, size_t* bytes) {
  LATTE_CHECK_ARG(bytes && n_loc > 0 && n_all >= n_loc && dim > 0);
  if (!latte_siglip_supported(dtype, dim)) return LATTE_ERR_UNSUPPORTED;
  *bytes = sig_layout(n_loc, n_all, dim, dtype, backward != 0, own_d_txt != 0).total + 256;
  return LATTE_OK;
}

extern "C" int latte_siglip_fwd(const void* img_loc, int64_t ld_img, const void* txt_all, int64_t ld_txt,
                                int dtype, int64_t n_loc, int64_t n_all, int64_t dim,
                                int64_t label_offset, const float* logit_scale,
                                const float* logit_bias, float* loss, void* workspace,
                                size_t workspace_bytes, void* stream) {
  LATTE_CHECK_ARG(img_loc && txt_all && logit_scale && loss && workspace);
  LATTE_CHECK_ARG(n_loc > 0 && n_all >= n_loc && dim > 0 && label_offset >= 0 &&
                  label_offset + n_loc <= n_all);
  if (!latte_siglip_supported(dtype, dim)) return LATTE_ERR_UNSUPPORTED;
  const SigLayout w = sig_layout(n_loc, n_all, dim, dtype, false, false);
  if (workspace_bytes < w.total + 256) return LATTE_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* ws = align_ws(workspace);
  const int parts = clip_pair_ds_count();
  LATTE_CUDA_OK(cudaMemsetAsync(ws + w.off_part0, 0, (size_t)parts * sizeof(float), st));
  PairSigArgs a = {};
  a.x = img_loc; a.ldx = ld_img; a.y = txt_all; a.ldy = ld_txt; a.dtype = dtype;
  a.n_loc = n_loc; a.n_all = n_all; a.dim = dim; a.label_offset = label_offset;
  a.logit_scale = logit_scale; a.logit_bias = logit_bias;
  a.g = nullptr; a.ds_partial = nullptr; a.aux_partial = ws + w.off_part0;
  int rc = clip_pair_sig_sweep(a, st);
  if (rc) return rc;
  sig_loss_kernel<<<1, 256, 0, st>>>(ws + w.off_part0, parts, n_loc, loss);
  LATTE_LAUNCH_OK();
  return LATTE_OK;
}

// d_txt: [n_all, dim] in grad_dtype when this rank owns every text row (world size 1), else NULL and
// exactly one of d_txt_partial (fp32 [n_all, dim], zeroed here, for the caller's reduce-scatter) or
// d_txt_peers (per-rank fp32 accumulators [n_all / n_peers, dim], zeroed and fenced by the caller).
extern "C" int latte_siglip_bwd(const void* img_loc, int64_t ld_img, const void* txt_all, int64_t ld_txt,
                                int dtype, int64_t n_loc, int64_t n_all, int64_t dim,
                                int64_t label_offset, const float* logit_scale,
                                const float* logit_bias, const float* grad_loss, void* d_img,
                                void* d_txt, int grad_dtype, int64_t ld_grad, float* d_txt_partial,
                                void* const* d_txt_peers, int n_peers, float* d_scale, float* d_bias,
                                void* workspace, size_t workspace_bytes, void* stream) {
  LATTE_CHECK_ARG(img_loc && txt_all && logit_scale && grad_loss && d_img && workspace);
  LATTE_CHECK_ARG(n_loc > 0 && n_all >= n_loc && dim > 0 && label_offset >= 0 &&
                  label_offset + n_loc <= n_all);
  LATTE_CHECK_ARG((d_txt != nullptr) + (d_txt_partial != nullptr) + (d_txt_peers != nullptr) == 1);
  LATTE_CHECK_ARG(!d_txt || n_loc == n_all);
  LATTE_CHECK_ARG(!d_txt_peers || (n_peers > 1 && n_peers <= 8 && n_all % n_peers == 0));
  if (!latte_siglip_supported(dtype, dim)) return LATTE_ERR_UNSUPPORTED;
  if ((ld_img % 8) || (ld_txt % 8)) return LATTE_ERR_UNSUPPORTED;
  const bool own = d_txt != nullptr;
  const SigLayout w = sig_layout(n_loc, n_all, dim, dtype, true, own);
  if (workspace_bytes < w.total + 256) return LATTE_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* ws = align_ws(workspace);
  const int parts = clip_pair_ds_count();
  float* out_scale = ws + w.off_scale;
  float* acc_i = ws + w.off_acc0;
  float* acc_t = ws + w.off_acc1;
  LATTE_CUDA_OK(cudaMemsetAsync(ws + w.off_part0, 0, 2 * w.parts * sizeof(float), st));
  if (d_txt_partial)
    LATTE_CUDA_OK(cudaMemsetAsync(d_txt_partial, 0, (size_t)n_all * (size_t)dim * sizeof(float), st));
  sig_scale_kernel<<<1, 1, 0, st>>>(grad_loss, logit_scale, n_loc, out_scale);
  LATTE_LAUNCH_OK();
  // fp16 operands of the gradient GEMMs (G is fp16 and kind::f16 takes one 16-bit format)
  const void* x16 = img_loc; int64_t ldx16 = ld_img;
  const void* y16 = txt_all; int64_t ldy16 = ld_txt;
  if (dtype == LATTE_BF16) {
    __half* xh = reinterpret_cast<__half*>(ws + w.off_x16);
    __half* yh = reinterpret_cast<__half*>(ws + w.off_y16);
    const int64_t per_row = dim / 8;
    sig_to_fp16_kernel<<<(unsigned)((n_loc * per_row + 255) / 256), 256, 0, st>>>(
        static_cast<const __nv_bfloat16*>(img_loc), xh, ld_img, (int64_t)w.ld16, n_loc, dim);
    LATTE_LAUNCH_OK();
    sig_to_fp16_kernel<<<(unsigned)((n_all * per_row + 255) / 256), 256, 0, st>>>(
        static_cast<const __nv_bfloat16*>(txt_all), yh, ld_txt, (int64_t)w.ld16, n_all, dim);
    LATTE_LAUNCH_OK();
    x16 = xh; ldx16 = (int64_t)w.ld16;
    y16 = yh; ldy16 = (int64_t)w.ld16;
  }
  PairSigArgs a = {};
  a.x = img_loc; a.ldx = ld_img; a.y = txt_all; a.ldy = ld_txt; a.dtype = dtype;
  a.n_loc = n_loc; a.n_all = n_all; a.dim = dim; a.label_offset = label_offset;
  a.logit_scale = logit_scale; a.logit_bias = logit_bias;
  a.g = ws + w.off_g; a.ds_partial = ws + w.off_part0; a.aux_partial = ws + w.off_part1;
  int rc = clip_pair_sig_sweep(a, st);
  if (rc) return rc;
  PairGemmArgs ga = {};
  ga.g = ws + w.off_g; ga.n_loc = n_loc; ga.n_all = n_all; ga.dim = dim; ga.ld32 = (int64_t)w.ld32;
  ga.y16 = y16; ga.ldy16 = ldy16; ga.x16 = x16; ga.ldx16 = ldx16;
  ga.feat_dtype = LATTE_F16;
  ga.out_dtype = grad_dtype; ga.ld_out = ld_grad; ga.out_scale = out_scale;
  ga.dx_out = d_img; ga.dy_out = own ? d_txt : nullptr;
  ga.dx32 = acc_i;
  ga.dy32 = acc_t; ga.ld_dy32 = (int64_t)w.ld32; ga.dy_scale = nullptr;
  ga.dy_peers = nullptr; ga.n_peers = 0;
  if (d_txt_partial) {
    ga.dy32 = d_txt_partial; ga.ld_dy32 = dim; ga.dy_scale = out_scale;
  } else if (d_txt_peers) {
    ga.dy32 = static_cast<float*>(d_txt_peers[0]); ga.ld_dy32 = dim; ga.dy_scale = out_scale;
    ga.dy_peers = reinterpret_cast<float* const*>(d_txt_peers);
    ga.n_peers = n_peers;
  }
  const bool direct_i = clip_pair_gemm_direct(ga, 0);
  const bool direct_t = own && clip_pair_gemm_direct(ga, 1);
  if (!direct_i) LATTE_CUDA_OK(cudaMemsetAsync(acc_i, 0, (size_t)n_loc * w.ld32 * sizeof(float), st));
  if (own && !direct_t)
    LATTE_CUDA_OK(cudaMemsetAsync(acc_t, 0, (size_t)n_all * w.ld32 * sizeof(float), st));
  rc = clip_pair_gemm_fixup(ga, 0, st);
  if (rc) return rc;
  rc = clip_pair_gemm(ga, st);
  if (rc) return rc;
  rc = clip_pair_gemm_fixup(ga, 1, st);
  if (rc) return rc;
  if (!direct_i) {
    rc = clip_pair_scale_cast(acc_i, nullptr, (int64_t)w.ld32, d_img, nullptr, grad_dtype, ld_grad,
                              n_loc, dim, out_scale, st);
    if (rc) return rc;
  }
  if (own && !direct_t) {
    rc = clip_pair_scale_cast(acc_t, nullptr, (int64_t)w.ld32, d_txt, nullptr, grad_dtype, ld_grad,
                              n_all, dim, out_scale, st);
    if (rc) return rc;
  }
  sig_scalar_grads_kernel<<<1, 256, 0, st>>>(ws + w.off_part0, ws + w.off_part1, parts, grad_loss,
                                             n_loc, d_scale, d_bias);
  LATTE_LAUNCH_OK();
  return LATTE_OK;
}

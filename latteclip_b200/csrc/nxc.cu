// N x C prototype similarity with fused row reductions (argmax, top-2 margin, top-k).
// Replaces train.py:410-411 (100*I@classifier, argmax), compute_text_weights
// (train.py:292-303: bmm + topk(2)) and zero_shot.py:14-20,40 (logits.topk) -- the
// [N, C] logit matrix is never written to HBM.  fp32 products and accumulation so that
// pseudo-labels are exact wherever the reference has no tie; ties resolve to the lowest
// class index like torch.argmax / torch.topk.
#include "latte_common.cuh"

namespace latte {
namespace {

constexpr int kR = 64;       // rows per CTA
constexpr int kC = 64;       // classes per tile
constexpr int kK = 16;       // feature chunk
constexpr int kT = 256;
constexpr int kMaxTopK = 16;

__device__ __forceinline__ float ldx(const void* base, int64_t idx, int dtype) {
  if (dtype == LATTE_F32) return __ldg(reinterpret_cast<const float*>(base) + idx);
  if (dtype == LATTE_BF16)
    return __bfloat162float(__ldg(reinterpret_cast<const __nv_bfloat16*>(base) + idx));
  return __half2float(__ldg(reinterpret_cast<const __half*>(base) + idx));
}

struct NxcArgs {
  const void* x; int64_t ldx; int dtype;
  const int64_t* row_index;
  int64_t n, dim;
  const float* protos; int64_t ldp; int64_t num_classes;
  float scale;
  int64_t* argmax_out; float* margin_out; float* top1_out;
  int k; int64_t* topk_idx; float* topk_val;
};

template <bool kTopK>
__global__ void __launch_bounds__(kT) nxc_kernel(NxcArgs a) {
  __shared__ float xs[kR][kK + 1];
  __shared__ float ps[kC][kK + 1];
  __shared__ float st[kR][kC + 1];
  __shared__ int64_t src_row[kR];
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const int64_t row0 = (int64_t)blockIdx.x * kR;

  if (tid < kR) {
    const int64_t gr = row0 + tid;
    int64_t s = gr < a.n ? gr : -1;
    if (s >= 0 && a.row_index) s = a.row_index[gr];
    src_row[tid] = s;
  }
  __syncthreads();

  // per-row running state, owned by threads 0..63
  float v1 = -INFINITY, v2 = -INFINITY;
  int64_t i1 = 0;
  float tv[kTopK ? kMaxTopK : 1];
  int ti[kTopK ? kMaxTopK : 1];
  if (kTopK) {
#pragma unroll
    for (int j = 0; j < kMaxTopK; ++j) { tv[j] = -INFINITY; ti[j] = 0x7fffffff; }
  }

  for (int64_t c0 = 0; c0 < a.num_classes; c0 += kC) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int64_t k0 = 0; k0 < a.dim; k0 += kK) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int idx = tid + e * kT;
        const int r = idx >> 4, k = idx & 15;
        const int64_t gk = k0 + k;
        const int64_t sr = src_row[r];
        const int64_t gc = c0 + r;
        xs[r][k] = (sr >= 0 && gk < a.dim) ? ldx(a.x, sr * a.ldx + gk, a.dtype) : 0.f;
        ps[r][k] = (gc < a.num_classes && gk < a.dim) ? __ldg(a.protos + gc * a.ldp + gk) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < kK; ++k) {
        float xa[4], pb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) xa[i] = xs[ty * 4 + i][k];
#pragma unroll
        for (int j = 0; j < 4; ++j) pb[j] = ps[tx * 4 + j][k];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xa[i], pb[j], acc[i][j]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) st[ty * 4 + i][tx * 4 + j] = acc[i][j];
    __syncthreads();
    if (tid < kR) {
      const int lim = (int)min((int64_t)kC, a.num_classes - c0);
      for (int j = 0; j < lim; ++j) {
        const float v = st[tid][j];
        if (kTopK) {
          // insert (v, c0 + j) into the descending list; equal values keep the lower index first
          if (v > tv[a.k - 1]) {
            int pos = a.k - 1;
            while (pos > 0 && v > tv[pos - 1]) { tv[pos] = tv[pos - 1]; ti[pos] = ti[pos - 1]; --pos; }
            tv[pos] = v; ti[pos] = (int)(c0 + j);
          }
        } else {
          if (v > v1) { v2 = v1; v1 = v; i1 = c0 + j; }
          else if (v > v2) { v2 = v; }
        }
      }
    }
    __syncthreads();
  }
  if (tid < kR) {
    const int64_t gr = row0 + tid;
    if (gr < a.n) {
      if (kTopK) {
        for (int j = 0; j < a.k; ++j) {
          a.topk_idx[gr * a.k + j] = ti[j];
          a.topk_val[gr * a.k + j] = a.scale * tv[j];
        }
      } else {
        if (a.argmax_out) a.argmax_out[gr] = i1;
        if (a.margin_out) a.margin_out[gr] = v1 - v2;
        if (a.top1_out) a.top1_out[gr] = a.scale * v1;
      }
    }
  }
}

}  // namespace
}  // namespace latte

namespace latte {
// tcgen05 path with operands split into 16-bit planes (nxc_tc.cu); LATTE_ERR_UNSUPPORTED -> use the SIMT tiles
int nxc_tc_run(const void* x, int64_t ldx, int x_dtype, const int64_t* row_index, int64_t n,
               int64_t dim, const float* protos, int64_t ldp, int64_t num_classes, float scale,
               int64_t* argmax_out, float* margin_out, float* top1_out, int k, int64_t* topk_idx,
               float* topk_val, void* workspace, size_t workspace_bytes, cudaStream_t st);
size_t nxc_tc_workspace_bytes(int x_dtype, bool gathered, int64_t n, int64_t dim, int64_t num_classes);
constexpr int64_t kTcMinRows = 128;      // below this the launch of the plane split dominates
}  // namespace latte

using namespace latte;

extern "C" int latte_nxc_workspace_bytes(int x_dtype, int gathered, int64_t n, int64_t dim,
                                         int64_t num_classes, size_t* bytes) {
  LATTE_CHECK_ARG(bytes && n >= 0 && dim > 0 && num_classes > 0);
  LATTE_CHECK_ARG(x_dtype >= LATTE_F32 && x_dtype <= LATTE_F16);
  *bytes = n >= kTcMinRows ? nxc_tc_workspace_bytes(x_dtype, gathered != 0, n, dim, num_classes) : 0;
  return LATTE_OK;
}

extern "C" int latte_nxc_argmax_margin(const void* x, int64_t ldx_, int x_dtype,
                                       const int64_t* row_index, int64_t n, int64_t dim,
                                       const float* protos, int64_t ldp, int64_t num_classes,
                                       float scale, int64_t* argmax_out, float* margin_out,
                                       float* top1_out, void* workspace, size_t workspace_bytes,
                                       void* stream) {
  LATTE_CHECK_ARG(x && protos && n >= 0 && dim > 0 && num_classes > 0);
  LATTE_CHECK_ARG(x_dtype >= LATTE_F32 && x_dtype <= LATTE_F16);
  LATTE_CHECK_ARG(ldx_ >= dim && ldp >= dim);
  if (n == 0) return LATTE_OK;
  if (n >= kTcMinRows) {
    const int rc = nxc_tc_run(x, ldx_, x_dtype, row_index, n, dim, protos, ldp, num_classes, scale,
                              argmax_out, margin_out, top1_out, 0, nullptr, nullptr, workspace,
                              workspace_bytes, static_cast<cudaStream_t>(stream));
    if (rc != LATTE_ERR_UNSUPPORTED) return rc;
  }
  NxcArgs a{x, ldx_, x_dtype, row_index, n, dim, protos, ldp, num_classes, scale,
            argmax_out, margin_out, top1_out, 0, nullptr, nullptr};
  dim3 grid((unsigned)((n + kR - 1) / kR));
  nxc_kernel<false><<<grid, kT, 0, static_cast<cudaStream_t>(stream)>>>(a);
  LATTE_LAUNCH_OK();
  return LATTE_OK;
}

extern "C" int latte_nxc_topk(const void* x, int64_t ldx_, int x_dtype, int64_t n, int64_t dim,
                              const float* protos, int64_t ldp, int64_t num_classes, float scale,
                              int k, int64_t* topk_idx, float* topk_val, void* workspace,
                              size_t workspace_bytes, void* stream) {
  LATTE_CHECK_ARG(x && protos && topk_idx && topk_val && n >= 0 && dim > 0 && num_classes > 0);
  LATTE_CHECK_ARG(x_dtype >= LATTE_F32 && x_dtype <= LATTE_F16);
  LATTE_CHECK_ARG(ldx_ >= dim && ldp >= dim);
  if (k < 1 || k > kMaxTopK || k > num_classes) return LATTE_ERR_UNSUPPORTED;
  if (n == 0) return LATTE_OK;
  if (n >= kTcMinRows) {
    const int rc = nxc_tc_run(x, ldx_, x_dtype, nullptr, n, dim, protos, ldp, num_classes, scale,
                              nullptr, nullptr, nullptr, k, topk_idx, topk_val, workspace,
                              workspace_bytes, static_cast<cudaStream_t>(stream));
    if (rc != LATTE_ERR_UNSUPPORTED) return rc;
  }
  NxcArgs a{x, ldx_, x_dtype, nullptr, n, dim, protos, ldp, num_classes, scale,
            nullptr, nullptr, nullptr, k, topk_idx, topk_val};
  dim3 grid((unsigned)((n + kR - 1) / kR));
  nxc_kernel<true><<<grid, kT, 0, static_cast<cudaStream_t>(stream)>>>(a);
  LATTE_LAUNCH_OK();
  return LATTE_OK;
}

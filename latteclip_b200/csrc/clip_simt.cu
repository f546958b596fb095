// ClipLoss row kernels, fp32 SIMT path: exact fp32 products and accumulation for fp32
// features (and for 16-bit features whose layout the TMA path cannot take).  Same
// contract as clip_tc.cu; logits are never stored.  open_clip/loss.py:109-116,126-129.
#include "latte_common.cuh"

namespace latte {
namespace {

constexpr int kRows = 64;     // x rows per CTA
constexpr int kCols = 64;     // y rows per tile
constexpr int kKc = 16;       // feature chunk
constexpr int kSlab = 128;    // dX feature columns per CTA in the backward
constexpr int kThreadsSimt = 256;

__device__ __forceinline__ float ld_elem(const void* base, int64_t idx, int dtype) {
  if (dtype == LATTE_F32) return __ldg(reinterpret_cast<const float*>(base) + idx);
  if (dtype == LATTE_BF16)
    return __bfloat162float(__ldg(reinterpret_cast<const __nv_bfloat16*>(base) + idx));
  return __half2float(__ldg(reinterpret_cast<const __half*>(base) + idx));
}

__device__ __forceinline__ void st_elem(void* base, int64_t idx, int dtype, float v) {
  if (dtype == LATTE_F32) reinterpret_cast<float*>(base)[idx] = v;
  else if (dtype == LATTE_BF16) reinterpret_cast<__nv_bfloat16*>(base)[idx] = __float2bfloat16_rn(v);
  else reinterpret_cast<__half*>(base)[idx] = __float2half_rn(v);
}

// S tile (64 x 64) of x_blk . y_tile^T; thread (ty, tx) owns rows 4ty..4ty+3, cols 4tx..4tx+3.
__device__ __forceinline__ void s_tile(const void* x, int64_t ldx, const void* y, int64_t ldy,
                                       int dtype, int64_t row0, int64_t col0, int64_t n_loc,
                                       int64_t n_all, int64_t dim, float (*xs)[kKc + 1],
                                       float (*ys)[kKc + 1], float (&acc)[4][4]) {
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  for (int64_t k0 = 0; k0 < dim; k0 += kKc) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * kThreadsSimt;   // 0..1023
      const int r = idx >> 4, k = idx & 15;
      const int64_t gk = k0 + k;
      const int64_t gr = row0 + r, gc = col0 + r;
      xs[r][k] = (gr < n_loc && gk < dim) ? ld_elem(x, gr * ldx + gk, dtype) : 0.f;
      ys[r][k] = (gc < n_all && gk < dim) ? ld_elem(y, gc * ldy + gk, dtype) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kKc; ++k) {
      float xa[4], yb[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) xa[a] = xs[ty * 4 + a][k];
#pragma unroll
      for (int b = 0; b < 4; ++b) yb[b] = ys[tx * 4 + b][k];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(xa[a], yb[b], acc[a][b]);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kThreadsSimt)
clip_fwd_simt_kernel(ClipFwdArgs a) {
  __shared__ float xs[kRows][kKc + 1];
  __shared__ float ys[kCols][kKc + 1];
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const int64_t row0 = (int64_t)blockIdx.x * kRows;
  const float c2 = __ldg(a.logit_scale) * kLog2e;
  float m[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { m[i] = -INFINITY; l[i] = 0.f; }
  for (int64_t col0 = 0; col0 < a.n_all; col0 += kCols) {
    float acc[4][4];
    s_tile(a.x, a.ldx, a.y, a.ldy, a.dtype, row0, col0, a.n_loc, a.n_all, a.dim, xs, ys, acc);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t grow = row0 + ty * 4 + i;
      const int64_t label = a.label_offset + grow;
      float tmax = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t gc = col0 + tx * 4 + j;
        if (gc >= a.n_all) acc[i][j] = -INFINITY;
        if (gc == label && grow < a.n_loc) a.diag[grow] = acc[i][j];
        tmax = fmaxf(tmax, acc[i][j]);
      }
      const float m_new = fmaxf(m[i], tmax * c2);
      if (m_new > -INFINITY) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) s += exp2f(fmaf(acc[i][j], c2, -m_new));
        l[i] = l[i] * exp2f(m[i] - m_new) + s;
        m[i] = m_new;
      }
    }
  }
  // combine the 16 threads (tx) that share a row: they are 16 consecutive lanes
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      const float mo = __shfl_xor_sync(0xffffffffu, m[i], o);
      const float lo = __shfl_xor_sync(0xffffffffu, l[i], o);
      const float mn = fmaxf(m[i], mo);
      if (mn > -INFINITY) {
        l[i] = l[i] * exp2f(m[i] - mn) + lo * exp2f(mo - mn);
        m[i] = mn;
      }
    }
    const int64_t grow = row0 + ty * 4 + i;
    if (tx == 0 && grow < a.n_loc) {
      a.part_max[grow] = m[i];
      a.part_sum[grow] = l[i];
    }
  }
}

__global__ void __launch_bounds__(kThreadsSimt)
clip_bwd_simt_kernel(ClipBwdArgs a, float cb, float cd) {
  extern __shared__ float sm[];
  float (*xs)[kKc + 1] = reinterpret_cast<float (*)[kKc + 1]>(sm);
  float (*ys)[kKc + 1] = reinterpret_cast<float (*)[kKc + 1]>(sm + kRows * (kKc + 1));
  float (*gs)[kCols + 1] = reinterpret_cast<float (*)[kCols + 1]>(sm + 2 * kRows * (kKc + 1));
  float (*yd)[kSlab] =
      reinterpret_cast<float (*)[kSlab]>(sm + 2 * kRows * (kKc + 1) + kRows * (kCols + 1));
  __shared__ float red[kThreadsSimt / 32];

  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const int64_t row0 = (int64_t)blockIdx.x * kRows;
  const int64_t d_base = (int64_t)blockIdx.y * kSlab;
  const float s = __ldg(a.logit_scale);
  const float c2 = s * kLog2e;

  float a2[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t grow = row0 + ty * 4 + i;
    a2[i] = grow < a.n_loc ? __ldg(a.lse_a2 + a.label_offset + grow) : 0.f;
  }
  float dacc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) dacc[i][j] = 0.f;
  float ds_acc = 0.f;

  for (int64_t col0 = 0; col0 < a.n_all; col0 += kCols) {
    float acc[4][4];
    s_tile(a.x, a.ldx, a.y, a.ldy, a.dtype, row0, col0, a.n_loc, a.n_all, a.dim, xs, ys, acc);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t grow = row0 + ty * 4 + i;
      const int64_t label = a.label_offset + grow;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t gc = col0 + tx * 4 + j;
        float g = 0.f;
        if (gc < a.n_all && grow < a.n_loc) {
          const float v = acc[i][j];
          const float ea = exp2f(fmaf(v, c2, -a2[i]));
          const float eb = exp2f(fmaf(v, c2, -__ldg(a.lse_b2 + gc)));
          g = fmaf(cb, eb, ea);
          ds_acc = fmaf(ea, v, ds_acc);
          if (gc == label) { g -= cd; ds_acc -= v; }
        }
        gs[ty * 4 + i][tx * 4 + j] = g;
      }
    }
    // y tile slab [64 j][128 d]
    for (int e = tid; e < kCols * kSlab; e += kThreadsSimt) {
      const int jj = e / kSlab, dd = e % kSlab;
      const int64_t gj = col0 + jj, gd = d_base + dd;
      yd[jj][dd] = (gj < a.n_all && gd < a.dim) ? ld_elem(a.y, gj * a.ldy + gd, a.dtype) : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int jj = 0; jj < kCols; ++jj) {
      float gv[4], yv[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) gv[i] = gs[ty * 4 + i][jj];
#pragma unroll
      for (int j = 0; j < 8; ++j) yv[j] = yd[jj][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) dacc[i][j] = fmaf(gv[i], yv[j], dacc[i][j]);
    }
    __syncthreads();
  }

  const float coef = __ldg(a.grad_loss) * a.grad_mult / (2.0f * (float)a.n_loc);
  const float cs = coef * s;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t grow = row0 + ty * 4 + i;
    if (grow >= a.n_loc) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int64_t gd = d_base + tx + 16 * j;
      if (gd < a.dim) st_elem(a.dx, grow * a.ld_dx + gd, a.grad_dtype, cs * dacc[i][j]);
    }
  }
  if (blockIdx.y == 0) {
    float v = ds_acc;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    if (tid == 0) {
      float tot = 0.f;
      for (int w = 0; w < kThreadsSimt / 32; ++w) tot += red[w];
      a.ds_partial[blockIdx.x] = tot;
    }
  }
}

}  // namespace

int clip_simt_ds_count(int64_t n_loc) { return (int)((n_loc + kRows - 1) / kRows); }

int clip_fwd_rows_simt(const ClipFwdArgs& a, cudaStream_t stream) {
  dim3 grid((unsigned)((a.n_loc + kRows - 1) / kRows));
  clip_fwd_simt_kernel<<<grid, kThreadsSimt, 0, stream>>>(a);
  LATTE_LAUNCH_OK();
  return LATTE_OK;
}

int clip_bwd_rows_simt(const ClipBwdArgs& a, cudaStream_t stream) {
  const size_t smem = sizeof(float) * (2 * kRows * (kKc + 1) + kRows * (kCols + 1) + kCols * kSlab);
  LATTE_CUDA_OK(cudaFuncSetAttribute(clip_bwd_simt_kernel,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((a.n_loc + kRows - 1) / kRows), (unsigned)((a.dim + kSlab - 1) / kSlab));
  clip_bwd_simt_kernel<<<grid, kThreadsSimt, smem, stream>>>(
      a, a.cross_terms ? 1.f : 0.f, a.cross_terms ? 2.f : 1.f);
  LATTE_LAUNCH_OK();
  return LATTE_OK;
}

}  // namespace latte

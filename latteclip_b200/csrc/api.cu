// C ABI of liblatte_b200: library info + the ClipLoss forward / backward orchestration
// (row kernels in clip_tc.cu / clip_simt.cu, small finalize kernels here).
#include "latte_common.cuh"
#include "tc_ptx.cuh"

#include <stdlib.h>

namespace latte {

int device_sm_count() {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 148;
  return sms > 0 ? sms : 148;
}

namespace {

// Optional per-stage CUDA-event timing (latte_clip_stage_times).  mark(id) closes the interval
// since the previous mark and charges it to stage `id`.
struct StageTimer {
  static constexpr int kMaxMarks = 64;
  cudaEvent_t ev[kMaxMarks];
  int id[kMaxMarks];
  int n = 0;
  bool failed = false;
  void mark(cudaStream_t st, int stage) {
    if (n >= kMaxMarks) { failed = true; return; }
    if (cudaEventCreate(&ev[n]) != cudaSuccess) { failed = true; return; }
    if (cudaEventRecord(ev[n], st) != cudaSuccess) failed = true;
    id[n++] = stage;
  }
  void collect(float* ms) {          // after a stream synchronise
    for (int k = 1; k < n; ++k) {
      float t = 0.f;
      if (cudaEventElapsedTime(&t, ev[k - 1], ev[k]) != cudaSuccess) failed = true;
      if (id[k] >= 0 && id[k] < LATTE_NUM_STAGES) ms[id[k]] += t;
    }
    for (int k = 0; k < n; ++k) cudaEventDestroy(ev[k]);
    n = 0;
  }
};
#define LATTE_MARK(stage) do { if (tm) tm->mark(st, (stage)); } while (0)

constexpr int kMaxParts = 16;

// Merge the per-(split, half) online-softmax partials of one row into a base-2 LSE.
// -> M (largest partial max) and log2 of the sum relative to it; LSE (base 2) = M + logL.
// Kept apart because the per-sample loss term (M - label logit) + logL is far more accurate
// than LSE - label logit when the label dominates (M == label logit, logL ~ 1e-4).
__device__ __forceinline__ void merge_parts2(const float* pmax, const float* psum, int nparts,
                                             int64_t n_loc, int64_t i, float& M, float& logL) {
  M = -INFINITY;
  for (int k = 0; k < nparts; ++k) M = fmaxf(M, pmax[(int64_t)k * n_loc + i]);
  float L = 0.f;
  for (int k = 0; k < nparts; ++k) {
    const float m = pmax[(int64_t)k * n_loc + i];
    if (m > -INFINITY) L += psum[(int64_t)k * n_loc + i] * exp2f(m - M);
  }
  logL = log2f(L);
}
__device__ __forceinline__ float merge_parts(const float* pmax, const float* psum, int nparts,
                                             int64_t n_loc, int64_t i) {
  float M, logL;
  merge_parts2(pmax, psum, nparts, n_loc, i, M, logL);
  return M + logL;
}

// row_lse / col_lse in natural-log units and per-block partial sums of the loss terms
// (loss.py:126-129); loss_reduce_kernel adds the partials.
constexpr int kFinalRows = 256;
__global__ void __launch_bounds__(kFinalRows)
clip_finalize_kernel(const float* pmax_r, const float* psum_r, const float* diag_r,
                     const float* pmax_c, const float* psum_c, const float* diag_c, int nparts,
                     int64_t n_loc, const float* logit_scale, float* row_lse, float* col_lse,
                     float* row_nll, float* col_nll, double* loss_partial) {
  __shared__ double red[kFinalRows / 32];
  const float s = __ldg(logit_scale);
  const float c2 = s * kLog2e;
  double acc = 0.0;
  const int64_t i = (int64_t)blockIdx.x * kFinalRows + threadIdx.x;
  if (i < n_loc) {
    float Mr, Lr, Mc, Lc;
    merge_parts2(pmax_r, psum_r, nparts, n_loc, i, Mr, Lr);
    merge_parts2(pmax_c, psum_c, nparts, n_loc, i, Mc, Lc);
    row_lse[i] = (Mr + Lr) * kLn2;
    col_lse[i] = (Mc + Lc) * kLn2;
    const float nr = ((Mr - c2 * diag_r[i]) + Lr) * kLn2;
    const float nc = ((Mc - c2 * diag_c[i]) + Lc) * kLn2;
    if (row_nll) row_nll[i] = nr;
    if (col_nll) col_nll[i] = nc;
    acc = (double)nr + (double)nc;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < kFinalRows / 32; ++w) tot += red[w];
    loss_partial[blockIdx.x] = tot;
  }
}

// Multi-rank forward: `gathered` is the all-gathered per-rank payload [world, stride] with
// col_ml [n_all, 2] | row_lse [n_loc] | row_nll [n_loc] | label_logit [n_loc] per rank.  Merges
// the column (max, sum) pairs into every column's LSE and per-sample loss term (same exactness
// check as above) and unpacks the row vectors into contiguous [n_all] arrays.
__global__ void __launch_bounds__(256)
col_merge_kernel(const float* gathered, int64_t stride, int world, int64_t n_loc, int64_t n_all,
                 int nblk_total, float* row_lse_all, float* row_nll_all, float* label_logit_all,
                 float* col_lse_all, float* col_nll_all, int* flag, const int* wait_flags,
                 int wait_gen) {
  if (wait_flags) {          // peer-memory exchange: every rank's payload row must have landed
    if ((int)threadIdx.x < world) ptx::flag_wait_ge(wait_flags + threadIdx.x, wait_gen);
    __syncthreads();
  }
  const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (j >= n_all) return;
  const float* own = gathered + (j / n_loc) * stride + 2 * n_all + (j % n_loc);
  const float label_logit = __ldcg(own + 2 * n_loc);
  row_lse_all[j] = __ldcg(own);
  row_nll_all[j] = __ldcg(own + n_loc);
  label_logit_all[j] = label_logit;
  float mx[LATTE_COMM_MAX_RANKS], lx[LATTE_COMM_MAX_RANKS];
#pragma unroll
  for (int w = 0; w < LATTE_COMM_MAX_RANKS; ++w) {
    // scalar loads: the rank stride 2 N + 3 n is odd for odd shard sizes
    mx[w] = w < world ? __ldcg(gathered + (int64_t)w * stride + 2 * j) : -INFINITY;
    lx[w] = w < world ? __ldcg(gathered + (int64_t)w * stride + 2 * j + 1) : 0.f;
  }
  float M = -INFINITY, L = 0.f;
#pragma unroll
  for (int w = 0; w < LATTE_COMM_MAX_RANKS; ++w) {
    if (!(mx[w] > -INFINITY)) continue;
    if (mx[w] > M) {
      L = L * exp2f(M - mx[w]) + lx[w];
      M = mx[w];
    } else {
      L = fmaf(lx[w], exp2f(mx[w] - M), L);
    }
  }
  const float logL = log2f(L);
  const float lse2 = M + logL;
  col_lse_all[j] = lse2 * kLn2;
  col_nll_all[j] = (fmaf(-kLog2e, label_logit, M) + logL) * kLn2;
  const bool ok = (M - lse2) + log2f((float)nblk_total) < 95.0f;
  if (!ok) atomicOr(flag, 1);
}

// Second (exact) round of the peer-memory forward, gated: merge the exact (max, sum) pairs of all
// ranks ([world][2 n_all]) into the column LSE and loss term.
__global__ void __launch_bounds__(256)
col_merge_exact_kernel(const int* gate, const float* exact, int world, int64_t n_all,
                       const float* label_logit_all, float* col_lse_all, float* col_nll_all,
                       const int* wait_flags, int wait_gen) {
  if (*gate == 0) return;
  if (wait_flags) {
    if ((int)threadIdx.x < world) ptx::flag_wait_ge(wait_flags + threadIdx.x, wait_gen);
    __syncthreads();
  }
  const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (j >= n_all) return;
  float M = -INFINITY, L = 0.f;
  for (int w = 0; w < world; ++w) {
    const float m = __ldcg(exact + (int64_t)w * 2 * n_all + 2 * j);
    const float l = __ldcg(exact + (int64_t)w * 2 * n_all + 2 * j + 1);
    if (!(m > -INFINITY)) continue;
    if (m > M) {
      L = L * exp2f(M - m) + l;
      M = m;
    } else {
      L = fmaf(l, exp2f(m - M), L);
    }
  }
  const float logL = log2f(L);
  col_lse_all[j] = (M + logL) * kLn2;
  col_nll_all[j] = (fmaf(-kLog2e, label_logit_all[j], M) + logL) * kLn2;
}

// ---- one launch for everything the pair forward sweep leaves to finish ------------------------
// Rows: merge the (slot, epilogue group, tile half) partials of every row.  Which slots hold data
// follows from the tile schedule alone (a group owns every second tile of its cluster's range), so
// the partial arrays need no "empty" fill beforehand.  Columns: per-128-row-block partial sums -> (max, sum) per column (see below).
struct PairFinalizeArgs {
  const float* pmax; const float* psum; int64_t n_loc; int col_tiles; int64_t total; int64_t ncl;
  const float* diag; const float* logit_scale;
  float* row_lse; float* row_nll; float* label_logit;         // label_logit nullable
  const float* col_part; const float* col_ref; int64_t ld; int nblk; int64_t n_all;   // col_part nullable
  float* col_lse; float* col_nll; float* col_ml; int* flag;
};

__device__ __forceinline__ bool pair_group_has_tile(int64_t cl, int64_t rb, int col_tiles, int64_t total,
                                                    int64_t ncl, int group) {
  const int64_t u0 = cl * total / ncl, u1 = (cl + 1) * total / ncl;
  const int64_t t0 = rb * col_tiles, t1 = t0 + col_tiles;
  const int64_t a = u0 > t0 ? u0 : t0, b = u1 < t1 ? u1 : t1;
  if (b - a >= 2) return true;
  if (b - a <= 0) return false;
  return ((a - u0) & 1) == group;
}

__device__ __forceinline__ void pair_finalize_row(const PairFinalizeArgs& a, int64_t i) {
  const int64_t rb = i / 256;
  const int64_t c0 = cluster_of_tile(rb * a.col_tiles, a.total, a.ncl);
  const int64_t c1 = cluster_of_tile(rb * a.col_tiles + a.col_tiles - 1, a.total, a.ncl);
  float M = -INFINITY;
  for (int64_t c = c0; c <= c1; ++c)
#pragma unroll
    for (int g = 0; g < 2; ++g)
      if (pair_group_has_tile(c, rb, a.col_tiles, a.total, a.ncl, g)) {
        const int64_t base = ((c - c0) * 2 + g) * 2 * a.n_loc + i;
        M = fmaxf(M, fmaxf(a.pmax[base], a.pmax[base + a.n_loc]));
      }
  float L = 0.f;
  for (int64_t c = c0; c <= c1; ++c)
#pragma unroll
    for (int g = 0; g < 2; ++g)
      if (pair_group_has_tile(c, rb, a.col_tiles, a.total, a.ncl, g)) {
        const int64_t base = ((c - c0) * 2 + g) * 2 * a.n_loc + i;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float m = a.pmax[base + h * a.n_loc];
          if (m > -INFINITY) L += a.psum[base + h * a.n_loc] * exp2f(m - M);
        }
      }
  const float logL = log2f(L);
  a.row_lse[i] = (M + logL) * kLn2;
  const float s = __ldg(a.logit_scale);
  // fma(-c2, dot, M): the exact residual the sweep's own fma(dot, c2, -M) saw for the label
  if (a.row_nll) a.row_nll[i] = (fmaf(-(s * kLog2e), a.diag[i], M) + logL) * kLn2;
  if (a.label_logit) a.label_logit[i] = s * a.diag[i];
}

__global__ void __launch_bounds__(256) pair_fwd_finalize_kernel(const PairFinalizeArgs a) {
  if (!a.col_part) {                       // rows only: 256 rows per CTA
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i < a.n_loc) pair_finalize_row(a, i);
    return;
  }
  // 64 columns x 4 interleaved groups of row blocks per CTA; loads are batched 8 deep so the
  // merge is not a chain of dependent L2 round trips.
  __shared__ float sm_m[4][64], sm_l[4][64];
  const int cx = threadIdx.x & 63, grp = threadIdx.x >> 6;
  const int64_t j = (int64_t)blockIdx.x * 64 + cx;
  const int64_t jc = j < a.n_all ? j : a.n_all - 1;
  const int64_t ht = jc / 32;
  float M = -INFINITY, L = 0.f;
  for (int b0 = grp; b0 < a.nblk; b0 += 32) {
    float mb[8], c[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int b = b0 + 4 * u;
      mb[u] = b < a.nblk ? a.col_ref[(int64_t)b * 4 * a.col_tiles + ht] : -INFINITY;
      c[u] = b < a.nblk ? a.col_part[(int64_t)b * a.ld + jc] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (!(mb[u] > -INFINITY)) continue;
      if (mb[u] > M) {
        L = L * exp2f(M - mb[u]) + c[u];
        M = mb[u];
      } else {
        L = fmaf(c[u], exp2f(mb[u] - M), L);
      }
    }
  }
  sm_m[grp][cx] = M;
  sm_l[grp][cx] = L;
  __syncthreads();
  if (grp == 1 && j < a.n_loc) pair_finalize_row(a, j);      // this CTA's 64 rows
  if (grp == 0 && j < a.n_all) {
    float Mt = fmaxf(fmaxf(sm_m[0][cx], sm_m[1][cx]), fmaxf(sm_m[2][cx], sm_m[3][cx]));
    float Lt = 0.f;
#pragma unroll
    for (int g = 0; g < 4; ++g)
      if (sm_m[g][cx] > -INFINITY) Lt = fmaf(sm_l[g][cx], exp2f(sm_m[g][cx] - Mt), Lt);
    if (a.col_ml) {
      a.col_ml[2 * j] = Mt;
      a.col_ml[2 * j + 1] = Lt;
      return;
    }
    const float logL = log2f(Lt);
    const float lse2 = Mt + logL;
    a.col_lse[j] = lse2 * kLn2;
    if (a.col_nll) a.col_nll[j] = (fmaf(-(__ldg(a.logit_scale) * kLog2e), a.diag[j], Mt) + logL) * kLn2;
    const bool ok = (Mt - lse2) + log2f((float)a.nblk) < 95.0f;    // false for NaN / -inf too
    if (!ok) atomicOr(a.flag, 1);
  }
}

// Last kernel of a forward: (gated) overwrite of the column LSE / loss term with the exact
// row-kernel result, then the loss = sum over rows [loss_off, loss_off + n_loss) of
// row_nll + col_nll, / (2 n_loss)  (loss.py:126-129).  Per-CTA partials, added in order by the
// last CTA to arrive (counter must be 0 on entry and is 0 again on exit).
__global__ void __launch_bounds__(256)
pair_fwd_finish_kernel(const int* gate, const float* pmax, const float* psum, int nparts,
                       int64_t n_merge, const float* label_dot, const float* logit_scale,
                       const float* row_lse, float* lse_out, float* nll_out, const float* row_nll,
                       const float* col_nll, int64_t loss_off, int64_t n_loss, double* loss_partial,
                       float* stat_partial, unsigned int* counter, float* loss, float* stats) {
  __shared__ double red[8];
  __shared__ float rlo[8], rhi[8], rnl[8];
  __shared__ bool last;
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int nb = (int)gridDim.x;
  double acc = 0.0;
  float lo = INFINITY, hi = -INFINITY, nmax = -INFINITY;
  if (i < n_merge) {
    float cn = col_nll[i];
    float cl = lse_out[i];
    if (gate && *gate != 0) {
      float M, logL;
      merge_parts2(pmax, psum, nparts, n_merge, i, M, logL);
      cl = (M + logL) * kLn2;
      lse_out[i] = cl;
      const float s = logit_scale ? __ldg(logit_scale) : 1.f;
      cn = ((M - s * kLog2e * label_dot[i]) + logL) * kLn2;
      nll_out[i] = cn;
    }
    const float rl = row_lse[i], rn = row_nll[i];
    lo = fminf(rl, cl); hi = fmaxf(rl, cl); nmax = fmaxf(rn, cn);
    if (i >= loss_off && i < loss_off + n_loss) acc = (double)rn + (double)cn;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    acc += __shfl_xor_sync(0xffffffffu, acc, o);
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    nmax = fmaxf(nmax, __shfl_xor_sync(0xffffffffu, nmax, o));
  }
  if ((threadIdx.x & 31) == 0) {
    const int w = threadIdx.x >> 5;
    red[w] = acc; rlo[w] = lo; rhi[w] = hi; rnl[w] = nmax;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < 8; ++w) {
      tot += red[w];
      lo = fminf(lo, rlo[w]); hi = fmaxf(hi, rhi[w]); nmax = fmaxf(nmax, rnl[w]);
    }
    loss_partial[blockIdx.x] = tot;
    stat_partial[blockIdx.x] = lo;
    stat_partial[nb + blockIdx.x] = hi;
    stat_partial[2 * nb + blockIdx.x] = nmax;
    __threadfence();
    last = atomicInc(counter, gridDim.x - 1) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  double v = 0.0;
  lo = INFINITY; hi = -INFINITY; nmax = -INFINITY;
  for (int b = threadIdx.x; b < nb; b += 256) {
    v += __ldcg(loss_partial + b);
    lo = fminf(lo, __ldcg(stat_partial + b));
    hi = fmaxf(hi, __ldcg(stat_partial + nb + b));
    nmax = fmaxf(nmax, __ldcg(stat_partial + 2 * nb + b));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v += __shfl_xor_sync(0xffffffffu, v, o);
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    nmax = fmaxf(nmax, __shfl_xor_sync(0xffffffffu, nmax, o));
  }
  if ((threadIdx.x & 31) == 0) {
    const int w = threadIdx.x >> 5;
    red[w] = v; rlo[w] = lo; rhi[w] = hi; rnl[w] = nmax;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < 8; ++w) {
      tot += red[w];
      lo = fminf(lo, rlo[w]); hi = fmaxf(hi, rhi[w]); nmax = fmaxf(nmax, rnl[w]);
    }
    *loss = (float)(tot / (2.0 * (double)n_loss));
    if (stats) { stats[0] = lo; stats[1] = hi; stats[2] = nmax; stats[3] = 0.f; }
  }
}

__global__ void __launch_bounds__(256)
loss_reduce_kernel(const double* partial, int count, int64_t n_loc, float* loss) {
  __shared__ double red[8];
  double acc = 0.0;
  for (int i = threadIdx.x; i < count; i += 256) acc += partial[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < 8; ++w) tot += red[w];
    *loss = (float)(tot / (2.0 * (double)n_loc));
  }
}

// One warp per gathered row k: fp16 copies of both feature rows (bf16 input only) and the label
// logit d_k = s * <img_k, txt_k>, from which u = max_k max(1 - P^row_kk, 1 - P^col_kk) bounds
// every |G_ij| by 2u (off-diagonal softmax weights of a row / column sum to 1 - P_kk).
__global__ void __launch_bounds__(256)
pair_prep_features_kernel(const void* img, int64_t ld_img, const void* txt, int64_t ld_txt,
                          int is_bf16, __half* img16, __half* txt16, int64_t ld16, int64_t rows,
                          int64_t dim, const float* logit_scale, const float* row_lse,
                          const float* col_lse, unsigned int* u_bits) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float s = __ldg(logit_scale);
  float umax = 0.f;
  for (int64_t k = warp0; k < rows; k += nwarps) {
    float dot = 0.f;
    for (int64_t c = lane * 8; c < dim; c += 256) {
      float fi[8], ft[8];
      if (is_bf16) {
        const uint4 ri = *reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(img) + k * ld_img + c);
        const uint4 rt = *reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(txt) + k * ld_txt + c);
        const __nv_bfloat162* vi = reinterpret_cast<const __nv_bfloat162*>(&ri);
        const __nv_bfloat162* vt = reinterpret_cast<const __nv_bfloat162*>(&rt);
        uint4 oi, ot;
        __half2* hi = reinterpret_cast<__half2*>(&oi);
        __half2* ht = reinterpret_cast<__half2*>(&ot);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 a = __bfloat1622float2(vi[e]), b = __bfloat1622float2(vt[e]);
          fi[2 * e] = a.x; fi[2 * e + 1] = a.y; ft[2 * e] = b.x; ft[2 * e + 1] = b.y;
          hi[e] = __float22half2_rn(a);
          ht[e] = __float22half2_rn(b);
        }
        if (img16) {
          *reinterpret_cast<uint4*>(img16 + k * ld16 + c) = oi;
          if (txt16 != img16) *reinterpret_cast<uint4*>(txt16 + k * ld16 + c) = ot;
        }
      } else {
        const uint4 ri = *reinterpret_cast<const uint4*>(static_cast<const __half*>(img) + k * ld_img + c);
        const uint4 rt = *reinterpret_cast<const uint4*>(static_cast<const __half*>(txt) + k * ld_txt + c);
        const __half2* vi = reinterpret_cast<const __half2*>(&ri);
        const __half2* vt = reinterpret_cast<const __half2*>(&rt);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 a = __half22float2(vi[e]), b = __half22float2(vt[e]);
          fi[2 * e] = a.x; fi[2 * e + 1] = a.y; ft[2 * e] = b.x; ft[2 * e + 1] = b.y;
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) dot = fmaf(fi[e], ft[e], dot);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    const float d = s * dot;
    const float u = fmaxf(-expm1f(d - row_lse[k]), -expm1f(d - col_lse[k]));
    umax = fmaxf(umax, u);            // NaN-free inputs assumed; fmaxf drops a NaN operand
  }
  if (lane == 0 && umax > 0.f) atomicMax(u_bits, __float_as_uint(umax));
}

// Statistics the backward's scaling needs: stats[0] = min, stats[1] = max of all LSE values
// (natural log), stats[2] = the largest per-sample loss term nll (bounds |G|: 1 - P_kk =
// -expm1(-nll_k)); without nll vectors stats[2] = -log1p(-u) from the label-logit bound u.
// The forward produces the same three numbers for free (pair_fwd_finish_kernel); this kernel
// serves callers that only hand over the LSE vectors.
__global__ void __launch_bounds__(1024)
lse_stats_kernel(const float* row_lse, const float* col_lse, int64_t n_all,
                 const unsigned int* u_bits, const float* row_nll, const float* col_nll,
                 float* stats) {
  __shared__ float smin[32], smax[32], snll[32];
  float lo = INFINITY, hi = -INFINITY, nmax = -INFINITY;
  for (int64_t i = threadIdx.x; i < n_all; i += 1024) {
    const float a = row_lse[i], b = col_lse[i];
    lo = fminf(lo, fminf(a, b));
    hi = fmaxf(hi, fmaxf(a, b));
    if (row_nll) nmax = fmaxf(nmax, fmaxf(row_nll[i], col_nll[i]));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    nmax = fmaxf(nmax, __shfl_xor_sync(0xffffffffu, nmax, o));
  }
  if ((threadIdx.x & 31) == 0) {
    smin[threadIdx.x >> 5] = lo; smax[threadIdx.x >> 5] = hi; snll[threadIdx.x >> 5] = nmax;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 0; w < 32; ++w) {
      lo = fminf(lo, smin[w]); hi = fmaxf(hi, smax[w]); nmax = fmaxf(nmax, snll[w]);
    }
    if (!row_nll) {
      // no loss terms: the label-logit bound u of the prep kernel, or no bound at all (u = 1)
      const float u = fminf(u_bits ? __uint_as_float(*u_bits) : 1.0f, 0.9999999f);
      nmax = -log1pf(-u);
    }
    stats[0] = lo; stats[1] = hi; stats[2] = nmax;
  }
}

// Scalars of the backward from the LSE statistics (every thread derives the same values):
//   rho = mid-point of the LSE range (base 2); fast = 1 when max - min <= 64 binary orders, so that
//   2^(lse - rho) and products of two such factors stay well inside fp32 range (clip_pair.cu's
//   one-ex2 epilogue), otherwise the sweep uses its two-ex2 epilogue;
//   gs = log2 of the fp16 scale of G: |G| <= 2u with u = max_k (1 - P_kk); the inputs carry ~5e-5 of
//   fp32 error, so u is padded by 1e-3: 2 (u + 1e-3) 2^gs <= 2^14 keeps fp16 G finite (gs = 13 for
//   u ~ 1, up to 22 when converged);  out_scale: gradient = out_scale * (accumulated G . features).
struct BwdScalars { float rho; int fast; float gs; float out_scale; };
__device__ __forceinline__ BwdScalars bwd_scalars(const float* stats, const float* grad_loss,
                                                  float grad_mult, const float* logit_scale,
                                                  int64_t n_loc) {
  // grad_loss == NULL: plain products (latte_distill_products), out_scale only undoes the fp16 scale
  BwdScalars o;
  const float lo = __ldg(stats) * kLog2e, hi = __ldg(stats + 1) * kLog2e;
  const bool ok = (hi - lo) <= 64.0f && hi < 3.0e38f && lo > -3.0e38f;   // also rejects NaN / inf
  o.rho = ok ? 0.5f * (lo + hi) : 0.f;
  o.fast = ok ? 1 : 0;
  // an nll can come out slightly negative; NaN -> u = 1
  float u_raw = -expm1f(-fmaxf(__ldg(stats + 2), 0.f));
  if (!(u_raw >= 0.f)) u_raw = 1.0f;
  const float u = fminf(u_raw, 1.0f) + 1.0e-3f;
  o.gs = fminf(fmaxf(13.0f - ceilf(log2f(u)), 13.0f), 22.0f);
  o.out_scale = grad_loss ? __ldg(grad_loss) * grad_mult / (2.0f * (float)n_loc) * __ldg(logit_scale) * exp2f(-o.gs)
                          : exp2f(-o.gs);
  return o;
}

// natural-log LSE -> base-2 units (zero padded to the tile grid) and the rank-one factors; the
// first thread also publishes the scalars (scal = {rho, fast flag, u bits (unused), gs, out_scale})
__global__ void lse_vectors_kernel(const float* row_lse, const float* col_lse, int64_t n_all,
                                   int64_t n_pad, const float* stats, const float* grad_loss,
                                   float grad_mult, const float* logit_scale, int64_t n_loc,
                                   float* scal, float* row2, float* col2,
                                   float* e_row, float* einv_row, float* e_col, float* einv_col) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  BwdScalars sc = {0.f, 0, 13.f, 0.f};
  if (stats) {
    sc = bwd_scalars(stats, grad_loss, grad_mult, logit_scale, n_loc);
    if (j == 0) {
      scal[0] = sc.rho;
      reinterpret_cast<int*>(scal)[1] = sc.fast;
      scal[3] = sc.gs;
      scal[4] = sc.out_scale;
    }
  }
  if (j >= n_pad) return;
  const bool in = j < n_all;
  const float r2 = in ? row_lse[j] * kLog2e : 0.f;
  const float c2 = in ? col_lse[j] * kLog2e : 0.f;
  row2[j] = r2;
  col2[j] = c2;
  if (e_row) {
    const float rho = sc.rho;
    e_row[j] = in ? exp2f(r2 - rho) : 0.f;
    einv_row[j] = in ? exp2f(rho - r2) : 0.f;
    e_col[j] = in ? exp2f(c2 - rho) : 0.f;
    einv_col[j] = in ? exp2f(rho - c2) : 0.f;
  }
}

__global__ void __launch_bounds__(256)
ds_reduce_kernel(const float* partial, int count, const float* grad_loss, float grad_mult,
                 int64_t n_loc, float* d_scale) {
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
    *d_scale = (float)(tot * (double)(__ldg(grad_loss) * grad_mult) / (2.0 * (double)n_loc));
  }
}

// Fused operand preparation (the reference normalises in the towers, model.py:415-418 / :420-437,
// and casts inside its autocast matmul, loss.py:109-116): one warp per feature row -- optional L2
// normalisation in fp32 (F.normalize: x / max(||x||, 1e-12)), rounding to the compute dtype
// (bf16 or fp16, what the reference's matmul would see) and storage as fp16, the one 16-bit format
// every tensor-core kernel of the path takes (a bf16 value is exact in fp16 down to 2^-17; below
// that the absolute error is < 2^-25).  Replaces the ATen cast, the backward's prep pass and its
// bf16 -> fp16 copies.  dim <= 768, dim % 8 == 0.
__global__ void __launch_bounds__(256)
prep_features_kernel(const void* x, int64_t ld, int in_dtype, int64_t rows, int64_t dim, int normalize,
                     int round_dtype, __half* out, int64_t ld_out, float* inv_norm) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  float v[24];
  float ss = 0.f;
#pragma unroll
  for (int it = 0; it < 3; ++it) {
    const int64_t c = (int64_t)it * 256 + lane * 8;
    if (c < dim) {
      if (in_dtype == LATTE_F32) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(x) + row * ld + c));
        const float4 b = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(x) + row * ld + c) + 1);
        v[it * 8 + 0] = a.x; v[it * 8 + 1] = a.y; v[it * 8 + 2] = a.z; v[it * 8 + 3] = a.w;
        v[it * 8 + 4] = b.x; v[it * 8 + 5] = b.y; v[it * 8 + 6] = b.z; v[it * 8 + 7] = b.w;
      } else {
        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(x) + row * ld + c));
        if (in_dtype == LATTE_BF16) {
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __bfloat1622float2(h[e]);
            v[it * 8 + 2 * e] = f.x; v[it * 8 + 2 * e + 1] = f.y;
          }
        } else {
          const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __half22float2(h[e]);
            v[it * 8 + 2 * e] = f.x; v[it * 8 + 2 * e + 1] = f.y;
          }
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) ss = fmaf(v[it * 8 + e], v[it * 8 + e], ss);
    }
  }
  float inv = 1.0f;
  if (normalize) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
    if (inv_norm && lane == 0) inv_norm[row] = inv;
  }
#pragma unroll
  for (int it = 0; it < 3; ++it) {
    const int64_t c = (int64_t)it * 256 + lane * 8;
    if (c < dim) {
      uint4 o;
      __half2* h = reinterpret_cast<__half2*>(&o);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float a = v[it * 8 + 2 * e], b = v[it * 8 + 2 * e + 1];
        if (normalize) { a *= inv; b *= inv; }
        if (round_dtype == LATTE_BF16) {
          a = __bfloat162float(__float2bfloat16_rn(a));
          b = __bfloat162float(__float2bfloat16_rn(b));
        }
        h[e] = __floats2half2_rn(a, b);
      }
      *reinterpret_cast<uint4*>(out + row * ld_out + c) = o;
    }
  }
}

// Backward of the fused normalisation: d_x = inv * (g - xh <xh, g>), xh = x * inv (F.normalize).
// One warp per row; x is the original input and d_x its gradient, both in `x_dtype`.
__global__ void __launch_bounds__(256)
normalize_bwd_kernel(const void* g, int64_t ld_g, int g_dtype, const void* x, int64_t ld_x, int x_dtype,
                     const float* inv_norm, int64_t rows, int64_t dim, void* d_x, int64_t ld_dx) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  auto ld1 = [](const void* base, int64_t idx, int dt) -> float {
    if (dt == LATTE_F32) return static_cast<const float*>(base)[idx];
    if (dt == LATTE_BF16) return __bfloat162float(static_cast<const __nv_bfloat16*>(base)[idx]);
    return __half2float(static_cast<const __half*>(base)[idx]);
  };
  const float inv = inv_norm[row];
  float dot = 0.f;
  for (int64_t c = lane; c < dim; c += 32)
    dot = fmaf(ld1(x, row * ld_x + c, x_dtype) * inv, ld1(g, row * ld_g + c, g_dtype), dot);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  for (int64_t c = lane; c < dim; c += 32) {
    const float xh = ld1(x, row * ld_x + c, x_dtype) * inv;
    const float r = inv * (ld1(g, row * ld_g + c, g_dtype) - xh * dot);
    if (x_dtype == LATTE_F32) static_cast<float*>(d_x)[row * ld_dx + c] = r;
    else if (x_dtype == LATTE_BF16) static_cast<__nv_bfloat16*>(d_x)[row * ld_dx + c] = __float2bfloat16_rn(r);
    else static_cast<__half*>(d_x)[row * ld_dx + c] = __float2half_rn(r);
  }
}

// bf16 -> fp16 copy of a feature matrix (second GEMM of the tc backward, see clip_tc.cu)
__global__ void bf16_to_fp16_kernel(const __nv_bfloat16* in0, __half* out0,
                                    const __nv_bfloat16* in1, __half* out1, int64_t ld_in,
                                    int64_t ld_out, int64_t rows, int64_t dim) {
  const __nv_bfloat16* in = blockIdx.y ? in1 : in0;
  __half* out = blockIdx.y ? out1 : out0;
  const int64_t per_row = (dim + 7) / 8;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * per_row) return;
  const int64_t r = idx / per_row, c = (idx % per_row) * 8;
  if (c + 8 <= dim && (ld_in % 8) == 0) {
    const uint4 raw = *reinterpret_cast<const uint4*>(in + r * ld_in + c);
    const __nv_bfloat162* v = reinterpret_cast<const __nv_bfloat162*>(&raw);
    uint4 o;
    __half2* h = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int k = 0; k < 4; ++k) h[k] = __float22half2_rn(__bfloat1622float2(v[k]));
    *reinterpret_cast<uint4*>(out + r * ld_out + c) = o;
  } else {
    for (int64_t k = c; k < dim && k < c + 8; ++k)
      out[r * ld_out + k] = __float2half_rn(__bfloat162float(in[r * ld_in + k]));
  }
}

// =========================================================================== peer-memory exchange
// Flag block of one slot on one rank (int32, LATTE_COMM_FLAG_INTS): entry [kind * 8 + src].
constexpr int kFlagLandedTxt = 0, kFlagLandedImg = 8, kFlagPayload = 16, kFlagPayload2 = 24,
              kFlagDone = 32, kFlagFree = 40, kFlagCounter = 48;
constexpr int kCntPush = 0, kCntPayload = 1, kCntPayload2 = 2, kCntGemm = 3, kCntFinish = 4, kCntPushImg = 5;

struct PeerPtrs {
  void* p[LATTE_COMM_MAX_RANKS];
};
struct PeerFlags {
  int* p[LATTE_COMM_MAX_RANKS];      // NULL entries: no signalling (single-process tests)
};

// All threads call this after their (peer) stores; returns true in every thread of the last CTA
// of the grid to get here.  `counter` wraps back to 0.
__device__ __forceinline__ bool grid_last_cta(unsigned int* counter) {
  __shared__ bool s_last;
  // the CTA barrier orders every thread's stores before thread 0's system-scope fence, which is
  // cumulative: one fence per CTA instead of one per thread (a MEMBAR.SYS per thread made these
  // kernels several times slower)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    s_last = atomicInc(counter, gridDim.x - 1) == gridDim.x - 1;
  }
  __syncthreads();
  return s_last;
}
// flag[kind + rank] = gen on every rank (called by the last CTA)
__device__ __forceinline__ void signal_all(const PeerFlags& f, int world, int kind, int rank, int gen) {
  if ((int)threadIdx.x < world && f.p[threadIdx.x]) {
    __threadfence_system();
    ptx::st_release_sys(f.p[threadIdx.x] + kind + rank, gen);
  }
}

// Push this rank's text (and image) shard into every rank's gathered buffer: the all-gather of
// loss.py:49-50 / :54-55 as one NVLink store kernel, 16-byte vectors.  Before storing, wait until
// every peer released the previous generation of this slot (its readers are done).
__global__ void __launch_bounds__(256)
comm_push_kernel(const uint4* shard, int64_t shard_vecs, int64_t dst_off_vecs, PeerPtrs dst,
                 PeerFlags flags, int* my_flags, int world, int rank, int gen, int kind, int cnt) {
  if (my_flags) {
    if ((int)threadIdx.x < world && kind == kFlagLandedTxt)
      ptx::flag_wait_ge(my_flags + kFlagFree + threadIdx.x, gen - 1);
    __syncthreads();
  }
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < shard_vecs; v += stride) {
    const uint4 val = __ldg(shard + v);
#pragma unroll
    for (int w = 0; w < LATTE_COMM_MAX_RANKS; ++w)
      if (w < world) static_cast<uint4*>(dst.p[w])[dst_off_vecs + v] = val;
  }
  if (!my_flags) return;
  if (grid_last_cta(reinterpret_cast<unsigned int*>(my_flags + kFlagCounter + cnt)))
    signal_all(flags, world, kind, rank, gen);
}

__global__ void comm_release_kernel(PeerFlags flags, int world, int rank, int gen) {
  signal_all(flags, world, kFlagFree, rank, gen);
}

// Store this rank's forward payload (already in its own block row `rank`) into row `rank` of every
// peer's payload block, then publish `kind` = gen.  With `gate` the kernel only acts when *gate != 0
// (second, exact round: every rank takes the same decision because the merged statistics are
// bit-identical on all ranks), after rebuilding the (max, sum) pairs from the exact row kernel.
__global__ void __launch_bounds__(256)
comm_payload_kernel(const int* gate, const float* pmax, const float* psum, int nparts, int64_t n_cols,
                    float* row_local, int64_t len, int64_t row_off, PeerPtrs blocks, PeerFlags flags,
                    int* my_flags, int world, int rank, int gen, int kind, int cnt) {
  if (gate && *gate == 0) return;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < len; v += stride) {
    float val;
    if (gate) {
      // exact column partial of this rank: merge the row kernel's parts of text row v / 2
      const int64_t j = v >> 1;
      float M = -INFINITY;
      for (int k = 0; k < nparts; ++k) M = fmaxf(M, pmax[(int64_t)k * n_cols + j]);
      float L = 0.f;
      for (int k = 0; k < nparts; ++k) {
        const float m = pmax[(int64_t)k * n_cols + j];
        if (m > -INFINITY) L += psum[(int64_t)k * n_cols + j] * exp2f(m - M);
      }
      val = (v & 1) ? L : M;
      row_local[v] = val;
    } else {
      val = row_local[v];
    }
#pragma unroll
    for (int w = 0; w < LATTE_COMM_MAX_RANKS; ++w)
      if (w < world && w != rank) static_cast<float*>(blocks.p[w])[row_off + v] = val;
  }
  if (!my_flags) return;
  if (grid_last_cta(reinterpret_cast<unsigned int*>(my_flags + kFlagCounter + cnt)))
    signal_all(flags, world, kind, rank, gen);
}

// Last step of the fused reduce-scatter: wait until every rank's GEMM has published its adds and
// turn this rank's accumulator into d_txt (the adders applied the scale).  The accumulator is then
// cleared for the slot's next generation by a memset node and the slot released by
// comm_release_kernel (a zero store into the lines just read, fused in here, made this kernel
// 3-5x slower than the three operations together).
__global__ void __launch_bounds__(256)
comm_acc_finish_kernel(const float* acc, int64_t n_loc, int64_t dim, void* d_txt, int out_dtype,
                       int64_t ld_out, const int* my_flags, int world, int gen) {
  if (my_flags) {
    if ((int)threadIdx.x < world) ptx::flag_wait_ge(my_flags + kFlagDone + threadIdx.x, gen);
    __syncthreads();
  }
  // the acquire above orders these reads after the peers' adds (which happen in this GPU's L2, the
  // point of coherence of its memory); .cg keeps them out of L1
  const int64_t per_row = dim / 4;
  const int64_t total = n_loc * per_row;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int64_t r = idx / per_row, c = (idx % per_row) * 4;
    const float4 v = __ldcg(reinterpret_cast<const float4*>(acc + r * dim + c));
    if (out_dtype == LATTE_F32) {
      *reinterpret_cast<float4*>(static_cast<float*>(d_txt) + r * ld_out + c) = v;
    } else if (out_dtype == LATTE_BF16) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
      *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(d_txt) + r * ld_out + c) =
          make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
    } else {
      __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
      *reinterpret_cast<uint2*>(static_cast<__half*>(d_txt) + r * ld_out + c) =
          make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
    }
  }
}

struct WsLayout {
  size_t part;      // floats per partial array (kMaxParts * n_loc)
  size_t off_pmax_r, off_psum_r, off_diag_r, off_pmax_c, off_psum_c, off_diag_c;
  size_t off_row2, off_col2, off_ds, off_lossp, off_nll_r, off_nll_c;
  size_t off_y16a, off_y16b;    // fp16 copies of txt_all / img_all (bf16 features only)
  size_t ld16;
  size_t n_pad, ds_cap;
  // CTA-pair backward scratch (bwd layout only): blocked fp16 G + two fp32 accumulators
  bool pair;
  size_t off_g, off_acc0, off_acc1, ld32;
  size_t off_rho, off_erow, off_einvrow, off_ecol, off_einvcol;
  // CTA-pair forward scratch (fwd layout only)
  bool pair_fwd;
  size_t off_pp_max_r, off_pp_sum_r, off_pp_max_c, off_pp_sum_c, off_colpart, off_colref, off_flag;
  size_t total;
};

bool pair_shape_ok(int dtype, int64_t dim) {
  return (dtype == LATTE_BF16 || dtype == LATTE_F16) && dim >= 8 && dim <= 768 && (dim % 8) == 0;
}

WsLayout ws_layout(int64_t n_loc, int64_t n_all, int64_t dim, int dtype, bool bwd = false) {
  WsLayout w;
  auto up = [](size_t x) { return (x + 63) / 64 * 64; };   // keep every array 256-byte aligned
  w.part = up((size_t)kMaxParts * (size_t)n_loc);
  const size_t nl = up((size_t)n_loc);
  size_t o = 0;
  w.off_pmax_r = o; o += w.part;
  w.off_psum_r = o; o += w.part;
  w.off_diag_r = o; o += nl;
  w.off_pmax_c = o; o += w.part;
  w.off_psum_c = o; o += w.part;
  w.off_diag_c = o; o += nl;
  w.n_pad = ((size_t)n_all + 255) / 256 * 256 + 128;
  w.off_row2 = o; o += up(w.n_pad);
  w.off_col2 = o; o += up(w.n_pad);
  w.ds_cap = up(2 * (((size_t)n_loc + 63) / 64) + 2 * 160);
  w.off_ds = o; o += w.ds_cap;
  // per finish-CTA partials: one double (loss) + three floats (LSE min / max, nll max)
  w.off_lossp = o; o += up(5 * (((size_t)n_loc + 255) / 256) + 8);
  w.off_nll_r = o; o += nl;
  w.off_nll_c = o; o += nl;
  w.ld16 = ((size_t)dim + 7) / 8 * 8;
  w.off_y16a = w.off_y16b = o;
  if (dtype == LATTE_BF16) {
    const size_t f = up(((size_t)n_all * w.ld16 + 1) / 2);   // fp16 elements counted in floats
    w.off_y16a = o; o += f;
    w.off_y16b = o; o += f;
  }
  w.pair = bwd && pair_shape_ok(dtype, dim);
  w.off_g = w.off_acc0 = w.off_acc1 = o;
  w.ld32 = ((size_t)dim + 3) / 4 * 4;
  w.off_rho = w.off_erow = w.off_einvrow = w.off_ecol = w.off_einvcol = o;
  if (w.pair) {
    w.off_rho = o; o += 64;
    w.off_erow = o; o += up(w.n_pad);
    w.off_einvrow = o; o += up(w.n_pad);
    w.off_ecol = o; o += up(w.n_pad);
    w.off_einvcol = o; o += up(w.n_pad);
    const PairGeom geo = clip_pair_geom(n_loc, n_all);
    w.off_g = o; o += up((geo.g_elems + 1) / 2);
    const size_t acc = up((size_t)n_loc * w.ld32);
    w.off_acc0 = o; o += acc;
    w.off_acc1 = o; o += acc;
  }
  w.pair_fwd = !bwd && pair_shape_ok(dtype, dim);
  w.off_pp_max_r = w.off_pp_sum_r = w.off_pp_max_c = w.off_pp_sum_c = o;
  w.off_colpart = w.off_colref = w.off_flag = o;
  if (w.pair_fwd) {
    const PairFwdGeom f = clip_pair_fwd_geom(n_loc, n_all);
    const size_t pp = up((size_t)4 * f.slots * (size_t)n_loc);
    w.off_pp_max_r = o; o += pp;
    w.off_pp_sum_r = o; o += pp;
    w.off_pp_max_c = o; o += pp;
    w.off_pp_sum_c = o; o += pp;
    w.off_colpart = o; o += up((size_t)2 * f.row_blocks * (size_t)f.ld_colpart);
    w.off_colref = o; o += up((size_t)2 * f.row_blocks * 4 * (size_t)f.col_tiles);
    w.off_flag = o; o += 64;
  }
  w.total = o * sizeof(float);
  return w;
}

bool use_tc(int dtype, int64_t dim, const void* xa, int64_t lda, const void* xb, int64_t ldb,
            const void* xc, int64_t ldc, const void* xd, int64_t ldd) {
  return clip_tc_supported(dtype, dim, lda, ldb, xa, xb) && clip_tc_supported(dtype, dim, ldc, ldd, xc, xd);
}

}  // namespace
}  // namespace latte

using namespace latte;

extern "C" const char* latte_version(void) { return "latte_b200 0.1.0 (sm_100a)"; }

extern "C" const char* latte_status_string(int status) {
  switch (status) {
    case LATTE_OK: return "ok";
    case LATTE_ERR_BAD_ARG: return "bad argument (null pointer, negative size or unknown enum)";
    case LATTE_ERR_UNSUPPORTED: return "unsupported shape, dtype or alignment";
    case LATTE_ERR_WORKSPACE: return "workspace too small";
    case LATTE_ERR_CUDA: return "CUDA runtime/driver call or kernel launch failed";
    case LATTE_ERR_NO_DEVICE: return "no sm_100 CUDA device";
    default: return "unknown status";
  }
}

extern "C" int latte_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return LATTE_ERR_NO_DEVICE;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return LATTE_ERR_NO_DEVICE;
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return LATTE_OK;
}

extern "C" int latte_clip_workspace_bytes(int64_t n_loc, int64_t n_all, int64_t dim, int dtype,
                                          size_t* bytes) {
  LATTE_CHECK_ARG(bytes && n_loc > 0 && n_all >= n_loc && dim > 0);
  LATTE_CHECK_ARG(dtype >= LATTE_F32 && dtype <= LATTE_F16);
  *bytes = ws_layout(n_loc, n_all, dim, dtype).total;
  return LATTE_OK;
}

extern "C" int latte_clip_bwd_workspace_bytes(int64_t n_loc, int64_t n_all, int64_t dim, int dtype,
                                              size_t* bytes) {
  LATTE_CHECK_ARG(bytes && n_loc > 0 && n_all >= n_loc && dim > 0);
  LATTE_CHECK_ARG(dtype >= LATTE_F32 && dtype <= LATTE_F16);
  *bytes = ws_layout(n_loc, n_all, dim, dtype, true).total;
  return LATTE_OK;
}

namespace {
bool comm_ok(const latte_comm_t* c) {
  if (!c || c->world < 2 || c->world > LATTE_COMM_MAX_RANKS || c->rank < 0 || c->rank >= c->world ||
      c->gen < 1)
    return false;
  return true;
}
PeerFlags comm_flags(const latte_comm_t* c) {
  PeerFlags f;
  for (int w = 0; w < LATTE_COMM_MAX_RANKS; ++w) f.p[w] = w < c->world ? c->flags[w] : nullptr;
  return f;
}
}  // namespace

static int clip_fwd_impl(const void* img_loc, int64_t ld_img_loc, const void* txt_loc,
                         int64_t ld_txt_loc, const void* img_all, int64_t ld_img_all,
                         const void* txt_all, int64_t ld_txt_all, int dtype, int64_t n_loc,
                         int64_t n_all, int64_t dim, int64_t label_offset,
                         const float* logit_scale, float* row_lse, float* col_lse,
                         float* row_nll, float* col_nll, float* loss, float* stats, void* workspace,
                         size_t workspace_bytes, void* stream, StageTimer* tm) {
  LATTE_CHECK_ARG(img_loc && txt_loc && img_all && txt_all && logit_scale && row_lse && col_lse &&
                  loss && workspace);
  LATTE_CHECK_ARG(n_loc > 0 && n_all >= n_loc && dim > 0);
  LATTE_CHECK_ARG(label_offset >= 0 && label_offset + n_loc <= n_all);
  LATTE_CHECK_ARG(dtype >= LATTE_F32 && dtype <= LATTE_F16);
  LATTE_CHECK_ARG(ld_img_loc >= dim && ld_txt_loc >= dim && ld_img_all >= dim && ld_txt_all >= dim);
  const WsLayout w = ws_layout(n_loc, n_all, dim, dtype);
  if (workspace_bytes < w.total) return LATTE_ERR_WORKSPACE;
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return LATTE_ERR_BAD_ARG;
  float* ws = static_cast<float*>(workspace);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!row_nll) row_nll = ws + w.off_nll_r;
  if (!col_nll) col_nll = ws + w.off_nll_c;

  const bool tc = use_tc(dtype, dim, img_loc, ld_img_loc, txt_all, ld_txt_all, txt_loc,
                         ld_txt_loc, img_all, ld_img_all);
  int nparts = 1;
  if (tc) {
    nparts = clip_tc_nparts(n_loc, n_all, device_sm_count());
    if (nparts > kMaxParts) nparts = kMaxParts;
  }
  ClipFwdArgs a;
  a.dtype = dtype; a.n_loc = n_loc; a.n_all = n_all; a.dim = dim;
  a.label_offset = label_offset; a.logit_scale = logit_scale; a.nparts = nparts;
  a.gate = nullptr;
  const unsigned rblocks = (unsigned)((n_loc + 255) / 256);
  double* lossp = reinterpret_cast<double*>(ws + w.off_lossp);

  // ---- CTA-pair path (clip_pair.cu): TS-mode sweep; for one rank the column sums come from
  // the same logit tiles as the row sums, so S is computed once.
  if (tc && w.pair_fwd &&
      clip_pair_supported(dtype, dim, ld_img_loc, ld_txt_all, img_loc, txt_all) &&
      clip_pair_supported(dtype, dim, ld_txt_loc, ld_img_all, txt_loc, img_all)) {
    const PairFwdGeom f = clip_pair_fwd_geom(n_loc, n_all);
    const bool single = n_loc == n_all && img_loc == img_all && txt_loc == txt_all;
    int* flag = reinterpret_cast<int*>(ws + w.off_flag);
    unsigned int* counter = reinterpret_cast<unsigned int*>(flag + 1);
    PairFwdArgs pa;
    pa.dtype = dtype; pa.n_loc = n_loc; pa.n_all = n_all; pa.dim = dim;
    pa.label_offset = label_offset; pa.logit_scale = logit_scale;
    pa.x = img_loc; pa.ldx = ld_img_loc; pa.y = txt_all; pa.ldy = ld_txt_all;
    pa.part_max = ws + w.off_pp_max_r; pa.part_sum = ws + w.off_pp_sum_r; pa.diag = ws + w.off_diag_r;
    pa.col_part = single ? ws + w.off_colpart : nullptr;
    pa.col_ref = ws + w.off_colref;
    pa.zero2 = flag;                       // the sweep clears the fallback flag and the loss counter
    LATTE_MARK(-1);
    int rc = clip_pair_fwd_sweep(pa, st);
    if (rc) return rc;
    LATTE_MARK(LATTE_STAGE_FWD_SWEEP);
    PairFinalizeArgs fa = {};
    fa.pmax = pa.part_max; fa.psum = pa.part_sum; fa.n_loc = n_loc; fa.col_tiles = f.col_tiles;
    fa.total = f.total; fa.ncl = f.ncl; fa.diag = pa.diag; fa.logit_scale = logit_scale;
    fa.row_lse = row_lse; fa.row_nll = row_nll; fa.label_logit = nullptr;
    if (single) {
      fa.col_part = ws + w.off_colpart; fa.col_ref = ws + w.off_colref; fa.ld = f.ld_colpart;
      fa.nblk = 2 * f.row_blocks; fa.n_all = n_all;
      fa.col_lse = col_lse; fa.col_nll = col_nll; fa.col_ml = nullptr; fa.flag = flag;
      pair_fwd_finalize_kernel<<<(unsigned)((n_all + 63) / 64), 256, 0, st>>>(fa);
      LATTE_LAUNCH_OK();
      // exact fallback for columns whose partial sums may have lost flushed terms: the row
      // kernel on the transposed problem; it and the merge return at once unless flagged
      a.gate = flag;
      a.x = txt_loc; a.ldx = ld_txt_loc; a.y = img_all; a.ldy = ld_img_all;
      a.part_max = ws + w.off_pmax_c; a.part_sum = ws + w.off_psum_c; a.diag = ws + w.off_diag_c;
      rc = clip_fwd_rows_tc(a, st);
      if (rc) return rc;
      pair_fwd_finish_kernel<<<rblocks, 256, 0, st>>>(
          flag, a.part_max, a.part_sum, nparts, n_loc, pa.diag, logit_scale, row_lse, col_lse, col_nll,
          row_nll, col_nll, 0, n_loc, lossp, reinterpret_cast<float*>(lossp + rblocks), counter, loss,
          stats);
      LATTE_LAUNCH_OK();
    } else {
      pair_fwd_finalize_kernel<<<rblocks, 256, 0, st>>>(fa);
      LATTE_LAUNCH_OK();
      LATTE_MARK(LATTE_STAGE_FWD_FINALIZE);
      pa.x = txt_loc; pa.ldx = ld_txt_loc; pa.y = img_all; pa.ldy = ld_img_all;
      pa.part_max = ws + w.off_pp_max_c; pa.part_sum = ws + w.off_pp_sum_c; pa.diag = ws + w.off_diag_c;
      pa.col_part = nullptr;
      pa.zero2 = nullptr;
      rc = clip_pair_fwd_sweep(pa, st);
      if (rc) return rc;
      LATTE_MARK(LATTE_STAGE_FWD_SWEEP);
      fa.pmax = pa.part_max; fa.psum = pa.part_sum; fa.diag = pa.diag;
      fa.row_lse = col_lse; fa.row_nll = col_nll;
      pair_fwd_finalize_kernel<<<rblocks, 256, 0, st>>>(fa);
      LATTE_LAUNCH_OK();
      // statistics of the local rows only: the caller's backward sees all-gathered vectors
      pair_fwd_finish_kernel<<<rblocks, 256, 0, st>>>(
          nullptr, nullptr, nullptr, 0, n_loc, nullptr, nullptr, row_lse, col_lse, nullptr, row_nll,
          col_nll, 0, n_loc, lossp, reinterpret_cast<float*>(lossp + rblocks), counter, loss, nullptr);
      LATTE_LAUNCH_OK();
    }
    LATTE_MARK(LATTE_STAGE_FWD_FINALIZE);
    return LATTE_OK;
  }
  LATTE_MARK(-1);

  // image -> text: rows of logits_per_image (loss.py:109 / :115)
  a.x = img_loc; a.ldx = ld_img_loc; a.y = txt_all; a.ldy = ld_txt_all;
  a.part_max = ws + w.off_pmax_r; a.part_sum = ws + w.off_psum_r; a.diag = ws + w.off_diag_r;
  int rc = tc ? clip_fwd_rows_tc(a, st) : clip_fwd_rows_simt(a, st);
  if (rc) return rc;
  // text -> image: rows of logits_per_text (loss.py:110 / :116)
  a.x = txt_loc; a.ldx = ld_txt_loc; a.y = img_all; a.ldy = ld_img_all;
  a.part_max = ws + w.off_pmax_c; a.part_sum = ws + w.off_psum_c; a.diag = ws + w.off_diag_c;
  rc = tc ? clip_fwd_rows_tc(a, st) : clip_fwd_rows_simt(a, st);
  if (rc) return rc;
  LATTE_MARK(LATTE_STAGE_FWD_SWEEP);
  const int fblocks = (int)((n_loc + kFinalRows - 1) / kFinalRows);
  clip_finalize_kernel<<<fblocks, kFinalRows, 0, st>>>(
      ws + w.off_pmax_r, ws + w.off_psum_r, ws + w.off_diag_r, ws + w.off_pmax_c,
      ws + w.off_psum_c, ws + w.off_diag_c, nparts, n_loc, logit_scale, row_lse, col_lse, row_nll,
      col_nll, lossp);
  LATTE_LAUNCH_OK();
  loss_reduce_kernel<<<1, 256, 0, st>>>(lossp, fblocks, n_loc, loss);
  LATTE_LAUNCH_OK();
  LATTE_MARK(LATTE_STAGE_FWD_FINALIZE);
  return LATTE_OK;
}

extern "C" int latte_clip_fwd(const void* img_loc, int64_t ld_img_loc, const void* txt_loc,
                              int64_t ld_txt_loc, const void* img_all, int64_t ld_img_all,
                              const void* txt_all, int64_t ld_txt_all, int dtype, int64_t n_loc,
                              int64_t n_all, int64_t dim, int64_t label_offset,
                              const float* logit_scale, float* row_lse, float* col_lse,
                              float* row_nll, float* col_nll, float* loss, float* stats,
                              void* workspace, size_t workspace_bytes, void* stream) {
  return clip_fwd_impl(img_loc, ld_img_loc, txt_loc, ld_txt_loc, img_all, ld_img_all, txt_all,
                       ld_txt_all, dtype, n_loc, n_all, dim, label_offset, logit_scale, row_lse,
                       col_lse, row_nll, col_nll, loss, stats, workspace, workspace_bytes, stream,
                       nullptr);
}

// ---- multi-rank forward with ONE logit sweep per rank ---------------------------------------
extern "C" int latte_clip_rank_sweep_supported(int dtype, int64_t dim) {
  return pair_shape_ok(dtype, dim) ? 1 : 0;
}

static int clip_fwd_rows_impl(const void* img_loc, int64_t ld_img_loc, const void* txt_all,
                                   int64_t ld_txt_all, int dtype, int64_t n_loc, int64_t n_all,
                                   int64_t dim, int64_t label_offset, const float* logit_scale,
                                   float* row_lse, float* row_nll, float* label_logit,
                                   float* col_ml, void* workspace, size_t workspace_bytes,
                                   void* stream, StageTimer* tm) {
  LATTE_CHECK_ARG(img_loc && txt_all && logit_scale && row_lse && row_nll && label_logit && col_ml &&
                  workspace);
  LATTE_CHECK_ARG(n_loc > 0 && n_all >= n_loc && dim > 0);
  LATTE_CHECK_ARG(label_offset >= 0 && label_offset + n_loc <= n_all);
  LATTE_CHECK_ARG(ld_img_loc >= dim && ld_txt_all >= dim);
  if (!pair_shape_ok(dtype, dim) ||
      !clip_pair_supported(dtype, dim, ld_img_loc, ld_txt_all, img_loc, txt_all))
    return LATTE_ERR_UNSUPPORTED;
  const WsLayout w = ws_layout(n_loc, n_all, dim, dtype);
  if (workspace_bytes < w.total) return LATTE_ERR_WORKSPACE;
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return LATTE_ERR_BAD_ARG;
  float* ws = static_cast<float*>(workspace);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const PairFwdGeom f = clip_pair_fwd_geom(n_loc, n_all);
  PairFwdArgs pa;
  pa.dtype = dtype; pa.n_loc = n_loc; pa.n_all = n_all; pa.dim = dim;
  pa.label_offset = label_offset; pa.logit_scale = logit_scale;
  pa.x = img_loc; pa.ldx = ld_img_loc; pa.y = txt_all; pa.ldy = ld_txt_all;
  pa.part_max = ws + w.off_pp_max_r; pa.part_sum = ws + w.off_pp_sum_r; pa.diag = ws + w.off_diag_r;
  pa.col_part = ws + w.off_colpart;
  pa.col_ref = ws + w.off_colref;
  pa.zero2 = nullptr;
  LATTE_MARK(-1);
  int rc = clip_pair_fwd_sweep(pa, st);
  if (rc) return rc;
  LATTE_MARK(LATTE_STAGE_FWD_SWEEP);
  PairFinalizeArgs fa = {};
  fa.pmax = pa.part_max; fa.psum = pa.part_sum; fa.n_loc = n_loc; fa.col_tiles = f.col_tiles;
  fa.total = f.total; fa.ncl = f.ncl; fa.diag = pa.diag; fa.logit_scale = logit_scale;
  fa.row_lse = row_lse; fa.row_nll = row_nll; fa.label_logit = label_logit;
  fa.col_part = ws + w.off_colpart; fa.col_ref = ws + w.off_colref; fa.ld = f.ld_colpart;
  fa.nblk = 2 * f.row_blocks; fa.n_all = n_all;
  fa.col_lse = nullptr; fa.col_nll = nullptr; fa.col_ml = col_ml; fa.flag = nullptr;
  pair_fwd_finalize_kernel<<<(unsigned)((n_all + 63) / 64), 256, 0, st>>>(fa);
  LATTE_LAUNCH_OK();
  LATTE_MARK(LATTE_STAGE_FWD_FINALIZE);
  return LATTE_OK;
}

extern "C" int latte_clip_fwd_rows(const void* img_loc, int64_t ld_img_loc, const void* txt_all,
                                   int64_t ld_txt_all, int dtype, int64_t n_loc, int64_t n_all,
                                   int64_t dim, int64_t label_offset, const float* logit_scale,
                                   float* row_lse, float* row_nll, float* label_logit,
                                   float* col_ml, void* workspace, size_t workspace_bytes,
                                   void* stream) {
  return clip_fwd_rows_impl(img_loc, ld_img_loc, txt_all, ld_txt_all, dtype, n_loc, n_all, dim,
                            label_offset, logit_scale, row_lse, row_nll, label_logit, col_ml,
                            workspace, workspace_bytes, stream, nullptr);
}

extern "C" int latte_clip_fwd_cols_workspace_bytes(int64_t n_all, int64_t dim, int dtype,
                                                   size_t* bytes) {
  LATTE_CHECK_ARG(bytes && n_all > 0 && dim > 0);
  LATTE_CHECK_ARG(dtype >= LATTE_F32 && dtype <= LATTE_F16);
  *bytes = ws_layout(n_all, n_all, dim, dtype).total;
  return LATTE_OK;
}

extern "C" int latte_clip_fwd_cols(const float* gathered, int64_t stride, int world,
                                   const void* img_all, int64_t ld_img_all, const void* txt_all,
                                   int64_t ld_txt_all, int dtype, int64_t n_loc, int64_t n_all,
                                   int64_t dim, int64_t label_offset, const float* logit_scale,
                                   float* row_lse_all, float* row_nll_all, float* col_lse_all,
                                   float* col_nll_all, float* loss, float* stats, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  LATTE_CHECK_ARG(gathered && img_all && txt_all && logit_scale && row_lse_all && row_nll_all &&
                  col_lse_all && col_nll_all && loss && workspace);
  LATTE_CHECK_ARG(world > 0 && n_loc > 0 && n_all == n_loc * world && dim > 0);
  LATTE_CHECK_ARG(stride >= 2 * n_all + 3 * n_loc);
  LATTE_CHECK_ARG(label_offset >= 0 && label_offset + n_loc <= n_all);
  if (!pair_shape_ok(dtype, dim) || !clip_tc_supported(dtype, dim, ld_txt_all, ld_img_all, txt_all, img_all))
    return LATTE_ERR_UNSUPPORTED;
  const WsLayout w = ws_layout(n_all, n_all, dim, dtype);
  if (workspace_bytes < w.total) return LATTE_ERR_WORKSPACE;
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return LATTE_ERR_BAD_ARG;
  float* ws = static_cast<float*>(workspace);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int* flag = reinterpret_cast<int*>(ws + w.off_flag);
  float* label_logit_all = ws + w.off_nll_r;           // [n_all] scratch of this layout
  LATTE_CUDA_OK(cudaMemsetAsync(flag, 0, 2 * sizeof(int), st));     // fallback flag + loss counter
  unsigned int* counter = reinterpret_cast<unsigned int*>(flag + 1);
  const int nblk_total = (int)((n_all + 127) / 128);
  col_merge_kernel<<<(unsigned)((n_all + 255) / 256), 256, 0, st>>>(
      gathered, stride, world, n_loc, n_all, nblk_total, row_lse_all, row_nll_all, label_logit_all,
      col_lse_all, col_nll_all, flag, nullptr, 0);
  LATTE_LAUNCH_OK();
  // exact fallback (every column, from the gathered features): gated on the flag
  int nparts = clip_tc_nparts(n_all, n_all, device_sm_count());
  if (nparts > kMaxParts) nparts = kMaxParts;
  ClipFwdArgs a;
  a.dtype = dtype; a.n_loc = n_all; a.n_all = n_all; a.dim = dim;
  a.label_offset = 0; a.logit_scale = logit_scale; a.nparts = nparts;
  a.gate = flag;
  a.x = txt_all; a.ldx = ld_txt_all; a.y = img_all; a.ldy = ld_img_all;
  a.part_max = ws + w.off_pmax_c; a.part_sum = ws + w.off_psum_c; a.diag = ws + w.off_diag_c;
  int rc = clip_fwd_rows_tc(a, st);
  if (rc) return rc;
  double* lossp = reinterpret_cast<double*>(ws + w.off_lossp);
  const unsigned fblocks = (unsigned)((n_all + 255) / 256);
  pair_fwd_finish_kernel<<<fblocks, 256, 0, st>>>(
      flag, a.part_max, a.part_sum, nparts, n_all, label_logit_all, nullptr, row_lse_all, col_lse_all,
      col_nll_all, row_nll_all, col_nll_all, label_offset, n_loc, lossp,
      reinterpret_cast<float*>(lossp + fblocks), counter, loss, stats);
  LATTE_LAUNCH_OK();
  return LATTE_OK;
}

// ---- multi-rank forward over peer memory --------------------------------------------------------
extern "C" int latte_clip_fwd_rank_workspace_bytes(int64_t n_loc, int64_t n_all, int64_t dim, int dtype,
                                                   size_t* bytes) {
  LATTE_CHECK_ARG(bytes && n_loc > 0 && n_all >= n_loc && dim > 0);
  LATTE_CHECK_ARG(dtype >= LATTE_F32 && dtype <= LATTE_F16);
  *bytes = ws_layout(n_loc, n_all, dim, dtype).total + ws_layout(n_all, n_all, dim, dtype).total;
  return LATTE_OK;
}

extern "C" int latte_clip_fwd_rank(const latte_comm_t* comm, const void* img_loc, int64_t ld_img_loc,
                                   const void* txt_all, int64_t ld_txt_all, int dtype, int64_t n_loc,
                                   int64_t n_all, int64_t dim, int64_t label_offset,
                                   const float* logit_scale, float* row_lse_all, float* row_nll_all,
                                   float* col_lse_all, float* col_nll_all, float* loss, float* stats,
                                   int phases, void* workspace, size_t workspace_bytes, void* stream) {
  LATTE_CHECK_ARG(comm_ok(comm) && img_loc && txt_all && logit_scale && row_lse_all && row_nll_all &&
                  col_lse_all && col_nll_all && loss && workspace);
  LATTE_CHECK_ARG(n_loc > 0 && n_all == n_loc * comm->world && dim > 0);
  LATTE_CHECK_ARG(label_offset == (int64_t)comm->rank * n_loc);
  LATTE_CHECK_ARG(ld_img_loc >= dim && ld_txt_all >= dim);
  LATTE_CHECK_ARG(comm->payload_stride >= 2 * n_all + 3 * n_loc);
  if (!pair_shape_ok(dtype, dim) ||
      !clip_pair_supported(dtype, dim, ld_img_loc, ld_txt_all, img_loc, txt_all) ||
      !clip_tc_supported(dtype, dim, ld_txt_all, ld_img_loc, txt_all, img_loc))
    return LATTE_ERR_UNSUPPORTED;
  const WsLayout wa = ws_layout(n_loc, n_all, dim, dtype);
  const WsLayout wb = ws_layout(n_all, n_all, dim, dtype);
  if (workspace_bytes < wa.total + wb.total) return LATTE_ERR_WORKSPACE;
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return LATTE_ERR_BAD_ARG;
  float* ws = static_cast<float*>(workspace);
  float* wsb = ws + wa.total / sizeof(float);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int world = comm->world, rank = comm->rank, gen = comm->gen;
  const int64_t stride = comm->payload_stride;
  float* my_block = comm->payload[rank];
  LATTE_CHECK_ARG(my_block != nullptr);
  float* my_row = my_block + (int64_t)rank * stride;          // col_ml | row_lse | row_nll | label_logit
  float* exact_block = my_block + (int64_t)world * stride;    // [world][2 n_all]
  int* my_flags = comm->flags[rank];                          // NULL: no waits / signals (tests)
  const PeerFlags pflags = comm_flags(comm);
  PeerPtrs blocks;
  for (int w = 0; w < LATTE_COMM_MAX_RANKS; ++w) blocks.p[w] = w < world ? comm->payload[w] : nullptr;
  int* gate = reinterpret_cast<int*>(wsb + wb.off_flag);
  unsigned int* counter = reinterpret_cast<unsigned int*>(gate + 1);
  float* label_logit_all = wsb + wb.off_nll_r;                // [n_all] scratch of this layout
  const PairFwdGeom f = clip_pair_fwd_geom(n_loc, n_all);
  int nparts = clip_tc_nparts(n_all, n_loc, device_sm_count());
  if (nparts > kMaxParts) nparts = kMaxParts;

  if (phases & 1) {
    PairFwdArgs pa;
    pa.dtype = dtype; pa.n_loc = n_loc; pa.n_all = n_all; pa.dim = dim;
    pa.label_offset = label_offset; pa.logit_scale = logit_scale;
    pa.x = img_loc; pa.ldx = ld_img_loc; pa.y = txt_all; pa.ldy = ld_txt_all;
    pa.part_max = ws + wa.off_pp_max_r; pa.part_sum = ws + wa.off_pp_sum_r; pa.diag = ws + wa.off_diag_r;
    pa.col_part = ws + wa.off_colpart;
    pa.col_ref = ws + wa.off_colref;
    pa.zero2 = gate;                         // clears the exactness flag and the loss counter
    pa.landed = my_flags ? my_flags + kFlagLandedTxt : nullptr;
    pa.landed_gen = gen;
    pa.rows_per_rank = n_loc;
    int rc = clip_pair_fwd_sweep(pa, st);
    if (rc) return rc;
    PairFinalizeArgs fa = {};
    fa.pmax = pa.part_max; fa.psum = pa.part_sum; fa.n_loc = n_loc; fa.col_tiles = f.col_tiles;
    fa.total = f.total; fa.ncl = f.ncl; fa.diag = pa.diag; fa.logit_scale = logit_scale;
    fa.col_ml = my_row;
    fa.row_lse = my_row + 2 * n_all; fa.row_nll = fa.row_lse + n_loc; fa.label_logit = fa.row_nll + n_loc;
    fa.col_part = ws + wa.off_colpart; fa.col_ref = ws + wa.off_colref; fa.ld = f.ld_colpart;
    fa.nblk = 2 * f.row_blocks; fa.n_all = n_all;
    pair_fwd_finalize_kernel<<<(unsigned)((n_all + 63) / 64), 256, 0, st>>>(fa);
    LATTE_LAUNCH_OK();
    const int64_t len = 2 * n_all + 3 * n_loc;
    comm_payload_kernel<<<(unsigned)((len + 1023) / 1024), 256, 0, st>>>(
        nullptr, nullptr, nullptr, 0, 0, my_row, len, (int64_t)rank * stride, blocks, pflags, my_flags,
        world, rank, gen, kFlagPayload, kCntPayload);
    LATTE_LAUNCH_OK();
  }
  if (phases & 2) {
    const int nblk_total = (int)((n_all + 127) / 128);
    col_merge_kernel<<<(unsigned)((n_all + 255) / 256), 256, 0, st>>>(
        my_block, stride, world, n_loc, n_all, nblk_total, row_lse_all, row_nll_all, label_logit_all,
        col_lse_all, col_nll_all, gate, my_flags ? my_flags + kFlagPayload : nullptr, gen);
    LATTE_LAUNCH_OK();
    // Gated second round (the merged statistics, hence the flag, are bit-identical on all ranks):
    // this rank's column partials again, exactly -- all texts x own images with the row kernel --
    // and their exchange.  The three kernels return at once when the flag is clear.
    ClipFwdArgs a;
    a.dtype = dtype; a.n_loc = n_all; a.n_all = n_loc; a.dim = dim;
    a.label_offset = n_loc;                  // no label column: the label logits are known already
    a.logit_scale = logit_scale; a.nparts = nparts; a.gate = gate;
    a.x = txt_all; a.ldx = ld_txt_all; a.y = img_loc; a.ldy = ld_img_loc;
    a.part_max = wsb + wb.off_pmax_c; a.part_sum = wsb + wb.off_psum_c; a.diag = wsb + wb.off_diag_c;
    int rc = clip_fwd_rows_tc(a, st);
    if (rc) return rc;
    const int64_t len2 = 2 * n_all;
    comm_payload_kernel<<<(unsigned)((len2 + 1023) / 1024), 256, 0, st>>>(
        gate, a.part_max, a.part_sum, nparts, n_all, exact_block + (int64_t)rank * len2, len2,
        (int64_t)world * stride + (int64_t)rank * len2, blocks, pflags, my_flags, world, rank, gen,
        kFlagPayload2, kCntPayload2);
    LATTE_LAUNCH_OK();
  }
  if (phases & 4) {
    col_merge_exact_kernel<<<(unsigned)((n_all + 255) / 256), 256, 0, st>>>(
        gate, exact_block, world, n_all, label_logit_all, col_lse_all, col_nll_all,
        my_flags ? my_flags + kFlagPayload2 : nullptr, gen);
    LATTE_LAUNCH_OK();
    double* lossp = reinterpret_cast<double*>(wsb + wb.off_lossp);
    const unsigned fblocks = (unsigned)((n_all + 255) / 256);
    pair_fwd_finish_kernel<<<fblocks, 256, 0, st>>>(
        nullptr, nullptr, nullptr, 0, n_all, nullptr, nullptr, row_lse_all, col_lse_all, nullptr,
        row_nll_all, col_nll_all, label_offset, n_loc, lossp, reinterpret_cast<float*>(lossp + fblocks),
        counter, loss, stats);
    LATTE_LAUNCH_OK();
  }
  return LATTE_OK;
}

// cast accumulator -> d_txt, clear it, release the slot (three stream-ordered operations)
static int comm_finish_reduce_scatter(const latte_comm_t* comm, int64_t n_loc, int64_t dim, void* d_txt,
                                      int grad_dtype, int64_t ld_grad, cudaStream_t st) {
  float* acc = comm->acc[comm->rank];
  comm_acc_finish_kernel<<<(unsigned)(8 * device_sm_count()), 256, 0, st>>>(
      acc, n_loc, dim, d_txt, grad_dtype, ld_grad, comm->flags[comm->rank], comm->world, comm->gen);
  LATTE_LAUNCH_OK();
  LATTE_CUDA_OK(cudaMemsetAsync(acc, 0, (size_t)n_loc * (size_t)dim * sizeof(float), st));
  if (comm->flags[comm->rank]) {
    comm_release_kernel<<<1, 32, 0, st>>>(comm_flags(comm), comm->world, comm->rank, comm->gen);
    LATTE_LAUNCH_OK();
  }
  return LATTE_OK;
}

static int clip_bwd_impl(const void* img_loc, int64_t ld_img_loc, const void* txt_loc,
                         int64_t ld_txt_loc, const void* img_all, int64_t ld_img_all,
                         const void* txt_all, int64_t ld_txt_all, int dtype, int64_t n_loc,
                         int64_t n_all, int64_t dim, int64_t label_offset,
                         const float* logit_scale, const float* row_lse_all,
                         const float* col_lse_all, const float* row_nll_all,
                         const float* col_nll_all, const float* lse_stats, const float* grad_loss,
                         float grad_mult, int cross_terms, void* d_img, void* d_txt, int grad_dtype,
                         int64_t ld_grad, float* d_txt_partial, const latte_comm_t* comm, int phases,
                         float* d_scale, void* workspace,
                         size_t workspace_bytes, void* stream, StageTimer* tm) {
  LATTE_CHECK_ARG(img_loc && txt_loc && txt_all && logit_scale && row_lse_all &&
                  col_lse_all && grad_loss && d_img && (d_txt || d_txt_partial) &&
                  d_scale && workspace);
  LATTE_CHECK_ARG(!(d_txt_partial && comm));
  LATTE_CHECK_ARG(!comm || (comm_ok(comm) && n_all == n_loc * comm->world && d_txt &&
                            comm->acc[comm->rank] && (dim % 4) == 0 && (ld_grad % 4) == 0));
  LATTE_CHECK_ARG((row_nll_all == nullptr) == (col_nll_all == nullptr));
  LATTE_CHECK_ARG(n_loc > 0 && n_all >= n_loc && dim > 0);
  LATTE_CHECK_ARG(label_offset >= 0 && label_offset + n_loc <= n_all);
  LATTE_CHECK_ARG(dtype >= LATTE_F32 && dtype <= LATTE_F16);
  LATTE_CHECK_ARG(grad_dtype >= LATTE_F32 && grad_dtype <= LATTE_F16);
  LATTE_CHECK_ARG(ld_img_loc >= dim && ld_txt_loc >= dim && ld_txt_all >= dim && ld_grad >= dim);
  // one-sweep multi-rank modes (fp32 partial for a reduce-scatter, or peer accumulators) read only
  // this rank's images: img_all may be NULL there
  const bool one_sweep = (d_txt_partial != nullptr || comm != nullptr) && cross_terms;
  if ((d_txt_partial || comm) && !one_sweep) return LATTE_ERR_BAD_ARG;
  if (!one_sweep) LATTE_CHECK_ARG(img_all && ld_img_all >= dim);
  if (one_sweep) { img_all = img_loc; ld_img_all = ld_img_loc; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (comm && !(phases & 1)) {
    // second phase only (single-process tests drive the ranks phase by phase)
    return comm_finish_reduce_scatter(comm, n_loc, dim, d_txt, grad_dtype, ld_grad, st);
  }
  const WsLayout w = ws_layout(n_loc, n_all, dim, dtype, true);
  if (workspace_bytes < w.total) return LATTE_ERR_WORKSPACE;
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return LATTE_ERR_BAD_ARG;
  float* ws = static_cast<float*>(workspace);

  float* row2 = ws + w.off_row2;
  float* col2 = ws + w.off_col2;
  LATTE_MARK(-1);
  float* rho = ws + w.off_rho;
  int* fast_flag = reinterpret_cast<int*>(ws + w.off_rho + 1);
  unsigned int* u_bits = reinterpret_cast<unsigned int*>(ws + w.off_rho + 2);
  float* gscale = ws + w.off_rho + 3;
  float* out_scale = ws + w.off_rho + 4;
  const bool pair_ok =
      w.pair && clip_pair_supported(dtype, dim, ld_img_loc, ld_txt_all, img_loc, txt_all) &&
      clip_pair_supported(dtype, dim, ld_txt_loc, ld_img_all, txt_loc, img_all) &&
      clip_pair_supported(dtype, dim, ld_img_all, ld_txt_all, img_all, txt_all);
  if (one_sweep && !pair_ok) return LATTE_ERR_UNSUPPORTED;
  // G is fp16 (a bf16 G would cost 2^-9 per weight) and tcgen05.mma kind::f16 wants A and B in ONE
  // 16-bit format (an fp16 x bf16 instruction descriptor raises "illegal instruction" on sm_100a,
  // measured in round 2), so bf16 features get fp16 copies for the gradient GEMMs; fp16 features
  // are used as they are
  const bool copies16 = pair_ok && dtype == LATTE_BF16;
  const bool have_nll = row_nll_all != nullptr;
  const bool have_bound = have_nll || lse_stats != nullptr;
  __half* ya = reinterpret_cast<__half*>(ws + w.off_y16a);                               // txt_all
  __half* yb = (!one_sweep && img_all == txt_all) ? ya : reinterpret_cast<__half*>(ws + w.off_y16b);
  if (pair_ok) {
    // bound on |G| -> fp16 scale of G: from the forward's statistics or per-sample loss terms when
    // the caller has them, else from the label logits (one pass over the gathered features;
    // one-sweep modes without either use the bound 1)
    const bool dots = !have_bound && !one_sweep;
    if (dots) {
      LATTE_CUDA_OK(cudaMemsetAsync(u_bits, 0, sizeof(unsigned int), st));
      const int64_t warps_needed = n_all;
      const unsigned blocks = (unsigned)((warps_needed * 32 + 255) / 256 < 4096
                                             ? (warps_needed * 32 + 255) / 256 : 4096);
      pair_prep_features_kernel<<<blocks, 256, 0, st>>>(img_all, ld_img_all, txt_all, ld_txt_all,
                                                        dtype == LATTE_BF16 ? 1 : 0,
                                                        copies16 ? yb : nullptr, copies16 ? ya : nullptr,
                                                        (int64_t)w.ld16, n_all, dim, logit_scale,
                                                        row_lse_all, col_lse_all, u_bits);
      LATTE_LAUNCH_OK();
    } else if (copies16) {
      // pure conversions: all gathered texts, and the images the GEMMs read (this rank's rows in
      // the one-sweep modes, all of them otherwise)
      const int64_t img_rows = one_sweep ? n_loc : n_all;
      const int64_t per_row = (dim + 7) / 8;
      if (!one_sweep && ld_img_all == ld_txt_all) {
        const bool same = img_all == txt_all;
        bf16_to_fp16_kernel<<<dim3((unsigned)((n_all * per_row + 255) / 256), same ? 1 : 2), 256, 0, st>>>(
            static_cast<const __nv_bfloat16*>(txt_all), ya, static_cast<const __nv_bfloat16*>(img_all),
            yb, ld_txt_all, (int64_t)w.ld16, n_all, dim);
        LATTE_LAUNCH_OK();
      } else {
        bf16_to_fp16_kernel<<<dim3((unsigned)((n_all * per_row + 255) / 256), 1), 256, 0, st>>>(
            static_cast<const __nv_bfloat16*>(txt_all), ya, nullptr, nullptr, ld_txt_all,
            (int64_t)w.ld16, n_all, dim);
        LATTE_LAUNCH_OK();
        bf16_to_fp16_kernel<<<dim3((unsigned)((img_rows * per_row + 255) / 256), 1), 256, 0, st>>>(
            static_cast<const __nv_bfloat16*>(img_all), yb, nullptr, nullptr, ld_img_all,
            (int64_t)w.ld16, img_rows, dim);
        LATTE_LAUNCH_OK();
      }
    }
    if (!lse_stats) {
      float* st_buf = ws + w.off_rho + 8;
      lse_stats_kernel<<<1, 1024, 0, st>>>(row_lse_all, col_lse_all, n_all, dots ? u_bits : nullptr,
                                           have_nll ? row_nll_all : nullptr,
                                           have_nll ? col_nll_all : nullptr, st_buf);
      LATTE_LAUNCH_OK();
      lse_stats = st_buf;
    }
  }
  lse_vectors_kernel<<<(unsigned)((w.n_pad + 255) / 256), 256, 0, st>>>(
      row_lse_all, col_lse_all, n_all, (int64_t)w.n_pad, pair_ok ? lse_stats : nullptr, grad_loss,
      grad_mult, logit_scale, n_loc, rho, row2, col2,
      pair_ok ? ws + w.off_erow : nullptr, ws + w.off_einvrow, ws + w.off_ecol, ws + w.off_einvcol);
  LATTE_LAUNCH_OK();

  const bool tc = use_tc(dtype, dim, img_loc, ld_img_loc, txt_all, ld_txt_all, txt_loc,
                         ld_txt_loc, img_all, ld_img_all);
  const int ds_count = tc ? clip_tc_ds_count(n_loc, dim) : clip_simt_ds_count(n_loc);
  if ((size_t)(2 * ds_count) > w.ds_cap) return LATTE_ERR_WORKSPACE;

  // fp16 operands for the second GEMM of the tc path
  const void* txt16 = txt_all; int64_t ld_txt16 = ld_txt_all;
  const void* img16 = img_all; int64_t ld_img16 = ld_img_all;
  if (pair_ok && dtype == LATTE_BF16) {
    txt16 = ya; ld_txt16 = (int64_t)w.ld16;
    img16 = yb; ld_img16 = (int64_t)w.ld16;
  } else if (tc && dtype == LATTE_BF16) {
    const int64_t work = n_all * ((dim + 7) / 8);
    const unsigned blocks = (unsigned)((work + 255) / 256);
    const bool same = img_all == txt_all;
    if (same || ld_img_all == ld_txt_all) {
      bf16_to_fp16_kernel<<<dim3(blocks, same ? 1 : 2), 256, 0, st>>>(
          static_cast<const __nv_bfloat16*>(txt_all), ya, static_cast<const __nv_bfloat16*>(img_all),
          yb, ld_txt_all, (int64_t)w.ld16, n_all, dim);
      LATTE_LAUNCH_OK();
    } else {
      bf16_to_fp16_kernel<<<dim3(blocks, 1), 256, 0, st>>>(
          static_cast<const __nv_bfloat16*>(txt_all), ya, nullptr, nullptr, ld_txt_all,
          (int64_t)w.ld16, n_all, dim);
      LATTE_LAUNCH_OK();
      bf16_to_fp16_kernel<<<dim3(blocks, 1), 256, 0, st>>>(
          static_cast<const __nv_bfloat16*>(img_all), yb, nullptr, nullptr, ld_img_all,
          (int64_t)w.ld16, n_all, dim);
      LATTE_LAUNCH_OK();
    }
    txt16 = ya; ld_txt16 = (int64_t)w.ld16;
    img16 = yb; ld_img16 = (int64_t)w.ld16;
  }

  // ---- CTA-pair path: one logit recompute -> G, then the gradient GEMMs (clip_pair.cu)
  if (tc && pair_ok && clip_pair_supported(LATTE_F16, dim, ld_txt16, ld_img16, txt16, img16)) {
    const int dsn = clip_pair_ds_count();
    if ((size_t)(2 * dsn) > w.ds_cap) return LATTE_ERR_WORKSPACE;
    float* dsp = ws + w.off_ds;
    __half* gbuf = reinterpret_cast<__half*>(ws + w.off_g);
    float* acc_i = ws + w.off_acc0;
    float* acc_t = ws + w.off_acc1;
    const size_t acc_bytes = (size_t)n_loc * w.ld32 * sizeof(float);
    LATTE_CUDA_OK(cudaMemsetAsync(dsp, 0, (size_t)2 * dsn * sizeof(float), st));
    const bool single = !one_sweep && n_loc == n_all && img_loc == img_all && txt_loc == txt_all &&
                        cross_terms;
    // one sweep per rank: the text-side product G^T . img_loc covers ALL columns; it is returned as
    // an fp32 partial for the caller to reduce-scatter, or added into the owners' accumulators
    // over NVLink from the GEMM epilogue (loss.py:49-50's backward)
    const bool rank_sweep = one_sweep;
    if (!rank_sweep && !d_txt) return LATTE_ERR_BAD_ARG;
    PairSweepArgs sa;
    sa.dtype = dtype; sa.n_loc = n_loc; sa.n_all = n_all; sa.dim = dim;
    sa.label_offset = label_offset; sa.logit_scale = logit_scale; sa.cross_terms = cross_terms;
    sa.g = gbuf; sa.ds_both = (single || rank_sweep) ? 1 : 0;
    sa.nll_a = row_nll_all; sa.nll_b = col_nll_all; sa.no_label = 0;
    PairGemmArgs ga = {};
    ga.g = gbuf; ga.n_loc = n_loc; ga.n_all = n_all; ga.dim = dim; ga.ld32 = (int64_t)w.ld32;
    ga.feat_dtype = LATTE_F16;
    // Gradients of tiles owned by one cluster are written by the GEMM epilogue itself (scaled, in
    // the gradient dtype); the fp32 accumulators only serve the tiles the schedule splits.
    ga.out_dtype = grad_dtype; ga.ld_out = ld_grad; ga.out_scale = out_scale;
    // image side: G[loc rows, :] and d_img = G . txt_all  (+ d_txt = G^T . img for one rank)
    sa.x = img_loc; sa.ldx = ld_img_loc; sa.y = txt_all; sa.ldy = ld_txt_all;
    sa.lse_a2 = row2; sa.lse_b2 = col2; sa.ds_partial = dsp;
    sa.e_a = ws + w.off_erow; sa.einv_b = ws + w.off_einvcol; sa.fast_flag = fast_flag;
    sa.gscale_log2 = gscale;
    ga.y16 = txt16; ga.ldy16 = ld_txt16;
    ga.x16 = single ? img16 : nullptr; ga.ldx16 = ld_img16;
    ga.dx32 = acc_i; ga.dy32 = acc_t;
    ga.ld_dy32 = (int64_t)w.ld32; ga.dy_scale = nullptr;
    ga.dy_peers = nullptr; ga.n_peers = 0;
    ga.dx_out = d_img; ga.dy_out = single ? d_txt : nullptr;
    if (rank_sweep) {
      ga.x16 = img16;                       // this rank's images (fp16 copy or the features themselves)
      ga.dy32 = d_txt_partial; ga.ld_dy32 = dim; ga.dy_scale = out_scale;
      if (comm) {
        ga.dy32 = comm->acc[comm->rank];                     // unused: every row has an owner
        ga.dy_peers = comm->acc;
        ga.n_peers = comm->world;
        // the GEMM publishes done = gen on every rank once all its adds are out
        ga.n_done = comm->flags[comm->rank] ? comm->world : 0;
        for (int q = 0; q < comm->world; ++q)
          ga.done_flags[q] = comm->flags[q] ? comm->flags[q] + kFlagDone : nullptr;
        ga.done_counter = comm->flags[comm->rank]
                              ? reinterpret_cast<unsigned int*>(comm->flags[comm->rank] + kFlagCounter + kCntGemm)
                              : nullptr;
        ga.done_gen = comm->gen;
        ga.done_slot = comm->rank;
      }
      if (d_txt_partial)
        LATTE_CUDA_OK(cudaMemsetAsync(d_txt_partial, 0, (size_t)n_all * (size_t)dim * sizeof(float), st));
    }
    const bool direct_i = clip_pair_gemm_direct(ga, 0);
    const bool direct_t = single && clip_pair_gemm_direct(ga, 1);
    if (!direct_i) LATTE_CUDA_OK(cudaMemsetAsync(acc_i, 0, acc_bytes, st));
    if (single && !direct_t) LATTE_CUDA_OK(cudaMemsetAsync(acc_t, 0, acc_bytes, st));
    int rc = clip_pair_gemm_fixup(ga, 0, st);
    if (rc) return rc;
    LATTE_MARK(LATTE_STAGE_BWD_PREP);
    rc = clip_pair_sweep(sa, st);
    if (rc) return rc;
    LATTE_MARK(LATTE_STAGE_BWD_SWEEP);
    rc = clip_pair_gemm(ga, st);
    if (rc) return rc;
    LATTE_MARK(LATTE_STAGE_BWD_GEMM);
    rc = clip_pair_gemm_fixup(ga, 1, st);
    if (rc) return rc;
    if (!direct_i) {
      rc = clip_pair_scale_cast(acc_i, nullptr, (int64_t)w.ld32, d_img, nullptr, grad_dtype, ld_grad,
                                n_loc, dim, out_scale, st);
      if (rc) return rc;
    }
    if (single && !direct_t) {
      rc = clip_pair_scale_cast(acc_t, nullptr, (int64_t)w.ld32, d_txt, nullptr, grad_dtype, ld_grad,
                                n_loc, dim, out_scale, st);
      if (rc) return rc;
    }
    bool direct_t2 = false;
    if (!single && !rank_sweep) {
      // text side: the transposed block G'[loc cols, :] and d_txt = G' . img_all
      LATTE_MARK(LATTE_STAGE_BWD_FINISH);
      sa.x = txt_loc; sa.ldx = ld_txt_loc; sa.y = img_all; sa.ldy = ld_img_all;
      sa.lse_a2 = col2; sa.lse_b2 = row2; sa.ds_partial = dsp + dsn;
      sa.e_a = ws + w.off_ecol; sa.einv_b = ws + w.off_einvrow;
      sa.nll_a = col_nll_all; sa.nll_b = row_nll_all;
      ga.y16 = img16; ga.ldy16 = ld_img16; ga.x16 = nullptr;
      ga.dx32 = acc_t; ga.dy32 = nullptr; ga.dy_peers = nullptr; ga.n_peers = 0;
      ga.dy_scale = nullptr; ga.dx_out = d_txt; ga.dy_out = nullptr;
      direct_t2 = clip_pair_gemm_direct(ga, 0);
      if (!direct_t2) LATTE_CUDA_OK(cudaMemsetAsync(acc_t, 0, acc_bytes, st));
      rc = clip_pair_gemm_fixup(ga, 0, st);
      if (rc) return rc;
      LATTE_MARK(LATTE_STAGE_BWD_PREP);
      rc = clip_pair_sweep(sa, st);
      if (rc) return rc;
      LATTE_MARK(LATTE_STAGE_BWD_SWEEP);
      rc = clip_pair_gemm(ga, st);
      if (rc) return rc;
      LATTE_MARK(LATTE_STAGE_BWD_GEMM);
      rc = clip_pair_gemm_fixup(ga, 1, st);
      if (rc) return rc;
      if (!direct_t2) {
        rc = clip_pair_scale_cast(acc_t, nullptr, (int64_t)w.ld32, d_txt, nullptr, grad_dtype, ld_grad,
                                  n_loc, dim, out_scale, st);
        if (rc) return rc;
      }
    }
    ds_reduce_kernel<<<1, 256, 0, st>>>(dsp, 2 * dsn, grad_loss, grad_mult, n_loc, d_scale);
    LATTE_LAUNCH_OK();
    if (comm && (phases & 2)) {
      rc = comm_finish_reduce_scatter(comm, n_loc, dim, d_txt, grad_dtype, ld_grad, st);
      if (rc) return rc;
    }
    LATTE_MARK(LATTE_STAGE_BWD_FINISH);
    return LATTE_OK;
  }
  LATTE_MARK(LATTE_STAGE_BWD_PREP);
  if (d_txt_partial || comm || !d_txt) return LATTE_ERR_UNSUPPORTED;

  ClipBwdArgs a;
  a.dtype = dtype; a.n_loc = n_loc; a.n_all = n_all; a.dim = dim;
  a.label_offset = label_offset; a.logit_scale = logit_scale;
  a.grad_loss = grad_loss; a.grad_mult = grad_mult; a.cross_terms = cross_terms;
  a.grad_dtype = grad_dtype; a.ld_dx = ld_grad; a.ds_count = ds_count;
  // d_img = coef*s * G[loc rows, :] @ txt_all
  a.x = img_loc; a.ldx = ld_img_loc; a.y = txt_all; a.ldy = ld_txt_all;
  a.y16 = txt16; a.ldy16 = ld_txt16;
  a.lse_a2 = row2; a.lse_b2 = col2; a.dx = d_img; a.ds_partial = ws + w.off_ds;
  int rc = tc ? clip_bwd_rows_tc(a, st) : clip_bwd_rows_simt(a, st);
  if (rc) return rc;
  // d_txt = coef*s * G[:, loc cols]^T @ img_all  (rows of the transposed problem)
  a.x = txt_loc; a.ldx = ld_txt_loc; a.y = img_all; a.ldy = ld_img_all;
  a.y16 = img16; a.ldy16 = ld_img16;
  a.lse_a2 = col2; a.lse_b2 = row2; a.dx = d_txt; a.ds_partial = ws + w.off_ds + ds_count;
  rc = tc ? clip_bwd_rows_tc(a, st) : clip_bwd_rows_simt(a, st);
  if (rc) return rc;
  ds_reduce_kernel<<<1, 256, 0, st>>>(ws + w.off_ds, 2 * ds_count, grad_loss, grad_mult, n_loc,
                                      d_scale);
  LATTE_LAUNCH_OK();
  LATTE_MARK(LATTE_STAGE_BWD_SWEEP);      // legacy row kernels: sweep and gradient GEMM are fused
  return LATTE_OK;
}

extern "C" int latte_clip_bwd(const void* img_loc, int64_t ld_img_loc, const void* txt_loc,
                              int64_t ld_txt_loc, const void* img_all, int64_t ld_img_all,
                              const void* txt_all, int64_t ld_txt_all, int dtype, int64_t n_loc,
                              int64_t n_all, int64_t dim, int64_t label_offset,
                              const float* logit_scale, const float* row_lse_all,
                              const float* col_lse_all, const float* row_nll_all,
                              const float* col_nll_all, const float* lse_stats,
                              const float* grad_loss, float grad_mult,
                              int cross_terms, void* d_img, void* d_txt, int grad_dtype,
                              int64_t ld_grad, float* d_txt_partial, const latte_comm_t* comm,
                              int phases, float* d_scale, void* workspace,
                              size_t workspace_bytes, void* stream) {
  return clip_bwd_impl(img_loc, ld_img_loc, txt_loc, ld_txt_loc, img_all, ld_img_all, txt_all,
                       ld_txt_all, dtype, n_loc, n_all, dim, label_offset, logit_scale, row_lse_all,
                       col_lse_all, row_nll_all, col_nll_all, lse_stats, grad_loss, grad_mult,
                       cross_terms, d_img,
                       d_txt, grad_dtype, ld_grad, d_txt_partial, comm, phases, d_scale,
                       workspace, workspace_bytes, stream, nullptr);
}

// One real backward (any mode, including the peer-memory reduce-scatter -- every rank then calls
// this for the same generation) with CUDA events around its stages; synchronises the stream and
// returns the milliseconds per stage of THIS call (forward stages stay 0).
extern "C" int latte_clip_bwd_stage_times(const void* img_loc, int64_t ld_img_loc, const void* txt_loc,
                                          int64_t ld_txt_loc, const void* img_all, int64_t ld_img_all,
                                          const void* txt_all, int64_t ld_txt_all, int dtype, int64_t n_loc,
                                          int64_t n_all, int64_t dim, int64_t label_offset,
                                          const float* logit_scale, const float* row_lse_all,
                                          const float* col_lse_all, const float* row_nll_all,
                                          const float* col_nll_all, const float* lse_stats,
                                          const float* grad_loss, float grad_mult,
                                          int cross_terms, void* d_img, void* d_txt, int grad_dtype,
                                          int64_t ld_grad, float* d_txt_partial, const latte_comm_t* comm,
                                          int phases, float* d_scale, void* workspace,
                                          size_t workspace_bytes, void* stream, float* stage_ms) {
  LATTE_CHECK_ARG(stage_ms);
  for (int k = 0; k < LATTE_NUM_STAGES; ++k) stage_ms[k] = 0.f;
  StageTimer tm;
  const int rc = clip_bwd_impl(img_loc, ld_img_loc, txt_loc, ld_txt_loc, img_all, ld_img_all, txt_all,
                               ld_txt_all, dtype, n_loc, n_all, dim, label_offset, logit_scale,
                               row_lse_all, col_lse_all, row_nll_all, col_nll_all, lse_stats, grad_loss,
                               grad_mult, cross_terms, d_img, d_txt, grad_dtype, ld_grad, d_txt_partial,
                               comm, phases, d_scale, workspace, workspace_bytes, stream, &tm);
  const cudaError_t e = cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
  tm.collect(stage_ms);
  if (rc) return rc;
  return (e != cudaSuccess || tm.failed) ? LATTE_ERR_CUDA : LATTE_OK;
}

extern "C" int latte_clip_stage_times(const void* img_loc, int64_t ld_img_loc, const void* txt_loc,
                                      int64_t ld_txt_loc, const void* img_all, int64_t ld_img_all,
                                      const void* txt_all, int64_t ld_txt_all, int dtype,
                                      int64_t n_loc, int64_t n_all, int64_t dim,
                                      int64_t label_offset, const float* logit_scale,
                                      const float* row_lse_all, const float* col_lse_all,
                                      float* row_lse, float* col_lse, float* loss,
                                      const float* grad_loss, float grad_mult, int cross_terms,
                                      void* d_img, void* d_txt, int grad_dtype, int64_t ld_grad,
                                      float* d_txt_partial, float* d_scale, void* fwd_workspace,
                                      size_t fwd_workspace_bytes,
                                      void* bwd_workspace, size_t bwd_workspace_bytes, void* stream,
                                      int reps, float* stage_ms) {
  LATTE_CHECK_ARG(stage_ms && reps > 0);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int k = 0; k < LATTE_NUM_STAGES; ++k) stage_ms[k] = 0.f;
  for (int r = 0; r < reps; ++r) {
    StageTimer tm;
    int rc;
    float* nll_r = nullptr;
    float* nll_c = nullptr;
    float* stats = nullptr;
    if (d_txt_partial) {
      // one-sweep multi-rank flow: step 1 of the forward (the merge step after the all-gather
      // is a few microseconds and is not timed here); its outputs land in the legacy partial
      // arrays of the forward workspace, which this flow does not use
      const WsLayout wl = ws_layout(n_loc, n_all, dim, dtype);
      if ((size_t)3 * n_loc + (size_t)2 * n_all > 2 * wl.part || fwd_workspace_bytes < wl.total)
        return LATTE_ERR_WORKSPACE;
      float* tmp = static_cast<float*>(fwd_workspace) + wl.off_pmax_r;
      rc = clip_fwd_rows_impl(img_loc, ld_img_loc, txt_all, ld_txt_all, dtype, n_loc, n_all, dim,
                              label_offset, logit_scale, tmp, tmp + n_loc, tmp + 2 * n_loc,
                              tmp + 3 * n_loc, fwd_workspace, fwd_workspace_bytes, stream, &tm);
    } else {
      // like the product path: the forward hands its per-sample loss terms and LSE statistics to
      // the backward (one rank: they describe every row and column)
      const WsLayout wl = ws_layout(n_loc, n_all, dim, dtype);
      if (fwd_workspace_bytes < wl.total) return LATTE_ERR_WORKSPACE;
      float* fws = static_cast<float*>(fwd_workspace);
      if (n_loc == n_all && wl.pair_fwd) {
        nll_r = fws + wl.off_nll_r; nll_c = fws + wl.off_nll_c; stats = fws + wl.off_flag + 8;
      }
      rc = clip_fwd_impl(img_loc, ld_img_loc, txt_loc, ld_txt_loc, img_all, ld_img_all, txt_all,
                         ld_txt_all, dtype, n_loc, n_all, dim, label_offset, logit_scale, row_lse,
                         col_lse, nll_r, nll_c, loss, stats, fwd_workspace, fwd_workspace_bytes, stream,
                         &tm);
    }
    if (rc == LATTE_OK)
      rc = clip_bwd_impl(img_loc, ld_img_loc, txt_loc, ld_txt_loc, img_all, ld_img_all, txt_all,
                         ld_txt_all, dtype, n_loc, n_all, dim, label_offset, logit_scale,
                         row_lse_all, col_lse_all, nll_r, nll_c, stats, grad_loss, grad_mult,
                         cross_terms, d_img, d_txt, grad_dtype, ld_grad, d_txt_partial, nullptr, 3,
                         d_scale, bwd_workspace, bwd_workspace_bytes, stream, &tm);
    const cudaError_t e = cudaStreamSynchronize(st);
    tm.collect(stage_ms);
    if (rc) return rc;
    if (e != cudaSuccess || tm.failed) return LATTE_ERR_CUDA;
  }
  for (int k = 0; k < LATTE_NUM_STAGES; ++k) stage_ms[k] /= (float)reps;
  return LATTE_OK;
}

extern "C" int latte_comm_push(const latte_comm_t* comm, const void* txt_shard, const void* img_shard,
                               int64_t shard_bytes, int64_t tensor_stride_bytes, void* stream) {
  LATTE_CHECK_ARG(comm_ok(comm) && txt_shard && shard_bytes > 0 && (shard_bytes % 16) == 0);
  LATTE_CHECK_ARG((tensor_stride_bytes % 16) == 0 && tensor_stride_bytes >= shard_bytes * comm->world);
  LATTE_CHECK_ARG((reinterpret_cast<uintptr_t>(txt_shard) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(img_shard) & 15) == 0);
  PeerPtrs dst;
  for (int w = 0; w < LATTE_COMM_MAX_RANKS; ++w) {
    dst.p[w] = w < comm->world ? comm->gather[w] : nullptr;
    if (w < comm->world)
      LATTE_CHECK_ARG(dst.p[w] && (reinterpret_cast<uintptr_t>(dst.p[w]) & 15) == 0);
  }
  const PeerFlags flags = comm_flags(comm);
  int* my_flags = comm->flags[comm->rank];
  const int64_t shard_vecs = shard_bytes / 16;
  int64_t blocks = (shard_vecs + 255) / 256;
  if (blocks > 4 * device_sm_count()) blocks = 4 * device_sm_count();
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // the text matrix first: it is what the forward sweep waits for
  comm_push_kernel<<<(unsigned)blocks, 256, 0, st>>>(
      static_cast<const uint4*>(txt_shard), shard_vecs,
      tensor_stride_bytes / 16 + (int64_t)comm->rank * shard_vecs, dst, flags, my_flags, comm->world,
      comm->rank, comm->gen, kFlagLandedTxt, kCntPush);
  LATTE_LAUNCH_OK();
  if (img_shard) {
    comm_push_kernel<<<(unsigned)blocks, 256, 0, st>>>(
        static_cast<const uint4*>(img_shard), shard_vecs, (int64_t)comm->rank * shard_vecs, dst, flags,
        my_flags, comm->world, comm->rank, comm->gen, kFlagLandedImg, kCntPushImg);
    LATTE_LAUNCH_OK();
  }
  return LATTE_OK;
}

extern "C" int latte_comm_release(const latte_comm_t* comm, void* stream) {
  LATTE_CHECK_ARG(comm_ok(comm));
  comm_release_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(comm_flags(comm), comm->world,
                                                                      comm->rank, comm->gen);
  LATTE_LAUNCH_OK();
  return LATTE_OK;
}

extern "C" int latte_prep_features(const void* x, int64_t ld, int in_dtype, int64_t rows, int64_t dim,
                                   int normalize, int round_dtype, void* out_fp16, int64_t ld_out,
                                   float* inv_norm, void* stream) {
  LATTE_CHECK_ARG(x && out_fp16 && rows >= 0 && dim > 0 && ld >= dim && ld_out >= dim);
  LATTE_CHECK_ARG(in_dtype >= LATTE_F32 && in_dtype <= LATTE_F16);
  LATTE_CHECK_ARG(round_dtype == LATTE_BF16 || round_dtype == LATTE_F16);
  if (dim > 768 || (dim % 8) != 0 || (ld % 8) != 0 || (ld_out % 8) != 0 ||
      (reinterpret_cast<uintptr_t>(x) & 15) != 0 || (reinterpret_cast<uintptr_t>(out_fp16) & 15) != 0)
    return LATTE_ERR_UNSUPPORTED;
  if (rows == 0) return LATTE_OK;
  prep_features_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, ld, in_dtype, rows, dim, normalize, round_dtype, static_cast<__half*>(out_fp16), ld_out, inv_norm);
  LATTE_LAUNCH_OK();
  return LATTE_OK;
}

extern "C" int latte_normalize_bwd(const void* g, int64_t ld_g, int g_dtype, const void* x, int64_t ld_x,
                                   int x_dtype, const float* inv_norm, int64_t rows, int64_t dim,
                                   void* d_x, int64_t ld_dx, void* stream) {
  LATTE_CHECK_ARG(g && x && inv_norm && d_x && rows >= 0 && dim > 0);
  LATTE_CHECK_ARG(g_dtype >= LATTE_F32 && g_dtype <= LATTE_F16 && x_dtype >= LATTE_F32 && x_dtype <= LATTE_F16);
  LATTE_CHECK_ARG(ld_g >= dim && ld_x >= dim && ld_dx >= dim);
  if (rows == 0) return LATTE_OK;
  normalize_bwd_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      g, ld_g, g_dtype, x, ld_x, x_dtype, inv_norm, rows, dim, d_x, ld_dx);
  LATTE_LAUNCH_OK();
  return LATTE_OK;
}

// =========================================================================== DistillClipLoss
// /root/reference/src/open_clip/loss.py:324-362.  With S = s I T^T (student), S' the teacher's logits,
// P / P' the row softmaxes and Q / Q' the column softmaxes:
//   loss = 1/(2N) [ sum_i lse_j S_ij + sum_j lse_i S_ij - sum_ij (P'_ij + Q'_ij) S_ij ]
//   dL/dS_ij = 1/(2N) [ (P_ij + Q_ij) - (P'_ij + Q'_ij) ]
// and sum_ij W_ij S_ij = s <I, W T>, so both the loss and its gradient reduce to the products
// W.T and W^T.I with W = P + Q of one model multiplied into the STUDENT features: the gradient sweep
// (without its label term) and the stream-K gradient GEMM of ClipLoss, fed with two different
// operand pairs.  The [N, N] logits of neither model are ever stored.
namespace latte {
namespace {

constexpr int kDistillBlocks = 1024;

// partial[b] = sum over this CTA's share of (a) the 2N LSE values, (b) img .* prod; the last CTA
// adds the partials in index order (deterministic) and writes
//   loss = (sum_lse - s * dot) / (2N)  and  dot_out = dot
__global__ void __launch_bounds__(256)
distill_loss_kernel(const float* row_lse, const float* col_lse, int64_t n, const __half* img, int64_t ld_img,
                    const float* prod, int64_t ld_prod, int64_t dim, const float* logit_scale,
                    double* partial, unsigned int* counter, float* loss, float* dot_out) {
  __shared__ double red_a[8], red_b[8];
  __shared__ bool last;
  double lse = 0.0, dot = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t i = t0; i < n; i += stride) lse += (double)row_lse[i] + (double)col_lse[i];
  const int64_t per_row = dim / 4;
  for (int64_t v = t0; v < n * per_row; v += stride) {
    const int64_t r = v / per_row, c = (v % per_row) * 4;
    const float4 a = *reinterpret_cast<const float4*>(prod + r * ld_prod + c);
    const uint2 raw = *reinterpret_cast<const uint2*>(img + r * ld_img + c);
    const float2 x0 = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
    const float2 x1 = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
    dot += (double)(a.x * x0.x + a.y * x0.y) + (double)(a.z * x1.x + a.w * x1.y);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lse += __shfl_xor_sync(0xffffffffu, lse, o);
    dot += __shfl_xor_sync(0xffffffffu, dot, o);
  }
  if ((threadIdx.x & 31) == 0) { red_a[threadIdx.x >> 5] = lse; red_b[threadIdx.x >> 5] = dot; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < 8; ++w) { a += red_a[w]; b += red_b[w]; }
    partial[2 * blockIdx.x] = a;
    partial[2 * blockIdx.x + 1] = b;
    __threadfence();
    last = atomicInc(counter, gridDim.x - 1) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last || threadIdx.x != 0) return;
  __threadfence();
  double a = 0.0, b = 0.0;
  for (unsigned k = 0; k < gridDim.x; ++k) {
    a += *reinterpret_cast<volatile double*>(partial + 2 * k);
    b += *reinterpret_cast<volatile double*>(partial + 2 * k + 1);
  }
  if (loss) *loss = (float)((a - (double)__ldg(logit_scale) * b) / (2.0 * (double)n));
  if (dot_out) *dot_out = (float)b;
}

// d_img = k (A_s - A_t), d_txt = k (B_s - B_t) with k = grad * s / (2N);
// d_scale = grad / (2N) * (<img, A_s> - dot_t)
__global__ void __launch_bounds__(256)
distill_combine_kernel(const float* a_s, const float* a_t, const float* b_s, const float* b_t, int64_t ld_prod,
                       const __half* img, int64_t ld_img, int64_t n, int64_t dim, const float* logit_scale,
                       const float* grad_loss, const float* dot_t, void* d_img, void* d_txt, int grad_dtype,
                       int64_t ld_grad, double* partial, unsigned int* counter, float* d_scale) {
  __shared__ double red[8];
  __shared__ bool last;
  const float k = __ldg(grad_loss) * __ldg(logit_scale) / (2.0f * (float)n);
  const int64_t per_row = dim / 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double dot = 0.0;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n * per_row; v += stride) {
    const int64_t r = v / per_row, c = (v % per_row) * 4;
    const float4 as = *reinterpret_cast<const float4*>(a_s + r * ld_prod + c);
    const float4 at = *reinterpret_cast<const float4*>(a_t + r * ld_prod + c);
    const float4 bs = *reinterpret_cast<const float4*>(b_s + r * ld_prod + c);
    const float4 bt = *reinterpret_cast<const float4*>(b_t + r * ld_prod + c);
    const uint2 raw = *reinterpret_cast<const uint2*>(img + r * ld_img + c);
    const float2 x0 = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
    const float2 x1 = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
    dot += (double)(as.x * x0.x + as.y * x0.y) + (double)(as.z * x1.x + as.w * x1.y);
    const float gi[4] = {k * (as.x - at.x), k * (as.y - at.y), k * (as.z - at.z), k * (as.w - at.w)};
    const float gt[4] = {k * (bs.x - bt.x), k * (bs.y - bt.y), k * (bs.z - bt.z), k * (bs.w - bt.w)};
    if (grad_dtype == LATTE_F32) {
      *reinterpret_cast<float4*>(static_cast<float*>(d_img) + r * ld_grad + c) = make_float4(gi[0], gi[1], gi[2], gi[3]);
      *reinterpret_cast<float4*>(static_cast<float*>(d_txt) + r * ld_grad + c) = make_float4(gt[0], gt[1], gt[2], gt[3]);
    } else if (grad_dtype == LATTE_BF16) {
      __nv_bfloat162 i0 = __floats2bfloat162_rn(gi[0], gi[1]), i1 = __floats2bfloat162_rn(gi[2], gi[3]);
      __nv_bfloat162 t0 = __floats2bfloat162_rn(gt[0], gt[1]), t1 = __floats2bfloat162_rn(gt[2], gt[3]);
      *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(d_img) + r * ld_grad + c) =
          make_uint2(*reinterpret_cast<uint32_t*>(&i0), *reinterpret_cast<uint32_t*>(&i1));
      *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(d_txt) + r * ld_grad + c) =
          make_uint2(*reinterpret_cast<uint32_t*>(&t0), *reinterpret_cast<uint32_t*>(&t1));
    } else {
      __half2 i0 = __floats2half2_rn(gi[0], gi[1]), i1 = __floats2half2_rn(gi[2], gi[3]);
      __half2 t0 = __floats2half2_rn(gt[0], gt[1]), t1 = __floats2half2_rn(gt[2], gt[3]);
      *reinterpret_cast<uint2*>(static_cast<__half*>(d_img) + r * ld_grad + c) =
          make_uint2(*reinterpret_cast<uint32_t*>(&i0), *reinterpret_cast<uint32_t*>(&i1));
      *reinterpret_cast<uint2*>(static_cast<__half*>(d_txt) + r * ld_grad + c) =
          make_uint2(*reinterpret_cast<uint32_t*>(&t0), *reinterpret_cast<uint32_t*>(&t1));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
  __syncthreads();
  if (threadIdx.x == 0) {
    double b = 0.0;
    for (int w = 0; w < 8; ++w) b += red[w];
    partial[blockIdx.x] = b;
    __threadfence();
    last = atomicInc(counter, gridDim.x - 1) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last || threadIdx.x != 0) return;
  __threadfence();
  double b = 0.0;
  for (unsigned q = 0; q < gridDim.x; ++q) b += *reinterpret_cast<volatile double*>(partial + q);
  *d_scale = (float)((double)__ldg(grad_loss) / (2.0 * (double)n) * (b - (double)__ldg(dot_t)));
}

}  // namespace
}  // namespace latte

extern "C" int latte_distill_aux_bytes(size_t* bytes) {
  LATTE_CHECK_ARG(bytes);
  *bytes = (size_t)kDistillBlocks * 2 * sizeof(double) + 256;
  return LATTE_OK;
}

extern "C" int latte_distill_products(const void* sweep_img, int64_t ld_sweep_img, const void* sweep_txt,
                                      int64_t ld_sweep_txt, const void* gemm_img, int64_t ld_gemm_img,
                                      const void* gemm_txt, int64_t ld_gemm_txt, int64_t n, int64_t dim,
                                      const float* logit_scale, const float* row_lse, const float* col_lse,
                                      float* out_img, float* out_txt, int64_t ld_out, void* workspace,
                                      size_t workspace_bytes, void* stream) {
  LATTE_CHECK_ARG(sweep_img && sweep_txt && gemm_img && gemm_txt && logit_scale && row_lse && col_lse &&
                  out_img && out_txt && workspace);
  LATTE_CHECK_ARG(n > 0 && dim > 0 && ld_sweep_img >= dim && ld_sweep_txt >= dim && ld_gemm_img >= dim &&
                  ld_gemm_txt >= dim && ld_out >= dim);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int dtype = LATTE_F16;
  const WsLayout w = ws_layout(n, n, dim, dtype, true);
  if (workspace_bytes < w.total) return LATTE_ERR_WORKSPACE;
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return LATTE_ERR_BAD_ARG;
  if (!w.pair || !clip_pair_supported(dtype, dim, ld_sweep_img, ld_sweep_txt, sweep_img, sweep_txt) ||
      !clip_pair_supported(dtype, dim, ld_gemm_img, ld_gemm_txt, gemm_img, gemm_txt) || (ld_out % 4) != 0 ||
      (reinterpret_cast<uintptr_t>(out_img) & 15) != 0 || (reinterpret_cast<uintptr_t>(out_txt) & 15) != 0)
    return LATTE_ERR_UNSUPPORTED;
  float* ws = static_cast<float*>(workspace);
  float* row2 = ws + w.off_row2;
  float* col2 = ws + w.off_col2;
  float* rho = ws + w.off_rho;
  int* fast_flag = reinterpret_cast<int*>(ws + w.off_rho + 1);
  float* gscale = ws + w.off_rho + 3;
  float* out_scale = ws + w.off_rho + 4;
  float* st_buf = ws + w.off_rho + 8;
  // LSE range for the one-ex2 epilogue; without nll vectors the |W| bound is 2 (fp16 scale 2^13)
  lse_stats_kernel<<<1, 1024, 0, st>>>(row_lse, col_lse, n, nullptr, nullptr, nullptr, st_buf);
  LATTE_LAUNCH_OK();
  lse_vectors_kernel<<<(unsigned)((w.n_pad + 255) / 256), 256, 0, st>>>(
      row_lse, col_lse, n, (int64_t)w.n_pad, st_buf, nullptr, 1.0f, logit_scale, n, rho, row2, col2,
      ws + w.off_erow, ws + w.off_einvrow, ws + w.off_ecol, ws + w.off_einvcol);
  LATTE_LAUNCH_OK();
  const int dsn = clip_pair_ds_count();
  if ((size_t)(2 * dsn) > w.ds_cap) return LATTE_ERR_WORKSPACE;
  float* dsp = ws + w.off_ds;
  __half* gbuf = reinterpret_cast<__half*>(ws + w.off_g);
  float* acc_i = ws + w.off_acc0;
  float* acc_t = ws + w.off_acc1;
  const size_t acc_bytes = (size_t)n * w.ld32 * sizeof(float);
  LATTE_CUDA_OK(cudaMemsetAsync(dsp, 0, (size_t)2 * dsn * sizeof(float), st));
  PairSweepArgs sa;
  sa.dtype = dtype; sa.n_loc = n; sa.n_all = n; sa.dim = dim;
  sa.label_offset = 0; sa.logit_scale = logit_scale; sa.cross_terms = 1; sa.no_label = 1;
  sa.g = gbuf; sa.ds_both = 1; sa.nll_a = nullptr; sa.nll_b = nullptr;
  sa.x = sweep_img; sa.ldx = ld_sweep_img; sa.y = sweep_txt; sa.ldy = ld_sweep_txt;
  sa.lse_a2 = row2; sa.lse_b2 = col2; sa.ds_partial = dsp;
  sa.e_a = ws + w.off_erow; sa.einv_b = ws + w.off_einvcol; sa.fast_flag = fast_flag;
  sa.gscale_log2 = gscale;
  PairGemmArgs ga = {};
  ga.g = gbuf; ga.n_loc = n; ga.n_all = n; ga.dim = dim; ga.ld32 = (int64_t)w.ld32;
  ga.feat_dtype = LATTE_F16;
  ga.out_dtype = LATTE_F32; ga.ld_out = ld_out; ga.out_scale = out_scale;
  ga.y16 = gemm_txt; ga.ldy16 = ld_gemm_txt;
  ga.x16 = gemm_img; ga.ldx16 = ld_gemm_img;
  ga.dx32 = acc_i; ga.dy32 = acc_t; ga.ld_dy32 = (int64_t)w.ld32; ga.dy_scale = nullptr;
  ga.dy_peers = nullptr; ga.n_peers = 0;
  ga.dx_out = out_img; ga.dy_out = out_txt;
  const bool direct_i = clip_pair_gemm_direct(ga, 0);
  const bool direct_t = clip_pair_gemm_direct(ga, 1);
  if (!direct_i) LATTE_CUDA_OK(cudaMemsetAsync(acc_i, 0, acc_bytes, st));
  if (!direct_t) LATTE_CUDA_OK(cudaMemsetAsync(acc_t, 0, acc_bytes, st));
  int rc = clip_pair_gemm_fixup(ga, 0, st);
  if (rc) return rc;
  rc = clip_pair_sweep(sa, st);
  if (rc) return rc;
  rc = clip_pair_gemm(ga, st);
  if (rc) return rc;
  rc = clip_pair_gemm_fixup(ga, 1, st);
  if (rc) return rc;
  if (!direct_i) {
    rc = clip_pair_scale_cast(acc_i, nullptr, (int64_t)w.ld32, out_img, nullptr, LATTE_F32, ld_out, n, dim,
                              out_scale, st);
    if (rc) return rc;
  }
  if (!direct_t) {
    rc = clip_pair_scale_cast(acc_t, nullptr, (int64_t)w.ld32, out_txt, nullptr, LATTE_F32, ld_out, n, dim,
                              out_scale, st);
    if (rc) return rc;
  }
  return LATTE_OK;
}

extern "C" int latte_distill_loss(const float* row_lse, const float* col_lse, int64_t n, const void* img,
                                  int64_t ld_img, const float* teacher_prod, int64_t ld_prod, int64_t dim,
                                  const float* logit_scale, float* loss, float* dot_out, void* aux,
                                  size_t aux_bytes, void* stream) {
  LATTE_CHECK_ARG(row_lse && col_lse && img && teacher_prod && logit_scale && loss && dot_out && aux);
  LATTE_CHECK_ARG(n > 0 && dim > 0 && (dim % 4) == 0 && (ld_img % 4) == 0 && (ld_prod % 4) == 0);
  size_t need;
  latte_distill_aux_bytes(&need);
  if (aux_bytes < need) return LATTE_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uintptr_t base = (reinterpret_cast<uintptr_t>(aux) + 15) / 16 * 16;
  unsigned int* counter = reinterpret_cast<unsigned int*>(base);
  double* partial = reinterpret_cast<double*>(base + 16);
  LATTE_CUDA_OK(cudaMemsetAsync(counter, 0, 16, st));
  int64_t blocks = (n * (dim / 4) + 255) / 256;
  if (blocks > kDistillBlocks - 2) blocks = kDistillBlocks - 2;
  distill_loss_kernel<<<(unsigned)blocks, 256, 0, st>>>(row_lse, col_lse, n, static_cast<const __half*>(img),
                                                        ld_img, teacher_prod, ld_prod, dim, logit_scale,
                                                        partial, counter, loss, dot_out);
  LATTE_LAUNCH_OK();
  return LATTE_OK;
}

extern "C" int latte_distill_bwd_combine(const float* a_s, const float* a_t, const float* b_s, const float* b_t,
                                         int64_t ld_prod, const void* img, int64_t ld_img, int64_t n,
                                         int64_t dim, const float* logit_scale, const float* grad_loss,
                                         const float* dot_t, void* d_img, void* d_txt, int grad_dtype,
                                         int64_t ld_grad, float* d_scale, void* aux, size_t aux_bytes,
                                         void* stream) {
  LATTE_CHECK_ARG(a_s && a_t && b_s && b_t && img && logit_scale && grad_loss && dot_t && d_img && d_txt &&
                  d_scale && aux);
  LATTE_CHECK_ARG(n > 0 && dim > 0 && (dim % 4) == 0 && (ld_img % 4) == 0 && (ld_prod % 4) == 0 &&
                  (ld_grad % 4) == 0);
  LATTE_CHECK_ARG(grad_dtype >= LATTE_F32 && grad_dtype <= LATTE_F16);
  size_t need;
  latte_distill_aux_bytes(&need);
  if (aux_bytes < need) return LATTE_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uintptr_t base = (reinterpret_cast<uintptr_t>(aux) + 15) / 16 * 16;
  unsigned int* counter = reinterpret_cast<unsigned int*>(base);
  double* partial = reinterpret_cast<double*>(base + 16);
  LATTE_CUDA_OK(cudaMemsetAsync(counter, 0, 16, st));
  int64_t blocks = (n * (dim / 4) + 255) / 256;
  if (blocks > kDistillBlocks - 2) blocks = kDistillBlocks - 2;
  distill_combine_kernel<<<(unsigned)blocks, 256, 0, st>>>(
      a_s, a_t, b_s, b_t, ld_prod, static_cast<const __half*>(img), ld_img, n, dim, logit_scale, grad_loss,
      dot_t, d_img, d_txt, grad_dtype, ld_grad, partial, counter, d_scale);
  LATTE_LAUNCH_OK();
  return LATTE_OK;
}

// Shared declarations for the liblatte_b200 translation units.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/latte_b200.h"

namespace latte {

#define LATTE_CHECK_ARG(cond) \
  do {                        \
    if (!(cond)) return LATTE_ERR_BAD_ARG; \
  } while (0)

#define LATTE_CUDA_OK(expr)                          \
  do {                                               \
    cudaError_t _e = (expr);                         \
    if (_e != cudaSuccess) return LATTE_ERR_CUDA;    \
  } while (0)

// Checked after every launch: cudaGetLastError does not synchronise.
#define LATTE_LAUNCH_OK()                            \
  do {                                               \
    cudaError_t _e = cudaGetLastError();             \
    if (_e != cudaSuccess) return LATTE_ERR_CUDA;    \
  } while (0)

inline size_t dtype_size(int dtype) { return dtype == LATTE_F32 ? 4 : 2; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// --------------------------------------------------------------------------------------
// Row kernels of ClipLoss.  "x" is the rank-local row operand [n_loc, dim], "y" the
// gathered column operand [n_all, dim]; S = x @ y^T is never stored.
// --------------------------------------------------------------------------------------
struct ClipFwdArgs {
  const void* x; int64_t ldx;
  const void* y; int64_t ldy;
  int dtype;
  int64_t n_loc, n_all, dim;
  int64_t label_offset;
  const float* logit_scale;   // device scalar
  // per-(split,row) partial online-softmax state in base-2 units of (s*log2e)*dot
  float* part_max;            // [nparts, n_loc]
  float* part_sum;            // [nparts, n_loc]
  float* diag;                // [n_loc]  raw dot of the label pair (unscaled)
  int nparts;                 // partial slots per row (tc: 2 * column splits, simt: 1)
  const int* gate;            // tc only, nullable: the kernel returns at once when *gate == 0
};

struct ClipBwdArgs {
  const void* x; int64_t ldx;
  const void* y; int64_t ldy;
  const void* y16; int64_t ldy16;   // fp16 copy of y (tc path; == y for fp16 features)
  int dtype;
  int64_t n_loc, n_all, dim;
  int64_t label_offset;
  const float* logit_scale;   // device scalar
  const float* lse_a2;        // [n_all] base-2 LSE of the x side (indexed label_offset + i)
  const float* lse_b2;        // [n_all] base-2 LSE of the y side (indexed j)
  const float* grad_loss;     // device scalar
  float grad_mult;
  int cross_terms;
  void* dx; int grad_dtype; int64_t ld_dx;
  float* ds_partial;          // [ds_count] per-CTA partial sums of d loss / d s
  int ds_count;
};

// tcgen05 path (bf16 / fp16 features).  Return LATTE_ERR_UNSUPPORTED when the shape
// cannot run on it (the caller then reports the error; there is no silent fallback).
int clip_fwd_rows_tc(const ClipFwdArgs& a, cudaStream_t stream);
int clip_bwd_rows_tc(const ClipBwdArgs& a, cudaStream_t stream);
int clip_tc_nparts(int64_t n_loc, int64_t n_all, int sm_count);
int clip_tc_ds_count(int64_t n_loc, int64_t dim);
bool clip_tc_supported(int dtype, int64_t dim, int64_t ldx, int64_t ldy, const void* x, const void* y);

// CTA-pair backward (clip_pair.cu): gradient-weight sweep + stream-K gradient GEMMs, dim <= 512.
struct PairGeom {
  int row_blocks;     // blocks of 256 rows of the x side
  int col_tiles;      // tiles of 128 columns (even)
  int ncb;            // 64-column G blocks per block row
  size_t g_elems;     // fp16 elements of the blocked G scratch
};
struct PairSweepArgs {
  const void* x; int64_t ldx;
  const void* y; int64_t ldy;
  int dtype;
  int64_t n_loc, n_all, dim;
  int64_t label_offset;
  const float* logit_scale;
  const float* lse_a2;
  const float* lse_b2;
  const float* e_a;           // 2^(lse_a2 - rho)
  const float* einv_b;        // 2^(rho - lse_b2)
  const int* fast_flag;       // device flag: LSE range small enough for the one-ex2 epilogue
  const float* gscale_log2;   // device scalar: log2 of the fp16 scale of G
  const float* nll_a;         // nullable: per-sample loss terms lse - label logit of the x side
  const float* nll_b;         //           and the y side (accurate 1 - P_label near convergence)
  int cross_terms;
  int ds_both;                // 1: d loss / d s takes both softmax terms from this sweep
  int no_label;               // 1: G = P_row + cb * P_col without the label term (soft-target losses)
  void* g;                    // blocked fp16 G scratch (PairGeom::g_elems)
  float* ds_partial;          // [clip_pair_ds_count()] zeroed by the caller
};
struct PairSigArgs {
  const void* x; int64_t ldx;       // [n_loc, dim] image rows of this rank
  const void* y; int64_t ldy;       // [n_all, dim] all text rows
  int dtype;
  int64_t n_loc, n_all, dim;
  int64_t label_offset;
  const float* logit_scale;
  const float* logit_bias;          // nullable
  void* g;                          // NULL: forward (loss partials); else blocked fp16 G scratch
  float* ds_partial;                // backward: [clip_pair_ds_count()] zeroed by the caller
  float* aux_partial;               // forward: loss partials; backward: d bias partials (same size)
};
int clip_pair_sig_sweep(const PairSigArgs& a, cudaStream_t stream);
struct PairGemmArgs {
  const void* g;
  const void* y16; int64_t ldy16;   // fp16 [n_all, dim]
  const void* x16; int64_t ldx16;   // fp16 [n_loc, dim] or NULL (no transposed product)
  int64_t n_loc, n_all, dim;
  float* dx32;                      // [n_loc, ld32] zeroed
  float* dy32;                      // [n_all, ld_dy32] zeroed (with x16)
  int64_t ld32;
  int64_t ld_dy32;
  const float* dy_scale;            // nullable device scalar multiplied into dy32's partials
  float* const* dy_peers;           // nullable HOST array: per-rank accumulators [n_all / n_peers, ld_dy32]
  int n_peers;                      //   (peer-mapped); rows of dy go to their owner rank instead of dy32
  int feat_dtype;                   // element type of y16 / x16: LATTE_F16 (G is fp16; kind::f16 needs one format)
  // Direct outputs (nullable): tiles owned by one cluster are written as out_scale * acc in out_dtype
  // straight from the GEMM epilogue; dx32 / dy32 then only serve the tiles split between clusters
  // (clip_pair_gemm_fixup before and after the GEMM).  dy_out is ignored with dy_scale / dy_peers.
  void* dx_out;
  void* dy_out;
  int out_dtype;
  int64_t ld_out;
  const float* out_scale;
  // With dy_peers: once every add of the launch is out (last CTA), done_flags[w][done_slot] = done_gen
  // is published on every rank w (system-scope release); n_done = 0 switches the signal off.
  int* done_flags[8];
  int n_done;
  unsigned int* done_counter;
  int done_gen;
  int done_slot;
};
// Forward on the same sweep: rows dealt to clusters as contiguous tile ranges; a row block
// split over several clusters gets one partial slot per cluster.
struct PairFwdGeom {
  int row_blocks, col_tiles;
  int64_t total;          // row_blocks * col_tiles
  int ncl;                // clusters launched
  int slots;              // upper bound of partial slots per row (x2 groups x2 tile halves)
  int64_t ld_colpart;     // row pitch of the column-partial matrix
};
struct PairFwdArgs {
  const void* x; int64_t ldx;
  const void* y; int64_t ldy;
  int dtype;
  int64_t n_loc, n_all, dim;
  int64_t label_offset;
  const float* logit_scale;
  float* part_max;            // [4 * slots, n_loc]; which slots are written follows from the schedule
  float* part_sum;
  float* diag;                // [n_loc]
  float* col_part;            // [2 * row_blocks, ld_colpart] or NULL (rows only)
  float* col_ref;             // [2 * row_blocks, 4 * col_tiles]
  int* zero2;                 // nullable: two ints the sweep clears (fallback flag, loss counter)
  // multi-rank: y is a gathered buffer filled shard by shard by the ranks' push kernels;
  // landed[w] >= landed_gen says rank w's shard (rows [w * rows_per_rank, ...)) may be read
  const int* landed = nullptr;
  int landed_gen = 0;
  int64_t rows_per_rank = 0;
};
PairFwdGeom clip_pair_fwd_geom(int64_t n_loc, int64_t n_all);
int clip_pair_fwd_sweep(const PairFwdArgs& a, cudaStream_t stream);
// cluster that owns tile t when `total` tiles are dealt to `ncl` clusters as [c*total/ncl, ...)
__host__ __device__ inline int64_t cluster_of_tile(int64_t t, int64_t total, int64_t ncl) {
  return ((t + 1) * ncl - 1) / total;
}

bool clip_pair_supported(int dtype, int64_t dim, int64_t ldx, int64_t ldy, const void* x, const void* y);
PairGeom clip_pair_geom(int64_t n_loc, int64_t n_all);
int clip_pair_ds_count();
int clip_pair_sweep(const PairSweepArgs& a, cudaStream_t stream);
int clip_pair_gemm(const PairGemmArgs& a, cudaStream_t stream);
int clip_pair_gemm_fixup(const PairGemmArgs& a, int cast, cudaStream_t stream);
bool clip_pair_gemm_direct(const PairGemmArgs& a, int product);
int clip_pair_scale_cast(const float* acc0, const float* acc1, int64_t ld_acc, void* out0, void* out1,
                         int out_dtype, int64_t ld_out, int64_t rows, int64_t dim,
                         const float* out_scale, cudaStream_t stream);

// fp32 SIMT path (fp32 features; exact fp32 products).
int clip_fwd_rows_simt(const ClipFwdArgs& a, cudaStream_t stream);
int clip_bwd_rows_simt(const ClipBwdArgs& a, cudaStream_t stream);
int clip_simt_ds_count(int64_t n_loc);

int device_sm_count();

}  // namespace latte

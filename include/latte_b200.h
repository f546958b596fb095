/*
 * latte_b200.h -- C ABI of the B200-native LatteCLIP loss head (liblatte_b200.so).
 *
 * The reference (astra-vision/LatteCLIP) is pure Python/PyTorch and has no FFI for this
 * path; every entry point below replaces a span of reference Python that today runs as
 * a chain of torch library calls.  The citation after each prototype is the reference
 * code it stands in for (paths relative to the reference repository root).
 *
 * Conventions
 *   - All pointers are DEVICE pointers owned by the caller (PyTorch caching allocator);
 *     the library never frees or retains them past the call.
 *   - Sizes are int64_t element counts, ld* are row strides in ELEMENTS, rows are
 *     row-major with unit stride along the feature axis.
 *   - `stream` is a cudaStream_t passed as void*; every call only enqueues work on it
 *     (no host synchronisation, no thread-local or global mutable state => re-entrant,
 *     callable from autograd worker threads).
 *   - Scalars that live on the device in the reference (logit_scale = model.logit_scale.exp(),
 *     src/training/train.py:405; the upstream gradient of the loss) are passed as device
 *     pointers so the host never has to .item() them.
 *   - Return value: 0 on success, a negative latte_status_t otherwise.  There is no CPU or
 *     PyTorch fallback behind any entry point.
 */
#ifndef LATTE_B200_H_
#define LATTE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  LATTE_OK = 0,
  LATTE_ERR_BAD_ARG = -1,        /* null pointer, negative size, bad enum                */
  LATTE_ERR_UNSUPPORTED = -2,    /* shape / dtype / alignment this build cannot run       */
  LATTE_ERR_WORKSPACE = -3,      /* caller workspace too small                            */
  LATTE_ERR_CUDA = -4,           /* a CUDA runtime / driver call or kernel launch failed  */
  LATTE_ERR_NO_DEVICE = -5       /* no sm_100 device visible                              */
} latte_status_t;

typedef enum {
  LATTE_F32 = 0,
  LATTE_BF16 = 1,
  LATTE_F16 = 2
} latte_dtype_t;

/* label_weight_axis of the text mixture, src/training/train.py:476,481 */
typedef enum {
  LATTE_LABEL_AXIS_ROW = 0,      /* w_lbl[i] scales row i    (the evident intent)                 */
  LATTE_LABEL_AXIS_QUIRK = 1     /* w_lbl[d] scales column d (the reference's literal broadcast,  */
                                 /* only defined when B == D)                                     */
} latte_label_axis_t;

/* ---- library ----------------------------------------------------------------------- */
const char* latte_version(void);
const char* latte_status_string(int status);
/* Fills SM count and compute capability of the current device. */
int latte_device_info(int* sm_count, int* cc_major, int* cc_minor);

/*
 * ---- multi-rank exchange over NVLink peer memory (one process per GPU) ----------------------
 * The reference's collectives on this path are gather_features' two all-gathers (loss.py:49-50 /
 * :54-55) and, in backward, the reduce-scatter of torch.distributed.nn.all_gather.  Here every
 * rank owns one symmetric allocation per slot that all ranks map (torch symmetric memory); the
 * kernels of this library store into / add into the peers' buffers directly and synchronise with
 * generation-numbered flags (int32, system-scope release / acquire) -- no host-side barrier, no
 * collective call.  A slot carries generation `gen` (1, 2, ... identical on all ranks):
 *   gather[w]   rank w's gathered-feature buffer [2][n_all, dim] (image matrix, then text matrix)
 *   payload[w]  rank w's forward payload block: [world][payload_stride] floats (first exchange)
 *               followed by [world][2 n_all] floats (second, exact round)
 *   acc[w]      rank w's fp32 text-gradient accumulator [n_loc, dim]; zero when a generation starts
 *   flags[w]    rank w's flag block, LATTE_COMM_FLAG_INTS int32 (zero-initialised once):
 *               landed_txt[8] | landed_img[8] | payload[8] | payload2[8] | done[8] | free[8] |
 *               counters[8] (local); entry [src] = latest generation for which rank src has
 *               finished that step towards THIS rank.
 * Ordering rules the caller keeps: all ranks issue the same sequence of calls; a slot's generation
 * g may start once generation g - 1 of the same slot was released by every rank (the push waits
 * for free[*] >= g - 1 by itself).
 */
#define LATTE_COMM_MAX_RANKS 8
#define LATTE_COMM_FLAG_INTS 64
typedef struct latte_comm {
  int32_t rank, world, gen, reserved;
  void* gather[LATTE_COMM_MAX_RANKS];
  float* payload[LATTE_COMM_MAX_RANKS];
  float* acc[LATTE_COMM_MAX_RANKS];
  int32_t* flags[LATTE_COMM_MAX_RANKS];
  int64_t payload_stride;          /* floats per rank row of the first payload exchange */
} latte_comm_t;

/*
 * The feature all-gather as ONE store kernel: waits until every peer released the slot's previous
 * generation, writes this rank's text shard ([n_loc, dim], shard_bytes) -- and image shard, if not
 * NULL -- into every rank's gather buffer (the text matrix starts tensor_stride_bytes after the
 * image matrix) and publishes landed_txt / landed_img = gen on every rank.  The forward sweep
 * starts reading a shard as soon as its flag shows up (latte_clip_fwd_rank), so no rank waits for
 * the slowest pusher before it starts.  comm == NULL-flags variant for tests: pass flags[] = NULL
 * to skip waits and signals.
 */
int latte_comm_push(const latte_comm_t* comm, const void* txt_shard, const void* img_shard,
                    int64_t shard_bytes, int64_t tensor_stride_bytes, void* stream);

/* Releases the slot's generation without a backward (forward-only call): free = gen on all ranks. */
int latte_comm_release(const latte_comm_t* comm, void* stream);

/* ---- operand preparation -------------------------------------------------------------- */
/*
 * Fused cast (+ opt-in L2 normalisation) of a feature matrix into the tensor-core operand format.
 * The reference normalises in the towers (src/open_clip/model.py:415-418, 420-437 -- under amp the
 * result is fp32) and casts inside its autocast matmuls (loss.py:109-116).  One pass: optional
 * x / max(||x||, 1e-12) in fp32, rounding to `round_dtype` (LATTE_BF16 or LATTE_F16: the reference's
 * autocast dtype) and storage as fp16 [rows, dim] -- the one 16-bit format all tensor-core kernels
 * of this library take (a bf16 value is exact in fp16 down to 2^-17).  inv_norm (nullable, [rows]):
 * 1 / max(||x||, 1e-12) for latte_normalize_bwd.  dim <= 768, dim % 8 == 0, 16-byte aligned rows.
 */
int latte_prep_features(const void* x, int64_t ld, int in_dtype, int64_t rows, int64_t dim,
                        int normalize, int round_dtype, void* out_fp16, int64_t ld_out,
                        float* inv_norm, void* stream);
/* Backward of the normalisation: d_x = inv * (g - xh <xh, g>) with xh = x * inv; d_x in x_dtype. */
int latte_normalize_bwd(const void* g, int64_t ld_g, int g_dtype, const void* x, int64_t ld_x,
                        int x_dtype, const float* inv_norm, int64_t rows, int64_t dim,
                        void* d_x, int64_t ld_dx, void* stream);

/* ---- ClipLoss: open_clip/loss.py ------------------------------------------------------ */

/* Bytes of scratch latte_clip_fwd / latte_clip_bwd need for these sizes. */
int latte_clip_workspace_bytes(int64_t n_loc, int64_t n_all, int64_t dim, int dtype,
                               size_t* bytes);

/* Bytes of scratch latte_clip_bwd needs: the forward scratch plus, for 16-bit features with
 * dim <= 768, the fp16 gradient-weight blocks G [n_loc, n_all] and two fp32 [n_loc, dim]
 * accumulators of the CTA-pair backward (csrc/clip_pair.cu). */
int latte_clip_bwd_workspace_bytes(int64_t n_loc, int64_t n_all, int64_t dim, int dtype,
                                   size_t* bytes);

/*
 * Forward of ClipLoss on one rank.  Replaces ClipLoss.get_logits + get_ground_truth + the
 * two F.cross_entropy calls (src/open_clip/loss.py:102-118, 89-100, 126-129) without
 * materialising the logits:
 *   row_lse[i] = logsumexp_j( s * <img_loc[i], txt_all[j]> )        i in [0, n_loc)
 *   col_lse[i] = logsumexp_j( s * <txt_loc[i], img_all[j]> )
 *   diag[i]    = s * <img_loc[i], txt_all[label_offset + i]>
 *   *loss      = ( mean_i(row_lse[i] - diag[i]) + mean_i(col_lse[i] - diag'[i]) ) / 2
 * label_offset = rank * n_loc under local_loss (loss.py:93-94), 0 for world_size 1.
 * For world_size 1 pass img_all = img_loc, txt_all = txt_loc, n_all = n_loc.
 * bf16 / fp16 features run on tcgen05 tensor cores (fp32 accumulate in TMEM); fp32
 * features run on an fp32 SIMT kernel.
 */
int latte_clip_fwd(const void* img_loc, int64_t ld_img_loc,
                   const void* txt_loc, int64_t ld_txt_loc,
                   const void* img_all, int64_t ld_img_all,
                   const void* txt_all, int64_t ld_txt_all,
                   int dtype, int64_t n_loc, int64_t n_all, int64_t dim,
                   int64_t label_offset,
                   const float* logit_scale,     /* device scalar s                       */
                   float* row_lse, float* col_lse, /* [n_loc] each                        */
                   float* row_nll, float* col_nll, /* [n_loc] each, nullable: the per-sample
                                                    loss terms lse - label logit, formed as
                                                    (max - label logit) + log(sum) so they keep
                                                    full relative accuracy near convergence  */
                   float* loss,                  /* device scalar out                     */
                   float* stats,                 /* nullable [4] out: min / max of all LSE values
                                                    and the largest nll -- what latte_clip_bwd
                                                    needs to scale G (valid as `lse_stats` there
                                                    when n_loc == n_all)                      */
                   void* workspace, size_t workspace_bytes, void* stream);

/*
 * Multi-rank forward with ONE logit sweep per rank (16-bit features, dim <= 768; ask
 * latte_clip_rank_sweep_supported).  Step 1 on each rank: rows of this rank against all
 * gathered text features ->
 *   row_lse / row_nll / label_logit [n_loc]   (label_logit[i] = s * <img_loc[i], txt_all[off+i]>)
 *   col_ml [n_all, 2]                         base-2 (max, sum) of every column over THIS rank's rows
 * The four outputs may be views of ONE packed buffer [2 n_all + 3 n_loc] in that order
 * (col_ml first); the caller all-gathers that buffer across ranks and step 2 merges it:
 *   col_lse_all / col_nll_all [n_all], *loss = (mean_i row_nll[i] + mean_i col_nll_all[off+i]) / 2.
 * If a column's partial sums may have lost flushed terms, step 2 recomputes every column exactly
 * from the gathered features (device-side decision, no host sync).
 */
int latte_clip_rank_sweep_supported(int dtype, int64_t dim);
int latte_clip_fwd_rows(const void* img_loc, int64_t ld_img_loc,
                        const void* txt_all, int64_t ld_txt_all,
                        int dtype, int64_t n_loc, int64_t n_all, int64_t dim,
                        int64_t label_offset, const float* logit_scale,
                        float* row_lse, float* row_nll, float* label_logit, float* col_ml,
                        void* workspace, size_t workspace_bytes, void* stream);
int latte_clip_fwd_cols_workspace_bytes(int64_t n_all, int64_t dim, int dtype, size_t* bytes);
int latte_clip_fwd_cols(const float* gathered /* [world, stride]: per rank col_ml [n_all, 2] |
                                                    row_lse | row_nll | label_logit [n_loc] each */,
                        int64_t stride, int world,
                        const void* img_all, int64_t ld_img_all,
                        const void* txt_all, int64_t ld_txt_all,
                        int dtype, int64_t n_loc, int64_t n_all, int64_t dim,
                        int64_t label_offset, const float* logit_scale,
                        float* row_lse_all, float* row_nll_all,      /* [n_all] unpacked     */
                        float* col_lse_all, float* col_nll_all,      /* [n_all] merged       */
                        float* loss,
                        float* stats,                 /* nullable [4] out, as latte_clip_fwd (here
                                                         they cover all n_all rows and columns)  */
                        void* workspace, size_t workspace_bytes, void* stream);

/*
 * The same multi-rank forward with every exchange done by the kernels themselves over peer memory
 * (after latte_comm_push of this generation): the sweep reads the gathered text matrix
 * comm->gather[rank] + tensor_stride as the shards land; the per-rank payload is stored into every
 * peer's payload block and merged as soon as all blocks arrived.  If (after the merge, identically
 * on every rank) a column's partial sums may have lost flushed terms, every rank recomputes its own
 * column partials exactly (all texts x own images -- no gathered images needed) and a second
 * exchange replaces them; without that flag the three kernels of the second round return at once.
 * Outputs as latte_clip_fwd_cols.  `phases` (bit mask, 7 = everything) exists so that a test can
 * drive all ranks from ONE process on one GPU, phase by phase: 1 = sweep + finalize + payload
 * store, 2 = merge (+ the gated exact recompute and its store), 4 = final merge + loss.
 */
int latte_clip_fwd_rank_workspace_bytes(int64_t n_loc, int64_t n_all, int64_t dim, int dtype,
                                        size_t* bytes);
int latte_clip_fwd_rank(const latte_comm_t* comm,
                        const void* img_loc, int64_t ld_img_loc,
                        const void* txt_all, int64_t ld_txt_all,
                        int dtype, int64_t n_loc, int64_t n_all, int64_t dim,
                        int64_t label_offset, const float* logit_scale,
                        float* row_lse_all, float* row_nll_all,
                        float* col_lse_all, float* col_nll_all,
                        float* loss, float* stats, int phases,
                        void* workspace, size_t workspace_bytes, void* stream);

/*
 * Backward of ClipLoss on one rank (the autograd of loss.py:109-116 + 126-129, and the
 * reduce-scatter in the backward of torch.distributed.nn.all_gather, loss.py:49-50,
 * replaced by an exchange of the LSE vectors).  With S_ij = s*<img_i, txt_j>:
 *   G_ij   = ca * exp(S_ij - row_lse[i]) + cb * exp(S_ij - col_lse[j]) - cd * [j == label(i)]
 *   d_img  = coef * s * G[loc rows, :] @ txt_all          (and symmetrically d_txt)
 *   d_scale= coef * sum_ij (exp(S_ij - row_lse[i]) - delta_ij) * S_ij / s   (+ text side)
 * coef = grad_loss / (2 * n_loc) * grad_mult.   cross_terms = 1 gives ca = cb = 1, cd = 2
 * (every mode except local_loss && !gather_with_grad); cross_terms = 0 gives ca = 1,
 * cb = 0, cd = 1 (loss.py:52-55 with local_loss: gathered features carry no gradient).
 * row_lse_all / col_lse_all are the all-gathered LSE vectors [n_all] (== the local ones
 * when world_size is 1).  d_img / d_txt are [n_loc, dim] in `grad_dtype`.
 * `workspace` is sized by latte_clip_bwd_workspace_bytes.
 * row_nll_all / col_nll_all (nullable, both or neither): the forward's per-sample loss terms
 * [n_all]; with them the label entry of G is expm1(-nll) instead of exp(S - lse) - 1.
 * d_txt_partial (nullable, fp32 [n_all, dim], 16-bit features with dim <= 768 and
 * cross_terms = 1 only): one-sweep multi-rank mode -- d_txt is not written; instead
 * d_txt_partial = coef * s * G[loc rows, :]^T @ img_loc for ALL columns, which the caller
 * reduce-scatters (SUM) over the ranks; *d_scale then covers this rank's rows x all columns
 * (a different partition of the same global sum than the reference's per-rank value).
 */
int latte_clip_bwd(const void* img_loc, int64_t ld_img_loc,
                   const void* txt_loc, int64_t ld_txt_loc,
                   const void* img_all, int64_t ld_img_all,
                   const void* txt_all, int64_t ld_txt_all,
                   int dtype, int64_t n_loc, int64_t n_all, int64_t dim,
                   int64_t label_offset,
                   const float* logit_scale,
                   const float* row_lse_all, const float* col_lse_all,   /* [n_all]  */
                   const float* row_nll_all, const float* col_nll_all,   /* [n_all], nullable */
                   const float* lse_stats,       /* nullable [4]: the forward's `stats` over ALL
                                                    n_all rows and columns; NULL = recomputed
                                                    from the vectors above                     */
                   const float* grad_loss,       /* device scalar dL/dloss                */
                   float grad_mult, int cross_terms,
                   void* d_img, void* d_txt, int grad_dtype, int64_t ld_grad,
                   float* d_txt_partial,         /* nullable, see above                   */
                   const latte_comm_t* comm,     /* nullable.  Fused reduce-scatter: instead of
                                                    d_txt_partial, every row of G^T @ img_loc is
                                                    added straight into its owner's accumulator
                                                    comm->acc[w] over NVLink from the GEMM
                                                    epilogue; the kernel publishes done = gen when
                                                    its adds are out, and a last kernel waits for
                                                    every rank's done flag, casts this rank's
                                                    accumulator into d_txt, clears it and releases
                                                    the slot (free = gen).  `phases`: 1 = up to the
                                                    GEMM, 2 = that last kernel (3 = both)        */
                   int phases,
                   float* d_scale,               /* device scalar out (d loss / d s)      */
                   void* workspace, size_t workspace_bytes, void* stream);



/*
 * SigLipLoss (open_clip loss.py:453-560; factory.py:337-342 selects it on args.siglip).
 * z_ij = logit_scale * <img_i, txt_j> + logit_bias;  per rank
 *   loss = sum_{i in own rows, j in ALL columns} -logsigmoid(label_ij * z_ij) / n_loc,
 * label_ij = +1 iff j == label_offset + i else -1 (loss.py:500-519; the reference reaches the
 * other ranks' texts by passing them round the ring, :521-558, with identical pairs).
 * 16-bit features, 8 <= dim <= 768, dim % 8 == 0 (latte_siglip_supported); logit_bias nullable.
 * Backward: d_img [n_loc, dim]; the text-side product G^T . img_loc either as d_txt [n_all, dim]
 * (only when n_loc == n_all), or as an fp32 partial over all n_all rows for the caller's
 * reduce-scatter (the backward of the ring exchange, loss.py:419-428), or added straight into
 * the owners' peer-mapped fp32 accumulators (d_txt_peers, as in latte_clip_bwd).  d_scale,
 * d_bias: this rank's d loss / d logit_scale and d loss / d logit_bias (nullable).
 */
int latte_siglip_supported(int dtype, int64_t dim);
int latte_siglip_workspace_bytes(int64_t n_loc, int64_t n_all, int64_t dim, int dtype, int backward,
                                 int own_d_txt, size_t* bytes);
int latte_siglip_fwd(const void* img_loc, int64_t ld_img, const void* txt_all, int64_t ld_txt,
                     int dtype, int64_t n_loc, int64_t n_all, int64_t dim, int64_t label_offset,
                     const float* logit_scale, const float* logit_bias, float* loss,
                     void* workspace, size_t workspace_bytes, void* stream);
int latte_siglip_bwd(const void* img_loc, int64_t ld_img, const void* txt_all, int64_t ld_txt,
                     int dtype, int64_t n_loc, int64_t n_all, int64_t dim, int64_t label_offset,
                     const float* logit_scale, const float* logit_bias, const float* grad_loss,
                     void* d_img, void* d_txt, int grad_dtype, int64_t ld_grad, float* d_txt_partial,
                     void* const* d_txt_peers, int n_peers, float* d_scale, float* d_bias,
                     void* workspace, size_t workspace_bytes, void* stream);

/*
 * Diagnostic for bench.py: runs latte_clip_fwd then latte_clip_bwd `reps` times on `stream`
 * with CUDA events recorded on that stream around every kernel stage, synchronises the
 * stream, and returns the mean milliseconds per stage in stage_ms[LATTE_NUM_STAGES] (host
 * memory).  Arguments are those of the two calls (row_lse_all / col_lse_all are the inputs of
 * the backward, row_lse / col_lse / loss the outputs of the forward).
 */
typedef enum {
  LATTE_STAGE_FWD_SWEEP = 0,     /* logit sweep(s) of the forward (tcgen05)                  */
  LATTE_STAGE_FWD_FINALIZE = 1,  /* partial merges, loss reduction, scratch initialisation   */
  LATTE_STAGE_BWD_PREP = 2,      /* LSE vectors, fp16 feature copies, accumulator zeroing    */
  LATTE_STAGE_BWD_SWEEP = 3,     /* logit recompute -> gradient weights G (tcgen05)          */
  LATTE_STAGE_BWD_GEMM = 4,      /* dI = G.T, dT = G^T.I stream-K GEMM (tcgen05)             */
  LATTE_STAGE_BWD_FINISH = 5,    /* scale + cast of the gradients, d loss / d logit_scale    */
  LATTE_NUM_STAGES = 6
} latte_stage_t;
int latte_clip_stage_times(const void* img_loc, int64_t ld_img_loc,
                           const void* txt_loc, int64_t ld_txt_loc,
                           const void* img_all, int64_t ld_img_all,
                           const void* txt_all, int64_t ld_txt_all,
                           int dtype, int64_t n_loc, int64_t n_all, int64_t dim,
                           int64_t label_offset, const float* logit_scale,
                           const float* row_lse_all, const float* col_lse_all,
                           float* row_lse, float* col_lse, float* loss,
                           const float* grad_loss, float grad_mult, int cross_terms,
                           void* d_img, void* d_txt, int grad_dtype, int64_t ld_grad,
                           float* d_txt_partial, float* d_scale,
                           void* fwd_workspace, size_t fwd_workspace_bytes,
                           void* bwd_workspace, size_t bwd_workspace_bytes,
                           void* stream, int reps, float* stage_ms);

/* One real latte_clip_bwd call (same arguments; with `comm` every rank calls it for the same
 * generation) with CUDA events around its stages: synchronises the stream and returns the
 * milliseconds of THIS call's backward stages in stage_ms[LATTE_NUM_STAGES]. */
int latte_clip_bwd_stage_times(const void* img_loc, int64_t ld_img_loc,
                               const void* txt_loc, int64_t ld_txt_loc,
                               const void* img_all, int64_t ld_img_all,
                               const void* txt_all, int64_t ld_txt_all,
                               int dtype, int64_t n_loc, int64_t n_all, int64_t dim,
                               int64_t label_offset, const float* logit_scale,
                               const float* row_lse_all, const float* col_lse_all,
                               const float* row_nll_all, const float* col_nll_all,
                               const float* lse_stats, const float* grad_loss,
                               float grad_mult, int cross_terms,
                               void* d_img, void* d_txt, int grad_dtype, int64_t ld_grad,
                               float* d_txt_partial, const latte_comm_t* comm, int phases,
                               float* d_scale, void* workspace, size_t workspace_bytes,
                               void* stream, float* stage_ms);

/* ---- prototype / pseudo-label path: src/training/train.py ---------------------------- */

/* out[c,:] = in[c,:] / max(||in[c,:]||_2, 1e-12)   (F.normalize(dim=1), train.py:388;
 * zero_shot.py:143).  fp32 [C, dim]. */
int latte_normalize_rows(const float* in, int64_t ld_in, float* out, int64_t ld_out,
                         int64_t rows, int64_t dim, void* stream);

/*
 * Fused N x C similarity + row reductions, logits never stored:
 *   sim[i,c] = <x[i,:], protos[c,:]>   (fp32 accumulate)
 *   argmax_out[i] = first c maximising sim[i,c]          (train.py:410-411; zero_shot.py:40)
 *   margin_out[i] = top1(sim[i,:]) - top2(sim[i,:])      (compute_text_weights, train.py:292-303)
 *   top1_out[i]   = scale * max_c sim[i,c]
 * Any of the three outputs may be NULL.  x is [n, dim] in x_dtype; protos is fp32 [C, dim].
 * If row_index is not NULL, row i of x is x[row_index[i], :] (class-text gather,
 * train.py:420-438).  workspace: caller-owned scratch of latte_nxc_workspace_bytes()
 * bytes (16-bit operand planes of the tensor-core path; 0 bytes below 128 rows).
 */
int latte_nxc_workspace_bytes(int x_dtype, int gathered, int64_t n, int64_t dim,
                              int64_t num_classes, size_t* bytes);
int latte_nxc_argmax_margin(const void* x, int64_t ldx, int x_dtype,
                            const int64_t* row_index,
                            int64_t n, int64_t dim,
                            const float* protos, int64_t ldp, int64_t num_classes,
                            float scale,
                            int64_t* argmax_out, float* margin_out, float* top1_out,
                            void* workspace, size_t workspace_bytes, void* stream);

/*
 * The step's N x C products as ONE HBM-bound launch (C <= 64, e.g. DTD's 47 classes; SURVEY 2c
 * K8/K9: the pseudo-label argmax of train.py:410-411 and the compute_text_weights margins of
 * train.py:292-303 / :444-449 share a launch).  Prototypes are split once per matrix into 16-bit
 * operand planes by latte_nxc_split_prototypes (three bf16 planes for bf16 rows; two fp16 planes,
 * v = h0 + 2^-11 h1, for fp16 / fp32 rows) -- optionally fused with the row normalisation of
 * train.py:384-389 (normalized_out, nullable, then receives F.normalize(protos) in fp32) -- and
 * every job streams its feature rows exactly once: fp32 rows are split into two fp16 planes inside
 * the kernel (no materialised copies), 16-bit rows are used as they are.  Results are as accurate
 * as an fp32 FMA loop (see latte_nxc_argmax_margin) for |values| < 65504.  All jobs of one call
 * must share the operand class (all 16-bit, or all fp32).  planes: latte_nxc_planes_bytes()
 * bytes, 128-byte aligned, caller-owned.
 */
typedef struct latte_nxc_job {
  const void* x; int64_t ldx; int x_dtype;     /* [n, dim] feature rows                          */
  int64_t n, dim;
  const void* planes; int64_t num_classes;     /* from latte_nxc_split_prototypes, C <= 64        */
  float scale;                                 /* top1_out = scale * max_c sim                    */
  int64_t* argmax_out; float* margin_out; float* top1_out;   /* [n] each, nullable                */
} latte_nxc_job_t;
int latte_nxc_planes_bytes(int64_t num_classes, int64_t dim, size_t* bytes);
int latte_nxc_split_prototypes(const float* protos, int64_t ld, int64_t num_classes, int64_t dim,
                               int normalize, void* planes, float* normalized_out, int64_t ld_norm,
                               void* stream);
int latte_nxc_multi(const latte_nxc_job_t* jobs, int njobs, void* stream);

/*
 * Fused N x C similarity + top-k class ids (k <= 16), for the zero-shot evaluator
 * (zero_shot.py:14-20,40: logits.topk(max(topk)); train.py:1352-1358).
 * topk_idx is [n, k] int64, topk_val [n, k] fp32 (scale * sim), sorted descending,
 * lowest index first among equal values.  workspace: latte_nxc_workspace_bytes(x_dtype, 0, ...).
 */
int latte_nxc_topk(const void* x, int64_t ldx, int x_dtype, int64_t n, int64_t dim,
                   const float* protos, int64_t ldp, int64_t num_classes, float scale,
                   int k, int64_t* topk_idx, float* topk_val, void* workspace,
                   size_t workspace_bytes, void* stream);

/*
 * Text mixture + EMA toward the memory-bank rows (train.py:472-488), gathers fused:
 *   L_ft = class_text[preds], L_zs = class_text[zs], M_ft = bank[preds], M_zs = bank[zs]
 *   t_ft = M_ft + alpha * ((w_lbl (*) L_ft + w_img*P + w_grp*G) / (w_lbl   + w_img + w_grp) - M_ft)
 *   t_zs = M_zs + alpha * ((w_lbl (*) L_zs + w_img*P + w_grp*G) / (w_lbl_zs+ w_img + w_grp) - M_zs)
 * (*) per label_axis.  class_text/per_image/per_group/t_ft/t_zs share `dtype`; bank and
 * the weight vectors are fp32.  Weights are the already flag-scaled, detached margins
 * (train.py:444-469).
 */
int latte_mix_ema_fwd(const void* class_text, int64_t ld_ct,
                      const void* per_image, int64_t ld_pi,
                      const void* per_group, int64_t ld_pg,
                      const float* bank, int64_t ld_bank,
                      const int64_t* preds, const int64_t* zs,
                      const float* w_lbl, const float* w_lbl_zs,
                      const float* w_img, const float* w_grp,
                      float alpha, int label_axis, int dtype,
                      int64_t batch, int64_t dim, int64_t num_classes,
                      void* t_ft, void* t_zs, int64_t ld_out, void* stream);

/*
 * Backward of latte_mix_ema_fwd w.r.t. the three text sources (weights are detached in
 * the reference, train.py:444-449).  d_class_text [C, dim] fp32 is ACCUMULATED with a
 * deterministic per-class segment sum (must be zeroed by the caller or hold a running
 * gradient); d_per_image / d_per_group are [batch, dim] in `dtype`.  d_bank (nullable,
 * fp32 [C, dim], accumulated) receives (1 - alpha) * d_t scattered by preds / zs -- the
 * reference sends that gradient into memory-bank Parameters that the optimizer no
 * longer owns, so callers normally pass NULL.
 * workspace: caller-owned scratch of latte_seg_workspace_bytes(batch, dim, num_classes) bytes
 * (class-sorted entry order + per-piece partial sums of the segment sums).
 */
int latte_seg_workspace_bytes(int64_t batch, int64_t dim, int64_t num_classes, size_t* bytes);
int latte_mix_ema_bwd(const void* d_t_ft, const void* d_t_zs, int64_t ld_dt,
                      const int64_t* preds, const int64_t* zs,
                      const float* w_lbl, const float* w_lbl_zs,
                      const float* w_img, const float* w_grp,
                      float alpha, int label_axis, int dtype,
                      int64_t batch, int64_t dim, int64_t num_classes,
                      float* d_class_text, int64_t ld_dct,
                      void* d_per_image, void* d_per_group, int64_t ld_dp,
                      float* d_bank, int64_t ld_dbank,
                      void* workspace, size_t workspace_bytes, void* stream);

/*
 * Memory-bank update, step 1 (train.py:508-526): per-class sums and counts, accumulated
 * in the reference's order (for each sample i ascending: t_zs[i] into class zs[i], then
 * t_ft[i] into class preds[i]).  sums [C, dim] fp32 and counts [C] fp32 are OVERWRITTEN.
 * Deterministic (no atomics).  On several ranks the caller all-reduces sums and counts
 * between step 1 and step 2.  workspace: latte_seg_workspace_bytes(batch, dim, num_classes).
 */
int latte_bank_accumulate(const void* t_ft, const void* t_zs, int64_t ld_t, int dtype,
                          const int64_t* preds, const int64_t* zs,
                          int64_t batch, int64_t dim, int64_t num_classes,
                          float* sums, int64_t ld_sums, float* counts,
                          void* workspace, size_t workspace_bytes, void* stream);

/*
 * Memory-bank update, step 2 (train.py:528-530): for every class with counts[c] > 0,
 * bank[c,:] = normalize(sums[c,:] / counts[c]); other rows are left untouched.
 */
int latte_bank_finalize(const float* sums, int64_t ld_sums, const float* counts,
                        float* bank, int64_t ld_bank,
                        int64_t dim, int64_t num_classes, void* stream);

/*
 * DistillClipLoss (src/open_clip/loss.py:324-362: soft-target cross-entropy of the student's logits
 * against the teacher's softmax, both directions), world size 1, fp16 operands [n, dim] with
 * dim <= 768 (latte_prep_features makes them).  With W(S) = softmax_rows(S) + softmax_cols(S):
 *   loss     = 1/(2n) [ sum row_lse_S + sum col_lse_S - s <I, W(S') T> ]      (S' = teacher logits)
 *   dL/dI    = s/(2n) ( W(S) - W(S') ) T,   dL/dT = s/(2n) ( W(S) - W(S') )^T I
 *   dL/ds    = 1/(2n) ( <I, W(S) T> - <I, W(S') T> )
 * latte_distill_products computes out_img = W(S) @ gemm_txt and out_txt = W(S)^T @ gemm_img (fp32) for
 * S = s * sweep_img @ sweep_txt^T without storing S or W beyond the blocked fp16 scratch of
 * latte_clip_bwd (workspace: latte_clip_bwd_workspace_bytes(n, n, dim, LATTE_F16)); row_lse / col_lse
 * are the LSE vectors of S from latte_clip_fwd.  The forward calls it with the teacher's features as
 * the sweep pair and the student's as the gemm pair, the backward with the student's for both.
 * latte_distill_loss / latte_distill_bwd_combine are the two streaming reductions around it (`aux`:
 * latte_distill_aux_bytes of scratch).
 */
int latte_distill_aux_bytes(size_t* bytes);
int latte_distill_products(const void* sweep_img, int64_t ld_sweep_img, const void* sweep_txt,
                           int64_t ld_sweep_txt, const void* gemm_img, int64_t ld_gemm_img,
                           const void* gemm_txt, int64_t ld_gemm_txt, int64_t n, int64_t dim,
                           const float* logit_scale, const float* row_lse, const float* col_lse,
                           float* out_img, float* out_txt, int64_t ld_out, void* workspace,
                           size_t workspace_bytes, void* stream);
int latte_distill_loss(const float* row_lse, const float* col_lse, int64_t n, const void* img,
                       int64_t ld_img, const float* teacher_prod, int64_t ld_prod, int64_t dim,
                       const float* logit_scale, float* loss, float* dot_out, void* aux,
                       size_t aux_bytes, void* stream);
int latte_distill_bwd_combine(const float* a_s, const float* a_t, const float* b_s, const float* b_t,
                              int64_t ld_prod, const void* img, int64_t ld_img, int64_t n, int64_t dim,
                              const float* logit_scale, const float* grad_loss, const float* dot_t,
                              void* d_img, void* d_txt, int grad_dtype, int64_t ld_grad, float* d_scale,
                              void* aux, size_t aux_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LATTE_B200_H_ */

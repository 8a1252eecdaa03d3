/*
 * htcn.h -- C ABI of the B200-native HierTCN hot path (libhtcn.so, sm_100a).
 *
 * The reference (JiaxuanYou/HierTCN, TensorFlow 1.6) has NO plugin / operator / FFI interface: its
 * only seams are Python call signatures and the feed_dict/fetch contract.  This
 * header therefore DEFINES the boundary: one entry point per op group of the reference graph, each
 * citing the reference code it replaces.  hiertcn_b200/_cabi.py binds it with ctypes; INTEGRATION.md
 * shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success or a negative htcn_status; htcn_last_error() gives the
 *     thread-local message.  No exceptions cross the boundary.
 *   - the CALLER owns every buffer.  All data pointers are DEVICE pointers unless a name ends in
 *     _host.  Nothing is allocated inside; nothing is retained after return.
 *   - every call only ENQUEUES work on `stream` (a cudaStream_t passed as void*; NULL = default
 *     stream) and is asynchronous with respect to the host.
 *   - row-major, channels-last.  D (embedding width) = C (TCN channels) = H (GRU units) = 128, the
 *     value the reference hard-codes (model_hier.py:50,85; model_tcn.py:35; args.py:51,56).
 *   - ids are int32, 0 = null item (model.py:59-61).
 *   - there is NO CPU fallback: on a machine without an sm_100 device every compute entry point
 *     returns HTCN_ERR_CUDA.
 */
#ifndef HTCN_H_
#define HTCN_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HTCN_ABI_VERSION 6
#define HTCN_DIM 128          /* D = C = H */
#define HTCN_MAX_SLOTS 64     /* S (args.max_session_num, default 10) */
#define HTCN_MAX_LEVELS 8     /* TCN levels (len(args.tcn_channel)) */
#define HTCN_MAX_GRU_LAYERS 4 /* args.num_layer */
#define HTCN_MAX_TOPK 128
#define HTCN_WT_PITCH_BF16 144 /* bf16 elements per W_out^T row in the bf16 tier (see htcn_prepare_wout) */

typedef enum {
  HTCN_OK = 0,
  HTCN_ERR_INVALID = -1,      /* bad argument (null pointer, unsupported size)        */
  HTCN_ERR_CUDA = -2,         /* CUDA runtime/driver error, message has the details   */
  HTCN_ERR_UNSUPPORTED = -3   /* valid request this build does not implement          */
} htcn_status;

/* HTCN_F32_W256: fp32 tier with 256-wide user embeddings (the last TCN level has 129..256 channels, e.g. the reference's
 * single-level default [128,128,128,128,256,256], args.py:310-311): hout is block-planar [2][Q][128], w_out_t rows are 256
 * floats.  Accepted as `precision` / dtype by htcn_prepare_wout, htcn_target_logit, htcn_score_ce_rank_topk (k <= 64 with
 * HTCN_SCORE_TOPK), htcn_score_logits (hout_dtype; w_dtype must be HTCN_F32), htcn_score_topk, htcn_score_ce_rank_topk_fused. */
typedef enum { HTCN_F32 = 0, HTCN_BF16 = 1, HTCN_F32_W256 = 2 } htcn_dtype;

/* flags of htcn_score_ce_rank_topk */
#define HTCN_SCORE_CE   1u    /* softmax cross-entropy partials            (loss.py:20-21)   */
#define HTCN_SCORE_RANK 2u    /* strict-greater rank count                 (loss.py:179)     */
#define HTCN_SCORE_TOPK 4u    /* per-row top-k (score desc, index asc)     (loss.py:120)     */

/* loss kinds of htcn_sampled_rank_loss (reference loss.py:22-71) */
typedef enum {
  HTCN_LOSS_NCE = 0, HTCN_LOSS_HINGE_SIGMOID = 1, HTCN_LOSS_HINGE_LOGSIGMOID = 2,
  HTCN_LOSS_HINGE_LINEAR = 3, HTCN_LOSS_BPR = 4
} htcn_sampled_loss;

int32_t htcn_abi_version(void);
const char* htcn_last_error(void);
/* 1 if the current device is compute capability 10.x, 0 otherwise (never errors) */
int32_t htcn_device_ok(void);

/* ---------------------------------------------------------------------------------------------
 * K1  embedding gather + session mean-pool.
 * Replaces: tf.one_hot(id,N)*sign(id) (model.py:59-61) followed by dense(...,'emb') -- a
 * [B*L,N]x[N,128] GEMM on a materialised one-hot (model_hier.py:50 -> customized_dense_layer.py:155-172)
 * -- and the mean-pool + dense-with-bias of the true items (model_hier.py:83-85).
 *   xe[b,p,:]  = E[x_id[b,p],:]  (id 0 -> zeros; bit-exact copy; optionally rounded to bf16)
 *   yp[s,b,:]  = (sum_{p in slot s, y>0} E[y_id[b,p],:]) / n_{b,s} + emb_bias      (NaN if n == 0, like 0/0 in TF)
 * x_id,y_id [B,T] int32; slot_off_host [S+1] int32 HOST array (slot s = columns slot_off[s]..slot_off[s+1]).
 * emb_table [item_num, emb_pitch] f32: emb_pitch floats per row (multiple of 4, <= 128; 128 = the reference's width).
 *   A narrower embedding (BASELINE config 2: 100-d) is stored packed -- 400 B rows -- and the output rows are zero
 *   from column emb_pitch on.  emb_bias [128] (zero beyond the embedding width).
 * xe may be NULL (skip the x gather) and yp may be NULL (skip the pool).
 * ------------------------------------------------------------------------------------------- */
int32_t htcn_gather_meanpool(const float* emb_table, int32_t emb_pitch, const float* emb_bias, int32_t item_num,
                             const int32_t* x_id, const int32_t* y_id, const int32_t* slot_off_host,
                             int32_t B, int32_t T, int32_t S,
                             void* xe, int32_t xe_dtype, float* yp, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K3  GRU over sessions (+ hoisted state half of the TCN in-projection).
 * Replaces: S Python-unrolled calls of rnn.MultiRNNCell([GRUCell(H)]*G, state_is_tuple=False)
 * (model_hier.py:30-37,91; formula customed_gru_cell.py:309-337, stacking :1050-1073, linear
 * :1187-1197), the reset state *= mask[i] (model_hier.py:93) and the state columns of
 * tile+concat+dense (model_hier.py:54-55 + model_tcn.py:35).
 *   for s in 0..S-1:  state_pre[s] = state;  sbias[s] = state @ w_in_state;
 *                     state = mask[s] * GRU_stack(yp[s], state)
 * yp [S,B,128]; mask [S,B]; state_in [B,G*128]; gate_w[g] [(in+128),256] (r first, then u),
 * gate_b[g] [256], cand_w[g] [(in+128),128], cand_b[g] [128]  (in = 128 for every layer);
 * w_in_state [G*128,128] = rows D.. of hier/tcn/emb/kernel.
 * Outputs: state_pre [S,B,G*128] (may be NULL), sbias [S,B,128] (may be NULL), state_out [B,G*128].
 * precision: HTCN_F32 = fp32 FFMA (1e-4 tier, any num_layer <= 4);
 *            HTCN_BF16 = tcgen05 tensor cores, bf16 operands / fp32 accumulate and fp32 state (num_layer == 2 only);
 *                        needs scratch >= HTCN_GRU_SCRATCH_BYTES(B) device bytes (bf16 weight tiles), else may be NULL.
 * ------------------------------------------------------------------------------------------- */
#define HTCN_GRU_SCRATCH_BYTES(B) (14 * 128 * 128 * 2 + 4096)   /* independent of B: the fp32 state lives on chip */
int32_t htcn_gru_sessions(const float* yp, const float* mask, const float* state_in,
                          const float* const* gate_w_host, const float* const* gate_b_host,
                          const float* const* cand_w_host, const float* const* cand_b_host,
                          int32_t num_layer, const float* w_in_state,
                          int32_t B, int32_t S, int32_t precision, float* scratch,
                          float* state_pre, float* sbias, float* state_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K2  low-level TCN: in-projection + dilated causal conv residual stack.
 * Replaces: model_tcn's dense(x,128,use_bias=False,'emb') (model_tcn.py:35) and
 * TemporalConvNet -> TemporalBlock -> CausalConv1D (customized_tcn_cell.py:46-49,109-127,147-161;
 * customized_convolution_layer.py:171-198): per level l, d = 2^l,
 *   a = relu(b_l + sum_k h[t-(K-1-k)d] W_l[k]);  h' = relu(a + res)       (one conv per block)
 *   res = h, or h @ Wds_l + bds_l for a level that changes the width (the 1x1 Dense of customized_tcn_cell.py:102-106,
 *   123-124): ds_w_host[l] -> [128,128] f32, ds_b_host[l] -> [128] f32, entry NULL = identity residual; both arrays may
 *   be NULL.  Widths below 128 (and emb / hidden sizes below 128) are run zero-padded to 128 by the host
 *   (hiertcn_b200.weights.to_device_layout): padded channels stay exactly zero through every level.
 * Rows are the flat [B*T] positions; a sequence is one (b, slot) run of slot_off[s+1]-slot_off[s]
 * positions (for the single-level model_tcn: S = 1, T = L).
 *   h0[b,p,:] = xe[b,p,:] @ w_in_x + sbias[slot(p), b, :]      (sbias may be NULL: plain model_tcn)
 * xe [B,T,128] of xe_dtype (f32 | bf16); w_in_x [128,128] f32 (rows 0..D-1 of hier/tcn/emb/kernel);
 * conv_w_host[l] -> [K,128,128] f32, conv_b_host[l] -> [128] f32 (HOST arrays of DEVICE pointers);
 * slot_off_host [S+1] HOST array.
 * precision: HTCN_F32 = FFMA tier (1e-4); HTCN_BF16 = tcgen05 tier (2e-2; operands rounded to bf16,
 *   fp32 accumulate, level activations kept on chip).
 * out_row [B*T] int32 or NULL: destination row of each position in hout (-1 = drop the row);
 *   used to compact away padded positions before catalog scoring.
 * hout [n_out_rows,128] of hout_dtype.  scratch (device, caller-owned): f32 tier 2*B*T*128 floats (level
 *   ping-pong; 3*B*T*128 when a level has a down-sample residual); bf16 tier HTCN_TCN_SCRATCH_BYTES(n_levels,
 *   kernel_size) bytes (bf16 weight tiles).
 * The bf16 tier requires xe and hout in bf16 and (K-1)*2^(n_levels-1) <= 32 rows of causal shift; sequences longer
 * than a 128-row tile are streamed chunk by chunk (the scratch also parks each level's last 32 input rows per tile chain, up to 592 chains).
 * ------------------------------------------------------------------------------------------- */
#define HTCN_TCN_SCRATCH_BYTES(n_levels, K) \
  ((1 + (n_levels) * ((K) + 1)) * 128 * 128 * 2 + 640 + 2 * 8 * 512 + 256 + 592 * 2 * (n_levels) * 8192)
int32_t htcn_tcn_forward(const void* xe, int32_t xe_dtype, int32_t precision,
                         const float* w_in_x, const float* sbias,
                         const float* const* conv_w_host, const float* const* conv_b_host,
                         const float* const* ds_w_host, const float* const* ds_b_host,
                         int32_t n_levels, int32_t kernel_size, const int32_t* slot_off_host,
                         int32_t B, int32_t T, int32_t S,
                         const int32_t* out_row, void* hout, int32_t hout_dtype, float* scratch,
                         void* stream);

/* The fp32 conv stack with levels up to 256 channels wide (customized_tcn_cell.py:109-127 with n_outputs in 129..256):
 * activations are block-planar [P][B*T][128]; level_planes_host[l] in {1, 2} = 128-wide planes of level l's output (the
 * in-projection always produces one); conv_w_host[l] -> [P_out, P_in*K, 128, 128] f32 (block (po, pi*K + tap) =
 * W[tap][pi*128.., po*128..] zero-padded), conv_b_host[l] -> [P_out*128]; ds_w_host[l] -> [P_out, P_in, 128, 128],
 * ds_b_host[l] -> [P_out*128] (required where the plane count changes) -- the layouts hiertcn_b200.weights.to_device_layout
 * produces.  hout: [P_last][hout_plane_rows][128] f32 (compacted through out_row).  scratch: 6*B*T*128 floats.
 * fp32 tier only; the tcgen05 tier runs levels up to 128 channels. */
int32_t htcn_tcn_forward_wide(const void* xe, int32_t xe_dtype, const float* w_in_x, const float* sbias,
                              const float* const* conv_w_host, const float* const* conv_b_host,
                              const float* const* ds_w_host, const float* const* ds_b_host,
                              const int32_t* level_planes_host, int32_t n_levels, int32_t kernel_size,
                              const int32_t* slot_off_host, int32_t B, int32_t T, int32_t S, const int32_t* out_row,
                              float* hout, int64_t hout_plane_rows, float* scratch, void* stream);

/* ---------------------------------------------------------------------------------------------
 * weight preparation for K4 (one-time, at load): w_out [128,N] f32 + b_out [N] f32 (the TF layout of
 * hier/tcn/dense/{kernel,bias}, model_tcn.py:41) -> w_out_t, K-major rows:
 *   HTCN_F32 : [N,128] f32 (b_out is not used; pass it to the scoring calls)
 *   HTCN_BF16: [N,HTCN_WT_PITCH_BF16] bf16 = 128 weights | bf16(b) | bf16(b - bf16(b)) | 14 zeros -- the bias is
 *              folded into the GEMM as a 9th K=16 step (b_hi + b_lo carry it to 2^-17 relative), so the bf16
 *              scoring kernels never read b_out.
 * ------------------------------------------------------------------------------------------- */
int32_t htcn_prepare_wout(const float* w_out, const float* b_out, int32_t N, void* w_out_t, int32_t dtype,
                          void* stream);

/* ---------------------------------------------------------------------------------------------
 * K4  full-catalog scoring with fused CE / rank / top-k epilogues.  Logits never reach HBM.
 * Replaces: dense(pred, units=output_dim) (model_tcn.py:41) -> [B,T,N] logits,
 * softmax_cross_entropy_with_logits (loss.py:20-21), calc_metric_fast's strict-greater rank
 * (loss.py:179) and tf.nn.top_k ordering (loss.py:120).
 *   z[q,j] = hout[q,:] . w_out_t[n0+j,:] + b_out[n0+j],   j in [0, n_items)   (one catalog shard)
 * hout [Q,128] (f32 or bf16 per `precision`), w_out_t = the shard's rows of htcn_prepare_wout's output in the
 * same dtype ([n_items,128] f32 or [n_items,144] bf16), b_out [n_items] f32 (read by the f32 tier only).
 * y_id [Q] int32 GLOBAL target ids (needed for CE/RANK; may be NULL for TOPK only).
 * target_logit [Q] f32: z[q, y_id[q]] -- INPUT when have_target != 0 (computed by the shard that
 *   owns the id and exchanged by the caller), else computed here (requires the shard to own every id).
 * n_split: number of catalog splits processed by separate CTAs (>=1); partial results per split:
 *   part_max [n_split,Q], part_sum [n_split,Q]   sum_j exp(z - part_max)        (CE)
 *   part_cnt [n_split,Q] int32                   #{j: z > target_logit}         (RANK)
 *   topk_val [n_split,Q,k] f32, topk_idx [n_split,Q,k] int32 (global ids, unsorted; -1 = empty)  (TOPK)
 * Merge them with htcn_score_finish / htcn_topk_merge (or exchange them across ranks first).
 * ------------------------------------------------------------------------------------------- */
int32_t htcn_score_ce_rank_topk(const void* hout, int32_t precision, int32_t Q,
                                const void* w_out_t, const float* b_out, int32_t n_items, int32_t n0,
                                const int32_t* y_id, float* target_logit, int32_t have_target,
                                uint32_t flags, int32_t k, int32_t n_split,
                                float* part_max, float* part_sum, int32_t* part_cnt,
                                float* topk_val, int32_t* topk_idx, void* stream);

/* full logits (small catalogs / debugging only -- this is the [B,T,N] tensor the reference materialises,
 * model_hier.py:76-79): logits[q, j] = hout[q,:] . w_out_t[j,:] + b_out[j], fp32 fmaf chain over k in
 * both tiers (bf16 operands are widened).  logits [Q, n_items] f32. */
int32_t htcn_score_logits(const void* hout, int32_t hout_dtype, int32_t Q, const void* w_out_t,
                          int32_t w_dtype, const float* b_out, int32_t n_items, float* logits, void* stream);

/* target logits only: target_logit[q] = hout[q,:] . w_out_t[y_id[q]-n0,:] + b_out[y_id[q]-n0] when the
 * shard [n0, n0+n_items) owns y_id[q], else the entry is left untouched (pre-fill with 0 and
 * all-reduce-sum across shards).  Same arithmetic as the sweep of htcn_score_ce_rank_topk. */
int32_t htcn_target_logit(const void* hout, int32_t precision, int32_t Q,
                          const void* w_out_t, const float* b_out, int32_t n_items, int32_t n0,
                          const int32_t* y_id, float* target_logit, void* stream);

/* bf16 tier, CE (+ RANK) over a replicated or sharded catalog with the difference to the target logit FOLDED INTO THE
 * TENSOR-CORE PRODUCT (the default loss sweep of HierTCN.score).  The call writes the negated embeddings (an exact copy) into
 * the workspace and gives row q the constant K chunk [-1, -1, 0, 0, t1, t2, t3] (t1 + t2 + t3 = target_logit[q] exactly)
 * against the table's [b_hi, b_lo, b_hi, b_lo, 1, 1, 1] columns 128..134, so the accumulator is d_j = z_y - z_j in the
 * sweep's own arithmetic: the rank adds the sign bit of d_j (one instruction per logit), the CE sum 2^(-log2e d_j).  The
 * target's own column holds the rounding residue of z_y; its sign bit (evaluated by the same product) is taken out of the
 * count, which is therefore exactly #{j != y : d_j < 0} on the swept accumulators -- the strict count of loss.py:179 up to
 * exact floating-point ties.  flags: HTCN_SCORE_CE or HTCN_SCORE_CE | HTCN_SCORE_RANK; partials as htcn_score_ce_rank_topk
 * (part_max = target_logit, the reference point); workspace: HTCN_SCORE_FOLD_WS_BYTES(Q). */
#define HTCN_SCORE_FOLD_WS_BYTES(Q) (((((int64_t)(Q)) * 128 * 2 + 255) / 256 * 256) + ((int64_t)(Q)) * 4)
int32_t htcn_score_ce_rank_folded(const void* hout, int32_t Q, const void* w_out_t, int32_t n_items, int32_t n0,
                                  const int32_t* y_id, const float* target_logit, uint32_t flags, int32_t n_split,
                                  float* part_max, float* part_sum, int32_t* part_cnt, void* workspace, void* stream);

/* merge CE / rank partials of n_part (= splits x shards) parts:
 *   loss_row[q] = log sum_j exp(z_j) - target_logit[q]   (0 where y_id == 0);  rank_row[q] = sum cnt */
int32_t htcn_score_finish(const float* part_max, const float* part_sum, const int32_t* part_cnt,
                          int32_t n_part, int32_t Q, const int32_t* y_id, const float* target_logit,
                          float* loss_row, float* rank_row, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Exact per-row top-k over one catalog shard in one call (config 4): out_val / out_idx [Q,k] sorted by
 * (score desc, GLOBAL index asc) = tf.nn.top_k order (loss.py:120); idx -1 / val -inf pad rows with < k items.
 *   bf16 tier, large shards: two tensor-core sweeps with one compare per logit each -- (1) per-row maxima of
 *     column groups, whose k-th largest is a lower bound T of the row's k-th best score; (2) every logit >= T is
 *     appended to the row's candidate list (~k entries) -- then an exact selection.  If a candidate list
 *     overflows (pathological ties), *overflow_rows (device int, may be NULL) counts the affected rows and the
 *     caller should redo them with htcn_score_ce_rank_topk(HTCN_SCORE_TOPK), the heap path, which is always exact.
 *   f32 tier / small shards: the heap sweep + htcn_topk_merge, internally.
 * workspace: htcn_topk_workspace_bytes(...) bytes of device memory, caller-owned.  Merge shards with
 * htcn_topk_merge(n_part = number of shards).
 * ------------------------------------------------------------------------------------------- */
int64_t htcn_topk_workspace_bytes(int32_t precision, int32_t Q, int32_t n_items, int32_t k, int32_t n_split);
int32_t htcn_score_topk(const void* hout, int32_t precision, int32_t Q, const void* w_out_t, const float* b_out,
                        int32_t n_items, int32_t n0, int32_t k, int32_t n_split, void* workspace,
                        int64_t workspace_bytes, float* out_val, int32_t* out_idx, int32_t* overflow_rows,
                        void* stream);

/* Loss + rank + top-k of one catalog shard in TWO sweeps (instead of the three of htcn_score_ce_rank_topk(CE|RANK) +
 * htcn_score_topk): the CE / rank sweep also records the per-row maxima of its column groups (one 3-input max per logit
 * pair on top of the loss epilogue), i.e. pass 1 of the two-pass top-k; the second sweep appends the candidates, then
 * the exact selection.  Arguments as in htcn_score_ce_rank_topk (have_target = 1: target_logit is an input) and
 * htcn_score_topk; workspace = htcn_topk_workspace_bytes(...).  Partials go to part_max / part_sum / part_cnt
 * [n_split, Q] (merge with htcn_score_finish), the sorted list to out_val / out_idx [Q, k].
 * fp32 tier and shards too small for the two-pass method run the separate sweeps internally. */
int32_t htcn_score_ce_rank_topk_fused(const void* hout, int32_t precision, int32_t Q, const void* w_out_t,
                                      const float* b_out, int32_t n_items, int32_t n0, const int32_t* y_id,
                                      const float* target_logit, int32_t k, int32_t n_split, void* workspace,
                                      int64_t workspace_bytes, float* part_max, float* part_sum, int32_t* part_cnt,
                                      float* out_val, int32_t* out_idx, int32_t* overflow_rows, void* stream);

/* Exact redo of the cross-entropy rows whose bf16-tier partial sum overflowed: the tensor-core sweep sums
 * 2^((z_j - z_y) log2e) with the target logit as reference point, which exceeds fp32 when some logit beats the target by
 * more than ~88 (target probability < e^-88); such rows leave htcn_score_finish as +inf.  This call recomputes them with
 * an online-max log-sum-exp over the whole (unsharded) catalog on the same bf16 operands and overwrites loss_row; finite
 * rows cost one load.  repaired (device int, may be NULL) is incremented per redone row.  No-op in the fp32 tier. */
int32_t htcn_score_ce_repair(const void* hout, int32_t precision, int32_t Q, const void* w_out_t, int32_t n_items,
                             const float* target_logit, float* loss_row, int32_t* repaired, void* stream);

/* The same repair for ONE SHARD of a sharded catalog, applied to the shard's partials BEFORE they are exchanged: a row
 * with a non-finite part_sum in any of the shard's n_split splits gets its exact (max_j z_j, sum_j exp(z_j - max)) over
 * the shard's n_items rows written to split 0 and the neutral element (-inf, 0) to the other splits; htcn_score_finish
 * merges parts with arbitrary reference points, so the cross-shard loss comes out finite.  No-op in the fp32 tier. */
int32_t htcn_score_ce_repair_shard(const void* hout, int32_t precision, int32_t Q, const void* w_out_t, int32_t n_items,
                                   float* part_max, float* part_sum, int32_t n_split, int32_t* repaired, void* stream);

/* ---------------------------------------------------------------------------------------------
 * l2-normalised scoring head (reference model_tcn.py:42-43: pred = tf.nn.l2_normalize(pred, dim=-1) over the N logits of
 * a position), without materialising the logits: the squared norm of a logits row is the quadratic form
 *     sum_j (h . w_j + b_j)^2 = h^T G h + 2 h . c + s,   G = sum_j w_j w_j^T,  c = sum_j b_j w_j,  s = sum_j b_j^2.
 *   htcn_catalog_gram: gram [128*128 + 128 + 4] f32 = G | c | s of the table w_out_t (layout of htcn_prepare_wout for
 *     `precision`; b_out read by the f32 tier); scratch: htcn_catalog_gram_scratch_floats() floats.  Deterministic.
 *     Sharded catalogs: sum the shards' gram arrays (all-reduce) before use.
 *   htcn_logit_rownorm: row_scale[q] = rsqrt(max(||z_q||^2, eps)) (eps = 1e-12 = tf.nn.l2_normalize's epsilon).
 *   htcn_score_ce_rank_l2norm: the CE | RANK sweep of htcn_score_ce_rank_topk on the normalised logits
 *     row_scale[q] * z[q, :] (the per-row scale multiplies the exponent; ranks compare raw logits -- the order is
 *     invariant).  target_logit [Q] (raw, input); target_scaled [Q] = row_scale * target_logit (output): pass it to
 *     htcn_score_finish as the target logit.  The partial sums cannot overflow (|scaled logit| <= 1): no repair pass.
 *   htcn_scale_rows: x[r, :] *= row_scale[r]  (x [R, C] f32; e.g. top-k values of the normalised head, or the carried
 *     state times a gap decay, model_hier.py:40-47).
 * ------------------------------------------------------------------------------------------- */
int64_t htcn_catalog_gram_scratch_floats(void);
int32_t htcn_catalog_gram(const void* w_out_t, int32_t precision, const float* b_out, int32_t n_items, float* scratch,
                          float* gram, void* stream);
int32_t htcn_logit_rownorm(const void* hout, int32_t precision, int32_t Q, const float* gram, float eps, float* row_scale,
                           void* stream);
int32_t htcn_score_ce_rank_l2norm(const void* hout, int32_t precision, int32_t Q, const void* w_out_t, const float* b_out,
                                  int32_t n_items, int32_t n0, const int32_t* y_id, const float* target_logit,
                                  const float* row_scale, int32_t n_split, float* part_max, float* part_sum,
                                  int32_t* part_cnt, float* target_scaled, void* stream);
int32_t htcn_scale_rows(float* x, const float* row_scale, int64_t R, int32_t C, void* stream);

/* k-way merge of per-part top-k lists -> [Q,k] sorted by (score desc, index asc) [TF top_k order] */
int32_t htcn_topk_merge(const float* part_val, const int32_t* part_idx, int32_t n_part, int32_t Q,
                        int32_t k, float* out_val, int32_t* out_idx, void* stream);

/* ---------------------------------------------------------------------------------------------
 * masked two-level means of model.py:111-117 and loss.py:190-219.
 * loss_row / rank_row [n_rows] are indexed through row_of [B*T] int32 (NULL = identity; -1 = padded
 * position).  y_id [B,T].  Outputs: loss_bt, ranks, ranks_float [B,T] (masked, 0 at padding; any may be
 * NULL), user_part [B,8] (per-user means, also scratch for the second phase) and
 * scalars[8] = {loss, recall@1, recall@5, recall@10, mrr, mrp, user_count, n_valid}.
 * Deterministic (fixed summation order).
 * ------------------------------------------------------------------------------------------- */
int32_t htcn_loss_metrics_reduce(const float* loss_row, const float* rank_row, const int32_t* row_of,
                                 const int32_t* y_id, int32_t B, int32_t T, int32_t item_num,
                                 float* loss_bt, float* ranks, float* ranks_float, float* user_part,
                                 float* scalars, void* stream);

/* ---------------------------------------------------------------------------------------------
 * sampled ranking losses (reference loss.py:22-71): pred [Q,128] is l2-normalised, scored by inner
 * product against the positive row table[pos_id[q]] and k negatives table[neg_id[q,j]].
 * table [N,128] f32.  loss_row [Q] (0 where pos_id == 0).  The hinge / bpr kinds average over the k
 * negatives supplied (reduce_mean, loss.py:40,50,60,70); nce divides the negative term by
 * num_neg_sample = args.num_neg_sample (loss.py:31), which may differ from k (<= 0: use k).
 * ------------------------------------------------------------------------------------------- */
int32_t htcn_sampled_rank_loss(const void* pred, int32_t precision, int32_t Q, const float* table,
                               const int32_t* pos_id, const int32_t* neg_id, int32_t k,
                               int32_t loss_kind, float hinge_delta, float nce_weight,
                               int32_t num_neg_sample, float* loss_row, void* stream);

/* The same gathering from the bf16 scoring table wt [N,HTCN_WT_PITCH_BF16] of htcn_prepare_wout (bf16 tier: 288 B rows
 * instead of 512 B; pred [Q,128] bf16).  Equal, bit for bit, to htcn_sampled_rank_loss on the widened table wt[:, :128]. */
int32_t htcn_sampled_rank_loss_wt(const void* pred, int32_t Q, const void* wt_bf16, const int32_t* pos_id,
                                  const int32_t* neg_id, int32_t k, int32_t loss_kind, float hinge_delta,
                                  float nce_weight, int32_t num_neg_sample, float* loss_row, void* stream);

/* Backward of htcn_sampled_rank_loss: g_row [Q] = dL/dloss_row.  d_pred [Q,128] f32 is overwritten (rows with pos_id == 0
 * get zeros), d_table [N,128] f32 (the gradient of the positive / negative rows of `table`) is accumulated with atomics;
 * the null item 0 owns no row.  Hinge kinds use the relu gradient where the hinge is strictly positive (TF's ReluGrad).
 * k <= 32. */
int32_t htcn_sampled_rank_loss_backward(const void* pred, int32_t precision, int32_t Q, const float* table,
                                        const int32_t* pos_id, const int32_t* neg_id, int32_t k, int32_t loss_kind,
                                        float hinge_delta, float nce_weight, int32_t num_neg_sample, const float* g_row,
                                        float* d_pred, float* d_table, void* stream);

/* calc_score (reference loss.py:76-105): score[q,j] of k candidate rows table[cand_id[q,j]] for every query;
 * rank_metric 0 = 'l2' (-||pred - y'||^2), 1 = 'inner_prod' (<pred, y'>); pred is not normalised (as in the
 * reference).  cand_id [Q,k] int32 (0 -> zero row), score [Q,k] f32. */
int32_t htcn_calc_score(const void* pred, int32_t precision, int32_t Q, const float* table,
                        const int32_t* cand_id, int32_t k, int32_t rank_metric, float* score, void* stream);

/* =============================================================================================
 * Training step: replaces tf.train.AdamOptimizer(lr).minimize(loss) of model.py:134-141,
 * i.e. TensorFlow's autodiff of the graph of model.py:54-117 + the Adam update, run by
 * sess.run([train_op, loss, state, ...]) in run_hier_xing.py:301-302.  fp32 throughout (the reference's
 * precision); every gradient buffer is ACCUMULATED into (+=) unless stated, so the caller zeroes them once
 * per step (htcn_adam_step can do it while it reads them).  Gradients are those of
 *     sum_b ( sum_t loss[b,t] / (n_b + 1e-6) )            (model.py:110-116)
 * and the 1/user_count of model.py:117 is applied by htcn_adam_step (grad_div), so that data-parallel ranks
 * can all-reduce(SUM) gradients and user_count first.  The carried state is a placeholder in the reference
 * (model.py:44): no gradient flows into state_in.
 * ============================================================================================= */

/* g_row[row_of[b,t]] = 1 / (n_b + 1e-6) for every scored position; y_id [B,T], row_of [B*T], g_row [Q]. */
int32_t htcn_loss_row_weights(const int32_t* y_id, const int32_t* row_of, int32_t B, int32_t T, float* g_row,
                              void* stream);

/* Backward of dense(pred, N) + softmax_cross_entropy_with_logits (model_tcn.py:41, loss.py:20-21) without
 * materialising [Q,N]: recomputes the logits tile by tile from hout [Q,128] (f32 or bf16) and the fp32 table
 * w_out_t [n_items,128] / b_out, using lse_row = loss_row + target_logit from the forward sweep.
 * d_hout [Q,128] is overwritten; d_w_out_t [n_items,128] and d_b_out [n_items] are accumulated. */
int32_t htcn_score_ce_backward(const void* hout, int32_t hout_dtype, int32_t Q, const float* w_out_t,
                               const float* b_out, int32_t n_items, int32_t n0, const int32_t* y_id,
                               const float* loss_row, const float* target_logit, const float* g_row,
                               float* d_hout, float* d_w_out_t, float* d_b_out, void* stream);

/* The same backward on the tensor cores (bf16 operands, fp32 accumulation and outputs): two tcgen05 passes, each
 * recomputing its logits tile in TMEM, turning it into dL/dZ in the epilogue and feeding it straight back to the tensor
 * core as the operand of the second product.  Operands (all bf16, built with htcn_cast_transpose_bf16 /
 * htcn_refresh_wout):  hout [Q,128], hout_t [128,q_pad] (its transpose), w_out_t [n_items,144] (scoring layout),
 * w_out [128,n_pad] (TF layout).  q_pad, n_pad: row pitches, multiples of 8.  workspace:
 * HTCN_CE_BWD_BF16_WS_FLOATS(Q, n_items) floats.  Outputs as htcn_score_ce_backward. */
#define HTCN_CE_BWD_BF16_WS_FLOATS(Q, N) (4 * ((((int64_t)(Q)) + 63) / 64 * 64) + ((((int64_t)(N)) + 63) / 64 * 64))
int32_t htcn_score_ce_backward_bf16(const void* hout, const void* hout_t, int64_t q_pad, int32_t Q,
                                    const void* w_out_t, const void* w_out, int64_t n_pad, const float* b_out,
                                    int32_t n_items, int32_t n0, const int32_t* y_id, const float* loss_row,
                                    const float* target_logit, const float* g_row, float* workspace, float* d_hout,
                                    float* d_w_out_t, float* d_b_out, void* stream);

/* Loss AND gradients of the head in two catalog sweeps instead of three (training without the rank metrics): with the
 * target logit as the reference point of the softmax sum (the forward sweep's convention) O_q = sum_j exp(z_j - z_y) w_j
 * and l_q = sum_j exp(z_j - z_y) need no running maximum, so ONE sweep shaped like htcn_score_ce_backward_bf16's pass A
 * yields loss_row[q] = log l_q (written here, [Q] f32) and d_hout[q] = g_q (O_q / l_q - w_y); the second sweep is pass B
 * (d_w_out_t, d_b_out).  A row whose sum leaves the fp32 range (a logit ~69 nats above the target's) is redone exactly
 * with a running maximum.  Same operands and workspace as htcn_score_ce_backward_bf16; replaces
 * htcn_score_ce_rank_topk(HTCN_SCORE_CE) + htcn_score_finish + htcn_score_ce_repair + htcn_score_ce_backward_bf16. */
int32_t htcn_score_ce_fwd_bwd_bf16(const void* hout, const void* hout_t, int64_t q_pad, int32_t Q,
                                   const void* w_out_t, const void* w_out, int64_t n_pad, const float* b_out,
                                   int32_t n_items, int32_t n0, const int32_t* y_id, const float* target_logit,
                                   const float* g_row, float* workspace, float* loss_row, float* d_hout,
                                   float* d_w_out_t, float* d_b_out, void* stream);

/* src [R,128] (f32 or bf16) -> dst_rows [R,128] bf16 (or NULL) and dst_t [128,r_pad] bf16 = its transpose (or NULL;
 * columns R..r_pad-1 are zero).  r_pad % 8 == 0. */
int32_t htcn_cast_transpose_bf16(const void* src, int32_t src_dtype, int64_t R, void* dst_rows, void* dst_t,
                                 int64_t r_pad, void* stream);

/* K2 forward that keeps what the backward needs: h_save [(n_levels+1), B*T, 128] (the in-projection output and
 * every level's output) and a_save [n_levels, B*T, 128] (relu(conv+bias) before the residual add);
 * hout [Q,128] f32 = rows of the last level compacted through out_row [B*T] (-1 = not scored).
 * dropout_scale [S, n_levels, 128] f32 or NULL: training dropout of customized_tcn_cell.py:100,119 -- tf.layers.Dropout with
 * noise_shape [1,1,C], i.e. ONE channel mask per dropout op, shared over batch and time; the graph has one op per session
 * slot and level.  Entries are 0 or 1/(1-rate) (drawn by the caller); relu(conv+b) is multiplied by them before the
 * residual add; a_save keeps the value before the scale.  The same array goes to htcn_tcn_backward.
 * tc_scratch (device, HTCN_K2TC_SCRATCH_BYTES) or NULL: when given, every level runs on the tensor cores with fp32-grade
 * products -- activations and weights split into two bf16 each, x w ~= x_hi w_hi + x_lo w_hi + x_hi w_lo (three tcgen05
 * MMAs, fp32 accumulation, ~1e-5 relative), so the saved activations and ReLU gates are those of the fp32 stack; levels the
 * tensor-core kernel does not take (a slot longer than 128 - (K-1)*d positions) and NULL run the FFMA kernel. */
#define HTCN_K2TC_SCRATCH_BYTES (8 * 2 * 128 * 128 * 2)
int32_t htcn_tcn_forward_train(const float* xe, const float* w_in_x, const float* sbias,
                               const float* const* conv_w_host, const float* const* conv_b_host,
                               const float* const* ds_w_host, const float* const* ds_b_host, int32_t n_levels,
                               int32_t kernel_size, const int32_t* slot_off_host, int32_t B, int32_t T, int32_t S,
                               const int32_t* out_row, const float* dropout_scale, float* h_save, float* a_save, float* hout,
                               void* tc_scratch, void* stream);

/* The same on the tensor cores: the fused tcgen05 conv stack (htcn_tcn_forward, HTCN_BF16) with every layer's output
 * (h_save [(n_levels+1), B*T, 128]) and every level's pre-residual activation (a_save [n_levels, B*T, 128]) written out
 * in bf16.  xe [B*T,128] bf16, hout [Q,128] bf16; scratch as htcn_tcn_forward's bf16 tier (HTCN_TCN_SCRATCH_BYTES). */
int32_t htcn_tcn_forward_train_bf16(const void* xe, const float* w_in_x, const float* sbias,
                                    const float* const* conv_w_host, const float* const* conv_b_host,
                                    const float* const* ds_w_host, const float* const* ds_b_host, int32_t n_levels,
                                    int32_t kernel_size, const int32_t* slot_off_host, int32_t B, int32_t T, int32_t S,
                                    const int32_t* out_row, const float* dropout_scale, void* h_save, void* a_save,
                                    void* hout, float* scratch, void* stream);

/* Backward of the conv stack + in-projection (customized_tcn_cell.py:109-127, model_tcn.py:35).
 * xe, h_save, a_save are of save_dtype (HTCN_F32 from htcn_tcn_forward_train, HTCN_BF16 from ..._train_bf16); the
 * gradients are fp32.  scratch: 2*B*T*128 floats (3*B*T*128 with a down-sample level).  d_conv_w[l] [K,128,128],
 * d_conv_b[l] [128], d_ds_w[l] [128,128], d_ds_b[l] [128] (levels with ds_w_host[l] != NULL), d_w_in_x [128,128]
 * accumulated; d_sbias [S,B,128] and d_xe [B*T,128] overwritten.
 * tc_scratch (device, htcn_tcn_backward_tc_scratch_bytes(...) bytes) or NULL: when given, the weight gradients
 * dW[tap] = h[shifted]^T dp run on the tensor cores (tcgen05, bf16 operands, fp32 accumulation) from zero-padded
 * transposed bf16 copies of the activations written into it, and the data gradients (transposed convolutions, dXe) as
 * plain bf16 tcgen05 products (k2_level_tc.cu; the ReLU gates are the saved forward's); NULL = fp32 split-K products and
 * FFMA levels (the 1e-4 tier). */
int64_t htcn_tcn_backward_tc_scratch_bytes(int32_t B, int32_t T, int32_t S, int32_t n_levels, int32_t kernel_size);
int32_t htcn_tcn_backward(const float* d_hout, const int32_t* out_row, const void* xe, int32_t save_dtype,
                          const float* w_in_x, const float* const* conv_w_host, const float* const* ds_w_host,
                          int32_t n_levels, int32_t kernel_size,
                          const int32_t* slot_off_host, int32_t B, int32_t T, int32_t S, const float* dropout_scale,
                          const void* h_save, const void* a_save, float* scratch, void* tc_scratch,
                          float* const* d_conv_w_host,
                          float* const* d_conv_b_host, float* const* d_ds_w_host, float* const* d_ds_b_host,
                          float* d_w_in_x, float* d_sbias, float* d_xe, void* stream);

/* K3 fp32 forward that also saves state_pre [S,B,G*128] and gates_save [S,G,3,B,128] (r, u, c of every cell). */
int32_t htcn_gru_sessions_train(const float* yp, const float* mask, const float* state_in,
                                const float* const* gate_w_host, const float* const* gate_b_host,
                                const float* const* cand_w_host, const float* const* cand_b_host, int32_t num_layer,
                                const float* w_in_state, int32_t B, int32_t S, float* state_pre, float* sbias,
                                float* state_out, float* gates_save, void* stream);

/* The same on the tensor cores (bf16 operands, fp32 state and accumulation): the users-on-N tcgen05 kernel of
 * htcn_gru_sessions' bf16 tier with r, u, c of every cell call written out (customed_gru_cell.py:309-337).  num_layer must
 * be 2; scratch: HTCN_GRU_SCRATCH_BYTES(B) bytes. */
int32_t htcn_gru_sessions_train_bf16(const float* yp, const float* mask, const float* state_in,
                                     const float* const* gate_w_host, const float* const* gate_b_host,
                                     const float* const* cand_w_host, const float* const* cand_b_host, int32_t num_layer,
                                     const float* w_in_state, int32_t B, int32_t S, float* scratch, float* state_pre,
                                     float* sbias, float* state_out, float* gates_save, void* stream);

/* Back-propagation through the S session steps (model_hier.py:30-37,91,93), truncated at the batch boundary.
 * d_sbias [S,B,128] comes from htcn_tcn_backward.  scratch: htcn_gru_backward_scratch_floats(B,S,G) floats.
 * Weight / bias gradients and d_w_in_state [G*128,128] are accumulated; d_yp [S,B,128] is overwritten.
 * Requires G*128 == 256. */
int64_t htcn_gru_backward_scratch_floats(int32_t B, int32_t S, int32_t num_layer);
int32_t htcn_gru_backward(const float* yp, const float* mask, const float* state_pre, const float* gates_save,
                          const float* const* gate_w_host, const float* const* cand_w_host, int32_t num_layer,
                          const float* w_in_state, int32_t B, int32_t S, const float* d_sbias, float* scratch,
                          float* const* d_gate_w_host, float* const* d_gate_b_host, float* const* d_cand_w_host,
                          float* const* d_cand_b_host, float* d_w_in_state, float* d_yp, void* stream);

/* Backward of K1 (model.py:59-61, model_hier.py:50,83-85): d_emb[x] += d_xe[b,t]; d_emb[y] += d_yp[s,b] / n_{b,s};
 * d_emb_bias += sum_{s,b} d_yp[s,b].  d_emb [item_num,128] is the DENSE gradient TensorFlow produces. */
int32_t htcn_gather_backward(const float* d_xe, const float* d_yp, const int32_t* x_id, const int32_t* y_id,
                             const int32_t* slot_off_host, int32_t B, int32_t T, int32_t S, int32_t item_num,
                             float* d_emb, float* d_emb_bias, void* stream);

/* tf.train.AdamOptimizer update over a flat parameter buffer of n floats (n % 4 == 0):
 *     g' = grad / *grad_div (grad_div: device scalar, e.g. the global user_count; NULL = 1)
 *     m = b1 m + (1-b1) g';  v = b2 v + (1-b2) g'^2;  param -= lr_t m / (sqrt(v) + eps)
 * with lr_t = lr sqrt(1 - b2^t) / (1 - b1^t) computed by the caller (t = 1-based step).  zero_grad != 0 clears
 * grad for the next step. */
int32_t htcn_adam_step(float* param, float* grad, float* m, float* v, int64_t n, float lr_t, float beta1, float beta2,
                       float eps, const float* grad_div, int32_t zero_grad, void* stream);

/* After an update of the fp32 master W_out^T [N,128] / b_out: rebuild the bf16 scoring layout [N,144]. */
int32_t htcn_refresh_wout(const float* w_out_t_f32, const float* b_out, int32_t N, void* w_out_t, int32_t dtype,
                          void* stream);

/* ---------------------------------------------------------------------------------------------
 * Device-side batch assembly.  Replaces the host loop of data_loader.py:233-272 (dequeue: pad the
 * sessions at the head of every slot's queue, x = y shifted by one, reset mask) for an interaction log resident in HBM:
 *   items [n_events] int32 (item ids, sessions contiguous), sess_off [n_sessions+1] int32,
 *   sched_sess [B, sched_pitch] int32 / sched_last [B, sched_pitch] uint8: for every user slot the session ids in
 *   dequeue order and whether a session is its user's last (the queue discipline of data_loader.py:170-231, replayed
 *   once on the host).  The batch takes sessions first_session .. first_session+S-1 of every slot.
 * Outputs: x_id, y_id [B, S*L] (every slot padded to L = max_activity_len), mask [S,B], and the compaction of the scored
 * positions: row_of [B*S*L] (-1 = padding), y_rows [<= B*S*L], n_valid[1] (= Q).  scratch: htcn_batcher_scratch_ints ints.
 * ------------------------------------------------------------------------------------------- */
int64_t htcn_batcher_scratch_ints(int32_t B, int32_t T);
int32_t htcn_assemble_batch(const int32_t* items, const int32_t* sess_off, const int32_t* sched_sess,
                            const uint8_t* sched_last, int32_t sched_pitch, int32_t first_session, int32_t B, int32_t S,
                            int32_t L, int32_t* x_id, int32_t* y_id, float* mask, int32_t* row_of, int32_t* y_rows,
                            int32_t* n_valid, int32_t* scratch, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Peer-memory exchanges of the catalog-sharded scoring path (BASELINE config 4; the reference has no distributed code).
 * One process per GPU; every rank allocates one buffer with the SAME layout (htcn_peer_alloc), publishes its CUDA IPC handle
 * (htcn_peer_export, HTCN_PEER_HANDLE_BYTES bytes -- exchange them with any host-side all-gather) and maps the other
 * ranks' buffers (htcn_peer_import).  An exchange is then ONE kernel: it stores rows straight into the peers' buffers over
 * NVLink, raises flag[kind][my rank] = epoch in every peer's buffer (system-scope release) and waits until every peer has
 * raised its flag here.  epoch must grow by one per call on every rank; *err (device int) is set to 1 if a peer did not
 * arrive within ~1.5 s.  done: HTCN_MAX_PEERS device uint32 counters, zero-initialised, private to the rank.
 *   htcn_peer_exchange   : n_peers * n_seg 2-D segments (index p * n_seg + s; n_rows rows of row_bytes, every size,
 *                          pitch and address a multiple of 16) -- the all-gather of the query rows and the all-to-all of
 *                          the per-shard partials (max, sum, count, top-k values, top-k indices);
 *   htcn_peer_bcast_owned: dst[p][q] = src[q] for the rows whose target id this shard owns (n0 <= y_id[q] < n1) --
 *                          completes the target-logit vector on every rank (the all-reduce of the NCCL path).
 * ------------------------------------------------------------------------------------------- */
#define HTCN_MAX_PEERS 8
#define HTCN_PEER_MAX_SEGS 6
#define HTCN_PEER_HANDLE_BYTES 64
int32_t htcn_peer_alloc(int64_t bytes, void** ptr);
int32_t htcn_peer_free(void* ptr);
int32_t htcn_peer_export(const void* ptr, uint8_t* handle);
int32_t htcn_peer_import(const uint8_t* handle, void** ptr);
int32_t htcn_peer_unimport(void* ptr);
int32_t htcn_peer_exchange(const void* const* src, void* const* dst, const int64_t* row_bytes, const int32_t* n_rows,
                           const int64_t* src_pitch, const int64_t* dst_pitch, int32_t n_seg, int32_t n_peers,
                           int32_t rank, void* const* flag_remote, void* flag_local, void* done, void* err,
                           uint32_t epoch, void* stream);
int32_t htcn_peer_bcast_owned(const float* src, const int32_t* y_id, int32_t Q, int32_t n0, int32_t n1,
                              void* const* dst, int32_t n_peers, void* const* flag_remote, void* flag_local,
                              void* done, void* err, uint32_t epoch, void* stream);

/* Data-parallel training step over peer memory (BASELINE config 5): gradient all-reduce, all-reduce of the loss / metric
 * scalars and the TF-Adam update of htcn_adam_step in ONE kernel.  grads[p] / reduced[p] / scalars[p]: rank p's flat gradient
 * buffer (n floats), reduced-gradient buffer (n floats) and scalars[8] = {loss, r@1, r@5, r@10, mrr, mrp, user_count,
 * n_valid}, all inside the ranks' symmetric buffers.  Rank r sums slice r of the gradient over the ranks in rank order (every
 * replica applies the same bits) and stores it into every rank's reduced buffer; after the second flag wait every rank
 * updates ALL parameters with reduced / sum(user_count) and clears its own gradient buffer.  scalars_out [8]: the global
 * means (weighted by user_count) and counts.  Two flag kinds (in, mid), epoch as above. */
int32_t htcn_peer_allreduce_adam(void* const* grads, void* const* reduced, void* const* scalars, int32_t n_peers,
                                 int32_t rank, float* param, float* m, float* v, int64_t n, float lr_t, float beta1,
                                 float beta2, float eps, float* scalars_out, void* const* flag_in_remote,
                                 void* const* flag_mid_remote, void* flag_in_local, void* flag_mid_local, void* done,
                                 void* err, uint32_t epoch, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HTCN_H_ */

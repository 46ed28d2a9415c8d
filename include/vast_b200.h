/*
 * vast_b200 -- C-ABI of the B200-native (sm_100a) cross-modal contrastive + retrieval-scoring
 * hot path of VAST.  Loaded with ctypes from Python (vast_b200/_lib.py); no torch types, no
 * C++ types, plain pointers and sizes only.
 *
 * The reference (DelusionalLogic/VAST) is pure Python/PyTorch and has no FFI of its own: the
 * "interface each entry point replaces" is therefore the block of reference Python lines cited
 * beside it (paths relative to the reference root).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - the library never allocates, frees or synchronises: outputs and workspace are caller
 *     allocated (query the *_workspace_bytes function), work is only enqueued on `stream`;
 *   - returns VAST_OK (0) or a negative vast_status; vast_last_error_string() describes the last
 *     failure on the calling thread.  Nothing throws, nothing exits;
 *   - stateless and re-entrant; one host thread per GPU (one process per GPU).
 */
#ifndef VAST_B200_H_
#define VAST_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VAST_B200_VERSION 100

#if defined(__GNUC__)
#define VAST_API __attribute__((visibility("default")))
#else
#define VAST_API
#endif

typedef struct CUstream_st* vast_stream_t; /* == cudaStream_t */

typedef enum { VAST_F32 = 0, VAST_BF16 = 1, VAST_F16 = 2 } vast_dtype;

typedef enum {
  VAST_OK = 0,
  VAST_ERR_INVALID = -1,     /* bad argument (null pointer, misaligned, inconsistent sizes) */
  VAST_ERR_UNSUPPORTED = -2, /* shape / dtype outside what the kernels support            */
  VAST_ERR_CUDA = -3,        /* a CUDA runtime / driver call failed                        */
  VAST_ERR_WORKSPACE = -4    /* workspace too small                                        */
} vast_status;

VAST_API int vast_version(void);
VAST_API const char* vast_last_error_string(void);
VAST_API int vast_sm_count(void);

/* Developer / benchmark aid (the only calls that synchronise): when enabled, every kernel launch of
 * the library is bracketed by CUDA events on its stream; vast_timing_read waits for them, returns how
 * many launches were recorded since the last read (<= max_entries) and fills ms_out[i] and the
 * 48-byte name slots names_out[48*i]. */
VAST_API int vast_timing_enable(int on);
VAST_API int vast_timing_read(float* ms_out, char* names_out, int max_entries);

/* ------------------------------------------------------------------------------------------
 * Feature build: pool -> concat -> (Linear stays cuBLAS) -> L2 normalise
 * ---------------------------------------------------------------------------------------- */

/* pool_vision_for_contra + pool_audio_for_contra + pool_text_for_contra + torch.cat(dim=1)
 * (model/general_module.py:426-449, model/vast.py:269-275) in one pass.
 * Each modality may be absent (NULL).  vision [bs, n_v, tok_v, c_v], audio [bs, n_a, tok_a, c_a],
 * subtitle [bs, tok_s, c_s], all contiguous, dtype `dtype`.
 * *_mode: 0 = take token 0 ("cls": clip/evaclip vision, ast audio, text), 1 = mean over tokens
 * (swin vision, beats audio); then mean over frames/clips.
 * out [bs, c_v + c_a + c_s] (dtype out_dtype, leading dimension ldo) = concat of the pooled rows. */
VAST_API int vast_pool_concat(const void* vision, int64_t n_v, int64_t tok_v, int64_t c_v, int vision_mode,
                     const void* audio, int64_t n_a, int64_t tok_a, int64_t c_a, int audio_mode,
                     const void* subtitle, int64_t tok_s, int64_t c_s, int dtype, int64_t bs,
                     void* out, int out_dtype, int64_t ldo, vast_stream_t stream);

/* Backward of vast_pool_concat: grad_out [bs, c_v+c_a+c_s] (f32) -> dense grads of the encoder
 * outputs (same shapes/dtype as the inputs; every element written, zeros where no gradient). */
VAST_API int vast_pool_concat_bwd(const float* grad_out, int64_t ldg, int64_t bs,
                         void* grad_vision, int64_t n_v, int64_t tok_v, int64_t c_v, int vision_mode,
                         void* grad_audio, int64_t n_a, int64_t tok_a, int64_t c_a, int audio_mode,
                         void* grad_subtitle, int64_t tok_s, int64_t c_s, int dtype, vast_stream_t stream);

/* F.normalize(x, dim=-1): y = x / max(||x||_2, eps)  (model/vast.py:225,232,239,246,257,267,278).
 * x [rows, dim] (dtype x_dtype, leading dim ldx).  Any of the outputs may be NULL:
 *   y_f32 [rows, dim] (ld ldy), y_16 (bf16, ld ld16; e.g. straight into the all-gather send
 *   slot), inv_norm [rows] (1 / max(||x||, eps), saved for backward). */
VAST_API int vast_l2norm(const void* x, int x_dtype, int64_t rows, int64_t dim, int64_t ldx, float eps,
                float* y_f32, int64_t ldy, void* y_16, int64_t ld16, float* inv_norm, vast_stream_t stream);

/* Backward of F.normalize: dx = inv_norm * (g - y * <g, y>)  (rows whose norm was clamped by eps
 * get dx = g * inv_norm, like ATen). */
VAST_API int vast_l2norm_bwd(const float* grad_y, int64_t ldg, const float* y, int64_t ldy, const float* inv_norm,
                    int64_t rows, int64_t dim, float eps, float* grad_x, int64_t ldgx, vast_stream_t stream);

/* Contra_head / fusion Linear + F.normalize in ONE tensor-core kernel (model/vast.py:221-279; general_module.py:26-31):
 *   y = x . W^T + bias ;  feat = y / max(|y|_2, eps)
 * x_op [rows, cols], w_op [dim_out, cols] are packed 16-bit operands (vast_sim_pack_operand: mode BF16 for 16-bit
 * features and weights, FP32X3 / FP32X2 for fp32-grade results from fp32 inputs; x as_query = 1, W as_query = 0).
 * The GEMM epilogue adds the bias, stores y and sums its squares; the last tile of every 128-row block to finish
 * normalises the block in place (still in L2) and writes y [rows, dim_out] f32 (ld ldy), the bf16 copy y16 (ld ld16,
 * e.g. straight into the all-gather slot; may be NULL) and inv_norm [rows] (may be NULL; saved for the backward,
 * vast_l2norm_bwd).  Replaces cuBLAS + two normalise kernels + their HBM round trips. */
VAST_API size_t vast_project_normalize_workspace_bytes(int64_t rows, int64_t dim_out);
VAST_API int vast_project_normalize(const void* x_op, const void* w_op, int64_t rows, int64_t dim_out, int64_t cols,
                           const float* bias, float eps, float* y, int64_t ldy, void* y16, int64_t ld16,
                           float* inv_norm, void* workspace, size_t workspace_bytes, vast_stream_t stream);

/* Pack the two local feature blocks into the 16-bit all-gather send buffer:
 * pack[b, 0:D] = bf16(feat_t[b]), pack[b, D:2D] = bf16(feat_cond[b])   (pack is [bs, 2D]).
 * Replaces the two separate concat_all_gather payloads of model/vast.py:395,404 by one. */
VAST_API int vast_pack_pair(const void* feat_t, const void* feat_cond, int dtype, int64_t bs, int64_t dim,
                   int64_t ld_in, void* pack_bf16, vast_stream_t stream);

/* Fused pack + all-gather over NVLink / NVSwitch (replaces vast_pack_pair followed by an all-gather call):
 * every 16-byte vector of this rank's packed rows is stored into the gathered [n_total, 2*dim] bf16 buffer of EVERY
 * rank, at rows [row_offset, row_offset + bs) -- with one multimem.st to `multicast_ptr` (the NVSwitch multicast
 * address of the symmetric buffer) when it is non-NULL, otherwise with one store per entry of `peer_ptrs[world]`
 * (the buffer's unicast address on each rank, this rank included).  The caller issues a cross-rank barrier on the
 * same stream afterwards (and must not let a rank run more than one step ahead of a buffer that is still being
 * read: alternate two buffers).  concat_all_gather semantics of utils/distributed.py:50-66. */
VAST_API int vast_pack_pair_push(const void* feat_t, const void* feat_cond, int dtype, int64_t bs, int64_t dim,
                        int64_t ld_in, int64_t row_offset, void* multicast_ptr, void* const* peer_ptrs, int world,
                        vast_stream_t stream);

/* The same with ARRIVAL FLAGS instead of a barrier after the kernel: once all of this rank's stores are performed
 * (system-scope fences, last block by ticket) the kernel release-stores its push epoch (1, 2, 3, ... per call) into
 * entry `my_rank` of every rank's flag array (`flag_ptrs[world]`: the unicast address of each rank's uint32[world]
 * array in symmetric memory, zeroed once and published before the first call).  `sync_state`: LOCAL int32[4], zeroed
 * once ([0] ticket, [1] push epoch, [2] consumer epoch).  Consumers wait with vast_wait_arrivals on the same stream.
 * The flags also make the two-buffer rule safe without a barrier: a rank can only signal step k+1 after its own
 * kernels of step k have completed, and nobody pushes step k+2 before having seen every rank's step k+1. */
VAST_API int vast_pack_pair_push_signal(const void* feat_t, const void* feat_cond, int dtype, int64_t bs, int64_t dim,
                               int64_t ld_in, int64_t row_offset, void* multicast_ptr, void* const* peer_ptrs,
                               int world, void* const* flag_ptrs, int my_rank, int* sync_state, vast_stream_t stream);

/* Wait (on the device, one warp) until every rank's push of this step has arrived in this rank's buffer: entry r of
 * `flags` (this rank's own uint32[world] array) >= consumer epoch + 1, then advance the consumer epoch.  Launched as a
 * programmatic dependent of the push kernel, so it spins while that kernel's stores drain; kernels after it on the
 * stream see the gathered rows.  A rank that never arrives trips a ~3 s watchdog (trap) instead of hanging. */
VAST_API int vast_wait_arrivals(const void* flags, int world, int* sync_state, vast_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * OMC / ITC contrastive loss + hard-negative sampling + backward   (model/vast.py:405-440)
 * ---------------------------------------------------------------------------------------- */

VAST_API size_t vast_omc_workspace_bytes(int64_t bs, int64_t n_total, int64_t dim, int need_sample, int need_grad);

/* vast_omc_step flags */
#define VAST_OMC_TWO_PASS 1 /* evaluate the logits twice (row max / sum-exp pass first) instead of once */
#define VAST_OMC_SEPARATE_ROW_STATS 2 /* run the row statistics + hard-negative draw as their own kernel instead of
                                         inside the dQ GEMM's epilogue (same results bit for bit; A/B aid) */
#define VAST_OMC_ASSUME_IN_RANGE 4 /* the caller guarantees that no negative beats its positive by more than
                                      16 ln2 = 11.09 nats of logit (always true for unit-norm features with
                                      contra_temp >= 0.181, since |s| <= 1): the two flag-gated fallback launches
                                      are not enqueued.  If the range is exceeded after all, loss and grad_temp are
                                      NaN (never silently wrong). */
#define VAST_OMC_WORKSPACE_CLEAN 8 /* the workspace was last used by a COMPLETED vast_omc_step / vast_omc_step_local
                                      call of the same shape (every step leaves its flag / ticket block zeroed), so
                                      the step does not clear it again: one launch less.  Never pass it for a fresh
                                      or recycled allocation. */

/* One fused contrastive step on this rank's rows.
 *   pack        [n_total, 2*dim] bf16, row n = (feat_t_all[n] | feat_cond_all[n]) in rank order
 *               (the output of all_gather_into_tensor over vast_pack_pair buffers); the local
 *               rows are rows [row_offset, row_offset + bs)   (row_offset = rank * bs).
 *   contra_temp_dev  optional DEVICE pointer to the temperature (the nn.Parameter itself): when
 *               non-NULL it is read by the kernels and `contra_temp` is ignored -- no host sync.
 *   loss        [1]  f32: (CE_eps(cond2t) + CE_eps(t2cond)) / 2          (vast.py:412-415)
 *   neg_idx     [2, bs] int64 or NULL: [0] = negative TEXT index per row drawn from
 *               softmax(sim_cond2t)+floor, [1] = negative CONDITION index drawn from
 *               softmax(sim_t2cond)+floor, own positive excluded  (vast.py:423-440).
 *               The draw has the distribution of torch.multinomial(w, 1).  torch draws
 *               argmax_j w_j / E_j (E ~ Exp(1)); here the same race runs between 32-column chunks
 *               (a chunk's min of E_j / w_j is Exp(sum of its w_j)), the column inside the winning
 *               chunk is drawn by inverse CDF, and the constant floor is the uniform component of a
 *               two-part mixture; all randomness is Philox4x32-10(seed, offset), restated by
 *               oracle/spec.py::hardneg_hier_sample.
 *   step_counter optional DEVICE uint64: the Philox offset actually used is offset + *step_counter, and the
 *               library adds 1 to *step_counter at the end of the step (stream-ordered).  A step captured
 *               in a CUDA graph therefore draws fresh noise on every replay.
 *   debug_noise [2, bs, n_total] f32 or NULL: caller-supplied Exp(1) variates; the draw is then
 *               literally argmax_j w_j / debug_noise_j (index 0 = cond2t) -- index-exact parity tests
 *               against the reference's own draws.  Implies VAST_OMC_TWO_PASS.
 *   flags       0 (default): logits evaluated once, exponent reference = the positive pair's logit,
 *               automatic on-device fallback to the two-pass form if a probability numerator would leave
 *               the fp16 range; VAST_OMC_TWO_PASS: always two passes.
 *   grad_cond, grad_t [bs, dim] f32, grad_temp [1] f32 (all three or none): d loss / d input
 *               (unit upstream gradient; gathered side carries no gradient, utils/distributed.py:50).
 *   lse         [2, bs] f32 or NULL: natural-log row log-sum-exp of (cond2t, t2cond).
 * The [bs, n_total] logit matrices are never written to HBM. */
VAST_API int vast_omc_step(const void* pack, int64_t bs, int64_t n_total, int64_t dim, int64_t row_offset,
                  float contra_temp, const float* contra_temp_dev, float label_smoothing, float weight_floor,
                  uint64_t seed, uint64_t offset, uint64_t* step_counter, const float* debug_noise, int flags,
                  float* loss, int64_t* neg_idx, float* grad_cond, float* grad_t, float* grad_temp,
                  float* lse, void* workspace, size_t workspace_bytes, vast_stream_t stream);

/* The same step for a single rank (world_size 1: n_total = bs, row_offset 0) straight from the two feature
 * blocks feat_t / feat_cond [bs, dim] (dtype f32 / bf16 / f16, row stride ld_in elements, 16-byte aligned,
 * ld_in % 8 == 0): vast_pack_pair and the step's first kernel run as ONE pass over the features -- with no
 * all-gather between them there is nothing to wait for.  pack_bf16 [bs, 2*dim] is written (the operand of the
 * similarity GEMMs; same contents as vast_pack_pair's output).  All other arguments as vast_omc_step; results
 * agree with vast_pack_pair + vast_omc_step up to the summation order of the target logits (fp32 rounding). */
VAST_API int vast_omc_step_local(const void* feat_t, const void* feat_cond, int dtype, int64_t ld_in, void* pack_bf16,
                        int64_t bs, int64_t dim, float contra_temp, const float* contra_temp_dev,
                        float label_smoothing, float weight_floor, uint64_t seed, uint64_t offset,
                        uint64_t* step_counter, const float* debug_noise, int flags, float* loss, int64_t* neg_idx,
                        float* grad_cond, float* grad_t, float* grad_temp, float* lse, void* workspace,
                        size_t workspace_bytes, vast_stream_t stream);

/* Negative gather + 3-way concat (model/vast.py:432-448):
 *   ids_out  [3bs, L]  = cat(ids_local, ids_local, ids_all[neg_text])          (int64)
 *   mask_out [3bs, L]  = cat(mask_local, mask_local, mask_all[neg_text])       (int64)
 *   cond_out [3bs, S*H] = cat(cond_local, cond_all[neg_cond], cond_local)      (elem_bytes each)
 * neg_text / neg_cond are the [bs] int64 rows of vast_omc_step's neg_idx ([0] / [1]).
 * row_bytes_cond = S*H*elem_bytes must be a multiple of 16. */
VAST_API int vast_gather_rows_concat3(const int64_t* ids_local, const int64_t* mask_local, const int64_t* ids_all,
                             const int64_t* mask_all, int64_t L, const void* cond_local, const void* cond_all,
                             int64_t row_bytes_cond, const int64_t* neg_text, const int64_t* neg_cond,
                             int64_t bs, int64_t n_total, int64_t* ids_out, int64_t* mask_out, void* cond_out,
                             vast_stream_t stream);

/* Negative gather + 3-way concat with the negative condition rows read STRAIGHT FROM THEIR OWNER RANKS over NVLink
 * peer memory (replaces `all_gather_with_grad(condition_feats)` of the whole [N, S, 768] tensor,
 * model/vast.py:422 + utils/distributed.py:12-47, followed by the index of :432-433):
 * cond_peers[world] (HOST array) = every rank's [rows_per_rank, S*H] block in symmetric memory (this rank's own
 * included); global row j lives on rank j / rows_per_rank.  Otherwise as vast_gather_rows_concat3. */
VAST_API int vast_gather_rows_concat3_peer(const int64_t* ids_local, const int64_t* mask_local, const int64_t* ids_all,
                                  const int64_t* mask_all, int64_t L, const void* cond_local, void* const* cond_peers,
                                  int world, int64_t rows_per_rank, int64_t row_bytes_cond, const int64_t* neg_text,
                                  const int64_t* neg_cond, int64_t bs, int64_t* ids_out, int64_t* mask_out,
                                  void* cond_out, vast_stream_t stream);

/* Its backward: out[l] = base_grad[l] (optional) + sum of grad_peers[e / bs][e % bs] over all requests e (ascending:
 * fixed fp32 summation order) with requests[e] == row0 + l.  requests [world * bs] int64 = the all-gathered negative
 * indices of all ranks; grad_peers[world] (HOST array) = every rank's [bs, S*H] gradient block of its fetched rows in
 * symmetric memory; rows of `dtype`, row_bytes % 16 == 0.  The gradients travel once, to the owner only (the
 * reference all-reduces a [W, bs, S, 768] tensor to keep one slice). */
VAST_API size_t vast_pull_row_grads_workspace_bytes(int64_t bs);
VAST_API int vast_pull_row_grads(const int64_t* requests, void* const* grad_peers, int world, int64_t bs, int64_t row0,
                        int64_t row_bytes, int dtype, const void* base_grad, void* out, void* workspace,
                        size_t workspace_bytes, vast_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Retrieval scoring   (evaluation/evaluation_mm.py:223,253-380)
 * ---------------------------------------------------------------------------------------- */

typedef enum {
  VAST_SIM_BF16 = 0,    /* bf16 inputs, fp32 accumulate on tensor cores                         */
  VAST_SIM_FP32X3 = 1,  /* fp32 inputs split into 3 bf16 terms (6 tensor-core products): fp32-  */
                        /* grade scores (dense fp32-grade score matrix, evaluation_mm.py:223)    */
  VAST_SIM_FP32X2 = 2   /* 2 bf16 terms, 3 products (error <= 3 * 2^-18 |q||k| + accumulation): */
                        /* half the tensor work of FP32X3; the shortlist of the exact pipeline   */
} vast_sim_mode;

/* Build the 16-bit tensor-core operand of `x` [rows, dim] (f32 or bf16 input).
 * mode BF16:   out [rows, dim] bf16.
 * mode FP32X3: out [rows, 6*dim] bf16; as_query != 0 lays the three split terms out as
 *              (hi, hi, mid, mid, hi, lo), otherwise as (hi, mid, hi, mid, lo, hi), so that the dot
 *              product of a query row and a key row sums the six leading cross terms.
 * mode FP32X2: out [rows, 3*dim] bf16; (hi, hi, mid) for queries, (hi, mid, hi) for keys. */
VAST_API int64_t vast_sim_operand_cols(int64_t dim, int mode);
VAST_API int vast_sim_pack_operand(const void* x, int x_dtype, int64_t rows, int64_t dim, int64_t ldx, int mode,
                          int as_query, void* out, vast_stream_t stream);

VAST_API size_t vast_sim_topk_workspace_bytes(int64_t n_q, int64_t n_k, int64_t cols, int64_t k);

/* Streaming similarity + top-k: for every query row i the k best key rows by
 * (score descending, index ascending) of  q_op[i] . k_op[j], never materialising [n_q, n_k]
 * (replaces evaluation_mm.py:223 + :257/:259 `score.topk(k)`).  q_op / k_op are packed operands
 * ([n_q, cols], [n_k, cols] bf16).  Key indices are reported as col_offset + j (column shards).
 * out_keys [n_q, k] uint64 sortable keys (see vast_topk_unpack); unused tail entries are 0. */
VAST_API int vast_sim_topk(const void* q_op, const void* k_op, int64_t n_q, int64_t n_k, int64_t cols, int64_t k,
                  int64_t col_offset, uint64_t* out_keys, void* workspace, size_t workspace_bytes,
                  vast_stream_t stream);

/* vast_sim_topk with per-row bounds: bounds_in [n_q] (or NULL) are orderable bits (the high word of a key) of a
 * PROVEN lower bound of each row's final k-th score over ALL columns of the problem (0 = none), e.g. the k-th
 * score of a scan of a column sample; only scores at or above the bound are ever listed, so a column shard that
 * starts warm does k instead of k (1 + ln(n / k)) list insertions per row.  A row's list may then hold fewer than
 * k keys (the rest 0); the merge over all shards is still the exact global top-k.  bounds_out [n_q] (or NULL)
 * receives the bounds proven by this call (max of bounds_in and the row's k-th listed score).  Workspace as for
 * vast_sim_topk. */
VAST_API int vast_sim_topk_bounded(const void* q_op, const void* k_op, int64_t n_q, int64_t n_k, int64_t cols, int64_t k,
                          int64_t col_offset, const uint32_t* bounds_in, uint32_t* bounds_out, uint64_t* out_keys,
                          void* workspace, size_t workspace_bytes, vast_stream_t stream);

/* k-way merge of `parts` candidate lists per row ([parts, n_q, k_in] keys, e.g. the all-gather
 * of every rank's vast_sim_topk output) into the global top-k_out (score desc, index asc). */
VAST_API int vast_topk_merge(const uint64_t* keys_in, int64_t parts, int64_t n_q, int64_t k_in, int64_t k_out,
                    uint64_t* keys_out, vast_stream_t stream);

/* keys -> (values f32, indices i32); empty slots give value -inf, index -1. */
VAST_API int vast_topk_unpack(const uint64_t* keys, int64_t count, float* values, int32_t* indices, vast_stream_t stream);

/* Exact re-scoring of candidate lists: score64[i, c] = sum_k (double)q[i,k] * (double)kk[idx[i,c],k]
 * in a fixed summation order (lane-strided partials, xor-butterfly), then each row is sorted by
 * (score desc, index asc).  idx entries < 0 are ignored (sorted last).  q, kk are fp32.
 * key_offset is subtracted from idx to address `kk` (column shards). */
VAST_API int vast_rescore_f64(const float* q, int64_t ldq, const float* kk, int64_t ldk, int64_t n_q, int64_t dim,
                     int32_t* idx /* in/out [n_q, k] */, int64_t k, int64_t key_offset,
                     double* score64 /* out [n_q, k] */, vast_stream_t stream);

/* Brute-force exact fp64 top-k for the listed rows (fallback of the exact pipeline when the
 * shortlist cannot be proven complete).  rows_list [n_rows] int32. */
VAST_API int vast_exact_topk_rows(const float* q, int64_t ldq, const float* kk, int64_t ldk, int64_t n_k, int64_t dim,
                         const int32_t* rows_list, int64_t n_rows, int64_t k, int64_t col_offset,
                         int32_t* idx_out /* [n_q, k] rows addressed by rows_list */, double* score_out,
                         vast_stream_t stream);

/* Streaming rank of the ground truth, no [n_q, n_k] matrix (evaluation_mm.py:333-338 `indice_matrix[i].index(gt)`
 * and :355-364 without the sort / .tolist()):
 *   rank_out[i] = #{ j in [col_lo, col_lo + n_k) : s_ij > s_i,gt  or  (s_ij == s_i,gt and j < gt_col[i]) }
 * with s the EXACT similarities of the fp32 features (fp64 accumulation in the summation order of
 * vast_rescore_f64).  q [n_q, dim], kk [n_k_total, dim] fp32 (the FULL key matrix: the ground truth of a row may
 * lie outside the shard); q_op [n_q, cols] / k_op [n_k, cols] are packed operands (vast_sim_pack_operand, mode
 * FP32X2 or FP32X3) of the queries and of key rows [col_lo, col_lo + n_k).  The tensor cores count the columns
 * surely above the ground truth; columns within delta_rel * |q_i| * max_j |k_j| of it are re-scored in fp64
 * (rows with more than 32 of them are recounted by brute force), so the result is exact and independent of tile
 * shapes and GPU count; delta_rel = 2^-11 bounds the split + fp32-accumulation error of both operand modes.
 * Column shards: the ranks of disjoint shards add up (all-reduce sum) to the rank over all columns.
 * gt_col [n_q] int32 global column index (gt_col < 0: every column counts). */
VAST_API size_t vast_rank_of_gt_workspace_bytes(int64_t n_q, int64_t n_k, int64_t cols);
VAST_API int vast_rank_of_gt(const float* q, int64_t ldq, const float* kk, int64_t ldk, int64_t n_q, int64_t n_k_total,
                    int64_t dim, const void* q_op, const void* k_op, int64_t cols, int64_t col_lo, int64_t n_k,
                    const int32_t* gt_col, float delta_rel, int32_t* rank_out, void* workspace,
                    size_t workspace_bytes, vast_stream_t stream);

/* top-k of an already materialised fp32 score matrix along rows (axis=1) or columns (axis=0),
 * ties broken by lower index -- the drop-in for `score_matrix.topk(k, dim)` at
 * evaluation_mm.py:257/:259 when the caller hands refine_score_matrix a dense matrix.
 * idx_out: axis=1 -> [n_rows, k]; axis=0 -> [k, n_cols]  (int32). */
VAST_API int vast_dense_topk(const float* score, int64_t n_rows, int64_t n_cols, int64_t ld, int64_t k, int axis,
                    int32_t* idx_out, float* val_out, vast_stream_t stream);

/* Rank of entry (gt_row[g], gt_col[g]) in a stable descending sort of its row (axis=1, over the
 * columns) or of its column (axis=0, over the rows) of an fp32 matrix, lower index first on ties
 * (evaluation_mm.py:333-338 and :355-365: sort + list.index). */
VAST_API int vast_dense_rank_of_gt(const float* score, int64_t n_rows, int64_t n_cols, int64_t ld, int axis,
                          const int32_t* gt_row, const int32_t* gt_col, int64_t n_gt, int32_t* rank_out,
                          vast_stream_t stream);

/* Candidate bookkeeping for the ITM re-rank (evaluation_mm.py:264-314 without the dense mask):
 * bucket the (text, video) candidate pairs by video.  text_idx/video_idx [n_pairs] int32
 * (video_idx < 0 = empty slot).  Outputs: counts/offsets [n_videos+1] (CSR), order [n_pairs]
 * = text indices grouped by video, ascending inside a video. */
VAST_API size_t vast_bucket_by_video_workspace_bytes(int64_t n_pairs, int64_t n_videos);
VAST_API int vast_bucket_by_video(const int32_t* text_idx, const int32_t* video_idx, int64_t n_pairs, int64_t n_videos,
                         int32_t* offsets, int32_t* texts_sorted, void* workspace, size_t workspace_bytes,
                         vast_stream_t stream);

/* out[text, video] = score for every pair (evaluation_mm.py:313); out is pre-zeroed by the caller. */
VAST_API int vast_scatter_scores(const int32_t* text_idx, const int32_t* video_idx, const float* scores, int64_t n_pairs,
                        float* out, int64_t ld, vast_stream_t stream);

/* Match_head + softmax[:, 1] (model/general_module.py:34-42 Linear -> GELU(erf) -> LayerNorm -> Linear(2);
 * model/vast.py:378 `F.softmax(self.itm_head(cls), dim=1)[:, 1]`) as one tensor-core GEMM with a fused epilogue:
 * cls_op [rows, cols], w1_op [hidden, cols] packed 16-bit operands of the cls tokens and of linear1.weight; b1 =
 * linear1.bias; u_c = linear2.weight[c] * layernorm.weight (c = 0, 1), sum_u_c their sums, v_c = linear2.weight[c] .
 * layernorm.bias + linear2.bias[c]; eps the LayerNorm epsilon.  LayerNorm followed by a 2-row Linear needs only four
 * sums per row, so the hidden activations never leave the accumulator.  score [rows] = P(match); logits [rows, 2]
 * optional. */
VAST_API int vast_match_head(const void* cls_op, const void* w1_op, int64_t rows, int64_t hidden, int64_t cols,
                    const float* b1, const float* u0, const float* u1, float sum_u0, float sum_u1, float v0, float v1,
                    float eps, float* score, float* logits, vast_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Generic tensor-core NT GEMM (unit-test / building block): C[m, n] = alpha * sum_k A[m,k] B[n,k]
 * A [M, K], B [N, K] 16-bit (dtype bf16 or f16) with leading dims lda/ldb (multiples of 8), C f32.
 * ---------------------------------------------------------------------------------------- */
VAST_API size_t vast_gemm_nt_workspace_bytes(int64_t M, int64_t N, int64_t K);
VAST_API int vast_gemm_nt(const void* A, int64_t lda, const void* B, int64_t ldb, int dtype, int64_t M, int64_t N,
                 int64_t K, float alpha, float* C, int64_t ldc, void* workspace, size_t workspace_bytes,
                 vast_stream_t stream);
/* C[m, n] = alpha * sum_k A[m, k] * B[k, n]: B row-major [K, N] (ldb >= N) is fed to the tensor cores as
 * the MN-major operand straight from its row-major layout (no transposed copy); dtype_a must equal
 * dtype_b (VAST_BF16 or VAST_F16).  Workspace as for vast_gemm_nt. */
VAST_API int vast_gemm_nn(const void* A, int64_t lda, int dtype_a, const void* B, int64_t ldb, int dtype_b, int64_t M,
                 int64_t N, int64_t K, float alpha, float* C, int64_t ldc, void* workspace, size_t workspace_bytes,
                 vast_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VAST_B200_H_ */

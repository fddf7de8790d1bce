/* c2dsr_b200 -- C ABI of the B200-native C2DSR hot path (libc2dsr_b200.so, sm_100a only).
 *
 * The reference (crystal22/C2DSR) is pure PyTorch and has no FFI; its "operator interface" for
 * this path is the set of torch calls listed below.  Each entry point names the reference call
 * it replaces (file:line in the upstream repo).  INTEGRATION.md shows the ctypes stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; the caller owns all
 *     buffers, including scratch (sizes from the *_workspace_bytes functions);
 *   - `stream` is a cudaStream_t passed as void*; no entry point synchronises, allocates or
 *     keeps global mutable state, so all of them may be captured in a CUDA graph;
 *   - return value 0 = OK, negative = error (-(cudaError_t) or a C2DSR_ERR_* code);
 *     c2dsr_last_error() gives a thread-local message;
 *   - matrices are row-major fp32, indices int64 (the reference uses LongTensor), CSR arrays int32;
 *   - dropout: `p` drop probability, `seed`/`tag` select a counter-based mask that the matching
 *     backward call regenerates from the same (seed, tag); p = 0 disables it;
 *   - there is no CPU fallback: on a device that is not compute capability 10.x every compute
 *     entry returns C2DSR_ERR_ARCH.
 */
#ifndef C2DSR_B200_H
#define C2DSR_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define C2DSR_ABI_VERSION 1

int c2dsr_abi_version(void);
const char* c2dsr_last_error(void);
/* 0 when the current device is a Blackwell sm_100 part, else C2DSR_ERR_ARCH. */
int c2dsr_device_check(void);
/* Number of CUDA kernels this library has launched in this process (for benchmark accounting). */
int64_t c2dsr_launch_count(void);

/* ---- K1: branch input = embedding gather --------------------------------------------------
 * x[t,:] = drop( scale * (hi[seq[t],:] + E[seq[t],:]) + P[pos[t],:] )
 * replaces F.embedding(seq, hi) + embed_i(seq), "*= sqrt(d)" (models/C2DSR.py:65-71,81-82) and
 * "seq_enc += pos_emb(pos); dropout" (models/encoders.py:30-31). */
int c2dsr_gather_fwd(const float* hi, const float* E, const float* P, const int64_t* seq, const int64_t* pos,
                     float* x, int64_t n_tok, int d, float scale, float p, uint64_t seed, uint64_t tag,
                     void* stream);
/* Backward of the above, deterministic (first-reference election with integer atomics + ordered sums,
 * no float atomics; n_rows = rows of the item tables, len_max = rows of P):
 *   g[t] = dx[t] * mask;  d_P[pos[t]] += g[t];  S[n] = scale * sum_{t: seq[t]=n} g[t];
 *   d_hi[n] += S[n];  d_E[n] += S[n] for n != pad_idx  (nn.Embedding padding_idx, C2DSR.py:20).
 * replaces embedding_dense_backward reached from loss.backward() (trainer.py:156). */
/* Evaluation: x[b] = the branch input of token (b, sel[b]) for b < n_seq, and x[n_seq] = that of one PAD token at
 * position 0 (what the pad-key shortcut needs); x is [n_seq + 1, d].  No dropout. */
int c2dsr_gather_select_fwd(const float* hi, const float* E, const float* P, const int64_t* seq, const int64_t* pos,
                            const int64_t* sel, float* x, int64_t n_seq, int L, int d, float scale, int64_t pad_idx,
                            void* stream);
int64_t c2dsr_gather_bwd_workspace_bytes(int64_t n_tok, int d, int64_t n_rows, int len_max);
int c2dsr_gather_bwd(const float* dx, const int64_t* seq, const int64_t* pos, float* d_hi, float* d_E, float* d_P,
                     int64_t n_tok, int d, int64_t n_rows, int len_max, int64_t pad_idx, float scale, float p,
                     uint64_t seed, uint64_t tag, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- K2: CSR SpMM for GCN propagation -----------------------------------------------------
 * drop_mode 0: out = alpha * A X + beta * Y + gamma * Z
 * drop_mode 1: out = alpha * A (m .* X) + ...      (mask indexed by the gathered row: forward)
 * drop_mode 2: out = alpha * m .* (A X) + ...      (mask indexed by the output row: backward, A = A^T)
 * Y, Z may be NULL; out may alias Y or Z (not X).  long_rows (optional, device) lists the rows that are
 * split across a CTA to bound the heavy tail.  replaces torch.spmm(adj, h) + stack/mean
 * (models/encoders.py:43-48) and its autograd transpose product.
 * Optional byte masks (NULL = none), both exact:  out_row_needed [n_rows]: rows with 0 are not computed and
 * are left zero (a training step only reads the rows of the items in its batch: hi[seq]);  x_row_nonzero
 * [columns of A]: rows of X flagged 0 are known to be all zero and their entries are skipped (the gradient
 * d_hi is non-zero only on the rows the batch touched).  c2dsr_mark_rows builds such a mask from item ids. */
int c2dsr_spmm(const int32_t* rowptr, const int32_t* col, const float* val, const int32_t* long_rows, int n_long,
               const float* X, const float* Y, const float* Z, float* out, int64_t n_rows, int d, float alpha,
               float beta, float gamma, int drop_mode, float p, uint64_t seed, uint64_t tag,
               const uint8_t* out_row_needed, const uint8_t* x_row_nonzero, void* stream);
/* mask[k] = 1 for every k in ids[0..n) (0 <= k < n_rows), 0 elsewhere */
int c2dsr_mark_rows(const int64_t* ids, int64_t n, int64_t n_rows, uint8_t* mask, void* stream);
/* Rows with more non-zeros than this should be listed in long_rows (they get a whole CTA each). */
int c2dsr_spmm_long_row_threshold(void);

/* ---- dense building blocks ----------------------------------------------------------------
 * C[M,N] = drop(act( alpha * op(A) op(B) + bias[N] )) + beta * C
 *   ta = 0: A is [M,K] (lda)   ta = 1: A is stored [K,M] (lda)
 *   tb = 0: B is [K,N] (ldb)   tb = 1: B is stored [N,K] (ldb)  -- the nn.Linear weight layout
 * act 0 = none, 1 = relu.  fp32 FFMA path (exact fp32 products); workspace enables split-K.
 * replaces nn.Linear / torch.mm on the path (encoder projections, nn.Bilinear, classifiers). */
int64_t c2dsr_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K);
int c2dsr_gemm(int ta, int tb, int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t lda,
               const float* B, int64_t ldb, float beta, float* C, int64_t ldc, const float* bias, int act,
               float p, uint64_t seed, uint64_t tag, void* workspace, int64_t workspace_bytes, void* stream);
/* Same product on tensor cores (tcgen05 + TMA + tensor memory): operands are split into bf16 hi / lo
 * (passes = 3: hi*hi + hi*lo + lo*hi, fp32-grade; passes = 1: hi only) and accumulated in fp32.
 * Transposed operands (ta = 1, tb = 0) are read MN-major from their row-major storage.  alpha is 1,
 * beta must be 0 or 1.  Short grids with a long K are cut into K slabs added in a fixed order. */
int64_t c2dsr_gemm_tc_workspace_bytes(int64_t M, int64_t N, int64_t K);
int c2dsr_gemm_tc(int ta, int tb, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B,
                  int64_t ldb, float beta, float* C, int64_t ldc, const float* bias, int act, float p, uint64_t seed,
                  uint64_t tag, int passes, void* workspace, int64_t workspace_bytes, void* stream);
/* out[N] (+)= sum over rows of X[M,N] (ldx), fixed summation order.  With a workspace of at least
 * ceil(M/128) * N floats the reduction runs in two phases over a 2-D grid. */
int c2dsr_colsum(const float* X, int64_t ldx, int64_t M, int64_t N, float* out, int accumulate, void* workspace,
                 int64_t workspace_bytes, void* stream);
/* out[0] = sum_i w[i] * x[i] (w may be NULL), single fixed-order reduction. */
int c2dsr_wsum(const float* x, const float* w, int64_t n, float* out, void* stream);

/* ---- K3: SASRec-style encoder -------------------------------------------------------------
 * replaces SelfAttention.forward -> nn.TransformerEncoder (models/encoders.py:23-33):
 * n_layers x { MHA with allow[i,j] = (j <= i) && (seq[j] == pad), out-proj, LayerNorm(eps),
 * FFN d->d (ReLU) ->d, LayerNorm }, then a final LayerNorm.  norm_first selects the pre-norm
 * layer.  Rows with no allowed key attend to nothing (output 0). */
typedef struct {
    const float *in_proj_w, *in_proj_b;   /* [3d,d], [3d] */
    const float *out_proj_w, *out_proj_b; /* [d,d], [d]   */
    const float *lin1_w, *lin1_b;         /* [d,d], [d]   */
    const float *lin2_w, *lin2_b;         /* [d,d], [d]   */
    const float *ln1_w, *ln1_b, *ln2_w, *ln2_b; /* [d] each */
} c2dsr_layer_weights;
typedef struct {
    float *in_proj_w, *in_proj_b, *out_proj_w, *out_proj_b, *lin1_w, *lin1_b, *lin2_w, *lin2_b;
    float *ln1_w, *ln1_b, *ln2_w, *ln2_b;
} c2dsr_layer_grads;

/* floats the forward saves for the backward (`saved`), and scratch bytes for either pass.
 * dense_passes selects the kernels of the dense layers: 0 = fp32 FFMA, 3 / 1 = tcgen05 (bf16 hi/lo split / bf16). */
int64_t c2dsr_encoder_saved_floats(int64_t n_tok, int d, int n_head, int n_layers);
int64_t c2dsr_encoder_workspace_bytes(int64_t n_tok, int d, int n_head, int dense_passes);
int c2dsr_encoder_fwd(const c2dsr_layer_weights* layers_host, int n_layers, const float* lnf_w, const float* lnf_b,
                      const float* x, const int64_t* seq, int64_t n_seq, int L, int d, int n_head, int64_t pad_idx,
                      int norm_first, int dense_passes, float eps, float p, uint64_t seed, uint64_t tag, float* out,
                      float* saved, void* workspace, int64_t workspace_bytes, void* stream);
/* Evaluation form for a ONE-layer encoder: only out[b, :] = encoder(x)[b, sel[b], :] is produced (the position
 * the ranking query reads, trainer.py:169-177).  Keys / values are projected for every token; the attention
 * runs for the single query sel[b]; output projection, residuals, LayerNorms and the feed-forward run on n_seq
 * rows instead of n_seq * L.  No dropout, nothing saved.  Same numbers as c2dsr_encoder_fwd + a row gather. */
int64_t c2dsr_encoder_select_workspace_bytes(int64_t n_seq, int L, int d, int dense_passes);
int c2dsr_encoder_fwd_select(const c2dsr_layer_weights* layers_host, int n_layers, const float* lnf_w,
                             const float* lnf_b, const float* x, const int64_t* seq, const int64_t* sel, int64_t n_seq,
                             int L, int d, int n_head, int64_t pad_idx, int norm_first, int dense_passes, float eps,
                             float* out, void* workspace, int64_t workspace_bytes, void* stream);
/* Pad-key shortcut of the single-query forward (evaluation, one encoder layer, no dropout).  Under the reference's
 * inverted key-padding mask (models/encoders.py:33) a query attends only to PAD tokens at or before it,
 * and every PAD token of a branch has the same input row x_pad = sqrt(d) (hi[PAD] + E[PAD]) + P[0] (the preprocessor
 * gives pad tokens position 0; the host checks that once per split).  Identical keys -> uniform soft-max ->
 * the attention output of any query with at least one allowed key is the value row of x_pad; with none it is 0.
 * So: x_sel [n_seq, d] = input rows of the selected tokens, x_pad [d]; seq / sel only decide "has an allowed key".
 * No QKV projection or attention over the n_seq * L tokens is computed at all.
 * c2dsr_encoder_padkeys_prepare computes y_pad [d] = out_proj(value_proj(x_pad)) once per propagation (it does not
 * depend on the batch; workspace >= 32 d bytes + 16 MiB); c2dsr_encoder_fwd_padkeys is the per-batch part. */
int64_t c2dsr_encoder_padkeys_workspace_bytes(int64_t n_seq, int d, int dense_passes);
int c2dsr_encoder_padkeys_prepare(const c2dsr_layer_weights* layers, int n_layers, const float* x_pad, int d,
                                  int norm_first, float eps, float* y_pad, void* workspace, int64_t workspace_bytes,
                                  void* stream);
int c2dsr_encoder_fwd_padkeys(const c2dsr_layer_weights* layers, int n_layers, const float* lnf_w, const float* lnf_b,
                              const float* x_sel, const float* y_pad, const int64_t* seq, const int64_t* sel,
                              int64_t n_seq, int L, int d, int n_head, int64_t pad_idx, int norm_first,
                              int dense_passes, float eps, float* out, void* workspace, int64_t workspace_bytes,
                              void* stream);
/* Gradients are ACCUMULATED (+=) into `grads` and lnf grads; dx is overwritten. */
int c2dsr_encoder_bwd(const c2dsr_layer_weights* layers_host, const c2dsr_layer_grads* grads_host, int n_layers,
                      const float* lnf_w, float* d_lnf_w, float* d_lnf_b, const float* d_out, const int64_t* seq,
                      int64_t n_seq, int L, int d, int n_head, int64_t pad_idx, int norm_first, int dense_passes,
                      float eps, float p, uint64_t seed, uint64_t tag, const float* saved, float* dx, void* workspace,
                      int64_t workspace_bytes, void* stream);

/* primitives of the encoder, exported for unit tests */
int c2dsr_attention_fwd(const float* qkv, const int64_t* seq, int64_t n_seq, int L, int d, int n_head,
                        int64_t pad_idx, float p, uint64_t seed, uint64_t tag, float* o, float* lse, void* stream);
int c2dsr_attention_bwd(const float* qkv, const float* o, const float* lse, const float* d_o, const int64_t* seq,
                        int64_t n_seq, int L, int d, int n_head, int64_t pad_idx, float p, uint64_t seed,
                        uint64_t tag, float* d_qkv, void* stream);
/* s = x + drop(y) (y may be NULL);  out = do_ln ? LayerNorm(s) * w + b : s;  stats[t] = (mean, rstd) */
int c2dsr_add_ln_fwd(const float* x, const float* y, const float* w, const float* b, float* s_out, float* out,
                     float* stats, int64_t n_tok, int d, int do_ln, float eps, float p, uint64_t seed, uint64_t tag,
                     void* stream);
/* ds = LayerNorm backward of d_out at (s, stats, w);  dx_out = (accumulate ? dx_out : 0) + ds */
int c2dsr_ln_bwd(const float* d_out, const float* s, const float* stats, const float* w, float* dx_out,
                 int accumulate, int64_t n_tok, int d, void* stream);
/* d_w[d] += sum_t d_out * xhat,  d_b[d] += sum_t d_out */
int c2dsr_ln_param_grad(const float* d_out, const float* s, const float* stats, float* d_w, float* d_b,
                        int64_t n_tok, int d, void* stream);

/* ---- K5: infomax discriminator ------------------------------------------------------------
 * replaces Trainer.cal_mask + masked pooling + nn.Bilinear + BCE-with-logits x4
 * (trainer.py:85-119, models/C2DSR.py:46-55), including the crossed masks.
 * pooled[6][B][d] = {x_mean, y_mean, share.wb, share.wa, neg_a.wa, neg_b.wb};
 * loss[0] = sum of the four batch-mean BCE terms (batch mean over inv_batch = 1/B_global). */
int64_t c2dsr_infomax_workspace_bytes(int64_t B, int d);
int c2dsr_infomax_fwd(const float* h_share, const float* hx, const float* hy, const float* h_neg_a,
                      const float* h_neg_b, const int64_t* gt_mask_a, const int64_t* gt_mask_b, const float* W_a,
                      const float* W_b, const float* bias_a, const float* bias_b, int64_t B, int L, int d,
                      float inv_batch, float* pooled, float* U, float* sims, float* loss, void* workspace,
                      int64_t workspace_bytes, void* stream);
/* upstream: d_loss (device scalar).  d_h_* are ACCUMULATED (+=); dW/dbias accumulated. */
int c2dsr_infomax_bwd(const float* d_loss, const float* pooled, const float* U, const float* sims,
                      const int64_t* gt_mask_a, const int64_t* gt_mask_b, const float* W_a, const float* W_b,
                      int64_t B, int L, int d, float inv_batch, float* d_h_share, float* d_hx, float* d_hy,
                      float* d_h_neg_a, float* d_h_neg_b, float* dW_a, float* dW_b, float* dbias_a, float* dbias_b,
                      void* workspace, int64_t workspace_bytes, void* stream);

/* ---- K4a: classifier logits + cross-entropy (training) ------------------------------------
 * For M rows H[M,d] against one domain's classifier W[N,d], b[N] plus the pad logit zpad[M]:
 *   lse[m] = logsumexp([H W^T + b | zpad]);  loss_row[m] = gt[m] == N ? 0 : lse[m] - z[m, gt[m]]
 * replaces classifier_x(h) + cat(classifier_pad) + F.cross_entropy(ignore_index=N)
 * (trainer.py:131-152).  Z[M, ldz] (ldz = c2dsr_score_ldz(N)) is scratch kept for the backward. */
int64_t c2dsr_score_ldz(int64_t N);
int c2dsr_score_ce_fwd(const float* H, const float* W, const float* bias, const float* zpad, const int64_t* gt,
                       int64_t M, int64_t N, int d, float* Z, float* lse, float* loss_row, void* workspace,
                       int64_t workspace_bytes, void* stream);
/* coef[m] = upstream * row weight.  Overwrites Z with dZ, then
 * dH[M,d] = dZ W;  dW[N,d] += dZ^T H;  dbias[N] += colsum(dZ);  dzpad[m] = (exp(zpad-lse)) * coef. */
int c2dsr_score_ce_bwd(const float* H, const float* W, const float* zpad, const int64_t* gt, const float* lse,
                       const float* coef, int64_t M, int64_t N, int d, float* Z, float* dH, float* dW,
                       float* dbias, float* dzpad, void* workspace, int64_t workspace_bytes, void* stream);

/* Rows whose target is the ignore class contribute neither loss nor gradient (F.cross_entropy ignore_index,
 * trainer.py:143-152).  perm[M] = stable partition of 0..M-1 with the non-ignored rows first, so that the caller
 * can run the calls above on the leading rows only. */
int c2dsr_compact_rows(const int64_t* gt, int64_t M, int64_t ignore, int64_t* perm, void* stream);

/* Tensor-core form of the two calls above (tcgen05 + TMA, bf16 hi/lo split with `passes` = 3 for
 * fp32-grade logits, 1 for plain bf16).  The fp32 logits are never written to HBM: the forward keeps
 * per-tile (max, sum-exp) pairs, the backward recomputes the logits, stores dZ as bf16 hi/lo and runs
 * dH = dZ W and dW += dZ^T H as two more tcgen05 GEMMs.  Both calls split their operands into
 * `workspace`; W_hi / W_lo (optional, NULL = split here): the bf16 hi / lo split of W [N, d] made once by
 * c2dsr_split_bf16 with ld_out = d, so that forward and backward of a step share it.  backward = 1 sizes the
 * workspace for the backward call. */
int64_t c2dsr_score_ce_tc_workspace_bytes(int64_t M, int64_t N, int d, int backward);
int c2dsr_score_ce_fwd_tc(const float* H, const float* W, const uint16_t* W_hi, const uint16_t* W_lo, const float* bias,
                          const float* zpad, const int64_t* gt, int64_t M, int64_t N, int d, int passes, float* lse,
                          float* loss_row, void* workspace, int64_t workspace_bytes, void* stream);
int c2dsr_score_ce_bwd_tc(const float* H, const float* W, const uint16_t* W_hi, const uint16_t* W_lo, const float* bias,
                          const float* zpad, const int64_t* gt, const float* lse, const float* coef, int64_t M,
                          int64_t N, int d, int passes, float* dH, float* dW, float* dbias, float* dzpad,
                          void* workspace, int64_t workspace_bytes, void* stream);

/* ---- K4b: full-catalogue scoring + rank count (evaluation) ---------------------------------
 * replaces the per-sample loop of Trainer.evaluate_batch (trainer.py:168-179):
 *   s = W q + b;  rank = 1 + #{ j in candidates : s_j > s_gt }   (strict '>', fp32)
 * Catalogue shard = items [n0, n1) of the domain (W, bias point at the shard's first row).
 * Step 1 (scores):  S[Q, lds] = q W_shard^T + b_shard              -> c2dsr_score_shard
 * Step 2 (target):  s_gt[i] = S[i, gt[i]-n0] if gt[i] in shard else 0 (sum across shards)
 * Step 3 (count):   counts[i] += #{ j in shard, j != gt[i], (neg == NULL or j in neg[i]) : S[i,j] > s_gt[i] }
 * rank = 1 + sum over shards of counts (integer all-reduce). */
int c2dsr_score_shard(const float* Q, const float* W, const float* bias, int64_t n_q, int64_t n_shard, int d,
                      float* S, int64_t lds, void* workspace, int64_t workspace_bytes, void* stream);
int c2dsr_pick_target(const float* S, int64_t lds, const int64_t* gt, int64_t n_q, int64_t n0, int64_t n1,
                      float* s_gt, void* stream);
/* bit-exact counting contract: integer result from given fp32 scores.  neg: [n_q, n_neg] domain-local
 * ids or NULL (full-catalogue mode). */
int c2dsr_rank_from_scores(const float* S, int64_t lds, const float* s_gt, const int64_t* gt, const int64_t* neg,
                           int64_t n_neg, int64_t n_q, int64_t n0, int64_t n1, int32_t* counts, void* stream);
/* Fused tensor-core path: tcgen05 GEMM (bf16x3 split, fp32 accumulate in TMEM) with the count
 * done in the epilogue; the score matrix never reaches HBM.  Two launches: target scores, then
 * the counting GEMM.  W_hi/W_lo, Q_hi/Q_lo are the bf16 split produced by c2dsr_split_bf16. */
int c2dsr_split_bf16(const float* X, int64_t rows, int d, int64_t ld_out, uint16_t* hi, uint16_t* lo,
                     void* stream);
int64_t c2dsr_score_tc_workspace_bytes(int64_t n_q, int64_t n_shard, int d);
/* n_q_limit (optional, device pointer): only the first min(n_q, *n_q_limit) query rows are computed; n_q is then the
 * capacity the buffers were sized for.  Lets one captured launch serve a per-batch row count known only on the
 * device (the number of queries of a domain, see c2dsr_eval_partition). */
int c2dsr_score_target_tc(const uint16_t* Q_hi, const uint16_t* Q_lo, const uint16_t* W_hi, const uint16_t* W_lo,
                          const float* bias, const int64_t* gt, int64_t n_q, int64_t n0, int64_t n1, int d,
                          int passes, const int* n_q_limit, float* s_gt, void* workspace, int64_t workspace_bytes,
                          void* stream);
/* Target scores of ALL queries from the full fp32 classifier (replicated on every rank): same arithmetic and bits as
 * c2dsr_score_target_tc on the owning shard, without the exchange between catalogue shards. */
int c2dsr_score_target_full_tc(const uint16_t* Q_hi, const uint16_t* Q_lo, const float* W, const float* bias,
                               const int64_t* gt, int64_t n_q, int64_t n_items, int d, int passes, const int* n_q_limit,
                               float* s_gt, void* workspace, int64_t workspace_bytes, void* stream);
int c2dsr_score_count_tc(const uint16_t* Q_hi, const uint16_t* Q_lo, const uint16_t* W_hi, const uint16_t* W_lo,
                         const float* bias, const float* s_gt, const int64_t* gt, int64_t n_q, int64_t n0,
                         int64_t n1, int d, int passes, const int* n_q_limit, int32_t* counts, float* S_debug,
                         int64_t lds, void* workspace, int64_t workspace_bytes, void* stream);
/* Evaluation batch on the device, no host round trip (trainer.py:168-179: the per-sample `if xory_last == 0` branch):
 * stable partition of the B query vectors by domain.  Queries with dom == 0 go, in batch order, to the front of
 * (QA_hi, QA_lo, gtA), the others to (QB_hi, QB_lo, gtB), already split into bf16 hi / lo (lo may be NULL for
 * passes == 1); slot[i] = position of query i inside its domain's buffer; n_ab[0 / 1] = queries per domain.  All
 * buffers have capacity B rows.  workspace: 4 * B + 64 bytes. */
int c2dsr_eval_partition(const float* q, const int64_t* dom, const int64_t* gt, int64_t B, int d, uint16_t* QA_hi,
                         uint16_t* QA_lo, uint16_t* QB_hi, uint16_t* QB_lo, int64_t* gtA, int64_t* gtB, int32_t* slot,
                         int32_t* n_ab, void* stream);
/* out[i] = (1 + counts of query i in its domain's buffer, dom[i] != 0) as two int32 [B] planes: ranks, domain. */
int c2dsr_eval_ranks(const int32_t* countsA, const int32_t* countsB, const int32_t* slot, const int64_t* dom, int64_t B,
                     int32_t* out, void* stream);

/* ---- optimiser -----------------------------------------------------------------------------
 * AdamW with amsgrad on an accumulated gradient (trainer.py:21-22,42,157-158):
 *   acc += g (if g != NULL);  p *= 1 - lr*wd;  m,v,vmax updated from acc;  p -= lr/bc1 * m / (sqrt(vmax)/sqrt(bc2) + eps)
 * One launch over a table of n tensors (device arrays of pointers / sizes). */
typedef struct {
    float* p; const float* g; float* acc; float* m; float* v; float* vmax; int64_t n;
} c2dsr_adam_tensor;
int c2dsr_adamw_amsgrad(const c2dsr_adam_tensor* table_dev, int n_tensors, int64_t max_n, float lr, float beta1,
                        float beta2, float eps, float weight_decay, int step, void* stream);

/* ---- K4a row assembly (trainer.py:122-152: slices of the last len_rec positions, cat, add, ignore rows) ----
 * The 2 B R virtual loss rows of one domain -- B R "share" rows (H = Hpad = h_share[b, l], target gt_share) then
 * B R "domain" rows (H = h_share + h_dom, Hpad = h_dom, target gt_dom), (b, l) over the last R positions -- are
 * stably partitioned (rows whose target != ignore first) and the first M are emitted: H [M, d], gt [M],
 * w [M] (*w_share for share rows, 1 / *n_dom for domain rows; device scalars) and the pad logit
 * zpad [M] = Hpad . wpad + bpad.  perm [2BR] / inv [2BR] (inv[perm[r]] = r) are kept for the backward, which
 * writes d h_share and d h_dom for EVERY token ([B, L, d], zeros outside the last R positions) from dH [M, d]
 * and dzpad [M] (d Hpad = dzpad wpad^T is never materialised) and d wpad [d], d bpad [1]. */
int c2dsr_loss_rows_fwd(const float* h_share, const float* h_dom, const int64_t* gt_share, const int64_t* gt_dom,
                        int64_t B, int L, int R, int d, int64_t ignore, int64_t M, const float* w_share,
                        const float* n_dom, const float* wpad, const float* bpad, int64_t* perm, int32_t* inv, float* H,
                        int64_t* gt, float* w, float* zpad, void* stream);
int64_t c2dsr_loss_rows_bwd_workspace_bytes(int64_t M, int d);
int c2dsr_loss_rows_bwd(const float* dH, const float* dzpad, const float* h_share, const float* h_dom, const float* wpad,
                        const int64_t* perm, const int32_t* inv, int64_t B, int L, int R, int d, int64_t M,
                        float* d_h_share, float* d_h_dom, float* dwpad, float* dbpad, void* workspace,
                        int64_t workspace_bytes, void* stream);

/* ---- adjacency builder (utils/graph.py:33-96: preprocess_graph + normalize) -------------------
 * Raw directed transitions (src[i] -> dst[i], duplicates allowed) to the CSR of A = D^-1 (summed counts)
 * (transpose = 0) or of A^T (transpose = 1; values still divided by the row sum of the SOURCE item).
 * rowptr [n_rows + 1], col / val with capacity n_edges, *nnz_out = number of distinct pairs (device int).
 * Rows are column-sorted; val = (1 / rowsum) * count with separately rounded fp32 operations, bit-identical
 * to the reference's normalisation.  Deterministic (LSD radix sort + scans, integer atomics only). */
int64_t c2dsr_graph_build_workspace_bytes(int64_t n_edges, int64_t n_rows);
int c2dsr_graph_build(const int32_t* src, const int32_t* dst, int64_t n_edges, int64_t n_rows, int transpose,
                      int32_t* rowptr, int32_t* col, float* val, int32_t* nnz_out, void* workspace,
                      int64_t workspace_bytes, void* stream);

/* ---- per-step device state (CUDA-graph replay of a whole training step) ------------------------
 * A captured step cannot take fresh scalars from the host, so the three things that change every step
 * live in a small device struct: the step number and learning rate (bias corrections and step size of
 * c2dsr_adamw_amsgrad_dyn) and two dropout key words.  Any entry point that takes (seed, tag) accepts
 * tag | C2DSR_SEED_INDIRECT, meaning ``seed`` is the device address of c2dsr_step_state.key: the mask
 * key is then hash(tag) XOR those words, read by the kernels themselves.
 *   c2dsr_step_begin:  step += 1;  key = hash(step, seed_base)      (call once after every optimiser step)
 * The eager path uses the same entries, so a replayed step and an eager one compute the same thing. */
typedef struct {
    uint64_t step;       /* optimiser step the next c2dsr_adamw_amsgrad_dyn call will apply */
    uint32_t key[2];     /* per-step dropout key words (offset 8) */
    float lr;
    float reserved[3];
} c2dsr_step_state;
#define C2DSR_SEED_INDIRECT (1ull << 63)
#define C2DSR_STEP_KEY_OFFSET 8
int c2dsr_step_state_bytes(void);
int c2dsr_step_state_set(void* state, int64_t step, float lr, void* stream);
int c2dsr_step_state_set_lr(void* state, float lr, void* stream);
int c2dsr_step_begin(void* state, uint64_t seed_base, void* stream);
/* background != 0 (here and in c2dsr_adamw_amsgrad_peer): the launch runs beside higher-priority work -- short-lived
 * CTAs instead of a resident grid-stride grid. */
int c2dsr_adamw_amsgrad_dyn(const c2dsr_adam_tensor* table_dev, int n_tensors, int64_t max_n, const void* state,
                            float beta1, float beta2, float eps, float weight_decay, int background, void* stream);

/* Data-parallel form of the same update over peer memory (NVLink / NVSwitch), one launch: the gradients of the large
 * tensors live at the same offsets of a "gradient block" on every rank, the parameters of a "parameter block"; all
 * blocks are peer-mapped into every process (map.grad[k], map.param[k] = rank k's blocks as seen from this process;
 * map.grad_mc / map.param_mc = multicast addresses of the blocks, or NULL).  For each tensor of the table this rank
 * updates its own slice: gradient = sum over ranks of grad[k][offset + i] (multimem.ld_reduce when the multicast
 * addresses are given, else peer loads in rank order), AdamW-amsgrad as above on t.p / t.acc / t.m / t.v / t.vmax
 * (t.p = this rank's slice inside its own parameter block, t.g ignored, t.n % 4 == 0), and the new values are
 * stored into every rank's parameter block (multimem.st / peer stores).  The caller brackets the launch with
 * barriers over all ranks: gradients complete before, nobody reads parameters until after. */
#define C2DSR_MAX_PEERS 16
typedef struct {
    c2dsr_adam_tensor t;
    int64_t offset;
} c2dsr_peer_tensor;
typedef struct {
    int world, rank;
    const float* grad[C2DSR_MAX_PEERS];
    float* param[C2DSR_MAX_PEERS];
    const float* grad_mc;
    float* param_mc;
} c2dsr_peer_map;
int c2dsr_adamw_amsgrad_peer(const c2dsr_peer_tensor* table_dev, int n_tensors, int64_t max_n,
                             const c2dsr_peer_map* map, const void* state, float beta1, float beta2, float eps,
                             float weight_decay, int background, void* stream);

/* elementwise glue: out = a*x + b*y (y may be NULL) */
int c2dsr_axpby(const float* x, const float* y, float* out, int64_t n, float a, float b, void* stream);

/* ---- preprocessor on the device (dataloader.py:60-228) -------------------------------------------------
 * items / offs: the time-sorted item lists of n_seq sequences, concatenated (sequence u = items[offs[u] .. offs[u+1]),
 * its last item is the final target).  The random ingredients come from the host, drawn from Python's `random` in
 * the reference's order: draws[offs[u] - u + i] = the corruption draw of input position i of sequence u
 * (dataloader.py:80,85); neg [n_seq, n_neg] holds, on entry, the sample of range(population) (dataloader.py:216-224)
 * and, on return, the negative ids (picks shifted past the target).  Outputs are the reference's fields:
 * fields [n_seq, 14, len_max] + keep [n_seq] (the reference drops sequences without a target in either domain);
 * six [n_seq, 6, len_max], four [n_seq, 4] = idx_last_a, idx_last_b, xory_last, gt_last.  All int64. */
int c2dsr_preprocess_train(const int64_t* items, const int64_t* offs, const int64_t* draws, int64_t n_seq,
                           int64_t n_item_a, int64_t n_item_b, int len_max, int64_t* fields, uint8_t* keep,
                           void* stream);
int c2dsr_preprocess_eval(const int64_t* items, const int64_t* offs, int64_t n_seq, int64_t n_item_a,
                          int64_t n_item_b, int len_max, int n_neg, int64_t* six, int64_t* four, int64_t* neg,
                          void* stream);

#ifdef __cplusplus
}
#endif
#endif /* C2DSR_B200_H */

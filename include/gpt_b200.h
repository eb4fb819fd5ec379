/* gpt_b200 -- C ABI of the B200 (sm_100a) GCN-over-pruned-trees hot path.
 *
 * The reference (gstoica27/gcn-over-pruned-trees) has no FFI: its hot path is Python calling numpy and ATen.
 * Each entry point below replaces a span of that Python; the span is cited as file:line into /root/reference.
 * INTEGRATION.md shows the ctypes binding and where each call goes in the reference's model/gcn.py.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer valid on `stream` unless marked host; the caller owns all buffers
 *   - stateless, no allocation, no synchronisation, safe under CUDA-graph capture
 *   - `stream` is a cudaStream_t passed as void*
 *   - return 0 on success, <0 for argument/support errors (GPT_ERR_*), >0 = cudaError_t of the launch
 *   - batch layout: B sentences padded to T tokens; row r = b*T + t; activations are row-major [B*T, H] fp32
 */
#ifndef GPT_B200_H
#define GPT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPT_OK 0
#define GPT_ERR_BAD_ARG (-1)
#define GPT_ERR_UNSUPPORTED (-2)
#define GPT_ERR_DRIVER (-3)

/* flags[b,t] bits (gpt_prune_csr output) */
#define GPT_FLAG_INTREE 1 /* token has >= 1 adjacency entry, i.e. is NOT pool-masked (model/gcn.py:262) */
#define GPT_FLAG_SUBJ 2   /* subj_pos == 0 (model/gcn.py:116) */
#define GPT_FLAG_OBJ 4    /* obj_pos == 0 */

/* err[b] bits (gpt_prune_csr output).  Bits 1..32 are fatal for the sentence: it gets an empty adjacency, where
 * the reference raises or never returns.  64 is a warning. */
#define GPT_TREE_HEAD_RANGE 1
#define GPT_TREE_NO_ROOT 2      /* model/tree.py:164 assert */
#define GPT_TREE_CYCLE 4        /* model/tree.py:91-94 loops forever */
#define GPT_TREE_EMPTY_SUBJ 8   /* model/tree.py:109 AttributeError */
#define GPT_TREE_DISJOINT 16    /* model/tree.py:121-127 UnboundLocalError */
#define GPT_TREE_DEPREL_RANGE 32
#define GPT_TREE_DEPREL_PAD 64  /* kept edge with deprel 0: forward entry is 0, adjacency not symmetric */

/* pooling type of gpt_pool3_* (model/gcn.py:473-483) */
#define GPT_POOL_MAX 0
#define GPT_POOL_AVG 1
#define GPT_POOL_SUM 2

/* GEMM arithmetic of gpt_linear_* */
#define GPT_GEMM_FP32 0   /* SIMT FFMA, fp32 accumulate: the 1e-5 parity mode */
#define GPT_GEMM_TF32 1   /* tcgen05 kind::tf32, one pass */
#define GPT_GEMM_TF32X3 2 /* tcgen05 kind::tf32, hi/lo split, three passes: fp32-grade accuracy on tensor cores */
#define GPT_GEMM_BF16 3   /* tcgen05 kind::f16 on bf16-rounded operands, fp32 accumulate */

int gpt_version(void);
/* host: number of kernels launched by this library so far in this process (bench.py's gpu_launches) */
unsigned long long gpt_launch_count(void);
/* host: static string for a return code of this library (cudaGetErrorString for positive codes) */
const char* gpt_error_string(int code);

/* prefetch [ptr, ptr + bytes) into L2 (no data is returned); used to warm the weights ahead of their first use */
int gpt_l2_prefetch(const void* ptr, long long bytes, void* stream);

/* K1. head_to_tree + tree_to_adj + the batch loop + adj!=0/denom/mask
 *     (model/tree.py:58-165, model/tree.py:167-204, model/gcn.py:96-110, model/gcn.py:260-262).
 * in : head, subj_pos, obj_pos, deprel  int64 [B,T] (loader layout, data/loader.py:110-121);
 *      pad_mask uint8/bool [B,T], nonzero on padding (the `masks` tensor)
 *      prune_k < 0: full tree, else path-centric pruning distance
 * out: rowptr int32 [B,T+1]  offsets into the sentence's segment of col/val (segment b starts at b*3*T)
 *      col    int32 [B,3*T]  column (token index inside the sentence), ascending inside a row
 *      val    uint8 [B,3*T]  the reference's adjacency value: deprel (parent->child), deprel+42 (child->parent), 84
 *      flags  uint8 [B,T]    GPT_FLAG_* ; denom float [B,T] = row nnz + 1 ; lens int32 [B] ; err int32 [B] */
int gpt_prune_csr(const int64_t* head, const int64_t* subj_pos, const int64_t* obj_pos, const int64_t* deprel,
                  const uint8_t* pad_mask, int B, int T, int prune_k, int32_t* rowptr, int32_t* col, uint8_t* val,
                  uint8_t* flags, float* denom, int32_t* lens, int32_t* err, void* stream);

/* K2 forward. One GCN layer after the projection y = h W^T (model/gcn.py:269-271, 390-393):
 *     out_i = dropout(relu((sum_{j in row i} y_j + y_i + 2*bias) / denom_i)); rows with flags == 0 are written 0.
 * use_adj = 0 is the --no_adj ablation (model/gcn.py:264-265).
 * act_mask (optional, uint32 [B, ceil(H/32), T]): one bit per element, set where out > 0, for the backward.
 * Element (b, row, col) is bit col%32 of word (b*ceil(H/32) + col/32)*T + row.
 * Dropout: drop_p > 0 draws Philox bits in-kernel from rng_state = {seed, step} (device uint64[2]) and `subseq`
 * (layer id); or drop_mask (pre-scaled float [B*T,H], tests) multiplies the result; or neither.
 * force_vec: 0 = auto, else 1/2/4 -> slice width 32/64/128 columns, one slice per CTA. */
int gpt_gcn_aggregate_fwd(const float* y, const int32_t* rowptr, const int32_t* col, const float* denom,
                          const uint8_t* flags, const float* bias, float* out, uint32_t* act_mask, int B, int T,
                          int H, int use_adj, float drop_p, const uint64_t* rng_state, uint32_t subseq,
                          const float* drop_mask, int force_vec, void* stream);

/* K2 backward (autograd of the lines above; the adjacency is symmetric so the same CSR is its transpose):
 *     g_i = gout_i * dropscale * [out_i > 0] / denom_i ; dy_j = g_j + sum_{i in row j} g_i ; dbias += 2*sum_i g_i
 * [out_i > 0] is read from act_mask when it is not NULL, else from `out` (one of the two is required).
 * dbias (float [H]) is accumulated atomically and must be zeroed by the caller; may be NULL. */
/*     Last layer fused with K4 (max pooling): same arithmetic, plus pooled [B,3H] = [h_out | subj_out | obj_out] and
 *     argmax int32 [B,3H] exactly as gpt_pool3_fwd(type max) would produce them from the layer output (ties: smallest
 *     row; empty pool: -1e12 / -1).  out may be NULL (the layer output is then never stored).  No dropout (the last
 *     layer has none, gcn.py:393).  gpt_gcn_aggregate_fwd_pool_supported(B,T,H) = 1 when the sentence tile fits the
 *     one-slice-per-CTA configuration this mode needs; otherwise the call returns GPT_ERR_UNSUPPORTED. */
int gpt_gcn_aggregate_fwd_pool(const float* y, const int32_t* rowptr, const int32_t* col, const float* denom,
                               const uint8_t* flags, const float* bias, float* out, uint32_t* act_mask, float* pooled,
                               int32_t* argmax, int B, int T, int H, int use_adj, void* stream);
int gpt_gcn_aggregate_fwd_pool_supported(int B, int T, int H);
/*     Last layer's backward fused with K4's backward (max pooling): takes d(pooled) [B,3H] and argmax [B,3H] instead of
 *     a [B,T,H] gradient; equals gpt_pool3_bwd_masked followed by gpt_gcn_aggregate_bwd_pre bit for bit. */
int gpt_gcn_aggregate_bwd_pool(const float* dpooled, const int32_t* argmax, const uint32_t* act_mask,
                               const int32_t* rowptr, const int32_t* col, const float* denom, float* dy, float* dbias,
                               int B, int T, int H, int use_adj, void* stream);
int gpt_gcn_aggregate_bwd(const float* gout, const float* out, const uint32_t* act_mask, const int32_t* rowptr,
                          const int32_t* col, const float* denom, float* dy, float* dbias, int B, int T, int H,
                          int use_adj, float drop_p, const float* drop_mask, int force_vec, void* stream);

/* K2 backward when the first step (g = gout * dropscale * [out > 0] / denom) was fused into the kernel that produced
 * the gradient (gpt_linear_dgrad_tf32x3_masked, gpt_pool3_bwd_masked): dy_j = g_j + sum_{i in row j} g_i,
 * dbias += 2 * sum_i g_i.  One shared-memory pass and one barrier per slice less than gpt_gcn_aggregate_bwd. */
int gpt_gcn_aggregate_bwd_pre(const float* g, const int32_t* rowptr, const int32_t* col, const float* denom, float* dy,
                              float* dbias, int B, int T, int H, int use_adj, int force_vec, void* stream);
/* The two calls above with a second, compact copy of dy: every row n with inv[n] >= 0 (gpt_live_rows) is also stored at
 * dy_compact[inv[n], :], so that the tensor-core weight gradient (gpt_linear_wgrad_tf32x3_rows) reads the live rows as one
 * dense block without a gather launch in between.  inv == dy_compact == NULL: exactly the calls above. */
int gpt_gcn_aggregate_bwd_pool_c(const float* dpooled, const int32_t* argmax, const uint32_t* act_mask,
                                 const int32_t* rowptr, const int32_t* col, const float* denom, float* dy, float* dbias,
                                 const int32_t* inv, float* dy_compact, int B, int T, int H, int use_adj, void* stream);
int gpt_gcn_aggregate_bwd_pre_c(const float* g, const int32_t* rowptr, const int32_t* col, const float* denom, float* dy,
                                float* dbias, const int32_t* inv, float* dy_compact, int B, int T, int H, int use_adj,
                                int force_vec, void* stream);

/* K4. three masked pools + cat (model/gcn.py:116-121, 473-483): out float [B,3H] = [h_out, subj_out, obj_out];
 * argmax int32 [B,3H] (token index or -1) is required for GPT_POOL_MAX. */
int gpt_pool3_fwd(const float* h, const uint8_t* flags, int B, int T, int H, int pool_type, float* out,
                  int32_t* argmax, void* stream);
int gpt_pool3_bwd(const float* gout, const int32_t* argmax, const uint8_t* flags, int B, int T, int H,
                  int pool_type, float* dh, void* stream);

/* K4 backward fused with the first step of the last layer's K2 backward (act / denom / drop_scale as in
 * gpt_gcn_aggregate_bwd): writes g instead of dh */
int gpt_pool3_bwd_masked(const float* gout, const int32_t* argmax, const uint8_t* flags, const uint32_t* act,
                         const float* denom, float drop_scale, int B, int T, int H, int pool_type, float* g,
                         void* stream);

/* K3. the W projection without bias (model/gcn.py:270-271: W(Ax) + W(h) == (A+I) (h W^T) + 2b) and its autograd.
 *     x [M,K], w [N,K] (nn.Linear layout), y/dy [M,N], dx [M,K], dw [N,K]; all row-major fp32. */
int gpt_linear_fwd_f32(const float* x, const float* w, float* y, int M, int N, int K, void* stream);
int gpt_linear_dgrad_f32(const float* dy, const float* w, float* dx, int M, int N, int K, void* stream);
int gpt_linear_wgrad_f32(const float* dy, const float* x, float* dw, int M, int N, int K, void* stream);
/* dw += dy^T x (no zero-fill launch: for gradient buffers the caller keeps zeroed between steps, see K7) */
int gpt_linear_wgrad_f32_acc(const float* dy, const float* x, float* dw, int M, int N, int K, void* stream);
/* dw += dy^T x over the rows with flags[m] != 0 only (flags = gpt_prune_csr's [B*T] output, NULL = all rows): the
 * other rows of dy are exactly zero (K2 backward), so they are never read */
int gpt_linear_wgrad_rows_f32(const float* dy, const float* x, const uint8_t* flags, float* dw, int M, int N, int K,
                              void* stream);
/*     the same on the tensor cores (csrc/wgrad_tcgen05.cu): both operands are taken MN-major exactly as they lie in
 *     HBM (TMA boxes of [32 fp32 x 16 rows] = canonical swizzle atoms, no transposed copy), 3xTF32 (fp32-grade) with the
 *     fp32 accumulator in tensor memory, the row range split over one CTA per SM, partial sums added into dw with
 *     vector reductions.  flags may be NULL (every row).  GPT_ERR_UNSUPPORTED when K > 512 or K, N are not multiples of
 *     4: use gpt_linear_wgrad_rows_f32. */
int gpt_linear_wgrad_tf32x3(const float* dy, const float* x, const uint8_t* flags, float* dw, long long M, int N, int K,
                            void* stream);
/* K3 on the tensor cores: tcgen05.mma kind::tf32, TMA-fed, accumulator in TMEM (GPT_GEMM_TF32, ~1e-3 relative).
 *     Needs K % 4 == 0 (and N % 4 == 0 for dgrad) and 16-byte aligned operands, else GPT_ERR_UNSUPPORTED.
 *     dgrad takes a float [K*N] workspace for the transposed weight. */
int gpt_linear_fwd_tf32(const float* x, const float* w, float* y, int M, int N, int K, void* stream);
int gpt_linear_dgrad_tf32(const float* dy, const float* w, float* dx, float* wt_workspace, int M, int N, int K,
                          void* stream);
/* K3, 3xTF32 (GPT_GEMM_TF32X3): A.B ~ A_hi.B_hi + A_lo.B_hi + A_hi.B_lo with hi = round_tf32(x), lo = x - hi;
 *     fp32-grade accuracy on the tensor cores (the default projection; passes the 1e-5 logits parity).
 *     gpt_weight_prep_tf32x3 splits / transposes the weight once per step into ws = float [4*N*K]
 *     ([w_hi | w_lo | w^T_hi | w^T_lo]); fwd and dgrad then take ws in place of w. */
int gpt_weight_prep_tf32x3(const float* w, float* ws, int N, int K, void* stream);
/*     the same for n_layers <= 8 weights in one launch (host arrays of device pointers / sizes) */
int gpt_weight_prep_tf32x3_batch(const float* const* w, float* const* ws, const int* N, const int* K, int n_layers,
                                 void* stream);
int gpt_linear_fwd_tf32x3(const float* x, const float* ws, float* y, int M, int N, int K, void* stream);
int gpt_linear_dgrad_tf32x3(const float* dy, const float* ws, float* dx, int M, int N, int K, void* stream);
/*     Large M (>= min_rows, default 65 536; BASELINE.json configs[4]) runs the persistent kernel of csrc/gemm_persist.cuh
 *     behind the same four entry points: CTA pairs (tcgen05.mma.cta_group::2, M = 256), two accumulators in tensor
 *     memory, TMA-store epilogue.  cta_group: 0 = never, 1 = single-CTA tiles, 2 = pairs (default); process-wide. */
int gpt_gemm_persist_config(int cta_group, long long min_rows);
/* K3, bf16 operands (GPT_GEMM_BF16; north_star's reduced-precision mode): tcgen05.mma kind::f16 with A and B rounded to
 *     bfloat16 (round-to-nearest-even), fp32 accumulation in TMEM -- ~4e-3 relative per product, logits within ~3e-2 of
 *     the fp32 path (tolerance written in tests/test_gpu_bf16.py).  X stays fp32 in HBM and is rounded in shared memory
 *     after the TMA load; gpt_weight_prep_bf16 writes ws = bf16 [2*N*K] ([bf16(w) | bf16(w^T)]) once per step.
 *     Needs K % 8 == 0 (fwd) / N % 8 == 0 (dgrad).  The weight gradient of this mode stays 3xTF32. */
int gpt_weight_prep_bf16(const float* w, void* ws, int N, int K, void* stream);
int gpt_linear_fwd_bf16(const float* x, const void* ws, float* y, int M, int N, int K, void* stream);
int gpt_linear_dgrad_bf16(const float* dy, const void* ws, float* dx, int M, int N, int K, void* stream);
/* dgrad with the previous layer's K2-backward prologue in the epilogue: g = (dy . w) * drop_scale_prev *
 * [out_prev > 0] / denom, act_prev in K2's bit layout over the K columns of dx, rows = B*T sentences of T tokens */
int gpt_linear_dgrad_tf32x3_masked(const float* dy, const float* ws, float* g, const uint32_t* act_prev,
                                   const float* denom, float drop_scale_prev, int T, int M, int N, int K, void* stream);

/* K5. input stage of GCN.forward (model/gcn.py:235-247): x[r] = dropout(cat[emb_w[words[r]], pos_w[pos[r]],
 *     ner_w[ner[r]]]) for the n_rows = B*T token slots; x is [n_rows, E+Dp+Dn].  pos/pos_w and ner/ner_w are NULL when
 *     Dp / Dn is 0 (ner: SemEval).  Dropout as in K2 (Philox from rng_state = {seed, step}). */
int gpt_embed_fwd(const int64_t* words, const int64_t* pos, const int64_t* ner, const float* emb_w, const float* pos_w,
                  const float* ner_w, float* x, int n_rows, int V, int E, int Dp, int Dn, float drop_p,
                  const uint64_t* rng_state, uint32_t subseq, void* stream);
/* gpt_embed_fwd and gpt_weight_prep_tf32x3_batch in ONE launch -- the front of a captured training step: as two root nodes
 * of the graph they start microseconds apart and the first projection waits for the later one across streams.  w / ws /
 * wN / wK: host arrays of n_layers (<= 8) entries, as gpt_weight_prep_tf32x3_batch takes them. */
int gpt_embed_fwd_prep(const int64_t* words, const int64_t* pos, const int64_t* ner, const float* emb_w, const float* pos_w,
                       const float* ner_w, float* x, int n_rows, int V, int E, int Dp, int Dn, float drop_p,
                       const uint64_t* rng_state, uint32_t subseq, const float* const* w, float* const* ws, const int* wN,
                       const int* wK, int n_layers, void* stream);
/* K5 backward: scatter-add dx (re-applying the same dropout mask) into the gradient tables (all accumulated with
 *     atomics, caller zeroes them; any of g_emb/g_pos/g_ner may be NULL).  Rows with flags == 0 are skipped (their
 *     gradient is exactly zero), word id 0 (padding_idx) and ids >= topn (model/gcn.py:83-86) get no gradient.
 *     owner (optional int32 [V], preset to INT_MAX): first token index of every touched word row. */
int gpt_embed_bwd(const float* dx, const uint8_t* flags, const int64_t* words, const int64_t* pos, const int64_t* ner,
                  float* g_emb, float* g_pos, float* g_ner, int32_t* owner, int n_rows, int V, int E, int Dp, int Dn,
                  int topn, float drop_p, const uint64_t* rng_state, uint32_t subseq, void* stream);
/* Row-sparse tail of clip_grad_norm_ + SGD for the word embedding (train.py:224-227), touching only the rows the
 *     batch used: sq += sum |G[w]|^2 over touched rows;  then  W[w] -= lr * min(1, max_norm/(sqrt(*total_sq)+1e-6)) *
 *     G[w], G[w] = 0, owner[w] = INT_MAX.  total_sq must hold the squared global gradient norm of ALL parameters. */
/*     The same gradients for LARGE batches without floating-point atomics on the word table: the live token rows are
 *     grouped by word (counting sort: count -> scan -> fill, int32 workspace of gpt_embed_bwd_grouped_workspace(n_rows, V)
 *     entries) and one warp per word sums its rows of dX (dropout mask re-derived per row) and adds the result to g_emb[w]
 *     once; the POS / NER columns are collected in shared memory per CTA.  5.5 -> 1.4 ms at 2 M tokens.  Needs E % 4 == 0,
 *     (E + Dp + Dn) % 4 == 0, E <= 512, else GPT_ERR_UNSUPPORTED (use gpt_embed_bwd). */
long long gpt_embed_bwd_grouped_workspace(int n_rows, int V);
int gpt_embed_bwd_grouped(const float* dx, const uint8_t* flags, const int64_t* words, const int64_t* pos,
                          const int64_t* ner, float* g_emb, float* g_pos, float* g_ner, int32_t* owner, int n_rows, int V,
                          int E, int Dp, int Dn, int topn, float drop_p, const uint64_t* rng_state, uint32_t subseq,
                          int32_t* workspace, void* stream);
int gpt_embed_rows_sqnorm(const int64_t* words, const int32_t* owner, const float* g_emb, int n_rows, int E, int topn,
                          float* sq, void* stream);
int gpt_embed_rows_sgd(const int64_t* words, int32_t* owner, float* g_emb, float* emb_w, int n_rows, int E, int topn,
                       const float* total_sq, float max_norm, float lr, void* stream);

/* K11. eval-side tail of GCNTrainer.predict (model/trainer.py:112-124) in one launch: mean CrossEntropy (:118), softmax
 *     (:119), argmax with numpy's first-maximum rule (:120) and the un-sort to the loader's original order (:121-123).
 *     dest[b] = position of batch row b in the original order (NULL: keep the batch order).  result: one packed device
 *     buffer of gpt_predict_result_bytes(B, C) bytes, [ probs f32 [B,C] | predictions i32 [B] | loss f32 ], rows already in
 *     the original order -- the host needs a single device-to-host copy. */
long long gpt_predict_result_bytes(int B, int C);
int gpt_predict_tail(const float* logits, const int64_t* labels, const int32_t* dest, int B, int C, void* result,
                     void* stream);

/* K6. classifier head for one batch of pooled vectors [B,3H] (K4 output): out_mlp (model/gcn.py:64-68,122: n_mlp x
 *     Linear+ReLU, w[0] [H,3H], w[l>0] [H,H]), classifier (model/gcn.py:21,29: wc [C,H]), and the loss of
 *     model/trainer.py:94-100 without the conv_l2 term: CrossEntropy(mean) + pooling_l2 * mean_b sum_h pooled[b,h<H]^2.
 *     w / b are HOST arrays of n_mlp device pointers.  Needs H % 4 == 0, n_mlp <= 4.
 *     out: logits [B,C]; loss_rows [B] (the loss is their sum).  With train != 0 also the backward of loss.backward()
 *     (train.py:221) down to dpooled [B,3H], plus acts / dacts [B,n_mlp,H] (post-ReLU activations and
 *     d loss / d pre-activations) and dlogits [B,C] for gpt_head_wgrad. */
int gpt_head_fwd_bwd(const float* pooled, const int64_t* labels, const float* const* w, const float* const* b,
                     int n_mlp, const float* wc, const float* bc, int B, int H, int C, float pooling_l2, int train,
                     float* logits, float* loss_rows, float* acts, float* dacts, float* dlogits, float* dpooled,
                     void* stream);
/* K6 weight gradients: dw[l] = dacts[:,l,:]^T . in_l, db[l] = column sums (in_0 = pooled, in_l = acts[:,l-1,:]);
 *     dwc / dbc from dlogits and the last activation; *loss = sum_b loss_rows[b].  Overwrites (no accumulation, no
 *     atomics: deterministic).  dw / db are HOST arrays of n_mlp device pointers. */
int gpt_head_wgrad(const float* pooled, const float* acts, const float* dacts, const float* dlogits,
                   const float* loss_rows, int B, int H, int C, int n_mlp, float* const* dw, float* const* db,
                   float* dwc, float* dbc, float* loss, void* stream);

/* K7. clip_grad_norm_(max_norm) + plain SGD + zero_grad (train.py:224-227, --optim sgd) over ONE flat fp32 parameter /
 *     gradient buffer of n elements (16-byte aligned) plus the word-embedding rows the batch touched (words int64
 *     [n_rows], owner / g_emb as written by gpt_embed_bwd; n_rows may be 0).
 *     gpt_update_partials (host) = number of floats `partials` must hold.  gpt_update_sqnorm writes per-CTA partial
 *     sums of g^2; gpt_update_apply re-adds them in a fixed order, norm = grad_scale * sqrt(sum),
 *     coef = min(1, max_norm / (norm + 1e-6)) (max_norm <= 0: no clipping), then p -= lr * coef * grad_scale * g and
 *     g = 0 for the flat buffer and the live rows (owner reset to INT_MAX).  total_norm (optional) receives norm;
 *     *step_counter (optional; word 1 of the {seed, step} dropout state) is incremented. */
int gpt_update_partials(long long n, int n_rows);
int gpt_update_sqnorm(const float* grad, long long n, const int64_t* words, const int32_t* owner, const float* g_emb,
                      int n_rows, int E, int topn, float* partials, void* stream);
int gpt_update_apply(float* param, float* grad, long long n, const int64_t* words, int32_t* owner, float* g_emb,
                     float* emb_w, int n_rows, int E, int topn, const float* partials, float max_norm, float lr,
                     float grad_scale, float* total_norm, uint64_t* step_counter, void* stream);

/* K8. data-parallel gradient exchange over NVLink peer memory, fused with K7 (csrc/dp.cu; the reference is
 *     single-device, so this replaces nothing: it is the one collective sentence sharding adds).
 *     Every rank owns a region of gpt_dp_region_bytes(); regions[q] is rank q's region as mapped in this process
 *     (own: gpt_dp_alloc, peers: gpt_dp_open on the 64-byte cudaIpc handle).  W <= 8; cap_rows >= n_rows of any step.
 *     Per step, in stream order on every rank:
 *       gpt_dp_push    stores this rank's flat gradient, live word ids / rows and word->slot map into every region;
 *                      clears g_emb rows and owner marks.  No fence inside: the flags are raised by the next launch
 *       gpt_dp_reduce  signal != 0: first raises flags[rank] = step in every region (the push grid has completed by
 *                      then, so its peer stores have been performed); waits for all W flags, sums in rank order
 *                      (bit-identical on every rank) into flat_g and the first-owner row slots, writes
 *                      gpt_dp_partials() partial sums of g^2.  signal == 0: only regions[rank] is read and the caller
 *                      raises the flags itself with gpt_dp_signal (several ranks driven from one stream: tests)
 *       gpt_dp_apply   K7 with the mean gradient (sum / W): clip, SGD, resets; advances the region's step and
 *                      *step_counter
 *     A rank that waits more than ~10 s for a peer traps (the launch fails) instead of spinning for ever. */
long long gpt_dp_region_bytes(int W, int cap_rows, int E, int V, long long n_flat);
int gpt_dp_partials(int W, int cap_rows, int E, int V, long long n_flat);
int gpt_dp_alloc(long long bytes, void** ptr, void* ipc_handle_out /* 64 bytes, host */);
int gpt_dp_open(const void* ipc_handle, void** ptr);
int gpt_dp_close(void* ptr);
int gpt_dp_free(void* ptr);
int gpt_dp_region_init(void* region, int W, int cap_rows, int E, int V, long long n_flat, void* stream);
int gpt_dp_push(void* const* regions /* host array of W device pointers */, int rank, int W, int cap_rows, int E, int V,
                long long n_flat, const float* flat_g, float* g_emb, int32_t* owner, const int64_t* words, int n_rows,
                int topn, void* stream);
/*     The same push through an NVSwitch multicast mapping of the W regions (`multicast`: e.g. the multicast_ptr of torch
 *     symmetric memory holding the regions): multimem.st, every byte leaves the GPU once instead of W - 1 times; plain
 *     data movement, the rank-ordered adds of gpt_dp_reduce stay local. */
int gpt_dp_push_multicast(void* const* regions, void* multicast, int rank, int W, int cap_rows, int E, int V,
                          long long n_flat, const float* flat_g, float* g_emb, int32_t* owner, const int64_t* words,
                          int n_rows, int topn, void* stream);
int gpt_dp_signal(void* const* regions, int rank, int W, int cap_rows, int E, int V, long long n_flat, void* stream);
int gpt_dp_reduce(void* const* regions, int rank, int signal, int W, int cap_rows, int E, int V, long long n_flat,
                  float* flat_g, float* partials, void* stream);
int gpt_dp_apply(void* region, int W, int cap_rows, int E, int V, long long n_flat, float* param, float* flat_g,
                 float* emb_w, const float* partials, float max_norm, float lr, float* total_norm,
                 uint64_t* step_counter, void* stream);

/* K9. device-resident batch builder (csrc/batch.cu): DataLoader.__getitem__ of the reference (data/loader.py:81-141,
 *     semeval_loader.py:75-119) from a token arena in HBM.  arena[7] = host array of device pointers to int32 arrays
 *     indexed by token {words, pos, ner, deprel, head, subj_pos, obj_pos} (ner NULL for 9-tuple batches), offsets int64
 *     [n_sentences+1], labels int32 [n_sentences], sel int32 [B] = the batch's sentence ids in output row order
 *     (the caller sorts by length, loader.py:176-180).  Writes out[7] int64 [B,T] (pad 0; 150 for the two position
 *     fields), masks uint8 [B,T] = (t >= len), rels int64 [B].  word_dropout > 0: a token != <UNK> becomes <UNK>
 *     with that probability (loader.py:181-188), Philox stream keyed by (seed, stream_id, sentence, token). */
int gpt_build_batch(const int32_t* const* arena, const int64_t* offsets, const int32_t* labels, const int32_t* sel,
                    int B, int T, float word_dropout, uint64_t seed, uint64_t stream_id, int64_t* const* out,
                    uint8_t* masks, int64_t* rels, void* stream);

/* K10. relation-aware GCN layers (csrc/deprel.cu): adj_type 'full_deprel' (model/gcn.py:296-386, 400-434) and
 *      'diagonal_deprel' (model/gcn.py:272-294), with edge dropout (:436-449), relation forgetting (:451-470),
 *      deprel_max_depth, deprel_directed, deprel_self_loop, and the common /denom, ReLU, gcn_drop (:390-393).
 *      The layer is: Z = x . Wmat^T with gpt_linear_fwd_* (Wmat[d*H+h, k] = W.weight.reshape(D,K,H)[d,k,h], :301),
 *      gpt_relmix_fwd, gpt_agg3_fwd; backward: gpt_agg3_bwd, gpt_relmix_bwd, gpt_colsum_acc, gpt_linear_dgrad/wgrad_*.
 *
 * gpt_relmix_fwd: F/R/S[n,:] = sum_d e[d] (Z[n,d,:] + bias[d,:]) with e = E[deprel[n]] / E[deprel[n]+42] / E[84]
 *     (traverse_deprel :400-415, traverse_self_loop :417-434).  E float [85,D]; Z float [N, D*H]; bias float [D*H];
 *     keep_f / keep_r (optional uint8 [N]): 0 = the token's relation vector of that direction is replaced by ones
 *     (maybe_forget_deprels); deep != 0 = every vector is ones (layer >= deprel_max_depth, :324-325,355-356,376-379).
 *     Rows with flags == 0 are written 0 (as in K2); entity tokens outside the tree are computed.
 * gpt_relmix_bwd: dZ [N, D*H] (zeros on rows with flags == 0) and dE [85,D] += (caller-zeroed, atomics; row 0 =
 *     padding_idx and forgotten / deep vectors get nothing) from dF, dR, dS [N,H]. */
int gpt_relmix_fwd(const float* Z, const float* bias, const float* E, const int64_t* deprel, const uint8_t* flags,
                   const uint8_t* keep_f, const uint8_t* keep_r, int N, int D, int H, int deep, float* F, float* R,
                   float* S, void* stream);
int gpt_relmix_bwd(const float* Z, const float* bias, const float* E, const int64_t* deprel, const uint8_t* flags,
                   const uint8_t* keep_f, const uint8_t* keep_r, const float* dF, const float* dR, const float* dS, int N,
                   int D, int H, int deep, float* dZ, float* dE, void* stream);
/* diagonal_deprel (model/gcn.py:272-294): F = E[deprel] * x, R = E[deprel+42] * x, S = E[84] * x, E float [85,H];
 * backward: dx [N,H], dE [85,H] += (caller-zeroed). */
int gpt_diagmix_fwd(const float* x, const float* E, const int64_t* deprel, const uint8_t* flags, int N, int H, float* F,
                    float* R, float* S, void* stream);
int gpt_diagmix_bwd(const float* x, const float* E, const int64_t* deprel, const uint8_t* flags, const float* dF,
                    const float* dR, const float* dS, int N, int H, float* dx, float* dE, void* stream);
/* gpt_agg3_fwd: out_i = dropout(relu((sum_{c: 0<val<42} keep_f[i,c] F_c + sum_{p: 42<val<84} keep_r[i,p] R_p + S_i) /
 *     denom_i)) over gpt_prune_csr's rows (forward_adj_matrix.bmm + reverse_adj_matrix.bmm + self loop, :308-386, and
 *     :390-393).  directed != 0 drops the R term (:339), self_loop == 0 the S term (:369).
 *     Edge dropout (maybe_drop_edges): keep_f / keep_r = optional dense uint8 [B,T,T] masks, one per direction; else
 *     edge_keep < 1 draws Bernoulli(edge_keep) per matrix entry in-kernel (Philox keyed by rng_state = {seed, step},
 *     layer, direction, entry) -- gpt_edge_keep_dense materialises exactly those decisions.  No rescaling (:445).
 *     Dropout: drop_mask (pre-scaled float [N,H]) or drop_p > 0 (in-kernel Philox) or neither.
 * gpt_agg3_bwd: dF, dR, dS [N,H] from gout and the forward's out (d out/d z is recovered from out != 0). */
int gpt_agg3_fwd(const float* F, const float* R, const float* S, const int32_t* rowptr, const int32_t* col,
                 const uint8_t* val, const float* denom, const uint8_t* flags, const uint8_t* keep_f,
                 const uint8_t* keep_r, float edge_keep, const void* rng_state, unsigned layer, int directed,
                 int self_loop, float drop_p, const float* drop_mask, int B, int T, int H, float* out, void* stream);
int gpt_agg3_bwd(const float* gout, const float* out, const int32_t* rowptr, const int32_t* col, const uint8_t* val,
                 const float* denom, const uint8_t* flags, const uint8_t* keep_f, const uint8_t* keep_r, float edge_keep,
                 const void* rng_state, unsigned layer, int directed, int self_loop, float drop_p,
                 const float* drop_mask, int B, int T, int H, float* dF, float* dR, float* dS, void* stream);
/* the in-kernel edge-dropout decisions of (layer, dir) as a dense uint8 [B,T,T]; dir 0 = parent->child matrix */
int gpt_edge_keep_dense(const void* rng_state, int B, int T, unsigned layer, int dir, float keep_prob, uint8_t* out,
                        void* stream);
/* relation forgetting: keep_f / keep_r uint8 [N] ~ Bernoulli(keep_prop), independent per direction (:451-470) */
int gpt_relation_keep_tokens(const void* rng_state, int N, unsigned layer, float keep_prop, uint8_t* keep_f,
                             uint8_t* keep_r, void* stream);
/* out[c] += sum_r a[r,c] (bias gradient of the shared projection); out is caller-zeroed */
int gpt_colsum_acc(const float* a, long long rows, int cols, float* out, void* stream);

/* K10 over the OBSERVABLE rows only.  The reference computes every [B,T] position of every layer (gcn.py:296-386) although
 * only rows inside a pruned tree, or subject / object tokens, ever reach a pool (gcn.py:116-120, 262); at prune_k = 1 that is
 * a quarter of a TACRED-shaped batch.  gpt_live_rows lists the rows with flags != 0 (gpt_prune_csr) in ascending order:
 * perm int32 [N] (first *count entries), inv int32 [N] (position of row n in perm, or -1), live uint8 [N] (i < count: the
 * row flags of the compact arrays, which the weight-gradient entry points take), count int32 [1] -- which stays on the
 * device: a captured step never learns it on the host, so every consumer below takes the POINTER.
 *   gpt_gather_rows       out[i,:] = x[perm[i],:]  for i < *count           (x, out float [N,K]; other rows of out untouched)
 *   gpt_scatter_rows      dx[n,:] = inv[n] >= 0 ? dxc[inv[n],:] : 0        for every n < N
 *   gpt_linear_{fwd,dgrad}_tf32x3_rows, gpt_linear_wgrad_tf32x3_rows:  the projections of gpt_linear_*_tf32x3 over the first
 *                         *m_live rows (row tiles beyond leave at once; output rows beyond keep what they held); the
 *                         forward takes an optional column bias (float [N], 16-byte aligned, N % 4 == 0) added in its
 *                         epilogue -- gpt_relmix_*_rows are then called with bias == NULL (the mix reads every projected
 *                         element once per direction: a bias left for it to add doubles its L2 traffic)
 *   gpt_relmix_{fwd,bwd}_rows   Z / dZ compact [count, D*H], everything per token (deprel, flags, keep_*, F/R/S, dF/dR/dS)
 *                         addressed through perm; perm == count == NULL is gpt_relmix_{fwd,bwd}.  Rows with flags == 0 are
 *                         not written in the compact form (nothing reads them: the aggregation gathers kept rows only)
 *   gpt_colsum_acc_rows   gpt_colsum_acc over rows r < *count */
int gpt_live_rows(const uint8_t* flags, int N, int32_t* perm, int32_t* inv, uint8_t* live, int32_t* count, void* stream);
int gpt_gather_rows(const float* x, const int32_t* perm, const int32_t* count, int N, int K, float* out, void* stream);
int gpt_scatter_rows(const float* dxc, const int32_t* inv, int N, int K, float* dx, void* stream);
int gpt_linear_fwd_tf32x3_rows(const float* x, const float* ws, const float* bias, float* y, int M, int N, int K,
                               const int32_t* m_live, void* stream);
int gpt_linear_dgrad_tf32x3_rows(const float* dy, const float* ws, float* dx, int M, int N, int K, const int32_t* m_live,
                                 void* stream);
int gpt_linear_wgrad_tf32x3_rows(const float* dy, const float* x, const uint8_t* flags, float* dw, long long M, int N, int K,
                                 const int32_t* m_live, void* stream);
int gpt_relmix_fwd_rows(const float* Z, const float* bias, const float* E, const int64_t* deprel, const uint8_t* flags,
                        const uint8_t* keep_f, const uint8_t* keep_r, const int32_t* perm, const int32_t* count, int N, int D,
                        int H, int deep, float* F, float* R, float* S, void* stream);
int gpt_relmix_bwd_rows(const float* Z, const float* bias, const float* E, const int64_t* deprel, const uint8_t* flags,
                        const uint8_t* keep_f, const uint8_t* keep_r, const int32_t* perm, const int32_t* count,
                        const float* dF, const float* dR, const float* dS, int N, int D, int H, int deep, float* dZ, float* dE,
                        void* stream);
int gpt_colsum_acc_rows(const float* a, long long rows, int cols, const int32_t* count, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GPT_B200_H */

/* ruart_b200 — C ABI of the B200 (sm_100a) kernels behind RUArt's per-question inference path.
 *
 * The reference (xiaojino/RUArt) is pure Python/PyTorch plus one CPython extension
 * (Utils/cphoc.c).  It has no FFI for the model path, so each entry point below cites the
 * reference *Python/C interface it replaces* (file:line relative to the reference tree).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary
 *   - every pointer is a DEVICE pointer unless the parameter name ends in `_host`
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream)
 *   - return value: 0 (RUART_OK) on success, otherwise an error code; the message is kept in a
 *     thread-local buffer readable with ruart_last_error()
 *   - no allocation inside unless stated; callers pass outputs / workspaces
 *   - all functions are asynchronous with respect to the host unless stated
 */
#ifndef RUART_B200_H_
#define RUART_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define RUART_API __attribute__((visibility("default")))
#else
#define RUART_API
#endif

#define RUART_PHOC_DIM 604

/* GEMM epilogues */
#define RUART_EPI_NONE 0       /* C = A W^T                                             */
#define RUART_EPI_BIAS 1       /* C = A W^T + b            nn.Linear                    */
#define RUART_EPI_BIAS_GELU 2  /* C = gelu_erf(A W^T + b)  modeling.py:52-57,286-289    */
#define RUART_EPI_RELU_SCALE 3 /* C = relu(A W^T) * d      Layers.py:226-231            */
#define RUART_EPI_BIAS_RELU 4  /* C = relu(A W^T + b)                                   */

/* ---------------------------------------------------------------- library management */
RUART_API const char* ruart_last_error(void);
RUART_API int ruart_version(void);
/* number of SMs of the current device (cached) */
RUART_API int ruart_num_sms(void);

/* ---------------------------------------------------------------- PHOC featuriser
 * Replaces Utils/cphoc.c:12-113 `build_phoc(str) -> float[604]` (bound in Python by
 * Utils/phoc.py:8-13).  Batch form: string i is chars[offsets[i] .. offsets[i+1]).
 * `err` is an 8-byte, 8-byte-aligned device buffer initialised by the callee to ~0; unknown
 * characters (outside [a-z0-9]) are reported like the reference's RuntimeError
 * (cphoc.c:45-50): that string's row is left all-zero and err receives the smallest key
 * (string_index << 24 | char_position << 8 | char) over all offending characters.           */
RUART_API int ruart_phoc_batch(const uint8_t* chars, const int32_t* offsets, int64_t n, float* out,
                     int32_t* err, void* stream);
/* Same, writing row i at out + i*out_pitch floats (out_pitch >= 604, multiple of 4; out 16-byte
 * aligned): the PHOC channel of an item embedding `[phoc604 | word300 | bert768 | ...]`
 * (Models/SDNet.py:441-446) computed from the strings in place of the `[V,604]` table lookup.   */
RUART_API int ruart_phoc_batch_pitched(const uint8_t* chars, const int32_t* offsets, int64_t n,
                             float* out, int64_t out_pitch, int32_t* err, void* stream);
/* Same, bit-packed: 19 uint32 words per string (bit b of word w = feature 32*w+b). */
RUART_API int ruart_phoc_batch_packed(const uint8_t* chars, const int32_t* offsets, int64_t n,
                            uint32_t* out_words, int32_t* err, void* stream);
/* Host-buffer convenience (synchronous): copies in, runs, copies out.  Returns
 * RUART_ERR_PHOC_CHAR and fills bad_index/bad_char when a string has an unknown unigram.  */
RUART_API int ruart_phoc_batch_host(const char* chars_host, const int32_t* offsets_host, int64_t n,
                          float* out_host, int64_t* bad_index, int32_t* bad_char);

/* ---------------------------------------------------------------- tcgen05 GEMM
 * C[M,N] = epi(A[M,K] W[N,K]^T).  A and W are bf16, K-major, possibly "split" into parts laid
 * side by side in each row ([rows, parts*Kp], Kp a multiple of 64, zero padded).
 * n_terms: 1 (plain bf16), 3 (two-part split, ~2^-16 rel) or 6 (three-part split, ~fp32).
 * Outputs: fp32 (out_f32, ldo_f32) and/or bf16 (out_bf16, ldo_bf16); the bf16 output may itself
 * be written as out_parts split parts, part p at column offset p*out_part_stride.
 * fast_gelu (RUART_EPI_BIAS_GELU): 0 = erff(); 1 = erf by Abramowitz-Stegun 7.1.26 on MUFU rcp/ex2
 * (|gelu error| < 1e-5); 2 = erf matched by a fitted tanh form on one MUFU.TANH (< 4e-5).
 * residual_bf16 (optional, plain-bf16-output mode with N % 64 == 0): a [M, N] bf16 matrix added in
 * fp32 before the output is rounded (dense(x) + input_tensor, modeling.py:263,302).
 * Replaces nn.Linear at modeling.py:225-227,261,287,300 and Layers.py:226-227,166 (W_ih).    */
RUART_API int ruart_gemm_bf16(const void* A, long long lda, int a_parts, const void* W, long long ldw,
                    int w_parts, int M, int N, int Kp, int n_terms, int epi, const float* bias,
                    const float* scale, int scale_len, float* out_f32, long long ldo_f32,
                    void* out_bf16, long long ldo_bf16, int out_parts, long long out_part_stride,
                    int fast_gelu, const void* residual_bf16, long long ld_res, void* stream);

/* The CTA-pair GEMM with a FOLDED BertLayerNorm (modeling.py:155-168 fused into its neighbours; bf16 encoder).
 * LayerNorm is affine per row, y = (v - mu) r gamma + beta, so the encoder keeps the rows v BEFORE each
 * LayerNorm (bf16) plus 8 float2 partial (sum, sum of squares) per row, and
 *   fold 2 (BertSelfOutput.dense / BertOutput.dense, modeling.py:260-264,299-303):
 *       out = A W^T + LayerNorm(residual) + b  written pre-LayerNorm, with its partial sums in out_stats;
 *       vec = beta + b, vec2 = gamma (of the residual's pending LayerNorm), in_stats = the residual's sums
 *   fold 1 (query/key/value, BertIntermediate.dense: the consumers of a LayerNorm output, :225-227,287):
 *       out = act(LayerNorm(A) W0^T + b) computed as r acc - r mu colsum + vec with W = W0 * gamma (host),
 *       vec = W0 beta + b, vec2 = colsum(W), in_stats = A's sums; epi = RUART_EPI_BIAS or RUART_EPI_BIAS_GELU
 * Requires M >= 2048, N % 256 == 0, Kp % 64 == 0 (the CTA-pair kernel); RUART_ERR_ARG otherwise.            */
RUART_API int ruart_gemm_bf16_fold(const void* A, long long lda, const void* W, long long ldw, int M, int N,
                                   int Kp, int fold, int epi, const float* vec, const float* vec2,
                                   const float* in_stats, float ln_eps, void* out_bf16, long long ldo,
                                   const void* residual_bf16, long long ld_res, float* out_stats,
                                   void* stream);

/* Fused query/key/value projection + self-attention of the folded bf16 encoder (BertSelfAttention.forward,
 * modeling.py:224-250): ctx = softmax(Q K^T / 8 + mask) V per sequence and head, with Q|K|V = LayerNorm(A) W0^T + b
 * computed tile by tile and never written to memory.  A = the stored pre-LayerNorm rows + in_stats (as fold 1 of
 * ruart_gemm_bf16_fold); W = (W0 * gamma) with rows permuted head-major [q_h / 8 ; k_h ; v_h] (192 rows per head),
 * vec / vec2 = W0 beta + b and colsum(W) permuted / scaled alike; tile_meta / tok_bounds from ruart_seq_tiles.
 * Needs M >= 2048, Kp == 768, every sequence <= 128 tokens.                                                   */
RUART_API int ruart_qkv_attention_fold(const void* A, long long lda, const void* W, long long ldw, int M, int Kp,
                                       int n_heads, const float* vec, const float* vec2, const float* in_stats,
                                       float ln_eps, const int32_t* tile_meta, const int32_t* tok_bounds,
                                       void* out_bf16, long long ldo, void* stream);
/* Row tiles for ruart_qkv_attention_fold from the sequence offsets cu_seq[S + 1] of the packed layout: whole
 * consecutive sequences, at most 128 rows per tile (greedy).  meta[0] = number of tiles n, meta[1 .. n] = first row
 * of each tile, meta[n + 1] = cu_seq[S]; capacity (ints) >= 2 * T / 128 + 8.  tok_bounds [T][2] = (first row, end
 * row) of every token's sequence.                                                                             */
RUART_API int ruart_seq_tiles(const int32_t* cu_seq, int S, int32_t* meta, int capacity, int32_t* tok_bounds,
                              void* stream);

/* ---------------------------------------------------------------- BERT (packed, pad-free rows)
 * Activations are [T, hidden] row-major over the T real wordpieces of all sequences; every
 * kernel accepts fp32 and/or bf16 pointers (exactly one input representation, any outputs).
 * bf16 outputs may be written as `out_parts` split parts ([T, parts*hidden]).                */

/* Folded-LayerNorm form of ruart_bert_embed_ln: the embedding sum BEFORE its LayerNorm (bf16 [T, hidden]) and
 * the row's (sum, sum of squares) in slot 0 of out_stats [T][8] float2 (slots 1-7 zero).                    */
RUART_API int ruart_bert_embed_raw(const int32_t* ids, const int32_t* pos, const float* word_emb,
                                   const float* pos_emb, const float* type_emb, int T, int hidden,
                                   void* out_bf16, float* out_stats, void* stream);
/* Coefficients of ruart_subword_avg_layers_fold: G[l][c] = softmax(alpha)_l * gamma * ln_gamma[l][c],
 * C[c] = sum_l softmax(alpha)_l * gamma * ln_beta[l][c]  (SDNet.linear_sum, SDNet.py:573-583, times each
 * layer's output LayerNorm weights).                                                                        */
RUART_API int ruart_subword_coef(const float* alpha, const float* gamma, int n_layers, const float* ln_gamma,
                                 const float* ln_beta, int hidden, float* G, float* C, void* stream);
/* ruart_subword_avg_layers over PRE-LayerNorm layer outputs + their partial sums (folded encoder):
 * dst[item, j] = sum_l G_l * mean_pieces(v r - mu r) + C.                                                   */
RUART_API int ruart_subword_avg_layers_fold(const void* h_bf16, long long layer_stride, const float* stats,
                                            long long stats_layer_stride, float ln_eps, const int32_t* words,
                                            int n_words, const int32_t* row_start, const uint8_t* x_mask, int W,
                                            float* dst, long long dst_stride, const float* G, const float* C,
                                            int n_layers, int hidden, void* stream);
/* BertEmbeddings.forward, modeling.py:185-199: LN(word[ids] + position[pos] + token_type[0]) */
RUART_API int ruart_bert_embed_ln(const int32_t* ids, const int32_t* pos, const float* word_emb,
                                  const float* pos_emb, const float* type_emb, const float* gamma,
                                  const float* beta, float eps, int T, int hidden, float* out_f32,
                                  void* out_bf16, int out_parts, void* stream);
/* BertSelfOutput / BertOutput tail, modeling.py:260-264,299-303: LN(x + residual); both residual
 * pointers NULL = x already contains the residual (fused into the GEMM epilogue)             */
RUART_API int ruart_add_layernorm(const float* x_f32, const void* x_bf16, const float* res_f32,
                                  const void* res_bf16, const float* gamma, const float* beta,
                                  float eps, int T, int hidden, float* out_f32, void* out_bf16,
                                  int out_parts, void* stream);
/* BertSelfAttention.forward core, modeling.py:229-250, on fused qkv [T, 3*hidden]; sequences are
 * rows cu_seqlens[s] .. cu_seqlens[s+1]; head_dim 64; max_len = longest sequence (host hint). */
RUART_API int ruart_bert_attention(const float* qkv_f32, const void* qkv_bf16,
                                   const int32_t* cu_seqlens, int n_seq, int n_heads, float scale,
                                   int max_len, float* out_f32, void* out_bf16, int out_parts,
                                   void* stream);
/* Bert.combine_forward's subword->word mean (Bert.py:149-165) fused with SDNet.linear_sum
 * (SDNet.py:573-583).  words = int32 [4][n_words]: item row, word slot j, st, ed (the reference's
 * x_bert_offset[item][j]); row_start[item] = packed index of the row's first token; x_mask =
 * uint8 [N, W] (words with 0 are skipped, Bert.py:155); the word's slot is
 * dst + (item*W + j)*dst_stride.  dst (+)= mean(h[st..ed)) * softmax(alpha)[layer] * gamma;
 * alpha == NULL means coefficient 1 (the per-layer outputs of plain Bert.forward).            */
RUART_API int ruart_subword_avg_accum(const float* h_f32, const void* h_bf16, const int32_t* words,
                                      int n_words, const int32_t* row_start, const uint8_t* x_mask,
                                      int W, float* dst, long long dst_stride, const float* alpha,
                                      int n_layers, const float* gamma, int layer, int first,
                                      int hidden, void* stream);
/* Same for all encoder layers at once: h is [n_layers][T][hidden] (layer_stride elements apart);
 * dst = sum_l mean(h_l[st..ed)) * softmax(alpha)[l] * gamma, written once, in layer order.    */
RUART_API int ruart_subword_avg_layers(const float* h_f32, const void* h_bf16,
                                       long long layer_stride, const int32_t* words, int n_words,
                                       const int32_t* row_start, const uint8_t* x_mask, int W,
                                       float* dst, long long dst_stride, const float* alpha,
                                       int n_layers, const float* gamma, int hidden, void* stream);
/* Sequence bookkeeping of the packed layout.  ruart_seq_lengths: per row of mask [N, L] the number of
 * real tokens (row_len[N]) and of each window of `window` columns (win_len[N * n_win]).
 * ruart_seq_scan (one CTA): exclusive prefix sums cu_rows[R+1], cu_seq[S+1] over all segments'
 * rows / windows, totals[0] = token count, totals[1+k] = longest window of segment k;
 * seg_row0 / seg_seq0 are HOST arrays of n_seg+1 boundaries (n_seg <= 8).                       */
RUART_API int ruart_seq_lengths(const uint8_t* mask, int N, int L, int window, int32_t* row_len,
                                int32_t* win_len, void* stream);
RUART_API int ruart_seq_scan(const int32_t* row_len, int R, const int32_t* win_len, int S, int n_seg,
                             const int32_t* seg_row0_host, const int32_t* seg_seq0_host,
                             int32_t* cu_rows, int32_t* cu_seq, int32_t* totals, int max_total,
                             void* stream);
/* Real (mask != 0) wordpieces of ids [N, L] (int64, the collate's dtype) -> packed int32 ids and
 * position ids (column % window) at out[row_start[r] ...]; replaces the padded [N, L] layout of
 * BertModel.forward's inputs (modeling.py:585-604).  `capacity` = length of out_ids / out_pos:
 * slots beyond it are not written.                                                            */
RUART_API int ruart_pack_tokens(const long long* ids, const uint8_t* mask, int N, int L,
                                const int32_t* row_start, int window, int32_t* out_ids,
                                int32_t* out_pos, int capacity, void* stream);
/* fp32 [*, K] (row pitch ld) -> bf16 split operand [rows, parts*Kp] for ruart_gemm_bf16; output
 * row r reads source row row_idx[r] (NULL = r)                                                 */
RUART_API int ruart_split_bf16(const float* src, long long ld, const int32_t* row_idx,
                               long long rows, int K, int Kp, int parts, void* dst, void* stream);

/* Same for the column-wise concatenation of n_src (<= 8) fp32 sources with equal row counts — `torch.cat(xs, 2)`
 * feeding an nn.Linear / nn.LSTM input (Layers.py:499-501,508,516; SDNet.py:350,380-390): the concatenation is
 * never materialised.  srcs/pitches/widths are HOST arrays of n_src entries; sum(widths) <= Kp.             */
RUART_API int ruart_split_concat_bf16(const float* const* srcs_host, const long long* pitches_host,
                                      const int* widths_host, int n_src, long long rows, int Kp, int parts,
                                      void* dst, void* stream);

/* ---------------------------------------------------------------- SDNet fusion stack (fp32)
 * Pitches are in floats.  Masks are uint8 (0 = pad), exactly the reference's ByteTensor /
 * bool masks.                                                                                */

/* dst[dst_idx[k]] = src[src_idx[k]] (D floats; optional second copy dst2): nn.Embedding lookups
 * (SDNet.py:447-492), pre-align pack/unpack (SDNet.py:504-520,540-550), slot scatter
 * (SDNet.py:300-318).  NULL index = identity; negative index = skip.                          */
RUART_API int ruart_gather_rows(const float* src, long long src_pitch, const void* src_idx,
                                float* dst, long long dst_pitch, const void* dst_idx, float* dst2,
                                long long dst2_pitch, long long n, int D, int idx_is_64,
                                void* stream);
/* F.layer_norm(x, x.size()) — one mean/var over the whole [rows, cols] block, in place
 * (Layers.py:167-168).  workspace: >= 2048 doubles.                                           */
RUART_API int ruart_whole_layernorm(float* x, long long rows, int cols, long long pitch, float eps,
                                    double* workspace, void* stream);
/* Attention.forward after the projections (Layers.py:272-288): out = softmax(mask(p1 p2^T)) x3.
 * split_parts = 2: both products on the tensor cores with every fp32 operand as hi + lo bf16 parts
 * (three terms per product, ~2^-16 relative, the rule of the 2-part split GEMMs; L2 <= 128);
 * any other value: fp32 CUDA-core kernel.                                                      */
RUART_API int ruart_attention_tail(const float* p1, long long p1_pitch, const float* p2,
                                   long long p2_pitch, int hidden, const uint8_t* mask,
                                   const float* x3, long long x3_pitch, int D3, float* out,
                                   long long out_pitch, int B, int L1, int L2, int add_to_out,
                                   int split_parts, void* stream);
/* The n_heads (<= 4) Attention modules of one DeepAttention call (Layers.py:493-524) in one launch: head z reads
 * columns [z hidden, (z+1) hidden) of the stacked projections p1 / p2 and x3s_host[z] ([B, L2, D3], common pitch),
 * and writes out[:, :, z D3 : (z+1) D3].  Tensor-core form only (L2 <= 128).                                  */
RUART_API int ruart_attention_tail_heads(const float* p1, long long p1_pitch, const float* p2,
                                         long long p2_pitch, int hidden, int n_heads, const uint8_t* mask,
                                         const float* const* x3s_host, long long x3_pitch, int D3, float* out,
                                         long long out_pitch, int B, int L1, int L2, void* stream);
/* LinearSelfAttn + weighted_avg (Layers.py:328-341,529-534): out[b] = softmax(mask(x w + b)) x  */
RUART_API int ruart_self_attn_pool(const float* x, long long x_pitch, int B, int L, int D,
                                   const uint8_t* mask, const float* w, const float* bias,
                                   float* out, long long out_pitch, void* stream);
/* GetFinalScores.forward (Layers.py:373-432) given wy = [attn | attn2 | noanswer_linear](h0)
 * [B, 3X]: probabilities [B, M+1] (and optional pre-softmax logits); nan_flag is set to 1 if any
 * probability is NaN (replaces the reference's isnan asserts, Layers.py:430,462,467).         */
RUART_API int ruart_final_scores(const float* x, long long x_pitch, int B, int M, int X,
                                 const float* wy, const uint8_t* mask, int es_len,
                                 const float* noans_w, const float* noans_b, float* probs,
                                 float* logits, int* nan_flag, void* stream);
/* One step of the step-synchronous uni-LSTM `multi2one` (SDNet.py:137,270-271,304,310).        */
RUART_API int ruart_lstm_cell(const float* gx, const int32_t* row_gx, const float* gh, float* c,
                              void* h_split, int parts, int Kp, int H, int n_rows,
                              const int32_t* last_step, int step, const long long* slot_off,
                              float* slots, void* stream);
/* Persistent (Bi)LSTM recurrence of StackedBRNN (Layers.py:137,166): xg = x W_ih^T + b_ih + b_hh
 * [B*L, ndir*4H] -> out[:, dir*H + j]; w_hh [ndir][4H][H]; H <= 128; pads are processed.
 * Runs on tensor cores (mma.sync on bf16 hi|lo splits of W_hh and h, ~1e-6 from the fp32 recurrence)
 * when xg is 16-byte aligned with xg_pitch % 4 == 0; otherwise the fp32 FMA kernel.              */
RUART_API int ruart_lstm_recurrence(const float* xg, long long xg_pitch, const float* w_hh,
                                    float* out, long long out_pitch, int B, int L, int H,
                                    int ndir, void* stream);
/* Same recurrence on the fp32 FMA kernel only (the fp32 mode of the stack, SDNET_precision 'fp32').   */
RUART_API int ruart_lstm_recurrence_f32(const float* xg, long long xg_pitch, const float* w_hh,
                                        float* out, long long out_pitch, int B, int L, int H,
                                        int ndir, void* stream);

/* Answer-index rule of SDNetTrainer.predict (SDNetTrainer.py:402-412): out_idx[b] = index of the
 * largest probability among {last column (if label_no_answer)} U {i < num_cnt[b] - 1}.          */
RUART_API int ruart_select_answers(const float* probs, const int32_t* num_cnt, int B, int M1,
                                   int label_no_answer, int32_t* out_idx, void* stream);

/* ---------------------------------------------------------------- training step, optimizer side
 * (SDNetTrainer.update, SDNetTrainer.py:363-365) over ONE flat fp32 buffer of all trainable parameters.
 * ruart_grad_sqnorm: *out_sq = sum g[i]^2 in double (deterministic; workspace >= 1024 doubles).
 * ruart_adamax_step: torch.optim.Adamax (no weight decay) on the gradient scaled by
 * min(1, max_norm / (sqrt(*grad_sq) + 1e-6)) — torch.nn.utils.clip_grad_norm_ — read from device
 * memory; grad_sq NULL or max_norm <= 0: no clipping.  `step` counts from 1.                    */
RUART_API int ruart_grad_sqnorm(const float* g, long long n, double* workspace, double* out_sq,
                                void* stream);
RUART_API int ruart_adamax_step(float* p, const float* g, float* exp_avg, float* exp_inf, long long n,
                                float lr, float beta1, float beta2, float eps, int step,
                                const double* grad_sq, float max_norm, void* stream);


/* ---------------------------------------------------------------- training step, differentiable path
 * (SURVEY.md §8 a-19).  The reference differentiates its forward with torch autograd
 * (SDNetTrainer.update, SDNetTrainer.py:337-362: forward, loss, `loss.backward()`); here every op of the
 * SDNet stack has a forward and a backward entry point, composed by ruart_b200/autograd_ops.py.  BERT is
 * locked (SDNet.py:91-94); only alphaBERT / gammaBERT receive a gradient from it.                        */

/* C[b] (+)= alpha * op(A[b]) op(B[b]), fp32 on the CUDA cores; op(A) is [M,K], op(B) is [K,N]; row-major with
 * leading dimensions ld*, batch strides stride_* (elements; 0 = shared operand).  The per-image products of
 * Attention.forward (Layers.py:237,244: x1_rep.bmm(x2_rep^T), alpha.bmm(x3)), BilinearSeqAttn (:459) and their
 * gradients.                                                                                               */
RUART_API int ruart_bmm_f32(const float* A, long long lda, long long stride_a, int trans_a, const float* B,
                            long long ldb, long long stride_b, int trans_b, float* C, long long ldc,
                            long long stride_c, int batch, int M, int N, int K, float alpha, int accumulate,
                            void* stream);
/* out = softmax over the last dim of x [B, L1, L2] with keys whose mask[b, j] == 0 set to -inf
 * (Layers.py:283-288; mask NULL = plain softmax, Layers.py:418); rows without a live key give NaN.  */
RUART_API int ruart_masked_softmax(const float* x, long long x_pitch, const uint8_t* mask, int B, int L1,
                                   int L2, float* out, long long out_pitch, void* stream);
/* dx = p * (dp - sum_j p_j dp_j) per row */
RUART_API int ruart_softmax_backward(const float* p, long long p_pitch, const float* dp, long long dp_pitch,
                                     float* dx, long long dx_pitch, long long rows, int cols, void* stream);
/* Element-wise helpers on pitched fp32 [rows, cols]: op 0 out = a*b; 1 out = a * (b > 0) (ReLU backward);
 * 2 out = a + b; 3 out = a * v[c] (v_len == 1: scalar) — the diagonal of AttentionScore (Layers.py:229-231);
 * 4 out = max(a, 0); 5 out = a + v[c].                                                                      */
RUART_API int ruart_eltwise(int op, const float* a, long long a_pitch, const float* b, long long b_pitch,
                            const float* v, int v_len, float* out, long long out_pitch, long long rows,
                            int cols, void* stream);
/* out[i] = mask[i] ? x[i] : value over n contiguous elements: `scores.data.masked_fill_(x_mask == 0, -inf)`
 * (Layers.py:283-284,339,426,463).                                                                          */
RUART_API int ruart_mask_fill(const float* x, const uint8_t* mask, long long n, float value, float* out,
                              void* stream);
/* out[c] (+)= sum_r x[r][c] (bias / diagonal gradients), deterministic; workspace >= 64 * cols doubles. */
RUART_API int ruart_colsum(const float* x, long long pitch, long long rows, int cols, double* workspace,
                           float* out, int accumulate, void* stream);
/* fp32 [rows, K] (row pitch ld) -> TRANSPOSED bf16 split operand dst [K, parts * rows_p] (rows_p a multiple
 * of 64 >= rows, zero padded): the operands of the weight gradient dW = dY^T X and of dX = dY W on
 * ruart_gemm_bf16.                                                                                        */
RUART_API int ruart_split_bf16_t(const float* src, long long ld, long long rows, int K, long long rows_p,
                                 int parts, void* dst, void* stream);
/* ruart_whole_layernorm that also writes (mean, rstd) to stats_out[2] */
RUART_API int ruart_whole_layernorm_stats(float* x, long long rows, int cols, long long pitch, float eps,
                                          double* workspace, float* stats_out, void* stream);
/* backward of F.layer_norm(x, x.size()) (Layers.py:167-168): y = normalised output, stats from the forward;
 * workspace >= 2048 doubles.                                                                              */
RUART_API int ruart_whole_layernorm_backward(const float* y, long long y_pitch, const float* dy,
                                             long long dy_pitch, long long rows, int cols,
                                             const float* stats, double* workspace, float* dx,
                                             long long dx_pitch, void* stream);
/* nn.Embedding weight gradient (the backward of SDNet.py:447-492's lookups): dW[v] (+)= sum of dy[k] over the
 * positions k with ids[k] == v, in an order that depends on the ids alone (deterministic, no floating-point
 * atomics).  All-zero gradient rows (the pad-word slots) and ids outside [0, V) are skipped.
 * workspace >= ruart_embedding_grad_workspace_bytes(n, V, D) selects the sorted form (counting sort of the live
 * positions by row, then one warp per 64 consecutive sorted positions: work ~ n, not V x n).  A smaller workspace
 * — at least round_up(n, 16) bytes, plus up to 64 * V * D floats of per-segment partial sums for V < 2048 — runs the
 * first form, one warp per vocabulary row scanning the id list in ascending k.                                  */
RUART_API long long ruart_embedding_grad_workspace_bytes(long long n, int V, int D);
RUART_API int ruart_embedding_grad(const void* ids, int idx_is_64, long long n, const float* dy,
                                   long long dy_pitch, int D, int V, uint8_t* workspace,
                                   long long workspace_bytes, float* dW, long long dw_pitch, int accumulate,
                                   void* stream);
/* Gradient of ruart_subword_avg_layers with respect to alpha [n_layers] and gamma [1] (the encoder itself is
 * locked): dy is the gradient of dst (same addressing); workspace >= 256 * n_layers doubles.              */
RUART_API int ruart_subword_layers_backward(const float* h_f32, const void* h_bf16, long long layer_stride,
                                            const int32_t* words, int n_words, const int32_t* row_start,
                                            const uint8_t* x_mask, int W, const float* dy,
                                            long long dy_stride, const float* alpha, int n_layers,
                                            const float* gamma, int hidden, double* workspace,
                                            float* dalpha, float* dgamma, int accumulate, void* stream);
/* ruart_lstm_recurrence that also saves the activated gates and cell state of every step:
 * gates [B*L, ndir*5*H] (columns dir*5H + {i,f,g,o,c}*H + j).                                            */
RUART_API int ruart_lstm_recurrence_train(const float* xg, long long xg_pitch, const float* w_hh,
                                          float* out, long long out_pitch, int B, int L, int H,
                                          int ndir, float* gates, long long gates_pitch, void* stream);
/* Back-propagation through time: dout = gradient of `out`; dxg [B*L, ndir*4H] = gradient of the gate
 * pre-activations (= of xg); dW_hh, dW_ih, db and dx follow from dxg by GEMMs / column sums.             */
RUART_API int ruart_lstm_recurrence_backward(const float* gates, long long gates_pitch, const float* w_hh,
                                             const float* dout, long long dout_pitch, float* dxg,
                                             long long dxg_pitch, int B, int L, int H, int ndir,
                                             void* stream);
/* ruart_lstm_cell (gx rows contiguous for the step) that also saves (i,f,g,o,c,h) per row: save [n_rows, 6H] */
RUART_API int ruart_lstm_cell_train(const float* gx, const float* gh, float* c, void* h_split, int parts,
                                    int Kp, int H, int n_rows, const int32_t* last_step, int step,
                                    const long long* slot_off, float* slots, float* save, void* stream);
/* Backward of one multi2one step: dgates [n_rows, 4H]; dc_carry [n_all, H] carried between steps (zeroed by
 * the caller); dh_rec [n_rec, H] = dgates(step+1) W_hh, NULL at the last step; save_prev NULL at step 0.  */
RUART_API int ruart_lstm_cell_backward(const float* save, const float* save_prev, const float* dslots,
                                       const long long* slot_off, const int32_t* last_step, int step,
                                       const float* dh_rec, int n_rec, float* dc_carry, float* dgates, int H,
                                       int n_rows, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RUART_B200_H_ */

/*
 * mpformer_b200 -- C ABI of the B200-native (sm_100a) MP-Former hot path.
 *
 * Drop-in boundary (SURVEY.md §8 b1'): plain pointers + sizes + a cudaStream_t passed as void*.
 * No torch types cross this boundary.  Every function
 *   - returns 0 on success, a negative MPF_ERR_* for argument errors, or a positive cudaError_t;
 *     mpf_last_error() gives a thread-local human-readable message;
 *   - never allocates device memory, never synchronises, launches on `stream`;
 *   - borrows all pointers for the duration of the asynchronous work (caller keeps them alive).
 * All pointers are DEVICE pointers unless the parameter name ends in `_host`.
 *
 * The reference interfaces each entry point replaces are cited as
 *   ref: <path under /root/reference>:<lines>.
 */
#ifndef MPFORMER_B200_H_
#define MPFORMER_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPF_ABI_VERSION 1

#define MPF_OK 0
#define MPF_ERR_BAD_ARG (-1)
#define MPF_ERR_UNSUPPORTED (-2)
#define MPF_ERR_NO_DEVICE (-3)

/* ABI version of the loaded library (== MPF_ABI_VERSION it was built with). */
int mpf_abi_version(void);
/* Thread-local message describing the last non-zero return on this thread ("" if none). */
const char* mpf_last_error(void);
/* Number of kernels this library has launched since load (all threads); used by bench.py for
 * the `gpu_launches` claim. */
uint64_t mpf_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Multi-scale deformable attention, forward.
 *   out[b,q,m,:] = sum_{l,p} attn_weight[b,q,m,l,p] *
 *                  bilinear_zero_pad(value[b, level l, :, m, :], sampling_loc[b,q,m,l,p,:])
 * with pixel coordinates (w_im, h_im) = loc * (W_l, H_l) - 0.5.
 *
 * ref: mask2former/modeling/pixel_decoder/ops/src/ms_deform_attn.h:25-45 (ms_deform_attn_forward),
 *      .../src/cuda/ms_deform_attn_cuda.cu:25-85, .../src/cuda/ms_deform_im2col_cuda.cuh:38-89,242-304
 *
 *   value             [batch, spatial_size, num_heads, channels]           contiguous
 *   spatial_shapes    [num_levels, 2] int64 (H_l, W_l)    DEVICE array, as in the reference
 *   level_start_index [num_levels]    int64               DEVICE array, as in the reference
 *   sampling_loc      [batch, num_query, num_heads, num_levels, num_point, 2]  (x, y) in [0,1]
 *   attn_weight       [batch, num_query, num_heads, num_levels, num_point]
 *   out               [batch, num_query, num_heads*channels]  fully overwritten (no pre-zero needed)
 * The reference's `im2col_step` batching is an implementation detail of its launcher
 * (cuda.cu:55-80) and has no effect on results; it is accepted and validated
 * (batch % min(batch, im2col_step) == 0) by the Python binding, not here.
 * ------------------------------------------------------------------------------------------- */
int mpf_msda_forward_f32(const float* value, const int64_t* spatial_shapes,
                         const int64_t* level_start_index, const float* sampling_loc,
                         const float* attn_weight, int batch, int spatial_size, int num_heads,
                         int channels, int num_levels, int num_query, int num_point, float* out,
                         void* stream);
/* Same as mpf_msda_forward_f32 plus an optional HOST copy of spatial_shapes ([num_levels,2] int64,
 * may be NULL).  When given and num_query == spatial_size the launcher orders the queries as
 * 16x8 spatial tiles per level (better L1 locality); results are identical either way. */
int mpf_msda_forward_f32_ex(const float* value, const int64_t* spatial_shapes,
                            const int64_t* level_start_index, const float* sampling_loc,
                            const float* attn_weight, int batch, int spatial_size, int num_heads,
                            int channels, int num_levels, int num_query, int num_point, float* out,
                            const int64_t* spatial_shapes_host, void* stream);
int mpf_msda_forward_f64(const double* value, const int64_t* spatial_shapes,
                         const int64_t* level_start_index, const double* sampling_loc,
                         const double* attn_weight, int batch, int spatial_size, int num_heads,
                         int channels, int num_levels, int num_query, int num_point, double* out,
                         void* stream);

/* Encoder-fused form: consumes the raw output of the offsets+logits projection and the reference points;
 * the softmax over the L*P logits and  loc = ref + off / (W_l, H_l)  (ref ops/modules/ms_deform_attn.py:
 * 102-109) are evaluated inside the kernel.
 *   offsets_logits [batch, num_query, M*L*P*3]: first M*L*P*2 offsets laid out (m, l, p, xy), then M*L*P
 *                  attention logits laid out (m, l, p)  (= cat(sampling_offsets(q), attention_weights(q)))
 *   reference_points [*, num_query, L, 2] with batch stride ref_batch_stride elements (0 = shared)
 * Supported: P == 4, channels in {16, 32, 64}, L <= 4; otherwise MPF_ERR_UNSUPPORTED (use the plain form).
 * The backward returns grad_value (zero-filled, then accumulated) and grad_offsets_logits (overwritten). */
int mpf_msda_enc_forward_f32(const float* value, const int64_t* spatial_shapes,
                             const int64_t* level_start_index, const float* offsets_logits,
                             const float* reference_points, long long ref_batch_stride, int batch,
                             int spatial_size, int num_heads, int channels, int num_levels, int num_query,
                             int num_point, float* out, const int64_t* spatial_shapes_host, void* stream);
int mpf_msda_enc_backward_f32(const float* grad_out, const float* value, const int64_t* spatial_shapes,
                              const int64_t* level_start_index, const float* offsets_logits,
                              const float* reference_points, long long ref_batch_stride, int batch,
                              int spatial_size, int num_heads, int channels, int num_levels, int num_query,
                              int num_point, float* grad_value, float* grad_offsets_logits,
                              const int64_t* spatial_shapes_host, void* stream);

/* Selects the kernels behind mpf_msda_enc_*: 1 = TMA-staged feature tiles in shared memory + on-chip
 * fixed-point accumulation of grad_value (csrc/msda_staged.cu) whenever the launch is the encoder's self-attention
 * over a pixel grid (num_query == spatial_size, host shapes given, 32 channels per head, 4 points, <= 4 levels);
 * 0 (default: measured faster on the B200, profiles/r2e_msda_enc_probe.jsonl) = always the L1-gather / per-corner
 * global-reduction kernels (csrc/msda.cu).  enabled < 0 only queries.
 * Returns the previous setting.  Results agree to fp32 rounding (forward: bit for bit).  Same semantics as the
 * reference op either way (ref ops/src/cuda/ms_deform_im2col_cuda.cuh:242-304, :92-164).  Process-wide; meant for
 * A/B measurements and tests (also: environment MPF_MSDA_STAGED=1 at load time). */
/* CTA-pair mode of mpf_gemm_bf16x3 / mpf_gemm_bf16x3_relubits / mpf_conv3x3_cl_bf16x3 (tcgen05 cta_group::2: clusters of two
 * CTAs execute one M = 256 UMMA, each staging half of the B tile): 0 (default) = never, 1 = when a launch has at least
 * one full wave of 256-row pair tiles, 2 = whenever the problem has two M tiles (tests).  mode < 0 only queries.
 * Returns the previous mode.  Results are identical in every mode (same products, same accumulation order per
 * element); measured SLOWER than single CTAs with the 3-MMA split arithmetic, hence off by default.
 * Also: environment MPF_GEMM_PAIR at load time. */
int mpf_gemm_bf16x3_set_pair_mode(int mode);

int mpf_msda_set_staged(int enabled);

/* Per-row top-k selection with payload gather: for every row r of scores [rows, n] (fp32, device) the k largest
 * entries are selected and payload[r, i, 0:payload_width] of the selected i is written to out [rows, k,
 * payload_width] in ascending order of i (deterministic; ties at the threshold go to the lower index).  Replaces
 * torch.topk + gather in PointRend's importance sampling of the criterion -- ref mask2former/modeling/criterion.py:
 * 165-172 (detectron2 get_uncertain_point_coords_with_randomness): only the SET of the k most uncertain points is
 * needed, so a 4-pass radix select over keys held in shared memory replaces a full segmented sort.
 * Limits: 0 < k <= n <= 49152, payload_width 1 or 2.  NaN scores rank above +inf (torch.topk's convention). */
int mpf_topk_gather_rows_f32(const float* scores, int rows, int n, int k, const float* payload, int payload_width,
                             float* out, void* stream);

/* Point-sampled mask losses of the criterion for `rows` matched (prediction, target) pairs at once -- ref
 * mask2former/modeling/criterion.py:25-43 (dice_loss) and :51-68 (sigmoid_ce_loss), called from loss_masks
 * (criterion.py:188-189) once per prediction head; here the rows of all heads go through one launch:
 *   bce[r]  = mean_p BCEWithLogits(x[r, p], y[r, p])
 *   dice[r] = 1 - (2 sum_p s y + 1) / (sum_p s + sum_p y + 1),  s = sigmoid(x[r, p])
 * x, y [rows, points] fp32 (device, contiguous); stats [rows, 2] keeps (numerator, denominator) of the dice ratio for
 * the backward, which writes gx[r, p] = g_bce[r] * d bce[r]/dx + g_dice[r] * d dice[r]/dx (every element: no
 * atomics, deterministic).  The division by num_masks and the sum over rows stay with the caller. */
int mpf_mask_loss_rows_fwd_f32(const float* x, const float* y, int rows, int points, float* bce, float* dice,
                               float* stats, void* stream);
int mpf_mask_loss_rows_bwd_f32(const float* x, const float* y, const float* stats, const float* g_bce,
                               const float* g_dice, int rows, int points, float* gx, void* stream);

/* Self-attention core of the decoder's SelfAttentionLayer (ref transformer_decoder/
 * mask2former_transformer_decoder.py:42-52; tgt_mask of the mask-piloted groups: decoder :1051-1059):
 *   out[b, :, h] = softmax(q_h k_h^T / sqrt(head_dim) + mask) v_h   with q | k | v = qkv[b, :, 0:E | E:2E | 2E:3E]
 * qkv [B, Qt, 3E] fp32 (the packed in-projection output), mask uint8 [Qt, Qt] or null (1 = not allowed, shared by
 * images and heads), out [B, Qt, E], lse [B, heads, Qt] (natural-log log-sum-exp of the scaled scores, kept for the
 * backward).  head_dim == 32, Qt <= 320.  The backward recomputes the probabilities from lse and writes every element
 * of d_qkv [B, Qt, 3E] (no atomics: deterministic). */
int mpf_self_attn_fwd_f32(const float* qkv, const uint8_t* mask, int B, int Qt, int heads, int head_dim, float* out,
                          float* lse, void* stream);
int mpf_self_attn_bwd_f32(const float* qkv, const uint8_t* mask, const float* d_out, const float* lse, int B, int Qt,
                          int heads, int head_dim, float* d_qkv, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Multi-scale deformable attention, backward.
 * ref: .../src/ms_deform_attn.h:47-66 (ms_deform_attn_backward), .../cuda/ms_deform_attn_cuda.cu:88-158,
 *      .../cuda/ms_deform_im2col_cuda.cuh:92-164 (col2im_bilinear), :306-925 (kernel family)
 *
 *   grad_out          [batch, num_query, num_heads*channels]
 *   grad_value        same shape as value        -- zero-filled by the callee, then accumulated
 *                                                    with fp atomics (order non-deterministic,
 *                                                    as in the reference)
 *   grad_sampling_loc same shape as sampling_loc -- fully overwritten
 *   grad_attn_weight  same shape as attn_weight  -- fully overwritten
 * ------------------------------------------------------------------------------------------- */
int mpf_msda_backward_f32(const float* grad_out, const float* value, const int64_t* spatial_shapes,
                          const int64_t* level_start_index, const float* sampling_loc,
                          const float* attn_weight, int batch, int spatial_size, int num_heads,
                          int channels, int num_levels, int num_query, int num_point,
                          float* grad_value, float* grad_sampling_loc, float* grad_attn_weight,
                          void* stream);
int mpf_msda_backward_f32_ex(const float* grad_out, const float* value,
                             const int64_t* spatial_shapes, const int64_t* level_start_index,
                             const float* sampling_loc, const float* attn_weight, int batch,
                             int spatial_size, int num_heads, int channels, int num_levels,
                             int num_query, int num_point, float* grad_value,
                             float* grad_sampling_loc, float* grad_attn_weight,
                             const int64_t* spatial_shapes_host, void* stream);
int mpf_msda_backward_f64(const double* grad_out, const double* value,
                          const int64_t* spatial_shapes, const int64_t* level_start_index,
                          const double* sampling_loc, const double* attn_weight, int batch,
                          int spatial_size, int num_heads, int channels, int num_levels,
                          int num_query, int num_point, double* grad_value,
                          double* grad_sampling_loc, double* grad_attn_weight, void* stream);

/* ---------------------------------------------------------------------------------------------
 * fp32-accurate tensor-core GEMM ("3xTF32", tcgen05 + TMA):
 *   C[b] = A[b] * B[b]^T (+ bias[N]) (ReLU optional), A [batch,M,K], B [batch,N,K], K contiguous.
 * B (the small operand: an nn.Linear weight [out,in] or the per-query mask embedding) is passed
 * pre-split by mpf_split_tf32 (hi = rn_tf32(x), lo = rn_tf32(x - hi)).
 * transpose_c == 0: C[b][m*ldc + n];  transpose_c != 0: C[b][n*ldc + m].
 * Replaces the library calls at: ref decoder :1865 (einsum "bqc,bchw->bqhw" with mask_features kept
 * channels-last: M = H*W, N = Q, K = C, transposed store), ref decoder :105-108 (K/V/Q in-projections
 * of nn.MultiheadAttention), ref ops/modules/ms_deform_attn.py:98,102-103,124 and
 * pixel_decoder/msdeformattn.py:116-120 (Linear layers of the encoder).
 * Requirements: K % 32 == 0; A/B 16-byte aligned, lda/ldb/batch strides multiples of 4 elements.
 * ------------------------------------------------------------------------------------------- */
int mpf_split_tf32(const float* x, float* hi, float* lo, long long n, void* stream);
int mpf_gemm_tf32x3(const float* A, long long lda, long long a_batch_stride, const float* B_hi,
                    const float* B_lo, long long ldb, long long b_batch_stride, const float* bias,
                    float* C, long long ldc, long long c_batch_stride, int batch, int M, int N, int K,
                    int relu, int transpose_c, void* stream);

/* Extended form of mpf_gemm_tf32x3:  x = (A*B^T + bias + resid) * alpha, optional ReLU.
 *   resid : optional addend with row stride resid_ld; output row r uses resid row (r % resid_rows)
 *           (resid_rows == 0: row r) and only columns < resid_cols receive it (0 = all).  Used to add the
 *           batch-independent term pos*Wk^T + bk to the key projection (K = (memory+pos) Wk^T + bk,
 *           ref decoder :105-107) without materialising memory+pos.
 *   C_lo  : if non-NULL the result is emitted pre-split for a following 3xTF32 consumer:
 *           C = rn_tf32(x), C_lo = rn_tf32(x - C)  (same layout as C). */
int mpf_gemm_tf32x3_ex(const float* A, long long lda, long long a_batch_stride, const float* B_hi,
                       const float* B_lo, long long ldb, long long b_batch_stride, const float* bias,
                       float* C, float* C_lo, long long ldc, long long c_batch_stride, const float* resid,
                       long long resid_ld, int resid_rows, int resid_cols, float alpha, int batch, int M,
                       int N, int K, int relu, int transpose_c, void* stream);

/* General form: either operand may be "MN-major" (a_mn_major / b_mn_major != 0), i.e. stored
 * [batch, K rows, M-or-N contiguous] -- the transpose of the default [batch, M-or-N rows, K contiguous] --
 * and B may be plain fp32 (B_lo == NULL: it is split into TF32 halves in shared memory like A).
 * With both operands MN-major this computes dW = dY^T X (weight gradient of an nn.Linear, reduction over
 * the token dimension; the batch dimension then enumerates K-splits whose partial products the caller
 * sums) and dF = dOut^T E of the mask-logit einsum without any transposed copy.  K need not be a
 * multiple of 32 (the tail is zero-filled by TMA).  lda / ldb are the row strides of the stored layout.
 * k_splits > 1 cuts the reduction into k_splits ranges handled by different CTAs; C must then hold
 * batch*k_splits slabs ([batch, k_splits, M, N]) of partial sums which the caller adds (no bias / ReLU /
 * residual / scaling in that mode).  gate (optional, [M, gate_ld], batch 1): the result is zeroed where
 * gate[m][n] <= 0 -- the ReLU backward of the layer below fused into this layer's input-gradient GEMM. */
int mpf_gemm_tf32x3_general(const float* A, int a_mn_major, long long lda, long long a_batch_stride,
                            const float* B, const float* B_lo, int b_mn_major, long long ldb,
                            long long b_batch_stride, const float* bias, float* C, float* C_lo,
                            long long ldc, long long c_batch_stride, const float* resid, long long resid_ld,
                            int resid_rows, int resid_cols, const float* gate, long long gate_ld, float alpha,
                            int batch, int M, int N, int K, int k_splits, int relu, int transpose_c,
                            void* stream);

/* ---------------------------------------------------------------------------------------------
 * "bf16x3" tensor-core GEMM (tcgen05.mma.kind::f16 + TMA loads AND TMA stores), the default for the same call
 * sites as mpf_gemm_tf32x3* when both operands are K-major:
 *   x = (A*B^T + bias + resid) * alpha, optional ReLU / gate;  C (and C_lo) as in mpf_gemm_tf32x3_general.
 * A [batch,M,K] is fp32 and is split into bf16 halves (hi = bf16_rn(x), lo = bf16_rn(x - hi)) inside the kernel;
 * B [batch,N,K] is passed pre-split by mpf_split_bf16 (b_batch_stride == 0: one B shared by the whole batch).
 * D += A_lo*B_hi + A_hi*B_lo + A_hi*B_hi keeps ~2^-16 relative error per product (contract: 1e-3) at twice the
 * tensor rate and half the operand bytes of 3xTF32.  C_lo (optional) still receives TF32 halves of the result
 * (consumed by the attention kernels).
 * Requirements: A/B/C 16-byte aligned; lda, ldc, batch strides multiples of 4 elements; ldb multiple of 8.
 * ------------------------------------------------------------------------------------------- */
int mpf_split_bf16(const float* x, uint16_t* hi, uint16_t* lo, long long n, void* stream);

/* All weight operands of a step in one launch: `table` (device) holds n_entries records
 *   { const float* src; bf16* hi; bf16* lo; int32 rows, cols; int64 ld; int32 transposed; int32 tile0; }   (48 bytes)
 * -- source [rows, cols] fp32 with row stride ld; destination halves [rows, cols] or, transposed, [cols, rows]
 * (contiguous bf16); tile0 = index of the record's first 32 x 32 tile in the grid of total_tiles tiles (records sorted by
 * tile0).  Replaces one mpf_split_bf16 (plus a transposing copy for the input-gradient operand) per use of every
 * nn.Linear weight of the path: ref pixel_decoder/msdeformattn.py:116-131, transformer_decoder/
 * mask2former_transformer_decoder.py:19-180 (the weights of self-attention, cross-attention, FFN and the heads). */
int mpf_split_weights_f32(const void* table, int n_entries, int total_tiles, void* stream);
/* C = A B^T (+ bias, + full residual, ReLU) for one [M, K] x [N, K] product with the ReLU pattern kept as ONE BIT per
 * element: relu_bits_out (with relu != 0) receives word [n / 32][m] whose bit n % 32 says C[m][n] > 0; gate_bits zeroes
 * the elements of C whose bit is clear (the backward of Linear -> ReLU -> Linear reads 1/32 of the bytes of the
 * activation: ref pixel_decoder/msdeformattn.py:116-120 under autograd).  N % 32 == 0. */
int mpf_gemm_bf16x3_relubits(const float* A, long long lda, const uint16_t* B_hi, const uint16_t* B_lo, long long ldb,
                             const float* bias, float* C, long long ldc, const float* resid, long long resid_ld,
                             int M, int N, int K, int relu, uint32_t* relu_bits_out, const uint32_t* gate_bits,
                             void* stream);
/* x [batch, R, Cc] fp32 -> bf16 halves of its transpose, hi / lo [batch, Cc, R] (R, Cc multiples of 4): the B operand
 * of a product that reduces over R when x is stored R-major (mask_features tokens in the batched dE of the
 * prediction heads, ref decoder :1865 under autograd). */
int mpf_transpose_split_bf16(const float* x, uint16_t* hi, uint16_t* lo, int batch, long long R, int Cc, void* stream);
int mpf_gemm_bf16x3(const float* A, long long lda, long long a_batch_stride, const uint16_t* B_hi,
                    const uint16_t* B_lo, long long ldb, long long b_batch_stride, const float* bias, float* C,
                    float* C_lo, long long ldc, long long c_batch_stride, const float* resid, long long resid_ld,
                    int resid_rows, int resid_cols, const float* gate, long long gate_ld, float alpha, int batch,
                    int M, int N, int K, int k_splits, int relu, int transpose_c, void* stream);

/* "TN" form in bf16x3: C[b] = A[b]^T * B[b] with A [batch, T, M] and B [batch, T, N] fp32 row-major (reduction over
 * the leading token dimension T) -- the weight gradient dW = dY^T X of an nn.Linear and dF = dOut^T E of the
 * mask-logit einsum (ref decoder :1865), without transposed copies: both operands are split into bf16 halves
 * inside the kernel.  k_splits > 1 cuts T into ranges handled by different CTAs; C then holds batch*k_splits
 * partial slabs [batch, k_splits, M, N] that the caller sums.  lda/ldb/ldc and batch strides: multiples of 4. */
int mpf_gemm_bf16x3_tn(const float* A, long long lda, long long a_batch_stride, const float* B, long long ldb,
                       long long b_batch_stride, float* C, long long ldc, long long c_batch_stride, int batch, int M,
                       int N, int T, int k_splits, void* stream);
/* Same product; accumulate != 0 adds it into C (TMA reduce-add, C must hold valid data -- e.g. zeros) instead of
 * overwriting, and then C is [batch, M, N] for ANY k_splits: the K-splits of a batch entry add into the same slab
 * (fp32 additions in arrival order, like the atomics of the MSDeformAttn backward).  Used for the weight gradients
 * (no partial-slab reduction kernel) and for the mask-feature gradient summed over the prediction heads
 * (ref decoder :1865 called at :1767,:1797). */
int mpf_gemm_bf16x3_tn_ex(const float* A, long long lda, long long a_batch_stride, const float* B, long long ldb,
                          long long b_batch_stride, float* C, long long ldc, long long c_batch_stride, int batch, int M,
                          int N, int T, int k_splits, int accumulate, void* stream);
/* mpf_gemm_bf16x3_tn with the column sums of A over each slab's tokens written to colsum (slab s at colsum +
 * s * colsum_slab_stride floats, M values; e.g. right behind slab s of C so that one reduction sums both): the bias
 * gradient sum_t dY[t, :] that accompanies every weight gradient dW = dY^T X of an nn.Linear (autograd of F.linear at
 * ref pixel_decoder/msdeformattn.py:116-131 and decoder :19-180), taken from the registers of the converter warps
 * instead of a second pass over dY. */
int mpf_gemm_bf16x3_tn_colsum(const float* A, long long lda, long long a_batch_stride, const float* B, long long ldb,
                              long long b_batch_stride, float* C, long long ldc, long long c_batch_stride, int batch,
                              int M, int N, int T, int k_splits, float* colsum, long long colsum_slab_stride,
                              void* stream);


/* ---------------------------------------------------------------------------------------------
 * Row-wise kernels of the encoder / decoder layers.
 *   mpf_add_layernorm_fwd_f32:  y = LayerNorm(x + r) * gamma + beta over the last dimension (C = 128, 256 or 512;
 *     r may be NULL), also writing the per-row mean and 1/sqrt(var + eps) for the backward.
 *     ref: pixel_decoder/msdeformattn.py:125-126,129 (norm1 / norm2 of the encoder layer), decoder :52,:112,:169.
 *   mpf_add_layernorm_bwd_f32:  dx (= gradient of both x and r) and per-CTA partial sums
 *     partial[p][0][c] = sum dy*xhat, partial[p][1][c] = sum dy, partial[p][2][c] = sum dx
 *     for p < mpf_add_layernorm_partials(rows)  (the caller adds the partials: dgamma, dbeta, and the bias
 *     gradient of the Linear layer that produced r).
 *   mpf_colsum_f32:  out[c] = sum_rows x[row*ld + c]  (bias gradients of the Linear layers; out is zeroed here).
 * ------------------------------------------------------------------------------------------- */
int mpf_add_layernorm_partials(long long rows);
int mpf_add_layernorm_fwd_f32(const float* x, const float* r, const float* gamma, const float* beta, float eps,
                              long long rows, int C, float* y, float* mean, float* rstd, void* stream);
int mpf_add_layernorm_bwd_f32(const float* dy, const float* x, const float* r, const float* gamma, const float* mean,
                              const float* rstd, long long rows, int C, float* dx, float* partial, void* stream);
int mpf_colsum_f32(const float* x, long long rows, int C, long long ld, float* out, void* stream);
/* GroupNorm over channels-last maps x [batch, HW, C] (groups of C/groups consecutive channels, 4..32 channels per
 * group), optionally fused with the ReLU that follows it.  ref: pixel_decoder/msdeformattn.py:216-219 (input
 * projections), :262-275 (lateral / output convs, norm "GN").  stats_ws: 2*batch*groups doubles of scratch (zeroed
 * here).  Forward also writes mean / rstd [batch, groups] for the backward; the backward returns dx and
 * dgamma_dbeta [2, C] (zeroed here, accumulated with fp32 atomics); with relu != 0 the gate is recomputed from x. */
int mpf_groupnorm_cl_fwd_f32(const float* x, const float* gamma, const float* beta, float eps, int batch,
                             long long HW, int C, int groups, int relu, float* y, float* mean, float* rstd,
                             double* stats_ws, void* stream);
int mpf_groupnorm_cl_bwd_f32(const float* dy, const float* x, const float* gamma, const float* beta,
                             const float* mean, const float* rstd, int batch, long long HW, int C, int groups,
                             int relu, float* dx, float* dgamma_dbeta, double* stats_ws, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Boolean stage of the prediction heads, bit-packed:
 *   bits[row][j] bit i = ( sigmoid( bilinear_resize(logits[row], (h,w), align_corners=False) )[32j+i] < 0.5 )
 * ref: transformer_decoder/mask2former_transformer_decoder.py:1869-1875 (F.interpolate, sigmoid, < 0.5;
 * the reference then repeats the map over the 8 heads -- here one bit serves all heads).
 *   logits [rows, H, W] fp32 with row stride `row_stride` elements; bits [rows, words_per_row] uint32,
 *   1 = masked; keys >= h*w (padding up to 32*words_per_row) are set to 1.
 * mpf_pack_bool_bits packs an existing bool map (uint8 0/1, [rows, n]) the same way (used for the
 * mask-piloted GT masks, ref decoder :986, :1038-1039).
 * ------------------------------------------------------------------------------------------- */
int mpf_attn_mask_bits_f32(const float* logits, long long row_stride, int rows, int H, int W, int h, int w,
                           uint32_t* bits, int words_per_row, void* stream);
int mpf_pack_bool_bits(const uint8_t* src, int rows, int n, uint32_t* bits, int words_per_row,
                       void* stream);
/* Mask-piloted (DN) attention masks straight from the ground-truth instance masks:
 *   bits[row] bit of cell (i,j) = ( F.interpolate(masks[row].float(), (h,w), mode="area")[i,j] <= 1e-8 )
 * ref: decoder :986-987 (prepare_for_dn_v5), :1593-1594 (gen_mask_dn).  masks [rows, H, W] uint8/bool (0/1);
 * the area mean of a 0/1 window is <= 1e-8 exactly when the window holds no set pixel (window area < 1e8). */
int mpf_gt_mask_area_bits(const uint8_t* masks, int rows, int H, int W, int h, int w, uint32_t* bits,
                          int words_per_row, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused masked multi-head cross-attention, forward (tcgen05 + TMA, 3xTF32):
 *   out[b,q,h*32:(h+1)*32] = softmax_k( Q_h[b,q,:] . K_h[b,k,:]  (+ -inf where masked) ) V_h[b,k,:]
 * ref: decoder :100-112 (nn.MultiheadAttention(query, key=memory+pos, value=memory, attn_mask)),
 *      decoder :1780 (rows whose keys are all masked attend everywhere -> row_open).
 *   q_hi/q_lo  [B, Qt, heads*32]  query projection, pre-scaled by log2(e)/sqrt(32), pre-split
 *   k_hi/k_lo  [B, HW, heads*32]  key projection, pre-split
 *   vt_hi/vt_lo[B, heads*32, HW]  value projection TRANSPOSED, pre-split
 *   mask_bits  [B, Qt, mask_words] (mask_words >= 2*ceil(HW/64)), 1 = masked
 *   row_open   [B, Qt] uint8, 1 = ignore the mask for this row (may be NULL)
 *   out        [B, Qt, heads*32];  lse2 [B, heads, Qt] = log2-sum-exp2 of the scaled scores (may be NULL)
 * Requirements: head_dim == 32, HW % 4 == 0.
 * ------------------------------------------------------------------------------------------- */
int mpf_masked_xattn_fwd_f32(const float* q_hi, const float* q_lo, const float* k_hi, const float* k_lo,
                             const float* vt_hi, const float* vt_lo, const uint32_t* mask_bits,
                             const uint8_t* row_open, float* out, float* lse2, int B, int Qt, int HW,
                             int heads, int head_dim, int mask_words, void* stream);

/* Backward of mpf_masked_xattn_fwd_f32 (two tcgen05 kernels, probabilities recomputed from lse2):
 *   dq[b,q,:]  = (1/sqrt(d)) * sum_k dS[q,k] K[k,:]      (gradient wrt the UNSCALED query projection)
 *   dk[b,k,:]  = ln2 * sum_q dS[q,k] Qs[q,:]             (Qs = the pre-scaled query the forward consumed)
 *   dv[b,k,:]  = sum_q P[q,k] dO[q,:]
 * with P = exp2(Qs K^T - lse2) (0 where masked), dS = P * (dO V^T - delta), delta = rowsum(dO * O).
 * Operands are TF32-split halves: q/qt (Qs and its transpose [B,E,qt_ld]), k/kt (K and K^T [B,E,HW]),
 * v (V, row-major [B,HW,E]), do/dot (dO and dO^T [B,E,qt_ld]); lse2, delta [B,heads,Qt].
 * ref: autograd backward of nn.MultiheadAttention in CrossAttentionLayer (decoder :100-112). */
int mpf_masked_xattn_bwd_f32(const float* q_hi, const float* q_lo, const float* qt_hi, const float* qt_lo,
                             const float* k_hi, const float* k_lo, const float* kt_hi, const float* kt_lo,
                             const float* v_hi, const float* v_lo, const float* do_hi, const float* do_lo,
                             const float* dot_hi, const float* dot_lo, const uint32_t* mask_bits,
                             const uint8_t* row_open, const float* lse2, const float* delta, float* dq,
                             float* dk, float* dv, int B, int Qt, int qt_ld, int HW, int heads, int head_dim,
                             int mask_words, void* stream);

/* Key-split variants of the two entry points above for small batches (B * heads CTAs do not fill 148 SMs, e.g. the
 * 2 images per GPU of BASELINE configs[2]): the key tiles of every (query tile, head, image) are divided over
 * `key_splits` CTAs.  Forward: each CTA leaves its unnormalised output and (max, sum) in ws_o [key_splits, B, Qt, E] /
 * ws_ml [key_splits, B, heads, Qt, 2]; a log-sum-exp merge kernel writes out / lse2.  Backward: the dQ kernel writes
 * partial sums to ws_dq [key_splits, B, Qt, E], added in a fixed order (deterministic); dK / dV are key-parallel
 * already.  key_splits == 1 is exactly the unsplit kernel (workspaces may be null).  Same reference semantics:
 * transformer_decoder/mask2former_transformer_decoder.py:100-112, :1780. */
int mpf_masked_xattn_fwd_f32_ex(const float* q_hi, const float* q_lo, const float* k_hi, const float* k_lo,
                                const float* vt_hi, const float* vt_lo, const uint32_t* mask_bits,
                                const uint8_t* row_open, float* out, float* lse2, int B, int Qt, int HW, int heads,
                                int head_dim, int mask_words, int key_splits, float* ws_o, float* ws_ml,
                                void* stream);
int mpf_masked_xattn_bwd_f32_ex(const float* q_hi, const float* q_lo, const float* qt_hi, const float* qt_lo,
                                const float* k_hi, const float* k_lo, const float* kt_hi, const float* kt_lo,
                                const float* v_hi, const float* v_lo, const float* do_hi, const float* do_lo,
                                const float* dot_hi, const float* dot_lo, const uint32_t* mask_bits,
                                const uint8_t* row_open, const float* lse2, const float* delta, float* dq, float* dk,
                                float* dv, int B, int Qt, int qt_ld, int HW, int heads, int head_dim, int mask_words,
                                int key_splits, float* ws_dq, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Hungarian matching on the device (the caller right after the prediction heads).
 * ref: mask2former/modeling/matcher.py:96-157 (HungarianMatcher.memory_efficient_forward), :15-30 (batch_dice_loss),
 *      :38-62 (batch_sigmoid_ce_loss); detectron2 point_sample (F.grid_sample at 2*c-1, bilinear, zeros padding,
 *      align_corners=False); scipy.optimize.linear_sum_assignment (matcher.py:151).
 *
 * mpf_match_cost_f32: for every image b, query q and target j of that image
 *     cost[Q*off[b] + q*n_b + j] = cost_mask * mean_p BCE(x_qp, t_jp) + cost_class * (-softmax(logits[b,q])[label_j])
 *                                  + cost_dice * (1 - (2 sum_p sig(x_qp) t_jp + 1) / (sum_p sig(x_qp) + sum_p t_jp + 1))
 *   with x_qp / t_jp the bilinear samples of pred_masks[b,q] / target mask j at point_coords[b,p] (one point set per
 *   image shared by all its masks, matcher.py:118-132).  One row-major [Q, n_b] matrix per image, images back to back.
 *     pred_logits   [B, Q, K+1] through (logits_img_stride, logits_q_stride), class dim contiguous
 *     pred_masks    [B, Q, H, W] through (masks_img_stride, masks_q_stride), each H x W map contiguous
 *     tgt_mask_ptrs [B] DEVICE array of device pointers; entry b -> [n_b, Hg, Wg] contiguous, uint8/bool (0/1) or,
 *                   with tgt_is_f32 != 0, float32 (the reference converts with `.to(out_mask)`, matcher.py:115)
 *     tgt_labels    [total_targets] int64, images back to back;  tgt_offsets [B+1] int32 prefix sums of n_b
 *     point_coords  [B, P, 2] (x, y) in [0, 1], 8-byte aligned
 *     workspace     >= mpf_match_cost_workspace_bytes(...) bytes (per-point-split partial sums; no atomics, so the
 *                   result is run-to-run deterministic)
 * mpf_lsap_f32: solves every image's [Q, n_b] assignment problem (one CTA each) with scipy's algorithm, scan order
 *   and tie rule in float64, and writes the min(Q, n_b) pairs of image b at offset sum_{b'<b} min(Q, n_b') of
 *   out_query / out_target (int64), ordered as scipy returns them (ascending query index).  *status (int32, zeroed by
 *   the caller) is set to b+1 when image b's matrix is infeasible / contains NaN (scipy raises ValueError there) and
 *   its pairs are written as -1.  max(Q, max_targets) <= 4266.
 * ------------------------------------------------------------------------------------------- */
long long mpf_match_cost_workspace_bytes(int batch, int num_queries, int total_targets, int max_targets,
                                         int num_points);
int mpf_match_cost_f32(const float* pred_logits, long long logits_img_stride, long long logits_q_stride,
                       int num_classes_p1, const float* pred_masks, long long masks_img_stride,
                       long long masks_q_stride, int H, int W, const void* const* tgt_mask_ptrs, int tgt_is_f32,
                       int Hg, int Wg, const int64_t* tgt_labels, const int32_t* tgt_offsets, int total_targets,
                       int max_targets, const float* point_coords, int batch, int num_queries, int num_points,
                       float cost_class, float cost_mask, float cost_dice, void* workspace,
                       long long workspace_bytes, float* cost, void* stream);
/* The matcher's prediction samples through a STREAMING pass instead of gathers (same reference lines: the
 * point_sample of out_mask at matcher.py:124-131): every [H, W] map maps[b, q] is read once, band by band, into shared
 * memory and evaluated at the image's P shared points -> out [B, Q, P] (fp32).  point_coords [B, P, 2] should be
 * ordered row-major by the top-left pixel of the bilinear footprint, with band_lo [B, n_bands + 1] (int32) holding
 * the index of the first point of every band of band_rows rows (band_lo[b][n_bands] = P); points whose footprint
 * lies outside their band are still sampled correctly (from global memory), only slower.  W % 4 == 0.
 * mpf_match_cost_presampled_f32 is mpf_match_cost_f32 with those samples in place of (pred_masks, strides). */
int mpf_sample_shared_points_f32(const float* maps, long long img_stride, long long q_stride, int H, int W,
                                 const float* point_coords, const int32_t* band_lo, int n_bands, int band_rows,
                                 int batch, int num_queries, int num_points, float* out, void* stream);
int mpf_match_cost_presampled_f32(const float* pred_logits, long long logits_img_stride, long long logits_q_stride,
                                  int num_classes_p1, const float* sampled, int H, int W,
                                  const void* const* tgt_mask_ptrs, int tgt_is_f32, int Hg, int Wg,
                                  const int64_t* tgt_labels, const int32_t* tgt_offsets, int total_targets,
                                  int max_targets, const float* point_coords, int batch, int num_queries,
                                  int num_points, float cost_class, float cost_mask, float cost_dice, void* workspace,
                                  long long workspace_bytes, float* cost, void* stream);

int mpf_lsap_f32(const float* cost, const int32_t* tgt_offsets, int batch, int num_queries, int max_targets,
                 int64_t* out_query, int64_t* out_target, int32_t* status, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Point sampling of mask maps for the criterion.
 * ref: mask2former/modeling/criterion.py:141-191 (SetCriterion.loss_masks), detectron2 point_sample
 *      (F.grid_sample at 2*c-1, bilinear, zeros padding, align_corners=False) and its autograd backward.
 *   mpf_point_sample_rows:         out[r, p] = bilinear(map_r, point_coords[r, p, :])   (neg_abs != 0: -|.|, the
 *                                  uncertainty score of criterion.py:73-87)
 *   mpf_point_sample_rows_bwd_f32: grad_map_r[corner] += w_corner * grad_out[r, p]  (fp32 atomics; the caller zeroes
 *                                  the gradient maps; rows may alias the same map)
 *     map_ptrs      [rows] DEVICE array of device pointers, each to one contiguous H x W map (uint8/bool 0/1, or
 *                   float32 with maps_are_f32 != 0): rows are sampled where they live -- no gather, no conversion
 *     point_coords  [rows, num_points, 2] (x, y) in [0, 1], 8-byte aligned;  out / grad_out [rows, num_points]
 * ------------------------------------------------------------------------------------------- */
int mpf_point_sample_rows(const void* const* map_ptrs, int maps_are_f32, int H, int W, const float* point_coords,
                          int rows, int num_points, int neg_abs, float* out, void* stream);
int mpf_point_sample_rows_bwd_f32(float* const* grad_map_ptrs, int H, int W, const float* point_coords, int rows,
                                  int num_points, const float* grad_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Instance-segmentation epilogue.
 * ref: mask2former/maskformer_model.py:239-244 (F.interpolate of pred_masks to the padded image size), :257-259 +
 *      detectron2 sem_seg_postprocess (crop to the image, resize to the output resolution), :365-401
 *      (instance_inference: top-k queries' maps, `> 0`, average foreground probability).
 * For row r and output pixel (y, x):  v = resize2(crop(resize1(mask_logits[query_index[r]])))[y, x]  with both
 * bilinear resizes (align_corners=False, ATen's index / weight arithmetic) evaluated on the fly,
 *     out_masks[r, y, x] = v > 0   (uint8, or float32 0/1 with out_is_f32 != 0)
 *     partial[r, k, 0] = sum over block k's pixels of sigmoid(v) * [v > 0];   partial[r, k, 1] = sum of [v > 0]
 * with k < mpf_instance_masks_blocks(out_h, out_w); the caller adds the blocks (fixed order: deterministic) and forms
 * the mask score  sum0 / (sum1 + 1e-6)  (:398).
 *   mask_logits [Q, h, w] of ONE image through query_stride (each map contiguous); query_index int64 [rows]
 * ------------------------------------------------------------------------------------------- */
int mpf_instance_masks_blocks(int out_h, int out_w);
int mpf_instance_masks_f32(const float* mask_logits, long long query_stride, int h, int w,
                           const int64_t* query_index, int rows, int padded_h, int padded_w, int image_h,
                           int image_w, int out_h, int out_w, void* out_masks, int out_is_f32, float* partial,
                           void* stream);

/* ---------------------------------------------------------------------------------------------
 * FPN stage of the pixel decoder (ref pixel_decoder/msdeformattn.py:343-351): the map changes layout twice around
 * the 3x3 convolution (library, NCHW); both crossings are fused into the elementwise work next to them.
 *   mpf_upsample2x_add_nchw_fwd_f32:  out[b,c,h,w] = cur[b,h,w,c] + bilinear_x2(prev)[b,h,w,c]
 *     (ref :347-349: cur_fpn + F.interpolate(out[-1], size=cur_fpn.shape[-2:], mode="bilinear", align_corners=False)
 *     for the exact-x2 case); cur [B,H,W,C], prev [B,H/2,W/2,C] channels-last; H even, W % 4 == 0, C % 64 == 0.
 *   mpf_upsample2x_add_nchw_bwd_f32:  g [B,C,H,W] -> g_cur [B,H,W,C] and g_prev [B,H/2,W/2,C] (adjoint of the x2 resize).
 *   mpf_groupnorm_nchw2cl_fwd_f32 / _bwd_f32:  GroupNorm (+ReLU) reading the convolution output in NCHW and writing
 *     channels-last tokens (ref :275 output conv norm + activation); arguments as mpf_groupnorm_cl_*, x / dx NCHW
 *     [B,C,HW], y / dy channels-last [B,HW,C]; C % 64 == 0, channels per group a multiple of 4 dividing 64, HW % 4 == 0. */
int mpf_upsample2x_add_nchw_fwd_f32(const float* cur, const float* prev, int batch, int H, int W, int C, float* out,
                                    void* stream);
int mpf_upsample2x_add_nchw_bwd_f32(const float* g, int batch, int H, int W, int C, float* g_cur, float* g_prev,
                                    void* stream);
int mpf_groupnorm_nchw2cl_fwd_f32(const float* x, const float* gamma, const float* beta, float eps, int batch,
                                  long long HW, int C, int groups, int relu, float* y, float* mean, float* rstd,
                                  double* stats_ws, void* stream);
int mpf_groupnorm_nchw2cl_bwd_f32(const float* dy, const float* x, const float* gamma, const float* beta,
                                  const float* mean, const float* rstd, int batch, long long HW, int C, int groups,
                                  int relu, float* dx, float* dgamma_dbeta, double* stats_ws, void* stream);

/* ---------------------------------------------------------------------------------------------
 * 3x3 convolution (stride 1, zero padding 1) of a channels-last map on the bf16x3 tensor-core GEMMs
 * (ref pixel_decoder/msdeformattn.py:268-275 output conv `layer_{i}`, 77 GFLOP per image at stride 4):
 * K = 9 * Cin runs over (tap, channel); the tap only shifts the TMA box, out-of-map pixels are zero-filled.
 *   mpf_conv3x3_cl_bf16x3:        y[b,h,w,co] = bias[co] + sum_{ky,kx,ci} x[b,h+ky-1,w+kx-1,ci] * Wm[co,(ky*3+kx)*Cin+ci]
 *                                 (Wm pre-split by mpf_split_bf16; also the input gradient, with the flipped /
 *                                 transposed weights);  any H, W; Cin % 32 == 0, Cout % 4 == 0.
 *   mpf_conv3x3_cl_wgrad_bf16x3:  dw[s][co][(ky*3+kx)*Cin+ci] = partial sums over the pixels of split s of
 *                                 dy[b,h,w,co] * x[b,h+ky-1,w+kx-1,ci]  (k_splits slabs, summed by the caller);
 *                                 any H, W; Cin % 64 == 0.
 *   mpf_upsample2x_add_cl_fwd_f32 / mpf_upsample2x_cl_bwd_f32: the FPN merge with both sides channels-last
 *                                 (the gradient of `cur` is the incoming gradient itself). */
int mpf_conv3x3_cl_bf16x3(const float* x, const uint16_t* w_hi, const uint16_t* w_lo, const float* bias, float* y,
                          int batch, int H, int W, int Cin, int Cout, int relu, void* stream);
int mpf_conv3x3_cl_wgrad_bf16x3(const float* dy, const float* x, float* dw, int batch, int H, int W, int Cin, int Cout,
                                int k_splits, void* stream);
int mpf_upsample2x_add_cl_fwd_f32(const float* cur, const float* prev, int batch, int H, int W, int C, float* out,
                                  void* stream);
int mpf_upsample2x_cl_bwd_f32(const float* g, int batch, int H, int W, int C, float* g_prev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MPFORMER_B200_H_ */

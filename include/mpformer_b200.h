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

#ifdef __cplusplus
}
#endif
#endif /* MPFORMER_B200_H_ */

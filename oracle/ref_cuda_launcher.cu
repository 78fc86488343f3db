// TEST/BENCH INFRASTRUCTURE ONLY.  Thin C-ABI launcher around the UNMODIFIED reference CUDA kernels
// (the "stock CUDA" comparator of SURVEY.md §2.2 / §8d).  The kernel header is #included from the
// reference tree where it lies at BUILD time (-I/root/reference/...); nothing is copied into this
// repository, and the resulting .so goes to the git-ignored oracle/_ref/.
// The reference's own host file (ms_deform_attn_cuda.cu) does not compile against torch 2.11
// (AT_DISPATCH_FLOATING_TYPES(value.type(), ...) at :69 and :139), hence this launcher, which calls
// the reference's launch functions ms_deformable_im2col_cuda / ms_deformable_col2im_cuda
// (ms_deform_im2col_cuda.cuh:928-959, :961-1332) exactly as that host file does (:66-80, :136-153).
#include "cuda/ms_deform_im2col_cuda.cuh"

extern "C" int ref_msda_forward_f32(const float* value, const int64_t* shapes, const int64_t* lstart,
                                    const float* loc, const float* aw, int B, int S, int M, int D,
                                    int L, int Lq, int P, float* out, void* stream) {
  ms_deformable_im2col_cuda<float>(static_cast<cudaStream_t>(stream), value, shapes, lstart, loc, aw,
                                   B, S, M, D, L, Lq, P, out);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int ref_msda_backward_f32(const float* grad_out, const float* value, const int64_t* shapes,
                                     const int64_t* lstart, const float* loc, const float* aw, int B,
                                     int S, int M, int D, int L, int Lq, int P, float* gv, float* gl,
                                     float* ga, void* stream) {
  ms_deformable_col2im_cuda<float>(static_cast<cudaStream_t>(stream), grad_out, value, shapes, lstart,
                                   loc, aw, B, S, M, D, L, Lq, P, gv, gl, ga);
  return static_cast<int>(cudaGetLastError());
}

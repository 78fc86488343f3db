// Instance-segmentation epilogue: mask logits -> full-resolution binary masks + mask scores.
//
//   ref: mask2former/maskformer_model.py:239-244 (F.interpolate of pred_masks to the padded image size, bilinear,
//        align_corners=False), :257-259 + detectron2 sem_seg_postprocess (crop to the image, bilinear resize to the
//        requested output resolution), :365-401 instance_inference (gather of the top-k queries' maps, `> 0`,
//        average foreground probability).
//
// The reference materialises, per image, Q upsampled fp32 maps (420 MB at Q=100, 1024x1024), crops and resizes them
// again, gathers the top-k rows (another copy), thresholds into a float map and makes two more passes for the score.
// Here one kernel reads the 256x256 logits of the selected queries (L2/L1 resident) and evaluates both resampling
// stages on the fly for every output pixel -- same index / weight arithmetic and association as ATen's
// upsample_bilinear2d (see maskbits.cu) -- thresholds, writes the mask once (uint8 or float) and accumulates
//   sum sigmoid(v) * [v > 0]   and   sum [v > 0]
// per row as per-block partials (summed by the caller in a fixed order: deterministic, no atomics).
// Bound: HBM write of the masks (topk * oh * ow bytes); one thread per output pixel, 16 cached gathers each.
#include "mpf_common.cuh"
#include "inference_math.cuh"

namespace mpf {

template <typename TO>
__global__ void __launch_bounds__(256)
instance_masks_kernel(const float* __restrict__ logits, long long q_stride, const long long* __restrict__ query_index,
                      const TwoStage ts, int oh, int ow, TO* __restrict__ out, float* __restrict__ partial) {
  const int r = blockIdx.y;
  const long long npix = static_cast<long long>(oh) * ow;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const float* L = logits + __ldg(query_index + r) * q_stride;
  float fg = 0.f, prob = 0.f;
  if (idx < npix) {
    const int y = static_cast<int>(idx / ow), x = static_cast<int>(idx - static_cast<long long>(y) * ow);
    const float v = two_stage_at(L, ts, y, x);
    const bool m = v > 0.f;                                        // maskformer_model.py:392
    out[static_cast<long long>(r) * npix + idx] = static_cast<TO>(m ? 1 : 0);
    if (m) {
      fg = 1.f;
      prob = 1.0f / (1.0f + expf(-v));                             // :398 (only foreground pixels contribute)
    }
  }
  // block sums -> partial[r, blockIdx.x, 0..1]
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    fg += __shfl_xor_sync(0xffffffffu, fg, o);
    prob += __shfl_xor_sync(0xffffffffu, prob, o);
  }
  __shared__ float s_fg[8], s_pr[8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { s_fg[wid] = fg; s_pr[wid] = prob; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { a += s_pr[k]; b += s_fg[k]; }
    float* dst = partial + (static_cast<long long>(r) * gridDim.x + blockIdx.x) * 2;
    dst[0] = a;
    dst[1] = b;
  }
}

}  // namespace mpf

extern "C" {

int mpf_instance_masks_blocks(int out_h, int out_w) {
  if (out_h <= 0 || out_w <= 0) return -1;
  const long long n = (static_cast<long long>(out_h) * out_w + 255) / 256;
  return n < (1ll << 31) ? static_cast<int>(n) : -1;
}

int mpf_instance_masks_f32(const float* mask_logits, long long query_stride, int h, int w,
                           const int64_t* query_index, int rows, int padded_h, int padded_w, int image_h,
                           int image_w, int out_h, int out_w, void* out_masks, int out_is_f32, float* partial,
                           void* stream) {
  using namespace mpf;
  clear_error();
  MPF_REQUIRE(rows >= 0 && h > 0 && w > 0 && padded_h > 0 && padded_w > 0 && image_h > 0 && image_w > 0 && out_h > 0 &&
                  out_w > 0,
              "instance_masks: sizes must be positive");
  MPF_REQUIRE(image_h <= padded_h && image_w <= padded_w, "instance_masks: image (%d x %d) larger than the padded "
              "size (%d x %d)", image_h, image_w, padded_h, padded_w);
  if (rows == 0) return MPF_OK;
  MPF_REQUIRE(mask_logits && query_index && out_masks && partial, "instance_masks: null pointer argument");
  MPF_REQUIRE(query_stride >= static_cast<long long>(h) * w, "instance_masks: query_stride < h*w");
  MPF_REQUIRE(rows <= 65535, "instance_masks: more than 65535 rows");
  const int blocks = mpf_instance_masks_blocks(out_h, out_w);
  MPF_REQUIRE(blocks > 0, "instance_masks: output too large");
  const TwoStage ts = make_two_stage(h, w, padded_h, padded_w, image_h, image_w, out_h, out_w);
  const dim3 grid(blocks, rows);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long* qi = reinterpret_cast<const long long*>(query_index);
  if (out_is_f32)
    instance_masks_kernel<float><<<grid, 256, 0, st>>>(mask_logits, query_stride, qi, ts, out_h, out_w,
                                                       static_cast<float*>(out_masks), partial);
  else
    instance_masks_kernel<uint8_t><<<grid, 256, 0, st>>>(mask_logits, query_stride, qi, ts, out_h, out_w,
                                                         static_cast<uint8_t*>(out_masks), partial);
  count_launch();
  return finish_launch("instance_masks");
}

}  // extern "C"

// Instance-segmentation epilogue: mask logits -> full-resolution binary masks + mask scores.
//
//   ref: mask2former/maskformer_model.py:236-243 (F.interpolate of pred_masks to the padded image size, bilinear,
//        align_corners=False), :256-260 + detectron2 sem_seg_postprocess (crop to the image, bilinear resize to the
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

namespace mpf {

__device__ __forceinline__ void bilinear_src(float scale, int dst, int in_size, int& i0, int& i1, float& l1) {
  float s = scale * (static_cast<float>(dst) + 0.5f) - 0.5f;
  s = s < 0.f ? 0.f : s;
  i0 = static_cast<int>(s);
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  l1 = s - static_cast<float>(i0);
}

// value of the stage-1 map (h x w logits -> Hp x Wp) at (y, x); products and sums separately rounded like ATen
__device__ __forceinline__ float stage1_at(const float* __restrict__ L, int h, int w, float sh, float sw, int y, int x) {
  int y0, y1, x0, x1;
  float ly, lx;
  bilinear_src(sh, y, h, y0, y1, ly);
  bilinear_src(sw, x, w, x0, x1, lx);
  const float hy = 1.f - ly, hx = 1.f - lx;
  const float p00 = __ldg(L + y0 * w + x0), p01 = __ldg(L + y0 * w + x1);
  const float p10 = __ldg(L + y1 * w + x0), p11 = __ldg(L + y1 * w + x1);
  const float top = __fadd_rn(__fmul_rn(hx, p00), __fmul_rn(lx, p01));
  const float bot = __fadd_rn(__fmul_rn(hx, p10), __fmul_rn(lx, p11));
  return __fadd_rn(__fmul_rn(hy, top), __fmul_rn(ly, bot));
}

template <typename TO>
__global__ void __launch_bounds__(256)
instance_masks_kernel(const float* __restrict__ logits, long long q_stride, int h, int w,
                      const long long* __restrict__ query_index, int Hp, int Wp, int ih, int iw, int oh, int ow,
                      float s1h, float s1w, float s2h, float s2w, int identity2, TO* __restrict__ out,
                      float* __restrict__ partial) {
  const int r = blockIdx.y;
  const long long npix = static_cast<long long>(oh) * ow;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const float* L = logits + __ldg(query_index + r) * q_stride;
  float fg = 0.f, prob = 0.f;
  if (idx < npix) {
    const int y = static_cast<int>(idx / ow), x = static_cast<int>(idx - static_cast<long long>(y) * ow);
    float v;
    if (identity2) {                 // output resolution == image size: the second resize is the identity
      v = stage1_at(L, h, w, s1h, s1w, y, x);
    } else {
      int y0, y1, x0, x1;
      float ly, lx;
      bilinear_src(s2h, y, ih, y0, y1, ly);
      bilinear_src(s2w, x, iw, x0, x1, lx);
      const float hy = 1.f - ly, hx = 1.f - lx;
      const float u00 = stage1_at(L, h, w, s1h, s1w, y0, x0), u01 = stage1_at(L, h, w, s1h, s1w, y0, x1);
      const float u10 = stage1_at(L, h, w, s1h, s1w, y1, x0), u11 = stage1_at(L, h, w, s1h, s1w, y1, x1);
      const float top = __fadd_rn(__fmul_rn(hx, u00), __fmul_rn(lx, u01));
      const float bot = __fadd_rn(__fmul_rn(hx, u10), __fmul_rn(lx, u11));
      v = __fadd_rn(__fmul_rn(hy, top), __fmul_rn(ly, bot));
    }
    const bool m = v > 0.f;                                        // maskformer_model.py:391
    out[static_cast<long long>(r) * npix + idx] = static_cast<TO>(m ? 1 : 0);
    if (m) {
      fg = 1.f;
      prob = 1.0f / (1.0f + expf(-v));                             // :397 (only foreground pixels contribute)
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
  // ATen: scale = (float)input_size / output_size  (area_pixel_compute_scale, align_corners=False)
  const float s1h = static_cast<float>(h) / static_cast<float>(padded_h);
  const float s1w = static_cast<float>(w) / static_cast<float>(padded_w);
  const float s2h = static_cast<float>(image_h) / static_cast<float>(out_h);
  const float s2w = static_cast<float>(image_w) / static_cast<float>(out_w);
  const int identity2 = (image_h == out_h && image_w == out_w) ? 1 : 0;
  const dim3 grid(blocks, rows);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long* qi = reinterpret_cast<const long long*>(query_index);
  if (out_is_f32)
    instance_masks_kernel<float><<<grid, 256, 0, st>>>(mask_logits, query_stride, h, w, qi, padded_h, padded_w, image_h,
                                                       image_w, out_h, out_w, s1h, s1w, s2h, s2w, identity2,
                                                       static_cast<float*>(out_masks), partial);
  else
    instance_masks_kernel<uint8_t><<<grid, 256, 0, st>>>(mask_logits, query_stride, h, w, qi, padded_h, padded_w,
                                                         image_h, image_w, out_h, out_w, s1h, s1w, s2h, s2w,
                                                         identity2, static_cast<uint8_t*>(out_masks), partial);
  count_launch();
  return finish_launch("instance_masks");
}

}  // extern "C"

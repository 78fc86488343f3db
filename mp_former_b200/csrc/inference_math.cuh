// Per-pixel arithmetic of the instance-mask kernel (inference.cu), host+device so that the exact source the kernel
// runs can be pinned on a CPU against torch's interpolate chain (tests/test_inference_cpu.py builds
// tests/host_emul/inference_host.cu with nvcc and calls it through ctypes).
#pragma once

#ifdef __CUDA_ARCH__
#define MPF_MUL_RN(a, b) __fmul_rn((a), (b))
#define MPF_ADD_RN(a, b) __fadd_rn((a), (b))
#define MPF_LDG(p) __ldg(p)
#else   // host build: plain IEEE operations (no contraction on the baseline x86-64 target)
#define MPF_MUL_RN(a, b) ((a) * (b))
#define MPF_ADD_RN(a, b) ((a) + (b))
#define MPF_LDG(p) (*(p))
#endif

namespace mpf {

// ATen upsample_bilinear2d(align_corners=False): src = scale * (dst + 0.5) - 0.5 clamped at 0, scale = in / out
__host__ __device__ __forceinline__ void bilinear_src(float scale, int dst, int in_size, int& i0, int& i1, float& l1) {
  float s = scale * (static_cast<float>(dst) + 0.5f) - 0.5f;
  s = s < 0.f ? 0.f : s;
  i0 = static_cast<int>(s);
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  l1 = s - static_cast<float>(i0);
}

// value of the stage-1 map (h x w logits -> padded size) at (y, x); products and sums separately rounded like ATen
__host__ __device__ __forceinline__ float stage1_at(const float* __restrict__ L, int h, int w, float sh, float sw, int y,
                                                    int x) {
  int y0, y1, x0, x1;
  float ly, lx;
  bilinear_src(sh, y, h, y0, y1, ly);
  bilinear_src(sw, x, w, x0, x1, lx);
  const float hy = 1.f - ly, hx = 1.f - lx;
  const float p00 = MPF_LDG(L + y0 * w + x0), p01 = MPF_LDG(L + y0 * w + x1);
  const float p10 = MPF_LDG(L + y1 * w + x0), p11 = MPF_LDG(L + y1 * w + x1);
  const float top = MPF_ADD_RN(MPF_MUL_RN(hx, p00), MPF_MUL_RN(lx, p01));
  const float bot = MPF_ADD_RN(MPF_MUL_RN(hx, p10), MPF_MUL_RN(lx, p11));
  return MPF_ADD_RN(MPF_MUL_RN(hy, top), MPF_MUL_RN(ly, bot));
}

struct TwoStage {        // scales of the two resizes and the crop in between
  int h, w, ih, iw;
  float s1h, s1w, s2h, s2w;
  int identity2;         // output resolution == image size: the second resize is the identity
};

__host__ __device__ inline TwoStage make_two_stage(int h, int w, int padded_h, int padded_w, int image_h, int image_w,
                                                   int out_h, int out_w) {
  TwoStage t;
  t.h = h; t.w = w; t.ih = image_h; t.iw = image_w;
  // ATen: scale = (float)input_size / output_size  (area_pixel_compute_scale, align_corners=False)
  t.s1h = static_cast<float>(h) / static_cast<float>(padded_h);
  t.s1w = static_cast<float>(w) / static_cast<float>(padded_w);
  t.s2h = static_cast<float>(image_h) / static_cast<float>(out_h);
  t.s2w = static_cast<float>(image_w) / static_cast<float>(out_w);
  t.identity2 = (image_h == out_h && image_w == out_w) ? 1 : 0;
  return t;
}

// resize2(crop(resize1(L)))[y, x]
__host__ __device__ __forceinline__ float two_stage_at(const float* __restrict__ L, const TwoStage& t, int y, int x) {
  if (t.identity2) return stage1_at(L, t.h, t.w, t.s1h, t.s1w, y, x);
  int y0, y1, x0, x1;
  float ly, lx;
  bilinear_src(t.s2h, y, t.ih, y0, y1, ly);
  bilinear_src(t.s2w, x, t.iw, x0, x1, lx);
  const float hy = 1.f - ly, hx = 1.f - lx;
  const float u00 = stage1_at(L, t.h, t.w, t.s1h, t.s1w, y0, x0), u01 = stage1_at(L, t.h, t.w, t.s1h, t.s1w, y0, x1);
  const float u10 = stage1_at(L, t.h, t.w, t.s1h, t.s1w, y1, x0), u11 = stage1_at(L, t.h, t.w, t.s1h, t.s1w, y1, x1);
  const float top = MPF_ADD_RN(MPF_MUL_RN(hx, u00), MPF_MUL_RN(lx, u01));
  const float bot = MPF_ADD_RN(MPF_MUL_RN(hx, u10), MPF_MUL_RN(lx, u11));
  return MPF_ADD_RN(MPF_MUL_RN(hy, top), MPF_MUL_RN(ly, bot));
}

}  // namespace mpf

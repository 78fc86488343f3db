// Bilinear point sampling shared by the matcher and the criterion kernels (matcher.cu, point_sample.cu).
#pragma once

#include "mpf_common.cuh"

namespace mpf {

// Corner offsets (or -1 outside the map: zeros padding) and weights of one point on an H x W map.
// Mirrors ATen's grid_sampler_2d (bilinear, zeros, align_corners=False) on grid = 2*c - 1:
//   ix = ((gx + 1) * W - 1) / 2, nw = (x1 - ix) * (y1 - iy), ne = (ix - x0) * (y1 - iy), sw = ..., se = ...
struct Corners {
  int o[4];
  float w[4];
};

__device__ __forceinline__ Corners point_corners(float cx, float cy, int H, int W) {
  const float gx = __fsub_rn(__fmul_rn(2.0f, cx), 1.0f);
  const float gy = __fsub_rn(__fmul_rn(2.0f, cy), 1.0f);
  const float ix = __fsub_rn(__fmul_rn(__fadd_rn(gx, 1.0f), static_cast<float>(W)), 1.0f) * 0.5f;
  const float iy = __fsub_rn(__fmul_rn(__fadd_rn(gy, 1.0f), static_cast<float>(H)), 1.0f) * 0.5f;
  const float fx0 = floorf(ix), fy0 = floorf(iy);
  const float fx1 = fx0 + 1.0f, fy1 = fy0 + 1.0f;
  Corners c;
  c.w[0] = __fmul_rn(fx1 - ix, fy1 - iy);   // nw
  c.w[1] = __fmul_rn(ix - fx0, fy1 - iy);   // ne
  c.w[2] = __fmul_rn(fx1 - ix, iy - fy0);   // sw
  c.w[3] = __fmul_rn(ix - fx0, iy - fy0);   // se
  // coordinates far outside the map (a caller's own point set) must not overflow the int conversion
  const bool finite = (ix > -2.0f) && (iy > -2.0f) && (ix < static_cast<float>(W) + 1.0f) &&
                      (iy < static_cast<float>(H) + 1.0f);
  const int x0 = finite ? static_cast<int>(fx0) : -2, y0 = finite ? static_cast<int>(fy0) : -2;
  const int x1 = x0 + 1, y1 = y0 + 1;
  const bool vx0 = x0 >= 0 && x0 < W, vx1 = x1 >= 0 && x1 < W;
  const bool vy0 = y0 >= 0 && y0 < H, vy1 = y1 >= 0 && y1 < H;
  c.o[0] = (vx0 && vy0) ? y0 * W + x0 : -1;
  c.o[1] = (vx1 && vy0) ? y0 * W + x1 : -1;
  c.o[2] = (vx0 && vy1) ? y1 * W + x0 : -1;
  c.o[3] = (vx1 && vy1) ? y1 * W + x1 : -1;
  return c;
}

template <typename T>
__device__ __forceinline__ float sample_map(const T* __restrict__ map, const Corners& c) {
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (c.o[k] >= 0) acc += static_cast<float>(__ldg(map + c.o[k])) * c.w[k];
  return acc;
}

}  // namespace mpf

// Point sampling of mask maps for the criterion ("criterion point sampling").
//
//   ref: mask2former/modeling/criterion.py:141-191 (SetCriterion.loss_masks: point_sample of the matched prediction
//        maps at 3 x 12544 candidate points for the uncertainty ranking, then of predictions and GT masks at the
//        12544 chosen points), detectron2 point_sample (= F.grid_sample at 2*c-1, bilinear, zeros padding,
//        align_corners=False) and its autograd backward.
//
// The reference first gathers the matched maps (`pred_masks[idx]`, a copy), converts ALL GT masks of the batch to
// float in a zero-padded [B, n_max, Hg, Wg] tensor (1.3 GB at the bench geometry, per loss call) and samples those
// copies.  Here a row is addressed through a device table of map pointers, so predictions are sampled where the
// prediction heads wrote them, GT masks where the data loader put them (uint8/bool, never converted or padded), and the
// backward adds straight into the dense mask-logit gradient the prediction heads' backward consumes.
//
//   point_sample_rows_kernel<T>      out[r, p] = bilinear(map_r, coords[r, p])   (or -|.|, the uncertainty score)
//   point_sample_rows_bwd_kernel     grad_map_r[corner] += w_corner * grad_out[r, p]   (fp32 atomics, like ATen's)
//
// Bound: gathers (two 32-byte sectors per sample); one thread per sample, grid (ceil(P / 256), rows).
#include "mpf_common.cuh"
#include "point_sample.cuh"

namespace mpf {

template <typename T>
__global__ void __launch_bounds__(256)
point_sample_rows_kernel(const void* const* __restrict__ map_ptrs, int H, int W, const float* __restrict__ coords,
                         int P, int neg_abs, float* __restrict__ out) {
  const int r = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const float2 xy = __ldg(reinterpret_cast<const float2*>(coords) + static_cast<long long>(r) * P + p);
  const Corners c = point_corners(xy.x, xy.y, H, W);
  const float v = sample_map(static_cast<const T*>(map_ptrs[r]), c);
  out[static_cast<long long>(r) * P + p] = neg_abs ? -fabsf(v) : v;
}

__global__ void __launch_bounds__(256)
point_sample_rows_bwd_kernel(float* const* __restrict__ grad_map_ptrs, int H, int W,
                             const float* __restrict__ coords, int P, const float* __restrict__ grad_out) {
  const int r = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const float2 xy = __ldg(reinterpret_cast<const float2*>(coords) + static_cast<long long>(r) * P + p);
  const Corners c = point_corners(xy.x, xy.y, H, W);
  const float g = __ldg(grad_out + static_cast<long long>(r) * P + p);
  float* gm = grad_map_ptrs[r];
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (c.o[k] >= 0) atomicAdd(gm + c.o[k], c.w[k] * g);
}

}  // namespace mpf

extern "C" {

int mpf_point_sample_rows(const void* const* map_ptrs, int maps_are_f32, int H, int W, const float* point_coords,
                          int rows, int num_points, int neg_abs, float* out, void* stream) {
  using namespace mpf;
  clear_error();
  MPF_REQUIRE(rows >= 0 && num_points >= 0 && H > 0 && W > 0, "point_sample_rows: bad sizes");
  if (rows == 0 || num_points == 0) return MPF_OK;
  MPF_REQUIRE(map_ptrs && point_coords && out, "point_sample_rows: null pointer argument");
  MPF_REQUIRE(static_cast<long long>(H) * W < (1ll << 31), "point_sample_rows: map too large");
  MPF_REQUIRE(rows <= 65535, "point_sample_rows: more than 65535 rows");
  MPF_REQUIRE((reinterpret_cast<uintptr_t>(point_coords) & 7u) == 0, "point_sample_rows: point_coords must be 8-byte aligned");
  const dim3 grid((num_points + 255) / 256, rows);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (maps_are_f32)
    point_sample_rows_kernel<float><<<grid, 256, 0, st>>>(map_ptrs, H, W, point_coords, num_points, neg_abs, out);
  else
    point_sample_rows_kernel<uint8_t><<<grid, 256, 0, st>>>(map_ptrs, H, W, point_coords, num_points, neg_abs, out);
  count_launch();
  return finish_launch("point_sample_rows");
}

int mpf_point_sample_rows_bwd_f32(float* const* grad_map_ptrs, int H, int W, const float* point_coords, int rows,
                                  int num_points, const float* grad_out, void* stream) {
  using namespace mpf;
  clear_error();
  MPF_REQUIRE(rows >= 0 && num_points >= 0 && H > 0 && W > 0, "point_sample_rows_bwd: bad sizes");
  if (rows == 0 || num_points == 0) return MPF_OK;
  MPF_REQUIRE(grad_map_ptrs && point_coords && grad_out, "point_sample_rows_bwd: null pointer argument");
  MPF_REQUIRE(static_cast<long long>(H) * W < (1ll << 31), "point_sample_rows_bwd: map too large");
  MPF_REQUIRE(rows <= 65535, "point_sample_rows_bwd: more than 65535 rows");
  MPF_REQUIRE((reinterpret_cast<uintptr_t>(point_coords) & 7u) == 0,
              "point_sample_rows_bwd: point_coords must be 8-byte aligned");
  const dim3 grid((num_points + 255) / 256, rows);
  point_sample_rows_bwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(grad_map_ptrs, H, W, point_coords,
                                                                                    num_points, grad_out);
  count_launch();
  return finish_launch("point_sample_rows_bwd");
}

}  // extern "C"

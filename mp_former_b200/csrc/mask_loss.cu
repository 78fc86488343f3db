// Point-sampled mask losses of the criterion, one CTA per matched (prediction, target) pair:
//   ref: mask2former/modeling/criterion.py:25-43  dice_loss       1 - (2 sum(p y) + 1) / (sum(p) + sum(y) + 1), p = sigmoid(x)
//        mask2former/modeling/criterion.py:51-68  sigmoid_ce_loss mean_p BCEWithLogits(x, y)
// x [R, P] are the logits of R masks sampled at P points, y [R, P] the ground truth sampled at the same points
// (criterion.py:174-186).  The reference evaluates this with ~15 elementwise / reduction launches per call and 20
// calls per step; here the rows of ALL prediction heads go through one forward and one backward launch.  Both kernels
// stream x and y once (HBM-bound: 8 B per point forward, 12 B per point backward).
#include "mpf_common.cuh"

namespace mpf {

constexpr int kLossThreads = 256;

__device__ __forceinline__ float block_sum(float v, float* s_red) {       // all threads receive the sum
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();                                                         // s_red may still be read from a previous call
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  float t = lane < kLossThreads / 32 ? s_red[lane] : 0.f;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
  return t;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(kLossThreads)
mask_loss_rows_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y, int P, float* __restrict__ bce,
                          float* __restrict__ dice, float* __restrict__ stats) {
  __shared__ float s_red[kLossThreads / 32];
  const long long r = blockIdx.x;
  const float* xr = x + r * P;
  const float* yr = y + r * P;
  float a_bce = 0.f, a_py = 0.f, a_p = 0.f, a_y = 0.f;
  for (int i = threadIdx.x; i < P; i += kLossThreads) {
    const float xv = __ldg(xr + i), yv = __ldg(yr + i);
    a_bce += fmaxf(xv, 0.f) - xv * yv + log1pf(expf(-fabsf(xv)));
    const float p = sigmoidf_(xv);
    a_py += p * yv;
    a_p += p;
    a_y += yv;
  }
  a_bce = block_sum(a_bce, s_red);
  a_py = block_sum(a_py, s_red);
  a_p = block_sum(a_p, s_red);
  a_y = block_sum(a_y, s_red);
  if (threadIdx.x == 0) {
    const float num = 2.f * a_py + 1.f, den = a_p + a_y + 1.f;
    bce[r] = a_bce / static_cast<float>(P);
    dice[r] = 1.f - num / den;
    stats[2 * r] = num;
    stats[2 * r + 1] = den;
  }
}

// d bce_r / d x_i  = (p_i - y_i) / P
// d dice_r / d x_i = -(2 y_i den - num) / den^2 * p_i (1 - p_i)
__global__ void __launch_bounds__(kLossThreads)
mask_loss_rows_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ stats,
                          const float* __restrict__ g_bce, const float* __restrict__ g_dice, int P,
                          float* __restrict__ gx) {
  const long long r = blockIdx.x;
  const float num = stats[2 * r], den = stats[2 * r + 1];
  const float gb = g_bce[r] / static_cast<float>(P);
  const float gd = g_dice[r] / (den * den);
  const float* xr = x + r * P;
  const float* yr = y + r * P;
  float* gr = gx + r * P;
  for (int i = threadIdx.x; i < P; i += kLossThreads) {
    const float xv = __ldg(xr + i), yv = __ldg(yr + i);
    const float p = sigmoidf_(xv);
    gr[i] = gb * (p - yv) - gd * (2.f * yv * den - num) * p * (1.f - p);
  }
}

}  // namespace mpf

extern "C" int mpf_mask_loss_rows_fwd_f32(const float* x, const float* y, int rows, int points, float* bce, float* dice,
                                          float* stats, void* stream) {
  using namespace mpf;
  clear_error();
  MPF_REQUIRE(rows >= 0 && points > 0, "mask_loss_rows: need rows >= 0 and points > 0 (rows=%d, points=%d)", rows, points);
  if (rows == 0) return MPF_OK;
  MPF_REQUIRE(x && y && bce && dice && stats, "mask_loss_rows: null pointer argument");
  mask_loss_rows_fwd_kernel<<<rows, kLossThreads, 0, static_cast<cudaStream_t>(stream)>>>(x, y, points, bce, dice, stats);
  count_launch();
  return finish_launch("mask_loss_rows_fwd");
}

extern "C" int mpf_mask_loss_rows_bwd_f32(const float* x, const float* y, const float* stats, const float* g_bce,
                                          const float* g_dice, int rows, int points, float* gx, void* stream) {
  using namespace mpf;
  clear_error();
  MPF_REQUIRE(rows >= 0 && points > 0, "mask_loss_rows: need rows >= 0 and points > 0 (rows=%d, points=%d)", rows, points);
  if (rows == 0) return MPF_OK;
  MPF_REQUIRE(x && y && stats && g_bce && g_dice && gx, "mask_loss_rows: null pointer argument");
  mask_loss_rows_bwd_kernel<<<rows, kLossThreads, 0, static_cast<cudaStream_t>(stream)>>>(x, y, stats, g_bce, g_dice,
                                                                                            points, gx);
  count_launch();
  return finish_launch("mask_loss_rows_bwd");
}

// Per-row top-k SELECTION with payload gather -- the "keep the most uncertain points" step of PointRend's importance
// sampling inside the criterion:
//   ref: mask2former/modeling/criterion.py:165-172 -> detectron2 get_uncertain_point_coords_with_randomness:
//        idx = topk(uncertainty[R, 3*12544], k = 9408).indices;  coords = candidate_coords[idx]
// torch.topk sorts every row (segmented radix sort + gather: 0.43 ms per call at the bench geometry, 20 calls per
// step).  Only the SET of the k largest is needed (the losses are sums over the points), so one CTA per row keeps the
// row's keys in shared memory, finds the k-th largest key with a 4-pass radix select (8 bits per pass, shared integer
// atomics for the histograms) and compacts the selected rows of the payload in index order (ballot + scan: the output
// is deterministic).  Ties at the threshold are resolved towards the lower index.
#include "mpf_common.cuh"

namespace mpf {

constexpr int kSelThreads = 1024;
constexpr int kSelMaxN = 49152;                 // keys of one row in shared memory (192 KB)

__device__ __forceinline__ unsigned sel_key(float f) {     // order-preserving: larger float <=> larger key
  const unsigned u = __float_as_uint(f);
  if ((u & 0x7FFFFFFFu) > 0x7F800000u) return 0xFFFFFFFFu;            // NaN of either sign ranks first (torch.topk)
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

template <int W>
__global__ void __launch_bounds__(kSelThreads, 1)
topk_gather_rows_kernel(const float* __restrict__ scores, int n, int k, const float* __restrict__ payload,
                        float* __restrict__ out) {
  extern __shared__ unsigned s_keys[];            // [n]
  __shared__ unsigned s_hist[256];
  __shared__ unsigned s_prefix, s_need, s_warp[2][kSelThreads / 32], s_base[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long row = blockIdx.x;
  const float* sc = scores + row * n;
  for (int i = tid; i < n; i += kSelThreads) s_keys[i] = sel_key(__ldg(sc + i));
  if (tid == 0) { s_prefix = 0u; s_need = static_cast<unsigned>(k); }
  __syncthreads();
  // radix select: after pass b the top 8*(b+1) bits of the k-th largest key are known
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    if (tid < 256) s_hist[tid] = 0u;
    __syncthreads();
    const unsigned prefix = s_prefix;
    const unsigned mask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
    for (int i = tid; i < n; i += kSelThreads) {
      const unsigned key = s_keys[i];
      if ((key & mask) == prefix) atomicAdd(&s_hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid == 0) {                                 // walk the bins from the top: 256 steps, negligible
      unsigned need = s_need, bin = 255u;
      for (;; --bin) {
        const unsigned c = s_hist[bin];
        if (c >= need || bin == 0u) break;
        need -= c;
      }
      s_prefix = prefix | (bin << shift);
      s_need = need;                                // how many keys EQUAL to the (partial) threshold are still wanted
    }
    __syncthreads();
  }
  const unsigned thr = s_prefix;
  const unsigned need_eq = s_need;                  // keys == thr to take (>= 1), lowest indices first
  if (tid < 2) s_base[tid] = 0u;
  __syncthreads();
  // ordered compaction: keys > thr and the first need_eq keys == thr, in index order
  float* dst = out + row * static_cast<long long>(k) * W;
  const float* pl = payload + row * static_cast<long long>(n) * W;
  for (int i0 = 0; i0 < n; i0 += kSelThreads) {
    const int i = i0 + tid;
    const unsigned key = i < n ? s_keys[i] : 0u;
    const bool gt = i < n && key > thr, eq = i < n && key == thr;
    const unsigned bg = __ballot_sync(0xffffffffu, gt), be = __ballot_sync(0xffffffffu, eq);
    if (lane == 0) { s_warp[0][warp] = __popc(bg); s_warp[1][warp] = __popc(be); }
    __syncthreads();
    if (warp < 2) {                                  // exclusive scan of the 32 warp counts, one warp per class
      const unsigned v = s_warp[warp][lane];
      unsigned x = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += y;
      }
      s_warp[warp][lane] = x - v;
      if (lane == 31) s_hist[warp] = x;               // s_hist[0 / 1] now hold the chunk totals of the two classes
    }
    __syncthreads();
    const unsigned lanemask = (1u << lane) - 1u;
    const unsigned eq_rank = s_base[1] + s_warp[1][warp] + __popc(be & lanemask);     // rank among the == thr keys
    const unsigned gt_rank = s_base[0] + s_warp[0][warp] + __popc(bg & lanemask);
    // position: all selected keys in index order = (# selected before i); selected = gt or (eq and eq_rank < need_eq)
    const unsigned eq_before = min(eq_rank, need_eq);
    if (gt || (eq && eq_rank < need_eq)) {
      const unsigned pos = gt_rank + eq_before;
      if (pos < static_cast<unsigned>(k)) {
#pragma unroll
        for (int w = 0; w < W; ++w) dst[static_cast<long long>(pos) * W + w] = __ldg(pl + static_cast<long long>(i) * W + w);
      }
    }
    __syncthreads();
    if (tid == 0) { s_base[0] += s_hist[0]; s_base[1] += s_hist[1]; }
    __syncthreads();
  }
}

}  // namespace mpf

extern "C" int mpf_topk_gather_rows_f32(const float* scores, int rows, int n, int k, const float* payload,
                                        int payload_width, float* out, void* stream) {
  using namespace mpf;
  clear_error();
  MPF_REQUIRE(rows >= 0 && n > 0 && k > 0 && k <= n, "topk_gather_rows: need 0 < k <= n (k=%d, n=%d)", k, n);
  MPF_REQUIRE(n <= kSelMaxN, "topk_gather_rows: n = %d exceeds the %d keys one CTA keeps in shared memory", n, kSelMaxN);
  MPF_REQUIRE(payload_width == 1 || payload_width == 2, "topk_gather_rows: payload width must be 1 or 2");
  if (rows == 0) return MPF_OK;
  MPF_REQUIRE(scores && payload && out, "topk_gather_rows: null pointer argument");
  const size_t smem = static_cast<size_t>(n) * 4;
  static unsigned long long seen1 = 0, seen2 = 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (payload_width == 2) {
    if (first_use_on_this_device(seen2))
      MPF_CUDA_OK(cudaFuncSetAttribute(topk_gather_rows_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       kSelMaxN * 4));
    topk_gather_rows_kernel<2><<<rows, kSelThreads, smem, st>>>(scores, n, k, payload, out);
  } else {
    if (first_use_on_this_device(seen1))
      MPF_CUDA_OK(cudaFuncSetAttribute(topk_gather_rows_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       kSelMaxN * 4));
    topk_gather_rows_kernel<1><<<rows, kSelThreads, smem, st>>>(scores, n, k, payload, out);
  }
  count_launch();
  return finish_launch("topk_gather_rows");
}

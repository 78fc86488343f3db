// All weight operands of a training step in ONE launch.
//
// Every nn.Linear / projection of the path feeds its weight to the bf16x3 GEMMs as a pre-split operand: bf16 hi / lo
// halves of W [N, K] for the forward (y = x W^T) and of W^T [K, N] for the input gradient (dx = dy W).  The reference's
// modules (pixel_decoder/msdeformattn.py:116-131, transformer_decoder/mask2former_transformer_decoder.py:19-180) hold
// ~170 weight matrices; splitting each where it is used costs a transposing copy and a split launch per use: ~300
// launches of a few microseconds per step, which is what bounds the 2-images-per-GPU step of the 8-GPU split.  Here a
// table describes every (weight, orientation) pair once, and one grid of 32 x 32 tiles writes all halves into an arena.
//   entry: source [rows, cols] fp32 with row stride ld (a row slice of a packed in-projection weight is fine);
//          destination hi / lo bf16, either [rows, cols] or -- transposed -- [cols, rows], contiguous.
#include "mpf_common.cuh"

#include <cuda_bf16.h>

namespace mpf {

struct WeightSplitEntry {          // 48 bytes, filled on the host (mp_former_b200/native.py WeightOperandCache)
  const float* src;
  __nv_bfloat16* hi;
  __nv_bfloat16* lo;
  int rows, cols;
  long long ld;
  int transposed;
  int tile0;                       // first tile of this entry in the grid; tiles_c = ceil(cols / 32)
};

__global__ void __launch_bounds__(256)
split_weights_kernel(const WeightSplitEntry* __restrict__ table, int n_entries) {
  __shared__ float tile[32][33];
  const int tid = threadIdx.x;
  // binary search: last entry with tile0 <= blockIdx.x
  int lo = 0, hi = n_entries - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (table[mid].tile0 <= static_cast<int>(blockIdx.x)) lo = mid; else hi = mid - 1;
  }
  const WeightSplitEntry e = table[lo];
  const int tiles_c = (e.cols + 31) >> 5;
  const int t = blockIdx.x - e.tile0;
  const int r0 = (t / tiles_c) << 5, c0 = (t % tiles_c) << 5;
  // load: thread -> row tid / 8, four columns
  const int lr = tid >> 3, lc = (tid & 7) << 2;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (r0 + lr < e.rows) {
    const float* s = e.src + static_cast<long long>(r0 + lr) * e.ld + c0 + lc;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (c0 + lc + i < e.cols) v[i] = __ldg(s + i);
  }
  if (!e.transposed) {
    if (r0 + lr < e.rows) {
      const long long o = static_cast<long long>(r0 + lr) * e.cols + c0 + lc;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (c0 + lc + i < e.cols) {
          const __nv_bfloat16 h = __float2bfloat16_rn(v[i]);
          e.hi[o + i] = h;
          e.lo[o + i] = __float2bfloat16_rn(v[i] - __bfloat162float(h));
        }
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) tile[lr][lc + i] = v[i];
  __syncthreads();
  // store: thread -> output row (source column) tid / 8, four source rows
  const int oc = tid >> 3, orr = (tid & 7) << 2;
  if (c0 + oc < e.cols) {
    const long long o = static_cast<long long>(c0 + oc) * e.rows + r0 + orr;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (r0 + orr + i < e.rows) {
        const float x = tile[orr + i][oc];
        const __nv_bfloat16 h = __float2bfloat16_rn(x);
        e.hi[o + i] = h;
        e.lo[o + i] = __float2bfloat16_rn(x - __bfloat162float(h));
      }
  }
}

}  // namespace mpf

extern "C" int mpf_split_weights_f32(const void* table, int n_entries, int total_tiles, void* stream) {
  using namespace mpf;
  clear_error();
  MPF_REQUIRE(n_entries >= 0 && total_tiles >= 0, "split_weights: negative sizes");
  if (n_entries == 0 || total_tiles == 0) return MPF_OK;
  MPF_REQUIRE(table != nullptr, "split_weights: null table");
  static_assert(sizeof(WeightSplitEntry) == 48, "host table layout");
  split_weights_kernel<<<total_tiles, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const WeightSplitEntry*>(table), n_entries);
  count_launch();
  return finish_launch("split_weights");
}

// Shared by msda.cu (gather kernels through L1) and msda_staged.cu (TMA-staged tile kernels): the query tiling of a
// launch and the two vector memory primitives.
#pragma once

#include "mpf_common.cuh"

namespace mpf {

constexpr int kMaxTiledLevels = 8;
constexpr int kChunkQ = 128;  // queries per work unit
constexpr int kTileW = 16;
constexpr int kTileH = 8;
constexpr int kThreads = 256;

struct MsdaTiling {
  int mode;  // 0: linear chunks of kChunkQ queries; 1: 16x8 tiles per level (num_query == S)
  int num_chunks;
  int L;
  int H[kMaxTiledLevels];
  int W[kMaxTiledLevels];
  int start[kMaxTiledLevels];
  int tiles_x[kMaxTiledLevels];
  int chunk_begin[kMaxTiledLevels + 1];
};

// chunk-local index j (0..127) -> query index, or -1 if the slot is padding.
__device__ __forceinline__ int query_of(const MsdaTiling& t, int chunk, int j, int num_query) {
  if (t.mode == 0) {
    int q = chunk * kChunkQ + j;
    return q < num_query ? q : -1;
  }
  int l = 0;
#pragma unroll
  for (int i = 1; i < kMaxTiledLevels; ++i)
    if (i < t.L && chunk >= t.chunk_begin[i]) l = i;
  int tt = chunk - t.chunk_begin[l];
  int ty = tt / t.tiles_x[l];
  int tx = tt - ty * t.tiles_x[l];
  int y = ty * kTileH + j / kTileW;
  int x = tx * kTileW + (j % kTileW);
  if (y >= t.H[l] || x >= t.W[l]) return -1;
  return t.start[l] + y * t.W[l] + x;
}

__device__ __forceinline__ float4 ldg4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b),
               "f"(c), "f"(d)
               : "memory");
}


// One sample's contribution to one output channel with the reference kernel's exact association, as nvcc compiles
// ref cuh:82-88 + :295 (checked in the SASS of oracle/_ref/libmsda_stock.so):
//   val = fma(w4, v4, fma(w3, v3, fma(w1, v1, w2 * v2)));   col = fma(weight, val, col)
// with w1 = hh*hw, w2 = hh*lw, w3 = lh*hw, w4 = lh*lw -- so the forward kernels are bit-identical to the stock CUDA op.
__device__ __forceinline__ float bilinear_acc(float acc, float a, float w1, float w2, float w3, float w4, float v1,
                                              float v2, float v3, float v4) {
  float t = __fmul_rn(w2, v2);
  t = __fmaf_rn(w1, v1, t);
  t = __fmaf_rn(w3, v3, t);
  t = __fmaf_rn(w4, v4, t);
  return __fmaf_rn(a, t, acc);
}

// shapes_host may be null (then linear chunking is used).
inline MsdaTiling make_tiling(const int64_t* shapes_host, int L, int S, int Lq) {
  MsdaTiling t;
  t.mode = 0;
  t.L = L;
  t.num_chunks = (Lq + kChunkQ - 1) / kChunkQ;
  for (int i = 0; i < kMaxTiledLevels; ++i) t.H[i] = t.W[i] = t.start[i] = t.tiles_x[i] = 0;
  for (int i = 0; i <= kMaxTiledLevels; ++i) t.chunk_begin[i] = 0;
  if (shapes_host == nullptr || Lq != S || L > kMaxTiledLevels) return t;
  long long total = 0;
  int chunks = 0;
  for (int l = 0; l < L; ++l) {
    const long long H = shapes_host[2 * l], W = shapes_host[2 * l + 1];
    if (H <= 0 || W <= 0 || H > (1 << 20) || W > (1 << 20)) return t;
    t.H[l] = static_cast<int>(H);
    t.W[l] = static_cast<int>(W);
    t.start[l] = static_cast<int>(total);
    t.tiles_x[l] = (t.W[l] + kTileW - 1) / kTileW;
    t.chunk_begin[l] = chunks;
    chunks += t.tiles_x[l] * ((t.H[l] + kTileH - 1) / kTileH);
    total += H * W;
  }
  if (total != S) return t;  // host shapes do not describe this value tensor: stay linear
  // Tiles waste slots when W < 16 or H < 8; fall back to linear if padding exceeds 2x.
  if (static_cast<long long>(chunks) * kChunkQ > 2ll * S + kChunkQ) return t;
  for (int l = L; l <= kMaxTiledLevels; ++l) t.chunk_begin[l] = chunks;
  t.num_chunks = chunks;
  t.mode = 1;
  return t;
}

}  // namespace mpf

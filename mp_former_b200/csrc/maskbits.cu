// Boolean stage of the prediction heads, bit-packed.
//
//   attn_mask = (sigmoid(bilinear_resize(outputs_mask, (h, w), align_corners=False)) < 0.5)
//   ref: transformer_decoder/mask2former_transformer_decoder.py:1869-1875
//
// The reference materialises this as bool [B*8, Q, h*w] (8 identical head copies).  Here it is ONE
// bit per (image, query, key): word j of a row holds keys 32j..32j+31, bit i = key 32j+i, 1 = masked
// (not allowed to attend).  The arithmetic reproduces ATen's so that the bits are identical to the
// reference's on identical logits:
//   * resize: upsample_bilinear2d, align_corners=False: src = scale*(dst+0.5)-0.5 clamped at 0,
//     scale = in/out (float), lambda = src - floor(src), value =
//     h0*(w0*p00 + w1*p01) + h1*(w0*p10 + w1*p11)   (same association as ATen)
//   * `sigmoid(x) < 0.5` in fp32, i.e. 1/(1+exp(-x)) < 0.5 with a correctly rounded exp, holds exactly
//     when x <= -0x1.7ffffep-23 (NOT `x < 0`: for -1.788e-7 < x < 0 the sigmoid rounds to exactly 0.5 and
//     the key stays unmasked).  The equivalence is checked exhaustively against torch on the CPU in
//     tests/test_host_logic_cpu.py; using the threshold keeps the bits independent of the last-ulp
//     behaviour of a device expf.
#include "mpf_common.cuh"

namespace mpf {

__device__ constexpr float kMaskLogitThreshold = -0x1.7ffffep-23f;

__device__ __forceinline__ void src_index(float scale, int dst, int in_size, int& i0, int& i1, float& l1) {
  float s = scale * (static_cast<float>(dst) + 0.5f) - 0.5f;
  s = s < 0.f ? 0.f : s;
  i0 = static_cast<int>(s);
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  l1 = s - static_cast<float>(i0);
}

// one warp per 32 consecutive output keys of one (b, q) row
__global__ void __launch_bounds__(256)
attn_mask_bits_kernel(const float* __restrict__ logits, long long row_stride, int rows, int H, int W, int h,
                      int w, float scale_h, float scale_w, uint32_t* __restrict__ bits, int words_per_row) {
  const int lane = threadIdx.x & 31;
  const long long warp_global = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long total = static_cast<long long>(rows) * words_per_row;
  if (warp_global >= total) return;
  const int row = static_cast<int>(warp_global / words_per_row);
  const int word = static_cast<int>(warp_global - static_cast<long long>(row) * words_per_row);
  const int key = word * 32 + lane;
  const int hw = h * w;
  bool masked = true;                      // padding keys (>= h*w) count as masked
  if (key < hw) {
    const int oy = key / w, ox = key - oy * w;
    int y0, y1, x0, x1;
    float ly, lx;
    src_index(scale_h, oy, H, y0, y1, ly);
    src_index(scale_w, ox, W, x0, x1, lx);
    const float hy = 1.f - ly, hx = 1.f - lx;
    const float* p = logits + static_cast<long long>(row) * row_stride;
    const float p00 = __ldg(p + static_cast<long long>(y0) * W + x0), p01 = __ldg(p + static_cast<long long>(y0) * W + x1);
    const float p10 = __ldg(p + static_cast<long long>(y1) * W + x0), p11 = __ldg(p + static_cast<long long>(y1) * W + x1);
    // no FMA contraction: ATen evaluates these products and sums separately rounded
    const float top = __fadd_rn(__fmul_rn(hx, p00), __fmul_rn(lx, p01));
    const float bot = __fadd_rn(__fmul_rn(hx, p10), __fmul_rn(lx, p11));
    const float v = __fadd_rn(__fmul_rn(hy, top), __fmul_rn(ly, bot));
    masked = v <= kMaskLogitThreshold;
  }
  const uint32_t m = __ballot_sync(0xffffffffu, masked);
  if (lane == 0) bits[static_cast<long long>(row) * words_per_row + word] = m;
}

// bool (uint8, 0/1) [rows, n] -> bits [rows, ceil(n/32)], padding bits set (masked)
__global__ void __launch_bounds__(256)
pack_bool_bits_kernel(const uint8_t* __restrict__ src, int rows, int n, uint32_t* __restrict__ bits,
                      int words_per_row) {
  const int lane = threadIdx.x & 31;
  const long long warp_global = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long total = static_cast<long long>(rows) * words_per_row;
  if (warp_global >= total) return;
  const int row = static_cast<int>(warp_global / words_per_row);
  const int word = static_cast<int>(warp_global - static_cast<long long>(row) * words_per_row);
  const int key = word * 32 + lane;
  const bool masked = key < n ? (src[static_cast<long long>(row) * n + key] != 0) : true;
  const uint32_t m = __ballot_sync(0xffffffffu, masked);
  if (lane == 0) bits[static_cast<long long>(row) * words_per_row + word] = m;
}

}  // namespace mpf

extern "C" {

int mpf_attn_mask_bits_f32(const float* logits, long long row_stride, int rows, int H, int W, int h, int w,
                           uint32_t* bits, int words_per_row, void* stream) {
  using namespace mpf;
  clear_error();
  MPF_REQUIRE(logits && bits, "attn_mask_bits: null pointer argument");
  MPF_REQUIRE(rows > 0 && H > 0 && W > 0 && h > 0 && w > 0, "attn_mask_bits: sizes must be positive");
  MPF_REQUIRE(words_per_row * 32 >= h * w, "attn_mask_bits: words_per_row (%d) too small for %dx%d keys",
              words_per_row, h, w);
  MPF_REQUIRE(row_stride >= static_cast<long long>(H) * W, "attn_mask_bits: row_stride < H*W");
  const float sh = static_cast<float>(H) / static_cast<float>(h);
  const float sw = static_cast<float>(W) / static_cast<float>(w);
  const long long warps = static_cast<long long>(rows) * words_per_row;
  const long long blocks = (warps * 32 + 255) / 256;
  MPF_REQUIRE(blocks < (1ll << 31), "attn_mask_bits: problem too large");
  attn_mask_bits_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      logits, row_stride, rows, H, W, h, w, sh, sw, bits, words_per_row);
  count_launch();
  return finish_launch("attn_mask_bits");
}

int mpf_pack_bool_bits(const uint8_t* src, int rows, int n, uint32_t* bits, int words_per_row, void* stream) {
  using namespace mpf;
  clear_error();
  MPF_REQUIRE(src && bits && rows > 0 && n > 0 && words_per_row * 32 >= n, "pack_bool_bits: bad arguments");
  const long long warps = static_cast<long long>(rows) * words_per_row;
  const long long blocks = (warps * 32 + 255) / 256;
  pack_bool_bits_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, rows, n, bits, words_per_row);
  count_launch();
  return finish_launch("pack_bool_bits");
}

}  // extern "C"

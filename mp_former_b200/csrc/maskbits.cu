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

// Mask-piloted (DN) attention masks: a key (cell of the h x w grid) is masked for a ground-truth instance
// when the cell holds no pixel of the instance,
//   F.interpolate(gt.float(), (h, w), mode="area") <= 1e-8            (ref decoder :986-987, :1593-1594)
// "area" is adaptive average pooling over the window [floor(i*H/h), ceil((i+1)*H/h)) (same for columns); the
// mean of a 0/1 window is <= 1e-8 exactly when no pixel is set (any set pixel gives >= 1/area >> 1e-8), so the
// kernel ORs the window's bytes instead of averaging floats.  One warp per 32 consecutive cells of one mask;
// the lanes' windows are adjacent in memory, 16-byte loads when the window allows.
__global__ void __launch_bounds__(256)
gt_mask_area_bits_kernel(const uint8_t* __restrict__ masks, int rows, int H, int W, int h, int w,
                         uint32_t* __restrict__ bits, int words_per_row) {
  const int lane = threadIdx.x & 31;
  const long long warp_global = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long total = static_cast<long long>(rows) * words_per_row;
  if (warp_global >= total) return;
  const int row = static_cast<int>(warp_global / words_per_row);
  const int word = static_cast<int>(warp_global - static_cast<long long>(row) * words_per_row);
  const int key = word * 32 + lane;
  bool masked = true;                      // padding keys (>= h*w) count as masked
  if (key < h * w) {
    const int oy = key / w, ox = key - oy * w;
    const int y0 = static_cast<int>((static_cast<long long>(oy) * H) / h);
    const int y1 = static_cast<int>((static_cast<long long>(oy + 1) * H + h - 1) / h);
    const int x0 = static_cast<int>((static_cast<long long>(ox) * W) / w);
    const int x1 = static_cast<int>((static_cast<long long>(ox + 1) * W + w - 1) / w);
    const uint8_t* base = masks + static_cast<long long>(row) * H * W;
    uint32_t any = 0;
    for (int y = y0; y < y1 && any == 0; ++y) {
      const uint8_t* p = base + static_cast<long long>(y) * W + x0;
      const uint8_t* e = p + (x1 - x0);
      while (p < e && (reinterpret_cast<uintptr_t>(p) & 15u)) any |= *p++;
      for (; p + 16 <= e; p += 16) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
        any |= v.x | v.y | v.z | v.w;
      }
      while (p < e) any |= *p++;
    }
    masked = any == 0;
  }
  const uint32_t m = __ballot_sync(0xffffffffu, masked);
  if (lane == 0) bits[static_cast<long long>(row) * words_per_row + word] = m;
}

}  // namespace mpf

extern "C" {

int mpf_gt_mask_area_bits(const uint8_t* masks, int rows, int H, int W, int h, int w, uint32_t* bits,
                          int words_per_row, void* stream) {
  using namespace mpf;
  clear_error();
  MPF_REQUIRE(masks && bits, "gt_mask_area_bits: null pointer argument");
  MPF_REQUIRE(rows > 0 && H > 0 && W > 0 && h > 0 && w > 0, "gt_mask_area_bits: sizes must be positive");
  MPF_REQUIRE(words_per_row * 32 >= h * w, "gt_mask_area_bits: words_per_row (%d) too small for %dx%d keys",
              words_per_row, h, w);
  const long long warps = static_cast<long long>(rows) * words_per_row;
  const long long blocks = (warps * 32 + 255) / 256;
  MPF_REQUIRE(blocks < (1ll << 31), "gt_mask_area_bits: problem too large");
  gt_mask_area_bits_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      masks, rows, H, W, h, w, bits, words_per_row);
  count_launch();
  return finish_launch("gt_mask_area_bits");
}

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

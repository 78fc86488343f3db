// MSDeformAttn of the pixel-decoder encoder (queries == pixels) with TMA-staged feature tiles and on-chip gradient
// accumulation -- the design BASELINE.json's north_star names, applied where the measurements say it pays.
//
//   ref: ops/src/cuda/ms_deform_im2col_cuda.cuh:242-304 (forward), :92-164 + :306-408 (backward),
//        ops/modules/ms_deform_attn.py:102-112 (softmax + sampling locations, fused here as in msda.cu's "enc" kernels)
//
// Why (profiles/r2b_lsu_patterns.jsonl, SM cycles per warp instruction with every SM saturated):
//   LDG.128 gather, 4 pixel rows per instruction (msda.cu)          9.0      (bound by L1 misses: 4 CTAs' windows
//                                                                            overflow the L1, 78 % hit rate)
//   LDS.128 gather from a shared-memory tile, 4 rows                3.9-4.4
//   RED.128 global (red.global.add.v4.f32), 4 rows                  16.5
//   4 x red.shared.add.s32 (one 128-byte contribution per 8 lanes)   6.1
// One CTA owns a 16x8 tile of queries of one level, one head, one image.  Its samples on level l fall into a small
// window around the tile's image on that level, so:
//   * phase A computes every sample of the tile for ALL levels (softmax over the L*P logits, loc = ref + off / (W, H),
//     bilinear setup), reduces the bounding box of the sampled pixels per level (REDUX + 12 shared atomics per warp)
//     and writes one 16-byte descriptor per sample;
//   * warp 0 loads the bounding boxes of all levels -- "regions" -- into shared memory with cp.async.bulk.tensor
//     (4-D tensor map [B, H_l, W_l, M*D] per level, box = 8 pixels x 32 channels of the CTA's head; coordinates outside
//     the map are zero-filled by the TMA unit, which IS the op's zero padding), one mbarrier per level, so level l+1
//     lands while level l is being consumed;
//   * phase B gathers with LDS.128 (8 lanes x float4 = the 32 channels of a pixel; conflict free for any 4 pixels);
//   * backward: the gradient of the region is accumulated in shared memory in fixed point (int32; shared-memory fp32
//     atomics are CAS loops, integer ones are native) with a per-CTA power-of-two scale derived from max|grad_out|
//     (a tile's contributions to one texel sum to at most 128 * max|g|: 7 bits of headroom, quantum <= 2^-21 * max|g|,
//     i.e. fp32-grade resolution relative to the largest gradient of the tile) and flushed ONCE per touched texel
//     with red.global.add.v4.f32 -- 5-17x fewer L2 reductions than one per corner;
//   * samples whose footprint leaves the region (huge offsets, or regions clipped by the shared-memory pool: coarse
//     query tiles looking at the finest level) take the L1 gather / global reduction path of msda.cu, per sample.
// Results do not depend on which path a sample takes beyond fp32 rounding of the gradient sums.
//
// MEASURED OUTCOME (round 2, B200, B=16 at 1024^2, profiles/r2e_*): correct on the first run (bit-identical forward,
// memcheck clean) but NOT faster than msda.cu: forward 1.44 ms vs 0.99 ms, backward 3.3 ms vs 2.98 ms.  ncu
// (r2e_ncu_msda_staged_fwd.txt): the LDS gathers cost what the micro-benchmark promised (150 M shared wavefronts =
// 0.54 ms at peak) but the kernel issues 34.6 K warp instructions per CTA, more than half of them in phase A (the
// per-sample softmax / location / bilinear setup that both designs need), and with 100 KB of shared memory only two
// CTAs (16 warps) share an SM: issue slots 45 % busy, LSU 48 %, dominant stalls long-scoreboard (phase A's global
// loads) and fixed-latency waits -- latency-bound, where the L1 kernels hide the same latencies with 32 warps per SM.
// The op is bound by instruction issue + LSU together, not by where the texels live, so staging cannot buy the 2x the
// LSU numbers alone suggest.  The kernels stay in the tree as an option (mpf_msda_set_staged / MPF_MSDA_STAGED=1) and
// as the measured answer to BASELINE.json's "TMA staging of per-level feature tiles"; the default path is msda.cu.
#include "msda_tiling.cuh"
#include "sm100_ptx.cuh"

#include <cuda.h>

#include <climits>
#include <cstdlib>
#include <mutex>

namespace mpf {
namespace stg {

using namespace ptx;

constexpr int kT = 256;                 // threads per CTA
constexpr int kItems = kChunkQ;         // 128 queries per CTA
constexpr int kMaxL = 4;
constexpr int kBoxPx = 8;               // pixels of one TMA box (one row segment, 1 KB)
constexpr int kMaxPitch = 64;           // widest region row in pixels
constexpr int kFwdCapPx = 576;          // forward: 72 KB of staged texels per CTA (2 CTAs / SM)
constexpr int kBwdCapPx = 304;          // backward: 38 KB of texels + 38 KB of gradient accumulators
constexpr int kLPI = 8;                 // lanes per item (D = 32)
constexpr int kD = 32;
constexpr float kMagic = 12582912.f;    // 1.5 * 2^23: float -> int by addition (|x| < 2^22)
constexpr int kMagicBits = 0x4B400000;

struct Maps {
  CUtensorMap m[kMaxL];
};

struct Region {
  int x0, y0, pitch, rows, base;      // texel (x, y) of the level lives at pool[base + (y - y0) * pitch + (x - x0)]
};

__host__ __device__ constexpr int desc_stride(int L) { return 4 * L + 1; }   // odd: conflict-free broadcast reads

__host__ __device__ constexpr size_t smem_bytes(int L, bool bwd) {
  return 128 /* alignment slack */ + static_cast<size_t>(kItems) * desc_stride(L) * 16 +
         static_cast<size_t>(bwd ? 2 * kBwdCapPx : kFwdCapPx) * 128 + (bwd ? kItems * 4 * kMaxL * 4 : 0) + 256;
}

struct Samp {          // one (item, point) sample on one level
  int hl, wl;          // low corner
  float lh, lw, a;     // bilinear fractions, attention weight
  bool gate;           // -1 < h_im < H and -1 < w_im < W (ref cuh:293)
};

// Everything phase A produces, shared by the forward and the backward kernel.
template <bool kBwd>
struct Ctx {
  uint4* desc;
  float4* pool;        // staged texels
  int4* acc;           // backward: gradient accumulators, same geometry as pool
  float* s_ga;         // backward: d(loss)/d(attention weight) [item][L*4]
  int* s_bbox;         // [L][4]: min h, max h, min w, max w of the low corners
  uint64_t* bars;      // [L]
  float* s_misc;       // [0] = max |grad_out| bits (backward)
  Region* s_reg;       // [L] regions, for the level loop of phase B
};

template <bool kBwd>
__device__ __forceinline__ Ctx<kBwd> carve(uint8_t* raw, int L) {
  uint8_t* p = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 127) & ~uintptr_t(127));
  Ctx<kBwd> c;
  c.pool = reinterpret_cast<float4*>(p);
  p += static_cast<size_t>(kBwd ? kBwdCapPx : kFwdCapPx) * 128;
  c.acc = reinterpret_cast<int4*>(p);
  if (kBwd) p += static_cast<size_t>(kBwdCapPx) * 128;
  c.desc = reinterpret_cast<uint4*>(p);
  p += static_cast<size_t>(kItems) * desc_stride(L) * 16;
  c.s_ga = reinterpret_cast<float*>(p);
  if (kBwd) p += kItems * 4 * kMaxL * 4;
  c.bars = reinterpret_cast<uint64_t*>(p);
  c.s_bbox = reinterpret_cast<int*>(p + 64);
  c.s_misc = reinterpret_cast<float*>(p + 64 + 64);
  c.s_reg = reinterpret_cast<Region*>(p + 64 + 64 + 16);
  return c;
}

// Phase A: samples of this thread's two (item, point) slots on every level -> descriptors, regions, TMA loads.
// Returns the regions in `reg` (identical in every thread).
template <bool kBwd>
__device__ __forceinline__ void phase_a(const Ctx<kBwd>& c, const Maps& maps, const MsdaTiling& tiling,
                                        const float* __restrict__ value_img, const float* __restrict__ ow_b,
                                        const float* __restrict__ ref_b, int chunk, int m, int b, int M, int L, int Lq,
                                        bool force_global, Region (&reg)[kMaxL]) {
  const int tid = threadIdx.x, lane = tid & 31;
  const int LP = L * 4, owc = M * LP * 3, MD = M * kD;
  const int cap = kBwd ? kBwdCapPx : kFwdCapPx;
  const int ds = desc_stride(L);

  if (tid < 4 * kMaxL) c.s_bbox[tid] = (tid & 1) ? INT_MIN : INT_MAX;     // (min, max, min, max) per level
  if (tid == 0) {
    for (int l = 0; l < L; ++l) mbar_init(&c.bars[l], 1);
    fence_mbar_init();
  }

  Samp s[2][kMaxL];
  int q[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int idx = tid + k * kT;
    const int p = idx & 3;
    q[k] = query_of(tiling, chunk, idx >> 2, Lq);
    // softmax over the item's L*4 logits: this thread holds point p of every level, its 3 lane neighbours the rest
    float v[kMaxL];
    float mx = -INFINITY;
#pragma unroll
    for (int l = 0; l < kMaxL; ++l) {
      v[l] = -INFINITY;
      if (l < L && q[k] >= 0)
        v[l] = __ldg(ow_b + static_cast<size_t>(q[k]) * owc + M * LP * 2 + m * LP + l * 4 + p);
      mx = fmaxf(mx, v[l]);
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    float sum = 0.f;
#pragma unroll
    for (int l = 0; l < kMaxL; ++l) {
      v[l] = (l < L && q[k] >= 0) ? expf(v[l] - mx) : 0.f;
      sum += v[l];
    }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
#pragma unroll
    for (int l = 0; l < kMaxL; ++l) {
      Samp& t = s[k][l];
      t.gate = false;
      t.hl = t.wl = 0;
      t.lh = t.lw = t.a = 0.f;
      if (l < L && q[k] >= 0) {
        const int H = tiling.H[l], W = tiling.W[l];
        const float2 off = __ldg(reinterpret_cast<const float2*>(ow_b + static_cast<size_t>(q[k]) * owc +
                                                                 (m * LP + l * 4 + p) * 2));
        const float2 rf = __ldg(reinterpret_cast<const float2*>(ref_b + (static_cast<size_t>(q[k]) * L + l) * 2));
        t.a = v[l] / sum;
        // loc = ref + off / (W, H), then pixel = loc * size - 0.5: the reference's operation order
        // (a power-of-two size divides exactly by multiplying with its reciprocal: same bits, ~60 instructions less)
        const float ox = (W & (W - 1)) == 0 ? off.x * (1.f / static_cast<float>(W))
                                            : __fdiv_rn(off.x, static_cast<float>(W));
        const float oy = (H & (H - 1)) == 0 ? off.y * (1.f / static_cast<float>(H))
                                            : __fdiv_rn(off.y, static_cast<float>(H));
        const float lx = rf.x + ox;
        const float ly = rf.y + oy;
        const float h_im = ly * H - 0.5f;
        const float w_im = lx * W - 0.5f;
        if ((h_im > -1.f) && (w_im > -1.f) && (h_im < H) && (w_im < W)) {
          const float hf = floorf(h_im), wf = floorf(w_im);
          t.gate = true;
          t.hl = static_cast<int>(hf);
          t.wl = static_cast<int>(wf);
          t.lh = h_im - hf;
          t.lw = w_im - wf;
        }
      }
    }
  }
  __syncthreads();                                   // bounding boxes initialised, barriers initialised
  // bounding box of the low corners per level: warp reduction, then one shared atomic per warp and bound
#pragma unroll
  for (int l = 0; l < kMaxL; ++l) {
    if (l < L) {
      int h0 = INT_MAX, h1 = INT_MIN, w0 = INT_MAX, w1 = INT_MIN;
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (s[k][l].gate) {
          h0 = min(h0, s[k][l].hl); h1 = max(h1, s[k][l].hl);
          w0 = min(w0, s[k][l].wl); w1 = max(w1, s[k][l].wl);
        }
      h0 = __reduce_min_sync(0xffffffffu, h0); h1 = __reduce_max_sync(0xffffffffu, h1);
      w0 = __reduce_min_sync(0xffffffffu, w0); w1 = __reduce_max_sync(0xffffffffu, w1);
      if (lane == 0 && h0 <= h1) {
        atomicMin(&c.s_bbox[l * 4 + 0], h0); atomicMax(&c.s_bbox[l * 4 + 1], h1);
        atomicMin(&c.s_bbox[l * 4 + 2], w0); atomicMax(&c.s_bbox[l * 4 + 3], w1);
      }
    }
  }
  __syncthreads();
  // regions: the bounding box plus the high corners, rows clipped by what is left of the pool
  int left = cap, base = 0;
#pragma unroll
  for (int l = 0; l < kMaxL; ++l) {
    Region r;
    r.x0 = r.y0 = 0; r.pitch = kBoxPx; r.rows = 0; r.base = base;
    if (l < L && !force_global) {
      const int h0 = c.s_bbox[l * 4 + 0], h1 = c.s_bbox[l * 4 + 1], w0 = c.s_bbox[l * 4 + 2], w1 = c.s_bbox[l * 4 + 3];
      if (h0 <= h1) {
        const int width = min(w1 - w0 + 2, kMaxPitch);
        r.x0 = w0; r.y0 = h0;
        r.pitch = (width + kBoxPx - 1) / kBoxPx * kBoxPx;
        r.rows = min(h1 - h0 + 2, left / r.pitch);
        if (r.rows < 2) r.rows = 0;
      }
    }
    base += r.pitch * r.rows;
    left -= r.pitch * r.rows;
    reg[l] = r;
    if (tid == 0 && l < L) c.s_reg[l] = r;
  }
  // TMA: warp 0 issues every box of every region; one barrier per level
  if (tid < 32) {
#pragma unroll
    for (int l = 0; l < kMaxL; ++l) {
      if (l < L && reg[l].rows > 0) {
        const int bpr = reg[l].pitch / kBoxPx, n = bpr * reg[l].rows;
        if (lane == 0) mbar_arrive_expect_tx(&c.bars[l], static_cast<uint32_t>(n) * kBoxPx * 128u);
        __syncwarp();
        const CUtensorMap* tm = l == 0 ? &maps.m[0] : l == 1 ? &maps.m[1] : l == 2 ? &maps.m[2] : &maps.m[3];
        for (int i = lane; i < n; i += 32) {
          const int y = i / bpr, bx = i - y * bpr;
          tma_load_4d(c.pool + static_cast<size_t>(reg[l].base + y * reg[l].pitch + bx * kBoxPx) * 8, tm, &c.bars[l],
                      m * kD, reg[l].x0 + bx * kBoxPx, reg[l].y0 + y, b);
        }
      }
    }
  }
  // descriptors: word 0 = 0 (no contribution) | ((pool index + 1) << 4) (staged) | (global offset & ~15) | corner flags
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int idx = tid + k * kT;
    const int j = idx >> 2, p = idx & 3;
#pragma unroll
    for (int l = 0; l < kMaxL; ++l) {
      if (l < L) {
        const Samp& t = s[k][l];
        uint4 d = make_uint4(0u, __float_as_uint(t.lh), __float_as_uint(t.lw), __float_as_uint(t.a));
        if (t.gate) {
          const Region& r = reg[l];
          const int xr = t.wl - r.x0, yr = t.hl - r.y0;
          if (r.rows > 0 && xr >= 0 && xr + 1 < r.pitch && yr >= 0 && yr + 1 < r.rows) {
            d.x = static_cast<uint32_t>(r.base + yr * r.pitch + xr + 1) << 4;
          } else {
            const int H = tiling.H[l], W = tiling.W[l];
            const bool top = t.hl >= 0, bot = t.hl + 1 <= H - 1, lft = t.wl >= 0, rgt = t.wl + 1 <= W - 1;
            const int o1 = (t.hl * W + t.wl) * MD;                         // multiple of 32: low bits free
            d.x = static_cast<uint32_t>(o1) | (top && lft ? 1u : 0u) | (top && rgt ? 2u : 0u) |
                  (bot && lft ? 4u : 0u) | (bot && rgt ? 8u : 0u);
          }
        } else if (kBwd) {
          d.w = __float_as_uint(t.a);           // weight still enters the softmax backward
        }
        c.desc[j * ds + l * 4 + p] = d;
      }
    }
  }
  (void)value_img;
  __syncthreads();                                   // descriptors visible
}

// the four corner texels of one sample for this lane (float4 = channels 4*li .. 4*li+3)
struct Corners {
  float4 v1, v2, v3, v4;
};

__device__ __forceinline__ Corners gather(uint32_t w0, const float4* __restrict__ pool, int pitch,
                                          const float* __restrict__ vl, int MD, int W, int li) {
  Corners c;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  if ((w0 & 15u) == 0u) {                            // staged: texels outside the map were zero-filled by the TMA
    const float4* t = pool + (static_cast<size_t>(w0 >> 4) - 1) * 8 + li;
    c.v1 = t[0];
    c.v2 = t[8];
    c.v3 = t[pitch * 8];
    c.v4 = t[pitch * 8 + 8];
  } else {
    const int o1 = static_cast<int>(w0 & ~15u);
    const int rs = W * MD;
    c.v1 = (w0 & 1u) ? ldg4(vl + o1) : z;
    c.v2 = (w0 & 2u) ? ldg4(vl + o1 + MD) : z;
    c.v3 = (w0 & 4u) ? ldg4(vl + o1 + rs) : z;
    c.v4 = (w0 & 8u) ? ldg4(vl + o1 + rs + MD) : z;
  }
  return c;
}

__global__ void __launch_bounds__(kT, 2)
msda_enc_fwd_staged_kernel(const __grid_constant__ Maps maps, const float* __restrict__ value,
                           const float* __restrict__ ow, const float* __restrict__ ref, long long ref_bstride, int S,
                           int M, int L, int Lq, float* __restrict__ out, const MsdaTiling tiling) {
  extern __shared__ uint8_t smem_raw[];
  const Ctx<false> c = carve<false>(smem_raw, L);
  const int chunk = blockIdx.x, m = blockIdx.y, b = blockIdx.z;
  const int slot = threadIdx.x / kLPI, li = threadIdx.x % kLPI;
  const int MD = M * kD, LP = L * 4, owc = M * LP * 3;
  const float* vimg = value + static_cast<size_t>(b) * S * MD + m * kD + li * 4;
  Region reg[kMaxL];
  phase_a<false>(c, maps, tiling, vimg, ow + static_cast<size_t>(b) * Lq * owc, ref + b * ref_bstride, chunk, m, b, M, L,
                 Lq, false, reg);

  constexpr int SLOTS = kT / kLPI, ITERS = kItems / SLOTS;
  const int ds = desc_stride(L);
  float4 acc[ITERS];
#pragma unroll
  for (int it = 0; it < ITERS; ++it) acc[it] = make_float4(0.f, 0.f, 0.f, 0.f);
  // (runtime loops over levels and points keep the kernel inside the instruction cache: fully unrolled, the two
  // gather paths made 9 K instructions here and 28 K in the backward, and both ran slower than the L1 kernels)
#pragma unroll 1
  for (int l = 0; l < L; ++l) {
    const Region r = c.s_reg[l];
    if (r.rows > 0) mbar_wait(&c.bars[l], 0);
    const float* vl = vimg + static_cast<size_t>(tiling.start[l]) * MD;
    const int pitch = r.pitch, W = tiling.W[l];
#pragma unroll 1
    for (int p = 0; p < 4; ++p) {
#pragma unroll
      for (int it = 0; it < ITERS; ++it) {
        const int j = it * SLOTS + slot;
        const uint4 d = c.desc[j * ds + l * 4 + p];
        if (d.x == 0u) continue;
        const float lh = __uint_as_float(d.y), lw = __uint_as_float(d.z), a = __uint_as_float(d.w);
        const float hh = 1.f - lh, hw = 1.f - lw;
        const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
        const Corners g = gather(d.x, c.pool, pitch, vl, MD, W, li);
        acc[it].x = bilinear_acc(acc[it].x, a, w1, w2, w3, w4, g.v1.x, g.v2.x, g.v3.x, g.v4.x);
        acc[it].y = bilinear_acc(acc[it].y, a, w1, w2, w3, w4, g.v1.y, g.v2.y, g.v3.y, g.v4.y);
        acc[it].z = bilinear_acc(acc[it].z, a, w1, w2, w3, w4, g.v1.z, g.v2.z, g.v3.z, g.v4.z);
        acc[it].w = bilinear_acc(acc[it].w, a, w1, w2, w3, w4, g.v1.w, g.v2.w, g.v3.w, g.v4.w);
      }
    }
  }
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int q = query_of(tiling, chunk, it * SLOTS + slot, Lq);
    if (q >= 0)
      *reinterpret_cast<float4*>(out + ((static_cast<size_t>(b) * Lq + q) * M + m) * kD + li * 4) = acc[it];
  }
}

__device__ __forceinline__ void reds_s32(uint32_t addr, int v) {
  asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

__global__ void __launch_bounds__(kT, 2)
msda_enc_bwd_staged_kernel(const __grid_constant__ Maps maps, const float* __restrict__ grad_out,
                           const float* __restrict__ value, const float* __restrict__ ow,
                           const float* __restrict__ ref, long long ref_bstride, int S, int M, int L, int Lq,
                           float* __restrict__ grad_value, float* __restrict__ grad_ow, const MsdaTiling tiling) {
  extern __shared__ uint8_t smem_raw[];
  const Ctx<true> c = carve<true>(smem_raw, L);
  const unsigned FULL = 0xffffffffu;
  const int chunk = blockIdx.x, m = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31;
  const int slot = tid / kLPI, li = tid % kLPI;
  const int MD = M * kD, LP = L * 4, owc = M * LP * 3;
  const size_t img_off = static_cast<size_t>(b) * S * MD + m * kD + li * 4;
  const float* vimg = value + img_off;
  float* gvimg = grad_value + img_off;
  constexpr int SLOTS = kT / kLPI, ITERS = kItems / SLOTS;
  const int ds = desc_stride(L);

  // incoming gradient of this lane's four items + the CTA-wide max |g| that fixes the fixed-point scale
  float4 g[ITERS];
  long long obase[ITERS];
  float gmax = 0.f;
  bool finite = true;
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int q = query_of(tiling, chunk, it * SLOTS + slot, Lq);
    obase[it] = q >= 0 ? (static_cast<long long>(b) * Lq + q) * owc : -1;
    g[it] = q >= 0 ? ldg4(grad_out + ((static_cast<size_t>(b) * Lq + q) * M + m) * kD + li * 4)
                   : make_float4(0.f, 0.f, 0.f, 0.f);
    const float mx = fmaxf(fmaxf(fabsf(g[it].x), fabsf(g[it].y)), fmaxf(fabsf(g[it].z), fabsf(g[it].w)));
    finite = finite && (mx <= 3.0e38f);             // false for inf and NaN
    gmax = fmaxf(gmax, mx);
  }
  // zero the accumulators (generic proxy only; the TMA never touches them)
  for (int i = tid; i < kBwdCapPx * 8; i += kT) c.acc[i] = make_int4(0, 0, 0, 0);
  if (tid == 0) c.s_misc[0] = 0.f, c.s_misc[1] = 0.f;
  __syncthreads();
  {
    unsigned bits = __float_as_uint(gmax);           // non-negative floats order like their bit patterns
    bits = __reduce_max_sync(FULL, bits);
    const unsigned bad = __reduce_or_sync(FULL, finite ? 0u : 1u);
    if (lane == 0) {
      atomicMax(reinterpret_cast<unsigned*>(&c.s_misc[0]), bits);
      if (bad) atomicOr(reinterpret_cast<unsigned*>(&c.s_misc[1]), 1u);
    }
  }
  __syncthreads();
  const float gmax_cta = c.s_misc[0];
  // a NaN / inf gradient must propagate like fp32 atomics would: such a tile takes the global path throughout
  const bool force_global = (__float_as_uint(c.s_misc[1]) != 0u) || !(gmax_cta > 0.f) || gmax_cta < 1e-30f;
  // scale = 2^(21 - e) with 2^e <= gmax < 2^(e+1): one contribution (< gmax) stays below 2^22 (exact float -> int by
  // magic addition), a texel's sum (< 128 gmax) below 2^29
  const int e = static_cast<int>((__float_as_uint(gmax_cta) >> 23) & 0xff) - 127;
  const float scale = force_global ? 1.f : __uint_as_float(static_cast<unsigned>(21 - e + 127) << 23);
  const float inv_scale = force_global ? 1.f : __uint_as_float(static_cast<unsigned>(e - 21 + 127) << 23);

  Region reg[kMaxL];
  phase_a<true>(c, maps, tiling, vimg, ow + static_cast<size_t>(b) * Lq * owc, ref + b * ref_bstride, chunk, m, b, M, L,
                Lq, force_global, reg);

  // bank rotation of the four scalar atomics: the 4 items of a warp instruction target 4 texels; item gi uses
  // channel (k ^ gi) of its lane quad in instruction k, so the four items never meet in a bank
  const int gi = lane >> 3;
  const uint32_t acc_base = smem_u32(c.acc) + static_cast<uint32_t>(li) * 16u;
  uint32_t rot[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) rot[k] = 4u * static_cast<uint32_t>(k ^ gi);

#pragma unroll 1
  for (int l = 0; l < L; ++l) {
    {
      const Region rl = c.s_reg[l];
      if (rl.rows > 0) mbar_wait(&c.bars[l], 0);
      const size_t loff = static_cast<size_t>(tiling.start[l]) * MD;
      const float* vl = vimg + loff;
      float* gvl = gvimg + loff;
      const int pitch = rl.pitch, W = tiling.W[l];
#pragma unroll
      for (int it = 0; it < ITERS; ++it) {
        const int j = it * SLOTS + slot;
        float mine_x = 0.f, mine_y = 0.f, mine_a = 0.f;
        // incoming gradient pre-scaled and permuted for the atomics: ts[k] = scale * g[k ^ gi]
        const float s0 = g[it].x * scale, s1 = g[it].y * scale, s2 = g[it].z * scale, s3 = g[it].w * scale;
        const float u0 = (gi & 1) ? s1 : s0, u1 = (gi & 1) ? s0 : s1, u2 = (gi & 1) ? s3 : s2, u3 = (gi & 1) ? s2 : s3;
        const float ts0 = (gi & 2) ? u2 : u0, ts1 = (gi & 2) ? u3 : u1, ts2 = (gi & 2) ? u0 : u2,
                    ts3 = (gi & 2) ? u1 : u3;
#pragma unroll 1
        for (int p = 0; p < 4; ++p) {
          const uint4 d = c.desc[j * ds + l * 4 + p];
          float gh = 0.f, gw = 0.f, ga = 0.f;
          if (d.x != 0u) {                                   // (uniform over the 8 lanes of an item)
            const float lh = __uint_as_float(d.y), lw = __uint_as_float(d.z), wgt = __uint_as_float(d.w);
            const float hh = 1.f - lh, hw = 1.f - lw;
            const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
            const Corners v = gather(d.x, c.pool, pitch, vl, MD, W, li);
            const float4 tg = make_float4(g[it].x * wgt, g[it].y * wgt, g[it].z * wgt, g[it].w * wgt);  // ref cuh:116
            if ((d.x & 15u) == 0u) {
              // fixed-point accumulation in the region: corner weight * attention weight folded into one factor
              const uint32_t a0 = acc_base + ((d.x >> 4) - 1u) * 128u;
              const float f1 = w1 * wgt, f2 = w2 * wgt, f3 = w3 * wgt, f4 = w4 * wgt;
#define MPF_CORNER(addr, f)                                                                     \
  {                                                                                             \
    reds_s32((addr) + rot[0], __float_as_int(__fmaf_rn((f), ts0, kMagic)) - kMagicBits);        \
    reds_s32((addr) + rot[1], __float_as_int(__fmaf_rn((f), ts1, kMagic)) - kMagicBits);        \
    reds_s32((addr) + rot[2], __float_as_int(__fmaf_rn((f), ts2, kMagic)) - kMagicBits);        \
    reds_s32((addr) + rot[3], __float_as_int(__fmaf_rn((f), ts3, kMagic)) - kMagicBits);        \
  }
              MPF_CORNER(a0, f1)
              MPF_CORNER(a0 + 128u, f2)
              MPF_CORNER(a0 + static_cast<uint32_t>(pitch) * 128u, f3)
              MPF_CORNER(a0 + static_cast<uint32_t>(pitch) * 128u + 128u, f4)
#undef MPF_CORNER
            } else {
              const int o1 = static_cast<int>(d.x & ~15u), rs = W * MD;
              if (d.x & 1u) red_add_v4(gvl + o1, w1 * tg.x, w1 * tg.y, w1 * tg.z, w1 * tg.w);
              if (d.x & 2u) red_add_v4(gvl + o1 + MD, w2 * tg.x, w2 * tg.y, w2 * tg.z, w2 * tg.w);
              if (d.x & 4u) red_add_v4(gvl + o1 + rs, w3 * tg.x, w3 * tg.y, w3 * tg.z, w3 * tg.w);
              if (d.x & 8u) red_add_v4(gvl + o1 + rs + MD, w4 * tg.x, w4 * tg.y, w4 * tg.z, w4 * tg.w);
            }
#define MPF_ACC(comp)                                                                                  \
  {                                                                                                    \
    const float ghw = -hw * v.v1.comp - lw * v.v2.comp + hw * v.v3.comp + lw * v.v4.comp;              \
    const float gww = -hh * v.v1.comp + hh * v.v2.comp - lh * v.v3.comp + lh * v.v4.comp;              \
    const float val = w1 * v.v1.comp + w2 * v.v2.comp + w3 * v.v3.comp + w4 * v.v4.comp;               \
    gh += ghw * tg.comp;                                                                               \
    gw += gww * tg.comp;                                                                               \
    ga += val * g[it].comp;                                                                            \
  }
            MPF_ACC(x) MPF_ACC(y) MPF_ACC(z) MPF_ACC(w)
#undef MPF_ACC
          }
#pragma unroll
          for (int off = kLPI / 2; off >= 1; off >>= 1) {
            gh += __shfl_xor_sync(FULL, gh, off);
            gw += __shfl_xor_sync(FULL, gw, off);
            ga += __shfl_xor_sync(FULL, ga, off);
          }
          // d/d(off) = d/d(loc) / (W, H) = (W*gw)/W ...: the W, H factors of ref cuh:162-163 cancel
          if (li == p) { mine_x = gw; mine_y = gh; mine_a = ga; }
        }
        if (obase[it] >= 0 && li < 4) {
          *reinterpret_cast<float2*>(grad_ow + obase[it] + (m * LP + l * 4 + li) * 2) = make_float2(mine_x, mine_y);
          c.s_ga[j * (4 * kMaxL) + l * 4 + li] = mine_a;
        }
      }
    }
  }
  __syncthreads();
  // flush: every touched texel of every region, once (texels of the halo outside the map are skipped)
#pragma unroll 1
  for (int l = 0; l < L; ++l) {
    const Region r = c.s_reg[l];
    if (r.rows > 0) {
      const int H = tiling.H[l], W = tiling.W[l];
      float* gvl = gvimg + static_cast<size_t>(tiling.start[l]) * MD;
      const int npx = r.pitch * r.rows;
      for (int i = slot; i < npx; i += SLOTS) {
        const int yr = i / r.pitch, xr = i - yr * r.pitch;
        const int y = r.y0 + yr, x = r.x0 + xr;
        const int4 a = c.acc[static_cast<size_t>(r.base + i) * 8 + li];
        if ((a.x | a.y | a.z | a.w) != 0 && y >= 0 && y < H && x >= 0 && x < W)
          red_add_v4(gvl + (static_cast<size_t>(y) * W + x) * MD, a.x * inv_scale, a.y * inv_scale, a.z * inv_scale,
                     a.w * inv_scale);
      }
    }
  }
  // softmax backward per item: d(logit_i) = aw_i * (ga_i - sum_j aw_j ga_j); the weights are the descriptors' .w
  if (tid < kItems) {
    const int j = tid;
    const int q = query_of(tiling, chunk, j, Lq);
    if (q >= 0) {
      float dot = 0.f;
      for (int i = 0; i < LP; ++i) dot += __uint_as_float(c.desc[j * ds + i].w) * c.s_ga[j * (4 * kMaxL) + i];
      float* gl = grad_ow + (static_cast<long long>(b) * Lq + q) * owc + M * LP * 2 + m * LP;
      for (int i = 0; i < LP; ++i)
        gl[i] = __uint_as_float(c.desc[j * ds + i].w) * (c.s_ga[j * (4 * kMaxL) + i] - dot);
    }
  }
}

// ---- host ------------------------------------------------------------------------------------------------------
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                              const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn encoder() {
  static EncodeFn enc = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      enc = reinterpret_cast<EncodeFn>(p);
  });
  static thread_local bool ctx_bound = false;     // the encoder needs a current context (autograd worker threads)
  if (!ctx_bound) {
    cudaFree(nullptr);
    ctx_bound = true;
  }
  return enc;
}

// value of level l as the 4-D map [B, H_l, W_l, M*D] (channels contiguous); box = 32 channels x 8 pixels of one row,
// no swizzle (texel p of a box lands at p * 128 bytes), out-of-map coordinates are zero-filled.
static int make_maps(Maps* maps, const float* value, const MsdaTiling& t, int B, int S, int M) {
  EncodeFn enc = encoder();
  if (enc == nullptr) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return MPF_ERR_UNSUPPORTED;
  }
  const cuuint64_t MD = static_cast<cuuint64_t>(M) * kD;
  for (int l = 0; l < t.L; ++l) {
    cuuint64_t dims[4] = {MD, static_cast<cuuint64_t>(t.W[l]), static_cast<cuuint64_t>(t.H[l]),
                          static_cast<cuuint64_t>(B)};
    cuuint64_t strides[3] = {MD * 4, MD * 4 * static_cast<cuuint64_t>(t.W[l]), MD * 4 * static_cast<cuuint64_t>(S)};
    cuuint32_t box[4] = {kD, kBoxPx, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&maps->m[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                     const_cast<float*>(value + static_cast<size_t>(t.start[l]) * MD), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("msda_staged: cuTensorMapEncodeTiled failed (CUresult %d) for level %d (%d x %d)", static_cast<int>(r),
                l, t.H[l], t.W[l]);
      return MPF_ERR_BAD_ARG;
    }
  }
  for (int l = t.L; l < kMaxL; ++l) maps->m[l] = maps->m[0];
  return MPF_OK;
}

}  // namespace stg

// True when the staged kernels cover this launch (encoder self-attention: queries == pixels with host shapes, D = 32,
// P = 4, L <= 4).  MPF_MSDA_STAGED=0 keeps the L1-gather kernels of msda.cu (A/B measurements).
// Default: OFF.  Measured on the B200 (profiles/r2e_msda_enc_probe.jsonl, B=16, 1024^2): staged forward 1.44 ms vs
// 0.99 ms, staged backward 3.3 ms vs 2.98 ms -- see the note at the end of the header comment.
static int g_staged = [] { const char* e = getenv("MPF_MSDA_STAGED"); return (e != nullptr && e[0] == '1') ? 1 : 0; }();

bool msda_staged_ok(const MsdaTiling& t, int D, int L, int P, int M, int B) {
  if (!g_staged || t.mode != 1 || D != stg::kD || P != 4 || L > stg::kMaxL || M > 65535 || B > 65535) return false;
  for (int l = 0; l < L; ++l)
    if (static_cast<long long>(t.H[l]) * t.W[l] * M * D >= (1ll << 31)) return false;   // 32-bit texel offsets
  return true;
}

int msda_enc_forward_staged(const float* value, const float* ow, const float* ref, long long ref_bstride, int B, int S,
                            int M, int L, int Lq, float* out, const MsdaTiling& t, cudaStream_t st) {
  stg::Maps maps;
  int rc = stg::make_maps(&maps, value, t, B, S, M);
  if (rc) return rc;
  const size_t smem = stg::smem_bytes(L, false);
  static unsigned long long configured_on = 0;
  if (first_use_on_this_device(configured_on))
    MPF_CUDA_OK(cudaFuncSetAttribute(stg::msda_enc_fwd_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(stg::smem_bytes(stg::kMaxL, false))));
  dim3 grid(t.num_chunks, M, B);
  stg::msda_enc_fwd_staged_kernel<<<grid, stg::kT, smem, st>>>(maps, value, ow, ref, ref_bstride, S, M, L, Lq, out, t);
  count_launch();
  return finish_launch("msda_enc_fwd_staged");
}

int msda_enc_backward_staged(const float* grad_out, const float* value, const float* ow, const float* ref,
                             long long ref_bstride, int B, int S, int M, int L, int Lq, float* gv, float* gow,
                             const MsdaTiling& t, cudaStream_t st) {
  stg::Maps maps;
  int rc = stg::make_maps(&maps, value, t, B, S, M);
  if (rc) return rc;
  const size_t smem = stg::smem_bytes(L, true);
  static unsigned long long configured_on = 0;
  if (first_use_on_this_device(configured_on))
    MPF_CUDA_OK(cudaFuncSetAttribute(stg::msda_enc_bwd_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(stg::smem_bytes(stg::kMaxL, true))));
  dim3 grid(t.num_chunks, M, B);
  stg::msda_enc_bwd_staged_kernel<<<grid, stg::kT, smem, st>>>(maps, grad_out, value, ow, ref, ref_bstride, S, M, L, Lq,
                                                               gv, gow, t);
  count_launch();
  return finish_launch("msda_enc_bwd_staged");
}

}  // namespace mpf

extern "C" int mpf_msda_set_staged(int enabled) {
  const int prev = mpf::g_staged;
  if (enabled >= 0) mpf::g_staged = enabled ? 1 : 0;
  return prev;
}

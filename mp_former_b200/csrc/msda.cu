// Multi-scale deformable attention (MSDeformAttn) forward / backward for sm_100a.
//
// Semantics follow the reference op (ref: mask2former/modeling/pixel_decoder/ops/src/cuda/
// ms_deform_im2col_cuda.cuh:38-89 bilinear + zero padding, :242-304 forward, :92-164 backward):
//   (w_im, h_im) = loc * (W_l, H_l) - 0.5 ; a sample contributes only if -1 < h_im < H_l and
//   -1 < w_im < W_l ; each of the 4 corners is individually zero-padded.
//
// Design (not a translation of the reference's 1-thread-per-scalar kernels):
//   * "vec" kernels: the D channels of one (b, q, m) item are covered by LPI = D/4 lanes, each
//     holding a float4, so every corner gather is one coalesced 16-byte-per-lane request
//     (LPI lanes = one 128-byte line for D = 32).  A 256-thread CTA therefore works on 256/LPI
//     items at a time.
//   * Per level, the coordinate arithmetic of every (item, point) sample is done ONCE by one thread
//     and staged in shared memory as a descriptor (corner offsets + weights); the item's lanes then
//     only gather and FMA (see the comment above the vectorised kernels).
//   * A CTA owns one (batch, head, chunk-of-128-queries) unit.  When the queries are known to be
//     the pixel grid itself (encoder self-attention: num_query == spatial_size and host shapes are
//     provided) chunks are 16x8 spatial tiles of one level, so the gather footprint of a CTA is a
//     compact 2-D window per level that stays L1-resident.  The order in which queries are
//     processed never changes results.
//   * Backward: per-sample partial d/d(loc), d/d(weight) are reduced over the item's lanes with
//     shuffles; grad_value uses 16-byte vector reductions (red.global.add.v4.f32).
//   * "generic" kernels (any D / L / P, float or double) keep one thread per output scalar and are
//     used for shapes outside the fast path (e.g. the reference's test.py geometries, fp64
//     gradcheck).
#include "msda_tiling.cuh"

#include <cstdlib>

namespace mpf {

// ------------------------------------------------------------------------------------------------
// Vectorised kernels: D = 4*LPI channels, P = 4 points, L <= 8 levels.
//
// Per level, phase A lets every thread turn two (item, point) samples of the CTA's 128-query chunk into
// a descriptor in shared memory -- four corner offsets (-1 = corner outside the map / sample outside
// the gate) plus the weights -- so the coordinate arithmetic is done once per sample instead of once
// per lane; phase B is then pure gather + FMA: two broadcast LDS.128 and four 16-byte gathers per
// sample and lane.  (ncu on the first version showed the kernel issue-bound, 68% of issue slots,
// with only 52% of the L1 data pipe in use: see profiles/r1a_*.)
// ------------------------------------------------------------------------------------------------
constexpr int kDescStride = 5;   // 16-byte slots per item: 4 points + 1 pad so that the 4 items a warp
                                 // reads in one LDS.128 fall into distinct bank groups

// The two (item, point) samples a thread prepares per level: thread t owns chunk samples t and t + 256.
// Their locations / weights for level l+1 are fetched while phase B of level l runs.
struct MySamples {
  long long sbase[2];      // ((b*Lq + q)*M + m)*LP + p, or -1 if the item is padding
  float2 xy[2];
  float a[2];
};

__device__ __forceinline__ void my_samples_init(MySamples& ms, const MsdaTiling& tiling, int chunk, int b, int m,
                                                int M, int Lq, int LP) {
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int idx = threadIdx.x + k * kThreads;
    const int q = query_of(tiling, chunk, idx >> 2, Lq);
    ms.sbase[k] = q >= 0 ? ((static_cast<long long>(b) * Lq + q) * M + m) * LP + (idx & 3) : -1;
    ms.xy[k] = make_float2(0.f, 0.f);
    ms.a[k] = 0.f;
  }
}

__device__ __forceinline__ void my_samples_fetch(MySamples& ms, const float* __restrict__ loc,
                                                 const float* __restrict__ aw, int l) {
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    if (ms.sbase[k] >= 0) {
      const long long s = ms.sbase[k] + l * 4;
      ms.xy[k] = __ldg(reinterpret_cast<const float2*>(loc + 2 * s));
      ms.a[k] = __ldg(aw + s);
    }
  }
}

template <bool kBackward>
__device__ __forceinline__ void build_descriptors(int4* so, float4* sw, const MySamples& ms, int H, int W, int MD) {
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int idx = threadIdx.x + k * kThreads;
    const int j = idx >> 2, p = idx & 3;
    int4 offs = make_int4(-1, -1, -1, -1);
    float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ms.sbase[k] >= 0) {
      const float2 xy = ms.xy[k];
      const float a = ms.a[k];
      const float h_im = xy.y * H - 0.5f;
      const float w_im = xy.x * W - 0.5f;
      if ((h_im > -1.f) && (w_im > -1.f) && (h_im < H) && (w_im < W)) {
        const float hf = floorf(h_im), wf = floorf(w_im);
        const int h_low = static_cast<int>(hf), w_low = static_cast<int>(wf);
        const float lh = h_im - hf, lw = w_im - wf;
        const bool top = h_low >= 0, bot = h_low + 1 <= H - 1, lft = w_low >= 0, rgt = w_low + 1 <= W - 1;
        const int rs = W * MD;
        const int o1 = h_low * rs + w_low * MD;
        offs.x = (top && lft) ? o1 : -1;
        offs.y = (top && rgt) ? o1 + MD : -1;
        offs.z = (bot && lft) ? o1 + rs : -1;
        offs.w = (bot && rgt) ? o1 + rs + MD : -1;
        wv = make_float4(lh, lw, a, 0.f);        // forward and backward rebuild the four corner weights from these
      } else if (kBackward) {
        wv.z = a;
      }
    }
    so[j * kDescStride + p] = offs;
    sw[j * kDescStride + p] = wv;
  }
}

template <int LPI>
__global__ void __launch_bounds__(kThreads, 4)
msda_fwd_vec_kernel(const float* __restrict__ value, const int64_t* __restrict__ shapes,
                    const int64_t* __restrict__ lstart, const float* __restrict__ loc,
                    const float* __restrict__ aw, int S, int M, int L, int Lq,
                    float* __restrict__ out, const MsdaTiling tiling) {
  constexpr int D = LPI * 4;
  constexpr int SLOTS = kThreads / LPI;
  constexpr int ITERS = kChunkQ / SLOTS;
  __shared__ int4 so[kChunkQ * kDescStride];
  __shared__ float4 sw[kChunkQ * kDescStride];

  const int chunk = blockIdx.x, m = blockIdx.y, b = blockIdx.z;
  const int slot = threadIdx.x / LPI, li = threadIdx.x % LPI;
  const int MD = M * D, LP = L * 4;
  const float* vimg = value + static_cast<size_t>(b) * S * MD + m * D + li * 4;

  float4 acc[ITERS];
#pragma unroll
  for (int it = 0; it < ITERS; ++it) acc[it] = make_float4(0.f, 0.f, 0.f, 0.f);

  MySamples ms;
  my_samples_init(ms, tiling, chunk, b, m, M, Lq, LP);
  my_samples_fetch(ms, loc, aw, 0);
  for (int l = 0; l < L; ++l) {
    const int H = static_cast<int>(shapes[2 * l]), W = static_cast<int>(shapes[2 * l + 1]);
    const float* vl = vimg + static_cast<size_t>(lstart[l]) * MD;
    if (l > 0) __syncthreads();          // phase B of the previous level has finished reading
    build_descriptors<false>(so, sw, ms, H, W, MD);
    if (l + 1 < L) my_samples_fetch(ms, loc, aw, l + 1);   // in flight during phase B
    __syncthreads();
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int j = it * SLOTS + slot;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int4 o = so[j * kDescStride + p];
        const float4 d = sw[j * kDescStride + p];          // (lh, lw, attention weight, -)
        if ((o.x & o.y & o.z & o.w) < 0) continue;         // sample outside the gate: the reference adds nothing
        const float hh = 1.f - d.x, hw = 1.f - d.y;
        const float w1 = hh * hw, w2 = hh * d.y, w3 = d.x * hw, w4 = d.x * d.y;
        const float4 v1 = o.x >= 0 ? ldg4(vl + o.x) : z;
        const float4 v2 = o.y >= 0 ? ldg4(vl + o.y) : z;
        const float4 v3 = o.z >= 0 ? ldg4(vl + o.z) : z;
        const float4 v4 = o.w >= 0 ? ldg4(vl + o.w) : z;
        acc[it].x = bilinear_acc(acc[it].x, d.z, w1, w2, w3, w4, v1.x, v2.x, v3.x, v4.x);
        acc[it].y = bilinear_acc(acc[it].y, d.z, w1, w2, w3, w4, v1.y, v2.y, v3.y, v4.y);
        acc[it].z = bilinear_acc(acc[it].z, d.z, w1, w2, w3, w4, v1.z, v2.z, v3.z, v4.z);
        acc[it].w = bilinear_acc(acc[it].w, d.z, w1, w2, w3, w4, v1.w, v2.w, v3.w, v4.w);
      }
    }
  }
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int q = query_of(tiling, chunk, it * SLOTS + slot, Lq);
    if (q >= 0)
      *reinterpret_cast<float4*>(out + ((static_cast<size_t>(b) * Lq + q) * M + m) * D + li * 4) = acc[it];
  }
}

template <int LPI>
__global__ void __launch_bounds__(kThreads, 3)
msda_bwd_vec_kernel(const float* __restrict__ grad_out, const float* __restrict__ value,
                    const int64_t* __restrict__ shapes, const int64_t* __restrict__ lstart,
                    const float* __restrict__ loc, const float* __restrict__ aw, int S, int M, int L,
                    int Lq, float* __restrict__ grad_value, float* __restrict__ grad_loc,
                    float* __restrict__ grad_aw, const MsdaTiling tiling) {
  constexpr int D = LPI * 4;
  constexpr int SLOTS = kThreads / LPI;
  constexpr int ITERS = kChunkQ / SLOTS;
  const unsigned FULL = 0xffffffffu;
  __shared__ int4 so[kChunkQ * kDescStride];
  __shared__ float4 sw[kChunkQ * kDescStride];

  const int chunk = blockIdx.x, m = blockIdx.y, b = blockIdx.z;
  const int slot = threadIdx.x / LPI, li = threadIdx.x % LPI;
  const int MD = M * D, LP = L * 4;
  const size_t img_off = static_cast<size_t>(b) * S * MD + m * D + li * 4;
  const float* vimg = value + img_off;
  float* gvimg = grad_value + img_off;

  float4 g[ITERS];
  size_t item[ITERS];
  bool active[ITERS];
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int q = query_of(tiling, chunk, it * SLOTS + slot, Lq);
    active[it] = q >= 0;
    item[it] = (static_cast<size_t>(b) * Lq + (active[it] ? q : 0)) * M + m;
    g[it] = active[it] ? ldg4(grad_out + item[it] * D + li * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  }

  MySamples ms;
  my_samples_init(ms, tiling, chunk, b, m, M, Lq, LP);
  my_samples_fetch(ms, loc, aw, 0);
  for (int l = 0; l < L; ++l) {
    const int H = static_cast<int>(shapes[2 * l]), W = static_cast<int>(shapes[2 * l + 1]);
    const size_t loff = static_cast<size_t>(lstart[l]) * MD;
    const float* vl = vimg + loff;
    float* gvl = gvimg + loff;
    if (l > 0) __syncthreads();
    build_descriptors<true>(so, sw, ms, H, W, MD);
    if (l + 1 < L) my_samples_fetch(ms, loc, aw, l + 1);
    __syncthreads();
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int j = it * SLOTS + slot;
      float mine_x = 0.f, mine_y = 0.f, mine_a = 0.f;      // results of point p == li (lanes 0..3)
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int4 o = so[j * kDescStride + p];
        const float4 d = sw[j * kDescStride + p];         // (lh, lw, attention weight, -)
        const float lh = d.x, lw = d.y, wgt = d.z;
        const float hh = 1.f - lh, hw = 1.f - lw;
        const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
        const float4 v1 = o.x >= 0 ? ldg4(vl + o.x) : z;
        const float4 v2 = o.y >= 0 ? ldg4(vl + o.y) : z;
        const float4 v3 = o.z >= 0 ? ldg4(vl + o.z) : z;
        const float4 v4 = o.w >= 0 ? ldg4(vl + o.w) : z;
        const float4 tg = make_float4(g[it].x * wgt, g[it].y * wgt, g[it].z * wgt, g[it].w * wgt);  // ref cuh:116
        if (o.x >= 0) red_add_v4(gvl + o.x, w1 * tg.x, w1 * tg.y, w1 * tg.z, w1 * tg.w);
        if (o.y >= 0) red_add_v4(gvl + o.y, w2 * tg.x, w2 * tg.y, w2 * tg.z, w2 * tg.w);
        if (o.z >= 0) red_add_v4(gvl + o.z, w3 * tg.x, w3 * tg.y, w3 * tg.z, w3 * tg.w);
        if (o.w >= 0) red_add_v4(gvl + o.w, w4 * tg.x, w4 * tg.y, w4 * tg.z, w4 * tg.w);
        float gh = 0.f, gw = 0.f, ga = 0.f;                // ref cuh:119-161, per channel then summed
#define MPF_ACC(comp)                                                                         \
  {                                                                                           \
    const float ghw = -hw * v1.comp - lw * v2.comp + hw * v3.comp + lw * v4.comp;             \
    const float gww = -hh * v1.comp + hh * v2.comp - lh * v3.comp + lh * v4.comp;             \
    const float val = w1 * v1.comp + w2 * v2.comp + w3 * v3.comp + w4 * v4.comp;              \
    gh += ghw * tg.comp;                                                                      \
    gw += gww * tg.comp;                                                                      \
    ga += val * g[it].comp;                                                                   \
  }
        MPF_ACC(x) MPF_ACC(y) MPF_ACC(z) MPF_ACC(w)
#undef MPF_ACC
#pragma unroll
        for (int off = LPI / 2; off >= 1; off >>= 1) {
          gh += __shfl_xor_sync(FULL, gh, off);
          gw += __shfl_xor_sync(FULL, gw, off);
          ga += __shfl_xor_sync(FULL, ga, off);
        }
        if (li == p) { mine_x = W * gw; mine_y = H * gh; mine_a = ga; }   // ref cuh:162-163
      }
      if (active[it] && li < 4) {
        const size_t s = item[it] * LP + l * 4 + li;
        *reinterpret_cast<float2*>(grad_loc + 2 * s) = make_float2(mine_x, mine_y);
        grad_aw[s] = mine_a;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Encoder-fused variants: the kernels consume the raw output ``ow`` of the offsets+logits projection
// ([B, Lq, M*L*P*3]: offsets first, then attention logits) and the reference points, i.e. the softmax over
// the L*P logits and the location arithmetic  loc = ref + off / (W_l, H_l)  (ref ops/modules/
// ms_deform_attn.py:102-109) happen in the descriptor builder instead of five elementwise passes, two
// slicing copies and their backward counterparts.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxEncLevels = 4;
constexpr int kMaxLP = 4 * kMaxEncLevels;     // L*P of the fused path (L <= 4, P = 4)

struct EncSamples {
  int q[2];                    // query of slot (threadIdx.x + k * kThreads) >> 2, or -1 if padding
  float2 off[2], ref[2];
};

__device__ __forceinline__ void enc_samples_init(EncSamples& es, const MsdaTiling& tiling, int chunk, int Lq) {
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int idx = threadIdx.x + k * kThreads;
    es.q[k] = query_of(tiling, chunk, idx >> 2, Lq);
    es.off[k] = es.ref[k] = make_float2(0.f, 0.f);
  }
}

// ow_b / ref_b: this image's slices of ow [Lq, owc] and ref [Lq, L, 2]
__device__ __forceinline__ void enc_samples_fetch(EncSamples& es, const float* __restrict__ ow_b,
                                                  const float* __restrict__ ref_b, int m, int L, int owc, int l) {
  const int LP = L * 4;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    if (es.q[k] >= 0) {
      const int p = (threadIdx.x + k * kThreads) & 3;
      es.off[k] = __ldg(reinterpret_cast<const float2*>(ow_b + static_cast<size_t>(es.q[k]) * owc +
                                                        (m * LP + l * 4 + p) * 2));
      es.ref[k] = __ldg(reinterpret_cast<const float2*>(ref_b + (static_cast<size_t>(es.q[k]) * L + l) * 2));
    }
  }
}

// softmax over each item's L*P logits -> s_aw[item][LP].  The four lanes that own an item's four points hold the
// logits of one point each (one per level, L <= 4) and combine maximum and sum with two shuffles.
__device__ __forceinline__ void enc_softmax(float* s_aw, const EncSamples& es, const float* __restrict__ ow_b, int m,
                                            int M, int L, int owc) {
  const int LP = L * 4;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int idx = threadIdx.x + k * kThreads;
    const int j = idx >> 2, p = idx & 3;
    float v[kMaxEncLevels];
    float mx = -INFINITY;
#pragma unroll
    for (int l = 0; l < kMaxEncLevels; ++l) {
      v[l] = -INFINITY;
      if (l < L && es.q[k] >= 0) v[l] = __ldg(ow_b + static_cast<size_t>(es.q[k]) * owc + M * LP * 2 + m * LP + l * 4 + p);
      mx = fmaxf(mx, v[l]);
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    float sum = 0.f;
#pragma unroll
    for (int l = 0; l < kMaxEncLevels; ++l) {
      v[l] = (l < L && es.q[k] >= 0) ? expf(v[l] - mx) : 0.f;
      sum += v[l];
    }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    if (es.q[k] >= 0) {
#pragma unroll
      for (int l = 0; l < kMaxEncLevels; ++l)
        if (l < L) s_aw[j * kMaxLP + l * 4 + p] = v[l] / sum;
    }
  }
}

template <bool kBackward>
__device__ __forceinline__ void enc_build_descriptors(int4* so, float4* sw, const EncSamples& es, const float* s_aw,
                                                      int l, int H, int W, int MD) {
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int idx = threadIdx.x + k * kThreads;
    const int j = idx >> 2, p = idx & 3;
    int4 offs = make_int4(-1, -1, -1, -1);
    float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (es.q[k] >= 0) {
      const float a = s_aw[j * kMaxLP + l * 4 + p];
      // loc = ref + off / (W, H), then pixel = loc * size - 0.5: the reference's operation order
      const float lx = es.ref[k].x + __fdiv_rn(es.off[k].x, static_cast<float>(W));
      const float ly = es.ref[k].y + __fdiv_rn(es.off[k].y, static_cast<float>(H));
      const float h_im = ly * H - 0.5f;
      const float w_im = lx * W - 0.5f;
      if ((h_im > -1.f) && (w_im > -1.f) && (h_im < H) && (w_im < W)) {
        const float hf = floorf(h_im), wf = floorf(w_im);
        const int h_low = static_cast<int>(hf), w_low = static_cast<int>(wf);
        const float lh = h_im - hf, lw = w_im - wf;
        const bool top = h_low >= 0, bot = h_low + 1 <= H - 1, lft = w_low >= 0, rgt = w_low + 1 <= W - 1;
        const int rs = W * MD;
        const int o1 = h_low * rs + w_low * MD;
        offs.x = (top && lft) ? o1 : -1;
        offs.y = (top && rgt) ? o1 + MD : -1;
        offs.z = (bot && lft) ? o1 + rs : -1;
        offs.w = (bot && rgt) ? o1 + rs + MD : -1;
        wv = make_float4(lh, lw, a, 0.f);        // forward and backward rebuild the four corner weights from these
      } else if (kBackward) {
        wv.z = a;
      }
    }
    so[j * kDescStride + p] = offs;
    sw[j * kDescStride + p] = wv;
  }
}

template <int LPI>
__global__ void __launch_bounds__(kThreads, LPI <= 8 ? 4 : 3)      // D <= 32: 64 registers without spilling
msda_enc_fwd_kernel(const float* __restrict__ value, const int64_t* __restrict__ shapes,
                    const int64_t* __restrict__ lstart, const float* __restrict__ ow, const float* __restrict__ ref,
                    long long ref_bstride, int S, int M, int L, int Lq, float* __restrict__ out,
                    const MsdaTiling tiling) {
  constexpr int D = LPI * 4;
  constexpr int SLOTS = kThreads / LPI;
  constexpr int ITERS = kChunkQ / SLOTS;
  __shared__ int4 so[kChunkQ * kDescStride];
  __shared__ float4 sw[kChunkQ * kDescStride];
  __shared__ float s_aw[kChunkQ * kMaxLP];

  const int chunk = blockIdx.x, m = blockIdx.y, b = blockIdx.z;
  const int slot = threadIdx.x / LPI, li = threadIdx.x % LPI;
  const int MD = M * D, LP = L * 4, owc = M * LP * 3;
  const float* vimg = value + static_cast<size_t>(b) * S * MD + m * D + li * 4;
  const float* ow_b = ow + static_cast<size_t>(b) * Lq * owc;
  const float* ref_b = ref + b * ref_bstride;

  float4 acc[ITERS];
#pragma unroll
  for (int it = 0; it < ITERS; ++it) acc[it] = make_float4(0.f, 0.f, 0.f, 0.f);

  EncSamples es;
  enc_samples_init(es, tiling, chunk, Lq);
  enc_samples_fetch(es, ow_b, ref_b, m, L, owc, 0);
  enc_softmax(s_aw, es, ow_b, m, M, L, owc);
  __syncthreads();
  for (int l = 0; l < L; ++l) {
    const int H = static_cast<int>(shapes[2 * l]), W = static_cast<int>(shapes[2 * l + 1]);
    const float* vl = vimg + static_cast<size_t>(lstart[l]) * MD;
    if (l > 0) __syncthreads();
    enc_build_descriptors<false>(so, sw, es, s_aw, l, H, W, MD);
    if (l + 1 < L) enc_samples_fetch(es, ow_b, ref_b, m, L, owc, l + 1);
    __syncthreads();
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int j = it * SLOTS + slot;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int4 o = so[j * kDescStride + p];
        const float4 d = sw[j * kDescStride + p];          // (lh, lw, attention weight, -)
        if ((o.x & o.y & o.z & o.w) < 0) continue;         // sample outside the gate: the reference adds nothing
        const float hh = 1.f - d.x, hw = 1.f - d.y;
        const float w1 = hh * hw, w2 = hh * d.y, w3 = d.x * hw, w4 = d.x * d.y;
        const float4 v1 = o.x >= 0 ? ldg4(vl + o.x) : z;
        const float4 v2 = o.y >= 0 ? ldg4(vl + o.y) : z;
        const float4 v3 = o.z >= 0 ? ldg4(vl + o.z) : z;
        const float4 v4 = o.w >= 0 ? ldg4(vl + o.w) : z;
        acc[it].x = bilinear_acc(acc[it].x, d.z, w1, w2, w3, w4, v1.x, v2.x, v3.x, v4.x);
        acc[it].y = bilinear_acc(acc[it].y, d.z, w1, w2, w3, w4, v1.y, v2.y, v3.y, v4.y);
        acc[it].z = bilinear_acc(acc[it].z, d.z, w1, w2, w3, w4, v1.z, v2.z, v3.z, v4.z);
        acc[it].w = bilinear_acc(acc[it].w, d.z, w1, w2, w3, w4, v1.w, v2.w, v3.w, v4.w);
      }
    }
  }
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int q = query_of(tiling, chunk, it * SLOTS + slot, Lq);
    if (q >= 0)
      *reinterpret_cast<float4*>(out + ((static_cast<size_t>(b) * Lq + q) * M + m) * D + li * 4) = acc[it];
  }
}

template <int LPI, int OCC = 3>
__global__ void __launch_bounds__(kThreads, OCC)
msda_enc_bwd_kernel(const float* __restrict__ grad_out, const float* __restrict__ value,
                    const int64_t* __restrict__ shapes, const int64_t* __restrict__ lstart,
                    const float* __restrict__ ow, const float* __restrict__ ref, long long ref_bstride, int S, int M,
                    int L, int Lq, float* __restrict__ grad_value, float* __restrict__ grad_ow,
                    const MsdaTiling tiling) {
  constexpr int D = LPI * 4;
  constexpr int SLOTS = kThreads / LPI;
  constexpr int ITERS = kChunkQ / SLOTS;
  const unsigned FULL = 0xffffffffu;
  __shared__ int4 so[kChunkQ * kDescStride];
  __shared__ float4 sw[kChunkQ * kDescStride];
  __shared__ float s_aw[kChunkQ * kMaxLP];
  __shared__ float s_ga[kChunkQ * kMaxLP];      // d(loss)/d(attention weight) per item and sample

  const int chunk = blockIdx.x, m = blockIdx.y, b = blockIdx.z;
  const int slot = threadIdx.x / LPI, li = threadIdx.x % LPI;
  const int MD = M * D, LP = L * 4, owc = M * LP * 3;
  const size_t img_off = static_cast<size_t>(b) * S * MD + m * D + li * 4;
  const float* vimg = value + img_off;
  float* gvimg = grad_value + img_off;

  float4 g[ITERS];
  long long obase[ITERS];
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int q = query_of(tiling, chunk, it * SLOTS + slot, Lq);
    obase[it] = q >= 0 ? (static_cast<long long>(b) * Lq + q) * owc : -1;
    g[it] = q >= 0 ? ldg4(grad_out + ((static_cast<size_t>(b) * Lq + q) * M + m) * D + li * 4)
                   : make_float4(0.f, 0.f, 0.f, 0.f);
  }

  const float* ow_b = ow + static_cast<size_t>(b) * Lq * owc;
  const float* ref_b = ref + b * ref_bstride;
  EncSamples es;
  enc_samples_init(es, tiling, chunk, Lq);
  enc_samples_fetch(es, ow_b, ref_b, m, L, owc, 0);
  enc_softmax(s_aw, es, ow_b, m, M, L, owc);
  __syncthreads();
  for (int l = 0; l < L; ++l) {
    const int H = static_cast<int>(shapes[2 * l]), W = static_cast<int>(shapes[2 * l + 1]);
    const size_t loff = static_cast<size_t>(lstart[l]) * MD;
    const float* vl = vimg + loff;
    float* gvl = gvimg + loff;
    if (l > 0) __syncthreads();
    enc_build_descriptors<true>(so, sw, es, s_aw, l, H, W, MD);
    if (l + 1 < L) enc_samples_fetch(es, ow_b, ref_b, m, L, owc, l + 1);
    __syncthreads();
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int j = it * SLOTS + slot;
      float mine_x = 0.f, mine_y = 0.f, mine_a = 0.f;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int4 o = so[j * kDescStride + p];
        const float4 d = sw[j * kDescStride + p];
        const float lh = d.x, lw = d.y, wgt = d.z;
        const float hh = 1.f - lh, hw = 1.f - lw;
        const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
        const float4 v1 = o.x >= 0 ? ldg4(vl + o.x) : z;
        const float4 v2 = o.y >= 0 ? ldg4(vl + o.y) : z;
        const float4 v3 = o.z >= 0 ? ldg4(vl + o.z) : z;
        const float4 v4 = o.w >= 0 ? ldg4(vl + o.w) : z;
        const float4 tg = make_float4(g[it].x * wgt, g[it].y * wgt, g[it].z * wgt, g[it].w * wgt);
        if (o.x >= 0) red_add_v4(gvl + o.x, w1 * tg.x, w1 * tg.y, w1 * tg.z, w1 * tg.w);
        if (o.y >= 0) red_add_v4(gvl + o.y, w2 * tg.x, w2 * tg.y, w2 * tg.z, w2 * tg.w);
        if (o.z >= 0) red_add_v4(gvl + o.z, w3 * tg.x, w3 * tg.y, w3 * tg.z, w3 * tg.w);
        if (o.w >= 0) red_add_v4(gvl + o.w, w4 * tg.x, w4 * tg.y, w4 * tg.z, w4 * tg.w);
        float gh = 0.f, gw = 0.f, ga = 0.f;
#define MPF_ACC(comp)                                                                         \
  {                                                                                           \
    const float ghw = -hw * v1.comp - lw * v2.comp + hw * v3.comp + lw * v4.comp;             \
    const float gww = -hh * v1.comp + hh * v2.comp - lh * v3.comp + lh * v4.comp;             \
    const float val = w1 * v1.comp + w2 * v2.comp + w3 * v3.comp + w4 * v4.comp;              \
    gh += ghw * tg.comp;                                                                      \
    gw += gww * tg.comp;                                                                      \
    ga += val * g[it].comp;                                                                   \
  }
        MPF_ACC(x) MPF_ACC(y) MPF_ACC(z) MPF_ACC(w)
#undef MPF_ACC
#pragma unroll
        for (int off = LPI / 2; off >= 1; off >>= 1) {
          gh += __shfl_xor_sync(FULL, gh, off);
          gw += __shfl_xor_sync(FULL, gw, off);
          ga += __shfl_xor_sync(FULL, ga, off);
        }
        // d/d(off) = d/d(loc) / (W, H) = (W*gw)/W ... : the W, H factors of ref cuh:162-163 cancel
        if (li == p) { mine_x = gw; mine_y = gh; mine_a = ga; }
      }
      if (obase[it] >= 0 && li < 4) {
        *reinterpret_cast<float2*>(grad_ow + obase[it] + (m * LP + l * 4 + li) * 2) = make_float2(mine_x, mine_y);
        s_ga[j * kMaxLP + l * 4 + li] = mine_a;
      }
    }
  }
  __syncthreads();
  // softmax backward per item: d(logit_i) = aw_i * (ga_i - sum_j aw_j ga_j)
  if (threadIdx.x < kChunkQ) {
    const int j = threadIdx.x;
    const int q = query_of(tiling, chunk, j, Lq);
    if (q >= 0) {
      float dot = 0.f;
      for (int i = 0; i < LP; ++i) dot += s_aw[j * kMaxLP + i] * s_ga[j * kMaxLP + i];
      float* gl = grad_ow + (static_cast<long long>(b) * Lq + q) * owc + M * LP * 2 + m * LP;
      for (int i = 0; i < LP; ++i) gl[i] = s_aw[j * kMaxLP + i] * (s_ga[j * kMaxLP + i] - dot);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Generic kernels: one thread per (b, q, m, c); any D / L / P; float or double.
// ------------------------------------------------------------------------------------------------
template <typename T>
struct Bilinear {
  bool inside, c1, c2, c3, c4;
  long long o1, o2, o3, o4;
  T w1, w2, w3, w4, lh, lw, hh, hw;
};

template <typename T>
__device__ __forceinline__ Bilinear<T> bilinear_setup(T loc_x, T loc_y, int H, int W, int MD) {
  Bilinear<T> r;
  const T h_im = loc_y * H - T(0.5);
  const T w_im = loc_x * W - T(0.5);
  r.inside = (h_im > T(-1)) && (w_im > T(-1)) && (h_im < T(H)) && (w_im < T(W));
  const T hf = r.inside ? floor(h_im) : T(0), wf = r.inside ? floor(w_im) : T(0);
  const int h_low = static_cast<int>(hf), w_low = static_cast<int>(wf);
  r.lh = h_im - hf;
  r.lw = w_im - wf;
  r.hh = T(1) - r.lh;
  r.hw = T(1) - r.lw;
  const bool top = r.inside && h_low >= 0, bot = r.inside && (h_low + 1 <= H - 1);
  const bool lft = w_low >= 0, rgt = (w_low + 1 <= W - 1);
  r.c1 = top && lft; r.c2 = top && rgt; r.c3 = bot && lft; r.c4 = bot && rgt;
  const long long rs = static_cast<long long>(W) * MD;
  r.o1 = h_low * rs + static_cast<long long>(w_low) * MD;
  r.o2 = r.o1 + MD;
  r.o3 = r.o1 + rs;
  r.o4 = r.o3 + MD;
  r.w1 = r.hh * r.hw; r.w2 = r.hh * r.lw; r.w3 = r.lh * r.hw; r.w4 = r.lh * r.lw;
  return r;
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
msda_fwd_generic_kernel(const T* __restrict__ value, const int64_t* __restrict__ shapes,
                        const int64_t* __restrict__ lstart, const T* __restrict__ loc,
                        const T* __restrict__ aw, long long n, int S, int M, int D, int L, int Lq,
                        int P, T* __restrict__ out) {
  const int MD = M * D;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < n;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % D);
    const long long item = idx / D;  // (b*Lq + q)*M + m
    const int m = static_cast<int>(item % M);
    const long long b = item / M / Lq;
    const T* vimg = value + b * S * MD + m * D + c;
    T acc = 0;
    for (int l = 0; l < L; ++l) {
      const int H = static_cast<int>(shapes[2 * l]), W = static_cast<int>(shapes[2 * l + 1]);
      const T* vl = vimg + lstart[l] * MD;
      for (int p = 0; p < P; ++p) {
        const long long s = item * L * P + l * P + p;
        const Bilinear<T> bi = bilinear_setup<T>(loc[2 * s], loc[2 * s + 1], H, W, MD);
        if (!bi.inside) continue;
        const T v1 = bi.c1 ? vl[bi.o1] : T(0), v2 = bi.c2 ? vl[bi.o2] : T(0);
        const T v3 = bi.c3 ? vl[bi.o3] : T(0), v4 = bi.c4 ? vl[bi.o4] : T(0);
        acc += aw[s] * (bi.w1 * v1 + bi.w2 * v2 + bi.w3 * v3 + bi.w4 * v4);
      }
    }
    out[idx] = acc;
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
msda_bwd_generic_kernel(const T* __restrict__ grad_out, const T* __restrict__ value,
                        const int64_t* __restrict__ shapes, const int64_t* __restrict__ lstart,
                        const T* __restrict__ loc, const T* __restrict__ aw, long long n, int S, int M,
                        int D, int L, int Lq, int P, T* __restrict__ grad_value,
                        T* __restrict__ grad_loc, T* __restrict__ grad_aw) {
  const int MD = M * D;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < n;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % D);
    const long long item = idx / D;
    const int m = static_cast<int>(item % M);
    const long long b = item / M / Lq;
    const long long img = b * S * MD + m * D + c;
    const T g = grad_out[idx];
    for (int l = 0; l < L; ++l) {
      const int H = static_cast<int>(shapes[2 * l]), W = static_cast<int>(shapes[2 * l + 1]);
      const long long loff = img + lstart[l] * MD;
      const T* vl = value + loff;
      T* gvl = grad_value + loff;
      for (int p = 0; p < P; ++p) {
        const long long s = item * L * P + l * P + p;
        const Bilinear<T> bi = bilinear_setup<T>(loc[2 * s], loc[2 * s + 1], H, W, MD);
        if (!bi.inside) continue;
        const T wgt = aw[s];
        const T tg = g * wgt;
        const T v1 = bi.c1 ? vl[bi.o1] : T(0), v2 = bi.c2 ? vl[bi.o2] : T(0);
        const T v3 = bi.c3 ? vl[bi.o3] : T(0), v4 = bi.c4 ? vl[bi.o4] : T(0);
        if (bi.c1) atomicAdd(gvl + bi.o1, bi.w1 * tg);
        if (bi.c2) atomicAdd(gvl + bi.o2, bi.w2 * tg);
        if (bi.c3) atomicAdd(gvl + bi.o3, bi.w3 * tg);
        if (bi.c4) atomicAdd(gvl + bi.o4, bi.w4 * tg);
        const T ghw = -bi.hw * v1 - bi.lw * v2 + bi.hw * v3 + bi.lw * v4;
        const T gww = -bi.hh * v1 + bi.hh * v2 - bi.lh * v3 + bi.lh * v4;
        const T val = bi.w1 * v1 + bi.w2 * v2 + bi.w3 * v3 + bi.w4 * v4;
        atomicAdd(grad_aw + s, g * val);
        atomicAdd(grad_loc + 2 * s, W * gww * tg);
        atomicAdd(grad_loc + 2 * s + 1, H * ghw * tg);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
static int check_dims(int batch, int spatial_size, int num_heads, int channels, int num_levels,
                      int num_query, int num_point) {
  MPF_REQUIRE(batch > 0 && spatial_size > 0 && num_heads > 0 && channels > 0 && num_levels > 0 &&
                  num_query > 0 && num_point > 0,
              "msda: all dimensions must be positive (batch=%d S=%d M=%d D=%d L=%d Lq=%d P=%d)",
              batch, spatial_size, num_heads, channels, num_levels, num_query, num_point);
  MPF_REQUIRE(static_cast<long long>(spatial_size) * num_heads * channels < (1ll << 31),
              "msda: one image of value must have < 2^31 elements");
  return MPF_OK;
}

static bool vec_path_ok(int D, int L, int P, int M, int B, int* lpi) {
  if (P != 4 || D % 4 != 0) return false;
  const int l = D / 4;
  if (l != 4 && l != 8 && l != 16) return false;
  if (L > kMaxTiledLevels) return false;
  if (M > 65535 || B > 65535) return false;
  *lpi = l;
  return true;
}

template <typename T>
static int launch_fwd_generic(const T* value, const int64_t* shapes, const int64_t* lstart,
                              const T* loc, const T* aw, int B, int S, int M, int D, int L, int Lq,
                              int P, T* out, cudaStream_t st) {
  const long long n = static_cast<long long>(B) * Lq * M * D;
  const long long blocks = (n + kThreads - 1) / kThreads;
  const int grid = static_cast<int>(blocks < (1 << 20) ? blocks : (1 << 20));
  msda_fwd_generic_kernel<T><<<grid, kThreads, 0, st>>>(value, shapes, lstart, loc, aw, n, S, M, D,
                                                        L, Lq, P, out);
  count_launch();
  return finish_launch("msda_fwd_generic");
}

template <typename T>
static int launch_bwd_generic(const T* grad_out, const T* value, const int64_t* shapes,
                              const int64_t* lstart, const T* loc, const T* aw, int B, int S, int M,
                              int D, int L, int Lq, int P, T* gv, T* gl, T* ga, cudaStream_t st) {
  const size_t nloc = static_cast<size_t>(B) * Lq * M * L * P;
  MPF_CUDA_OK(cudaMemsetAsync(gl, 0, nloc * 2 * sizeof(T), st));
  MPF_CUDA_OK(cudaMemsetAsync(ga, 0, nloc * sizeof(T), st));
  const long long n = static_cast<long long>(B) * Lq * M * D;
  const long long blocks = (n + kThreads - 1) / kThreads;
  const int grid = static_cast<int>(blocks < (1 << 20) ? blocks : (1 << 20));
  msda_bwd_generic_kernel<T><<<grid, kThreads, 0, st>>>(grad_out, value, shapes, lstart, loc, aw, n,
                                                        S, M, D, L, Lq, P, gv, gl, ga);
  count_launch();
  return finish_launch("msda_bwd_generic");
}

int msda_forward_f32(const float* value, const int64_t* shapes, const int64_t* lstart,
                     const float* loc, const float* aw, int B, int S, int M, int D, int L, int Lq,
                     int P, float* out, const int64_t* shapes_host, cudaStream_t st) {
  int lpi = 0;
  const bool vec = vec_path_ok(D, L, P, M, B, &lpi) && aligned16(value) && aligned16(loc) &&
                   aligned16(aw) && aligned16(out);
  if (!vec) return launch_fwd_generic<float>(value, shapes, lstart, loc, aw, B, S, M, D, L, Lq, P, out, st);
  const MsdaTiling t = make_tiling(shapes_host, L, S, Lq);
  dim3 grid(t.num_chunks, M, B);
  switch (lpi) {
    case 4: msda_fwd_vec_kernel<4><<<grid, kThreads, 0, st>>>(value, shapes, lstart, loc, aw, S, M, L, Lq, out, t); break;
    case 8: msda_fwd_vec_kernel<8><<<grid, kThreads, 0, st>>>(value, shapes, lstart, loc, aw, S, M, L, Lq, out, t); break;
    default: msda_fwd_vec_kernel<16><<<grid, kThreads, 0, st>>>(value, shapes, lstart, loc, aw, S, M, L, Lq, out, t); break;
  }
  count_launch();
  return finish_launch("msda_fwd_vec");
}

int msda_backward_f32(const float* grad_out, const float* value, const int64_t* shapes,
                      const int64_t* lstart, const float* loc, const float* aw, int B, int S, int M,
                      int D, int L, int Lq, int P, float* gv, float* gl, float* ga,
                      const int64_t* shapes_host, cudaStream_t st) {
  MPF_CUDA_OK(cudaMemsetAsync(gv, 0, static_cast<size_t>(B) * S * M * D * sizeof(float), st));
  int lpi = 0;
  const bool vec = vec_path_ok(D, L, P, M, B, &lpi) && aligned16(value) && aligned16(loc) &&
                   aligned16(aw) && aligned16(grad_out) && aligned16(gv) && aligned16(gl) &&
                   aligned16(ga);
  if (!vec)
    return launch_bwd_generic<float>(grad_out, value, shapes, lstart, loc, aw, B, S, M, D, L, Lq, P,
                                     gv, gl, ga, st);
  const MsdaTiling t = make_tiling(shapes_host, L, S, Lq);
  dim3 grid(t.num_chunks, M, B);
  switch (lpi) {
    case 4: msda_bwd_vec_kernel<4><<<grid, kThreads, 0, st>>>(grad_out, value, shapes, lstart, loc, aw, S, M, L, Lq, gv, gl, ga, t); break;
    case 8: msda_bwd_vec_kernel<8><<<grid, kThreads, 0, st>>>(grad_out, value, shapes, lstart, loc, aw, S, M, L, Lq, gv, gl, ga, t); break;
    default: msda_bwd_vec_kernel<16><<<grid, kThreads, 0, st>>>(grad_out, value, shapes, lstart, loc, aw, S, M, L, Lq, gv, gl, ga, t); break;
  }
  count_launch();
  return finish_launch("msda_bwd_vec");
}

// msda_staged.cu
bool msda_staged_ok(const MsdaTiling& t, int D, int L, int P, int M, int B);
int msda_enc_forward_staged(const float* value, const float* ow, const float* ref, long long ref_bstride, int B, int S,
                            int M, int L, int Lq, float* out, const MsdaTiling& t, cudaStream_t st);
int msda_enc_backward_staged(const float* grad_out, const float* value, const float* ow, const float* ref,
                             long long ref_bstride, int B, int S, int M, int L, int Lq, float* gv, float* gow,
                             const MsdaTiling& t, cudaStream_t st);

static bool enc_path_ok(int D, int L, int P, int M, int B, int* lpi) {
  return vec_path_ok(D, L, P, M, B, lpi) && L * P <= kMaxLP;
}

int msda_enc_forward_f32(const float* value, const int64_t* shapes, const int64_t* lstart, const float* ow,
                         const float* ref, long long ref_bstride, int B, int S, int M, int D, int L, int Lq, int P,
                         float* out, const int64_t* shapes_host, cudaStream_t st) {
  int lpi = 0;
  if (!(enc_path_ok(D, L, P, M, B, &lpi) && aligned16(value) && aligned16(out) && aligned16(ow) && aligned16(ref))) {
    set_error("msda_enc_forward: unsupported geometry for the fused path (D=%d L=%d P=%d)", D, L, P);
    return MPF_ERR_UNSUPPORTED;
  }
  const MsdaTiling t = make_tiling(shapes_host, L, S, Lq);
  if (msda_staged_ok(t, D, L, P, M, B))
    return msda_enc_forward_staged(value, ow, ref, ref_bstride, B, S, M, L, Lq, out, t, st);
  dim3 grid(t.num_chunks, M, B);
  switch (lpi) {
    case 4: msda_enc_fwd_kernel<4><<<grid, kThreads, 0, st>>>(value, shapes, lstart, ow, ref, ref_bstride, S, M, L, Lq, out, t); break;
    case 8: msda_enc_fwd_kernel<8><<<grid, kThreads, 0, st>>>(value, shapes, lstart, ow, ref, ref_bstride, S, M, L, Lq, out, t); break;
    default: msda_enc_fwd_kernel<16><<<grid, kThreads, 0, st>>>(value, shapes, lstart, ow, ref, ref_bstride, S, M, L, Lq, out, t); break;
  }
  count_launch();
  return finish_launch("msda_enc_fwd");
}

int msda_enc_backward_f32(const float* grad_out, const float* value, const int64_t* shapes, const int64_t* lstart,
                          const float* ow, const float* ref, long long ref_bstride, int B, int S, int M, int D, int L,
                          int Lq, int P, float* gv, float* gow, const int64_t* shapes_host, cudaStream_t st) {
  int lpi = 0;
  if (!(enc_path_ok(D, L, P, M, B, &lpi) && aligned16(value) && aligned16(grad_out) && aligned16(ow) &&
        aligned16(ref) && aligned16(gv) && aligned16(gow))) {
    set_error("msda_enc_backward: unsupported geometry for the fused path (D=%d L=%d P=%d)", D, L, P);
    return MPF_ERR_UNSUPPORTED;
  }
  MPF_CUDA_OK(cudaMemsetAsync(gv, 0, static_cast<size_t>(B) * S * M * D * sizeof(float), st));
  const MsdaTiling t = make_tiling(shapes_host, L, S, Lq);
  if (msda_staged_ok(t, D, L, P, M, B))
    return msda_enc_backward_staged(grad_out, value, ow, ref, ref_bstride, B, S, M, L, Lq, gv, gow, t, st);
  dim3 grid(t.num_chunks, M, B);
  switch (lpi) {
    case 4: msda_enc_bwd_kernel<4><<<grid, kThreads, 0, st>>>(grad_out, value, shapes, lstart, ow, ref, ref_bstride, S, M, L, Lq, gv, gow, t); break;
    case 8:
      if (getenv("MPF_MSDA_BWD_OCC4") != nullptr)      // tuning aid: 64 registers (spilling ~300 B) at 4 CTAs per SM
        msda_enc_bwd_kernel<8, 4><<<grid, kThreads, 0, st>>>(grad_out, value, shapes, lstart, ow, ref, ref_bstride, S, M, L, Lq, gv, gow, t);
      else
        msda_enc_bwd_kernel<8><<<grid, kThreads, 0, st>>>(grad_out, value, shapes, lstart, ow, ref, ref_bstride, S, M, L, Lq, gv, gow, t);
      break;
    default: msda_enc_bwd_kernel<16><<<grid, kThreads, 0, st>>>(grad_out, value, shapes, lstart, ow, ref, ref_bstride, S, M, L, Lq, gv, gow, t); break;
  }
  count_launch();
  return finish_launch("msda_enc_bwd");
}

}  // namespace mpf

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int mpf_msda_forward_f32_ex(const float* value, const int64_t* spatial_shapes,
                            const int64_t* level_start_index, const float* sampling_loc,
                            const float* attn_weight, int batch, int spatial_size, int num_heads,
                            int channels, int num_levels, int num_query, int num_point, float* out,
                            const int64_t* spatial_shapes_host, void* stream) {
  mpf::clear_error();
  int rc = mpf::check_dims(batch, spatial_size, num_heads, channels, num_levels, num_query, num_point);
  if (rc) return rc;
  MPF_REQUIRE(value && spatial_shapes && level_start_index && sampling_loc && attn_weight && out,
              "msda_forward: null pointer argument");
  return mpf::msda_forward_f32(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                               batch, spatial_size, num_heads, channels, num_levels, num_query,
                               num_point, out, spatial_shapes_host,
                               static_cast<cudaStream_t>(stream));
}

int mpf_msda_forward_f32(const float* value, const int64_t* spatial_shapes,
                         const int64_t* level_start_index, const float* sampling_loc,
                         const float* attn_weight, int batch, int spatial_size, int num_heads,
                         int channels, int num_levels, int num_query, int num_point, float* out,
                         void* stream) {
  return mpf_msda_forward_f32_ex(value, spatial_shapes, level_start_index, sampling_loc,
                                 attn_weight, batch, spatial_size, num_heads, channels, num_levels,
                                 num_query, num_point, out, nullptr, stream);
}

int mpf_msda_forward_f64(const double* value, const int64_t* spatial_shapes,
                         const int64_t* level_start_index, const double* sampling_loc,
                         const double* attn_weight, int batch, int spatial_size, int num_heads,
                         int channels, int num_levels, int num_query, int num_point, double* out,
                         void* stream) {
  mpf::clear_error();
  int rc = mpf::check_dims(batch, spatial_size, num_heads, channels, num_levels, num_query, num_point);
  if (rc) return rc;
  MPF_REQUIRE(value && spatial_shapes && level_start_index && sampling_loc && attn_weight && out,
              "msda_forward: null pointer argument");
  return mpf::launch_fwd_generic<double>(value, spatial_shapes, level_start_index, sampling_loc,
                                         attn_weight, batch, spatial_size, num_heads, channels,
                                         num_levels, num_query, num_point, out,
                                         static_cast<cudaStream_t>(stream));
}

int mpf_msda_backward_f32_ex(const float* grad_out, const float* value,
                             const int64_t* spatial_shapes, const int64_t* level_start_index,
                             const float* sampling_loc, const float* attn_weight, int batch,
                             int spatial_size, int num_heads, int channels, int num_levels,
                             int num_query, int num_point, float* grad_value,
                             float* grad_sampling_loc, float* grad_attn_weight,
                             const int64_t* spatial_shapes_host, void* stream) {
  mpf::clear_error();
  int rc = mpf::check_dims(batch, spatial_size, num_heads, channels, num_levels, num_query, num_point);
  if (rc) return rc;
  MPF_REQUIRE(grad_out && value && spatial_shapes && level_start_index && sampling_loc &&
                  attn_weight && grad_value && grad_sampling_loc && grad_attn_weight,
              "msda_backward: null pointer argument");
  return mpf::msda_backward_f32(grad_out, value, spatial_shapes, level_start_index, sampling_loc,
                                attn_weight, batch, spatial_size, num_heads, channels, num_levels,
                                num_query, num_point, grad_value, grad_sampling_loc,
                                grad_attn_weight, spatial_shapes_host,
                                static_cast<cudaStream_t>(stream));
}

int mpf_msda_backward_f32(const float* grad_out, const float* value, const int64_t* spatial_shapes,
                          const int64_t* level_start_index, const float* sampling_loc,
                          const float* attn_weight, int batch, int spatial_size, int num_heads,
                          int channels, int num_levels, int num_query, int num_point,
                          float* grad_value, float* grad_sampling_loc, float* grad_attn_weight,
                          void* stream) {
  return mpf_msda_backward_f32_ex(grad_out, value, spatial_shapes, level_start_index, sampling_loc,
                                  attn_weight, batch, spatial_size, num_heads, channels, num_levels,
                                  num_query, num_point, grad_value, grad_sampling_loc,
                                  grad_attn_weight, nullptr, stream);
}

int mpf_msda_backward_f64(const double* grad_out, const double* value,
                          const int64_t* spatial_shapes, const int64_t* level_start_index,
                          const double* sampling_loc, const double* attn_weight, int batch,
                          int spatial_size, int num_heads, int channels, int num_levels,
                          int num_query, int num_point, double* grad_value,
                          double* grad_sampling_loc, double* grad_attn_weight, void* stream) {
  mpf::clear_error();
  int rc = mpf::check_dims(batch, spatial_size, num_heads, channels, num_levels, num_query, num_point);
  if (rc) return rc;
  MPF_REQUIRE(grad_out && value && spatial_shapes && level_start_index && sampling_loc &&
                  attn_weight && grad_value && grad_sampling_loc && grad_attn_weight,
              "msda_backward: null pointer argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MPF_CUDA_OK(cudaMemsetAsync(
      grad_value, 0, static_cast<size_t>(batch) * spatial_size * num_heads * channels * sizeof(double), st));
  return mpf::launch_bwd_generic<double>(grad_out, value, spatial_shapes, level_start_index,
                                         sampling_loc, attn_weight, batch, spatial_size, num_heads,
                                         channels, num_levels, num_query, num_point, grad_value,
                                         grad_sampling_loc, grad_attn_weight, st);
}

int mpf_msda_enc_forward_f32(const float* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                             const float* offsets_logits, const float* reference_points, long long ref_batch_stride,
                             int batch, int spatial_size, int num_heads, int channels, int num_levels, int num_query,
                             int num_point, float* out, const int64_t* spatial_shapes_host, void* stream) {
  mpf::clear_error();
  int rc = mpf::check_dims(batch, spatial_size, num_heads, channels, num_levels, num_query, num_point);
  if (rc) return rc;
  MPF_REQUIRE(value && spatial_shapes && level_start_index && offsets_logits && reference_points && out,
              "msda_enc_forward: null pointer argument");
  return mpf::msda_enc_forward_f32(value, spatial_shapes, level_start_index, offsets_logits, reference_points,
                                   ref_batch_stride, batch, spatial_size, num_heads, channels, num_levels, num_query,
                                   num_point, out, spatial_shapes_host, static_cast<cudaStream_t>(stream));
}

int mpf_msda_enc_backward_f32(const float* grad_out, const float* value, const int64_t* spatial_shapes,
                              const int64_t* level_start_index, const float* offsets_logits,
                              const float* reference_points, long long ref_batch_stride, int batch, int spatial_size,
                              int num_heads, int channels, int num_levels, int num_query, int num_point,
                              float* grad_value, float* grad_offsets_logits, const int64_t* spatial_shapes_host,
                              void* stream) {
  mpf::clear_error();
  int rc = mpf::check_dims(batch, spatial_size, num_heads, channels, num_levels, num_query, num_point);
  if (rc) return rc;
  MPF_REQUIRE(grad_out && value && spatial_shapes && level_start_index && offsets_logits && reference_points &&
                  grad_value && grad_offsets_logits,
              "msda_enc_backward: null pointer argument");
  return mpf::msda_enc_backward_f32(grad_out, value, spatial_shapes, level_start_index, offsets_logits,
                                    reference_points, ref_batch_stride, batch, spatial_size, num_heads, channels,
                                    num_levels, num_query, num_point, grad_value, grad_offsets_logits,
                                    spatial_shapes_host, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

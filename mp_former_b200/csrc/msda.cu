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
//   * Sampling locations / attention weights of an item are loaded once as float4 by the item's
//     own lanes and handed round with warp shuffles (no shared memory, no __syncthreads).
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

// ------------------------------------------------------------------------------------------------
// Forward, vectorised: D = 4*LPI channels, P = 4 points, L <= LPI/2 levels.
// ------------------------------------------------------------------------------------------------
template <int LPI>
__global__ void __launch_bounds__(kThreads, 3)
msda_fwd_vec_kernel(const float* __restrict__ value, const int64_t* __restrict__ shapes,
                    const int64_t* __restrict__ lstart, const float* __restrict__ loc,
                    const float* __restrict__ aw, int S, int M, int L, int Lq,
                    float* __restrict__ out, const MsdaTiling tiling) {
  constexpr int D = LPI * 4;
  constexpr int P = 4;
  constexpr int SLOTS = kThreads / LPI;
  constexpr int MAXL = (LPI / 2) < 8 ? (LPI / 2) : 8;
  const unsigned FULL = 0xffffffffu;

  const int chunk = blockIdx.x, m = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x;
  const int slot = tid / LPI;
  const int li = tid % LPI;                       // which float4 of the D channels
  const int lane = tid & 31;
  const int slot_lane0 = lane - li;               // first lane of this item's lane group
  const int LP = L * P;
  const int n_loc4 = LP / 2;                      // float4s of (x, y) pairs per item
  const int n_aw4 = LP / 4;
  const int MD = M * D;

  int Hs[MAXL], Ws[MAXL], St[MAXL];
#pragma unroll
  for (int l = 0; l < MAXL; ++l) {
    if (l < L) {
      Hs[l] = static_cast<int>(shapes[2 * l]);
      Ws[l] = static_cast<int>(shapes[2 * l + 1]);
      St[l] = static_cast<int>(lstart[l]);
    } else {
      Hs[l] = Ws[l] = St[l] = 0;
    }
  }
  const float* vimg = value + static_cast<size_t>(b) * S * MD + m * D + li * 4;

  for (int it = 0; it < kChunkQ / SLOTS; ++it) {
    const int j = it * SLOTS + slot;
    int q = query_of(tiling, chunk, j, Lq);
    const bool active = q >= 0;
    if (__all_sync(FULL, !active)) continue;       // warp-uniform skip keeps shuffles convergent
    q = active ? q : 0;
    const size_t item = (static_cast<size_t>(b) * Lq + q) * M + m;
    float4 locv = make_float4(0.f, 0.f, 0.f, 0.f), awv = locv;
    if (li < n_loc4) locv = ldg4(loc + item * LP * 2 + li * 4);
    if (li < n_aw4) awv = ldg4(aw + item * LP + li * 4);

    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int l = 0; l < MAXL; ++l) {
      if (l < L) {
        const int H = Hs[l], W = Ws[l];
        const float* vl = vimg + static_cast<size_t>(St[l]) * MD;
        const int row_stride = W * MD;
#pragma unroll
        for (int p = 0; p < P; ++p) {
          const int src = slot_lane0 + 2 * l + (p >> 1);
          const float lx = __shfl_sync(FULL, (p & 1) ? locv.z : locv.x, src);
          const float ly = __shfl_sync(FULL, (p & 1) ? locv.w : locv.y, src);
          const float wgt = __shfl_sync(
              FULL, p == 0 ? awv.x : (p == 1 ? awv.y : (p == 2 ? awv.z : awv.w)), slot_lane0 + l);
          const float h_im = ly * H - 0.5f;
          const float w_im = lx * W - 0.5f;
          const bool inside = (h_im > -1.f) && (w_im > -1.f) && (h_im < H) && (w_im < W);
          const float hf = inside ? floorf(h_im) : 0.f, wf = inside ? floorf(w_im) : 0.f;
          const int h_low = static_cast<int>(hf), w_low = static_cast<int>(wf);
          const float lh = h_im - hf, lw = w_im - wf;
          const float hh = 1.f - lh, hw = 1.f - lw;
          const bool top = inside && h_low >= 0, bot = inside && (h_low + 1 <= H - 1);
          const bool lft = w_low >= 0, rgt = (w_low + 1 <= W - 1);
          const int o1 = h_low * row_stride + w_low * MD;
          const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
          const float4 v1 = (top && lft) ? ldg4(vl + o1) : z;
          const float4 v2 = (top && rgt) ? ldg4(vl + o1 + MD) : z;
          const float4 v3 = (bot && lft) ? ldg4(vl + o1 + row_stride) : z;
          const float4 v4 = (bot && rgt) ? ldg4(vl + o1 + row_stride + MD) : z;
          const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
          acc.x += wgt * (w1 * v1.x + w2 * v2.x + w3 * v3.x + w4 * v4.x);
          acc.y += wgt * (w1 * v1.y + w2 * v2.y + w3 * v3.y + w4 * v4.y);
          acc.z += wgt * (w1 * v1.z + w2 * v2.z + w3 * v3.z + w4 * v4.z);
          acc.w += wgt * (w1 * v1.w + w2 * v2.w + w3 * v3.w + w4 * v4.w);
        }
      }
    }
    if (active) *reinterpret_cast<float4*>(out + item * D + li * 4) = acc;
  }
}

// ------------------------------------------------------------------------------------------------
// Backward, vectorised (same mapping as the forward).
// ------------------------------------------------------------------------------------------------
template <int LPI>
__global__ void __launch_bounds__(kThreads, 3)
msda_bwd_vec_kernel(const float* __restrict__ grad_out, const float* __restrict__ value,
                    const int64_t* __restrict__ shapes, const int64_t* __restrict__ lstart,
                    const float* __restrict__ loc, const float* __restrict__ aw, int S, int M, int L,
                    int Lq, float* __restrict__ grad_value, float* __restrict__ grad_loc,
                    float* __restrict__ grad_aw, const MsdaTiling tiling) {
  constexpr int D = LPI * 4;
  constexpr int P = 4;
  constexpr int SLOTS = kThreads / LPI;
  constexpr int MAXL = (LPI / 2) < 8 ? (LPI / 2) : 8;
  const unsigned FULL = 0xffffffffu;

  const int chunk = blockIdx.x, m = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x;
  const int slot = tid / LPI;
  const int li = tid % LPI;
  const int lane = tid & 31;
  const int slot_lane0 = lane - li;
  const int LP = L * P;
  const int n_loc4 = LP / 2;
  const int n_aw4 = LP / 4;
  const int MD = M * D;

  int Hs[MAXL], Ws[MAXL], St[MAXL];
#pragma unroll
  for (int l = 0; l < MAXL; ++l) {
    if (l < L) {
      Hs[l] = static_cast<int>(shapes[2 * l]);
      Ws[l] = static_cast<int>(shapes[2 * l + 1]);
      St[l] = static_cast<int>(lstart[l]);
    } else {
      Hs[l] = Ws[l] = St[l] = 0;
    }
  }
  const size_t img_off = static_cast<size_t>(b) * S * MD + m * D + li * 4;
  const float* vimg = value + img_off;
  float* gvimg = grad_value + img_off;

  for (int it = 0; it < kChunkQ / SLOTS; ++it) {
    const int j = it * SLOTS + slot;
    int q = query_of(tiling, chunk, j, Lq);
    const bool active = q >= 0;
    if (__all_sync(FULL, !active)) continue;
    q = active ? q : 0;
    const size_t item = (static_cast<size_t>(b) * Lq + q) * M + m;
    float4 locv = make_float4(0.f, 0.f, 0.f, 0.f), awv = locv;
    if (li < n_loc4) locv = ldg4(loc + item * LP * 2 + li * 4);
    if (li < n_aw4) awv = ldg4(aw + item * LP + li * 4);
    float4 g = ldg4(grad_out + item * D + li * 4);
    if (!active) g = make_float4(0.f, 0.f, 0.f, 0.f);

    float4 gloc_out = make_float4(0.f, 0.f, 0.f, 0.f);  // this lane's float4 of grad_loc
    float4 gaw_out = make_float4(0.f, 0.f, 0.f, 0.f);   // this lane's float4 of grad_aw
#pragma unroll
    for (int l = 0; l < MAXL; ++l) {
      if (l < L) {
        const int H = Hs[l], W = Ws[l];
        const size_t loff = static_cast<size_t>(St[l]) * MD;
        const float* vl = vimg + loff;
        float* gvl = gvimg + loff;
        const int row_stride = W * MD;
#pragma unroll
        for (int p = 0; p < P; ++p) {
          const int src = slot_lane0 + 2 * l + (p >> 1);
          const float lx = __shfl_sync(FULL, (p & 1) ? locv.z : locv.x, src);
          const float ly = __shfl_sync(FULL, (p & 1) ? locv.w : locv.y, src);
          const float wgt = __shfl_sync(
              FULL, p == 0 ? awv.x : (p == 1 ? awv.y : (p == 2 ? awv.z : awv.w)), slot_lane0 + l);
          const float h_im = ly * H - 0.5f;
          const float w_im = lx * W - 0.5f;
          const bool inside = (h_im > -1.f) && (w_im > -1.f) && (h_im < H) && (w_im < W);
          const float hf = inside ? floorf(h_im) : 0.f, wf = inside ? floorf(w_im) : 0.f;
          const int h_low = static_cast<int>(hf), w_low = static_cast<int>(wf);
          const float lh = h_im - hf, lw = w_im - wf;
          const float hh = 1.f - lh, hw = 1.f - lw;
          const bool top = inside && h_low >= 0, bot = inside && (h_low + 1 <= H - 1);
          const bool lft = w_low >= 0, rgt = (w_low + 1 <= W - 1);
          const int o1 = h_low * row_stride + w_low * MD;
          const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
          const bool c1 = top && lft, c2 = top && rgt, c3 = bot && lft, c4 = bot && rgt;
          const float4 v1 = c1 ? ldg4(vl + o1) : z;
          const float4 v2 = c2 ? ldg4(vl + o1 + MD) : z;
          const float4 v3 = c3 ? ldg4(vl + o1 + row_stride) : z;
          const float4 v4 = c4 ? ldg4(vl + o1 + row_stride + MD) : z;
          const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
          // top_grad_value = grad_out * attn_weight  (ref cuh:116)
          const float4 tg = make_float4(g.x * wgt, g.y * wgt, g.z * wgt, g.w * wgt);
          if (active) {
            if (c1) red_add_v4(gvl + o1, w1 * tg.x, w1 * tg.y, w1 * tg.z, w1 * tg.w);
            if (c2) red_add_v4(gvl + o1 + MD, w2 * tg.x, w2 * tg.y, w2 * tg.z, w2 * tg.w);
            if (c3) red_add_v4(gvl + o1 + row_stride, w3 * tg.x, w3 * tg.y, w3 * tg.z, w3 * tg.w);
            if (c4)
              red_add_v4(gvl + o1 + row_stride + MD, w4 * tg.x, w4 * tg.y, w4 * tg.z, w4 * tg.w);
          }
          // d/d(h_im), d/d(w_im) of the bilinear value, per channel (ref cuh:119-157)
          float gh = 0.f, gw = 0.f, ga = 0.f;
#define MPF_ACC(comp)                                                                         \
  {                                                                                           \
    const float ghw = -hw * v1.comp - lw * v2.comp + hw * v3.comp + lw * v4.comp;             \
    const float gww = -hh * v1.comp + hh * v2.comp - lh * v3.comp + lh * v4.comp;             \
    const float val = w1 * v1.comp + w2 * v2.comp + w3 * v3.comp + w4 * v4.comp;              \
    gh += ghw * tg.comp;                                                                      \
    gw += gww * tg.comp;                                                                      \
    ga += val * g.comp;                                                                       \
  }
          MPF_ACC(x) MPF_ACC(y) MPF_ACC(z) MPF_ACC(w)
#undef MPF_ACC
#pragma unroll
          for (int o = LPI / 2; o >= 1; o >>= 1) {
            gh += __shfl_xor_sync(FULL, gh, o);
            gw += __shfl_xor_sync(FULL, gw, o);
            ga += __shfl_xor_sync(FULL, ga, o);
          }
          const float glx = W * gw, gly = H * gh;  // ref cuh:162-163
          if (li == 2 * l + (p >> 1)) {
            if (p & 1) { gloc_out.z = glx; gloc_out.w = gly; }
            else       { gloc_out.x = glx; gloc_out.y = gly; }
          }
          if (li == l) {
            if (p == 0) gaw_out.x = ga;
            else if (p == 1) gaw_out.y = ga;
            else if (p == 2) gaw_out.z = ga;
            else gaw_out.w = ga;
          }
        }
      }
    }
    if (active) {
      if (li < n_loc4) *reinterpret_cast<float4*>(grad_loc + item * LP * 2 + li * 4) = gloc_out;
      if (li < n_aw4) *reinterpret_cast<float4*>(grad_aw + item * LP + li * 4) = gaw_out;
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
  if (2 * L > l || L > kMaxTiledLevels) return false;
  if (M > 65535 || B > 65535) return false;
  *lpi = l;
  return true;
}

// shapes_host may be null (then linear chunking is used).
static MsdaTiling make_tiling(const int64_t* shapes_host, int L, int S, int Lq) {
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

}  // extern "C"

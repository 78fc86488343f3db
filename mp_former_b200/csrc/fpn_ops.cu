// Layout-crossing kernels of the pixel decoder's FPN stage (ref pixel_decoder/msdeformattn.py:343-351):
//
//   lateral 1x1 conv (token GEMM, channels-last) -> GroupNorm -> + bilinear x2 upsample of the coarser map
//   -> 3x3 conv (cuDNN, wants NCHW: its fp32 NHWC path is 7x slower) -> GroupNorm + ReLU -> 1x1 mask conv (token GEMM)
//
// so the map has to change layout twice.  Instead of separate transposition copies around library ops, the two
// crossings are fused into the elementwise work that sits there anyway:
//
//   mpf_upsample2x_add_nchw_*   out[b,c,h,w] = cur[b,h,w,c] + bilinear_x2(prev)[b,h,w,c]   (CL, CL -> NCHW)
//                               backward: g (NCHW) -> g_cur (CL, a transposition) and g_prev (CL, the x2 adjoint)
//   mpf_groupnorm_nchw2cl_*     y[b,hw,c] = relu?(GroupNorm(x[b,c,hw]))                      (NCHW -> CL)
//                               backward: dy (CL), x (NCHW) -> dx (NCHW), dgamma, dbeta
//
// Every kernel moves a [64 channels x 64 pixels] tile through shared memory: the NCHW side is read/written as
// 16-byte vectors along the pixels (a half warp covers 256 contiguous bytes of one channel), the channels-last side
// as 16-byte vectors along the channels (a half warp covers the 256 contiguous bytes of one pixel).  HBM-bound.
#include "mpf_common.cuh"

namespace mpf {
namespace fpn {

constexpr int kThreads = 256;
constexpr int kTC = 64;            // channels per tile
constexpr int kTP = 64;            // pixels per tile
constexpr int kLd = kTP + 1;       // shared row stride (floats): scalar accesses in both directions, <= 2-way conflicts

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }

// ------------------------------------------------------------------------------------------------
// bilinear x2 upsample (align_corners=False, ATen's upsample_bilinear2d association) + add, CL -> NCHW
//   source index of output i: max(0.5 * (i + 0.5) - 0.5, 0); second tap = first + (first < n - 1)
// grid: (tiles_w * H, C / 64, B)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
upsample2x_add_nchw_fwd_kernel(const float* __restrict__ cur, const float* __restrict__ prev, int H, int W, int C,
                               float* __restrict__ out) {
  __shared__ float tile[kTC * kLd];
  const int Hp = H >> 1, Wp = W >> 1;
  const int tiles_w = (W + kTP - 1) / kTP;
  const int h = blockIdx.x / tiles_w;
  const int x0 = (blockIdx.x % tiles_w) * kTP;
  const int c0 = blockIdx.y * kTC;
  const int b = blockIdx.z;
  const int t = threadIdx.x;

  const float ys = fmaxf(0.5f * (static_cast<float>(h) + 0.5f) - 0.5f, 0.f);
  const int y0 = static_cast<int>(ys);
  const int y1 = y0 + (y0 < Hp - 1 ? 1 : 0);
  const float ly1 = ys - static_cast<float>(y0), ly0 = 1.f - ly1;
  const int c4 = t & 15;
  const float* cur_row = cur + ((static_cast<long long>(b) * H + h) * W) * C + c0 + 4 * c4;
  const float* p0 = prev + ((static_cast<long long>(b) * Hp + y0) * Wp) * C + c0 + 4 * c4;
  const float* p1 = prev + ((static_cast<long long>(b) * Hp + y1) * Wp) * C + c0 + 4 * c4;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int px = (t >> 4) + 16 * k;
    const int x = x0 + px;
    if (x < W) {
      const float xs = fmaxf(0.5f * (static_cast<float>(x) + 0.5f) - 0.5f, 0.f);
      const int xa = static_cast<int>(xs);
      const int xb = xa + (xa < Wp - 1 ? 1 : 0);
      const float lx1 = xs - static_cast<float>(xa), lx0 = 1.f - lx1;
      const float4 a = ld4(cur_row + static_cast<long long>(x) * C);
      const float4 v00 = ld4(p0 + static_cast<long long>(xa) * C), v01 = ld4(p0 + static_cast<long long>(xb) * C);
      const float4 v10 = ld4(p1 + static_cast<long long>(xa) * C), v11 = ld4(p1 + static_cast<long long>(xb) * C);
      float* s = tile + (4 * c4) * kLd + px;
      s[0 * kLd] = a.x + (ly0 * (lx0 * v00.x + lx1 * v01.x) + ly1 * (lx0 * v10.x + lx1 * v11.x));
      s[1 * kLd] = a.y + (ly0 * (lx0 * v00.y + lx1 * v01.y) + ly1 * (lx0 * v10.y + lx1 * v11.y));
      s[2 * kLd] = a.z + (ly0 * (lx0 * v00.z + lx1 * v01.z) + ly1 * (lx0 * v10.z + lx1 * v11.z));
      s[3 * kLd] = a.w + (ly0 * (lx0 * v00.w + lx1 * v01.w) + ly1 * (lx0 * v10.w + lx1 * v11.w));
    }
  }
  __syncthreads();
  const int p4 = t & 15;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int ch = (t >> 4) + 16 * k;
    const int x = x0 + 4 * p4;
    if (x < W) {                                     // W % 4 == 0: whole vectors
      const float* s = tile + ch * kLd + 4 * p4;
      st4(out + ((static_cast<long long>(b) * C + c0 + ch) * H + h) * W + x, make_float4(s[0], s[1], s[2], s[3]));
    }
  }
}

// backward: CTA = 2 low-resolution rows x 32 low-resolution pixels x 32 channels.
//   g_cur[b, 2i..2i+3 rows... ] : the 4 fine rows x 64 fine pixels it covers, transposed to channels-last
//   g_prev[b,i,j,:] = sum_{dy,dx in -1..2} wy[dy] wx[dx] g[b,:,clamp(2i+dy),clamp(2j+dx)],  w = (.25,.75,.75,.25)
// (clamping the index reproduces the border taps of the forward exactly)
constexpr int kBC = 32;                         // channels per CTA
constexpr int kBRows = 6;                       // fine rows staged: 2*i0-1 .. 2*i0+4
// fine columns staged: 2*j0-1 .. 2*j0+64 (66 of them) in rows of kBRowLd floats
constexpr int kBRowLd = 72;
constexpr int kBChLd = kBRows * kBRowLd + 1;    // odd: channel-quad reads hit distinct banks
constexpr int kBSmem = kBC * kBChLd * 4;

__global__ void __launch_bounds__(kThreads)
upsample2x_add_nchw_bwd_kernel(const float* __restrict__ g, int H, int W, int C, float* __restrict__ g_cur,
                               float* __restrict__ g_prev) {
  extern __shared__ float gs[];
  const int Hp = H >> 1, Wp = W >> 1;
  const int tiles_w = (Wp + 31) / 32;
  const int i0 = (blockIdx.x / tiles_w) * 2;
  const int j0 = (blockIdx.x % tiles_w) * 32;
  const int c0 = blockIdx.y * kBC;
  const int b = blockIdx.z;
  const int t = threadIdx.x;
  const float* gb = g + (static_cast<long long>(b) * C + c0) * H * W;

  // stage: column index cc in [0, 66) <-> fine column 2*j0 - 1 + cc (clamped); row rr in [0, 6) <-> 2*i0 - 1 + rr.
  // interior columns cc = 1..64 (fine columns 2*j0 .. 2*j0+63, 16-byte aligned since W % 4 == 0) as float4 loads,
  // 16 per (channel, row); the two halo columns as scalars.
  for (int i = t; i < kBC * kBRows * 16; i += kThreads) {
    const int v = i & 15;
    const int rr = (i >> 4) % kBRows;
    const int ch = (i >> 4) / kBRows;
    int row = 2 * i0 - 1 + rr;
    row = row < 0 ? 0 : (row > H - 1 ? H - 1 : row);
    const int col = 2 * j0 + 4 * v;
    float* d = gs + ch * kBChLd + rr * kBRowLd + 1 + 4 * v;
    const float* src = gb + (static_cast<long long>(ch) * H + row) * W;
    if (col + 3 < W) {
      const float4 q = ld4(src + col);
      d[0] = q.x; d[1] = q.y; d[2] = q.z; d[3] = q.w;
    } else {                                   // right edge of the map: clamp (only the first is ever read back)
      const float e = __ldg(src + W - 1);
#pragma unroll
      for (int u = 0; u < 4; ++u) d[u] = col + u < W ? __ldg(src + col + u) : e;
    }
  }
  for (int i = t; i < kBC * kBRows * 2; i += kThreads) {
    const int side = i & 1;
    const int rr = (i >> 1) % kBRows;
    const int ch = (i >> 1) / kBRows;
    int row = 2 * i0 - 1 + rr;
    row = row < 0 ? 0 : (row > H - 1 ? H - 1 : row);
    int col = side ? 2 * j0 + 64 : 2 * j0 - 1;
    col = col < 0 ? 0 : (col > W - 1 ? W - 1 : col);
    gs[ch * kBChLd + rr * kBRowLd + (side ? 65 : 0)] = __ldg(gb + (static_cast<long long>(ch) * H + row) * W + col);
  }
  __syncthreads();
  const int c4 = t & 7;
  const int pq = t >> 3;                         // 0..31
  // g_cur: fine rows 2*i0 .. 2*i0+3 (staged rows 1..4), fine columns 2*j0 .. 2*j0+63 (staged columns 1..64)
#pragma unroll
  for (int rr = 1; rr <= 4; ++rr) {
    const int row = 2 * i0 - 1 + rr;
    if (row >= H) break;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int cc = 1 + pq + 32 * half;
      const int col = 2 * j0 - 1 + cc;
      if (col < W) {
        const float* s = gs + (4 * c4) * kBChLd + rr * kBRowLd + cc;
        st4(g_cur + ((static_cast<long long>(b) * H + row) * W + col) * C + c0 + 4 * c4,
            make_float4(s[0], s[kBChLd], s[2 * kBChLd], s[3 * kBChLd]));
      }
    }
  }
  // g_prev: low-res rows i0, i0+1; low-res pixel j0 + pq
  const int j = j0 + pq;
  if (j < Wp) {
    const float w4[4] = {0.25f, 0.75f, 0.75f, 0.25f};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int i = i0 + r;
      if (i >= Hp) break;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int dy = 0; dy < 4; ++dy) {
        float rowacc[4] = {0.f, 0.f, 0.f, 0.f};
        const float* s = gs + (4 * c4) * kBChLd + (2 * r + dy) * kBRowLd + 2 * pq;
#pragma unroll
        for (int dx = 0; dx < 4; ++dx) {
#pragma unroll
          for (int e = 0; e < 4; ++e) rowacc[e] += w4[dx] * s[e * kBChLd + dx];
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[e] += w4[dy] * rowacc[e];
      }
      st4(g_prev + ((static_cast<long long>(b) * Hp + i) * Wp + j) * C + c0 + 4 * c4,
          make_float4(acc[0], acc[1], acc[2], acc[3]));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// The same merge with BOTH sides channels-last, for the path where the 3x3 convolution runs on this library's
// tensor-core GEMM (mpf_conv3x3_cl_bf16x3) and no layout change is needed:
//   fwd: out[b,h,w,:] = cur[b,h,w,:] + bilinear_x2(prev)[b,h,w,:]        one float4 of channels per thread
//   bwd: g_cur = g (the same tensor);  g_prev[b,i,j,:] = 4x4 clamped stencil of g     one float4 per thread
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
upsample2x_add_cl_fwd_kernel(const float* __restrict__ cur, const float* __restrict__ prev, int H, int W, int C,
                             long long total4, float* __restrict__ out) {
  const int Hp = H >> 1, Wp = W >> 1, c4n = C >> 2;
  for (long long i = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; i < total4;
       i += static_cast<long long>(gridDim.x) * kThreads) {
    const int c4 = static_cast<int>(i % c4n);
    long long p = i / c4n;
    const int x = static_cast<int>(p % W);
    p /= W;
    const int h = static_cast<int>(p % H);
    const int b = static_cast<int>(p / H);
    const float ys = fmaxf(0.5f * (static_cast<float>(h) + 0.5f) - 0.5f, 0.f);
    const int y0 = static_cast<int>(ys), y1 = y0 + (y0 < Hp - 1 ? 1 : 0);
    const float ly1 = ys - static_cast<float>(y0), ly0 = 1.f - ly1;
    const float xs = fmaxf(0.5f * (static_cast<float>(x) + 0.5f) - 0.5f, 0.f);
    const int xa = static_cast<int>(xs), xb = xa + (xa < Wp - 1 ? 1 : 0);
    const float lx1 = xs - static_cast<float>(xa), lx0 = 1.f - lx1;
    const float* pb = prev + static_cast<long long>(b) * Hp * Wp * C + 4 * c4;
    const float4 a = ld4(cur + i * 4);
    const float4 v00 = ld4(pb + (static_cast<long long>(y0) * Wp + xa) * C), v01 = ld4(pb + (static_cast<long long>(y0) * Wp + xb) * C);
    const float4 v10 = ld4(pb + (static_cast<long long>(y1) * Wp + xa) * C), v11 = ld4(pb + (static_cast<long long>(y1) * Wp + xb) * C);
    float4 o;
    o.x = a.x + (ly0 * (lx0 * v00.x + lx1 * v01.x) + ly1 * (lx0 * v10.x + lx1 * v11.x));
    o.y = a.y + (ly0 * (lx0 * v00.y + lx1 * v01.y) + ly1 * (lx0 * v10.y + lx1 * v11.y));
    o.z = a.z + (ly0 * (lx0 * v00.z + lx1 * v01.z) + ly1 * (lx0 * v10.z + lx1 * v11.z));
    o.w = a.w + (ly0 * (lx0 * v00.w + lx1 * v01.w) + ly1 * (lx0 * v10.w + lx1 * v11.w));
    st4(out + i * 4, o);
  }
}

__global__ void __launch_bounds__(kThreads)
upsample2x_cl_bwd_kernel(const float* __restrict__ g, int H, int W, int C, long long total4,
                         float* __restrict__ g_prev) {
  const int Hp = H >> 1, Wp = W >> 1, c4n = C >> 2;
  for (long long idx = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; idx < total4;
       idx += static_cast<long long>(gridDim.x) * kThreads) {
    const int c4 = static_cast<int>(idx % c4n);
    long long p = idx / c4n;
    const int j = static_cast<int>(p % Wp);
    p /= Wp;
    const int i = static_cast<int>(p % Hp);
    const int b = static_cast<int>(p / Hp);
    const float* gb = g + static_cast<long long>(b) * H * W * C + 4 * c4;
    const float w4[4] = {0.25f, 0.75f, 0.75f, 0.25f};
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int dy = 0; dy < 4; ++dy) {
      int row = 2 * i - 1 + dy;
      row = row < 0 ? 0 : (row > H - 1 ? H - 1 : row);
      float4 ra = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int dx = 0; dx < 4; ++dx) {
        int col = 2 * j - 1 + dx;
        col = col < 0 ? 0 : (col > W - 1 ? W - 1 : col);
        const float4 v = ld4(gb + (static_cast<long long>(row) * W + col) * C);
        ra.x += w4[dx] * v.x; ra.y += w4[dx] * v.y; ra.z += w4[dx] * v.z; ra.w += w4[dx] * v.w;
      }
      acc.x += w4[dy] * ra.x; acc.y += w4[dy] * ra.y; acc.z += w4[dy] * ra.z; acc.w += w4[dy] * ra.w;
    }
    st4(g_prev + idx * 4, acc);
  }
}

// ------------------------------------------------------------------------------------------------
// GroupNorm (+ReLU), NCHW in -> channels-last out
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// statistics: group (b, g) = cpg * HW contiguous floats; grid (splits, B * G); stats[(b*G+g)*2 + {0,1}] += (sum, sumsq)
__global__ void __launch_bounds__(kThreads)
groupnorm_nchw_stats_kernel(const float* __restrict__ x, long long n_group, long long per_cta,
                            double* __restrict__ stats) {
  const float* xg = x + static_cast<long long>(blockIdx.y) * n_group;
  const long long lo = static_cast<long long>(blockIdx.x) * per_cta;
  long long hi = lo + per_cta;
  if (hi > n_group) hi = n_group;
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  long long i = lo + 4ll * threadIdx.x;
  // 4 independent 16-byte loads in flight per thread
  for (; i + 3ll * 4 * kThreads < hi; i += 4ll * 4 * kThreads) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = ld4(xg + i + 4ll * kThreads * u);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      s1[u] += (v[u].x + v[u].y) + (v[u].z + v[u].w);
      s2[u] += (v[u].x * v[u].x + v[u].y * v[u].y) + (v[u].z * v[u].z + v[u].w * v[u].w);
    }
  }
  for (; i < hi; i += 4ll * kThreads) {
    const float4 v = ld4(xg + i);
    s1[0] += (v.x + v.y) + (v.z + v.w);
    s2[0] += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  double d1 = (static_cast<double>(s1[0]) + s1[1]) + (static_cast<double>(s1[2]) + s1[3]);
  double d2 = (static_cast<double>(s2[0]) + s2[1]) + (static_cast<double>(s2[2]) + s2[3]);
  d1 = warp_sum_d(d1);
  d2 = warp_sum_d(d2);
  __shared__ double red[2][kThreads / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = d1; red[1][warp] = d2; }
  __syncthreads();
  if (threadIdx.x < 2) {
    double s = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) s += red[threadIdx.x][w];
    atomicAdd(stats + 2ll * blockIdx.y + threadIdx.x, s);
  }
}

struct GnChan {          // per-channel constants of one tile, in shared memory
  float mean[kTC], rstd[kTC], gamma[kTC], beta[kTC], m1[kTC], m2[kTC];
};

__device__ __forceinline__ void gn_load_channels(GnChan& cs, const double* stats, const float* gamma, const float* beta,
                                                 int b, int c0, int C, int cpg, long long HW, float eps,
                                                 float* mean_out, float* rstd_out, bool publish) {
  const int t = threadIdx.x;
  if (t < kTC) {
    const int c = c0 + t, G = C / cpg, g = c / cpg;
    const double inv_n = 1.0 / (static_cast<double>(HW) * cpg);
    const double* st = stats + (static_cast<long long>(b) * G + g) * 2;
    const double m = st[0] * inv_n;
    double var = st[1] * inv_n - m * m;
    if (var < 0.0) var = 0.0;
    const float mean = static_cast<float>(m);
    const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    cs.mean[t] = mean;
    cs.rstd[t] = rstd;
    cs.gamma[t] = __ldg(gamma + c);
    cs.beta[t] = __ldg(beta + c);
    if (publish && (c % cpg) == 0) {
      mean_out[b * G + g] = mean;
      rstd_out[b * G + g] = rstd;
    }
  }
}

// grid: (ceil(HW / 64), C / 64, B)
__global__ void __launch_bounds__(kThreads)
groupnorm_nchw2cl_apply_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                               const float* __restrict__ beta, const double* __restrict__ stats, long long HW, int C,
                               int cpg, float eps, int relu, float* __restrict__ y, float* __restrict__ mean_out,
                               float* __restrict__ rstd_out) {
  __shared__ float tile[kTC * kLd];
  __shared__ GnChan cs;
  const long long p0 = static_cast<long long>(blockIdx.x) * kTP;
  const int c0 = blockIdx.y * kTC, b = blockIdx.z, t = threadIdx.x;
  gn_load_channels(cs, stats, gamma, beta, b, c0, C, cpg, HW, eps, mean_out, rstd_out, blockIdx.x == 0);
  __syncthreads();
  const int p4 = t & 15;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int ch = (t >> 4) + 16 * k;
    const long long p = p0 + 4 * p4;
    if (p < HW) {
      const float4 v = ld4(x + (static_cast<long long>(b) * C + c0 + ch) * HW + p);
      const float mean = cs.mean[ch], a = cs.rstd[ch], gm = cs.gamma[ch], bt = cs.beta[ch];
      float o[4] = {(v.x - mean) * a * gm + bt, (v.y - mean) * a * gm + bt, (v.z - mean) * a * gm + bt,
                    (v.w - mean) * a * gm + bt};
      float* s = tile + ch * kLd + 4 * p4;
#pragma unroll
      for (int e = 0; e < 4; ++e) s[e] = relu ? fmaxf(o[e], 0.f) : o[e];
    }
  }
  __syncthreads();
  const int c4 = t & 15;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int px = (t >> 4) + 16 * k;
    const long long p = p0 + px;
    if (p < HW) {
      const float* s = tile + (4 * c4) * kLd + px;
      st4(y + (static_cast<long long>(b) * HW + p) * C + c0 + 4 * c4,
          make_float4(s[0], s[kLd], s[2 * kLd], s[3 * kLd]));
    }
  }
}

// backward statistics.  grid: (ceil(HW / (64 * tiles_per_cta)), C / 64, B).
//   gstats[(b*G+g)*2 + {0,1}] += sum gamma*dy', sum gamma*dy'*xhat      (doubles)
//   dgb[c] += sum dy'*xhat;  dgb[C + c] += sum dy'                      (floats)
// dy' = dy masked by the fused ReLU (y > 0 recomputed from x)
__global__ void __launch_bounds__(kThreads)
groupnorm_nchw2cl_bwd_stats_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   const float* __restrict__ mean_in, const float* __restrict__ rstd_in, long long HW,
                                   int C, int cpg, int relu, int tiles_per_cta, double* __restrict__ gstats,
                                   float* __restrict__ dgb) {
  __shared__ float tile[kTC * kLd];
  __shared__ float red[kThreads / 32][16][10];
  const int c0 = blockIdx.y * kTC, b = blockIdx.z, t = threadIdx.x;
  const int G = C / cpg;
  const int c4 = t & 15, p4 = t & 15;
  float gm[4], bt[4], mean, rstd;
  {
    const int c = c0 + 4 * c4, g = c / cpg;
    mean = __ldg(mean_in + b * G + g);
    rstd = __ldg(rstd_in + b * G + g);
#pragma unroll
    for (int e = 0; e < 4; ++e) { gm[e] = __ldg(gamma + c + e); bt[e] = __ldg(beta + c + e); }
  }
  float dg[4] = {0.f, 0.f, 0.f, 0.f}, db[4] = {0.f, 0.f, 0.f, 0.f}, s1 = 0.f, s2 = 0.f;
  for (int tt = 0; tt < tiles_per_cta; ++tt) {
    const long long p0 = (static_cast<long long>(blockIdx.x) * tiles_per_cta + tt) * kTP;
    if (p0 >= HW) break;
    __syncthreads();                                   // previous tile fully consumed
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int ch = (t >> 4) + 16 * k;
      const long long p = p0 + 4 * p4;
      if (p < HW) {
        const float4 v = ld4(x + (static_cast<long long>(b) * C + c0 + ch) * HW + p);
        float* s = tile + ch * kLd + 4 * p4;
        s[0] = v.x; s[1] = v.y; s[2] = v.z; s[3] = v.w;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int px = (t >> 4) + 16 * k;
      const long long p = p0 + px;
      if (p < HW) {
        const float4 d4 = ld4(dy + (static_cast<long long>(b) * HW + p) * C + c0 + 4 * c4);
        const float d[4] = {d4.x, d4.y, d4.z, d4.w};
        const float* s = tile + (4 * c4) * kLd + px;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float xh = (s[e * kLd] - mean) * rstd;
          float de = d[e];
          if (relu && !(xh * gm[e] + bt[e] > 0.f)) de = 0.f;
          dg[e] += de * xh;
          db[e] += de;
          const float gg = de * gm[e];
          s1 += gg;
          s2 += gg * xh;
        }
      }
    }
  }
  // lanes l and l ^ 16 share the channel quad; then the 8 warps
  float vals[10] = {dg[0], dg[1], dg[2], dg[3], db[0], db[1], db[2], db[3], s1, s2};
#pragma unroll
  for (int v = 0; v < 10; ++v) vals[v] += __shfl_xor_sync(0xffffffffu, vals[v], 16);
  const int warp = t >> 5, lane = t & 31;
  if (lane < 16) {
#pragma unroll
    for (int v = 0; v < 10; ++v) red[warp][lane][v] = vals[v];
  }
  __syncthreads();
  if (t < 160) {
    const int q = t / 10, v = t % 10;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) s += red[w][q][v];
    const int c = c0 + 4 * q;
    if (v < 4) atomicAdd(dgb + c + v, s);
    else if (v < 8) atomicAdd(dgb + C + c + (v - 4), s);
    else atomicAdd(gstats + (static_cast<long long>(b) * G + c / cpg) * 2 + (v - 8), static_cast<double>(s));
  }
}

// backward apply: dx[b,c,hw] = rstd * (gamma*dy' - mean_grp(gamma*dy') - xhat * mean_grp(gamma*dy'*xhat))  -> NCHW
__global__ void __launch_bounds__(kThreads)
groupnorm_nchw2cl_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                                   const double* __restrict__ gstats, long long HW, int C, int cpg, int relu,
                                   float* __restrict__ dx) {
  __shared__ float tile[kTC * kLd];
  __shared__ GnChan cs;
  const long long p0 = static_cast<long long>(blockIdx.x) * kTP;
  const int c0 = blockIdx.y * kTC, b = blockIdx.z, t = threadIdx.x;
  const int G = C / cpg;
  if (t < kTC) {
    const int c = c0 + t, g = c / cpg;
    const double inv_n = 1.0 / (static_cast<double>(HW) * cpg);
    const double* st = gstats + (static_cast<long long>(b) * G + g) * 2;
    cs.mean[t] = __ldg(mean_in + b * G + g);
    cs.rstd[t] = __ldg(rstd_in + b * G + g);
    cs.gamma[t] = __ldg(gamma + c);
    cs.beta[t] = __ldg(beta + c);
    cs.m1[t] = static_cast<float>(st[0] * inv_n);
    cs.m2[t] = static_cast<float>(st[1] * inv_n);
  }
  const int c4 = t & 15;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int px = (t >> 4) + 16 * k;
    const long long p = p0 + px;
    if (p < HW) {
      const float4 d = ld4(dy + (static_cast<long long>(b) * HW + p) * C + c0 + 4 * c4);
      float* s = tile + (4 * c4) * kLd + px;
      s[0] = d.x; s[kLd] = d.y; s[2 * kLd] = d.z; s[3 * kLd] = d.w;
    }
  }
  __syncthreads();
  const int p4 = t & 15;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int ch = (t >> 4) + 16 * k;
    const long long p = p0 + 4 * p4;
    if (p < HW) {
      const long long off = (static_cast<long long>(b) * C + c0 + ch) * HW + p;
      const float4 v4 = ld4(x + off);
      const float v[4] = {v4.x, v4.y, v4.z, v4.w};
      const float mean = cs.mean[ch], rstd = cs.rstd[ch], gm = cs.gamma[ch], bt = cs.beta[ch];
      const float m1 = cs.m1[ch], m2 = cs.m2[ch];
      const float* s = tile + ch * kLd + 4 * p4;
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float xh = (v[e] - mean) * rstd;
        float d = s[e];
        if (relu && !(xh * gm + bt > 0.f)) d = 0.f;
        o[e] = rstd * (d * gm - m1 - xh * m2);
      }
      st4(dx + off, make_float4(o[0], o[1], o[2], o[3]));
    }
  }
}

static bool gn_shape_ok(int C, int G, long long HW) {
  if (G <= 0 || C % G) return false;
  const int cpg = C / G;
  return C % kTC == 0 && cpg % 4 == 0 && kTC % cpg == 0 && HW % 4 == 0;
}

}  // namespace fpn
}  // namespace mpf

extern "C" {

int mpf_upsample2x_add_nchw_fwd_f32(const float* cur, const float* prev, int batch, int H, int W, int C, float* out,
                                    void* stream) {
  using namespace mpf;
  using namespace mpf::fpn;
  clear_error();
  MPF_REQUIRE(cur && prev && out, "upsample2x_add_fwd: null pointer argument");
  MPF_REQUIRE(batch > 0 && H >= 2 && W >= 4 && H % 2 == 0 && W % 4 == 0 && C > 0 && C % kTC == 0,
              "upsample2x_add_fwd: needs even H, W %% 4 == 0 and C %% 64 == 0 (H=%d W=%d C=%d)", H, W, C);
  MPF_REQUIRE(aligned16(cur) && aligned16(prev) && aligned16(out), "upsample2x_add_fwd: 16-byte alignment");
  MPF_REQUIRE(batch <= 65535 && C / kTC <= 65535, "upsample2x_add_fwd: grid too large");
  const int tiles_w = (W + kTP - 1) / kTP;
  dim3 grid(static_cast<unsigned>(tiles_w) * H, C / kTC, batch);
  upsample2x_add_nchw_fwd_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(cur, prev, H, W, C, out);
  count_launch();
  return finish_launch("upsample2x_add_fwd");
}

int mpf_upsample2x_add_nchw_bwd_f32(const float* g, int batch, int H, int W, int C, float* g_cur, float* g_prev,
                                    void* stream) {
  using namespace mpf;
  using namespace mpf::fpn;
  clear_error();
  MPF_REQUIRE(g && g_cur && g_prev, "upsample2x_add_bwd: null pointer argument");
  MPF_REQUIRE(batch > 0 && H >= 2 && W >= 4 && H % 2 == 0 && W % 4 == 0 && C > 0 && C % kBC == 0,
              "upsample2x_add_bwd: needs even H, W %% 4 == 0 and C %% 32 == 0 (H=%d W=%d C=%d)", H, W, C);
  MPF_REQUIRE(aligned16(g_cur) && aligned16(g_prev), "upsample2x_add_bwd: 16-byte alignment");
  MPF_REQUIRE(batch <= 65535 && C / kBC <= 65535, "upsample2x_add_bwd: grid too large");
  static unsigned long long configured_on = 0;
  if (first_use_on_this_device(configured_on)) {
    MPF_CUDA_OK(cudaFuncSetAttribute(upsample2x_add_nchw_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBSmem));
  }
  const int Hp = H / 2, Wp = W / 2;
  dim3 grid(static_cast<unsigned>((Wp + 31) / 32) * ((Hp + 1) / 2), C / kBC, batch);
  upsample2x_add_nchw_bwd_kernel<<<grid, kThreads, kBSmem, static_cast<cudaStream_t>(stream)>>>(g, H, W, C, g_cur, g_prev);
  count_launch();
  return finish_launch("upsample2x_add_bwd");
}

int mpf_upsample2x_add_cl_fwd_f32(const float* cur, const float* prev, int batch, int H, int W, int C, float* out,
                                  void* stream) {
  using namespace mpf;
  using namespace mpf::fpn;
  clear_error();
  MPF_REQUIRE(cur && prev && out, "upsample2x_add_cl_fwd: null pointer argument");
  MPF_REQUIRE(batch > 0 && H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0 && C > 0 && C % 4 == 0,
              "upsample2x_add_cl_fwd: needs even H, W and C %% 4 == 0 (H=%d W=%d C=%d)", H, W, C);
  MPF_REQUIRE(aligned16(cur) && aligned16(prev) && aligned16(out), "upsample2x_add_cl_fwd: 16-byte alignment");
  const long long total4 = static_cast<long long>(batch) * H * W * (C / 4);
  long long blocks = (total4 + kThreads - 1) / kThreads;
  if (blocks > 148ll * 32) blocks = 148ll * 32;
  upsample2x_add_cl_fwd_kernel<<<static_cast<int>(blocks), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      cur, prev, H, W, C, total4, out);
  count_launch();
  return finish_launch("upsample2x_add_cl_fwd");
}

int mpf_upsample2x_cl_bwd_f32(const float* g, int batch, int H, int W, int C, float* g_prev, void* stream) {
  using namespace mpf;
  using namespace mpf::fpn;
  clear_error();
  MPF_REQUIRE(g && g_prev, "upsample2x_cl_bwd: null pointer argument");
  MPF_REQUIRE(batch > 0 && H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0 && C > 0 && C % 4 == 0,
              "upsample2x_cl_bwd: needs even H, W and C %% 4 == 0 (H=%d W=%d C=%d)", H, W, C);
  MPF_REQUIRE(aligned16(g) && aligned16(g_prev), "upsample2x_cl_bwd: 16-byte alignment");
  const long long total4 = static_cast<long long>(batch) * (H / 2) * (W / 2) * (C / 4);
  long long blocks = (total4 + kThreads - 1) / kThreads;
  if (blocks > 148ll * 32) blocks = 148ll * 32;
  upsample2x_cl_bwd_kernel<<<static_cast<int>(blocks), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(g, H, W, C,
                                                                                                      total4, g_prev);
  count_launch();
  return finish_launch("upsample2x_cl_bwd");
}

int mpf_groupnorm_nchw2cl_fwd_f32(const float* x, const float* gamma, const float* beta, float eps, int batch,
                                  long long HW, int C, int groups, int relu, float* y, float* mean, float* rstd,
                                  double* stats_ws, void* stream) {
  using namespace mpf;
  using namespace mpf::fpn;
  clear_error();
  MPF_REQUIRE(x && gamma && beta && y && mean && rstd && stats_ws, "groupnorm_nchw2cl_fwd: null pointer argument");
  MPF_REQUIRE(batch > 0 && HW > 0 && gn_shape_ok(C, groups, HW),
              "groupnorm_nchw2cl_fwd: unsupported shape (C=%d, groups=%d, HW=%lld)", C, groups, HW);
  MPF_REQUIRE(aligned16(x) && aligned16(y), "groupnorm_nchw2cl_fwd: 16-byte alignment");
  MPF_REQUIRE(batch <= 65535 && static_cast<long long>(batch) * groups <= 65535, "groupnorm_nchw2cl_fwd: grid too large");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int cpg = C / groups;
  MPF_CUDA_OK(cudaMemsetAsync(stats_ws, 0, sizeof(double) * 2 * batch * groups, st));
  const long long n_group = static_cast<long long>(cpg) * HW;
  long long splits = (148ll * 8 + batch * groups - 1) / (batch * groups);
  const long long min_per = 4ll * kThreads * 4;                  // one unrolled round per thread
  if (splits > (n_group + min_per - 1) / min_per) splits = (n_group + min_per - 1) / min_per;
  if (splits < 1) splits = 1;
  long long per_cta = (n_group + splits - 1) / splits;
  per_cta = (per_cta + 3) / 4 * 4;
  splits = (n_group + per_cta - 1) / per_cta;
  groupnorm_nchw_stats_kernel<<<dim3(static_cast<unsigned>(splits), batch * groups), kThreads, 0, st>>>(x, n_group,
                                                                                                      per_cta, stats_ws);
  dim3 grid(static_cast<unsigned>((HW + kTP - 1) / kTP), C / kTC, batch);
  groupnorm_nchw2cl_apply_kernel<<<grid, kThreads, 0, st>>>(x, gamma, beta, stats_ws, HW, C, cpg, eps, relu, y, mean, rstd);
  count_launch(2);
  return finish_launch("groupnorm_nchw2cl_fwd");
}

int mpf_groupnorm_nchw2cl_bwd_f32(const float* dy, const float* x, const float* gamma, const float* beta,
                                  const float* mean, const float* rstd, int batch, long long HW, int C, int groups,
                                  int relu, float* dx, float* dgamma_dbeta, double* stats_ws, void* stream) {
  using namespace mpf;
  using namespace mpf::fpn;
  clear_error();
  MPF_REQUIRE(dy && x && gamma && beta && mean && rstd && dx && dgamma_dbeta && stats_ws,
              "groupnorm_nchw2cl_bwd: null pointer argument");
  MPF_REQUIRE(batch > 0 && HW > 0 && gn_shape_ok(C, groups, HW),
              "groupnorm_nchw2cl_bwd: unsupported shape (C=%d, groups=%d, HW=%lld)", C, groups, HW);
  MPF_REQUIRE(aligned16(x) && aligned16(dy) && aligned16(dx), "groupnorm_nchw2cl_bwd: 16-byte alignment");
  MPF_REQUIRE(batch <= 65535, "groupnorm_nchw2cl_bwd: grid too large");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int cpg = C / groups;
  MPF_CUDA_OK(cudaMemsetAsync(stats_ws, 0, sizeof(double) * 2 * batch * groups, st));
  MPF_CUDA_OK(cudaMemsetAsync(dgamma_dbeta, 0, sizeof(float) * 2 * C, st));
  const long long tiles = (HW + kTP - 1) / kTP;
  int tiles_per_cta = 8;
  while (tiles_per_cta > 1 && (tiles / tiles_per_cta) * (C / kTC) * batch < 148 * 4) tiles_per_cta >>= 1;
  dim3 sgrid(static_cast<unsigned>((tiles + tiles_per_cta - 1) / tiles_per_cta), C / kTC, batch);
  groupnorm_nchw2cl_bwd_stats_kernel<<<sgrid, kThreads, 0, st>>>(dy, x, gamma, beta, mean, rstd, HW, C, cpg, relu,
                                                                 tiles_per_cta, stats_ws, dgamma_dbeta);
  dim3 grid(static_cast<unsigned>(tiles), C / kTC, batch);
  groupnorm_nchw2cl_bwd_apply_kernel<<<grid, kThreads, 0, st>>>(dy, x, gamma, beta, mean, rstd, stats_ws, HW, C, cpg,
                                                                relu, dx);
  count_launch(2);
  return finish_launch("groupnorm_nchw2cl_bwd");
}

}  // extern "C"

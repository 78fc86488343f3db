// Row-wise HBM-bound kernels of the encoder / decoder layers (C = 256 channels per token):
//
//   y = LayerNorm(x + r) * gamma + beta            (ref: pixel_decoder/msdeformattn.py:125-126,129 norm1/norm2 of the
//                                                    encoder layer; decoder :52,:112,:169 post-norm blocks)
//   its backward (dx = ds for both addends, dgamma, dbeta)
//   column sums of a [rows, C] matrix               (bias gradients of the nn.Linear layers)
//
// The reference runs `src + src2` and `LayerNorm` as separate PyTorch kernels (5 passes over the token matrix
// forward, 5 backward); here the sum is never written: forward = read x, r / write y (3 passes), backward = read
// x, r, dy / write dx (4 passes) with the per-channel dgamma / dbeta partial sums kept in registers across the rows a
// warp owns.  One warp per token row: 32 lanes x VPT float4 = C channels, 16-byte coalesced accesses, mean / variance
// by warp shuffles (two-pass: mean, then sum of squared deviations, both in fp32).
#include "mpf_common.cuh"

namespace mpf {

constexpr int kRowThreads = 256;           // 8 warps = 8 rows in flight per CTA
constexpr int kRowWarps = kRowThreads / 32;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int VPT>   // float4 per lane: C = 128 * VPT
__global__ void __launch_bounds__(kRowThreads)
add_layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ r, const float* __restrict__ gamma,
                         const float* __restrict__ beta, float eps, long long rows, float* __restrict__ y,
                         float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  constexpr int C = 128 * VPT;
  const int lane = threadIdx.x & 31;
  const long long warp0 = static_cast<long long>(blockIdx.x) * kRowWarps + (threadIdx.x >> 5);
  const long long nwarps = static_cast<long long>(gridDim.x) * kRowWarps;
  float4 gm[VPT], bt[VPT];
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    gm[i] = __ldg(reinterpret_cast<const float4*>(gamma) + i * 32 + lane);
    bt[i] = __ldg(reinterpret_cast<const float4*>(beta) + i * 32 + lane);
  }
  for (long long row = warp0; row < rows; row += nwarps) {
    const float4* xr = reinterpret_cast<const float4*>(x + row * C);
    float4 s[VPT];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      s[i] = __ldg(xr + i * 32 + lane);
      if (r != nullptr) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(r + row * C) + i * 32 + lane);
        s[i].x += t.x; s[i].y += t.y; s[i].z += t.z; s[i].w += t.w;
      }
      sum += (s[i].x + s[i].y) + (s[i].z + s[i].w);
    }
    const float mean = warp_sum(sum) * (1.f / C);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const float a = s[i].x - mean, b = s[i].y - mean, c = s[i].z - mean, d = s[i].w - mean;
      sq += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(sq) * (1.f / C) + eps);
    float4* yr = reinterpret_cast<float4*>(y + row * C);
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      float4 o;
      o.x = (s[i].x - mean) * rstd * gm[i].x + bt[i].x;
      o.y = (s[i].y - mean) * rstd * gm[i].y + bt[i].y;
      o.z = (s[i].z - mean) * rstd * gm[i].z + bt[i].z;
      o.w = (s[i].w - mean) * rstd * gm[i].w + bt[i].w;
      yr[i * 32 + lane] = o;
    }
    if (lane == 0) {
      mean_out[row] = mean;
      rstd_out[row] = rstd;
    }
  }
}

// dx = rstd * (g - mean_c(g) - xhat * mean_c(g * xhat)),  g = dy * gamma,  xhat = (x + r - mean) * rstd
// partial[blockIdx.x][0][c] = sum over this CTA's rows of dy * xhat, [1][c] = sum of dy, [2][c] = sum of dx (the bias
// gradient of the Linear layer that produced the normalised sum's second addend).
template <int VPT>
__global__ void __launch_bounds__(kRowThreads)
add_layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ r,
                         const float* __restrict__ gamma, const float* __restrict__ mean_in,
                         const float* __restrict__ rstd_in, long long rows, float* __restrict__ dx,
                         float* __restrict__ partial) {
  constexpr int C = 128 * VPT;
  __shared__ float4 red[kRowWarps][VPT][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long warp0 = static_cast<long long>(blockIdx.x) * kRowWarps + w;
  const long long nwarps = static_cast<long long>(gridDim.x) * kRowWarps;
  float4 gm[VPT], dg[VPT], db[VPT], ds[VPT];
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    gm[i] = __ldg(reinterpret_cast<const float4*>(gamma) + i * 32 + lane);
    dg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    ds[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (long long row = warp0; row < rows; row += nwarps) {
    const float mean = __ldg(mean_in + row), rstd = __ldg(rstd_in + row);
    float4 xh[VPT], g[VPT];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      float4 s = __ldg(reinterpret_cast<const float4*>(x + row * C) + i * 32 + lane);
      if (r != nullptr) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(r + row * C) + i * 32 + lane);
        s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
      }
      const float4 d = __ldg(reinterpret_cast<const float4*>(dy + row * C) + i * 32 + lane);
      xh[i] = make_float4((s.x - mean) * rstd, (s.y - mean) * rstd, (s.z - mean) * rstd, (s.w - mean) * rstd);
      g[i] = make_float4(d.x * gm[i].x, d.y * gm[i].y, d.z * gm[i].z, d.w * gm[i].w);
      s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
      s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
      dg[i].x += d.x * xh[i].x; dg[i].y += d.y * xh[i].y; dg[i].z += d.z * xh[i].z; dg[i].w += d.w * xh[i].w;
      db[i].x += d.x; db[i].y += d.y; db[i].z += d.z; db[i].w += d.w;
    }
    const float m1 = warp_sum(s1) * (1.f / C), m2 = warp_sum(s2) * (1.f / C);
    float4* dr = reinterpret_cast<float4*>(dx + row * C);
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      float4 o;
      o.x = rstd * (g[i].x - m1 - xh[i].x * m2);
      o.y = rstd * (g[i].y - m1 - xh[i].y * m2);
      o.z = rstd * (g[i].z - m1 - xh[i].z * m2);
      o.w = rstd * (g[i].w - m1 - xh[i].w * m2);
      dr[i * 32 + lane] = o;
      ds[i].x += o.x; ds[i].y += o.y; ds[i].z += o.z; ds[i].w += o.w;
    }
  }
  // three rounds through one [warps][VPT][32] staging array: dgamma, dbeta, column sums of dx
#pragma unroll
  for (int which = 0; which < 3; ++which) {
    if (which) __syncthreads();
#pragma unroll
    for (int i = 0; i < VPT; ++i) red[w][i][lane] = which == 0 ? dg[i] : (which == 1 ? db[i] : ds[i]);
    __syncthreads();
    for (int slot = threadIdx.x; slot < VPT * 32; slot += kRowThreads) {
      const int i = slot >> 5, l = slot & 31;
      float4 a = red[0][i][l];
#pragma unroll
      for (int ww = 1; ww < kRowWarps; ++ww) {
        const float4 b = red[ww][i][l];
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      }
      reinterpret_cast<float4*>(partial + (static_cast<long long>(blockIdx.x) * 3 + which) * C)[i * 32 + l] = a;
    }
  }
}

// out[c] += sum over rows of x[row][c]   (out zeroed by the launcher); C % 4 == 0, C <= 4096.
// blockDim = (C/4 rounded up to a multiple of 32 lanes... ) -- simple layout: thread t owns float4 column chunk
// t % chunks and row phase t / chunks; partial sums meet in shared memory, one atomicAdd per column per CTA.
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ x, long long rows, int C, long long ld, float* __restrict__ out) {
  extern __shared__ float4 sh[];
  const int chunks = C >> 2;
  const int rpb = 256 / chunks > 0 ? 256 / chunks : 1;          // row phases per CTA (chunks <= 256)
  const int t = threadIdx.x;
  const int cchunk = t % chunks, phase = t / chunks;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (phase < rpb) {
    const long long step = static_cast<long long>(gridDim.x) * rpb;
    long long row = static_cast<long long>(blockIdx.x) * rpb + phase;
    // eight independent 16-byte loads in flight per thread: the grid is kept small (4 CTAs per SM) because every
    // CTA ends with atomics on the same C addresses, and ~1200 CTAs finishing together serialise there for about as
    // long as the reads take (measured 3.3 TB/s with 8 CTAs per SM and 4 loads)
    float4 a4[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) a4[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (; row + 7 * step < rows; row += 8 * step) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldg(reinterpret_cast<const float4*>(x + (row + u * step) * ld) + cchunk);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        a4[u & 3].x += v[u].x; a4[u & 3].y += v[u].y; a4[u & 3].z += v[u].z; a4[u & 3].w += v[u].w;
      }
    }
    for (; row < rows; row += step) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + row * ld) + cchunk);
      a4[0].x += v.x; a4[0].y += v.y; a4[0].z += v.z; a4[0].w += v.w;
    }
    acc.x = (a4[0].x + a4[1].x) + (a4[2].x + a4[3].x);
    acc.y = (a4[0].y + a4[1].y) + (a4[2].y + a4[3].y);
    acc.z = (a4[0].z + a4[1].z) + (a4[2].z + a4[3].z);
    acc.w = (a4[0].w + a4[1].w) + (a4[2].w + a4[3].w);
  }
  sh[t] = acc;
  __syncthreads();
  if (t < chunks) {
    float4 a = sh[t];
    for (int p = 1; p < rpb; ++p) {
      const float4 b = sh[p * chunks + t];
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    atomicAdd(out + 4 * t + 0, a.x);
    atomicAdd(out + 4 * t + 1, a.y);
    atomicAdd(out + 4 * t + 2, a.z);
    atomicAdd(out + 4 * t + 3, a.w);
  }
}

static int row_grid(long long rows) {
  long long blocks = (rows + kRowWarps - 1) / kRowWarps;
  const long long cap = 148 * 8;                 // persistent: 8 CTAs of 256 threads per SM
  return static_cast<int>(blocks < cap ? blocks : cap);
}

}  // namespace mpf

extern "C" {

int mpf_add_layernorm_partials(long long rows) { return mpf::row_grid(rows); }

int mpf_add_layernorm_fwd_f32(const float* x, const float* r, const float* gamma, const float* beta, float eps,
                              long long rows, int C, float* y, float* mean, float* rstd, void* stream) {
  using namespace mpf;
  clear_error();
  MPF_REQUIRE(x && gamma && beta && y && mean && rstd, "add_layernorm_fwd: null pointer argument");
  MPF_REQUIRE(rows > 0, "add_layernorm_fwd: rows must be positive");
  MPF_REQUIRE(aligned16(x) && aligned16(y) && aligned16(gamma) && aligned16(beta) && (r == nullptr || aligned16(r)),
              "add_layernorm_fwd: pointers must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = row_grid(rows);
  switch (C) {
    case 128: add_layernorm_fwd_kernel<1><<<grid, kRowThreads, 0, st>>>(x, r, gamma, beta, eps, rows, y, mean, rstd); break;
    case 256: add_layernorm_fwd_kernel<2><<<grid, kRowThreads, 0, st>>>(x, r, gamma, beta, eps, rows, y, mean, rstd); break;
    case 512: add_layernorm_fwd_kernel<4><<<grid, kRowThreads, 0, st>>>(x, r, gamma, beta, eps, rows, y, mean, rstd); break;
    default:
      set_error("add_layernorm_fwd: C = %d is not supported (128, 256, 512)", C);
      return MPF_ERR_UNSUPPORTED;
  }
  count_launch();
  return finish_launch("add_layernorm_fwd");
}

int mpf_add_layernorm_bwd_f32(const float* dy, const float* x, const float* r, const float* gamma, const float* mean,
                              const float* rstd, long long rows, int C, float* dx, float* partial, void* stream) {
  using namespace mpf;
  clear_error();
  MPF_REQUIRE(dy && x && gamma && mean && rstd && dx && partial, "add_layernorm_bwd: null pointer argument");
  MPF_REQUIRE(rows > 0, "add_layernorm_bwd: rows must be positive");
  MPF_REQUIRE(aligned16(dy) && aligned16(x) && aligned16(dx) && aligned16(gamma) && aligned16(partial) &&
                  (r == nullptr || aligned16(r)),
              "add_layernorm_bwd: pointers must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = row_grid(rows);
  switch (C) {
    case 128: add_layernorm_bwd_kernel<1><<<grid, kRowThreads, 0, st>>>(dy, x, r, gamma, mean, rstd, rows, dx, partial); break;
    case 256: add_layernorm_bwd_kernel<2><<<grid, kRowThreads, 0, st>>>(dy, x, r, gamma, mean, rstd, rows, dx, partial); break;
    case 512: add_layernorm_bwd_kernel<4><<<grid, kRowThreads, 0, st>>>(dy, x, r, gamma, mean, rstd, rows, dx, partial); break;
    default:
      set_error("add_layernorm_bwd: C = %d is not supported (128, 256, 512)", C);
      return MPF_ERR_UNSUPPORTED;
  }
  count_launch();
  return finish_launch("add_layernorm_bwd");
}

int mpf_colsum_f32(const float* x, long long rows, int C, long long ld, float* out, void* stream) {
  using namespace mpf;
  clear_error();
  MPF_REQUIRE(x && out && rows > 0 && C > 0, "colsum: bad arguments");
  MPF_REQUIRE(C % 4 == 0 && C <= 1024 && ld % 4 == 0 && ld >= C && aligned16(x),
              "colsum: C must be a multiple of 4 (<= 1024), ld a multiple of 4, x 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MPF_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(float) * C, st));
  const int chunks = C / 4;
  const int rpb = 256 / chunks > 0 ? 256 / chunks : 1;
  long long blocks = (rows + rpb - 1) / rpb;
  if (blocks > 148 * 4) blocks = 148 * 4;
  colsum_kernel<<<static_cast<int>(blocks), 256, 256 * sizeof(float4), st>>>(x, rows, C, ld, out);
  count_launch();
  return finish_launch("colsum");
}

}  // extern "C"

namespace mpf {

// ------------------------------------------------------------------------------------------------
// GroupNorm over channels-last maps (x [B, HW, C], groups of C/G consecutive channels; the pixel decoder uses
// G = 32, C = 256), optionally fused with the ReLU that follows it (ref pixel_decoder/msdeformattn.py:216-219 input
// projections, :262-275 lateral / output convs with norm "GN").  PyTorch's kernel runs 1.7 ms per [16,256,256,256]
// map; here: one statistics pass (fp32 partial sums per thread, combined in
// fp64 atomics so E[x^2] - E[x]^2 does not cancel) and one normalise pass, 16-byte coalesced accesses.
//   stats[b][g] = (sum, sum of squares) as doubles, zeroed by the launcher.
// ------------------------------------------------------------------------------------------------
constexpr int kGnThreads = 256;

// thread t: channel quad q = t % (C/4), row phase t / (C/4); rows [row0, row1) of image b = blockIdx.y
__global__ void __launch_bounds__(kGnThreads)
groupnorm_stats_kernel(const float* __restrict__ x, long long HW, int C, int cpg, int rows_per_cta,
                       double* __restrict__ stats) {
  const int quads = C >> 2;
  const int rpb = kGnThreads / quads;
  const int q = threadIdx.x % quads, phase = threadIdx.x / quads;
  const int b = blockIdx.y;
  const long long row0 = static_cast<long long>(blockIdx.x) * rows_per_cta;
  long long row1 = row0 + rows_per_cta;
  if (row1 > HW) row1 = HW;
  float s1 = 0.f, s2 = 0.f;
  if (phase < rpb) {
    const float* xb = x + (static_cast<long long>(b) * HW) * C + 4 * q;
    for (long long r = row0 + phase; r < row1; r += rpb) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(xb + r * C));
      s1 += (v.x + v.y) + (v.z + v.w);
      s2 += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    }
  }
  // combine the quads of one group inside the warp when a group spans several lanes (cpg = 8 -> 2 lanes)
  const int lanes_per_group = cpg >> 2;          // cpg is a multiple of 4
  for (int o = 1; o < lanes_per_group; o <<= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if (phase < rpb && (q % lanes_per_group) == 0) {
    const int g = (4 * q) / cpg;
    double* st = stats + (static_cast<long long>(b) * (C / cpg) + g) * 2;
    atomicAdd(st, static_cast<double>(s1));
    atomicAdd(st + 1, static_cast<double>(s2));
  }
}

__global__ void __launch_bounds__(kGnThreads)
groupnorm_apply_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                       const double* __restrict__ stats, long long HW, int C, int cpg, float eps, int relu,
                       float* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  const int quads = C >> 2;
  const int b = blockIdx.y;
  const long long total = HW * quads;
  const int G = C / cpg;
  const double inv_n = 1.0 / (static_cast<double>(HW) * cpg);
  for (long long i = static_cast<long long>(blockIdx.x) * kGnThreads + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * kGnThreads) {
    const int q = static_cast<int>(i % quads);
    const int g = (4 * q) / cpg;
    const double* st = stats + (static_cast<long long>(b) * G + g) * 2;
    const double m = st[0] * inv_n;
    double var = st[1] * inv_n - m * m;
    if (var < 0.0) var = 0.0;
    const float mean = static_cast<float>(m);
    const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    if (i < quads && (q % (cpg >> 2)) == 0) {        // first row of the image: publish the statistics once
      mean_out[b * G + g] = mean;
      rstd_out[b * G + g] = rstd;
    }
    const long long off = (static_cast<long long>(b) * HW) * C + i * 4;
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + off));
    const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + q);
    const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + q);
    float4 o;
    o.x = (v.x - mean) * rstd * gm.x + bt.x;
    o.y = (v.y - mean) * rstd * gm.y + bt.y;
    o.z = (v.z - mean) * rstd * gm.z + bt.z;
    o.w = (v.w - mean) * rstd * gm.w + bt.w;
    if (relu) {
      o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
    }
    *reinterpret_cast<float4*>(y + off) = o;
  }
}

// backward pass 1: per (b, g) sums of g = gamma*dy' and g*xhat (doubles), per channel sums of dy'*xhat and dy'
// (floats, [2][C]); dy' = dy masked by the fused ReLU (y > 0 recomputed from x).
__global__ void __launch_bounds__(kGnThreads)
groupnorm_bwd_stats_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                           const float* __restrict__ beta, const float* __restrict__ mean_in,
                           const float* __restrict__ rstd_in, long long HW, int C, int cpg, int relu,
                           int rows_per_cta, double* __restrict__ gstats, float* __restrict__ dgb) {
  const int quads = C >> 2;
  const int rpb = kGnThreads / quads;
  const int q = threadIdx.x % quads, phase = threadIdx.x / quads;
  const int b = blockIdx.y;
  const int G = C / cpg;
  const long long row0 = static_cast<long long>(blockIdx.x) * rows_per_cta;
  long long row1 = row0 + rows_per_cta;
  if (row1 > HW) row1 = HW;
  float s1 = 0.f, s2 = 0.f;
  float4 dg = make_float4(0.f, 0.f, 0.f, 0.f), db = make_float4(0.f, 0.f, 0.f, 0.f);
  if (phase < rpb) {
    const int g = (4 * q) / cpg;
    const float mean = __ldg(mean_in + b * G + g), rstd = __ldg(rstd_in + b * G + g);
    const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + q);
    const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + q);
    const long long base = (static_cast<long long>(b) * HW) * C + 4 * q;
    for (long long r = row0 + phase; r < row1; r += rpb) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + base + r * C));
      float4 d = __ldg(reinterpret_cast<const float4*>(dy + base + r * C));
      const float4 xh = make_float4((v.x - mean) * rstd, (v.y - mean) * rstd, (v.z - mean) * rstd, (v.w - mean) * rstd);
      if (relu) {
        if (!(xh.x * gm.x + bt.x > 0.f)) d.x = 0.f;
        if (!(xh.y * gm.y + bt.y > 0.f)) d.y = 0.f;
        if (!(xh.z * gm.z + bt.z > 0.f)) d.z = 0.f;
        if (!(xh.w * gm.w + bt.w > 0.f)) d.w = 0.f;
      }
      const float4 gg = make_float4(d.x * gm.x, d.y * gm.y, d.z * gm.z, d.w * gm.w);
      s1 += (gg.x + gg.y) + (gg.z + gg.w);
      s2 += (gg.x * xh.x + gg.y * xh.y) + (gg.z * xh.z + gg.w * xh.w);
      dg.x += d.x * xh.x; dg.y += d.y * xh.y; dg.z += d.z * xh.z; dg.w += d.w * xh.w;
      db.x += d.x; db.y += d.y; db.z += d.z; db.w += d.w;
    }
  }
  const int lanes_per_group = cpg >> 2;
  for (int o = 1; o < lanes_per_group; o <<= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if (phase < rpb && (q % lanes_per_group) == 0) {
    const int g = (4 * q) / cpg;
    double* st = gstats + (static_cast<long long>(b) * G + g) * 2;
    atomicAdd(st, static_cast<double>(s1));
    atomicAdd(st + 1, static_cast<double>(s2));
  }
  // dgamma / dbeta: combine the CTA's row phases in shared memory first -- one atomic per channel and CTA instead of
  // one per thread (every CTA of every image targets the same 2*C addresses)
  __shared__ float4 sdg[kGnThreads], sdb[kGnThreads];
  sdg[threadIdx.x] = dg;
  sdb[threadIdx.x] = db;
  __syncthreads();
  if (threadIdx.x < quads) {
    float4 a = sdg[threadIdx.x], c = sdb[threadIdx.x];
    for (int p = 1; p < rpb; ++p) {
      const float4 a2 = sdg[p * quads + threadIdx.x], c2 = sdb[p * quads + threadIdx.x];
      a.x += a2.x; a.y += a2.y; a.z += a2.z; a.w += a2.w;
      c.x += c2.x; c.y += c2.y; c.z += c2.z; c.w += c2.w;
    }
    float* pg = dgb + 4 * threadIdx.x;
    atomicAdd(pg + 0, a.x); atomicAdd(pg + 1, a.y); atomicAdd(pg + 2, a.z); atomicAdd(pg + 3, a.w);
    float* pb = dgb + C + 4 * threadIdx.x;
    atomicAdd(pb + 0, c.x); atomicAdd(pb + 1, c.y); atomicAdd(pb + 2, c.z); atomicAdd(pb + 3, c.w);
  }
}

// backward pass 2: dx = rstd * (g - mean_grp(g) - xhat * mean_grp(g * xhat))
__global__ void __launch_bounds__(kGnThreads)
groupnorm_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                           const float* __restrict__ beta, const float* __restrict__ mean_in,
                           const float* __restrict__ rstd_in, const double* __restrict__ gstats, long long HW, int C,
                           int cpg, int relu, float* __restrict__ dx) {
  const int quads = C >> 2;
  const int b = blockIdx.y;
  const int G = C / cpg;
  const long long total = HW * quads;
  const double inv_n = 1.0 / (static_cast<double>(HW) * cpg);
  for (long long i = static_cast<long long>(blockIdx.x) * kGnThreads + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * kGnThreads) {
    const int q = static_cast<int>(i % quads);
    const int g = (4 * q) / cpg;
    const float mean = __ldg(mean_in + b * G + g), rstd = __ldg(rstd_in + b * G + g);
    const double* st = gstats + (static_cast<long long>(b) * G + g) * 2;
    const float m1 = static_cast<float>(st[0] * inv_n), m2 = static_cast<float>(st[1] * inv_n);
    const long long off = (static_cast<long long>(b) * HW) * C + i * 4;
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + off));
    float4 d = __ldg(reinterpret_cast<const float4*>(dy + off));
    const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + q);
    const float4 xh = make_float4((v.x - mean) * rstd, (v.y - mean) * rstd, (v.z - mean) * rstd, (v.w - mean) * rstd);
    if (relu) {
      const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + q);
      if (!(xh.x * gm.x + bt.x > 0.f)) d.x = 0.f;
      if (!(xh.y * gm.y + bt.y > 0.f)) d.y = 0.f;
      if (!(xh.z * gm.z + bt.z > 0.f)) d.z = 0.f;
      if (!(xh.w * gm.w + bt.w > 0.f)) d.w = 0.f;
    }
    float4 o;
    o.x = rstd * (d.x * gm.x - m1 - xh.x * m2);
    o.y = rstd * (d.y * gm.y - m1 - xh.y * m2);
    o.z = rstd * (d.z * gm.z - m1 - xh.z * m2);
    o.w = rstd * (d.w * gm.w - m1 - xh.w * m2);
    *reinterpret_cast<float4*>(dx + off) = o;
  }
}

static bool gn_shape_ok(int C, int G) {
  if (G <= 0 || C % G) return false;
  const int cpg = C / G;
  // a group is 4, 8, 16 or 32 channels (1..8 lanes of one warp), C/4 threads per row divide the CTA
  return C % 4 == 0 && (cpg == 4 || cpg == 8 || cpg == 16 || cpg == 32) && (C >> 2) <= kGnThreads &&
         32 % (cpg >> 2) == 0 && (C >> 2) % (cpg >> 2) == 0;
}

}  // namespace mpf

extern "C" {

int mpf_groupnorm_cl_fwd_f32(const float* x, const float* gamma, const float* beta, float eps, int batch,
                             long long HW, int C, int groups, int relu, float* y, float* mean, float* rstd,
                             double* stats_ws, void* stream) {
  using namespace mpf;
  clear_error();
  MPF_REQUIRE(x && gamma && beta && y && mean && rstd && stats_ws, "groupnorm_fwd: null pointer argument");
  MPF_REQUIRE(batch > 0 && HW > 0 && gn_shape_ok(C, groups), "groupnorm_fwd: unsupported shape (C=%d, groups=%d)", C, groups);
  MPF_REQUIRE(aligned16(x) && aligned16(y) && aligned16(gamma) && aligned16(beta), "groupnorm_fwd: 16-byte alignment");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int cpg = C / groups;
  MPF_CUDA_OK(cudaMemsetAsync(stats_ws, 0, sizeof(double) * 2 * batch * groups, st));
  int ctas_x = static_cast<int>((148ll * 8 + batch - 1) / batch);
  long long rows_per_cta = (HW + ctas_x - 1) / ctas_x;
  if (rows_per_cta < 64) rows_per_cta = 64;
  ctas_x = static_cast<int>((HW + rows_per_cta - 1) / rows_per_cta);
  groupnorm_stats_kernel<<<dim3(ctas_x, batch), kGnThreads, 0, st>>>(x, HW, C, cpg, static_cast<int>(rows_per_cta),
                                                                     stats_ws);
  long long total = HW * (C / 4);
  int ax = static_cast<int>((total + kGnThreads - 1) / kGnThreads);
  const int cap = (148 * 8 + batch - 1) / batch;
  if (ax > cap) ax = cap;
  groupnorm_apply_kernel<<<dim3(ax, batch), kGnThreads, 0, st>>>(x, gamma, beta, stats_ws, HW, C, cpg, eps, relu, y,
                                                                 mean, rstd);
  count_launch(2);
  return finish_launch("groupnorm_fwd");
}

int mpf_groupnorm_cl_bwd_f32(const float* dy, const float* x, const float* gamma, const float* beta,
                             const float* mean, const float* rstd, int batch, long long HW, int C, int groups,
                             int relu, float* dx, float* dgamma_dbeta, double* stats_ws, void* stream) {
  using namespace mpf;
  clear_error();
  MPF_REQUIRE(dy && x && gamma && beta && mean && rstd && dx && dgamma_dbeta && stats_ws,
              "groupnorm_bwd: null pointer argument");
  MPF_REQUIRE(batch > 0 && HW > 0 && gn_shape_ok(C, groups), "groupnorm_bwd: unsupported shape (C=%d, groups=%d)", C, groups);
  MPF_REQUIRE(aligned16(x) && aligned16(dy) && aligned16(dx) && aligned16(gamma) && aligned16(beta),
              "groupnorm_bwd: 16-byte alignment");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int cpg = C / groups;
  MPF_CUDA_OK(cudaMemsetAsync(stats_ws, 0, sizeof(double) * 2 * batch * groups, st));
  MPF_CUDA_OK(cudaMemsetAsync(dgamma_dbeta, 0, sizeof(float) * 2 * C, st));
  int ctas_x = static_cast<int>((148ll * 8 + batch - 1) / batch);
  long long rows_per_cta = (HW + ctas_x - 1) / ctas_x;
  if (rows_per_cta < 64) rows_per_cta = 64;
  ctas_x = static_cast<int>((HW + rows_per_cta - 1) / rows_per_cta);
  groupnorm_bwd_stats_kernel<<<dim3(ctas_x, batch), kGnThreads, 0, st>>>(
      dy, x, gamma, beta, mean, rstd, HW, C, cpg, relu, static_cast<int>(rows_per_cta), stats_ws, dgamma_dbeta);
  long long total = HW * (C / 4);
  int ax = static_cast<int>((total + kGnThreads - 1) / kGnThreads);
  const int cap = (148 * 8 + batch - 1) / batch;
  if (ax > cap) ax = cap;
  groupnorm_bwd_apply_kernel<<<dim3(ax, batch), kGnThreads, 0, st>>>(dy, x, gamma, beta, mean, rstd, stats_ws, HW, C,
                                                                     cpg, relu, dx);
  count_launch(2);
  return finish_launch("groupnorm_bwd");
}

}  // extern "C"

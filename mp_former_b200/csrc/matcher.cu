// Hungarian matching on the device (the step right after the decoder's prediction heads).
//
//   ref: mask2former/modeling/matcher.py:96-157 (HungarianMatcher.memory_efficient_forward), :15-62 (the two cost
//        terms), detectron2 point_sample (= F.grid_sample at 2*c-1, bilinear, zeros padding, align_corners=False),
//        scipy.optimize.linear_sum_assignment (matcher.py:151).
//
// The reference runs, per image and per prediction head, two grid_samples, two softplus maps, three einsums over
// the 12544 sampled points, a device->host copy of the cost matrix (a stream sync) and a CPU solve: 16 x 10 syncs per
// step at the bench geometry.  Here the whole batch is three launches with no host round trip:
//
//   match_cost_partial_kernel   grid (point splits, query tiles x target tiles, images).  A CTA owns 64 queries x
//                               <=32 targets of one image and walks its share of the points in chunks of 128:
//                               phase 1 samples the prediction logits (4 gathers per query and point; the corner
//                               offsets and weights of a point are computed once per thread and reused for all its
//                               queries) and the GT masks (uint8 or float) and leaves softplus(-x), softplus(x),
//                               sigmoid(x) and t in shared memory, point-major with an odd row pitch so both the
//                               phase-1 stores (lanes = points) and the phase-2 loads (lanes = queries) are
//                               conflict-free; phase 2 is a 2x4 register tile per thread of the three point sums
//                               sum pos*t, sum neg*(1-t), sum sig*t.  Partial sums per point split go to a workspace.
//   match_cost_finish_kernel    one thread per (query, target): adds the splits in a fixed order (deterministic, no
//                               atomics), forms the cross-entropy and dice costs, the class cost -softmax(logits)[label]
//                               and the weighted total, written as one row-major [Q, n_b] matrix per image.
//   lsap_kernel                 one CTA per image: shortest-augmenting-path LSAP in float64 with scipy's scan order and
//                               tie rule (restated and pinned in oracle/lsap_oracle.py), the column scan parallel over
//                               the CTA with an order-independent key, so the INDICES equal scipy's on equal costs.
//
// Bound: the gathers (Q*P*2 32-byte sectors per image and head out of a 26 MB logit map that was just written and is
// L2-resident); arithmetic is ~1 GFLOP per head.
#include "mpf_common.cuh"
#include "point_sample.cuh"
#include "sm100_ptx.cuh"

#include <math_constants.h>

namespace mpf {

constexpr int MC_PT = 64;        // points per chunk (58 KB of shared memory per CTA: three CTAs per SM)
constexpr int MC_QT = 64;        // queries per CTA
constexpr int MC_NT = 32;        // targets per CTA
constexpr int MC_THREADS = 256;
constexpr int MC_QS = MC_QT + 1; // shared-memory row pitches (odd: conflict-free both ways)
constexpr int MC_TS = MC_NT + 1;
constexpr size_t MC_SMEM = (3 * MC_PT * MC_QS + MC_PT * MC_TS) * sizeof(float);

struct MatchCostArgs {
  const float* pred_masks;        // [B, Q, H, W] through the two strides below
  long long masks_img_stride, masks_q_stride;
  const void* const* tgt_mask_ptrs;  // [B] device pointers, each [n_b, Hg, Wg] contiguous (uint8 0/1 or float)
  const int* tgt_offsets;         // [B + 1] prefix sums of n_b
  const float* point_coords;      // [B, P, 2] (x, y) in [0, 1]
  const float* sampled;           // optional [B, Q, P]: the prediction logits already sampled at the points
  float* part3;                   // [S, Q, ntot, 3]
  float* part_sig;                // [S, B, Q]
  float* part_t;                  // [S, ntot]
  int B, Q, P, H, W, Hg, Wg, ntot, qtiles, nchunks;
};

template <typename TM>
__global__ void __launch_bounds__(MC_THREADS, 3) match_cost_partial_kernel(const MatchCostArgs a) {
  extern __shared__ float smem[];
  float* s_pos = smem;
  float* s_neg = s_pos + MC_PT * MC_QS;
  float* s_sig = s_neg + MC_PT * MC_QS;
  float* s_t = s_sig + MC_PT * MC_QS;

  const int b = blockIdx.z;
  const int n0 = a.tgt_offsets[b];
  const int nb = a.tgt_offsets[b + 1] - n0;
  const int qt = blockIdx.y % a.qtiles, tt = blockIdx.y / a.qtiles;
  const int j0 = tt * MC_NT;
  if (j0 >= nb) return;                         // CTA-uniform (also images without targets)
  const int q0 = qt * MC_QT;
  const int nq = min(MC_QT, a.Q - q0), nj = min(MC_NT, nb - j0);
  const int S = gridDim.x, s = blockIdx.x;
  const int tid = threadIdx.x;

  const float* pm = a.pred_masks + b * a.masks_img_stride + q0 * a.masks_q_stride;
  const TM* tm = static_cast<const TM*>(a.tgt_mask_ptrs[b]) + static_cast<long long>(j0) * a.Hg * a.Wg;
  const float* pc = a.point_coords + static_cast<long long>(b) * a.P * 2;

  // phase 1 mapping: a thread owns one point of the chunk and every fourth query / target
  const int p1 = tid & (MC_PT - 1), part = tid / MC_PT;
  constexpr int PARTS = MC_THREADS / MC_PT;
  // phase 2 mapping: lanes = queries (lane, lane + 32), warp = four targets
  const int lane = tid & 31, wj = tid >> 5;

  float accP[2][4], accN[2][4], accD[2][4], sumS[2] = {0.f, 0.f}, sumT[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int k = 0; k < 4; ++k) accP[i][k] = accN[i][k] = accD[i][k] = 0.f;

  for (int c = s; c < a.nchunks; c += S) {
    const int pidx = c * MC_PT + p1;
    const bool valid = pidx < a.P;
    float cx = 0.f, cy = 0.f;
    if (valid) {
      const float2 xy = __ldg(reinterpret_cast<const float2*>(pc) + pidx);
      cx = xy.x;
      cy = xy.y;
    }
    {
      // The kernel is bound by the latency of these random gathers (ncu, profiles/r2b_ncu_match_cost.txt: 8 warps
      // per SM, long-scoreboard stalls 4.3 per issue, issue slots 20 % busy): four queries' sixteen corner loads are
      // put in flight before the first transcendental, and three CTAs share an SM.
      const Corners cp = point_corners(cx, cy, a.H, a.W);
#pragma unroll 1
      for (int q4 = part; q4 < MC_QT; q4 += 4 * PARTS) {
        float x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int qq = q4 + u * PARTS;
          if (!(valid && qq < nq)) x[u] = 0.f;
          else if (a.sampled != nullptr)
            x[u] = __ldg(a.sampled + (static_cast<long long>(b) * a.Q + q0 + qq) * a.P + pidx);   // lanes = points: coalesced
          else x[u] = sample_map(pm + qq * a.masks_q_stride, cp);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int qq = q4 + u * PARTS;
          float pos = 0.f, neg = 0.f, sig = 0.f;
          if (valid && qq < nq) {
            // BCE-with-logits against 1 and against 0 (matcher.py:52-57): max(-+x, 0) + log1p(exp(-|x|))
            const float sp = log1pf(expf(-fabsf(x[u])));
            pos = fmaxf(-x[u], 0.f) + sp;
            neg = fmaxf(x[u], 0.f) + sp;
            sig = 1.0f / (1.0f + expf(-x[u]));
          }
          if (qq < MC_QT) {
            s_pos[p1 * MC_QS + qq] = pos;
            s_neg[p1 * MC_QS + qq] = neg;
            s_sig[p1 * MC_QS + qq] = sig;
          }
        }
      }
    }
    {
      const Corners ct = point_corners(cx, cy, a.Hg, a.Wg);
      float t[MC_NT / PARTS];
#pragma unroll
      for (int u = 0; u < MC_NT / PARTS; ++u) {
        const int jj = part + u * PARTS;
        t[u] = (valid && jj < nj) ? sample_map(tm + static_cast<long long>(jj) * a.Hg * a.Wg, ct) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < MC_NT / PARTS; ++u) s_t[p1 * MC_TS + part + u * PARTS] = t[u];
    }
    __syncthreads();
    // a warp owns four targets: with fewer targets in the tile (COCO averages 7 instances per image against the
    // tile's 32) the warps past the last one have nothing to add and leave the issue slots to the others
    if (wj * 4 < nj) {
#pragma unroll 4
    for (int p = 0; p < MC_PT; ++p) {
      const float ps[2] = {s_pos[p * MC_QS + lane], s_pos[p * MC_QS + lane + 32]};
      const float ng[2] = {s_neg[p * MC_QS + lane], s_neg[p * MC_QS + lane + 32]};
      const float sg[2] = {s_sig[p * MC_QS + lane], s_sig[p * MC_QS + lane + 32]};
      float t[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) t[k] = s_t[p * MC_TS + wj * 4 + k];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        sumS[i] += sg[i];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          accP[i][k] += ps[i] * t[k];
          accN[i][k] += ng[i] * (1.0f - t[k]);
          accD[i][k] += sg[i] * t[k];
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) sumT[k] += t[k];
    }
    }
    __syncthreads();
  }

  // partial sums of this point split
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int ql = lane + 32 * i;
    if (ql >= nq) continue;
    const int q = q0 + ql;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int jl = wj * 4 + k;
      if (jl >= nj) continue;
      float* dst = a.part3 + ((static_cast<long long>(s) * a.Q + q) * a.ntot + n0 + j0 + jl) * 3;
      dst[0] = accP[i][k];
      dst[1] = accN[i][k];
      dst[2] = accD[i][k];
    }
    if (tt == 0 && wj == 0) a.part_sig[(static_cast<long long>(s) * a.B + b) * a.Q + q] = sumS[i];
  }
  if (qt == 0 && lane == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int jl = wj * 4 + k;
      if (jl < nj) a.part_t[static_cast<long long>(s) * a.ntot + n0 + j0 + jl] = sumT[k];
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Streaming sampler for the matcher's SHARED points: all Q maps of an image are sampled at the same P points
// (matcher.py:120-131), and 12544 points with their 2x2 footprints touch most of a 256 x 256 map.  Gathering them
// costs a 32-byte DRAM sector per corner once the logits of a head (419 MB at 16 images) exceed the L2 -- 661 MB of
// sector traffic and, worse, the latency of dependent random loads (0.45 ms per head).  Here a CTA owns one map,
// streams it through shared memory band by band with coalesced 16-byte loads (each byte read once: the map's own
// size) and evaluates the points whose footprint starts in the band from shared memory.  The points arrive sorted
// row-major (HungarianMatcher.row_major_order) with the first point of every band in band_lo; a point whose
// footprint is not inside the staged band after all (a caller's own order, rounding at a band edge) is sampled from
// global memory instead, so the result never depends on the order.
//   grid (Q, B), 512 threads, shared memory (rows + 1) * W floats.   out [B, Q, P]
// ------------------------------------------------------------------------------------------------------------------
constexpr int kSspThreads = 512;

__global__ void __launch_bounds__(kSspThreads)
sample_shared_points_kernel(const float* __restrict__ maps, long long img_stride, long long q_stride, int H, int W,
                            const float* __restrict__ coords, const int* __restrict__ band_lo, int n_bands, int rows,
                            int P, int Q, float* __restrict__ out) {
  extern __shared__ __align__(16) float s_band[];   // rows [r0, r0 + rows] of the map (one halo row below)
  __shared__ uint64_t s_bar;
  const int q = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const float* map = maps + b * img_stride + q * q_stride;
  const float2* pc = reinterpret_cast<const float2*>(coords) + static_cast<long long>(b) * P;
  float* dst = out + (static_cast<long long>(b) * Q + q) * P;
  const int* lo = band_lo + b * (n_bands + 1);
  if (tid == 0) {
    ptx::mbar_init(&s_bar, 1);
    ptx::fence_mbar_init();
  }
  __syncthreads();
  uint32_t phase = 0;
  for (int k = 0; k < n_bands; ++k) {
    const int p0 = lo[k], p1 = lo[k + 1];
    if (p1 <= p0) continue;                       // CTA-uniform
    const int r0 = k * rows;
    const int nr = min(rows + 1, H - r0);         // rows present in the map (rows past its end are never addressed:
                                                  // point_corners marks those corners invalid)
    if (tid == 0) {
      // the band is one contiguous piece of the map: a single bulk copy (TMA engine), completion on the mbarrier
      const uint32_t bytes = static_cast<uint32_t>(nr) * W * 4u;
      ptx::mbar_arrive_expect_tx(&s_bar, bytes);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       ptx::smem_u32(s_band)),
                   "l"(map + static_cast<long long>(r0) * W), "r"(bytes), "r"(ptx::smem_u32(&s_bar))
                   : "memory");
    }
    ptx::mbar_wait(&s_bar, phase);
    phase ^= 1u;
    const int base = r0 * W, limit = nr * W;
    // two points per thread and iteration: both coordinate loads are in flight before either is used (the loop is
    // otherwise one dependent L2 round trip per point: 282 us per head instead of the ~70 us the map read takes)
    for (int p = p0 + tid; p < p1; p += 2 * kSspThreads) {
      const int pb = p + kSspThreads;
      const float2 xa = __ldg(pc + p);
      const float2 xb = pb < p1 ? __ldg(pc + pb) : xa;
      const Corners ca = point_corners(xa.x, xa.y, H, W), cb = point_corners(xb.x, xb.y, H, W);
      float va[4], vb[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int oa = ca.o[j] - base, ob = cb.o[j] - base;
        va[j] = ca.o[j] < 0 ? 0.f : ((oa >= 0 && oa < limit) ? s_band[oa] : __ldg(map + ca.o[j]));
        vb[j] = cb.o[j] < 0 ? 0.f : ((ob >= 0 && ob < limit) ? s_band[ob] : __ldg(map + cb.o[j]));
      }
      float acca = 0.f, accb = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (ca.o[j] >= 0) acca += va[j] * ca.w[j];
        if (cb.o[j] >= 0) accb += vb[j] * cb.w[j];
      }
      dst[p] = acca;
      if (pb < p1) dst[pb] = accb;
    }
    __syncthreads();                              // all reads of the band done before the next copy lands in it
  }
}

// One WARP per (query, target) cell: the lanes share the point splits (up to 111 of them when a couple of images
// leave the partial kernel with many splits: a single thread walking them took 54 us per launch at two images) and
// the K + 1 classes of the soft-max; fixed lane-strided order + xor-shuffle tree -> deterministic.
__global__ void __launch_bounds__(256)
match_cost_finish_kernel(const float* __restrict__ part3, const float* __restrict__ part_sig,
                         const float* __restrict__ part_t, const float* __restrict__ logits,
                         long long logits_img_stride, long long logits_q_stride, int K1,
                         const long long* __restrict__ labels, const int* __restrict__ offsets, int B, int Q, int ntot,
                         int S, int P, float w_class, float w_mask, float w_dice, float* __restrict__ cost) {
  const long long idx = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (idx >= static_cast<long long>(Q) * ntot) return;          // warp-uniform
  const int q = static_cast<int>(idx / ntot), jg = static_cast<int>(idx - static_cast<long long>(q) * ntot);
  int b = 0;
  while (b + 1 < B && __ldg(offsets + b + 1) <= jg) ++b;
  const int n0 = __ldg(offsets + b), nb = __ldg(offsets + b + 1) - n0;
  float sp = 0.f, sn = 0.f, sd = 0.f, ss = 0.f, st = 0.f;
  for (int s = lane; s < S; s += 32) {
    const float* p3 = part3 + ((static_cast<long long>(s) * Q + q) * ntot + jg) * 3;
    sp += p3[0];
    sn += p3[1];
    sd += p3[2];
    ss += part_sig[(static_cast<long long>(s) * B + b) * Q + q];
    st += part_t[static_cast<long long>(s) * ntot + jg];
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    sp += __shfl_xor_sync(0xffffffffu, sp, d);
    sn += __shfl_xor_sync(0xffffffffu, sn, d);
    sd += __shfl_xor_sync(0xffffffffu, sd, d);
    ss += __shfl_xor_sync(0xffffffffu, ss, d);
    st += __shfl_xor_sync(0xffffffffu, st, d);
  }
  const float c_mask = (sp + sn) / static_cast<float>(P);                 // matcher.py:59-61
  const float c_dice = 1.0f - (2.0f * sd + 1.0f) / (ss + st + 1.0f);      // matcher.py:26-29
  // class cost: -softmax(logits[b, q])[label]   (matcher.py:105,111)
  const float* row = logits + b * logits_img_stride + q * logits_q_stride;
  const long long lab = __ldg(labels + jg);
  float c_class;
  if (lab < 0 || lab >= K1) {
    c_class = CUDART_NAN_F;                                               // surfaces as "invalid" in the solver
  } else {
    float m = -CUDART_INF_F;
    for (int k = lane; k < K1; k += 32) m = fmaxf(m, __ldg(row + k));
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
    float den = 0.f;
    for (int k = lane; k < K1; k += 32) den += expf(__ldg(row + k) - m);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) den += __shfl_xor_sync(0xffffffffu, den, d);
    c_class = -(expf(__ldg(row + lab) - m) / den);
  }
  if (lane == 0)
    cost[static_cast<long long>(Q) * n0 + static_cast<long long>(q) * nb + (jg - n0)] =
        (w_mask * c_mask + w_class * c_class) + w_dice * c_dice;          // matcher.py:142-146
}

// ------------------------------------------------------------------------------------------------------------------
// Rectangular LSAP, one CTA per image.  Algorithm and tie rules: oracle/lsap_oracle.py (scipy's, restated).
// ------------------------------------------------------------------------------------------------------------------
constexpr int LSAP_THREADS = 128;

struct ScanKey {       // minimum of (value, cls, ord) == the column scipy's sequential scan selects
  double value;
  int cls, ord, it;
};

__device__ __forceinline__ bool key_less(const ScanKey& x, const ScanKey& y) {
  if (x.value != y.value) return x.value < y.value;
  if (x.cls != y.cls) return x.cls < y.cls;
  return x.ord < y.ord;
}

__device__ __forceinline__ ScanKey key_shfl_xor(const ScanKey& k, int m) {
  ScanKey r;
  r.value = __shfl_xor_sync(0xffffffffu, k.value, m);
  r.cls = __shfl_xor_sync(0xffffffffu, k.cls, m);
  r.ord = __shfl_xor_sync(0xffffffffu, k.ord, m);
  r.it = __shfl_xor_sync(0xffffffffu, k.it, m);
  return r;
}

__global__ void __launch_bounds__(LSAP_THREADS)
lsap_kernel(const float* __restrict__ cost, const int* __restrict__ offsets, int Q, int dim, int cache_cost,
            long long* __restrict__ out_q, long long* __restrict__ out_t, int* __restrict__ status) {
  extern __shared__ double lsap_smem[];
  double* u = lsap_smem;            // [dim] rows
  double* v = u + dim;              // [dim] cols
  double* spc = v + dim;            // [dim] shortest path costs
  int* path = reinterpret_cast<int*>(spc + dim);
  int* col4row = path + dim;
  int* row4col = col4row + dim;
  int* remaining = row4col + dim;
  int* SR = remaining + dim;
  int* SC = SR + dim;
  float* c_smem = reinterpret_cast<float*>(SC + dim);          // [Q * n] when cache_cost
  __shared__ ScanKey s_key[LSAP_THREADS / 32];
  __shared__ int s_i, s_sink, s_num_remaining, s_fail;
  __shared__ double s_min;

  const int b = blockIdx.x, tid = threadIdx.x;
  const int n0 = offsets[b], n = offsets[b + 1] - n0;
  // matches of the images before this one
  long long out0 = 0;
  for (int bb = 0; bb < b; ++bb) out0 += min(Q, offsets[bb + 1] - offsets[bb]);
  if (n <= 0) return;
  const float* Cg = cost + static_cast<long long>(Q) * n0;   // [Q, n] row-major
  // The search below reads ONE cost entry per candidate column and step, each on the critical path of a serial
  // algorithm: from global memory that is an L2 round trip per step (the solve of a 100 x 20 matrix took 150 us,
  // scipy on a host core ~20 us).  The matrix (<= 40 KB for 100 queries x 100 targets) is therefore copied to shared
  // memory by the validity pre-scan, which reads every entry anyway.
  const float* C = cache_cost ? c_smem : Cg;
  const bool transposed = n < Q;                              // scipy: solve the transpose of a tall matrix
  const int nr = transposed ? n : Q, nc = transposed ? Q : n;
  const long long si = transposed ? 1 : n, sj = transposed ? n : 1;   // cost(i, j) = C[i * si + j * sj]

  for (int k = tid; k < nr; k += LSAP_THREADS) { u[k] = 0.0; col4row[k] = -1; }
  for (int k = tid; k < nc; k += LSAP_THREADS) { v[k] = 0.0; row4col[k] = -1; path[k] = -1; }
  if (tid == 0) s_fail = 0;
  __syncthreads();
  // scipy's linear_sum_assignment rejects a matrix with ANY NaN or -inf entry ("matrix contains invalid numeric
  // entries") before it solves; the augmenting-path search alone would only notice a row without a finite column.
  {
    int bad = 0;
    const long long total = static_cast<long long>(Q) * n;
    for (long long k = tid; k < total; k += LSAP_THREADS) {
      const float c = __ldg(Cg + k);
      if (cache_cost) c_smem[k] = c;
      bad |= (c != c) || (c == -CUDART_INF_F);
    }
    if (bad) s_fail = 1;
  }
  __syncthreads();

  for (int cur = 0; cur < nr && !s_fail; ++cur) {
    for (int k = tid; k < nr; k += LSAP_THREADS) SR[k] = 0;
    for (int k = tid; k < nc; k += LSAP_THREADS) { SC[k] = 0; spc[k] = CUDART_INF; remaining[k] = nc - k - 1; }
    if (tid == 0) { s_i = cur; s_sink = -1; s_num_remaining = nc; s_min = 0.0; }
    __syncthreads();
    while (true) {
      const int i = s_i, num_remaining = s_num_remaining;
      const double min_val = s_min, ui = u[i];
      ScanKey best;
      best.value = CUDART_INF; best.cls = 2; best.ord = 0; best.it = -1;
      for (int it = tid; it < num_remaining; it += LSAP_THREADS) {
        const int j = remaining[it];
        const double r = min_val + static_cast<double>(C[i * si + j * sj]) - ui - v[j];
        double d = spc[j];
        if (r < d) { path[j] = i; spc[j] = r; d = r; }
        if (d < CUDART_INF) {
          ScanKey k;
          k.value = d; k.it = it;
          if (row4col[j] == -1) { k.cls = 0; k.ord = -it; } else { k.cls = 1; k.ord = it; }
          if (best.it < 0 || key_less(k, best)) best = k;
        }
      }
#pragma unroll
      for (int m = 16; m > 0; m >>= 1) {
        const ScanKey o = key_shfl_xor(best, m);
        if (o.it >= 0 && (best.it < 0 || key_less(o, best))) best = o;
      }
      if ((tid & 31) == 0) s_key[tid >> 5] = best;
      __syncthreads();
      if (tid == 0) {
        ScanKey w = s_key[0];
        for (int k = 1; k < LSAP_THREADS / 32; ++k) {
          const ScanKey o = s_key[k];
          if (o.it >= 0 && (w.it < 0 || key_less(o, w))) w = o;
        }
        SR[i] = 1;
        if (w.it < 0) {                    // infeasible / NaN costs (scipy raises ValueError)
          s_fail = 1;
        } else {
          s_min = w.value;
          const int j = remaining[w.it];
          if (row4col[j] == -1) s_sink = j; else s_i = row4col[j];
          SC[j] = 1;
          remaining[w.it] = remaining[num_remaining - 1];
          s_num_remaining = num_remaining - 1;
        }
      }
      __syncthreads();
      if (s_fail || s_sink >= 0) break;
    }
    if (s_fail) break;
    // dual variables
    const double min_val = s_min;
    for (int k = tid; k < nr; k += LSAP_THREADS)
      if (SR[k] && k != cur) u[k] += min_val - spc[col4row[k]];
    for (int k = tid; k < nc; k += LSAP_THREADS)
      if (SC[k]) v[k] -= min_val - spc[k];
    __syncthreads();
    if (tid == 0) {
      u[cur] += min_val;
      int j = s_sink;
      while (true) {                         // augment along the alternating path
        const int i = path[j];
        row4col[j] = i;
        const int t = col4row[i];
        col4row[i] = j;
        j = t;
        if (i == cur) break;
      }
    }
    __syncthreads();
  }

  if (tid == 0) {
    const int m = min(Q, n);
    if (s_fail) {
      // failure is reported through ``status`` (the reference: scipy's ValueError); the pairs written here only have
      // to be IN RANGE, so that a caller that consumes them before it looks at the status (the criterion gathers and
      // scatters through them without a host round trip) cannot leave its buffers
      for (int k = 0; k < m; ++k) { out_q[out0 + k] = k; out_t[out0 + k] = k; }
      atomicMax(status, b + 1);
    } else if (transposed) {                 // rows = targets, cols = queries: emit pairs in query order
      int cnt = 0;
      for (int q = 0; q < Q; ++q)
        if (row4col[q] >= 0) { out_q[out0 + cnt] = q; out_t[out0 + cnt] = row4col[q]; ++cnt; }
    } else {
      for (int q = 0; q < Q; ++q) { out_q[out0 + q] = q; out_t[out0 + q] = col4row[q]; }
    }
  }
}

static int match_split(int B, int qtiles, int ttiles, int nchunks) {
  // as many point splits as fill ONE wave of 148 SMs x 3 resident CTAs without spilling into a second, mostly
  // empty one (round 1 launched 320 one-per-SM CTAs = 3 rounds for 2.16 rounds of work); at most one chunk per split
  const long long per_split = static_cast<long long>(B) * qtiles * ttiles;
  long long s = (3 * 148) / per_split;
  if (s < 1) s = 1;
  if (s > nchunks) s = nchunks;
  return static_cast<int>(s);
}

}  // namespace mpf

extern "C" {

long long mpf_match_cost_workspace_bytes(int batch, int num_queries, int total_targets, int max_targets,
                                         int num_points) {
  using namespace mpf;
  if (batch <= 0 || num_queries <= 0 || total_targets < 0 || max_targets < 0 || num_points <= 0) return -1;
  const int qtiles = (num_queries + MC_QT - 1) / MC_QT, ttiles = max(1, (max_targets + MC_NT - 1) / MC_NT);
  const int nchunks = (num_points + MC_PT - 1) / MC_PT;
  const long long S = match_split(batch, qtiles, ttiles, nchunks);
  const long long floats = S * num_queries * total_targets * 3 + S * batch * num_queries + S * total_targets;
  return (floats + 4) * static_cast<long long>(sizeof(float));
}

int mpf_sample_shared_points_f32(const float* maps, long long img_stride, long long q_stride, int H, int W,
                                 const float* point_coords, const int32_t* band_lo, int n_bands, int band_rows,
                                 int batch, int num_queries, int num_points, float* out, void* stream) {
  using namespace mpf;
  clear_error();
  MPF_REQUIRE(batch > 0 && num_queries > 0 && num_points > 0 && H > 0 && W > 0, "sample_shared_points: sizes must be positive");
  MPF_REQUIRE(maps && point_coords && band_lo && out, "sample_shared_points: null pointer argument");
  MPF_REQUIRE(W % 4 == 0 && aligned16(maps) && img_stride % 4 == 0 && q_stride % 4 == 0,
              "sample_shared_points: W and the strides must be multiples of 4 and the maps 16-byte aligned");
  MPF_REQUIRE(n_bands > 0 && band_rows > 0 && static_cast<long long>(n_bands) * band_rows >= H,
              "sample_shared_points: the bands must cover the map (%d bands of %d rows, H = %d)", n_bands, band_rows, H);
  MPF_REQUIRE(batch <= 65535, "sample_shared_points: batch > 65535");
  MPF_REQUIRE((reinterpret_cast<uintptr_t>(point_coords) & 7u) == 0, "sample_shared_points: point_coords must be 8-byte aligned");
  const size_t smem = static_cast<size_t>(band_rows + 1) * W * sizeof(float);
  MPF_REQUIRE(smem <= 200 * 1024, "sample_shared_points: a band of %d rows x %d columns exceeds the shared memory", band_rows + 1, W);
  static unsigned long long seen = 0;
  if (first_use_on_this_device(seen))
    MPF_CUDA_OK(cudaFuncSetAttribute(sample_shared_points_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const dim3 grid(num_queries, batch);
  sample_shared_points_kernel<<<grid, kSspThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      maps, img_stride, q_stride, H, W, point_coords, band_lo, n_bands, band_rows, num_points, num_queries, out);
  count_launch();
  return finish_launch("sample_shared_points");
}

static int match_cost_impl(const float* pred_logits, long long logits_img_stride, long long logits_q_stride,
                           int num_classes_p1, const float* pred_masks, long long masks_img_stride,
                           long long masks_q_stride, int H, int W, const void* const* tgt_mask_ptrs, int tgt_is_f32,
                           int Hg, int Wg, const int64_t* tgt_labels, const int32_t* tgt_offsets, int total_targets,
                           int max_targets, const float* point_coords, int batch, int num_queries, int num_points,
                           float cost_class, float cost_mask, float cost_dice, void* workspace,
                           long long workspace_bytes, float* cost, const float* sampled, void* stream) {
  using namespace mpf;
  clear_error();
  MPF_REQUIRE(batch > 0 && num_queries > 0 && num_points > 0 && H > 0 && W > 0 && Hg > 0 && Wg > 0 &&
                  num_classes_p1 > 0,
              "match_cost: sizes must be positive");
  MPF_REQUIRE(total_targets >= 0 && max_targets >= 0 && max_targets <= total_targets,
              "match_cost: bad target counts (total %d, max %d)", total_targets, max_targets);
  if (total_targets == 0) return MPF_OK;       // nothing to match
  MPF_REQUIRE(pred_logits && (pred_masks || sampled) && tgt_mask_ptrs && tgt_labels && tgt_offsets && point_coords &&
                  workspace && cost,
              "match_cost: null pointer argument");
  MPF_REQUIRE(static_cast<long long>(H) * W < (1ll << 31) && static_cast<long long>(Hg) * Wg < (1ll << 31),
              "match_cost: map too large");
  MPF_REQUIRE(batch <= 65535, "match_cost: batch > 65535");
  const long long need = mpf_match_cost_workspace_bytes(batch, num_queries, total_targets, max_targets, num_points);
  MPF_REQUIRE(workspace_bytes >= need, "match_cost: workspace of %lld bytes, need %lld", workspace_bytes, need);
  MPF_REQUIRE((reinterpret_cast<uintptr_t>(point_coords) & 7u) == 0, "match_cost: point_coords must be 8-byte aligned");

  const int qtiles = (num_queries + MC_QT - 1) / MC_QT, ttiles = max(1, (max_targets + MC_NT - 1) / MC_NT);
  const int nchunks = (num_points + MC_PT - 1) / MC_PT;
  const int S = match_split(batch, qtiles, ttiles, nchunks);
  MPF_REQUIRE(static_cast<long long>(qtiles) * ttiles <= 65535, "match_cost: too many query x target tiles");

  MatchCostArgs a;
  a.pred_masks = pred_masks;
  a.masks_img_stride = masks_img_stride;
  a.masks_q_stride = masks_q_stride;
  a.tgt_mask_ptrs = tgt_mask_ptrs;
  a.tgt_offsets = tgt_offsets;
  a.point_coords = point_coords;
  a.sampled = sampled;
  float* ws = static_cast<float*>(workspace);
  a.part3 = ws;
  a.part_sig = a.part3 + static_cast<long long>(S) * num_queries * total_targets * 3;
  a.part_t = a.part_sig + static_cast<long long>(S) * batch * num_queries;
  a.B = batch; a.Q = num_queries; a.P = num_points; a.H = H; a.W = W; a.Hg = Hg; a.Wg = Wg;
  a.ntot = total_targets; a.qtiles = qtiles; a.nchunks = nchunks;

  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const dim3 grid(S, qtiles * ttiles, batch);
  static unsigned long long seen_u8 = 0, seen_f32 = 0;
  if (tgt_is_f32) {
    if (first_use_on_this_device(seen_f32))
      MPF_CUDA_OK(cudaFuncSetAttribute(match_cost_partial_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(MC_SMEM)));
    match_cost_partial_kernel<float><<<grid, MC_THREADS, MC_SMEM, st>>>(a);
  } else {
    if (first_use_on_this_device(seen_u8))
      MPF_CUDA_OK(cudaFuncSetAttribute(match_cost_partial_kernel<uint8_t>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(MC_SMEM)));
    match_cost_partial_kernel<uint8_t><<<grid, MC_THREADS, MC_SMEM, st>>>(a);
  }
  count_launch();
  int rc = finish_launch("match_cost (partial sums)");
  if (rc != MPF_OK) return rc;

  const long long cells = static_cast<long long>(num_queries) * total_targets;
  const long long blocks = (cells + 7) / 8;          // a warp per cell, eight cells per CTA
  MPF_REQUIRE(blocks < (1ll << 31), "match_cost: problem too large");
  match_cost_finish_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(
      a.part3, a.part_sig, a.part_t, pred_logits, logits_img_stride, logits_q_stride, num_classes_p1,
      reinterpret_cast<const long long*>(tgt_labels), tgt_offsets, batch, num_queries, total_targets, S, num_points,
      cost_class, cost_mask, cost_dice, cost);
  count_launch();
  return finish_launch("match_cost (finish)");
}

int mpf_match_cost_f32(const float* pred_logits, long long logits_img_stride, long long logits_q_stride,
                       int num_classes_p1, const float* pred_masks, long long masks_img_stride,
                       long long masks_q_stride, int H, int W, const void* const* tgt_mask_ptrs, int tgt_is_f32,
                       int Hg, int Wg, const int64_t* tgt_labels, const int32_t* tgt_offsets, int total_targets,
                       int max_targets, const float* point_coords, int batch, int num_queries, int num_points,
                       float cost_class, float cost_mask, float cost_dice, void* workspace,
                       long long workspace_bytes, float* cost, void* stream) {
  return match_cost_impl(pred_logits, logits_img_stride, logits_q_stride, num_classes_p1, pred_masks, masks_img_stride,
                         masks_q_stride, H, W, tgt_mask_ptrs, tgt_is_f32, Hg, Wg, tgt_labels, tgt_offsets,
                         total_targets, max_targets, point_coords, batch, num_queries, num_points, cost_class,
                         cost_mask, cost_dice, workspace, workspace_bytes, cost, nullptr, stream);
}

int mpf_match_cost_presampled_f32(const float* pred_logits, long long logits_img_stride, long long logits_q_stride,
                                  int num_classes_p1, const float* sampled, int H, int W,
                                  const void* const* tgt_mask_ptrs, int tgt_is_f32, int Hg, int Wg,
                                  const int64_t* tgt_labels, const int32_t* tgt_offsets, int total_targets,
                                  int max_targets, const float* point_coords, int batch, int num_queries,
                                  int num_points, float cost_class, float cost_mask, float cost_dice, void* workspace,
                                  long long workspace_bytes, float* cost, void* stream) {
  MPF_REQUIRE(sampled != nullptr, "match_cost_presampled: null pointer argument");
  return match_cost_impl(pred_logits, logits_img_stride, logits_q_stride, num_classes_p1, nullptr, 0, 0, H, W,
                         tgt_mask_ptrs, tgt_is_f32, Hg, Wg, tgt_labels, tgt_offsets, total_targets, max_targets,
                         point_coords, batch, num_queries, num_points, cost_class, cost_mask, cost_dice, workspace,
                         workspace_bytes, cost, sampled, stream);
}

int mpf_lsap_f32(const float* cost, const int32_t* tgt_offsets, int batch, int num_queries, int max_targets,
                 int64_t* out_query, int64_t* out_target, int32_t* status, void* stream) {
  using namespace mpf;
  clear_error();
  MPF_REQUIRE(batch > 0 && num_queries > 0 && max_targets >= 0, "lsap: bad sizes");
  if (max_targets == 0) return MPF_OK;
  MPF_REQUIRE(cost && tgt_offsets && out_query && out_target && status, "lsap: null pointer argument");
  const int dim = max(num_queries, max_targets);
  size_t smem = static_cast<size_t>(dim) * (3 * sizeof(double) + 6 * sizeof(int));
  MPF_REQUIRE(smem <= 200 * 1024, "lsap: max(queries, targets) = %d exceeds the shared-memory solver's limit (4266)",
              dim);
  const size_t matrix = static_cast<size_t>(num_queries) * max_targets * sizeof(float);
  const int cache_cost = smem + matrix <= 200 * 1024 ? 1 : 0;       // the largest image's matrix fits next to the state
  if (cache_cost) smem += matrix;
  static unsigned long long seen = 0;
  if (first_use_on_this_device(seen))
    MPF_CUDA_OK(cudaFuncSetAttribute(lsap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  lsap_kernel<<<batch, LSAP_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(
      cost, tgt_offsets, num_queries, dim, cache_cost, reinterpret_cast<long long*>(out_query),
      reinterpret_cast<long long*>(out_target), status);
  count_launch();
  return finish_launch("lsap");
}

}  // extern "C"

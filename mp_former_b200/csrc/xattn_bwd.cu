// Fused masked multi-head cross-attention, backward (two kernels; probabilities are recomputed from the
// saved log-sum-exp, nothing of size [B*heads, Q, HW] ever exists in HBM).
//
//   forward (xattn.cu):  S2 = Qs K^T (log2 domain, Qs = q * log2e/sqrt(d)),  P = exp2(S2 - lse2) (0 where
//                        masked),  O = P V
//   given dO:   dP = dO V^T,   D = rowsum(dO * O),   dS = P * (dP - D)        (gradient wrt the natural logits)
//               dq = dS K / sqrt(d)          dk = dS^T Qs * ln2          dv = P^T dO
//
//   kernel A  masked_xattn_bwd_dq_kernel : CTA = (128-query tile, head, image), walks the keys in tiles of 64;
//             S and dP land in TMEM (double buffered), the softmax warps turn them into dS (hi/lo, shared
//             memory, UMMA K-major layout), dQ += dS K accumulates in TMEM over all key tiles.
//   kernel B  masked_xattn_bwd_dkv_kernel: CTA = (128-key tile, head, image), walks the queries in chunks of 64;
//             computes the TRANSPOSED tiles S^T = K Qs^T and dP^T = V dO^T directly (TMEM lanes = keys), so P^T
//             and dS^T are written row-wise by their owner threads and feed dV += P^T dO, dK += dS^T Qs
//             as ordinary K-major operands -- no transposition through shared memory is ever needed.
// All products are 3xTF32 (hi/lo split operands, fp32 accumulation in TMEM) like the forward.
// ref: the backward of nn.MultiheadAttention in CrossAttentionLayer (decoder :100-112) under autograd.
#include "mpf_common.cuh"
#include "sm100_ptx.cuh"
#include "tmap.cuh"

namespace mpf {

using namespace ptx;

namespace xb {
constexpr int kD = 32;            // head dim
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ float rn(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

// Writes 32 fp32 values of one row as hi/lo TF32 halves into a [rows x 32] K-major SWIZZLE_128B atom.
__device__ __forceinline__ void store_row_atom(uint8_t* atom_hi, uint8_t* atom_lo, int r, const float (&x)[32]) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int sw = r * 128 + ((c ^ (r & 7)) << 4);
    float4 h, l;
    h.x = rn(x[4 * c]); h.y = rn(x[4 * c + 1]); h.z = rn(x[4 * c + 2]); h.w = rn(x[4 * c + 3]);
    l.x = rn(x[4 * c] - h.x); l.y = rn(x[4 * c + 1] - h.y); l.z = rn(x[4 * c + 2] - h.z); l.w = rn(x[4 * c + 3] - h.w);
    *reinterpret_cast<float4*>(atom_hi + sw) = h;
    *reinterpret_cast<float4*>(atom_lo + sw) = l;
  }
}

// Same for 16 values: columns [16 * half, 16 * half + 16) of the row (two threads share a row).
__device__ __forceinline__ void store_half_row_atom(uint8_t* atom_hi, uint8_t* atom_lo, int r, int half,
                                                    const float (&x)[16]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int sw = r * 128 + (((4 * half + c) ^ (r & 7)) << 4);
    float4 h, l;
    h.x = rn(x[4 * c]); h.y = rn(x[4 * c + 1]); h.z = rn(x[4 * c + 2]); h.w = rn(x[4 * c + 3]);
    l.x = rn(x[4 * c] - h.x); l.y = rn(x[4 * c + 1] - h.y); l.z = rn(x[4 * c + 2] - h.z); l.w = rn(x[4 * c + 3] - h.w);
    *reinterpret_cast<float4*>(atom_hi + sw) = h;
    *reinterpret_cast<float4*>(atom_lo + sw) = l;
  }
}

// Same for 8 values: columns [8 * part, 8 * part + 8) of the row (four threads share a row).
__device__ __forceinline__ void store_quarter_row_atom(uint8_t* atom_hi, uint8_t* atom_lo, int r, int part,
                                                       const float (&x)[8]) {
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int sw = r * 128 + (((2 * part + c) ^ (r & 7)) << 4);
    float4 h, l;
    h.x = rn(x[4 * c]); h.y = rn(x[4 * c + 1]); h.z = rn(x[4 * c + 2]); h.w = rn(x[4 * c + 3]);
    l.x = rn(x[4 * c] - h.x); l.y = rn(x[4 * c + 1] - h.y); l.z = rn(x[4 * c + 2] - h.z); l.w = rn(x[4 * c + 3] - h.w);
    *reinterpret_cast<float4*>(atom_hi + sw) = h;
    *reinterpret_cast<float4*>(atom_lo + sw) = l;
  }
}

// three-pass 3xTF32 product step: D (+)= A B^T with split operands
__device__ __forceinline__ void mma3(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                     uint32_t idesc, bool accumulate) {
  mma_tf32_ss(d, smem_desc_sw128_kmajor(a_lo), smem_desc_sw128_kmajor(b_hi), idesc, accumulate ? 1u : 0u);
  mma_tf32_ss(d, smem_desc_sw128_kmajor(a_hi), smem_desc_sw128_kmajor(b_lo), idesc, 1u);
  mma_tf32_ss(d, smem_desc_sw128_kmajor(a_hi), smem_desc_sw128_kmajor(b_hi), idesc, 1u);
}
}  // namespace xb

struct XbwdArgs {
  const uint32_t* bits;     // [B, Qt, words]
  const uint8_t* row_open;  // [B, Qt] or null
  const float* lse2;        // [B, heads, Qt]
  const float* delta;       // [B, heads, Qt]   D = rowsum(dO * O)
  float* dq;                // kernel A: [B, Qt, E]
  float* dk;                // kernel B: [B, HW, E]
  float* dv;                // kernel B: [B, HW, E]
  int B, Qt, HW, E, heads, words;
  float inv_sqrt_d;
  // kernel A key split (see XattnArgs in xattn.cu): partial dQ of key range `split` -> part_dq[split], summed by
  // xattn_sum_splits_kernel in a fixed order (deterministic, unlike atomics)
  int splits, tiles_per_split;
  float* part_dq;           // [splits, B, Qt, E]
};

// ================================================================================================
// kernel A: dQ
// ================================================================================================
namespace xa {
constexpr int kQ = 128, kK = 64;
constexpr int kQBytes = kQ * 32 * 4;            // 16 KiB (one of Q_hi, Q_lo, dO_hi, dO_lo)
constexpr int kKBytes = kK * 32 * 4;            // 8 KiB  (K_hi / K_lo / V_hi / V_lo tile)
constexpr int kKtBytes = 32 * kK * 4;           // 8 KiB  (K^T hi or lo: two [32 x 32] atoms)
constexpr int kStage = 4 * kKBytes + 2 * kKtBytes;   // 48 KiB
constexpr int kAtom = kQ * 32 * 4;              // 16 KiB: [128 rows x 32 keys]
constexpr int kDsBytes = 2 * kAtom;             // 32 KiB (one of dS_hi / dS_lo)
constexpr int kSmem = 4 * kQBytes + 2 * kStage + 2 * kDsBytes + 256 + 1024;
constexpr uint32_t kTmemCols = 512;
constexpr int kTS = 0, kTP = 128, kTQ = 256;    // S: 2x64, dP: 2x64, dQ: 32
constexpr int kThreads = 384;                   // warps 0 TMA, 1 MMA, 2 TMEM allocation, 4-11 softmax
}  // namespace xa

__global__ void __launch_bounds__(xa::kThreads, 1)
masked_xattn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQh, const __grid_constant__ CUtensorMap tmQl,
                           const __grid_constant__ CUtensorMap tmDh, const __grid_constant__ CUtensorMap tmDl,
                           const __grid_constant__ CUtensorMap tmKh, const __grid_constant__ CUtensorMap tmKl,
                           const __grid_constant__ CUtensorMap tmVh, const __grid_constant__ CUtensorMap tmVl,
                           const __grid_constant__ CUtensorMap tmKth, const __grid_constant__ CUtensorMap tmKtl,
                           const XbwdArgs g) {
  using namespace xa;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                         // Q_hi | Q_lo | dO_hi | dO_lo
  uint8_t* sKV = smem + 4 * kQBytes;          // stage: K_hi | K_lo | V_hi | V_lo | Kt_hi | Kt_lo
  uint8_t* sDS = sKV + 2 * kStage;            // dS_hi | dS_lo
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDS + 2 * kDsBytes);
  uint64_t* qdo_full = bars;
  uint64_t* kv_full = bars + 1;    // [2]
  uint64_t* kv_empty = bars + 3;   // [2]
  uint64_t* sdp_full = bars + 5;   // [2]
  uint64_t* sdp_empty = bars + 7;  // [2]
  uint64_t* ds_full = bars + 9;
  uint64_t* ds_empty = bars + 10;
  uint64_t* dq_full = bars + 11;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = blockIdx.x % g.splits;
  const int q0 = (blockIdx.x / g.splits) * kQ, head = blockIdx.y, b = blockIdx.z;
  const int jb = split * g.tiles_per_split;                                  // first key tile of this CTA
  const int T = min((g.HW + kK - 1) / kK, jb + g.tiles_per_split) - jb;      // its number of key tiles (>= 1)

  if (warp == 0 && lane == 0) {
    mbar_init(qdo_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
      mbar_init(&sdp_full[i], 1);
      mbar_init(&sdp_empty[i], 8);
    }
    mbar_init(ds_full, 8);
    mbar_init(ds_empty, 1);
    mbar_init(dq_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(qdo_full, 4 * kQBytes);
      tma_load_3d(sQ, &tmQh, qdo_full, head * 32, q0, b);
      tma_load_3d(sQ + kQBytes, &tmQl, qdo_full, head * 32, q0, b);
      tma_load_3d(sQ + 2 * kQBytes, &tmDh, qdo_full, head * 32, q0, b);
      tma_load_3d(sQ + 3 * kQBytes, &tmDl, qdo_full, head * 32, q0, b);
      for (int j = 0; j < T; ++j) {
        const int st = j & 1;
        mbar_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1);
        uint8_t* s = sKV + st * kStage;
        mbar_arrive_expect_tx(&kv_full[st], kStage);
        const int key0 = (jb + j) * kK;
        tma_load_3d(s, &tmKh, &kv_full[st], head * 32, key0, b);
        tma_load_3d(s + kKBytes, &tmKl, &kv_full[st], head * 32, key0, b);
        tma_load_3d(s + 2 * kKBytes, &tmVh, &kv_full[st], head * 32, key0, b);
        tma_load_3d(s + 3 * kKBytes, &tmVl, &kv_full[st], head * 32, key0, b);
        tma_load_3d(s + 4 * kKBytes, &tmKth, &kv_full[st], key0, head * 32, b);
        tma_load_3d(s + 4 * kKBytes + kKtBytes / 2, &tmKth, &kv_full[st], key0 + 32, head * 32, b);
        tma_load_3d(s + 4 * kKBytes + kKtBytes, &tmKtl, &kv_full[st], key0, head * 32, b);
        tma_load_3d(s + 4 * kKBytes + kKtBytes + kKtBytes / 2, &tmKtl, &kv_full[st], key0 + 32, head * 32, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = idesc_tf32(kQ, kK);     // 128 x 64
      constexpr uint32_t idesc_q = idesc_tf32(kQ, 32);     // 128 x 32
      const uint32_t q_hi = smem_u32(sQ), q_lo = q_hi + kQBytes, do_hi = q_hi + 2 * kQBytes, do_lo = q_hi + 3 * kQBytes;
      const uint32_t ds_hi = smem_u32(sDS), ds_lo = ds_hi + kDsBytes;
      auto issue_sdp = [&](int j) {
        const int st = j & 1;
        mbar_wait(&kv_full[st], (j >> 1) & 1);
        mbar_wait(&sdp_empty[st], ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t k_hi = smem_u32(sKV + st * kStage), k_lo = k_hi + kKBytes;
        const uint32_t v_hi = k_hi + 2 * kKBytes, v_lo = k_hi + 3 * kKBytes;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          xb::mma3(tmem_base + kTS + st * kK, q_hi + k * 32, q_lo + k * 32, k_hi + k * 32, k_lo + k * 32, idesc_s, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          xb::mma3(tmem_base + kTP + st * kK, do_hi + k * 32, do_lo + k * 32, v_hi + k * 32, v_lo + k * 32, idesc_s, k > 0);
        mma_commit(&sdp_full[st]);
      };
      mbar_wait(qdo_full, 0);
      issue_sdp(0);
      for (int j = 0; j < T; ++j) {
        if (j + 1 < T) issue_sdp(j + 1);
        const int st = j & 1;
        mbar_wait(ds_full, j & 1);
        tc_fence_after();
        const uint32_t kt_hi = smem_u32(sKV + st * kStage + 4 * kKBytes), kt_lo = kt_hi + kKtBytes;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t ao = (k >> 2) * kAtom + (k & 3) * 32;
          const uint32_t bo = (k >> 2) * (kKtBytes / 2) + (k & 3) * 32;
          xb::mma3(tmem_base + kTQ, ds_hi + ao, ds_lo + ao, kt_hi + bo, kt_lo + bo, idesc_q, (j | k) != 0);
        }
        mma_commit(ds_empty);
        mma_commit(&kv_empty[st]);
      }
      mma_commit(dq_full);
    }
  } else if (warp >= 4) {
    // 8 softmax warps: warps w and w + 4 share a TMEM lane quarter (a query row belongs to two threads); each takes
    // one 32-key half of the 64-key tile = one [128 x 32] atom of dS
    const int ew = (warp - 4) & 3;
    const int h = (warp - 4) >> 2;
    const int r = ew * 32 + lane;
    const int q = q0 + r;
    const bool q_ok = q < g.Qt;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16);
    const uint32_t* brow = g.bits + (static_cast<long long>(b) * g.Qt + (q_ok ? q : 0)) * g.words;
    const bool open = !q_ok || (g.row_open != nullptr && g.row_open[static_cast<long long>(b) * g.Qt + q] != 0);
    const long long rowi = (static_cast<long long>(b) * g.heads + head) * g.Qt + (q_ok ? q : 0);
    const float lse = q_ok ? g.lse2[rowi] : 0.f;
    const float dlt = q_ok ? g.delta[rowi] : 0.f;

    for (int j = 0; j < T; ++j) {
      const int st = j & 1;
      const int key0 = (jb + j) * kK + 32 * h;
      uint32_t w = 0u;
      if (!open) w = brow[2 * (jb + j) + h];
      if (key0 + 32 > g.HW) w |= (key0 >= g.HW) ? 0xFFFFFFFFu : (0xFFFFFFFFu << (g.HW - key0));
      if (!q_ok) w = 0xFFFFFFFFu;                                  // padding rows contribute nothing
      mbar_wait(&sdp_full[st], (j >> 1) & 1);
      tc_fence_after();
      float ds[32];
      {
        uint32_t sv[32], pv[32];
        tmem_ld_32x32(lane_addr + kTS + st * kK + h * 32, sv);
        tmem_ld_32x32(lane_addr + kTP + st * kK + h * 32, pv);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float p = ((w >> i) & 1u) ? 0.f : exp2f(__uint_as_float(sv[i]) - lse);
          ds[i] = p * (__uint_as_float(pv[i]) - dlt);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sdp_empty[st]);
      mbar_wait(ds_empty, (j & 1) ^ 1);
      xb::store_row_atom(sDS + h * kAtom, sDS + kDsBytes + h * kAtom, r, ds);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_full);
    }
    mbar_wait(dq_full, 0);
    tc_fence_after();
    // the row's two threads store 16 of the 32 head-dim values each
    uint32_t v[16];
    tmem_ld_32x16(lane_addr + kTQ + 16 * h, v);
    tmem_ld_wait();
    if (q_ok) {
      float* base = g.splits == 1 ? g.dq : g.part_dq + static_cast<long long>(split) * g.B * g.Qt * g.E;
      float* dst = base + (static_cast<long long>(b) * g.Qt + q) * g.E + head * 32 + 16 * h;
#pragma unroll
      for (int i = 0; i < 16; i += 4)
        *reinterpret_cast<float4*>(dst + i) =
            make_float4(__uint_as_float(v[i]) * g.inv_sqrt_d, __uint_as_float(v[i + 1]) * g.inv_sqrt_d,
                        __uint_as_float(v[i + 2]) * g.inv_sqrt_d, __uint_as_float(v[i + 3]) * g.inv_sqrt_d);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ================================================================================================
// kernel B: dK, dV
// ================================================================================================
namespace xk {
// CTA = (128-key tile, head, image); the queries are walked in chunks of 32 through a software pipeline:
//   T1(c)  MMA      S^T(c) = K Qs(c)^T,  dP^T(c) = V dO(c)^T                 -> TMEM buffer c & 1
//   T2(c)  softmax  P^T(c), dS^T(c) from the TMEM buffer -> shared memory (hi/lo, UMMA K-major)
//   T3(c)  MMA      dV += P^T(c) dO(c),  dK += dS^T(c) Qs(c)
// The MMA warp issues T1(c+1) BEFORE T3(c), so the tensor pipe works on the next chunk's scores while the softmax
// warps turn the current ones into operands; chunk loads (Q, dO and their transposes) run two stages ahead.
// (The first version used 64-query chunks, one stage and one shared P/dS buffer: every step of the chain waited for
// the previous one -- 7 % tensor-pipe, 8 % DRAM at one 192 KB CTA per SM, profiles/r1x_ncu_xattn_bwd_dkv.txt.)
constexpr int kKeys = 128, kQc = 32;
constexpr int kKVBytes = kKeys * 32 * 4;        // 16 KiB (one of K_hi, K_lo, V_hi, V_lo)
constexpr int kQcBytes = kQc * 32 * 4;          // 4 KiB  (Q / dO chunk, hi or lo: [32 q x 32 d])
constexpr int kQtBytes = 32 * kQc * 4;          // 4 KiB  (Q^T / dO^T chunk, hi or lo: one [32 d x 32 q] atom)
constexpr int kChunk = 4 * kQcBytes + 4 * kQtBytes;     // 32 KiB: Q_hi Q_lo dO_hi dO_lo | Qt_hi Qt_lo dOt_hi dOt_lo
constexpr int kStages = 2;
constexpr int kAtom = kKeys * 32 * 4;           // 16 KiB: [128 keys x 32 q]
constexpr int kPBytes = 2 * kAtom;              // P^T hi | lo  (same for dS^T)
constexpr int kSmem = 4 * kKVBytes + kStages * kChunk + 2 * kPBytes + 256 + 1024;
constexpr uint32_t kTmemCols = 256;
constexpr int kTS = 0, kTP = 64, kTV = 128, kTK = 160;   // S^T 2 x 32, dP^T 2 x 32, dV 32, dK 32 columns
// 20 warps: 0 TMA, 1 MMA, 2 TMEM allocation, 4-19 softmax.  A key row (TMEM lane) is shared by FOUR threads -- warps
// w, w + 4, w + 8, w + 12 address the same lane quarter -- each taking 8 of the chunk's 32 query columns: the exp2 /
// split / store work of a chunk is the critical path of the tile, and with one softmax warp per scheduler it ran at
// the issue latency of a single dependent instruction stream (one warp per row: 9.8 ms per step, two: 6.6 ms).
constexpr int kThreads = 640;
}  // namespace xk

__global__ void __launch_bounds__(xk::kThreads, 1)
masked_xattn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmKh, const __grid_constant__ CUtensorMap tmKl,
                            const __grid_constant__ CUtensorMap tmVh, const __grid_constant__ CUtensorMap tmVl,
                            const __grid_constant__ CUtensorMap tmQh, const __grid_constant__ CUtensorMap tmQl,
                            const __grid_constant__ CUtensorMap tmDh, const __grid_constant__ CUtensorMap tmDl,
                            const __grid_constant__ CUtensorMap tmQth, const __grid_constant__ CUtensorMap tmQtl,
                            const __grid_constant__ CUtensorMap tmDth, const __grid_constant__ CUtensorMap tmDtl,
                            const XbwdArgs g) {
  using namespace xk;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sKV = smem;                        // K_hi | K_lo | V_hi | V_lo
  uint8_t* sC = smem + 4 * kKVBytes;          // kStages query-chunk stages
  uint8_t* sP = sC + kStages * kChunk;        // P^T: hi | lo
  uint8_t* sD = sP + kPBytes;                 // dS^T: hi | lo
  uint64_t* bars = reinterpret_cast<uint64_t*>(sD + kPBytes);
  uint64_t* kv_full = bars;
  uint64_t* qc_full = bars + 1;               // [2] chunk stage landed
  uint64_t* qc_empty = bars + 3;              // [2] commit after the chunk's dV / dK MMAs: stage free
  uint64_t* s_full = bars + 5;                // [2] commit after S^T / dP^T MMAs: TMEM buffer ready
  uint64_t* s_empty = bars + 7;               // [2] count 16: softmax warps have read the TMEM buffer
  uint64_t* p_full = bars + 9;                // count 16: P^T and dS^T written
  uint64_t* p_empty = bars + 10;              // commit after the dV / dK MMAs: P^T / dS^T buffers free
  uint64_t* acc_full = bars + 11;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int key0 = blockIdx.x * kKeys, head = blockIdx.y, b = blockIdx.z;
  const int NC = (g.Qt + kQc - 1) / kQc;

  if (warp == 0 && lane == 0) {
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&qc_full[s], 1);
      mbar_init(&qc_empty[s], 1);
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], 16);
    }
    mbar_init(p_full, 16);
    mbar_init(p_empty, 1);
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(kv_full, 4 * kKVBytes);
      tma_load_3d(sKV, &tmKh, kv_full, head * 32, key0, b);
      tma_load_3d(sKV + kKVBytes, &tmKl, kv_full, head * 32, key0, b);
      tma_load_3d(sKV + 2 * kKVBytes, &tmVh, kv_full, head * 32, key0, b);
      tma_load_3d(sKV + 3 * kKVBytes, &tmVl, kv_full, head * 32, key0, b);
      for (int c = 0; c < NC; ++c) {
        const int s = c & 1;
        mbar_wait(&qc_empty[s], ((c >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&qc_full[s], kChunk);
        const int qc0 = c * kQc;
        uint8_t* st = sC + s * kChunk;
        tma_load_3d(st, &tmQh, &qc_full[s], head * 32, qc0, b);
        tma_load_3d(st + kQcBytes, &tmQl, &qc_full[s], head * 32, qc0, b);
        tma_load_3d(st + 2 * kQcBytes, &tmDh, &qc_full[s], head * 32, qc0, b);
        tma_load_3d(st + 3 * kQcBytes, &tmDl, &qc_full[s], head * 32, qc0, b);
        uint8_t* t = st + 4 * kQcBytes;
        tma_load_3d(t, &tmQth, &qc_full[s], qc0, head * 32, b);
        tma_load_3d(t + kQtBytes, &tmQtl, &qc_full[s], qc0, head * 32, b);
        tma_load_3d(t + 2 * kQtBytes, &tmDth, &qc_full[s], qc0, head * 32, b);
        tma_load_3d(t + 3 * kQtBytes, &tmDtl, &qc_full[s], qc0, head * 32, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = idesc_tf32(kKeys, kQc);   // 128 x 32
      constexpr uint32_t idesc_o = idesc_tf32(kKeys, 32);    // 128 x 32
      const uint32_t k_hi = smem_u32(sKV), k_lo = k_hi + kKVBytes, v_hi = k_hi + 2 * kKVBytes, v_lo = k_hi + 3 * kKVBytes;
      const uint32_t p_hi = smem_u32(sP), p_lo = p_hi + kAtom, d_hi = smem_u32(sD), d_lo = d_hi + kAtom;
      auto issue_scores = [&](int c) {          // T1(c)
        const int s = c & 1;
        mbar_wait(&qc_full[s], (c >> 1) & 1);
        mbar_wait(&s_empty[s], ((c >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t q_hi = smem_u32(sC + s * kChunk), q_lo = q_hi + kQcBytes, do_hi = q_hi + 2 * kQcBytes,
                       do_lo = q_hi + 3 * kQcBytes;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          xb::mma3(tmem_base + kTS + s * 32, k_hi + k * 32, k_lo + k * 32, q_hi + k * 32, q_lo + k * 32, idesc_s, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          xb::mma3(tmem_base + kTP + s * 32, v_hi + k * 32, v_lo + k * 32, do_hi + k * 32, do_lo + k * 32, idesc_s, k > 0);
        mma_commit(&s_full[s]);
      };
      mbar_wait(kv_full, 0);
      issue_scores(0);
      for (int c = 0; c < NC; ++c) {
        if (c + 1 < NC) issue_scores(c + 1);
        const int s = c & 1;
        const uint32_t qt_hi = smem_u32(sC + s * kChunk) + 4 * kQcBytes, qt_lo = qt_hi + kQtBytes,
                       dt_hi = qt_hi + 2 * kQtBytes, dt_lo = qt_hi + 3 * kQtBytes;
        mbar_wait(p_full, c & 1);               // T3(c)
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          xb::mma3(tmem_base + kTV, p_hi + k * 32, p_lo + k * 32, dt_hi + k * 32, dt_lo + k * 32, idesc_o, (c | k) != 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          xb::mma3(tmem_base + kTK, d_hi + k * 32, d_lo + k * 32, qt_hi + k * 32, qt_lo + k * 32, idesc_o, (c | k) != 0);
        mma_commit(p_empty);
        mma_commit(&qc_empty[s]);
      }
      mma_commit(acc_full);
    }
  } else if (warp >= 4) {
    const int ew = (warp - 4) & 3;                // TMEM lane quarter
    const int part = (warp - 4) >> 2;             // which 8 of the chunk's 32 query columns (0..3)
    const int r = ew * 32 + lane;                 // key row inside the tile == TMEM lane
    const int key = key0 + r;
    const bool key_ok = key < g.HW;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16);
    const uint32_t bit = 1u << (key & 31);
    const float* lse_b = g.lse2 + (static_cast<long long>(b) * g.heads + head) * g.Qt;
    const float* dlt_b = g.delta + (static_cast<long long>(b) * g.heads + head) * g.Qt;
    // per-chunk column data (log-sum-exp, delta, the mask word of each of the tile's four 32-key groups), staged by
    // the first 128 softmax threads; double buffered by chunk parity, one named barrier (512 threads) per chunk
    __shared__ float s_lse[2][kQc], s_dlt[2][kQc];
    __shared__ uint32_t s_mw[2][kQc][4];
    const int st_id = threadIdx.x - 128;          // 0..511

    for (int c = 0; c < NC; ++c) {
      const int qc0 = c * kQc;
      const int pb = c & 1;
      if (st_id < kQc) {
        const int q = qc0 + st_id;
        s_lse[pb][st_id] = q < g.Qt ? __ldg(lse_b + q) : 0.f;
        s_dlt[pb][st_id] = q < g.Qt ? __ldg(dlt_b + q) : 0.f;
      }
      if (st_id < 4 * kQc) {
        const int ql = st_id >> 2, wg = st_id & 3;            // kQc * 4 == 128 threads: one word each
        const int q = qc0 + ql;
        uint32_t w = 0xFFFFFFFFu;                  // queries beyond Qt contribute nothing
        if (q < g.Qt) {
          const long long qi = static_cast<long long>(b) * g.Qt + q;
          const bool open = g.row_open != nullptr && g.row_open[qi] != 0;
          const int wi = (key0 >> 5) + wg;
          w = open ? 0u : (wi < g.words ? g.bits[qi * g.words + wi] : 0xFFFFFFFFu);
        }
        s_mw[pb][ql][wg] = w;
      }
      asm volatile("bar.sync 1, 512;" ::: "memory");
      mbar_wait(&s_full[pb], (c >> 1) & 1);
      tc_fence_after();
      float pt[8], ds[8];
      {
        uint32_t sv[8], pv[8];
        tmem_ld_32x8(lane_addr + kTS + pb * 32 + part * 8, sv);
        tmem_ld_32x8(lane_addr + kTP + pb * 32 + part * 8, pv);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int ql = part * 8 + i;
          const bool masked = !key_ok || (s_mw[pb][ql][ew] & bit) != 0u;
          const float p = masked ? 0.f : exp2f(__uint_as_float(sv[i]) - s_lse[pb][ql]);
          pt[i] = p;
          ds[i] = p * (__uint_as_float(pv[i]) - s_dlt[pb][ql]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[pb]);
      // the P^T / dS^T buffers are free once the previous chunk's dV / dK MMAs retired
      if (c > 0) mbar_wait(p_empty, (c - 1) & 1);
      xb::store_quarter_row_atom(sP, sP + kAtom, r, part, pt);
      xb::store_quarter_row_atom(sD, sD + kAtom, r, part, ds);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    mbar_wait(acc_full, 0);
    tc_fence_after();
    // the four threads of a row share the final store: parts 0, 1 write the halves of dV, parts 2, 3 those of dK
    uint32_t vv[16];
    const int which = part >> 1, hh = part & 1;
    tmem_ld_32x16(lane_addr + (which == 0 ? kTV : kTK) + 16 * hh, vv);
    tmem_ld_wait();
    if (key_ok) {
      float* dst = (which == 0 ? g.dv : g.dk) + (static_cast<long long>(b) * g.HW + key) * g.E + head * 32 + 16 * hh;
      const float sc = which == 0 ? 1.f : xb::kLn2;
#pragma unroll
      for (int i = 0; i < 16; i += 4)
        *reinterpret_cast<float4*>(dst + i) = make_float4(__uint_as_float(vv[i]) * sc, __uint_as_float(vv[i + 1]) * sc,
                                                          __uint_as_float(vv[i + 2]) * sc, __uint_as_float(vv[i + 3]) * sc);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// out[i] = sum_s part[s][i]: the key splits' partial dQ in a fixed order
__global__ void __launch_bounds__(256)
xattn_sum_splits_kernel(const float4* __restrict__ part, int splits, long long n4, float4* __restrict__ out) {
  const long long i = blockIdx.x * 256ll + threadIdx.x;
  if (i >= n4) return;
  float4 a = __ldg(part + i);
  for (int s = 1; s < splits; ++s) {
    const float4 v = __ldg(part + s * n4 + i);
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
  }
  out[i] = a;
}

}  // namespace mpf

extern "C" {

int mpf_masked_xattn_bwd_f32_ex(const float* q_hi, const float* q_lo, const float* qt_hi, const float* qt_lo,
                                const float* k_hi, const float* k_lo, const float* kt_hi, const float* kt_lo,
                                const float* v_hi, const float* v_lo, const float* do_hi, const float* do_lo,
                                const float* dot_hi, const float* dot_lo, const uint32_t* mask_bits,
                                const uint8_t* row_open, const float* lse2, const float* delta, float* dq, float* dk,
                                float* dv, int B, int Qt, int qt_ld, int HW, int heads, int head_dim, int mask_words,
                                int key_splits, float* ws_dq, void* stream);

int mpf_masked_xattn_bwd_f32(const float* q_hi, const float* q_lo, const float* qt_hi, const float* qt_lo,
                             const float* k_hi, const float* k_lo, const float* kt_hi, const float* kt_lo,
                             const float* v_hi, const float* v_lo, const float* do_hi, const float* do_lo,
                             const float* dot_hi, const float* dot_lo, const uint32_t* mask_bits,
                             const uint8_t* row_open, const float* lse2, const float* delta, float* dq, float* dk,
                             float* dv, int B, int Qt, int qt_ld, int HW, int heads, int head_dim, int mask_words,
                             void* stream) {
  return mpf_masked_xattn_bwd_f32_ex(q_hi, q_lo, qt_hi, qt_lo, k_hi, k_lo, kt_hi, kt_lo, v_hi, v_lo, do_hi, do_lo,
                                     dot_hi, dot_lo, mask_bits, row_open, lse2, delta, dq, dk, dv, B, Qt, qt_ld, HW,
                                     heads, head_dim, mask_words, 1, nullptr, stream);
}

int mpf_masked_xattn_bwd_f32_ex(const float* q_hi, const float* q_lo, const float* qt_hi, const float* qt_lo,
                                const float* k_hi, const float* k_lo, const float* kt_hi, const float* kt_lo,
                                const float* v_hi, const float* v_lo, const float* do_hi, const float* do_lo,
                                const float* dot_hi, const float* dot_lo, const uint32_t* mask_bits,
                                const uint8_t* row_open, const float* lse2, const float* delta, float* dq, float* dk,
                                float* dv, int B, int Qt, int qt_ld, int HW, int heads, int head_dim, int mask_words,
                                int key_splits, float* ws_dq, void* stream) {
  using namespace mpf;
  clear_error();
  MPF_REQUIRE(q_hi && q_lo && qt_hi && qt_lo && k_hi && k_lo && kt_hi && kt_lo && v_hi && v_lo && do_hi && do_lo &&
                  dot_hi && dot_lo && mask_bits && lse2 && delta && dq && dk && dv,
              "masked_xattn_bwd: null pointer");
  MPF_REQUIRE(B > 0 && Qt > 0 && HW > 0 && heads > 0, "masked_xattn_bwd: sizes must be positive");
  MPF_REQUIRE(head_dim == xb::kD, "masked_xattn_bwd: head_dim must be 32 (got %d)", head_dim);
  MPF_REQUIRE(HW % 4 == 0 && qt_ld % 4 == 0 && qt_ld >= Qt,
              "masked_xattn_bwd: HW (%d) and the row stride of Q^T / dO^T (%d, >= Qt = %d) must be multiples of 4",
              HW, qt_ld, Qt);
  MPF_REQUIRE(mask_words >= 2 * ((HW + 63) / 64), "masked_xattn_bwd: mask_words too small");
  MPF_REQUIRE(heads <= 65535 && B <= 65535, "masked_xattn_bwd: grid too large");
  const int Ta = (HW + xa::kK - 1) / xa::kK;
  MPF_REQUIRE(key_splits >= 1 && key_splits <= Ta, "masked_xattn_bwd: key_splits (%d) must be in [1, %d key tiles]",
              key_splits, Ta);
  MPF_REQUIRE(key_splits == 1 || ws_dq, "masked_xattn_bwd: key_splits > 1 needs the dQ workspace");
  const int tiles_per_split = (Ta + key_splits - 1) / key_splits;
  const int splits = (Ta + tiles_per_split - 1) / tiles_per_split;
  const int E = heads * head_dim;
  const long long qs = static_cast<long long>(Qt) * E, ks = static_cast<long long>(HW) * E;
  CUtensorMap tq_h, tq_l, td_h, td_l, tk_h, tk_l, tv_h, tv_l, tkt_h, tkt_l;
  int rc;
  // ---- kernel A maps: Q/dO [128 x 32], K/V [64 x 32], K^T atoms [32 d x 32 keys]
  if ((rc = make_tmap_f32_3d(&tq_h, q_hi, E, Qt, B, E, qs, 32, xa::kQ))) return rc;
  if ((rc = make_tmap_f32_3d(&tq_l, q_lo, E, Qt, B, E, qs, 32, xa::kQ))) return rc;
  if ((rc = make_tmap_f32_3d(&td_h, do_hi, E, Qt, B, E, qs, 32, xa::kQ))) return rc;
  if ((rc = make_tmap_f32_3d(&td_l, do_lo, E, Qt, B, E, qs, 32, xa::kQ))) return rc;
  if ((rc = make_tmap_f32_3d(&tk_h, k_hi, E, HW, B, E, ks, 32, xa::kK))) return rc;
  if ((rc = make_tmap_f32_3d(&tk_l, k_lo, E, HW, B, E, ks, 32, xa::kK))) return rc;
  if ((rc = make_tmap_f32_3d(&tv_h, v_hi, E, HW, B, E, ks, 32, xa::kK))) return rc;
  if ((rc = make_tmap_f32_3d(&tv_l, v_lo, E, HW, B, E, ks, 32, xa::kK))) return rc;
  if ((rc = make_tmap_f32_3d(&tkt_h, kt_hi, HW, E, B, HW, ks, 32, 32))) return rc;
  if ((rc = make_tmap_f32_3d(&tkt_l, kt_lo, HW, E, B, HW, ks, 32, 32))) return rc;
  static unsigned long long configured_on = 0;
  if (first_use_on_this_device(configured_on)) {
    MPF_CUDA_OK(cudaFuncSetAttribute(masked_xattn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, xa::kSmem));
    MPF_CUDA_OK(cudaFuncSetAttribute(masked_xattn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, xk::kSmem));
  }
  XbwdArgs g;
  g.bits = mask_bits; g.row_open = row_open; g.lse2 = lse2; g.delta = delta; g.dq = dq; g.dk = dk; g.dv = dv;
  g.B = B; g.Qt = Qt; g.HW = HW; g.E = E; g.heads = heads; g.words = mask_words;
  g.inv_sqrt_d = 1.0f / sqrtf(static_cast<float>(head_dim));
  g.splits = splits; g.tiles_per_split = tiles_per_split; g.part_dq = ws_dq;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid_a(((Qt + xa::kQ - 1) / xa::kQ) * splits, heads, B);
  masked_xattn_bwd_dq_kernel<<<grid_a, xa::kThreads, xa::kSmem, st>>>(tq_h, tq_l, td_h, td_l, tk_h, tk_l, tv_h, tv_l,
                                                                      tkt_h, tkt_l, g);
  count_launch();
  if ((rc = finish_launch("masked_xattn_bwd_dq"))) return rc;
  if (splits > 1) {
    const long long n4 = static_cast<long long>(B) * Qt * E / 4;
    xattn_sum_splits_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<const float4*>(ws_dq), splits, n4, reinterpret_cast<float4*>(dq));
    count_launch();
    if ((rc = finish_launch("masked_xattn_bwd_dq (split sum)"))) return rc;
  }
  // ---- kernel B maps: K/V [128 x 32], Q/dO chunks [64 x 32], Q^T/dO^T atoms [32 d x 32 q]
  CUtensorMap bk_h, bk_l, bv_h, bv_l, bq_h, bq_l, bd_h, bd_l, bqt_h, bqt_l, bdt_h, bdt_l;
  if ((rc = make_tmap_f32_3d(&bk_h, k_hi, E, HW, B, E, ks, 32, xk::kKeys))) return rc;
  if ((rc = make_tmap_f32_3d(&bk_l, k_lo, E, HW, B, E, ks, 32, xk::kKeys))) return rc;
  if ((rc = make_tmap_f32_3d(&bv_h, v_hi, E, HW, B, E, ks, 32, xk::kKeys))) return rc;
  if ((rc = make_tmap_f32_3d(&bv_l, v_lo, E, HW, B, E, ks, 32, xk::kKeys))) return rc;
  if ((rc = make_tmap_f32_3d(&bq_h, q_hi, E, Qt, B, E, qs, 32, xk::kQc))) return rc;
  if ((rc = make_tmap_f32_3d(&bq_l, q_lo, E, Qt, B, E, qs, 32, xk::kQc))) return rc;
  if ((rc = make_tmap_f32_3d(&bd_h, do_hi, E, Qt, B, E, qs, 32, xk::kQc))) return rc;
  if ((rc = make_tmap_f32_3d(&bd_l, do_lo, E, Qt, B, E, qs, 32, xk::kQc))) return rc;
  const long long qts = static_cast<long long>(qt_ld) * E;
  if ((rc = make_tmap_f32_3d(&bqt_h, qt_hi, Qt, E, B, qt_ld, qts, 32, 32))) return rc;
  if ((rc = make_tmap_f32_3d(&bqt_l, qt_lo, Qt, E, B, qt_ld, qts, 32, 32))) return rc;
  if ((rc = make_tmap_f32_3d(&bdt_h, dot_hi, Qt, E, B, qt_ld, qts, 32, 32))) return rc;
  if ((rc = make_tmap_f32_3d(&bdt_l, dot_lo, Qt, E, B, qt_ld, qts, 32, 32))) return rc;
  dim3 grid_b((HW + xk::kKeys - 1) / xk::kKeys, heads, B);
  masked_xattn_bwd_dkv_kernel<<<grid_b, xk::kThreads, xk::kSmem, st>>>(bk_h, bk_l, bv_h, bv_l, bq_h, bq_l, bd_h, bd_l,
                                                                       bqt_h, bqt_l, bdt_h, bdt_l, g);
  count_launch();
  return finish_launch("masked_xattn_bwd_dkv");
}

}  // extern "C"

// Fused masked multi-head cross-attention, forward:   O = softmax(Q K^T + mask) V   per (image, head)
//
//   ref: transformer_decoder/mask2former_transformer_decoder.py:100-112 (nn.MultiheadAttention with a
//        boolean attn_mask), :1780 (a query whose keys are ALL masked attends to every key).
//
// The reference materialises fp32 scores, a float -inf mask and probabilities of shape [B*8, Q, HW]
// (three times) in HBM; here scores live in TMEM, the mask is one bit per (image, query, key) shared
// by the 8 heads and applied in registers before an online softmax, and probabilities go straight
// back to the tensor core through shared memory.
//
// Operands (all fp32, pre-split x = hi + lo into TF32-exact halves by the projection GEMM epilogues,
// "3xTF32": S and O are computed as lo*hi + hi*lo + hi*hi with fp32 accumulation in TMEM):
//   Q   [B, Qt, E]      query projection, ALREADY scaled by log2(e)/sqrt(head_dim)
//   K   [B, HW, E]      key projection
//   Vt  [B, E, HW]      value projection, transposed (keys contiguous) so P*V is a K-major UMMA
//   bits [B, Qt, W32]   1 = key masked;  row_open [B, Qt] = 1 if the row must ignore the mask
// Outputs: O [B, Qt, E] (normalised), lse2 [B, heads, Qt] = m + log2(l) (log2 domain, for backward).
//
// One CTA = (128-query tile, head, image).  Keys are walked in tiles of 64:
//   warp 0   TMA producer (Q once; K / Vt tiles, 2-stage ring)
//   warp 1   MMA issuer:  S(j) = Q K(j)^T -> TMEM (double buffered);  O_tile(j) = P(j) Vt(j)^T -> TMEM
//   warps 4-7 softmax (thread = query row = TMEM lane): tcgen05.ld S, mask, online max / sum with
//            exp2, split P into hi/lo and store them to smem in the UMMA K-major SWIZZLE_128B
//            layout, then accumulate the finished O_tile(j-1) into registers with the rescale factor.
#include "mpf_common.cuh"
#include "sm100_ptx.cuh"
#include "tmap.cuh"

namespace mpf {

using namespace ptx;

constexpr int kXQ = 128;                 // queries per CTA (UMMA M)
constexpr int kXK = 64;                  // keys per tile
constexpr int kXD = 32;                  // head dim (= one 128-byte swizzle row of fp32)
constexpr int kXThreads = 384;            // warps 0 TMA, 1 MMA, 2 TMEM allocation, 4-11 softmax (two threads per row)
constexpr int kQBytes = kXQ * kXD * 4;   // 16 KiB  (one of Q_hi / Q_lo)
constexpr int kKBytes = kXK * kXD * 4;   // 8 KiB   (one of K_hi / K_lo)
constexpr int kVBytes = kXD * kXK * 4;   // 8 KiB   (one of Vt_hi / Vt_lo): two 32x32 atoms of 4 KiB
constexpr int kPAtom = kXQ * 32 * 4;     // 16 KiB  (128 rows x 32 keys)
constexpr int kPBytes = 2 * kPAtom;      // 32 KiB  (one of P_hi / P_lo)
constexpr int kKVStage = 2 * kKBytes + 2 * kVBytes;          // 32 KiB
constexpr int kKVStages = 3;             // K / Vt ring: loads run two tiles ahead of the MMAs
constexpr int kXSmem = 2 * kQBytes + kKVStages * kKVStage + 2 * kPBytes + 256 + 1024;
constexpr uint32_t kTmemColsX = 256;     // S: 2 x 64, O_tile: 2 x 32
constexpr int kTmemS = 0, kTmemO = 128;

struct XattnArgs {
  const uint32_t* bits;      // [B, Qt, words]
  const uint8_t* row_open;   // [B, Qt] 1 = ignore mask for this row (all keys masked), may be null
  float* out;                // [B, Qt, E]
  float* lse2;               // [B, heads, Qt]
  int B, Qt, HW, E, heads, words;
  // key split (small batches: B * heads CTAs do not fill 148 SMs): blockIdx.x = q_tile * splits + split; a CTA walks
  // the key tiles [split * tiles_per_split, ...) and, when splits > 1, leaves its UNNORMALISED output with the
  // running maximum and sum in the workspace for xattn_merge_splits_kernel (log-sum-exp merge).
  int splits, tiles_per_split;
  float* part_o;             // [splits, B, Qt, E]
  float* part_ml;            // [splits, B, heads, Qt, 2]  (m, l) in the log2 domain
};

__device__ __forceinline__ float rn_tf32x(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

__global__ void __launch_bounds__(kXThreads, 1)
masked_xattn_fwd_kernel(const __grid_constant__ CUtensorMap tmQh, const __grid_constant__ CUtensorMap tmQl,
                        const __grid_constant__ CUtensorMap tmKh, const __grid_constant__ CUtensorMap tmKl,
                        const __grid_constant__ CUtensorMap tmVh, const __grid_constant__ CUtensorMap tmVl,
                        const XattnArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                   // Q_hi | Q_lo
  uint8_t* sKV = smem + 2 * kQBytes;                    // stage s: K_hi | K_lo | Vt_hi | Vt_lo
  uint8_t* sP = sKV + kKVStages * kKVStage;             // P_hi | P_lo
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kPBytes);
  uint64_t* q_full = bars;          // [1]
  uint64_t* kv_full = bars + 1;     // [kKVStages]
  uint64_t* kv_empty = bars + 4;    // [kKVStages]
  uint64_t* s_full = bars + 7;      // [2]
  uint64_t* s_empty = bars + 9;     // [2]
  uint64_t* p_full = bars + 11;     // [1]
  uint64_t* p_empty = bars + 12;    // [1]
  uint64_t* o_full = bars + 13;     // [2]
  uint64_t* o_empty = bars + 15;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = blockIdx.x % g.splits;
  const int q0 = (blockIdx.x / g.splits) * kXQ, head = blockIdx.y, b = blockIdx.z;
  const int jb = split * g.tiles_per_split;                       // first key tile of this CTA
  const int T = min((g.HW + kXK - 1) / kXK, jb + g.tiles_per_split) - jb;   // its number of key tiles (>= 1)

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQh); prefetch_tmap(&tmQl); prefetch_tmap(&tmKh);
    prefetch_tmap(&tmKl); prefetch_tmap(&tmVh); prefetch_tmap(&tmVl);
    mbar_init(q_full, 1);
    for (int i = 0; i < kKVStages; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 8);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_empty[i], 8);
    }
    mbar_init(p_full, 8);
    mbar_init(p_empty, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, kTmemColsX);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, 2 * kQBytes);
      tma_load_3d(sQ, &tmQh, q_full, head * kXD, q0, b);
      tma_load_3d(sQ + kQBytes, &tmQl, q_full, head * kXD, q0, b);
      for (int j = 0; j < T; ++j) {
        const int st = j % kKVStages;
        const uint32_t ph = (j / kKVStages) & 1;
        mbar_wait(&kv_empty[st], ph ^ 1);
        uint8_t* s = sKV + st * kKVStage;
        mbar_arrive_expect_tx(&kv_full[st], kKVStage);
        const int key0 = (jb + j) * kXK;
        tma_load_3d(s, &tmKh, &kv_full[st], head * kXD, key0, b);
        tma_load_3d(s + kKBytes, &tmKl, &kv_full[st], head * kXD, key0, b);
        tma_load_3d(s + 2 * kKBytes, &tmVh, &kv_full[st], key0, head * kXD, b);
        tma_load_3d(s + 2 * kKBytes + kVBytes / 2, &tmVh, &kv_full[st], key0 + 32, head * kXD, b);
        tma_load_3d(s + 2 * kKBytes + kVBytes, &tmVl, &kv_full[st], key0, head * kXD, b);
        tma_load_3d(s + 2 * kKBytes + kVBytes + kVBytes / 2, &tmVl, &kv_full[st], key0 + 32, head * kXD, b);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (whole warp, one elected lane issues) =================
    {
      constexpr uint32_t idesc_s = idesc_tf32(kXQ, kXK);   // S: 128 x 64
      constexpr uint32_t idesc_o = idesc_tf32(kXQ, kXD);   // O: 128 x 32
      const uint32_t q_hi = smem_u32(sQ), q_lo = q_hi + kQBytes;
      const uint32_t p_hi = smem_u32(sP), p_lo = p_hi + kPBytes;
      auto issue_s = [&](int j) {
        const int st = j & 1, kvs = j % kKVStages;
        mbar_wait(&kv_full[kvs], (j / kKVStages) & 1);
        mbar_wait(&s_empty[st], ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t k_hi = smem_u32(sKV + kvs * kKVStage), k_lo = k_hi + kKBytes;
        const uint32_t d = tmem_base + kTmemS + st * kXK;
#pragma unroll
        for (int k = 0; k < kXD / 8; ++k) {
          const uint32_t ko = k * 32;
          mma_tf32_ss_elect(d, smem_desc_sw128_kmajor(q_lo + ko), smem_desc_sw128_kmajor(k_hi + ko), idesc_s, k ? 1u : 0u);
          mma_tf32_ss_elect(d, smem_desc_sw128_kmajor(q_hi + ko), smem_desc_sw128_kmajor(k_lo + ko), idesc_s, 1u);
          mma_tf32_ss_elect(d, smem_desc_sw128_kmajor(q_hi + ko), smem_desc_sw128_kmajor(k_hi + ko), idesc_s, 1u);
        }
        mma_commit_elect(&s_full[st]);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < T; ++j) {
        if (j + 1 < T) issue_s(j + 1);
        const int st = j & 1;
        mbar_wait(p_full, j & 1);
        mbar_wait(&o_empty[st], ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        const int kvs = j % kKVStages;
        const uint32_t v_hi = smem_u32(sKV + kvs * kKVStage + 2 * kKBytes), v_lo = v_hi + kVBytes;
        const uint32_t d = tmem_base + kTmemO + st * kXD;
#pragma unroll
        for (int k = 0; k < kXK / 8; ++k) {
          const uint32_t po = (k >> 2) * kPAtom + (k & 3) * 32;        // 32-key atoms of P
          const uint32_t vo = (k >> 2) * (kVBytes / 2) + (k & 3) * 32; // 32-key atoms of Vt
          mma_tf32_ss_elect(d, smem_desc_sw128_kmajor(p_lo + po), smem_desc_sw128_kmajor(v_hi + vo), idesc_o, k ? 1u : 0u);
          mma_tf32_ss_elect(d, smem_desc_sw128_kmajor(p_hi + po), smem_desc_sw128_kmajor(v_lo + vo), idesc_o, 1u);
          mma_tf32_ss_elect(d, smem_desc_sw128_kmajor(p_hi + po), smem_desc_sw128_kmajor(v_hi + vo), idesc_o, 1u);
        }
        mma_commit_elect(&o_full[st]);
        mma_commit_elect(p_empty);
        mma_commit_elect(&kv_empty[kvs]);
      }
    }
  } else if (warp >= 4) {
    // ================= softmax / output warps =================
    // A query row (TMEM lane) belongs to TWO threads: warps w and w + 4 address the same lane quarter; thread h takes
    // the 32-key half h of every 64-key tile (one atom of P) and 16 of the 32 output columns.  The row maximum of a
    // tile is exchanged through shared memory (one named barrier per tile), so both threads derive identical
    // rescale factors; the row sums stay per thread until the end.  With one softmax warp per scheduler the exp2 /
    // split / store stream ran at single-warp issue latency and was the critical path of the tile.
    const int ew = (warp - 4) & 3;
    const int h = (warp - 4) >> 2;
    const int r = ew * 32 + lane;                 // query row inside the tile == TMEM lane
    const int q = q0 + r;
    const bool q_ok = q < g.Qt;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16);
    const uint32_t* brow = g.bits + (static_cast<long long>(b) * g.Qt + (q_ok ? q : 0)) * g.words;
    const bool open = !q_ok || (g.row_open != nullptr && g.row_open[static_cast<long long>(b) * g.Qt + q] != 0);
    __shared__ float s_xch[2][2][kXQ];            // [tile parity][half][row]

    float m = -INFINITY, l = 0.f, m_o = -INFINITY, m_prev = -INFINITY;
    float o[kXD / 2];
#pragma unroll
    for (int i = 0; i < kXD / 2; ++i) o[i] = 0.f;

    auto accumulate_o = [&](int j, float m_j) {
      const int st = j & 1;
      mbar_wait(&o_full[st], (j >> 1) & 1);
      tc_fence_after();
      uint32_t v[16];
      tmem_ld_32x16(lane_addr + kTmemO + st * kXD + 16 * h, v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_empty[st]);
      if (m_j != -INFINITY) {              // a fully masked tile contributes exactly zero
        const float sc = (m_o == -INFINITY) ? 0.f : exp2f(m_o - m_j);   // m_j >= m_o: never overflows
#pragma unroll
        for (int i = 0; i < kXD / 2; ++i) o[i] = o[i] * sc + __uint_as_float(v[i]);
        m_o = m_j;
      }
    };

    for (int j = 0; j < T; ++j) {
      const int st = j & 1;
      mbar_wait(&s_full[st], (j >> 1) & 1);
      tc_fence_after();
      uint32_t s0[32];
      tmem_ld_32x32(lane_addr + kTmemS + st * kXK + 32 * h, s0);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[st]);

      // mask word for keys [64j + 32h, 64j + 32h + 32): bit = 1 -> masked.  Keys >= HW are always masked.
      const int key0 = (jb + j) * kXK + 32 * h;
      uint32_t w0 = 0u;
      if (!open) w0 = brow[2 * (jb + j) + h];
      if (key0 + 32 > g.HW) w0 |= (key0 >= g.HW) ? 0xFFFFFFFFu : (0xFFFFFFFFu << (g.HW - key0));

      float tmax = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float a = ((w0 >> i) & 1u) ? -INFINITY : __uint_as_float(s0[i]);
        s0[i] = __float_as_uint(a);
        tmax = fmaxf(tmax, a);
      }
      s_xch[st][h][r] = tmax;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      tmax = fmaxf(tmax, s_xch[st][h ^ 1][r]);
      const float m_new = fmaxf(m, tmax);
      const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = (m == -INFINITY) ? 0.f : exp2f(m - m_new);
      float psum = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float a = exp2f(__uint_as_float(s0[i]) - m_safe);
        psum += a;
        s0[i] = __float_as_uint(a);
      }
      l = l * alpha + psum;                 // this thread's half of the row sum

      // P(j) -> smem (UMMA A operand, K-major, 128B swizzle): row r, 16-byte chunk c of atom a lives at
      // a*16K + r*128 + ((c ^ (r & 7)) << 4); this thread writes atom h
      mbar_wait(p_empty, (j & 1) ^ 1);
      uint8_t* prow = sP + h * kPAtom + r * 128;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int sw = ((c ^ (r & 7)) << 4);
        float4 h0, l0;
        const float a0 = __uint_as_float(s0[4 * c]), a1 = __uint_as_float(s0[4 * c + 1]);
        const float a2 = __uint_as_float(s0[4 * c + 2]), a3 = __uint_as_float(s0[4 * c + 3]);
        h0.x = rn_tf32x(a0); h0.y = rn_tf32x(a1); h0.z = rn_tf32x(a2); h0.w = rn_tf32x(a3);
        l0.x = rn_tf32x(a0 - h0.x); l0.y = rn_tf32x(a1 - h0.y); l0.z = rn_tf32x(a2 - h0.z); l0.w = rn_tf32x(a3 - h0.w);
        *reinterpret_cast<float4*>(prow + sw) = h0;                          // P_hi
        *reinterpret_cast<float4*>(prow + kPBytes + sw) = l0;                // P_lo
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);

      if (j > 0) accumulate_o(j - 1, m_prev);
      m_prev = m_new;       // running maximum the probabilities of tile j were computed against
      m = m_new;
    }
    accumulate_o(T - 1, m_prev);

    // total row sum = the two threads' halves (both were rescaled by identical factors)
    s_xch[T & 1][h][r] = l;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    l += s_xch[T & 1][h ^ 1][r];
    if (q_ok && g.splits == 1) {
      const float inv = 1.f / l;
      float* dst = g.out + (static_cast<long long>(b) * g.Qt + q) * g.E + head * kXD + 16 * h;
#pragma unroll
      for (int i = 0; i < kXD / 2; i += 4)
        *reinterpret_cast<float4*>(dst + i) = make_float4(o[i] * inv, o[i + 1] * inv, o[i + 2] * inv, o[i + 3] * inv);
      if (g.lse2 != nullptr && h == 0)
        g.lse2[(static_cast<long long>(b) * g.heads + head) * g.Qt + q] = m + log2f(l);
    } else if (q_ok) {
      // partial result of this key range: o is relative to the running maximum m_o (== m unless the last tiles
      // were fully masked, in which case o was not rescaled: bring it to m here)
      const float fix = (m_o == -INFINITY || m_o == m) ? 1.f : exp2f(m_o - m);
      float* dst = g.part_o + ((static_cast<long long>(split) * g.B + b) * g.Qt + q) * g.E + head * kXD + 16 * h;
#pragma unroll
      for (int i = 0; i < kXD / 2; i += 4)
        *reinterpret_cast<float4*>(dst + i) = make_float4(o[i] * fix, o[i + 1] * fix, o[i + 2] * fix, o[i + 3] * fix);
      if (h == 0) {
        float* ml = g.part_ml + ((((static_cast<long long>(split) * g.B + b) * g.heads + head) * g.Qt + q) << 1);
        ml[0] = m;
        ml[1] = l;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemColsX);
  }
}

// out[b, q, head*32 + d] = sum_s o_s * 2^(m_s - m*) / sum_s l_s * 2^(m_s - m*),  lse2 = m* + log2(sum ...): the
// log-sum-exp merge of the key splits.  One warp per (image, query, head), lane = head-dim channel.
__global__ void __launch_bounds__(256)
xattn_merge_splits_kernel(const float* __restrict__ part_o, const float* __restrict__ part_ml, int splits, int B,
                          int Qt, int heads, float* __restrict__ out, float* __restrict__ lse2) {
  const long long wid = (blockIdx.x * 256ll + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= static_cast<long long>(B) * Qt * heads) return;
  const int head = static_cast<int>(wid % heads);
  const long long bq = wid / heads;                     // b * Qt + q
  const int b = static_cast<int>(bq / Qt), q = static_cast<int>(bq - static_cast<long long>(b) * Qt);
  const int E = heads * kXD;
  float mstar = -INFINITY;
  for (int s = 0; s < splits; ++s)
    mstar = fmaxf(mstar, __ldg(part_ml + ((((static_cast<long long>(s) * B + b) * heads + head) * Qt + q) << 1)));
  float acc = 0.f, lsum = 0.f;
  for (int s = 0; s < splits; ++s) {
    const float* ml = part_ml + ((((static_cast<long long>(s) * B + b) * heads + head) * Qt + q) << 1);
    const float ms = __ldg(ml);
    if (ms == -INFINITY) continue;                       // this key range was fully masked for the row
    const float w = exp2f(ms - mstar);
    lsum += __ldg(ml + 1) * w;
    acc += __ldg(part_o + ((static_cast<long long>(s) * B + b) * Qt + q) * E + head * kXD + lane) * w;
  }
  out[(static_cast<long long>(b) * Qt + q) * E + head * kXD + lane] = acc / lsum;
  if (lse2 != nullptr && lane == 0) lse2[(static_cast<long long>(b) * heads + head) * Qt + q] = mstar + log2f(lsum);
}

}  // namespace mpf

extern "C" {

int mpf_masked_xattn_fwd_f32_ex(const float* q_hi, const float* q_lo, const float* k_hi, const float* k_lo,
                                const float* vt_hi, const float* vt_lo, const uint32_t* mask_bits,
                                const uint8_t* row_open, float* out, float* lse2, int B, int Qt, int HW, int heads,
                                int head_dim, int mask_words, int key_splits, float* ws_o, float* ws_ml,
                                void* stream);

int mpf_masked_xattn_fwd_f32(const float* q_hi, const float* q_lo, const float* k_hi, const float* k_lo,
                             const float* vt_hi, const float* vt_lo, const uint32_t* mask_bits,
                             const uint8_t* row_open, float* out, float* lse2, int B, int Qt, int HW,
                             int heads, int head_dim, int mask_words, void* stream) {
  return mpf_masked_xattn_fwd_f32_ex(q_hi, q_lo, k_hi, k_lo, vt_hi, vt_lo, mask_bits, row_open, out, lse2, B, Qt, HW,
                                     heads, head_dim, mask_words, 1, nullptr, nullptr, stream);
}

int mpf_masked_xattn_fwd_f32_ex(const float* q_hi, const float* q_lo, const float* k_hi, const float* k_lo,
                                const float* vt_hi, const float* vt_lo, const uint32_t* mask_bits,
                                const uint8_t* row_open, float* out, float* lse2, int B, int Qt, int HW, int heads,
                                int head_dim, int mask_words, int key_splits, float* ws_o, float* ws_ml,
                                void* stream) {
  using namespace mpf;
  clear_error();
  MPF_REQUIRE(q_hi && q_lo && k_hi && k_lo && vt_hi && vt_lo && mask_bits && out, "masked_xattn: null pointer");
  MPF_REQUIRE(B > 0 && Qt > 0 && HW > 0 && heads > 0, "masked_xattn: sizes must be positive");
  MPF_REQUIRE(head_dim == kXD, "masked_xattn: head_dim must be %d (got %d)", kXD, head_dim);
  MPF_REQUIRE(HW % 4 == 0, "masked_xattn: HW (%d) must be a multiple of 4 (TMA row stride of V^T)", HW);
  const int T = (HW + kXK - 1) / kXK;
  MPF_REQUIRE(mask_words >= 2 * T, "masked_xattn: mask_words (%d) must cover %d key tiles of 64", mask_words, T);
  MPF_REQUIRE(heads <= 65535 && B <= 65535, "masked_xattn: grid too large");
  MPF_REQUIRE(key_splits >= 1 && key_splits <= T, "masked_xattn: key_splits (%d) must be in [1, %d key tiles]",
              key_splits, T);
  MPF_REQUIRE(key_splits == 1 || (ws_o && ws_ml), "masked_xattn: key_splits > 1 needs the two workspaces");
  const int tiles_per_split = (T + key_splits - 1) / key_splits;
  const int splits = (T + tiles_per_split - 1) / tiles_per_split;       // no empty split
  const int E = heads * head_dim;
  CUtensorMap tqh, tql, tkh, tkl, tvh, tvl;
  int rc;
  if ((rc = make_tmap_f32_3d(&tqh, q_hi, E, Qt, B, E, static_cast<long long>(Qt) * E, kXD, kXQ))) return rc;
  if ((rc = make_tmap_f32_3d(&tql, q_lo, E, Qt, B, E, static_cast<long long>(Qt) * E, kXD, kXQ))) return rc;
  if ((rc = make_tmap_f32_3d(&tkh, k_hi, E, HW, B, E, static_cast<long long>(HW) * E, kXD, kXK))) return rc;
  if ((rc = make_tmap_f32_3d(&tkl, k_lo, E, HW, B, E, static_cast<long long>(HW) * E, kXD, kXK))) return rc;
  if ((rc = make_tmap_f32_3d(&tvh, vt_hi, HW, E, B, HW, static_cast<long long>(HW) * E, 32, kXD))) return rc;
  if ((rc = make_tmap_f32_3d(&tvl, vt_lo, HW, E, B, HW, static_cast<long long>(HW) * E, 32, kXD))) return rc;
  static unsigned long long configured_on = 0;
  if (first_use_on_this_device(configured_on)) {
    MPF_CUDA_OK(cudaFuncSetAttribute(masked_xattn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kXSmem));
  }
  XattnArgs g;
  g.bits = mask_bits; g.row_open = row_open; g.out = out; g.lse2 = lse2;
  g.B = B; g.Qt = Qt; g.HW = HW; g.E = E; g.heads = heads; g.words = mask_words;
  g.splits = splits; g.tiles_per_split = tiles_per_split; g.part_o = ws_o; g.part_ml = ws_ml;
  dim3 grid(((Qt + kXQ - 1) / kXQ) * splits, heads, B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  masked_xattn_fwd_kernel<<<grid, kXThreads, kXSmem, st>>>(tqh, tql, tkh, tkl, tvh, tvl, g);
  count_launch();
  rc = finish_launch("masked_xattn_fwd");
  if (rc != MPF_OK || splits == 1) return rc;
  const long long warps = static_cast<long long>(B) * Qt * heads;
  xattn_merge_splits_kernel<<<static_cast<unsigned>((warps * 32 + 255) / 256), 256, 0, st>>>(ws_o, ws_ml, splits, B, Qt,
                                                                                             heads, out, lse2);
  count_launch();
  return finish_launch("masked_xattn_fwd (split merge)");
}

}  // extern "C"

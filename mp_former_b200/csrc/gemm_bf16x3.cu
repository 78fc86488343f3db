// C[b] = A[b] * B[b]^T (+ bias, + residual, * alpha, ReLU, gate) with "bf16x3" split arithmetic on the 5th-gen
// tensor cores: every fp32 operand value is written x = hi + lo with hi = bf16_rn(x), lo = bf16_rn(x - hi) (16
// significand bits together) and
//     D += A_lo * B_hi + A_hi * B_lo + A_hi * B_hi                        (fp32 accumulate in TMEM)
// which keeps a relative error of ~2^-16 per product (the path's contract is 1e-3 against the reference's fp32
// results) at TWICE the tensor rate and HALF the shared-memory operand bytes of the 3xTF32 kernel
// (gemm_tf32x3.cu, kept for the MN-major / in-kernel-split variants and as the higher-precision option).
//
//   A : [batch, M, K] fp32, K contiguous -- loaded raw by TMA, split into bf16 hi/lo by converter warps
//   B : [batch or 1, N, K] PRE-SPLIT bf16 hi/lo (mpf_split_bf16): weights / per-query mask embedding
//   C : row-major [slab, M, ldc] or transposed [slab, N, ldc]; written by TMA STORES from a shared-memory staging
//       tile, so the LSU never sees the 128-rows-x-16-bytes scatter the 3xTF32 epilogue issues (the timing
//       decomposition in profiles/r1i_gemm_debug_probe.jsonl shows those stores cost 30% of a K=256 GEMM).
//
// One persistent CTA per SM, 16 warps:
//   warp 0      TMA producer   raw A (fp32, SWIZZLE_128B) + B_hi + B_lo (bf16, SWIZZLE_64B) per 32-wide k-block
//   warps 8-11  converters     thread = tile row: 8 x LDS.128 raw -> 4 + 4 x STS.128 bf16 hi / lo (SWIZZLE_64B K-major)
//   warp 1      MMA issuer     2 k-steps x 3 tcgen05.mma.kind::f16 (M=128, N=BN, K=16) per k-block
//   warps 4-7, 12-15  epilogue (two groups, alternate 32-column chunks)
//                              tcgen05.ld -> bias/residual/scale/ReLU/gate -> staging tile -> cp.async.bulk.tensor store
//   warp 2      TMEM allocation (2 accumulators x 256 columns, double-buffered across tiles)
#include "mpf_common.cuh"
#include "sm100_ptx.cuh"
#include "tmap.cuh"

#include <cuda_bf16.h>

#include <cstdlib>
#include <mutex>

namespace mpf {

using namespace ptx;

namespace bf3 {

constexpr int kBM = 128;
constexpr int kBK = 32;                        // k-block: 32 fp32 = 128 B raw rows, 32 bf16 = 64 B operand rows
constexpr int kRawABytes = kBM * kBK * 4;      // 16 KiB
constexpr int kOpABytes = kBM * kBK * 2;       // 8 KiB each for A_hi, A_lo
constexpr int kStagingBytes = kBM * 32 * 4;    // 16 KiB: 128 rows x 32 fp32 (or 32 x 128 transposed)
constexpr int kThreads = 512;
constexpr int kMaxStages = 6;
constexpr int kSmemBudget = 232448;            // 227 KiB opt-in maximum per CTA
constexpr int kTmemCols = 512;

struct Args {
  const float* bias;
  const float* resid;
  long long resid_ld;
  int resid_rows, resid_cols;
  const float* gate;
  long long gate_ld;
  // the same ReLU gate as ONE BIT per element (batch 1, row-major C, N % 32 == 0): word [n / 32][m], bit n % 32.
  // relu_bits_out is written by a GEMM with relu (bit = output > 0); gate_bits replaces `gate` in the backward GEMM:
  // 1/32 of the bytes of reading the activation back (1.4 GB -> 44 MB for the encoder FFN at the bench geometry).
  uint32_t* relu_bits_out;
  const uint32_t* gate_bits;
  float alpha;
  int batch, M, N, K;
  int bn;                                     // N tile (multiple of 32, <= 256)
  int sbufs;                                  // staging tiles per epilogue group (1 or 2)
  int tiles_m, tiles_n;
  int k_splits, k_per_split;
  int stages, stage_bytes;
  int relu, transpose_c, split_out, vec_aux;
  int b_broadcast;                            // B has no batch dimension (shared weight)
  int debug;
  // 3x3 convolution mode (conv_wt > 0): A is a channels-last map [batch, H, W, Cin] behind a 4-D tensor map with
  // box {32 channels, 128 pixels of one row}; an M tile is 128 consecutive pixels of a row, K = 9 * Cin runs over
  // (tap, channel block) and the tap only shifts the box: pixels outside the map are zero-filled by TMA.
  int conv_wt;                                // tiles per image row (ceil(W / 128))
  int conv_cb;                                // channel blocks per tap (Cin / 32)
};

__device__ __forceinline__ void tma_load_3d_hint(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                 int c2) {
  tma_load_3d(smem_dst, m, bar, c0, c1, c2);
}

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void epi_bar(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

// ---- CTA-pair ("cta_group::2") primitives: two CTAs of a cluster on the two SMs of a TPC execute one UMMA with
// M = 256 (each CTA holds 128 rows of A and of the accumulator, and HALF of the B tile), so a CTA moves half the B
// bytes through its shared memory per MMA.  Only the leader (cluster rank 0) issues MMAs; barriers that both CTAs feed
// live in the leader and are reached through the shared::cluster window (rank bit 24 cleared).
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the LEADER's copy of `bar` (works from both CTAs of the pair)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask)
               : "memory");
}
// TMA load into THIS CTA's shared memory whose bytes are counted on the LEADER's barrier
__device__ __forceinline__ void tma_load_3d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                 int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0),
      "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void mma_bf16_ss_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// completion of all prior MMAs of the issuing thread -> arrive on `bar` in BOTH CTAs of the pair
__device__ __forceinline__ void mma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(static_cast<uint16_t>(3))
               : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {   // one warp in EACH CTA
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T, bf16 inputs, fp32 accumulate.  One thread issues.
__device__ __forceinline__ void mma_bf16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// K-major operand tile, SWIZZLE_64B: rows of 64 bytes (32 bf16), 8-row groups of 512 B stacked along M/N.
//   [0,14) start >> 4   [16,30) LBO (unused for swizzled K-major; 1)   [32,46) SBO = 512 >> 4
//   [46,48) version 1   [61,64) layout 4 (SWIZZLE_64B)
__device__ __forceinline__ uint64_t smem_desc_sw64_kmajor(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(512 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(4) << 61;
  return d;
}

// Instruction descriptor, kind::f16: [4,6) D = F32 (1)  [7,10) A = BF16 (1)  [10,13) B = BF16 (1)
//   [15] A major K  [16] B major K  [17,23) N >> 3  [24,29) M >> 4
__device__ __forceinline__ uint32_t idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {      // a -> low half (lower address)
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo_f32(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi_f32(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }
__device__ __forceinline__ float rn_tf32(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

// kPair: the CTA-pair variant (launched as clusters of two CTAs).  A pair walks "pair tiles" of 256 rows: rank r
// loads / converts / drains M tile 2 * pair_m + r and loads the B rows [r * BN/2, (r + 1) * BN/2) of the N tile.
template <bool kPair>
__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16x3_kernel_t(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi,
                     const __grid_constant__ CUtensorMap tmBlo, const __grid_constant__ CUtensorMap tmC,
                     const __grid_constant__ CUtensorMap tmClo, const Args g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int S = g.stages;
  uint8_t* staging = smem + S * g.stage_bytes;                 // 2 groups x sbufs x 16 KiB
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + 2 * g.sbufs * kStagingBytes);
  uint64_t* full = bars;                         // [S] TMA bytes landed
  uint64_t* conv = bars + kMaxStages;            // [S] A_hi / A_lo written
  uint64_t* empty = bars + 2 * kMaxStages;       // [S] MMAs reading the stage retired
  uint64_t* tfull = bars + 3 * kMaxStages;       // [2] accumulator complete
  uint64_t* tempty = bars + 3 * kMaxStages + 2;  // [2] accumulator drained by the epilogue
  uint64_t* fullA = bars + 3 * kMaxStages + 4;   // [S] pair mode: this CTA's raw A tile landed (full = B of both CTAs)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 * kMaxStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int BN = g.bn;
  const int rank = kPair ? static_cast<int>(cluster_ctarank()) : 0;
  const int b_rows = kPair ? BN / 2 : BN;        // B rows this CTA holds
  const int b_bytes = b_rows * kBK * 2;          // one of B_hi / B_lo per stage
  const int worker = kPair ? blockIdx.x >> 1 : blockIdx.x;        // index of this CTA (pair) in the persistent loop
  const int workers = kPair ? gridDim.x >> 1 : gridDim.x;
  const int tiles_m = kPair ? (g.tiles_m + 1) / 2 : g.tiles_m;     // pair tiles along M

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmBhi);
    prefetch_tmap(&tmBlo);
    prefetch_tmap(&tmC);
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&conv[s], kPair ? 8 : 4);          // pair mode: the converter warps of BOTH CTAs arrive at the leader
      mbar_init(&empty[s], 1);
      mbar_init(&fullA[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], kPair ? 16 : 8);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    if (kPair) tmem_alloc_pair(tmem_slot, kTmemCols); else tmem_alloc(tmem_slot, kTmemCols);
  }
  tc_fence_before();
  if (kPair) cluster_sync_all(); else __syncthreads();   // barriers of BOTH CTAs exist before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_tiles = g.batch * g.k_splits * tiles_m * g.tiles_n;
  const int kblocks = g.k_per_split / kBK;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const bool ldA = !(g.debug & 32), ldB = !(g.debug & 16);
      for (int tile = worker; tile < num_tiles; tile += workers) {
        const int n_t = tile % g.tiles_n;
        int rest = tile / g.tiles_n;
        const int m_t = (rest % tiles_m) * (kPair ? 2 : 1) + rank;
        rest /= tiles_m;
        const int ks = rest % g.k_splits;
        const int b = rest / g.k_splits;
        const int kb0 = ks * kblocks;
        const int bb = g.b_broadcast ? 0 : b;
        for (int kbi = 0; kbi < kblocks; ++kbi) {
          const int kb = kb0 + kbi;
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* st = smem + stage * g.stage_bytes;
          if (kPair) {
            // own A tile -> own barrier; the B halves of both CTAs are counted on the leader's `full`
            mbar_arrive_expect_tx(&fullA[stage], kRawABytes);
            if (rank == 0) mbar_arrive_expect_tx(&full[stage], 4 * b_bytes);
          } else {
            mbar_arrive_expect_tx(&full[stage], (ldA ? kRawABytes : 0) + (ldB ? 2 * b_bytes : 0));
          }
          uint64_t* const barA = kPair ? &fullA[stage] : &full[stage];
          if (ldA) {
            if (g.conv_wt) {
              const int y = m_t / g.conv_wt, x0 = (m_t - y * g.conv_wt) * kBM;
              const int tap = kb / g.conv_cb, c0 = (kb - tap * g.conv_cb) * kBK;
              const int ty = tap / 3;
              tma_load_4d(st, &tmA, barA, c0, x0 + (tap - 3 * ty) - 1, y + ty - 1, b);
            } else {
              tma_load_3d(st, &tmA, barA, kb * kBK, m_t * kBM, b);
            }
          }
          if (ldB) {
            uint8_t* bs = st + kRawABytes;
            if (kPair) {
              tma_load_3d_pair(bs, &tmBhi, &full[stage], kb * kBK, n_t * BN + rank * b_rows, bb);
              tma_load_3d_pair(bs + b_bytes, &tmBlo, &full[stage], kb * kBK, n_t * BN + rank * b_rows, bb);
            } else {
              tma_load_3d(bs, &tmBhi, &full[stage], kb * kBK, n_t * BN, bb);
              tma_load_3d(bs + b_bytes, &tmBlo, &full[stage], kb * kBK, n_t * BN, bb);
            }
          }
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0 && rank == 0) {                  // pair mode: the leader issues for both CTAs
      const uint32_t idesc = idesc_bf16(kPair ? 2 * kBM : kBM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = worker; tile < num_tiles; tile += workers) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * 256);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full[stage], phase);
          mbar_wait(&conv[stage], phase);
          tc_fence_after();
          const uint32_t st = smem_u32(smem + stage * g.stage_bytes);
          const uint32_t a_hi = st;                  // converted IN PLACE over the raw fp32 tile
          const uint32_t a_lo = a_hi + kOpABytes;
          const uint32_t b_hi = a_lo + kOpABytes;
          const uint32_t b_lo = b_hi + b_bytes;
#pragma unroll
          for (int k = 0; k < 2; ++k) {              // 16 bf16 = 32 bytes per MMA K step inside the 64-B swizzle row
            const uint64_t dah = smem_desc_sw64_kmajor(a_hi + k * 32);
            const uint64_t dal = smem_desc_sw64_kmajor(a_lo + k * 32);
            const uint64_t dbh = smem_desc_sw64_kmajor(b_hi + k * 32);
            const uint64_t dbl = smem_desc_sw64_kmajor(b_lo + k * 32);
            if (kPair) {
              mma_bf16_ss_pair(d_tmem, dal, dbh, idesc, (kb | k) ? 1u : 0u);
              mma_bf16_ss_pair(d_tmem, dah, dbl, idesc, 1u);
              mma_bf16_ss_pair(d_tmem, dah, dbh, idesc, 1u);
              continue;
            }
            if (g.debug & 8) {
              mma_bf16_ss(d_tmem, dah, dbh, idesc, (kb | k) ? 1u : 0u);
              continue;
            }
            mma_bf16_ss(d_tmem, dal, dbh, idesc, (kb | k) ? 1u : 0u);
            mma_bf16_ss(d_tmem, dah, dbl, idesc, 1u);
            mma_bf16_ss(d_tmem, dah, dbh, idesc, 1u);
          }
          if (kPair) {
            mma_commit_pair(&empty[stage]);          // frees the stage in BOTH CTAs
            if (kb == kblocks - 1) mma_commit_pair(&tfull[acc]);
          } else {
            mma_commit(&empty[stage]);
            if (kb == kblocks - 1) mma_commit(&tfull[acc]);
          }
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 8 && warp < 12) {
    // ================= converters (128 threads, thread = tile row) =================
    const int r = threadIdx.x - 256;
    const int rsw = r & 7;                 // SWIZZLE_128B: 16-byte chunk index ^ (row % 8)
    const int osw = (r >> 1) & 3;          // SWIZZLE_64B:  16-byte chunk index ^ ((row / 2) % 4)
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = worker; tile < num_tiles; tile += workers) {
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(kPair ? &fullA[stage] : &full[stage], phase);
        uint8_t* st = smem + stage * g.stage_bytes;
        if (!(g.debug & 4)) {
          // The bf16 halves (2 x 8 KiB) replace the raw fp32 tile (16 KiB) IN PLACE: every thread reads its raw row
          // into registers, the 128 converter threads meet at a named barrier, then the halves are written.  A
          // stage is 48 KiB instead of 64 KiB at BN = 256, i.e. FOUR stages in flight instead of three: with three
          // the main loop waited for loads (probe: 0.70 ms with two stages, 0.56 ms with three, profiles/r2m_*).
          const uint8_t* raw = st + r * 128;
          uint8_t* ohi = st + r * 64;
          uint8_t* olo = ohi + kOpABytes;
          float4 x[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) x[c] = *reinterpret_cast<const float4*>(raw + ((c ^ rsw) << 4));
          epi_bar(3);
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const float4 u = x[2 * p], v = x[2 * p + 1];
            uint4 h, l;
            h.x = pack_bf16x2(u.x, u.y); h.y = pack_bf16x2(u.z, u.w);
            h.z = pack_bf16x2(v.x, v.y); h.w = pack_bf16x2(v.z, v.w);
            l.x = pack_bf16x2(u.x - bf16_lo_f32(h.x), u.y - bf16_hi_f32(h.x));
            l.y = pack_bf16x2(u.z - bf16_lo_f32(h.y), u.w - bf16_hi_f32(h.y));
            l.z = pack_bf16x2(v.x - bf16_lo_f32(h.z), v.y - bf16_hi_f32(h.z));
            l.w = pack_bf16x2(v.z - bf16_lo_f32(h.w), v.w - bf16_hi_f32(h.w));
            const int off = (p ^ osw) << 4;
            *reinterpret_cast<uint4*>(ohi + off) = h;
            *reinterpret_cast<uint4*>(olo + off) = l;
          }
          fence_proxy_async_smem();
        }
        __syncwarp();
        if (lane == 0) {
          if (kPair) mbar_arrive_leader(&conv[stage]); else mbar_arrive(&conv[stage]);
        }
        if (++stage == S) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue: two groups of 4 warps (4..7 and 12..15), warp % 4 -> TMEM lane group =========
    // Group eg drains the 32-column chunks c = eg, eg + 2, ... of every tile through its own staging buffer and
    // named barrier, so two chunks are in flight (tcgen05.ld / math / staging / TMA store) at any time.
    const int eg = warp >= 12 ? 1 : 0;
    const int ew = warp & 3;
    const int trow = ew * 32 + lane;               // row of the tile this thread drains
    const bool issuer = (ew == 0 && lane == 0);
    const int bar_id = 1 + eg;
    // g.sbufs staging tiles per group, used in turn: with two, the TMA store of a chunk still reads its tile while
    // the group fills the other one (the wait below then only covers the store before the previous one)
    uint8_t* const sbuf0 = staging + eg * g.sbufs * kStagingBytes;
    uint8_t* sbuf = sbuf0;
    int scur = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    const int nchunks = BN / 32;
    for (int tile = worker; tile < num_tiles; tile += workers) {
      const int n_t = tile % g.tiles_n;
      int rest = tile / g.tiles_n;
      const int m_t = (rest % tiles_m) * (kPair ? 2 : 1) + rank;
      const int slab = rest / tiles_m;              // = batch * k_splits + split: index of the output slab
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const int row = m_t * kBM + trow;
      const bool row_ok = row < g.M && m_t < g.tiles_m;
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + static_cast<uint32_t>(acc * 256);
      const float* grow = (g.gate != nullptr && row_ok) ? g.gate + static_cast<long long>(row) * g.gate_ld : nullptr;
      const float* rrow = nullptr;
      if (g.resid != nullptr && row_ok)
        rrow = g.resid + static_cast<long long>(g.resid_rows > 0 ? row % g.resid_rows : row) * g.resid_ld;
      int live = 0;                                 // chunks of this tile that hold output columns
      for (int c = 0; c < nchunks; ++c)
        if (n_t * BN + c * 32 < g.N) live = c + 1;
      if (g.debug & 2) live = 0;
      const int my_last = live - 1 - ((live - 1 - eg) & 1);     // last chunk of this group (< eg: none)
      if (live <= eg) {                             // nothing to drain for this group: release the accumulator
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (kPair) mbar_arrive_leader(&tempty[acc]); else mbar_arrive(&tempty[acc]);
        }
      }
#pragma unroll 1
      for (int c = eg; c < live; c += 2) {
        const int n0 = n_t * BN + c * 32;
        uint32_t v[32];
        tmem_ld_32x32(t_addr + c * 32, v);
        tmem_ld_wait();
        if (c == my_last) {                         // this group's part of the accumulator is read
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
          if (kPair) mbar_arrive_leader(&tempty[acc]); else mbar_arrive(&tempty[acc]);
        }
        }
        if (g.debug & 64) continue;
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        const bool fullc = n0 + 32 <= g.N;
        if (g.bias != nullptr) {
          if (fullc && g.vec_aux) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 t = __ldg(reinterpret_cast<const float4*>(g.bias + n0 + j));
              f[j] += t.x; f[j + 1] += t.y; f[j + 2] += t.z; f[j + 3] += t.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + j < g.N) f[j] += __ldg(g.bias + n0 + j);
          }
        }
        if (rrow != nullptr) {
          if (n0 + 32 <= g.resid_cols && g.vec_aux) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 t = __ldg(reinterpret_cast<const float4*>(rrow + n0 + j));
              f[j] += t.x; f[j + 1] += t.y; f[j + 2] += t.z; f[j + 3] += t.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + j < g.resid_cols) f[j] += __ldg(rrow + n0 + j);
          }
        }
        if (g.alpha != 1.0f) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] *= g.alpha;
        }
        if (g.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
          if (g.relu_bits_out != nullptr && row_ok) {
            uint32_t w = 0u;
#pragma unroll
            for (int j = 0; j < 32; ++j) w |= (f[j] > 0.f ? 1u : 0u) << j;
            g.relu_bits_out[static_cast<long long>(n0 >> 5) * g.M + row] = w;
          }
        }
        if (g.gate_bits != nullptr && row_ok) {
          const uint32_t w = __ldg(g.gate_bits + static_cast<long long>(n0 >> 5) * g.M + row);
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (!((w >> j) & 1u)) f[j] = 0.f;
        }
        if (grow != nullptr) {
          if (fullc && g.vec_aux) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 t = __ldg(reinterpret_cast<const float4*>(grow + n0 + j));
              if (!(t.x > 0.f)) f[j] = 0.f;
              if (!(t.y > 0.f)) f[j + 1] = 0.f;
              if (!(t.z > 0.f)) f[j + 2] = 0.f;
              if (!(t.w > 0.f)) f[j + 3] = 0.f;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + j < g.N && !(__ldg(grow + n0 + j) > 0.f)) f[j] = 0.f;
          }
        }
        if (g.debug & 1) continue;

        // ---- registers -> staging tile -> TMA store (clipped at M / N by the tensor map) ----
        // row-major C: staging is [128 rows][32 fp32] with the 128-byte swizzle (16-byte chunk ^ row % 8);
        // transposed C: staging is [32 n-rows][128 m] plain (a warp writes 128 contiguous bytes).
        auto put = [&](const float (&val)[32]) {
          if (g.transpose_c) {
            float* bt = reinterpret_cast<float*>(sbuf) + trow;
#pragma unroll
            for (int j = 0; j < 32; ++j) bt[j * kBM] = val[j];
          } else {
            uint8_t* br = sbuf + trow * 128;
            const int sw = trow & 7;
#pragma unroll
            for (int q = 0; q < 8; ++q)
              *reinterpret_cast<float4*>(br + ((q ^ sw) << 4)) =
                  make_float4(val[4 * q], val[4 * q + 1], val[4 * q + 2], val[4 * q + 3]);
          }
        };
        auto flush = [&](const CUtensorMap* tm) {   // staging tile of this group -> global
          fence_proxy_async_smem();
          epi_bar(bar_id);
          if (issuer) {
            if (g.debug & 128) {
            } else if (g.conv_wt) {               // 4-D map [batch, H, W, N]: the row segment is clipped at W
              const int y = m_t / g.conv_wt;
              tma_store_4d(tm, sbuf, n0, (m_t - y * g.conv_wt) * kBM, y, slab);
            } else if (g.transpose_c) {
              tma_store_3d(tm, sbuf, m_t * kBM, n0, slab);
            } else {
              tma_store_3d(tm, sbuf, n0, m_t * kBM, slab);
            }
            bulk_commit();
          }
        };
        auto next_tile = [&]() {                    // the store that last used this staging tile has read it
          sbuf = sbuf0 + scur * kStagingBytes;
          if (issuer) {
            if (g.sbufs == 2) bulk_wait_read<1>(); else bulk_wait_read<0>();
          }
          if (g.sbufs == 2) scur ^= 1;
          epi_bar(bar_id);
        };
        next_tile();
        if (!g.split_out) {
          put(f);
          flush(&tmC);
        } else {                       // emit the result pre-split (TF32 halves) for the attention kernels
          float lo[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float h = rn_tf32(f[j]);
            lo[j] = rn_tf32(f[j] - h);
            f[j] = h;
          }
          put(f);
          flush(&tmC);
          next_tile();
          put(lo);
          flush(&tmClo);
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (issuer) bulk_wait_all();
  }

  tc_fence_before();
  if (kPair) cluster_sync_all(); else __syncthreads();     // the peer may still be signalling this CTA's barriers
  if (warp == 2) {
    tc_fence_after();
    if (kPair) tmem_dealloc_pair(tmem_base, kTmemCols); else tmem_dealloc(tmem_base, kTmemCols);
  }
}

__global__ void __launch_bounds__(256)
split_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                  long long n) {
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * 256) {
    const float v = x[i];
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

// x [batch, R, Cc] fp32 -> hi, lo [batch, Cc, R] bf16: the split of mpf_split_bf16 fused with a transposition, for a
// big activation that a later product needs as its K-major B operand (mask_features in the batched dE of the
// prediction heads: reduction over the H*W pixels).  64 x 64 tile through shared memory; 16-byte loads along Cc,
// 8-byte (4 x bf16) stores along R.  grid (ceil(R/64), ceil(Cc/64), batch); R % 4 == 0, Cc % 4 == 0.
__global__ void __launch_bounds__(256)
transpose_split_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi,
                            __nv_bfloat16* __restrict__ lo, long long R, int Cc) {
  __shared__ float tile[64 * 65];
  const long long r0 = static_cast<long long>(blockIdx.x) * 64;
  const int c0 = blockIdx.y * 64, b = blockIdx.z, t = threadIdx.x;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = t + 256 * k;
    const int r = i >> 4, c4 = i & 15;
    if (r0 + r < R && c0 + 4 * c4 < Cc) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + (static_cast<long long>(b) * R + r0 + r) * Cc + c0) + c4);
      float* s = tile + (4 * c4) * 65 + r;
      s[0] = v.x; s[65] = v.y; s[130] = v.z; s[195] = v.w;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = t + 256 * k;
    const int c = i >> 4, r4 = i & 15;
    if (c0 + c < Cc && r0 + 4 * r4 < R) {
      const float* s = tile + c * 65 + 4 * r4;
      __nv_bfloat16 h[4], l[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        h[e] = __float2bfloat16_rn(s[e]);
        l[e] = __float2bfloat16_rn(s[e] - __bfloat162float(h[e]));
      }
      const long long o = (static_cast<long long>(b) * Cc + c0 + c) * R + r0 + 4 * r4;
      *reinterpret_cast<uint2*>(hi + o) = *reinterpret_cast<const uint2*>(h);
      *reinterpret_cast<uint2*>(lo + o) = *reinterpret_cast<const uint2*>(l);
    }
  }
}

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
  static thread_local bool ctx_bound = false;   // see make_tmap_f32_3d
  if (!ctx_bound) {
    cudaFree(nullptr);
    ctx_bound = true;
  }
  return enc;
}

// 3-D tensor [d2, d1, d0] (d0 contiguous) of `esize`-byte elements; strides in elements.
static int make_tmap_3d(CUtensorMap* m, CUtensorMapDataType dt, int esize, const void* base, long long d0,
                        long long d1, long long d2, long long ld1, long long ld2, int box0, int box1,
                        CUtensorMapSwizzle sw, const char* what) {
  EncodeFn enc = encoder();
  if (enc == nullptr) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return MPF_ERR_UNSUPPORTED;
  }
  if ((ld1 * esize) % 16 != 0 || (ld2 * esize) % 16 != 0 || !aligned16(base)) {
    set_error("gemm_bf16x3: %s needs a 16-byte aligned base and byte strides that are multiples of 16", what);
    return MPF_ERR_BAD_ARG;
  }
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(d0), static_cast<cuuint64_t>(d1), static_cast<cuuint64_t>(d2)};
  if (ld2 < d1 * ld1) ld2 = d1 * ld1;
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld1) * esize, static_cast<cuuint64_t>(ld2) * esize};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(box0), static_cast<cuuint32_t>(box1), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, dt, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(%s) failed (CUresult %d): dims=(%lld,%lld,%lld) ld=(%lld,%lld) box=(%d,%d)", what,
              static_cast<int>(r), d0, d1, d2, ld1, ld2, box0, box1);
    return MPF_ERR_BAD_ARG;
  }
  return MPF_OK;
}

// channels-last fp32 map [d3, d2, d1, d0 = channels] (dense), box {box0 channels, box1 pixels of one row, 1, 1},
// SWIZZLE_128B (box0 * 4 == 128); coordinates outside the map read as zero.
static int make_tmap_f32_4d(CUtensorMap* m, const float* base, long long d0, long long d1, long long d2, long long d3,
                            int box0, int box1, const char* what) {
  EncodeFn enc = encoder();
  if (enc == nullptr) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return MPF_ERR_UNSUPPORTED;
  }
  if (box0 * 4 != 128 || d0 % 4 != 0 || !aligned16(base)) {
    set_error("gemm_bf16x3: %s needs a 16-byte aligned base, channels %% 4 == 0 and a 128-byte box row", what);
    return MPF_ERR_BAD_ARG;
  }
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(d0), static_cast<cuuint64_t>(d1), static_cast<cuuint64_t>(d2),
                        static_cast<cuuint64_t>(d3)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(d0) * 4, static_cast<cuuint64_t>(d0) * d1 * 4,
                           static_cast<cuuint64_t>(d0) * d1 * d2 * 4};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(box0), static_cast<cuuint32_t>(box1), 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(%s, 4-D) failed (CUresult %d): dims=(%lld,%lld,%lld,%lld)", what,
              static_cast<int>(r), d0, d1, d2, d3);
    return MPF_ERR_BAD_ARG;
  }
  return MPF_OK;
}

// 0 (default): never use CTA pairs, 1: when a launch has at least one full wave of pair tiles, 2: whenever possible.
// MEASURED (round 2, profiles/r2j_*): the pair kernel is bit-identical to the single-CTA kernel and SLOWER -- encoder
// FFN1 0.73 vs 0.55 ms, FFN2 0.78 vs 0.51 ms, 3x3 convolution 4.07 vs 3.01 ms, whole step 162.1 vs 152.3 ms.  The
// single-CTA main loop already runs at 77 % (K = 1024) to 90 % (K = 2304) of the sustained bf16 rate counting the three
// MMAs per product, so shared-memory operand traffic was not the limiter round 1's analysis took it for; with
// cta_group::2 every SM reads half of each B tile from its PEER's shared memory, three times per product in the split
// arithmetic, and the leader's issue thread waits on cross-SM barriers.  Kept as an option and as a measured answer.
static int g_pair_mode = [] { const char* e = getenv("MPF_GEMM_PAIR"); return e ? atoi(e) : 0; }();

// measurement knobs (environment, read once): staging tiles per epilogue group and a cap on the pipeline stages
static int g_staging_tiles = [] { const char* e = getenv("MPF_GEMM_SBUFS"); return e && atoi(e) == 2 ? 2 : 1; }();
static int g_max_stages = [] { const char* e = getenv("MPF_GEMM_STAGES"); return e ? atoi(e) : 0; }();
static int g_small_bn = [] { const char* e = getenv("MPF_GEMM_SMALL_BN"); return e ? atoi(e) : 2; }();

static int pick_bn(int N) {
  if (N <= 64) return 64;
  for (int bn = 256; bn >= 64; bn -= 32) {
    const int padded = (N + bn - 1) / bn * bn;
    if ((padded - N) * 8 <= N) return bn;          // <= 12.5 % of the tile columns wasted
  }
  return 64;
}

}  // namespace bf3
}  // namespace mpf

extern "C" {

int mpf_gemm_bf16x3_set_pair_mode(int mode) {
  const int prev = mpf::bf3::g_pair_mode;
  if (mode >= 0 && mode <= 2) mpf::bf3::g_pair_mode = mode;
  return prev;
}

int mpf_split_bf16(const float* x, uint16_t* hi, uint16_t* lo, long long n, void* stream) {
  mpf::clear_error();
  MPF_REQUIRE(x && hi && lo && n > 0, "split_bf16: bad arguments");
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  mpf::bf3::split_bf16_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo), n);
  mpf::count_launch();
  return mpf::finish_launch("split_bf16");
}

int mpf_transpose_split_bf16(const float* x, uint16_t* hi, uint16_t* lo, int batch, long long R, int Cc, void* stream) {
  using namespace mpf;
  clear_error();
  MPF_REQUIRE(x && hi && lo && batch > 0 && R > 0 && Cc > 0, "transpose_split_bf16: bad arguments");
  MPF_REQUIRE(R % 4 == 0 && Cc % 4 == 0 && aligned16(x) && (reinterpret_cast<uintptr_t>(hi) & 7u) == 0 &&
                  (reinterpret_cast<uintptr_t>(lo) & 7u) == 0,
              "transpose_split_bf16: R and Cc must be multiples of 4, x 16-byte and hi / lo 8-byte aligned");
  MPF_REQUIRE(batch <= 65535 && (Cc + 63) / 64 <= 65535, "transpose_split_bf16: grid too large");
  dim3 grid(static_cast<unsigned>((R + 63) / 64), (Cc + 63) / 64, batch);
  bf3::transpose_split_bf16_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo), R, Cc);
  count_launch();
  return finish_launch("transpose_split_bf16");
}

// conv_H > 0: 3x3 convolution of the channels-last map A [batch, conv_H, conv_W, conv_C] (M = conv_H * conv_W,
// K = 9 * conv_C; lda / a_batch_stride unused)
static int gemm_bf16x3_impl(const float* A, long long lda, long long a_batch_stride, const uint16_t* B_hi,
                            const uint16_t* B_lo, long long ldb, long long b_batch_stride, const float* bias, float* C,
                            float* C_lo, long long ldc, long long c_batch_stride, const float* resid,
                            long long resid_ld, int resid_rows, int resid_cols, const float* gate, long long gate_ld,
                            float alpha, int batch, int M, int N, int K, int k_splits, int relu, int transpose_c,
                            int conv_H, int conv_W, int conv_C, uint32_t* relu_bits_out, const uint32_t* gate_bits,
                            void* stream) {
  using namespace mpf;
  using namespace mpf::bf3;
  clear_error();
  MPF_REQUIRE(A && B_hi && B_lo && C, "gemm_bf16x3: null pointer argument");
  MPF_REQUIRE(batch > 0 && M > 0 && N > 0 && K > 0 && k_splits >= 1, "gemm_bf16x3: dimensions must be positive");
  MPF_REQUIRE(gate == nullptr || (batch == 1 && !transpose_c), "gemm_bf16x3: gate needs batch 1, row-major C");
  MPF_REQUIRE(k_splits == 1 || (!bias && !resid && !relu && !C_lo && !gate && alpha == 1.0f),
              "gemm_bf16x3: split-K produces partial sums; no epilogue operations are allowed");
  MPF_REQUIRE(lda >= K && ldb >= K, "gemm_bf16x3: row stride too small");
  Args g;
  g.bn = pick_bn(N);
  if (const char* force = getenv("MPF_GEMM_BN")) {
    const int f = atoi(force);
    if (f >= 32 && f <= 256 && f % 32 == 0) g.bn = f;
  }
  g.tiles_m = conv_H > 0 ? conv_H * ((conv_W + kBM - 1) / kBM) : (M + kBM - 1) / kBM;
  // A product with a handful of output tiles (the decoder's linears on B * 120 query rows: 4 - 28 M tiles) leaves most
  // SMs idle while each busy one drains an epilogue of BN / 32 chunks after a short main loop: narrower N tiles spread
  // the same work over four times as many CTAs (the A tile is converted once per N tile, which is cheap at this size).
  if (g_small_bn && !getenv("MPF_GEMM_BN") && N % 64 == 0 && N >= 128 &&
      static_cast<long long>(batch) * k_splits * g.tiles_m * ((N + g.bn - 1) / g.bn) * 4 <= sm_count()) {
    g.bn = 64;
    // second tier: 32-column tiles when even those leave 3/4 of the SMs idle (two images: 28.2 -> 27.5 ms per step;
    // MPF_GEMM_SMALL_BN=1 keeps the 64-column tier only, 0 disables both)
    if (g_small_bn >= 2 && static_cast<long long>(batch) * k_splits * g.tiles_m * (N / 64) * 4 <= sm_count()) g.bn = 32;
  }
  g.tiles_n = (N + g.bn - 1) / g.bn;
  MPF_REQUIRE(static_cast<long long>(batch) * k_splits * g.tiles_m * g.tiles_n < (1ll << 31), "gemm_bf16x3: too many tiles");
  const int kblocks_total = (K + kBK - 1) / kBK;
  g.k_splits = k_splits;
  g.k_per_split = (kblocks_total + k_splits - 1) / k_splits * kBK;
  // CTA-pair mode (cta_group::2): worth it when there is a full wave of 256-row pair tiles; each CTA then stages only
  // half of the B tile.  MPF_GEMM_PAIR=0 disables it (A/B measurements).
  const long long pair_tiles = static_cast<long long>(batch) * k_splits * ((g.tiles_m + 1) / 2) * g.tiles_n;
  const bool pair = g_pair_mode != 0 && g.tiles_m >= 2 && sm_count() % 2 == 0 &&
                    (g_pair_mode == 2 || pair_tiles >= sm_count() / 2);
  const int b_rows = pair ? g.bn / 2 : g.bn;
  g.stage_bytes = kRawABytes + 2 * b_rows * kBK * 2;      // A: raw fp32, converted in place to bf16 hi | lo
  g.sbufs = g_staging_tiles;
  int avail = kSmemBudget - 1024 - 2 * g.sbufs * kStagingBytes - 512;
  if (g.sbufs == 2 && avail / g.stage_bytes < 2) {       // not with this stage size: back to one staging tile per group
    g.sbufs = 1;
    avail = kSmemBudget - 1024 - 2 * kStagingBytes - 512;
  }
  g.stages = avail / g.stage_bytes;
  if (g.stages > kMaxStages) g.stages = kMaxStages;
  if (g_max_stages >= 2 && g.stages > g_max_stages) g.stages = g_max_stages;
  MPF_REQUIRE(g.stages >= 2, "gemm_bf16x3: not enough shared memory for two stages");
  const int smem_bytes = g.stages * g.stage_bytes + 2 * g.sbufs * kStagingBytes + 512 + 1024;
  g.b_broadcast = (b_batch_stride == 0) ? 1 : 0;
  const int slabs = batch * k_splits;

  CUtensorMap ta, tbh, tbl, tc, tcl;
  int rc;
  g.conv_wt = g.conv_cb = 0;
  if (conv_H > 0) {
    g.conv_wt = (conv_W + kBM - 1) / kBM;
    g.conv_cb = conv_C / kBK;
    rc = make_tmap_f32_4d(&ta, A, conv_C, conv_W, conv_H, batch, kBK, kBM, "conv input");
  } else {
    rc = make_tmap_3d(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, A, K, M, batch, lda, a_batch_stride, kBK, kBM,
                      CU_TENSOR_MAP_SWIZZLE_128B, "A");
  }
  if (rc) return rc;
  const long long nb = g.b_broadcast ? 1 : batch;
  rc = make_tmap_3d(&tbh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, B_hi, K, N, nb, ldb, b_batch_stride, kBK, b_rows,
                    CU_TENSOR_MAP_SWIZZLE_64B, "B_hi");
  if (rc) return rc;
  rc = make_tmap_3d(&tbl, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, B_lo, K, N, nb, ldb, b_batch_stride, kBK, b_rows,
                    CU_TENSOR_MAP_SWIZZLE_64B, "B_lo");
  if (rc) return rc;
  if (conv_H > 0) {
    MPF_REQUIRE(!transpose_c && !C_lo && k_splits == 1 && !resid && !gate && ldc == N,
                "conv3x3_cl: plain row-major output only");
    rc = make_tmap_f32_4d(&tc, C, N, conv_W, conv_H, batch, 32, kBM, "conv output");
    tcl = tc;
  } else if (transpose_c) {
    rc = make_tmap_3d(&tc, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, C, M, N, slabs, ldc, c_batch_stride, kBM, 32,
                      CU_TENSOR_MAP_SWIZZLE_NONE, "C^T");
    if (rc) return rc;
    tcl = tc;
    if (C_lo) rc = make_tmap_3d(&tcl, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, C_lo, M, N, slabs, ldc, c_batch_stride, kBM,
                                32, CU_TENSOR_MAP_SWIZZLE_NONE, "C_lo^T");
  } else {
    rc = make_tmap_3d(&tc, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, C, N, M, slabs, ldc, c_batch_stride, 32, kBM,
                      CU_TENSOR_MAP_SWIZZLE_128B, "C");
    if (rc) return rc;
    tcl = tc;
    if (C_lo) rc = make_tmap_3d(&tcl, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, C_lo, N, M, slabs, ldc, c_batch_stride, 32,
                                kBM, CU_TENSOR_MAP_SWIZZLE_128B, "C_lo");
  }
  if (rc) return rc;

  g.bias = bias;
  g.resid = resid; g.resid_ld = resid_ld; g.resid_rows = resid_rows;
  g.resid_cols = resid ? (resid_cols > 0 ? resid_cols : N) : 0;
  g.gate = gate; g.gate_ld = gate_ld;
  MPF_REQUIRE((relu_bits_out == nullptr && gate_bits == nullptr) ||
                  (batch == 1 && !transpose_c && k_splits == 1 && N % 32 == 0 && conv_H == 0),
              "gemm_bf16x3: ReLU bit masks need batch 1, row-major C and N %% 32 == 0");
  MPF_REQUIRE(relu_bits_out == nullptr || relu, "gemm_bf16x3: relu_bits_out without relu");
  g.relu_bits_out = relu_bits_out; g.gate_bits = gate_bits;
  g.alpha = alpha;
  g.batch = batch; g.M = M; g.N = N; g.K = K;
  g.relu = relu; g.transpose_c = transpose_c; g.split_out = C_lo ? 1 : 0;
  g.vec_aux = ((bias == nullptr || aligned16(bias)) && (resid == nullptr || (aligned16(resid) && resid_ld % 4 == 0)) &&
               (gate == nullptr || (aligned16(gate) && gate_ld % 4 == 0))) ? 1 : 0;
  g.debug = 0;
  if (const char* dbg = getenv("MPF_GEMM_DEBUG")) g.debug = atoi(dbg);

  static unsigned long long configured_on = 0;
  if (first_use_on_this_device(configured_on)) {
    MPF_CUDA_OK(cudaFuncSetAttribute(gemm_bf16x3_kernel_t<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
    MPF_CUDA_OK(cudaFuncSetAttribute(gemm_bf16x3_kernel_t<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
  }
  if (pair) {
    const int clusters = static_cast<int>(pair_tiles < sm_count() / 2 ? pair_tiles : sm_count() / 2);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = static_cast<cudaStream_t>(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    MPF_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_bf16x3_kernel_t<true>, ta, tbh, tbl, tc, tcl, g));
    count_launch();
    return finish_launch("gemm_bf16x3 (CTA pairs)");
  }
  const long long tiles = static_cast<long long>(slabs) * g.tiles_m * g.tiles_n;
  const int grid = static_cast<int>(tiles < sm_count() ? tiles : sm_count());
  gemm_bf16x3_kernel_t<false><<<grid, kThreads, smem_bytes, static_cast<cudaStream_t>(stream)>>>(ta, tbh, tbl, tc, tcl, g);
  count_launch();
  return finish_launch("gemm_bf16x3");
}

int mpf_gemm_bf16x3(const float* A, long long lda, long long a_batch_stride, const uint16_t* B_hi,
                    const uint16_t* B_lo, long long ldb, long long b_batch_stride, const float* bias, float* C,
                    float* C_lo, long long ldc, long long c_batch_stride, const float* resid, long long resid_ld,
                    int resid_rows, int resid_cols, const float* gate, long long gate_ld, float alpha, int batch,
                    int M, int N, int K, int k_splits, int relu, int transpose_c, void* stream) {
  return gemm_bf16x3_impl(A, lda, a_batch_stride, B_hi, B_lo, ldb, b_batch_stride, bias, C, C_lo, ldc, c_batch_stride,
                          resid, resid_ld, resid_rows, resid_cols, gate, gate_ld, alpha, batch, M, N, K, k_splits, relu,
                          transpose_c, 0, 0, 0, nullptr, nullptr, stream);
}

int mpf_gemm_bf16x3_relubits(const float* A, long long lda, const uint16_t* B_hi, const uint16_t* B_lo, long long ldb,
                             const float* bias, float* C, long long ldc, const float* resid, long long resid_ld,
                             int M, int N, int K, int relu, uint32_t* relu_bits_out, const uint32_t* gate_bits,
                             void* stream) {
  return gemm_bf16x3_impl(A, lda, static_cast<long long>(M) * lda, B_hi, B_lo, ldb, 0, bias, C, nullptr, ldc,
                          static_cast<long long>(M) * ldc, resid, resid_ld, 0, 0, nullptr, 0, 1.0f, 1, M, N, K, 1, relu, 0, 0,
                          0, 0, relu_bits_out, gate_bits, stream);
}

int mpf_conv3x3_cl_bf16x3(const float* x, const uint16_t* w_hi, const uint16_t* w_lo, const float* bias, float* y,
                          int batch, int H, int W, int Cin, int Cout, int relu, void* stream) {
  using namespace mpf;
  clear_error();
  MPF_REQUIRE(x && w_hi && w_lo && y, "conv3x3_cl: null pointer argument");
  MPF_REQUIRE(batch > 0 && H > 0 && W > 0 && Cin > 0 && Cin % 32 == 0 && Cout > 0 && Cout % 4 == 0,
              "conv3x3_cl: needs Cin %% 32 == 0, Cout %% 4 == 0 (H=%d W=%d Cin=%d Cout=%d)", H, W, Cin, Cout);
  MPF_REQUIRE(static_cast<long long>(H) * W < (1ll << 31) / 2, "conv3x3_cl: map too large");
  const int M = H * W, K = 9 * Cin;
  return gemm_bf16x3_impl(x, K, 0, w_hi, w_lo, K, 0, bias, y, nullptr, Cout, static_cast<long long>(M) * Cout, nullptr, 0,
                          0, 0, nullptr, 0, 1.0f, batch, M, Cout, K, 1, relu, 0, H, W, Cin, nullptr, nullptr, stream);
}

}  // extern "C"

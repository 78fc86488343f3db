// C[b] = A[b]^T * B[b]  ("TN" GEMM, reduction over the leading token dimension) in bf16x3 split arithmetic:
//   A : [batch, T, M] fp32, M contiguous      B : [batch, T, N] fp32, N contiguous      C : [batch*k_splits, M, N]
// This is the weight gradient of an nn.Linear (dW = dY^T X, T = tokens) and dF = dOut^T E of the mask-logit einsum:
// both operands are big activations stored "MN-major" for this product, so BOTH are loaded raw (fp32) by TMA and
// split into bf16 hi / lo halves in shared memory by converter warps, directly in the tensor core's MN-major
// SWIZZLE_128B operand layout -- no transposed copies, no pre-split pass over HBM.
//   D += A_lo*B_hi + A_hi*B_lo + A_hi*B_hi  (tcgen05.mma.kind::f16, fp32 accumulate in TMEM; ~2^-16 per product).
//
// Pipeline (one persistent CTA per SM, 16 warps), ONE ring of slots that hold a 32-token block first raw, then split:
//   warp 0       TMA producer    [32 tokens x 32 columns] fp32 boxes (SWIZZLE_128B), 4 for A, BN/32 for B, per slot
//   warps 8-15   converters      read the raw boxes into registers, meet at a named barrier, write the bf16 hi/lo
//                                halves IN PLACE (same bytes: MN-major SW128, rows = token, 64 columns per 128-byte
//                                row, 4 KiB per 64-column group).  Round 2: with separate raw / operand rings only
//                                two loads were in flight per SM and the main loop waited for them (~1200 cycles
//                                per block against 768 of MMA); in place, 48 KiB slots give four blocks in flight.
//   warp 1       MMA issuer      2 k-steps x 3 MMAs (M=128, N=BN, K=16) per 32-token block; commit frees the slot
//   warps 4-7    epilogue        tcgen05.ld -> staging tile -> TMA store (split-K partial slabs, summed by the caller)
//   warp 2       TMEM allocation
#include "mpf_common.cuh"
#include "sm100_ptx.cuh"
#include "tmap.cuh"

#include <cuda_bf16.h>

#include <cstdlib>
#include <mutex>

namespace mpf {

using namespace ptx;

namespace bf3tn {

constexpr int kBM = 128;
constexpr int kBK = 32;                        // tokens per k-block
constexpr int kBoxBytes = 32 * 32 * 4;         // one raw box: 32 tokens x 32 columns fp32 = 4 KiB
constexpr int kGroupBytes = kBK * 128;         // one operand group: 32 tokens x 64 bf16 columns = 4 KiB
constexpr int kStagingBytes = kBM * 32 * 4;
constexpr int kThreads = 512;
constexpr int kConvThreads = 256;
constexpr int kSmemBudget = 232448;
constexpr int kTmemCols = 512;
constexpr int kMaxRing = 6;                    // slots of the ring (as many as fit: 4 at BN = 256)

struct Args {
  int batch, M, N, T;
  int bn;                                      // 64, 128, 192 or 256
  int tiles_m, tiles_n;
  int k_splits, k_per_split;                   // tokens per split (multiple of 32)
  int raw_bytes, op_bytes;                     // per ring slot (equal: the split halves replace the raw boxes)
  int ring;                                    // slots
  int debug;
  int accumulate;                              // 1: C += product (TMA reduce-add store) instead of C = product
  long long colsum_ld;                         // slab stride of colsum (floats)
  float* colsum;                               // optional [slabs, M]: column sums of A over the slab's tokens (the bias
                                               // gradient that goes with a weight gradient), written by the converters
  // weight gradient of a 3x3 convolution over channels-last maps (conv_W > 0): the reduction index is the pixel
  // (b, y, x), A = dY tokens [T, Cout]; B is the INPUT map [batch, H, W, Cin] behind a 4-D tensor map and the N tile
  // selects (tap, channel range): the tap only shifts the 32-pixel box, pixels outside the map read as zero.
  int conv_H, conv_ws;                         // H and row segments of 32 pixels per image row (ceil(W / 32)):
                                               // k-block index = (image * H + y) * conv_ws + segment
  int conv_nt;                                 // N tiles per tap (Cin / bn)
};

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// same box, but added into global memory by the TMA unit (f32 add taken from the tensor map's data type)
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ void conv_bar() { asm volatile("bar.sync 2, 256;" ::: "memory"); }

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

// MN-major operand, 16-bit elements, SWIZZLE_128B: rows of 128 bytes = 64 consecutive M (or N) elements of ONE
// token, consecutive tokens in consecutive rows (8-row atoms of 1 KiB), 16-byte chunks XOR-ed with (row % 8).
//   LBO = byte stride between 64-element MN groups (4096), SBO = byte stride between 8-token groups (1024).
// One MMA (K = 16) consumes two 8-token groups; the next K step starts 2048 bytes further.
__device__ __forceinline__ uint64_t smem_desc_sw128_mnmajor16(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(kGroupBytes >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// kind::f16, BF16 x BF16 -> F32, both operands MN-major (bits 15, 16)
__device__ __forceinline__ uint32_t idesc_bf16_mn(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo_f32(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi_f32(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }

__device__ __forceinline__ void split8(const float4 u, const float4 v, uint4& h, uint4& l) {
  h.x = pack_bf16x2(u.x, u.y); h.y = pack_bf16x2(u.z, u.w);
  h.z = pack_bf16x2(v.x, v.y); h.w = pack_bf16x2(v.z, v.w);
  l.x = pack_bf16x2(u.x - bf16_lo_f32(h.x), u.y - bf16_hi_f32(h.x));
  l.y = pack_bf16x2(u.z - bf16_lo_f32(h.y), u.w - bf16_hi_f32(h.y));
  l.z = pack_bf16x2(v.x - bf16_lo_f32(h.z), v.y - bf16_hi_f32(h.z));
  l.w = pack_bf16x2(v.z - bf16_lo_f32(h.w), v.w - bf16_hi_f32(h.w));
}

__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16x3_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmC, const Args g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int kRing = g.ring;
  uint8_t* raw_ring = smem;
  uint8_t* op_ring = smem;                                     // the same slots: converted in place
  uint8_t* staging = smem + kRing * g.raw_bytes;               // 2 x 16 KiB
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + 2 * kStagingBytes);
  uint64_t* raw_full = bars;                 // [ring] TMA bytes landed
  uint64_t* op_full = bars + kMaxRing;       // [ring] operand halves written (8 warps)
  uint64_t* op_empty = bars + 2 * kMaxRing;  // [ring] MMAs reading the slot retired
  uint64_t* tfull = bars + 3 * kMaxRing;     // [2]
  uint64_t* tempty = bars + 3 * kMaxRing + 2;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * kMaxRing + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int BN = g.bn;
  const int b_boxes = BN / 32;
  const int a_op_bytes = 2 * kGroupBytes;            // 128 columns = 2 groups, per half
  const int b_op_bytes = (BN / 64) * kGroupBytes;    // per half

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmC);
    for (int s = 0; s < kRing; ++s) {
      mbar_init(&raw_full[s], 1);
      mbar_init(&op_full[s], kConvThreads / 32);
      mbar_init(&op_empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_tiles = g.batch * g.k_splits * g.tiles_m * g.tiles_n;
  const int kblocks = g.k_per_split / kBK;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int slot = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n_t = tile % g.tiles_n;
        int rest = tile / g.tiles_n;
        const int m_t = rest % g.tiles_m;
        rest /= g.tiles_m;
        const int ks = rest % g.k_splits;
        const int b = rest / g.k_splits;
        for (int kbi = 0; kbi < kblocks; ++kbi) {
          const int t0 = ks * g.k_per_split + kbi * kBK;
          mbar_wait(&op_empty[slot], phase ^ 1);
          uint8_t* st = raw_ring + slot * g.raw_bytes;
          mbar_arrive_expect_tx(&raw_full[slot], (4 + b_boxes) * kBoxBytes);
          if (g.conv_ws) {
            // both operands through 4-D maps over [image, y, x, channel]: dY at the segment, the input shifted by the tap
            const int kbg = t0 / kBK;
            const int seg = kbg % g.conv_ws, row = kbg / g.conv_ws;
            const int y = row % g.conv_H, img = row / g.conv_H;
            const int x0 = seg * kBK;
            const int tap = n_t / g.conv_nt, c0 = (n_t - tap * g.conv_nt) * BN;
            const int ty = tap / 3;
#pragma unroll
            for (int i = 0; i < 4; ++i)
              tma_load_4d(st + i * kBoxBytes, &tmA, &raw_full[slot], m_t * kBM + i * 32, x0, y, img);
            for (int j = 0; j < b_boxes; ++j)
              tma_load_4d(st + (4 + j) * kBoxBytes, &tmB, &raw_full[slot], c0 + j * 32, x0 + (tap - 3 * ty) - 1,
                          y + ty - 1, img);
            if (++slot == kRing) { slot = 0; phase ^= 1; }
            continue;
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) tma_load_3d(st + i * kBoxBytes, &tmA, &raw_full[slot], m_t * kBM + i * 32, t0, b);
          {
            for (int j = 0; j < b_boxes; ++j)
              tma_load_3d(st + (4 + j) * kBoxBytes, &tmB, &raw_full[slot], n_t * BN + j * 32, t0, b);
          }
          if (++slot == kRing) { slot = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = idesc_bf16_mn(kBM, BN);
      int slot = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * 256);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&op_full[slot], phase);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(op_ring + slot * g.op_bytes);
          const uint32_t a_lo = a_hi + a_op_bytes;
          const uint32_t b_hi = a_lo + a_op_bytes;
          const uint32_t b_lo = b_hi + b_op_bytes;
#pragma unroll
          for (int k = 0; k < 2; ++k) {               // 16 tokens = two 8-row atoms = 2048 bytes per K step
            const uint64_t dah = smem_desc_sw128_mnmajor16(a_hi + k * 2048);
            const uint64_t dal = smem_desc_sw128_mnmajor16(a_lo + k * 2048);
            const uint64_t dbh = smem_desc_sw128_mnmajor16(b_hi + k * 2048);
            const uint64_t dbl = smem_desc_sw128_mnmajor16(b_lo + k * 2048);
            if (g.debug & 8) {
              mma_bf16_ss(d_tmem, dah, dbh, idesc, (kb | k) ? 1u : 0u);
              continue;
            }
            mma_bf16_ss(d_tmem, dal, dbh, idesc, (kb | k) ? 1u : 0u);
            mma_bf16_ss(d_tmem, dah, dbl, idesc, 1u);
            mma_bf16_ss(d_tmem, dah, dbh, idesc, 1u);
          }
          mma_commit(&op_empty[slot]);
          if (kb == kblocks - 1) mma_commit(&tfull[acc]);
          if (++slot == kRing) { slot = 0; phase ^= 1; }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 8) {
    // ================= converters (256 threads) =================
    // thread -> token row k = t % 32, 8-column chunk mc = (t / 32) % 4 of a 32-column box, boxes of parity t / 128.
    // The 8 lanes of a shared-memory phase touch 8 consecutive rows, i.e. 8 different XOR patterns: conflict-free
    // reads (raw, SWIZZLE_128B) and writes (operand, SWIZZLE_128B).
    const int t = threadIdx.x - 256;
    const int k = t & 31;
    const int mc = (t >> 5) & 3;
    const int par = t >> 7;
    const int sw = k & 7;
    const int n_boxes = 4 + b_boxes;
    int slot = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      // column sums of A (= dY^T 1, the bias gradient next to the weight gradient dY^T X): every converter thread
      // already holds its raw A values in registers; it adds them up over the tile's token blocks (its token row k,
      // its 8 columns of the boxes par and par + 2) and the warp folds the 32 token rows once per tile.  Only the
      // first N tile of an M tile does it.
      const bool do_cs = g.colsum != nullptr && (tile % g.tiles_n) == 0;
      float cs[2][8];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) cs[i][j] = 0.f;
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(&raw_full[slot], phase);
        if (!(g.debug & 4)) {
          const uint8_t* raw = raw_ring + slot * g.raw_bytes + k * 128;
          uint8_t* opa = op_ring + slot * g.op_bytes + k * 128;
          uint8_t* opb = opa + 2 * a_op_bytes;
          // raw boxes of this thread -> registers; all 256 converter threads; then the halves over the same bytes
          float4 u[6], v[6];
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            const int box = par + 2 * i;
            if (box < n_boxes) {
              const uint8_t* rb = raw + box * kBoxBytes;
              u[i] = *reinterpret_cast<const float4*>(rb + (((2 * mc) ^ sw) << 4));
              v[i] = *reinterpret_cast<const float4*>(rb + (((2 * mc + 1) ^ sw) << 4));
            }
          }
          if (do_cs) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {                        // boxes par, par + 2: the A boxes of this thread
              cs[i][0] += u[i].x; cs[i][1] += u[i].y; cs[i][2] += u[i].z; cs[i][3] += u[i].w;
              cs[i][4] += v[i].x; cs[i][5] += v[i].y; cs[i][6] += v[i].z; cs[i][7] += v[i].w;
            }
          }
          conv_bar();
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            const int box = par + 2 * i;
            if (box < n_boxes) {
              uint4 h, l;
              split8(u[i], v[i], h, l);
              const bool is_a = box < 4;
              const int bi = is_a ? box : box - 4;
              uint8_t* dst = (is_a ? opa : opb) + (bi >> 1) * kGroupBytes + (((((bi & 1) << 2) + mc) ^ sw) << 4);
              const int half_bytes = is_a ? a_op_bytes : b_op_bytes;
              *reinterpret_cast<uint4*>(dst) = h;
              *reinterpret_cast<uint4*>(dst + half_bytes) = l;
            }
          }
          fence_proxy_async_smem();
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&op_full[slot]);
        if (++slot == kRing) { slot = 0; phase ^= 1; }
      }
      if (do_cs) {
        const int rest = tile / g.tiles_n;
        const int m_t = rest % g.tiles_m, slab = rest / g.tiles_m;
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float x = cs[i][j];
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
            const int col = m_t * kBM + (par + 2 * i) * 32 + mc * 8 + j;
            if (lane == 0 && col < g.M) g.colsum[static_cast<long long>(slab) * g.colsum_ld + col] = x;
          }
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue (warps 4..7 -> TMEM lane groups 0..3) =================
    const int ew = warp - 4;
    const int trow = ew * 32 + lane;
    const bool issuer = (threadIdx.x == 128);
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t chunk_ctr = 0;
    const int nchunks = BN / 32;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int n_t = tile % g.tiles_n;
      int rest = tile / g.tiles_n;
      const int m_t = rest % g.tiles_m;
      const int slab = rest / g.tiles_m;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + static_cast<uint32_t>(acc * 256);
      int live = 0;
      for (int c = 0; c < nchunks; ++c)
        if (n_t * BN + c * 32 < g.N) live = c + 1;
#pragma unroll 1
      for (int c = 0; c < live; ++c) {
        const int n0 = n_t * BN + c * 32;
        uint32_t v[32];
        tmem_ld_32x32(t_addr + c * 32, v);
        tmem_ld_wait();
        if (c == live - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[acc]);
        }
        uint8_t* buf = staging + (chunk_ctr & 1) * kStagingBytes;
        if (issuer) bulk_wait_read<1>();
        epi_bar();
        uint8_t* br = buf + trow * 128;
        const int sw = trow & 7;
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<uint4*>(br + ((q ^ sw) << 4)) = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        fence_proxy_async_smem();
        epi_bar();
        if (issuer) {
          // accumulate mode: all K-splits of a batch entry add into ONE output slab (zeroed or pre-filled by the caller)
          if (g.accumulate) tma_reduce_add_3d(&tmC, buf, n0, m_t * kBM, slab / g.k_splits);
          else tma_store_3d(&tmC, buf, n0, m_t * kBM, slab);
          bulk_commit();
        }
        ++chunk_ctr;
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (issuer) bulk_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
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
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    cudaFree(nullptr);
    ctx_bound = true;
  }
  return enc;
}

// fp32 tensor [d2, d1, d0] (d0 contiguous), strides in elements, box [1, box1, 32], SWIZZLE_128B
static int make_tmap(CUtensorMap* m, const float* base, long long d0, long long d1, long long d2, long long ld1,
                     long long ld2, int box1, const char* what) {
  EncodeFn enc = encoder();
  if (enc == nullptr) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return MPF_ERR_UNSUPPORTED;
  }
  if (ld1 % 4 != 0 || ld2 % 4 != 0 || !aligned16(base)) {
    set_error("gemm_bf16x3_tn: %s needs a 16-byte aligned base and strides that are multiples of 4 elements", what);
    return MPF_ERR_BAD_ARG;
  }
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(d0), static_cast<cuuint64_t>(d1), static_cast<cuuint64_t>(d2)};
  if (ld2 < d1 * ld1) ld2 = d1 * ld1;
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld1) * 4, static_cast<cuuint64_t>(ld2) * 4};
  cuuint32_t box[3] = {32, static_cast<cuuint32_t>(box1), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(%s) failed (CUresult %d): dims=(%lld,%lld,%lld) ld=(%lld,%lld)", what,
              static_cast<int>(r), d0, d1, d2, ld1, ld2);
    return MPF_ERR_BAD_ARG;
  }
  return MPF_OK;
}

// channels-last fp32 map [d3, d2, d1, d0 = channels], box {32 channels, 32 pixels of one row, 1, 1}, SWIZZLE_128B
static int make_tmap_4d(CUtensorMap* m, const float* base, long long d0, long long d1, long long d2, long long d3,
                        const char* what) {
  EncodeFn enc = encoder();
  if (enc == nullptr) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return MPF_ERR_UNSUPPORTED;
  }
  if (d0 % 4 != 0 || !aligned16(base)) {
    set_error("gemm_bf16x3_tn: %s needs a 16-byte aligned base and channels %% 4 == 0", what);
    return MPF_ERR_BAD_ARG;
  }
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(d0), static_cast<cuuint64_t>(d1), static_cast<cuuint64_t>(d2),
                        static_cast<cuuint64_t>(d3)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(d0) * 4, static_cast<cuuint64_t>(d0) * d1 * 4,
                           static_cast<cuuint64_t>(d0) * d1 * d2 * 4};
  cuuint32_t box[4] = {32, static_cast<cuuint32_t>(kBK), 1, 1};
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

}  // namespace bf3tn
}  // namespace mpf

extern "C" {

// conv_H > 0: weight gradient of a 3x3 convolution; B = the channels-last input map [conv_B, conv_H, conv_W, conv_C],
// N = 9 * conv_C, T = conv_B * conv_H * conv_W, batch = 1 (ldb / b_batch_stride unused)
static int gemm_bf16x3_tn_impl(const float* A, long long lda, long long a_batch_stride, const float* B, long long ldb,
                               long long b_batch_stride, float* C, long long ldc, long long c_batch_stride, int batch,
                               int M, int N, int T, int k_splits, int accumulate, int conv_B, int conv_H, int conv_W,
                               int conv_C, float* colsum, long long colsum_ld, void* stream) {
  using namespace mpf;
  using namespace mpf::bf3tn;
  clear_error();
  MPF_REQUIRE(A && B && C, "gemm_bf16x3_tn: null pointer argument");
  MPF_REQUIRE(batch > 0 && M > 0 && N > 0 && T > 0 && k_splits >= 1, "gemm_bf16x3_tn: dimensions must be positive");
  MPF_REQUIRE(lda >= M && ldb >= N && ldc >= N, "gemm_bf16x3_tn: row stride too small");
  Args g;
  // N tile: every n-tile repeats the conversion of the 128-column A slab, so score = tiles_n * (bn + 128)
  g.bn = 64;
  long long best_cost = -1;
  for (int bn = 64; bn <= 256; bn += 64) {
    if (conv_H > 0 && conv_C % bn != 0) continue;          // an N tile must not straddle two taps
    const long long cost = static_cast<long long>((N + bn - 1) / bn) * (bn + 128);
    if (best_cost < 0 || cost <= best_cost) { best_cost = cost; g.bn = bn; }
  }
  if (const char* force = getenv("MPF_GEMM_BN")) {
    const int f = atoi(force);
    if (f >= 64 && f <= 256 && f % 64 == 0 && (conv_H == 0 || conv_C % f == 0)) g.bn = f;
  }
  g.batch = batch; g.M = M; g.N = N; g.T = T;
  g.tiles_m = (M + kBM - 1) / kBM;
  g.tiles_n = (N + g.bn - 1) / g.bn;
  const int kblocks_total = (T + kBK - 1) / kBK;
  if (k_splits > kblocks_total) k_splits = kblocks_total;
  g.k_splits = k_splits;
  g.k_per_split = (kblocks_total + k_splits - 1) / k_splits * kBK;
  g.raw_bytes = (4 + g.bn / 32) * kBoxBytes;
  g.op_bytes = 2 * (2 + g.bn / 64) * kGroupBytes;
  MPF_REQUIRE(g.raw_bytes == g.op_bytes, "gemm_bf16x3_tn: raw and split slot sizes differ");
  g.ring = (kSmemBudget - 1024 - 512 - 2 * kStagingBytes) / g.raw_bytes;
  if (g.ring > kMaxRing) g.ring = kMaxRing;
  if (const char* e = getenv("MPF_GEMM_TN_RING")) {          // measurement knob
    const int r = atoi(e);
    if (r >= 2 && r < g.ring) g.ring = r;
  }
  MPF_REQUIRE(g.ring >= 2, "gemm_bf16x3_tn: shared-memory budget exceeded");
  const int smem_bytes = g.ring * g.raw_bytes + 2 * kStagingBytes + 512 + 1024;
  MPF_REQUIRE(static_cast<long long>(batch) * k_splits * g.tiles_m * g.tiles_n < (1ll << 31), "gemm_bf16x3_tn: too many tiles");
  g.debug = 0;
  if (const char* dbg = getenv("MPF_GEMM_DEBUG")) g.debug = atoi(dbg);
  g.accumulate = accumulate ? 1 : 0;
  MPF_REQUIRE(colsum == nullptr || (!accumulate && conv_H == 0), "gemm_bf16x3_tn: column sums need plain slabs");
  g.colsum = colsum;
  g.colsum_ld = colsum_ld;

  CUtensorMap ta, tb, tc;
  int rc = make_tmap(&ta, A, M, T, batch, lda, a_batch_stride, kBK, "A");
  if (rc) return rc;
  g.conv_H = g.conv_ws = g.conv_nt = 0;
  if (conv_H > 0) {
    g.conv_H = conv_H;
    g.conv_ws = (conv_W + kBK - 1) / kBK;
    g.conv_nt = conv_C / g.bn;
    rc = make_tmap_4d(&ta, A, M, conv_W, conv_H, conv_B, "conv output gradient");
    if (rc) return rc;
    rc = make_tmap_4d(&tb, B, conv_C, conv_W, conv_H, conv_B, "conv input");
  } else {
    rc = make_tmap(&tb, B, N, T, batch, ldb, b_batch_stride, kBK, "B");
  }
  if (rc) return rc;
  rc = make_tmap(&tc, C, N, M, accumulate ? batch : static_cast<long long>(batch) * k_splits, ldc, c_batch_stride, kBM,
                 "C");
  if (rc) return rc;

  static unsigned long long configured_on = 0;
  if (first_use_on_this_device(configured_on)) {
    MPF_CUDA_OK(cudaFuncSetAttribute(gemm_bf16x3_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
  }
  const long long tiles = static_cast<long long>(batch) * k_splits * g.tiles_m * g.tiles_n;
  const int grid = static_cast<int>(tiles < sm_count() ? tiles : sm_count());
  gemm_bf16x3_tn_kernel<<<grid, kThreads, smem_bytes, static_cast<cudaStream_t>(stream)>>>(ta, tb, tc, g);
  count_launch();
  return finish_launch("gemm_bf16x3_tn");
}

int mpf_gemm_bf16x3_tn_ex(const float* A, long long lda, long long a_batch_stride, const float* B, long long ldb,
                          long long b_batch_stride, float* C, long long ldc, long long c_batch_stride, int batch, int M,
                          int N, int T, int k_splits, int accumulate, void* stream) {
  return gemm_bf16x3_tn_impl(A, lda, a_batch_stride, B, ldb, b_batch_stride, C, ldc, c_batch_stride, batch, M, N, T,
                             k_splits, accumulate, 0, 0, 0, 0, nullptr, 0, stream);
}

int mpf_gemm_bf16x3_tn_colsum(const float* A, long long lda, long long a_batch_stride, const float* B, long long ldb,
                              long long b_batch_stride, float* C, long long ldc, long long c_batch_stride, int batch,
                              int M, int N, int T, int k_splits, float* colsum, long long colsum_slab_stride,
                              void* stream) {
  MPF_REQUIRE(colsum != nullptr && colsum_slab_stride >= M, "gemm_bf16x3_tn_colsum: bad column-sum output");
  return gemm_bf16x3_tn_impl(A, lda, a_batch_stride, B, ldb, b_batch_stride, C, ldc, c_batch_stride, batch, M, N, T,
                             k_splits, 0, 0, 0, 0, 0, colsum, colsum_slab_stride, stream);
}

int mpf_conv3x3_cl_wgrad_bf16x3(const float* dy, const float* x, float* dw, int batch, int H, int W, int Cin, int Cout,
                                int k_splits, void* stream) {
  using namespace mpf;
  clear_error();
  MPF_REQUIRE(dy && x && dw, "conv3x3_cl_wgrad: null pointer argument");
  MPF_REQUIRE(batch > 0 && H > 0 && W > 0 && Cin > 0 && Cin % 64 == 0 && Cout > 0 && Cout % 4 == 0,
              "conv3x3_cl_wgrad: needs Cin %% 64 == 0, Cout %% 4 == 0 (H=%d W=%d Cin=%d Cout=%d)", H, W, Cin, Cout);
  // the reduction runs over row segments of 32 pixels (the last one of a row is zero-filled beyond W)
  const long long T = static_cast<long long>(batch) * H * ((W + 31) / 32) * 32;
  MPF_REQUIRE(T < (1ll << 31), "conv3x3_cl_wgrad: too many pixels");
  const int N = 9 * Cin;
  return gemm_bf16x3_tn_impl(dy, Cout, 0, x, N, 0, dw, N, static_cast<long long>(Cout) * N, 1, Cout, N,
                             static_cast<int>(T), k_splits, 0, batch, H, W, Cin, nullptr, 0, stream);
}

int mpf_gemm_bf16x3_tn(const float* A, long long lda, long long a_batch_stride, const float* B, long long ldb,
                       long long b_batch_stride, float* C, long long ldc, long long c_batch_stride, int batch, int M,
                       int N, int T, int k_splits, void* stream) {
  return mpf_gemm_bf16x3_tn_ex(A, lda, a_batch_stride, B, ldb, b_batch_stride, C, ldc, c_batch_stride, batch, M, N, T,
                               k_splits, 0, stream);
}

}  // extern "C"

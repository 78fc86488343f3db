// C[b] = A[b] * B[b]^T (+ bias) with fp32-accurate "3xTF32" arithmetic on the 5th-gen tensor cores.
//
//   A : [batch, M, K] fp32, K contiguous (activations / pixel features, channels-last)
//   B : [batch, N, K] fp32, K contiguous (nn.Linear weight [out,in], or the per-query mask embedding)
//       supplied PRE-SPLIT as B_hi, B_lo (mpf_split_tf32) because it is small and reused by every CTA
//   C : row-major [batch, M, ldc] or transposed [batch, N, ldc_t] (used for the mask logits
//       out[b,q,hw] = sum_c E[b,q,c] * F[b,hw,c], ref decoder :1865)
//
// Why 3xTF32: the path's contract is fp32 parity (1e-3) with the reference's fp32 cuBLAS results; a
// single TF32 pass (10-bit mantissa) gives ~5e-3 absolute error at K = 256.  Each operand is written
// as x = hi + lo with hi = rn_tf32(x), lo = rn_tf32(x - hi) (both exactly representable in TF32), and
// D += A_lo*B_hi + A_hi*B_lo + A_hi*B_hi accumulates in fp32 in TMEM; the dropped lo*lo term is
// ~2^-22 relative.
//
// Structure (one CTA per SM, persistent over output tiles, warp-specialised):
//   warp 0      TMA producer: A tile (fp32), B_hi, B_lo tiles, SWIZZLE_128B, mbarrier complete_tx
//   warps 8-11  splitters: rewrite the landed A tile in place as A_hi and write A_lo next to it
//               (same swizzled positions), fence.proxy.async, arrive
//   warp 1      MMA issuer: 3 x tcgen05.mma.kind::tf32 per 8-wide K step, accumulators in TMEM
//               (double-buffered across tiles), tcgen05.commit frees the smem stage
//   warps 4-7   epilogue: tcgen05.ld TMEM -> registers -> bias / ReLU -> global (row-major or
//               transposed), overlapped with the next tile's main loop
#include "mpf_common.cuh"
#include "sm100_ptx.cuh"
#include "tmap.cuh"

#include <cstdlib>
#include <mutex>

namespace mpf {

using namespace ptx;

constexpr int kBM = 128;
constexpr int kBK = 32;                       // 32 fp32 = 128 B = one swizzle row
constexpr int kUmmaK = 8;                     // tf32: 32 bytes per MMA K step
constexpr int kABytes = kBM * kBK * 4;        // 16 KiB
constexpr int kGemmThreads = 384;

struct GemmArgs {
  float* C;
  const float* bias;                          // [N] or null
  long long c_batch_stride;                   // elements
  long long ldc;                              // row stride of C (row-major: of M rows; transposed: of N rows)
  int batch, M, N, K;
  int tiles_m, tiles_n;
  int relu;
  int transpose_c;
  int vec_store;                              // row-major float4 stores are legal
  int vec_aux;                                // bias / resid / gate may be read as float4
  float* C_lo;                                // if set: C receives rn_tf32(x), C_lo rn_tf32(x - hi)
  const float* resid;                         // optional addend [resid_rows, resid_ld]; row % resid_rows
  long long resid_ld;
  int resid_rows, resid_cols;                 // resid_rows == 0: one row per output row; cols < resid_cols get it
  float alpha;                                // x = (acc + bias + resid) * alpha, then ReLU
  const float* gate;                          // optional [M, gate_ld]: x = gate[row][n] > 0 ? x : 0 (ReLU backward)
  long long gate_ld;
  int k_splits;                               // K is cut into k_splits ranges of k_per_split (multiple of 32);
  int k_per_split;                            // split s of batch b writes partial sums to C[(b*k_splits+s)]
  int debug;                                  // MPF_GEMM_DEBUG bit mask (timing experiments only; results invalid):
                                              //  1 no global stores  2 no epilogue work  4 no A/B splitting
                                              //  8 one MMA of three  16 no B loads  32 no A loads  64 epilogue = tcgen05.ld only
};

template <int BN>
struct GemmCfg {
  static constexpr int kBBytes = BN * kBK * 4;
  static constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;
  static constexpr int kStages = (BN == 256) ? 2 : (BN == 128 ? 3 : 4);
  static constexpr int kBarrierBytes = 256;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarrierBytes + 1024;  // + alignment slack
  static constexpr int kTmemCols = 2 * BN;     // power of two >= 32 for BN in {64,128,256}
};

__device__ __forceinline__ float rn_tf32(float x) {
  // round-to-nearest (ties away) to a 10-bit mantissa; result has its 13 low bits clear
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

// A_MN / B_MN: the operand is "MN-major" -- stored [K rows][M (or N) contiguous] in global memory, i.e.
// the transpose of the K-major case -- and is fed to the tensor core through MN-major shared-memory
// descriptors, so weight gradients (dW = dY^T X) and the like need no transposed copies.
// SPLIT_B: B arrives as plain fp32 and is split into hi/lo in shared memory like A (tmBlo unused).
template <int BN, bool A_MN, bool B_MN, bool SPLIT_B>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi,
                   const __grid_constant__ CUtensorMap tmBlo, const GemmArgs g) {
  using Cfg = GemmCfg<BN>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * Cfg::kStageBytes);
  uint64_t* full = bars;               // [S] TMA bytes landed
  uint64_t* split = bars + S;          // [S] A_hi / A_lo written
  uint64_t* empty = bars + 2 * S;      // [S] MMAs reading the stage retired
  uint64_t* tfull = bars + 3 * S;      // [2] accumulator complete
  uint64_t* tempty = bars + 3 * S + 2; // [2] accumulator drained by the epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * S + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmBhi);
    prefetch_tmap(&tmBlo);
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&split[s], 4);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_tiles = g.batch * g.k_splits * g.tiles_m * g.tiles_n;
  const int kblocks = g.k_per_split / kBK;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n_t = tile % g.tiles_n;
        int rest = tile / g.tiles_n;
        const int m_t = rest % g.tiles_m;
        rest /= g.tiles_m;
        const int ks = rest % g.k_splits;
        const int b = rest / g.k_splits;
        const int kb0 = ks * kblocks;
        for (int kbi = 0; kbi < kblocks; ++kbi) {
          const int kb = kb0 + kbi;
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* st = smem + stage * Cfg::kStageBytes;
          const bool ldA = !(g.debug & 32), ldB = !(g.debug & 16);
          mbar_arrive_expect_tx(&full[stage], (ldA ? kABytes : 0) + (ldB ? (SPLIT_B ? 1 : 2) * Cfg::kBBytes : 0));
          if (!ldA) {
          } else if (A_MN) {          // [32 k-rows x 32 m] boxes of 4 KiB, one per 32-wide slice of M
#pragma unroll
            for (int i = 0; i < kBM / 32; ++i)
              tma_load_3d(st + i * 4096, &tmA, &full[stage], m_t * kBM + i * 32, kb * kBK, b);
          } else {
            tma_load_3d(st, &tmA, &full[stage], kb * kBK, m_t * kBM, b);
          }
          if (!ldB) {
          } else if (B_MN) {
#pragma unroll
            for (int i = 0; i < BN / 32; ++i) {
              tma_load_3d(st + 2 * kABytes + i * 4096, &tmBhi, &full[stage], n_t * BN + i * 32, kb * kBK, b);
              if (!SPLIT_B)
                tma_load_3d(st + 2 * kABytes + Cfg::kBBytes + i * 4096, &tmBlo, &full[stage], n_t * BN + i * 32,
                            kb * kBK, b);
            }
          } else {
            tma_load_3d(st + 2 * kABytes, &tmBhi, &full[stage], kb * kBK, n_t * BN, b);
            if (!SPLIT_B) tma_load_3d(st + 2 * kABytes + Cfg::kBBytes, &tmBlo, &full[stage], kb * kBK, n_t * BN, b);
          }
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_tf32(kBM, BN) | (A_MN ? (1u << 15) : 0u) | (B_MN ? (1u << 16) : 0u);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full[stage], phase);
          mbar_wait(&split[stage], phase);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t a_lo = a_hi + kABytes;
          const uint32_t b_hi = a_hi + 2 * kABytes;
          const uint32_t b_lo = b_hi + Cfg::kBBytes;
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k) {
            // K-major: advance 32 bytes inside the 128-B swizzle row; MN-major: one 8-row (1 KiB) atom
            const uint32_t koa = A_MN ? k * 1024 : k * kUmmaK * 4;
            const uint32_t kob = B_MN ? k * 1024 : k * kUmmaK * 4;
            const uint64_t dah = A_MN ? smem_desc_sw128_mnmajor(a_hi + koa) : smem_desc_sw128_kmajor(a_hi + koa);
            const uint64_t dal = A_MN ? smem_desc_sw128_mnmajor(a_lo + koa) : smem_desc_sw128_kmajor(a_lo + koa);
            const uint64_t dbh = B_MN ? smem_desc_sw128_mnmajor(b_hi + kob) : smem_desc_sw128_kmajor(b_hi + kob);
            const uint64_t dbl = B_MN ? smem_desc_sw128_mnmajor(b_lo + kob) : smem_desc_sw128_kmajor(b_lo + kob);
            if (g.debug & 8) {
              mma_tf32_ss(d_tmem, dah, dbh, idesc, (kb | k) ? 1u : 0u);
              continue;
            }
            mma_tf32_ss(d_tmem, dal, dbh, idesc, (kb | k) ? 1u : 0u);
            mma_tf32_ss(d_tmem, dah, dbl, idesc, 1u);
            mma_tf32_ss(d_tmem, dah, dbh, idesc, 1u);
          }
          mma_commit(&empty[stage]);
          if (kb == kblocks - 1) mma_commit(&tfull[acc]);
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 8) {
    // ================= splitters (128 threads) =================
    const int st_id = threadIdx.x - 256;
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(&full[stage], phase);
        uint8_t* a = smem + stage * Cfg::kStageBytes;
        if (g.debug & 4) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&split[stage]);
          if (++stage == S) { stage = 0; phase ^= 1; }
          continue;
        }
#pragma unroll
        for (int i = 0; i < kABytes / 16 / 128; ++i) {
          const int off = (st_id + i * 128) * 16;
          const float4 x = *reinterpret_cast<const float4*>(a + off);
          float4 hi, lo;
          hi.x = rn_tf32(x.x); hi.y = rn_tf32(x.y); hi.z = rn_tf32(x.z); hi.w = rn_tf32(x.w);
          lo.x = rn_tf32(x.x - hi.x); lo.y = rn_tf32(x.y - hi.y);
          lo.z = rn_tf32(x.z - hi.z); lo.w = rn_tf32(x.w - hi.w);
          *reinterpret_cast<float4*>(a + off) = hi;
          *reinterpret_cast<float4*>(a + kABytes + off) = lo;
        }
        if (SPLIT_B) {
          uint8_t* bt = a + 2 * kABytes;
#pragma unroll
          for (int i = 0; i < Cfg::kBBytes / 16 / 128; ++i) {
            const int off = (st_id + i * 128) * 16;
            const float4 x = *reinterpret_cast<const float4*>(bt + off);
            float4 hi, lo;
            hi.x = rn_tf32(x.x); hi.y = rn_tf32(x.y); hi.z = rn_tf32(x.z); hi.w = rn_tf32(x.w);
            lo.x = rn_tf32(x.x - hi.x); lo.y = rn_tf32(x.y - hi.y);
            lo.z = rn_tf32(x.z - hi.z); lo.w = rn_tf32(x.w - hi.w);
            *reinterpret_cast<float4*>(bt + off) = hi;
            *reinterpret_cast<float4*>(bt + Cfg::kBBytes + off) = lo;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&split[stage]);
        if (++stage == S) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue (warps 4..7 -> TMEM lane groups 0..3) =================
    const int ew = warp - 4;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int n_t = tile % g.tiles_n;
      int rest = tile / g.tiles_n;
      const int m_t = rest % g.tiles_m;
      const int b = rest / g.tiles_m;             // = batch * k_splits + split: index of the output slab
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const int row = m_t * kBM + ew * 32 + lane;
      const bool row_ok = row < g.M;
      const long long boff = static_cast<long long>(b) * g.c_batch_stride;
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + static_cast<uint32_t>(acc * BN);
      const float* grow = (g.gate != nullptr && row_ok) ? g.gate + static_cast<long long>(row) * g.gate_ld : nullptr;
      const float* rrow = nullptr;
      if (g.resid != nullptr && row_ok)
        rrow = g.resid + static_cast<long long>(g.resid_rows > 0 ? row % g.resid_rows : row) * g.resid_ld;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int n0 = n_t * BN + c * 32;
        if (n0 >= g.N || (g.debug & 2)) break;
        uint32_t v[32];
        tmem_ld_32x32(t_addr + c * 32, v);
        tmem_ld_wait();
        if (g.debug & 64) continue;
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        const bool full = n0 + 32 <= g.N;
        // addends / gate are read 16 bytes at a time when the chunk is full and the pointers allow it
        if (g.bias != nullptr) {
          if (full && g.vec_aux) {
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
        }
        if (grow != nullptr) {
          if (full && g.vec_aux) {
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
        auto store = [&](float* base, const float (&val)[32]) {
          float* cb = base + boff;
          if (g.transpose_c) {
            if (row_ok) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (n0 + j < g.N) cb[static_cast<long long>(n0 + j) * g.ldc + row] = val[j];
            }
          } else if (row_ok) {
            float* dst = cb + static_cast<long long>(row) * g.ldc + n0;
            if (g.vec_store && n0 + 32 <= g.N) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(dst + j) = make_float4(val[j], val[j + 1], val[j + 2], val[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (n0 + j < g.N) dst[j] = val[j];
            }
          }
        };
        if (g.debug & 1) {
        } else if (g.C_lo == nullptr) {
          store(g.C, f);
        } else {                       // emit the result pre-split for a following 3xTF32 consumer
          float lo[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float h = rn_tf32(f[j]);
            lo[j] = rn_tf32(f[j] - h);
            f[j] = h;
          }
          store(g.C, f);
          store(g.C_lo, lo);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

__global__ void __launch_bounds__(256)
split_tf32_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, long long n) {
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * 256) {
    const float v = x[i];
    const float h = rn_tf32(v);
    hi[i] = h;
    lo[i] = rn_tf32(v - h);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
int make_tmap_f32_3d(CUtensorMap* m, const float* base, long long d0, long long d1, long long d2,
                     long long ld1, long long ld2, int box0, int box1, bool atom32) {
  using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn enc = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      enc = reinterpret_cast<EncodeFn>(p);
  });
  if (enc == nullptr) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return MPF_ERR_UNSUPPORTED;
  }
  // The encoder needs a current CUDA context; a fresh thread (e.g. PyTorch's autograd worker) may not have
  // one bound yet.  cudaFree(0) binds the device's primary context to this thread (no-op afterwards).
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    cudaFree(nullptr);
    ctx_bound = true;
  }
  if (box0 * 4 != 128 || ld1 % 4 != 0 || ld2 % 4 != 0 || !aligned16(base)) {
    set_error("make_tmap: need a 128-byte box row, 16-byte aligned base and strides (ld1=%lld ld2=%lld)", ld1, ld2);
    return MPF_ERR_BAD_ARG;
  }
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(d0), static_cast<cuuint64_t>(d1), static_cast<cuuint64_t>(d2)};
  if (d2 == 1 && ld2 < d1 * ld1) ld2 = d1 * ld1;
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld1) * 4, static_cast<cuuint64_t>(ld2) * 4};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(box0), static_cast<cuuint32_t>(box1), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d): base=%p dims=(%lld,%lld,%lld) ld=(%lld,%lld) box=(%d,%d)",
              static_cast<int>(r), static_cast<const void*>(base), d0, d1, d2, ld1, ld2, box0, box1);
    return MPF_ERR_BAD_ARG;
  }
  return MPF_OK;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int BN, bool A_MN, bool B_MN, bool SPLIT_B>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tbh, const CUtensorMap& tbl, GemmArgs g,
                       cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  static unsigned long long configured_on = 0;
  if (first_use_on_this_device(configured_on)) {
    MPF_CUDA_OK(cudaFuncSetAttribute(gemm_tf32x3_kernel<BN, A_MN, B_MN, SPLIT_B>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
  }
  g.tiles_m = (g.M + kBM - 1) / kBM;
  g.tiles_n = (g.N + BN - 1) / BN;
  const long long tiles = static_cast<long long>(g.batch) * g.k_splits * g.tiles_m * g.tiles_n;
  const int grid = static_cast<int>(tiles < sm_count() ? tiles : sm_count());
  gemm_tf32x3_kernel<BN, A_MN, B_MN, SPLIT_B><<<grid, kGemmThreads, Cfg::kSmemBytes, st>>>(ta, tbh, tbl, g);
  count_launch();
  return finish_launch("gemm_tf32x3");
}

}  // namespace mpf

extern "C" {

int mpf_split_tf32(const float* x, float* hi, float* lo, long long n, void* stream) {
  mpf::clear_error();
  MPF_REQUIRE(x && hi && lo && n > 0, "split_tf32: bad arguments");
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  mpf::split_tf32_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, hi, lo, n);
  mpf::count_launch();
  return mpf::finish_launch("split_tf32");
}

int mpf_gemm_tf32x3_general(const float* A, int a_mn_major, long long lda, long long a_batch_stride,
                            const float* B, const float* B_lo, int b_mn_major, long long ldb,
                            long long b_batch_stride, const float* bias, float* C, float* C_lo, long long ldc,
                            long long c_batch_stride, const float* resid, long long resid_ld, int resid_rows,
                            int resid_cols, const float* gate, long long gate_ld, float alpha, int batch, int M,
                            int N, int K, int k_splits, int relu, int transpose_c, void* stream) {
  using namespace mpf;
  clear_error();
  MPF_REQUIRE(A && B && C, "gemm_tf32x3: null pointer argument");
  MPF_REQUIRE(k_splits >= 1, "gemm_tf32x3: k_splits must be >= 1");
  MPF_REQUIRE(gate == nullptr || (batch == 1 && !transpose_c), "gemm_tf32x3: gate needs batch 1, row-major C");
  MPF_REQUIRE(k_splits == 1 || (!bias && !resid && !relu && !C_lo && !gate && alpha == 1.0f),
              "gemm_tf32x3: split-K produces partial sums; bias / residual / ReLU / scale / split output are not allowed");
  MPF_REQUIRE(batch > 0 && M > 0 && N > 0 && K > 0, "gemm_tf32x3: dimensions must be positive");
  MPF_REQUIRE(lda % 4 == 0 && ldb % 4 == 0 && a_batch_stride % 4 == 0 && b_batch_stride % 4 == 0 &&
                  aligned16(A) && aligned16(B) && (B_lo == nullptr || aligned16(B_lo)),
              "gemm_tf32x3: A / B must be 16-byte aligned with strides that are multiples of 4 elements");
  MPF_REQUIRE(lda >= (a_mn_major ? M : K) && ldb >= (b_mn_major ? N : K), "gemm_tf32x3: row stride too small");
  MPF_REQUIRE(static_cast<long long>(batch) * k_splits * ((M + kBM - 1) / kBM) * ((N + 63) / 64) < (1ll << 31),
              "gemm_tf32x3: too many tiles");
  const int kblocks_total = (K + kBK - 1) / kBK;  // the tail k-block is zero-filled by TMA
  const int k_per_split = (kblocks_total + k_splits - 1) / k_splits * kBK;
  int bn = 64;
  if (N > 64) {
    const int waste128 = (N + 127) / 128 * 128 - N, waste256 = (N + 255) / 256 * 256 - N;
    bn = (N <= 128 || waste128 < waste256) ? 128 : 256;
  }
  if (const char* force = getenv("MPF_GEMM_BN")) {        // tuning aid (benchmarks/kernel_probe.py)
    const int f = atoi(force);
    if (f == 64 || f == 128 || f == 256) bn = f;
  }
  const bool split_b = B_lo == nullptr;
  CUtensorMap ta, tbh, tbl;
  int rc;
  if (a_mn_major) rc = make_tmap_f32_3d(&ta, A, M, K, batch, lda, a_batch_stride, 32, kBK, true);
  else rc = make_tmap_f32_3d(&ta, A, K, M, batch, lda, a_batch_stride, kBK, kBM);
  if (rc) return rc;
  if (b_mn_major) rc = make_tmap_f32_3d(&tbh, B, N, K, batch, ldb, b_batch_stride, 32, kBK, true);
  else rc = make_tmap_f32_3d(&tbh, B, K, N, batch, ldb, b_batch_stride, kBK, bn);
  if (rc) return rc;
  tbl = tbh;
  if (!split_b) {
    if (b_mn_major) rc = make_tmap_f32_3d(&tbl, B_lo, N, K, batch, ldb, b_batch_stride, 32, kBK, true);
    else rc = make_tmap_f32_3d(&tbl, B_lo, K, N, batch, ldb, b_batch_stride, kBK, bn);
    if (rc) return rc;
  }
  GemmArgs g;
  g.C = C; g.C_lo = C_lo; g.bias = bias; g.c_batch_stride = c_batch_stride; g.ldc = ldc;
  g.batch = batch; g.M = M; g.N = N; g.K = K; g.tiles_m = g.tiles_n = 0;
  g.k_splits = k_splits; g.k_per_split = k_per_split;
  g.gate = gate; g.gate_ld = gate_ld;
  g.relu = relu; g.transpose_c = transpose_c;
  g.resid = resid; g.resid_ld = resid_ld; g.resid_rows = resid_rows;
  g.resid_cols = resid ? (resid_cols > 0 ? resid_cols : N) : 0;
  g.alpha = alpha;
  g.debug = 0;
  if (const char* dbg = getenv("MPF_GEMM_DEBUG")) g.debug = atoi(dbg);
  g.vec_store = (!transpose_c && ldc % 4 == 0 && c_batch_stride % 4 == 0 && aligned16(C) &&
                 (C_lo == nullptr || aligned16(C_lo))) ? 1 : 0;
  g.vec_aux = ((bias == nullptr || aligned16(bias)) && (resid == nullptr || (aligned16(resid) && resid_ld % 4 == 0)) &&
               (gate == nullptr || (aligned16(gate) && gate_ld % 4 == 0))) ? 1 : 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define MPF_DISPATCH_BN(AM, BM_, SB)                                              \
  do {                                                                            \
    if (bn == 64) return launch_gemm<64, AM, BM_, SB>(ta, tbh, tbl, g, st);       \
    if (bn == 128) return launch_gemm<128, AM, BM_, SB>(ta, tbh, tbl, g, st);     \
    return launch_gemm<256, AM, BM_, SB>(ta, tbh, tbl, g, st);                    \
  } while (0)
  if (!a_mn_major && !b_mn_major && !split_b) MPF_DISPATCH_BN(false, false, false);
  if (!a_mn_major && !b_mn_major && split_b) MPF_DISPATCH_BN(false, false, true);
  if (!a_mn_major && b_mn_major && split_b) MPF_DISPATCH_BN(false, true, true);
  if (a_mn_major && b_mn_major && split_b) MPF_DISPATCH_BN(true, true, true);
  if (a_mn_major && !b_mn_major && split_b) MPF_DISPATCH_BN(true, false, true);
#undef MPF_DISPATCH_BN
  set_error("gemm_tf32x3: unsupported operand combination (a_mn=%d b_mn=%d presplit_b=%d)", a_mn_major,
            b_mn_major, !split_b);
  return MPF_ERR_UNSUPPORTED;
}

int mpf_gemm_tf32x3_ex(const float* A, long long lda, long long a_batch_stride, const float* B_hi,
                       const float* B_lo, long long ldb, long long b_batch_stride, const float* bias, float* C,
                       float* C_lo, long long ldc, long long c_batch_stride, const float* resid,
                       long long resid_ld, int resid_rows, int resid_cols, float alpha, int batch, int M, int N,
                       int K, int relu, int transpose_c, void* stream) {
  if (B_lo == nullptr) {
    mpf::set_error("gemm_tf32x3_ex: B_lo is required (use mpf_gemm_tf32x3_general for in-kernel splitting)");
    return MPF_ERR_BAD_ARG;
  }
  return mpf_gemm_tf32x3_general(A, 0, lda, a_batch_stride, B_hi, B_lo, 0, ldb, b_batch_stride, bias, C, C_lo, ldc,
                                 c_batch_stride, resid, resid_ld, resid_rows, resid_cols, nullptr, 0, alpha, batch, M,
                                 N, K, 1, relu, transpose_c, stream);
}

int mpf_gemm_tf32x3(const float* A, long long lda, long long a_batch_stride, const float* B_hi,
                    const float* B_lo, long long ldb, long long b_batch_stride, const float* bias, float* C,
                    long long ldc, long long c_batch_stride, int batch, int M, int N, int K, int relu,
                    int transpose_c, void* stream) {
  return mpf_gemm_tf32x3_ex(A, lda, a_batch_stride, B_hi, B_lo, ldb, b_batch_stride, bias, C, nullptr, ldc,
                            c_batch_stride, nullptr, 0, 0, 0, 1.0f, batch, M, N, K, relu, transpose_c, stream);
}

}  // extern "C"

// Self-attention core of the decoder's SelfAttentionLayer over the (few hundred) queries, forward and backward:
//   ref: transformer_decoder/mask2former_transformer_decoder.py:42-52 (nn.MultiheadAttention(tgt, tgt, tgt) with the
//        optional boolean tgt_mask of the mask-piloted groups, decoder :1051-1059; True = not allowed)
//   O = softmax(Q K^T / sqrt(d) + mask) V   per (image, head);  Q, K, V are the slices of the packed in-projection.
// Qt <= 320 queries x 32 head channels: 3.3 MFLOP per (image, head) -- far too small for the tensor pipeline to pay
// (one 128 x 64 UMMA tile is already 20 % of a row block); the kernels are plain fp32 CUDA-core code, exact to fp32
// rounding, one CTA per (image, head):
//   forward : Q, K, V of the head in shared memory (row pitch 36 floats: LDS.128 by key is conflict free); a warp
//             processes FOUR query rows together -- lanes = keys for the scores (each K row is read once for the four
//             rows), shuffles for max / sum, lanes = channels for P V with the probabilities read four keys at a time
//             as broadcast LDS.128; the log-sum-exp is kept for the backward.
//   backward: probabilities recomputed from the log-sum-exp; rows are processed in chunks of 16 (two per warp): the
//             warps write P and dS = P (dP - delta) of the chunk TRANSPOSED to shared memory and produce dQ, then every
//             thread adds the chunk into ITS fixed (key, channel) entries of dK and dV held in registers (40 each at
//             Qt = 320) -- no atomics, deterministic.  (The first version -- one row per warp, scalar LDS -- was
//             LSU-bound at 180 us per call, profiles/r2h_kernels_step_b16.txt.)
// This replaces the library scaled_dot_product_attention call (PyTorch's memory-efficient kernel) of round 1.
#include "mpf_common.cuh"

namespace mpf {
namespace sa {

constexpr int kD = 32;
constexpr int kPitch = 36;                      // floats per row: 16-byte aligned rows, conflict-free LDS.128 by key
constexpr int kP4 = kPitch / 4;
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxQ = 320;
constexpr int kKeysPerLane = kMaxQ / 32;        // 10
constexpr int kFR = 4;                          // forward: query rows a warp processes together (K / V rows reused)
constexpr int kBR = 2;                          // backward: rows per warp
constexpr int kChunk = kWarps * kBR;            // backward: 16 query rows per chunk
constexpr int kCP = kChunk + 2;                 // pitch of the transposed chunk buffers: lanes = keys store with 2-way
                                                // bank conflicts (16-way at pitch 16), rows stay 8-byte aligned
constexpr int kOwn = kMaxQ * kD / kThreads;     // 40 (key, channel) entries of dK / dV per thread

__host__ __device__ constexpr int pad32(int q) { return (q + 31) / 32 * 32; }
__host__ __device__ constexpr size_t fwd_smem(int Qt) {
  return (3 * static_cast<size_t>(Qt) * kPitch + static_cast<size_t>(kWarps) * kFR * pad32(Qt)) * 4;
}
__host__ __device__ constexpr size_t bwd_smem(int Qt) {
  return (4 * static_cast<size_t>(Qt) * kPitch + 2 * static_cast<size_t>(pad32(Qt)) * kCP) * 4;
}

__device__ __forceinline__ float dot4(const float4 a, const float4 b) {
  return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
}

// qkv: [B, Qt, 3E] (q | k | v), mask: uint8 [Qt, Qt] or null (1 = not allowed), out: [B, Qt, E], lse: [B, heads, Qt]
__global__ void __launch_bounds__(kThreads)
self_attn_fwd_kernel(const float* __restrict__ qkv, const uint8_t* __restrict__ mask, int Qt, int heads, float scale,
                     float* __restrict__ out, float* __restrict__ lse) {
  extern __shared__ float4 sm4[];
  float* sm = reinterpret_cast<float*>(sm4);
  float* sQ = sm;
  float* sK = sQ + Qt * kPitch;
  float* sV = sK + Qt * kPitch;
  const int QP = pad32(Qt);
  float* sP = sV + Qt * kPitch;                   // [warps][kFR][QP]
  const int head = blockIdx.x, b = blockIdx.y;
  const int E = heads * kD, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* base = qkv + static_cast<long long>(b) * Qt * 3 * E + head * kD;
  for (int i = tid; i < Qt * kD; i += kThreads) {
    const int r = i / kD, c = i - r * kD;
    const float* row = base + static_cast<long long>(r) * 3 * E + c;
    sQ[r * kPitch + c] = __ldg(row) * scale;
    sK[r * kPitch + c] = __ldg(row + E);
    sV[r * kPitch + c] = __ldg(row + 2 * E);
  }
  __syncthreads();
  const float4* sQ4 = reinterpret_cast<const float4*>(sQ);
  const float4* sK4 = reinterpret_cast<const float4*>(sK);
  float* p = sP + warp * kFR * QP;
  for (int r0 = warp * kFR; r0 < Qt; r0 += kWarps * kFR) {
    int rr[kFR];
#pragma unroll
    for (int i = 0; i < kFR; ++i) rr[i] = min(r0 + i, Qt - 1);       // rows past the end repeat the last one (not stored)
    float mx[kFR], sum[kFR];
#pragma unroll
    for (int i = 0; i < kFR; ++i) mx[i] = -INFINITY, sum[i] = 0.f;
    // scores of the kFR rows against this lane's keys; kept in shared memory (p) between the two softmax passes
    for (int u = 0; u * 32 < Qt; ++u) {
      const int k = lane + 32 * u;
      float acc[kFR];
#pragma unroll
      for (int i = 0; i < kFR; ++i) acc[i] = 0.f;
      if (k < Qt) {
#pragma unroll
        for (int c4 = 0; c4 < kD / 4; ++c4) {
          const float4 kk = sK4[k * kP4 + c4];
#pragma unroll
          for (int i = 0; i < kFR; ++i) acc[i] += dot4(sQ4[rr[i] * kP4 + c4], kk);
        }
      }
#pragma unroll
      for (int i = 0; i < kFR; ++i) {
        const bool dead = k >= Qt || (mask != nullptr && mask[static_cast<long long>(rr[i]) * Qt + k]);
        const float v = dead ? -INFINITY : acc[i];
        if (k < QP) p[i * QP + k] = v;
        mx[i] = fmaxf(mx[i], v);
      }
    }
#pragma unroll
    for (int i = 0; i < kFR; ++i)
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], o));
    for (int u = 0; u * 32 < Qt; ++u) {
      const int k = lane + 32 * u;
#pragma unroll
      for (int i = 0; i < kFR; ++i) {
        const float v = p[i * QP + k];
        const float e = (v == -INFINITY) ? 0.f : expf(v - mx[i]);
        p[i * QP + k] = e;
        sum[i] += e;
      }
    }
#pragma unroll
    for (int i = 0; i < kFR; ++i)
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) sum[i] += __shfl_xor_sync(0xffffffffu, sum[i], o);
    __syncwarp();
    // O = P V with lane = channel; the probabilities are read four keys at a time (broadcast LDS.128)
    float o[kFR];
#pragma unroll
    for (int i = 0; i < kFR; ++i) o[i] = 0.f;
    const float4* p4 = reinterpret_cast<const float4*>(p);
    for (int k4 = 0; k4 * 4 < Qt; ++k4) {
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = (4 * k4 + j < Qt) ? sV[(4 * k4 + j) * kPitch + lane] : 0.f;
#pragma unroll
      for (int i = 0; i < kFR; ++i) {
        const float4 pp = p4[i * (QP / 4) + k4];
        o[i] += pp.x * v[0] + pp.y * v[1] + pp.z * v[2] + pp.w * v[3];
      }
    }
#pragma unroll
    for (int i = 0; i < kFR; ++i) {
      if (r0 + i < Qt) {
        // (a fully masked row gives NaN exactly like the reference's softmax)
        out[(static_cast<long long>(b) * Qt + r0 + i) * E + head * kD + lane] = o[i] / sum[i];
        if (lane == 0) lse[(static_cast<long long>(b) * heads + head) * Qt + r0 + i] = mx[i] + logf(sum[i]);
      }
    }
    __syncwarp();
  }
}

// d_out: [B, Qt, E]; d_qkv: [B, Qt, 3E] (every element written)
__global__ void __launch_bounds__(kThreads)
self_attn_bwd_kernel(const float* __restrict__ qkv, const uint8_t* __restrict__ mask, const float* __restrict__ d_out,
                     const float* __restrict__ lse, int Qt, int heads, float scale, float* __restrict__ d_qkv) {
  extern __shared__ float4 sm4[];
  float* sm = reinterpret_cast<float*>(sm4);
  float* sQ = sm;                                  // scaled by 1/sqrt(d)
  float* sK = sQ + Qt * kPitch;
  float* sV = sK + Qt * kPitch;
  float* sG = sV + Qt * kPitch;                   // dO
  const int QP = pad32(Qt);
  float* sPT = sG + Qt * kPitch;                  // [QP keys][kCP]  probabilities of the chunk's rows, transposed
  float* sST = sPT + QP * kCP;                    // [QP keys][kCP]  dS, transposed
  const int head = blockIdx.x, b = blockIdx.y;
  const int E = heads * kD, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long img = static_cast<long long>(b) * Qt;
  const float* base = qkv + img * 3 * E + head * kD;
  for (int i = tid; i < Qt * kD; i += kThreads) {
    const int r = i / kD, c = i - r * kD;
    const float* row = base + static_cast<long long>(r) * 3 * E + c;
    sQ[r * kPitch + c] = __ldg(row) * scale;
    sK[r * kPitch + c] = __ldg(row + E);
    sV[r * kPitch + c] = __ldg(row + 2 * E);
    sG[r * kPitch + c] = __ldg(d_out + (img + r) * E + head * kD + c);
  }
  const float4* sQ4 = reinterpret_cast<const float4*>(sQ);
  const float4* sK4 = reinterpret_cast<const float4*>(sK);
  const float4* sV4 = reinterpret_cast<const float4*>(sV);
  const float4* sG4 = reinterpret_cast<const float4*>(sG);
  // this thread's fixed entries of dK / dV: flat index e = tid + kThreads * i -> key e / 32, channel e % 32 (= lane)
  float dK[kOwn], dV[kOwn];
#pragma unroll
  for (int i = 0; i < kOwn; ++i) dK[i] = dV[i] = 0.f;
  __syncthreads();
  for (int c0 = 0; c0 < Qt; c0 += kChunk) {
    // ---- rows c0 + warp * kBR + {0 .. kBR-1}: P and dS into the transposed chunk buffers, dQ to global
    const int w0 = warp * kBR;                     // row slot of this warp inside the chunk
    int rr[kBR];
    float l[kBR], delta[kBR];
#pragma unroll
    for (int i = 0; i < kBR; ++i) {
      rr[i] = min(c0 + w0 + i, Qt - 1);
      l[i] = __ldg(lse + (static_cast<long long>(b) * heads + head) * Qt + rr[i]);
      delta[i] = 0.f;
    }
    for (int u = 0; u * 32 < Qt; ++u) {
      const int k = lane + 32 * u;
      float s[kBR], g[kBR];
#pragma unroll
      for (int i = 0; i < kBR; ++i) s[i] = g[i] = 0.f;
      if (k < Qt) {
#pragma unroll
        for (int c4 = 0; c4 < kD / 4; ++c4) {
          const float4 kk = sK4[k * kP4 + c4], vv = sV4[k * kP4 + c4];
#pragma unroll
          for (int i = 0; i < kBR; ++i) {
            s[i] += dot4(sQ4[rr[i] * kP4 + c4], kk);
            g[i] += dot4(sG4[rr[i] * kP4 + c4], vv);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < kBR; ++i) {
        const bool live = k < Qt && c0 + w0 + i < Qt && !(mask != nullptr && mask[static_cast<long long>(rr[i]) * Qt + k]);
        const float pv = live ? expf(s[i] - l[i]) : 0.f;
        sPT[k * kCP + w0 + i] = pv;
        sST[k * kCP + w0 + i] = pv * g[i];        // dP weighted; delta is subtracted below
        delta[i] += pv * g[i];
      }
    }
#pragma unroll
    for (int i = 0; i < kBR; ++i)
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) delta[i] += __shfl_xor_sync(0xffffffffu, delta[i], o);
    for (int u = 0; u * 32 < Qt; ++u) {
      const int k = lane + 32 * u;
#pragma unroll
      for (int i = 0; i < kBR; ++i) sST[k * kCP + w0 + i] -= sPT[k * kCP + w0 + i] * delta[i];   // dS = P (dP - delta)
    }
    __syncwarp();
    {
      float dq[kBR];
#pragma unroll
      for (int i = 0; i < kBR; ++i) dq[i] = 0.f;
      for (int k = 0; k < Qt; ++k) {
        const float kv = sK[k * kPitch + lane];
#pragma unroll
        for (int i = 0; i < kBR; ++i) dq[i] += sST[k * kCP + w0 + i] * kv;
      }
#pragma unroll
      for (int i = 0; i < kBR; ++i)
        if (c0 + w0 + i < Qt) d_qkv[(img + c0 + w0 + i) * 3 * E + head * kD + lane] = dq[i] * scale;
    }
    __syncthreads();
    // ---- every thread adds the chunk's rows into its entries of dK (dS^T Qs) and dV (P^T dO)
    float qv[kChunk], gv[kChunk];
#pragma unroll
    for (int w = 0; w < kChunk; ++w) {
      const int r = min(c0 + w, Qt - 1);            // rows past the end carry P = dS = 0
      qv[w] = sQ[r * kPitch + lane];
      gv[w] = sG[r * kPitch + lane];
    }
    const float2* sST2 = reinterpret_cast<const float2*>(sST);
    const float2* sPT2 = reinterpret_cast<const float2*>(sPT);
#pragma unroll
    for (int i = 0; i < kOwn; ++i) {
      const int k = (tid + kThreads * i) >> 5;
      if (k < Qt) {
#pragma unroll
        for (int w2 = 0; w2 < kChunk / 2; ++w2) {
          const float2 ds = sST2[k * (kCP / 2) + w2], pp = sPT2[k * (kCP / 2) + w2];
          dK[i] += ds.x * qv[2 * w2] + ds.y * qv[2 * w2 + 1];
          dV[i] += pp.x * gv[2 * w2] + pp.y * gv[2 * w2 + 1];
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < kOwn; ++i) {
    const int k = (tid + kThreads * i) >> 5;
    if (k < Qt) {
      float* row = d_qkv + (img + k) * 3 * E + head * kD + lane;
      row[E] = dK[i];                              // (sQ already carries the 1/sqrt(d) factor)
      row[2 * E] = dV[i];
    }
  }
}

}  // namespace sa
}  // namespace mpf

extern "C" {

int mpf_self_attn_fwd_f32(const float* qkv, const uint8_t* mask, int B, int Qt, int heads, int head_dim, float* out,
                          float* lse, void* stream) {
  using namespace mpf;
  clear_error();
  MPF_REQUIRE(B > 0 && Qt > 0 && heads > 0, "self_attn: sizes must be positive");
  MPF_REQUIRE(head_dim == sa::kD, "self_attn: head_dim must be %d (got %d)", sa::kD, head_dim);
  MPF_REQUIRE(Qt <= sa::kMaxQ, "self_attn: at most %d queries (got %d)", sa::kMaxQ, Qt);
  MPF_REQUIRE(heads <= 65535 && B <= 65535, "self_attn: grid too large");
  MPF_REQUIRE(qkv && out && lse, "self_attn: null pointer argument");
  static unsigned long long seen = 0;
  if (first_use_on_this_device(seen))
    MPF_CUDA_OK(cudaFuncSetAttribute(sa::self_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(sa::fwd_smem(sa::kMaxQ))));
  sa::self_attn_fwd_kernel<<<dim3(heads, B), sa::kThreads, sa::fwd_smem(Qt), static_cast<cudaStream_t>(stream)>>>(
      qkv, mask, Qt, heads, 1.0f / sqrtf(static_cast<float>(head_dim)), out, lse);
  count_launch();
  return finish_launch("self_attn_fwd");
}

int mpf_self_attn_bwd_f32(const float* qkv, const uint8_t* mask, const float* d_out, const float* lse, int B, int Qt,
                          int heads, int head_dim, float* d_qkv, void* stream) {
  using namespace mpf;
  clear_error();
  MPF_REQUIRE(B > 0 && Qt > 0 && heads > 0, "self_attn_bwd: sizes must be positive");
  MPF_REQUIRE(head_dim == sa::kD, "self_attn_bwd: head_dim must be %d (got %d)", sa::kD, head_dim);
  MPF_REQUIRE(Qt <= sa::kMaxQ, "self_attn_bwd: at most %d queries (got %d)", sa::kMaxQ, Qt);
  MPF_REQUIRE(heads <= 65535 && B <= 65535, "self_attn_bwd: grid too large");
  MPF_REQUIRE(qkv && d_out && lse && d_qkv, "self_attn_bwd: null pointer argument");
  static unsigned long long seen = 0;
  if (first_use_on_this_device(seen))
    MPF_CUDA_OK(cudaFuncSetAttribute(sa::self_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(sa::bwd_smem(sa::kMaxQ))));
  sa::self_attn_bwd_kernel<<<dim3(heads, B), sa::kThreads, sa::bwd_smem(Qt), static_cast<cudaStream_t>(stream)>>>(
      qkv, mask, d_out, lse, Qt, heads, 1.0f / sqrtf(static_cast<float>(head_dim)), d_qkv);
  count_launch();
  return finish_launch("self_attn_bwd");
}

}  // extern "C"

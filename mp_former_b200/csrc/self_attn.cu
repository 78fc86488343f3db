// Self-attention core of the decoder's SelfAttentionLayer over the (few hundred) queries, forward and backward:
//   ref: transformer_decoder/mask2former_transformer_decoder.py:42-52 (nn.MultiheadAttention(tgt, tgt, tgt) with the
//        optional boolean tgt_mask of the mask-piloted groups, decoder :1051-1059; True = not allowed)
//   O = softmax(Q K^T / sqrt(d) + mask) V   per (image, head);  Q, K, V are the slices of the packed in-projection.
// Qt <= 320 queries x 32 head channels: 3.3 MFLOP per (image, head) -- far too small for the tensor pipeline to pay
// (one 128 x 64 UMMA tile is already 20 % of a row block); the kernels are plain fp32 CUDA-core code, exact to fp32
// rounding, one CTA per (image, head):
//   forward : K, V (and Q) of the head in shared memory (row pitch 33: conflict free by key and by channel); a warp
//             owns a query row at a time -- lanes = keys for the scores, shuffles for max / sum, lanes = channels for
//             P V with the probabilities broadcast from shared memory; the log-sum-exp is kept for the backward.
//   backward: probabilities recomputed from the log-sum-exp; rows are processed in chunks of 8: the warps write P
//             and dS = P (dP - delta) of the chunk to shared memory and produce dQ, then every thread adds the chunk
//             into ITS fixed (key, channel) entries of dK and dV held in registers (40 each at Qt = 320) -- no atomics.
// This replaces the library scaled_dot_product_attention call (PyTorch's memory-efficient kernel) of round 1.
#include "mpf_common.cuh"

namespace mpf {
namespace sa {

constexpr int kD = 32;
constexpr int kPitch = kD + 1;
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxQ = 320;
constexpr int kKeysPerLane = kMaxQ / 32;        // 10
constexpr int kChunk = 8;                       // backward: query rows per chunk (= one per warp)
constexpr int kOwn = kMaxQ * kD / kThreads;     // 40 (key, channel) entries of dK / dV per thread

__host__ __device__ constexpr size_t fwd_smem(int Qt) { return (3 * static_cast<size_t>(Qt) * kPitch + kWarps * kMaxQ) * 4; }
__host__ __device__ constexpr size_t bwd_smem(int Qt) {
  return (4 * static_cast<size_t>(Qt) * kPitch + 2 * kChunk * kMaxQ) * 4;
}

// qkv: [B, Qt, 3E] (q | k | v), mask: uint8 [Qt, Qt] or null (1 = not allowed), out: [B, Qt, E], lse: [B, heads, Qt]
__global__ void __launch_bounds__(kThreads)
self_attn_fwd_kernel(const float* __restrict__ qkv, const uint8_t* __restrict__ mask, int Qt, int heads, float scale,
                     float* __restrict__ out, float* __restrict__ lse) {
  extern __shared__ float sm[];
  float* sQ = sm;
  float* sK = sQ + Qt * kPitch;
  float* sV = sK + Qt * kPitch;
  float* sP = sV + Qt * kPitch;                   // [warps][kMaxQ]
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
  float* p = sP + warp * kMaxQ;
  for (int r = warp; r < Qt; r += kWarps) {
    float s[kKeysPerLane];
    float mx = -INFINITY;
#pragma unroll
    for (int u = 0; u < kKeysPerLane; ++u) {
      const int k = lane + 32 * u;
      float acc = -INFINITY;
      if (k < Qt && !(mask != nullptr && mask[static_cast<long long>(r) * Qt + k])) {
        acc = 0.f;
#pragma unroll
        for (int c = 0; c < kD; ++c) acc += sQ[r * kPitch + c] * sK[k * kPitch + c];
      }
      s[u] = acc;
      mx = fmaxf(mx, acc);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
#pragma unroll
    for (int u = 0; u < kKeysPerLane; ++u) {
      s[u] = (s[u] == -INFINITY) ? 0.f : expf(s[u] - mx);
      sum += s[u];
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.f / sum;                  // (a fully masked row gives NaN exactly like the reference's softmax)
#pragma unroll
    for (int u = 0; u < kKeysPerLane; ++u) {
      const int k = lane + 32 * u;
      if (k < Qt) p[k] = s[u] * inv;
    }
    __syncwarp();
    float o = 0.f;
    for (int k = 0; k < Qt; ++k) o += p[k] * sV[k * kPitch + lane];
    out[(static_cast<long long>(b) * Qt + r) * E + head * kD + lane] = o;
    if (lane == 0) lse[(static_cast<long long>(b) * heads + head) * Qt + r] = mx + logf(sum);
    __syncwarp();
  }
}

// d_out: [B, Qt, E]; d_qkv: [B, Qt, 3E] (every element written)
__global__ void __launch_bounds__(kThreads)
self_attn_bwd_kernel(const float* __restrict__ qkv, const uint8_t* __restrict__ mask, const float* __restrict__ d_out,
                     const float* __restrict__ lse, int Qt, int heads, float scale, float* __restrict__ d_qkv) {
  extern __shared__ float sm[];
  float* sQ = sm;                                  // scaled by 1/sqrt(d)
  float* sK = sQ + Qt * kPitch;
  float* sV = sK + Qt * kPitch;
  float* sG = sV + Qt * kPitch;                   // dO
  float* sP = sG + Qt * kPitch;                   // [chunk][kMaxQ] probabilities
  float* sS = sP + kChunk * kMaxQ;                // [chunk][kMaxQ] dS
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
  // this thread's fixed entries of dK / dV: flat index e = tid + kThreads * i -> key e / 32, channel e % 32 (= lane)
  float dK[kOwn], dV[kOwn];
#pragma unroll
  for (int i = 0; i < kOwn; ++i) dK[i] = dV[i] = 0.f;
  __syncthreads();
  for (int r0 = 0; r0 < Qt; r0 += kChunk) {
    const int r = r0 + warp;                       // one row per warp (kChunk == kWarps)
    float* p = sP + warp * kMaxQ;
    float* ds = sS + warp * kMaxQ;
    if (r < Qt) {
      const float l = __ldg(lse + (static_cast<long long>(b) * heads + head) * Qt + r);
      float pv[kKeysPerLane], dp[kKeysPerLane];
      float delta = 0.f;
#pragma unroll
      for (int u = 0; u < kKeysPerLane; ++u) {
        const int k = lane + 32 * u;
        pv[u] = dp[u] = 0.f;
        if (k < Qt && !(mask != nullptr && mask[static_cast<long long>(r) * Qt + k])) {
          float s = 0.f, g = 0.f;
#pragma unroll
          for (int c = 0; c < kD; ++c) {
            s += sQ[r * kPitch + c] * sK[k * kPitch + c];
            g += sG[r * kPitch + c] * sV[k * kPitch + c];
          }
          pv[u] = expf(s - l);
          dp[u] = g;
          delta += pv[u] * g;
        }
      }
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) delta += __shfl_xor_sync(0xffffffffu, delta, o);
#pragma unroll
      for (int u = 0; u < kKeysPerLane; ++u) {
        const int k = lane + 32 * u;
        if (k < Qt) {
          p[k] = pv[u];
          ds[k] = pv[u] * (dp[u] - delta);
        }
      }
      __syncwarp();
      float dq = 0.f;                              // lane = channel
      for (int k = 0; k < Qt; ++k) dq += ds[k] * sK[k * kPitch + lane];
      d_qkv[(img + r) * 3 * E + head * kD + lane] = dq * scale;
    } else {
      for (int k = lane; k < Qt; k += 32) p[k] = ds[k] = 0.f;
    }
    __syncthreads();
    // every thread adds the chunk's rows into its entries of dK (dS^T Qs) and dV (P^T dO)
#pragma unroll
    for (int i = 0; i < kOwn; ++i) {
      const int k = (tid + kThreads * i) >> 5;
      if (k < Qt) {
#pragma unroll
        for (int w = 0; w < kChunk; ++w) {
          const int rr = r0 + w;
          if (rr < Qt) {
            dK[i] += sS[w * kMaxQ + k] * sQ[rr * kPitch + lane];
            dV[i] += sP[w * kMaxQ + k] * sG[rr * kPitch + lane];
          }
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

// Micro-benchmark of the memory-instruction shapes the MSDeformAttn kernels can be built from (B200, sm_100a):
// how many SM cycles one warp instruction costs on the l1tex / shared / L2-atomic paths when every SM is saturated
// with the same pattern.  Stand-alone (no torch):  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lsu_patterns
// lsu_patterns.cu && ./lsu_patterns          -> one JSON line per pattern.
//
// Geometry mimics one (image, head) of the kernels: "pixels" are 128-byte rows (32 fp32 channels of one head) at a
// pitch of 1024 bytes (8 heads interleaved, value layout [S, M, D]) or 128 bytes (head-major / shared-memory tile);
// a warp instruction addresses 1, 2 or 4 different pixels.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int kThreads = 256;
constexpr int kIters = 512;           // instructions of the measured kind per thread
constexpr int kWin = 512;             // pixels of the window a CTA gathers from (64 KB at 128 B: L1-resident)

__device__ __forceinline__ unsigned hash32(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

// pattern 0: LDG.128, 8 lanes per pixel, 4 pixels per instruction, pixel pitch `pitch` floats (current kernels)
// pattern 1: LDG.32, 32 lanes = one pixel per instruction
// pattern 2: LDG.128, 16 lanes = two ADJACENT pixels (256 contiguous bytes), 2 such pairs per instruction
template <int PATTERN>
__global__ void __launch_bounds__(kThreads) ldg_kernel(const float* __restrict__ base, int pitch, int win_pixels,
                                                        float* __restrict__ out, long long* cycles) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* win = base + static_cast<size_t>(blockIdx.x % 64) * win_pixels * pitch;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  unsigned seed = blockIdx.x * 977u + warp * 131u;
  const long long t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < kIters; ++i) {
    seed = seed * 1664525u + 1013904223u;
    if (PATTERN == 0) {
      const unsigned px = hash32(seed + (lane >> 3)) % win_pixels;
      const float4 v = __ldg(reinterpret_cast<const float4*>(win + static_cast<size_t>(px) * pitch + (lane & 7) * 4));
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    } else if (PATTERN == 1) {
      const unsigned px = hash32(seed) % win_pixels;
      acc.x += __ldg(win + static_cast<size_t>(px) * pitch + lane);
    } else {
      const unsigned px = hash32(seed + (lane >> 4)) % (win_pixels - 1);
      const float4 v = __ldg(reinterpret_cast<const float4*>(win + static_cast<size_t>(px) * pitch + (lane & 15) * 4));
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  out[blockIdx.x * kThreads + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

// LDS.128 gather from a shared-memory tile: 8 lanes per pixel row (128 B), 4 rows per instruction
__global__ void __launch_bounds__(kThreads) lds_kernel(float* __restrict__ out, long long* cycles, int pad_floats) {
  extern __shared__ float4 tile[];
  const int row_f4 = 8 + pad_floats / 4;
  for (int i = threadIdx.x; i < kWin * row_f4; i += kThreads) tile[i] = make_float4(i, 1.f, 2.f, 3.f);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  unsigned seed = blockIdx.x * 977u + warp * 131u;
  const long long t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < kIters; ++i) {
    seed = seed * 1664525u + 1013904223u;
    const unsigned px = hash32(seed + (lane >> 3)) % kWin;
    const float4 v = tile[px * row_f4 + (lane & 7)];
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  out[blockIdx.x * kThreads + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

// pattern 0: RED.128 (red.global.add.v4.f32), 8 lanes per pixel, 4 pixels per instruction (current backward)
// pattern 1: RED.32, 32 lanes = one pixel
// pattern 2: RED.128, 16 lanes = two adjacent pixels
template <int PATTERN>
__global__ void __launch_bounds__(kThreads) red_kernel(float* __restrict__ base, int pitch, int win_pixels,
                                                        long long* cycles) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* win = base + static_cast<size_t>(blockIdx.x % 64) * win_pixels * pitch;
  unsigned seed = blockIdx.x * 977u + warp * 131u;
  const long long t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < kIters; ++i) {
    seed = seed * 1664525u + 1013904223u;
    if (PATTERN == 0) {
      const unsigned px = hash32(seed + (lane >> 3)) % win_pixels;
      float* a = win + static_cast<size_t>(px) * pitch + (lane & 7) * 4;
      asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(a), "f"(1.0f) : "memory");
    } else if (PATTERN == 1) {
      const unsigned px = hash32(seed) % win_pixels;
      float* a = win + static_cast<size_t>(px) * pitch + lane;
      asm volatile("red.global.add.f32 [%0], %1;" ::"l"(a), "f"(1.0f) : "memory");
    } else {
      const unsigned px = hash32(seed + (lane >> 4)) % (win_pixels - 1);
      float* a = win + static_cast<size_t>(px) * pitch + (lane & 15) * 4;
      asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(a), "f"(1.0f) : "memory");
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// shared-memory accumulation of a 128-byte contribution per 8-lane group, 4 groups (pixels) per instruction group:
// pattern 0: 4 x red.shared.add.s32 per lane (fixed point), channel rotation per group so banks do not collide
// pattern 1: plain LDS.128 + FADD + STS.128 read-modify-write (NOT race free across warps: cost reference only)
template <int PATTERN>
__global__ void __launch_bounds__(kThreads) smem_acc_kernel(float* __restrict__ out, long long* cycles) {
  extern __shared__ float4 tile[];
  for (int i = threadIdx.x; i < kWin * 8; i += kThreads) tile[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned seed = blockIdx.x * 977u + warp * 131u;
  const long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < kIters; ++i) {
    seed = seed * 1664525u + 1013904223u;
    const unsigned px = hash32(seed + (lane >> 3)) % kWin;
    if (PATTERN == 0) {
      int* row = reinterpret_cast<int*>(tile + px * 8 + (lane & 7));
      const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(row));
      const int g = lane >> 3;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(a + 4u * ((k + g) & 3)), "r"(i + k) : "memory");
    } else {
      float4 v = tile[px * 8 + (lane & 7)];
      v.x += 1.f; v.y += 2.f; v.z += 3.f; v.w += 4.f;
      tile[px * 8 + (lane & 7)] = v;
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  out[blockIdx.x * kThreads + threadIdx.x] = tile[threadIdx.x].x;
}

static double run_report(const char* name, const char* what, int ctas_per_sm, int n_sm, long long* d_cycles,
                         float ms, int instr_per_iter) {
  const int grid = ctas_per_sm * n_sm;
  long long* h = static_cast<long long*>(malloc(grid * sizeof(long long)));
  CK(cudaMemcpy(h, d_cycles, grid * sizeof(long long), cudaMemcpyDeviceToHost));
  double avg = 0;
  for (int i = 0; i < grid; ++i) avg += h[i];
  avg /= grid;
  free(h);
  // per SM: ctas_per_sm CTAs x 8 warps each issue kIters*instr_per_iter instructions within `avg` cycles
  const double cyc_per_instr = avg / (static_cast<double>(kIters) * instr_per_iter * 8 * ctas_per_sm);
  printf("{\"pattern\": \"%s\", \"what\": \"%s\", \"ctas_per_sm\": %d, \"sm_cycles_per_warp_instruction\": %.2f, "
         "\"kernel_ms\": %.4f}\n", name, what, ctas_per_sm, cyc_per_instr, ms);
  return cyc_per_instr;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int n_sm = prop.multiProcessorCount;
  const int ctas = 4;
  const int grid = ctas * n_sm;
  const size_t buf_floats = static_cast<size_t>(64) * kWin * 256;      // 64 windows x 512 pixels x 1024 B = 32 MB
  float *buf, *out;
  long long* cyc;
  CK(cudaMalloc(&buf, buf_floats * 4));
  CK(cudaMemset(buf, 0, buf_floats * 4));
  CK(cudaMalloc(&out, static_cast<size_t>(grid) * kThreads * 4));
  CK(cudaMalloc(&cyc, grid * sizeof(long long)));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float ms;
#define TIME(launch)                                   \
  launch; CK(cudaDeviceSynchronize());                 \
  CK(cudaEventRecord(e0)); launch; CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); \
  CK(cudaGetLastError()); CK(cudaEventElapsedTime(&ms, e0, e1));

  TIME((ldg_kernel<0><<<grid, kThreads>>>(buf, 256, kWin, out, cyc)));
  run_report("ldg128_4px_pitch1024", "LDG.128, 8 lanes/pixel, 4 pixels per instruction, [S,M,D] pitch (current)", ctas, n_sm, cyc, ms, 1);
  TIME((ldg_kernel<0><<<grid, kThreads>>>(buf, 32, kWin, out, cyc)));
  run_report("ldg128_4px_pitch128", "LDG.128, 8 lanes/pixel, 4 pixels per instruction, head-major pitch", ctas, n_sm, cyc, ms, 1);
  TIME((ldg_kernel<1><<<grid, kThreads>>>(buf, 256, kWin, out, cyc)));
  run_report("ldg32_1px", "LDG.32, 32 lanes = one pixel per instruction", ctas, n_sm, cyc, ms, 1);
  TIME((ldg_kernel<2><<<grid, kThreads>>>(buf, 32, kWin, out, cyc)));
  run_report("ldg128_2pairs_pitch128", "LDG.128, 16 lanes = 2 adjacent pixels (256 B), 2 pairs per instruction", ctas, n_sm, cyc, ms, 1);

  CK(cudaFuncSetAttribute(lds_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  for (int occ = 1; occ <= 3; ++occ) {
    const int g = occ * n_sm;
    TIME((lds_kernel<<<g, kThreads, kWin * 128>>>(out, cyc, 0)));
    run_report("lds128_4rows", "LDS.128 from a 64 KB shared tile, 8 lanes/row, 4 rows per instruction", occ, n_sm, cyc, ms, 1);
  }
  TIME((lds_kernel<<<2 * n_sm, kThreads, kWin * 144>>>(out, cyc, 4)));
  run_report("lds128_4rows_pad16", "same, rows padded to 144 B", 2, n_sm, cyc, ms, 1);

  TIME((red_kernel<0><<<grid, kThreads>>>(buf, 256, kWin, cyc)));
  run_report("red128_4px_pitch1024", "RED.128 v4.f32, 8 lanes/pixel, 4 pixels per instruction (current backward)", ctas, n_sm, cyc, ms, 1);
  TIME((red_kernel<0><<<grid, kThreads>>>(buf, 32, kWin, cyc)));
  run_report("red128_4px_pitch128", "RED.128 v4.f32, head-major pitch", ctas, n_sm, cyc, ms, 1);
  TIME((red_kernel<1><<<grid, kThreads>>>(buf, 256, kWin, cyc)));
  run_report("red32_1px", "RED.32, 32 lanes = one pixel per instruction", ctas, n_sm, cyc, ms, 1);
  TIME((red_kernel<2><<<grid, kThreads>>>(buf, 32, kWin, cyc)));
  run_report("red128_2pairs_pitch128", "RED.128, 16 lanes = 2 adjacent pixels, 2 pairs per instruction", ctas, n_sm, cyc, ms, 1);

  CK(cudaFuncSetAttribute(smem_acc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  CK(cudaFuncSetAttribute(smem_acc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  TIME((smem_acc_kernel<0><<<2 * n_sm, kThreads, kWin * 128>>>(out, cyc)));
  run_report("atoms_s32_x4", "4 x red.shared.add.s32 per lane = one 128 B contribution per 8 lanes (cycles per ATOMS instruction)", 2, n_sm, cyc, ms, 4);
  TIME((smem_acc_kernel<1><<<2 * n_sm, kThreads, kWin * 128>>>(out, cyc)));
  run_report("lds_fadd_sts", "LDS.128 + 4 FADD + STS.128 read-modify-write (racy; cost reference), per LDS+STS pair", 2, n_sm, cyc, ms, 1);
  return 0;
}

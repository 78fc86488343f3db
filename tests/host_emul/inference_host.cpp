// TEST INFRASTRUCTURE ONLY: runs the per-pixel arithmetic of csrc/inference.cu (the very header the kernel includes,
// mp_former_b200/csrc/inference_math.cuh) on the host, so that its index / weight / crop logic is pinned against torch's
// interpolate chain without a GPU.  Built by tests/test_inference_cpu.py with
//   g++ -O2 -ffp-contract=off -shared -fPIC -D__host__= -D__device__= -D__forceinline__=inline
#include <cmath>

#include "../../mp_former_b200/csrc/inference_math.cuh"

extern "C" void host_instance_masks(const float* logits, long long q_stride, int h, int w, const long long* query_index,
                                    int rows, int padded_h, int padded_w, int image_h, int image_w, int out_h,
                                    int out_w, unsigned char* out_masks, float* values, double* sums) {
  const mpf::TwoStage ts = mpf::make_two_stage(h, w, padded_h, padded_w, image_h, image_w, out_h, out_w);
  for (int r = 0; r < rows; ++r) {
    const float* L = logits + query_index[r] * q_stride;
    double prob = 0.0, fg = 0.0;
    for (int y = 0; y < out_h; ++y)
      for (int x = 0; x < out_w; ++x) {
        const float v = mpf::two_stage_at(L, ts, y, x);
        const long long o = (static_cast<long long>(r) * out_h + y) * out_w + x;
        values[o] = v;
        out_masks[o] = v > 0.f ? 1 : 0;
        if (v > 0.f) {
          fg += 1.0;
          prob += 1.0 / (1.0 + std::exp(-static_cast<double>(v)));
        }
      }
    sums[2 * r] = prob;
    sums[2 * r + 1] = fg;
  }
}

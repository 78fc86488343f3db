// Host-side creation of TMA tensor maps (cuTensorMapEncodeTiled obtained through the runtime's
// driver entry point, so the library has no link-time dependency on libcuda).
#pragma once

#include <cuda.h>

#include "mpf_common.cuh"

namespace mpf {

// 3-D fp32 tensor [d2, d1, d0] with d0 contiguous; strides in ELEMENTS (ld1 between d1 rows, ld2 between
// d2 slabs); box [1, box1, box0]; 128-byte swizzle (box0 * 4 bytes must be 128); OOB reads are zero.
// atom32 = true selects CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B (needed for MN-major tf32 MMA operands).
int make_tmap_f32_3d(CUtensorMap* m, const float* base, long long d0, long long d1, long long d2,
                     long long ld1, long long ld2, int box0, int box1, bool atom32 = false);

int sm_count();

}  // namespace mpf

// Shared device/host helpers for libalignsdf_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/alignsdf_b200.h"

namespace asdf {

void set_error(const char* fmt, ...);

#define ASDF_CUDA_CHECK(expr)                                                         \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      asdf::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return ASDF_ERR_CUDA;                                                           \
    }                                                                                 \
  } while (0)

#define ASDF_REQUIRE(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      asdf::set_error(__VA_ARGS__);        \
      return ASDF_ERR_ARG;                 \
    }                                      \
  } while (0)

// Query coordinates of linear grid index i, bit-exact w.r.t. the reference's torch-CPU
// expressions (utils/mesh.py:32-40, 86-94):  mode 0 reproduces the true-division shear,
//   c2 = float(i % N); q = float(i)/float(N); c1 = fmod(q, N); c0 = fmod(q/N, N)
// then c*voxel + origin as a separate multiply and add (no FMA contraction).
__device__ __forceinline__ void grid_point(int64_t i, int N, int mode, float voxel,
                                           float o0, float o1, float o2,
                                           float& x0, float& x1, float& x2) {
  float c0, c1, c2;
  c2 = (float)(int)(i % N);
  if (mode == ASDF_QUERY_GRID_REFERENCE) {
    const float fN = (float)N;
    const float q = __fdiv_rn(__ll2float_rn(i), fN);
    c1 = fmodf(q, fN);
    c0 = fmodf(__fdiv_rn(q, fN), fN);
  } else {
    c1 = (float)(int)((i / N) % N);
    c0 = (float)(int)((i / ((int64_t)N * N)) % N);
  }
  x0 = __fadd_rn(__fmul_rn(c0, voxel), o0);
  x1 = __fadd_rn(__fmul_rn(c1, voxel), o1);
  x2 = __fadd_rn(__fmul_rn(c2, voxel), o2);
}

// NeRF positional encoding of one coordinate triple (utils/utils.py:433-463): feature f of
// [x(3), sin(x 2^0)(3), cos(x 2^0)(3), sin(x 2^1)(3), ...]; x * freq is a separate fp32 multiply like
// torch's, sinf / cosf are the accurate (range-reducing) versions.
__device__ __forceinline__ float nerf_feature(int f, float x0, float x1, float x2) {
  const int a = f % 3;
  const float x = a == 0 ? x0 : (a == 1 ? x1 : x2);
  if (f < 3) return x;
  const int g = (f - 3) / 3;                       // 0: sin 2^0, 1: cos 2^0, 2: sin 2^1, ...
  const float arg = __fmul_rn(x, (float)(1 << (g >> 1)));
  return (g & 1) ? cosf(arg) : sinf(arg);
}

// Warp-aggregated update of the 6-int bounding box of negative samples
// (utils/mesh.py:207-247: nonzero(sdf<0) -> per-axis min/max over the *unravelled* index).
__device__ __forceinline__ void bbox_update(int32_t* bbox, bool neg, int64_t i, int N) {
  const unsigned full = 0xffffffffu;
  int a0 = 0x7fffffff, a1 = 0x7fffffff, a2 = 0x7fffffff, b0 = -1, b1 = -1, b2 = -1;
  if (neg) {
    a2 = b2 = (int)(i % N);
    a1 = b1 = (int)((i / N) % N);
    a0 = b0 = (int)(i / ((int64_t)N * N));
  }
  if (!__any_sync(full, neg)) return;
  a0 = __reduce_min_sync(full, a0); a1 = __reduce_min_sync(full, a1); a2 = __reduce_min_sync(full, a2);
  b0 = __reduce_max_sync(full, b0); b1 = __reduce_max_sync(full, b1); b2 = __reduce_max_sync(full, b2);
  if ((threadIdx.x & 31) == 0) {
    atomicMin(bbox + 0, a0); atomicMin(bbox + 1, a1); atomicMin(bbox + 2, a2);
    atomicMax(bbox + 3, b0); atomicMax(bbox + 4, b1); atomicMax(bbox + 5, b2);
  }
}

}  // namespace asdf

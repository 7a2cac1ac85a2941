// Per-sample set-up of the tensor-core path on the device (no host numpy, no host round trips):
//
//   asdf_tc_bind   latent + pose-align affine of S samples -> the per-sample "P tile" blocks k1_tc.cu streams:
//                  fold the latent columns of layers 0 / 2 into biases and the pose-align feature columns into
//                  [512,3] point matrices (float64; SURVEY.md App. A -- what utils/utils.py:376-430,561-572 and the
//                  first / skip layer of networks/model.py:285-350 recompute for every query point), choose the
//                  power-of-two operand scales, split into fp16 hi + lo and write the swizzled tiles.
//   asdf_regrid    bounding boxes of pass 1 -> lattice of pass 2 (utils/mesh.py:198-256 get_higher_res_cube), the
//                  same f32 arithmetic on 4 scalars, written where pass 2 and marching cubes read it.
#include "common.cuh"
#include <cuda_fp16.h>
#include <math.h>

namespace asdf {
namespace bind {

constexpr int kTileBytes = 64 * 64 * 2;
constexpr int kPTiles = 14;
constexpr int64_t kSampleTileBytes = (int64_t)2 * 2 * kPTiles * kTileBytes;
constexpr double kF16Safe = 16384.0;

// fold[s][d][j][n][4] = (M[n][0..2], B[n]) of layer 0 (j = 0) / layer 2 (j = 1), float64
__global__ void __launch_bounds__(256) fold_kernel(const asdf_tc_bind_desc desc, const double* __restrict__ stat,
                                                   const float* __restrict__ latent, const double* __restrict__ affine,
                                                   double* __restrict__ fold) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + warp;                 // output row
  const int j = blockIdx.y;                            // 0: layer 0, 1: layer 2
  const int s = blockIdx.z / desc.n_decoders, d = blockIdx.z % desc.n_decoders;
  const int L = desc.latent_size, nf = desc.n_features[d];
  const double* base = stat + (int64_t)d * desc.decoder_stride;
  const double* wz = base + ((int64_t)j * 512 + n) * L;
  const double* wf = base + (int64_t)2 * 512 * L + ((int64_t)j * 512 + n) * ASDF_MAX_POINT_DIM;
  const double* b = base + (int64_t)2 * 512 * L + (int64_t)2 * 512 * ASDF_MAX_POINT_DIM + (int64_t)(2 * j) * 512;
  const float* z = latent + (int64_t)s * L;
  const double* ac = affine + (int64_t)s * ASDF_MAX_POINT_DIM * 4;
  double m0 = 0, m1 = 0, m2 = 0, bb = 0;
  for (int k = lane; k < L; k += 32) bb = fma(wz[k], (double)z[k], bb);
  for (int f = lane; f < nf; f += 32) {
    const double w = wf[f];
    const double* r = ac + 4 * desc.feature_index[d][f];
    m0 = fma(w, r[0], m0); m1 = fma(w, r[1], m1); m2 = fma(w, r[2], m2); bb = fma(w, r[3], bb);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m0 += __shfl_xor_sync(0xffffffffu, m0, o); m1 += __shfl_xor_sync(0xffffffffu, m1, o);
    m2 += __shfl_xor_sync(0xffffffffu, m2, o); bb += __shfl_xor_sync(0xffffffffu, bb, o);
  }
  if (lane == 0) {
    double* o = fold + ((((int64_t)s * desc.n_decoders + d) * 2 + j) * 512 + n) * 4;
    o[0] = m0; o[1] = m1; o[2] = m2; o[3] = bb + b[n];
  }
}

__device__ __forceinline__ double pow2_floor(double x) { int e; frexp(x, &e); return ldexp(1.0, e - 1); }
__device__ __forceinline__ double pow2_ceil(double x) { int e; const double m = frexp(x, &e); return ldexp(1.0, m == 0.5 ? e - 1 : e); }

// one block per sample: operand scales (the rule of tc_pack.choose_point_scales), fp16 split, swizzled P tiles
__global__ void __launch_bounds__(512) tiles_kernel(const asdf_tc_bind_desc desc, const double* __restrict__ stat,
                                                    const double* __restrict__ fold, uint8_t* __restrict__ samples,
                                                    int64_t sample_stride, int32_t* __restrict__ status) {
  const int s = blockIdx.x, tid = threadIdx.x, nd = desc.n_decoders, L = desc.latent_size;
  __shared__ double red[2][6][16];      // [decoder][maxM0, maxB0, maxM2, maxB2, maxB1, maxB3][warp]
  __shared__ double sc[8];              // cp, c1, S0[2], ok
  const double t = (double)desc.act_scale;
  auto layer_row = [&](int d, int l, int n, double& m0, double& m1, double& m2, double& b) {
    if (l == 0 || l == 2) {
      const double* r = fold + ((((int64_t)s * nd + d) * 2 + (l >> 1)) * 512 + n) * 4;
      m0 = r[0]; m1 = r[1]; m2 = r[2]; b = r[3];
    } else {
      const double* bs = stat + (int64_t)d * desc.decoder_stride + (int64_t)2 * 512 * L + (int64_t)2 * 512 * ASDF_MAX_POINT_DIM;
      m0 = m1 = m2 = 0.0; b = bs[(int64_t)l * 512 + n];       // b1 is zero padded to 256 rows
    }
  };
  // ---- maxima ----
  for (int d = 0; d < nd; ++d) {
    double mx[6] = {0, 0, 0, 0, 0, 0};
    for (int n = tid; n < 512; n += 512) {
      double m0, m1, m2, b;
      layer_row(d, 0, n, m0, m1, m2, b);
      mx[0] = fmax(fmax(fabs(m0), fabs(m1)), fabs(m2)); mx[1] = fabs(b);
      layer_row(d, 2, n, m0, m1, m2, b);
      mx[2] = fmax(fmax(fabs(m0), fabs(m1)), fabs(m2)); mx[3] = fabs(b);
      if (n < 256) { layer_row(d, 1, n, m0, m1, m2, b); mx[4] = fabs(b); }
      layer_row(d, 3, n, m0, m1, m2, b); mx[5] = fabs(b);
    }
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      double v = mx[q];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
      if ((tid & 31) == 0) red[d][q][tid >> 5] = v;
    }
  }
  __syncthreads();
  if (tid == 0) {
    double mxs[2][6];
    for (int d = 0; d < nd; ++d)
      for (int q = 0; q < 6; ++q) { double v = 0; for (int w = 0; w < 16; ++w) v = fmax(v, red[d][q][w]); mxs[d][q] = v; }
    const double p_absmax = fmax((double)desc.p_absmax, 1e-3);
    const double cp_max = pow2_floor(60000.0 / p_absmax);
    double cp_min = 1.0, c1_min = 1.0;
    for (int d = 0; d < nd; ++d) {
      const double S1 = t * desc.w_scale[d][0], S2 = t * desc.w_scale[d][1], S3 = t * desc.w_scale[d][2];
      if (mxs[d][2] > 0) cp_min = fmax(cp_min, S2 * mxs[d][2] / kF16Safe);
      if (mxs[d][4] > 0) c1_min = fmax(c1_min, S1 * mxs[d][4] / kF16Safe);
      if (mxs[d][3] > 0) c1_min = fmax(c1_min, S2 * mxs[d][3] / kF16Safe);
      if (mxs[d][5] > 0) c1_min = fmax(c1_min, S3 * mxs[d][5] / kF16Safe);
    }
    double cp = pow2_ceil(cp_min), c1 = pow2_ceil(c1_min);
    bool ok = isfinite(cp) && isfinite(c1) && cp <= cp_max && c1 <= 32768.0;
    cp = fmax(cp, fmin(cp_max, 1024.0));        // prefer a large cp: more headroom for the lo part of p
    c1 = fmax(c1, 1024.0);
    sc[0] = cp; sc[1] = c1;
    for (int d = 0; d < nd; ++d) {
      const double lim = fmin(kF16Safe * cp / fmax(mxs[d][0], 1e-30), kF16Safe * c1 / fmax(mxs[d][1], 1e-30));
      ok = ok && isfinite(lim) && lim > 0;
      sc[2 + d] = ok ? pow2_floor(lim) : 1.0;
    }
    sc[4] = ok ? 1.0 : 0.0;
  }
  __syncthreads();
  const double cp = sc[0], c1 = sc[1];
  bool ok = sc[4] != 0.0;
  uint8_t* blk = samples + (int64_t)s * sample_stride;
  // ---- tiles: 1792 rows per decoder ----
  for (int d = 0; d < nd; ++d) {
    for (int idx = tid; idx < 1792; idx += 512) {
      int l, n, g0;
      if (idx < 512) { l = 0; n = idx; g0 = 0; }
      else if (idx < 768) { l = 1; n = idx - 512; g0 = 4; }
      else if (idx < 1280) { l = 2; n = idx - 768; g0 = 6; }
      else { l = 3; n = idx - 1280; g0 = 10; }
      double m[3], b;
      layer_row(d, l, n, m[0], m[1], m[2], b);
      const double S = l == 0 ? sc[2 + d] : t * desc.w_scale[d][l - 1];
      __half hh[4], hl[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const double v = q < 3 ? (S / cp) * m[q] : (S / c1) * b;
        if (!(fabs(v) <= 60000.0)) ok = false;
        hh[q] = __double2half(v);
        hl[q] = __double2half(v - (double)__half2float(hh[q]));
      }
      const int g = g0 + (n >> 7), c = (n & 127) >> 6, r = n & 63;
      uint8_t* row = blk + ((int64_t)(d * 2 + c) * kPTiles + g) * kTileBytes + (r >> 3) * 1024 + (r & 7) * 128;
      auto u16 = [](__half h) { return (uint32_t)__half_as_ushort(h); };
      const uint4 c0 = make_uint4(u16(hh[0]) | (u16(hh[1]) << 16), u16(hh[2]) | (u16(hh[3]) << 16),
                                  u16(hh[0]) | (u16(hh[1]) << 16), u16(hh[2]));
      const uint4 c1v = make_uint4(u16(hl[0]) | (u16(hl[1]) << 16), u16(hl[2]) | (u16(hl[3]) << 16), 0u, 0u);
      *reinterpret_cast<uint4*>(row + ((0 ^ (r & 7)) << 4)) = c0;       // k 0..7 : M_h B_h | M_h 0
      *reinterpret_cast<uint4*>(row + ((1 ^ (r & 7)) << 4)) = c1v;      // k 8..15: M_l B_l | 0
    }
  }
  if (tid == 0) {
    float* scal = reinterpret_cast<float*>(blk + kSampleTileBytes);
    scal[0] = (float)(t / sc[2]); scal[1] = (float)(t / sc[3]);
    scal[2] = (float)cp; scal[3] = (float)c1;
  }
  if (!ok) atomicOr(status, 2);
}

// utils/mesh.py:198-256 on the device.  box: int32[12] (hand min/max, object min/max) as asdf_*_eval leaves it.
__global__ void regrid_kernel(const int32_t* __restrict__ bbox, int n_samples, int mask, int N, float voxel,
                              float* __restrict__ grid, float* __restrict__ minmax) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_samples) return;
  const int32_t* b = bbox + 12 * s;
  float mn[3], mx[3];
  bool first = true;
  for (int o = 0; o < 2; ++o) {
    if (!(mask >> o & 1)) continue;
    float lo[3], hi[3];
    const bool empty = b[6 * o + 3] < 0;                  // no negative sample: the reference uses zeros (:209-211)
    for (int k = 0; k < 3; ++k) { lo[k] = empty ? 0.f : (float)b[6 * o + k]; hi[k] = empty ? 0.f : (float)b[6 * o + 3 + k]; }
    for (int k = 0; k < 3; ++k) {
      mn[k] = first ? lo[k] : fminf(mn[k], lo[k]);
      mx[k] = first ? hi[k] : fmaxf(mx[k], hi[k]);
    }
    first = false;
  }
  // new_cube_size = (max(max - min) + 4) * voxel;  new_voxel = new_cube_size / (N - 1);  new_origin = (min - 2) * voxel - 1
  const float ext = fmaxf(fmaxf(__fsub_rn(mx[0], mn[0]), __fsub_rn(mx[1], mn[1])), __fsub_rn(mx[2], mn[2]));
  const float cube = __fmul_rn(__fadd_rn(ext, 4.f), voxel);
  float4 g;
  g.x = __fdiv_rn(cube, (float)(N - 1));
  g.y = __fsub_rn(__fmul_rn(__fsub_rn(mn[0], 2.f), voxel), 1.f);
  g.z = __fsub_rn(__fmul_rn(__fsub_rn(mn[1], 2.f), voxel), 1.f);
  g.w = __fsub_rn(__fmul_rn(__fsub_rn(mn[2], 2.f), voxel), 1.f);
  reinterpret_cast<float4*>(grid)[s] = g;
  if (minmax) for (int k = 0; k < 3; ++k) { minmax[6 * s + k] = mn[k]; minmax[6 * s + 3 + k] = mx[k]; }
}

}  // namespace bind
}  // namespace asdf

extern "C" int64_t asdf_tc_bind_static_doubles(int32_t n_decoders, int32_t latent_size) {
  return (int64_t)n_decoders * ((int64_t)2 * 512 * latent_size + (int64_t)2 * 512 * ASDF_MAX_POINT_DIM + 4 * 512);
}

extern "C" int asdf_tc_bind(const asdf_tc_bind_desc* desc, const double* static_dev, const float* latent_dev,
                            const double* affine_dev, int32_t n_samples, double* fold_scratch_dev,
                            void* samples_dev, int64_t sample_stride, int32_t* status_dev, void* stream) {
  using namespace asdf;
  ASDF_REQUIRE(desc && static_dev && latent_dev && affine_dev && fold_scratch_dev && samples_dev && status_dev,
               "asdf_tc_bind: null argument");
  ASDF_REQUIRE(desc->n_decoders == 1 || desc->n_decoders == 2, "asdf_tc_bind: n_decoders must be 1 or 2");
  ASDF_REQUIRE(desc->latent_size >= 1 && desc->latent_size <= 512, "asdf_tc_bind: bad latent size");
  for (int d = 0; d < desc->n_decoders; ++d) {
    ASDF_REQUIRE(desc->n_features[d] >= 1 && desc->n_features[d] <= ASDF_MAX_POINT_DIM, "asdf_tc_bind: bad feature count");
    for (int f = 0; f < desc->n_features[d]; ++f)
      ASDF_REQUIRE(desc->feature_index[d][f] >= 0 && desc->feature_index[d][f] < ASDF_MAX_POINT_DIM, "asdf_tc_bind: bad feature index");
  }
  ASDF_REQUIRE(desc->decoder_stride * desc->n_decoders == asdf_tc_bind_static_doubles(desc->n_decoders, desc->latent_size),
               "asdf_tc_bind: decoder_stride does not match the static layout");
  ASDF_REQUIRE(n_samples >= 0 && sample_stride >= bind::kSampleTileBytes + 64 && (sample_stride & 15) == 0 &&
               ((uintptr_t)samples_dev & 15) == 0, "asdf_tc_bind: bad sample blocks");
  if (n_samples == 0) return ASDF_OK;
  const cudaStream_t st = (cudaStream_t)stream;
  bind::fold_kernel<<<dim3(64, 2, (unsigned)(n_samples * desc->n_decoders)), 256, 0, st>>>(*desc, static_dev, latent_dev,
                                                                                          affine_dev, fold_scratch_dev);
  ASDF_CUDA_CHECK(cudaGetLastError());
  bind::tiles_kernel<<<(unsigned)n_samples, 512, 0, st>>>(*desc, static_dev, fold_scratch_dev, (uint8_t*)samples_dev,
                                                         sample_stride, status_dev);
  ASDF_CUDA_CHECK(cudaGetLastError());
  return ASDF_OK;
}

extern "C" int asdf_regrid(const int32_t* bbox_dev, int32_t n_samples, int32_t branch_mask, int32_t N, float voxel,
                           float* grid_dev, float* minmax_dev, void* stream) {
  using namespace asdf;
  ASDF_REQUIRE(bbox_dev && grid_dev && n_samples >= 0 && N >= 2, "asdf_regrid: bad argument");
  ASDF_REQUIRE((branch_mask & 3) != 0, "asdf_regrid: at least one of the hand / object branches must be set");
  ASDF_REQUIRE(((uintptr_t)grid_dev & 15) == 0, "asdf_regrid: grid_dev must be 16-byte aligned");
  if (n_samples == 0) return ASDF_OK;
  bind::regrid_kernel<<<(unsigned)((n_samples + 63) / 64), 64, 0, (cudaStream_t)stream>>>(bbox_dev, n_samples, branch_mask, N,
                                                                                          voxel, grid_dev, minmax_dev);
  ASDF_CUDA_CHECK(cudaGetLastError());
  return ASDF_OK;
}

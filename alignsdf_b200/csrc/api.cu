// Library-level entry points: error reporting, device probe, affine pose-align embedding.
#include "common.cuh"

namespace asdf {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

namespace {
// feats[p][f] = A[f][0..2] . xyz[p] + A[f][3]   (utils/utils.py:376-430 folded)
__global__ void embed_kernel(const float* __restrict__ xyz, int64_t P, const float* __restrict__ aff,
                             int pf, float* __restrict__ feats) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P * pf) return;
  const int64_t p = idx / pf;
  const int f = (int)(idx % pf);
  const float x = __ldg(xyz + p * 3), y = __ldg(xyz + p * 3 + 1), z = __ldg(xyz + p * 3 + 2);
  const float4 a = *reinterpret_cast<const float4*>(aff + 4 * f);
  feats[idx] = fmaf(a.x, x, fmaf(a.y, y, fmaf(a.z, z, a.w)));
}
// feats[p][f] for the public get_nerf_embedder() API
__global__ void nerf_embed_kernel(const float* __restrict__ xyz, int64_t P, int pf, float* __restrict__ feats) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P * pf) return;
  const int64_t p = idx / pf;
  feats[idx] = nerf_feature((int)(idx % pf), __ldg(xyz + p * 3), __ldg(xyz + p * 3 + 1), __ldg(xyz + p * 3 + 2));
}
__global__ void grid_points_kernel(asdf_query q, float* __restrict__ xyz) {
  const int64_t i = q.begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= q.end) return;
  float x0, x1, x2;
  grid_point(i, q.N, q.mode, q.voxel, q.origin[0], q.origin[1], q.origin[2], x0, x1, x2);
  float* o = xyz + (i - q.begin) * 3;
  o[0] = x0; o[1] = x1; o[2] = x2;
}
}  // namespace
}  // namespace asdf

extern "C" int asdf_nerf_embed(const float* xyz_dev, int64_t P, int32_t n_freqs, float* feats_dev, void* stream) {
  using namespace asdf;
  ASDF_REQUIRE(P >= 0 && n_freqs >= 0 && n_freqs <= 24, "asdf_nerf_embed: bad sizes");
  if (P == 0) return ASDF_OK;
  ASDF_REQUIRE(xyz_dev && feats_dev, "asdf_nerf_embed: null argument");
  const int pf = 3 + 6 * n_freqs;
  const int64_t n = P * pf;
  nerf_embed_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(xyz_dev, P, pf, feats_dev);
  ASDF_CUDA_CHECK(cudaGetLastError());
  return ASDF_OK;
}

extern "C" int asdf_grid_points(const asdf_query* q, float* xyz_dev, void* stream) {
  using namespace asdf;
  ASDF_REQUIRE(q && xyz_dev, "asdf_grid_points: null argument");
  ASDF_REQUIRE(q->mode == ASDF_QUERY_GRID_REFERENCE || q->mode == ASDF_QUERY_GRID_REGULAR, "bad query mode");
  ASDF_REQUIRE(q->N >= 2 && q->begin >= 0 && q->end >= q->begin && q->end <= (int64_t)q->N * q->N * q->N, "bad range");
  const int64_t n = q->end - q->begin;
  if (n == 0) return ASDF_OK;
  grid_points_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*q, xyz_dev);
  ASDF_CUDA_CHECK(cudaGetLastError());
  return ASDF_OK;
}

extern "C" int asdf_embed_points(const float* xyz_dev, int64_t P, const float* affine_dev, int32_t pf,
                                 float* feats_dev, void* stream) {
  using namespace asdf;
  ASDF_REQUIRE(xyz_dev && affine_dev && feats_dev && P >= 0 && pf >= 1, "asdf_embed_points: bad argument");
  if (P == 0) return ASDF_OK;
  const int64_t total = P * pf;
  embed_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(xyz_dev, P, affine_dev, pf, feats_dev);
  ASDF_CUDA_CHECK(cudaGetLastError());
  return ASDF_OK;
}

extern "C" int asdf_abi_version(void) { return ASDF_ABI_VERSION; }

extern "C" const char* asdf_last_error(void) { return asdf::g_err; }

extern "C" int asdf_device_ok(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    asdf::set_error("no CUDA device visible");
    return 0;
  }
  int dev = 0, major = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10) {
    asdf::set_error("device compute capability %d.x is not sm_100", major);
    return 0;
  }
  return 1;
}

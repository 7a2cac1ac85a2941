// Exact nearest neighbour between two point sets in float64 (brute force) -- the primitive behind the step after
// the hot path: ICP scale / translation alignment (deep_sdf/metrics/icp_trans_scale.py:32-113: two KDTree queries
// per iteration over 30 k sampled points) and the symmetric Chamfer distance (deep_sdf/metrics/chamfer.py:217-231).
// The reference builds sklearn / scipy KD-trees on the CPU; 30 k x 30 k squared distances are 0.9 G pairs, ~6 fp64
// instructions each: a fraction of a millisecond of the FP64 pipe, so no tree is built.
//
//   block = 64 queries x 4 parts (256 threads); the reference points stream through shared memory in tiles of
//   1024 (24 KiB of doubles, read as broadcasts); part p of a query scans every 4th point of a tile; the four
//   partial results meet by shuffles.  Ties go to the smallest index.
#include "common.cuh"

namespace asdf {
namespace {

constexpr int NN_QUERIES = 64, NN_PARTS = 4, NN_TILE = 1024;

__global__ void __launch_bounds__(NN_QUERIES * NN_PARTS) nn_search_kernel(const double* __restrict__ query, int64_t nq,
                                                                          const double* __restrict__ ref, int64_t nr,
                                                                          int32_t* __restrict__ idx_out,
                                                                          double* __restrict__ d2_out) {
  __shared__ double tile[NN_TILE * 3];
  const int part = threadIdx.x & (NN_PARTS - 1);
  const int64_t q = (int64_t)blockIdx.x * NN_QUERIES + (threadIdx.x >> 2);
  const bool live = q < nq;
  const double qx = live ? query[q * 3 + 0] : 0.0, qy = live ? query[q * 3 + 1] : 0.0, qz = live ? query[q * 3 + 2] : 0.0;
  double best = __longlong_as_double(0x7ff0000000000000ll);      // +inf
  int32_t besti = 0x7fffffff;
  for (int64_t base = 0; base < nr; base += NN_TILE) {
    const int n = (int)min((int64_t)NN_TILE, nr - base);
    __syncthreads();
    for (int i = threadIdx.x; i < n * 3; i += NN_QUERIES * NN_PARTS) tile[i] = ref[base * 3 + i];
    __syncthreads();
    for (int j = part; j < n; j += NN_PARTS) {
      const double dx = __dsub_rn(qx, tile[3 * j]), dy = __dsub_rn(qy, tile[3 * j + 1]), dz = __dsub_rn(qz, tile[3 * j + 2]);
      // (x - y)^2 summed in coordinate order, like the trees' reduced distance
      const double d = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
      if (d < best) { best = d; besti = (int32_t)(base + j); }
    }
  }
#pragma unroll
  for (int o = 1; o < NN_PARTS; o <<= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int32_t oi = __shfl_xor_sync(0xffffffffu, besti, o);
    if (ob < best || (ob == best && oi < besti)) { best = ob; besti = oi; }
  }
  if (live && part == 0) {
    idx_out[q] = besti;
    if (d2_out) d2_out[q] = best;
  }
}

}  // namespace
}  // namespace asdf

extern "C" int asdf_nn_search(const double* query_dev, int64_t n_query, const double* ref_dev, int64_t n_ref,
                              int32_t* idx_dev, double* dist2_dev, void* stream) {
  using namespace asdf;
  ASDF_REQUIRE(n_query >= 0 && n_ref >= 1 && n_ref < ((int64_t)1 << 31), "asdf_nn_search: bad sizes");
  ASDF_REQUIRE(n_query == 0 || (query_dev && ref_dev && idx_dev), "asdf_nn_search: null argument");
  if (n_query == 0) return ASDF_OK;
  const unsigned blocks = (unsigned)((n_query + NN_QUERIES - 1) / NN_QUERIES);
  nn_search_kernel<<<blocks, NN_QUERIES * NN_PARTS, 0, (cudaStream_t)stream>>>(query_dev, n_query, ref_dev, n_ref,
                                                                              idx_dev, dist2_dev);
  ASDF_CUDA_CHECK(cudaGetLastError());
  return ASDF_OK;
}

// Marching cubes over a dense f32 field.
//
// Replaces skimage.measure.marching_cubes_lewiner at utils/mesh.py:354 and
// deep_sdf/mesh.py:81 (CPU, single-threaded Cython) plus the vertex shift of
// utils/mesh.py:360-363.  Topology rule, vertex placement and output ordering are specified
// in alignsdf_b200/mc_tables.py and restated independently in oracle/mc_oracle.py.
//
// Passes (all HBM-bound; the field is read twice):
//   mc_classify : 1 thread / grid point -> flags byte (3 edge-crossing bits, centre bit,
//                 4-bit triangle count), per-block sums, field min/max
//   mc_scan_blocks : exclusive scan of the block sums (single block)
//   mc_offsets  : per-point exclusive vertex / triangle offsets
//   mc_emit     : vertices (fp64 inverse-distance interpolation), keys, faces
#include "common.cuh"
#include "mc_tables.inc"

namespace asdf {
namespace {

constexpr int MC_BLOCK = 256;

struct McDims {
  int n0, n1, n2;
  int64_t n_pts;
  int n_blocks;
};

__host__ __device__ inline McDims mc_dims(const asdf_mc_params& p) {
  McDims d;
  d.n0 = p.n0; d.n1 = p.n1; d.n2 = p.n2;
  d.n_pts = (int64_t)p.n0 * p.n1 * p.n2;
  d.n_blocks = (int)((d.n_pts + MC_BLOCK - 1) / MC_BLOCK);
  return d;
}

// scratch layout: flags u8[n_pts] | pad | voff u32[n_pts] | toff u32[n_pts] | bsum u32[2][n_blocks]
struct McScratch {
  uint8_t* flags;
  uint32_t* voff;
  uint32_t* toff;
  uint32_t* bsum_v;
  uint32_t* bsum_t;
};

__host__ __device__ inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

__host__ __device__ inline McScratch mc_scratch(void* base, const McDims& d) {
  McScratch s;
  uint8_t* b = (uint8_t*)base;
  s.flags = b; b += align256((size_t)d.n_pts);
  s.voff = (uint32_t*)b; b += align256((size_t)d.n_pts * 4);
  s.toff = (uint32_t*)b; b += align256((size_t)d.n_pts * 4);
  s.bsum_v = (uint32_t*)b; b += align256((size_t)d.n_blocks * 4);
  s.bsum_t = (uint32_t*)b;
  return s;
}

__device__ __forceinline__ int ordered_int(float f) {
  const int b = __float_as_int(f);
  return b >= 0 ? b : b ^ 0x7fffffff;
}

// config + decider variant -> table entry
__device__ __forceinline__ int cell_entry(const float* f /* 8 values minus iso */) {
  int config = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c) config |= (f[c] < 0.f) << c;
  if (config == 0 || config == 255) return -1;
  const int amb = kMcAmbMask[config];
  int variant = 0, bit = 0;
  if (amb) {
#pragma unroll
    for (int face = 0; face < 6; ++face) {
      if (amb >> face & 1) {
        const unsigned char* q = kMcFaceCorners + face * 4;
        const float p02 = __fmul_rn(f[q[0]], f[q[2]]);
        const float p13 = __fmul_rn(f[q[1]], f[q[3]]);
        const bool joined = f[q[0]] < 0.f ? (p02 > p13) : (p13 > p02);
        variant |= (int)joined << bit;
        ++bit;
      }
    }
  }
  return kMcVarOffset[config] + variant;
}

__device__ __forceinline__ void load_cell(const float* __restrict__ vol, int64_t p, int n1, int n2,
                                          float iso, float* f) {
  const int64_t s0 = (int64_t)n1 * n2;
#pragma unroll
  for (int c = 0; c < 8; ++c)
    f[c] = __fsub_rn(__ldg(vol + p + ((c >> 2) & 1) * s0 + ((c >> 1) & 1) * n2 + (c & 1)), iso);
}

__global__ void __launch_bounds__(MC_BLOCK) mc_classify(const float* __restrict__ vol, asdf_mc_params prm,
                                                        McScratch s, int* minmax) {
  const McDims d = mc_dims(prm);
  const int64_t p = (int64_t)blockIdx.x * MC_BLOCK + threadIdx.x;
  int nv = 0, nt = 0;
  float v = 0.f;
  const bool live = p < d.n_pts;
  if (live) {
    const int k = (int)(p % d.n2), j = (int)((p / d.n2) % d.n1), i = (int)(p / ((int64_t)d.n1 * d.n2));
    v = __ldg(vol + p);
    const bool in0 = __fsub_rn(v, prm.iso) < 0.f;
    int flags = 0;
    const bool e0 = i + 1 < d.n0, e1 = j + 1 < d.n1, e2 = k + 1 < d.n2;
    if (e0) flags |= (in0 != (__fsub_rn(__ldg(vol + p + (int64_t)d.n1 * d.n2), prm.iso) < 0.f)) << 0;
    if (e1) flags |= (in0 != (__fsub_rn(__ldg(vol + p + d.n2), prm.iso) < 0.f)) << 1;
    if (e2) flags |= (in0 != (__fsub_rn(__ldg(vol + p + 1), prm.iso) < 0.f)) << 2;
    if (e0 && e1 && e2) {
      float f[8];
      load_cell(vol, p, d.n1, d.n2, prm.iso, f);
      const int e = cell_entry(f);
      if (e >= 0) {
        const int t = kMcNumTris[e];
        nt = t & 0x7f;
        flags |= (t >> 7) << 3;
      }
    }
    nv = __popc(flags & 0xf);
    s.flags[p] = (uint8_t)(flags | (nt << 4));
  }
  // block sums + field range
  __shared__ int sv[MC_BLOCK / 32], st[MC_BLOCK / 32], smin[MC_BLOCK / 32], smax[MC_BLOCK / 32];
  const unsigned full = 0xffffffffu;
  int wv = __reduce_add_sync(full, nv), wt = __reduce_add_sync(full, nt);
  int wmin = __reduce_min_sync(full, live ? ordered_int(v) : 0x7fffffff);
  int wmax = __reduce_max_sync(full, live ? ordered_int(v) : (int)0x80000000);
  if ((threadIdx.x & 31) == 0) {
    sv[threadIdx.x >> 5] = wv; st[threadIdx.x >> 5] = wt;
    smin[threadIdx.x >> 5] = wmin; smax[threadIdx.x >> 5] = wmax;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int a = 0, b = 0, mn = 0x7fffffff, mx = (int)0x80000000;
    for (int w = 0; w < MC_BLOCK / 32; ++w) { a += sv[w]; b += st[w]; mn = min(mn, smin[w]); mx = max(mx, smax[w]); }
    s.bsum_v[blockIdx.x] = a; s.bsum_t[blockIdx.x] = b;
    atomicMin(minmax, mn); atomicMax(minmax + 1, mx);
  }
}

// Single block: exclusive scan of both block-sum arrays in place; totals -> int64[2].
__global__ void __launch_bounds__(1024) mc_scan_blocks(McScratch s, int n_blocks, int64_t* totals) {
  __shared__ unsigned long long carry_v, carry_t;
  __shared__ unsigned wsum_v[32], wsum_t[32];
  if (threadIdx.x == 0) { carry_v = 0; carry_t = 0; }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < n_blocks; base += 1024) {
    const int idx = base + threadIdx.x;
    unsigned v = idx < n_blocks ? s.bsum_v[idx] : 0u, t = idx < n_blocks ? s.bsum_t[idx] : 0u;
    unsigned iv = v, it = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned a = __shfl_up_sync(0xffffffffu, iv, o), b = __shfl_up_sync(0xffffffffu, it, o);
      if (lane >= o) { iv += a; it += b; }
    }
    if (lane == 31) { wsum_v[warp] = iv; wsum_t[warp] = it; }
    __syncthreads();
    if (warp == 0) {
      unsigned a = wsum_v[lane], b = wsum_t[lane];
      unsigned ia = a, ib = b;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned x = __shfl_up_sync(0xffffffffu, ia, o), y = __shfl_up_sync(0xffffffffu, ib, o);
        if (lane >= o) { ia += x; ib += y; }
      }
      wsum_v[lane] = ia - a; wsum_t[lane] = ib - b;   // exclusive warp prefixes
    }
    __syncthreads();
    const unsigned long long ev = carry_v + wsum_v[warp] + (iv - v);
    const unsigned long long et = carry_t + wsum_t[warp] + (it - t);
    if (idx < n_blocks) { s.bsum_v[idx] = (unsigned)ev; s.bsum_t[idx] = (unsigned)et; }
    __syncthreads();
    if (threadIdx.x == 1023) { carry_v = ev + v; carry_t = et + t; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { totals[0] = (int64_t)carry_v; totals[1] = (int64_t)carry_t; }
}

__global__ void __launch_bounds__(MC_BLOCK) mc_offsets(asdf_mc_params prm, McScratch s) {
  const McDims d = mc_dims(prm);
  const int64_t p = (int64_t)blockIdx.x * MC_BLOCK + threadIdx.x;
  const int fl = p < d.n_pts ? s.flags[p] : 0;
  const unsigned nv = __popc(fl & 0xf), nt = fl >> 4;
  unsigned iv = nv, it = nt;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned a = __shfl_up_sync(0xffffffffu, iv, o), b = __shfl_up_sync(0xffffffffu, it, o);
    if (lane >= o) { iv += a; it += b; }
  }
  __shared__ unsigned wv[MC_BLOCK / 32], wt[MC_BLOCK / 32];
  if (lane == 31) { wv[warp] = iv; wt[warp] = it; }
  __syncthreads();
  unsigned pv = 0, pt = 0;
  for (int w = 0; w < warp; ++w) { pv += wv[w]; pt += wt[w]; }
  if (p < d.n_pts) {
    s.voff[p] = s.bsum_v[blockIdx.x] + pv + iv - nv;
    s.toff[p] = s.bsum_t[blockIdx.x] + pt + it - nt;
  }
}

// interpolation parameter along an edge whose end values (minus iso) are f0 (lower point), f1
__device__ __forceinline__ double edge_t(float f0, float f1) {
  const double eps = 1.1920928955078125e-07;   // FLT_EPSILON
  const double w0 = __ddiv_rn(1.0, __dadd_rn(eps, fabs((double)f0)));
  const double w1 = __ddiv_rn(1.0, __dadd_rn(eps, fabs((double)f1)));
  return __ddiv_rn(w1, __dadd_rn(w0, w1));
}

__global__ void __launch_bounds__(MC_BLOCK) mc_emit(const float* __restrict__ vol, asdf_mc_params prm,
                                                    McScratch s, float* __restrict__ verts,
                                                    float* __restrict__ points, int32_t* __restrict__ faces,
                                                    unsigned long long* __restrict__ keys) {
  const McDims d = mc_dims(prm);
  const int64_t p = (int64_t)blockIdx.x * MC_BLOCK + threadIdx.x;
  if (p >= d.n_pts) return;
  const int fl = s.flags[p];
  if (fl == 0) return;
  const int k = (int)(p % d.n2), j = (int)((p / d.n2) % d.n1), i = (int)(p / ((int64_t)d.n1 * d.n2));
  const int64_t s0 = (int64_t)d.n1 * d.n2;
  const int64_t stride[3] = {s0, (int64_t)d.n2, 1};
  const double gidx[3] = {(double)(i + prm.index0_offset), (double)j, (double)k};
  const unsigned long long gkey = (unsigned long long)(p + prm.index0_offset * s0) * 4ull;
  unsigned vo = s.voff[p];

  auto put_vertex = [&](unsigned slot, double q0, double q1, double q2, unsigned long long key) {
    const float x0 = (float)__dmul_rn(q0, prm.spacing[0]);
    const float x1 = (float)__dmul_rn(q1, prm.spacing[1]);
    const float x2 = (float)__dmul_rn(q2, prm.spacing[2]);
    verts[(size_t)slot * 3 + 0] = x0; verts[(size_t)slot * 3 + 1] = x1; verts[(size_t)slot * 3 + 2] = x2;
    if (points) {
      points[(size_t)slot * 3 + 0] = __fadd_rn(prm.origin[0], x0);
      points[(size_t)slot * 3 + 1] = __fadd_rn(prm.origin[1], x1);
      points[(size_t)slot * 3 + 2] = __fadd_rn(prm.origin[2], x2);
    }
    if (keys) keys[slot] = key;
  };

  // ---- vertices on the three edges owned by this grid point
  if (fl & 7) {
    const float f0 = __fsub_rn(__ldg(vol + p), prm.iso);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      if (fl >> a & 1) {
        const float f1 = __fsub_rn(__ldg(vol + p + stride[a]), prm.iso);
        double q[3] = {gidx[0], gidx[1], gidx[2]};
        q[a] = __dadd_rn(q[a], edge_t(f0, f1));
        put_vertex(vo++, q[0], q[1], q[2], gkey + a);
      }
    }
  }
  const int nt = fl >> 4;
  if (nt == 0) return;

  // ---- this point's cell
  float f[8];
  load_cell(vol, p, d.n1, d.n2, prm.iso, f);
  const int e = cell_entry(f);
  const unsigned char* tri = kMcTriEdges + 3 * (int)kMcTriStart[e];

  // global vertex index of local vertex id (edge 0..11 or centre 12)
  auto vindex = [&](int id) -> int32_t {
    if (id == 12) return (int32_t)(s.voff[p] + __popc(fl & 7));
    const int c0 = kMcEdgeCorner[id * 2];
    const int a = id >> 2;
    const int64_t owner = p + ((c0 >> 2) & 1) * s0 + ((c0 >> 1) & 1) * d.n2 + (c0 & 1);
    const int ofl = s.flags[owner];
    return (int32_t)(s.voff[owner] + __popc(ofl & ((1 << a) - 1)));
  };

  if (fl & 8) {   // centre vertex: mean (fp64, loop order) of the loop's vertices in index space
    double acc[3] = {0.0, 0.0, 0.0};
    int cnt = 0;
    for (int t = 0; t < nt; ++t) {
      if (tri[3 * t] != 12) continue;
      const int id = tri[3 * t + 1];
      const int c0 = kMcEdgeCorner[id * 2], c1 = kMcEdgeCorner[id * 2 + 1];
      const int a = id >> 2;
      double q[3] = {gidx[0] + ((c0 >> 2) & 1), gidx[1] + ((c0 >> 1) & 1), gidx[2] + (c0 & 1)};
      q[a] = __dadd_rn(q[a], edge_t(f[c0], f[c1]));
      acc[0] = __dadd_rn(acc[0], q[0]); acc[1] = __dadd_rn(acc[1], q[1]); acc[2] = __dadd_rn(acc[2], q[2]);
      ++cnt;
    }
    const double n = (double)cnt;
    put_vertex(s.voff[p] + __popc(fl & 7), __ddiv_rn(acc[0], n), __ddiv_rn(acc[1], n), __ddiv_rn(acc[2], n), gkey + 3);
  }
  const unsigned to = s.toff[p];
  for (int t = 0; t < nt; ++t) {
    faces[(size_t)(to + t) * 3 + 0] = vindex(tri[3 * t + 0]);
    faces[(size_t)(to + t) * 3 + 1] = vindex(tri[3 * t + 1]);
    faces[(size_t)(to + t) * 3 + 2] = vindex(tri[3 * t + 2]);
  }
}

__global__ void mc_init_totals(int64_t* totals, int* minmax) {
  totals[0] = totals[1] = 0;
  minmax[0] = 0x7fffffff; minmax[1] = (int)0x80000000;
}

__global__ void mc_finish_totals(int64_t* totals, const int* minmax) {
  totals[2] = minmax[0]; totals[3] = minmax[1];
}

int check_params(const asdf_mc_params* p) {
  ASDF_REQUIRE(p, "null mc params");
  ASDF_REQUIRE(p->n0 >= 2 && p->n1 >= 2 && p->n2 >= 2, "marching cubes needs at least 2 samples per axis");
  ASDF_REQUIRE((int64_t)p->n0 * p->n1 * p->n2 < ((int64_t)1 << 31) * 2 - 1024, "volume too large for 32-bit offsets");
  ASDF_REQUIRE(p->full1 == p->n1 && p->full2 == p->n2, "slabs must span axes 1 and 2 completely");
  return ASDF_OK;
}

}  // namespace
}  // namespace asdf

extern "C" size_t asdf_mc_scratch_bytes(const asdf_mc_params* p) {
  using namespace asdf;
  if (!p) return 0;
  const McDims d = mc_dims(*p);
  return align256((size_t)d.n_pts) + 2 * align256((size_t)d.n_pts * 4) + 2 * align256((size_t)d.n_blocks * 4) + 256;
}

extern "C" int asdf_mc_count(const float* vol_dev, const asdf_mc_params* p, void* scratch_dev,
                             int64_t* totals_dev, void* stream) {
  using namespace asdf;
  if (int rc = check_params(p)) return rc;
  ASDF_REQUIRE(vol_dev && scratch_dev && totals_dev, "asdf_mc_count: null argument");
  const McDims d = mc_dims(*p);
  McScratch s = mc_scratch(scratch_dev, d);
  cudaStream_t st = (cudaStream_t)stream;
  int* minmax = (int*)((uint8_t*)s.bsum_t + align256((size_t)d.n_blocks * 4));
  mc_init_totals<<<1, 1, 0, st>>>(totals_dev, minmax);
  mc_classify<<<d.n_blocks, MC_BLOCK, 0, st>>>(vol_dev, *p, s, minmax);
  mc_scan_blocks<<<1, 1024, 0, st>>>(s, d.n_blocks, totals_dev);
  mc_offsets<<<d.n_blocks, MC_BLOCK, 0, st>>>(*p, s);
  mc_finish_totals<<<1, 1, 0, st>>>(totals_dev, minmax);
  ASDF_CUDA_CHECK(cudaGetLastError());
  return ASDF_OK;
}

extern "C" int asdf_mc_emit(const float* vol_dev, const asdf_mc_params* p, const void* scratch_dev,
                            float* verts_dev, float* points_dev, int32_t* faces_dev,
                            uint64_t* keys_dev, void* stream) {
  using namespace asdf;
  if (int rc = check_params(p)) return rc;
  ASDF_REQUIRE(vol_dev && scratch_dev && verts_dev && faces_dev, "asdf_mc_emit: null argument");
  const McDims d = mc_dims(*p);
  McScratch s = mc_scratch(const_cast<void*>(scratch_dev), d);
  mc_emit<<<d.n_blocks, MC_BLOCK, 0, (cudaStream_t)stream>>>(vol_dev, *p, s, verts_dev, points_dev, faces_dev,
                                                             (unsigned long long*)keys_dev);
  ASDF_CUDA_CHECK(cudaGetLastError());
  return ASDF_OK;
}

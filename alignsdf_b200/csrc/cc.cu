// K3: connected components of a marching-cubes mesh + per-component statistics on the GPU.
//
// Replaces the CPU stage trimesh.graph.split + "keep the largest-area piece iff the mesh splits"
// of utils/mesh.py:371-381.  Semantics are those of
// alignsdf_b200/trimesh_lite.py::largest_watertight_component_mc, which tests check against the
// generic edge-adjacency split: a marching-cubes mesh is a closed manifold except where it leaves
// the volume, so a component is watertight iff none of its triangle edges lies in a boundary plane
// of the volume.
//
//   cc_init    parent[v] = v
//   cc_union   lock-free union-find over the face edges (a,b), (b,c): roots hook onto the SMALLER
//              index with atomicMin, so the final label of a component is its smallest vertex id
//              (deterministic whatever the thread order)
//   cc_flatten parent[v] = root(v)
//   cc_stats   per face: component = parent[a]; area (fp64), face count, "open" flag (an edge in a
//              boundary plane), first face index -> per-root accumulators (atomics)
//   cc_mark / cc_gather  compaction of the selected component (prefix sums by the caller)
// All HBM/L2-latency bound and tiny (V ~ 2e5, F ~ 4e5 at 256^3: a few tens of microseconds).
#include "common.cuh"

namespace asdf {
namespace {

constexpr int CC_BLOCK = 256;

__device__ __forceinline__ int cc_find(const int32_t* parent, int v) {
  int p = parent[v];
  while (p != v) { v = p; p = parent[v]; }
  return v;
}

__global__ void cc_init(int32_t* parent, int64_t V) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v < V) parent[v] = (int32_t)v;
}

__device__ __forceinline__ void cc_unite(int32_t* parent, int a, int b) {
  int ra = cc_find(parent, a), rb = cc_find(parent, b);
  while (ra != rb) {
    if (ra < rb) { const int t = ra; ra = rb; rb = t; }          // hook the larger root onto the smaller
    const int old = atomicMin(parent + ra, rb);
    if (old == ra) break;                                          // ra was still a root: done
    ra = cc_find(parent, old);                                     // somebody re-parented ra meanwhile
    rb = cc_find(parent, rb);
  }
}

__global__ void cc_union(const int32_t* __restrict__ faces, int64_t F, int32_t* parent) {
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  const int a = faces[3 * f], b = faces[3 * f + 1], c = faces[3 * f + 2];
  cc_unite(parent, a, b);
  cc_unite(parent, b, c);
}

__global__ void cc_flatten(int32_t* parent, int64_t V) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v < V) parent[v] = cc_find(parent, (int)v);
}

// bit 2k: on plane 0 of axis k, bit 2k+1: on the last plane of axis k (exact float compares: lattice
// planes are hit exactly by the emit kernel's vertices)
__device__ __forceinline__ unsigned plane_bits(const float* __restrict__ vl, int v, const float* last) {
  unsigned m = 0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float x = vl[3 * v + k];
    m |= (x == 0.0f ? 1u : 0u) << (2 * k);
    m |= (x == last[k] ? 1u : 0u) << (2 * k + 1);
  }
  return m;
}

struct CcStatsArgs {
  const int32_t* faces;
  int64_t F;
  const float* verts_local;     // [V,3] raw marching-cubes vertices (array-axis order x spacing)
  const float* points;          // [V,3] final mesh points (area is measured on these)
  const int32_t* parent;        // flattened labels
  float last[3];                // (dims[k] - 1) * spacing[k] as the emit kernel rounds it
  double* area;                 // [V] per-root accumulators, zero-initialised by the caller
  int32_t* nfaces;              // [V]
  int32_t* open;                // [V]
  int32_t* first_face;          // [V], initialised to INT_MAX
};

__global__ void cc_stats(const CcStatsArgs a) {
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= a.F) return;
  const int ia = a.faces[3 * f], ib = a.faces[3 * f + 1], ic = a.faces[3 * f + 2];
  const int root = a.parent[ia];
  const unsigned fa = plane_bits(a.verts_local, ia, a.last), fb = plane_bits(a.verts_local, ib, a.last),
                 fc = plane_bits(a.verts_local, ic, a.last);
  const double ax = a.points[3 * ia], ay = a.points[3 * ia + 1], az = a.points[3 * ia + 2];
  const double e1x = a.points[3 * ib] - ax, e1y = a.points[3 * ib + 1] - ay, e1z = a.points[3 * ib + 2] - az;
  const double e2x = a.points[3 * ic] - ax, e2y = a.points[3 * ic + 1] - ay, e2z = a.points[3 * ic + 2] - az;
  const double cx = e1y * e2z - e1z * e2y, cy = e1z * e2x - e1x * e2z, cz = e1x * e2y - e1y * e2x;
  double area = 0.5 * sqrt(cx * cx + cy * cy + cz * cz);
  const bool open = ((fa & fb) | (fb & fc) | (fc & fa)) != 0u;
  // Almost every warp sees a single component: aggregate in the warp and issue one set of atomics (thousands of
  // faces hammering the same four addresses serialise in L2 otherwise).
  const unsigned active = __activemask();
  const unsigned same = __match_any_sync(active, root);
  if (same == active) {
    const int lane = threadIdx.x & 31, leader = __ffs(active) - 1;
    int cnt = __popc(active);
    const int any_open = __any_sync(active, open);
    const int first = __reduce_min_sync(active, (int)f);
    for (int off = 16; off > 0; off >>= 1) {
      const double o = __shfl_down_sync(active, area, off);
      if (lane + off < 32 && (active >> (lane + off) & 1u)) area += o;
    }
    if (lane == leader) {
      atomicAdd(a.area + root, area);
      atomicAdd(a.nfaces + root, cnt);
      atomicMin(a.first_face + root, first);
      if (any_open) atomicOr(a.open + root, 1);
    }
  } else {
    if (open) atomicOr(a.open + root, 1);
    atomicAdd(a.area + root, area);
    atomicAdd(a.nfaces + root, 1);
    atomicMin(a.first_face + root, (int)f);
  }
}

// keep_v[v] = parent[v] == best;  keep_f[f] = parent[faces[f][0]] == best   (int32 0/1 for the caller's scan)
__global__ void cc_mark(const int32_t* __restrict__ parent, int64_t V, const int32_t* __restrict__ faces, int64_t F,
                        int best, int32_t* keep_v, int32_t* keep_f) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < V) keep_v[i] = parent[i] == best ? 1 : 0;
  if (i < F) keep_f[i] = parent[faces[3 * i]] == best ? 1 : 0;
}

// scan_v / scan_f: INCLUSIVE prefix sums of keep_v / keep_f
__global__ void cc_gather(const float* __restrict__ points, const int32_t* __restrict__ faces, int64_t V, int64_t F,
                          const int32_t* __restrict__ keep_v, const int32_t* __restrict__ scan_v,
                          const int32_t* __restrict__ keep_f, const int32_t* __restrict__ scan_f,
                          float* out_points, int32_t* out_faces) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < V && keep_v[i]) {
    const int64_t o = scan_v[i] - 1;
    out_points[3 * o] = points[3 * i]; out_points[3 * o + 1] = points[3 * i + 1]; out_points[3 * o + 2] = points[3 * i + 2];
  }
  if (i < F && keep_f[i]) {
    const int64_t o = scan_f[i] - 1;
#pragma unroll
    for (int k = 0; k < 3; ++k) out_faces[3 * o + k] = scan_v[faces[3 * i + k]] - 1;
  }
}

inline unsigned blocks_for(int64_t n) { return (unsigned)((n + CC_BLOCK - 1) / CC_BLOCK); }

}  // namespace
}  // namespace asdf

extern "C" int asdf_cc_label(const int32_t* faces_dev, int64_t F, int64_t V, int32_t* parent_dev, void* stream) {
  using namespace asdf;
  ASDF_REQUIRE(F >= 0 && V >= 0 && V < (int64_t)0x7fffffff, "asdf_cc_label: bad sizes");
  if (V == 0) return ASDF_OK;
  ASDF_REQUIRE(parent_dev && (F == 0 || faces_dev), "asdf_cc_label: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  cc_init<<<blocks_for(V), CC_BLOCK, 0, st>>>(parent_dev, V);
  if (F > 0) cc_union<<<blocks_for(F), CC_BLOCK, 0, st>>>(faces_dev, F, parent_dev);
  cc_flatten<<<blocks_for(V), CC_BLOCK, 0, st>>>(parent_dev, V);
  ASDF_CUDA_CHECK(cudaGetLastError());
  return ASDF_OK;
}

extern "C" int asdf_cc_stats(const int32_t* faces_dev, int64_t F, const float* verts_local_dev,
                             const float* points_dev, int64_t V, const int32_t* parent_dev,
                             const float last_plane[3], double* area_dev, int32_t* nfaces_dev,
                             int32_t* open_dev, int32_t* first_face_dev, void* stream) {
  using namespace asdf;
  ASDF_REQUIRE(F >= 0 && V >= 0, "asdf_cc_stats: bad sizes");
  if (F == 0) return ASDF_OK;
  ASDF_REQUIRE(faces_dev && verts_local_dev && points_dev && parent_dev && last_plane && area_dev && nfaces_dev &&
               open_dev && first_face_dev, "asdf_cc_stats: null argument");
  CcStatsArgs a;
  a.faces = faces_dev; a.F = F; a.verts_local = verts_local_dev; a.points = points_dev; a.parent = parent_dev;
  for (int k = 0; k < 3; ++k) a.last[k] = last_plane[k];
  a.area = area_dev; a.nfaces = nfaces_dev; a.open = open_dev; a.first_face = first_face_dev;
  cc_stats<<<blocks_for(F), CC_BLOCK, 0, (cudaStream_t)stream>>>(a);
  ASDF_CUDA_CHECK(cudaGetLastError());
  return ASDF_OK;
}

extern "C" int asdf_cc_mark(const int32_t* parent_dev, int64_t V, const int32_t* faces_dev, int64_t F,
                            int32_t best_label, int32_t* keep_v_dev, int32_t* keep_f_dev, void* stream) {
  using namespace asdf;
  ASDF_REQUIRE(parent_dev && faces_dev && keep_v_dev && keep_f_dev && V > 0 && F > 0, "asdf_cc_mark: bad argument");
  cc_mark<<<blocks_for(V > F ? V : F), CC_BLOCK, 0, (cudaStream_t)stream>>>(parent_dev, V, faces_dev, F, best_label,
                                                                            keep_v_dev, keep_f_dev);
  ASDF_CUDA_CHECK(cudaGetLastError());
  return ASDF_OK;
}

extern "C" int asdf_cc_gather(const float* points_dev, const int32_t* faces_dev, int64_t V, int64_t F,
                              const int32_t* keep_v_dev, const int32_t* scan_v_dev, const int32_t* keep_f_dev,
                              const int32_t* scan_f_dev, float* out_points_dev, int32_t* out_faces_dev, void* stream) {
  using namespace asdf;
  ASDF_REQUIRE(points_dev && faces_dev && keep_v_dev && scan_v_dev && keep_f_dev && scan_f_dev && out_points_dev &&
               out_faces_dev && V > 0 && F > 0, "asdf_cc_gather: bad argument");
  cc_gather<<<blocks_for(V > F ? V : F), CC_BLOCK, 0, (cudaStream_t)stream>>>(points_dev, faces_dev, V, F, keep_v_dev,
                                                                              scan_v_dev, keep_f_dev, scan_f_dev,
                                                                              out_points_dev, out_faces_dev);
  ASDF_CUDA_CHECK(cudaGetLastError());
  return ASDF_OK;
}

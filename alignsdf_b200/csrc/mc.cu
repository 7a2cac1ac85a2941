// Marching cubes over a dense f32 field -- HBM bound: the field is read ONCE.
//
// Replaces skimage.measure.marching_cubes_lewiner at utils/mesh.py:354 and
// deep_sdf/mesh.py:81 (CPU, single-threaded Cython) plus the vertex shift of
// utils/mesh.py:360-363.  Topology rule, vertex placement and output ordering are specified
// in alignsdf_b200/mc_tables.py and restated independently in oracle/mc_oracle.py.
//
// Unit of bookkeeping: a SEGMENT = 32 consecutive points along axis 2 at fixed (axis 0, axis 1) = one warp.
// Output order is grid order, i.e. (segment, lane) order.
//
//   mc_classify   tiles of 32 x 8 points marching 8 planes along axis 0 with the previous plane in registers
//                 (every field value is fetched from DRAM once; 2 loads per point).  Per point a flags byte
//                 (3 edge-crossing bits, centre bit, 4-bit triangle count); per segment one word
//                 (vertices | triangles << 16); the per-point words (flags | vertex prefix inside the segment << 8)
//                 are stored only for segments that own something.
//                 Field min / max for the "level outside the data range" check.
//   mc_scan_*     exclusive scan of the segment words (4096 segments per block, then the block sums)
//   mc_emit       one warp per batch of 32 segment words; only non-empty segments are touched again: flags of the
//                 segment and of its three (axis 0 / axis 1) neighbours -> vertex indices of all 12 cell edges by
//                 warp shuffles, vertices (fp64 inverse-distance interpolation), keys, faces.
// Traffic ~ 4 N^3 (field) + N^3 / 8 (segment words) + N^3 / 4 (offsets) + a few % for the surface itself.
#include "common.cuh"
#include "mc_tables.inc"

namespace asdf {
namespace {

constexpr int MC_ZC = 16;             // planes per classify block (tile: 32 x 8 x 16 points + halo in shared memory)
constexpr int MC_SCAN_ITEMS = 4;      // segments per thread of the scan
constexpr int MC_SCAN_BLOCK = 1024 * MC_SCAN_ITEMS;

struct McDims {
  int n0, n1, n2;
  int nsx;                            // segments per row
  int64_t n_pts;
  int64_t n_seg;
  int n_scan_blocks;
};

__host__ __device__ inline McDims mc_dims(const asdf_mc_params& p) {
  McDims d;
  d.n0 = p.n0; d.n1 = p.n1; d.n2 = p.n2;
  d.nsx = (p.n2 + 31) / 32;
  d.n_pts = (int64_t)p.n0 * p.n1 * p.n2;
  d.n_seg = (int64_t)p.n0 * p.n1 * d.nsx;
  d.n_scan_blocks = (int)((d.n_seg + MC_SCAN_BLOCK - 1) / MC_SCAN_BLOCK);
  return d;
}

// scratch layout: segw u32[n_seg] | info u16[n_seg][32] | offv u32[n_seg] | offt u32[n_seg] | list u32[n_seg] |
//                 bsum u32[3][n_scan_blocks] | minmax
struct McScratch {
  uint32_t* segw;      // vertices | triangles << 16 of the segment
  uint16_t* info;      // [segment][lane]: flags | (vertices owned by the earlier lanes of the segment) << 8; valid only where segw != 0
  uint32_t* offv;      // exclusive offsets inside the segment's scan block
  uint32_t* offt;
  uint32_t* list;      // the non-empty segments, in order
  uint32_t* bsum_v;    // per scan block: totals, then (after mc_scan_blocks) exclusive bases
  uint32_t* bsum_t;
  uint32_t* bsum_c;    // same for the number of non-empty segments
  int* minmax;
};

__host__ __device__ inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

__host__ __device__ inline McScratch mc_scratch(void* base, const McDims& d) {
  McScratch s;
  uint8_t* b = (uint8_t*)base;
  s.segw = (uint32_t*)b; b += align256((size_t)d.n_seg * 4);
  s.info = (uint16_t*)b; b += align256((size_t)d.n_seg * 64);
  s.offv = (uint32_t*)b; b += align256((size_t)d.n_seg * 4);
  s.offt = (uint32_t*)b; b += align256((size_t)d.n_seg * 4);
  s.list = (uint32_t*)b; b += align256((size_t)d.n_seg * 4);
  s.bsum_v = (uint32_t*)b; b += align256((size_t)d.n_scan_blocks * 4);
  s.bsum_t = (uint32_t*)b; b += align256((size_t)d.n_scan_blocks * 4);
  s.bsum_c = (uint32_t*)b; b += align256((size_t)d.n_scan_blocks * 4);
  s.minmax = (int*)b;
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

// flags byte of a point from the 8 corner values (minus iso) of its cell: f[c], c = 4 d0 + 2 d1 + d2
__device__ __forceinline__ int point_flags(const float* f, bool e0, bool e1, bool e2) {
  const bool in0 = f[0] < 0.f;
  int flags = 0;
  if (e0) flags |= (in0 != (f[4] < 0.f)) << 0;
  if (e1) flags |= (in0 != (f[2] < 0.f)) << 1;
  if (e2) flags |= (in0 != (f[1] < 0.f)) << 2;
  if (e0 && e1 && e2) {
    const int e = cell_entry(f);
    if (e >= 0) {
      const int t = kMcNumTris[e];
      flags |= ((t >> 7) << 3) | ((t & 0x7f) << 4);
    }
  }
  return flags;
}

__global__ void __launch_bounds__(256) mc_classify(const float* __restrict__ vol, asdf_mc_params prm, McScratch s) {
  const McDims d = mc_dims(prm);
  // tile of 32 x 8 x MC_ZC points + one halo layer on the upper side of every axis, staged in shared memory with
  // all loads of the block in flight at once (the kernel is a pure stream over the field: latency must be hidden
  // by memory-level parallelism, not by arithmetic).  While staging, every row of 32 points leaves its inside /
  // outside bits (one ballot): a segment whose own row and three neighbour rows agree in all bits owns nothing,
  // which is the case for ~98 % of them, and is dismissed with a handful of warp-uniform instructions.
  __shared__ float tile[MC_ZC + 1][9][33];
  __shared__ uint32_t rowbits[MC_ZC + 1][9];                      // bit x: value < iso
  __shared__ uint8_t halobit[MC_ZC + 1][9];                       // the same for the halo column x0 + 32
  const int lane = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + lane;
  const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 8, z0 = blockIdx.z * MC_ZC;
  const int nx = min(33, d.n2 - x0), ny = min(9, d.n1 - y0), nz = min(MC_ZC + 1, d.n0 - z0);
  const int64_t plane = (int64_t)d.n1 * d.n2;
  const unsigned full = 0xffffffffu;
  const float iso = prm.iso;
  int mn = 0x7fffffff, mx = (int)0x80000000;
  const float* base = vol + (int64_t)z0 * plane + (int64_t)y0 * d.n2 + x0;
  if (nx == 33 && ny == 9 && nz == MC_ZC + 1) {
    // interior tile: compile-time trip counts, every load of the thread independent of the others
    float vmin = __int_as_float(0x7f800000), vmax = __int_as_float(0xff800000);
    auto stage_row = [&](int py) {
      const float* colp = base + (int64_t)py * d.n2 + lane;
      float v[MC_ZC + 1];
#pragma unroll
      for (int pz = 0; pz <= MC_ZC; ++pz) v[pz] = __ldg(colp + (int64_t)pz * plane);
#pragma unroll
      for (int pz = 0; pz <= MC_ZC; ++pz) {
        tile[pz][py][lane] = v[pz];
        vmin = fminf(vmin, v[pz]); vmax = fmaxf(vmax, v[pz]);
        const uint32_t bits = __ballot_sync(full, v[pz] < iso);
        if (lane == 0) rowbits[pz][py] = bits;
      }
    };
    stage_row(ty);
    if (ty == 0) stage_row(8);
    if (tid >= 64 && tid < 64 + (MC_ZC + 1) * 9) {                 // the halo column x0 + 32 (warps 2..6)
      const int r = tid - 64, pz = r / 9, py = r - pz * 9;
      const float v = __ldg(base + (int64_t)pz * plane + (int64_t)py * d.n2 + 32);
      tile[pz][py][32] = v;
      halobit[pz][py] = v < iso;
    }
    mn = ordered_int(vmin); mx = ordered_int(vmax);
  } else {
    const int rows = nz * ny;
    // 32-wide part: one row per warp and step
    for (int r = ty; r < rows; r += 8) {
      const int pz = r / ny, py = r - pz * ny;
      float v = 0.f;
      const bool ok = lane < nx;
      if (ok) {
        v = __ldg(base + (int64_t)pz * plane + (int64_t)py * d.n2 + lane);
        tile[pz][py][lane] = v;
        const int o = ordered_int(v);
        mn = min(mn, o); mx = max(mx, o);
      }
      const uint32_t bits = __ballot_sync(full, ok && v < iso);
      if (lane == 0) rowbits[pz][py] = bits;
    }
    // the halo column x0 + 32
    if (nx == 33)
      for (int r = tid; r < rows; r += 256) {
        const int pz = r / ny, py = r - pz * ny;
        const float v = __ldg(base + (int64_t)pz * plane + (int64_t)py * d.n2 + 32);
        tile[pz][py][32] = v;
        halobit[pz][py] = v < iso;
      }
  }
  __syncthreads();
  const int x = x0 + lane, y = y0 + ty;
  const bool vx = x < d.n2, vy = y < d.n1;
  const bool e1 = y + 1 < d.n1, e2 = x + 1 < d.n2;
  if (vy) {                                                      // warp-uniform
    const int zc = min(MC_ZC, d.n0 - z0);
    const uint32_t valid = nx >= 32 ? 0xffffffffu : ((1u << nx) - 1u);
    const bool has_halo = nx == 33;
    const int ty1 = e1 ? ty + 1 : ty;
    for (int pz = 0; pz < zc; ++pz) {
      const bool e0 = z0 + pz + 1 < d.n0;
      const int pz1 = e0 ? pz + 1 : pz;
      const int64_t seg = ((int64_t)(z0 + pz) * d.n1 + y) * d.nsx + blockIdx.x;
      {   // all inside or all outside over the segment's own row and its neighbour rows (and the halo column)?
        const uint32_t a = rowbits[pz][ty], b = rowbits[pz][ty1], c = rowbits[pz1][ty], e = rowbits[pz1][ty1];
        uint32_t any = (a | b | c | e) & valid, all = (a & b & c & e) & valid;
        if (has_halo) {
          const uint32_t ha = halobit[pz][ty], hb = halobit[pz][ty1], hc = halobit[pz1][ty], he = halobit[pz1][ty1];
          const uint32_t hany = ha | hb | hc | he, hall = ha & hb & hc & he;
          if (any == 0u && hany == 0u) { if (lane == 0) s.segw[seg] = 0u; continue; }
          if (all == valid && hall != 0u) { if (lane == 0) s.segw[seg] = 0u; continue; }
        } else {
          if (any == 0u || all == valid) { if (lane == 0) s.segw[seg] = 0u; continue; }
        }
      }
      int fl = 0;
      if (vx) {
        float f[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const bool ok = (!(c & 4) || e0) && (!(c & 2) || e1) && (!(c & 1) || e2);
          f[c] = ok ? __fsub_rn(tile[pz + ((c >> 2) & 1)][ty + ((c >> 1) & 1)][lane + (c & 1)], iso) : 0.f;
        }
        fl = point_flags(f, e0, e1, e2);
      }
      const int wv = __reduce_add_sync(full, __popc(fl & 0xf)), wt = __reduce_add_sync(full, fl >> 4);
      if (lane == 0) s.segw[seg] = (uint32_t)wv | ((uint32_t)wt << 16);
      if (wv | wt) {
        const int inc = __popc(fl & 0xf);
        int sc = inc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int a = __shfl_up_sync(full, sc, o); if (lane >= o) sc += a; }
        s.info[seg * 32 + lane] = (uint16_t)(fl | ((sc - inc) << 8));
      }
    }
  }
  // field range: one atomic pair per block, and only if it can still change the result
  __shared__ int smin[8], smax[8];
  mn = __reduce_min_sync(full, mn); mx = __reduce_max_sync(full, mx);
  if (lane == 0) { smin[ty] = mn; smax[ty] = mx; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 8; ++w) { mn = min(mn, smin[w]); mx = max(mx, smax[w]); }
    if (mn < *(volatile int*)s.minmax) atomicMin(s.minmax, mn);
    if (mx > *(volatile int*)(s.minmax + 1)) atomicMax(s.minmax + 1, mx);
  }
}

// exclusive scan of the segment words inside blocks of MC_SCAN_BLOCK segments; block totals -> bsum
__global__ void __launch_bounds__(1024) mc_scan_local(McScratch s, int64_t n_seg) {
  const int64_t base = (int64_t)blockIdx.x * MC_SCAN_BLOCK + (int64_t)threadIdx.x * MC_SCAN_ITEMS;
  uint32_t v[MC_SCAN_ITEMS], t[MC_SCAN_ITEMS];
  uint32_t sv = 0, st = 0, sc = 0;
  if (base + MC_SCAN_ITEMS <= n_seg) {
    const uint4 w = *reinterpret_cast<const uint4*>(s.segw + base);
    const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int i = 0; i < MC_SCAN_ITEMS; ++i) { v[i] = ws[i] & 0xffffu; t[i] = ws[i] >> 16; sc += ws[i] != 0u; }
  } else {
#pragma unroll
    for (int i = 0; i < MC_SCAN_ITEMS; ++i) {
      const uint32_t w = base + i < n_seg ? s.segw[base + i] : 0u;
      v[i] = w & 0xffffu; t[i] = w >> 16; sc += w != 0u;
    }
  }
#pragma unroll
  for (int i = 0; i < MC_SCAN_ITEMS; ++i) { sv += v[i]; st += t[i]; }
  {   // block total of the non-empty count (its prefix is only needed by mc_compact, which recomputes it)
    const uint32_t wc = __reduce_add_sync(0xffffffffu, sc);
    __shared__ uint32_t tot_c;
    if (threadIdx.x == 0) tot_c = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && wc) atomicAdd(&tot_c, wc);
    __syncthreads();
    if (threadIdx.x == 0) s.bsum_c[blockIdx.x] = tot_c;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t iv = sv, it = st;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t a = __shfl_up_sync(0xffffffffu, iv, o), b = __shfl_up_sync(0xffffffffu, it, o);
    if (lane >= o) { iv += a; it += b; }
  }
  __shared__ uint32_t wsv[32], wst[32];
  if (lane == 31) { wsv[warp] = iv; wst[warp] = it; }
  __syncthreads();
  if (warp == 0) {
    const uint32_t a = wsv[lane], b = wst[lane];
    uint32_t ia = a, ib = b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t xx = __shfl_up_sync(0xffffffffu, ia, o), yy = __shfl_up_sync(0xffffffffu, ib, o);
      if (lane >= o) { ia += xx; ib += yy; }
    }
    wsv[lane] = ia - a; wst[lane] = ib - b;
    if (lane == 31) { s.bsum_v[blockIdx.x] = ia; s.bsum_t[blockIdx.x] = ib; }
  }
  __syncthreads();
  uint32_t ev = wsv[warp] + iv - sv, et = wst[warp] + it - st;
  if (base + MC_SCAN_ITEMS <= n_seg) {
    uint4 ov, ot;
    ov.x = ev; ov.y = ev + v[0]; ov.z = ov.y + v[1]; ov.w = ov.z + v[2];
    ot.x = et; ot.y = et + t[0]; ot.z = ot.y + t[1]; ot.w = ot.z + t[2];
    *reinterpret_cast<uint4*>(s.offv + base) = ov;
    *reinterpret_cast<uint4*>(s.offt + base) = ot;
  } else {
#pragma unroll
    for (int i = 0; i < MC_SCAN_ITEMS; ++i) {
      if (base + i < n_seg) { s.offv[base + i] = ev; s.offt[base + i] = et; }
      ev += v[i]; et += t[i];
    }
  }
}

// single block: block totals -> exclusive bases; grand totals + field range -> totals[5]
__global__ void __launch_bounds__(1024) mc_scan_blocks(McScratch s, int n_blocks, int64_t* totals) {
  __shared__ unsigned long long carry[3];
  __shared__ unsigned wsum[3][32];
  if (threadIdx.x < 3) carry[threadIdx.x] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t* arr[3] = {s.bsum_v, s.bsum_t, s.bsum_c};
  for (int base = 0; base < n_blocks; base += 1024) {
    const int idx = base + threadIdx.x;
    unsigned v[3], iv[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      v[q] = idx < n_blocks ? arr[q][idx] : 0u;
      iv[q] = v[q];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const unsigned a = __shfl_up_sync(0xffffffffu, iv[q], o); if (lane >= o) iv[q] += a; }
      if (lane == 31) wsum[q][warp] = iv[q];
    }
    __syncthreads();
    if (warp < 3) {
      const unsigned a = wsum[warp][lane];
      unsigned ia = a;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const unsigned x = __shfl_up_sync(0xffffffffu, ia, o); if (lane >= o) ia += x; }
      wsum[warp][lane] = ia - a;                     // exclusive warp prefixes
    }
    __syncthreads();
    unsigned long long ex[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      ex[q] = carry[q] + wsum[q][warp] + (iv[q] - v[q]);
      if (idx < n_blocks) arr[q][idx] = (unsigned)ex[q];
    }
    __syncthreads();
    if (threadIdx.x == 1023)
      for (int q = 0; q < 3; ++q) carry[q] = ex[q] + v[q];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    totals[0] = (int64_t)carry[0]; totals[1] = (int64_t)carry[1];
    totals[2] = s.minmax[0]; totals[3] = s.minmax[1];
    totals[4] = (int64_t)carry[2];
  }
}

// list[k] = k-th non-empty segment
__global__ void __launch_bounds__(1024) mc_compact(McScratch s, int64_t n_seg) {
  const int64_t base = (int64_t)blockIdx.x * MC_SCAN_BLOCK + (int64_t)threadIdx.x * MC_SCAN_ITEMS;
  bool ne[MC_SCAN_ITEMS];
  uint32_t sc = 0;
#pragma unroll
  for (int i = 0; i < MC_SCAN_ITEMS; ++i) { ne[i] = base + i < n_seg && s.segw[base + i] != 0u; sc += ne[i]; }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t ic = sc;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const uint32_t a = __shfl_up_sync(0xffffffffu, ic, o); if (lane >= o) ic += a; }
  __shared__ uint32_t ws[32];
  if (lane == 31) ws[warp] = ic;
  __syncthreads();
  if (warp == 0) {
    const uint32_t a = ws[lane];
    uint32_t ia = a;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, ia, o); if (lane >= o) ia += x; }
    ws[lane] = ia - a;
  }
  __syncthreads();
  uint32_t k = s.bsum_c[blockIdx.x] + ws[warp] + ic - sc;
#pragma unroll
  for (int i = 0; i < MC_SCAN_ITEMS; ++i)
    if (ne[i]) s.list[k++] = (uint32_t)(base + i);
}

// interpolation parameter along an edge whose end values (minus iso) are f0 (lower point), f1
__device__ __forceinline__ double edge_t(float f0, float f1) {
  const double eps = 1.1920928955078125e-07;   // FLT_EPSILON
  const double w0 = __ddiv_rn(1.0, __dadd_rn(eps, fabs((double)f0)));
  const double w1 = __ddiv_rn(1.0, __dadd_rn(eps, fabs((double)f1)));
  return __ddiv_rn(w1, __dadd_rn(w0, w1));
}

constexpr int MC_EMIT_WARPS = 8;

// one warp per non-empty segment, one lane per grid point
__global__ void __launch_bounds__(32 * MC_EMIT_WARPS) mc_emit(const float* __restrict__ vol, asdf_mc_params prm,
                                                             McScratch s, float* __restrict__ verts,
                                                             float* __restrict__ points, int32_t* __restrict__ faces,
                                                             unsigned long long* __restrict__ keys, int64_t n_list) {
  const McDims d = mc_dims(prm);
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t plane = (int64_t)d.n1 * d.n2;
  const int64_t li = (int64_t)blockIdx.x * MC_EMIT_WARPS + wib;
  if (li >= n_list) return;
  const uint32_t segu = s.list[li];
  const int64_t seg = segu;
  const uint32_t row = segu / (uint32_t)d.nsx, xs = segu - row * (uint32_t)d.nsx;
  const int z = (int)(row / (uint32_t)d.n1), y = (int)(row - (uint32_t)z * (uint32_t)d.n1);
  const int x = (int)xs * 32 + lane;
  const int own = x < d.n2 ? s.info[seg * 32 + lane] : 0;
  const int fl = own & 0xff;
  // exclusive triangle prefix inside the segment
  const int nt = fl >> 4;
  int tsc = nt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int a = __shfl_up_sync(full, tsc, o); if (lane >= o) tsc += a; }
  if (!fl) return;
  const uint32_t basev = s.bsum_v[seg / MC_SCAN_BLOCK] + s.offv[seg];
  const int64_t p = ((int64_t)z * d.n1 + y) * d.n2 + x;
  const double gidx[3] = {(double)(z + prm.index0_offset), (double)y, (double)x};
  const unsigned long long gkey = (unsigned long long)(p + prm.index0_offset * plane) * 4ull;
  const int64_t stride[3] = {plane, (int64_t)d.n2, 1};
  unsigned vo = basev + (own >> 8);
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
  if (!nt) return;
  // ---- this point's cell
  float f[8];
#pragma unroll
  for (int c = 0; c < 8; ++c)
    f[c] = __fsub_rn(__ldg(vol + p + ((c >> 2) & 1) * plane + ((c >> 1) & 1) * d.n2 + (c & 1)), prm.iso);
  const int e = cell_entry(f);
  const unsigned char* tri = kMcTriEdges + 3 * (int)kMcTriStart[e];
  const int32_t centre = (int32_t)(basev + (own >> 8) + __popc(fl & 7));
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
    put_vertex((unsigned)centre, __ddiv_rn(acc[0], n), __ddiv_rn(acc[1], n), __ddiv_rn(acc[2], n), gkey + 3);
  }
  // vertex index of the vertex on cell edge `id`: owned by the edge's lower corner point, possibly in a
  // neighbouring segment (its per-point word holds the flags and the vertex prefix inside that segment)
  auto vindex = [&](int id) -> int32_t {
    if (id == 12) return centre;
    const int c0 = kMcEdgeCorner[id * 2], a = id >> 2;
    int64_t so = seg + ((c0 >> 2) & 1) * ((int64_t)d.n1 * d.nsx) + ((c0 >> 1) & 1) * d.nsx;
    int lo = lane + (c0 & 1);
    if (lo == 32) { lo = 0; ++so; }
    const int w = s.info[so * 32 + lo];
    return (int32_t)(s.bsum_v[so / MC_SCAN_BLOCK] + s.offv[so] + (w >> 8) + __popc(w & ((1 << a) - 1)));
  };
  const unsigned to = s.bsum_t[seg / MC_SCAN_BLOCK] + s.offt[seg] + (unsigned)(tsc - nt);
  for (int t = 0; t < nt; ++t) {
    faces[(size_t)(to + t) * 3 + 0] = vindex(tri[3 * t + 0]);
    faces[(size_t)(to + t) * 3 + 1] = vindex(tri[3 * t + 1]);
    faces[(size_t)(to + t) * 3 + 2] = vindex(tri[3 * t + 2]);
  }
}

__global__ void mc_init_minmax(int* minmax) {
  minmax[0] = 0x7fffffff; minmax[1] = (int)0x80000000;
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
  return align256((size_t)d.n_seg * 4) * 4 + align256((size_t)d.n_seg * 64) + 3 * align256((size_t)d.n_scan_blocks * 4) + 256;
}

extern "C" int asdf_mc_count(const float* vol_dev, const asdf_mc_params* p, void* scratch_dev,
                             int64_t* totals_dev, void* stream) {
  using namespace asdf;
  if (int rc = check_params(p)) return rc;
  ASDF_REQUIRE(vol_dev && scratch_dev && totals_dev, "asdf_mc_count: null argument");
  ASDF_REQUIRE(((uintptr_t)scratch_dev & 255) == 0, "asdf_mc_count: scratch must be 256-byte aligned");
  const McDims d = mc_dims(*p);
  McScratch s = mc_scratch(scratch_dev, d);
  cudaStream_t st = (cudaStream_t)stream;
  mc_init_minmax<<<1, 1, 0, st>>>(s.minmax);
  const dim3 grid((unsigned)d.nsx, (unsigned)((d.n1 + 7) / 8), (unsigned)((d.n0 + MC_ZC - 1) / MC_ZC));
  mc_classify<<<grid, dim3(32, 8), 0, st>>>(vol_dev, *p, s);
  mc_scan_local<<<d.n_scan_blocks, 1024, 0, st>>>(s, d.n_seg);
  mc_scan_blocks<<<1, 1024, 0, st>>>(s, d.n_scan_blocks, totals_dev);
  mc_compact<<<d.n_scan_blocks, 1024, 0, st>>>(s, d.n_seg);
  ASDF_CUDA_CHECK(cudaGetLastError());
  return ASDF_OK;
}

extern "C" int asdf_mc_emit(const float* vol_dev, const asdf_mc_params* p, const void* scratch_dev,
                            int64_t n_segments, float* verts_dev, float* points_dev, int32_t* faces_dev,
                            uint64_t* keys_dev, void* stream) {
  using namespace asdf;
  if (int rc = check_params(p)) return rc;
  ASDF_REQUIRE(vol_dev && scratch_dev && verts_dev && faces_dev, "asdf_mc_emit: null argument");
  const McDims d = mc_dims(*p);
  ASDF_REQUIRE(n_segments >= 0 && n_segments <= d.n_seg, "asdf_mc_emit: n_segments must be totals[4] of asdf_mc_count");
  if (n_segments == 0) return ASDF_OK;
  McScratch s = mc_scratch(const_cast<void*>(scratch_dev), d);
  mc_emit<<<(unsigned)((n_segments + MC_EMIT_WARPS - 1) / MC_EMIT_WARPS), 32 * MC_EMIT_WARPS, 0, (cudaStream_t)stream>>>(
      vol_dev, *p, s, verts_dev, points_dev, faces_dev, (unsigned long long*)keys_dev, n_segments);
  ASDF_CUDA_CHECK(cudaGetLastError());
  return ASDF_OK;
}

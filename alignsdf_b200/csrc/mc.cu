// Marching cubes over a dense f32 field -- HBM bound: the field is read ONCE.
//
// Replaces skimage.measure.marching_cubes_lewiner at utils/mesh.py:354 and
// deep_sdf/mesh.py:81 (CPU, single-threaded Cython) plus the vertex shift of
// utils/mesh.py:360-363.  Topology rule, vertex placement and output ordering are specified
// in alignsdf_b200/mc_tables.py and restated independently in oracle/mc_oracle.py.
//
// Unit of bookkeeping: a SEGMENT = 32 consecutive points along axis 2 at fixed (axis 0, axis 1).
// Output order is grid order, i.e. (segment, lane) order.
//
//   mc_classify   a pure stream over the field: tiles of 128 x 8 x 16 points (+ one halo layer on the upper side of
//                 every axis), one 16-byte load per lane and row (6 in flight per lane), two ballots per row ("some
//                 value below the level", "some value not below") -> 2 x 32 bits per row of 128 points in shared
//                 memory; nothing else is kept and the instruction stream per row is ~14 instructions (the kernel
//                 is issue bound long before it is bandwidth bound otherwise).  One THREAD per row then tests
//                 its own row and its three neighbour rows: all below or all not below => the segment owns
//                 nothing (~98 % of them) and gets a zero word.  The few others are queued in shared memory and
//                 classified by whole warps, one lane per point, re-reading their 2 x 2 rows through L1 / L2
//                 (an earlier variant that staged the tile in shared memory with cp.async was slower: 80 KiB
//                 per block leave two blocks per SM, not enough to overlap the copy with the arithmetic): per
//                 point a word (3 edge-crossing bits, centre bit, 4-bit triangle count | vertex prefix inside the
//                 segment << 8 | case-table entry << 16), per segment one word (vertices | triangles << 16).
//   mc_scan_local exclusive scan of the segment words inside blocks of 4096 segments
//   mc_compact    block bases (every block sums the totals of the blocks before it), the list of non-empty
//                 segments in order with their output bases and per-point words copied next to it (so that
//                 mc_emit starts with independent, coalesced loads), grand totals
//   mc_emit       one warp per non-empty segment, 8 segments per block; the block's vertex jobs are pooled and dealt
//                 out to its threads (fp64 inverse-distance placement on dense warps), (triangle, corner) pairs
//                 are dealt out to the lanes of the segment's warp (coalesced face stores).
// Traffic ~ 4 N^3 (field) + N^3 / 8 (segment words, three passes) + a few % for the surface itself.
#include <type_traits>

#include "common.cuh"
#include "mc_tables.inc"

namespace asdf {
namespace {

constexpr int MC_TX = 128;            // classify tile: points along axis 2 (one 16-byte copy per lane)
constexpr int MC_TY = 8;              //                rows along axis 1
constexpr int MC_ZC = 16;             //                planes along axis 0
constexpr int MC_TILE_SEGS = MC_ZC * MC_TY * (MC_TX / 32);
constexpr int MC_CLASSIFY_THREADS = 32 * (MC_TY + 1);
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

struct McScratch {
  uint32_t* segw;      // [n_seg] vertices | triangles << 16 of the segment
  uint32_t* info;      // [n_seg][32] per point: flags | (vertices owned by the earlier lanes) << 8 | table entry << 16; valid only where segw != 0
  uint32_t* offv;      // [n_seg] exclusive offsets inside the segment's scan block; valid only where segw != 0
  uint32_t* offt;
  uint32_t* list;      // [n_seg] the non-empty segments, in order
  uint2* cbase;        // [n_seg] per list entry: (first vertex, first triangle) of the segment
  uint32_t* cinfo;     // [n_seg][32] per list entry: copy of the segment's per-point words
  uint32_t* bsum_v;    // per scan block: totals
  uint32_t* bsum_t;
  uint32_t* bsum_c;    // same for the number of non-empty segments
  uint32_t* base_v;    // per scan block: exclusive bases (mc_compact)
  uint32_t* base_t;
};

__host__ __device__ inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

__host__ __device__ inline McScratch mc_scratch(void* base, const McDims& d) {
  McScratch s;
  uint8_t* b = (uint8_t*)base;
  s.segw = (uint32_t*)b; b += align256((size_t)d.n_seg * 4);
  s.info = (uint32_t*)b; b += align256((size_t)d.n_seg * 128);
  s.offv = (uint32_t*)b; b += align256((size_t)d.n_seg * 4);
  s.offt = (uint32_t*)b; b += align256((size_t)d.n_seg * 4);
  s.list = (uint32_t*)b; b += align256((size_t)d.n_seg * 4);
  s.cbase = (uint2*)b; b += align256((size_t)d.n_seg * 8);
  s.cinfo = (uint32_t*)b; b += align256((size_t)d.n_seg * 128);
  s.bsum_v = (uint32_t*)b; b += align256((size_t)d.n_scan_blocks * 4);
  s.bsum_t = (uint32_t*)b; b += align256((size_t)d.n_scan_blocks * 4);
  s.bsum_c = (uint32_t*)b; b += align256((size_t)d.n_scan_blocks * 4);
  s.base_v = (uint32_t*)b; b += align256((size_t)d.n_scan_blocks * 4);
  s.base_t = (uint32_t*)b;
  return s;
}

__host__ __device__ inline size_t mc_scratch_size(const McDims& d) {
  return align256((size_t)d.n_seg * 4) * 4 + align256((size_t)d.n_seg * 8) + 2 * align256((size_t)d.n_seg * 128) +
         5 * align256((size_t)d.n_scan_blocks * 4);
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

// two ballots for one row of 128 points held as 4 values per lane: lanes with a value below iso / not below iso
__device__ __forceinline__ uint2 row_mask(const float4 v, const float iso) {
  uint32_t bl, nb;
  asm("{\n\t.reg .pred b, n;\n\t"
      "setp.lt.f32 b, %2, %6;\n\t"
      "setp.lt.or.f32 b, %3, %6, b;\n\t"
      "setp.lt.or.f32 b, %4, %6, b;\n\t"
      "setp.lt.or.f32 b, %5, %6, b;\n\t"
      "setp.lt.f32 n, %2, %6;\n\t"
      "setp.lt.and.f32 n, %3, %6, n;\n\t"
      "setp.lt.and.f32 n, %4, %6, n;\n\t"
      "setp.lt.and.f32 n, %5, %6, n;\n\t"
      "vote.sync.ballot.b32 %0, b, 0xffffffff;\n\t"
      "vote.sync.ballot.b32 %1, !n, 0xffffffff;\n\t}"
      : "=r"(bl), "=r"(nb)
      : "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "f"(iso));
  return make_uint2(bl, nb);
}

__global__ void __launch_bounds__(MC_CLASSIFY_THREADS) mc_classify(const float* __restrict__ vol, asdf_mc_params prm,
                                                                   McScratch s) {
  const McDims d = mc_dims(prm);
  __shared__ uint2 rowm[MC_ZC + 1][MC_TY + 1];                   // per row: (below, notbelow), bit = lane = 4 points
  __shared__ uint8_t halo[MC_ZC + 1][MC_TY + 1];                 // column x0 + 128: 1 below, 2 not below, 0 absent
  __shared__ uint16_t queue[MC_TILE_SEGS];
  __shared__ int n_queue;
  const unsigned full = 0xffffffffu;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int x0 = blockIdx.x * MC_TX, y0 = blockIdx.y * MC_TY, z0 = blockIdx.z * MC_ZC;
  const int nx = min(MC_TX, d.n2 - x0), ny = min(MC_TY + 1, d.n1 - y0), nz = min(MC_ZC + 1, d.n0 - z0);
  const bool has_halo = x0 + MC_TX < d.n2;
  const int64_t plane = (int64_t)d.n1 * d.n2;
  const float iso = prm.iso;
  const float* base = vol + (int64_t)z0 * plane + (int64_t)y0 * d.n2 + x0;
  const int rows = nz * ny;
  if (tid == 0) n_queue = 0;
  // ---- phase A: stream the tile, keep 2 x 32 bits per row.  Warp w owns row y0 + w (8 rows + the halo row) on
  // every plane and walks the planes with a constant pointer step.
  const bool fast = nx == MC_TX && (d.n2 & 3) == 0 && (((uintptr_t)vol) & 15) == 0;
  if (warp < ny) {
    uint2* rm = &rowm[0][warp];
    if (fast && nz == MC_ZC + 1) {                               // interior tile: 17 planes as 6 + 6 + 5 loads in flight
      const char* rowp = reinterpret_cast<const char*>(base + (int64_t)warp * d.n2 + 4 * lane);
      const int64_t pstep = plane * 4;
      auto batch = [&](auto cnt, int pz0) {
        constexpr int U = decltype(cnt)::value;
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { v[u] = __ldg(reinterpret_cast<const float4*>(rowp)); rowp += pstep; }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const uint2 m = row_mask(v[u], iso);
          if (lane == 0) rm[(pz0 + u) * (MC_TY + 1)] = m;
        }
      };
      batch(std::integral_constant<int, 6>{}, 0);
      batch(std::integral_constant<int, 6>{}, 6);
      batch(std::integral_constant<int, 5>{}, 12);
    } else {                                                     // border tiles / unaligned rows: guarded scalar loads
      const float* rowp = base + (int64_t)warp * d.n2;
      for (int pz = 0; pz < nz; ++pz) {
        bool any_b = false, any_n = false;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (4 * lane + c < nx) {
            const bool b = __ldg(rowp + 4 * lane + c) < iso;
            any_b |= b; any_n |= !b;
          }
        }
        const uint32_t bl = __ballot_sync(full, any_b), nb = __ballot_sync(full, any_n);
        if (lane == 0) rm[pz * (MC_TY + 1)] = make_uint2(bl, nb);
        rowp += plane;
      }
    }
  }
  if (tid < rows) {                                              // the halo column (the next tile's first)
    const int pz = tid / ny, py = tid - pz * ny;
    uint8_t h = 0;
    if (has_halo) h = __ldg(base + (int64_t)pz * plane + (int64_t)py * d.n2 + MC_TX) < iso ? 1 : 2;
    halo[pz][py] = h;
  }
  __syncthreads();
  // ---- phase B: one thread per row of the tile's own 16 x 8 rows: which of its 4 segments have anything to do?
  const int zc = min(MC_ZC, d.n0 - z0), yc = min(MC_TY, d.n1 - y0), xsegs = (nx + 31) / 32;
  if (tid < MC_ZC * MC_TY) {
    const int py = tid & 7, pz = tid >> 3;
    if (pz < zc && py < yc) {
      const int py1 = (y0 + py + 1 < d.n1) ? py + 1 : py, pz1 = (z0 + pz + 1 < d.n0) ? pz + 1 : pz;
      const uint2 a = rowm[pz][py], b = rowm[pz][py1], c = rowm[pz1][py], e = rowm[pz1][py1];
      const int h = halo[pz][py] | halo[pz][py1] | halo[pz1][py] | halo[pz1][py1];
      // per segment its 8 lanes + the next lane (whose first value is the segment's x + 1 neighbour; conservative),
      // for the last segment the halo column
      const uint64_t below = (uint64_t)(a.x | b.x | c.x | e.x) | ((uint64_t)(h & 1) << 32);
      const uint64_t notbelow = (uint64_t)(a.y | b.y | c.y | e.y) | ((uint64_t)((h >> 1) & 1) << 32);
      uint32_t* segw = s.segw + ((int64_t)(z0 + pz) * d.n1 + y0 + py) * d.nsx + (x0 >> 5);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j < xsegs) {
          const bool busy = ((below >> (8 * j)) & 0x1ffu) != 0u && ((notbelow >> (8 * j)) & 0x1ffu) != 0u;
          if (busy) queue[atomicAdd(&n_queue, 1)] = (uint16_t)(tid * 4 + j);
          else segw[j] = 0u;
        }
      }
    }
  }
  __syncthreads();
  // ---- phase C: a warp per queued segment, a lane per point; the 2 x 2 rows come back through L1 / L2
  const int nq = n_queue;
  for (int k = warp; k < nq; k += MC_TY + 1) {
    const int idx = queue[k];
    const int j = idx & 3, py = (idx >> 2) & 7, pz = idx >> 5;
    const int z = z0 + pz, y = y0 + py, x = x0 + 32 * j + lane;
    const bool vx = x < d.n2, e0 = z + 1 < d.n0, e1 = y + 1 < d.n1, e2 = x + 1 < d.n2;
    const float* p = vol + (int64_t)z * plane + (int64_t)y * d.n2 + x;
    float g[4], h[4];                                            // g: (d0, d1) at x;  h: at x + 1
    g[0] = vx ? __ldg(p) : 0.f;
    g[1] = (vx && e1) ? __ldg(p + d.n2) : 0.f;
    g[2] = (vx && e0) ? __ldg(p + plane) : 0.f;
    g[3] = (vx && e0 && e1) ? __ldg(p + plane + d.n2) : 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) h[q] = __shfl_down_sync(full, g[q], 1);
    if (lane == 31) {
      h[0] = e2 ? __ldg(p + 1) : 0.f;
      h[1] = (e2 && e1) ? __ldg(p + d.n2 + 1) : 0.f;
      h[2] = (e2 && e0) ? __ldg(p + plane + 1) : 0.f;
      h[3] = (e2 && e0 && e1) ? __ldg(p + plane + d.n2 + 1) : 0.f;
    }
    int fl = 0, ent = 0;
    if (vx) {
      float f[8];                                                // corner c = 4 d0 + 2 d1 + d2, 0 where the corner does not exist
      f[0] = __fsub_rn(g[0], iso);
      f[1] = e2 ? __fsub_rn(h[0], iso) : 0.f;
      f[2] = e1 ? __fsub_rn(g[1], iso) : 0.f;
      f[3] = (e1 && e2) ? __fsub_rn(h[1], iso) : 0.f;
      f[4] = e0 ? __fsub_rn(g[2], iso) : 0.f;
      f[5] = (e0 && e2) ? __fsub_rn(h[2], iso) : 0.f;
      f[6] = (e0 && e1) ? __fsub_rn(g[3], iso) : 0.f;
      f[7] = (e0 && e1 && e2) ? __fsub_rn(h[3], iso) : 0.f;
      const bool in0 = f[0] < 0.f;
      if (e0) fl |= (in0 != (f[4] < 0.f)) << 0;
      if (e1) fl |= (in0 != (f[2] < 0.f)) << 1;
      if (e2) fl |= (in0 != (f[1] < 0.f)) << 2;
      if (e0 && e1 && e2) {
        const int e = cell_entry(f);
        if (e >= 0) {
          const int nt = kMcNumTris[e];
          fl |= ((nt >> 7) << 3) | ((nt & 0x7f) << 4);
          ent = e;
        }
      }
    }
    const int inc = __popc(fl & 0xf);
    const int wv = __reduce_add_sync(full, inc), wt = __reduce_add_sync(full, fl >> 4);
    const int64_t seg = ((int64_t)z * d.n1 + y) * d.nsx + (x0 >> 5) + j;
    if (lane == 0) s.segw[seg] = (uint32_t)wv | ((uint32_t)wt << 16);
    if (wv | wt) {
      int sc = inc;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int a = __shfl_up_sync(full, sc, o); if (lane >= o) sc += a; }
      s.info[seg * 32 + lane] = (uint32_t)fl | ((uint32_t)(sc - inc) << 8) | ((uint32_t)ent << 16);
    }
  }
}

// exclusive scan of the segment words inside blocks of MC_SCAN_BLOCK segments; block totals -> bsum
__global__ void __launch_bounds__(1024) mc_scan_local(McScratch s, int64_t n_seg) {
  const int64_t base = (int64_t)blockIdx.x * MC_SCAN_BLOCK + (int64_t)threadIdx.x * MC_SCAN_ITEMS;
  uint32_t ws[MC_SCAN_ITEMS];
  if (base + MC_SCAN_ITEMS <= n_seg) {
    const uint4 w = *reinterpret_cast<const uint4*>(s.segw + base);
    ws[0] = w.x; ws[1] = w.y; ws[2] = w.z; ws[3] = w.w;
  } else {
#pragma unroll
    for (int i = 0; i < MC_SCAN_ITEMS; ++i) ws[i] = base + i < n_seg ? s.segw[base + i] : 0u;
  }
  uint32_t sv = 0, st = 0, sc = 0;
#pragma unroll
  for (int i = 0; i < MC_SCAN_ITEMS; ++i) { sv += ws[i] & 0xffffu; st += ws[i] >> 16; sc += ws[i] != 0u; }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // inclusive warp scans of (vertices, triangles), warp sum of the non-empty count
  uint32_t iv = sv, it = st;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t a = __shfl_up_sync(0xffffffffu, iv, o), b = __shfl_up_sync(0xffffffffu, it, o);
    if (lane >= o) { iv += a; it += b; }
  }
  const uint32_t wc = __reduce_add_sync(0xffffffffu, sc);
  __shared__ uint32_t wsv[32], wst[32], wsc[32];
  if (lane == 31) { wsv[warp] = iv; wst[warp] = it; }
  if (lane == 0) wsc[warp] = wc;
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
    const uint32_t tc = __reduce_add_sync(0xffffffffu, wsc[lane]);
    if (lane == 31) { s.bsum_v[blockIdx.x] = ia; s.bsum_t[blockIdx.x] = ib; s.bsum_c[blockIdx.x] = tc; }
  }
  __syncthreads();
  if (sc == 0u) return;                                          // offsets are only ever read for non-empty segments
  uint32_t ev = wsv[warp] + iv - sv, et = wst[warp] + it - st;
#pragma unroll
  for (int i = 0; i < MC_SCAN_ITEMS; ++i) {
    if (ws[i] != 0u) { s.offv[base + i] = ev; s.offt[base + i] = et; }
    ev += ws[i] & 0xffffu; et += ws[i] >> 16;
  }
}

// block bases (sum of the totals of the blocks before this one); list[k] = k-th non-empty segment together with its
// output bases and a copy of its per-point words; grand totals
__global__ void __launch_bounds__(1024) mc_compact(McScratch s, int64_t n_seg, int n_blocks, int64_t* totals) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ uint32_t red[3][32];
  __shared__ uint32_t ws[32];
  __shared__ uint32_t found[MC_SCAN_BLOCK];                      // the block's non-empty segments, in order
  __shared__ uint32_t n_found;
  uint32_t bv = 0, bt = 0, bc = 0;
  for (int i = threadIdx.x; i < (int)blockIdx.x; i += 1024) { bv += s.bsum_v[i]; bt += s.bsum_t[i]; bc += s.bsum_c[i]; }
  bv = __reduce_add_sync(0xffffffffu, bv); bt = __reduce_add_sync(0xffffffffu, bt); bc = __reduce_add_sync(0xffffffffu, bc);
  if (lane == 0) { red[0][warp] = bv; red[1][warp] = bt; red[2][warp] = bc; }
  const int64_t base = (int64_t)blockIdx.x * MC_SCAN_BLOCK + (int64_t)threadIdx.x * MC_SCAN_ITEMS;
  bool ne[MC_SCAN_ITEMS];
  uint32_t sc = 0;
  if (base + MC_SCAN_ITEMS <= n_seg) {
    const uint4 w = *reinterpret_cast<const uint4*>(s.segw + base);
    ne[0] = w.x != 0u; ne[1] = w.y != 0u; ne[2] = w.z != 0u; ne[3] = w.w != 0u;
  } else {
#pragma unroll
    for (int i = 0; i < MC_SCAN_ITEMS; ++i) ne[i] = base + i < n_seg && s.segw[base + i] != 0u;
  }
#pragma unroll
  for (int i = 0; i < MC_SCAN_ITEMS; ++i) sc += ne[i];
  uint32_t ic = sc;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const uint32_t a = __shfl_up_sync(0xffffffffu, ic, o); if (lane >= o) ic += a; }
  if (lane == 31) ws[warp] = ic;
  __syncthreads();
  if (warp == 0) {
    const uint32_t a = ws[lane];
    uint32_t ia = a;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, ia, o); if (lane >= o) ia += x; }
    ws[lane] = ia - a;
    if (lane == 31) n_found = ia;
    bv = __reduce_add_sync(0xffffffffu, red[0][lane]); bt = __reduce_add_sync(0xffffffffu, red[1][lane]);
    bc = __reduce_add_sync(0xffffffffu, red[2][lane]);
    if (lane == 0) {
      red[0][0] = bv; red[1][0] = bt; red[2][0] = bc;
      s.base_v[blockIdx.x] = bv; s.base_t[blockIdx.x] = bt;
      if ((int)blockIdx.x == n_blocks - 1) {
        totals[0] = (int64_t)bv + s.bsum_v[blockIdx.x]; totals[1] = (int64_t)bt + s.bsum_t[blockIdx.x];
        totals[2] = 0; totals[3] = 0;
        totals[4] = (int64_t)bc + s.bsum_c[blockIdx.x];
      }
    }
  }
  __syncthreads();
  const uint32_t kbase = red[2][0];
  if (sc != 0u) {
    uint32_t k = ws[warp] + ic - sc;
#pragma unroll
    for (int i = 0; i < MC_SCAN_ITEMS; ++i)
      if (ne[i]) {
        const int64_t seg = base + i;
        found[k] = (uint32_t)seg;
        s.list[kbase + k] = (uint32_t)seg;
        s.cbase[kbase + k] = make_uint2(red[0][0] + s.offv[seg], red[1][0] + s.offt[seg]);
        ++k;
      }
  }
  __syncthreads();
  const uint32_t nf = n_found;
  for (uint32_t i = warp; i < nf; i += 32)
    s.cinfo[(int64_t)(kbase + i) * 32 + lane] = s.info[(int64_t)found[i] * 32 + lane];
}

// interpolation parameter along an edge whose end values (minus iso) are f0 (lower point), f1
__device__ __forceinline__ double edge_t(float f0, float f1) {
  const double eps = 1.1920928955078125e-07;   // FLT_EPSILON
  const double w0 = __ddiv_rn(1.0, __dadd_rn(eps, fabs((double)f0)));
  const double w1 = __ddiv_rn(1.0, __dadd_rn(eps, fabs((double)f1)));
  return __ddiv_rn(w1, __dadd_rn(w0, w1));
}

constexpr int MC_EMIT_WARPS = 8;

// One warp per non-empty segment, 8 segments per block.  Everything a warp needs first (segment id, per-point words,
// output bases) sits at its list position: three independent coalesced loads.
__global__ void __launch_bounds__(32 * MC_EMIT_WARPS, 5) mc_emit(const float* __restrict__ vol, asdf_mc_params prm,
                                                                McScratch s, float* __restrict__ verts,
                                                                float* __restrict__ points, int32_t* __restrict__ faces,
                                                                unsigned long long* __restrict__ keys, int64_t n_list) {
  __shared__ uint32_t own_s[MC_EMIT_WARPS][32];                  // the segments' per-point words
  __shared__ uint32_t seg_s[MC_EMIT_WARPS], basev_s[MC_EMIT_WARPS];
  __shared__ uint16_t tpre_s[MC_EMIT_WARPS][32];                 // exclusive triangle prefix per point
  __shared__ uint8_t tjob_s[MC_EMIT_WARPS][384];                 // triangle -> lane
  __shared__ uint16_t vjob_s[MC_EMIT_WARPS * 96];                // the block's vertex jobs: warp << 8 | axis << 5 | lane
  __shared__ int n_vjobs;
  const McDims d = mc_dims(prm);
  const unsigned full = 0xffffffffu;
  const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
  const int64_t plane = (int64_t)d.n1 * d.n2;
  const int64_t li = (int64_t)blockIdx.x * MC_EMIT_WARPS + wib;
  const bool live = li < n_list;
  if (tid == 0) n_vjobs = 0;
  uint32_t own = 0, segu = 0;
  uint2 cb = make_uint2(0u, 0u);
  if (live) {
    segu = __ldg(s.list + li);
    own = __ldg(s.cinfo + li * 32 + lane);
    cb = __ldg(s.cbase + li);
  }
  const int64_t seg = segu;
  const uint32_t basev = cb.x, baset = cb.y;
  const int fl = own & 0xff, nt = fl >> 4;
  own_s[wib][lane] = own;
  if (lane == 0) { seg_s[wib] = segu; basev_s[wib] = basev; }
  int tsc = nt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int a = __shfl_up_sync(full, tsc, o); if (lane >= o) tsc += a; }
  const int T = __shfl_sync(full, tsc, 31);
  tpre_s[wib][lane] = (uint16_t)(tsc - nt);
  for (int t = 0; t < nt; ++t) tjob_s[wib][tsc - nt + t] = (uint8_t)lane;
  const uint32_t m0 = __ballot_sync(full, fl & 1), m1 = __ballot_sync(full, fl & 2), m2 = __ballot_sync(full, fl & 4);
  const int c0 = __popc(m0), c1 = __popc(m1), nvj = c0 + c1 + __popc(m2);
  __syncthreads();                                               // n_vjobs = 0 visible
  {
    int start = 0;
    if (lane == 0 && nvj) start = atomicAdd(&n_vjobs, nvj);
    start = __shfl_sync(full, start, 0);
    const uint32_t lt = (1u << lane) - 1u;
    const int me = (wib << 8) | lane;
    if (fl & 1) vjob_s[start + __popc(m0 & lt)] = (uint16_t)me;
    if (fl & 2) vjob_s[start + c0 + __popc(m1 & lt)] = (uint16_t)(me | 32);
    if (fl & 4) vjob_s[start + c0 + c1 + __popc(m2 & lt)] = (uint16_t)(me | 64);
  }
  __syncthreads();
  double spacing[3] = {prm.spacing[0], prm.spacing[1], prm.spacing[2]};
  float origin[3] = {prm.origin[0], prm.origin[1], prm.origin[2]};
  if (prm.grid_dev) {                                            // lattice produced on the device (asdf_regrid)
    const float v = __ldg(prm.grid_dev);
    spacing[0] = spacing[1] = spacing[2] = (double)v;
    origin[0] = __ldg(prm.grid_dev + 1); origin[1] = __ldg(prm.grid_dev + 2); origin[2] = __ldg(prm.grid_dev + 3);
  }
  auto put_vertex = [&](unsigned slot, double q0, double q1, double q2, unsigned long long key) {
    const float x0 = (float)__dmul_rn(q0, spacing[0]);
    const float x1 = (float)__dmul_rn(q1, spacing[1]);
    const float x2 = (float)__dmul_rn(q2, spacing[2]);
    verts[(size_t)slot * 3 + 0] = x0; verts[(size_t)slot * 3 + 1] = x1; verts[(size_t)slot * 3 + 2] = x2;
    if (points) {
      points[(size_t)slot * 3 + 0] = __fadd_rn(origin[0], x0);
      points[(size_t)slot * 3 + 1] = __fadd_rn(origin[1], x1);
      points[(size_t)slot * 3 + 2] = __fadd_rn(origin[2], x2);
    }
    if (keys) keys[slot] = key;
  };
  // ---- vertices on the edges owned by the block's points: the pooled jobs, one per thread
  const int nj = n_vjobs;
  for (int i = tid; i < nj; i += 32 * MC_EMIT_WARPS) {
    const int job = vjob_s[i], w = job >> 8, a = (job >> 5) & 3, l = job & 31;
    const uint32_t ow = own_s[w][l], sg = seg_s[w];
    const uint32_t row = sg / (uint32_t)d.nsx, xs = sg - row * (uint32_t)d.nsx;
    const int z = (int)(row / (uint32_t)d.n1), y = (int)(row - (uint32_t)z * (uint32_t)d.n1), x = (int)xs * 32 + l;
    const int64_t p = ((int64_t)z * d.n1 + y) * d.n2 + x;
    const float f0 = __fsub_rn(__ldg(vol + p), prm.iso);
    const float f1 = __fsub_rn(__ldg(vol + p + (a == 0 ? plane : a == 1 ? (int64_t)d.n2 : (int64_t)1)), prm.iso);
    double q0 = (double)(z + prm.index0_offset), q1 = (double)y, q2 = (double)x;
    const double t = edge_t(f0, f1);
    if (a == 0) q0 = __dadd_rn(q0, t); else if (a == 1) q1 = __dadd_rn(q1, t); else q2 = __dadd_rn(q2, t);
    const unsigned slot = basev_s[w] + ((ow >> 8) & 0xffu) + __popc(ow & ((1u << a) - 1u));
    put_vertex(slot, q0, q1, q2, (unsigned long long)(p + prm.index0_offset * plane) * 4ull + a);
  }
  if (!live) return;
  const uint32_t row = segu / (uint32_t)d.nsx, xs = segu - row * (uint32_t)d.nsx;
  const int z = (int)(row / (uint32_t)d.n1), y = (int)(row - (uint32_t)z * (uint32_t)d.n1);
  // ---- centre vertices (rare): mean (fp64, loop order) of the loop's vertices in index space
  if (fl & 8) {
    const int64_t p = ((int64_t)z * d.n1 + y) * d.n2 + (int64_t)xs * 32 + lane;
    float f[8];
#pragma unroll
    for (int c = 0; c < 8; ++c)
      f[c] = __fsub_rn(__ldg(vol + p + ((c >> 2) & 1) * plane + ((c >> 1) & 1) * d.n2 + (c & 1)), prm.iso);
    const unsigned char* tri = kMcTriEdges + 3 * (int)kMcTriStart[own >> 16];
    const double gidx[3] = {(double)(z + prm.index0_offset), (double)y, (double)((int)xs * 32 + lane)};
    double acc[3] = {0.0, 0.0, 0.0};
    int cnt = 0;
    for (int t = 0; t < nt; ++t) {
      if (tri[3 * t] != 12) continue;
      const int id = tri[3 * t + 1];
      const int k0 = kMcEdgeCorner[id * 2], k1 = kMcEdgeCorner[id * 2 + 1];
      const int a = id >> 2;
      double q[3] = {gidx[0] + ((k0 >> 2) & 1), gidx[1] + ((k0 >> 1) & 1), gidx[2] + (k0 & 1)};
      const double tt = edge_t(f[k0], f[k1]);
      if (a == 0) q[0] = __dadd_rn(q[0], tt); else if (a == 1) q[1] = __dadd_rn(q[1], tt); else q[2] = __dadd_rn(q[2], tt);
      acc[0] = __dadd_rn(acc[0], q[0]); acc[1] = __dadd_rn(acc[1], q[1]); acc[2] = __dadd_rn(acc[2], q[2]);
      ++cnt;
    }
    const double n = (double)cnt;
    put_vertex(basev + ((own >> 8) & 0xffu) + __popc(fl & 7), __ddiv_rn(acc[0], n), __ddiv_rn(acc[1], n),
               __ddiv_rn(acc[2], n), (unsigned long long)(p + prm.index0_offset * plane) * 4ull + 3ull);
  }
  // ---- faces: one (triangle, corner) pair per lane, consecutive lanes write consecutive ints
  const int64_t seg_dz = (int64_t)d.n1 * d.nsx;
  for (int i = lane; i < 3 * T; i += 32) {
    const int g = i / 3, c = i - 3 * g;
    const int l = tjob_s[wib][g];
    const uint32_t ow = own_s[wib][l];
    const int t = g - (int)tpre_s[wib][l];
    const int id = kMcTriEdges[3 * ((int)kMcTriStart[ow >> 16] + t) + c];
    int32_t vi;
    if (id == 12) {
      vi = (int32_t)(basev + ((ow >> 8) & 0xffu) + __popc(ow & 7u));
    } else {
      // the vertex on cell edge id = 4 a + 2 o_b + o_c (axes b < c other than a) is owned by the edge's lower
      // corner point, possibly in a neighbouring segment
      const int a = id >> 2, ob = (id >> 1) & 1, oc = id & 1;
      const int dz = a == 0 ? 0 : ob, dy = a == 0 ? ob : (a == 1 ? 0 : oc), dx = a == 2 ? 0 : oc;
      int lo = l + dx;
      int64_t so = seg + dz * seg_dz + dy * d.nsx;
      if (lo == 32) { lo = 0; ++so; }
      uint32_t w, bse;
      if (so == seg) { w = own_s[wib][lo]; bse = basev; }
      else { w = s.info[so * 32 + lo]; bse = s.base_v[so / MC_SCAN_BLOCK] + s.offv[so]; }
      vi = (int32_t)(bse + ((w >> 8) & 0xffu) + __popc(w & ((1u << a) - 1u)));
    }
    faces[(size_t)baset * 3 + i] = vi;
  }
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
  return mc_scratch_size(mc_dims(*p));
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
  const dim3 grid((unsigned)((d.n2 + MC_TX - 1) / MC_TX), (unsigned)((d.n1 + MC_TY - 1) / MC_TY),
                  (unsigned)((d.n0 + MC_ZC - 1) / MC_ZC));
  mc_classify<<<grid, MC_CLASSIFY_THREADS, 0, st>>>(vol_dev, *p, s);
  mc_scan_local<<<d.n_scan_blocks, 1024, 0, st>>>(s, d.n_seg);
  mc_compact<<<d.n_scan_blocks, 1024, 0, st>>>(s, d.n_seg, d.n_scan_blocks, totals_dev);
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

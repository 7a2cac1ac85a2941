// Generic fp32 decoder evaluation (CUDA cores).  Handles every decoder topology the reference
// can express after folding (any widths <= 512, any skip layout, CombinedDecoder with classifier
// head, feature-mode queries).  The shipped 2 x 5-layer x 512 topology normally runs on the
// tcgen05 kernel in k1_tc.cu; this kernel is the exact-fp32 path for everything else.
//
// Replaces, per chunk of points (SURVEY.md §2b K1/K1'):
//   utils/mesh.py:24-63,82-115        grid construction + chunk loop + H2D/D2H
//   utils/utils.py:376-430,561-572    kinematic_embedding + latent expand/cat
//   networks/model.py:139-188,285-350 decoder forward
//   utils/mesh.py:207-247             nonzero(sdf<0) bounding box (fused as atomics)
#include "common.cuh"

namespace asdf {
namespace {

constexpr int PT = 32;        // points per tile
constexpr int PTS = 36;       // padded row stride (floats) of the k-major activation tiles
constexpr int NT = 256;       // threads per block
constexpr int MAXW = 512;     // widest layer supported

struct SimtArgs {
  asdf_simt_desc d;
  asdf_query q;
  const float* stat;
  const float* samp;
  const float* cls;
  float* out_hand;
  float* out_obj;
  int32_t* out_cls;
  float* out_logits;
  int32_t* bbox;
};

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// out[n][p] = relu( sum_k WxT[k][n] in[k][p] + sum_d MB[n][d] u[d][p] + MB[n][D] )   (no relu when `raw`:
// a LayerNorm follows)
// ``pa_map`` != nullptr: this layer sees the per-point PixelAlign latent -- add the bicubic combination
// sum_t tap_w[t][p] * pa_map[tap_i[t][p]][n] of the projected feature map's rows (asdf_pixel_align).
__device__ void hidden_layer(const float* __restrict__ wxt, const float* __restrict__ mb,
                             int h, int n, int npad, int has_m, int D,
                             const float* __restrict__ in, const float* __restrict__ u,
                             float* __restrict__ out, bool raw, const float* __restrict__ pa_map = nullptr,
                             const int* __restrict__ tap_i = nullptr, const float* __restrict__ tap_w = nullptr) {
  const int tn = threadIdx.x & 63;
  const int p0 = (threadIdx.x >> 6) * 8;
  float acc[8][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int nn = tn + 64 * j;
    const float b = nn < npad ? __ldg(mb + (size_t)nn * (D + 1) + D) : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[j][i] = b;
  }
  if (pa_map) {
    for (int t = 0; t < 16; ++t) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float w = tap_w[t * PT + p0 + i];
        if (w != 0.f) {                              // uniform over the warp (all its lanes share p0)
          const float* row = pa_map + (size_t)tap_i[t * PT + p0 + i] * npad + tn;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (tn + 64 * j < npad) acc[j][i] = fmaf(w, __ldg(row + 64 * j), acc[j][i]);
        }
      }
    }
  }
  if (has_m) {
    for (int dd = 0; dd < D; ++dd) {
      const float4 ua = ld4(u + dd * PTS + p0), ub = ld4(u + dd * PTS + p0 + 4);
      const float x[8] = {ua.x, ua.y, ua.z, ua.w, ub.x, ub.y, ub.z, ub.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int nn = tn + 64 * j;
        const float w = nn < npad ? __ldg(mb + (size_t)nn * (D + 1) + dd) : 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[j][i] = fmaf(w, x[i], acc[j][i]);
      }
    }
  }
  const int jmax = (npad - tn + 63) / 64;   // number of live feature slots of this thread
  // The weights of the NEXT block of KU k-rows are fetched (L1 / L2) while the current block is multiplied: with one
  // 256-thread block per SM only two warps share a scheduler, too few to hide the load latency by themselves
  // (unroll 2 without the prefetch: 1.65 x slower, tools/variant_timing.py).  Same fmaf sequence per accumulator.
  constexpr int KU = 4;
  const int hb = (h / KU) * KU;
  {
    float wn[KU][8];
    auto fetch = [&](int k0) {
#pragma unroll
      for (int uu = 0; uu < KU; ++uu) {
        const float* wr = wxt + (size_t)(k0 + uu) * npad + tn;
#pragma unroll
        for (int j = 0; j < 8; ++j) wn[uu][j] = j < jmax ? __ldg(wr + 64 * j) : 0.f;
      }
    };
    if (hb > 0) fetch(0);
    for (int k0 = 0; k0 < hb; k0 += KU) {
      float wc[KU][8];
#pragma unroll
      for (int uu = 0; uu < KU; ++uu)
#pragma unroll
        for (int j = 0; j < 8; ++j) wc[uu][j] = wn[uu][j];
      if (k0 + KU < hb) fetch(k0 + KU);
#pragma unroll
      for (int uu = 0; uu < KU; ++uu) {
        const float4 xa = ld4(in + (k0 + uu) * PTS + p0), xb = ld4(in + (k0 + uu) * PTS + p0 + 4);
        const float x[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (j < jmax) {
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[j][i] = fmaf(wc[uu][j], x[i], acc[j][i]);
          }
        }
      }
    }
  }
#pragma unroll KU
  for (int k = hb; k < h; ++k) {       // the h % KU rows left over
    const float4 xa = ld4(in + k * PTS + p0), xb = ld4(in + k * PTS + p0 + 4);
    const float x[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
    const float* wr = wxt + (size_t)k * npad + tn;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j < jmax) {
        const float w = __ldg(wr + 64 * j);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[j][i] = fmaf(w, x[i], acc[j][i]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int nn = tn + 64 * j;
    if (nn < n) {
      const float lo = raw ? -3.402823466e38f : 0.f;
      float4 a = make_float4(fmaxf(acc[j][0], lo), fmaxf(acc[j][1], lo), fmaxf(acc[j][2], lo), fmaxf(acc[j][3], lo));
      float4 b = make_float4(fmaxf(acc[j][4], lo), fmaxf(acc[j][5], lo), fmaxf(acc[j][6], lo), fmaxf(acc[j][7], lo));
      *reinterpret_cast<float4*>(out + nn * PTS + p0) = a;
      *reinterpret_cast<float4*>(out + nn * PTS + p0 + 4) = b;
    }
  }
}

// x[k][p] <- relu(LayerNorm_k(x[.][p]) * gamma[k] + beta[k]) over the n features of every point of the
// tile (nn.LayerNorm, eps 1e-5, biased variance, two-pass moments): one warp per 4 points.
__device__ void layer_norm_relu(float* __restrict__ x, int n, const float* __restrict__ gamma,
                                const float* __restrict__ beta) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* col = x + warp * 4;
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k = lane; k < n; k += 32) {
    const float4 v = ld4(col + k * PTS);
    s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
  }
  float mean[4], rstd[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s[i] += __shfl_xor_sync(0xffffffffu, s[i], off);
    mean[i] = s[i] / (float)n;
    s[i] = 0.f;
  }
  for (int k = lane; k < n; k += 32) {
    const float4 v = ld4(col + k * PTS);
    const float d0 = v.x - mean[0], d1 = v.y - mean[1], d2 = v.z - mean[2], d3 = v.w - mean[3];
    s[0] = fmaf(d0, d0, s[0]); s[1] = fmaf(d1, d1, s[1]); s[2] = fmaf(d2, d2, s[2]); s[3] = fmaf(d3, d3, s[3]);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s[i] += __shfl_xor_sync(0xffffffffu, s[i], off);
    rstd[i] = 1.0f / sqrtf(s[i] / (float)n + 1e-5f);
  }
  for (int k = lane; k < n; k += 32) {
    const float g = __ldg(gamma + k), b = __ldg(beta + k);
    float4 v = ld4(col + k * PTS);
    v.x = fmaxf(fmaf((v.x - mean[0]) * rstd[0], g, b), 0.f);
    v.y = fmaxf(fmaf((v.y - mean[1]) * rstd[1], g, b), 0.f);
    v.z = fmaxf(fmaf((v.z - mean[2]) * rstd[2], g, b), 0.f);
    v.w = fmaxf(fmaf((v.w - mean[3]) * rstd[3], g, b), 0.f);
    *reinterpret_cast<float4*>(col + k * PTS) = v;
  }
}

// Small-width head: y[o][p] = sum_k W[k*ldw + o] in[k][p] (+ M u + B), one warp per 4 points.
// Result (before any activation) is left in res[o*PT + p].
__device__ void small_head(const float* __restrict__ wxt, int ldw, const float* __restrict__ mb,
                           int h, int n, int has_m, int D, const float* __restrict__ in,
                           const float* __restrict__ u, float* __restrict__ res) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int o = 0; o < n; ++o) {
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = lane; k < h; k += 32) {
      const float w = __ldg(wxt + (size_t)k * ldw + o);
      const float4 x = ld4(in + k * PTS + warp * 4);
      s[0] = fmaf(w, x.x, s[0]); s[1] = fmaf(w, x.y, s[1]);
      s[2] = fmaf(w, x.z, s[2]); s[3] = fmaf(w, x.w, s[3]);
    }
    if (has_m) {
      for (int dd = lane; dd < D; dd += 32) {
        const float w = __ldg(mb + (size_t)o * (D + 1) + dd);
        const float4 x = ld4(u + dd * PTS + warp * 4);
        s[0] = fmaf(w, x.x, s[0]); s[1] = fmaf(w, x.y, s[1]);
        s[2] = fmaf(w, x.z, s[2]); s[3] = fmaf(w, x.w, s[3]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) s[i] += __shfl_xor_sync(0xffffffffu, s[i], off);
    }
    if (lane == 0) {
      const float b = mb ? __ldg(mb + (size_t)o * (D + 1) + D) : 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) res[o * PT + warp * 4 + i] = s[i] + b;
    }
  }
}

// torch's bicubic convolution coefficients (A = -0.75) for the taps at offsets -1, 0, 1, 2 around floor(x)
__device__ __forceinline__ void cubic_coeffs(float t, float* c) {
  const float A = -0.75f;
  const float x0 = t + 1.f, x1 = t, x2 = 1.f - t, x3 = 2.f - t;
  c[0] = ((A * x0 - 5.f * A) * x0 + 8.f * A) * x0 - 4.f * A;
  c[1] = ((A + 2.f) * x1 - (A + 3.f)) * x1 * x1 + 1.f;
  c[2] = ((A + 2.f) * x2 - (A + 3.f)) * x2 * x2 + 1.f;
  c[3] = ((A * x3 - 5.f * A) * x3 + 8.f * A) * x3 - 4.f * A;
}

// PixelAlign taps of query i (utils/utils.py:536-558): project, normalise to [-1,1], bicubic grid_sample with
// align_corners=True and zero padding; a point whose projection leaves the image takes the mean feature (map row
// fh*fw) instead.  ti / tw: this point's column of the [16][PT] tap tables.
__device__ void pixel_align_taps(const asdf_pixel_align& pa, const asdf_query& q, int64_t i, int* ti, float* tw) {
#pragma unroll
  for (int t = 0; t < 16; ++t) { ti[t * PT] = 0; tw[t * PT] = 0.f; }
  if (i >= q.end) return;
  float x0, x1, x2;
  if (q.mode == ASDF_QUERY_POINTS) {
    const float* row = q.points_dev + (size_t)i * q.point_stride;
    x0 = __ldg(row); x1 = __ldg(row + 1); x2 = __ldg(row + 2);
  } else {
    grid_point(i, q.N, q.mode, q.voxel, q.origin[0], q.origin[1], q.origin[2], x0, x1, x2);
  }
  float c[3], h[3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
    c[r] = pa.point_affine[4 * r] * x0 + pa.point_affine[4 * r + 1] * x1 + pa.point_affine[4 * r + 2] * x2 + pa.point_affine[4 * r + 3];
#pragma unroll
  for (int r = 0; r < 3; ++r) h[r] = pa.cam[4 * r] * c[0] + pa.cam[4 * r + 1] * c[1] + pa.cam[4 * r + 2] * c[2] + pa.cam[4 * r + 3];
  const float u = h[0] / h[2] / pa.image_size * 2.f - 1.f, v = h[1] / h[2] / pa.image_size * 2.f - 1.f;
  if (!(u >= -1.f && u <= 1.f && v >= -1.f && v <= 1.f)) {       // also NaN: the mean feature (utils/utils.py:553-555)
    ti[0] = pa.fh * pa.fw; tw[0] = 1.f;
    return;
  }
  const float ix = (u + 1.f) * 0.5f * (float)(pa.fw - 1), iy = (v + 1.f) * 0.5f * (float)(pa.fh - 1);
  const float fx = floorf(ix), fy = floorf(iy);
  float cx[4], cy[4];
  cubic_coeffs(ix - fx, cx);
  cubic_coeffs(iy - fy, cy);
  const int bx = (int)fx - 1, by = (int)fy - 1;
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int xx = bx + k, yy = by + j;
      if (xx >= 0 && xx < pa.fw && yy >= 0 && yy < pa.fh) {
        ti[(4 * j + k) * PT] = yy * pa.fw + xx;
        tw[(4 * j + k) * PT] = cy[j] * cx[k];
      }
    }
}

__global__ void __launch_bounds__(NT, 1) simt_eval_kernel(const SimtArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* act0 = smem;
  float* act1 = act0 + MAXW * PTS;
  float* u = act1 + MAXW * PTS;                       // [ASDF_MAX_POINT_DIM][PTS]
  float* res = u + ASDF_MAX_POINT_DIM * PTS;          // [8][PT] head outputs
  float* sdf = res + 8 * PT;                          // [2][PT]
  int* cls_s = reinterpret_cast<int*>(sdf + 2 * PT);  // [PT]
  int* tap_i = cls_s + PT;                            // [16][PT] PixelAlign: map rows of the bicubic taps
  float* tap_w = reinterpret_cast<float*>(tap_i + 16 * PT);   // [16][PT] and their weights

  const asdf_simt_desc& d = a.d;
  const asdf_query& q = a.q;
  const int64_t total = q.end - q.begin;
  const int64_t n_tiles = (total + PT - 1) / PT;
  const int L = d.n_layers;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t base = q.begin + tile * PT;
    if (d.pa.enabled) {
      __syncthreads();
      if (threadIdx.x < PT) pixel_align_taps(d.pa, q, base + threadIdx.x, tap_i + threadIdx.x, tap_w + threadIdx.x);
    }
    for (int b = 0; b < d.n_branches; ++b) {
      const int D = d.point_dim[b];
      __syncthreads();
      // ---- per-point input vector u[d][p]
      if (threadIdx.x < PT) {
        const int p = threadIdx.x;
        const int64_t i = base + p;
        if (i < q.end) {
          if (d.nerf_freqs > 0) {                  // positional encoding of xyz, fused (never materialised)
            float x0, x1, x2;
            if (q.mode == ASDF_QUERY_POINTS) {
              const float* row = q.points_dev + (size_t)i * q.point_stride;
              x0 = __ldg(row); x1 = __ldg(row + 1); x2 = __ldg(row + 2);
            } else {
              grid_point(i, q.N, q.mode, q.voxel, q.origin[0], q.origin[1], q.origin[2], x0, x1, x2);
            }
            for (int dd = 0; dd < D; ++dd) u[dd * PTS + p] = nerf_feature(dd, x0, x1, x2);
          } else if (q.mode == ASDF_QUERY_POINTS) {
            const float* row = q.points_dev + (size_t)i * q.point_stride;
            for (int dd = 0; dd < D; ++dd) u[dd * PTS + p] = __ldg(row + d.point_index[b][dd]);
          } else {
            float x0, x1, x2;
            grid_point(i, q.N, q.mode, q.voxel, q.origin[0], q.origin[1], q.origin[2], x0, x1, x2);
            u[0 * PTS + p] = x0; u[1 * PTS + p] = x1; u[2 * PTS + p] = x2;
          }
        } else {
          for (int dd = 0; dd < D; ++dd) u[dd * PTS + p] = 0.f;
        }
      }
      __syncthreads();
      float* in = act0;
      float* out = act1;
      for (int l = 0; l < L - 1; ++l) {
        const int32_t* t = d.table[b][l];
        const float* pa_map = nullptr;
        if (d.pa.enabled && (l == d.pa.layer[0] || l == d.pa.layer[1]))
          pa_map = d.pa.maps_dev + (size_t)(b * 2 + (l == d.pa.layer[0] ? 0 : 1)) * d.pa.slot_stride;
        hidden_layer(a.stat + t[4], a.samp + t[5], t[0], t[1], t[2], t[3], D, in, u, out, t[6] >= 0, pa_map, tap_i, tap_w);
        __syncthreads();
        if (t[6] >= 0) {
          layer_norm_relu(out, t[1], a.stat + t[6], a.stat + t[6] + t[1]);
          __syncthreads();
        }
        float* tmp = in; in = out; out = tmp;
      }
      const int32_t* t = d.table[b][L - 1];
      small_head(a.stat + t[4], t[2], a.samp + t[5], t[0], t[1], t[3], D, in, u, res);
      if (d.n_class > 0 && b == 0) {
        __syncthreads();
        if (threadIdx.x < PT) {     // stash the sdf head before res is reused for the logits
          for (int o = 0; o < t[1]; ++o) sdf[o * PT + threadIdx.x] = res[o * PT + threadIdx.x];
        }
        __syncthreads();
        // classifier weights: [n_class][h+1] row-major -> treat as ldw=1 per class
        for (int c = 0; c < d.n_class; ++c)
          small_head(a.cls + (size_t)c * (t[0] + 1), 1, nullptr, t[0], 1, 0, 0, in, u, res + c * PT);
        __syncthreads();
        if (threadIdx.x < PT) {
          int best = 0;
          float bv = res[threadIdx.x] + __ldg(a.cls + t[0]);
          const int64_t ip = base + threadIdx.x;
          float* lg = (a.out_logits && ip < q.end) ? a.out_logits + (size_t)(ip - q.begin) * d.n_class : nullptr;
          if (lg) lg[0] = bv;
          for (int c = 1; c < d.n_class; ++c) {
            const float v = res[c * PT + threadIdx.x] + __ldg(a.cls + (size_t)c * (t[0] + 1) + t[0]);
            if (lg) lg[c] = v;
            if (v > bv) { bv = v; best = c; }
          }
          cls_s[threadIdx.x] = best;
          for (int o = 0; o < t[1]; ++o) res[o * PT + threadIdx.x] = sdf[o * PT + threadIdx.x];
        }
      }
      __syncthreads();
      if (threadIdx.x < PT) {
        const int p = threadIdx.x;
        for (int o = 0; o < t[1]; ++o) {
          float v = res[o * PT + p];
          if (d.pre_tanh) v = tanhf(v);
          v = tanhf(v);
          sdf[(d.n_branches == 2 ? b : o) * PT + p] = v;
        }
      }
    }
    __syncthreads();
    if (threadIdx.x < PT) {
      const int p = threadIdx.x;
      const int64_t i = base + p;
      const bool live = i < q.end;
      const bool two = d.n_branches == 2 || d.n_outputs == 2;
      float vh = 1.f, vo = 1.f;
      if (live) {
        vh = sdf[p];
        a.out_hand[i - q.begin] = vh;
        if (two) { vo = sdf[PT + p]; if (a.out_obj) a.out_obj[i - q.begin] = vo; }
        if (a.out_cls) a.out_cls[i - q.begin] = d.n_class > 0 ? cls_s[p] : 0;
      }
      if (a.bbox && q.mode != ASDF_QUERY_POINTS) {
        if (q.bbox_mask & 1) bbox_update(a.bbox, live && vh < 0.f, i, q.N);
        if (q.bbox_mask & 2) bbox_update(a.bbox + 6, live && two && vo < 0.f, i, q.N);
      }
    }
  }
}

}  // namespace
}  // namespace asdf

extern "C" int asdf_simt_eval(const asdf_simt_desc* desc, const float* static_dev,
                              const float* sample_dev, const float* cls_dev, const asdf_query* q,
                              float* out_hand_dev, float* out_obj_dev, int32_t* out_cls_dev,
                              float* out_logits_dev, int32_t* bbox_dev, void* stream) {
  using namespace asdf;
  ASDF_REQUIRE(desc && q && static_dev && sample_dev && out_hand_dev, "asdf_simt_eval: null argument");
  ASDF_REQUIRE(desc->n_branches == 1 || desc->n_branches == 2, "n_branches must be 1 or 2");
  ASDF_REQUIRE(desc->n_layers >= 2 && desc->n_layers <= ASDF_MAX_LAYERS, "bad n_layers %d", desc->n_layers);
  ASDF_REQUIRE(desc->n_outputs >= 1 && desc->n_outputs <= 2, "n_outputs must be 1 or 2");
  ASDF_REQUIRE(desc->n_class >= 0 && desc->n_class <= 8, "n_class must be in [0,8]");
  ASDF_REQUIRE(desc->n_class == 0 || cls_dev, "classifier weights missing");
  ASDF_REQUIRE(desc->nerf_freqs >= 0 && 3 + 6 * desc->nerf_freqs <= ASDF_MAX_POINT_DIM, "bad nerf_freqs");
  if (desc->nerf_freqs > 0)
    for (int b = 0; b < desc->n_branches; ++b)
      ASDF_REQUIRE(desc->point_dim[b] == 3 + 6 * desc->nerf_freqs, "point_dim must be 3 + 6 nerf_freqs");
  ASDF_REQUIRE(q->end >= q->begin, "empty or negative query range");
  for (int b = 0; b < desc->n_branches; ++b) {
    ASDF_REQUIRE(desc->point_dim[b] >= 1 && desc->point_dim[b] <= ASDF_MAX_POINT_DIM, "bad point_dim");
    for (int l = 0; l < desc->n_layers; ++l) {
      const int32_t* t = desc->table[b][l];
      ASDF_REQUIRE(t[0] >= 0 && t[0] <= MAXW && t[2] >= t[1] && t[2] <= MAXW && t[2] % 8 == 0,
                   "layer %d of branch %d: widths (h=%d n=%d npad=%d) outside the supported range", l, b, t[0], t[1], t[2]);
      if (l < desc->n_layers - 1) ASDF_REQUIRE(l == 0 ? t[0] == 0 : t[0] > 0, "layer %d: bad input width", l);
    }
    ASDF_REQUIRE(desc->table[b][desc->n_layers - 1][1] == desc->n_outputs, "last layer width != n_outputs");
  }
  if (q->mode == ASDF_QUERY_POINTS) {
    ASDF_REQUIRE(q->points_dev && q->point_stride >= (desc->nerf_freqs > 0 ? 3 : 1), "points query without points");
  } else {
    ASDF_REQUIRE(q->mode == ASDF_QUERY_GRID_REFERENCE || q->mode == ASDF_QUERY_GRID_REGULAR, "bad query mode");
    ASDF_REQUIRE(q->N >= 2 && q->end <= (int64_t)q->N * q->N * q->N && q->begin >= 0, "grid range outside N^3");
    ASDF_REQUIRE(desc->nerf_freqs > 0 || (desc->point_dim[0] == 3 && (desc->n_branches == 1 || desc->point_dim[1] == 3)),
                 "grid queries need xyz-folded weights (point_dim 3) or the in-kernel NeRF encoding");
  }
  if (q->end == q->begin) return ASDF_OK;
  SimtArgs a;
  a.d = *desc; a.q = *q; a.stat = static_dev; a.samp = sample_dev; a.cls = cls_dev;
  a.out_hand = out_hand_dev; a.out_obj = out_obj_dev; a.out_cls = out_cls_dev; a.out_logits = out_logits_dev; a.bbox = bbox_dev;
  if (desc->pa.enabled) {
    ASDF_REQUIRE(desc->pa.maps_dev && desc->pa.fh >= 1 && desc->pa.fw >= 1, "PixelAlign: missing feature maps");
    ASDF_REQUIRE(desc->pa.slot_stride >= (int64_t)(desc->pa.fh * desc->pa.fw + 1) * 8, "PixelAlign: bad slot_stride");
    ASDF_REQUIRE(q->mode != ASDF_QUERY_POINTS || q->point_stride >= 3, "PixelAlign: point rows need >= 3 columns");
  }
  const size_t smem = (size_t)(2 * MAXW * PTS + ASDF_MAX_POINT_DIM * PTS + 8 * PT + 2 * PT + PT + 32 * PT) * sizeof(float);
  int dev = 0, sms = 0;
  ASDF_CUDA_CHECK(cudaGetDevice(&dev));
  ASDF_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t n_tiles = (q->end - q->begin + PT - 1) / PT;
  const int grid = (int)(n_tiles < sms ? n_tiles : sms);
  ASDF_CUDA_CHECK(cudaFuncSetAttribute(simt_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  simt_eval_kernel<<<grid, NT, smem, (cudaStream_t)stream>>>(a);
  ASDF_CUDA_CHECK(cudaGetLastError());
  return ASDF_OK;
}

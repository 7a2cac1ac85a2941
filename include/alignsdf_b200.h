/* alignsdf_b200 -- C ABI of the B200-native AlignSDF reconstruction hot path.
 *
 * The reference has no FFI layer: its hot path is plain Python calling ATen
 * (SURVEY.md §8b).  This header is the boundary a maintainer binds instead
 * (ctypes stub in INTEGRATION.md); every entry point names the reference code it
 * replaces (paths relative to the reference repository root).
 *
 * Conventions
 *  - every pointer named *_dev is a DEVICE pointer owned by the caller (a torch
 *    tensor); the library never frees or retains it beyond the call;
 *  - `stream` is a cudaStream_t passed as void*; all calls are asynchronous on it;
 *  - return value 0 = OK, negative = error; asdf_last_error() returns the message
 *    of the last failure on the calling thread;
 *  - no hidden global state; no CPU fallback: if no CUDA device is usable the
 *    calls fail with ASDF_ERR_CUDA.
 */
#ifndef ALIGNSDF_B200_H_
#define ALIGNSDF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ASDF_ABI_VERSION 4
#define ASDF_MAX_LAYERS 8
#define ASDF_MAX_POINT_DIM 64

#define ASDF_OK 0
#define ASDF_ERR_ARG (-1)
#define ASDF_ERR_CUDA (-2)
#define ASDF_ERR_UNSUPPORTED (-3)

/* Query set.  Replaces the CPU grid construction + per-chunk H2D copy of
 * utils/mesh.py:24-48,82-96 (deep_sdf/mesh.py:21-46). */
#define ASDF_QUERY_GRID_REFERENCE 0 /* sheared lattice the reference really evaluates (true division) */
#define ASDF_QUERY_GRID_REGULAR 1   /* floor-division lattice */
#define ASDF_QUERY_POINTS 2         /* explicit rows of `point_stride` floats */
typedef struct {
  int32_t mode;
  int32_t N;             /* grid resolution (grid modes) */
  int64_t begin, end;    /* linear index range [begin,end) of the N^3 grid, or [0,P) for points */
  float voxel;           /* grid spacing  */
  float origin[3];       /* column k of the query uses origin[k] */
  const float* points_dev; /* ASDF_QUERY_POINTS: [P, point_stride] */
  int32_t point_stride;
  int32_t bbox_mask;     /* bit0: hand field feeds the bbox, bit1: object field */
} asdf_query;

/* Generic fp32 decoder description (any widths <= 512, any skip layout).
 * Replaces networks/model.py:79-188 (CombinedDecoder.forward) and :285-350
 * (SeparateDecoder.forward) together with utils/utils.py:376-430 (kinematic_embedding)
 * and :561-572 (decode_sdf_multi_output) after host-side folding (alignsdf_b200/packer.py). */
/* PixelAlign (specs['PixelAlign'], utils/utils.py:536-558,563-566): the latent of a query point is the bicubic
 * sample (grid_sample, align_corners=True, zero padding; the mean feature for points that project outside the image)
 * of an image feature map at the point's projection.  Sampling is linear in the map, so the latent columns of the
 * two layers that see the latent (layer 0 and the skip layer) are applied to the MAP once per sample:
 *   maps_dev[branch][slot][pixel][npad]   pixel < fh*fw: (W_z F)[:, pixel];  pixel == fh*fw: W_z mean(F)
 * and the kernel adds the 16-tap bicubic combination of those rows to the layer's pre-activations.
 * xyz_cam = point_affine[3x4] . [x; 1] with x = the query's first three features (POINTS mode: columns 0..2 of the
 * row, like queries[:, :3] at utils/utils.py:564) or the grid point (grid modes; the pose-align affine of the first
 * three features folded in by the host);  [px py pz] = cam[3x4] . [xyz_cam; 1];  uv = (px, py) / pz / image_size * 2 - 1. */
typedef struct {
  int32_t enabled;
  int32_t fh, fw;                           /* feature map height / width */
  int32_t layer[2];                         /* layers receiving the latent (slot 0, 1); -1 = unused slot */
  float point_affine[12];
  float cam[12];
  float image_size;
  int32_t reserved;
  int64_t slot_stride;                      /* floats between consecutive [pixel][npad] maps: (fh*fw + 1) * npad */
  const float* maps_dev;
} asdf_pixel_align;

typedef struct {
  int32_t n_branches;                       /* 2: separate hand/object MLPs, 1: one MLP with n_outputs */
  int32_t n_layers;                         /* linear layers per branch */
  int32_t n_outputs;                        /* width of the last layer (1 or 2) */
  int32_t pre_tanh;                         /* NetworkSpecs.use_tanh */
  int32_t n_class;                          /* classifier logits on the penultimate activations, 0 = none */
  int32_t nerf_freqs;                       /* > 0: u = NeRF positional encoding of the query's xyz, computed in the
                                               kernel: [x, sin(2^f x), cos(2^f x)] for f < nerf_freqs (utils/utils.py:
                                               433-463,521-533; utils/mesh.py:54-55); point_dim must be 3 + 6 nerf_freqs */
  int32_t point_dim[2];                     /* D per branch: 3 (xyz) or the branch's feature count */
  int32_t point_index[2][ASDF_MAX_POINT_DIM]; /* column of the query row feeding u[d] */
  int32_t table[2][ASDF_MAX_LAYERS][8];     /* h, n, npad, has_M, off_static, off_sample (in floats),
                                               off_layernorm (gamma[n] | beta[n] in static_dev, -1 = none), 0 */
  asdf_pixel_align pa;                      /* pa.enabled = 0: the latent is folded into the sample block */
} asdf_simt_desc;

/* y_l = relu(LN_l(WxT_l^T x + M_l u + B_l)) ... tanh  (LN_l = LayerNorm, eps 1e-5, only where off_layernorm >= 0:
 * networks/model.py:254-255,317-319);   static_dev: WxT blocks, sample_dev: [npad][D+1] blocks,
 * cls_dev: [n_class][h_last+1] (weights then bias) or NULL.
 * out_hand_dev / out_obj_dev: [end-begin] f32.  out_cls_dev: [end-begin] int32 argmax or NULL.
 * out_logits_dev: [end-begin][n_class] f32 raw classifier logits (what the reference's decoder returns as its
 * third output, networks/model.py:161-162,188) or NULL.
 * bbox_dev: int32[12] = hand {min0,min1,min2,max0,max1,max2} then object {...}, updated with
 * atomicMin/Max over the unravelled indices of every point whose field is < 0
 * (utils/mesh.py:198-247); the caller initialises it to {INT_MAX x3, -1 x3} x2.  May be NULL. */
int asdf_simt_eval(const asdf_simt_desc* desc, const float* static_dev, const float* sample_dev,
                   const float* cls_dev, const asdf_query* q, float* out_hand_dev, float* out_obj_dev,
                   int32_t* out_cls_dev, float* out_logits_dev, int32_t* bbox_dev, void* stream);

/* Tensor-core path (csrc/k1_tc.cu) for the shipped topology -- 5 linear layers, 512 wide, skip into layer 2:
 * SeparateDecoder (two MLPs, networks/model.py:285-350) or CombinedDecoder without xyz_in_all (one MLP, two outputs,
 * networks/model.py:139-188) -- replacing the hot loops of utils/mesh.py:46-63,96-115 for a BATCH of samples per
 * launch: tcgen05 / TMEM, fp32 accumulation, split precision.
 *   ASDF_TC_F16X3   x.W ~= hi16(x).hi16(W) + lo16(x).hi16(W) + hi16(x).lo16(W): within the 1e-5 contract for any decoder
 *   ASDF_TC_F16_F8  the two correction products in e4m3 (2/3 of the tensor time): only for decoders whose
 *                   calibration run shows them inside the contract (alignsdf_b200/engine.py)
 * static_dev: asdf_tc_static_bytes(kind, n_decoders) bytes (alignsdf_b200/tc_pack.py, one stream per kind);
 * samples_dev: per sample asdf_tc_sample_bytes() bytes written by asdf_tc_bind (sample_stride apart);
 * grid_dev: NULL or float[n_samples][4] = {voxel, origin0, origin1, origin2} overriding q->voxel / q->origin per
 *   sample (e.g. written by asdf_regrid: pass 2 then needs no host round trip);
 * out_*_dev: NULL (bounding-box-only pass: utils/mesh.py:46-80 only uses pass 1 for the box) or [n_samples][out_stride];
 * bbox_dev: NULL or int32[n_samples][12] (initialised by the caller like asdf_simt_eval's);
 * status_dev: int32[1], zeroed by the caller; bit 0 is raised when an activation left the range of the operand
 *   format (e4m3: 448, fp16: 60000 / 16) -- the outputs of that launch must then be discarded and the query re-run
 *   through the next safer kernel (F16_F8 -> F16X3 -> asdf_simt_eval). */
#define ASDF_TC_F16X3 0
#define ASDF_TC_F16_F8 1
/* Bounding-box-only kind for pass 1 (the reference uses that pass for nothing but the box, utils/mesh.py:46-80): the
 * fp16 main product ALONE (static / sample blocks of kind ASDF_TC_F16_F8; half its tensor time).  Its error (~1e-4 x
 * output range) cannot decide the sign of values near zero, so the launch takes a threshold: points with
 * val < -bbox_tau enter the box, points with |val| <= bbox_tau are appended to amb_dev -- (x, y, z, bits = grid index |
 * output << 30) per entry, amb_count_dev[sample] counts the entries WANTED -- and the caller re-evaluates those
 * exactly (ASDF_QUERY_POINTS through an exact kind) and merges their signs into the box: the result equals the
 * exact kind's box whenever bbox_tau bounds this kind's error (the caller calibrates it per sample). */
#define ASDF_TC_F16X1 2
typedef struct {
  int32_t kind;
  int32_t n_decoders;
  int32_t n_samples;
  int32_t reserved;
  const void* static_dev;
  const void* samples_dev;
  int64_t sample_stride;
  const float* grid_dev;
  float* out_hand_dev;
  float* out_obj_dev;
  int64_t out_stride;
  int32_t* bbox_dev;
  int32_t* status_dev;
  float bbox_tau;           /* with amb_dev: sign threshold of the bounding boxes (see ASDF_TC_F16X1); any kind */
  int32_t amb_capacity;     /* entries per sample */
  void* amb_dev;            /* NULL or float4[n_samples][amb_capacity] */
  int32_t* amb_count_dev;   /* int32[n_samples], zeroed by the caller */
} asdf_tc_launch;
int asdf_tc_eval(const asdf_tc_launch* l, const asdf_query* q, void* stream);
int64_t asdf_tc_static_bytes(int32_t kind, int32_t n_decoders);
int64_t asdf_tc_sample_bytes(void);

/* Per-sample set-up of the tensor-core path on the device (csrc/bind.cu).  Replaces what the reference recomputes
 * for every query point: the latent columns of the first / skip layer (utils/utils.py:561-572 cat + networks/model.py:
 * 304-311 lin0 / lin2) fold into biases, the pose-align features (utils/utils.py:376-430) into [512,3] point matrices
 * (float64), which are then scaled, split into fp16 hi + lo and written as the "P tiles" asdf_tc_eval streams.
 * static_dev: per decoder (decoder_stride doubles apart)  Wz[2][512][latent_size] | Wf[2][512][ASDF_MAX_POINT_DIM] |
 *   b[4][512]  (latent / feature columns of layers 0 and 2; biases of layers 0..3, b1 zero padded);
 * latent_dev: float[n_samples][latent_size];  affine_dev: double[n_samples][ASDF_MAX_POINT_DIM][4] = rows (A, c) of
 *   features = A.xyz + c;  feature_index[d][f]: which affine row feeds feature f of decoder d (networks/model.py:288-299);
 * fold_scratch_dev: double[n_samples][n_decoders][2][512][4];  samples_dev: n_samples blocks of
 *   asdf_tc_sample_bytes() bytes, zero-initialised once by the caller;
 * status_dev: bit 1 is raised when the operands of a sample do not fit fp16 (caller falls back to asdf_simt_eval). */
typedef struct {
  int32_t n_decoders;
  int32_t latent_size;
  int32_t n_features[2];
  int32_t feature_index[2][ASDF_MAX_POINT_DIM];
  int64_t decoder_stride;
  float act_scale;       /* t: 16 for ASDF_TC_F16X3, 1 for ASDF_TC_F16_F8 */
  float p_absmax;        /* bound on |xyz| of the queries the block will be used for */
  double w_scale[2][3];  /* power-of-two scales of the packed weights of layers 1..3 (tc_pack.py) */
} asdf_tc_bind_desc;
int asdf_tc_bind(const asdf_tc_bind_desc* desc, const double* static_dev, const float* latent_dev,
                 const double* affine_dev, int32_t n_samples, double* fold_scratch_dev, void* samples_dev,
                 int64_t sample_stride, int32_t* status_dev, void* stream);
int64_t asdf_tc_bind_static_doubles(int32_t n_decoders, int32_t latent_size);

/* Bounding boxes of pass 1 -> lattice of pass 2 (utils/mesh.py:198-256 get_higher_res_cube, same f32 arithmetic):
 * bbox_dev int32[n_samples][12], branch_mask bit0 hand / bit1 object, voxel = spacing of pass 1 (origin -1);
 * grid_dev float[n_samples][4] = {new_voxel, new_origin[3]}, minmax_dev NULL or float[n_samples][6]. */
int asdf_regrid(const int32_t* bbox_dev, int32_t n_samples, int32_t branch_mask, int32_t N, float voxel,
                float* grid_dev, float* minmax_dev, void* stream);

/* Query coordinates only (tests / debugging): xyz_dev [end-begin,3], bit-exact w.r.t. the
 * reference's torch expressions at utils/mesh.py:32-40,86-94. */
int asdf_grid_points(const asdf_query* q, float* xyz_dev, void* stream);

/* NeRF positional encoding for the public get_nerf_embedder() API (utils/utils.py:521-533):
 * feats[P, 3 + 6 n_freqs] = [x, sin(x 2^0), cos(x 2^0), sin(x 2^1), ...]. */
int asdf_nerf_embed(const float* xyz_dev, int64_t P, int32_t n_freqs, float* feats_dev, void* stream);

/* Pose-align embedding as an affine map: feats[P,pf] = xyz[P,3] A^T + c.
 * Replaces utils/utils.py:376-430 for the public kinematic_embedding() API. */
int asdf_embed_points(const float* xyz_dev, int64_t P, const float* affine_dev /* [pf][4] */,
                      int32_t pf, float* feats_dev, void* stream);

/* Marching cubes over a [n0,n1,n2] f32 field.  Replaces the
 * skimage.measure.marching_cubes_lewiner call at utils/mesh.py:354 / deep_sdf/mesh.py:81
 * plus the vertex affine of utils/mesh.py:360-369. */
typedef struct {
  int32_t n0, n1, n2;      /* local volume (a z-slab including its halo plane) */
  int32_t full1, full2;    /* extents of axes 1,2 of the full grid (== n1,n2) */
  int64_t index0_offset;   /* global axis-0 index of local plane 0 (vertex keys and coordinates) */
  float iso;
  double spacing[3];
  float origin[3];         /* points = origin + verts (utils/mesh.py:360-363) */
  const float* grid_dev;   /* optional: f32[4] = {voxel, origin x3} on the device (a row of asdf_regrid's output); when
                            * non-NULL it replaces spacing (all three axes) and origin, so that marching cubes can be
                            * queued behind the grid passes without a host round trip */
} asdf_mc_params;
size_t asdf_mc_scratch_bytes(const asdf_mc_params* p);
/* Pass 1: classify + count + scan (the field is read once).  totals_dev: int64[5] = {n_verts, n_tris, 0, 0,
 * n_segments} (n_segments = runs of 32 grid points that own a vertex or a triangle, handed back to asdf_mc_emit).
 * n_verts == 0 means the field never crosses iso; whether iso lies outside [min, max] (skimage's ValueError, caught
 * at utils/mesh.py:353-358) is then the caller's reduction -- the streaming pass carries none. */
int asdf_mc_count(const float* vol_dev, const asdf_mc_params* p, void* scratch_dev,
                  int64_t* totals_dev, void* stream);
/* Pass 2: emit.  verts_dev [V,3] f32 (array-axis order * spacing, what marching_cubes returns),
 * points_dev [V,3] f32 (origin + verts) or NULL, faces_dev [F,3] int32,
 * keys_dev [V] uint64 global vertex keys (for slab stitching) or NULL. */
int asdf_mc_emit(const float* vol_dev, const asdf_mc_params* p, const void* scratch_dev, int64_t n_segments,
                 float* verts_dev, float* points_dev, int32_t* faces_dev, uint64_t* keys_dev,
                 void* stream);

/* Connected components of a marching-cubes mesh and per-component statistics (K3).  Replaces
 * trimesh.graph.split + the largest-area selection of utils/mesh.py:371-381.
 * asdf_cc_label: parent_dev[V] <- smallest vertex index of the vertex's component (faces [F,3] int32).
 * asdf_cc_stats: per-root accumulators (area_dev f64[V], nfaces_dev/open_dev i32[V] zeroed by the
 *   caller, first_face_dev i32[V] initialised to INT_MAX): area on points_dev, number of faces,
 *   open = some triangle edge lies in a boundary plane of the volume (verts_local_dev coordinate == 0
 *   or == last_plane[k]), index of the component's first face.
 * asdf_cc_mark / asdf_cc_gather: compaction of the component `best_label` around the caller's
 *   inclusive prefix sums of keep_v / keep_f (vertex and face order are preserved). */
int asdf_cc_label(const int32_t* faces_dev, int64_t F, int64_t V, int32_t* parent_dev, void* stream);
int asdf_cc_stats(const int32_t* faces_dev, int64_t F, const float* verts_local_dev, const float* points_dev,
                  int64_t V, const int32_t* parent_dev, const float last_plane[3], double* area_dev,
                  int32_t* nfaces_dev, int32_t* open_dev, int32_t* first_face_dev, void* stream);
int asdf_cc_mark(const int32_t* parent_dev, int64_t V, const int32_t* faces_dev, int64_t F, int32_t best_label,
                 int32_t* keep_v_dev, int32_t* keep_f_dev, void* stream);
int asdf_cc_gather(const float* points_dev, const int32_t* faces_dev, int64_t V, int64_t F,
                   const int32_t* keep_v_dev, const int32_t* scan_v_dev, const int32_t* keep_f_dev,
                   const int32_t* scan_f_dev, float* out_points_dev, int32_t* out_faces_dev, void* stream);

/* Exact nearest neighbour (float64, brute force): for every query point the index of the closest reference point
 * (ties: smallest index) and, when dist2_dev != NULL, the squared distance.  Replaces the KDTree queries of
 * deep_sdf/metrics/icp_trans_scale.py:36-57 (ICP scale / translation alignment, what --eval_mode runs after every
 * mesh, utils/mesh.py:385-395) and deep_sdf/metrics/chamfer.py:217-229 (symmetric Chamfer distance).
 * query_dev [n_query,3], ref_dev [n_ref,3] float64; idx_dev int32[n_query]; dist2_dev float64[n_query] or NULL. */
int asdf_nn_search(const double* query_dev, int64_t n_query, const double* ref_dev, int64_t n_ref,
                   int32_t* idx_dev, double* dist2_dev, void* stream);

int asdf_abi_version(void);
const char* asdf_last_error(void);
/* 1 if a CUDA device with compute capability 10.x is present, else 0 (never falls back). */
int asdf_device_ok(void);

#ifdef __cplusplus
}
#endif
#endif /* ALIGNSDF_B200_H_ */

/*
 * durf_b200.h -- C ABI of libdurf_b200.so: the B200 (sm_100a) implementation of DURF's per-ray
 * Mip-NeRF hot path (FelTris/durf: internal/mip.py, mip360.py, math.py, box_helpers.py,
 * obbpose_model.py and the loss block of train_boxpose.py).
 *
 * The reference has no plugin/FFI interface of its own (it is pure Python/JAX); the seam this
 * library replaces is the Python call surface of MipNerfModel.__call__ (obbpose_model.py:105-254).
 * Every entry point below names the reference function(s) it replaces (file:line).
 *
 * Conventions
 *   - All pointers are DEVICE pointers unless the name ends in _host.  float = IEEE fp32, row-major.
 *   - The caller owns every buffer: inputs are const, outputs and workspaces are pre-allocated by the
 *     caller.  The library never allocates device memory, never synchronises and never creates
 *     streams: it only enqueues kernels on `stream`.  Calls are re-entrant (one host thread per GPU)
 *     and CUDA-graph capturable.
 *   - Return value: DURF_OK (0) or a negative DURF_E_* code; durf_last_error() returns a thread-local
 *     message for the last failing call.
 *   - `stream` is a cudaStream_t passed as void* (so that this header needs no CUDA headers).
 *   - Nullable arguments are marked [opt].
 */
#ifndef DURF_B200_H_
#define DURF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DURF_OK              0
#define DURF_E_INVALID      -1   /* bad argument (null pointer, unsupported shape) */
#define DURF_E_LAUNCH       -2   /* CUDA launch / runtime error */
#define DURF_E_WORKSPACE    -3   /* workspace too small */
#define DURF_E_UNSUPPORTED  -4   /* valid request this build cannot serve (e.g. tensor-core path with N != 128) */

typedef void* durf_stream_t;

/* ---- library ------------------------------------------------------------------------------- */
const char* durf_version(void);
const char* durf_last_error(void);
/* Number of kernels this library has launched from the calling thread since the last reset
 * (bench.py reports it as gpu_launches). */
int64_t durf_launch_count(void);
void    durf_reset_launch_count(void);

/* ---- MLP topology (obbpose_model.py:294-303 MLP, :358-367 BoxMLP) --------------------------- */
typedef struct DurfMlpTopology {
  int32_t in_dim;      /* 60 (IPE) or 63 (mean + weighted IPE) */
  int32_t width;       /* net_width: 256 (MLP) / 128 (BoxMLP) */
  int32_t depth;       /* net_depth: 8 */
  int32_t skip;        /* skip_layer: 4 -> input re-concatenated after layer 4 */
  int32_t cond_dim;    /* 27 = 3 + 2*3*deg_view */
  int32_t cond_width;  /* net_width_condition: 128 */
} DurfMlpTopology;

/* Parameter blob layout (fp32): for i = 0 .. depth+3 in flax creation order
 * (Dense_0..Dense_{depth-1} trunk, Dense_depth density, +1 bottleneck, +2 condition, +3 rgb):
 *   kernel_i [in_i, out_i] row-major, then bias_i [out_i].                                     */
int64_t durf_mlp_param_count(const DurfMlpTopology* topo);
/* Offset (in floats) of kernel_i inside the blob; the bias follows the kernel. */
int64_t durf_mlp_param_offset(const DurfMlpTopology* topo, int32_t layer, int32_t* in_dim, int32_t* out_dim);

#define DURF_PREC_FP32 0   /* CUDA-core fp32 GEMMs: parity mode (<=1e-5 vs the fp32 oracle) */
#define DURF_PREC_BF16 1   /* tcgen05 bf16 x bf16 -> fp32 (TMEM accumulators), fused layer chain */

/* ---- K0: OBB front-end ------------------------------------------------------------------- */
/* box_helpers.aa2matrix (box_helpers.py:148-167): angles [K,3] -> R [K,3,3]. */
int durf_aa2matrix_fwd(durf_stream_t stream, int32_t K, const float* angles, float* R);

/* box_helpers.world2object_rpy(pts, dirs, pose, rot) (box_helpers.py:286-341; dim=None, inverse=False) with explicit
 * rotation matrices: pts, dirs [B,3]; pose [K,3] or [B,K,3] (pose_per_ray); rot [K,3,3] or [B,K,3,3] (rot_per_ray)
 * -> pts_o, dirs_o [B,K,3] (dirs_o unit length, box_helpers.py:340). */
int durf_world2object_fwd(durf_stream_t stream, int32_t B, int32_t K, const float* pts, const float* dirs,
                          const float* pose, int32_t pose_per_ray, const float* rot, int32_t rot_per_ray,
                          float* pts_o, float* dirs_o);

/* box_helpers.ray_box_intersection(ray_o, ray_d, aabb_min, aabb_max) (box_helpers.py:59-106) on n (ray, box) pairs:
 * ray_o, ray_d, aabb_min, aabb_max [n,3] (NULL bounds = the unit box) -> z_in, z_out [n], intersection [n] int32. */
int durf_ray_box_intersection_fwd(durf_stream_t stream, int64_t n, const float* ray_o, const float* ray_d,
                                  const float* aabb_min, const float* aabb_max, float* z_in, float* z_out,
                                  int32_t* intersection);

/* box_helpers.world2object_rpy (box_helpers.py:286-341) + ray_box_intersection (:59-106) +
 * the scene-graph merge of MipNerfModel.__call__ (obbpose_model.py:99-131).
 *   origins, directions [B,3]; box [K,6] = box_centers[ts] (xyz + axis-angle); ext [K,3] half-extents.
 * Outputs (all caller-allocated):
 *   origins_s, dirs_s [B,3]; hit [B,K] int32; zi, zo [B,K]; zo_ret [B]; nhit [B] (= sum_k hit, float);
 *   [opt] origins_o, dirs_o [B,K,3] (object-frame rays, only if non-null).                      */
int durf_obb_frontend_fwd(durf_stream_t stream, int32_t B, int32_t K,
                          const float* origins, const float* directions,
                          const float* box, const float* ext,
                          float* origins_s, float* dirs_s, int32_t* hit, float* zi, float* zo,
                          float* zo_ret, float* nhit, float* origins_o, float* dirs_o);

/* Backward of the front-end for the joint box-pose optimisation (obbpose_model.py:99-122 under
 * jax.value_and_grad, train_boxpose.py:251): given dL/d origins_s and dL/d dirs_s [B,3] accumulate
 * dL/d box [K,6] (fp32 atomics into d_box, which the caller zeroes).  `hit` is the stop_gradient'ed mask. */
int durf_obb_frontend_bwd(durf_stream_t stream, int32_t B, int32_t K,
                          const float* origins, const float* directions, const float* box,
                          const int32_t* hit, const float* d_origins_s, const float* d_dirs_s,
                          int32_t pose_grad, int32_t rot_grad, float* d_box);

/* raw_rgb[ray_index[m]] += src_rgb[m], raw_density[ray_index[m]] += src_density[m] for m < *count (M when count is NULL):
 * the second half of `raw += mask_k * BoxMLP_k(...)` (obbpose_model.py:203-204, 233-234) for a network evaluated with
 * DurfMlpArgs.accumulate == 2.  Merging the objects in index order gives the sums of accumulate == 1 bit for bit. */
int durf_mlp_merge_raw(durf_stream_t stream, int32_t M, int32_t N, const int32_t* ray_index, const int32_t* count,
                       const float* src_rgb, const float* src_density, float* raw_rgb, float* raw_density);

/* Compaction of the rays that hit object k (the reference evaluates every BoxMLP on every ray and
 * multiplies by the 0/1 mask, obbpose_model.py:174-201; evaluating only hit rays is result-identical).
 * ray_index [B] receives the indices (in NO particular order: warp-aggregated atomics; every consumer scatters its
 * results back per ray) and *count their number. */
int durf_compact_hits(durf_stream_t stream, int32_t B, int32_t K, int32_t k, const int32_t* hit,
                      int32_t* ray_index, int32_t* count);
/* All K objects in one launch: ray_index [K,B] (row k = the list of object k), count [K]. */
int durf_compact_hits_all(durf_stream_t stream, int32_t B, int32_t K, const int32_t* hit, int32_t* ray_index, int32_t* count);

/* ---- N2: pinhole ray generation ------------------------------------------------------------ */
typedef struct DurfCamera {
  int32_t width, height;   /* pixels */
  float focal;             /* pixels */
  float c2w[12];           /* camera-to-world [3,4], row-major (HOST values) */
  float near, far;
  int32_t use_principal_point; /* 0: image centre (width/2, height/2), the Carla loader (obbpose_dataset.py:627-629);
                                  1: (cx, cy) below, the Waymo loader (obbpose_dataset.py:1863, 1882-1885) */
  float cx, cy;            /* pixels */
} DurfCamera;
/* Carla/Waymo._generate_rays_multi (internal/obbpose_dataset.py:613-661, 1868-1917) for the pixel rows [row0,row1) of one camera,
 * row-major: origins, directions (un-normalised), viewdirs [n,3]; radii, lossmult (=1), near, far [n]; n = (row1-row0)*width.
 * `cam` is a HOST pointer (read during the call). */
int durf_generate_rays(durf_stream_t stream, const DurfCamera* cam, int32_t row0, int32_t row1, float* origins,
                       float* directions, float* viewdirs, float* radii, float* lossmult, float* near, float* far);

/* ---- K1: ray-march (sample -> conical frustum Gaussian -> contraction -> IPE) ------------- */
#define DURF_RM_SAMPLE        (1u << 0)  /* generate t_vals from near/far (mip.sample_along_rays, mip.py:351-368) */
#define DURF_RM_RANDOMIZED    (1u << 1)  /* stratified jitter with the explicit t_rand buffer (mip.py:360-365) */
#define DURF_RM_CONTRACT      (1u << 2)  /* mip360.new_space (mip360.py:63-79) before encoding */
#define DURF_RM_WEIGHTED      (1u << 3)  /* mip.weighted_ipe (mip.py:182-223): [mean, w*enc], 63 features */
#define DURF_RM_CYLINDER      (1u << 4)  /* ray_shape == 'cylinder' (mip.py:133-152) */
#define DURF_RM_NO_INTEGRATE  (1u << 5)  /* disable_integration: zero covariances (obbpose_model.py:164-165) */
#define DURF_RM_OUT_BF16_TILE (1u << 6)  /* features as bf16 128x64 SWIZZLE_128B tile images (input of the tcgen05 MLP) */
#define DURF_RM_MULT_IS_NHIT  (1u << 8)  /* ray_mult holds the front-end's nhit (number of boxes the ray hits): the multiplier is
                                            1 - nhit, the background's `1 - sum_k mask_k` (obbpose_model.py:205) */
#define DURF_RM_NO_TVALS_OUT  (1u << 7)  /* fused_raymarch only, with DURF_RM_SAMPLE: the fenceposts are formed but not stored (t_vals may
                                            be NULL): an object network evaluated next to the background network, which stores them */

typedef struct DurfRaymarchArgs {
  int32_t B;            /* rays in the buffers */
  int32_t N;            /* samples per ray (t_vals has N+1 fenceposts) */
  int32_t min_deg, max_deg;
  uint32_t flags;
  float alpha;          /* BARF coarse-to-fine alpha (weighted_ipe) */
  const float* origins; /* [B,3] origins_s */
  const float* dirs;    /* [B,3] dirs_s */
  const float* radii;   /* [B] */
  const float* near;    /* [B] (DURF_RM_SAMPLE) */
  const float* far;     /* [B] (DURF_RM_SAMPLE) */
  const float* t_rand;  /* [B,N+1] U[0,1) (DURF_RM_RANDOMIZED) */
  float* t_vals;        /* [B,N+1]: written when DURF_RM_SAMPLE, otherwise read */
  const float* ray_mult;   /* [opt] [B] multiplier on mean and cov (hit mask / 1 - sum(hit), obbpose_model.py:179-180, 205-208) */
  const int32_t* ray_index;/* [opt] [M] compacted ray list: output row m <- ray ray_index[m] */
  const int32_t* count;    /* [opt] device count of valid entries in ray_index (else M = B) */
  void* features;       /* fp32 [M,N,F] (F = 60 or 63)  or  bf16 tile images [M*N/128][128x64] */
  float* means;         /* [opt] [M,N,3] cast_rays means (after ray_mult / contraction), parity tests */
  float* cov_diag;      /* [opt] [M,N,3] covariance diagonal fed to the encoding */
  const float* alpha_dev;  /* [opt] device scalar overriding `alpha` (lets a captured CUDA graph follow alpha_rate_fn) */
} DurfRaymarchArgs;

/* mip.sample_along_rays / cast_rays / mip360.new_space / integrated_pos_enc / weighted_ipe
 * (mip.py:330-370, 155-179, 99-130, 226-282, 182-223; mip360.py:47-79). */
int durf_raymarch_fwd(durf_stream_t stream, const DurfRaymarchArgs* args);

/* Backward of the object-frame ray-march (weighted IPE, no contraction) into the ray origin and
 * direction, for the box-pose gradient: d_features fp32 [M,N,63] -> d_origins_s, d_dirs_s [B,3]
 * (written for the rays in ray_index, others untouched). */
int durf_raymarch_bwd(durf_stream_t stream, const DurfRaymarchArgs* args, const float* d_features,
                      float* d_origins_s, float* d_dirs_s);

/* mip.pos_enc(viewdirs, 0, deg, append_identity=True) (mip.py:36-45): [B,3] -> [B, 3+6*deg]. */
int durf_viewdir_enc_fwd(durf_stream_t stream, int32_t B, int32_t deg, const float* viewdirs, float* enc);

/* ---- K2: radiance/density MLP ---------------------------------------------------------------- */
/* Bytes of the tensor-core weight image (bf16, pre-tiled and pre-swizzled per 64x128 chunk). */
int64_t durf_mlp_packed_bytes(const DurfMlpTopology* topo);
/* fp32 parameter blob -> tensor-core weight image.  Call again after every optimizer step. */
int durf_mlp_pack_weights(durf_stream_t stream, const DurfMlpTopology* topo, const float* params, void* packed);
/* The same for n networks (the background MLP and every BoxMLP after an optimizer step) in ONE kernel launch:
 * topos[n] (an array of structs), params[n], packed[n]. */
int durf_mlp_pack_weights_multi(durf_stream_t stream, int32_t n, const DurfMlpTopology* topos, const float* const* params,
                                void* const* packed);

typedef struct DurfMlpArgs {
  DurfMlpTopology topo;
  int32_t precision;        /* DURF_PREC_* */
  int32_t M;                /* rays (rows = M*N) */
  int32_t N;                /* samples per ray */
  const void* features;     /* fp32 [M*N, in_dim] (FP32) or bf16 tile images (BF16) */
  const float* cond;        /* [B,cond_dim] view encoding, indexed by ray (through ray_index when given) */
  const float* params;      /* fp32 blob (always needed: biases, heads) */
  const void* packed;       /* tensor-core weight image (BF16) */
  const int32_t* ray_index; /* [opt] [M] output rows go to ray ray_index[m] */
  const int32_t* count;     /* [opt] device count of valid rays */
  int32_t accumulate;       /* 0: write, 1: add into raw_rgb/raw_density (object MLPs, obbpose_model.py:203-204,233-234),
                               2: write row m of the call (NOT ray ray_index[m]): compact outputs [M,N,3] / [M,N] that
                               durf_mlp_merge_raw adds into the per-ray buffers later, so that the call does not depend on them */
  float* raw_rgb;           /* [B,N,3] */
  float* raw_density;       /* [B,N] */
  void* saved;              /* [opt] kept for the backward pass (durf_mlp_saved_bytes): FP32 the layer outputs; BF16 every
                               layer's bf16 activations as SWIZZLE_128B block images, per layer ordered [sample half][64-column
                               block][64 rows], followed by the trunk layers' 1-bit ReLU masks (an opaque record: only
                               durf_mlp_bwd* read it).  The
                               backward call must pass the same buffer with the same M. */
  void* workspace;
  size_t workspace_bytes;
  const struct DurfRaymarchArgs* fused_raymarch;
                            /* [opt] BF16 only, SURVEY N1: the kernel GENERATES its input tiles (mip.sample_along_rays / cast_rays /
                               mip360.new_space / integrated_pos_enc / weighted_ipe, mip.py:155-282, mip360.py:47-79) from these
                               ray-march arguments instead of reading `features`: the 16 KB/ray-level tile image never touches HBM.
                               B, N = 128, min_deg / max_deg (10 degrees) and the flags are taken from it; ray_index / count are this
                               struct's; t_vals is written when DURF_RM_SAMPLE.  `features` may then be NULL; if it is not, the
                               generated tiles are ALSO stored there (training: the weight-gradient kernel reads them). */
  int32_t saved_tile_offset;  /* [opt] BF16 training: `saved` (and `features`, when the generated tiles are stored) are buffers for  */
  int32_t saved_total_tiles;  /* saved_total_tiles >= M ray-levels and this call fills the records saved_tile_offset .. + M - 1, so
                                 that several forward calls (the levels of obbpose_model.py:133-256) share ONE backward call with
                                 M = saved_total_tiles (one data-gradient and one weight-gradient launch for all levels).
                                 0 / 0 = the buffer is this call's alone.  `features` itself is NOT offset: pass the pointer of
                                 the first tile this call writes. */
} DurfMlpArgs;

size_t durf_mlp_workspace_bytes(const DurfMlpTopology* topo, int32_t precision, int32_t M, int32_t N, int32_t training);
size_t durf_mlp_saved_bytes(const DurfMlpTopology* topo, int32_t precision, int32_t M, int32_t N);

/* MLP.__call__ / BoxMLP.__call__ (obbpose_model.py:305-354, 369-418). */
int durf_mlp_fwd(durf_stream_t stream, const DurfMlpArgs* args);

/* Reverse of durf_mlp_fwd: d_raw_rgb [B,N,3], d_raw_density [B,N] -> d_params (fp32 blob, ACCUMULATED:
 * caller zeroes) and [opt] d_features fp32 [M*N,in_dim] (needed only for the pose gradient). */
int durf_mlp_bwd(durf_stream_t stream, const DurfMlpArgs* args, const float* d_raw_rgb,
                 const float* d_raw_density, float* d_params, float* d_features);

/* The same backward (DURF_PREC_BF16 only) as its two kernels, for callers that run them CONCURRENTLY on two streams:
 * durf_mlp_bwd_data walks the dZ chain (jax.grad through obbpose_model.py:326-353 w.r.t. the activations) and writes
 * every layer's dZ into args->workspace; durf_mlp_bwd_weights forms dW = A^T dZ, db (ACCUMULATED into d_params) from
 * those records.  `tile_done` is an int32 array of durf_mlp_bwd_flags_bytes(topo, M) bytes that the caller ZEROES
 * (stream-ordered before both launches): the data kernel release-increments tile_done[tile][layer] as each dZ block
 * becomes complete in global memory, the weight kernel acquires it before fetching the block, so the dZ records are
 * consumed out of L2 while the chain is still running.  `max_ctas` caps each kernel's grid (0 = every SM): the two
 * caps must add up to at most the SM count, and the data kernel must be able to start (it never waits for the weight
 * kernel; the weight kernel traps, never hangs, if a tile does not arrive within ~2 s).  With tile_done = NULL the
 * weight kernel must be stream-ordered after the data kernel (that is what durf_mlp_bwd does). */
size_t durf_mlp_bwd_flags_bytes(const DurfMlpTopology* topo, int32_t M);
int durf_mlp_bwd_data(durf_stream_t stream, const DurfMlpArgs* args, const float* d_raw_rgb,
                      const float* d_raw_density, float* d_features, int32_t* tile_done, int32_t max_ctas);
int durf_mlp_bwd_weights(durf_stream_t stream, const DurfMlpArgs* args, const float* d_raw_rgb,
                         const float* d_raw_density, float* d_params, const int32_t* tile_done, int32_t max_ctas);

/* ---- K3: activations + alpha compositing ---------------------------------------------------- */
typedef struct DurfCompositeArgs {
  int32_t B, N;
  int32_t white_bkgd, rand_bkgd;
  int32_t activated;           /* 0: inputs are raw (sigmoid / softplus(x + density_bias) applied here); 1: inputs are rgb / density (mip.volumetric_rendering's own signature) */
  float density_bias;          /* -1 (obbpose_model.py:58) */
  /* N = 128: raw_rgb, raw_density, weights, t_mids, t_dists (and the gradient buffers of durf_composite_bwd) are moved with
   * 16-byte loads / stores and must be 16-byte aligned (DURF_E_INVALID otherwise); t_vals rows are 516 bytes, 4-byte aligned. */
  const float* raw_rgb;        /* [B,N,3] summed raw colour (obbpose_model.py:233) */
  const float* raw_density;    /* [B,N]   summed raw density (+ optional noise already added) */
  const float* t_vals;         /* [B,N+1] */
  const float* dirs;           /* [B,3] dirs_s */
  float* comp_rgb;             /* [B,3] */
  float* depth;                /* [B] un-normalised sum(w * t_mid): the model's "distance" (mip.py:317,327) */
  float* acc;                  /* [B] */
  float* weights;              /* [B,N] */
  float* t_mids;               /* [opt] [B,N] */
  float* t_dists;              /* [opt] [B,N] */
} DurfCompositeArgs;

/* rgb/density activations (obbpose_model.py:243-245) + mip.volumetric_rendering (mip.py:285-327). */
int durf_composite_fwd(durf_stream_t stream, const DurfCompositeArgs* args);

/* Backward: d_comp_rgb [B,3], d_depth [B], d_acc [opt][B], d_weights [B,N] ->
 * d_raw_rgb [B,N,3], d_raw_density [B,N]; [opt] d_dirs [B,3] (the analytically-zero pose path via
 * delta = t_dists * |dirs|, mip.py:304). */
int durf_composite_bwd(durf_stream_t stream, const DurfCompositeArgs* args,
                       const float* d_comp_rgb, const float* d_depth, const float* d_acc,
                       const float* d_weights, float* d_raw_rgb, float* d_raw_density, float* d_dirs);

/* ---- K4: hierarchical resampling --------------------------------------------------------------- */
/* mip.resample_along_rays blur-pool (mip.py:393-404) + math.sorted_piecewise_constant_pdf
 * (math.py:222-284).  t_vals [B,N+1] bins, weights [B,N], output [B,num_samples] (the model uses N+1).
 * u_rand [opt] [B,num_samples] U[0,1) => randomized.  blurpool = 0 skips the blur-pool and the padding, which is
 * exactly math.sorted_piecewise_constant_pdf.  No backward: the result is stop_gradient'ed (mip.py:413-414). */
int durf_resample_fwd(durf_stream_t stream, int32_t B, int32_t N, const float* t_vals, const float* weights,
                      const float* u_rand, float resample_padding, int32_t blurpool, int32_t num_samples,
                      float* new_t_vals);

/* ---- KL: losses (train_boxpose.py:94-220) ------------------------------------------------------ */
typedef struct DurfLossArgs {
  int32_t B, N;
  int32_t level;               /* 0 = coarse ... num_levels-1 = fine */
  int32_t num_levels;
  float eps;                   /* line-of-sight half-width, eps_rate_fn (train_boxpose.py:355-361) */
  float coarse_loss_mult, box_loss_mult, depth_loss_mult, near_loss_mult, empty_loss_mult, sky_loss_mult;
  float distortion_mult;       /* 1e-6 hard-coded at train_boxpose.py:220 */
  const float* comp_rgb;       /* [B,3] */
  const float* depth;          /* [B] */
  const float* weights;        /* [B,N] */
  const float* t_vals;         /* [B,N+1] */
  const float* pixels;         /* [B,3] */
  const float* depth_gt;       /* [B] */
  const float* sky;            /* [B] */
  const float* lossmult;       /* [B] */
  const float* dyn_mask;       /* [B] nhit */
  const float* zo;             /* [B] zo_ret */
  float* depth_mask;           /* [B] in/out: accumulates across levels (train_boxpose.py:140); level 0 initialises it */
  float* partials;             /* [num_levels * 8] sums of the level are WRITTEN (not accumulated) -- see DURF_LP_* */
  float* d_comp_rgb;           /* [B,3] gradient of the total loss */
  float* d_depth;              /* [B] */
  float* d_weights;            /* [B,N]; 16-byte aligned rows use vector stores, otherwise scalar stores */
  float* reduce_ws;            /* durf_losses_reduce_ws_floats(B) floats, zero-initialised ONCE by the caller: per-block partial
                                  sums + a ticket; the last block adds them in a fixed order (deterministic loss value) and
                                  re-arms the ticket */
  const float* eps_dev;        /* [opt] device scalar overriding `eps` (lets a captured CUDA graph follow eps_rate_fn) */
} DurfLossArgs;

/* Slots of `partials` (per level: base = level * 8): sums before normalisation. */
#define DURF_LP_RGB    0   /* sum w_rgb (rgb-px)^2 */
#define DURF_LP_DEPTH  1   /* sum m_d (depth-z)^2 */
#define DURF_LP_NEAR   2
#define DURF_LP_EMPTY  3
#define DURF_LP_SKY    4
#define DURF_LP_DISTR  5
#define DURF_LP_OBJ    6   /* sum dyn_mask (rgb-px)^2   (obj_losses numerator, train_boxpose.py:192) */
#define DURF_LP_DYN    7   /* sum dyn_mask              (obj_losses denominator) */
#define DURF_LP_STRIDE 8

/* Pass 1: masks and normalisers (sum lossmult, sum depth_mask, sum sky_mask) into norms[4].
 * Pass 2 (durf_losses_fwd_bwd): loss partial sums + gradients, using the normalisers. */
int64_t durf_losses_reduce_ws_floats(int32_t B);
int durf_losses_prepare(durf_stream_t stream, const DurfLossArgs* args, float* norms);
int durf_losses_fwd_bwd(durf_stream_t stream, const DurfLossArgs* args, const float* norms);

/* The scalar tail of loss_fn (train_boxpose.py:196-220) on the device: per-level losses from the sums and normalisers
 * of the two passes above, and the weighted total.  stats = [num_levels][DURF_LS_STRIDE] {losses, d_losses, n_losses,
 * e_losses, s_losses, distr_losses, obj_losses, tv_losses} followed by {loss, weight_l2}. */
#define DURF_LS_STRIDE 8
typedef struct DurfLossFinalizeArgs {
  int32_t num_levels;
  float coarse_loss_mult, depth_loss_mult, near_loss_mult, empty_loss_mult, sky_loss_mult, tv_loss_mult, distortion_mult;
  const float* partials;       /* [num_levels * DURF_LP_STRIDE] */
  const float* norms;          /* [num_levels * 4] */
  const float* tv;             /* [opt] [num_levels] tv_losses = sum (pose - prev)^2 (train_boxpose.py:136) */
  const float* weight_l2;      /* [opt] scalar (train_boxpose.py:72-74) */
  float* stats;                /* [num_levels * DURF_LS_STRIDE + 2] */
} DurfLossFinalizeArgs;
int durf_losses_finalize(durf_stream_t stream, const DurfLossFinalizeArgs* args);

/* ---- KA: gradient post-processing + Adam (train_boxpose.py:262-288) --------------------------- */
/* nan_to_num(posinf=0) -> clip to +-max_val -> sum of squares into sumsq[0] (caller zeroes). */
int durf_grad_sanitize(durf_stream_t stream, int64_t n, float* grad, float max_val, float grad_scale, float* sumsq);
/* mult = min(1, max_norm / (1e-7 + sqrt(sumsq))) applied on the fly, then flax.optim.Adam.  beta1/beta2/eps are
 * doubles because they are Python floats in the reference: (1. - beta) is formed in double before it meets fp32. */
int durf_adam_step(durf_stream_t stream, int64_t n, float* params, const float* grad, float* m, float* v,
                   const float* sumsq, float max_norm, float lr, double beta1, double beta2, double eps, int32_t step);
/* The same update with the learning rate and the step counter read from DEVICE memory (a captured CUDA graph replays
 * with a new lr / step without re-capture): lr_dev[0] = learning rate, step_dev[0] = 0-based step, incremented by the
 * kernel when `advance_step` != 0.  The bias corrections 1 - beta^t are formed on the device in double like the host. */
int durf_adam_step_dev(durf_stream_t stream, int64_t n, float* params, const float* grad, float* m, float* v,
                       const float* sumsq, float max_norm, const float* lr_dev, int32_t* step_dev, int32_t advance_step,
                       double beta1, double beta2, double eps);

#ifdef __cplusplus
}
#endif
#endif  /* DURF_B200_H_ */

/*
 * dqo_b200.h — C-ABI of the B200-native DQO-MAP hot path (libdqomap_b200.so).
 *
 * Every entry point takes only POD: device pointers, sizes, floats and a CUDA stream handle
 * (cudaStream_t passed as void*).  No torch types, no C++ exceptions, no std::function.
 * All functions are asynchronous with respect to the host (they only enqueue work on `stream`)
 * and return 0 on success or a negative DQO_ERR_* / positive cudaError_t code.
 *
 * Each declaration cites the reference interface it replaces (paths relative to the reference
 * repository root; RAST = submodules/diff-gaussian-rasterizer-depth, KNN = submodules/simple-knn,
 * CU = submodules/cuda_utils).
 */
#ifndef DQO_B200_H_
#define DQO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DQO_OK 0
#define DQO_ERR_INVALID_ARG (-1)
#define DQO_ERR_NO_DEVICE (-2)
#define DQO_ERR_WORKSPACE (-3)

/* ABI version; bumped whenever a signature or struct layout changes. */
int dqo_abi_version(void);
/* Human-readable description of the last error on this thread (never NULL). */
const char *dqo_last_error(void);
/* Instrumentation (bench.py): number of this library's own kernel launches so far (every kernel is this library's: no
 * CUB / Thrust / cuBLAS call anywhere), and optional per-stage CUDA-event timing of the rasterizer on the launching
 * stream, per calling thread.  Stage order: 0 begin-fwd, 1 preprocess, 2 depth sort, 3 (unused: the scan is fused into the
 * emission), 4 scan + emit, 5 tile sort, 6 ranges, 7 front blend and 8 back-phase binning (two-phase binning only),
 * 9 compact, 10 render-fwd, 11 begin-bwd, 12 render-bwd, 13 gaussian-bwd; each value is the time since the previous
 * recorded mark.  dqo_profile_read synchronises on the recorded events. */
long long dqo_launch_count(void);
void dqo_profile_enable(int on);
int dqo_profile_read(float *ms_out, int n);

/* ------------------------------------------------------------------------------------------------
 * Rasterizer settings.  Mirrors GaussianRasterizationSettings
 * (RAST/diff_gaussian_rasterization_depth/__init__.py:288-307) minus the tensors, plus the sizes
 * the pybind layer derives from tensor shapes (RAST/rasterize_points.cu:72-109).
 * ---------------------------------------------------------------------------------------------- */
typedef struct dqo_rast_settings {
    int32_t P;          /* number of Gaussians (means3D.size(0)) */
    int32_t D;          /* active SH degree (sh_degree) */
    int32_t M;          /* SH coefficients per channel stored in `shs` (0 when colors are precomputed) */
    int32_t W, H;       /* image_width, image_height */
    float tanfovx, tanfovy;
    float cx, cy;
    float scale_modifier;
    float color_sigma;        /* default 3.0 */
    float opaque_threshold;
    float depth_threshold;    /* "hit_depth_threshold" of the kernels */
    float normal_threshold;   /* cos of the angle */
    float T_threshold;        /* default 1e-4 */
    int32_t prefiltered;
    int32_t debug;            /* nonzero: synchronise after every stage and report the failing one */
    int32_t need_n_touched;   /* 1 = reference behaviour (count per-Gaussian touches, forward.cu:833-835) */
    /* Occlusion-aware two-phase binning (results identical to single-phase).  0 = single phase: all R
     * instances are binned and sorted, `instance_capacity` must hold R.  > 0: only the nearest Gaussians (in depth-rank
     * order) whose instances fit into `front_instances` are binned first; tiles whose pixels all terminated are done;
     * the remaining Gaussians are then binned only into the unfinished tiles, into `back_instances` slots.
     * front_instances (a multiple of 256) + back_instances <= instance_capacity; DQO_ST_OVERFLOW is raised when the
     * back phase needs more than back_instances.  The same values must be passed to the backward pass. */
    int32_t front_instances;
    int32_t back_instances;
    /* 1: the caller guarantees that `geom_buffer` was initialised once with dqo_rast_geom_init() after its allocation and
     * has since been written by this library only.  The backward pass leaves the per-Gaussian gradient accumulators
     * inside it zeroed (it clears exactly the records it consumed), so with this promise it skips clearing all
     * 128 B x P of them at its start.  0 (default): the accumulators are cleared on every backward call.
     * 2: as 1, and additionally the nine gradient output tensors of dqo_rast_backward are the SAME buffers in every call,
     * zero-initialised by the caller before the first one and not written by anyone else: rows that were zero after the
     * previous call and are zero again (most of a large map) are then not rewritten (~300 B per Gaussian). */
    int32_t geom_clean;
} dqo_rast_settings;

/* status words written on the device by the forward pass (int32[8]) */
#define DQO_ST_NUM_RENDERED 0 /* R: number of (Gaussian, tile) instances (rasterizer_impl.cu:307) */
#define DQO_ST_TILE_NUM 1     /* number of non-empty tiles (rasterizer_impl.cu:365) */
#define DQO_ST_OVERFLOW 2     /* 1 if R exceeded the instance capacity: outputs are invalid, re-run larger */
#define DQO_ST_NUM_VISIBLE 3  /* Gaussians with radii > 0 */
#define DQO_ST_R_FRONT 4      /* two-phase: instances binned in the front phase */
#define DQO_ST_R_BACK 5       /* two-phase: instances needed by the back phase */
#define DQO_ST_WALKED 6       /* list entries staged by the forward blend before early termination (all phases) */
#define DQO_ST_UNFINISHED 7   /* two-phase: rendered tiles still unfinished after the front phase */
#define DQO_ST_WORDS 8

/* Workspace sizing (replaces the three resize callbacks of CudaRasterizer::Rasterizer::forward,
 * RAST/cuda_rasterizer/rasterizer.h:31-33 and rasterizer_impl.h:68-74 `required<T>`). */
size_t dqo_rast_geom_bytes(int32_t P);
/* One-time initialisation of a freshly allocated geometry buffer (see dqo_rast_settings.geom_clean). */
int dqo_rast_geom_init(int32_t P, void *geom_buffer, void *stream);
size_t dqo_rast_binning_bytes(int64_t instance_capacity);
size_t dqo_rast_image_bytes(int32_t W, int32_t H);

/* Forward pass.  Replaces CudaRasterizer::Rasterizer::forward (RAST/cuda_rasterizer/rasterizer.h:30-76,
 * rasterizer_impl.cu:205-441) as called by RasterizeGaussiansCUDA (RAST/rasterize_points.cu:37-155).
 * Optional inputs are NULL when absent (the reference detects empty tensors via null data_ptr).
 * Every output image is fully written by the call (including the fill values of
 * rasterize_points.cu:79-89 for tiles that are not rendered), so callers may pass uninitialised memory.
 * `status` is device int32[DQO_ST_WORDS].  No host synchronisation is performed.  All work is ordered on `stream`;
 * two small stages run on a helper stream owned by the pair (device, `stream`) that is forked from and joined back into
 * `stream` with events before the call returns (capturable in a CUDA graph; callers on different streams share nothing). */
int dqo_rast_forward(const dqo_rast_settings *s,
                     const float *background,     /* [3] */
                     const float *means3D,        /* [P,3] */
                     const float *shs,            /* [P,M,3] or NULL */
                     const float *colors_precomp, /* [P,3] or NULL */
                     const float *opacities,      /* [P] */
                     const float *scales,         /* [P,3] or NULL */
                     const float *rotations,      /* [P,4] or NULL */
                     const float *cov3D_precomp,  /* [P,6] or NULL */
                     const float *viewmatrix,     /* [16] column-major (transposed W2C) */
                     const float *projmatrix,     /* [16] column-major */
                     const float *campos,         /* [3] */
                     const int32_t *tile_mask,    /* [ceil(H/16), ceil(W/16)] */
                     void *geom_buffer, void *binning_buffer, int64_t instance_capacity, void *image_buffer,
                     int32_t *tile_indices,       /* [>= tiles] compact list of non-empty tiles, rest -1 */
                     float *out_color,            /* [3,H,W] */
                     float *out_depth,            /* [H,W] */
                     int32_t *out_hit_depth,      /* [H,W] index of the Gaussian fixing depth, -1 none */
                     int32_t *out_hit_color,      /* [H,W] index of the max-weight Gaussian, -1 none */
                     float *out_hit_color_weight, /* [H,W] */
                     float *out_hit_depth_weight, /* [H,W] */
                     float *out_T,                /* [H,W] */
                     int32_t *radii,              /* [P] */
                     int32_t *n_touched,          /* [P] */
                     int32_t *status, void *stream);

/* Backward pass.  Replaces CudaRasterizer::Rasterizer::backward (rasterizer.h:78-105,
 * rasterizer_impl.cu:445-564) as called by RasterizeGaussiansBackwardCUDA (rasterize_points.cu:157-249).
 * All nine gradient outputs are fully written (no pre-zeroing needed, cf. rasterize_points.cu:198-206).
 * dL_dconic is [P,4] (x, y, unused, w) exactly as the reference's [P,2,2] tensor is used. */
int dqo_rast_backward(const dqo_rast_settings *s, const float *background, const float *means3D, const float *shs,
                      const float *colors_precomp, const float *scales, const float *rotations,
                      const float *cov3D_precomp, const float *viewmatrix, const float *projmatrix,
                      const float *campos, const int32_t *radii, void *geom_buffer /* holds the gradient accumulators */,
                      const void *binning_buffer, int64_t instance_capacity, const void *image_buffer,
                      const int32_t *status,
                      const float *dL_dout_color,   /* [3,H,W] */
                      const float *dL_dout_depth,   /* [H,W] */
                      const int32_t *hit_image,     /* [H,W] = out_hit_depth of the forward */
                      float *dL_dmeans2D,           /* [P,3] */
                      float *dL_dconic,             /* [P,4] */
                      float *dL_dopacity,           /* [P] */
                      float *dL_dcolors,            /* [P,3] */
                      float *dL_dmeans3D,           /* [P,3] */
                      float *dL_dcov3D,             /* [P,6] */
                      float *dL_dsh,                /* [P,M,3] or NULL when M == 0 */
                      float *dL_dscales,            /* [P,3] */
                      float *dL_drotations,         /* [P,4] */
                      void *stream);

/* Frustum test.  Replaces CudaRasterizer::Rasterizer::markVisible (rasterizer.h:23-28,
 * rasterizer_impl.cu:145-157; pybind `mark_visible`, rasterize_points.cu:251-270). */
int dqo_mark_visible(int32_t P, const float *means3D, const float *viewmatrix, const float *projmatrix,
                     uint8_t *present, void *stream);

/* Re-blend of the view binned by the last dqo_rast_forward with another colour per Gaussian.
 * Replaces the second / third full rasterizer call of Renderer.render (SLAM/render.py:227-262, colors_precomp =
 * semantic or instance colours): same lists, same alpha / hit / termination logic, so `out_color` [3,H,W] equals the
 * colour image a full forward with colors_precomp = `colors` [P,3] would return, at the cost of one blend pass.
 * Forward only (the reference consumes these images without gradient: mapper.py:944-969). */
int dqo_rast_blend_extra(const dqo_rast_settings *s, const float *background, const float *colors,
                         const void *geom_buffer, const void *binning_buffer, int64_t instance_capacity,
                         const void *image_buffer, const int32_t *status, float *out_color, void *stream);
/* Backward of dqo_rast_blend_extra (the reference back-propagates through its second full rasterizer call, SLAM/render.py:227-262
 * + rasterizer_impl.cu:445-564): given dL/d(extra image) [3,H,W], writes the gradient w.r.t. the extra colours [P,3] and the
 * extra image's contribution to the gradients of the geometric inputs of the view the workspaces hold (same tensors and
 * meaning as dqo_rast_backward; autograd adds them to the main render's).  `color_acc`: f64[4P] device scratch, zero on entry
 * (left zero).  Uses the per-Gaussian accumulators inside geom_buffer like dqo_rast_backward does. */
int dqo_rast_blend_extra_backward(const dqo_rast_settings *s, const float *background, const float *colors /* [P,3] */,
                                  const float *means3D, const float *scales, const float *rotations,
                                  const float *cov3D_precomp, const float *viewmatrix, const float *projmatrix,
                                  const float *campos, const int32_t *radii, void *geom_buffer, const void *binning_buffer,
                                  int64_t instance_capacity, const void *image_buffer, const int32_t *status,
                                  const float *dL_dextra_image, void *color_acc, float *dL_dcolors, float *dL_dmeans2D,
                                  float *dL_dconic, float *dL_dopacity, float *dL_dmeans3D, float *dL_dcov3D,
                                  float *dL_dscales, float *dL_drotations, void *stream);

/* Introspection for parity tests: re-materialises the reference's binning artefacts from the
 * private workspace in the reference's own formats (rasterizer_impl.h:29-66): 64-bit sorted keys
 * (tile<<32 | depth bits), sorted Gaussian ids, per-tile ranges, and row-major per-pixel / per-Gaussian state.
 * Any output pointer may be NULL.  keys/point_list must hold instance_capacity entries. */
int dqo_rast_export_state(const dqo_rast_settings *s, const void *geom_buffer, const void *binning_buffer,
                          int64_t instance_capacity, const void *image_buffer, const int32_t *status,
                          uint64_t *sorted_keys, uint32_t *point_list, uint32_t *ranges /* [tiles,2] */,
                          uint32_t *n_contrib /* [H,W] */, float *final_T /* [H,W] */,
                          float *means2D /* [P,2] */, float *depths /* [P] */, float *conic_opacity /* [P,4] */,
                          float *rgb /* [P,3] */, uint32_t *tiles_touched /* [P] */, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Stable LSD radix sort of (key, value) pairs with a DEVICE-side item count (csrc/sort.cu).  Replaces the library
 * sorts of the reference's binning and kNN: cub::DeviceRadixSort::SortPairs at RAST/cuda_rasterizer/rasterizer_impl.cu:
 * 327-336 and KNN/simple_knn.cu:241-244.  The rasterizer and dqo_knn3 use it internally; it is exported for tests and
 * for callers that bin by other keys.
 *   n = min(*count, capacity) (count == NULL: capacity); nothing is sorted when *skip != 0 (skip may be NULL).
 *   Bits [0, key_bits) of the keys are sorted, 8 per pass; the data ping-pongs between the (a) and (b) buffers and ends
 *   up in (a) after an even number of passes (key_bits in 9..16 or 25..32), in (b) after an odd one.
 *   implicit_vals != 0: values are the input positions 0..n-1; vals_a is then scratch only (may be NULL for one pass).
 *   temp: dqo_sort_pairs_temp_bytes(capacity, key_bits) bytes of device memory.
 * ---------------------------------------------------------------------------------------------- */
size_t dqo_sort_pairs_temp_bytes(int64_t capacity, int32_t key_bits);
int dqo_sort_pairs_u32(uint32_t *keys_a, uint32_t *keys_b, uint32_t *vals_a, uint32_t *vals_b, int32_t implicit_vals,
                       const int32_t *count, const int32_t *skip, int64_t capacity, int32_t key_bits, void *temp,
                       void *stream);
int dqo_sort_pairs_u16(uint16_t *keys_a, uint16_t *keys_b, uint32_t *vals_a, uint32_t *vals_b, int32_t implicit_vals,
                       const int32_t *count, const int32_t *skip, int64_t capacity, int32_t key_bits, void *temp,
                       void *stream);

/* ------------------------------------------------------------------------------------------------
 * simple-knn.  Replaces SimpleKNN::knn (KNN/simple_knn.cu:216-252) behind distCUDA2 (KNN/spatial.cu:15-28):
 * mean squared distance to the 3 nearest neighbours and their original indices (ascending distance,
 * ties resolved by Morton-sorted position, INT_MAX / FLT_MAX when fewer than 3 neighbours exist).
 * ---------------------------------------------------------------------------------------------- */
size_t dqo_knn_workspace_bytes(int32_t P);
int dqo_knn3(int32_t P, const float *points /* [P,3] */, float *mean_dist2 /* [P] */, int32_t *knn_idx /* [P,3] */,
             void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Map geometry around distCUDA2 (inline PyTorch in the reference).
 *   dqo_bbox_mask       bbox_filter (SLAM/utils.py:801-808): mask[i] = total_xyz[i] strictly inside the bounding box of
 *                       local_xyz grown by `padding`; the box never visits the host.  count (device int, may be NULL) =
 *                       mask.sum().  workspace32: 32 bytes of device memory.
 *   dqo_gaussian_radius GaussianPointCloud.get_radius (SLAM/gaussian_pointcloud.py:739-743): (sum(exp s) - min(exp s)) / 2
 *   dqo_scale_init      GaussianPointCloud.update_geometry after the kNN (gaussian_pointcloud.py:540-569): the first n_new
 *                       rows of xyz_total are the new points, knn_idx [n_new,3] their neighbours in xyz_total (dqo_knn3);
 *                       d_k = |p - p_k| - 3 radius_k, invalid = any d_k < 0 (the rows update_geometry deletes),
 *                       s = clip(sqrt(mean d_k^2), min_radius, max_radius),
 *                       log_scales = log(scale_factor * s * xyz_factor); valid_count (device int, may be NULL) = survivors
 *                       (update_geometry keeps the old scales when it is 0).
 *   dqo_knn_cross3      squared distances (ascending) and indices of the 3 nearest points of `ref` for every point of
 *                       `query`: what temp_points_filter asks pytorch3d.ops.knn_points for (mapper.py:1366-1372, K = 3);
 *                       fewer than 3 reference points: padded with zeros like pytorch3d (index -1 when n_ref == 0);
 *                       equal distances: lower reference index first.
 *   dqo_inside_mask     (sqrt(dist2) < ratio * ref_radius[idx]).any(-1)   (mapper.py:1376-1377, ratio 0.6)
 * ---------------------------------------------------------------------------------------------- */
int dqo_bbox_mask(int32_t n_local, const float *local_xyz, int32_t n_total, const float *total_xyz, float padding,
                  uint8_t *mask /* [n_total] */, int32_t *count, void *workspace32, void *stream);
int dqo_gaussian_radius(int32_t n, const float *log_scales /* [n,3] */, float *radius /* [n] */, void *stream);
int dqo_scale_init(int32_t n_new, int32_t n_total, const float *xyz_total /* [n_total,3] */,
                   const float *radius_total /* [n_total] */, const int32_t *knn_idx /* [n_new,3] */, float min_radius,
                   float max_radius, float scale_factor, float xyz_factor_x, float xyz_factor_y, float xyz_factor_z,
                   float *log_scales /* [n_new,3] */, uint8_t *invalid /* [n_new] */, int32_t *valid_count, void *stream);
size_t dqo_knn_cross3_workspace_bytes(int32_t n_query, int32_t n_ref);
int dqo_knn_cross3(int32_t n_query, const float *query /* [n_query,3] */, int32_t n_ref, const float *ref /* [n_ref,3] */,
                   float *dist2 /* [n_query,3] */, int32_t *idx /* [n_query,3] */, void *workspace, size_t workspace_bytes,
                   void *stream);
int dqo_inside_mask(int32_t n, const float *dist2, const int32_t *idx, const float *ref_radius, float ratio,
                    uint8_t *mask /* [n] */, void *stream);

/* ------------------------------------------------------------------------------------------------
 * cuda_utils.  Replaces accumulate_gaussian_error_impl (CU/map_process.cu:194-245) behind
 * accumulate_gaussian_error (CU/cuda_utils.cu:17-60).  All seven [P] outputs are fully written.
 * ---------------------------------------------------------------------------------------------- */
int dqo_accumulate_error(int32_t W, int32_t H, int32_t P, const float *color_err, const float *depth_err,
                         const float *normal_err, const int32_t *color_index, const int32_t *depth_index,
                         float color_thr, float depth_thr, float normal_thr, int32_t check_max,
                         float *gs_color_error, float *gs_depth_error, float *gs_normal_error,
                         int32_t *color_counter, int32_t *depth_counter, int32_t *normal_counter,
                         float *rescale_counter, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Mask builders and error maps of the mapper (inline PyTorch in the reference).
 * Tiles are the rasterizer's 16x16 blocks (the callers pass stride 16: mapper.py:956,968,985).
 *   dqo_render_range          render_mask = (T_map != 1) and transmission2tilemask(render_mask, 16, ratio)
 *                             (mapper.py:983-987, SLAM/utils.py:752-763); *render_count = render_mask.sum()
 *   dqo_pixelmask_to_tilemask transmission2tilemask on a caller-made bool mask (SLAM/utils.py:752-763)
 *   dqo_color_error_map       sum_c |render - gt|, 0 where render.sum(c) == 0; CHW in, HW out (mapper.py:948-955)
 *   dqo_topk_tilemask         colorerror2tilemask (SLAM/utils.py:765-796): tile means in avg_pool2d's accumulation
 *                             order, the k = int(n_tiles * top_ratio) largest set to 1 (ties: lower tile index);
 *                             or_into != 0 ORs into an existing mask (mapper.py:969); render_mask, when given, is the
 *                             x16 nearest upsampling cropped to the image (mapper.py:971-980)
 *   dqo_tilemask_to_pixelmask that upsampling alone
 *   dqo_render_error_maps     colour / depth / normal error images of error_gaussians_remove (mapper.py:1008-1025);
 *                             gt maps are HWC / HW (frame_map), renders CHW; feeds dqo_accumulate_error
 * ---------------------------------------------------------------------------------------------- */
int dqo_render_range(int32_t W, int32_t H, const float *T_map, float tile_mask_ratio, uint8_t *render_mask /* [H,W] or NULL */,
                     int32_t *tile_mask /* [ceil(H/16), ceil(W/16)] */, int32_t *render_count /* [1] or NULL */, void *stream);
int dqo_pixelmask_to_tilemask(int32_t W, int32_t H, const uint8_t *pixelmask, float tile_mask_ratio, int32_t *tile_mask,
                              int32_t *pixel_count /* [1] or NULL */, void *stream);
int dqo_color_error_map(int32_t W, int32_t H, const float *render /* [3,H,W] */, const float *gt /* [3,H,W] */,
                        float *color_error /* [H,W] */, void *stream);
size_t dqo_topk_tilemask_workspace_bytes(int32_t W, int32_t H);
int dqo_topk_tilemask(int32_t W, int32_t H, const float *error /* [H,W] */, int32_t k, int32_t or_into, int32_t *tile_mask,
                      uint8_t *render_mask /* [H,W] or NULL */, void *workspace, void *stream);
int dqo_tilemask_to_pixelmask(int32_t W, int32_t H, const int32_t *tile_mask, uint8_t *render_mask, void *stream);
int dqo_render_error_maps(int32_t W, int32_t H, const float *render_color /* [3,H,W] */, const float *render_depth /* [H,W] */,
                          const float *gt_color_hwc /* [H,W,3] */, const float *gt_depth /* [H,W] */,
                          const int32_t *depth_index /* [H,W] */, float *color_error, float *depth_error,
                          float *normal_error /* zeros; NULL to skip */, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Tracker front-end (inline PyTorch in the reference).
 *   dqo_depth_maxpool       one level of ImagePyramids(.., "max"): MaxPool2d(1 << level) of the depth (SLAM/icp.py:342-360)
 *   dqo_vertex_normal_map   compute_vertex_map + compute_normal_map fused (SLAM/utils.py:65-125): vertex [H,W,3] from depth
 *                           and K, normal = cross(sobel_y, sobel_x) of the vertex map (replicate border), normalised,
 *                           zero where depth <= min or >= max; normal may be NULL
 *   dqo_normal_map          compute_normal_map on an existing vertex map
 *   dqo_icp_level           ICP.icp (SLAM/icp.py:33-48): `iterations` Gauss-Newton steps of projective point-to-plane ICP
 *                           at one pyramid level; frame 0 is the template.  pose10 is a device float[16] (row-major 4x4),
 *                           updated in place -- no host round trip for the 6x6 solve (icp.py:313-335 inverts on the CPU);
 *                           valid_ratio (device float, may be NULL) = valid correspondences / (H*W) of the last iteration.
 * ---------------------------------------------------------------------------------------------- */
int dqo_depth_maxpool(int32_t W, int32_t H, int32_t level, const float *depth, float *out /* [H>>level, W>>level] */,
                      void *stream);
size_t dqo_vertex_normal_workspace_bytes(void);
int dqo_vertex_normal_map(int32_t W, int32_t H, const float *depth, float fx, float fy, float cx, float cy,
                          float *vertex, float *normal, void *workspace, void *stream);
int dqo_normal_map(int32_t W, int32_t H, const float *vertex, float *normal, void *workspace, void *stream);
size_t dqo_icp_workspace_bytes(void);
int dqo_icp_level(int32_t W, int32_t H, const float *vertex0, const float *vertex1, const float *normal0,
                  const float *normal1, float fx, float fy, float cx, float cy, float distance_threshold,
                  float normal_threshold_cos, float damping, int32_t iterations, float *pose10, float *valid_ratio,
                  void *workspace, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Mapping step (no native boundary exists in the reference: SLAM/multiprocess/mapper.py:799-928 is
 * inline PyTorch).  Masked L1 colour + depth loss and its image gradients in two launches.
 *   colour: mean |img - gt| over render_mask pixels x 3 channels          (mapper.py:847, loss_utils.py:27-31)
 *   depth : mean |d - gt_d| over hit != -1 & gt_d > 0 & (d-gt_d) < depth_err_thres & render_mask (mapper.py:849-857)
 *   total = depth_weight * depth + color_weight * colour
 * `loss_out` is device float[4] = {total, colour, depth, unused}; `counts_out` device int32[2].
 * Empty selections give NaN means exactly like torch (0/0) and zero gradients.
 * ---------------------------------------------------------------------------------------------- */
size_t dqo_loss_workspace_bytes(int32_t W, int32_t H);
int dqo_masked_l1_loss(int32_t W, int32_t H, const float *image /* [3,H,W] */, const float *depth /* [H,W] */,
                       const int32_t *hit_depth /* [H,W] */, const float *gt_color /* [H,W,3] */,
                       const float *gt_depth /* [H,W] */, const uint8_t *render_mask /* [H,W] or NULL */,
                       float color_weight, float depth_weight, float depth_err_thres,
                       float *dL_dimage /* [3,H,W] */, float *dL_ddepth /* [H,W] */, float *loss_out,
                       int32_t *counts_out, void *workspace, void *stream);

/* Fused multi-tensor Adam.  Replaces torch.optim.Adam(l, lr=0.0, eps=1e-15).step() over the parameter groups
 * of GaussianPointCloud.parametrize (SLAM/gaussian_pointcloud.py:331-378; mapper.py:548,906) with one
 * launch, plus the confidence bump of mapper.py:909-910 when `confidence`/`conf_tensor` are given.
 * Semantics: torch Adam, amsgrad=False, weight_decay=0, maximize=False. */
typedef struct dqo_adam_tensor {
    float *param;
    const float *grad;
    float *exp_avg;
    float *exp_avg_sq;
    int64_t numel;
    double lr;
    int32_t row_width; /* trailing elements per Gaussian (for the confidence bump), 0 = n/a */
    int32_t reserved;
} dqo_adam_tensor;
#define DQO_ADAM_MAX_TENSORS 16
/* betas / eps / lr are doubles like the Python floats torch receives ((float)(1 - beta2) != 1 - (float)beta2). */
/* SSIM term of the mask-less global pass of loss_update: ssim_weight * (1 - ssim(image, gt)) (mapper.py:839-841, :874;
 * ssim = utils/loss_utils.py:61-99: 11x11 Gaussian window, sigma 1.5, zero padding, per channel, C1 = 0.01^2,
 * C2 = 0.03^2, mean over 3*H*W).  Two launches produce the value and d(weight * (1 - ssim)) / d image, which is written
 * to (accumulate = 0) or added to (accumulate = 1) `dL_dimage`, e.g. on top of dqo_masked_l1_loss's colour gradient.
 * loss_out: device float[2] = {1 - ssim, weight * (1 - ssim)}; deterministic (fixed-order fp64 reduction). */
size_t dqo_ssim_workspace_bytes(int32_t W, int32_t H);
/* The 11 taps of the 1-D window (host call, no GPU): the bits torch produces for loss_utils.py:42-49. */
void dqo_ssim_window(float *window11);
int dqo_ssim_loss(int32_t W, int32_t H, const float *image /* [3,H,W] */, const float *gt_color /* [H,W,3] */,
                  float weight, float *dL_dimage /* [3,H,W] */, int32_t accumulate, float *loss_out, void *workspace,
                  void *stream);

int dqo_adam_step(const dqo_adam_tensor *tensors /* host array */, int32_t n_tensors, int32_t step, double beta1,
                  double beta2, double eps, float *confidence /* [P] or NULL */, int32_t conf_tensor, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Fused mapping iteration (opt-in).  One call enqueues, on one stream and without any host
 * synchronisation, what one iteration of Mapping.local_optimize does (SLAM/multiprocess/mapper.py:568-599 and
 * loss_update :799-928: masked L1 colour + depth loss, the attach term, the SSIM term of the mask-less final global pass
 * :841 and the semantic colour term :877-880 -- see dqo_keyframe; NOT the normal term (weight 0 in every shipped config);
 * the instance term :881-902 (Method 2) is an L1 on T_map, whose gradient the rasterizer's backward drops
 * (diff_gaussian_rasterization_depth/__init__.py:184), i.e. a reported number without any effect on the parameters):
 * activations of the RAW parameters (exp / sigmoid / normalize,
 * SLAM/gaussian_pointcloud.py:20-30, 724-826; the torch.cat of f_dc / f_rest is never materialised), rasterize
 * forward, masked L1 colour + depth loss, rasterize backward, activation backward and Adam (eps / betas / lrs as in
 * GaussianPointCloud.parametrize :331-378) on the raw parameters in place, plus the confidence bump (:909-910).
 * Parameter order everywhere: 0 xyz [P,3], 1 f_dc [P,1,3], 2 f_rest [P,M-1,3], 3 opacity [P,1], 4 scaling [P,3],
 * 5 rotation [P,4].  M must be 16 (f_dc + f_rest) or 1 (f_dc only).
 * ---------------------------------------------------------------------------------------------- */
typedef struct dqo_map_params {
    float *param[6];
    float *exp_avg[6];
    float *exp_avg_sq[6];
    double lr[6];
    float *confidence; /* [P] or NULL */
    /* [P] bytes or NULL.  ever[i] becomes 1 when Gaussian i first receives a non-zero gradient.  A Gaussian that never
     * has is a fixed point of Adam (zero gradient on zero moments): its gradients are then neither written by the backward
     * nor read by the optimiser.  Owned by the caller, zero-initialised TOGETHER WITH the moments (set to 1 wherever moments
     * are restored non-zero). */
    uint8_t *ever;
    /* Optional, with `ever`: compact list of the Gaussians whose flag is set (u32[P], append order) and its length (device
     * int), zero-initialised together with `ever`.  The backward appends a Gaussian when its flag flips; the optimiser then
     * walks the list instead of the whole cloud.  NULL: the optimiser scans all P Gaussians and skips by flag. */
    uint32_t *ever_list;
    int32_t *ever_count;
    /* Attach term of loss_update (mapper.py:810-829): Gaussians whose INITIAL opacity sigmoid(init_opacity) is below
     * attach_opacity_thres (0.9) are anchored to their initial raw xyz / scaling / rotation by
     * attach_weight (1000) * (mse + mse + mse); the gradient is added to the rasterizer's before Adam, the value lands in
     * loss_out[3] (the reference's "scale_loss" report).  init_* are the history_stat tensors of local_optimize
     * (mapper.py:535-545); attach_count is a device int written by dqo_attach_count.  init_opacity == NULL disables it. */
    const float *init_xyz, *init_scaling, *init_rotation, *init_opacity;
    const int32_t *attach_count;
    float attach_weight, attach_opacity_thres;
    /* device int32[4] or NULL.  [0] = Adam steps taken so far (the bias corrections use [0] + 1), [1] = steps SKIPPED because
     * the forward overflowed its instance capacity; both are maintained on the device, so the call has no per-step host
     * argument (CUDA-graph replayable) and a skipped step can neither be missed (sticky until the caller clears [1]) nor
     * advance the bias correction.  NULL: the `step` argument is used and nothing is counted. */
    int32_t *step_state;
    /* Semantic colours [P,3] (GaussianPointCloud._semantics, parameter group "semantics_color" with lr semantic_lr,
     * gaussian_pointcloud.py:371-378), their Adam moments and learning rate.  Used when the keyframe carries a semantic
     * target (dqo_keyframe.gt_semantic); NULL otherwise. */
    float *semantics, *semantics_exp_avg, *semantics_exp_avg_sq;
    double lr_semantics;
    /* The DQO_STEP_TERM_* mask the step workspace was sized and initialised with: a keyframe that asks for a term the
     * workspace has no room for is refused (DQO_ERR_WORKSPACE). */
    int32_t workspace_terms;
    int32_t reserved;
} dqo_map_params;
typedef struct dqo_keyframe {
    const float *gt_color;      /* [H,W,3] */
    const float *gt_depth;      /* [H,W] */
    const uint8_t *render_mask; /* [H,W] or NULL */
    const int32_t *tile_mask;   /* [ceil(H/16), ceil(W/16)] */
    const float *viewmatrix, *projmatrix, *campos, *background;
    float color_weight, depth_weight, depth_err_thres;
    /* Optional terms of loss_update (0 / NULL = off).
     * ssim_weight: adds ssim_weight * (1 - ssim(image, gt_color)); like the reference only when render_mask == NULL
     *   (mapper.py:839-841, the final global pass); ignored for masked keyframes.
     * gt_semantic + semantic_weight: adds semantic_weight * mean |semantic_seg - gt_semantic| over the render_mask pixels
     *   (mapper.py:877-880).  The semantic image is the blend of dqo_map_params.semantics over the lists of the main render
     *   (the reference runs the whole rasterizer a second time with colors_precomp, SLAM/render.py:227-246); its gradient
     *   reaches the semantic colours and, through alpha and the 2-D geometry, every geometric parameter. */
    float ssim_weight;
    float semantic_weight;
    const float *gt_semantic;   /* [H,W,3] or NULL */
} dqo_keyframe;
/* Optional loss terms a step workspace has room for (their scratch is 36 B per pixel for SSIM, 28 B per pixel + 44 B per
 * Gaussian for the semantic term: not carried by workspaces that never use them). */
#define DQO_STEP_TERM_SSIM 1
#define DQO_STEP_TERM_SEMANTIC 2
size_t dqo_mapping_step_workspace_bytes(int32_t P, int32_t M, int32_t W, int32_t H, int64_t instance_capacity,
                                        int32_t terms /* DQO_STEP_TERM_* mask */);
/* One-time initialisation of a freshly allocated step workspace: lets the step run with dqo_rast_settings.geom_clean = 1. */
int dqo_mapping_step_workspace_init(int32_t P, int32_t M, int32_t W, int32_t H, int64_t instance_capacity, int32_t terms,
                                    void *workspace, void *stream);
/* loss_out: device float[8] {total, colour, depth, attach, ssim (1 - ssim), semantic, 0, 0} -- `total` is the reference's
 * `loss` (mapper.py:870-880, :904: every weighted term except the attach term); counts_out: device int32[2]; status: device int32[DQO_ST_WORDS]
 * (check DQO_ST_OVERFLOW together with the loss read-back; on overflow the render is invalid and the Adam update is
 * skipped on the device -- parameters and moments are untouched, repeat the step with a larger capacity and the same
 * `step` number: see mapping.FusedMappingStep.check).  * The list of non-empty tiles is not produced by the step: status[DQO_ST_TILE_NUM] stays 0.
 */
int dqo_mapping_step(const dqo_rast_settings *s, const dqo_map_params *p, const dqo_keyframe *kf, int32_t step,
                     double beta1, double beta2, double eps, void *workspace, int64_t instance_capacity,
                     float *loss_out, int32_t *counts_out, int32_t *status, void *stream);
/* attach_count = number of Gaussians with sigmoid(init_opacity) < opacity_thres (device int; mapper.py:810-812). */
int dqo_attach_count(int32_t P, const float *init_opacity, float opacity_thres, int32_t *count, void *stream);
/* Device pointers of the images rendered by the last step inside `workspace` (any output may be NULL). */
int dqo_mapping_step_outputs(int32_t P, int32_t M, int32_t W, int32_t H, int64_t instance_capacity, void *workspace,
                             float **color, float **depth, int32_t **hit_depth, float **T_map);

/* ------------------------------------------------------------------------------------------------
 * Dual quadrics (SLAM/multiprocess/quadrics.py).  Batched over objects.
 *   dqo_quadric_init   : Object.__init__ single-view construction (quadrics.py:451-487)
 *   dqo_quadric_project: Ellipsoid.project + Ellipse.ComputeBbox (quadrics.py:388-425,148-248) in fp64
 *   dqo_quadric_refine : Object_Optimize_only's 20-iteration IoU-Adam loop (quadrics.py:2234-2298) with
 *                        Ellipsoid_tensor.forward / Ellipse_tensor (quadrics.py:2018-2225) in fp32,
 *                        analytic gradients; the per-iteration view choice is supplied by the caller.
 * ---------------------------------------------------------------------------------------------- */
int dqo_quadric_init(int32_t n, const double *bboxes /* [n,4] */, const double *depth_stats /* [n,2] avg,diff */,
                     const double *K /* [3,3] */, const double *Rts /* [n,3,4] */, double *axes /* [n,3] */,
                     double *R /* [n,3,3] */, double *center /* [n,3] */, void *stream);
int dqo_quadric_project(int32_t n, const double *axes, const double *R, const double *center,
                        const double *P /* [n,3,4] */, double *bbox /* [n,4] */, double *ellipse /* [n,5] ax0,ax1,angle,cx,cy */,
                        void *stream);
int dqo_quadric_refine(int32_t n, int32_t iters, int32_t max_views, const int32_t *n_views /* [n] */,
                       const float *obs_bboxes /* [n,max_views,4] */, const float *Ps /* [n,max_views,3,4] */,
                       const int32_t *view_choice /* [n,iters] */, float lr_axes, float lr_center, float lr_R,
                       float *axes /* [n,3] in/out */, float *R /* [n,9] in/out */, float *center /* [n,3] in/out */,
                       float *last_loss /* [n] or NULL */, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DQO_B200_H_ */

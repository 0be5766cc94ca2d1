/*
 * fnx.h -- C ABI of libfnx.so, the B200 (sm_100a) hot path of FluidNexus' FluidDynamics stage.
 *
 * Everything here takes raw DEVICE pointers (unless a parameter says "host"), plain sizes and a
 * cudaStream_t (passed as void*), and returns an int status (FNX_OK == 0).  After a non-zero
 * status `fnx_last_error()` returns a thread-local description.  No torch types cross this line.
 *
 * Each entry point names the reference interface it replaces (paths relative to the reference
 * repository root; R3 = FluidDynamics/submodules/gaussian_rasterization_ch3, R1 = ..._ch1 (same
 * code, NUM_CHANNELS 1), KNN = FluidDynamics/submodules/simple-knn, FD = FluidDynamics).
 *
 * Conventions shared with the reference:
 *   - view / proj matrices: 16 floats each, the row-major storage of the TRANSPOSED matrices, exactly
 *     what Camera.world_view_transform / full_proj_transform hold (FD/scene/camera.py:90-108).
 *   - quaternions are (r,x,y,z) and are NOT normalised inside (R3/cuda_rasterizer/forward.cu:121).
 *   - colours are "precomputed" [P,C] on every FluidNexus pipe (FD/renderer/pipe_fluid.py:107-118); the reference module's
 *     other colour input, spherical harmonics (sh [P,M,3] + degree + campos), is accepted by the plain forward / backward for
 *     3 channels, like the reference.
 *   - all scratch is caller-owned and grown through an allocation callback, like the reference's
 *     resize lambdas (R3/rasterize_points.cu:27-33); its layout is private to the library.
 */
#ifndef FNX_H_INCLUDED
#define FNX_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FNX_OK 0
#define FNX_ERR_INVALID 1     /* bad argument (shape / null / range) */
#define FNX_ERR_CUDA 2        /* a CUDA runtime call or kernel launch failed */
#define FNX_ERR_UNSUPPORTED 3 /* feature of the reference interface that is dead on the FluidNexus path */
#define FNX_ERR_CAPACITY 4    /* FNX_NO_HOST_SYNC was given and the instance capacity was too small */
#define FNX_ERR_ALLOC 5       /* allocation callback returned NULL */

#define FNX_ABI_VERSION 1

typedef void *fnx_stream_t; /* cudaStream_t */

/* Allocation callback: must return a device pointer to at least `bytes` bytes (256-B aligned) that stays
 * valid until the paired backward has run.  `ctx` is passed through.  Mirrors resizeFunctional,
 * R3/rasterize_points.cu:27-33. */
typedef void *(*fnx_alloc_fn)(void *ctx, size_t bytes);

int fnx_abi_version(void);
const char *fnx_last_error(void);
/* Compiled-for architecture string, e.g. "sm_100a". */
const char *fnx_build_arch(void);
/* Number of libfnx's own kernels launched so far by this process (library calls such as CUB sorts not counted). */
unsigned long long fnx_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Rasterizer  (replaces R3|R1 `_C.rasterize_gaussians`, `_C.rasterize_gaussians_backward`,
 * `_C.mark_visible`: R3/rasterize_points.h:18-64, R3/ext.cpp:15-19, host orchestration
 * R3/cuda_rasterizer/rasterizer_impl.cu:131-414, kernels forward.cu / backward.cu)
 * ---------------------------------------------------------------------------------------------- */

/* flags */
#define FNX_NO_HOST_SYNC 1u /* never block the host and use no events (the call can be captured into a CUDA graph):
                               trust `instance_capacity_hint`; on overflow nothing is rendered (images = background),
                               *num_rendered_host is left at -1 and fnx_raster_check() / num_rendered_pinned report it */
#define FNX_EXACT_RECT 2u   /* bin with the reference's full 3-sigma tile rectangle (no opacity-aware tile
                               culling).  Results are identical either way; this exists for A/B tests. */

#define FNX_BIN_ONLY 4u     /* forward: stop after the record stream is packed (no blending; out_color/out_depth may be
                               NULL).  Used to build a stream that fnx_raster_blend_merged consumes. */
#define FNX_ALL_FROZEN 8u   /* every Gaussian of this call is frozen (no gradients): marks all its records */
#define FNX_STATIC_TILE_CACHE 16u /* fnx_raster_blend_merged only: tiles without any dynamic instance keep the pixels that are
                               already in out_color / out_depth instead of being blended again -- the frozen set and the
                               cameras are fixed, so those pixels cannot change.  Which tiles of out_color / out_depth hold valid
                               static-only pixels is tracked per 8x8 patch in the DYNAMIC scratch's image buffer: reset with
                               fnx_raster_tile_cache_set(valid = 0) for fresh buffers, or (valid = 1) after copying the
                               static-only render of fnx_raster_static_prepare into them.  The SAME out_color / out_depth
                               buffers and image scratch must be passed to every call. */

#define FNX_BUCKET_BINNING 32u /* no global sorts: instances are histogrammed per tile and dropped into per-tile buckets, each
                               bucket is sorted by (depth, index) in shared memory.  Results are identical to the sorted
                               path.  With FNX_BIN_ONLY (the DYNAMIC set of merged streams; needs FNX_NO_HOST_SYNC, C == 3)
                               the buckets are left unsorted and fnx_raster_blend_merged (same flags) sorts them while it
                               merges. */

typedef struct fnx_raster_args {
    /* sizes */
    int32_t P;        /* Gaussians */
    int32_t V;        /* cameras rendered by this call (>=1). The reference renders one; V>1 batches views */
    int32_t C;        /* colour channels: 1 (R1) or 3 (R3) */
    int32_t W, H;     /* image size */
    /* per-Gaussian inputs, shared by all V views (reference arg names in brackets) */
    const float *means3D;       /* [P,3] */
    const float *colors;        /* [P,C]  (colors_precomp) */
    const float *opacities;     /* [P]    (opacity [P,1]) */
    const float *scales;        /* [P,3] or NULL when cov3D_precomp is given */
    const float *rotations;     /* [P,4] or NULL */
    const float *cov3D_precomp; /* [P,6] or NULL */
    const float *sh;            /* [P, sh_coeffs, 3] spherical-harmonics coefficients INSTEAD of colors (C == 3 only; colours are
                                   max(0, 0.5 + sum_k B_k(dir) sh_k), R3/cuda_rasterizer/forward.cu:20-67); plain forward / backward
                                   only (FNX_ERR_UNSUPPORTED with workspaces / merged streams).  NULL on every FluidNexus pipe. */
    /* per-view inputs */
    const float *view_matrix;   /* [V,16] */
    const float *proj_matrix;   /* [V,16] */
    const float *bg;            /* [C] (R1 reads bg[0]) */
    float tan_fov_x, tan_fov_y, scale_modifier;
    int32_t prefiltered;        /* accepted, unused (reference only traps on inconsistency) */
    uint32_t flags;
    int64_t instance_capacity_hint; /* 0 = size exactly (one host sync, like the reference) */
    int64_t *num_rendered_pinned;   /* optional PINNED HOST int64 the device writes the instance count to (it can be
                                       polled at any later time; with FNX_NO_HOST_SYNC it is the only read-back) */
    int32_t grad_begin, grad_end;   /* Gaussians [grad_begin, grad_end) receive gradients in the backward; the others are
                                       FROZEN (e.g. the background set that FD/renderer/pipe_dynamics.py:51-57 concatenates
                                       behind the fluid particles): they still occlude, but their gradient rows are
                                       written as zeros.  grad_end <= grad_begin means "all" (the reference's behaviour). */
    const uint32_t *tile_order;     /* optional DEVICE array of V * tiles entries (tiles = ceil(W/16) * ceil(H/16)): a permutation
                                       of the (view, tile) units, view * tiles + tile, in the order the blend kernels should start
                                       them (fnx_raster_tile_order: longest first).  NULL = natural order.  Scheduling only:
                                       results do not depend on it. */
    const int32_t *static_view_map; /* fnx_raster_blend_merged only: DEVICE array of V entries, static_view_map[v] = index of this
                                       call's view v among the `static_views` cameras the static stream was built for, so that ONE
                                       static stream per frame serves every subset of its cameras (the reference draws
                                       random.sample(cur_viewpoint_set, batch) cameras per iteration,
                                       FD/entries_fluid_nexus/train_physical_particle.py:337).  NULL = identity (static_views == V). */
    int32_t static_views;           /* cameras of the static stream (0 = V) */
    int32_t sh_degree;              /* with sh: active degree 0..3 (`degree` of R3/rasterize_points.h:18-37) */
    int32_t sh_coeffs;              /* with sh: coefficients per Gaussian M (sh is [P, M, 3]), (sh_degree+1)^2 <= M <= 16 */
    int32_t reserved0;
    const float *campos;            /* with sh: camera centres [V,3] (`campos`) */
} fnx_raster_args;

/* Opaque handles to the three scratch buffers of one forward (what the reference returns as
 * geomBuffer / binningBuffer / imgBuffer and hands back to backward). */
typedef struct fnx_raster_scratch {
    void *geom;    size_t geom_bytes;
    void *binning; size_t binning_bytes;
    void *image;   size_t image_bytes;
    int64_t binning_capacity; /* instances the binning buffer was sized for */
    int32_t check_slot;       /* private: which read-back slot carries this forward's instance count (a ring of 256 per device:
                                 call fnx_raster_check before 256 later forwards on that device) */
    int32_t reserved;         /* private: the device of that ring */
} fnx_raster_scratch;

/* Sizes, for callers that pre-allocate instead of using callbacks. */
size_t fnx_raster_geom_bytes(int32_t P, int32_t V);
size_t fnx_raster_image_bytes(int32_t W, int32_t H, int32_t V);
size_t fnx_raster_binning_bytes(int64_t instance_capacity, int32_t C);

/* Forward.  Outputs: out_color [V,C,H,W], out_depth [V,1,H,W] (median depth, default 15), radii [V,P] int32.
 * `*num_rendered_host` receives the number of (tile, Gaussian) instances binned (after tile culling, so it
 * can be smaller than the reference's).  The three allocators are called at most a few times each; the final
 * pointers/sizes are reported in *scratch.  Replaces RasterizeGaussiansCUDA, R3/rasterize_points.cu:35-115. */
int fnx_raster_forward(const fnx_raster_args *a, fnx_alloc_fn alloc_geom, void *geom_ctx, fnx_alloc_fn alloc_binning,
                       void *binning_ctx, fnx_alloc_fn alloc_image, void *image_ctx, float *out_color,
                       float *out_depth, int32_t *radii, int64_t *num_rendered_host, fnx_raster_scratch *scratch,
                       fnx_stream_t stream);
/* Channel-count-specific names, one per reference extension module. */
int fnx_raster_forward_ch1(const fnx_raster_args *a, fnx_alloc_fn ag, void *cg, fnx_alloc_fn ab, void *cb, fnx_alloc_fn ai,
                           void *ci, float *out_color, float *out_depth, int32_t *radii, int64_t *num_rendered_host,
                           fnx_raster_scratch *scratch, fnx_stream_t stream);
int fnx_raster_forward_ch3(const fnx_raster_args *a, fnx_alloc_fn ag, void *cg, fnx_alloc_fn ab, void *cb, fnx_alloc_fn ai,
                           void *ci, float *out_color, float *out_depth, int32_t *radii, int64_t *num_rendered_host,
                           fnx_raster_scratch *scratch, fnx_stream_t stream);

/* Gradient outputs of the backward.  Any pointer may be NULL (that gradient is then not written).
 * All non-NULL outputs are fully overwritten (no pre-zeroing needed), summed over the V views. */
typedef struct fnx_raster_grads {
    float *dL_dmeans3D;  /* [P,3] */
    float *dL_dmeans2D;  /* [V,P,3]: screen-space gradient, z == 0, x,y scaled by 0.5*W, 0.5*H (backward.cu:444-445) */
    float *dL_dcolors;   /* [P,C] */
    float *dL_dopacity;  /* [P] */
    float *dL_dscales;   /* [P,3] (written only when scales were given) */
    float *dL_drotations;/* [P,4] raw un-normalised quaternion gradient (backward.cu:326) */
    float *dL_dcov3D;    /* [P,6] */
    float *dL_dsh;       /* [P,sh_coeffs,3] (written only when sh was given) */
} fnx_raster_grads;

/* Backward.  dL_dout_color [V,C,H,W].  `a` must equal the forward's args (`opacities` may be NULL here: like the reference's
 * backward, which is not handed them, it reads them from the forward's buffers); `scratch` is what forward reported;
 * `num_rendered` what it returned.  Depth has no gradient (R3/README.md:13).
 * Replaces RasterizeGaussiansBackwardCUDA, R3/rasterize_points.cu:117-194. */
int fnx_raster_backward(const fnx_raster_args *a, const fnx_raster_scratch *scratch, int64_t num_rendered,
                        const int32_t *radii, const float *dL_dout_color, const fnx_raster_grads *g,
                        fnx_stream_t stream);
int fnx_raster_backward_ch1(const fnx_raster_args *a, const fnx_raster_scratch *scratch, int64_t num_rendered,
                            const int32_t *radii, const float *dL_dout_color, const fnx_raster_grads *g,
                            fnx_stream_t stream);
int fnx_raster_backward_ch3(const fnx_raster_args *a, const fnx_raster_scratch *scratch, int64_t num_rendered,
                            const int32_t *radii, const float *dL_dout_color, const fnx_raster_grads *g,
                            fnx_stream_t stream);

/* Static + dynamic streams (3 channels).  The frozen background set that FD/renderer/pipe_dynamics.py:51-57 concatenates
 * behind the fluid particles does not change within a frame (nor do the cameras), so it is binned ONCE with
 * fnx_raster_forward(FNX_BIN_ONLY | FNX_ALL_FROZEN) -> `stat`; every iteration bins only the moving Gaussians
 * (FNX_BIN_ONLY) -> `dyn`, then this call merges the two depth-sorted streams per tile (dynamic first on equal depth,
 * which reproduces the reference's order for the concatenated array [dynamic; static]) and blends.  Results are
 * identical to one forward over the concatenated set.  merged_records: 48 * (dyn capacity + static instances) bytes.
 * Same V, W, H, cameras and bg for both sets.  out_color [V,3,H,W], out_depth [V,1,H,W]. */
int fnx_raster_blend_merged(const fnx_raster_args *dyn_args, const fnx_raster_scratch *dyn, const fnx_raster_scratch *stat,
                            int32_t P_static, void *merged_records, float *out_color, float *out_depth, fnx_stream_t stream);
/* Optional, once per static stream (after the FNX_BIN_ONLY | FNX_ALL_FROZEN forward, `static_args` = its args): blends
 * the static stream alone into out_color / out_depth and records, per tile, how deep that blend reaches.  Afterwards
 *   - merges copy only the static records a blend can reach (inserting occluders in front can only make a pixel
 *     terminate earlier), and
 *   - fnx_raster_blend_merged(FNX_STATIC_TILE_CACHE) skips tiles that hold no dynamic instance.
 * Results of blend_merged / backward_merged are unchanged. */
int fnx_raster_static_prepare(const fnx_raster_args *static_args, const fnx_raster_scratch *stat, float *out_color,
                              float *out_depth, fnx_stream_t stream);
/* Backward of fnx_raster_blend_merged: gradients for the DYNAMIC Gaussians only (g sized for dyn_args->P); radii = the dynamic
 * forward's radii. */
int fnx_raster_backward_merged(const fnx_raster_args *dyn_args, const fnx_raster_scratch *dyn, const fnx_raster_scratch *stat,
                               const void *merged_records, const int32_t *radii, const float *dL_dout_color,
                               const fnx_raster_grads *g, fnx_stream_t stream);

/* Device address of the "instance capacity overflowed" flag (int32, 0 / 1) that every forward on `scratch` (its geom buffer must
 * be set) writes: lets later kernels of the same stream -- fnx_adam_step_dev_gated -- drop the work of a void forward without a
 * host round trip. */
int fnx_raster_overflow_flag(const fnx_raster_scratch *scratch, const int32_t **flag_dev);

/* Marks every tile of the out_color / out_depth buffers that go with the dynamic scratch `dyn` as holding (valid != 0) or not
 * holding (0) its static-only pixels; see FNX_STATIC_TILE_CACHE.  `dyn->image` must be set. */
int fnx_raster_tile_cache_set(const fnx_raster_scratch *dyn, int32_t W, int32_t H, int32_t V, int32_t valid, fnx_stream_t stream);

/* Longest-first order of the (view, tile) units by the work of the LAST forward on `scratch` (for merged streams pass the
 * static stream's scratch as `stat`, else NULL), for fnx_raster_args.tile_order of later forward / backward calls on the same
 * scene.  order: V * tiles uint32 (device).  The reference dispatches one block per tile in raster order
 * (R3/cuda_rasterizer/forward.cu:389); with a few hundred long tiles among thousands of short ones that leaves the SMs idle
 * for the last ~20 % of the kernel. */
int fnx_raster_tile_order(const fnx_raster_scratch *scratch, const fnx_raster_scratch *stat, int32_t W, int32_t H, int32_t V,
                          uint32_t *order, fnx_stream_t stream);

/* After a FNX_NO_HOST_SYNC forward: blocks until the forward's instance count is known and returns FNX_OK or
 * FNX_ERR_CAPACITY; *num_rendered_host is set either way. */
int fnx_raster_check(const fnx_raster_scratch *scratch, int64_t *num_rendered_host, fnx_stream_t stream);

/* present[i] = (view-space z of means3D[i] > 0.2).  present is uint8/bool [P].
 * Replaces markVisible, R3/rasterize_points.cu:196-215 (checkFrustum, rasterizer_impl.cu:52-63). */
int fnx_mark_visible(int32_t P, const float *means3D, const float *view_matrix, const float *proj_matrix,
                     uint8_t *present, fnx_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Fixed-radius neighbour grid + particle physics  (replaces torch_cluster.radius / radius_graph
 * [torch-cluster 1.6.3] and torch_scatter.scatter_min [torch-scatter 2.1.2] as called from
 * FD/gaussian_splatting/gm_fluid.py:873,909,1076,1088,1114,1140,1206,1256,1272,1301, the poly6 gather /
 * index_add_ chains of gm_fluid.py:846-862,1107-1158,1291-1336, FD/utils/loss_utils.py:98-121 (distance_loss),
 * simple_knn distCUDA2 KNN/spatial.h:14, and torch.optim.Adam at gm_fluid.py:349)
 * ---------------------------------------------------------------------------------------------- */

/* A hashed uniform grid over n points [n,3] with the given cell size, built into caller-owned scratch of
 * fnx_grid_bytes(n) bytes.  Searches need radius <= cell. */
size_t fnx_grid_bytes(int32_t n);
int fnx_grid_build(const float *pts, int32_t n, float cell, void *grid, fnx_stream_t stream);

/* torch_cluster.radius(x, y, r, max_num_neighbors) semantics: for query y_c the neighbours are the x_j with
 * sum((x_j-y_c)^2) < r^2 (strict), truncated to the first `max_num_neighbors` in x-index order.
 * counts[c] = number of neighbours kept (int32 [ny], may be NULL); kth[c] = largest x index kept when the cap
 * binds, INT32_MAX otherwise (int32 [ny], may be NULL): "j is kept  <=>  j within r and j <= kth[c]". */
int fnx_radius_count(const void *grid_x, int32_t nx, float cell, const float *y, int32_t ny, float r,
                     int32_t max_num_neighbors, int32_t *counts, int32_t *kth, fnx_stream_t stream);
/* Edge list like torch_cluster: offsets = exclusive prefix sum of counts (int64 [ny]); writes edge_query (index into
 * y) and edge_x (index into x, ascending per query), int64 each. */
int fnx_radius_fill(const void *grid_x, int32_t nx, float cell, const float *y, int32_t ny, float r, const int32_t *kth,
                    const int64_t *offsets, int64_t *edge_query, int64_t *edge_x, fnx_stream_t stream);

/* P2/P3 density constraint (gm_fluid.py:1107-1158): p_ratio[i] = sum_j poly6(|X_i-X_j|^2) / imass[i] / p0 over the
 * radius_graph(X, H, loop=True, K) edges (kth from fnx_radius_count(grid(X), X)); grid must be built on X with cell H.
 * bwd: dL_dX (+)= d/dX of sum_i dL_dpratio[i] * p_ratio[i]   (deterministic gather, no atomics). */
int fnx_pbf_density_fwd(const void *grid, const float *X, int32_t N, const float *imass, const int32_t *kth, float H,
                        float p0, float *p_ratio, fnx_stream_t stream);
int fnx_pbf_density_bwd(const void *grid, const float *X, int32_t N, const float *imass, const int32_t *kth, float H,
                        float p0, const float *dL_dpratio, float *dL_dX, int32_t accumulate, fnx_stream_t stream);
/* fnx_pbf_ratio_loss + fnx_pbf_density_bwd in one call for the training step's loss = weight * mean((p_ratio - 1)^2)
 * (FD/entries_scalar_real/train_physical_particle.py:336-342): *loss (device scalar) receives the un-weighted mean, dL_dpratio [N]
 * the gradient w.r.t. p_ratio (scratch the caller may inspect), dL_dX as fnx_pbf_density_bwd. */
int fnx_pbf_density_bwd_ratio(const void *grid, const float *X, int32_t N, const float *imass, const int32_t *kth, float H,
                              float p0, const float *p_ratio, float weight, float *loss, float *dL_dpratio, float *dL_dX,
                              int32_t accumulate, fnx_stream_t stream);
/* fnx_radius_count(grid(X), X, K) + fnx_pbf_density_fwd in one neighbour walk: kth_out [N] is written for the backward;
 * cap_flag is one int32 of device scratch (raised when some particle has more than K neighbours; p_ratio is then
 * recomputed with the cut-offs, so the result always equals the two-call sequence). */
int fnx_pbf_density_fwd_counted(const void *grid, const float *X, int32_t N, const float *imass, int32_t max_num_neighbors,
                                float H, float p0, int32_t *kth_out, float *p_ratio, int32_t *cap_flag, fnx_stream_t stream);

/* P1 (gm_fluid.py:1291-1336): A = visual + secs * sum_j w u_j / max(sum_j w, 1e-8), w = poly6(|visual-X_j|^2),
 * u_j = (X_j - xyz_j)/secs over radius(x=X, y=visual, H, K) edges; visual_out = A / out_div (out_div = 1, or the
 * scale factor 100 to get render units directly, FD/renderer/pipe_fluid.py:44-45).  grid_hidden is built on X
 * (cell H); kthV from fnx_radius_count(grid_hidden, visual).  num_out [V,3] / den_out [V] are saved for the backward.
 * bwd gathers per hidden particle over grid_visual (built on `visual`, cell H):
 *   dL_dX (+)= J_A^T ( g_scale * (dL_dvisual_out + dL_dvisual_out2) ),  dL_dvisual_out2 may be NULL. */
int fnx_visual_advect_fwd(const void *grid_hidden, const float *X, const float *xyz, int32_t N, const float *visual,
                          int32_t V, const int32_t *kthV, float H, float secs, float out_div, float *visual_out,
                          float *num_out, float *den_out, fnx_stream_t stream);
/* fnx_radius_count(grid_hidden, visual, K) + fnx_visual_advect_fwd in one neighbour walk; kthV_out [V] is written for
 * the backward. */
int fnx_visual_advect_fwd_counted(const void *grid_hidden, const float *X, const float *xyz, int32_t N, const float *visual,
                                  int32_t V, int32_t max_num_neighbors, float H, float secs, float out_div, float *visual_out,
                                  float *num_out, float *den_out, int32_t *kthV_out, fnx_stream_t stream);
int fnx_visual_advect_bwd(const void *grid_visual, const float *X, const float *xyz, int32_t N, int32_t V,
                          const int32_t *kthV, const float *num, const float *den, const float *dL_dvisual_out,
                          const float *dL_dvisual_out2, float g_scale, float H, float secs, float *dL_dX,
                          int32_t accumulate, fnx_stream_t stream);

/* P5 distance_loss (loss_utils.py:98-121): *loss = sum_{i != j, d_ij < thr} (thr - d_ij)^2 (device scalar),
 * dL_dpts [n,3] = grad_scale * dloss/dpts (may be NULL).  grid built on pts with cell >= threshold. */
int fnx_pair_distance_loss(const void *grid, const float *pts, int32_t n, float cell, float threshold, float grad_scale,
                           float *loss, float *dL_dpts, fnx_stream_t stream);

/* distCUDA2 (KNN/simple_knn.cu:134-202): mean squared distance to the 3 nearest other points, [n]. */
int fnx_knn3_mean_dist2(const void *grid, const float *pts, int32_t n, float cell, float *mean_dist2, fnx_stream_t stream);

/* P3 map (gm_fluid.py:846-862): X = scale*e, Y = X + secs*((X-xyz)/secs + b*secs + secs*force),
 * b = buoyancy*(1 - e_y/buoyancy_max_y) if buoyancy_max_y > 0 else buoyancy.  X or Y may be NULL. */
int fnx_pbf_next_tick_fwd(int32_t N, const float *e, const float *xyz, const float *buoyancy, const float *force,
                          float secs, float buoyancy_max_y, float scale_factor, float *X, float *Y, fnx_stream_t stream);
/* dL_de = scale*dL_dX + (dY/de)^T dL_dY + lambda_exyz * d/de mean((scale*e - estimate_xyz)^2)  (P4,
 * train_physical_particle.py:333-334); *exyz_loss (device scalar, may be NULL) receives the un-weighted mean.
 * dL_dX, dL_dY, estimate_xyz may be NULL. */
int fnx_pbf_combine_grad(int32_t N, const float *e, const float *buoyancy, float secs, float buoyancy_max_y,
                         float scale_factor, const float *dL_dX, const float *dL_dY, const float *estimate_xyz,
                         float lambda_exyz, float *dL_de, float *exyz_loss, fnx_stream_t stream);
/* fnx_pbf_combine_grad followed by fnx_adam_step_dev_gated(grad = dL_de, grad_scale = 1) on `e` itself, in one element-wise pass
 * (set_batch_gradient_current + optimizer.step of FD/entries_scalar_real/train_physical_particle.py:379-380 for a step whose views
 * all ran in this process).  dL_de is still written (callers log / reduce it). */
int fnx_pbf_combine_grad_adam(int32_t N, float *e, const float *buoyancy, float secs, float buoyancy_max_y, float scale_factor,
                              const float *dL_dX, const float *dL_dY, const float *estimate_xyz, float lambda_exyz, float *dL_de,
                              float *exyz_loss, float *exp_avg, float *exp_avg_sq, float lr, float beta1, float beta2, float eps,
                              int32_t *step_dev, float *bc_dev, const int32_t *skip_flag, fnx_stream_t stream);
/* *loss = mean((p_ratio-1)^2) (l2_loss vs ones, train_physical_particle.py:336-342); dL_dpratio = weight * d loss. */
int fnx_pbf_ratio_loss(int32_t N, const float *p_ratio, float weight, float *loss, float *dL_dpratio, fnx_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * No-grad PBF solver tick  (replaces the torch chains of gm_fluid.py:809-844
 * guess_hidden_particles, :896-1021 project_gas_constraints, :1160-1175 confirm_guess_hidden_particles and
 * :1197-1239 update_visual_particles).  All state arrays are [N,3] / [N] device fp32 and updated in place.
 * ---------------------------------------------------------------------------------------------- */
/* velocity += (gravity*alpha) * coeff * secs + secs*force (+ clamp(pow(y/scale, wind_power)*wind_force, 0, wind_force_max)*secs),
 * coeff = 1 - y/(buoyancy_max_y*scale_factor) if buoyancy_max_y > 0 else 1; buoyancy = gravity*alpha (* decay if decay > 0);
 * force = 0; estimate_xyz = xyz + secs*velocity; counts = 0.  gravity3 / wind_force3 are HOST pointers to 3 floats. */
int fnx_pbf_guess_hidden(int32_t N, const float *xyz, float *velocity, float *buoyancy, float *force, float *estimate_xyz,
                         float *counts, const float *gravity3_host, float alpha, float secs, float buoyancy_max_y,
                         float scale_factor, float buoyancy_decay_rate, int32_t use_wind, const float *wind_force3_host,
                         float wind_power, float wind_force_max, fnx_stream_t stream);
/* One solver iteration on estimate_xyz (in place): grid + radius_graph(exyz, H, loop=True, K) semantics, density
 * p_ratio, lambda = -(p_ratio-1)/(sum|grad|^2 + |sum grad|^2 + relaxation), force += velocity*(1-p_ratio)*(-k)
 * (skipped when force == NULL), exyz += sum (lambda_i+lambda_j+s_corr) spiky / p0 / (neighbours + counts),
 * s_corr = -K_P (poly6/poly6(DQ_P^2 H^2))^E_P.  grid: fnx_grid_bytes(N) scratch; kth/lambda/nlen: [N] scratch;
 * p_ratio_out [N] optional. */
int fnx_pbf_project_gas_constraints(void *grid, float *estimate_xyz, int32_t N, const float *imass, const float *velocity,
                                    float *force, const float *counts, float H, float p0, float k, int32_t max_num_neighbors,
                                    float relaxation, float K_P, int32_t E_P, float DQ_P, int32_t *kth_scratch,
                                    float *lambda_scratch, float *nlen_scratch, float *p_ratio_out, fnx_stream_t stream);
/* degree[i] = bincount(row)[i] of radius_graph(X, r, loop, max_num_neighbors) (row = neighbour index): what
 * remove_invalid_particles (gm_fluid.py:864-891) thresholds with min_neighbors.  grid: fnx_grid_bytes(N) scratch. */
int fnx_radius_graph_degree(void *grid, const float *X, int32_t N, float r, int32_t loop, int32_t max_num_neighbors,
                            int32_t *kth_scratch, int32_t *degree, fnx_stream_t stream);
/* Rigid coupling (gm_fluid.py:1058-1105 project_rigid_body_constraints, :1241-1289 ..._for_visual_particles): every point
 * of xyz [N,3] that lies inside the body is moved onto the nearest of the rigid-body samples rigid_xyz [M,3] found by
 * radius(x=rigid_xyz, y=point, r, max_num_neighbors) (<= 0: no cap; ties: smaller sample index).  body 0 = cuboid
 * (params = half edge lengths), 1 = sphere (params[0] = radius), 2 = z-axis cylinder (params = radius, half height);
 * center3 / params3 are HOST pointers.  grid: fnx_grid_bytes(M) scratch; *n_inside (device int32, may be NULL) counts the
 * points that were inside. */
int fnx_rigid_project(void *grid, const float *rigid_xyz, int32_t M, float *xyz, int32_t N, int32_t body,
                      const float *center3_host, const float *params3_host, float r, int32_t max_num_neighbors,
                      int32_t *n_inside, fnx_stream_t stream);
/* velocity = (estimate_xyz - xyz)/secs, zeroed (and xyz kept) where |estimate_xyz - xyz| < 1e-8, else xyz = estimate_xyz. */
int fnx_pbf_confirm_guess(int32_t N, float *xyz, const float *estimate_xyz, float *velocity, float secs, fnx_stream_t stream);
/* visual += secs * sum_j poly6 v_j / max(sum_j poly6, 1e-8) over radius(x=estimate_xyz, y=visual, H, K) (in place). */
int fnx_pbf_update_visual(void *grid, const float *estimate_xyz, const float *velocity, int32_t N, float *visual, int32_t V,
                          int32_t max_num_neighbors, float H, float secs, int32_t *kthV_scratch, fnx_stream_t stream);

/* torch.optim.Adam step on a flat fp32 tensor (gm_fluid.py:349: eps 1e-15); grad is multiplied by grad_scale first
 * (= 1/batch of set_batch_gradient_*, gm_fluid.py:428-430).  step >= 1 is the step count AFTER this update. */
int fnx_adam_step(int64_t n, float *param, const float *grad, float *exp_avg, float *exp_avg_sq, float grad_scale,
                  float lr, float beta1, float beta2, float eps, int32_t step, fnx_stream_t stream);

/* Same update with the step count kept on the device (*step_dev is incremented first; bc_dev is 2 floats of scratch):
 * safe to capture once into a CUDA graph and replay. */
int fnx_adam_step_dev(int64_t n, float *param, const float *grad, float *exp_avg, float *exp_avg_sq, float grad_scale,
                      float lr, float beta1, float beta2, float eps, int32_t *step_dev, float *bc_dev, fnx_stream_t stream);

/* The same, gated: when skip_flag (device int32, may be NULL) is non-zero at execution time the call leaves parameters, moments and
 * *step_dev untouched.  The fused step passes the rasterizer's overflow flag (fnx_raster_overflow_flag): a forward whose instance
 * capacity overflowed renders only the background, its gradient is void, and the update must not be applied. */
int fnx_adam_step_dev_gated(int64_t n, float *param, const float *grad, float *exp_avg, float *exp_avg_sq, float grad_scale,
                            float lr, float beta1, float beta2, float eps, int32_t *step_dev, float *bc_dev,
                            const int32_t *skip_flag, fnx_stream_t stream);

/* torch_scatter.scatter_min(src, index, dim_size=n_out) for 1-D fp32 src and int64 index (gm_fluid.py:1088,1272):
 * out [n_out] (0 for empty groups), arg [n_out] int64 (n for empty groups). */
int fnx_scatter_min(int64_t n, const float *src, const int64_t *index, int32_t n_out, float *out, int64_t *arg,
                    fnx_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Static-background training stage, per-Gaussian kernels  (replaces the activations of
 * gm_background.py:29-37,90-107 and their autograd twins, the scaling regulariser of train_background.py:194-201, the
 * densification statistics of train_background.py:238-243 / gm_background.py:472-476 and torch.optim.Adam over the five
 * parameter groups of gm_background.py:155-168)
 * ---------------------------------------------------------------------------------------------- */
/* scales = exp(raw_scaling) [P,3], opacity = sigmoid(raw_opacity) [P], rotation = raw_rotation / max(|.|, 1e-12) [P,4]. */
int fnx_gs_activate(int32_t P, const float *raw_scaling, const float *raw_opacity, const float *raw_rotation, float *scales,
                    float *opacity, float *rotation, fnx_stream_t stream);

typedef struct fnx_gs_state {   /* raw (pre-activation) parameters, Adam moments, densification statistics; all updated in place */
    float *xyz, *color, *opacity, *scaling, *rotation;          /* [P,3] [P,C] [P] [P,3] [P,4] */
    float *m_xyz, *v_xyz, *m_color, *v_color, *m_opacity, *v_opacity, *m_scaling, *v_scaling, *m_rotation, *v_rotation;
    float *max_radii2D, *xyz_gradient_accum, *denom;            /* [P] each; may be NULL when update_stats == 0 */
} fnx_gs_state;
typedef struct fnx_gs_grads {   /* what fnx_raster_backward produced (w.r.t. the ACTIVATED attributes); any may be NULL = zero */
    const float *dL_dmeans3D, *dL_dmeans2D, *dL_dcolors, *dL_dopacity, *dL_dscales, *dL_drotations;
} fnx_gs_grads;
typedef struct fnx_gs_hparams {
    float lr_xyz, lr_color, lr_opacity, lr_scaling, lr_rotation;
    float beta1, beta2, eps;
    int32_t step;                 /* Adam step count AFTER this update (>= 1) */
    int32_t update_stats;         /* != 0: max_radii2D = max(., radii), xyz_gradient_accum += |dL_dmeans2D.xy|, denom += 1 where radii > 0 */
    float lambda_reg_scaling;     /* > 0: adds lambda * mean(max(s_max/s_min - reg_ratio_threshold, 0)) to the loss being minimised */
    float reg_ratio_threshold;
} fnx_gs_hparams;
/* Chains the gradients through the activations, adds the regulariser's gradient, updates the statistics and applies one
 * Adam step to all five raw tensors -- one launch.  radii [P] int32 of the rendered view; *reg_loss (device scalar, may
 * be NULL) receives the un-weighted regulariser value. */
int fnx_gs_update(int32_t P, int32_t C, const fnx_gs_state *state, const fnx_gs_grads *grads, const fnx_gs_hparams *hp,
                  const int32_t *radii, float *reg_loss, fnx_stream_t stream);

/* Level-two ("visual particle") stage, FD/entries_fluid_nexus/train_visual_particle.py:133-222 (ScalarReal twin :129-218): the
 * positions are fixed; colour / opacity / scales / rotation of the V visual particles are trained (one Adam group each,
 * gm_dynamics.py:380-397) on the image loss + lambda_consistency_X * mse(X[:prev_num], prev_X) on the RAW tensors
 * (l2_loss_consistency, loss_utils.py:138-146) + the scaling regulariser.  One launch chains the rasterizer's gradients through
 * the activations, adds the consistency / regulariser gradients and applies Adam to the fitted tensors.  `state`: only color /
 * opacity / scaling / rotation and their moments are used (color is [V, color_channels]); `grads`: what the rasterizer's backward
 * produced for the V particles, dL_dcolors [V, render_channels] (summed over the channels when the colour parameter has one).
 * losses5 (device, may be NULL) receives {color, opacity, scales, rotation consistency, scaling regulariser}, un-weighted. */
typedef struct fnx_gs_level_two {
    const float *prev_color, *prev_opacity, *prev_scales, *prev_rotation;   /* raw tensors of the previous frame (NULL: no term) */
    int32_t prev_num;               /* rows of the prev_* tensors (particles are appended over time: prev_num <= V) */
    int32_t color_channels;         /* channels of the colour PARAMETER: 1 (grey particles) or 3 */
    int32_t fit_color, fit_opacity, fit_scales, fit_rotation;   /* which tensors are trained (configs: fit_*) */
    float lambda_consistency_color, lambda_consistency_opacity, lambda_consistency_scales, lambda_consistency_rotation;
    float lambda_reg_scaling, reg_ratio_threshold;
    float lr_color, lr_opacity, lr_scaling, lr_rotation;
    float beta1, beta2, eps;
    int32_t step;                   /* Adam step count AFTER this update (>= 1) */
} fnx_gs_level_two;
int fnx_gs_update_level_two(int32_t V, int32_t render_channels, const fnx_gs_state *state, const fnx_gs_grads *grads,
                            const fnx_gs_level_two *hp, float *losses5, fnx_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Fused image loss  (replaces l1_loss + ssim of FD/utils/loss_utils.py:9-64, the grey conversion of
 * FD/entries_fluid_nexus/train_physical_particle.py:356-360 and the weighting of
 * FD/entries_scalar_real/train_physical_particle.py:346-347)
 * ---------------------------------------------------------------------------------------------- */
size_t fnx_image_loss_bytes(int32_t V, int32_t C, int32_t H, int32_t W);
/* img, gt [V,C,H,W].  l1_mean[v] = mean|img-gt|, ssim_mean[v] = mean SSIM (window 11, sigma 1.5, zero padding).
 * If grey != 0 both images are first replaced by their channel mean repeated C times (FluidNexus entries).
 * dL_dimg [V,C,H,W] (may be NULL) = d/dimg of  sum_v ( w_l1 * l1_mean[v] + w_ssim * (1 - ssim_mean[v]) ). */
int fnx_image_loss(int32_t V, int32_t C, int32_t H, int32_t W, const float *img, const float *gt, int32_t grey, float w_l1,
                   float w_ssim, float *dL_dimg, float *l1_mean, float *ssim_mean, void *scratch, fnx_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Section profiler (measurement only): CUDA events recorded around the library's own kernel launches on the
 * launching stream.  Sections: fnx_profile_section_name(0..fnx_profile_sections()-1).
 * ---------------------------------------------------------------------------------------------- */
int fnx_profile_sections(void);
const char *fnx_profile_section_name(int32_t section);
int fnx_profile_enable(uint32_t section_mask);            /* 0 disables */
int fnx_profile_collect(float *total_ms, int32_t *launches); /* arrays of fnx_profile_sections() entries; blocks */

/* Introspection for parity tests: device-to-device copies of the forward's intermediate state.  Any destination may
 * be NULL.  xy [V,P,2], depth [V,P], conic_opacity [V,P,4] (R3 GeometryState means2D/depths/conic_opacity,
 * rasterizer_impl.h:28-46), tiles_touched [V,P] (after tile culling); final_T / n_contrib [V,H,W]
 * (ImageState accum_alpha / n_contrib; n_contrib indexes the culled tile list). */
int fnx_raster_read_geom(const fnx_raster_scratch *scratch, int32_t P, int32_t V, float *xy, float *depth,
                         float *conic_opacity, uint32_t *tiles_touched, fnx_stream_t stream);
int fnx_raster_read_image(const fnx_raster_scratch *scratch, int32_t W, int32_t H, int32_t V, float *final_T,
                          uint32_t *n_contrib, fnx_stream_t stream);
/* Per-tile state [V, tiles] of the last forward (measurement / tests): ranges uint32 [.,2] (begin, end of the tile's
 * span; merged != 0: in the merged stream of fnx_raster_blend_merged), tile_last uint32 [.,4] (records of the span
 * that the blend used, per 8x8 pixel patch of the tile), and for merged streams tile_src (1: blended straight from the static stream) and tile_dyn_last (1 + span
 * index of the last dynamic record: where the backward starts). */
int fnx_raster_read_tiles(const fnx_raster_scratch *scratch, int32_t W, int32_t H, int32_t V, int32_t merged, uint32_t *ranges,
                          uint32_t *tile_last, uint32_t *tile_src, uint32_t *tile_dyn_last, fnx_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FNX_H_INCLUDED */

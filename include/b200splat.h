/*
 * b200splat.h -- C ABI of libb200splat.so, the B200 (sm_100a) Gaussian-splatting rasterizer.
 *
 * Drop-in boundary for the hot path of lizhiqi49/threestudio-3dgs.  The reference calls two
 * un-vendored pip packages (README.md:17-20):
 *     from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
 *                                   (renderer/diff_gaussian_rasterizer.py:8-11, :83-131)
 *     from simple_knn._C import distCUDA2        (geometry/gaussian_base.py:25, :434-437)
 * whose pybind layer binds rasterize_gaussians / rasterize_gaussians_backward / mark_visible /
 * distCUDA2 [UPSTREAM-RECALL: ext.cpp, rasterize_points.cu of ashawkey/diff-gaussian-rasterization;
 * ext.cpp of DSaurus/simple-knn].  Each entry point below names the binding it replaces.
 *
 * Conventions: plain pointers and sizes only (no torch / C++ types); every pointer marked "device"
 * is a CUDA device pointer owned by the caller (PyTorch's caching allocator in the shipped Python
 * host side); every call takes the cudaStream_t to launch on; return value 0 = ok, < 0 = error
 * (b200splat_last_error() gives the message; no exceptions cross the boundary).  All floating
 * point is fp32.  Matrices are the reference's transposed (row-vector) 4x4s, 16 floats,
 * (world_view_transform, full_proj_transform: renderer/gaussian_batch_renderer.py:39-49).
 */
#ifndef B200SPLAT_H
#define B200SPLAT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200SPLAT_ABI_VERSION 7

#define B200SPLAT_OK 0
#define B200SPLAT_ERR_INVALID -1   /* bad argument                                  */
#define B200SPLAT_ERR_CUDA -2      /* a CUDA runtime call or kernel launch failed   */
#define B200SPLAT_ERR_NOMEM -3     /* a caller-provided buffer is too small         */

#define B200SPLAT_MAX_EXTRA 4      /* extra feature channels one raster pass can carry */

typedef void* b200splat_stream; /* cudaStream_t */

/* Growable-buffer callback: replaces upstream's std::function<char*(size_t)> resizeFunctional
 * (rasterize_points.cu).  Must return a device pointer to >= bytes bytes, 256-byte aligned,
 * that stays valid until the matching backward has run; NULL = failure. */
typedef void* (*b200splat_alloc_fn)(void* user, size_t bytes);

/* GaussianRasterizationSettings (renderer/diff_gaussian_rasterizer.py:83-96). */
typedef struct b200splat_camera {
    int32_t image_height;
    int32_t image_width;
    float tanfovx;
    float tanfovy;
    float scale_modifier;
    int32_t sh_degree;        /* active degree; clamped to what M coefficients hold */
    int32_t prefiltered;
    int32_t debug;            /* !=0: synchronise + check after every kernel        */
    const float* bg;          /* device, 3  */
    const float* viewmatrix;  /* device, 16 */
    const float* projmatrix;  /* device, 16 */
    const float* campos;      /* device, 3  */
    /* Optional: the camera's derived scalars already on the device, 4 floats (focal_x = W / (2 tanfovx),
     * focal_y = H / (2 tanfovy), limx = 1.3 tanfovx, limy = 1.3 tanfovy).  When set, tanfovx / tanfovy above are
     * ignored: a caller whose field of view lives in a device tensor (the reference's batch["fovy"],
     * renderer/diff_gaussian_rasterizer.py:80-81 reads it back with math.tan(tensor) -- one D2H sync per view) never
     * has to bring it to the host.  NULL: computed from tanfovx / tanfovy on the host. */
    const float* scalars_dev;
} b200splat_camera;

/* ---- buffer sizing (bytes) --------------------------------------------------------------- */
size_t b200splat_geom_bytes(int32_t P);                  /* per-Gaussian state kept for backward */
size_t b200splat_image_bytes(int32_t H, int32_t W);      /* tile ranges + n_contrib + final T    */
size_t b200splat_binning_bytes(int64_t num_rendered);    /* key/value ping-pong + sort scratch   */
size_t b200splat_backward_scratch_bytes(int32_t P);      /* packed 2-D stage gradients (+ extra channels) */

/* ---- forward: replaces rasterize_gaussians (RasterizeGaussiansCUDA) ------------------------
 * Inputs (device): means3D (P,3); shs (P,M,3) or NULL; colors_precomp (P,3) or NULL; opacities
 * (P,1); scales (P,3) + rotations (P,4) (r,x,y,z) or cov3D_precomp (P,6).  Outputs (device):
 * out_color (3,H,W), out_depth (1,H,W), out_alpha (1,H,W), radii (P) int32.
 * geom/image buffers are caller-allocated (sizes above).  The binning buffer depends on
 * num_rendered = sum(tiles_touched), known only after the scan: either pass binning_buffer with
 * binning_bytes large enough, or pass binning_alloc and the library calls it once with the exact
 * size (one stream synchronise, as upstream).  *num_rendered_out (host) receives num_rendered,
 * *binning_out (host) the binning pointer actually used. */
typedef struct b200splat_forward_args {
    b200splat_camera cam;
    int32_t P;
    int32_t M; /* SH coefficients per channel present in shs (0 when shs == NULL) */
    const float* means3D;
    const float* shs;
    const float* colors_precomp;
    const float* opacities;
    const float* scales;
    const float* rotations;
    const float* cov3D_precomp;
    float* out_color;
    float* out_depth;
    float* out_alpha;
    int32_t* radii;
    void* geom_buffer;
    size_t geom_bytes;
    void* image_buffer;
    size_t image_bytes;
    void* binning_buffer;
    size_t binning_bytes;
    b200splat_alloc_fn binning_alloc;
    void* alloc_user;
    b200splat_stream stream;
    int64_t* num_rendered_out;
    void** binning_out;
    /* Optional extra feature channels rendered by the same pass (new, non-breaking): extra_features (P,n_extra)
     * fp32, n_extra in 0..B200SPLAT_MAX_EXTRA; out_extra (n_extra,H,W) = sum_i e_i alpha_i T_i (no background
     * term).  Replaces the reference's second rasterizer call with the per-Gaussian normals as colours
     * (renderer/diff_gaussian_rasterizer_shading.py:177-187, ..._normal.py:175-185): same alphas, one pass. */
    const float* extra_features;
    int32_t n_extra;
    float* out_extra;
} b200splat_forward_args;

int b200splat_forward(const b200splat_forward_args* args);

/* ---- backward: replaces rasterize_gaussians_backward (RasterizeGaussiansBackwardCUDA) -------
 * Takes the forward's inputs, radii and the three buffers unchanged (they are read
 * only, so backward may run twice on one forward: system/gaussian_splatting.py:129,137-138),
 * plus dL/dout_color (3,H,W), dL/dout_depth (1,H,W), dL/dout_alpha (1,H,W) (contiguous; any may
 * be NULL = zeros).  Writes dense gradients for every Gaussian (zeros where radii == 0):
 * dL_dmeans3D (P,3), dL_dmeans2D (P,3; NDC units, z = 0; geometry/gaussian_base.py:815-819 reads
 * it), dL_dopacity (P,1), and, when the matching input was given, dL_dshs (P,M,3),
 * dL_dcolors (P,3), dL_dscales (P,3), dL_drotations (P,4), dL_dcov3D (P,6).  When accumulate != 0
 * the results are ADDED to the output tensors instead of overwriting them (multi-view batches).
 */
typedef struct b200splat_backward_args {
    b200splat_camera cam;
    int32_t P;
    int32_t M;
    int64_t num_rendered;
    const float* means3D;
    const float* shs;
    const float* colors_precomp;
    const float* opacities;
    const float* scales;
    const float* rotations;
    const float* cov3D_precomp;
    const int32_t* radii;
    const float* out_alpha; /* accepted for signature parity with upstream, not read: the exact final
                               transmittance is kept in image_buffer (1 - alpha cancels catastrophically) */
    const void* geom_buffer;
    const void* binning_buffer;
    const void* image_buffer;
    const float* dL_dout_color;
    const float* dL_dout_depth;
    const float* dL_dout_alpha;
    float* dL_dmeans3D;
    float* dL_dmeans2D;
    float* dL_dshs;
    float* dL_dcolors;
    float* dL_dopacity;
    float* dL_dscales;
    float* dL_drotations;
    float* dL_dcov3D;
    void* scratch;
    size_t scratch_bytes;
    int32_t accumulate;
    /* Optional fused densification statistics (geometry/gaussian_base.py:815-819, :846-851), each (P) fp32 or
     * NULL; updated for Gaussians with radii > 0 of THIS view:  stat_grad_accum += ||dL_dmeans2D.xy||,
     * stat_denom += 1, stat_max_radii = max(stat_max_radii, radii). */
    float* stat_grad_accum;
    float* stat_denom;
    float* stat_max_radii;
    b200splat_stream stream;
    /* extra feature channels of the forward (same extra_features / n_extra): dL_dout_extra (n_extra,H,W) or NULL,
     * dL_dextra (P,n_extra) written (added when accumulate != 0); the geometry gradients include the extra
     * channels' contribution through alpha -- except dL_dmeans2D and stat_grad_accum, which see the colour / depth /
     * alpha terms only: the reference renders extra channels with a second rasterizer call whose means2D is a
     * gradient-free zeros tensor (renderer/diff_gaussian_rasterizer_shading.py:177-187) */
    const float* extra_features;
    int32_t n_extra;
    const float* dL_dout_extra;
    float* dL_dextra;
} b200splat_backward_args;

int b200splat_backward(const b200splat_backward_args* args);

/* ---- view-batched forward / backward ------------------------------------------------------------
 * The reference renders the B views of a step one after the other with the same Gaussians
 * (renderer/gaussian_batch_renderer.py:21-54).  These entry points take up to B200SPLAT_MAX_VIEWS views at
 * once: the Gaussian parameters are read once per phase for all views, every phase of the pipeline is ONE
 * launch for the whole batch (so the long per-tile tails of one view are filled by the others), and
 * num_rendered never visits the host -- each view's binning buffer has a caller-chosen capacity
 * (b200splat_binning_capacity(binning_bytes) pairs); if a view needs more, its result is invalid and
 * overflow_out[v] (sync != 0) / the status word (b200splat_forward_views_get) says so: retry with a larger
 * buffer.  All views share P, M, the image size, scale_modifier and sh_degree (taken from cams[0]).
 * Arrays of V device pointers live on the host. */
#define B200SPLAT_MAX_VIEWS 8
int64_t b200splat_binning_capacity(size_t binning_bytes);

typedef struct b200splat_batch_forward_args {
    int32_t V;
    const b200splat_camera* cams; /* V */
    int32_t P;
    int32_t M;
    const float* means3D;
    const float* shs;
    const float* colors_precomp;
    const float* opacities;
    const float* scales;
    const float* rotations;
    float* const* out_color;      /* V x (3,H,W) */
    float* const* out_depth;
    float* const* out_alpha;
    int32_t* const* radii;        /* V x (P) */
    void* const* geom_buffer;     /* V, each >= b200splat_geom_bytes(P) */
    void* const* image_buffer;    /* V, each >= b200splat_image_bytes(H, W) */
    void* const* binning_buffer;  /* V, each binning_bytes */
    size_t binning_bytes;
    b200splat_stream stream;
    int32_t sync;                 /* != 0: synchronise the stream and fill the two host arrays below */
    int64_t* num_rendered_out;    /* V */
    int32_t* overflow_out;        /* V */
    const float* extra_features;  /* (P,n_extra) or NULL: see b200splat_forward_args */
    int32_t n_extra;
    float* const* out_extra;      /* V x (n_extra,H,W) */
    /* Optional early overflow notice without a stream synchronisation: V 64-bit words of PINNED HOST memory
     * (device-accessible by the same pointer: cudaHostAlloc / torch pin_memory).  A small reduction right behind the
     * preprocess kernel -- the first point of the forward at which a view's pair count is known -- stores
     * (notify_epoch << 32) | num_rendered of view v into pairs_notify[v] while the depth sort, the scan, the key
     * duplication, the tile partition and the render are still queued behind it; the host polls the word
     * until it carries its epoch and compares the count with the capacity (count > capacity: the view's result is
     * invalid, re-run with a larger buffer).  NULL: no notice. */
    uint64_t* pairs_notify;
    uint32_t notify_epoch;
} b200splat_batch_forward_args;

int b200splat_forward_batched(const b200splat_batch_forward_args* args);

/* Gradients of the parameters are summed over the V views (plus the previous contents when
 * accumulate != 0); dL_dmeans2D (optional) is per view, as are the upstream pixel gradients. */
typedef struct b200splat_batch_backward_args {
    int32_t V;
    const b200splat_camera* cams;
    int32_t P;
    int32_t M;
    const float* means3D;
    const float* shs;
    const float* colors_precomp;
    const float* opacities;
    const float* scales;
    const float* rotations;
    const int32_t* const* radii;
    const void* const* geom_buffer;
    const void* const* image_buffer;
    const void* const* binning_buffer;
    size_t binning_bytes;
    const float* const* dL_dout_color; /* V entries, any may be NULL */
    const float* const* dL_dout_depth;
    const float* const* dL_dout_alpha;
    float* const* dL_dmeans2D;         /* NULL or V entries (each NULL or (P,3)) */
    float* dL_dmeans3D;
    float* dL_dshs;
    float* dL_dcolors;
    float* dL_dopacity;
    float* dL_dscales;
    float* dL_drotations;
    void* const* scratch;              /* V, each >= b200splat_backward_scratch_bytes(P) */
    int32_t accumulate;
    /* != 0: the caller guarantees every scratch buffer is all-zero on entry (e.g. a persistent workspace
     * allocated zeroed); the library then skips its own clear and leaves the buffers all-zero on return
     * (the consumer kernel zeroes each record it has read), so the next call can pass 1 again */
    int32_t scratch_clean;
    /* phase: 0 = whole backward; 1 = render backward only (pixel gradients -> per-view 2-D gradient records);
     * 2 = preprocess backward only, for the Gaussians [g_begin, g_end) (g_end <= 0: P) -- lets the caller exchange
     * the finished gradients of one Gaussian range with its peers while the next range is computed */
    int32_t phase;
    int32_t g_begin;
    int32_t g_end;
    float* stat_grad_accum;
    float* stat_denom;
    float* stat_max_radii;
    b200splat_stream stream;
    const float* extra_features;          /* see b200splat_backward_args */
    int32_t n_extra;
    const float* const* dL_dout_extra;    /* NULL or V entries, any may be NULL */
    float* dL_dextra;                     /* (P,n_extra), summed over the views */
    /* optional (ABI 7): one byte per Gaussian of [g_begin, g_end), 1 if the Gaussian received a gradient in at least
     * one view of this call (its parameter-gradient rows can be non-zero), else 0 (its rows are exactly zero) -- the
     * live map of the row-sparse exchange, b200splat_p2p_args.seg_row_floats */
    uint8_t* live_map;
} b200splat_batch_backward_args;

int b200splat_backward_batched(const b200splat_batch_backward_args* args);

/* ---- fused per-pixel post-ops of the renderer variants (forward + backward), for V views at once -------------
 * What the reference does in ~25 PyTorch kernels per view after the rasterizer call:
 *   PLAIN       render = clamp(image, 0, 1)                          (renderer/diff_gaussian_rasterizer_advanced.py:139-146)
 *   BACKGROUND  render = clamp(image + (1 - alpha) * bg)             (..._background.py:130-141)
 *   NORMAL      + normal = normalize(Depth2Normal(rays_o + depth * rays_d)) * 0.5 * alpha + 0.5, with the
 *               gradients of normal and depth cut where alpha <= 0.99 (..._normal.py:172-201)
 *   SHADING     + point-light Lambert shading of albedo = image / (alpha + 1e-6) and the composite
 *               fg * alpha + (1 - alpha) * bg (..._shading.py:174-213; material/gaussian_material.py:70-104);
 *               pred_normal (the rendered per-Gaussian normals, used detached) replaces the depth normal
 *               as the shading normal when given (..._shading.py:195-197).
 * Layouts: image / pred_normal / render / normal / their gradients (V,3,H,W); depth, alpha (V,1,H,W);
 * rays_o, rays_d, bg, d_bg (V,H,W,3); light (V,3); all device fp32, contiguous.  Backward recomputes the forward
 * from the same inputs; g_* may be NULL (= zeros); d_bg may be NULL. */
#define B200SPLAT_POST_PLAIN 0
#define B200SPLAT_POST_BACKGROUND 1
#define B200SPLAT_POST_NORMAL 2
#define B200SPLAT_POST_SHADING 3
#define B200SPLAT_SHADE_ALBEDO 0
#define B200SPLAT_SHADE_TEXTURELESS 1
#define B200SPLAT_SHADE_DIFFUSE 2
typedef struct b200splat_postprocess_args {
    int32_t V;
    int32_t H;
    int32_t W;
    int32_t mode;    /* B200SPLAT_POST_*  */
    int32_t shading; /* B200SPLAT_SHADE_* (mode SHADING) */
    const float* image;
    const float* depth;
    const float* alpha;
    const float* rays_o;
    const float* rays_d;
    const float* bg;
    const float* light;
    const float* pred_normal;
    float ambient[3]; /* material ambient_light_color */
    float diffuse[3]; /* material diffuse_light_color */
    /* forward outputs */
    float* render;
    float* normal;    /* modes NORMAL, SHADING */
    float* depth_out; /* optional copy of depth (the output whose gradient is masked) */
    /* backward: upstream gradients in, gradients of the rasterizer outputs and of bg out */
    const float* g_render;
    const float* g_normal;
    const float* g_depth;
    float* d_image;
    float* d_depth;
    float* d_alpha;
    float* d_bg;
    void* scratch; /* modes NORMAL, SHADING: >= b200splat_postprocess_scratch_bytes(V,H,W) */
    size_t scratch_bytes;
    b200splat_stream stream;
    const int32_t* shading_per_view; /* HOST array of V B200SPLAT_SHADE_* (V <= 64) overriding `shading`, or NULL:
                                        the reference's material draws the mode per view in training */
    const float* lights_per_view;    /* HOST array (V,6) = (ambient rgb, diffuse rgb) per view (V <= 64) overriding
                                        ambient / diffuse, or NULL: the material's soft_shading draws a new ambient
                                        ratio per view in training (material/gaussian_material.py:58-63) */
} b200splat_postprocess_args;
size_t b200splat_postprocess_scratch_bytes(int32_t V, int32_t H, int32_t W);
int b200splat_postprocess_forward(const b200splat_postprocess_args* args);
int b200splat_postprocess_backward(const b200splat_postprocess_args* args);

/* ---- fused activation-backward + Adam step on the rasterizer's gradient buffer --------------------------------
 * The reference optimises RAW parameters with torch.optim.Adam(lr=0, eps=1e-15), one group per tensor
 * (geometry/gaussian_base.py:470-525; per-step lr schedule :539-572), through the activations exp (scaling),
 * sigmoid (opacity), F.normalize (rotation) and clip(-color_clip, color_clip) (features_dc)
 * (geometry/gaussian_base.py:240-248, :371-400).  This entry point takes the gradients with respect to the
 * ACTIVATED values -- the backward's outputs, i.e. the fields of the packed all-reduce buffer -- applies the
 * activation Jacobians and one Adam step (amsgrad off, no weight decay) to the raw parameters and their
 * exp_avg (m_*) / exp_avg_sq (v_*) states, in place.  lr order: xyz, f_dc, f_rest, opacity, scaling, rotation.
 * step is the 1-based step count of this update.  rotation and g_rotations must be 16-byte aligned. */
typedef struct b200splat_adam_args {
    int32_t P;
    int32_t M; /* SH coefficients per channel of g_shs: features_dc holds 1, features_rest M-1 */
    float* xyz;           /* (P,3)     */
    float* features_dc;   /* (P,1,3)   */
    float* features_rest; /* (P,M-1,3), NULL when M == 1 */
    float* opacity;       /* (P,1) raw */
    float* scaling;       /* (P,3) raw */
    float* rotation;      /* (P,4) raw */
    float* m_xyz;
    float* m_features_dc;
    float* m_features_rest;
    float* m_opacity;
    float* m_scaling;
    float* m_rotation;
    float* v_xyz;
    float* v_features_dc;
    float* v_features_rest;
    float* v_opacity;
    float* v_scaling;
    float* v_rotation;
    const float* g_means3D;   /* (P,3)   */
    const float* g_shs;       /* (P,M,3) */
    const float* g_opacities; /* (P,1)   */
    const float* g_scales;    /* (P,3)   */
    const float* g_rotations; /* (P,4)   */
    double lr[6]; /* doubles, like the Python floats torch.optim.Adam computes 1 - beta and lr / (1 - beta^t) from */
    double beta1;
    double beta2;
    float eps;
    float color_clip;
    int32_t step;
    b200splat_stream stream;
} b200splat_adam_args;
int b200splat_adam_step(const b200splat_adam_args* args);

/* ---- mark_visible: replaces markVisible (GaussianRasterizer.markVisible) ------------------- */
int b200splat_mark_visible(int32_t P, const float* means3D, const float* viewmatrix,
                           const float* projmatrix, uint8_t* present, b200splat_stream stream);

/* ---- distCUDA2: replaces simple_knn._C.distCUDA2 (geometry/gaussian_base.py:434-437) ---------
 * out[i] = mean squared distance from point i to its 3 nearest other points (exact).
 * workspace: >= b200splat_dist2_workspace_bytes(P) device bytes. */
size_t b200splat_dist2_workspace_bytes(int32_t P);
int b200splat_dist2(int32_t P, const float* points, float* out, void* workspace,
                    size_t workspace_bytes, b200splat_stream stream);

/* ---- stage-level entry points (parity tests and reuse) ------------------------------------- */
/* Stable LSD radix sort of (u64 key, u32 value) pairs over key bits [0, end_bit); onesweep.
 * workspace >= b200splat_sort_workspace_bytes(n).  keys_alt/vals_alt are ping-pong storage;
 * *result_in_alt (host) is set to 1 when the sorted data ended up in the alt arrays. */
size_t b200splat_sort_workspace_bytes(int64_t n);
int b200splat_sort_pairs(int64_t n, int32_t end_bit, uint64_t* keys, uint32_t* vals,
                         uint64_t* keys_alt, uint32_t* vals_alt, void* workspace,
                         size_t workspace_bytes, int32_t* result_in_alt, b200splat_stream stream);
/* Inclusive prefix sum of n uint32 (decoupled look-back). workspace >= ..._scan_workspace_bytes. */
size_t b200splat_scan_workspace_bytes(int64_t n);
int b200splat_inclusive_scan_u32(int64_t n, const uint32_t* in, uint32_t* out, void* workspace,
                                 size_t workspace_bytes, b200splat_stream stream);

/* Views into the buffers of a finished forward (device pointers into the caller's buffers), for
 * the bit-exact parity checks: tiles_touched (P) u32, point_offsets (P) u32, depths (P) f32,
 * sorted pair words (R) u64, ranges (T,2) u32, n_contrib (H*W) u32 (1-based index of the
 * last blended list entry), n_visited (H*W) u32 (list entries traversed before the pixel stopped). */
typedef struct b200splat_forward_views {
    const uint32_t* tiles_touched;
    const uint32_t* point_offsets;
    const float* depths;
    const float* gauss2d; /* (P,12): x, y, conic a, conic b | conic c, opacity, depth, r | g, b, -, - */
    const float* cov3D;   /* always NULL: Sigma3 is recomputed in backward instead of stored */
    const uint64_t* keys_sorted;
    const uint32_t* point_list;
    const uint32_t* ranges;
    const uint32_t* n_contrib;
    const uint32_t* n_visited;
    const uint32_t* status; /* [0] != 0: binning capacity overflow */
    /* keys_sorted[i] holds the sorted pair WORD (tile_id << packed_idx_bits) | gaussian_index (packed_idx_bits = 32);
     * upstream's 64-bit key of entry i is (tile_id << 32) | float_bits(depths[gaussian_index]); point_list is unused */
    int64_t packed_idx_bits;
    const uint64_t* gaussian_order; /* (P) words (depth_bits << 32 | index) sorted by depth: the per-view depth order */
} b200splat_forward_views;

int b200splat_forward_views_get(int32_t P, int32_t H, int32_t W, int64_t num_rendered,
                                const void* geom_buffer, const void* binning_buffer,
                                const void* image_buffer, b200splat_forward_views* out);

/* ---- per-kernel timing (bench.py roofline): CUDA events recorded on the launch stream around each
 * kernel family while enabled.  Families: 0 preprocess, 1 scan, 2 duplicateWithKeys, 3 sort,
 * 4 tile ranges, 5 render fwd, 6 render bwd, 7 preprocess bwd, 8 dist2. */
#define B200SPLAT_NUM_FAMILIES 9
int b200splat_profile_enable(int32_t on);
/* Synchronises the recorded events; fills ms[f] (total) and count[f] (launch groups) per family and
 * clears the record.  Arrays must hold B200SPLAT_NUM_FAMILIES entries. */
int b200splat_profile_read(float* ms, int64_t* count);

/* ---- all-reduce over NVLink peer memory (new capability: the reference is single-GPU) ----------
 * One process per GPU.  Each rank allocates its exchange buffer and its signal words with
 * b200splat_p2p_alloc (cudaMalloc + CUDA IPC export, zero-filled), sends the 64-byte handles to its peers
 * (any host channel: torch.distributed all_gather_object), maps the peers' with b200splat_p2p_open, and then
 * calls b200splat_p2p_allreduce on its own stream: every listed segment of every rank's buffer becomes its sum (or
 * maximum) over the ranks, in place, bit-identical on all ranks.  Segment offsets / counts are multiples of 4 floats;
 * epoch must increase by 1 per call, equally on all ranks (calls of one group are stream-ordered on each rank);
 * epoch = 0: the library keeps the call counter itself, in the rank's signal words -- no launch argument changes
 * from call to call, so the exchange can be captured in a CUDA graph and replayed (do not mix the two modes).
 * Returns B200SPLAT_OK immediately (stream-ordered); b200splat_p2p_error reads the rank's error flag
 * (peer timeout) -- it synchronises nothing itself, call it after the stream is known to be idle. */
#define B200SPLAT_P2P_MAX_RANKS 8
#define B200SPLAT_P2P_HANDLE_BYTES 64
#define B200SPLAT_P2P_SIGNAL_BYTES 256
#define B200SPLAT_P2P_MAX_SEGMENTS 16
#define B200SPLAT_P2P_SUM 0
#define B200SPLAT_P2P_MAX 1
typedef struct b200splat_p2p_args {
    int32_t rank, world;
    void* bufs[B200SPLAT_P2P_MAX_RANKS];    /* device pointers valid on THIS rank: own buffer + mapped peers */
    void* signals[B200SPLAT_P2P_MAX_RANKS]; /* same for the B200SPLAT_P2P_SIGNAL_BYTES signal areas */
    int32_t n_segments;                     /* 1..16 disjoint float ranges of the buffer, reduced in one kernel */
    int64_t seg_offset[B200SPLAT_P2P_MAX_SEGMENTS]; /* in floats, multiple of 4 */
    int64_t seg_count[B200SPLAT_P2P_MAX_SEGMENTS];  /* in floats, multiple of 4 */
    int32_t seg_op[B200SPLAT_P2P_MAX_SEGMENTS];     /* B200SPLAT_P2P_SUM | B200SPLAT_P2P_MAX */
    uint32_t epoch;
    b200splat_stream stream;
    /* Row-sparse SUM segments (ABI 7).  seg_row_floats[i] > 0 (a multiple of 4): segment i is an array of rows of that
     * many floats, row j belongs to index seg_row0[i] + j (a Gaussian), and a row is exchanged only if the byte of its
     * index in the LIVE MAP is non-zero on at least one rank; every other row must be all zeros on every rank (it is
     * left alone).  The live map is one byte per index at byte offset live_offset of every rank's buffer (outside the
     * exchanged segments); b200splat_backward_batched writes it (live_map).  seg_row0 is a multiple of 4.
     * All zero: every segment is dense. */
    int32_t seg_row_floats[B200SPLAT_P2P_MAX_SEGMENTS];
    int64_t seg_row0[B200SPLAT_P2P_MAX_SEGMENTS];
    int64_t live_offset;
} b200splat_p2p_args;
int b200splat_p2p_alloc(size_t bytes, void** ptr, void* handle_out);
int b200splat_p2p_open(const void* handle, void** ptr);
int b200splat_p2p_close(void* mapped_ptr);
int b200splat_p2p_free(void* ptr);
int b200splat_p2p_allreduce(const b200splat_p2p_args* args);
int b200splat_p2p_error(const void* own_signals, int32_t* flag_out);

/* ---- all-reduce through NVSwitch multicast (in-switch reduction) -----------------------------------------------
 * Same contract as b200splat_p2p_allreduce (segments, ops, epoch, in place, bit-identical on all ranks), but the data
 * plane is the switch: every rank's buffer is bound to ONE multicast object (CUDA VMM multicast; the shipped Python
 * host side gets it from torch.distributed._symmetric_memory), rank r pulls its slice of every segment with
 * multimem.ld_reduce (the switch adds / maxes the N copies on the way) and pushes the result to all ranks with
 * multimem.st (the switch replicates it): N/N of the buffer per NVLink direction per GPU instead of 2 (N-1)/N.
 * mc_buffer: the multicast address of the buffer; signals[k]: rank k's signal words, peer-mapped as for
 * b200splat_p2p_allreduce (B200SPLAT_P2P_SIGNAL_BYTES each, zero before the first call).  MAX segments hold
 * non-negative floats (radii): they are reduced as unsigned integers (same order), exactly. */
typedef struct b200splat_mc_args {
    int32_t rank, world;
    void* mc_buffer;
    void* signals[B200SPLAT_P2P_MAX_RANKS];
    int32_t n_segments;
    int64_t seg_offset[B200SPLAT_P2P_MAX_SEGMENTS]; /* in floats, multiple of 4 */
    int64_t seg_count[B200SPLAT_P2P_MAX_SEGMENTS];  /* in floats, multiple of 4 */
    int32_t seg_op[B200SPLAT_P2P_MAX_SEGMENTS];     /* B200SPLAT_P2P_SUM | B200SPLAT_P2P_MAX */
    uint32_t epoch;
    b200splat_stream stream;
    /* row-sparse SUM segments, as in b200splat_p2p_args (the union of the ranks' live maps is read through the
     * multicast address: one multimem.ld_reduce.add.u32 sums four indices' bytes of all ranks) */
    int32_t seg_row_floats[B200SPLAT_P2P_MAX_SEGMENTS];
    int64_t seg_row0[B200SPLAT_P2P_MAX_SEGMENTS];
    int64_t live_offset;
} b200splat_mc_args;
int b200splat_mc_allreduce(const b200splat_mc_args* args);

/* ---- misc ----------------------------------------------------------------------------------- */
/* How the render kernels stage a tile's Gaussian records into shared memory: 0 = per-entry 16-byte cp.async
 * (LDGSTS) arriving on the stage's mbarrier, 1 = cp.async.bulk (UBLKCP) with complete_tx on the mbarrier,
 * -1 = the default (environment variable B200SPLAT_STAGING = "ldgsts" | "bulk", else the measured-faster one).
 * Process-wide; takes effect at the next launch.  Returns the previous setting. */
int b200splat_set_staging(int32_t mode);
int b200splat_abi_version(void);
const char* b200splat_last_error(void);
/* Number of kernels this library launched since process start (all streams). */
uint64_t b200splat_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* B200SPLAT_H */

// Shared definitions of the sm_100a splatting kernels (internal; the public surface is
// include/b200splat.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace b200splat {

constexpr int BLOCK_X = 16;
constexpr int BLOCK_Y = 16;
constexpr int BLOCK_SIZE = BLOCK_X * BLOCK_Y;
constexpr int NUM_SMS = 148;

constexpr float NEAR_CULL = 0.2f;
constexpr float FOV_CLAMP = 1.3f;
constexpr float DILATION = 0.3f;
constexpr float LAMBDA_FLOOR = 0.1f;
constexpr float ALPHA_MAX = 0.99f;
constexpr float ALPHA_MIN = 1.0f / 255.0f;
constexpr float T_MIN = 0.0001f;
constexpr float PW_EPS = 0.0000001f;

// per-Gaussian 2-D record written by preprocess and gathered by the render kernels (48 B, 16 B aligned)
//   q0 = (x, y, conic_a, conic_b)   q1 = (conic_c, opacity, depth, r)   q2 = (g, b, cull_thr, 0)
//   cull_thr = 2 ln(255 opacity) + margin: the largest value of the conic quadratic at which alpha >= 1/255
constexpr int REC_FLOATS = 12;
// packed 2-D stage gradients accumulated by render-backward (48 B per Gaussian)
//   g0 = (dx, dy, dconic_a, dconic_b)  g1 = (dconic_c, dopacity, dr, dg)  g2 = (db, ddepth, 0, 0)
constexpr int GRAD2D_FLOATS = 12;

struct CameraParams {
    int H, W;
    int grid_x, grid_y;
    float tanfovx, tanfovy;
    float focal_x, focal_y;
    float limx, limy;
    float scale_modifier;
    int sh_degree;   // effective (clamped) degree
    int M;           // coefficients per channel in shs
    const float* bg;
    const float* view;
    const float* proj;
    const float* campos;
};

// ---- geometry buffer layout (per Gaussian, SoA) ---------------------------------------------
struct GeomViews {
    float* rec;              // P * 12
    float* depths;           // P
    float* cov3D;            // P * 6
    uint8_t* clamped;        // P   (bit c set: channel c clamped at 0)
    uint32_t* tiles_touched; // P
    uint32_t* point_offsets; // P
    void* scan_ws;
    size_t scan_ws_bytes;
};
struct BinningViews {
    uint64_t* keys[2];
    uint32_t* vals[2];
    int32_t* final_sel;      // device int: which of the two holds the sorted result (host mirrors)
    void* sort_ws;
    size_t sort_ws_bytes;
};
struct ImageViews {
    uint32_t* ranges;        // T * 2
    uint32_t* n_contrib;     // H * W
    float* final_T;          // H * W  (transmittance after the last blended entry; 1 - alpha loses bits)
    uint32_t* n_visited;     // H * W  (list entries traversed by the pixel in forward)
    uint32_t* tile_order;    // T      (tiles by decreasing list length: CTA i renders tile_order[i])
};

size_t geom_layout(int P, void* base, GeomViews* v);
size_t binning_layout(int64_t R, void* base, BinningViews* v);
size_t image_layout(int H, int W, void* base, ImageViews* v);

// ---- launchers (each returns cudaError_t from the launch) -----------------------------------
cudaError_t launch_preprocess(int P, const CameraParams& cam, const float* means3D, const float* scales,
                              const float* rotations, const float* opacities, const float* shs,
                              const float* colors_precomp, const float* cov3D_precomp, int32_t* radii,
                              const GeomViews& g, cudaStream_t st);
cudaError_t launch_duplicate(int P, const CameraParams& cam, const int32_t* radii, const GeomViews& g,
                             uint64_t* keys, uint32_t* vals, uint32_t* hist, int end_bit, cudaStream_t st);
cudaError_t launch_mark_visible(int P, const float* means3D, const float* view, const float* proj,
                                uint8_t* present, cudaStream_t st);

size_t scan_workspace_bytes(int64_t n);
cudaError_t launch_inclusive_scan(int64_t n, const uint32_t* in, uint32_t* out, void* ws, cudaStream_t st);
size_t sort_workspace_bytes(int64_t n);
// returns index (0/1) of the buffer pair holding the result through *sel
cudaError_t launch_sort_pairs(int64_t n, int end_bit, uint64_t* keys[2], uint32_t* vals[2], void* ws,
                              int* sel, cudaStream_t st, bool hist_ready = false);
// zero the sort workspace / where a fused producer (duplicateWithKeys) accumulates the digit histograms
cudaError_t sort_prepare(int64_t n, int end_bit, void* ws, cudaStream_t st);
uint32_t* sort_histogram_ptr(void* ws);
cudaError_t launch_tile_ranges(int64_t R, int T, const uint64_t* keys_sorted, uint32_t* ranges, cudaStream_t st);
cudaError_t launch_tile_order(int T, const uint32_t* ranges, uint32_t* order, cudaStream_t st);

cudaError_t launch_render_forward(const CameraParams& cam, const uint32_t* ranges, const uint32_t* tile_order,
                                  const uint32_t* point_list,
                                  const float* rec, uint32_t* n_contrib, uint32_t* n_visited, float* final_T,
                                  float* out_color, float* out_depth, float* out_alpha, cudaStream_t st);
cudaError_t launch_render_backward(const CameraParams& cam, const uint32_t* ranges, const uint32_t* tile_order,
                                   const uint32_t* point_list,
                                   const float* rec, const uint32_t* n_contrib, const float* final_T,
                                   const float* dL_dcolor, const float* dL_ddepth, const float* dL_dalpha,
                                   float* grad2d, cudaStream_t st);
cudaError_t launch_preprocess_backward(int P, const CameraParams& cam, const float* means3D, const float* scales,
                                       const float* rotations, const float* shs, const float* cov3D_precomp,
                                       const int32_t* radii, const GeomViews& g, const float* grad2d,
                                       float* dL_dmeans3D, float* dL_dmeans2D, float* dL_dshs, float* dL_dcolors,
                                       float* dL_dopacity, float* dL_dscales, float* dL_drotations, float* dL_dcov3D,
                                       float* stat_grad_accum, float* stat_denom, float* stat_max_radii,
                                       int accumulate, cudaStream_t st);

size_t dist2_workspace_bytes(int P);
cudaError_t launch_dist2(int P, const float* points, float* out, void* ws, cudaStream_t st);

void count_launch(int n = 1);

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace b200splat

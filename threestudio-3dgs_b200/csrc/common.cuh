// Shared definitions of the sm_100a splatting kernels (internal; the public surface is
// include/b200splat.h).
//
// Every kernel of the pipeline takes ONE BatchTab by value (__grid_constant__): the per-view device
// pointers of up to MAX_VIEWS views that share the Gaussian parameters and the image size.  A phase of
// the pipeline (preprocess, scan, key duplication, each sort pass, tile ranges, render, ...) is ONE
// launch for the whole view batch (blockIdx.y or the tile-order entry selects the view); the single-view
// entry points are the V = 1 case of the same kernels.  num_rendered never has to visit the host: kernels
// read it from point_offsets[P-1] and are launched over the binning *capacity*.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace b200splat {

constexpr int BLOCK_X = 16;
constexpr int BLOCK_Y = 16;
constexpr int BLOCK_SIZE = BLOCK_X * BLOCK_Y;
constexpr int NUM_SMS = 148;
constexpr int MAX_VIEWS = 8;

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
//   g0 = (dx, dy, 2 dconic_a, dconic_b)  g1 = (2 dconic_c, dopacity, dr, dg)  g2 = (db, ddepth, 0, 0)
//   (the consumer applies the factor 1/2 of the two diagonal conic terms)
constexpr int GRAD2D_FLOATS = 12;

constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int MAX_PASSES = 8;

// status words of a view (in its image buffer)
constexpr int STATUS_OVERFLOW = 0;   // != 0: num_rendered exceeded the binning capacity (results invalid)
constexpr int STATUS_WORDS = 4;
// buckets of the longest-first work orders (tile order of render forward, block order of render backward): a monotone
// 11-bit key of the list length (float exponent + 3 mantissa bits: 8 buckets per octave)
constexpr int ORDER_BUCKETS = 2048;
constexpr uint32_t BLOCK_CODE_NONE = 0xffffffffu;
__host__ __device__ __forceinline__ uint32_t order_bucket(uint32_t len) {   // descending: longest -> smallest bucket index
#ifdef __CUDA_ARCH__
    const uint32_t k = __float_as_uint((float)len) >> 20;
#else
    const float f = (float)len;
    uint32_t k;
    memcpy(&k, &f, 4);
    k >>= 20;
#endif
    return (uint32_t)(ORDER_BUCKETS - 1) - (k < (uint32_t)(ORDER_BUCKETS - 1) ? k : (uint32_t)(ORDER_BUCKETS - 1));
}      // [1], [2]: pair-count sum / ticket of pair_count_kernel (scan_sort.cu)

struct ViewTab {
    // camera (device pointers to the reference's transposed 4x4s) + host-derived fp32 scalars
    const float* view;
    const float* proj;
    const float* campos;
    const float* bg;
    float tanfovx, tanfovy, focal_x, focal_y, limx, limy;
    const float* scalars;   // optional device copy of (focal_x, focal_y, limx, limy): overrides the four host values
    // per-Gaussian state of this view (geometry buffer)
    int32_t* radii;
    float* rec;
    float* depths;
    uint8_t* clamped;
    uint32_t* tiles_touched;
    ushort4* rect;             // tile rectangle (x0, y0, x1, y1) of the Gaussian; empty when culled
    uint32_t* point_offsets;
    uint32_t* scan_ticket;
    uint64_t* scan_desc;
    // depth order of the Gaussians: words (depth_bits << 32 | index), ping-pong, + its sort workspace
    uint64_t* gwords[2];
    uint32_t* ghist;
    uint32_t* gtickets;
    uint32_t* gdesc;
    // binning buffer
    uint64_t* keys[2];
    uint32_t* vals[2];
    uint32_t* hist;      // [MAX_PASSES][256]
    uint32_t* tickets;   // [MAX_PASSES]
    uint32_t* desc;      // [passes][tiles_cap][256]
    // image buffer
    uint32_t* ranges;
    uint32_t* n_contrib;
    uint32_t* n_visited;
    float* final_T;
    uint32_t* status;
    uint32_t* tile_count;  // [T] pairs per tile, accumulated by duplicateWithKeys (ranges = its exclusive scan)
    uint32_t* block_last;  // [T * 8] largest n_contrib of every 8x4 pixel block (written by render forward)
    float* out_color;
    float* out_depth;
    float* out_alpha;
    // backward
    const float* dL_dcolor;
    const float* dL_ddepth;
    const float* dL_dalpha;
    float* grad2d;
    float* dL_dmeans2D;   // optional per-view output (P,3)
    // extra feature channels blended by the same pass (normals, ...): BatchTab.n_extra > 0
    float* out_extra;         // (n_extra, H, W): sum_i e_i alpha_i T_i, no background term
    const float* dL_dextra;   // (n_extra, H, W) pixel gradients, or NULL
    float* gradext;           // [P][4] per-Gaussian gradient of the padded extra record (atomics, behind grad2d)
    uint8_t* touched;         // [P] set to 1 by render backward when it adds to the Gaussian's record (behind gradext)
};

struct BatchTab {
    int V;
    int P, M, sh_degree;       // sh_degree: effective (clamped) degree; -1 when colours are precomputed
    int W, H, grid_x, grid_y;
    float scale_modifier;
    uint32_t capacity;         // pairs each view's key/value arrays can hold
    int end_bit;               // tile-id bits: pair words (tile << 32 | index) are sorted on bits [32, 32 + end_bit)
    int sort_tiles_cap;        // ceil(capacity / SORT_TILE)
    int pt_words;              // > 0: scan + duplicateWithKeys also count the pairs per (partition tile of pt_words
                               // words, image tile) for the look-back-free pair partition (scan_sort.cu K4d)
    int digit_passes;          // 8-bit passes of the pair sort; 0 = one wide pass binned by the per-tile counts
    int idx_bits;              // index bits of a pair word (32)
    int clean_scratch;         // preprocess backward zeroes every grad2d record it has read (self-cleaning scratch)
    int n_extra;               // 0..4 extra per-Gaussian feature channels rendered next to the colour
    const float4* ext4;        // [P] the extra features padded to 16 B (view independent; lives in view 0's geometry)
    uint8_t* live_map;         // optional [P]: preprocess backward's scan writes 1 for a Gaussian with a gradient, else 0
    uint32_t* tile_order;      // [V * T] entries (view * T + tile), longest list first
    uint32_t* block_order;     // [4 + V * T * 8]: count, then the batch's non-empty 8x4 blocks ((view * T + tile) * 8 + block),
                               // most list entries to walk first: the work items of render backward
    uint32_t* block_hist;      // [ORDER_BUCKETS] blocks per walk-length bucket: zeroed by tile_order_kernel, counted by
                               // render forward (one atomic per warp, whose return value is the block's rank in its bucket)
    uint32_t* block_code;      // [V * T * 8] bucket << 20 | rank of every block (0xffffffff: nothing to walk)
    ViewTab v[MAX_VIEWS];
};

// ---- buffer layouts ---------------------------------------------------------------------------
struct GeomViews {
    float* rec;              // P * 12
    uint64_t* gwords[2];     // P each
    void* gsort_ws;
    float* depths;           // P
    uint8_t* clamped;        // P   (bit c set: channel c clamped at 0)
    uint32_t* tiles_touched; // P
    ushort4* rect;           // P
    uint32_t* point_offsets; // P
    void* scan_ws;
    size_t scan_ws_bytes;
    float4* ext4;            // P   (extra feature channels padded to 16 B; only written when n_extra > 0)
};
struct BinningViews {
    uint64_t* keys[2];
    uint32_t* vals[2];
    void* sort_ws;
    size_t sort_ws_bytes;
};
struct ImageViews {
    uint32_t* ranges;        // T * 2
    uint32_t* n_contrib;     // H * W
    float* final_T;          // H * W  (transmittance after the last blended entry; 1 - alpha loses bits)
    uint32_t* n_visited;     // H * W  (list entries traversed by the pixel in forward)
    uint32_t* tile_order;    // MAX_VIEWS * T (a batch uses view 0's copy)
    uint32_t* block_last;    // T * 8
    uint32_t* block_order;   // 4 + MAX_VIEWS * T * 8 (a batch uses view 0's copy)
    uint32_t* block_hist;    // ORDER_BUCKETS (view 0's copy)
    uint32_t* block_code;    // MAX_VIEWS * T * 8 (view 0's copy)
    uint32_t* status;        // STATUS_WORDS
    uint32_t* tile_count;    // T
};

size_t geom_layout(int P, void* base, GeomViews* v);
size_t binning_layout(int64_t capacity, void* base, BinningViews* v);
size_t image_layout(int H, int W, void* base, ImageViews* v);

// ---- launchers (each returns cudaError_t from the launch) -----------------------------------
cudaError_t launch_preprocess(const BatchTab& tab, const float* means3D, const float* scales, const float* rotations,
                              const float* opacities, const float* shs, const float* colors_precomp,
                              const float* cov3D_precomp, cudaStream_t st);
cudaError_t launch_duplicate(const BatchTab& tab, cudaStream_t st);
// scan + duplicateWithKeys in one kernel (scan_sort.cu); needs the scan work area AND the binning work areas cleared
bool scan_duplicate_supported(const BatchTab& tab);
int partition_direct_words(const BatchTab& tab);   // words per partition tile of the look-back-free partition, or 0
cudaError_t launch_scan_duplicate(const BatchTab& tab, cudaStream_t st);
cudaError_t launch_mark_visible(int P, const float* means3D, const float* view, const float* proj,
                                uint8_t* present, cudaStream_t st);

size_t scan_workspace_bytes(int64_t n);
// notify / epoch: optional early notice of every view's pair count in host-mapped memory (see b200splat.h)
cudaError_t launch_scan_batch(const BatchTab& tab, cudaStream_t st, bool cleared = false, uint64_t* notify = nullptr,
                              uint32_t epoch = 0);
// zero the small per-view work areas of a forward in one launch (with_binning: also the pair sort's and the tile counts)
cudaError_t launch_clear_batch(const BatchTab& tab, bool with_binning, cudaStream_t st);   // tiles_touched (in depth order) -> point_offsets
cudaError_t launch_gaussian_sort(const BatchTab& tab, cudaStream_t st, bool cleared = false);   // gwords[0] sorted by depth bits (stable)
cudaError_t launch_inclusive_scan(int64_t n, const uint32_t* in, uint32_t* out, void* ws, cudaStream_t st);
size_t sort_workspace_bytes(int64_t n);
void sort_workspace_views(void* ws, uint32_t** hist, uint32_t** tickets, uint32_t** desc);
size_t sort_workspace_zero_bytes(int64_t capacity, int end_bit);
size_t pair_sort_zero_bytes(int64_t capacity, int end_bit);   // same for the pipeline's pair sort (wide or 8-bit)
int pair_sort_digit_passes(int end_bit);   // 0: single wide pass (tile ids of <= 10 bits)
int pair_sort_result_sel(int end_bit);     // which of keys[0/1] holds the sorted words
int sort_tiles_for(int64_t n);
// stand-alone sort (stage-level entry point / KNN): histogram kernel + passes; *sel = buffer holding the result
cudaError_t launch_sort_pairs(int64_t n, int end_bit, uint64_t* keys[2], uint32_t* vals[2], void* ws, int* sel,
                              cudaStream_t st);
// pipeline sort: histograms already accumulated by duplicateWithKeys, n read from the device per view
cudaError_t launch_partition_offsets(const BatchTab& tab, cudaStream_t st);   // look-back-free partition: before ranges
cudaError_t launch_sort_batch(const BatchTab& tab, cudaStream_t st);
cudaError_t launch_tile_ranges_batch(const BatchTab& tab, int sel, cudaStream_t st);   // + tile order
cudaError_t launch_block_order(const BatchTab& tab, cudaStream_t st);   // work order of render backward (after forward)

cudaError_t launch_render_forward(const BatchTab& tab, int sel, cudaStream_t st);
cudaError_t launch_render_backward(const BatchTab& tab, int sel, cudaStream_t st);
cudaError_t launch_preprocess_backward(const BatchTab& tab, const float* means3D, const float* scales,
                                       const float* rotations, const float* shs, const float* cov3D_precomp,
                                       float* dL_dmeans3D, float* dL_dshs, float* dL_dcolors, float* dL_dopacity,
                                       float* dL_dscales, float* dL_drotations, float* dL_dcov3D,
                                       float* stat_grad_accum, float* stat_denom, float* stat_max_radii,
                                       int accumulate, cudaStream_t st, int g_begin = 0, int g_end = 0);

// extra feature channels (extra.cu): pad (P, n_extra) to float4 records; sum the views' per-Gaussian gradients
constexpr int EXT_FLOATS = 4;
cudaError_t launch_pad_extra(int P, int n_extra, const float* extra, float4* ext4, cudaStream_t st);
cudaError_t launch_extra_backward(const BatchTab& tab, float* dL_dextra, int accumulate, cudaStream_t st,
                                  int g_begin = 0, int g_end = 0);
size_t grad2d_bytes(int P);   // offset of gradext inside the backward scratch

size_t dist2_workspace_bytes(int P);
cudaError_t launch_dist2(int P, const float* points, float* out, void* ws, cudaStream_t st);

void count_launch(int n = 1);
cudaError_t launch_pair_count(const BatchTab& tab, uint64_t* notify, uint32_t epoch, cudaStream_t st);   // scan_sort.cu
int set_staging_mode(int mode);   // render.cu; returns the previous mode

// ---- all-reduce over NVLink peer memory (p2p.cu) -------------------------------------------------
constexpr int P2P_MAX_RANKS = 8;
constexpr int P2P_MAX_SEG = 16;
constexpr int P2P_ERROR_WORD = 2 * P2P_MAX_RANKS;       // signal row layout: [ready x 8][done x 8][error][counter][epoch]
constexpr int P2P_COUNTER_WORD = 2 * P2P_MAX_RANKS + 1;
constexpr int P2P_EPOCH_WORD = 2 * P2P_MAX_RANKS + 2;     // the rank's call counter (epoch argument 0: see p2p.cu)
constexpr int P2P_SIGNAL_WORDS = 64;
struct P2PTab {
    int rank, world;
    uint32_t epoch;
    int n_seg;                         // disjoint float4 ranges reduced by this call
    int64_t seg_first4[P2P_MAX_SEG], seg_n4[P2P_MAX_SEG];
    uint32_t seg_max_mask;             // bit i: segment i is max-reduced (else summed)
    int seg_row4[P2P_MAX_SEG];         // > 0: row-sparse SUM segment, float4s per row
    int64_t seg_row0[P2P_MAX_SEG];     // live-map index of the segment's first row
    int64_t live_off;                  // byte offset of the live map (one byte per index) in every rank's buffer
    float* bufs[P2P_MAX_RANKS];        // every rank's buffer (own + IPC-mapped peers)
    uint32_t* signals[P2P_MAX_RANKS];  // every rank's signal words
};
cudaError_t launch_p2p_allreduce(const P2PTab& t, cudaStream_t st);
// the same through the NVSwitch multicast address `mc` of the buffers (t.bufs unused)
cudaError_t launch_mc_allreduce(const P2PTab& t, float* mc, cudaStream_t st);

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// per-thread asynchronous 16-byte copies global -> shared (LDGSTS), groups committed / awaited by the same thread
__device__ __forceinline__ void cpa16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src)
                 : "memory");
}
__device__ __forceinline__ void cpa4(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src)
                 : "memory");
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cpa_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace b200splat

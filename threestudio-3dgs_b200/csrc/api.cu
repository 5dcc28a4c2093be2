// extern "C" surface of libb200splat.so (declared in include/b200splat.h): buffer layouts and the
// forward / backward pipelines.  Replaces upstream rasterize_points.cu (RasterizeGaussiansCUDA,
// RasterizeGaussiansBackwardCUDA, markVisible) + cuda_rasterizer/rasterizer_impl.cu
// (CudaRasterizer::Rasterizer::forward/backward) and simple-knn's distCUDA2 [UPSTREAM-RECALL];
// reference call sites renderer/diff_gaussian_rasterizer.py:98-131, geometry/gaussian_base.py:434-437.
#include "../../include/b200splat.h"
#include "common.cuh"

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstring>
#include <cstdlib>
#include <mutex>
#include <vector>

namespace b200splat {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(expr)                                                                                         \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return fail(B200SPLAT_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                        __LINE__);                                                                       \
    } while (0)

#define DEBUG_SYNC(cam, st, what)                                                                        \
    do {                                                                                                 \
        if ((cam).debug) {                                                                               \
            cudaError_t _e = cudaStreamSynchronize(st);                                                  \
            if (_e != cudaSuccess)                                                                       \
                return fail(B200SPLAT_ERR_CUDA, "kernel %s failed: %s", what, cudaGetErrorString(_e));   \
        }                                                                                                \
    } while (0)

// ---- layouts --------------------------------------------------------------------------------------
template <typename T>
static T* carve(char*& p, size_t count) {
    T* r = reinterpret_cast<T*>(p);
    p += align_up(count * sizeof(T), 256);
    return r;
}

size_t geom_layout(int P, void* base, GeomViews* v) {
    char* p = reinterpret_cast<char*>(base);
    char* p0 = p;
    GeomViews g;
    const size_t n = (size_t)(P > 0 ? P : 1);
    g.rec = carve<float>(p, n * REC_FLOATS);
    g.gwords[0] = carve<uint64_t>(p, n);
    g.gwords[1] = carve<uint64_t>(p, n);
    g.gsort_ws = carve<char>(p, sort_workspace_bytes((int64_t)n));
    g.depths = carve<float>(p, n);
    g.clamped = carve<uint8_t>(p, n);
    g.tiles_touched = carve<uint32_t>(p, n);
    g.rect = carve<ushort4>(p, n);
    g.point_offsets = carve<uint32_t>(p, n);
    g.scan_ws_bytes = scan_workspace_bytes((int64_t)n);
    g.scan_ws = carve<char>(p, g.scan_ws_bytes);
    g.ext4 = carve<float4>(p, n);
    if (v) *v = g;
    return (size_t)(p - p0);
}

size_t binning_layout(int64_t capacity, void* base, BinningViews* v) {
    char* p = reinterpret_cast<char*>(base);
    char* p0 = p;
    BinningViews b;
    const size_t n = (size_t)(capacity > 0 ? capacity : 1);
    b.keys[0] = carve<uint64_t>(p, n);
    b.keys[1] = carve<uint64_t>(p, n);
    b.vals[0] = carve<uint32_t>(p, n);
    b.vals[1] = carve<uint32_t>(p, n);
    b.sort_ws_bytes = sort_workspace_bytes((int64_t)n);
    b.sort_ws = carve<char>(p, b.sort_ws_bytes);
    if (v) *v = b;
    return (size_t)(p - p0);
}

size_t image_layout(int H, int W, void* base, ImageViews* v) {
    char* p = reinterpret_cast<char*>(base);
    char* p0 = p;
    ImageViews im;
    const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
    im.ranges = carve<uint32_t>(p, (size_t)gx * gy * 2);
    im.n_contrib = carve<uint32_t>(p, (size_t)H * W);
    im.final_T = carve<float>(p, (size_t)H * W);
    im.n_visited = carve<uint32_t>(p, (size_t)H * W);
    im.tile_order = carve<uint32_t>(p, (size_t)gx * gy * MAX_VIEWS);
    im.status = carve<uint32_t>(p, STATUS_WORDS);
    im.tile_count = carve<uint32_t>(p, (size_t)gx * gy);
    im.block_last = carve<uint32_t>(p, (size_t)gx * gy * (BLOCK_SIZE / 32));
    im.block_order = carve<uint32_t>(p, 4 + (size_t)gx * gy * (BLOCK_SIZE / 32) * MAX_VIEWS);
    im.block_hist = carve<uint32_t>(p, ORDER_BUCKETS);
    im.block_code = carve<uint32_t>(p, (size_t)gx * gy * (BLOCK_SIZE / 32) * MAX_VIEWS);
    if (v) *v = im;
    return (size_t)(p - p0);
}

// largest capacity (pairs) whose layout fits in `bytes`
static int64_t capacity_for_bytes(size_t bytes) {
    int64_t lo = 0, hi = (int64_t)(bytes / 24) + 1;
    if (hi > (1ll << 30) - 1) hi = (1ll << 30) - 1;
    while (lo < hi) {
        const int64_t mid = (lo + hi + 1) / 2;
        if (binning_layout(mid, nullptr, nullptr) <= bytes) lo = mid; else hi = mid - 1;
    }
    return lo;
}

static int higher_msb(uint32_t n) {
    uint32_t msb = sizeof(n) * 4;
    uint32_t step = msb;
    while (step > 1) {
        step /= 2;
        if (n >> msb) msb += step; else msb -= step;
    }
    if (n >> msb) msb++;
    return (int)msb;
}

// parity of the pass count decides which ping-pong half holds the sorted data
static int tile_id_bits(int T) {
    int b = 1;
    while (b < 31 && (1 << b) < T) ++b;
    return b;
}
static int sorted_sel_for(int T) {
    return pair_sort_result_sel(tile_id_bits(T));
}

// shared (view independent) part of the batch table from the first camera
static int init_table(const b200splat_camera& c, int P, int M, bool has_sh, BatchTab* tab) {
    if (c.image_height <= 0 || c.image_width <= 0) return fail(B200SPLAT_ERR_INVALID, "image size must be positive");
    memset(tab, 0, sizeof(*tab));
    tab->P = P, tab->M = M;
    tab->H = c.image_height, tab->W = c.image_width;
    tab->grid_x = (tab->W + BLOCK_X - 1) / BLOCK_X, tab->grid_y = (tab->H + BLOCK_Y - 1) / BLOCK_Y;
    tab->scale_modifier = c.scale_modifier;
    int deg = c.sh_degree < 0 ? 0 : c.sh_degree;
    if (has_sh) {
        int cap = (int)std::floor(std::sqrt((double)(M > 0 ? M : 1)) + 1e-9) - 1;  // (P,1,3) "SH" with sh_degree > 0
        if (deg > cap) deg = cap;
        if (deg > 3) deg = 3;
    } else {
        deg = -1;
    }
    tab->sh_degree = deg;
    // pair words are (tile << 32 | gaussian index): only the tile bits are sorted (the Gaussians are emitted in
    // depth order), upstream's bit count 32 + getHigherMsb(T) minus the 32 depth bits
    // (upstream sorts 32 + getHigherMsb(T) bits; the tile ids are < T, so the bits of T - 1 give the same order)
    tab->end_bit = tile_id_bits(tab->grid_x * tab->grid_y);
    tab->idx_bits = 32;
    tab->clean_scratch = 0;
    tab->digit_passes = pair_sort_digit_passes(tab->end_bit);
    return B200SPLAT_OK;
}

static int fill_camera(const b200splat_camera& c, const BatchTab& tab, ViewTab* vt) {
    if (!c.bg || !c.viewmatrix || !c.projmatrix || !c.campos)
        return fail(B200SPLAT_ERR_INVALID, "camera device pointers (bg, viewmatrix, projmatrix, campos) must be set");
    if (c.image_height != tab.H || c.image_width != tab.W)
        return fail(B200SPLAT_ERR_INVALID, "all views of a batch must share the image size");
    vt->view = c.viewmatrix, vt->proj = c.projmatrix, vt->campos = c.campos, vt->bg = c.bg;
    vt->tanfovx = c.tanfovx, vt->tanfovy = c.tanfovy;
    vt->focal_x = (float)tab.W / (2.0f * c.tanfovx);
    vt->focal_y = (float)tab.H / (2.0f * c.tanfovy);
    vt->limx = FOV_CLAMP * c.tanfovx;
    vt->limy = FOV_CLAMP * c.tanfovy;
    vt->scalars = c.scalars_dev;
    return B200SPLAT_OK;
}

static void fill_geom(int P, void* geom, ViewTab* vt, const float4** ext4 = nullptr) {
    GeomViews g;
    geom_layout(P, geom, &g);
    if (ext4) *ext4 = g.ext4;
    vt->rec = g.rec, vt->depths = g.depths, vt->clamped = g.clamped;
    vt->tiles_touched = g.tiles_touched, vt->point_offsets = g.point_offsets;
    vt->rect = g.rect;
    vt->gwords[0] = g.gwords[0], vt->gwords[1] = g.gwords[1];
    sort_workspace_views(g.gsort_ws, &vt->ghist, &vt->gtickets, &vt->gdesc);
    vt->scan_ticket = reinterpret_cast<uint32_t*>(g.scan_ws);
    vt->scan_desc = reinterpret_cast<uint64_t*>(reinterpret_cast<char*>(g.scan_ws) + 16);
}

static void fill_image(int H, int W, void* image, ViewTab* vt, BatchTab* batch) {   // batch: takes the per-batch areas
    ImageViews im;
    image_layout(H, W, image, &im);
    vt->ranges = im.ranges, vt->n_contrib = im.n_contrib, vt->n_visited = im.n_visited, vt->final_T = im.final_T;
    vt->status = im.status;
    vt->tile_count = im.tile_count;
    vt->block_last = im.block_last;
    if (batch) {
        batch->tile_order = im.tile_order, batch->block_order = im.block_order;
        batch->block_hist = im.block_hist, batch->block_code = im.block_code;
    }
}

static void fill_binning(int64_t capacity, void* binning, ViewTab* vt) {
    BinningViews b;
    binning_layout(capacity, binning, &b);
    vt->keys[0] = b.keys[0], vt->keys[1] = b.keys[1], vt->vals[0] = b.vals[0], vt->vals[1] = b.vals[1];
    sort_workspace_views(b.sort_ws, &vt->hist, &vt->tickets, &vt->desc);
}

// ---- per-family CUDA-event timing ---------------------------------------------------------------
struct ProfRecord {
    int family;
    bool counted;   // false: time is added to the family but it is not a separate launch group
    cudaEvent_t a, b;
};
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<ProfRecord> g_prof;
static std::vector<cudaEvent_t> g_event_pool;

static cudaEvent_t get_event() {
    if (!g_event_pool.empty()) {
        cudaEvent_t e = g_event_pool.back();
        g_event_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

struct ProfScope {
    bool on = false;
    ProfRecord r{};
    cudaStream_t st;
    ProfScope(int family, cudaStream_t s, bool counted = true) : st(s) {
        std::lock_guard<std::mutex> lk(g_prof_mu);
        if (!g_prof_on) return;
        on = true;
        r.family = family;
        r.counted = counted;
        r.a = get_event();
        r.b = get_event();
        cudaEventRecord(r.a, st);
    }
    ~ProfScope() {
        if (!on) return;
        cudaEventRecord(r.b, st);
        std::lock_guard<std::mutex> lk(g_prof_mu);
        g_prof.push_back(r);
    }
};

// A second stream + two events per (host thread, device): lets a forward put kernels that do not depend on each other
// on parallel branches (fork: side waits for an event recorded on the caller's stream; join: the caller's stream waits
// for an event recorded on side).  Works the same under stream capture (the branch becomes a parallel path of the graph).
struct SideLane {
    cudaStream_t s = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
};
static SideLane* side_lane() {
    static const bool enabled = [] {
        const char* e = getenv("B200SPLAT_FORK");
        return !(e && e[0] == '0');
    }();
    if (!enabled) return nullptr;
    {
        std::lock_guard<std::mutex> lk(g_prof_mu);
        if (g_prof_on) return nullptr;      // per-family event timing wants the families back to back on one stream
    }
    constexpr int MAX_DEV = 32;
    static thread_local SideLane lanes[MAX_DEV];
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEV) return nullptr;
    SideLane& l = lanes[dev];
    if (!l.s) {
        cudaStream_t s = nullptr;
        cudaEvent_t a = nullptr, b = nullptr;
        if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&a, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&b, cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        l.s = s, l.fork = a, l.join = b;
    }
    return &l;
}

struct PinnedSlot {
    int64_t* host = nullptr;
    uint32_t* host_u32 = nullptr;
};
static PinnedSlot& pinned() {
    static thread_local PinnedSlot s;
    if (!s.host_u32) {
        void* p = nullptr;
        if (cudaHostAlloc(&p, 64, cudaHostAllocDefault) == cudaSuccess) s.host_u32 = reinterpret_cast<uint32_t*>(p);
    }
    return s;
}

}  // namespace b200splat

using namespace b200splat;

// backward scratch of a view: grad2d records | extra-channel records | one byte per Gaussian that render backward sets
// when it adds a gradient to the Gaussian's record (what preprocess backward's scan looks at instead of the records)
static size_t touched_offset(int P) {
    return grad2d_bytes(P) + align_up((size_t)(P > 0 ? P : 1) * EXT_FLOATS * sizeof(float), 256);
}

// error reporting for the other translation units that export C-ABI entry points (postops.cu, adam.cu)
int b200splat_set_error(int code, const char* msg) { return fail(code, "%s", msg); }

// rotations and dL/drotations are accessed as float4 (one 16-byte access per Gaussian)
static int check_align16(const void* p, const char* what) {
    if (p && (reinterpret_cast<uintptr_t>(p) & 15))
        return fail(B200SPLAT_ERR_INVALID, "%s must be 16-byte aligned (it is accessed as float4 per Gaussian)", what);
    return B200SPLAT_OK;
}

static int check_extra(int n_extra, const void* features, bool has_out) {
    if (n_extra < 0 || n_extra > EXT_FLOATS)
        return fail(B200SPLAT_ERR_INVALID, "n_extra must be in [0, %d]", EXT_FLOATS);
    if (n_extra > 0 && (!features || !has_out))
        return fail(B200SPLAT_ERR_INVALID, "n_extra > 0 needs extra_features and the matching output");
    return B200SPLAT_OK;
}

// binning + render part of the forward, common to the single-view and the batched entry points
static int forward_tail(BatchTab& tab, int debug, cudaStream_t st, bool binning_cleared = false,
                        bool duplicated = false) {
    const int T = tab.grid_x * tab.grid_y;
    const int sel = sorted_sel_for(T);
    b200splat_camera dbg{};
    dbg.debug = debug;
    if (tab.P > 0 && tab.capacity > 0) {
        tab.sort_tiles_cap = sort_tiles_for(tab.capacity);
        if (!duplicated) {
            ProfScope ps(2, st);
            for (int v = 0; v < tab.V && !binning_cleared; ++v) {
                CU(cudaMemsetAsync(tab.v[v].hist, 0, pair_sort_zero_bytes(tab.capacity, tab.end_bit), st));
                CU(cudaMemsetAsync(tab.v[v].tile_count, 0, (size_t)T * sizeof(uint32_t), st));
            }
            CU(launch_duplicate(tab, st));
        }
        DEBUG_SYNC(dbg, st, "duplicateWithKeys");
    }
    // The tile ranges (exclusive scan of duplicateWithKeys' per-tile counts) and the tile order of the render do not
    // depend on the sorted pair words: they run on a parallel branch next to the pair partition (12 us per step off
    // the critical path).  Above 8192 tiles the ranges are read off the sorted words and the branch is not taken.
    { ProfScope ps(3, st, /*counted=*/false);
    CU(launch_partition_offsets(tab, st)); }   // (look-back-free partition only) also writes the per-tile counts
    SideLane* lane = (tab.P > 0 && tab.capacity > 0 && T <= 8192 && !debug) ? side_lane() : nullptr;
    if (lane) {
        CU(cudaEventRecord(lane->fork, st));
        CU(cudaStreamWaitEvent(lane->s, lane->fork, 0));
        CU(launch_tile_ranges_batch(tab, sel, lane->s));
        CU(cudaEventRecord(lane->join, lane->s));
    }
    if (tab.P > 0 && tab.capacity > 0) {
        { ProfScope ps(3, st);
        CU(launch_sort_batch(tab, st)); }
        DEBUG_SYNC(dbg, st, "sort");
    }
    if (lane) {
        CU(cudaStreamWaitEvent(st, lane->join, 0));
    } else {
        ProfScope ps(4, st);
        CU(launch_tile_ranges_batch(tab, sel, st));
    }
    DEBUG_SYNC(dbg, st, "identifyTileRanges");
    { ProfScope ps(5, st);
    CU(launch_render_forward(tab, sel, st)); }
    DEBUG_SYNC(dbg, st, "render");
    return B200SPLAT_OK;
}

extern "C" {

int b200splat_abi_version(void) { return B200SPLAT_ABI_VERSION; }
int b200splat_set_staging(int32_t mode) { return set_staging_mode(mode); }

// ---- all-reduce over NVLink peer memory ----------------------------------------------------------
int b200splat_p2p_alloc(size_t bytes, void** ptr, void* handle_out) {
    if (!ptr || !handle_out || bytes == 0) return fail(B200SPLAT_ERR_INVALID, "bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == B200SPLAT_P2P_HANDLE_BYTES, "IPC handle size");
    void* p = nullptr;
    CU(cudaMalloc(&p, bytes));
    CU(cudaMemset(p, 0, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return fail(B200SPLAT_ERR_CUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    }
    memcpy(handle_out, &h, sizeof(h));
    *ptr = p;
    return B200SPLAT_OK;
}
int b200splat_p2p_open(const void* handle, void** ptr) {
    if (!handle || !ptr) return fail(B200SPLAT_ERR_INVALID, "bad argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    CU(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return B200SPLAT_OK;
}
int b200splat_p2p_close(void* mapped_ptr) {
    if (mapped_ptr) CU(cudaIpcCloseMemHandle(mapped_ptr));
    return B200SPLAT_OK;
}
int b200splat_p2p_free(void* ptr) {
    if (ptr) CU(cudaFree(ptr));
    return B200SPLAT_OK;
}
int b200splat_p2p_allreduce(const b200splat_p2p_args* a) {
    if (!a) return fail(B200SPLAT_ERR_INVALID, "null args");
    if (a->world < 1 || a->world > P2P_MAX_RANKS || a->rank < 0 || a->rank >= a->world)
        return fail(B200SPLAT_ERR_INVALID, "rank %d / world %d out of range (max %d)", a->rank, a->world, P2P_MAX_RANKS);
    if (a->n_segments < 1 || a->n_segments > P2P_MAX_SEG)
        return fail(B200SPLAT_ERR_INVALID, "n_segments must be in [1, %d]", P2P_MAX_SEG);
    static_assert(P2P_SIGNAL_WORDS * 4 == B200SPLAT_P2P_SIGNAL_BYTES, "signal area size");
    static_assert(P2P_MAX_SEG == B200SPLAT_P2P_MAX_SEGMENTS, "segment count");
    P2PTab t;
    t.rank = a->rank, t.world = a->world, t.epoch = a->epoch;
    t.n_seg = a->n_segments, t.seg_max_mask = 0u;
    for (int i = 0; i < P2P_MAX_SEG; ++i) t.seg_first4[i] = t.seg_n4[i] = 0, t.seg_row4[i] = 0, t.seg_row0[i] = 0;
    for (int i = 0; i < a->n_segments; ++i) {
        if (a->seg_offset[i] < 0 || a->seg_count[i] < 0 || (a->seg_offset[i] & 3) || (a->seg_count[i] & 3))
            return fail(B200SPLAT_ERR_INVALID, "segment %d: offset and count must be non-negative multiples of 4", i);
        if (a->seg_op[i] != B200SPLAT_P2P_SUM && a->seg_op[i] != B200SPLAT_P2P_MAX)
            return fail(B200SPLAT_ERR_INVALID, "segment %d: unknown op %d", i, a->seg_op[i]);
        t.seg_first4[i] = a->seg_offset[i] / 4, t.seg_n4[i] = a->seg_count[i] / 4;
        if (a->seg_op[i] == B200SPLAT_P2P_MAX) t.seg_max_mask |= 1u << i;
        const int rf = a->seg_row_floats[i];
        if (rf != 0) {
            if (rf < 0 || (rf & 3) || a->seg_op[i] != B200SPLAT_P2P_SUM || (a->seg_count[i] % rf) != 0 ||
                a->seg_row0[i] < 0 || (a->seg_row0[i] & 3) || a->live_offset < 0 || (a->live_offset & 3))
                return fail(B200SPLAT_ERR_INVALID, "segment %d: bad row-sparse description", i);
            t.seg_row4[i] = rf / 4, t.seg_row0[i] = a->seg_row0[i];
        }
    }
    t.live_off = a->live_offset;
    for (int k = 0; k < P2P_MAX_RANKS; ++k) {
        t.bufs[k] = k < a->world ? reinterpret_cast<float*>(a->bufs[k]) : nullptr;
        t.signals[k] = k < a->world ? reinterpret_cast<uint32_t*>(a->signals[k]) : nullptr;
        if (k < a->world && (!t.bufs[k] || !t.signals[k] || (reinterpret_cast<uintptr_t>(t.bufs[k]) & 15)))
            return fail(B200SPLAT_ERR_INVALID, "buffer / signal pointer of rank %d missing or misaligned", k);
    }
    if (a->world == 1) return B200SPLAT_OK;
    CU(launch_p2p_allreduce(t, reinterpret_cast<cudaStream_t>(a->stream)));
    return B200SPLAT_OK;
}
int b200splat_mc_allreduce(const b200splat_mc_args* a) {
    if (!a) return fail(B200SPLAT_ERR_INVALID, "null args");
    if (a->world < 1 || a->world > P2P_MAX_RANKS || a->rank < 0 || a->rank >= a->world)
        return fail(B200SPLAT_ERR_INVALID, "rank %d / world %d out of range (max %d)", a->rank, a->world, P2P_MAX_RANKS);
    if (a->n_segments < 1 || a->n_segments > P2P_MAX_SEG)
        return fail(B200SPLAT_ERR_INVALID, "n_segments must be in [1, %d]", P2P_MAX_SEG);
    if (!a->mc_buffer || (reinterpret_cast<uintptr_t>(a->mc_buffer) & 15))
        return fail(B200SPLAT_ERR_INVALID, "multicast buffer pointer missing or misaligned");
    P2PTab t;
    t.rank = a->rank, t.world = a->world, t.epoch = a->epoch;
    t.n_seg = a->n_segments, t.seg_max_mask = 0u;
    for (int i = 0; i < P2P_MAX_SEG; ++i) t.seg_first4[i] = t.seg_n4[i] = 0, t.seg_row4[i] = 0, t.seg_row0[i] = 0;
    for (int i = 0; i < a->n_segments; ++i) {
        if (a->seg_offset[i] < 0 || a->seg_count[i] < 0 || (a->seg_offset[i] & 3) || (a->seg_count[i] & 3))
            return fail(B200SPLAT_ERR_INVALID, "segment %d: offset and count must be non-negative multiples of 4", i);
        if (a->seg_op[i] != B200SPLAT_P2P_SUM && a->seg_op[i] != B200SPLAT_P2P_MAX)
            return fail(B200SPLAT_ERR_INVALID, "segment %d: unknown op %d", i, a->seg_op[i]);
        t.seg_first4[i] = a->seg_offset[i] / 4, t.seg_n4[i] = a->seg_count[i] / 4;
        if (a->seg_op[i] == B200SPLAT_P2P_MAX) t.seg_max_mask |= 1u << i;
        const int rf = a->seg_row_floats[i];
        if (rf != 0) {
            if (rf < 0 || (rf & 3) || a->seg_op[i] != B200SPLAT_P2P_SUM || (a->seg_count[i] % rf) != 0 ||
                a->seg_row0[i] < 0 || (a->seg_row0[i] & 3) || a->live_offset < 0 || (a->live_offset & 3))
                return fail(B200SPLAT_ERR_INVALID, "segment %d: bad row-sparse description", i);
            t.seg_row4[i] = rf / 4, t.seg_row0[i] = a->seg_row0[i];
        }
    }
    t.live_off = a->live_offset;
    for (int k = 0; k < P2P_MAX_RANKS; ++k) {
        t.bufs[k] = nullptr;
        t.signals[k] = k < a->world ? reinterpret_cast<uint32_t*>(a->signals[k]) : nullptr;
        if (k < a->world && !t.signals[k]) return fail(B200SPLAT_ERR_INVALID, "signal pointer of rank %d missing", k);
    }
    if (a->world == 1) return B200SPLAT_OK;
    CU(launch_mc_allreduce(t, reinterpret_cast<float*>(a->mc_buffer), reinterpret_cast<cudaStream_t>(a->stream)));
    return B200SPLAT_OK;
}
int b200splat_p2p_error(const void* own_signals, int32_t* flag_out) {
    if (!own_signals || !flag_out) return fail(B200SPLAT_ERR_INVALID, "bad argument");
    uint32_t w = 0;
    CU(cudaMemcpy(&w, reinterpret_cast<const uint32_t*>(own_signals) + P2P_ERROR_WORD, 4, cudaMemcpyDeviceToHost));
    *flag_out = (int32_t)w;
    return B200SPLAT_OK;
}
const char* b200splat_last_error(void) { return g_err; }
uint64_t b200splat_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

size_t b200splat_geom_bytes(int32_t P) { return geom_layout(P, nullptr, nullptr); }
size_t b200splat_image_bytes(int32_t H, int32_t W) { return image_layout(H, W, nullptr, nullptr); }
size_t b200splat_binning_bytes(int64_t R) { return binning_layout(R, nullptr, nullptr); }
int64_t b200splat_binning_capacity(size_t bytes) { return capacity_for_bytes(bytes); }
size_t b200splat_backward_scratch_bytes(int32_t P) {   // grad2d records, then the extra-channel records
    return touched_offset(P) + align_up((size_t)(P > 0 ? P : 1), 256);   // + one "has a gradient" byte per Gaussian
}
size_t b200splat_sort_workspace_bytes(int64_t n) { return sort_workspace_bytes(n); }
size_t b200splat_scan_workspace_bytes(int64_t n) { return scan_workspace_bytes(n); }
size_t b200splat_dist2_workspace_bytes(int32_t P) { return dist2_workspace_bytes(P); }

int b200splat_forward(const b200splat_forward_args* a) {
    if (!a) return fail(B200SPLAT_ERR_INVALID, "null args");
    const int P = a->P;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(a->stream);
    if (P < 0) return fail(B200SPLAT_ERR_INVALID, "P < 0");
    const bool has_sh = a->shs != nullptr;
    if (has_sh == (a->colors_precomp != nullptr))
        return fail(B200SPLAT_ERR_INVALID, "provide exactly one of shs / colors_precomp");
    const bool has_sr = a->scales != nullptr && a->rotations != nullptr;
    if (has_sr == (a->cov3D_precomp != nullptr))
        return fail(B200SPLAT_ERR_INVALID, "provide exactly one of (scales, rotations) / cov3D_precomp");
    if (has_sh && a->M < 1) return fail(B200SPLAT_ERR_INVALID, "shs given but M < 1");
    if (int rca = check_align16(a->rotations, "rotations")) return rca;
    if (!a->out_color || !a->out_depth || !a->out_alpha) return fail(B200SPLAT_ERR_INVALID, "null output image");
    BatchTab tab;
    int rc = init_table(a->cam, P, a->M, has_sh, &tab);
    if (rc) return rc;
    tab.V = 1;
    ViewTab& vt = tab.v[0];
    rc = fill_camera(a->cam, tab, &vt);
    if (rc) return rc;
    if (!a->image_buffer || a->image_bytes < image_layout(tab.H, tab.W, nullptr, nullptr))
        return fail(B200SPLAT_ERR_NOMEM, "image_buffer too small");
    fill_image(tab.H, tab.W, a->image_buffer, &vt, &tab);
    vt.out_color = a->out_color, vt.out_depth = a->out_depth, vt.out_alpha = a->out_alpha;
    vt.radii = a->radii;
    rc = check_extra(a->n_extra, a->extra_features, a->out_extra != nullptr);
    if (rc) return rc;
    tab.n_extra = P > 0 ? a->n_extra : 0;
    vt.out_extra = a->out_extra;
    if (P <= 0 && a->n_extra > 0)
        CU(cudaMemsetAsync(a->out_extra, 0, (size_t)a->n_extra * tab.H * tab.W * sizeof(float), st));
    if (a->num_rendered_out) *a->num_rendered_out = 0;
    if (a->binning_out) *a->binning_out = a->binning_buffer;
    if (P <= 0) CU(cudaMemsetAsync(vt.status, 0, STATUS_WORDS * sizeof(uint32_t), st));

    int64_t R = 0;
    if (P > 0) {
        if (!a->means3D || !a->opacities || !a->radii) return fail(B200SPLAT_ERR_INVALID, "null per-Gaussian input");
        if (!a->geom_buffer || a->geom_bytes < geom_layout(P, nullptr, nullptr))
            return fail(B200SPLAT_ERR_NOMEM, "geom_buffer too small");
        fill_geom(P, a->geom_buffer, &vt, &tab.ext4);
        { ProfScope ps(0, st);
        CU(launch_clear_batch(tab, /*with_binning=*/false, st));   // binning buffer: sized after the scan
        CU(launch_preprocess(tab, a->means3D, a->scales, a->rotations, a->opacities, a->shs, a->colors_precomp,
                             a->cov3D_precomp, st));
        CU(launch_pad_extra(P, tab.n_extra, a->extra_features, const_cast<float4*>(tab.ext4), st)); }
        DEBUG_SYNC(a->cam, st, "preprocess");
        { ProfScope ps(3, st, /*counted=*/false);
        CU(launch_gaussian_sort(tab, st, /*cleared=*/true)); }
        DEBUG_SYNC(a->cam, st, "depth sort");
        { ProfScope ps(1, st);
        CU(launch_scan_batch(tab, st, /*cleared=*/true)); }
        DEBUG_SYNC(a->cam, st, "scan");
        // the one host<->device round trip of the single-view forward: num_rendered sizes the binning buffer
        // (the batched entry point works on a caller-chosen capacity instead and never waits)
        PinnedSlot& slot = pinned();
        if (!slot.host_u32) return fail(B200SPLAT_ERR_CUDA, "cudaHostAlloc failed");
        CU(cudaMemcpyAsync(slot.host_u32, vt.point_offsets + (P - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        R = (int64_t)*slot.host_u32;
        if (R >= (1ll << 30)) return fail(B200SPLAT_ERR_INVALID, "num_rendered %lld exceeds 2^30", (long long)R);
    }
    if (a->num_rendered_out) *a->num_rendered_out = R;
    if (R > 0) {
        const size_t need = binning_layout(R, nullptr, nullptr);
        void* bbuf = a->binning_buffer;
        if (!bbuf || a->binning_bytes < need) {
            if (!a->binning_alloc) return fail(B200SPLAT_ERR_NOMEM, "binning_buffer too small and no allocator given");
            bbuf = a->binning_alloc(a->alloc_user, need);
            if (!bbuf) return fail(B200SPLAT_ERR_NOMEM, "binning allocator returned NULL for %zu bytes", need);
        }
        if (a->binning_out) *a->binning_out = bbuf;
        fill_binning(R, bbuf, &vt);
        tab.capacity = (uint32_t)R;
    }
    return forward_tail(tab, a->cam.debug, st);
}

int b200splat_forward_batched(const b200splat_batch_forward_args* a) {
    if (!a) return fail(B200SPLAT_ERR_INVALID, "null args");
    const int P = a->P, V = a->V;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(a->stream);
    if (V < 1 || V > MAX_VIEWS) return fail(B200SPLAT_ERR_INVALID, "V must be in [1, %d]", MAX_VIEWS);
    if (P <= 0) return fail(B200SPLAT_ERR_INVALID, "P must be positive");
    if (!a->cams || !a->out_color || !a->out_depth || !a->out_alpha || !a->radii || !a->geom_buffer ||
        !a->image_buffer || !a->binning_buffer)
        return fail(B200SPLAT_ERR_INVALID, "null per-view array");
    const bool has_sh = a->shs != nullptr;
    if (has_sh == (a->colors_precomp != nullptr))
        return fail(B200SPLAT_ERR_INVALID, "provide exactly one of shs / colors_precomp");
    if (!a->scales || !a->rotations || !a->means3D || !a->opacities)
        return fail(B200SPLAT_ERR_INVALID, "means3D, opacities, scales and rotations are required");
    if (has_sh && a->M < 1) return fail(B200SPLAT_ERR_INVALID, "shs given but M < 1");
    if (int rca = check_align16(a->rotations, "rotations")) return rca;
    BatchTab tab;
    int rc = init_table(a->cams[0], P, a->M, has_sh, &tab);
    if (rc) return rc;
    tab.V = V;
    const int64_t cap = capacity_for_bytes(a->binning_bytes);
    if (cap < 1) return fail(B200SPLAT_ERR_NOMEM, "binning_bytes too small");
    tab.capacity = (uint32_t)cap;
    rc = check_extra(a->n_extra, a->extra_features, a->out_extra != nullptr);
    if (rc) return rc;
    tab.n_extra = a->n_extra;
    for (int v = 0; v < V; ++v) {
        ViewTab& vt = tab.v[v];
        rc = fill_camera(a->cams[v], tab, &vt);
        if (rc) return rc;
        if (!a->geom_buffer[v] || !a->image_buffer[v] || !a->binning_buffer[v] || !a->radii[v] || !a->out_color[v] ||
            !a->out_depth[v] || !a->out_alpha[v] || (tab.n_extra > 0 && !a->out_extra[v]))
            return fail(B200SPLAT_ERR_INVALID, "null buffer for view %d", v);
        vt.out_extra = tab.n_extra > 0 ? a->out_extra[v] : nullptr;
        fill_geom(P, a->geom_buffer[v], &vt, v == 0 ? &tab.ext4 : nullptr);
        fill_image(tab.H, tab.W, a->image_buffer[v], &vt, v == 0 ? &tab : nullptr);
        fill_binning(cap, a->binning_buffer[v], &vt);
        vt.out_color = a->out_color[v], vt.out_depth = a->out_depth[v], vt.out_alpha = a->out_alpha[v];
        vt.radii = a->radii[v];
    }
    { ProfScope ps(0, st);
    CU(launch_clear_batch(tab, /*with_binning=*/true, st));
    CU(launch_preprocess(tab, a->means3D, a->scales, a->rotations, a->opacities, a->shs, a->colors_precomp, nullptr,
                         st));
    CU(launch_pad_extra(P, tab.n_extra, a->extra_features, const_cast<float4*>(tab.ext4), st));
    // the views' pair counts for a host that waits for them (pairs_notify): known here, ~5 us of work, with the
    // depth sort, the scan, the key duplication, the tile partition and the render still to be queued behind it
    CU(launch_pair_count(tab, a->pairs_notify, a->notify_epoch, st)); }
    DEBUG_SYNC(a->cams[0], st, "preprocess");
    { ProfScope ps(3, st, /*counted=*/false);
    CU(launch_gaussian_sort(tab, st, /*cleared=*/true)); }
    DEBUG_SYNC(a->cams[0], st, "depth sort");
    // scan of the pair counts in depth order + duplicateWithKeys: one fused kernel (family "duplicate" of the profile)
    // when the tile histogram fits in shared memory, else the two stand-alone kernels
    tab.sort_tiles_cap = sort_tiles_for(tab.capacity);
    const bool fused = scan_duplicate_supported(tab);
    tab.pt_words = fused ? partition_direct_words(tab) : 0;
    if (fused) {
        ProfScope ps(2, st);
        CU(launch_scan_duplicate(tab, st));
    } else {
        ProfScope ps(1, st);
        CU(launch_scan_batch(tab, st, /*cleared=*/true));
    }
    DEBUG_SYNC(a->cams[0], st, "scan");
    rc = forward_tail(tab, a->cams[0].debug, st, /*binning_cleared=*/true, /*duplicated=*/fused);
    if (rc) return rc;
    if (a->sync) {
        static thread_local uint32_t* host = nullptr;
        if (!host && cudaHostAlloc(reinterpret_cast<void**>(&host), 2 * MAX_VIEWS * sizeof(uint32_t),
                                   cudaHostAllocDefault) != cudaSuccess)
            return fail(B200SPLAT_ERR_CUDA, "cudaHostAlloc failed");
        for (int v = 0; v < V; ++v) {
            CU(cudaMemcpyAsync(host + 2 * v, tab.v[v].point_offsets + (P - 1), 4, cudaMemcpyDeviceToHost, st));
            CU(cudaMemcpyAsync(host + 2 * v + 1, tab.v[v].status + STATUS_OVERFLOW, 4, cudaMemcpyDeviceToHost, st));
        }
        CU(cudaStreamSynchronize(st));
        for (int v = 0; v < V; ++v) {
            if (a->num_rendered_out) a->num_rendered_out[v] = (int64_t)host[2 * v];
            if (a->overflow_out) a->overflow_out[v] = (int32_t)(host[2 * v + 1] != 0 || host[2 * v] > tab.capacity);
        }
    }
    return B200SPLAT_OK;
}

static int backward_run(BatchTab& tab, const float* means3D, const float* scales, const float* rotations,
                        const float* shs, const float* cov3D_precomp, float* dL_dmeans3D, float* dL_dshs,
                        float* dL_dcolors, float* dL_dopacity, float* dL_dscales, float* dL_drotations,
                        float* dL_dcov3D, float* sa, float* sd, float* sm, int accumulate, int debug, bool any_pairs,
                        cudaStream_t st, bool scratch_clean = false, int phase = 0, int g_begin = 0, int g_end = 0,
                        float* dL_dextra = nullptr) {
    b200splat_camera dbg{};
    dbg.debug = debug;
    const int sel = sorted_sel_for(tab.grid_x * tab.grid_y);
    tab.clean_scratch = scratch_clean ? 1 : 0;
    if (phase != 2) {
        ProfScope ps(6, st);
        if (!scratch_clean)
            for (int v = 0; v < tab.V; ++v) {
                CU(cudaMemsetAsync(tab.v[v].grad2d, 0, (size_t)tab.P * GRAD2D_FLOATS * sizeof(float), st));
                CU(cudaMemsetAsync(tab.v[v].touched, 0, (size_t)tab.P, st));
                if (tab.n_extra > 0)
                    CU(cudaMemsetAsync(tab.v[v].gradext, 0, (size_t)tab.P * EXT_FLOATS * sizeof(float), st));
            }
        if (any_pairs) {
            CU(launch_block_order(tab, st));
            CU(launch_render_backward(tab, sel, st));
        }
    }
    DEBUG_SYNC(dbg, st, "render backward");
    if (phase != 1) {
        ProfScope ps(7, st);
        CU(launch_preprocess_backward(tab, means3D, scales, rotations, shs, cov3D_precomp, dL_dmeans3D, dL_dshs,
                                      dL_dcolors, dL_dopacity, dL_dscales, dL_drotations, dL_dcov3D, sa, sd, sm,
                                      accumulate, st, g_begin, g_end));
        CU(launch_extra_backward(tab, dL_dextra, accumulate, st, g_begin, g_end));
    }
    DEBUG_SYNC(dbg, st, "preprocess backward");
    return B200SPLAT_OK;
}

int b200splat_backward(const b200splat_backward_args* a) {
    if (!a) return fail(B200SPLAT_ERR_INVALID, "null args");
    const int P = a->P;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(a->stream);
    if (P <= 0) return B200SPLAT_OK;
    const bool has_sh = a->shs != nullptr;
    BatchTab tab;
    int rc = init_table(a->cam, P, a->M, has_sh, &tab);
    if (rc) return rc;
    tab.V = 1;
    ViewTab& vt = tab.v[0];
    rc = fill_camera(a->cam, tab, &vt);
    if (rc) return rc;
    if (!a->dL_dmeans3D || !a->dL_dmeans2D || !a->dL_dopacity)
        return fail(B200SPLAT_ERR_INVALID, "dL_dmeans3D / dL_dmeans2D / dL_dopacity must be given");
    if (has_sh && !a->dL_dshs) return fail(B200SPLAT_ERR_INVALID, "shs given but dL_dshs is NULL");
    if (int rca = check_align16(a->rotations, "rotations")) return rca;
    if (int rca = check_align16(a->dL_drotations, "dL_drotations")) return rca;
    if (!a->scratch || a->scratch_bytes < b200splat_backward_scratch_bytes(P))
        return fail(B200SPLAT_ERR_NOMEM, "scratch too small");
    if (!a->geom_buffer || !a->image_buffer || !a->radii) return fail(B200SPLAT_ERR_INVALID, "null saved buffer");
    rc = check_extra(a->n_extra, a->extra_features, a->dL_dextra != nullptr);
    if (rc) return rc;
    tab.n_extra = a->n_extra;
    fill_geom(P, const_cast<void*>(a->geom_buffer), &vt, &tab.ext4);
    fill_image(tab.H, tab.W, const_cast<void*>(a->image_buffer), &vt, &tab);
    vt.radii = const_cast<int32_t*>(a->radii);
    if (a->num_rendered > 0) {
        if (!a->binning_buffer) return fail(B200SPLAT_ERR_INVALID, "null binning buffer");
        fill_binning(a->num_rendered, const_cast<void*>(a->binning_buffer), &vt);
        tab.capacity = (uint32_t)a->num_rendered;
    }
    vt.dL_dcolor = a->dL_dout_color, vt.dL_ddepth = a->dL_dout_depth, vt.dL_dalpha = a->dL_dout_alpha;
    vt.grad2d = reinterpret_cast<float*>(a->scratch);
    vt.gradext = reinterpret_cast<float*>(reinterpret_cast<char*>(a->scratch) + grad2d_bytes(P));
    vt.touched = reinterpret_cast<uint8_t*>(a->scratch) + touched_offset(P);
    vt.dL_dextra = a->dL_dout_extra;
    vt.dL_dmeans2D = a->dL_dmeans2D;
    return backward_run(tab, a->means3D, a->scales, a->rotations, a->shs, a->cov3D_precomp, a->dL_dmeans3D, a->dL_dshs,
                        a->dL_dcolors, a->dL_dopacity, a->dL_dscales, a->dL_drotations, a->dL_dcov3D, a->stat_grad_accum,
                        a->stat_denom, a->stat_max_radii, a->accumulate, a->cam.debug, a->num_rendered > 0, st,
                        /*scratch_clean=*/false, /*phase=*/0, 0, 0, a->dL_dextra);
}

int b200splat_backward_batched(const b200splat_batch_backward_args* a) {
    if (!a) return fail(B200SPLAT_ERR_INVALID, "null args");
    const int P = a->P, V = a->V;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(a->stream);
    if (V < 1 || V > MAX_VIEWS) return fail(B200SPLAT_ERR_INVALID, "V must be in [1, %d]", MAX_VIEWS);
    if (P <= 0) return fail(B200SPLAT_ERR_INVALID, "P must be positive");
    if (!a->cams || !a->radii || !a->geom_buffer || !a->image_buffer || !a->binning_buffer || !a->scratch)
        return fail(B200SPLAT_ERR_INVALID, "null per-view array");
    const bool has_sh = a->shs != nullptr;
    if (!a->dL_dmeans3D || !a->dL_dopacity || !a->dL_dscales || !a->dL_drotations)
        return fail(B200SPLAT_ERR_INVALID, "dL_dmeans3D / dL_dopacity / dL_dscales / dL_drotations must be given");
    if (has_sh && !a->dL_dshs) return fail(B200SPLAT_ERR_INVALID, "shs given but dL_dshs is NULL");
    if (int rca = check_align16(a->rotations, "rotations")) return rca;
    if (int rca = check_align16(a->dL_drotations, "dL_drotations")) return rca;
    BatchTab tab;
    int rc = init_table(a->cams[0], P, a->M, has_sh, &tab);
    if (rc) return rc;
    tab.V = V;
    const int64_t cap = capacity_for_bytes(a->binning_bytes);
    if (cap < 1) return fail(B200SPLAT_ERR_NOMEM, "binning_bytes too small");
    tab.capacity = (uint32_t)cap;
    rc = check_extra(a->n_extra, a->extra_features, a->dL_dextra != nullptr);
    if (rc) return rc;
    tab.n_extra = a->n_extra;
    for (int v = 0; v < V; ++v) {
        ViewTab& vt = tab.v[v];
        rc = fill_camera(a->cams[v], tab, &vt);
        if (rc) return rc;
        if (!a->geom_buffer[v] || !a->image_buffer[v] || !a->binning_buffer[v] || !a->radii[v] || !a->scratch[v])
            return fail(B200SPLAT_ERR_INVALID, "null buffer for view %d", v);
        fill_geom(P, const_cast<void*>(a->geom_buffer[v]), &vt, v == 0 ? &tab.ext4 : nullptr);
        vt.gradext = reinterpret_cast<float*>(reinterpret_cast<char*>(a->scratch[v]) + grad2d_bytes(P));
        vt.touched = reinterpret_cast<uint8_t*>(a->scratch[v]) + touched_offset(P);
        vt.dL_dextra = (tab.n_extra > 0 && a->dL_dout_extra) ? a->dL_dout_extra[v] : nullptr;
        fill_image(tab.H, tab.W, const_cast<void*>(a->image_buffer[v]), &vt, v == 0 ? &tab : nullptr);
        fill_binning(cap, const_cast<void*>(a->binning_buffer[v]), &vt);
        vt.radii = const_cast<int32_t*>(a->radii[v]);
        vt.dL_dcolor = a->dL_dout_color ? a->dL_dout_color[v] : nullptr;
        vt.dL_ddepth = a->dL_dout_depth ? a->dL_dout_depth[v] : nullptr;
        vt.dL_dalpha = a->dL_dout_alpha ? a->dL_dout_alpha[v] : nullptr;
        vt.grad2d = reinterpret_cast<float*>(a->scratch[v]);
        vt.dL_dmeans2D = a->dL_dmeans2D ? a->dL_dmeans2D[v] : nullptr;
    }
    tab.live_map = a->live_map;
    return backward_run(tab, a->means3D, a->scales, a->rotations, a->shs, nullptr, a->dL_dmeans3D, a->dL_dshs,
                        a->dL_dcolors, a->dL_dopacity, a->dL_dscales, a->dL_drotations, nullptr, a->stat_grad_accum,
                        a->stat_denom, a->stat_max_radii, a->accumulate, a->cams[0].debug, true, st, a->scratch_clean != 0,
                        a->phase, a->g_begin, a->g_end, a->dL_dextra);
}

int b200splat_mark_visible(int32_t P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                           uint8_t* present, b200splat_stream stream) {
    if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !present))) return fail(B200SPLAT_ERR_INVALID, "bad argument");
    CU(launch_mark_visible(P, means3D, viewmatrix, projmatrix, present, reinterpret_cast<cudaStream_t>(stream)));
    return B200SPLAT_OK;
}

int b200splat_dist2(int32_t P, const float* points, float* out, void* workspace, size_t workspace_bytes,
                    b200splat_stream stream) {
    if (P < 0 || (P > 0 && (!points || !out))) return fail(B200SPLAT_ERR_INVALID, "bad argument");
    if (P > 0 && (!workspace || workspace_bytes < dist2_workspace_bytes(P)))
        return fail(B200SPLAT_ERR_NOMEM, "dist2 workspace too small");
    { ProfScope ps(8, reinterpret_cast<cudaStream_t>(stream));
    CU(launch_dist2(P, points, out, workspace, reinterpret_cast<cudaStream_t>(stream))); }
    return B200SPLAT_OK;
}

int b200splat_sort_pairs(int64_t n, int32_t end_bit, uint64_t* keys, uint32_t* vals, uint64_t* keys_alt,
                         uint32_t* vals_alt, void* workspace, size_t workspace_bytes, int32_t* result_in_alt,
                         b200splat_stream stream) {
    if (n < 0 || n >= (1ll << 30)) return fail(B200SPLAT_ERR_INVALID, "n out of range");
    if (n > 0 && (!keys || !vals || !keys_alt || !vals_alt || !workspace))
        return fail(B200SPLAT_ERR_INVALID, "null buffer");
    if (n > 0 && workspace_bytes < sort_workspace_bytes(n)) return fail(B200SPLAT_ERR_NOMEM, "sort workspace too small");
    uint64_t* k[2] = {keys, keys_alt};
    uint32_t* v[2] = {vals, vals_alt};
    int sel = 0;
    CU(launch_sort_pairs(n, end_bit, k, v, workspace, &sel, reinterpret_cast<cudaStream_t>(stream)));
    if (result_in_alt) *result_in_alt = sel;
    return B200SPLAT_OK;
}

int b200splat_inclusive_scan_u32(int64_t n, const uint32_t* in, uint32_t* out, void* workspace, size_t workspace_bytes,
                                 b200splat_stream stream) {
    if (n < 0) return fail(B200SPLAT_ERR_INVALID, "n < 0");
    if (n > 0 && (!in || !out || !workspace || workspace_bytes < scan_workspace_bytes(n)))
        return fail(B200SPLAT_ERR_NOMEM, "scan workspace too small");
    CU(launch_inclusive_scan(n, in, out, workspace, reinterpret_cast<cudaStream_t>(stream)));
    return B200SPLAT_OK;
}

int b200splat_forward_views_get(int32_t P, int32_t H, int32_t W, int64_t num_rendered, const void* geom_buffer,
                                const void* binning_buffer, const void* image_buffer, b200splat_forward_views* out) {
    if (!out) return fail(B200SPLAT_ERR_INVALID, "null out");
    memset(out, 0, sizeof(*out));
    if (geom_buffer && P > 0) {
        GeomViews g;
        geom_layout(P, const_cast<void*>(geom_buffer), &g);
        out->tiles_touched = g.tiles_touched;
        out->point_offsets = g.point_offsets;
        out->depths = g.depths;
        out->gauss2d = g.rec;
        out->gaussian_order = g.gwords[0];
    }
    if (binning_buffer && num_rendered > 0) {   // num_rendered = the capacity the buffer was laid out for
        BinningViews b;
        binning_layout(num_rendered, const_cast<void*>(binning_buffer), &b);
        const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
        const int sel = sorted_sel_for(gx * gy);
        out->keys_sorted = b.keys[sel];
        out->point_list = b.vals[sel];
        out->packed_idx_bits = 32;
    }
    if (image_buffer) {
        ImageViews im;
        image_layout(H, W, const_cast<void*>(image_buffer), &im);
        out->ranges = im.ranges;
        out->n_contrib = im.n_contrib;
        out->n_visited = im.n_visited;
        out->status = im.status;
    }
    return B200SPLAT_OK;
}

int b200splat_profile_enable(int32_t on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_on = on != 0;
    return B200SPLAT_OK;
}

int b200splat_profile_read(float* ms, int64_t* count) {
    if (!ms || !count) return fail(B200SPLAT_ERR_INVALID, "null out");
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (int f = 0; f < B200SPLAT_NUM_FAMILIES; ++f) ms[f] = 0.f, count[f] = 0;
    for (const ProfRecord& r : g_prof) {
        float t = 0.f;
        cudaError_t e = cudaEventSynchronize(r.b);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&t, r.a, r.b);
        if (e != cudaSuccess) return fail(B200SPLAT_ERR_CUDA, "profile event: %s", cudaGetErrorString(e));
        if (r.family >= 0 && r.family < B200SPLAT_NUM_FAMILIES) ms[r.family] += t, count[r.family] += r.counted ? 1 : 0;
        g_event_pool.push_back(r.a);
        g_event_pool.push_back(r.b);
    }
    g_prof.clear();
    return B200SPLAT_OK;
}

}  // extern "C"

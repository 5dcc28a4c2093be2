// K6 render forward, K7 render backward: per-16x16-tile alpha blending.
//
// Replaces upstream forward.cu renderCUDA and backward.cu renderCUDA (ashawkey variant with depth
// and alpha channels) [UPSTREAM-RECALL]; semantics restated in oracle/torch_oracle.py
// (_blend_tile); reference consumers renderer/diff_gaussian_rasterizer_advanced.py:122-146.
//
// B200 design:
//  * forward: one CTA (8 warps, one 8x4 pixel block each) per tile shares batches of 256 staged entries,
//    double buffered (measured faster than warp-private staging for the forward: 64 vs 75 us/view);
//  * backward: the unit of work is ONE WARP = one 8x4 pixel block of a tile, launched as its own 32-thread CTA:
//    the eight warps of a tile used to wait for each other at a barrier per batch (40 % of the stall
//    samples: an edge-of-object block walks ten times more survivors than its neighbours), and a finished
//    warp kept its registers until the slowest one was done.  Independent warps stop as soon as their own
//    32 pixels are saturated (forward) / start at their own last contributor (backward), and the SM back-
//    fills with the next work item;
//  * a warp consumes its tile's Gaussian list in batches of 64 entries through a 3-stage shared-memory
//    ring behind mbarriers: each lane gathers two 48-byte records per batch (common.cuh REC layout) either
//    with bulk async copies (cp.async.bulk -> UBLKCP, complete_tx on the mbarrier) or with 16-byte
//    cp.async (LDGSTS) whose completion arrives on the same mbarrier, two batches ahead of the blend;
//  * warp-cooperative culling: for every 32 staged entries, lane l bounds entry l's best-case alpha
//    over the warp's 8x4 pixel rectangle (exact minimum of the conic quadratic over the rectangle
//    edges); only entries that can reach alpha >= 1/255 somewhere in the rectangle are evaluated by
//    the 32 pixels.  Culling is conservative, so the set of blended (pixel, Gaussian) pairs -- and
//    therefore the image, n_contrib and the gradients -- are exactly those of the unculled loop;
//  * forward terminates a tile as soon as every pixel is saturated (T' < 1e-4);
//  * backward walks back to front starting at the tile's largest n_contrib, and reduces the 10
//    per-Gaussian partial gradients across the warp with a 12-shuffle transposed butterfly before
//    one 10-lane global atomic per (warp, Gaussian).
#include "common.cuh"

#include <atomic>
#include <cstdlib>
#include <cstring>

namespace b200splat {

constexpr int BATCH = 64;    // list entries per warp batch (2 per lane)
constexpr int STAGES = 3;    // ring depth: two batches in flight while one is blended
constexpr int WARPS_PER_TILE = BLOCK_SIZE / 32;
constexpr int ILP = 4;   // survivors whose alpha is evaluated together (hides the LDS/MUFU latency chain)

// ---- mbarrier / async-copy PTX ----------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// pixel owned by this lane: block w of the tile -> 8x4 pixels at (w%2, w/2), lane -> (l%8, l/8)
__device__ __forceinline__ void thread_pixel(int wblock, int& lx, int& ly) {
    const int lane = threadIdx.x & 31;
    lx = (wblock & 1) * 8 + (lane & 7);
    ly = (wblock >> 1) * 4 + (lane >> 3);
}

struct PointList {   // sorted list: packed words (key << idx_bits | idx) or a plain index array
    const uint64_t* words;
    const uint32_t* vals;
    uint32_t mask;
    __device__ __forceinline__ uint32_t at(int64_t pos) const {
        return words ? (uint32_t)(__ldg(words + pos)) & mask : __ldg(vals + pos);
    }
};

// gather the record of list entry `pos` into buf[slot] (and, EXT, its extra-feature record into extbuf[slot]);
// completion arrives on `bar`
template <bool BULK, bool EXT = false>
__device__ __forceinline__ void stage_entry(float4* buf, uint32_t* ids, uint64_t* bar, const float* __restrict__ rec,
                                            const PointList point_list, int slot, int64_t pos, bool valid,
                                            float4* extbuf = nullptr, const float4* __restrict__ ext4 = nullptr) {
    if (valid) {
        const uint32_t id = point_list.at(pos);
        if (ids) ids[slot] = id;
        const float* src = rec + (size_t)id * REC_FLOATS;
        float4* dst = buf + 3 * slot;
        if (BULK) {
            mbar_arrive_expect_tx(bar, REC_FLOATS * 4 + (EXT ? 16 : 0));
            bulk_g2s(dst, src, REC_FLOATS * 4, bar);
            if (EXT) bulk_g2s(extbuf + slot, ext4 + id, 16, bar);
        } else {
            cp_async16(dst, src);
            cp_async16(dst + 1, src + 4);
            cp_async16(dst + 2, src + 8);
            if (EXT) cp_async16(extbuf + slot, ext4 + id);
            cp_async_arrive_noinc(bar);
        }
    } else {
        mbar_arrive(bar);
    }
}

// Conservative test: can entry (x, y, conic A,B,C, threshold thr = 2 ln(255 o) + margin) reach
// alpha >= 1/255 at some point of the rectangle [X0,X1] x [Y0,Y1]?  f = A dx^2 + 2 B dx dy + C dy^2 is
// minimised exactly over the four edges (convex when the conic is positive definite; otherwise keep).
__device__ __forceinline__ bool cull_keep(const float4 q0, const float C, const float thr, float X0, float X1,
                                          float Y0, float Y1) {
    const float A = q0.z, B = q0.w;
    const float dx0 = q0.x - X0, dx1 = q0.x - X1, dy0 = q0.y - Y0, dy1 = q0.y - Y1;
    const bool convex = (A > 0.f) && (C > 0.f) && (A * C - B * B > 0.f);
    const bool inside = (dx0 >= 0.f) && (dx1 <= 0.f) && (dy0 >= 0.f) && (dy1 <= 0.f);
    const float iC = __fdividef(1.f, C), iA = __fdividef(1.f, A);
    // lower bound of f at (dx,dy): value minus a 64-ulp bound on the rounding of both this evaluation and
    // the per-pixel one (terms can cancel for elongated Gaussians)
    auto lower = [&](float dx, float dy) {
        const float t1 = A * dx * dx, t2 = 2.f * B * dx * dy, t3 = C * dy * dy;
        return (t1 + t2 + t3) - 8e-6f * (t1 + fabsf(t2) + t3);
    };
    float fmin = lower(dx0, fminf(dy0, fmaxf(dy1, -B * dx0 * iC)));
    fmin = fminf(fmin, lower(dx1, fminf(dy0, fmaxf(dy1, -B * dx1 * iC))));
    fmin = fminf(fmin, lower(fminf(dx0, fmaxf(dx1, -B * dy0 * iA)), dy0));
    fmin = fminf(fmin, lower(fminf(dx0, fmaxf(dx1, -B * dy1 * iA)), dy1));
    return !convex || inside || !(fmin > thr);
}

// ---- CTA-per-tile helpers of the forward kernel (8 warps share one staged batch of 256 entries) --------
constexpr int FWD_BATCH = 256;
constexpr int FWD_STAGES = 2;
__device__ __forceinline__ void thread_pixel_cta(int& lx, int& ly) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    lx = (warp & 1) * 8 + (lane & 7);
    ly = (warp >> 1) * 4 + (lane >> 3);
}
template <bool BULK, bool EXT>
__device__ __forceinline__ void stage_batch_cta(float4* buf, uint32_t* ids, uint64_t* bar, const float* __restrict__ rec,
                                                const PointList point_list, int64_t pos, bool valid, float4* extbuf,
                                                const float4* __restrict__ ext4) {
    stage_entry<BULK, EXT>(buf, ids, bar, rec, point_list, (int)threadIdx.x, pos, valid, extbuf, ext4);
}

// ============================================================================================
// K6 forward
// ============================================================================================
// EXT: up to 4 extra feature channels per Gaussian (one more 16-byte record per staged entry) are blended with the
// same weights as the colour and written to out_extra (no background term)
template <bool BULK, bool EXT>
__global__ void __launch_bounds__(BLOCK_SIZE)
render_forward_kernel(const __grid_constant__ BatchTab tab, int sel) {
    __shared__ __align__(16) float4 s_rec[FWD_STAGES][FWD_BATCH * 3];
    __shared__ __align__(16) float4 s_ext[FWD_STAGES][EXT ? FWD_BATCH : 1];
    __shared__ __align__(8) uint64_t s_bar[FWD_STAGES];
    __shared__ __align__(16) uint8_t s_surv[BLOCK_SIZE / 32][FWD_BATCH + 16];

    const int W = tab.W, H = tab.H, grid_x = tab.grid_x;
    const int n_tiles = grid_x * tab.grid_y;
    const uint32_t entry = tab.tile_order[blockIdx.x];   // longest lists of the whole view batch first (LPT)
    const ViewTab& vt = tab.v[entry / n_tiles];
    const int tile = (int)(entry % n_tiles);
    const int tile_x = tile % grid_x, tile_y = tile / grid_x;
    const uint32_t* __restrict__ ranges = vt.ranges;
    const PointList point_list{tab.idx_bits ? vt.keys[sel] : nullptr, vt.vals[sel],
                               tab.idx_bits ? (uint32_t)((1ull << tab.idx_bits) - 1ull) : 0xffffffffu};
    const float* __restrict__ rec = vt.rec;
    const float4* __restrict__ ext4 = tab.ext4;
    const float* __restrict__ bg = vt.bg;
    uint32_t* __restrict__ n_contrib = vt.n_contrib;
    uint32_t* __restrict__ n_visited = vt.n_visited;
    float* __restrict__ final_T = vt.final_T;
    float* __restrict__ out_color = vt.out_color;
    float* __restrict__ out_depth = vt.out_depth;
    float* __restrict__ out_alpha = vt.out_alpha;
    int lx, ly;
    thread_pixel_cta(lx, ly);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const int pxi = tile_x * BLOCK_X + lx, pyi = tile_y * BLOCK_Y + ly;
    const bool inside = pxi < W && pyi < H;
    const float pixx = (float)pxi, pixy = (float)pyi;
    const float X0 = (float)(tile_x * BLOCK_X + (warp & 1) * 8), X1 = X0 + 7.f;
    const float Y0 = (float)(tile_y * BLOCK_Y + (warp >> 1) * 4), Y1 = Y0 + 3.f;
    const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
    const int total = (int)(r1 - r0);
    const int rounds = (total + FWD_BATCH - 1) / FWD_BATCH;

    if (threadIdx.x == 0) {
        mbar_init(&s_bar[0], BLOCK_SIZE);
        mbar_init(&s_bar[1], BLOCK_SIZE);
        mbar_fence_init();
    }
    __syncthreads();

    bool done = !inside;
    float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, Wt = 0.f, D = 0.f;
    float E0 = 0.f, E1 = 0.f, E2 = 0.f, E3 = 0.f;
    uint32_t last_contributor = 0, visited = 0;
    int traversed = 0;

    if (rounds > 0)
        stage_batch_cta<BULK, EXT>(s_rec[0], nullptr, &s_bar[0], rec, point_list, (int64_t)r0 + threadIdx.x,
                                   (int)threadIdx.x < total, s_ext[0], ext4);
    for (int b = 0; b < rounds; ++b) {
        // also orders "everyone finished reading stage (b+1)&1" before it is refilled
        const int num_done = __syncthreads_count(done);
        const int s = b & 1;
        if (num_done == BLOCK_SIZE) {
            // batch b is already in flight: it must land before the CTA's shared memory is released
            mbar_wait(&s_bar[s], (b >> 1) & 1);
            break;
        }
        if (b + 1 < rounds) {
            const int nb = (b + 1) * FWD_BATCH + threadIdx.x;
            stage_batch_cta<BULK, EXT>(s_rec[s ^ 1], nullptr, &s_bar[s ^ 1], rec, point_list, (int64_t)r0 + nb,
                                       nb < total, s_ext[s ^ 1], ext4);
        }
        mbar_wait(&s_bar[s], (b >> 1) & 1);
        const int count = min(FWD_BATCH, total - b * FWD_BATCH);
        traversed = b * FWD_BATCH + count;
        const float4* __restrict__ buf = s_rec[s];
        const float4* __restrict__ ebuf = s_ext[s];
        // batch-level cull: 8 independent tests per lane; survivors compacted (in list order) into the
        // warp's private index list
        uint8_t* __restrict__ surv = s_surv[warp];
        int nsurv = 0;
        // a warp whose 32 pixels have all stopped only helps staging from here on: no cull, no evaluation (the tile
        // runs until its LAST pixel stops -- a silhouette pixel walks the whole list -- so most warps are in this state
        // for most batches)
        const bool warp_done = __all_sync(0xffffffffu, done);
#pragma unroll
        for (int c = 0; c < FWD_BATCH / 32; ++c) {
            if (warp_done) break;
            const int e = c * 32 + lane;
            bool keep = false;
            if (e < count) {
                const float4 q0 = buf[3 * e];
                const float4 q1 = buf[3 * e + 1];
                const float thr = buf[3 * e + 2].z;
                keep = cull_keep(q0, q1.x, thr, X0, X1, Y0, Y1);
            }
            const uint32_t bal = __ballot_sync(0xffffffffu, keep);
            if (keep) surv[nsurv + __popc(bal & lt_mask)] = (uint8_t)e;
            nsurv += __popc(bal);
        }
        __syncwarp();
        for (int i = 0; i < nsurv; i += ILP) {
            if (__all_sync(0xffffffffu, done)) break;
            // ILP survivors: alpha evaluated independently, then blended in list order (predicated, branch free)
            const uint32_t packed = *reinterpret_cast<const uint32_t*>(surv + i);
            int j[ILP];
            float alpha[ILP], cr[ILP], cg[ILP], cb[ILP], cd[ILP];
            float4 ex[ILP];
            bool ok[ILP];
#pragma unroll
            for (int k = 0; k < ILP; ++k) {
                const bool has = i + k < nsurv;
                j[k] = has ? (int)((packed >> (8 * k)) & 0xffu) : (int)(packed & 0xffu);
                const float4 q0 = buf[3 * j[k]];
                const float4 q1 = buf[3 * j[k] + 1];
                const float2 q2 = *reinterpret_cast<const float2*>(buf + 3 * j[k] + 2);
                if (EXT) ex[k] = ebuf[j[k]];
                const float dx = q0.x - pixx, dy = q0.y - pixy;
                const float power = -0.5f * (q0.z * dx * dx + q1.x * dy * dy) - q0.w * dx * dy;
                alpha[k] = fminf(ALPHA_MAX, q1.y * __expf(power));
                ok[k] = has && (power <= 0.0f) && (alpha[k] >= ALPHA_MIN);
                cr[k] = q1.w, cg[k] = q2.x, cb[k] = q2.y, cd[k] = q1.z;
            }
#pragma unroll
            for (int k = 0; k < ILP; ++k) {
                const bool act = ok[k] && !done;
                const float test_T = T * (1.0f - alpha[k]);
                const bool stop = act && (test_T < T_MIN);
                const bool use = act && !stop;
                const uint32_t position = (uint32_t)(b * FWD_BATCH + j[k] + 1);
                if (use) {
                    const float w = alpha[k] * T;
                    C0 += cr[k] * w;
                    C1 += cg[k] * w;
                    C2 += cb[k] * w;
                    Wt += w;
                    D += cd[k] * w;
                    if (EXT) E0 += ex[k].x * w, E1 += ex[k].y * w, E2 += ex[k].z * w, E3 += ex[k].w * w;
                    T = test_T;
                    last_contributor = position;
                }
                if (stop) {
                    done = true;
                    visited = position;
                }
            }
        }
    }
    if (inside) {
        const int pix = pyi * W + pxi;
        const size_t HW = (size_t)H * W;
        n_contrib[pix] = last_contributor;
        n_visited[pix] = visited ? visited : (uint32_t)traversed;
        final_T[pix] = T;
        out_color[pix] = C0 + T * bg[0];
        out_color[HW + pix] = C1 + T * bg[1];
        out_color[2 * HW + pix] = C2 + T * bg[2];
        out_depth[pix] = D;
        out_alpha[pix] = Wt;
        if (EXT) {
            const float E[EXT_FLOATS] = {E0, E1, E2, E3};
#pragma unroll
            for (int c = 0; c < EXT_FLOATS; ++c)
                if (c < tab.n_extra) vt.out_extra[c * HW + pix] = E[c];
        }
    }
}

// ============================================================================================
// K7 backward
// ============================================================================================

// Sum 10 per-lane values over the 32 lanes with 12 shuffles (transposed butterfly): after the call the
// lane whose slot (see reduce_slot) is k holds the warp total of v[k] in the return value.
__device__ __forceinline__ float warp_reduce10(const float v[10], int lane) {
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
    float a[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const float send = b4 ? v[k] : v[k + 5];
        const float recv = __shfl_xor_sync(0xffffffffu, send, 16);
        a[k] = (b4 ? v[k + 5] : v[k]) + recv;
    }
    float c[3];
    {
        // keep (a0,a1,a2) when b3 == 0, (a3,a4,-) when b3 == 1
        float send = b3 ? a[0] : a[3];
        float recv = __shfl_xor_sync(0xffffffffu, send, 8);
        c[0] = (b3 ? a[3] : a[0]) + recv;
        send = b3 ? a[1] : a[4];
        recv = __shfl_xor_sync(0xffffffffu, send, 8);
        c[1] = (b3 ? a[4] : a[1]) + recv;
        send = b3 ? a[2] : 0.f;
        recv = __shfl_xor_sync(0xffffffffu, send, 8);
        c[2] = (b3 ? 0.f : a[2]) + recv;
    }
    float d[2];
    {
        // keep (c0,c1) when b2 == 0, (c2,-) when b2 == 1
        float send = b2 ? c[0] : c[2];
        float recv = __shfl_xor_sync(0xffffffffu, send, 4);
        d[0] = (b2 ? c[2] : c[0]) + recv;
        send = b2 ? c[1] : 0.f;
        recv = __shfl_xor_sync(0xffffffffu, send, 4);
        d[1] = (b2 ? 0.f : c[1]) + recv;
    }
    float e;
    {
        const float send = b1 ? d[0] : d[1];
        const float recv = __shfl_xor_sync(0xffffffffu, send, 2);
        e = (b1 ? d[1] : d[0]) + recv;
    }
    e += __shfl_xor_sync(0xffffffffu, e, 1);
    return e;
}
// index k (0..9) of the value a lane ends up holding, or -1
__device__ __forceinline__ int reduce_slot(int lane) {
    if (lane & 1) return -1;
    const int b4 = (lane >> 4) & 1, b3 = (lane >> 3) & 1, b2 = (lane >> 2) & 1, b1 = (lane >> 1) & 1;
    int k;
    if (!b3) {
        if (!b2) k = b1; else k = b1 ? -1 : 2;
    } else {
        if (!b2) k = 3 + b1; else k = -1;
    }
    return k < 0 ? -1 : 5 * b4 + k;
}

// Sum 4 per-lane values over the warp with 6 shuffles; afterwards the lanes with (lane & 7) == 0 hold the total of
// v[2 * bit4 + bit3] (reduce_slot4).
__device__ __forceinline__ float warp_reduce4(const float v[4], int lane) {
    const bool b4 = lane & 16, b3 = lane & 8;
    float a0, a1;
    {
        float send = b4 ? v[0] : v[2];
        float recv = __shfl_xor_sync(0xffffffffu, send, 16);
        a0 = (b4 ? v[2] : v[0]) + recv;
        send = b4 ? v[1] : v[3];
        recv = __shfl_xor_sync(0xffffffffu, send, 16);
        a1 = (b4 ? v[3] : v[1]) + recv;
    }
    float c;
    {
        const float send = b3 ? a0 : a1;
        const float recv = __shfl_xor_sync(0xffffffffu, send, 8);
        c = (b3 ? a1 : a0) + recv;
    }
    c += __shfl_xor_sync(0xffffffffu, c, 4);
    c += __shfl_xor_sync(0xffffffffu, c, 2);
    c += __shfl_xor_sync(0xffffffffu, c, 1);
    return c;
}
__device__ __forceinline__ int reduce_slot4(int lane) {
    return (lane & 7) ? -1 : 2 * ((lane >> 4) & 1) + ((lane >> 3) & 1);
}

template <bool BULK, bool EXT>
__global__ void __launch_bounds__(32, EXT ? 16 : 1)   // EXT: cap at 128 registers (150 uncapped -> 13 warps per SM)
render_backward_kernel(const __grid_constant__ BatchTab tab, int sel) {
    __shared__ __align__(16) float4 s_rec[STAGES][BATCH * 3];
    __shared__ __align__(16) float4 s_ext[STAGES][EXT ? BATCH : 1];
    __shared__ uint32_t s_ids[STAGES][BATCH];
    __shared__ __align__(8) uint64_t s_bar[STAGES];
    __shared__ __align__(16) uint8_t s_surv[BATCH + 16];

    const int W = tab.W, H = tab.H, grid_x = tab.grid_x;
    const int n_tiles = grid_x * tab.grid_y;
    const uint32_t entry = tab.tile_order[blockIdx.x / WARPS_PER_TILE];   // longest lists of the batch first (LPT)
    const int wblock = blockIdx.x % WARPS_PER_TILE;
    const ViewTab& vt = tab.v[entry / n_tiles];
    const int tile = (int)(entry % n_tiles);
    const int tile_x = tile % grid_x, tile_y = tile / grid_x;
    const PointList point_list{tab.idx_bits ? vt.keys[sel] : nullptr, vt.vals[sel],
                               tab.idx_bits ? (uint32_t)((1ull << tab.idx_bits) - 1ull) : 0xffffffffu};
    const float* __restrict__ rec = vt.rec;
    const float4* __restrict__ ext4 = tab.ext4;
    const float* __restrict__ bg = vt.bg;
    float* __restrict__ grad2d = vt.grad2d;
    float* __restrict__ gradext = vt.gradext;
    int lx, ly;
    thread_pixel(wblock, lx, ly);
    const int lane = threadIdx.x;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const int pxi = tile_x * BLOCK_X + lx, pyi = tile_y * BLOCK_Y + ly;
    const bool inside = pxi < W && pyi < H;
    const float pixx = (float)pxi, pixy = (float)pyi;
    const float X0 = (float)(tile_x * BLOCK_X + (wblock & 1) * 8), X1 = X0 + 7.f;
    const float Y0 = (float)(tile_y * BLOCK_Y + (wblock >> 1) * 4), Y1 = Y0 + 3.f;
    const int pix = pyi * W + pxi;
    const size_t HW = (size_t)H * W;
    const uint32_t r0 = vt.ranges[2 * tile];

    const uint32_t my_last = inside ? vt.n_contrib[pix] : 0u;
    uint32_t warp_last = my_last;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) warp_last = max(warp_last, __shfl_xor_sync(0xffffffffu, warp_last, o));
    const int total = (int)warp_last;  // entries [0,total) of the tile's list can have contributed to this block
    if (total == 0) return;
    const int rounds = (total + BATCH - 1) / BATCH;
    const int slot = reduce_slot(lane);
    const int slot4 = reduce_slot4(lane);

    const float T_final = inside ? vt.final_T[pix] : 0.0f;
    float T = T_final;
    float gC0 = 0.f, gC1 = 0.f, gC2 = 0.f, gD = 0.f, gA = 0.f;
    if (inside) {
        if (vt.dL_dcolor) gC0 = vt.dL_dcolor[pix], gC1 = vt.dL_dcolor[HW + pix], gC2 = vt.dL_dcolor[2 * HW + pix];
        if (vt.dL_ddepth) gD = vt.dL_ddepth[pix];
        if (vt.dL_dalpha) gA = vt.dL_dalpha[pix];
    }
    float gE[EXT_FLOATS] = {0.f, 0.f, 0.f, 0.f}, accE[EXT_FLOATS] = {0.f, 0.f, 0.f, 0.f},
          lE[EXT_FLOATS] = {0.f, 0.f, 0.f, 0.f};
    if (EXT && inside && vt.dL_dextra) {
#pragma unroll
        for (int c = 0; c < EXT_FLOATS; ++c)
            if (c < tab.n_extra) gE[c] = vt.dL_dextra[c * HW + pix];
    }
    const float bg_dot = bg[0] * gC0 + bg[1] * gC1 + bg[2] * gC2;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, accD = 0.f, accA = 0.f;
    float last_alpha = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f, lD = 0.f;

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&s_bar[s], BATCH);
        mbar_fence_init();
    }
    __syncwarp();
    // batch b holds list positions total-1-(b*64+t), t = 0..63 (back to front)
    auto stage = [&](int b) {
        const int s = b % STAGES;
#pragma unroll
        for (int u = 0; u < BATCH / 32; ++u) {
            const int sl = lane + 32 * u;
            const int p = b * BATCH + sl;
            stage_entry<BULK, EXT>(s_rec[s], s_ids[s], &s_bar[s], rec, point_list, sl, (int64_t)r0 + (total - 1 - p),
                                   p < total, s_ext[s], ext4);
        }
    };
    stage(0);
    if (rounds > 1) stage(1);
    for (int b = 0; b < rounds; ++b) {
        __syncwarp();   // every lane finished with the slot refilled below (batch b-1's)
        if (b + 2 < rounds) stage(b + 2);
        const int s = b % STAGES;
        mbar_wait(&s_bar[s], (b / STAGES) & 1);
        const int count = min(BATCH, total - b * BATCH);
        const float4* __restrict__ buf = s_rec[s];
        int nsurv = 0;
#pragma unroll
        for (int c = 0; c < BATCH / 32; ++c) {
            const int e = c * 32 + lane;
            bool keep = false;
            if (e < count) {
                const float4 q0 = buf[3 * e];
                const float4 q1 = buf[3 * e + 1];
                const float thr = buf[3 * e + 2].z;
                keep = cull_keep(q0, q1.x, thr, X0, X1, Y0, Y1);
            }
            const uint32_t bal = __ballot_sync(0xffffffffu, keep);
            if (keep) s_surv[nsurv + __popc(bal & lt_mask)] = (uint8_t)e;
            nsurv += __popc(bal);
        }
        __syncwarp();
        for (int i = 0; i < nsurv; i += ILP) {
            const uint32_t packed = *reinterpret_cast<const uint32_t*>(s_surv + i);
            int j[ILP];
            float v[ILP][10];
            bool hit[ILP];
            float G[ILP], alpha[ILP], inv1m[ILP], ddx[ILP], ddy[ILP], cA[ILP], cB[ILP], cC[ILP], op[ILP];
            float cr[ILP], cg[ILP], cb[ILP], cd[ILP];
            float ve[ILP][EXT_FLOATS];
            float4 ex[ILP];
#pragma unroll
            for (int k = 0; k < ILP; ++k) {
                const bool has = i + k < nsurv;
                j[k] = has ? (int)((packed >> (8 * k)) & 0xffu) : (int)(packed & 0xffu);
                const float4 q0 = buf[3 * j[k]];
                const float4 q1 = buf[3 * j[k] + 1];
                const float2 q2 = *reinterpret_cast<const float2*>(buf + 3 * j[k] + 2);
                if (EXT) ex[k] = s_ext[s][j[k]];
                const uint32_t q = (uint32_t)(total - 1 - (b * BATCH + j[k]));
                ddx[k] = q0.x - pixx, ddy[k] = q0.y - pixy;
                cA[k] = q0.z, cB[k] = q0.w, cC[k] = q1.x, op[k] = q1.y;
                const float power = -0.5f * (cA[k] * ddx[k] * ddx[k] + cC[k] * ddy[k] * ddy[k]) - cB[k] * ddx[k] * ddy[k];
                G[k] = __expf(power);
                alpha[k] = fminf(ALPHA_MAX, op[k] * G[k]);
                inv1m[k] = __fdividef(1.0f, 1.0f - alpha[k]);   // MUFU.RCP + FMUL; 1 - alpha is in [0.01, 1]
                hit[k] = has && (q < my_last) && (power <= 0.0f) && (alpha[k] >= ALPHA_MIN);
                cr[k] = q1.w, cg[k] = q2.x, cb[k] = q2.y, cd[k] = q1.z;
            }
            // sequential recurrences, predicated per lane
#pragma unroll
            for (int k = 0; k < ILP; ++k) {
#pragma unroll
                for (int t = 0; t < 10; ++t) v[k][t] = 0.f;
                if (EXT) {
#pragma unroll
                    for (int c = 0; c < EXT_FLOATS; ++c) ve[k][c] = 0.f;
                }
                if (hit[k]) {
                    const float a = alpha[k], dx = ddx[k], dy = ddy[k];
                    T = T * inv1m[k];
                    const float w = a * T;
                    const float om = 1.0f - last_alpha;
                    float dL_da = 0.f;
                    acc0 = last_alpha * lc0 + om * acc0;
                    lc0 = cr[k];
                    dL_da += (cr[k] - acc0) * gC0;
                    acc1 = last_alpha * lc1 + om * acc1;
                    lc1 = cg[k];
                    dL_da += (cg[k] - acc1) * gC1;
                    acc2 = last_alpha * lc2 + om * acc2;
                    lc2 = cb[k];
                    dL_da += (cb[k] - acc2) * gC2;
                    accD = last_alpha * lD + om * accD;
                    lD = cd[k];
                    dL_da += (cd[k] - accD) * gD;
                    accA = last_alpha + om * accA;
                    dL_da += (1.0f - accA) * gA;
                    if (EXT) {
                        const float e[EXT_FLOATS] = {ex[k].x, ex[k].y, ex[k].z, ex[k].w};
#pragma unroll
                        for (int c = 0; c < EXT_FLOATS; ++c) {
                            accE[c] = last_alpha * lE[c] + om * accE[c];
                            lE[c] = e[c];
                            dL_da += (e[c] - accE[c]) * gE[c];
                            ve[k][c] = w * gE[c];
                        }
                    }
                    dL_da *= T;
                    last_alpha = a;
                    dL_da += (-T_final * inv1m[k]) * bg_dot;
                    const float dL_dG = op[k] * dL_da;
                    const float gdx = G[k] * dx, gdy = G[k] * dy;
                    // power = -0.5 (A dx^2 + C dy^2) - B dx dy, d = mean - pixel
                    v[k][0] = dL_dG * (-gdx * cA[k] - gdy * cB[k]);
                    v[k][1] = dL_dG * (-gdy * cC[k] - gdx * cB[k]);
                    v[k][2] = -0.5f * gdx * dx * dL_dG;
                    v[k][3] = -gdx * dy * dL_dG;
                    v[k][4] = -0.5f * gdy * dy * dL_dG;
                    v[k][5] = G[k] * dL_da;
                    v[k][6] = w * gC0, v[k][7] = w * gC1, v[k][8] = w * gC2, v[k][9] = w * gD;
                }
            }
            // ILP independent warp reductions (interleaved by the scheduler), one 10-lane atomic each
            uint32_t anyhit = 0;
#pragma unroll
            for (int k = 0; k < ILP; ++k) anyhit |= (__any_sync(0xffffffffu, hit[k]) ? 1u : 0u) << k;
            if (anyhit) {
                float r[ILP];
#pragma unroll
                for (int k = 0; k < ILP; ++k) r[k] = warp_reduce10(v[k], lane);
#pragma unroll
                for (int k = 0; k < ILP; ++k)
                    if (slot >= 0 && ((anyhit >> k) & 1u))
                        atomicAdd(grad2d + (size_t)s_ids[s][j[k]] * GRAD2D_FLOATS + slot, r[k]);
                if (EXT) {
                    float re[ILP];
#pragma unroll
                    for (int k = 0; k < ILP; ++k) re[k] = warp_reduce4(ve[k], lane);
#pragma unroll
                    for (int k = 0; k < ILP; ++k)
                        if (slot4 >= 0 && slot4 < tab.n_extra && ((anyhit >> k) & 1u))
                            atomicAdd(gradext + (size_t)s_ids[s][j[k]] * EXT_FLOATS + slot4, re[k]);
                }
            }
        }
    }
}

// staging mode: -1 = default (environment B200SPLAT_STAGING, else LDGSTS), 0 = LDGSTS, 1 = bulk (UBLKCP)
static std::atomic<int> g_staging{-1};
int set_staging_mode(int mode) { return g_staging.exchange(mode < 0 ? -1 : (mode ? 1 : 0)); }
static bool use_bulk_staging() {
    const int m = g_staging.load(std::memory_order_relaxed);
    if (m >= 0) return m == 1;
    static const int env_mode = [] {
        const char* e = getenv("B200SPLAT_STAGING");
        return (e && strcmp(e, "bulk") == 0) ? 1 : 0;
    }();
    return env_mode == 1;
}

cudaError_t launch_render_forward(const BatchTab& tab, int sel, cudaStream_t st) {
    const unsigned grid = (unsigned)(tab.V * tab.grid_x * tab.grid_y);
    const bool ext = tab.n_extra > 0;
    if (use_bulk_staging()) {
        if (ext) render_forward_kernel<true, true><<<grid, BLOCK_SIZE, 0, st>>>(tab, sel);
        else render_forward_kernel<true, false><<<grid, BLOCK_SIZE, 0, st>>>(tab, sel);
    } else {
        if (ext) render_forward_kernel<false, true><<<grid, BLOCK_SIZE, 0, st>>>(tab, sel);
        else render_forward_kernel<false, false><<<grid, BLOCK_SIZE, 0, st>>>(tab, sel);
    }
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_render_backward(const BatchTab& tab, int sel, cudaStream_t st) {
    const unsigned grid = (unsigned)(tab.V * tab.grid_x * tab.grid_y * WARPS_PER_TILE);
    const bool ext = tab.n_extra > 0;
    if (use_bulk_staging()) {
        if (ext) render_backward_kernel<true, true><<<grid, 32, 0, st>>>(tab, sel);
        else render_backward_kernel<true, false><<<grid, 32, 0, st>>>(tab, sel);
    } else {
        if (ext) render_backward_kernel<false, true><<<grid, 32, 0, st>>>(tab, sel);
        else render_backward_kernel<false, false><<<grid, 32, 0, st>>>(tab, sel);
    }
    count_launch();
    return cudaGetLastError();
}

}  // namespace b200splat

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
//  * a backward warp consumes its tile's Gaussian list in batches of 64 entries through a 2-stage shared-memory
//    ring behind mbarriers: each lane gathers two 48-byte records per batch (common.cuh REC layout) either
//    with 16-byte cp.async (LDGSTS, default) whose completion arrives on the stage's mbarrier, or with bulk async
//    copies (cp.async.bulk -> UBLKCP, complete_tx on the mbarrier; measured 20 % slower for 48-byte granules);
//  * warp-cooperative culling: for every 32 staged entries, lane l bounds entry l's best-case alpha
//    over the warp's 8x4 pixel rectangle (exact minimum of the conic quadratic over the rectangle
//    edges); only entries that can reach alpha >= 1/255 somewhere in the rectangle are evaluated by
//    the 32 pixels.  Culling is conservative, so the set of blended (pixel, Gaussian) pairs -- and
//    therefore the image, n_contrib and the gradients -- are exactly those of the unculled loop;
//  * forward terminates a tile as soon as every pixel is saturated (T' < 1e-4), and enters each of its 8x4 blocks
//    into the counting sort of the backward's work order (longest walk first);
//  * backward walks back to front starting at the block's largest n_contrib, with ONE scalar suffix-sum recurrence
//    per pixel, and reduces the 10 per-Gaussian partial gradients across the warp through a shared-memory
//    transpose (see K7 below) before three 16-byte vector atomics per (warp, Gaussian).
#include "common.cuh"

#include <atomic>
#include <cstdlib>
#include <cstring>

namespace b200splat {

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
template <bool BULK, bool EXT, int MINB = 0>
__global__ void __launch_bounds__(BLOCK_SIZE, MINB)
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
    {   // the block's largest n_contrib: how far render backward has to walk the list for these 32 pixels; the block
        // enters the counting sort of render backward's work order here (bucket of the walk length, rank in the bucket)
        const uint32_t wl = __reduce_max_sync(0xffffffffu, inside ? last_contributor : 0u);
        if (lane == 0) {
            vt.block_last[tile * WARPS_PER_TILE + warp] = wl;
            uint32_t code = BLOCK_CODE_NONE;
            if (wl) {
                const uint32_t bucket = order_bucket(wl);
                code = bucket << 20 | atomicAdd(tab.block_hist + bucket, 1u);
            }
            tab.block_code[(size_t)entry * WARPS_PER_TILE + warp] = code;
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
//
// Per pixel, walking its tile's list back to front over the entries that were blended in the forward:
//
//     T_i   = T_{i+1} / (1 - alpha_i)                          transmittance in front of entry i
//     dot_i = g_C . c_i + g_D z_i + g_A                        (+ g_E . e_i with extra channels)
//     dL/dalpha_i = dot_i T_i - S_i / (1 - alpha_i),           S_i = sum_{j > i} dot_j alpha_j T_j + T_final (g_C . bg)
//
// i.e. ONE scalar suffix sum S instead of upstream's five normalised per-channel recurrences
// (accum_rec = last_alpha last_c + (1 - last_alpha) accum_rec; (c - accum_rec) g T): the same derivative of
// C = sum_j c_j alpha_j T_j + T_final bg, written with the pixel gradient contracted first.  It costs 5 dependent
// operations per blended pair instead of ~25.
//
// The 10 (14 with extra channels) partial gradients of a Gaussian are summed over the warp's 32 pixels through shared
// memory: every lane stores its partials of the ILP survivors of a group as rows of a [32 lanes][ILP * 12 (16)] matrix
// (3 or 4 STS.128 per survivor), then lane (c, h) sums float4-column c over the 16 rows of half h (16 LDS.128), the two
// halves are combined with 4 shuffles, and the lanes holding a column issue ONE 16-byte vector reduction
// (red.global.add.v4.f32) each into the Gaussian's 48-byte gradient record.  Per survivor that is ~27 instructions
// and 3 vector atomics; the 12-shuffle transposed butterfly it replaces cost 46 (12 SHFL + 22 SEL + 12 FADD) and 10
// scalar atomics.

__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void red_add_v4(float* addr, const float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

// packed fp32 pairs (Blackwell FADD2 / FMUL2 / FFMA2: two fp32 operations per issue slot)
__device__ __forceinline__ float2 add2(const float2 a, const float2 b) {
    float2 r;
    asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc; }"
        : "=f"(r.x), "=f"(r.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 mul2(const float2 a, const float2 b) {
    float2 r;
    asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mul.rn.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc; }"
        : "=f"(r.x), "=f"(r.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}

constexpr int BWD_BATCH = 64;   // list entries per warp batch (two per lane)
// BWD_STAGES: ring depth (3: two batches in flight while one is processed; 2: one in flight, 3.4 KB less shared memory)
template <bool BULK, bool EXT, int BWD_STAGES>
__global__ void __launch_bounds__(32, EXT ? 12 : (BWD_STAGES == 2 ? 16 : 13))
render_backward_kernel(const __grid_constant__ BatchTab tab, int sel) {
    constexpr int F4 = EXT ? 4 : 3;        // float4s of a survivor's reduced record: 10 gradients + 2 pad (+ 4 extra)
    constexpr int NC4 = ILP * F4;          // float4 columns of a group's matrix
    constexpr int RS = NC4 * 4 + 4;        // row stride in floats (52 / 68: STS.128 of 8 neighbouring lanes hit 8 bank groups)
    constexpr int PER_LANE = BWD_BATCH / 32;
    __shared__ __align__(16) float4 s_rec[BWD_STAGES][BWD_BATCH * 3];
    __shared__ __align__(16) float4 s_ext[BWD_STAGES][EXT ? BWD_BATCH : 1];
    __shared__ uint32_t s_ids[BWD_STAGES][BWD_BATCH];
    __shared__ __align__(8) uint64_t s_bar[BWD_STAGES];
    __shared__ __align__(16) float s_red[32 * RS];
    __shared__ __align__(4) uint8_t s_surv[BWD_BATCH + 4];

    const int W = tab.W, H = tab.H, grid_x = tab.grid_x;
    const int n_tiles = grid_x * tab.grid_y;
    // work item: a non-empty 8x4 block of the batch, longest walk first (block_order_kernel); launched over all blocks
    if (blockIdx.x >= tab.block_order[0]) return;
    const uint32_t item = tab.block_order[4 + blockIdx.x];
    const uint32_t entry = item / WARPS_PER_TILE;
    const int wblock = (int)(item % WARPS_PER_TILE);
    const ViewTab& vt = tab.v[entry / n_tiles];
    const int tile = (int)(entry % n_tiles);
    const int tile_x = tile % grid_x, tile_y = tile / grid_x;
    // the sorted list: low halves of the packed pair words (tile << 32 | gaussian), or a plain index array
    const uint32_t* __restrict__ list = tab.idx_bits ? reinterpret_cast<const uint32_t*>(vt.keys[sel]) : vt.vals[sel];
    const int list_stride = tab.idx_bits ? 2 : 1;
    const float* __restrict__ rec = vt.rec;
    const float4* __restrict__ ext4 = tab.ext4;
    const float* __restrict__ bg = vt.bg;
    float* __restrict__ grad2d = vt.grad2d;
    float* __restrict__ gradext = vt.gradext;
    uint8_t* __restrict__ touched = vt.touched;
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
    const int total = (int)__reduce_max_sync(0xffffffffu, my_last);  // entries [0,total) can have contributed to this block
    if (total == 0) return;
    const int rounds = (total + BWD_BATCH - 1) / BWD_BATCH;

    // Batch b holds list positions total-1-(b*64+t), t = 0..63 (back to front), slot t = lane + 32 u.  The list word
    // (-> Gaussian index) of a batch is loaded one iteration before its record is gathered and is not touched until
    // then, so the gather's address is in a register when it is issued: the dependent chain word -> record is split
    // over two iterations instead of stalling the warp in the middle of one.
    auto load_ids = [&](int bb, uint32_t (&id)[PER_LANE]) {
#pragma unroll
        for (int u = 0; u < PER_LANE; ++u) {
            const int p = bb * BWD_BATCH + lane + 32 * u;
            id[u] = 0xffffffffu;
            if (bb < rounds && p < total) id[u] = __ldg(list + (size_t)list_stride * ((size_t)r0 + (size_t)(total - 1 - p)));
        }
    };
    auto stage = [&](int bb, const uint32_t (&id)[PER_LANE]) {
        const int st = bb % BWD_STAGES;
        uint64_t* bar = &s_bar[st];
#pragma unroll
        for (int u = 0; u < PER_LANE; ++u) {
            const int sl = lane + 32 * u;
            if (id[u] != 0xffffffffu) {
                s_ids[st][sl] = id[u];
                const float* src = rec + (size_t)id[u] * REC_FLOATS;
                float4* dst = s_rec[st] + 3 * sl;
                if (BULK) {
                    mbar_arrive_expect_tx(bar, REC_FLOATS * 4 + (EXT ? 16 : 0));
                    bulk_g2s(dst, src, REC_FLOATS * 4, bar);
                    if (EXT) bulk_g2s(s_ext[st] + sl, ext4 + id[u], 16, bar);
                } else {
                    cp_async16(dst, src);
                    cp_async16(dst + 1, src + 4);
                    cp_async16(dst + 2, src + 8);
                    if (EXT) cp_async16(s_ext[st] + sl, ext4 + id[u]);
                    cp_async_arrive_noinc(bar);
                }
            } else {
                mbar_arrive(bar);
            }
        }
    };
    if (lane == 0) {
#pragma unroll
        for (int st = 0; st < BWD_STAGES; ++st) mbar_init(&s_bar[st], BWD_BATCH);
        mbar_fence_init();
    }
    uint32_t id_a[PER_LANE], id_b[PER_LANE], id_next[PER_LANE];
    load_ids(0, id_a);
    load_ids(1, id_b);
    if (BWD_STAGES == 3) load_ids(2, id_next);

    const float T_final = inside ? vt.final_T[pix] : 0.0f;
    float T = T_final;
    float gC0 = 0.f, gC1 = 0.f, gC2 = 0.f, gD = 0.f, gA = 0.f;
    if (inside) {
        if (vt.dL_dcolor) gC0 = vt.dL_dcolor[pix], gC1 = vt.dL_dcolor[HW + pix], gC2 = vt.dL_dcolor[2 * HW + pix];
        if (vt.dL_ddepth) gD = vt.dL_ddepth[pix];
        if (vt.dL_dalpha) gA = vt.dL_dalpha[pix];
    }
    float gE[EXT_FLOATS] = {0.f, 0.f, 0.f, 0.f};
    if (EXT && inside && vt.dL_dextra) {
#pragma unroll
        for (int c = 0; c < EXT_FLOATS; ++c)
            if (c < tab.n_extra) gE[c] = vt.dL_dextra[c * HW + pix];
    }
    // suffix sum of dot_j alpha_j T_j over the entries behind the current one; the background is the last "entry"
    float S = T_final * (bg[0] * gC0 + bg[1] * gC1 + bg[2] * gC2);

    // reduction roles: lane (c4, half) sums float4-column c4 of s_red over rows [16 half, 16 half + 16).  Half 1 walks
    // its rows starting 4 further down (rows 20..31, 16..19): row r of half 0 and row 16 + (r + 4) % 16 of half 1 are
    // 16 banks apart, so the quarter-warp that mixes columns of both halves (lanes 8..15 without extra channels) is
    // conflict free.
    const int c4 = lane % NC4, half = lane / NC4;             // half >= 2: idle in the column sums (non-EXT lanes 24..31)
    const bool red_active = half < 2;
    const float4* red_src = reinterpret_cast<const float4*>(s_red + (red_active && half ? 20 : 0) * RS) + c4;   // steps 0..11
    const float4* red_src2 = red_src + (half == 1 ? -4 : 12) * (RS / 4);                                         // steps 12..15
    float4* my_row = reinterpret_cast<float4*>(s_red + lane * RS);
    const int red_k = c4 / F4, red_part = c4 % F4;            // survivor of the group / float4 of its record this lane owns
    // the two pad floats of every record are zero for the whole kernel (only .xy of its third float4 is ever stored)
#pragma unroll
    for (int k = 0; k < ILP; ++k) my_row[k * F4 + 2] = make_float4(0.f, 0.f, 0.f, 0.f);

    // per-survivor results of phase 1 (registers reused from group to group)
    float a_h[ILP], og_h[ILP], G_h[ILP], inv1m[ILP], ddx[ILP], ddy[ILP], cA[ILP], cB[ILP], cC[ILP], dot[ILP];
    float4 ex[ILP];
    float dot0[ILP];          // EXT: dot without the extra channels' terms
    float S0 = S;             // EXT: the suffix sum of dot0
    uint32_t jpack = 0, anyhit = 0;   // slots of the group's survivors (one byte each); which of them hit some pixel

    __syncwarp();
    constexpr int AHEAD = BWD_STAGES - 1;   // batches in flight
    stage(0, id_a);
    if (AHEAD == 2) {
        if (rounds > 1) stage(1, id_b);
    } else {
#pragma unroll
        for (int u = 0; u < PER_LANE; ++u) id_next[u] = id_b[u];
    }
    for (int b = 0; b < rounds; ++b) {
        __syncwarp();   // every lane finished with the slot refilled below (batch b-1's)
        if (b + AHEAD < rounds) stage(b + AHEAD, id_next);
        load_ids(b + AHEAD + 1, id_next);
        const int st = b % BWD_STAGES;
        mbar_wait(&s_bar[st], (b / BWD_STAGES) & 1);
        const int count = min(BWD_BATCH, total - b * BWD_BATCH);
        const float4* __restrict__ buf = s_rec[st];
        // survivors of the batch, compacted in list order into the warp's byte list (one LDS.32 then hands a group its
        // four slots; walking a 64-bit mask with ffs cost 14 % of the kernel's stall samples)
        int nsurv = 0, gi = 0;
#pragma unroll
        for (int u = 0; u < PER_LANE; ++u) {
            const int e = lane + 32 * u;
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
        // ---- phase 1: ILP survivors evaluated independently (LDS / MUFU latencies overlap) -----------------------
        auto eval_group = [&]() {
            jpack = *reinterpret_cast<const uint32_t*>(s_surv + gi), anyhit = 0;
#pragma unroll
            for (int k = 0; k < ILP; ++k) {
                const bool has = gi + k < nsurv;
                const int j = has ? (int)((jpack >> (8 * k)) & 0xffu) : 0;
                const float4 q0 = buf[3 * j];
                const float4 q1 = buf[3 * j + 1];
                const float2 q2 = *reinterpret_cast<const float2*>(buf + 3 * j + 2);
                const uint32_t q = (uint32_t)(total - 1 - (b * BWD_BATCH + j));
                ddx[k] = q0.x - pixx, ddy[k] = q0.y - pixy;
                cA[k] = q0.z, cB[k] = q0.w, cC[k] = q1.x;
                const float power = -0.5f * (cA[k] * ddx[k] * ddx[k] + cC[k] * ddy[k] * ddy[k]) - cB[k] * ddx[k] * ddy[k];
                const float G = __expf(power);
                const float og = q1.y * G;
                const float alpha = fminf(ALPHA_MAX, og);
                const bool hit = has && (q < my_last) && (power <= 0.0f) && (alpha >= ALPHA_MIN);
                // a pair that was not blended leaves every state variable and every sum unchanged: alpha = 0, 1/(1-alpha) = 1
                a_h[k] = hit ? alpha : 0.f;
                og_h[k] = hit ? og : 0.f;
                G_h[k] = hit ? G : 0.f;
                inv1m[k] = hit ? rcp_approx(1.0f - alpha) : 1.0f;   // MUFU.RCP; 1 - alpha is in [0.01, 1]
                float d = fmaf(gD, q1.z, gA);
                d = fmaf(gC2, q2.y, d), d = fmaf(gC1, q2.x, d), d = fmaf(gC0, q1.w, d);
                if (EXT) {
                    dot0[k] = d;   // without the extra channels: what reaches dL/dmeans2D (see phase 2)
                    ex[k] = s_ext[st][j];
                    d = fmaf(gE[0], ex[k].x, d), d = fmaf(gE[1], ex[k].y, d);
                    d = fmaf(gE[2], ex[k].z, d), d = fmaf(gE[3], ex[k].w, d);
                }
                dot[k] = d;
                anyhit |= (__any_sync(0xffffffffu, hit) ? 1u : 0u) << k;
            }
            gi += ILP;
        };
        if (nsurv > 0) eval_group();
        else anyhit = 0u;
        bool more = true;
        while (more) {
            const uint32_t hit_g = anyhit, jpack_g = jpack;
            if (hit_g) {
                // ---- phase 2: the recurrence, in list order; each survivor's partial gradients go to this lane's row
#pragma unroll
                for (int k = 0; k < ILP; ++k) {
                    // a survivor that hit no pixel of the block (warp-uniform: anyhit is a vote) leaves T and S as they
                    // are and contributes nothing: its row is not written and its column is neither summed nor sent
                    if (!((hit_g >> k) & 1u)) continue;
                    T *= inv1m[k];
                    const float w = a_h[k] * T;
                    const float dL_da = fmaf(dot[k], T, -(S * inv1m[k]));
                    S = fmaf(dot[k], w, S);
                    // alpha = o G (straight through the min), G = exp(power), power = -0.5 (A dx^2 + C dy^2) - B dx dy,
                    // d = mean - pixel
                    const float qq = -og_h[k] * dL_da;
                    const float2 qxy = mul2(make_float2(qq, qq), make_float2(ddx[k], ddy[k]));
                    const float2 dia = mul2(qxy, make_float2(ddx[k], ddy[k]));       // (2 d conic a, 2 d conic c)
                    const float2 wc = mul2(make_float2(w, w), make_float2(gC0, gC1));
                    const float2 wd = mul2(make_float2(w, w), make_float2(gC2, gD));
                    float4* row = my_row + k * F4;
                    // (d mean x, d mean y, 2 d conic a, d conic b) (2 d conic c, d opacity, d r, d g) (d b, d depth, 0, 0):
                    // the factor 1/2 of the two diagonal conic terms is applied by the consumer (preprocess backward)
                    row[0] = make_float4(fmaf(cA[k], qxy.x, cB[k] * qxy.y), fmaf(cC[k], qxy.y, cB[k] * qxy.x), dia.x,
                                         qxy.x * ddy[k]);
                    row[1] = make_float4(dia.y, G_h[k] * dL_da, wc.x, wc.y);
                    if (EXT) {
                        // The reference renders the extra channels with a second rasterizer call whose means2D is a
                        // gradient-free zeros tensor (renderer/diff_gaussian_rasterizer_shading.py:177-187): the extra
                        // channels reach dL/dmeans3D but NOT dL/dmeans2D, the densification statistic
                        // (geometry/gaussian_base.py:815-819).  The mean gradient without their terms rides in the
                        // record's two spare floats; preprocess backward writes THAT to dL/dmeans2D / grad_accum.
                        const float dL_da0 = fmaf(dot0[k], T, -(S0 * inv1m[k]));
                        S0 = fmaf(dot0[k], w, S0);
                        const float q0x = -og_h[k] * dL_da0 * ddx[k], q0y = -og_h[k] * dL_da0 * ddy[k];
                        row[2] = make_float4(wd.x, wd.y, fmaf(cA[k], q0x, cB[k] * q0y), fmaf(cC[k], q0y, cB[k] * q0x));
                        row[3] = make_float4(w * gE[0], w * gE[1], w * gE[2], w * gE[3]);
                    } else {
                        *reinterpret_cast<float2*>(row + 2) = wd;
                    }
                }
                __syncwarp();
            }
            // the next group's phase 1 is independent of this group's reduction: issued first, they overlap
            more = gi < nsurv;
            if (more) eval_group();
            if (hit_g) {
                // ---- phase 3: column sums over the 32 rows, one vector reduction per float4 of a record -----------
                float2 s0 = make_float2(0.f, 0.f), s1 = s0, s2 = s0, s3 = s0;   // (xy, zw) of the even / the odd rows
                if (red_active && ((hit_g >> red_k) & 1u)) {
#pragma unroll
                    for (int r = 0; r < 16; r += 2) {
                        const float4 t = r < 12 ? red_src[r * (RS / 4)] : red_src2[(r - 12) * (RS / 4)];
                        const float4 u = r < 12 ? red_src[(r + 1) * (RS / 4)] : red_src2[(r - 11) * (RS / 4)];
                        s0 = add2(s0, make_float2(t.x, t.y)), s1 = add2(s1, make_float2(t.z, t.w));
                        s2 = add2(s2, make_float2(u.x, u.y)), s3 = add2(s3, make_float2(u.z, u.w));
                    }
                }
                s0 = add2(s0, s2), s1 = add2(s1, s3);
                float4 acc = make_float4(s0.x, s0.y, s1.x, s1.y);
                acc.x += __shfl_down_sync(0xffffffffu, acc.x, NC4);
                acc.y += __shfl_down_sync(0xffffffffu, acc.y, NC4);
                acc.z += __shfl_down_sync(0xffffffffu, acc.z, NC4);
                acc.w += __shfl_down_sync(0xffffffffu, acc.w, NC4);
                if (lane < NC4 && ((hit_g >> red_k) & 1u)) {
                    const size_t id = s_ids[st][(jpack_g >> (8 * red_k)) & 0xffu];
                    if (!EXT || red_part < 3) red_add_v4(grad2d + id * GRAD2D_FLOATS + 4 * red_part, acc);
                    else red_add_v4(gradext + id * EXT_FLOATS, acc);
                    // "this Gaussian has a gradient in this view": what preprocess backward's scan reads instead of
                    // the 48-byte record (a plain byte store; every writer stores the same value)
                    if (red_part == 0) touched[id] = 1;
                }
                __syncwarp();   // the rows are rewritten by the next group
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
        static const int minb = [] {   // B200SPLAT_FWD_MINB=6: cap the registers for 6 CTAs per SM (A/B)
            const char* e = getenv("B200SPLAT_FWD_MINB");
            return e ? atoi(e) : 1;
        }();
        if (ext) render_forward_kernel<false, true><<<grid, BLOCK_SIZE, 0, st>>>(tab, sel);
        else if (minb == 6) render_forward_kernel<false, false, 6><<<grid, BLOCK_SIZE, 0, st>>>(tab, sel);
        else if (minb == 7) render_forward_kernel<false, false, 7><<<grid, BLOCK_SIZE, 0, st>>>(tab, sel);
        else render_forward_kernel<false, false><<<grid, BLOCK_SIZE, 0, st>>>(tab, sel);
    }
    count_launch();
    return cudaGetLastError();
}

static int backward_stages() {
    static const int n = [] {
        const char* e = getenv("B200SPLAT_BWD_STAGES");
        return (e && atoi(e) == 3) ? 3 : 2;   // measured on B200 (headline, 4 views): 2 stages 108.6, 3 stages 110.6 us/view
    }();
    return n;
}

cudaError_t launch_render_backward(const BatchTab& tab, int sel, cudaStream_t st) {
    const unsigned grid = (unsigned)(tab.V * tab.grid_x * tab.grid_y * WARPS_PER_TILE);
    const bool ext = tab.n_extra > 0;
    const bool bulk = use_bulk_staging();
#define LAUNCH_BWD(B, E, S) render_backward_kernel<B, E, S><<<grid, 32, 0, st>>>(tab, sel)
    if (backward_stages() == 2) {
        if (bulk) { if (ext) LAUNCH_BWD(true, true, 2); else LAUNCH_BWD(true, false, 2); }
        else { if (ext) LAUNCH_BWD(false, true, 2); else LAUNCH_BWD(false, false, 2); }
    } else {
        if (bulk) { if (ext) LAUNCH_BWD(true, true, 3); else LAUNCH_BWD(true, false, 3); }
        else { if (ext) LAUNCH_BWD(false, true, 3); else LAUNCH_BWD(false, false, 3); }
    }
#undef LAUNCH_BWD
    count_launch();
    return cudaGetLastError();
}

}  // namespace b200splat

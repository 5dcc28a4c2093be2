// K6 render forward, K7 render backward: per-16x16-tile alpha blending.
//
// Replaces upstream forward.cu renderCUDA and backward.cu renderCUDA (ashawkey variant with depth
// and alpha channels) [UPSTREAM-RECALL]; semantics restated in oracle/torch_oracle.py
// (_blend_tile); reference consumers renderer/diff_gaussian_rasterizer_advanced.py:122-146.
//
// B200 design:
//  * one CTA per tile, 8 warps, each warp owns an 8x4 pixel block (coherent skip tests);
//  * the tile's Gaussian list is consumed in batches of 256 entries; each thread stages one
//    48-byte record (csrc/common.cuh REC layout) with ONE bulk async copy (cp.async.bulk ->
//    UBLKCP) completing on an mbarrier, double buffered, so the gather of batch k+1 overlaps
//    the blend of batch k;
//  * forward terminates a tile as soon as every pixel is saturated (T' < 1e-4);
//  * backward walks back to front starting at the tile's largest n_contrib, skips entries no
//    lane of the warp blended, and reduces the 10 per-Gaussian partial gradients across the
//    warp with shuffles before a single set of global atomics per (warp, Gaussian).
#include "common.cuh"

namespace b200splat {

constexpr int BATCH = 256;
constexpr int STAGES = 2;

// ---- mbarrier / bulk-copy PTX ------------------------------------------------------------------
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

// pixel owned by this thread: warp w -> 8x4 block (w%2, w/2), lane -> (l%8, l/8)
__device__ __forceinline__ void thread_pixel(int& lx, int& ly) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    lx = (warp & 1) * 8 + (lane & 7);
    ly = (warp >> 1) * 4 + (lane >> 3);
}

// stage one batch: thread t copies the record of list entry (first + t) into buf[t]
__device__ __forceinline__ void stage_batch(float4* buf, uint32_t* ids, uint64_t* bar, const float* __restrict__ rec,
                                            const uint32_t* __restrict__ point_list, int64_t pos, bool valid) {
    if (valid) {
        const uint32_t id = __ldg(point_list + pos);
        if (ids) ids[threadIdx.x] = id;
        mbar_arrive_expect_tx(bar, REC_FLOATS * 4);
        bulk_g2s(buf + 3 * threadIdx.x, rec + (size_t)id * REC_FLOATS, REC_FLOATS * 4, bar);
    } else {
        mbar_arrive(bar);
    }
}

// ============================================================================================
// K6 forward
// ============================================================================================
__global__ void __launch_bounds__(BLOCK_SIZE)
render_forward_kernel(int W, int H, int grid_x, const uint32_t* __restrict__ ranges,
                      const uint32_t* __restrict__ point_list, const float* __restrict__ rec,
                      const float* __restrict__ bg, uint32_t* __restrict__ n_contrib, uint32_t* __restrict__ n_visited,
                      float* __restrict__ final_T, float* __restrict__ out_color, float* __restrict__ out_depth, float* __restrict__ out_alpha) {
    __shared__ __align__(16) float4 s_rec[STAGES][BATCH * 3];
    __shared__ __align__(8) uint64_t s_bar[STAGES];

    const int tile = blockIdx.y * grid_x + blockIdx.x;
    int lx, ly;
    thread_pixel(lx, ly);
    const int pxi = blockIdx.x * BLOCK_X + lx, pyi = blockIdx.y * BLOCK_Y + ly;
    const bool inside = pxi < W && pyi < H;
    const float pixx = (float)pxi, pixy = (float)pyi;
    const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
    const int total = (int)(r1 - r0);
    const int rounds = (total + BATCH - 1) / BATCH;

    if (threadIdx.x == 0) {
        mbar_init(&s_bar[0], BLOCK_SIZE);
        mbar_init(&s_bar[1], BLOCK_SIZE);
        mbar_fence_init();
    }
    __syncthreads();

    bool done = !inside;
    float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, Wt = 0.f, D = 0.f;
    uint32_t contributor = 0, last_contributor = 0;

    if (rounds > 0) stage_batch(s_rec[0], nullptr, &s_bar[0], rec, point_list, (int64_t)r0 + threadIdx.x,
                                (int)threadIdx.x < total);
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
            const int nb = (b + 1) * BATCH + threadIdx.x;
            stage_batch(s_rec[s ^ 1], nullptr, &s_bar[s ^ 1], rec, point_list, (int64_t)r0 + nb, nb < total);
        }
        mbar_wait(&s_bar[s], (b >> 1) & 1);
        const int count = min(BATCH, total - b * BATCH);
        const float4* __restrict__ buf = s_rec[s];
        for (int j = 0; !done && j < count; ++j) {
            ++contributor;
            const float4 q0 = buf[3 * j];
            const float4 q1 = buf[3 * j + 1];
            const float dx = q0.x - pixx, dy = q0.y - pixy;
            const float power = -0.5f * (q0.z * dx * dx + q1.x * dy * dy) - q0.w * dx * dy;
            if (power > 0.0f) continue;
            const float alpha = fminf(ALPHA_MAX, q1.y * __expf(power));
            if (alpha < ALPHA_MIN) continue;
            const float test_T = T * (1.0f - alpha);
            if (test_T < T_MIN) {
                done = true;
                continue;
            }
            const float4 q2 = buf[3 * j + 2];
            const float w = alpha * T;
            C0 += q1.w * w;
            C1 += q2.x * w;
            C2 += q2.y * w;
            Wt += w;
            D += q1.z * w;
            T = test_T;
            last_contributor = contributor;
        }
    }
    if (inside) {
        const int pix = pyi * W + pxi;
        const size_t HW = (size_t)H * W;
        n_contrib[pix] = last_contributor;
        final_T[pix] = T;
        n_visited[pix] = contributor;
        out_color[pix] = C0 + T * bg[0];
        out_color[HW + pix] = C1 + T * bg[1];
        out_color[2 * HW + pix] = C2 + T * bg[2];
        out_depth[pix] = D;
        out_alpha[pix] = Wt;
    }
}

// ============================================================================================
// K7 backward
// ============================================================================================
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(BLOCK_SIZE)
render_backward_kernel(int W, int H, int grid_x, const uint32_t* __restrict__ ranges,
                       const uint32_t* __restrict__ point_list, const float* __restrict__ rec,
                       const float* __restrict__ bg, const uint32_t* __restrict__ n_contrib,
                       const float* __restrict__ final_T, const float* __restrict__ dL_dcolor,
                       const float* __restrict__ dL_ddepth, const float* __restrict__ dL_dalpha,
                       float* __restrict__ grad2d) {
    __shared__ __align__(16) float4 s_rec[STAGES][BATCH * 3];
    __shared__ uint32_t s_ids[STAGES][BATCH];
    __shared__ __align__(8) uint64_t s_bar[STAGES];
    __shared__ uint32_t s_max;

    const int tile = blockIdx.y * grid_x + blockIdx.x;
    int lx, ly;
    thread_pixel(lx, ly);
    const int lane = threadIdx.x & 31;
    const int pxi = blockIdx.x * BLOCK_X + lx, pyi = blockIdx.y * BLOCK_Y + ly;
    const bool inside = pxi < W && pyi < H;
    const float pixx = (float)pxi, pixy = (float)pyi;
    const int pix = pyi * W + pxi;
    const size_t HW = (size_t)H * W;
    const uint32_t r0 = ranges[2 * tile];

    if (threadIdx.x == 0) {
        mbar_init(&s_bar[0], BLOCK_SIZE);
        mbar_init(&s_bar[1], BLOCK_SIZE);
        mbar_fence_init();
        s_max = 0;
    }
    __syncthreads();
    const uint32_t my_last = inside ? n_contrib[pix] : 0u;
    {
        uint32_t m = my_last;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0 && m) atomicMax(&s_max, m);
    }
    __syncthreads();
    const int total = (int)s_max;  // entries [0,total) of the tile's list can have contributed
    if (total == 0) return;
    const int rounds = (total + BATCH - 1) / BATCH;

    const float T_final = inside ? final_T[pix] : 0.0f;
    float T = T_final;
    float gC0 = 0.f, gC1 = 0.f, gC2 = 0.f, gD = 0.f, gA = 0.f;
    if (inside) {
        if (dL_dcolor) gC0 = dL_dcolor[pix], gC1 = dL_dcolor[HW + pix], gC2 = dL_dcolor[2 * HW + pix];
        if (dL_ddepth) gD = dL_ddepth[pix];
        if (dL_dalpha) gA = dL_dalpha[pix];
    }
    const float bg_dot = bg[0] * gC0 + bg[1] * gC1 + bg[2] * gC2;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, accD = 0.f, accA = 0.f;
    float last_alpha = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f, lD = 0.f;

    // batch b holds list positions total-1-(b*256+t), t = 0..255 (back to front)
    {
        const int p = (int)threadIdx.x;
        stage_batch(s_rec[0], s_ids[0], &s_bar[0], rec, point_list, (int64_t)r0 + (total - 1 - p), p < total);
    }
    for (int b = 0; b < rounds; ++b) {
        __syncthreads();
        const int s = b & 1;
        if (b + 1 < rounds) {
            const int p = (b + 1) * BATCH + threadIdx.x;
            stage_batch(s_rec[s ^ 1], s_ids[s ^ 1], &s_bar[s ^ 1], rec, point_list, (int64_t)r0 + (total - 1 - p),
                        p < total);
        }
        mbar_wait(&s_bar[s], (b >> 1) & 1);
        const int count = min(BATCH, total - b * BATCH);
        const float4* __restrict__ buf = s_rec[s];
        for (int j = 0; j < count; ++j) {
            const uint32_t q = (uint32_t)(total - 1 - (b * BATCH + j));  // 0-based list position
            const float4 q0 = buf[3 * j];
            const float4 q1 = buf[3 * j + 1];
            const float dx = q0.x - pixx, dy = q0.y - pixy;
            const float power = -0.5f * (q0.z * dx * dx + q1.x * dy * dy) - q0.w * dx * dy;
            const float G = __expf(power);
            const float alpha = fminf(ALPHA_MAX, q1.y * G);
            const bool hit = (q < my_last) && (power <= 0.0f) && (alpha >= ALPHA_MIN);
            if (!__any_sync(0xffffffffu, hit)) continue;
            float v_dx = 0.f, v_dy = 0.f, v_ca = 0.f, v_cb = 0.f, v_cc = 0.f, v_op = 0.f, v_r = 0.f, v_g = 0.f,
                  v_b = 0.f, v_d = 0.f;
            if (hit) {
                const float4 q2 = buf[3 * j + 2];
                T = T / (1.0f - alpha);
                const float w = alpha * T;
                float dL_da = 0.f;
                acc0 = last_alpha * lc0 + (1.0f - last_alpha) * acc0;
                lc0 = q1.w;
                dL_da += (q1.w - acc0) * gC0;
                acc1 = last_alpha * lc1 + (1.0f - last_alpha) * acc1;
                lc1 = q2.x;
                dL_da += (q2.x - acc1) * gC1;
                acc2 = last_alpha * lc2 + (1.0f - last_alpha) * acc2;
                lc2 = q2.y;
                dL_da += (q2.y - acc2) * gC2;
                accD = last_alpha * lD + (1.0f - last_alpha) * accD;
                lD = q1.z;
                dL_da += (q1.z - accD) * gD;
                accA = last_alpha + (1.0f - last_alpha) * accA;
                dL_da += (1.0f - accA) * gA;
                dL_da *= T;
                last_alpha = alpha;
                dL_da += (-T_final / (1.0f - alpha)) * bg_dot;
                v_r = w * gC0, v_g = w * gC1, v_b = w * gC2, v_d = w * gD;
                const float dL_dG = q1.y * dL_da;
                const float gdx = G * dx, gdy = G * dy;
                // power = -0.5 (A dx^2 + C dy^2) - B dx dy, d = mean - pixel
                v_dx = dL_dG * (-gdx * q0.z - gdy * q0.w);
                v_dy = dL_dG * (-gdy * q1.x - gdx * q0.w);
                v_ca = -0.5f * gdx * dx * dL_dG;
                v_cb = -gdx * dy * dL_dG;
                v_cc = -0.5f * gdy * dy * dL_dG;
                v_op = G * dL_da;
            }
            v_dx = warp_sum(v_dx), v_dy = warp_sum(v_dy), v_ca = warp_sum(v_ca), v_cb = warp_sum(v_cb);
            v_cc = warp_sum(v_cc), v_op = warp_sum(v_op), v_r = warp_sum(v_r), v_g = warp_sum(v_g);
            v_b = warp_sum(v_b), v_d = warp_sum(v_d);
            if (lane < 10) {
                float v = v_dx;
                v = lane == 1 ? v_dy : v;
                v = lane == 2 ? v_ca : v;
                v = lane == 3 ? v_cb : v;
                v = lane == 4 ? v_cc : v;
                v = lane == 5 ? v_op : v;
                v = lane == 6 ? v_r : v;
                v = lane == 7 ? v_g : v;
                v = lane == 8 ? v_b : v;
                v = lane == 9 ? v_d : v;
                atomicAdd(grad2d + (size_t)s_ids[s][j] * GRAD2D_FLOATS + lane, v);
            }
        }
    }
}

cudaError_t launch_render_forward(const CameraParams& cam, const uint32_t* ranges, const uint32_t* point_list,
                                  const float* rec, uint32_t* n_contrib, uint32_t* n_visited, float* final_T,
                                  float* out_color, float* out_depth, float* out_alpha, cudaStream_t st) {
    dim3 grid(cam.grid_x, cam.grid_y);
    render_forward_kernel<<<grid, BLOCK_SIZE, 0, st>>>(cam.W, cam.H, cam.grid_x, ranges, point_list, rec, cam.bg,
                                                       n_contrib, n_visited, final_T, out_color, out_depth, out_alpha);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_render_backward(const CameraParams& cam, const uint32_t* ranges, const uint32_t* point_list,
                                   const float* rec, const uint32_t* n_contrib, const float* final_T,
                                   const float* dL_dcolor, const float* dL_ddepth, const float* dL_dalpha,
                                   float* grad2d, cudaStream_t st) {
    dim3 grid(cam.grid_x, cam.grid_y);
    render_backward_kernel<<<grid, BLOCK_SIZE, 0, st>>>(cam.W, cam.H, cam.grid_x, ranges, point_list, rec, cam.bg,
                                                        n_contrib, final_T, dL_dcolor, dL_ddepth, dL_dalpha,
                                                        grad2d);
    count_launch();
    return cudaGetLastError();
}

}  // namespace b200splat

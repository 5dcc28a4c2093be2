// K8 preprocess backward: 2-D stage gradients -> parameter gradients, one fused kernel for a whole
// view batch.
//
// Replaces upstream backward.cu computeCov2DCUDA + preprocessCUDA (+ computeColorFromSH and
// computeCov3D backward) [UPSTREAM-RECALL; SURVEY.md Appendix A]; the gradients it produces are
// the ones geometry/gaussian_base.py:815-819 (means2D.grad) and the Adam groups
// (geometry/gaussian_base.py:470-525) consume.  Conventions kept from upstream (Appendix A.1):
// the FOV clamp zeroes dL/dt.x|y and treats the clamped t.x|y as constant w.r.t. t.z; the conic
// gradient uses 1/(det^2 + 1e-7); culled Gaussians (radii == 0) get exactly-zero gradients;
// dL/dmeans2D is the gradient w.r.t. the NDC position (pixel gradient x W/2, H/2), z = 0.
//
// View batching: one thread owns one Gaussian and loops over the V views of the step, so the parameters
// (and the SH block) are read once, the gradients of all views are summed in registers and written once
// -- instead of V read-modify-write passes -- and Sigma3's backward runs once on the summed dL/dSigma.
// The reference's per-view densification statistics (geometry/gaussian_base.py:815-819, 846-851) are an
// optional fused epilogue.
#include "common.cuh"

namespace b200splat {

constexpr int SH_ROW_F4 = 13;   // staged SH row: 12 float4 + 1 of padding

constexpr float SH_C0 = 0.28209479177387814f;
constexpr float SH_C1 = 0.4886025119029199f;
constexpr float SH_C2_0 = 1.0925484305920792f;
constexpr float SH_C2_1 = -1.0925484305920792f;
constexpr float SH_C2_2 = 0.31539156525252005f;
constexpr float SH_C2_3 = -1.0925484305920792f;
constexpr float SH_C2_4 = 0.5462742152960396f;
constexpr float SH_C3_0 = -0.5900435899266435f;
constexpr float SH_C3_1 = 2.890611442640554f;
constexpr float SH_C3_2 = -0.4570457994644658f;
constexpr float SH_C3_3 = 0.3731763325901154f;
constexpr float SH_C3_4 = -0.4570457994644658f;
constexpr float SH_C3_5 = 1.445305721320277f;
constexpr float SH_C3_6 = -0.5900435899266435f;

// SH basis values and their gradients w.r.t. the unit direction for coefficients 4C .. 4C+3
template <int C>
__device__ __forceinline__ void sh_basis_chunk(float x, float y, float z, float B[4], float Bx[4], float By[4],
                                               float Bz[4]) {
#pragma unroll
    for (int t = 0; t < 4; ++t) B[t] = Bx[t] = By[t] = Bz[t] = 0.f;
    const float xx = x * x, yy = y * y, zz = z * z;
    if (C == 0) {
        B[0] = SH_C0;
        B[1] = -SH_C1 * y, By[1] = -SH_C1;
        B[2] = SH_C1 * z, Bz[2] = SH_C1;
        B[3] = -SH_C1 * x, Bx[3] = -SH_C1;
    } else if (C == 1) {
        B[0] = SH_C2_0 * x * y, Bx[0] = SH_C2_0 * y, By[0] = SH_C2_0 * x;
        B[1] = SH_C2_1 * y * z, By[1] = SH_C2_1 * z, Bz[1] = SH_C2_1 * y;
        B[2] = SH_C2_2 * (2.f * zz - xx - yy), Bx[2] = SH_C2_2 * -2.f * x, By[2] = SH_C2_2 * -2.f * y,
        Bz[2] = SH_C2_2 * 4.f * z;
        B[3] = SH_C2_3 * x * z, Bx[3] = SH_C2_3 * z, Bz[3] = SH_C2_3 * x;
    } else if (C == 2) {
        B[0] = SH_C2_4 * (xx - yy), Bx[0] = SH_C2_4 * 2.f * x, By[0] = SH_C2_4 * -2.f * y;
        B[1] = SH_C3_0 * y * (3.f * xx - yy), Bx[1] = SH_C3_0 * 6.f * x * y, By[1] = SH_C3_0 * (3.f * xx - 3.f * yy);
        B[2] = SH_C3_1 * x * y * z, Bx[2] = SH_C3_1 * y * z, By[2] = SH_C3_1 * x * z, Bz[2] = SH_C3_1 * x * y;
        B[3] = SH_C3_2 * y * (4.f * zz - xx - yy), Bx[3] = SH_C3_2 * -2.f * x * y,
        By[3] = SH_C3_2 * (4.f * zz - xx - 3.f * yy), Bz[3] = SH_C3_2 * 8.f * y * z;
    } else {
        B[0] = SH_C3_3 * z * (2.f * zz - 3.f * xx - 3.f * yy), Bx[0] = SH_C3_3 * -6.f * x * z,
        By[0] = SH_C3_3 * -6.f * y * z, Bz[0] = SH_C3_3 * (6.f * zz - 3.f * xx - 3.f * yy);
        B[1] = SH_C3_4 * x * (4.f * zz - xx - yy), Bx[1] = SH_C3_4 * (4.f * zz - 3.f * xx - yy),
        By[1] = SH_C3_4 * -2.f * x * y, Bz[1] = SH_C3_4 * 8.f * x * z;
        B[2] = SH_C3_5 * z * (xx - yy), Bx[2] = SH_C3_5 * 2.f * x * z, By[2] = SH_C3_5 * -2.f * y * z,
        Bz[2] = SH_C3_5 * (xx - yy);
        B[3] = SH_C3_6 * x * (xx - 3.f * yy), Bx[3] = SH_C3_6 * (3.f * xx - 3.f * yy), By[3] = SH_C3_6 * -6.f * x * y;
    }
}

// one chunk of 4 coefficients, one view: acc[12] += basis * g, and the chunk's part of dL/ddir
template <int C, int K>
__device__ __forceinline__ void sh_chunk_view(const float sv[12], float x, float y, float z, const float g[3],
                                              float acc[12], float& ddx, float& ddy, float& ddz) {
    float B[4], Bx[4], By[4], Bz[4];
    sh_basis_chunk<C>(x, y, z, B, Bx, By, Bz);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        if (4 * C + t < K) {
            const float gs = g[0] * sv[3 * t] + g[1] * sv[3 * t + 1] + g[2] * sv[3 * t + 2];
            ddx += Bx[t] * gs, ddy += By[t] * gs, ddz += Bz[t] * gs;
            acc[3 * t] += B[t] * g[0];
            acc[3 * t + 1] += B[t] * g[1];
            acc[3 * t + 2] += B[t] * g[2];
        }
    }
}

// Phase B of the kernel for chunk C: take the chunk's 12 SH floats once (from the thread's staged row in shared
// memory, or from global), loop over the views (their masked colour gradient and unit direction sit in the
// thread's shared-memory slot), store the chunk's dL/dsh, accumulate each view's dL/ddir in the slot.
// Slot of (view, thread): 3 float4 = (g.r, g.g, g.b, dir.x) (dir.y, dir.z, 1/|d|, -) (ddir.x, ddir.y, ddir.z, -).
template <int C, int K, bool ACC, bool VEC>
__device__ __forceinline__ void sh_chunk_all_views(const float* __restrict__ sh, const float4* sh_row,
                                                   float* __restrict__ dst, int M, int V, uint32_t vis,
                                                   float4* s_slots /* [V][128][3] */, int col) {
    if (4 * C >= K) return;
    float sv[12];
    if (VEC) {
        float4 a, b, c;
        if (sh_row != nullptr) {
            a = sh_row[3 * C], b = sh_row[3 * C + 1], c = sh_row[3 * C + 2];
        } else {
            const float4* in4 = reinterpret_cast<const float4*>(sh) + 3 * C;
            a = __ldg(in4), b = __ldg(in4 + 1), c = __ldg(in4 + 2);
        }
        sv[0] = a.x, sv[1] = a.y, sv[2] = a.z, sv[3] = a.w, sv[4] = b.x, sv[5] = b.y, sv[6] = b.z, sv[7] = b.w;
        sv[8] = c.x, sv[9] = c.y, sv[10] = c.z, sv[11] = c.w;
    } else {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) sv[3 * t + ch] = (4 * C + t < K) ? __ldg(sh + 3 * (4 * C + t) + ch) : 0.f;
        }
    }
    float acc[12];
#pragma unroll
    for (int t = 0; t < 12; ++t) acc[t] = 0.f;
    for (int v = 0; v < V; ++v) {
        if (!((vis >> v) & 1u)) continue;
        float4* r = s_slots + ((size_t)v * 128 + col) * 3;
        const float4 ra = r[0], rb = r[1];
        float4 rc = r[2];
        const float g[3] = {ra.x, ra.y, ra.z};
        sh_chunk_view<C, K>(sv, ra.w, rb.x, rb.y, g, acc, rc.x, rc.y, rc.z);
        r[2] = rc;
    }
    if (VEC) {
        float4* d4 = reinterpret_cast<float4*>(dst) + 3 * C;
        if (ACC) {
            const float4 p0 = d4[0], p1 = d4[1], p2 = d4[2];
            d4[0] = make_float4(p0.x + acc[0], p0.y + acc[1], p0.z + acc[2], p0.w + acc[3]);
            d4[1] = make_float4(p1.x + acc[4], p1.y + acc[5], p1.z + acc[6], p1.w + acc[7]);
            d4[2] = make_float4(p2.x + acc[8], p2.y + acc[9], p2.z + acc[10], p2.w + acc[11]);
        } else {
            d4[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
            d4[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
            d4[2] = make_float4(acc[8], acc[9], acc[10], acc[11]);
        }
    } else {
#pragma unroll
        for (int t = 0; t < 12; ++t) {
            if (4 * C + t / 3 < K) {
                float* p = dst + 12 * C + t;
                if (ACC) *p += acc[t]; else *p = acc[t];
            }
        }
    }
}

template <bool ACC>
__device__ __forceinline__ void put(float* p, float v) {
    if (ACC) *p += v; else *p = v;
}

// ---------------------------------------------------------------------------------------------------------------
// Scan kernel: one thread per Gaussian, streaming, full occupancy.  Which views saw the Gaussian (radii > 0), and in
// which of those did render backward leave a gradient at all (its `touched` byte)?  Most pairs of a dense scene sit
// behind the saturation point of their pixels: the Gaussian is visible, its 48-byte gradient record is all zeros, and
// every term it would add to the parameter gradients is an exact zero.  Such a view is dropped here.  A Gaussian without any gradient
// only gets its zero outputs and its statistics written -- no parameters, no SH row, no arithmetic; the others are
// appended (index, live views, clamp bits) to the work list of preprocess_backward_kernel, whose warps are then full
// of Gaussians that do have work.  (In the headline scene ~1/6 of the Gaussians of a 4-view step carry a gradient.)
//   * zero outputs of EVERY Gaussian of the range are written here, warp-cooperatively for the SH block (the 32
//     Gaussians of a warp own 32 x 12 M contiguous bytes: full-line stores); the worker overwrites its Gaussians;
//   * densification statistics that need no gradient (visibility count, largest radius) are applied here.
// ---------------------------------------------------------------------------------------------------------------
template <bool ACC>
__global__ void __launch_bounds__(256)
preprocess_backward_scan_kernel(const __grid_constant__ BatchTab tab, float* __restrict__ dL_dmeans3D,
                                float* __restrict__ dL_dshs, float* __restrict__ dL_dcolors,
                                float* __restrict__ dL_dopacity, float* __restrict__ dL_dscales,
                                float* __restrict__ dL_drotations, float* __restrict__ dL_dcov3D,
                                float* __restrict__ stat_denom, float* __restrict__ stat_max_radii, int has_clamp,
                                int sh_vec, uint2* __restrict__ live_list, uint32_t* __restrict__ live_count,
                                int g_begin, int g_end) {
    const int V = tab.V, M = tab.M;
    const int lane = threadIdx.x & 31;
    const int warp_first = g_begin + (int)(blockIdx.x * blockDim.x + (threadIdx.x & ~31u));
    const int idx = warp_first + lane;
    const bool in_range = idx < g_end;
    uint32_t vis = 0, live = 0, clamp3 = 0;
    if (in_range) {
        // every load of the thread goes out before anything is consumed (one round of memory latency instead of three:
        // the kernel is a stream of independent 1-16 byte reads and zero stores).  The bytes of a view that did not
        // see the Gaussian are zero, so they are read unconditionally.
        int rad[MAX_VIEWS];
        uint8_t tch[MAX_VIEWS], clp[MAX_VIEWS];
#pragma unroll
        for (int v = 0; v < MAX_VIEWS; ++v) {
            rad[v] = v < V ? tab.v[v].radii[idx] : 0;
            tch[v] = v < V ? tab.v[v].touched[idx] : (uint8_t)0;
            clp[v] = (v < V && has_clamp) ? tab.v[v].clamped[idx] : (uint8_t)0;
        }
        const float sd = stat_denom ? stat_denom[idx] : 0.f;
        const float sm = stat_max_radii ? stat_max_radii[idx] : 0.f;
        int max_radius = 0;
#pragma unroll
        for (int v = 0; v < MAX_VIEWS; ++v) {
            if (rad[v] > 0) vis |= 1u << v;
            max_radius = max(max_radius, rad[v]);
            // which of the views left a gradient: render backward's byte per (view, Gaussian); consumed here (a set
            // byte is cleared, after it has been read: self-cleaning like the records)
            if (tch[v]) {
                live |= 1u << v;
                clamp3 |= ((uint32_t)clp[v] & 7u) << (3 * v);
                tab.v[v].touched[idx] = 0;
            }
        }
        // the row-sparse exchange's live map; an accumulating call (a later view group of the same step) only adds
        if (tab.live_map && (!ACC || live)) tab.live_map[idx] = live ? (uint8_t)1 : (uint8_t)0;
        if (stat_denom && vis) stat_denom[idx] = sd + (float)__popc(vis);
        if (stat_max_radii && vis) stat_max_radii[idx] = fmaxf(sm, (float)max_radius);
        if (!ACC) {
            for (int v = 0; v < V; ++v) {      // views without a gradient: dL/dmeans2D = 0
                float* m2 = tab.v[v].dL_dmeans2D;
                if (m2 && !((live >> v) & 1u)) m2[3 * idx] = m2[3 * idx + 1] = m2[3 * idx + 2] = 0.f;
            }
            if (live == 0) {
                dL_dmeans3D[3 * idx] = dL_dmeans3D[3 * idx + 1] = dL_dmeans3D[3 * idx + 2] = 0.f;
                dL_dopacity[idx] = 0.f;
                if (dL_dshs && !sh_vec) for (int k = 0; k < 3 * M; ++k) dL_dshs[(size_t)idx * 3 * M + k] = 0.f;
                if (dL_dcolors) dL_dcolors[3 * idx] = dL_dcolors[3 * idx + 1] = dL_dcolors[3 * idx + 2] = 0.f;
                if (dL_dscales) dL_dscales[3 * idx] = dL_dscales[3 * idx + 1] = dL_dscales[3 * idx + 2] = 0.f;
                if (dL_drotations) reinterpret_cast<float4*>(dL_drotations)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (dL_dcov3D) for (int k = 0; k < 6; ++k) dL_dcov3D[(size_t)idx * 6 + k] = 0.f;
            }
        }
    }
    // the SH gradients of the warp's Gaussians: one contiguous block, zeroed with full-width vector stores (the worker
    // overwrites the rows of the Gaussians that have a gradient; it runs after this kernel)
    if (!ACC && dL_dshs && sh_vec && warp_first < g_end) {
        const int rows = min(32, g_end - warp_first);
        float4* d4 = reinterpret_cast<float4*>(dL_dshs + (size_t)warp_first * 3 * M);
        const int n4 = rows * (3 * M / 4);
        for (int k = lane; k < n4; k += 32) d4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // append the Gaussians that have work: ONE atomic per CTA (one per warp -- 31 K returning atomics on one address per
    // launch -- was half of this kernel's stall samples)
    __shared__ uint32_t s_wcnt[8], s_base;
    const int warp = threadIdx.x >> 5;
    const uint32_t m = __ballot_sync(0xffffffffu, live != 0);
    if (lane == 0) s_wcnt[warp] = (uint32_t)__popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t total = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) total += s_wcnt[w];
        s_base = total ? atomicAdd(live_count, total) : 0u;
    }
    __syncthreads();
    if (live) {
        uint32_t base = s_base;
        for (int w = 0; w < warp; ++w) base += s_wcnt[w];
        live_list[base + __popc(m & ((1u << lane) - 1u))] = make_uint2((uint32_t)idx, live | (clamp3 << 8));
    }
}

// DEG = -1: colours precomputed
template <int DEG, bool ACC, bool VEC>
__global__ void __launch_bounds__(128)
preprocess_backward_kernel(const __grid_constant__ BatchTab tab, const float* __restrict__ means3D,
                           const float* __restrict__ scales, const float* __restrict__ rotations,
                           const float* __restrict__ shs, const float* __restrict__ cov3D_precomp,
                           float* __restrict__ dL_dmeans3D, float* __restrict__ dL_dshs,
                           float* __restrict__ dL_dcolors, float* __restrict__ dL_dopacity,
                           float* __restrict__ dL_dscales, float* __restrict__ dL_drotations,
                           float* __restrict__ dL_dcov3D, float* __restrict__ stat_grad_accum,
                           const uint2* __restrict__ live_list, const uint32_t* __restrict__ live_count) {
    __shared__ float sV[MAX_VIEWS][16], sP[MAX_VIEWS][16], sC[MAX_VIEWS][4], sS[MAX_VIEWS][4];
    // dynamic shared memory: one 48-byte slot per (view, thread) -- first the landing zone of the view's gradient
    // record (cp.async), then the SH phase's per-view state -- and, on the vector path, the thread's SH row
    // (row stride 13 float4: LDS.128 of neighbouring threads fall in different bank groups)
    extern __shared__ float4 s_dyn4[];
    if (blockIdx.x * blockDim.x >= *live_count) return;
    float4* s_slots = s_dyn4;   // [V][128][3]
    const int V = tab.V;
    float4* sh_row = (VEC && DEG >= 0) ? s_dyn4 + (size_t)V * 128 * 3 + threadIdx.x * SH_ROW_F4 : nullptr;
    for (int i = threadIdx.x; i < V * 16; i += blockDim.x) {
        sV[i >> 4][i & 15] = tab.v[i >> 4].view[i & 15];
        sP[i >> 4][i & 15] = tab.v[i >> 4].proj[i & 15];
    }
    for (int i = threadIdx.x; i < V * 3; i += blockDim.x) sC[i / 3][i % 3] = tab.v[i / 3].campos[i % 3];
    for (int i = threadIdx.x; i < V * 4; i += blockDim.x) {   // (focal_x, focal_y, limx, limy), host or device copy
        const ViewTab& t = tab.v[i >> 2];
        const float host[4] = {t.focal_x, t.focal_y, t.limx, t.limy};
        sS[i >> 2][i & 3] = t.scalars ? t.scalars[i & 3] : host[i & 3];
    }
    __syncthreads();
    const int M = tab.M;
    constexpr int K = DEG < 0 ? 0 : (DEG + 1) * (DEG + 1);
    constexpr int NCH = (K + 3) / 4;          // live chunks of 4 coefficients
    // work item: a Gaussian with a gradient in at least one view, from the list preprocess_backward_scan_kernel wrote
    // (x = index, y = views with a gradient | clamp bits of view v << (8 + 3 v)); launched over the whole range
    // a resident grid (launch_preprocess_backward: a few CTAs per SM) walks the list
    const uint32_t n_items = *live_count;
    for (uint32_t item_no = blockIdx.x * blockDim.x + threadIdx.x; item_no < n_items; item_no += gridDim.x * blockDim.x) {
    const uint2 item = live_list[item_no];
    const int idx = (int)item.x;
    const uint32_t vis = item.y & 0xffu;           // from here on "vis" = the views with a gradient
    const uint32_t clamp3 = item.y >> 8;           // 3 clamp bits per view
    const int src = threadIdx.x;                   // slot column of this thread
    // the running statistic is read now and written at the end (a read-modify-write at the end of the chain stalled
    // every thread for a DRAM round trip: 14 % of the kernel's attributed stall samples)
    const float ga_prev = stat_grad_accum ? stat_grad_accum[idx] : 0.f;
    // asynchronous copies into the thread's own shared-memory slots: group 0 = the live views' 48-byte gradient
    // records, group 1 = the SH row.  They land while Sigma3 / the per-view chains run.
    for (int v = 0; v < V; ++v) {
        if ((vis >> v) & 1u) {
            const float4* gp = reinterpret_cast<const float4*>(tab.v[v].grad2d) + 3 * (size_t)idx;
            float4* slot = s_slots + ((size_t)v * 128 + src) * 3;
            cpa16(slot, gp), cpa16(slot + 1, gp + 1), cpa16(slot + 2, gp + 2);
        }
    }
    cpa_commit();
    const float* sh = DEG >= 0 ? shs + (size_t)idx * M * 3 : nullptr;
    if (DEG >= 0) {   // the SH row goes out first (asynchronous, into the thread's own row), then the parameters
        if (VEC) {
            const float4* s4 = reinterpret_cast<const float4*>(sh);
#pragma unroll
            for (int j = 0; j < 3 * NCH; ++j) cpa16(sh_row + j, s4 + j);
        } else {
            prefetch_l2(sh);
            if (K * 12 > 128) prefetch_l2(reinterpret_cast<const char*>(sh) + 128);
        }
    }
    cpa_commit();
    const float x = means3D[3 * idx], y = means3D[3 * idx + 1], z = means3D[3 * idx + 2];
    const bool has_sr = (scales != nullptr) && (cov3D_precomp == nullptr);
    float sc_in[3] = {0.f, 0.f, 0.f};
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    if (has_sr) {
        sc_in[0] = scales[3 * idx], sc_in[1] = scales[3 * idx + 1], sc_in[2] = scales[3 * idx + 2];
        q = reinterpret_cast<const float4*>(rotations)[idx];
    }
    // Sigma3 (view independent); R and s kept for its backward
    float c0, c1, c2, c3, c4, c5;
    float R[3][3], s[3];
    if (!has_sr) {
        const float* cs = cov3D_precomp + 6 * (size_t)idx;
        c0 = cs[0], c1 = cs[1], c2 = cs[2], c3 = cs[3], c4 = cs[4], c5 = cs[5];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            s[i] = 0.f;
#pragma unroll
            for (int j = 0; j < 3; ++j) R[i][j] = 0.f;
        }
    } else {
        const float mod = tab.scale_modifier;
        s[0] = mod * sc_in[0], s[1] = mod * sc_in[1], s[2] = mod * sc_in[2];
        const float r = q.x, qx = q.y, qy = q.z, qz = q.w;
        R[0][0] = 1.f - 2.f * (qy * qy + qz * qz), R[0][1] = 2.f * (qx * qy - r * qz), R[0][2] = 2.f * (qx * qz + r * qy);
        R[1][0] = 2.f * (qx * qy + r * qz), R[1][1] = 1.f - 2.f * (qx * qx + qz * qz), R[1][2] = 2.f * (qy * qz - r * qx);
        R[2][0] = 2.f * (qx * qz - r * qy), R[2][1] = 2.f * (qy * qz + r * qx), R[2][2] = 1.f - 2.f * (qx * qx + qy * qy);
        float L[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) L[i][j] = R[i][j] * s[j];
        // same association as the forward kernel (preprocess.cu); FMA contraction differs by <= 1 ulp
        c0 = L[0][0] * L[0][0] + L[0][1] * L[0][1] + L[0][2] * L[0][2];
        c1 = L[0][0] * L[1][0] + L[0][1] * L[1][1] + L[0][2] * L[1][2];
        c2 = L[0][0] * L[2][0] + L[0][1] * L[2][1] + L[0][2] * L[2][2];
        c3 = L[1][0] * L[1][0] + L[1][1] * L[1][1] + L[1][2] * L[1][2];
        c4 = L[1][0] * L[2][0] + L[1][1] * L[2][1] + L[1][2] * L[2][2];
        c5 = L[2][0] * L[2][0] + L[2][1] * L[2][1] + L[2][2] * L[2][2];
    }

    float dmx = 0.f, dmy = 0.f, dmz = 0.f, dop = 0.f;
    float dS[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float dcol[3] = {0.f, 0.f, 0.f};
    float st_norm = 0.f;

    cpa_wait<1>();   // the gradient records have landed (each thread reads only what it copied itself)
    if (tab.clean_scratch) {   // self-cleaning scratch: the record is consumed, leave zeros for the next backward
        for (int v = 0; v < V; ++v) {   // (the stores are issued after the loads have completed: DESIGN.md 4b)
            if ((vis >> v) & 1u) {
                float4* gp = reinterpret_cast<float4*>(tab.v[v].grad2d) + 3 * (size_t)idx;
                gp[0] = gp[1] = gp[2] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
    for (int v = 0; v < V; ++v) {
        const ViewTab& vt = tab.v[v];
        if (!((vis >> v) & 1u)) continue;   // (its dL/dmeans2D was zeroed by the scan kernel)
        const float* mV = sV[v];
        const float* mP = sP[v];
        float4* slot = s_slots + ((size_t)v * 128 + src) * 3;
        const float4 g0 = slot[0], g1 = slot[1], g2 = slot[2];
        // render backward leaves the two diagonal conic gradients without their factor 1/2 (render.cu K7)
        const float g_px = g0.x, g_py = g0.y, g_ca = 0.5f * g0.z, g_cb = g0.w, g_cc = 0.5f * g1.x, g_op = g1.y;
        float g_rgb[3] = {g1.z, g1.w, g2.x};
        const float g_depth = g2.y;

        // ---- conic -> cov2D -> (Sigma3, t) ------------------------------------------------------
        const float tvx = mV[0] * x + mV[4] * y + mV[8] * z + mV[12];
        const float tvy = mV[1] * x + mV[5] * y + mV[9] * z + mV[13];
        const float tvz = mV[2] * x + mV[6] * y + mV[10] * z + mV[14];
        const float txtz = tvx / tvz, tytz = tvy / tvz;
        const float focal_x = sS[v][0], focal_y = sS[v][1], limx = sS[v][2], limy = sS[v][3];
        const float tx = fminf(limx, fmaxf(-limx, txtz)) * tvz;
        const float ty = fminf(limy, fmaxf(-limy, tytz)) * tvz;
        const float xmul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
        const float ymul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
        const float itz = 1.0f / tvz, itz2 = itz * itz, itz3 = itz2 * itz;
        const float J00 = focal_x * itz, J02 = -focal_x * tx * itz2;
        const float J11 = focal_y * itz, J12 = -focal_y * ty * itz2;
        const float M0[3] = {J00 * mV[0] + J02 * mV[2], J00 * mV[4] + J02 * mV[6], J00 * mV[8] + J02 * mV[10]};
        const float M1[3] = {J11 * mV[1] + J12 * mV[2], J11 * mV[5] + J12 * mV[6], J11 * mV[9] + J12 * mV[10]};
        const float S0[3] = {c0 * M0[0] + c1 * M0[1] + c2 * M0[2], c1 * M0[0] + c3 * M0[1] + c4 * M0[2],
                             c2 * M0[0] + c4 * M0[1] + c5 * M0[2]};  // Sigma M0
        const float S1[3] = {c0 * M1[0] + c1 * M1[1] + c2 * M1[2], c1 * M1[0] + c3 * M1[1] + c4 * M1[2],
                             c2 * M1[0] + c4 * M1[1] + c5 * M1[2]};  // Sigma M1
        const float a = M0[0] * S0[0] + M0[1] * S0[1] + M0[2] * S0[2] + DILATION;
        const float b = M0[0] * S1[0] + M0[1] * S1[1] + M0[2] * S1[2];
        const float c = M1[0] * S1[0] + M1[1] * S1[1] + M1[2] * S1[2] + DILATION;
        const float det = a * c - b * b;
        const float d2i = 1.0f / (det * det + 0.0000001f);
        const float dLa = d2i * (-c * c * g_ca + b * c * g_cb + (det - a * c) * g_cc);
        const float dLc = d2i * (-a * a * g_cc + a * b * g_cb + (det - a * c) * g_ca);
        const float dLb = d2i * (2.f * b * c * g_ca - (det + 2.f * b * b) * g_cb + 2.f * a * b * g_cc);
        dS[0] += M0[0] * M0[0] * dLa + M0[0] * M1[0] * dLb + M1[0] * M1[0] * dLc;
        dS[3] += M0[1] * M0[1] * dLa + M0[1] * M1[1] * dLb + M1[1] * M1[1] * dLc;
        dS[5] += M0[2] * M0[2] * dLa + M0[2] * M1[2] * dLb + M1[2] * M1[2] * dLc;
        dS[1] += 2.f * M0[0] * M0[1] * dLa + (M0[0] * M1[1] + M0[1] * M1[0]) * dLb + 2.f * M1[0] * M1[1] * dLc;
        dS[2] += 2.f * M0[0] * M0[2] * dLa + (M0[0] * M1[2] + M0[2] * M1[0]) * dLb + 2.f * M1[0] * M1[2] * dLc;
        dS[4] += 2.f * M0[1] * M0[2] * dLa + (M0[1] * M1[2] + M0[2] * M1[1]) * dLb + 2.f * M1[1] * M1[2] * dLc;
        {
            float dM0[3], dM1[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                dM0[k] = 2.f * dLa * S0[k] + dLb * S1[k];
                dM1[k] = 2.f * dLc * S1[k] + dLb * S0[k];
            }
            const float dJ00 = dM0[0] * mV[0] + dM0[1] * mV[4] + dM0[2] * mV[8];
            const float dJ02 = dM0[0] * mV[2] + dM0[1] * mV[6] + dM0[2] * mV[10];
            const float dJ11 = dM1[0] * mV[1] + dM1[1] * mV[5] + dM1[2] * mV[9];
            const float dJ12 = dM1[0] * mV[2] + dM1[1] * mV[6] + dM1[2] * mV[10];
            const float dtx = xmul * -focal_x * itz2 * dJ02;
            const float dty = ymul * -focal_y * itz2 * dJ12;
            const float dtz = -focal_x * itz2 * dJ00 - focal_y * itz2 * dJ11 +
                              2.f * focal_x * tx * itz3 * dJ02 + 2.f * focal_y * ty * itz3 * dJ12;
            dmx += mV[0] * dtx + mV[1] * dty + mV[2] * dtz;
            dmy += mV[4] * dtx + mV[5] * dty + mV[6] * dtz;
            dmz += mV[8] * dtx + mV[9] * dty + mV[10] * dtz;
        }
        // ---- mean2D (NDC) and depth -------------------------------------------------------------
        const float gnx = g_px * 0.5f * (float)tab.W, gny = g_py * 0.5f * (float)tab.H;
        // with extra channels the record's spare floats carry the mean gradient WITHOUT the extra channels' terms: that
        // is what the reference's means2D sees (its second rasterizer call gets a gradient-free means2D)
        const float gnx2 = tab.n_extra > 0 ? g2.z * 0.5f * (float)tab.W : gnx;
        const float gny2 = tab.n_extra > 0 ? g2.w * 0.5f * (float)tab.H : gny;
        {
            const float hx = mP[0] * x + mP[4] * y + mP[8] * z + mP[12];
            const float hy = mP[1] * x + mP[5] * y + mP[9] * z + mP[13];
            const float hw = mP[3] * x + mP[7] * y + mP[11] * z + mP[15];
            const float pw = 1.0f / (hw + PW_EPS);
            const float mul1 = hx * pw * pw, mul2 = hy * pw * pw;
            dmx += (mP[0] * pw - mP[3] * mul1) * gnx + (mP[1] * pw - mP[3] * mul2) * gny;
            dmy += (mP[4] * pw - mP[7] * mul1) * gnx + (mP[5] * pw - mP[7] * mul2) * gny;
            dmz += (mP[8] * pw - mP[11] * mul1) * gnx + (mP[9] * pw - mP[11] * mul2) * gny;
            dmx += mV[2] * g_depth;
            dmy += mV[6] * g_depth;
            dmz += mV[10] * g_depth;
        }
        if (vt.dL_dmeans2D) {
            put<ACC>(vt.dL_dmeans2D + 3 * idx, gnx2);
            put<ACC>(vt.dL_dmeans2D + 3 * idx + 1, gny2);
            if (!ACC) vt.dL_dmeans2D[3 * idx + 2] = 0.f;
        }
        // densification statistics of this view (geometry/gaussian_base.py:815-819)
        st_norm += sqrtf(gnx2 * gnx2 + gny2 * gny2);
        dop += g_op;
        // ---- colour -----------------------------------------------------------------------------
        if (DEG >= 0) {
            const uint32_t bits = (clamp3 >> (3 * v)) & 7u;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch)
                if (bits & (1u << ch)) g_rgb[ch] = 0.f;
            float dx = x - sC[v][0], dy = y - sC[v][1], dz = z - sC[v][2];
            const float n = sqrtf(dx * dx + dy * dy + dz * dz);
            const float in = 1.0f / n;
            slot[0] = make_float4(g_rgb[0], g_rgb[1], g_rgb[2], dx * in);
            slot[1] = make_float4(dy * in, dz * in, in, 0.f);
            slot[2] = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            dcol[0] += g_rgb[0], dcol[1] += g_rgb[1], dcol[2] += g_rgb[2];
        }
    }

    // ---- SH: chunk-outer / view-inner (each 48-byte chunk of the SH block is loaded once) ---------------
    if (DEG >= 0) {
        float* dst = dL_dshs + (size_t)idx * M * 3;
        cpa_wait<0>();   // the SH row has landed
        sh_chunk_all_views<0, K, ACC, VEC>(sh, sh_row, dst, M, V, vis, s_slots, src);
        sh_chunk_all_views<1, K, ACC, VEC>(sh, sh_row, dst, M, V, vis, s_slots, src);
        sh_chunk_all_views<2, K, ACC, VEC>(sh, sh_row, dst, M, V, vis, s_slots, src);
        sh_chunk_all_views<3, K, ACC, VEC>(sh, sh_row, dst, M, V, vis, s_slots, src);
        if (!ACC) {   // coefficients above the active degree get zero gradients
            if (VEC) {
                float4* d4 = reinterpret_cast<float4*>(dst);
                for (int i = 3 * NCH; i < 3 * (M >> 2); ++i) d4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                // dead coefficients inside the last live chunk were written as zeros by the chunk store
            } else {
                for (int k = 3 * K; k < 3 * M; ++k) dst[k] = 0.f;
            }
        }
        // through dir = d / |d|, per view
        for (int v = 0; v < V; ++v) {
            if (!((vis >> v) & 1u)) continue;
            const float4* r = s_slots + ((size_t)v * 128 + src) * 3;
            const float4 ra = r[0], rb = r[1], rc = r[2];
            const float dx = ra.w, dy = rb.x, dz = rb.y, in = rb.z;
            const float ddx = rc.x, ddy = rc.y, ddz = rc.z;
            const float dot = dx * ddx + dy * ddy + dz * ddz;
            dmx += (ddx - dx * dot) * in;
            dmy += (ddy - dy * dot) * in;
            dmz += (ddz - dz * dot) * in;
        }
    } else if (dL_dcolors) {
        put<ACC>(dL_dcolors + 3 * idx, dcol[0]);
        put<ACC>(dL_dcolors + 3 * idx + 1, dcol[1]);
        put<ACC>(dL_dcolors + 3 * idx + 2, dcol[2]);
    }
    // ---- Sigma3 -> scale / rotation (once, on the dL/dSigma summed over the views) ------------------
    if (has_sr && dL_dscales != nullptr) {
        const float mod = tab.scale_modifier;
        const float r = q.x, qx = q.y, qy = q.z, qz = q.w;
        // G = dL/dSigma as a full symmetric matrix; dL/dL = 2 G L, L = R diag(s)
        const float Gm[3][3] = {{dS[0], 0.5f * dS[1], 0.5f * dS[2]},
                                {0.5f * dS[1], dS[3], 0.5f * dS[4]},
                                {0.5f * dS[2], 0.5f * dS[4], dS[5]}};
        float dLm[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)
                dLm[i][j] = 2.f * (Gm[i][0] * R[0][j] + Gm[i][1] * R[1][j] + Gm[i][2] * R[2][j]) * s[j];
        float ds[3], dR[3][3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            ds[j] = (dLm[0][j] * R[0][j] + dLm[1][j] * R[1][j] + dLm[2][j] * R[2][j]) * mod;
#pragma unroll
            for (int i = 0; i < 3; ++i) dR[i][j] = dLm[i][j] * s[j];
        }
        put<ACC>(dL_dscales + 3 * idx, ds[0]);
        put<ACC>(dL_dscales + 3 * idx + 1, ds[1]);
        put<ACC>(dL_dscales + 3 * idx + 2, ds[2]);
        const float dr = 2.f * (-qz * dR[0][1] + qy * dR[0][2] + qz * dR[1][0] - qx * dR[1][2] - qy * dR[2][0] + qx * dR[2][1]);
        const float dqx = 2.f * (qy * dR[0][1] + qz * dR[0][2] + qy * dR[1][0] - 2.f * qx * dR[1][1] - r * dR[1][2] +
                                 qz * dR[2][0] + r * dR[2][1] - 2.f * qx * dR[2][2]);
        const float dqy = 2.f * (-2.f * qy * dR[0][0] + qx * dR[0][1] + r * dR[0][2] + qx * dR[1][0] + qz * dR[1][2] -
                                 r * dR[2][0] + qz * dR[2][1] - 2.f * qy * dR[2][2]);
        const float dqz = 2.f * (-2.f * qz * dR[0][0] - r * dR[0][1] + qx * dR[0][2] + r * dR[1][0] -
                                 2.f * qz * dR[1][1] + qy * dR[1][2] + qx * dR[2][0] + qy * dR[2][1]);
        float4* o = reinterpret_cast<float4*>(dL_drotations) + idx;
        if (ACC) {
            float4 p = *o;
            *o = make_float4(p.x + dr, p.y + dqx, p.z + dqy, p.w + dqz);
        } else {
            *o = make_float4(dr, dqx, dqy, dqz);
        }
    } else if (dL_dcov3D != nullptr) {
#pragma unroll
        for (int k = 0; k < 6; ++k) put<ACC>(dL_dcov3D + (size_t)idx * 6 + k, dS[k]);
    }
    put<ACC>(dL_dmeans3D + 3 * idx, dmx);
    put<ACC>(dL_dmeans3D + 3 * idx + 1, dmy);
    put<ACC>(dL_dmeans3D + 3 * idx + 2, dmz);
    put<ACC>(dL_dopacity + idx, dop);
    // fused densification statistics (geometry/gaussian_base.py:815-819, 846-851)
    // (denominator and largest radius: the scan kernel; a view without a gradient adds a zero norm)
    if (stat_grad_accum) stat_grad_accum[idx] = ga_prev + st_norm;
    }   // work-list loop
}

cudaError_t launch_preprocess_backward(const BatchTab& tab, const float* means3D, const float* scales,
                                       const float* rotations, const float* shs, const float* cov3D_precomp,
                                       float* dL_dmeans3D, float* dL_dshs, float* dL_dcolors, float* dL_dopacity,
                                       float* dL_dscales, float* dL_drotations, float* dL_dcov3D,
                                       float* stat_grad_accum, float* stat_denom, float* stat_max_radii,
                                       int accumulate, cudaStream_t st, int g_begin, int g_end) {
    if (tab.P <= 0) return cudaSuccess;
    if (g_end <= 0 || g_end > tab.P) g_end = tab.P;
    if (g_begin < 0) g_begin = 0;
    if (g_begin >= g_end) return cudaSuccess;
    const int grid = (g_end - g_begin + 127) / 128;
    const bool vec = tab.sh_degree >= 0 && (tab.M & 3) == 0 && tab.M <= 16 &&
                     ((reinterpret_cast<uintptr_t>(shs) | reinterpret_cast<uintptr_t>(dL_dshs)) & 15) == 0;
    // work list of the Gaussians that carry a gradient: view 0's second depth-sort buffer (dead after the forward) and
    // its sort tickets hold the list and its length
    uint2* live_list = reinterpret_cast<uint2*>(tab.v[0].gwords[1]);
    uint32_t* live_count = tab.v[0].gtickets;
    cudaError_t e0 = cudaMemsetAsync(live_count, 0, sizeof(uint32_t), st);
    if (e0 != cudaSuccess) return e0;
    {
        const int sgrid = (g_end - g_begin + 255) / 256;
        // zero rows of dL_dshs are written warp-wide when a row is a whole number of float4s at a 16-byte aligned base
        const int sh_vec = (dL_dshs != nullptr && ((3 * tab.M) & 3) == 0 &&
                            (reinterpret_cast<uintptr_t>(dL_dshs) & 15) == 0) ? 1 : 0;
        if (accumulate)
            preprocess_backward_scan_kernel<true><<<sgrid, 256, 0, st>>>(
                tab, dL_dmeans3D, dL_dshs, dL_dcolors, dL_dopacity, dL_dscales, dL_drotations, dL_dcov3D, stat_denom,
                stat_max_radii, tab.sh_degree >= 0 ? 1 : 0, sh_vec, live_list, live_count, g_begin, g_end);
        else
            preprocess_backward_scan_kernel<false><<<sgrid, 256, 0, st>>>(
                tab, dL_dmeans3D, dL_dshs, dL_dcolors, dL_dopacity, dL_dscales, dL_drotations, dL_dcov3D, stat_denom,
                stat_max_radii, tab.sh_degree >= 0 ? 1 : 0, sh_vec, live_list, live_count, g_begin, g_end);
        count_launch();
        e0 = cudaGetLastError();
        if (e0 != cudaSuccess) return e0;
    }
    const size_t smem = ((size_t)tab.V * 128 * 3 + (vec ? 128 * SH_ROW_F4 : 0)) * sizeof(float4);
#define LAUNCH_PB(D, A, VC)                                                                                        \
    {                                                                                                              \
        auto kfn = preprocess_backward_kernel<D, A, VC>;                                                           \
        static size_t attr = 48 * 1024;                                                                            \
        if (smem > attr) {                                                                                         \
            cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
            if (e != cudaSuccess) return e;                                                                        \
            attr = smem;                                                                                           \
        }                                                                                                          \
        kfn<<<grid < NUM_SMS * 4 ? grid : NUM_SMS * 4, 128, smem, st>>>(                                           \
            tab, means3D, scales, rotations, shs, cov3D_precomp, dL_dmeans3D, dL_dshs,                             \
                                     dL_dcolors, dL_dopacity, dL_dscales, dL_drotations, dL_dcov3D,                \
                                     stat_grad_accum, live_list, live_count);                                      \
    }
#define DISPATCH_DEG(A, VC)                                                                                        \
    switch (tab.sh_degree) {                                                                                       \
        case -1: LAUNCH_PB(-1, A, false); break;                                                                   \
        case 0: LAUNCH_PB(0, A, VC); break;                                                                        \
        case 1: LAUNCH_PB(1, A, VC); break;                                                                        \
        case 2: LAUNCH_PB(2, A, VC); break;                                                                        \
        default: LAUNCH_PB(3, A, VC); break;                                                                       \
    }
    if (accumulate) {
        if (vec) { DISPATCH_DEG(true, true) } else { DISPATCH_DEG(true, false) }
    } else {
        if (vec) { DISPATCH_DEG(false, true) } else { DISPATCH_DEG(false, false) }
    }
#undef DISPATCH_DEG
#undef LAUNCH_PB
    count_launch();
    return cudaGetLastError();
}

}  // namespace b200splat

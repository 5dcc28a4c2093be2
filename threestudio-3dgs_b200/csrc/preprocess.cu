// K1 preprocess, K3 duplicateWithKeys, K9 markVisible.
//
// Replaces upstream forward.cu preprocessCUDA (+computeCov3D/computeCov2D/computeColorFromSH/
// in_frustum/getRect) and rasterizer_impl.cu duplicateWithKeys/checkFrustum [UPSTREAM-RECALL];
// reference call site renderer/diff_gaussian_rasterizer.py:122-131.
//
// THIS TRANSLATION UNIT IS COMPILED WITH -fmad=false: every quantity feeding the bit-exact
// outputs (radii, tile rectangle, tiles_touched, depth bits of the sort key) is computed with
// one IEEE fp32 operation per source operator, in the association order written here, which is
// the order oracle/torch_oracle.py spells out.  Division and sqrt are nvcc's IEEE defaults.
// The kernel is HBM-bound (104-284 B/Gaussian), so the lost FMA contraction is free.
//
// View batching: the Gaussian parameters (44 + 12 M bytes) are read ONCE and projected into every
// view of the batch (the reference renders the views of a step one after the other,
// renderer/gaussian_batch_renderer.py:21-54, re-reading them each time).
#include "common.cuh"

namespace b200splat {

__device__ __forceinline__ float ndc2pix(float v, int S) { return ((v + 1.0f) * (float)S - 1.0f) * 0.5f; }

__device__ __forceinline__ void get_rect(float px, float py, float radius_f, int gx, int gy, int& x0, int& y0,
                                         int& x1, int& y1) {
    x0 = min(gx, max(0, (int)((px - radius_f) / (float)BLOCK_X)));
    y0 = min(gy, max(0, (int)((py - radius_f) / (float)BLOCK_Y)));
    x1 = min(gx, max(0, (int)((px + radius_f + (float)(BLOCK_X - 1)) / (float)BLOCK_X)));
    y1 = min(gy, max(0, (int)((py + radius_f + (float)(BLOCK_Y - 1)) / (float)BLOCK_Y)));
}

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

template <int DEG>
__device__ __forceinline__ float sh_channel(const float* s, float x, float y, float z) {
    // s[k] = coefficient k of this channel
    float res = SH_C0 * s[0];
    if (DEG > 0) {
        res = res - SH_C1 * y * s[1] + SH_C1 * z * s[2] - SH_C1 * x * s[3];
        if (DEG > 1) {
            float xx = x * x, yy = y * y, zz = z * z;
            float xy = x * y, yz = y * z, xz = x * z;
            res = res + SH_C2_0 * xy * s[4] + SH_C2_1 * yz * s[5] + SH_C2_2 * (2.0f * zz - xx - yy) * s[6] +
                  SH_C2_3 * xz * s[7] + SH_C2_4 * (xx - yy) * s[8];
            if (DEG > 2) {
                res = res + SH_C3_0 * y * (3.0f * xx - yy) * s[9] + SH_C3_1 * xy * z * s[10] +
                      SH_C3_2 * y * (4.0f * zz - xx - yy) * s[11] +
                      SH_C3_3 * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * s[12] +
                      SH_C3_4 * x * (4.0f * zz - xx - yy) * s[13] + SH_C3_5 * z * (xx - yy) * s[14] +
                      SH_C3_6 * x * (xx - 3.0f * yy) * s[15];
            }
        }
    }
    return res;
}

// DEG = -1: colours precomputed; 0..3: SH degree
template <int DEG>
__global__ void __launch_bounds__(256)
preprocess_kernel(const __grid_constant__ BatchTab tab, const float* __restrict__ means3D,
                  const float* __restrict__ scales, const float* __restrict__ rotations,
                  const float* __restrict__ opacities, const float* __restrict__ shs,
                  const float* __restrict__ colors_precomp, const float* __restrict__ cov3D_precomp) {
    __shared__ float sV[MAX_VIEWS][16], sP[MAX_VIEWS][16], sC[MAX_VIEWS][4], sS[MAX_VIEWS][4];
    const int V = tab.V;
    for (int i = threadIdx.x; i < V * 16; i += blockDim.x) {
        sV[i >> 4][i & 15] = tab.v[i >> 4].view[i & 15];
        sP[i >> 4][i & 15] = tab.v[i >> 4].proj[i & 15];
    }
    for (int i = threadIdx.x; i < V * 3; i += blockDim.x) sC[i / 3][i % 3] = tab.v[i / 3].campos[i % 3];
    // (focal_x, focal_y, limx, limy): from the host, or from the device when the field of view never left it
    for (int i = threadIdx.x; i < V * 4; i += blockDim.x) {
        const ViewTab& t = tab.v[i >> 2];
        const float host[4] = {t.focal_x, t.focal_y, t.limx, t.limy};
        sS[i >> 2][i & 3] = t.scalars ? t.scalars[i & 3] : host[i & 3];
    }
    __syncthreads();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= tab.P) return;
    constexpr int K = DEG < 0 ? 1 : (DEG + 1) * (DEG + 1);

    const float x = __ldg(means3D + 3 * idx), y = __ldg(means3D + 3 * idx + 1), z = __ldg(means3D + 3 * idx + 2);
    // the SH block (up to 2 lines) goes to L2 while the cheap near-plane tests run
    if (DEG >= 0) {
        const char* shp = reinterpret_cast<const char*>(shs + (size_t)idx * tab.M * 3);
        prefetch_l2(shp);
        if (K * 12 > 128) prefetch_l2(shp + 128);
    }
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    float s3[3] = {0.f, 0.f, 0.f};
    if (cov3D_precomp == nullptr) {
        q = __ldg(reinterpret_cast<const float4*>(rotations) + idx);
        s3[0] = __ldg(scales + 3 * idx), s3[1] = __ldg(scales + 3 * idx + 1), s3[2] = __ldg(scales + 3 * idx + 2);
    }
    const float opac = __ldg(opacities + idx);

    uint32_t front = 0;
    for (int v = 0; v < V; ++v) {
        const float tvz = sV[v][2] * x + sV[v][6] * y + sV[v][10] * z + sV[v][14];
        if (tvz > NEAR_CULL) front |= 1u << v;
    }
    if (front == 0) {
        for (int v = 0; v < V; ++v) {
            tab.v[v].radii[idx] = 0;
            tab.v[v].tiles_touched[idx] = 0;
            tab.v[v].rect[idx] = make_ushort4(0, 0, 0, 0);
            tab.v[v].gwords[0][idx] = 0xFFFFFFFF00000000ull | (uint64_t)(uint32_t)idx;   // culled: sorts last
        }
        return;
    }
    // Sigma3 (view independent)
    float c0, c1, c2, c3, c4, c5;
    if (cov3D_precomp != nullptr) {
        const float* c = cov3D_precomp + 6 * (size_t)idx;
        c0 = __ldg(c), c1 = __ldg(c + 1), c2 = __ldg(c + 2), c3 = __ldg(c + 3), c4 = __ldg(c + 4), c5 = __ldg(c + 5);
    } else {
        const float mod = tab.scale_modifier;
        const float sx = mod * s3[0], sy = mod * s3[1], sz = mod * s3[2];
        const float r = q.x, qx = q.y, qy = q.z, qz = q.w;
        const float R00 = 1.0f - 2.0f * (qy * qy + qz * qz);
        const float R01 = 2.0f * (qx * qy - r * qz);
        const float R02 = 2.0f * (qx * qz + r * qy);
        const float R10 = 2.0f * (qx * qy + r * qz);
        const float R11 = 1.0f - 2.0f * (qx * qx + qz * qz);
        const float R12 = 2.0f * (qy * qz - r * qx);
        const float R20 = 2.0f * (qx * qz - r * qy);
        const float R21 = 2.0f * (qy * qz + r * qx);
        const float R22 = 1.0f - 2.0f * (qx * qx + qy * qy);
        const float L00 = R00 * sx, L01 = R01 * sy, L02 = R02 * sz;
        const float L10 = R10 * sx, L11 = R11 * sy, L12 = R12 * sz;
        const float L20 = R20 * sx, L21 = R21 * sy, L22 = R22 * sz;
        c0 = L00 * L00 + L01 * L01 + L02 * L02;
        c1 = L00 * L10 + L01 * L11 + L02 * L12;
        c2 = L00 * L20 + L01 * L21 + L02 * L22;
        c3 = L10 * L10 + L11 * L11 + L12 * L12;
        c4 = L10 * L20 + L11 * L21 + L12 * L22;
        c5 = L20 * L20 + L21 * L21 + L22 * L22;
    }
    // colour inputs (view independent): SH coefficients per channel, or the precomputed colour
    float coef[3][K];
    if (DEG < 0) {
        coef[0][0] = __ldg(colors_precomp + 3 * idx);
        coef[1][0] = __ldg(colors_precomp + 3 * idx + 1);
        coef[2][0] = __ldg(colors_precomp + 3 * idx + 2);
    } else {
        const float* sh = shs + (size_t)idx * tab.M * 3;
        if (K == 16) {
            const float4* v4 = reinterpret_cast<const float4*>(sh);
#pragma unroll
            for (int i = 0; i < 12; ++i) {
                const float4 t = __ldg(v4 + i);
                const float e[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int f = i * 4 + j;
                    coef[f % 3][f / 3] = e[j];
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < K; ++k) {
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) coef[ch][k] = __ldg(sh + k * 3 + ch);
            }
        }
    }
    // cull threshold of the render kernels: alpha >= 1/255 needs A dx^2 + 2B dx dy + C dy^2 <= 2 ln(255 o)
    const float thr = opac > 0.0f ? 2.0f * __logf(255.0f * opac) + 0.002f : -1.0f;

    for (int v = 0; v < V; ++v) {
        const ViewTab& vt = tab.v[v];
        int my_radius = 0;
        uint32_t my_tiles = 0;
        ushort4 my_rect = make_ushort4(0, 0, 0, 0);
        float my_depth = 0.f;
        if ((front >> v) & 1u) {
            const float* mV = sV[v];
            const float* mP = sP[v];
            // view / projection (transformPoint4x3 / 4x4 on the transposed matrices)
            const float tvx = mV[0] * x + mV[4] * y + mV[8] * z + mV[12];
            const float tvy = mV[1] * x + mV[5] * y + mV[9] * z + mV[13];
            const float tvz = mV[2] * x + mV[6] * y + mV[10] * z + mV[14];
            const float hx = mP[0] * x + mP[4] * y + mP[8] * z + mP[12];
            const float hy = mP[1] * x + mP[5] * y + mP[9] * z + mP[13];
            const float hw = mP[3] * x + mP[7] * y + mP[11] * z + mP[15];
            const float pw = 1.0f / (hw + PW_EPS);
            const float ndcx = hx * pw, ndcy = hy * pw;
            // EWA projection
            const float txtz = tvx / tvz, tytz = tvy / tvz;
            const float focal_x = sS[v][0], focal_y = sS[v][1], limx = sS[v][2], limy = sS[v][3];
            const float tx = fminf(limx, fmaxf(-limx, txtz)) * tvz;
            const float ty = fminf(limy, fmaxf(-limy, tytz)) * tvz;
            const float J00 = focal_x / tvz;
            const float J02 = -(focal_x * tx) / (tvz * tvz);
            const float J11 = focal_y / tvz;
            const float J12 = -(focal_y * ty) / (tvz * tvz);
            const float M00 = J00 * mV[0] + J02 * mV[2];
            const float M01 = J00 * mV[4] + J02 * mV[6];
            const float M02 = J00 * mV[8] + J02 * mV[10];
            const float M10 = J11 * mV[1] + J12 * mV[2];
            const float M11 = J11 * mV[5] + J12 * mV[6];
            const float M12 = J11 * mV[9] + J12 * mV[10];
            const float N00 = M00 * c0 + M01 * c1 + M02 * c2;
            const float N01 = M00 * c1 + M01 * c3 + M02 * c4;
            const float N02 = M00 * c2 + M01 * c4 + M02 * c5;
            const float N10 = M10 * c0 + M11 * c1 + M12 * c2;
            const float N11 = M10 * c1 + M11 * c3 + M12 * c4;
            const float N12 = M10 * c2 + M11 * c4 + M12 * c5;
            const float a = N00 * M00 + N01 * M01 + N02 * M02 + DILATION;
            const float b = N00 * M10 + N01 * M11 + N02 * M12;
            const float c = N10 * M10 + N11 * M11 + N12 * M12 + DILATION;
            const float det = a * c - b * b;
            if (det != 0.0f) {
                const float det_inv = 1.0f / det;
                const float mid = 0.5f * (a + c);
                const float disc = sqrtf(fmaxf(mid * mid - det, LAMBDA_FLOOR));
                const float lam = fmaxf(mid + disc, mid - disc);
                const float radius_f = ceilf(3.0f * sqrtf(lam));
                const float px = ndc2pix(ndcx, tab.W), py = ndc2pix(ndcy, tab.H);
                int x0, y0, x1, y1;
                get_rect(px, py, radius_f, tab.grid_x, tab.grid_y, x0, y0, x1, y1);
                const int area = (x1 - x0) * (y1 - y0);
                if (area > 0) {
                    float rgb[3];
                    uint8_t bits = 0;
                    if (DEG < 0) {
                        rgb[0] = coef[0][0], rgb[1] = coef[1][0], rgb[2] = coef[2][0];
                    } else {
                        float dx = x - sC[v][0], dy = y - sC[v][1], dz = z - sC[v][2];
                        // the view direction feeds the colour only (1e-4 bar), nothing bit-exact: MUFU.RSQ and three
                        // multiplies instead of an IEEE square root and three IEEE divisions (10 % of the kernel's
                        // instructions)
                        const float in = rsqrtf(dx * dx + dy * dy + dz * dz);
                        dx = dx * in, dy = dy * in, dz = dz * in;
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch) {
                            const float rr = sh_channel<(DEG < 0 ? 0 : DEG)>(coef[ch], dx, dy, dz) + 0.5f;
                            if (rr < 0.0f) bits |= (uint8_t)(1u << ch);
                            rgb[ch] = fmaxf(rr, 0.0f);
                        }
                    }
                    my_radius = (int)radius_f;
                    my_tiles = (uint32_t)area;
                    my_rect = make_ushort4((unsigned short)x0, (unsigned short)y0, (unsigned short)x1, (unsigned short)y1);
                    float4* o = reinterpret_cast<float4*>(vt.rec) + 3 * (size_t)idx;
                    o[0] = make_float4(px, py, c * det_inv, -b * det_inv);
                    o[1] = make_float4(a * det_inv, opac, tvz, rgb[0]);
                    o[2] = make_float4(rgb[1], rgb[2], thr, 0.0f);
                    vt.depths[idx] = tvz;
                    my_depth = tvz;
                    vt.clamped[idx] = bits;
                }
            }
        }
        vt.radii[idx] = my_radius;
        vt.tiles_touched[idx] = my_tiles;
        vt.rect[idx] = my_rect;
        // word of the per-view depth sort of the Gaussians (culled ones sort last and emit nothing)
        vt.gwords[0][idx] = ((uint64_t)(my_radius > 0 ? __float_as_uint(my_depth) : 0xFFFFFFFFu) << 32) |
                            (uint64_t)(uint32_t)idx;
    }
}

// duplicateWithKeys over the Gaussians IN DEPTH ORDER: position i of the depth-sorted list emits one word
// (tile_id << 32 | gaussian) per tile of its rectangle (tiles y-major then x) at point_offsets[i-1].  The words
// of one tile are therefore already in (depth, index) order; a STABLE sort on the tile bits alone (2 passes for
// 1024 tiles instead of 6 passes over 43 key bits) finishes the job and yields exactly the order upstream's sort
// of (tile << 32 | depth) keys gives.  Fused: per-tile pair counts (-> tile ranges by an exclusive scan, and the
// digit histograms of the tile-bit sort).  blockIdx.y = view.  Pairs beyond the binning capacity are dropped and
// flagged (STATUS_OVERFLOW).
constexpr int DUP_MAX_TILES = 8192;   // tile histogram kept in shared memory up to this many tiles
__global__ void __launch_bounds__(256)
duplicate_kernel(const __grid_constant__ BatchTab tab) {
    extern __shared__ uint32_t s_dyn[];   // [passes*256] digit histograms | [T] tile histogram (if it fits)
    const ViewTab& vt = tab.v[blockIdx.y];
    const int end_bit = tab.end_bit;
    const int passes = tab.digit_passes;   // 0: the sort bins by the per-tile counts directly
    const int gx = tab.grid_x, gy = tab.grid_y;
    const int T = gx * gy;
    const bool tile_hist = T <= DUP_MAX_TILES;
    uint32_t* s_hist = s_dyn;
    uint32_t* s_tile = s_dyn + passes * 256;
    for (int i = threadIdx.x; i < passes * 256 + (tile_hist ? T : 0); i += blockDim.x) s_dyn[i] = 0;
    __syncthreads();
    const ushort4* __restrict__ rect = vt.rect;
    const uint32_t* __restrict__ point_offsets = vt.point_offsets;
    const uint64_t* __restrict__ order = vt.gwords[0];
    uint64_t* __restrict__ words = vt.keys[0];
    // 4 Gaussians per thread and iteration: the order words / offsets of all four, then the four dependent rectangle
    // gathers, are in flight together (the kernel was bound by the order -> rect load chain)
    constexpr int DUP_ILP = 4;
    const int stride = gridDim.x * blockDim.x;
    for (int i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < tab.P; i0 += DUP_ILP * stride) {
        uint32_t g[DUP_ILP], off[DUP_ILP];
        ushort4 rc[DUP_ILP];
#pragma unroll
        for (int u = 0; u < DUP_ILP; ++u) {
            const int i = i0 + u * stride;
            const bool live = i < tab.P;
            g[u] = live ? (uint32_t)__ldg(order + i) : 0u;
            off[u] = (live && i > 0) ? __ldg(point_offsets + i - 1) : 0u;
        }
#pragma unroll
        for (int u = 0; u < DUP_ILP; ++u)
            rc[u] = (i0 + u * stride < tab.P) ? __ldg(rect + g[u]) : make_ushort4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < DUP_ILP; ++u) {
            const int x0 = rc[u].x, y0 = rc[u].y, x1 = rc[u].z, y1 = rc[u].w;
            const uint32_t ntiles = (uint32_t)((x1 - x0) * (y1 - y0));
            if (ntiles == 0) continue;
            uint32_t o = off[u];
            if (o + ntiles > tab.capacity) {
                atomicOr(vt.status + STATUS_OVERFLOW, 1u);
                continue;
            }
            for (int ty = y0; ty < y1; ++ty) {
                for (int tx = x0; tx < x1; ++tx) {
                    const uint32_t tile = (uint32_t)(ty * gx + tx);
                    words[o] = ((uint64_t)tile << 32) | (uint64_t)g[u];
                    ++o;
                    if (tile_hist) {
                        atomicAdd(&s_tile[tile], 1u);   // digit histograms of the tile bytes are derived at flush time
                    } else {
                        for (int p = 0; p < passes; ++p) {
                            const int bits = min(8, end_bit - 8 * p);
                            atomicAdd(&s_hist[p * 256 + ((tile >> (8 * p)) & ((1u << bits) - 1u))], 1u);
                        }
                    }
                }
            }
        }
    }
    __syncthreads();
    if (tile_hist) {
        for (int t = threadIdx.x; t < T; t += blockDim.x) {
            const uint32_t c = s_tile[t];
            if (c) {
                atomicAdd(&vt.tile_count[t], c);
                for (int p = 0; p < passes; ++p) {
                    const int bits = min(8, end_bit - 8 * p);
                    atomicAdd(&s_hist[p * 256 + (((uint32_t)t >> (8 * p)) & ((1u << bits) - 1u))], c);
                }
            }
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < passes * 256; i += blockDim.x) {
        const uint32_t c = s_hist[i];
        if (c) atomicAdd(&vt.hist[i], c);
    }
}

__global__ void mark_visible_kernel(int P, const float* __restrict__ means3D, const float* __restrict__ view,
                                    uint8_t* __restrict__ present) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const float x = means3D[3 * idx], y = means3D[3 * idx + 1], z = means3D[3 * idx + 2];
    const float tvz = view[2] * x + view[6] * y + view[10] * z + view[14];
    present[idx] = tvz > NEAR_CULL ? 1 : 0;
}

cudaError_t launch_preprocess(const BatchTab& tab, const float* means3D, const float* scales, const float* rotations,
                              const float* opacities, const float* shs, const float* colors_precomp,
                              const float* cov3D_precomp, cudaStream_t st) {
    if (tab.P <= 0) return cudaSuccess;
    const int grid = (tab.P + 255) / 256;
#define LAUNCH_PRE(D)                                                                                             \
    preprocess_kernel<D><<<grid, 256, 0, st>>>(tab, means3D, scales, rotations, opacities, shs, colors_precomp, \
                                                cov3D_precomp)
    switch (tab.sh_degree) {
        case -1: LAUNCH_PRE(-1); break;
        case 0: LAUNCH_PRE(0); break;
        case 1: LAUNCH_PRE(1); break;
        case 2: LAUNCH_PRE(2); break;
        default: LAUNCH_PRE(3); break;
    }
#undef LAUNCH_PRE
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_duplicate(const BatchTab& tab, cudaStream_t st) {
    if (tab.P <= 0) return cudaSuccess;
    const int T = tab.grid_x * tab.grid_y;
    const int passes = tab.digit_passes;
    const size_t smem = ((size_t)passes * 256 + (T <= DUP_MAX_TILES ? T : 0)) * sizeof(uint32_t);
    static size_t attr = 0;
    if (smem > 48 * 1024 && smem > attr) {
        cudaError_t e = cudaFuncSetAttribute(duplicate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr = smem;
    }
    // few, fat CTAs: every CTA flushes its histograms with global atomics
    const int bx = min((tab.P + 255) / 256, max(1, NUM_SMS * 8 / tab.V));
    duplicate_kernel<<<dim3(bx, tab.V), 256, smem, st>>>(tab);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_mark_visible(int P, const float* means3D, const float* view, const float* proj,
                                uint8_t* present, cudaStream_t st) {
    (void)proj;
    if (P <= 0) return cudaSuccess;
    mark_visible_kernel<<<(P + 255) / 256, 256, 0, st>>>(P, means3D, view, present);
    count_launch();
    return cudaGetLastError();
}

}  // namespace b200splat

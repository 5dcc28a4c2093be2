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

template <int DEG>
__device__ __forceinline__ void sh_to_rgb(const float* __restrict__ sh, float dx, float dy, float dz, float* rgb,
                                          uint8_t* clamped_bits) {
    // sh: (M,3) interleaved for this Gaussian; load the (DEG+1)^2 coefficients of each channel
    constexpr int K = (DEG + 1) * (DEG + 1);
    float c[3][K];
    if (K == 16) {
        const float4* v = reinterpret_cast<const float4*>(sh);
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            float4 q = __ldg(v + i);
            float e[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int f = i * 4 + j;
                c[f % 3][f / 3] = e[j];
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < K; ++k) {
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) c[ch][k] = __ldg(sh + k * 3 + ch);
        }
    }
    uint8_t bits = 0;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        float r = sh_channel<DEG>(c[ch], dx, dy, dz) + 0.5f;
        if (r < 0.0f) bits |= (uint8_t)(1u << ch);
        rgb[ch] = fmaxf(r, 0.0f);
    }
    *clamped_bits = bits;
}

__global__ void __launch_bounds__(256)
preprocess_kernel(int P, CameraParams cam, const float* __restrict__ means3D, const float* __restrict__ scales,
                  const float* __restrict__ rotations, const float* __restrict__ opacities,
                  const float* __restrict__ shs, const float* __restrict__ colors_precomp,
                  const float* __restrict__ cov3D_precomp, int32_t* __restrict__ radii, float* __restrict__ rec,
                  float* __restrict__ depths, float* __restrict__ cov3D_out, uint8_t* __restrict__ clamped,
                  uint32_t* __restrict__ tiles_touched) {
    __shared__ float sV[16], sP[16], sC[3];
    if (threadIdx.x < 16) {
        sV[threadIdx.x] = cam.view[threadIdx.x];
        sP[threadIdx.x] = cam.proj[threadIdx.x];
    }
    if (threadIdx.x < 3) sC[threadIdx.x] = cam.campos[threadIdx.x];
    __syncthreads();
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;

    // start the long-latency, visibility-dependent loads now: the SH block (2 lines) goes to L2 while the
    // projection runs, instead of a third serialized DRAM round trip after the cull tests
    if (shs != nullptr) {
        const char* shp = reinterpret_cast<const char*>(shs + (size_t)idx * cam.M * 3);
        prefetch_l2(shp);
        if (cam.M * 12 > 128) prefetch_l2(shp + 128);
    }
    float4 q_early = make_float4(0.f, 0.f, 0.f, 0.f);
    float s_early[3] = {0.f, 0.f, 0.f};
    if (cov3D_precomp == nullptr) {
        q_early = __ldg(reinterpret_cast<const float4*>(rotations) + idx);
        s_early[0] = __ldg(scales + 3 * idx), s_early[1] = __ldg(scales + 3 * idx + 1), s_early[2] = __ldg(scales + 3 * idx + 2);
    }
    const float opac = __ldg(opacities + idx);
    int my_radius = 0;
    uint32_t my_tiles = 0;
    const float x = __ldg(means3D + 3 * idx), y = __ldg(means3D + 3 * idx + 1), z = __ldg(means3D + 3 * idx + 2);
    // view / projection (transformPoint4x3 / 4x4 on the transposed matrices)
    const float tvx = sV[0] * x + sV[4] * y + sV[8] * z + sV[12];
    const float tvy = sV[1] * x + sV[5] * y + sV[9] * z + sV[13];
    const float tvz = sV[2] * x + sV[6] * y + sV[10] * z + sV[14];
    if (tvz > NEAR_CULL) {
        const float hx = sP[0] * x + sP[4] * y + sP[8] * z + sP[12];
        const float hy = sP[1] * x + sP[5] * y + sP[9] * z + sP[13];
        const float hw = sP[3] * x + sP[7] * y + sP[11] * z + sP[15];
        const float pw = 1.0f / (hw + PW_EPS);
        const float ndcx = hx * pw, ndcy = hy * pw;
        // Sigma3
        float c0, c1, c2, c3, c4, c5;
        if (cov3D_precomp != nullptr) {
            const float* c = cov3D_precomp + 6 * (size_t)idx;
            c0 = __ldg(c), c1 = __ldg(c + 1), c2 = __ldg(c + 2), c3 = __ldg(c + 3), c4 = __ldg(c + 4), c5 = __ldg(c + 5);
        } else {
            const float mod = cam.scale_modifier;
            const float sx = mod * s_early[0], sy = mod * s_early[1], sz = mod * s_early[2];
            const float4 q = q_early;
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
        // EWA projection
        const float txtz = tvx / tvz, tytz = tvy / tvz;
        const float tx = fminf(cam.limx, fmaxf(-cam.limx, txtz)) * tvz;
        const float ty = fminf(cam.limy, fmaxf(-cam.limy, tytz)) * tvz;
        const float J00 = cam.focal_x / tvz;
        const float J02 = -(cam.focal_x * tx) / (tvz * tvz);
        const float J11 = cam.focal_y / tvz;
        const float J12 = -(cam.focal_y * ty) / (tvz * tvz);
        const float M00 = J00 * sV[0] + J02 * sV[2];
        const float M01 = J00 * sV[4] + J02 * sV[6];
        const float M02 = J00 * sV[8] + J02 * sV[10];
        const float M10 = J11 * sV[1] + J12 * sV[2];
        const float M11 = J11 * sV[5] + J12 * sV[6];
        const float M12 = J11 * sV[9] + J12 * sV[10];
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
            const float px = ndc2pix(ndcx, cam.W), py = ndc2pix(ndcy, cam.H);
            int x0, y0, x1, y1;
            get_rect(px, py, radius_f, cam.grid_x, cam.grid_y, x0, y0, x1, y1);
            const int area = (x1 - x0) * (y1 - y0);
            if (area > 0) {
                float rgb[3];
                uint8_t bits = 0;
                if (colors_precomp != nullptr) {
                    rgb[0] = __ldg(colors_precomp + 3 * idx);
                    rgb[1] = __ldg(colors_precomp + 3 * idx + 1);
                    rgb[2] = __ldg(colors_precomp + 3 * idx + 2);
                } else {
                    float dx = x - sC[0], dy = y - sC[1], dz = z - sC[2];
                    const float n = sqrtf(dx * dx + dy * dy + dz * dz);
                    dx = dx / n, dy = dy / n, dz = dz / n;
                    const float* sh = shs + (size_t)idx * cam.M * 3;
                    switch (cam.sh_degree) {
                        case 0: sh_to_rgb<0>(sh, dx, dy, dz, rgb, &bits); break;
                        case 1: sh_to_rgb<1>(sh, dx, dy, dz, rgb, &bits); break;
                        case 2: sh_to_rgb<2>(sh, dx, dy, dz, rgb, &bits); break;
                        default: sh_to_rgb<3>(sh, dx, dy, dz, rgb, &bits); break;
                    }
                }
                my_radius = (int)radius_f;
                my_tiles = (uint32_t)area;
                float4* o = reinterpret_cast<float4*>(rec) + 3 * (size_t)idx;
                o[0] = make_float4(px, py, c * det_inv, -b * det_inv);
                o[1] = make_float4(a * det_inv, opac, tvz, rgb[0]);
                // cull threshold of the render kernels: alpha >= 1/255 needs A dx^2 + 2B dx dy + C dy^2 <= 2 ln(255 o)
                const float thr = opac > 0.0f ? 2.0f * __logf(255.0f * opac) + 0.002f : -1.0f;
                o[2] = make_float4(rgb[1], rgb[2], thr, 0.0f);
                depths[idx] = tvz;
                float2* co = reinterpret_cast<float2*>(cov3D_out) + 3 * (size_t)idx;
                co[0] = make_float2(c0, c1);
                co[1] = make_float2(c2, c3);
                co[2] = make_float2(c4, c5);
                clamped[idx] = bits;
            }
        }
    }
    radii[idx] = my_radius;
    tiles_touched[idx] = my_tiles;
}

// key = (tile_id << 32) | float_bits(depth); value = Gaussian index; tiles y-major then x.
// Fused: the per-digit-place histograms the onesweep sort needs (hist[pass][256]) are accumulated here --
// all keys of one Gaussian share their low 32 bits, so the four depth digits cost one weighted shared-memory
// atomic per Gaussian instead of one per key, and the separate histogram read of the key array disappears.
constexpr int DUP_MAX_PASSES = 8;
__global__ void __launch_bounds__(256)
duplicate_kernel(int P, int gx, int gy, const int32_t* __restrict__ radii, const float* __restrict__ rec,
                 const float* __restrict__ depths, const uint32_t* __restrict__ point_offsets,
                 uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, uint32_t* __restrict__ hist, int end_bit) {
    __shared__ uint32_t s_hist[DUP_MAX_PASSES * 256];
    const int passes = (end_bit + 7) / 8;
    for (int i = threadIdx.x; i < passes * 256; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < P; idx += gridDim.x * blockDim.x) {
        const int rad = radii[idx];
        if (rad <= 0) continue;
        uint32_t off = (idx == 0) ? 0u : point_offsets[idx - 1];
        const float2 xy = *reinterpret_cast<const float2*>(rec + (size_t)idx * REC_FLOATS);
        int x0, y0, x1, y1;
        get_rect(xy.x, xy.y, (float)rad, gx, gy, x0, y0, x1, y1);
        const uint32_t d32 = __float_as_uint(depths[idx]);
        const uint64_t dbits = (uint64_t)d32;
        const uint32_t ntiles = (uint32_t)((x1 - x0) * (y1 - y0));
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            if (p < passes) {
                const int bits = min(8, end_bit - 8 * p);
                atomicAdd(&s_hist[p * 256 + ((d32 >> (8 * p)) & ((1u << bits) - 1u))], ntiles);
            }
        }
        for (int ty = y0; ty < y1; ++ty) {
            for (int tx = x0; tx < x1; ++tx) {
                const uint32_t tile = (uint32_t)(ty * gx + tx);
                keys[off] = ((uint64_t)tile << 32) | dbits;
                vals[off] = (uint32_t)idx;
                ++off;
                for (int p = 4; p < passes; ++p) {
                    const int bits = min(8, end_bit - 8 * p);
                    atomicAdd(&s_hist[p * 256 + ((tile >> (8 * (p - 4))) & ((1u << bits) - 1u))], 1u);
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * 256; i += blockDim.x) {
        const uint32_t c = s_hist[i];
        if (c) atomicAdd(&hist[i], c);
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

cudaError_t launch_preprocess(int P, const CameraParams& cam, const float* means3D, const float* scales,
                              const float* rotations, const float* opacities, const float* shs,
                              const float* colors_precomp, const float* cov3D_precomp, int32_t* radii,
                              const GeomViews& g, cudaStream_t st) {
    if (P <= 0) return cudaSuccess;
    preprocess_kernel<<<(P + 255) / 256, 256, 0, st>>>(P, cam, means3D, scales, rotations, opacities, shs,
                                                        colors_precomp, cov3D_precomp, radii, g.rec, g.depths,
                                                        g.cov3D, g.clamped, g.tiles_touched);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_duplicate(int P, const CameraParams& cam, const int32_t* radii, const GeomViews& g,
                             uint64_t* keys, uint32_t* vals, uint32_t* hist, int end_bit, cudaStream_t st) {
    if (P <= 0) return cudaSuccess;
    const int blocks = min((P + 255) / 256, NUM_SMS * 8);
    duplicate_kernel<<<blocks, 256, 0, st>>>(P, cam.grid_x, cam.grid_y, radii, g.rec, g.depths,
                                              g.point_offsets, keys, vals, hist, end_bit);
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

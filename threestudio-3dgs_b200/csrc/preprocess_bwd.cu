// K8 preprocess backward: 2-D stage gradients -> parameter gradients, one fused kernel.
//
// Replaces upstream backward.cu computeCov2DCUDA + preprocessCUDA (+ computeColorFromSH and
// computeCov3D backward) [UPSTREAM-RECALL; SURVEY.md Appendix A]; the gradients it produces are
// the ones geometry/gaussian_base.py:815-819 (means2D.grad) and the Adam groups
// (geometry/gaussian_base.py:470-525) consume.  Conventions kept from upstream (Appendix A.1):
// the FOV clamp zeroes dL/dt.x|y and treats the clamped t.x|y as constant w.r.t. t.z; the conic
// gradient uses 1/(det^2 + 1e-7); culled Gaussians (radii == 0) get exactly-zero gradients;
// dL/dmeans2D is the gradient w.r.t. the NDC position (pixel gradient x W/2, H/2), z = 0.
#include "common.cuh"

namespace b200splat {

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

template <bool ACC>
__device__ __forceinline__ void put(float* p, float v) {
    if (ACC) *p += v; else *p = v;
}

// SH backward for one Gaussian.  g[c] = dL/drgb_c already masked by the clamp flags.
// Writes dL/dsh (all M coefficients; zeros above the active degree) and returns dL/ddir.
template <int DEG, bool ACC>
__device__ __forceinline__ void sh_backward(const float* __restrict__ sh, float* __restrict__ dsh, int M,
                                            float x, float y, float z, const float g[3], float dd[3]) {
    constexpr int K = (DEG + 1) * (DEG + 1);
    float B[K], Bx[K], By[K], Bz[K];
    B[0] = SH_C0, Bx[0] = By[0] = Bz[0] = 0.f;
    if (DEG > 0) {
        B[1] = -SH_C1 * y, Bx[1] = 0.f, By[1] = -SH_C1, Bz[1] = 0.f;
        B[2] = SH_C1 * z, Bx[2] = 0.f, By[2] = 0.f, Bz[2] = SH_C1;
        B[3] = -SH_C1 * x, Bx[3] = -SH_C1, By[3] = 0.f, Bz[3] = 0.f;
    }
    if (DEG > 1) {
        const float xx = x * x, yy = y * y, zz = z * z;
        B[4] = SH_C2_0 * x * y, Bx[4] = SH_C2_0 * y, By[4] = SH_C2_0 * x, Bz[4] = 0.f;
        B[5] = SH_C2_1 * y * z, Bx[5] = 0.f, By[5] = SH_C2_1 * z, Bz[5] = SH_C2_1 * y;
        B[6] = SH_C2_2 * (2.f * zz - xx - yy), Bx[6] = SH_C2_2 * -2.f * x, By[6] = SH_C2_2 * -2.f * y,
        Bz[6] = SH_C2_2 * 4.f * z;
        B[7] = SH_C2_3 * x * z, Bx[7] = SH_C2_3 * z, By[7] = 0.f, Bz[7] = SH_C2_3 * x;
        B[8] = SH_C2_4 * (xx - yy), Bx[8] = SH_C2_4 * 2.f * x, By[8] = SH_C2_4 * -2.f * y, Bz[8] = 0.f;
    }
    if (DEG > 2) {
        const float xx = x * x, yy = y * y, zz = z * z;
        B[9] = SH_C3_0 * y * (3.f * xx - yy), Bx[9] = SH_C3_0 * 6.f * x * y, By[9] = SH_C3_0 * (3.f * xx - 3.f * yy),
        Bz[9] = 0.f;
        B[10] = SH_C3_1 * x * y * z, Bx[10] = SH_C3_1 * y * z, By[10] = SH_C3_1 * x * z, Bz[10] = SH_C3_1 * x * y;
        B[11] = SH_C3_2 * y * (4.f * zz - xx - yy), Bx[11] = SH_C3_2 * -2.f * x * y,
        By[11] = SH_C3_2 * (4.f * zz - xx - 3.f * yy), Bz[11] = SH_C3_2 * 8.f * y * z;
        B[12] = SH_C3_3 * z * (2.f * zz - 3.f * xx - 3.f * yy), Bx[12] = SH_C3_3 * -6.f * x * z,
        By[12] = SH_C3_3 * -6.f * y * z, Bz[12] = SH_C3_3 * (6.f * zz - 3.f * xx - 3.f * yy);
        B[13] = SH_C3_4 * x * (4.f * zz - xx - yy), Bx[13] = SH_C3_4 * (4.f * zz - 3.f * xx - yy),
        By[13] = SH_C3_4 * -2.f * x * y, Bz[13] = SH_C3_4 * 8.f * x * z;
        B[14] = SH_C3_5 * z * (xx - yy), Bx[14] = SH_C3_5 * 2.f * x * z, By[14] = SH_C3_5 * -2.f * y * z,
        Bz[14] = SH_C3_5 * (xx - yy);
        B[15] = SH_C3_6 * x * (xx - 3.f * yy), Bx[15] = SH_C3_6 * (3.f * xx - 3.f * yy),
        By[15] = SH_C3_6 * -6.f * x * y, Bz[15] = 0.f;
    }
    float ddx = 0.f, ddy = 0.f, ddz = 0.f;
    if ((M & 3) == 0 && M <= 16) {
        // 16-byte path: 4 coefficients (12 floats = 3 float4) per step; per-Gaussian base is 16 B aligned
        const float4* __restrict__ in4 = reinterpret_cast<const float4*>(sh);
        float4* __restrict__ out4 = reinterpret_cast<float4*>(dsh);
#pragma unroll
        for (int k0 = 0; k0 < 16; k0 += 4) {
            if (k0 >= M) break;
            float o[12];
            if (k0 < K) {
                const float4 a = __ldg(in4 + 3 * (k0 >> 2)), b = __ldg(in4 + 3 * (k0 >> 2) + 1),
                             c = __ldg(in4 + 3 * (k0 >> 2) + 2);
                const float sv[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    // K is a compile-time constant: the k < K test folds after unrolling of the caller's DEG
                    const int kk = (k0 + t < K) ? k0 + t : 0;   // compile-time after unrolling
                    const bool live = k0 + t < K;
                    const float Bk = live ? B[kk] : 0.f, Bxk = live ? Bx[kk] : 0.f, Byk = live ? By[kk] : 0.f,
                                Bzk = live ? Bz[kk] : 0.f;
                    const float gs = g[0] * sv[3 * t] + g[1] * sv[3 * t + 1] + g[2] * sv[3 * t + 2];
                    ddx += Bxk * gs, ddy += Byk * gs, ddz += Bzk * gs;
                    o[3 * t] = Bk * g[0], o[3 * t + 1] = Bk * g[1], o[3 * t + 2] = Bk * g[2];
                }
            } else {
#pragma unroll
                for (int t = 0; t < 12; ++t) o[t] = 0.f;
            }
            float4* dst = out4 + 3 * (k0 >> 2);
            if (ACC) {
                if (k0 < K) {
                    const float4 p0 = dst[0], p1 = dst[1], p2 = dst[2];
                    dst[0] = make_float4(p0.x + o[0], p0.y + o[1], p0.z + o[2], p0.w + o[3]);
                    dst[1] = make_float4(p1.x + o[4], p1.y + o[5], p1.z + o[6], p1.w + o[7]);
                    dst[2] = make_float4(p2.x + o[8], p2.y + o[9], p2.z + o[10], p2.w + o[11]);
                }
            } else {
                dst[0] = make_float4(o[0], o[1], o[2], o[3]);
                dst[1] = make_float4(o[4], o[5], o[6], o[7]);
                dst[2] = make_float4(o[8], o[9], o[10], o[11]);
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float s0 = __ldg(sh + 3 * k), s1 = __ldg(sh + 3 * k + 1), s2 = __ldg(sh + 3 * k + 2);
            const float gs = g[0] * s0 + g[1] * s1 + g[2] * s2;
            ddx += Bx[k] * gs, ddy += By[k] * gs, ddz += Bz[k] * gs;
            put<ACC>(dsh + 3 * k, B[k] * g[0]);
            put<ACC>(dsh + 3 * k + 1, B[k] * g[1]);
            put<ACC>(dsh + 3 * k + 2, B[k] * g[2]);
        }
        if (!ACC) {
            for (int k = 3 * K; k < 3 * M; ++k) dsh[k] = 0.f;
        }
    }
    dd[0] = ddx, dd[1] = ddy, dd[2] = ddz;
}

template <bool ACC>
__global__ void __launch_bounds__(256)
preprocess_backward_kernel(int P, CameraParams cam, const float* __restrict__ means3D,
                           const float* __restrict__ scales, const float* __restrict__ rotations,
                           const float* __restrict__ shs, const float* __restrict__ cov3D_precomp,
                           const int32_t* __restrict__ radii, const float* __restrict__ cov3D_saved,
                           const uint8_t* __restrict__ clamped, const float* __restrict__ grad2d,
                           float* __restrict__ dL_dmeans3D, float* __restrict__ dL_dmeans2D,
                           float* __restrict__ dL_dshs, float* __restrict__ dL_dcolors,
                           float* __restrict__ dL_dopacity, float* __restrict__ dL_dscales,
                           float* __restrict__ dL_drotations, float* __restrict__ dL_dcov3D,
                           float* __restrict__ stat_grad_accum, float* __restrict__ stat_denom,
                           float* __restrict__ stat_max_radii) {
    __shared__ float sV[16], sP[16], sC[3];
    if (threadIdx.x < 16) {
        sV[threadIdx.x] = cam.view[threadIdx.x];
        sP[threadIdx.x] = cam.proj[threadIdx.x];
    }
    if (threadIdx.x < 3) sC[threadIdx.x] = cam.campos[threadIdx.x];
    __syncthreads();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const int M = cam.M;
    if (shs != nullptr) {   // SH block of this Gaussian towards L2 while the covariance chain runs
        const char* shp = reinterpret_cast<const char*>(shs + (size_t)idx * M * 3);
        prefetch_l2(shp);
        if (M * 12 > 128) prefetch_l2(shp + 128);
        if (ACC) {
            const char* dp = reinterpret_cast<const char*>(dL_dshs + (size_t)idx * M * 3);
            prefetch_l2(dp);
            if (M * 12 > 128) prefetch_l2(dp + 128);
        }
    }
    const int my_radius = radii[idx];
    if (my_radius <= 0) {
        if (!ACC) {
            dL_dmeans3D[3 * idx] = dL_dmeans3D[3 * idx + 1] = dL_dmeans3D[3 * idx + 2] = 0.f;
            dL_dmeans2D[3 * idx] = dL_dmeans2D[3 * idx + 1] = dL_dmeans2D[3 * idx + 2] = 0.f;
            dL_dopacity[idx] = 0.f;
            if (dL_dshs) for (int k = 0; k < 3 * M; ++k) dL_dshs[(size_t)idx * 3 * M + k] = 0.f;
            if (dL_dcolors) dL_dcolors[3 * idx] = dL_dcolors[3 * idx + 1] = dL_dcolors[3 * idx + 2] = 0.f;
            if (dL_dscales) dL_dscales[3 * idx] = dL_dscales[3 * idx + 1] = dL_dscales[3 * idx + 2] = 0.f;
            if (dL_drotations)
                reinterpret_cast<float4*>(dL_drotations)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (dL_dcov3D) for (int k = 0; k < 6; ++k) dL_dcov3D[(size_t)idx * 6 + k] = 0.f;
        }
        return;
    }
    const float4* gp = reinterpret_cast<const float4*>(grad2d) + 3 * (size_t)idx;
    const float4 g0 = gp[0], g1 = gp[1], g2 = gp[2];
    const float g_px = g0.x, g_py = g0.y, g_ca = g0.z, g_cb = g0.w, g_cc = g1.x, g_op = g1.y;
    float g_rgb[3] = {g1.z, g1.w, g2.x};
    const float g_depth = g2.y;

    const float x = means3D[3 * idx], y = means3D[3 * idx + 1], z = means3D[3 * idx + 2];
    float dmx = 0.f, dmy = 0.f, dmz = 0.f;

    // ---- conic -> cov2D -> (Sigma3, t) ----------------------------------------------------------
    const float tvx = sV[0] * x + sV[4] * y + sV[8] * z + sV[12];
    const float tvy = sV[1] * x + sV[5] * y + sV[9] * z + sV[13];
    const float tvz = sV[2] * x + sV[6] * y + sV[10] * z + sV[14];
    const float txtz = tvx / tvz, tytz = tvy / tvz;
    const float tx = fminf(cam.limx, fmaxf(-cam.limx, txtz)) * tvz;
    const float ty = fminf(cam.limy, fmaxf(-cam.limy, tytz)) * tvz;
    const float xmul = (txtz < -cam.limx || txtz > cam.limx) ? 0.f : 1.f;
    const float ymul = (tytz < -cam.limy || tytz > cam.limy) ? 0.f : 1.f;
    const float itz = 1.0f / tvz, itz2 = itz * itz, itz3 = itz2 * itz;
    const float J00 = cam.focal_x * itz, J02 = -cam.focal_x * tx * itz2;
    const float J11 = cam.focal_y * itz, J12 = -cam.focal_y * ty * itz2;
    const float M0[3] = {J00 * sV[0] + J02 * sV[2], J00 * sV[4] + J02 * sV[6], J00 * sV[8] + J02 * sV[10]};
    const float M1[3] = {J11 * sV[1] + J12 * sV[2], J11 * sV[5] + J12 * sV[6], J11 * sV[9] + J12 * sV[10]};
    float c0, c1, c2, c3, c4, c5;
    {
        const float* cs = (cov3D_precomp ? cov3D_precomp : cov3D_saved) + 6 * (size_t)idx;
        c0 = cs[0], c1 = cs[1], c2 = cs[2], c3 = cs[3], c4 = cs[4], c5 = cs[5];
    }
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
    float dS[6];
    dS[0] = M0[0] * M0[0] * dLa + M0[0] * M1[0] * dLb + M1[0] * M1[0] * dLc;
    dS[3] = M0[1] * M0[1] * dLa + M0[1] * M1[1] * dLb + M1[1] * M1[1] * dLc;
    dS[5] = M0[2] * M0[2] * dLa + M0[2] * M1[2] * dLb + M1[2] * M1[2] * dLc;
    dS[1] = 2.f * M0[0] * M0[1] * dLa + (M0[0] * M1[1] + M0[1] * M1[0]) * dLb + 2.f * M1[0] * M1[1] * dLc;
    dS[2] = 2.f * M0[0] * M0[2] * dLa + (M0[0] * M1[2] + M0[2] * M1[0]) * dLb + 2.f * M1[0] * M1[2] * dLc;
    dS[4] = 2.f * M0[1] * M0[2] * dLa + (M0[1] * M1[2] + M0[2] * M1[1]) * dLb + 2.f * M1[1] * M1[2] * dLc;
    {
        float dM0[3], dM1[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            dM0[k] = 2.f * dLa * S0[k] + dLb * S1[k];
            dM1[k] = 2.f * dLc * S1[k] + dLb * S0[k];
        }
        const float dJ00 = dM0[0] * sV[0] + dM0[1] * sV[4] + dM0[2] * sV[8];
        const float dJ02 = dM0[0] * sV[2] + dM0[1] * sV[6] + dM0[2] * sV[10];
        const float dJ11 = dM1[0] * sV[1] + dM1[1] * sV[5] + dM1[2] * sV[9];
        const float dJ12 = dM1[0] * sV[2] + dM1[1] * sV[6] + dM1[2] * sV[10];
        const float dtx = xmul * -cam.focal_x * itz2 * dJ02;
        const float dty = ymul * -cam.focal_y * itz2 * dJ12;
        const float dtz = -cam.focal_x * itz2 * dJ00 - cam.focal_y * itz2 * dJ11 +
                          2.f * cam.focal_x * tx * itz3 * dJ02 + 2.f * cam.focal_y * ty * itz3 * dJ12;
        dmx += sV[0] * dtx + sV[1] * dty + sV[2] * dtz;
        dmy += sV[4] * dtx + sV[5] * dty + sV[6] * dtz;
        dmz += sV[8] * dtx + sV[9] * dty + sV[10] * dtz;
    }
    // ---- mean2D (NDC) and depth -----------------------------------------------------------------
    const float gnx = g_px * 0.5f * (float)cam.W, gny = g_py * 0.5f * (float)cam.H;
    {
        const float hx = sP[0] * x + sP[4] * y + sP[8] * z + sP[12];
        const float hy = sP[1] * x + sP[5] * y + sP[9] * z + sP[13];
        const float hw = sP[3] * x + sP[7] * y + sP[11] * z + sP[15];
        const float pw = 1.0f / (hw + PW_EPS);
        const float mul1 = hx * pw * pw, mul2 = hy * pw * pw;
        dmx += (sP[0] * pw - sP[3] * mul1) * gnx + (sP[1] * pw - sP[3] * mul2) * gny;
        dmy += (sP[4] * pw - sP[7] * mul1) * gnx + (sP[5] * pw - sP[7] * mul2) * gny;
        dmz += (sP[8] * pw - sP[11] * mul1) * gnx + (sP[9] * pw - sP[11] * mul2) * gny;
        dmx += sV[2] * g_depth;
        dmy += sV[6] * g_depth;
        dmz += sV[10] * g_depth;
    }
    // ---- colour ---------------------------------------------------------------------------------
    if (shs != nullptr) {
        const uint8_t bits = clamped[idx];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch)
            if (bits & (1u << ch)) g_rgb[ch] = 0.f;
        float dx = x - sC[0], dy = y - sC[1], dz = z - sC[2];
        const float n = sqrtf(dx * dx + dy * dy + dz * dz);
        const float in = 1.0f / n;
        dx *= in, dy *= in, dz *= in;
        float dd[3];
        const float* sh = shs + (size_t)idx * M * 3;
        float* dsh = dL_dshs + (size_t)idx * M * 3;
        switch (cam.sh_degree) {
            case 0: sh_backward<0, ACC>(sh, dsh, M, dx, dy, dz, g_rgb, dd); break;
            case 1: sh_backward<1, ACC>(sh, dsh, M, dx, dy, dz, g_rgb, dd); break;
            case 2: sh_backward<2, ACC>(sh, dsh, M, dx, dy, dz, g_rgb, dd); break;
            default: sh_backward<3, ACC>(sh, dsh, M, dx, dy, dz, g_rgb, dd); break;
        }
        // through dir = d / |d|
        const float dot = dx * dd[0] + dy * dd[1] + dz * dd[2];
        dmx += (dd[0] - dx * dot) * in;
        dmy += (dd[1] - dy * dot) * in;
        dmz += (dd[2] - dz * dot) * in;
    } else if (dL_dcolors) {
        put<ACC>(dL_dcolors + 3 * idx, g_rgb[0]);
        put<ACC>(dL_dcolors + 3 * idx + 1, g_rgb[1]);
        put<ACC>(dL_dcolors + 3 * idx + 2, g_rgb[2]);
    }
    // ---- Sigma3 -> scale / rotation -------------------------------------------------------------
    if (scales != nullptr && dL_dscales != nullptr) {
        const float mod = cam.scale_modifier;
        const float sx = mod * scales[3 * idx], sy = mod * scales[3 * idx + 1], sz = mod * scales[3 * idx + 2];
        const float4 q = reinterpret_cast<const float4*>(rotations)[idx];
        const float r = q.x, qx = q.y, qy = q.z, qz = q.w;
        const float R[3][3] = {{1.f - 2.f * (qy * qy + qz * qz), 2.f * (qx * qy - r * qz), 2.f * (qx * qz + r * qy)},
                               {2.f * (qx * qy + r * qz), 1.f - 2.f * (qx * qx + qz * qz), 2.f * (qy * qz - r * qx)},
                               {2.f * (qx * qz - r * qy), 2.f * (qy * qz + r * qx), 1.f - 2.f * (qx * qx + qy * qy)}};
        const float s[3] = {sx, sy, sz};
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
    put<ACC>(dL_dmeans2D + 3 * idx, gnx);
    put<ACC>(dL_dmeans2D + 3 * idx + 1, gny);
    if (!ACC) dL_dmeans2D[3 * idx + 2] = 0.f;
    put<ACC>(dL_dopacity + idx, g_op);
    // fused densification statistics of this view (geometry/gaussian_base.py:815-819, 846-851)
    if (stat_grad_accum) stat_grad_accum[idx] += sqrtf(gnx * gnx + gny * gny);
    if (stat_denom) stat_denom[idx] += 1.0f;
    if (stat_max_radii) stat_max_radii[idx] = fmaxf(stat_max_radii[idx], (float)my_radius);
}

cudaError_t launch_preprocess_backward(int P, const CameraParams& cam, const float* means3D, const float* scales,
                                       const float* rotations, const float* shs, const float* cov3D_precomp,
                                       const int32_t* radii, const GeomViews& g, const float* grad2d,
                                       float* dL_dmeans3D, float* dL_dmeans2D, float* dL_dshs, float* dL_dcolors,
                                       float* dL_dopacity, float* dL_dscales, float* dL_drotations, float* dL_dcov3D,
                                       float* stat_grad_accum, float* stat_denom, float* stat_max_radii,
                                       int accumulate, cudaStream_t st) {
    if (P <= 0) return cudaSuccess;
    const int grid = (P + 255) / 256;
    if (accumulate)
        preprocess_backward_kernel<true><<<grid, 256, 0, st>>>(
            P, cam, means3D, scales, rotations, shs, cov3D_precomp, radii, g.cov3D, g.clamped, grad2d, dL_dmeans3D,
            dL_dmeans2D, dL_dshs, dL_dcolors, dL_dopacity, dL_dscales, dL_drotations, dL_dcov3D, stat_grad_accum,
            stat_denom, stat_max_radii);
    else
        preprocess_backward_kernel<false><<<grid, 256, 0, st>>>(
            P, cam, means3D, scales, rotations, shs, cov3D_precomp, radii, g.cov3D, g.clamped, grad2d, dL_dmeans3D,
            dL_dmeans2D, dL_dshs, dL_dcolors, dL_dopacity, dL_dscales, dL_drotations, dL_dcov3D, stat_grad_accum,
            stat_denom, stat_max_radii);
    count_launch();
    return cudaGetLastError();
}

}  // namespace b200splat

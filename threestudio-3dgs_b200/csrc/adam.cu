// Fused "activation backward + Adam step" on the rasterizer's gradient buffer.
//
// The reference keeps RAW parameters (_xyz, _features_dc, _features_rest, _opacity, _scaling, _rotation) in six
// Adam parameter groups (geometry/gaussian_base.py:470-525: torch.optim.Adam(l, lr=0.0, eps=1e-15), one lr per group,
// scheduled per step :539-572) and feeds the rasterizer ACTIVATED values (exp, sigmoid, F.normalize, clip:
// geometry/gaussian_base.py:240-248, :371-411).  After backward, autograd walks the activations back (~10 kernels) and
// Adam's foreach path runs ~12 more over 4 tensors per group.  Here one pass does both: each thread takes one
// Gaussian's gradients with respect to the activated values -- exactly what preprocess-backward wrote into the packed
// buffer (and what the all-reduce left there) -- applies the activation Jacobians in registers, and updates
// exp_avg / exp_avg_sq / parameter in place.  A second, purely elementwise kernel handles the 3(M-1) higher-order SH
// coefficients per Gaussian.  HBM-bound: 4 (13 + 3M) + 7 * 4 (11 + 3M) bytes per Gaussian.
//
// Arithmetic mirrors torch.optim.Adam's single-tensor path (amsgrad off, weight decay 0, maximize off):
//   exp_avg.lerp_(g, 1 - b1); exp_avg_sq.mul_(b2).addcmul_(g, g, 1 - b2);
//   denom = sqrt(exp_avg_sq) / sqrt(1 - b2^t) + eps; p -= (lr / (1 - b1^t)) * exp_avg / denom
#include "common.cuh"

#include "../../include/b200splat.h"

namespace b200splat {

struct AdamTab {
    int P, M;
    float *xyz, *f_dc, *f_rest, *opacity, *scaling, *rotation;
    float *m_xyz, *m_f_dc, *m_f_rest, *m_opacity, *m_scaling, *m_rotation;
    float *v_xyz, *v_f_dc, *v_f_rest, *v_opacity, *v_scaling, *v_rotation;
    const float *g_means3D, *g_shs, *g_opacities, *g_scales, *g_rotations;
    float step_size[6];   // lr / (1 - b1^t): xyz, f_dc, f_rest, opacity, scaling, rotation
    float b2, omb1, omb2, eps, inv_bc2_sqrt, color_clip;   // omb = 1 - beta, rounded from double as torch does
};

__device__ __forceinline__ void adam1(float& p, float& m, float& v, float g, float step_size, const AdamTab& t) {
    m = m + (g - m) * t.omb1;
    v = v * t.b2 + t.omb2 * g * g;
    const float denom = sqrtf(v) * t.inv_bc2_sqrt + t.eps;
    p = p - step_size * (m / denom);
}

template <int N>
__device__ __forceinline__ void adam_vec(float* p, float* m, float* v, const float (&g)[N], size_t base, float step_size,
                                         const AdamTab& t) {
#pragma unroll
    for (int c = 0; c < N; ++c) {
        float pp = p[base + c], mm = m[base + c], vv = v[base + c];
        adam1(pp, mm, vv, g[c], step_size, t);
        p[base + c] = pp, m[base + c] = mm, v[base + c] = vv;
    }
}

// one thread per Gaussian: xyz 3 | f_dc 3 | opacity 1 | scaling 3 | rotation 4
__global__ void __launch_bounds__(256) adam_gaussian_kernel(const __grid_constant__ AdamTab t) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= t.P) return;
    {   // position: identity activation
        const float g[3] = {__ldg(t.g_means3D + 3 * (size_t)i), __ldg(t.g_means3D + 3 * (size_t)i + 1),
                            __ldg(t.g_means3D + 3 * (size_t)i + 2)};
        adam_vec<3>(t.xyz, t.m_xyz, t.v_xyz, g, 3 * (size_t)i, t.step_size[0], t);
    }
    {   // SH DC: features_dc.clip(-c, c) -> the gradient passes where |raw| <= c
        const float* gs = t.g_shs + (size_t)i * 3 * t.M;
        float g[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float raw = t.f_dc[3 * (size_t)i + c];
            g[c] = (raw >= -t.color_clip && raw <= t.color_clip) ? __ldg(gs + c) : 0.f;
        }
        adam_vec<3>(t.f_dc, t.m_f_dc, t.v_f_dc, g, 3 * (size_t)i, t.step_size[1], t);
    }
    {   // opacity = sigmoid(raw)
        const float raw = t.opacity[i];
        const float o = 1.0f / (1.0f + expf(-raw));
        const float g[1] = {__ldg(t.g_opacities + i) * o * (1.0f - o)};
        adam_vec<1>(t.opacity, t.m_opacity, t.v_opacity, g, (size_t)i, t.step_size[3], t);
    }
    {   // scaling = exp(raw)
        float g[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) g[c] = __ldg(t.g_scales + 3 * (size_t)i + c) * expf(t.scaling[3 * (size_t)i + c]);
        adam_vec<3>(t.scaling, t.m_scaling, t.v_scaling, g, 3 * (size_t)i, t.step_size[4], t);
    }
    {   // rotation = q / max(|q|, 1e-12)
        const float4 q = *reinterpret_cast<const float4*>(t.rotation + 4 * (size_t)i);
        const float4 gq = __ldg(reinterpret_cast<const float4*>(t.g_rotations + 4 * (size_t)i));
        const float n = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
        float g[4];
        if (n >= 1e-12f) {
            const float inv = 1.0f / n;
            const float hx = q.x * inv, hy = q.y * inv, hz = q.z * inv, hw = q.w * inv;
            const float d = hx * gq.x + hy * gq.y + hz * gq.z + hw * gq.w;
            g[0] = (gq.x - hx * d) * inv, g[1] = (gq.y - hy * d) * inv, g[2] = (gq.z - hz * d) * inv,
            g[3] = (gq.w - hw * d) * inv;
        } else {
            g[0] = gq.x * 1e12f, g[1] = gq.y * 1e12f, g[2] = gq.z * 1e12f, g[3] = gq.w * 1e12f;
        }
        adam_vec<4>(t.rotation, t.m_rotation, t.v_rotation, g, 4 * (size_t)i, t.step_size[5], t);
    }
}

// higher-order SH coefficients: element e of features_rest (P, M-1, 3) <- g_shs (P, M, 3)[:, 1:, :]
__global__ void __launch_bounds__(256) adam_sh_rest_kernel(const __grid_constant__ AdamTab t) {
    const int per = 3 * (t.M - 1);
    const size_t n = (size_t)t.P * per;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
        const size_t i = e / per;
        const int j = (int)(e - i * per);
        const float g = __ldg(t.g_shs + i * 3 * t.M + 3 + j);
        float p = t.f_rest[e], m = t.m_f_rest[e], v = t.v_f_rest[e];
        adam1(p, m, v, g, t.step_size[2], t);
        t.f_rest[e] = p, t.m_f_rest[e] = m, t.v_f_rest[e] = v;
    }
}

}  // namespace b200splat

using namespace b200splat;
extern int b200splat_set_error(int code, const char* msg);

extern "C" int b200splat_adam_step(const b200splat_adam_args* a) {
    if (!a) return b200splat_set_error(B200SPLAT_ERR_INVALID, "null args");
    if (a->P < 0 || a->M < 1) return b200splat_set_error(B200SPLAT_ERR_INVALID, "P >= 0 and M >= 1 required");
    if (a->step < 1) return b200splat_set_error(B200SPLAT_ERR_INVALID, "step is 1-based");
    if (a->P == 0) return B200SPLAT_OK;
    const bool rest = a->M > 1;
    const void* need[] = {a->xyz, a->features_dc, a->opacity, a->scaling, a->rotation, a->m_xyz, a->m_features_dc,
                          a->m_opacity, a->m_scaling, a->m_rotation, a->v_xyz, a->v_features_dc, a->v_opacity,
                          a->v_scaling, a->v_rotation, a->g_means3D, a->g_shs, a->g_opacities, a->g_scales,
                          a->g_rotations};
    for (const void* p : need)
        if (!p) return b200splat_set_error(B200SPLAT_ERR_INVALID, "null parameter / state / gradient pointer");
    if (rest && (!a->features_rest || !a->m_features_rest || !a->v_features_rest))
        return b200splat_set_error(B200SPLAT_ERR_INVALID, "M > 1 needs features_rest and its state");
    if ((reinterpret_cast<uintptr_t>(a->rotation) | reinterpret_cast<uintptr_t>(a->g_rotations)) & 15)
        return b200splat_set_error(B200SPLAT_ERR_INVALID, "rotation / g_rotations must be 16-byte aligned");
    AdamTab t;
    t.P = a->P, t.M = a->M;
    t.xyz = a->xyz, t.f_dc = a->features_dc, t.f_rest = a->features_rest, t.opacity = a->opacity;
    t.scaling = a->scaling, t.rotation = a->rotation;
    t.m_xyz = a->m_xyz, t.m_f_dc = a->m_features_dc, t.m_f_rest = a->m_features_rest, t.m_opacity = a->m_opacity;
    t.m_scaling = a->m_scaling, t.m_rotation = a->m_rotation;
    t.v_xyz = a->v_xyz, t.v_f_dc = a->v_features_dc, t.v_f_rest = a->v_features_rest, t.v_opacity = a->v_opacity;
    t.v_scaling = a->v_scaling, t.v_rotation = a->v_rotation;
    t.g_means3D = a->g_means3D, t.g_shs = a->g_shs, t.g_opacities = a->g_opacities, t.g_scales = a->g_scales;
    t.g_rotations = a->g_rotations;
    // bias corrections in double on the host, as torch does with Python floats
    const double bc1 = 1.0 - pow(a->beta1, (double)a->step);
    const double bc2 = 1.0 - pow(a->beta2, (double)a->step);
    for (int k = 0; k < 6; ++k) t.step_size[k] = (float)(a->lr[k] / bc1);
    t.b2 = (float)a->beta2, t.omb1 = (float)(1.0 - a->beta1), t.omb2 = (float)(1.0 - a->beta2);
    t.eps = a->eps, t.inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
    t.color_clip = a->color_clip;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(a->stream);
    adam_gaussian_kernel<<<(a->P + 255) / 256, 256, 0, st>>>(t);
    count_launch();
    if (rest) {
        const size_t n = (size_t)a->P * 3 * (a->M - 1);
        const unsigned blocks = (unsigned)((n + 255) / 256 < (size_t)NUM_SMS * 16 ? (n + 255) / 256 : (size_t)NUM_SMS * 16);
        adam_sh_rest_kernel<<<blocks, 256, 0, st>>>(t);
        count_launch();
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return b200splat_set_error(B200SPLAT_ERR_CUDA, cudaGetErrorString(e));
    return B200SPLAT_OK;
}

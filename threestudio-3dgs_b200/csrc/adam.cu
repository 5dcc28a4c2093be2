// Fused "activation backward + Adam step" on the rasterizer's gradient buffer.
//
// The reference keeps RAW parameters (_xyz, _features_dc, _features_rest, _opacity, _scaling, _rotation) in six
// Adam parameter groups (geometry/gaussian_base.py:470-525: torch.optim.Adam(l, lr=0.0, eps=1e-15), one lr per group,
// scheduled per step :539-572) and feeds the rasterizer ACTIVATED values (exp, sigmoid, F.normalize, clip:
// geometry/gaussian_base.py:240-248, :371-411).  After backward, autograd walks the activations back (~10 kernels) and
// Adam's foreach path runs ~12 more over 4 tensors per group.  Here one pass does both: it takes the gradients with
// respect to the activated values -- exactly what preprocess-backward wrote into the packed buffer (and what the
// all-reduce left there) -- applies the activation Jacobians in registers, and updates exp_avg / exp_avg_sq /
// parameter in place.  Two launches: one elementwise kernel over every tensor whose activation is
// elementwise, one quaternion kernel.  HBM-bound: 7 * 4 * (11 + 3M) bytes per Gaussian.
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

// Every parameter but the quaternion has an ELEMENTWISE activation (identity, clip, sigmoid, exp), so the (P,3) / (P,1)
// / (P,K,3) tensors are walked as flat arrays, one element per thread, fully coalesced: a thread-per-Gaussian layout
// reads and writes them with a 12-byte stride and measured 3x slower (194 vs ~65 us for 1 M Gaussians).
//   element space: [xyz 3P | f_dc 3P | opacity P | scaling 3P | f_rest 3(M-1)P]
__global__ void __launch_bounds__(256) adam_elementwise_kernel(const __grid_constant__ AdamTab t) {
    const size_t P = (size_t)t.P;
    const size_t n_xyz = 3 * P, n_dc = 3 * P, n_op = P, n_sc = 3 * P, per = 3 * (size_t)(t.M - 1), n_rest = per * P;
    const size_t total = n_xyz + n_dc + n_op + n_sc + n_rest;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        float *p, *m, *v;
        float g, step;
        size_t k = e;
        if (k < n_xyz) {                         // position: identity
            p = t.xyz + k, m = t.m_xyz + k, v = t.v_xyz + k;
            g = __ldg(t.g_means3D + k), step = t.step_size[0];
        } else if ((k -= n_xyz) < n_dc) {        // SH DC: features_dc.clip(-c, c) passes the gradient where |raw| <= c
            p = t.f_dc + k, m = t.m_f_dc + k, v = t.v_f_dc + k;
            const size_t i = k / 3;
            const float raw = *p;
            g = (raw >= -t.color_clip && raw <= t.color_clip) ? __ldg(t.g_shs + i * 3 * t.M + (k - 3 * i)) : 0.f;
            step = t.step_size[1];
        } else if ((k -= n_dc) < n_op) {         // opacity = sigmoid(raw)
            p = t.opacity + k, m = t.m_opacity + k, v = t.v_opacity + k;
            const float o = 1.0f / (1.0f + expf(-*p));
            g = __ldg(t.g_opacities + k) * o * (1.0f - o), step = t.step_size[3];
        } else if ((k -= n_op) < n_sc) {         // scaling = exp(raw)
            p = t.scaling + k, m = t.m_scaling + k, v = t.v_scaling + k;
            g = __ldg(t.g_scales + k) * expf(*p), step = t.step_size[4];
        } else {                                 // higher-order SH: features_rest (P, M-1, 3) <- g_shs (P, M, 3)[:, 1:, :]
            k -= n_sc;
            p = t.f_rest + k, m = t.m_f_rest + k, v = t.v_f_rest + k;
            const size_t i = k / per;
            g = __ldg(t.g_shs + i * 3 * t.M + 3 + (k - i * per)), step = t.step_size[2];
        }
        float pp = *p, mm = *m, vv = *v;
        adam1(pp, mm, vv, g, step, t);
        *p = pp, *m = mm, *v = vv;
    }
}

// rotation = q / max(|q|, 1e-12): one thread per quaternion, 16-byte accesses
__global__ void __launch_bounds__(256) adam_rotation_kernel(const __grid_constant__ AdamTab t) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= t.P) return;
    float4* qp = reinterpret_cast<float4*>(t.rotation) + i;
    float4* mp = reinterpret_cast<float4*>(t.m_rotation) + i;
    float4* vp = reinterpret_cast<float4*>(t.v_rotation) + i;
    const float4 q = *qp;
    const float4 gq = __ldg(reinterpret_cast<const float4*>(t.g_rotations) + i);
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
    float4 pp = q, mm = *mp, vv = *vp;
    adam1(pp.x, mm.x, vv.x, g[0], t.step_size[5], t);
    adam1(pp.y, mm.y, vv.y, g[1], t.step_size[5], t);
    adam1(pp.z, mm.z, vv.z, g[2], t.step_size[5], t);
    adam1(pp.w, mm.w, vv.w, g[3], t.step_size[5], t);
    *qp = pp, *mp = mm, *vp = vv;
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
    if ((reinterpret_cast<uintptr_t>(a->rotation) | reinterpret_cast<uintptr_t>(a->g_rotations) |
         reinterpret_cast<uintptr_t>(a->m_rotation) | reinterpret_cast<uintptr_t>(a->v_rotation)) & 15)
        return b200splat_set_error(B200SPLAT_ERR_INVALID, "rotation, its state and g_rotations must be 16-byte aligned");
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
    const size_t total = (size_t)a->P * (10 + 3 * (size_t)(a->M - 1));
    const size_t want = (total + 255) / 256, cap = (size_t)NUM_SMS * 32;
    adam_elementwise_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(t);
    adam_rotation_kernel<<<(a->P + 255) / 256, 256, 0, st>>>(t);
    count_launch(2);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return b200splat_set_error(B200SPLAT_ERR_CUDA, cudaGetErrorString(e));
    return B200SPLAT_OK;
}

// Fused per-pixel post-ops of the reference's renderer variants, forward and backward, for a batch of views.
//
// After the rasterizer call the reference runs ~25 small PyTorch kernels per view (and as many again in backward):
//   xyz_map = rays_o + depth * rays_d; Depth2Normal (two 3x3 convolutions + cross product); F.normalize;
//   point-light Lambert shading on albedo = image / (alpha + 1e-6); composite over the background; the normal map
//   scaled by alpha; the alpha > 0.99 gradient masks of normal and depth; clamp(0, 1)
//   (renderer/diff_gaussian_rasterizer_shading.py:174-213,222; ..._normal.py:172-201; ..._background.py:130-141;
//   material/gaussian_material.py:70-104).  Semantics restated in oracle/postops.py, which is pinned to those files.
//
// Here: ONE forward kernel and TWO backward kernels for all V views of a step (blockIdx.z = view), one thread per
// pixel, every intermediate in registers.  HBM-bound streaming work: forward reads 14 floats and writes 7 per
// pixel (84 B), backward reads 21 and writes 8 + 9 scratch (the stencil's transpose needs the neighbours'
// dL/d(dx), dL/d(dy), hence the second, gather-only kernel).  Neighbour reads of the 3x3 stencil hit L1/L2.
#include "common.cuh"

#include <cstring>

namespace b200splat {

constexpr int POST_MAX_PER_VIEW = 64;
struct PostTab {
    int V, H, W, mode, shading;
    const float* image;   // (V,3,H,W)
    const float* depth;   // (V,1,H,W)
    const float* alpha;   // (V,1,H,W)
    const float* rays_o;  // (V,H,W,3)
    const float* rays_d;  // (V,H,W,3)
    const float* bg;      // (V,H,W,3)
    const float* light;   // (V,3)
    const float* pred;    // (V,3,H,W) or null
    float ambient[3], diffuse[3];
    int per_view_shading;              // != 0: shading_v[view] instead of `shading`
    uint8_t shading_v[POST_MAX_PER_VIEW];
    int per_view_light;                // != 0: light_v[view] = (ambient rgb, diffuse rgb) instead of ambient / diffuse
    float light_v[POST_MAX_PER_VIEW][6];
    // forward outputs
    float* render;        // (V,3,H,W)
    float* normal;        // (V,3,H,W)
    float* depth_out;     // (V,1,H,W) copy of depth (its gradient is masked in backward)
    // backward
    const float* g_render;
    const float* g_normal;
    const float* g_depth;
    float* d_image;
    float* d_depth;
    float* d_alpha;
    float* d_bg;
    float* scratch;       // (V,H,W,9): dL/d(dx) 3 | dL/d(dy) 3 | direct dL/dxyz 3
};

constexpr int POST_PLAIN = 0, POST_BACKGROUND = 1, POST_NORMAL = 2, POST_SHADING = 3;
constexpr int SHADE_ALBEDO = 0, SHADE_TEXTURELESS = 1, SHADE_DIFFUSE = 2;
constexpr float NORM_EPS = 1e-12f;

struct V3 {
    float x, y, z;
};
__device__ __forceinline__ V3 v3(float x, float y, float z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross3(V3 a, V3 b) {
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ V3 ld3(const float* p) { return v3(__ldg(p), __ldg(p + 1), __ldg(p + 2)); }

// xyz_map at (y, x) of view v; zero outside the image (the convolutions pad with zeros)
__device__ __forceinline__ V3 xyz_at(const PostTab& t, size_t vbase, int y, int x) {
    if (x < 0 || y < 0 || x >= t.W || y >= t.H) return v3(0.f, 0.f, 0.f);
    const size_t p = vbase + (size_t)y * t.W + x;
    const float d = __ldg(t.depth + p);
    return ld3(t.rays_o + 3 * p) + ld3(t.rays_d + 3 * p) * d;
}

// y = x / max(|x|, eps) and its vector-Jacobian product (F.normalize)
__device__ __forceinline__ V3 normalize3(V3 x, float& norm) {
    norm = sqrtf(dot3(x, x));
    return x * (1.0f / fmaxf(norm, NORM_EPS));
}
__device__ __forceinline__ V3 normalize3_vjp(V3 y, float norm, V3 g) {
    if (norm >= NORM_EPS) return (g - y * dot3(y, g)) * (1.0f / norm);
    return g * (1.0f / NORM_EPS);
}

struct Pixel {   // forward intermediates of one pixel
    V3 xyz, a, b, nh, sn, ld, alb, tl, fg, pre;
    float nlen, vlen, dotv, alpha;
};

__device__ __forceinline__ void pixel_forward(const PostTab& t, int v, int y, int x, Pixel& px) {
    const size_t HW = (size_t)t.H * t.W;
    const size_t vbase = (size_t)v * HW;
    const size_t p = vbase + (size_t)y * t.W + x;
    const size_t c0 = (size_t)v * 3 * HW + (size_t)y * t.W + x;
    const V3 img = v3(__ldg(t.image + c0), __ldg(t.image + c0 + HW), __ldg(t.image + c0 + 2 * HW));
    px.alpha = __ldg(t.alpha + p);
    px.nh = v3(0.f, 0.f, 0.f);
    if (t.mode >= POST_NORMAL) {
        px.xyz = xyz_at(t, vbase, y, x);
        px.a = xyz_at(t, vbase, y, x + 1) - xyz_at(t, vbase, y, x - 1);
        px.b = xyz_at(t, vbase, y + 1, x) - xyz_at(t, vbase, y - 1, x);
        const V3 n = cross3(px.a, px.b) * -1.0f;
        px.nh = normalize3(n, px.nlen);
    }
    if (t.mode == POST_SHADING) {
        const float inv = 1.0f / (px.alpha + 1e-6f);
        px.alb = img * inv;
        const V3 lv = ld3(t.light + 3 * v) - px.xyz;
        px.ld = normalize3(lv, px.vlen);
        if (t.pred) {
            float pn;
            px.sn = normalize3(v3(__ldg(t.pred + c0) * 2.f - 1.f, __ldg(t.pred + c0 + HW) * 2.f - 1.f,
                                  __ldg(t.pred + c0 + 2 * HW) * 2.f - 1.f), pn);
        } else {
            px.sn = px.nh;
        }
        px.dotv = dot3(px.sn, px.ld);
        const float dl = fmaxf(px.dotv, 0.f);
        const float* amb = t.per_view_light ? t.light_v[v] : t.ambient;
        const float* dif = t.per_view_light ? t.light_v[v] + 3 : t.diffuse;
        px.tl = v3(dl * dif[0] + amb[0], dl * dif[1] + amb[1], dl * dif[2] + amb[2]);
        const V3 albc = v3(fminf(fmaxf(px.alb.x, 0.f), 1.f), fminf(fmaxf(px.alb.y, 0.f), 1.f),
                           fminf(fmaxf(px.alb.z, 0.f), 1.f));
        const int shading = t.per_view_shading ? (int)t.shading_v[v] : t.shading;
        if (shading == SHADE_ALBEDO) px.fg = px.alb;
        else if (shading == SHADE_TEXTURELESS) px.fg = px.tl;
        else px.fg = v3(albc.x * px.tl.x, albc.y * px.tl.y, albc.z * px.tl.z);
        const V3 bgv = ld3(t.bg + 3 * p);
        px.pre = px.fg * px.alpha + bgv * (1.0f - px.alpha);
    } else if (t.mode == POST_BACKGROUND) {
        px.pre = img + ld3(t.bg + 3 * p) * (1.0f - px.alpha);
    } else {
        px.pre = img;
    }
}

__global__ void __launch_bounds__(256) postprocess_forward_kernel(const __grid_constant__ PostTab t) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y, v = blockIdx.z;
    if (x >= t.W || y >= t.H) return;
    Pixel px;
    pixel_forward(t, v, y, x, px);
    const size_t HW = (size_t)t.H * t.W;
    const size_t p = (size_t)v * HW + (size_t)y * t.W + x;
    const size_t c0 = (size_t)v * 3 * HW + (size_t)y * t.W + x;
    t.render[c0] = fminf(fmaxf(px.pre.x, 0.f), 1.f);
    t.render[c0 + HW] = fminf(fmaxf(px.pre.y, 0.f), 1.f);
    t.render[c0 + 2 * HW] = fminf(fmaxf(px.pre.z, 0.f), 1.f);
    if (t.mode >= POST_NORMAL) {
        const float s = 0.5f * px.alpha;
        t.normal[c0] = px.nh.x * s + 0.5f;
        t.normal[c0 + HW] = px.nh.y * s + 0.5f;
        t.normal[c0 + 2 * HW] = px.nh.z * s + 0.5f;
    }
    if (t.depth_out) t.depth_out[p] = __ldg(t.depth + p);
}

// backward, pass 1: everything that is local to the pixel; the stencil's transpose is left in `scratch`
__global__ void __launch_bounds__(256) postprocess_backward_local_kernel(const __grid_constant__ PostTab t) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y, v = blockIdx.z;
    if (x >= t.W || y >= t.H) return;
    Pixel px;
    pixel_forward(t, v, y, x, px);
    const size_t HW = (size_t)t.H * t.W;
    const size_t p = (size_t)v * HW + (size_t)y * t.W + x;
    const size_t c0 = (size_t)v * 3 * HW + (size_t)y * t.W + x;
    const float al = px.alpha;
    V3 g_pre = v3(0.f, 0.f, 0.f);
    if (t.g_render) {
        const V3 g = v3(__ldg(t.g_render + c0), __ldg(t.g_render + c0 + HW), __ldg(t.g_render + c0 + 2 * HW));
        g_pre = v3((px.pre.x >= 0.f && px.pre.x <= 1.f) ? g.x : 0.f, (px.pre.y >= 0.f && px.pre.y <= 1.f) ? g.y : 0.f,
                   (px.pre.z >= 0.f && px.pre.z <= 1.f) ? g.z : 0.f);
    }
    V3 d_img = g_pre, d_bg = v3(0.f, 0.f, 0.f), g_nh = v3(0.f, 0.f, 0.f), g_xyz = v3(0.f, 0.f, 0.f);
    float d_al = 0.f;
    if (t.mode == POST_BACKGROUND) {
        const V3 bgv = ld3(t.bg + 3 * p);
        d_al = -dot3(g_pre, bgv);
        d_bg = g_pre * (1.0f - al);
    } else if (t.mode == POST_SHADING) {
        const V3 bgv = ld3(t.bg + 3 * p);
        const V3 g_fg = g_pre * al;
        d_al = dot3(g_pre, px.fg - bgv);
        d_bg = g_pre * (1.0f - al);
        V3 g_alb = v3(0.f, 0.f, 0.f), g_tl = v3(0.f, 0.f, 0.f);
        const int shading = t.per_view_shading ? (int)t.shading_v[v] : t.shading;
        if (shading == SHADE_ALBEDO) {
            g_alb = g_fg;
        } else if (shading == SHADE_TEXTURELESS) {
            g_tl = g_fg;
        } else {
            const V3 albc = v3(fminf(fmaxf(px.alb.x, 0.f), 1.f), fminf(fmaxf(px.alb.y, 0.f), 1.f),
                               fminf(fmaxf(px.alb.z, 0.f), 1.f));
            g_alb = v3((px.alb.x >= 0.f && px.alb.x <= 1.f) ? g_fg.x * px.tl.x : 0.f,
                       (px.alb.y >= 0.f && px.alb.y <= 1.f) ? g_fg.y * px.tl.y : 0.f,
                       (px.alb.z >= 0.f && px.alb.z <= 1.f) ? g_fg.z * px.tl.z : 0.f);
            g_tl = v3(g_fg.x * albc.x, g_fg.y * albc.y, g_fg.z * albc.z);
        }
        const float inv = 1.0f / (al + 1e-6f);
        d_img = g_alb * inv;                          // albedo = image / (alpha + 1e-6)
        d_al -= dot3(g_alb, px.alb) * inv;
        const float* dif = t.per_view_light ? t.light_v[v] + 3 : t.diffuse;
        const float g_dot = px.dotv >= 0.f ? g_tl.x * dif[0] + g_tl.y * dif[1] + g_tl.z * dif[2] : 0.f;
        const V3 g_ld = px.sn * g_dot;
        if (!t.pred) g_nh = px.ld * g_dot;             // shading normal = the depth-derived normal
        g_xyz = normalize3_vjp(px.ld, px.vlen, g_ld) * -1.0f;   // light direction = normalize(light - xyz)
    }
    V3 dLda = v3(0.f, 0.f, 0.f), dLdb = v3(0.f, 0.f, 0.f);
    if (t.mode >= POST_NORMAL) {
        if (t.g_normal && al > 0.99f) {                // normal = nh * 0.5 * alpha + 0.5, detached where alpha <= 0.99
            const V3 g = v3(__ldg(t.g_normal + c0), __ldg(t.g_normal + c0 + HW), __ldg(t.g_normal + c0 + 2 * HW));
            g_nh = g_nh + g * (0.5f * al);
            d_al += 0.5f * dot3(g, px.nh);
        }
        const V3 G = normalize3_vjp(px.nh, px.nlen, g_nh);   // dL/dn, n = -(a x b)
        dLda = cross3(px.b, G) * -1.0f;
        dLdb = cross3(G, px.a) * -1.0f;
        float* s = t.scratch + 9 * p;
        s[0] = dLda.x, s[1] = dLda.y, s[2] = dLda.z;
        s[3] = dLdb.x, s[4] = dLdb.y, s[5] = dLdb.z;
        s[6] = g_xyz.x, s[7] = g_xyz.y, s[8] = g_xyz.z;
    } else if (t.d_depth) {
        t.d_depth[p] = t.g_depth ? __ldg(t.g_depth + p) : 0.f;
    }
    t.d_image[c0] = d_img.x, t.d_image[c0 + HW] = d_img.y, t.d_image[c0 + 2 * HW] = d_img.z;
    t.d_alpha[p] = d_al;
    if (t.d_bg) {
        float* o = t.d_bg + 3 * p;
        o[0] = d_bg.x, o[1] = d_bg.y, o[2] = d_bg.z;
    }
}

// backward, pass 2 (modes with a normal map): gather the stencil's transpose -> dL/dxyz -> dL/ddepth
__global__ void __launch_bounds__(256) postprocess_backward_gather_kernel(const __grid_constant__ PostTab t) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y, v = blockIdx.z;
    if (x >= t.W || y >= t.H) return;
    const size_t HW = (size_t)t.H * t.W;
    const size_t vbase = (size_t)v * HW;
    const size_t p = vbase + (size_t)y * t.W + x;
    const float* s = t.scratch + 9 * p;
    V3 gx = ld3(s + 6);
    // xyz(p) enters dx(q) = xyz(q + 1x) - xyz(q - 1x) of its left neighbour with +, of its right neighbour with -
    if (x > 0) gx = gx + ld3(s - 9);
    if (x + 1 < t.W) gx = gx - ld3(s + 9);
    if (y > 0) gx = gx + ld3(s - 9 * (size_t)t.W + 3);
    if (y + 1 < t.H) gx = gx - ld3(s + 9 * (size_t)t.W + 3);
    float d = dot3(gx, ld3(t.rays_d + 3 * p));
    if (t.g_depth && __ldg(t.alpha + p) > 0.99f) d += __ldg(t.g_depth + p);   // depth output detached where alpha <= 0.99
    t.d_depth[p] = d;
}

}  // namespace b200splat

// ---- C ABI ----------------------------------------------------------------------------------------------------
#include "../../include/b200splat.h"
using namespace b200splat;

extern int b200splat_set_error(int code, const char* msg);

static int fill_tab(const b200splat_postprocess_args* a, PostTab* t) {
    if (!a) return b200splat_set_error(B200SPLAT_ERR_INVALID, "null args");
    if (a->V < 1 || a->H < 1 || a->W < 1) return b200splat_set_error(B200SPLAT_ERR_INVALID, "V, H, W must be positive");
    if (a->mode < POST_PLAIN || a->mode > POST_SHADING) return b200splat_set_error(B200SPLAT_ERR_INVALID, "unknown mode");
    if (a->shading < SHADE_ALBEDO || a->shading > SHADE_DIFFUSE)
        return b200splat_set_error(B200SPLAT_ERR_INVALID, "unknown shading");
    if (!a->image || !a->depth || !a->alpha) return b200splat_set_error(B200SPLAT_ERR_INVALID, "null rasterizer output");
    if (a->mode >= POST_NORMAL && (!a->rays_o || !a->rays_d))
        return b200splat_set_error(B200SPLAT_ERR_INVALID, "normal / shading modes need rays_o and rays_d");
    if ((a->mode == POST_BACKGROUND || a->mode == POST_SHADING) && !a->bg)
        return b200splat_set_error(B200SPLAT_ERR_INVALID, "background / shading modes need bg");
    if (a->mode == POST_SHADING && !a->light) return b200splat_set_error(B200SPLAT_ERR_INVALID, "shading mode needs light");
    memset(t, 0, sizeof(*t));
    t->V = a->V, t->H = a->H, t->W = a->W, t->mode = a->mode, t->shading = a->shading;
    t->image = a->image, t->depth = a->depth, t->alpha = a->alpha, t->rays_o = a->rays_o, t->rays_d = a->rays_d;
    t->bg = a->bg, t->light = a->light, t->pred = a->mode == POST_SHADING ? a->pred_normal : nullptr;
    for (int c = 0; c < 3; ++c) t->ambient[c] = a->ambient[c], t->diffuse[c] = a->diffuse[c];
    if (a->lights_per_view && a->mode == POST_SHADING) {
        if (a->V > POST_MAX_PER_VIEW)
            return b200splat_set_error(B200SPLAT_ERR_INVALID, "lights_per_view supports at most 64 views per call");
        t->per_view_light = 1;
        for (int v = 0; v < a->V; ++v)
            for (int c = 0; c < 6; ++c) t->light_v[v][c] = a->lights_per_view[6 * v + c];
    }
    if (a->shading_per_view && a->mode == POST_SHADING) {
        if (a->V > POST_MAX_PER_VIEW)
            return b200splat_set_error(B200SPLAT_ERR_INVALID, "shading_per_view supports at most 64 views per call");
        t->per_view_shading = 1;
        for (int v = 0; v < a->V; ++v) {
            if (a->shading_per_view[v] < SHADE_ALBEDO || a->shading_per_view[v] > SHADE_DIFFUSE)
                return b200splat_set_error(B200SPLAT_ERR_INVALID, "unknown per-view shading");
            t->shading_v[v] = (uint8_t)a->shading_per_view[v];
        }
    }
    return B200SPLAT_OK;
}

static dim3 post_grid(const PostTab& t) { return dim3((t.W + 31) / 32, (t.H + 7) / 8, t.V); }

extern "C" {

size_t b200splat_postprocess_scratch_bytes(int32_t V, int32_t H, int32_t W) {
    return (size_t)(V > 0 ? V : 0) * (H > 0 ? H : 0) * (W > 0 ? W : 0) * 9 * sizeof(float);
}

int b200splat_postprocess_forward(const b200splat_postprocess_args* a) {
    PostTab t;
    int rc = fill_tab(a, &t);
    if (rc) return rc;
    if (!a->render || (a->mode >= POST_NORMAL && !a->normal))
        return b200splat_set_error(B200SPLAT_ERR_INVALID, "null output");
    t.render = a->render, t.normal = a->normal, t.depth_out = a->depth_out;
    postprocess_forward_kernel<<<post_grid(t), dim3(32, 8), 0, reinterpret_cast<cudaStream_t>(a->stream)>>>(t);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return b200splat_set_error(B200SPLAT_ERR_CUDA, cudaGetErrorString(e));
    return B200SPLAT_OK;
}

int b200splat_postprocess_backward(const b200splat_postprocess_args* a) {
    PostTab t;
    int rc = fill_tab(a, &t);
    if (rc) return rc;
    if (!a->d_image || !a->d_alpha || !a->d_depth) return b200splat_set_error(B200SPLAT_ERR_INVALID, "null gradient output");
    if (a->mode >= POST_NORMAL &&
        (!a->scratch || a->scratch_bytes < b200splat_postprocess_scratch_bytes(a->V, a->H, a->W)))
        return b200splat_set_error(B200SPLAT_ERR_NOMEM, "postprocess scratch too small");
    t.g_render = a->g_render, t.g_normal = a->g_normal, t.g_depth = a->g_depth;
    t.d_image = a->d_image, t.d_depth = a->d_depth, t.d_alpha = a->d_alpha, t.d_bg = a->d_bg;
    t.scratch = reinterpret_cast<float*>(a->scratch);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(a->stream);
    postprocess_backward_local_kernel<<<post_grid(t), dim3(32, 8), 0, st>>>(t);
    count_launch();
    if (a->mode >= POST_NORMAL) {
        postprocess_backward_gather_kernel<<<post_grid(t), dim3(32, 8), 0, st>>>(t);
        count_launch();
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return b200splat_set_error(B200SPLAT_ERR_CUDA, cudaGetErrorString(e));
    return B200SPLAT_OK;
}

}  // extern "C"

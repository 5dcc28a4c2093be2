// Extra per-Gaussian feature channels rendered by the SAME raster pass as the colour.
//
// The reference's normal / shading variants rasterize twice per view: once for the colour and once more with
// the per-Gaussian normals in place of the colour (renderer/diff_gaussian_rasterizer_shading.py:177-187,
// renderer/diff_gaussian_rasterizer_normal.py:175-185) -- a second preprocess, sort and blend of the same
// Gaussians with the same alphas.  Here up to 4 extra channels ride along the colour pass: the render
// kernels stage one more 16-byte record per list entry and blend it with the alpha they already computed
// (render.cu, EXT = true), so the extra image costs one FMA per channel per blended pair instead of a pass.
//
//   pad_extra_kernel       (P, n_extra) fp32 -> [P] float4 records (the 16-byte granule the render kernels gather)
//   extra_backward_kernel  per Gaussian: sum the views' atomically accumulated record gradients -> dL/dextra
//
// Both are HBM streaming kernels (16 + 4 n_extra bytes resp. 16 V + 4 n_extra bytes per Gaussian).
#include "common.cuh"

namespace b200splat {

size_t grad2d_bytes(int P) { return align_up((size_t)(P > 0 ? P : 1) * GRAD2D_FLOATS * sizeof(float), 256); }

__global__ void __launch_bounds__(256)
pad_extra_kernel(int P, int n_extra, const float* __restrict__ extra, float4* __restrict__ ext4) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    float e[EXT_FLOATS] = {0.f, 0.f, 0.f, 0.f};
    const float* src = extra + (size_t)i * n_extra;
#pragma unroll
    for (int c = 0; c < EXT_FLOATS; ++c)
        if (c < n_extra) e[c] = __ldg(src + c);
    ext4[i] = make_float4(e[0], e[1], e[2], e[3]);
}

// flat over the 4 P floats of the padded records: thread e sums element e of every view's record array (and zeroes it
// again under the self-cleaning protocol); fully coalesced 4-byte accesses
__global__ void __launch_bounds__(256)
extra_backward_kernel(const __grid_constant__ BatchTab tab, float* __restrict__ dL_dextra, int accumulate, int g_begin,
                      int g_end) {
    const size_t e0 = (size_t)g_begin * EXT_FLOATS, e1 = (size_t)g_end * EXT_FLOATS;
    const size_t e = e0 + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= e1) return;
    float s = 0.f;
    for (int v = 0; v < tab.V; ++v) s += tab.v[v].gradext[e];
    if (tab.clean_scratch) {
        // The zero that is stored is derived from the loaded sum (bits(s) & 0, opaque to the compiler), so the stores
        // issue only after the loads have COMPLETED.  A store issued right behind a load of the same address that is
        // still in flight is serialised by the memory system: measured on B200 (scripts/ubench/readback.cu) 237 us
        // instead of 30 us for this very access pattern (4 x 16 MB).
        float z;
        asm volatile("and.b32 %0, %1, 0;" : "=f"(z) : "f"(s));
        for (int v = 0; v < tab.V; ++v) tab.v[v].gradext[e] = z;
    }
    const int c = (int)(e & (EXT_FLOATS - 1));
    if (c < tab.n_extra) {
        float* dst = dL_dextra + (e >> 2) * tab.n_extra + c;
        *dst = accumulate ? *dst + s : s;
    }
}

cudaError_t launch_pad_extra(int P, int n_extra, const float* extra, float4* ext4, cudaStream_t st) {
    if (P <= 0 || n_extra <= 0) return cudaSuccess;
    pad_extra_kernel<<<(P + 255) / 256, 256, 0, st>>>(P, n_extra, extra, ext4);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_extra_backward(const BatchTab& tab, float* dL_dextra, int accumulate, cudaStream_t st, int g_begin,
                                  int g_end) {
    if (tab.n_extra <= 0 || !dL_dextra) return cudaSuccess;
    if (g_end <= 0 || g_end > tab.P) g_end = tab.P;
    if (g_begin < 0) g_begin = 0;
    if (g_begin >= g_end) return cudaSuccess;
    const size_t n = (size_t)(g_end - g_begin) * EXT_FLOATS;
    extra_backward_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(tab, dL_dextra, accumulate, g_begin, g_end);
    count_launch();
    return cudaGetLastError();
}

}  // namespace b200splat

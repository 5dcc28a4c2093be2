// K10 distCUDA2: mean squared distance to the 3 nearest other points (exact).
//
// Replaces simple_knn._C.distCUDA2 (DSaurus/simple-knn simple_knn.cu, [UPSTREAM-RECALL]); reference call
// site geometry/gaussian_base.py:434-437 (P = 4096 random points at init in every shipped config) and
// system/gaussian_splatting.py:214-223 (P identical points on checkpoint load).
//
// Upstream's Morton-box search is exact (boxes only prune), so any exact 3-NN gives the same result.
// This version: queries are sorted into a uniform grid (cell = one Morton-free linear id, our own
// onesweep sort), each thread answers one query by scanning cells in growing Chebyshev rings until the
// ring's distance lower bound exceeds its current 3rd-best.  Degenerate inputs (all points identical ->
// one cell) fall back to the exhaustive scan inside that cell, like upstream's box scan.
#include "common.cuh"
#include <float.h>

namespace b200splat {

__device__ __forceinline__ void insert3(float d, float& b0, float& b1, float& b2) {
    if (d < b2) {
        if (d < b1) {
            b2 = b1;
            if (d < b0) {
                b1 = b0;
                b0 = d;
            } else {
                b1 = d;
            }
        } else {
            b2 = d;
        }
    }
}

// ---- small P: exhaustive, shared-memory tiled ---------------------------------------------------
constexpr int KNN_TILE = 1024;
__global__ void __launch_bounds__(256)
dist2_bruteforce_kernel(int P, const float* __restrict__ pts, float* __restrict__ out) {
    __shared__ float sx[KNN_TILE], sy[KNN_TILE], sz[KNN_TILE];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (i < P) qx = pts[3 * i], qy = pts[3 * i + 1], qz = pts[3 * i + 2];
    float b0 = FLT_MAX, b1 = FLT_MAX, b2 = FLT_MAX;
    for (int base = 0; base < P; base += KNN_TILE) {
        const int cnt = min(KNN_TILE, P - base);
        __syncthreads();
        for (int t = threadIdx.x; t < cnt; t += blockDim.x) {
            sx[t] = pts[3 * (base + t)], sy[t] = pts[3 * (base + t) + 1], sz[t] = pts[3 * (base + t) + 2];
        }
        __syncthreads();
        if (i < P) {
            for (int t = 0; t < cnt; ++t) {
                if (base + t == i) continue;
                const float dx = qx - sx[t], dy = qy - sy[t], dz = qz - sz[t];
                insert3(dx * dx + dy * dy + dz * dz, b0, b1, b2);
            }
        }
    }
    if (i < P) {
        float s = 0.f;
        if (b0 < FLT_MAX) s += b0;
        if (b1 < FLT_MAX) s += b1;
        if (b2 < FLT_MAX) s += b2;
        out[i] = s / 3.0f;
    }
}

// ---- large P: uniform grid ------------------------------------------------------------------------
struct GridParams {
    float minx, miny, minz;
    float inv_cell;
    int nx, ny, nz;
};

__global__ void bbox_kernel(int P, const float* __restrict__ pts, float* __restrict__ bbox /* min3,max3 as ordered ints */) {
    // ordered-int trick for float atomicMin/Max
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (; i < P; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float v = pts[3 * i + k];
            mn[k] = fminf(mn[k], v);
            mx[k] = fmaxf(mx[k], v);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
            mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
        int* b = reinterpret_cast<int*>(bbox);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            int a = __float_as_int(mn[k]);
            a = a >= 0 ? a : a ^ 0x7fffffff;
            atomicMin(b + k, a);
            int c = __float_as_int(mx[k]);
            c = c >= 0 ? c : c ^ 0x7fffffff;
            atomicMax(b + 3 + k, c);
        }
    }
}

__device__ __forceinline__ float ordered_to_float(int a) { return __int_as_float(a >= 0 ? a : a ^ 0x7fffffff); }

__global__ void grid_params_kernel(int P, int64_t max_cells, const float* bbox, GridParams* gp) {
    const int* b = reinterpret_cast<const int*>(bbox);
    float mn[3], mx[3];
    for (int k = 0; k < 3; ++k) mn[k] = ordered_to_float(b[k]), mx[k] = ordered_to_float(b[3 + k]);
    const float ex = fmaxf(mx[0] - mn[0], 1e-12f), ey = fmaxf(mx[1] - mn[1], 1e-12f), ez = fmaxf(mx[2] - mn[2], 1e-12f);
    const float emax = fmaxf(ex, fmaxf(ey, ez));
    // ~2 points per cell on average; at most 1024 cells per axis and max_cells in total
    float cell = fmaxf(cbrtf(ex * ey * ez * 2.0f / (float)P), emax / 1000.0f);
    int nx, ny, nz;
    for (int it = 0; it < 64; ++it) {
        nx = max(1, min(1024, (int)(ex / cell) + 1));
        ny = max(1, min(1024, (int)(ey / cell) + 1));
        nz = max(1, min(1024, (int)(ez / cell) + 1));
        if ((int64_t)nx * ny * nz <= max_cells) break;
        cell *= 1.25f;
    }
    gp->minx = mn[0], gp->miny = mn[1], gp->minz = mn[2];
    gp->inv_cell = 1.0f / cell;
    gp->nx = nx, gp->ny = ny, gp->nz = nz;
}

__device__ __forceinline__ void cell_of(const GridParams& g, float x, float y, float z, int& cx, int& cy, int& cz) {
    cx = min(g.nx - 1, max(0, (int)((x - g.minx) * g.inv_cell)));
    cy = min(g.ny - 1, max(0, (int)((y - g.miny) * g.inv_cell)));
    cz = min(g.nz - 1, max(0, (int)((z - g.minz) * g.inv_cell)));
}

__global__ void cell_keys_kernel(int P, const float* __restrict__ pts, const GridParams* gp, uint64_t* keys,
                                 uint32_t* vals) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const GridParams g = *gp;
    int cx, cy, cz;
    cell_of(g, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], cx, cy, cz);
    keys[i] = (uint64_t)(((uint32_t)cz * g.ny + cy) * g.nx + cx);
    vals[i] = (uint32_t)i;
}

// cell_start[c] = first sorted position of cell c, cell_start[ncell] = P (filled by lower-bound search)
__global__ void cell_start_kernel(int P, const uint64_t* __restrict__ keys_sorted, const GridParams* gp,
                                  uint32_t* __restrict__ cell_start) {
    const GridParams g = *gp;
    const int ncell = g.nx * g.ny * g.nz;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > ncell) return;
    int lo = 0, hi = P;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (keys_sorted[mid] < (uint64_t)c) lo = mid + 1; else hi = mid;
    }
    cell_start[c] = (uint32_t)lo;
}

__global__ void gather_sorted_kernel(int P, const float* __restrict__ pts, const uint32_t* __restrict__ order,
                                     float4* __restrict__ sorted) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const uint32_t o = order[i];
    sorted[i] = make_float4(pts[3 * o], pts[3 * o + 1], pts[3 * o + 2], __uint_as_float(o));
}

__global__ void __launch_bounds__(128)
dist2_grid_kernel(int P, const float4* __restrict__ sorted, const uint32_t* __restrict__ cell_start,
                  const GridParams* gp, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const GridParams g = *gp;
    const float4 q = sorted[i];
    int cx, cy, cz;
    cell_of(g, q.x, q.y, q.z, cx, cy, cz);
    const float cell = 1.0f / g.inv_cell;
    float b0 = FLT_MAX, b1 = FLT_MAX, b2 = FLT_MAX;
    const int rmax = max(g.nx, max(g.ny, g.nz));
    for (int ring = 0; ring <= rmax; ++ring) {
        // every unvisited point is at least (ring-1)*cell + (distance of q to its own cell's faces) away;
        // use the conservative bound (ring-1)*cell
        if (ring >= 2) {
            const float lb = (float)(ring - 1) * cell;
            if (b2 < FLT_MAX && lb * lb > b2) break;
        }
        const int z0 = max(0, cz - ring), z1 = min(g.nz - 1, cz + ring);
        const int y0 = max(0, cy - ring), y1 = min(g.ny - 1, cy + ring);
        const int x0 = max(0, cx - ring), x1 = min(g.nx - 1, cx + ring);
        for (int zz = z0; zz <= z1; ++zz) {
            for (int yy = y0; yy <= y1; ++yy) {
                const bool shell_zy = (abs(zz - cz) == ring) || (abs(yy - cy) == ring);
                if (shell_zy) {
                    // whole x-run [x0,x1] is on the shell: cells are contiguous in the sorted order
                    const uint32_t cbase = ((uint32_t)zz * g.ny + yy) * g.nx;
                    const uint32_t s = cell_start[cbase + x0], e = cell_start[cbase + x1 + 1];
                    for (uint32_t t = s; t < e; ++t) {
                        if ((int)t == i) continue;
                        const float4 p = sorted[t];
                        const float dx = q.x - p.x, dy = q.y - p.y, dz = q.z - p.z;
                        insert3(dx * dx + dy * dy + dz * dz, b0, b1, b2);
                    }
                } else {
                    // only the two end cells x = cx-ring, cx+ring
                    for (int side = 0; side < 2; ++side) {
                        const int xx = side ? cx + ring : cx - ring;
                        if (xx < 0 || xx >= g.nx || (side == 1 && ring == 0)) continue;
                        const uint32_t cid = ((uint32_t)zz * g.ny + yy) * g.nx + xx;
                        const uint32_t s = cell_start[cid], e = cell_start[cid + 1];
                        for (uint32_t t = s; t < e; ++t) {
                            if ((int)t == i) continue;
                            const float4 p = sorted[t];
                            const float dx = q.x - p.x, dy = q.y - p.y, dz = q.z - p.z;
                            insert3(dx * dx + dy * dy + dz * dz, b0, b1, b2);
                        }
                    }
                }
            }
        }
    }
    float s = 0.f;
    if (b0 < FLT_MAX) s += b0;
    if (b1 < FLT_MAX) s += b1;
    if (b2 < FLT_MAX) s += b2;
    out[__float_as_uint(q.w)] = s / 3.0f;
}

constexpr int KNN_BRUTE_MAX = 16384;

static inline int64_t knn_cell_bound(int P) { return (int64_t)P * 2 + 4096; }

// workspace: bbox (32 B) | GridParams (64 B) | keys[2] | vals[2] | sorted float4 | cell_start | sort ws
struct KnnLayout {
    float* bbox;
    GridParams* gp;
    uint64_t* keys[2];
    uint32_t* vals[2];
    float4* sorted;
    uint32_t* cell_start;
    void* sort_ws;
    size_t total;
};
static KnnLayout knn_layout(int P, void* base) {
    KnnLayout L;
    char* p = reinterpret_cast<char*>(base);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char* r = p + off;
        off = align_up(off + bytes, 256);
        return r;
    };
    L.bbox = reinterpret_cast<float*>(take(32));
    L.gp = reinterpret_cast<GridParams*>(take(64));
    L.keys[0] = reinterpret_cast<uint64_t*>(take((size_t)P * 8));
    L.keys[1] = reinterpret_cast<uint64_t*>(take((size_t)P * 8));
    L.vals[0] = reinterpret_cast<uint32_t*>(take((size_t)P * 4));
    L.vals[1] = reinterpret_cast<uint32_t*>(take((size_t)P * 4));
    L.sorted = reinterpret_cast<float4*>(take((size_t)P * 16));
    L.cell_start = reinterpret_cast<uint32_t*>(take(((size_t)knn_cell_bound(P) + 2) * 4));
    L.sort_ws = take(sort_workspace_bytes(P));
    L.total = off;
    return L;
}

size_t dist2_workspace_bytes(int P) {
    if (P <= KNN_BRUTE_MAX) return 256;
    return knn_layout(P, nullptr).total;
}

cudaError_t launch_dist2(int P, const float* points, float* out, void* ws, cudaStream_t st) {
    if (P <= 0) return cudaSuccess;
    if (P <= KNN_BRUTE_MAX) {
        dist2_bruteforce_kernel<<<(P + 255) / 256, 256, 0, st>>>(P, points, out);
        count_launch();
        return cudaGetLastError();
    }
    KnnLayout L = knn_layout(P, ws);
    const int64_t bound = knn_cell_bound(P);
    // bbox init: min slots = +inf (ordered int), max slots = -inf
    const int init[8] = {0x7f7fffff, 0x7f7fffff, 0x7f7fffff, (int)0x80800000, (int)0x80800000, (int)0x80800000, 0, 0};
    cudaError_t e = cudaMemcpyAsync(L.bbox, init, sizeof(init), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return e;
    const int g = (P + 255) / 256;
    bbox_kernel<<<min(g, NUM_SMS * 8), 256, 0, st>>>(P, points, L.bbox);
    grid_params_kernel<<<1, 1, 0, st>>>(P, bound, L.bbox, L.gp);
    cell_keys_kernel<<<g, 256, 0, st>>>(P, points, L.gp, L.keys[0], L.vals[0]);
    count_launch(3);
    int sel = 0;
    e = launch_sort_pairs(P, 32, L.keys, L.vals, L.sort_ws, &sel, st);
    if (e != cudaSuccess) return e;
    cell_start_kernel<<<(unsigned)((bound + 1 + 255) / 256), 256, 0, st>>>(P, L.keys[sel], L.gp, L.cell_start);
    gather_sorted_kernel<<<g, 256, 0, st>>>(P, points, L.vals[sel], L.sorted);
    dist2_grid_kernel<<<(P + 127) / 128, 128, 0, st>>>(P, L.sorted, L.cell_start, L.gp, out);
    count_launch(3);
    return cudaGetLastError();
}

}  // namespace b200splat

// Micro-benchmark: why is "sum V per-view arrays and zero them" slow?  Variants isolate (a) zeroing stores,
// (b) arrays last written by RED atomics vs plain stores, (c) base-address spacing of the V arrays.
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
struct Tab { float* a[4]; };
template <bool ZERO>
__global__ void sum4(Tab t, float* __restrict__ dst, size_t n) {
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    float s = 0.f;
#pragma unroll
    for (int v = 0; v < 4; ++v) { s += t.a[v][e]; if (ZERO) t.a[v][e] = 0.f; }
    dst[e] = s;
}
// variants of the zeroing: 1 = store right behind each load (as above), 2 = all loads first, stores after the sum
// has consumed them, 3 = stores depend on the loaded value (x - x), 4 = zero through a different thread (e ^ 32),
// 5 = st.global.cs, 6 = loads with ld.global.cv (no L1 allocation)
template <int MODE>
__global__ void sum4v(Tab t, float* __restrict__ dst, size_t n) {
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    float x[4];
    float s = 0.f;
    if (MODE == 6) {
#pragma unroll
        for (int v = 0; v < 4; ++v) { asm volatile("ld.global.cv.f32 %0, [%1];" : "=f"(x[v]) : "l"(t.a[v] + e)); t.a[v][e] = 0.f; }
    } else {
#pragma unroll
        for (int v = 0; v < 4; ++v) x[v] = t.a[v][e];
    }
#pragma unroll
    for (int v = 0; v < 4; ++v) s += x[v];
    if (MODE == 2) {
        dst[e] = s;
        if (s != 12345.f) {
#pragma unroll
            for (int v = 0; v < 4; ++v) t.a[v][e] = 0.f;
        }
        return;
    }
    if (MODE == 3) {
#pragma unroll
        for (int v = 0; v < 4; ++v) t.a[v][e] = x[v] - x[v];
    }
    if (MODE == 4) {
#pragma unroll
        for (int v = 0; v < 4; ++v) t.a[v][e ^ 32] = 0.f;
    }
    if (MODE == 5) {
#pragma unroll
        for (int v = 0; v < 4; ++v) asm volatile("st.global.cs.f32 [%0], %1;" :: "l"(t.a[v] + e), "f"(0.f));
    }
    dst[e] = s;
}
__global__ void red_fill(Tab t, size_t n, int reps) {   // scattered RED.ADD like render backward
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int r = 0; r < reps; ++r) {
        size_t e = (i * 2654435761ull + r * 40503ull) % n;
        for (int v = 0; v < 4; ++v) atomicAdd(t.a[v] + e, 1.0f);
    }
}
__global__ void plain_fill(Tab t, size_t n) {
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) for (int v = 0; v < 4; ++v) t.a[v][e] = 1.0f;
}
int main() {
    const size_t n = 4u << 20;   // 4 M floats = 16 MB per array, as gradext at P = 1 M
    float* pool; CK(cudaMalloc(&pool, 6 * (n * 4 + (64u << 20))));
    float* dst; CK(cudaMalloc(&dst, n * 4));
    char* big; CK(cudaMalloc(&big, 256u << 20));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const unsigned grid = (unsigned)((n + 255) / 256);
    const size_t spacings[3] = {n * 4, n * 4 + 64000256 - 16000000, n * 4 + 4352};   // packed | as in the workspace | odd
    for (int sp = 0; sp < 3; ++sp) {
        Tab t; for (int v = 0; v < 4; ++v) t.a[v] = (float*)((char*)pool + v * spacings[sp]);
        for (int fill = 0; fill < 2; ++fill) for (int zero = 0; zero < 2; ++zero) {
            float best = 1e9f;
            for (int it = 0; it < 5; ++it) {
                if (fill) { for (int v = 0; v < 4; ++v) cudaMemsetAsync(t.a[v], 0, n * 4); red_fill<<<grid / 4, 256>>>(t, n, 8); }
                else plain_fill<<<grid, 256>>>(t, n);
                cudaMemsetAsync(big, 1, 256u << 20);   // push the arrays out of L2
                cudaEventRecord(e0);
                if (zero) sum4<true><<<grid, 256>>>(t, dst, n); else sum4<false><<<grid, 256>>>(t, dst, n);
                cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
                float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
            }
            printf("spacing %zu B  last writer %-5s  zeroing %d : %.1f us  (%.0f GB/s read)\n", spacings[sp], fill ? "RED" : "store",
                   zero, best * 1e3, 4.0 * n * 4 / (best * 1e-3) / 1e9);
        }
    }
    {
        Tab t; for (int v = 0; v < 4; ++v) t.a[v] = (float*)((char*)pool + v * spacings[1]);
        for (int mode = 2; mode <= 6; ++mode) {
            float best = 1e9f;
            for (int it = 0; it < 5; ++it) {
                plain_fill<<<grid, 256>>>(t, n);
                cudaMemsetAsync(big, 1, 256u << 20);
                cudaEventRecord(e0);
                switch (mode) {
                    case 2: sum4v<2><<<grid, 256>>>(t, dst, n); break;
                    case 3: sum4v<3><<<grid, 256>>>(t, dst, n); break;
                    case 4: sum4v<4><<<grid, 256>>>(t, dst, n); break;
                    case 5: sum4v<5><<<grid, 256>>>(t, dst, n); break;
                    case 6: sum4v<6><<<grid, 256>>>(t, dst, n); break;
                }
                cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
                float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
            }
            printf("mode %d : %.1f us\n", mode, best * 1e3);
        }
        // zeroing by a separate memset pass
        float best = 1e9f;
        for (int it = 0; it < 5; ++it) {
            plain_fill<<<grid, 256>>>(t, n);
            cudaMemsetAsync(big, 1, 256u << 20);
            cudaEventRecord(e0);
            sum4<false><<<grid, 256>>>(t, dst, n);
            for (int v = 0; v < 4; ++v) cudaMemsetAsync(t.a[v], 0, n * 4);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
        }
        printf("sum then 4 memsets : %.1f us\n", best * 1e3);
    }
    return 0;
}

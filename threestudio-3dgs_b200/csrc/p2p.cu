// All-reduce of a rank's packed gradient buffer over NVLink peer memory, one kernel per rank.
//
// The view-data-parallel step (one process per GPU, Gaussians replicated) ends with SUM(packed gradients +
// densification accumulators) and MAX(max_radii) over the ranks (b200splat/dist.py).  Every rank's buffer lives
// in cudaMalloc memory exported with CUDA IPC and mapped by all peers, so one kernel per rank does the whole
// exchange with plain loads / stores through NVSwitch:
//
//   phase 0  rank r tells every peer "my buffer is complete" (release store of the epoch into the peer's signal
//            row) and waits for the same from every peer;
//   reduce   r owns the r-th slice of the float4 index space: it loads that slice from ALL ranks (its own
//            included) in rank order 0..N-1 -- so every rank ends up with bit-identical sums -- and stores the
//            result into ALL ranks' buffers.  Only r ever touches slice r, anywhere, so the exchange is in place;
//            inbound (peer loads) and outbound (peer stores) traffic run in opposite NVLink directions at once:
//            (N-1)/N of the buffer each way, against 2 (N-1)/N each way for a ring all-reduce;
//   phase 1  the last CTA to finish tells every peer "my stores have landed" (after a system fence) and waits for
//            every peer's: when the kernel completes the local buffer is fully reduced and no peer reads it any more.
//
// Spins are bounded (about 4 s of SM clock): a dead peer turns into an error flag, not a hung GPU.
#include "common.cuh"
#include <cstdlib>

namespace b200splat {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_cg4(const float4* p) {
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_cg4(float4* p, float4 v) {
    asm volatile("st.global.cg.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

struct float8 {
    float v[8];
};
// 256-bit global accesses (LDG/STG.E.ENL2.256, sm_100): half as many NVLink requests as 128-bit ones
__device__ __forceinline__ float8 ld_cg8(const float* p) {
    float8 r;
    asm volatile("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]),
                   "=f"(r.v[7])
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_cg8(float* p, const float8& r) {
    asm volatile("st.global.cg.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]),
                 "f"(r.v[3]), "f"(r.v[4]), "f"(r.v[5]), "f"(r.v[6]), "f"(r.v[7])
                 : "memory");
}

constexpr long long P2P_SPIN_LIMIT = 8000000000ll;   // SM clocks

// The call's epoch.  t.epoch != 0: given by the host (must increase by 1 per call).  t.epoch == 0: the rank keeps its
// own call counter in its signal row (word P2P_EPOCH_WORD), read here by every CTA and advanced by the last CTA of the
// kernel -- nothing in the launch arguments changes from call to call, so the exchange can be captured in a CUDA graph
// and replayed (launches of one rank are stream-ordered; every rank makes the same sequence of calls).
__device__ __forceinline__ uint32_t call_epoch(const P2PTab& t) {
    return t.epoch ? t.epoch : ld_acquire_sys(t.signals[t.rank] + P2P_EPOCH_WORD) + 1u;
}

// wait until every peer's word in my signal row `row` reached `epoch`; returns false on timeout
__device__ __forceinline__ bool wait_row(const P2PTab& t, int row, int tid, uint32_t epoch) {
    bool ok = true;
    if (tid < t.world) {
        const uint32_t* flag = t.signals[t.rank] + row * P2P_MAX_RANKS + tid;
        const long long t0 = clock64();
        while ((int32_t)(ld_acquire_sys(flag) - epoch) < 0) {
            if (clock64() - t0 > P2P_SPIN_LIMIT) {
                ok = false;
                break;
            }
        }
        if (!ok) atomicExch(t.signals[t.rank] + P2P_ERROR_WORD, 1u);
    }
    return ok;
}

template <bool IS_MAX, int U>
__device__ __forceinline__ void reduce_slice(const P2PTab& t, int64_t first4, int64_t n4, int mode) {
    // float4 elements [first4, first4 + n4) of every rank's buffer; this rank's share of them
    const int64_t lo = first4 + n4 * t.rank / t.world, hi = first4 + n4 * (t.rank + 1) / t.world;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i0 = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < hi; i0 += U * stride) {
        float4 acc[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + u * stride;
            acc[u] = i < hi ? ld_cg4(reinterpret_cast<const float4*>(t.bufs[0]) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int k = 1; k < t.world && mode != 2; ++k) {
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t i = i0 + u * stride;
                v[u] = i < hi ? ld_cg4(reinterpret_cast<const float4*>(t.bufs[k]) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (IS_MAX) {
                    acc[u].x = fmaxf(acc[u].x, v[u].x), acc[u].y = fmaxf(acc[u].y, v[u].y);
                    acc[u].z = fmaxf(acc[u].z, v[u].z), acc[u].w = fmaxf(acc[u].w, v[u].w);
                } else {
                    acc[u].x += v[u].x, acc[u].y += v[u].y, acc[u].z += v[u].z, acc[u].w += v[u].w;
                }
            }
        }
        for (int k = 0; k < (mode == 1 ? 1 : t.world); ++k) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t i = i0 + u * stride;
                if (i < hi) st_cg4(reinterpret_cast<float4*>(t.bufs[k]) + i, acc[u]);
            }
        }
    }
}

// the same over 32-byte elements (first8 / n8 in units of 8 floats)
template <bool IS_MAX, int U>
__device__ __forceinline__ void reduce_slice8(const P2PTab& t, int64_t first8, int64_t n8) {
    const int64_t lo = first8 + n8 * t.rank / t.world, hi = first8 + n8 * (t.rank + 1) / t.world;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i0 = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < hi; i0 += U * stride) {
        float8 acc[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < hi) acc[u] = ld_cg8(t.bufs[0] + 8 * i);
        }
        for (int k = 1; k < t.world; ++k) {
            float8 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t i = i0 + u * stride;
                if (i < hi) v[u] = ld_cg8(t.bufs[k] + 8 * i);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (i0 + u * stride < hi) {
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        acc[u].v[c] = IS_MAX ? fmaxf(acc[u].v[c], v[u].v[c]) : acc[u].v[c] + v[u].v[c];
                }
            }
        }
        for (int k = 0; k < t.world; ++k) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t i = i0 + u * stride;
                if (i < hi) st_cg8(t.bufs[k] + 8 * i, acc[u]);
            }
        }
    }
}

// ---- the same exchange through the NVSwitch multicast address of the buffers ---------------------------------------
// multimem.ld_reduce: one load whose value is the switch's reduction of the word at that offset in every bound buffer;
// multimem.st: one store the switch replicates into every bound buffer.
__device__ __forceinline__ float4 mc_ld_add4(const float* p) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p)
                 : "memory");
    return v;
}
__device__ __forceinline__ void mc_st4(float* p, const float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ uint32_t mc_ld_max_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.max.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void mc_st_u32(uint32_t* p, uint32_t v) {
    asm volatile("multimem.st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ uint32_t mc_ld_add_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_cg_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// Row-sparse SUM segment: rows of `row4` float4s, row j = live-map index row0 + j.  Most Gaussians of a dense scene
// carry no gradient on any rank (preprocess backward's scan knows, and writes one byte per Gaussian into the live map
// next to the packed buffer): their rows are zero everywhere and stay untouched.  Rank r owns the r-th slice of the
// ROWS; a CTA takes a contiguous run of them in pieces of SP_ROWS rows: (1) union of the ranks' live bytes (16-byte
// loads: 16 rows each), compacted into a shared-memory list; (2) the float4s of the listed rows are loaded from all
// ranks in rank order, summed and stored to all ranks, as in reduce_slice.  MC: the union is ONE
// multimem.ld_reduce.add.u32 per four rows (the switch adds the ranks' bytes; at most 8 ranks, no carry).
constexpr int SP_ROWS = 2048;
template <bool MC>
__device__ __forceinline__ void reduce_rows(const P2PTab& t, float* mc, int64_t first4, int64_t n4, int row4,
                                            int64_t row0, uint16_t* s_list, uint32_t* s_count) {
    const int64_t rows = n4 / row4;
    const int64_t lo = rows * t.rank / t.world / 16 * 16;                              // slices on 16-row boundaries
    const int64_t hi = t.rank + 1 == t.world ? rows : rows * (t.rank + 1) / t.world / 16 * 16;
    const int64_t per_cta = ((hi - lo + gridDim.x - 1) / gridDim.x + 15) / 16 * 16;
    const int64_t c0 = lo + (int64_t)blockIdx.x * per_cta, c1 = min(hi, c0 + per_cta);
    const int tid = threadIdx.x, lane = tid & 31;
    for (int64_t p0 = c0; p0 < c1; p0 += SP_ROWS) {
        const int n_rows = (int)min((int64_t)SP_ROWS, c1 - p0);
        if (tid == 0) *s_count = 0;
        __syncthreads();
        // (1) union of the live bytes, 4 rows per thread and step
        for (int qb = (tid & ~31) * 4; qb < n_rows; qb += blockDim.x * 4) {   // warp-uniform bound (shuffles below)
            const int q = qb + lane * 4;
            const int64_t idx = row0 + p0 + q;     // multiple of 4 (row0, lo, per_cta, SP_ROWS, q all are)
            uint32_t u = 0;
            if (q < n_rows) {
                if (MC) {
                    u = mc_ld_add_u32(reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(mc) + t.live_off + idx));
                } else {
                    for (int k = 0; k < t.world; ++k)
                        u |= ld_cg_u32(reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(t.bufs[k]) + t.live_off + idx));
                }
            }
            // rows q .. q+3 (beyond n_rows only in the segment's last piece: their bytes are padding zeros or belong
            // to rows of another piece -- mask them off)
            uint32_t m = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b)
                if (((u >> (8 * b)) & 0xffu) && q + b < n_rows) m |= 1u << b;
            const int cnt = __popc(m);
            // warp-aggregated append
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int x = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += x;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            uint32_t base = 0;
            if (lane == 31 && total) base = atomicAdd(s_count, (uint32_t)total);
            base = __shfl_sync(0xffffffffu, base, 31);
            uint32_t pos = base + (uint32_t)(incl - cnt);
#pragma unroll
            for (int b = 0; b < 4; ++b)
                if ((m >> b) & 1u) s_list[pos++] = (uint16_t)(q + b);
        }
        __syncthreads();
        // (2) the listed rows.  (The list is in no particular order: every element is handled independently.)
        const int64_t work = (int64_t)(*s_count) * row4;
        for (int64_t j = tid; j < work; j += blockDim.x) {
            const int r = (int)(j / row4), part = (int)(j % row4);
            const int64_t i = first4 + (p0 + s_list[r]) * row4 + part;
            float4 acc;
            if (MC) {
                acc = mc_ld_add4(mc + 4 * i);
                mc_st4(mc + 4 * i, acc);
            } else {
                acc = ld_cg4(reinterpret_cast<const float4*>(t.bufs[0]) + i);
                for (int k = 1; k < t.world; ++k) {
                    const float4 v = ld_cg4(reinterpret_cast<const float4*>(t.bufs[k]) + i);
                    acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
                }
                for (int k = 0; k < t.world; ++k) st_cg4(reinterpret_cast<float4*>(t.bufs[k]) + i, acc);
            }
        }
        __syncthreads();
    }
}

template <int U>
__global__ void __launch_bounds__(256)
p2p_allreduce_kernel(const __grid_constant__ P2PTab t, int mode) {
    __shared__ int s_flag;
    __shared__ uint16_t s_list[SP_ROWS];
    __shared__ uint32_t s_count;
    const int tid = threadIdx.x;
    const uint32_t epoch = call_epoch(t);
    // phase 0
    if (blockIdx.x == 0 && tid < t.world) st_release_sys(t.signals[tid] + 0 * P2P_MAX_RANKS + t.rank, epoch);
    if (tid == 0) s_flag = 1;
    __syncthreads();
    if (!wait_row(t, 0, tid, epoch)) s_flag = 0;
    __syncthreads();
    if (s_flag) {
        const bool aligned32 = (reinterpret_cast<uintptr_t>(t.bufs[t.rank]) & 31) == 0;
        for (int sgm = 0; sgm < t.n_seg; ++sgm) {
            const int64_t f4 = t.seg_first4[sgm], n4 = t.seg_n4[sgm];
            const bool is_max = (t.seg_max_mask >> sgm) & 1u;
            if (t.seg_row4[sgm] > 0 && mode == 0) {
                reduce_rows<false>(t, nullptr, f4, n4, t.seg_row4[sgm], t.seg_row0[sgm], s_list, &s_count);
                continue;
            }
            if (mode == 0 && aligned32 && ((f4 | n4) & 1) == 0) {
                if (is_max) reduce_slice8<true, (U > 2 ? U / 2 : 1)>(t, f4 / 2, n4 / 2);
                else reduce_slice8<false, (U > 2 ? U / 2 : 1)>(t, f4 / 2, n4 / 2);
            } else {
                if (is_max) reduce_slice<true, U>(t, f4, n4, mode);
                else reduce_slice<false, U>(t, f4, n4, mode);
            }
        }
    }
    // phase 1.  One system fence per CTA, by the thread that then counts the CTA as finished (the barrier makes
    // the other threads' stores part of what the fence orders); a fence per thread cost ~0.3 ms per call: every
    // warp waited for its own NVLink round trip
    __syncthreads();
    if (tid == 0) {
        __threadfence_system();
        uint32_t* counter = t.signals[t.rank] + P2P_COUNTER_WORD;
        const uint32_t prev = atomicAdd(counter, 1u);
        s_flag = (prev == gridDim.x - 1) ? 2 : 0;
        if (s_flag == 2) {
            *counter = 0u;
            __threadfence_system();   // acquire side: every other CTA's fenced stores precede the signal below
        }
    }
    __syncthreads();
    if (s_flag == 2) {
        if (tid < t.world) st_release_sys(t.signals[tid] + 1 * P2P_MAX_RANKS + t.rank, epoch);
        wait_row(t, 1, tid, epoch);
        // every CTA of this launch has read the call counter (they all passed the count above): advance it
        __syncthreads();
        if (tid == 0 && t.epoch == 0u) st_release_sys(t.signals[t.rank] + P2P_EPOCH_WORD, epoch);
    }
}

template <int U>
__global__ void __launch_bounds__(256)
mc_allreduce_kernel(const __grid_constant__ P2PTab t, float* mc) {
    __shared__ int s_flag;
    __shared__ uint16_t s_list[SP_ROWS];
    __shared__ uint32_t s_count;
    const int tid = threadIdx.x;
    const uint32_t epoch = call_epoch(t);
    // phase 0: "my buffer is complete" to every peer; wait for theirs
    if (blockIdx.x == 0 && tid < t.world) st_release_sys(t.signals[tid] + 0 * P2P_MAX_RANKS + t.rank, epoch);
    if (tid == 0) s_flag = 1;
    __syncthreads();
    if (!wait_row(t, 0, tid, epoch)) s_flag = 0;
    __syncthreads();
    if (s_flag) {
        const int64_t stride = (int64_t)gridDim.x * blockDim.x;
        for (int sgm = 0; sgm < t.n_seg; ++sgm) {
            const int64_t first4 = t.seg_first4[sgm], n4 = t.seg_n4[sgm];
            if (t.seg_row4[sgm] > 0) {
                reduce_rows<true>(t, mc, first4, n4, t.seg_row4[sgm], t.seg_row0[sgm], s_list, &s_count);
                continue;
            }
            const int64_t lo = first4 + n4 * t.rank / t.world, hi = first4 + n4 * (t.rank + 1) / t.world;   // my slice
            if ((t.seg_max_mask >> sgm) & 1u) {
                uint32_t* w = reinterpret_cast<uint32_t*>(mc);
                for (int64_t i = 4 * lo + (int64_t)blockIdx.x * blockDim.x + tid; i < 4 * hi; i += stride)
                    mc_st_u32(w + i, mc_ld_max_u32(w + i));
                continue;
            }
            for (int64_t i0 = lo + (int64_t)blockIdx.x * blockDim.x + tid; i0 < hi; i0 += U * stride) {
                float4 v[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int64_t i = i0 + u * stride;
                    if (i < hi) v[u] = mc_ld_add4(mc + 4 * i);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int64_t i = i0 + u * stride;
                    if (i < hi) mc_st4(mc + 4 * i, v[u]);
                }
            }
        }
    }
    // phase 1: my replicated stores have landed everywhere -> tell every peer, wait for theirs (see p2p_allreduce_kernel)
    __syncthreads();
    if (tid == 0) {
        __threadfence_system();
        uint32_t* counter = t.signals[t.rank] + P2P_COUNTER_WORD;
        const uint32_t prev = atomicAdd(counter, 1u);
        s_flag = (prev == gridDim.x - 1) ? 2 : 0;
        if (s_flag == 2) {
            *counter = 0u;
            __threadfence_system();
        }
    }
    __syncthreads();
    if (s_flag == 2) {
        if (tid < t.world) st_release_sys(t.signals[tid] + 1 * P2P_MAX_RANKS + t.rank, epoch);
        wait_row(t, 1, tid, epoch);
        // every CTA of this launch has read the call counter (they all passed the count above): advance it
        __syncthreads();
        if (tid == 0 && t.epoch == 0u) st_release_sys(t.signals[t.rank] + P2P_EPOCH_WORD, epoch);
    }
}

cudaError_t launch_mc_allreduce(const P2PTab& t, float* mc, cudaStream_t st) {
    static const int blocks = getenv("B200SPLAT_MC_BLOCKS") ? atoi(getenv("B200SPLAT_MC_BLOCKS")) : NUM_SMS;
    static const int unroll = getenv("B200SPLAT_MC_UNROLL") ? atoi(getenv("B200SPLAT_MC_UNROLL")) : 4;
    if (unroll == 8) mc_allreduce_kernel<8><<<blocks, 256, 0, st>>>(t, mc);
    else if (unroll == 2) mc_allreduce_kernel<2><<<blocks, 256, 0, st>>>(t, mc);
    else mc_allreduce_kernel<4><<<blocks, 256, 0, st>>>(t, mc);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_p2p_allreduce(const P2PTab& t, cudaStream_t st) {
    static const int blocks = getenv("B200SPLAT_P2P_BLOCKS") ? atoi(getenv("B200SPLAT_P2P_BLOCKS")) : NUM_SMS * 2;   // 2 per SM: same time as 4 or 8 (NVLink-bound), leaves room for the compute it overlaps
    static const int unroll = getenv("B200SPLAT_P2P_UNROLL") ? atoi(getenv("B200SPLAT_P2P_UNROLL")) : 4;
    static const int mode = getenv("B200SPLAT_P2P_MODE") ? atoi(getenv("B200SPLAT_P2P_MODE")) : 0;   // 1/2: timing experiments
    if (unroll == 8) p2p_allreduce_kernel<8><<<blocks, 256, 0, st>>>(t, mode);
    else if (unroll == 2) p2p_allreduce_kernel<2><<<blocks, 256, 0, st>>>(t, mode);
    else p2p_allreduce_kernel<4><<<blocks, 256, 0, st>>>(t, mode);
    count_launch();
    return cudaGetLastError();
}

}  // namespace b200splat

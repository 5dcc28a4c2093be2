// K2 inclusive scan (decoupled look-back), K2+K3 fused scan + duplicateWithKeys, K4 onesweep LSD radix sort of (u64 key,
// u32 value) pairs, K4d look-back-free pair partition (count matrix -> offsets -> partition), K5 identifyTileRanges +
// longest-list-first tile order, block order of render backward.  Hand-written; no CUB.
//
// Replaces cub::DeviceScan::InclusiveSum, cub::DeviceRadixSort::SortPairs and
// rasterizer_impl.cu identifyTileRanges as called by upstream CudaRasterizer::Rasterizer::forward
// [UPSTREAM-RECALL]; reference call site renderer/diff_gaussian_rasterizer.py:122-131.
// All integer work: results are bit-exact by construction (stable sort, exact sums).
// Every kernel handles a whole view batch (blockIdx.y = view) and reads the pair count of a view from
// device memory (point_offsets[P-1]), so the host never has to wait for it.
#include "common.cuh"

#include <cstddef>
#include <cstdlib>

namespace b200splat {

__device__ __forceinline__ uint64_t ld_volatile_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_volatile_u64(uint64_t* p, uint64_t v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_volatile_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ============================================================================================
// K2: single-pass inclusive scan of uint32 with decoupled look-back
// ============================================================================================
// 8192 elements per CTA: the look-back chain of a 1M-element scan is 123 tiles (4 warp-wide windows) deep;
// with 2048-element tiles it was 489 tiles / 15 windows and the launch took 38 us for 4 views
constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

constexpr uint64_t FLAG_AGG = 1ull << 32;
constexpr uint64_t FLAG_INC = 2ull << 32;

struct ScanTab {
    int V;
    const uint64_t* perm[MAX_VIEWS];   // optional: element i of the scan is in[(uint32)perm[i]] (depth order)
    const uint32_t* in[MAX_VIEWS];
    uint32_t* out[MAX_VIEWS];
    uint32_t* ticket[MAX_VIEWS];
    uint64_t* desc[MAX_VIEWS];
    uint64_t* notify;   // optional host-mapped words [V]: (epoch << 32 | total) of each view as soon as it is known
    uint32_t epoch;
};

__global__ void __launch_bounds__(SCAN_THREADS)
scan_lookback_kernel(int64_t n, const __grid_constant__ ScanTab tab) {
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_warp[SCAN_THREADS / 32];
    __shared__ uint32_t s_excl;
    const int view = blockIdx.y;
    const uint32_t* __restrict__ in = tab.in[view];
    uint32_t* __restrict__ out = tab.out[view];
    uint64_t* desc = tab.desc[view];
    if (threadIdx.x == 0) s_tile = atomicAdd(tab.ticket[view], 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const int64_t base = (int64_t)tile * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    const uint64_t* __restrict__ perm = tab.perm[view];
    // gathered input: coalesced (striped) loads of the order words, dependent gathers, then a transpose through
    // shared memory to the blocked arrangement the scan wants (index + index/32: conflict-free both ways).  The
    // blocked loads (each lane 128 bytes apart) throttled the load/store unit: 32 requests per instruction.
    __shared__ uint32_t s_x[SCAN_TILE + SCAN_TILE / 32];
    const int64_t tile_base = (int64_t)tile * SCAN_TILE;
    if (perm != nullptr) {
        uint32_t g[SCAN_ITEMS];
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i) {
            const int64_t e = tile_base + i * SCAN_THREADS + threadIdx.x;
            g[i] = e < n ? (uint32_t)__ldg(perm + e) : 0xffffffffu;
        }
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i) {
            const int idx = i * SCAN_THREADS + threadIdx.x;
            s_x[idx + (idx >> 5)] = g[i] != 0xffffffffu ? __ldg(in + g[i]) : 0u;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i) {
            const int idx = threadIdx.x * SCAN_ITEMS + i;
            v[i] = s_x[idx + (idx >> 5)];
        }
    } else if (base + SCAN_ITEMS <= n) {
        const uint4* p = reinterpret_cast<const uint4*>(in + base);
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS / 4; ++i) {
            const uint4 a = __ldg(p + i);
            v[4 * i] = a.x, v[4 * i + 1] = a.y, v[4 * i + 2] = a.z, v[4 * i + 3] = a.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i) v[i] = (base + i < n) ? in[base + i] : 0u;
    }
    uint32_t tsum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        tsum += v[i];
        v[i] = tsum;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t warp_off = 0, block_total = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; ++w) {
        uint32_t t = s_warp[w];
        if (w < warp) warp_off += t;
        block_total += t;
    }
    // look-back by warp 0
    if (warp == 0) {
        uint32_t excl = 0;
        if (tile == 0) {
            if (lane == 0) st_volatile_u64(desc + 0, FLAG_INC | block_total);
        } else {
            if (lane == 0) st_volatile_u64(desc + tile, FLAG_AGG | block_total);
            int64_t look = (int64_t)tile - 1;
            while (true) {
                const int64_t idx = look - lane;
                uint64_t d = FLAG_INC;  // virtual predecessor of tile 0: inclusive prefix 0
                if (idx >= 0) {
                    do {
                        d = ld_volatile_u64(desc + idx);
                    } while ((d >> 32) == 0);
                }
                const uint32_t inc_mask = __ballot_sync(0xffffffffu, (d >> 32) == 2);
                uint32_t val = (uint32_t)d;
                if (inc_mask) {
                    const int first = __ffs(inc_mask) - 1;
                    if (lane > first) val = 0;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
                excl += val;
                if (inc_mask) break;
                look -= 32;
            }
            if (lane == 0) st_volatile_u64(desc + tile, FLAG_INC | (uint64_t)(excl + block_total));
        }
        if (lane == 0) s_excl = excl;
        // the last tile knows the view's total (= num_rendered in the pipeline): tell the host right away
        if (lane == 0 && tab.notify != nullptr && (int64_t)(tile + 1) * SCAN_TILE >= n)
        {
            st_volatile_u64(tab.notify + view, ((uint64_t)tab.epoch << 32) | (uint64_t)(excl + block_total));
            __threadfence_system();
        }
    }
    __syncthreads();
    const uint32_t off = s_excl + warp_off + (inc - tsum);
    if (perm != nullptr) {   // back through shared memory: coalesced stores
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i) {
            const int idx = threadIdx.x * SCAN_ITEMS + i;
            s_x[idx + (idx >> 5)] = v[i] + off;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i) {
            const int idx = i * SCAN_THREADS + threadIdx.x;
            if (tile_base + idx < n) out[tile_base + idx] = s_x[idx + (idx >> 5)];
        }
    } else if (base + SCAN_ITEMS <= n) {
        uint4* p = reinterpret_cast<uint4*>(out + base);
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS / 4; ++i)
            p[i] = make_uint4(v[4 * i] + off, v[4 * i + 1] + off, v[4 * i + 2] + off, v[4 * i + 3] + off);
    } else {
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i)
            if (base + i < n) out[base + i] = v[i] + off;
    }
}

// Early notice of every view's pair count (= sum of tiles_touched) for the host: run right after preprocess, long
// before the scan produces the same number, so that a host waiting for it (b200splat_batch_forward_args.pairs_notify)
// gets it while most of the forward is still queued.  Grid (blocks, V); the last CTA of a view writes the word.
constexpr int STATUS_PAIR_SUM = 1, STATUS_PAIR_TICKET = 2;
__global__ void __launch_bounds__(256)
pair_count_kernel(const __grid_constant__ BatchTab tab, uint64_t* notify, uint32_t epoch) {
    const ViewTab& vt = tab.v[blockIdx.y];
    const uint4* __restrict__ t4 = reinterpret_cast<const uint4*>(vt.tiles_touched);
    const int n4 = tab.P / 4;
    uint32_t sum = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
        const uint4 a = __ldg(t4 + i);
        sum += a.x + a.y + a.z + a.w;
    }
    if (blockIdx.x == 0 && threadIdx.x < (tab.P & 3)) sum += vt.tiles_touched[4 * n4 + threadIdx.x];
    sum = __reduce_add_sync(0xffffffffu, sum);
    __shared__ uint32_t s_sum;
    if (threadIdx.x == 0) s_sum = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && sum) atomicAdd(&s_sum, sum);
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_sum) atomicAdd(vt.status + STATUS_PAIR_SUM, s_sum);
        __threadfence();
        if (atomicAdd(vt.status + STATUS_PAIR_TICKET, 1u) == gridDim.x - 1) {
            __threadfence();
            const uint32_t total = atomicAdd(vt.status + STATUS_PAIR_SUM, 0u);
            st_volatile_u64(notify + blockIdx.y, ((uint64_t)epoch << 32) | (uint64_t)total);
            __threadfence_system();
        }
    }
}
cudaError_t launch_pair_count(const BatchTab& tab, uint64_t* notify, uint32_t epoch, cudaStream_t st) {
    if (tab.P <= 0 || notify == nullptr) return cudaSuccess;
    const int blocks = min(NUM_SMS, (tab.P / 4 + 255) / 256 + 1);
    pair_count_kernel<<<dim3(blocks, tab.V), 256, 0, st>>>(tab, notify, epoch);
    count_launch();
    return cudaGetLastError();
}

// ============================================================================================
// K2+K3 fused: inclusive scan of tiles_touched in depth order AND duplicateWithKeys in one kernel
// ============================================================================================
// The stand-alone scan gathered tiles_touched[order[i]], wrote point_offsets, and duplicateWithKeys then re-read the
// order words and the offsets and gathered the tile rectangle of the same Gaussians.  Here a CTA takes 8192 consecutive
// ranks of the depth order: order words (coalesced) -> ONE gather of the 8-byte rectangle (its area is the count) ->
// CTA scan + decoupled look-back (the aggregate is published BEFORE the emission, so successors never wait for it) ->
// every thread emits the pair words (tile << 32 | gaussian) of its 16 Gaussians at their scanned offsets, and the per-tile
// pair counts (global bins of the tile partition, tile ranges) are accumulated in shared memory and flushed once per CTA.
// Same outputs, bit for bit, as scan_lookback_kernel + duplicate_kernel: point_offsets, pair words, tile_count, hist.
constexpr int SD_MAX_TILES = 8192;
constexpr int PO_CHUNKS_WORDS = 32 * 1024;   // chunk bases of partition_offsets_kernel: [PO_CHUNKS][WIDE_BINS]
constexpr int SD_NPT = 8;          // partition tiles of a CTA's pair range counted in shared memory (pt mode)
constexpr int WIDE_BITS = 10;      // the single-pass pair partition handles tile ids of up to 10 bits (512 x 512 images)
constexpr int WIDE_BINS = 1 << WIDE_BITS;
template <int SD_T>   // threads per CTA (512: one CTA per SM; 256: two, whose phases overlap)
__global__ void __launch_bounds__(SD_T, SD_T == 256 ? 2 : 1)
scan_duplicate_kernel(const __grid_constant__ BatchTab tab) {
    constexpr int SD_TILE = SD_T * SCAN_ITEMS;
    extern __shared__ uint32_t s_dyn[];                 // [SD_TILE + SD_TILE/32] exchange | [passes*256] | [T]
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_warp[SD_T / 32];
    __shared__ uint32_t s_excl;
    const ViewTab& vt = tab.v[blockIdx.y];
    const int64_t n = tab.P;
    const int T = tab.grid_x * tab.grid_y, gx = tab.grid_x;
    const int passes = tab.digit_passes, end_bit = tab.end_bit;
    uint32_t* s_x = s_dyn;
    uint32_t* s_hist = s_dyn + SD_TILE + SD_TILE / 32;
    uint32_t* s_cnt = s_hist + passes * 256;
    // pt mode (look-back-free pair partition, K4d): the pairs are counted per (partition tile of pt_words words of the
    // pair stream, image tile) -- SD_NPT partition tiles of this CTA's part of the stream in shared memory, any further
    // ones (huge splats) straight in global memory -- and flushed into the matrix vt.desc[partition tile][1024]
    const uint32_t pt_words = (uint32_t)tab.pt_words;
    const int cnt_words = pt_words ? SD_NPT * T : T;
    for (int i = threadIdx.x; i < passes * 256 + cnt_words; i += SD_T) s_hist[i] = 0;
    if (threadIdx.x == 0) s_tile = atomicAdd(vt.scan_ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const int64_t tile_base = (int64_t)tile * SD_TILE;
    const uint64_t* __restrict__ order = vt.gwords[0];
    const ushort4* __restrict__ rect = vt.rect;
    uint64_t* desc = vt.scan_desc;
    // ---- striped loads: order word, then the rectangle of that Gaussian ---------------------------------
    uint32_t g[SCAN_ITEMS];
    ushort4 rc[SCAN_ITEMS];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        const int64_t e = tile_base + i * SD_T + threadIdx.x;
        g[i] = e < n ? (uint32_t)__ldg(order + e) : 0xffffffffu;
    }
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i)
        rc[i] = g[i] != 0xffffffffu ? __ldg(rect + g[i]) : make_ushort4(0, 0, 0, 0);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        const int idx = i * SD_T + threadIdx.x;
        s_x[idx + (idx >> 5)] = (uint32_t)((rc[i].z - rc[i].x) * (rc[i].w - rc[i].y));
    }
    __syncthreads();
    // ---- blocked scan (as scan_lookback_kernel) --------------------------------------------------------------
    uint32_t v[SCAN_ITEMS];
    uint32_t tsum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        const int idx = threadIdx.x * SCAN_ITEMS + i;
        tsum += s_x[idx + (idx >> 5)];
        v[i] = tsum;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t warp_off = 0, block_total = 0;
#pragma unroll
    for (int w = 0; w < SD_T / 32; ++w) {
        const uint32_t t = s_warp[w];
        if (w < warp) warp_off += t;
        block_total += t;
    }
    if (warp == 0) {
        uint32_t excl = 0;
        if (tile == 0) {
            if (lane == 0) st_volatile_u64(desc + 0, FLAG_INC | block_total);
        } else {
            if (lane == 0) st_volatile_u64(desc + tile, FLAG_AGG | block_total);
            int64_t look = (int64_t)tile - 1;
            while (true) {
                const int64_t idx = look - lane;
                uint64_t d = FLAG_INC;
                if (idx >= 0) {
                    do {
                        d = ld_volatile_u64(desc + idx);
                    } while ((d >> 32) == 0);
                }
                const uint32_t inc_mask = __ballot_sync(0xffffffffu, (d >> 32) == 2);
                uint32_t val = (uint32_t)d;
                if (inc_mask) {
                    const int first = __ffs(inc_mask) - 1;
                    if (lane > first) val = 0;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
                excl += val;
                if (inc_mask) break;
                look -= 32;
            }
            if (lane == 0) st_volatile_u64(desc + tile, FLAG_INC | (uint64_t)(excl + block_total));
        }
        if (lane == 0) s_excl = excl;
    }
    __syncthreads();
    const uint32_t off = s_excl + warp_off + (inc - tsum);
    // ---- inclusive offsets back to the striped arrangement: point_offsets (coalesced) and the emission ------
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        const int idx = threadIdx.x * SCAN_ITEMS + i;
        s_x[idx + (idx >> 5)] = v[i] + off;
    }
    __syncthreads();
    uint32_t* __restrict__ out = vt.point_offsets;
    uint64_t* __restrict__ words = vt.keys[0];
    const uint32_t pt_first = pt_words ? s_excl / pt_words : 0u;   // partition tile of this CTA's first pair
    uint32_t* __restrict__ matrix = vt.desc;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        const int idx = i * SD_T + threadIdx.x;
        const uint32_t incl = s_x[idx + (idx >> 5)];
        if (tile_base + idx < n) out[tile_base + idx] = incl;
        const int x0 = rc[i].x, y0 = rc[i].y, x1 = rc[i].z, y1 = rc[i].w;
        const uint32_t ntiles = (uint32_t)((x1 - x0) * (y1 - y0));
        if (ntiles == 0) continue;
        if (incl > tab.capacity) {                       // pairs beyond the binning capacity are dropped and flagged
            atomicOr(vt.status + STATUS_OVERFLOW, 1u);
            continue;
        }
        uint32_t o = incl - ntiles;
        if (pt_words) {
            uint32_t pt = o / pt_words;
            uint32_t next = (pt + 1u) * pt_words;        // first pair of the next partition tile
            pt -= pt_first;
            if (incl <= next && pt < (uint32_t)SD_NPT) {   // all pairs of the Gaussian in one counted partition tile
                uint32_t* __restrict__ row = s_cnt + pt * T;
                for (int ty = y0; ty < y1; ++ty) {
                    for (int tx = x0; tx < x1; ++tx) {
                        const uint32_t t = (uint32_t)(ty * gx + tx);
                        words[o++] = ((uint64_t)t << 32) | (uint64_t)g[i];
                        atomicAdd(&row[t], 1u);
                    }
                }
            } else {
                for (int ty = y0; ty < y1; ++ty) {
                    for (int tx = x0; tx < x1; ++tx) {
                        const uint32_t t = (uint32_t)(ty * gx + tx);
                        if (o == next) ++pt, next += pt_words;
                        words[o++] = ((uint64_t)t << 32) | (uint64_t)g[i];
                        if (pt < (uint32_t)SD_NPT) atomicAdd(&s_cnt[pt * T + t], 1u);
                        else atomicAdd(&matrix[(size_t)(pt_first + pt) * WIDE_BINS + t], 1u);
                    }
                }
            }
        } else {
            for (int ty = y0; ty < y1; ++ty) {
                for (int tx = x0; tx < x1; ++tx) {
                    const uint32_t t = (uint32_t)(ty * gx + tx);
                    words[o++] = ((uint64_t)t << 32) | (uint64_t)g[i];
                    atomicAdd(&s_cnt[t], 1u);
                }
            }
        }
    }
    __syncthreads();
    if (pt_words) {   // flush the CTA's part of the count matrix; the per-tile totals are summed by partition_offsets_kernel
        // rows this CTA's pairs can have reached: from its first pair's partition tile to its last pair's
        const uint32_t last_pair = s_excl + block_total;   // one past
        const int rows = min(SD_NPT, (int)((last_pair + pt_words - 1) / pt_words - pt_first));
        for (int r = 0; r < rows; ++r) {
            uint32_t* __restrict__ dst = matrix + (size_t)(pt_first + (uint32_t)r) * WIDE_BINS;
            for (int t = threadIdx.x; t < T; t += SD_T) {
                const uint32_t c = s_cnt[r * T + t];
                if (c) atomicAdd(&dst[t], c);
            }
        }
        return;
    }
    // ---- flush: per-tile counts (and the digit histograms of the 8-bit passes, derived from them) -----------
    for (int t = threadIdx.x; t < T; t += SD_T) {
        const uint32_t c = s_cnt[t];
        if (c) {
            atomicAdd(&vt.tile_count[t], c);
            for (int p = 0; p < passes; ++p) {
                const int bits = min(8, end_bit - 8 * p);
                atomicAdd(&s_hist[p * 256 + (((uint32_t)t >> (8 * p)) & ((1u << bits) - 1u))], c);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * 256; i += SD_T) {
        const uint32_t c = s_hist[i];
        if (c) atomicAdd(&vt.hist[i], c);
    }
}

// true: the fused kernel covers this configuration (the tile histogram fits in shared memory)
bool scan_duplicate_supported(const BatchTab& tab) {
    static const bool on = [] {
        const char* e = getenv("B200SPLAT_FUSED_SCAN_DUP");
        return !(e && e[0] == '0');
    }();
    return on && tab.grid_x * tab.grid_y <= SD_MAX_TILES;
}

static int sd_threads() {   // B200SPLAT_SD_THREADS = 256 | 512
    static const int t = [] {
        const char* e = getenv("B200SPLAT_SD_THREADS");
        return (e && atoi(e) == 512) ? 512 : 256;   // measured (headline, 4 views): 256 -> 1.2187, 512 -> 1.2244 ms per step
    }();
    return t;
}
cudaError_t launch_scan_duplicate(const BatchTab& tab, cudaStream_t st) {
    const int64_t n = tab.P;
    if (n <= 0) return cudaSuccess;
    const int threads = sd_threads();
    const int64_t tile = (int64_t)threads * SCAN_ITEMS;
    const int64_t tiles = (n + tile - 1) / tile;
    const int T = tab.grid_x * tab.grid_y;
    const size_t smem = ((size_t)tile + tile / 32 + (size_t)tab.digit_passes * 256 +
                         (size_t)(tab.pt_words ? SD_NPT * T : T)) * sizeof(uint32_t);
    static size_t attr = 0;
    if (smem > 48 * 1024 && smem > attr) {
        cudaError_t e = cudaFuncSetAttribute(scan_duplicate_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(scan_duplicate_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr = smem;
    }
    if (threads == 256) scan_duplicate_kernel<256><<<dim3((unsigned)tiles, tab.V), 256, smem, st>>>(tab);
    else scan_duplicate_kernel<512><<<dim3((unsigned)tiles, tab.V), 512, smem, st>>>(tab);
    count_launch();
    return cudaGetLastError();
}

size_t scan_workspace_bytes(int64_t n) {
    int64_t tiles = (n + SCAN_TILE / 2 - 1) / (SCAN_TILE / 2);   // (the fused kernel may run with half-size tiles)
    return align_up(16 + (size_t)tiles * 8, 256);
}

// One launch zeroes every small per-view work area of a forward (status words, scan descriptors, the depth
// sort's and the pair sort's histograms / tickets / look-back descriptors, per-tile counts) for all views --
// instead of 5 memset nodes per view in the stream.
constexpr int CLEAR_REGIONS = 5;
struct ClearTab {
    uint32_t* p[MAX_VIEWS][CLEAR_REGIONS];
    uint32_t words[CLEAR_REGIONS];
};
__global__ void __launch_bounds__(256)
clear_regions_kernel(const __grid_constant__ ClearTab tab) {
    uint32_t* __restrict__ p = tab.p[blockIdx.z][blockIdx.y];
    const uint32_t n = tab.words[blockIdx.y];
    if (p == nullptr) return;
    const uint32_t n4 = ((reinterpret_cast<uintptr_t>(p) & 15) == 0) ? n / 4 : 0;
    uint4* p4 = reinterpret_cast<uint4*>(p);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x)
        p4[i] = make_uint4(0u, 0u, 0u, 0u);
    for (uint32_t i = 4 * n4 + blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = 0u;
}
size_t pair_sort_zero_bytes(int64_t capacity, int end_bit);
cudaError_t launch_clear_batch(const BatchTab& tab, bool with_binning, cudaStream_t st) {
    if (tab.P <= 0) return cudaSuccess;
    ClearTab c;
    const int T = tab.grid_x * tab.grid_y;
    c.words[0] = STATUS_WORDS;
    c.words[1] = (uint32_t)(scan_workspace_bytes(tab.P) / 4);
    c.words[2] = (uint32_t)(sort_workspace_zero_bytes(tab.P, 32) / 4);
    c.words[3] = with_binning ? (uint32_t)(pair_sort_zero_bytes(tab.capacity, tab.end_bit) / 4) : 0u;
    c.words[4] = with_binning ? (uint32_t)T : 0u;
    for (int v = 0; v < MAX_VIEWS; ++v) {
        const bool live = v < tab.V;
        c.p[v][0] = live ? tab.v[v].status : nullptr;
        c.p[v][1] = live ? tab.v[v].scan_ticket : nullptr;
        c.p[v][2] = live ? tab.v[v].ghist : nullptr;
        c.p[v][3] = live && with_binning ? tab.v[v].hist : nullptr;
        c.p[v][4] = live && with_binning ? tab.v[v].tile_count : nullptr;
    }
    clear_regions_kernel<<<dim3(48, CLEAR_REGIONS, tab.V), 256, 0, st>>>(c);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_scan_batch(const BatchTab& tab, cudaStream_t st, bool cleared, uint64_t* notify, uint32_t epoch) {
    const int64_t n = tab.P;
    if (n <= 0) return cudaSuccess;
    const int64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    ScanTab s;
    s.V = tab.V;
    s.notify = notify, s.epoch = epoch;
    for (int v = 0; v < tab.V; ++v) {
        s.perm[v] = tab.v[v].gwords[0];
        s.in[v] = tab.v[v].tiles_touched;
        s.out[v] = tab.v[v].point_offsets;
        s.ticket[v] = tab.v[v].scan_ticket;
        s.desc[v] = tab.v[v].scan_desc;
        if (!cleared) {
            cudaError_t e = cudaMemsetAsync(tab.v[v].scan_ticket, 0, scan_workspace_bytes(n), st);
            if (e != cudaSuccess) return e;
        }
    }
    scan_lookback_kernel<<<dim3((unsigned)tiles, tab.V), SCAN_THREADS, 0, st>>>(n, s);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_inclusive_scan(int64_t n, const uint32_t* in, uint32_t* out, void* ws, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const int64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    cudaError_t e = cudaMemsetAsync(ws, 0, scan_workspace_bytes(n), st);
    if (e != cudaSuccess) return e;
    ScanTab s;
    s.V = 1;
    s.notify = nullptr, s.epoch = 0;
    s.perm[0] = nullptr;
    s.in[0] = in;
    s.out[0] = out;
    s.ticket[0] = reinterpret_cast<uint32_t*>(ws);
    s.desc[0] = reinterpret_cast<uint64_t*>(reinterpret_cast<char*>(ws) + 16);
    scan_lookback_kernel<<<dim3((unsigned)tiles, 1), SCAN_THREADS, 0, st>>>(n, s);
    count_launch();
    return cudaGetLastError();
}

// ============================================================================================
// K4: onesweep LSD radix sort, 8-bit digits, (u64 key, u32 value)
// ============================================================================================
constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_ITEMS_MAX = 24;
constexpr int SORT_TILE = SORT_THREADS * 16;  // default 4096 pairs per CTA (workspace sizing assumes >= 2048)

constexpr uint32_t DESC_AGG = 1u << 30;
constexpr uint32_t DESC_INC = 2u << 30;
constexpr uint32_t DESC_VAL = (1u << 30) - 1;

// items per thread of the keys-only passes (tile = 256 * items); B200SPLAT_SORT_ITEMS = 8 | 16 | 24
static int sort_items() {
    static const int it = [] {
        const char* e = getenv("B200SPLAT_SORT_ITEMS");
        const int v = e ? atoi(e) : 24;   // measured on B200 (1M Gaussians, 4 views): 8 -> 202, 16 -> 171, 24 -> 163 us/view
        return (v == 8 || v == 16 || v == 512) ? v : 24;   // 512 = 512 threads x 12 items
    }();
    return it;
}
static int gsort_items() {   // tile size of the Gaussian (depth) sort: P keys only, so smaller tiles spread better
    static const int it = [] {
        const char* e = getenv("B200SPLAT_GSORT_ITEMS");
        const int v = e ? atoi(e) : 24;
        return (v == 8 || v == 16) ? v : 24;
    }();
    return it;
}
int sort_tiles_for(int64_t n) {   // tiles of the pipeline's (keys-only or pair) passes
    const int64_t tile = sort_items() == 512 ? 6144 : (int64_t)SORT_THREADS * sort_items();
    return (int)((n + tile - 1) / tile);
}
static int sort_tiles_pairs(int64_t n) { return (int)((n + SORT_TILE - 1) / SORT_TILE); }

// words per partition tile when the look-back-free partition (K4d) applies: one wide pass, the fused kernel, the
// count matrix of a CTA in shared memory.  B200SPLAT_DIRECT_PARTITION=0 selects the look-back kernel (A/B).
int partition_direct_words(const BatchTab& tab) {
    static const bool on = [] {
        const char* e = getenv("B200SPLAT_DIRECT_PARTITION");
        return !(e && e[0] == '0');
    }();
    if (!on || tab.digit_passes != 0 || !scan_duplicate_supported(tab)) return 0;
    if (tab.grid_x * tab.grid_y > WIDE_BINS) return 0;
    const int items = sort_items();
    return SORT_THREADS * (items == 8 || items == 16 ? items : 24);
}


// Histogram of every digit place in one read of the keys (stand-alone sort only; the pipeline gets its
// histograms from duplicateWithKeys).  hist: [passes][256] u32 (zeroed).
struct HistTab {
    const uint64_t* keys[MAX_VIEWS];
    uint32_t* hist[MAX_VIEWS];
};
template <int PASSES>   // 0: run-time pass count (stand-alone sort); the pipeline's depth sort uses 4
__global__ void __launch_bounds__(256)
radix_histogram_kernel(int64_t n, int passes_rt, int end_bit, int shift_base, const __grid_constant__ HistTab tab) {
    __shared__ uint32_t s_hist[MAX_PASSES * RADIX];
    const int passes = PASSES ? PASSES : passes_rt;
    const uint64_t* __restrict__ keys = tab.keys[blockIdx.y];
    uint32_t* __restrict__ hist = tab.hist[blockIdx.y];
    for (int i = threadIdx.x; i < passes * RADIX; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    constexpr int UN = 4;                              // independent loads in flight per thread
    // top digit (the exponent byte of depth): a view holds two or three distinct values, and 256 threads adding to the
    // same two shared-memory words serialise -- every thread counts its two most recent values in registers instead
    uint32_t da = 0xffffffffu, db = 0xffffffffu, ca = 0, cb = 0;
    const int top_shift = (passes - 1) * RADIX_BITS;
    const uint32_t top_mask = (1u << min(RADIX_BITS, end_bit - top_shift)) - 1u;
    uint32_t* s_top = s_hist + (passes - 1) * RADIX;
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += UN * stride) {
        uint64_t kk[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) kk[u] = (i0 + u * stride < n) ? __ldg(keys + i0 + u * stride) : 0ull;
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            if (i0 + u * stride >= n) continue;
            const uint64_t k = kk[u] >> shift_base;
#pragma unroll
            for (int p = 0; p < (PASSES ? PASSES - 1 : MAX_PASSES - 1); ++p) {   // low digits are spread: plain atomics
                if (p + 1 < passes) {
                    const int shift = p * RADIX_BITS;
                    atomicAdd(&s_hist[p * RADIX + ((uint32_t)(k >> shift) & (RADIX - 1))], 1u);
                }
            }
            const uint32_t d = (uint32_t)(k >> top_shift) & top_mask;
            if (d == da) {
                ++ca;
            } else if (d == db) {
                ++cb;
            } else {                 // evict the older entry
                if (cb) atomicAdd(&s_top[db], cb);
                db = da, cb = ca;
                da = d, ca = 1;
            }
        }
    }
    if (ca) atomicAdd(&s_top[da], ca);
    if (cb) atomicAdd(&s_top[db], cb);
    __syncthreads();
    for (int i = threadIdx.x; i < passes * RADIX; i += blockDim.x) {
        const uint32_t c = s_hist[i];
        if (c) atomicAdd(&hist[i], c);
    }
}

struct SortView {
    const uint32_t* n_ptr;       // device pair count (nullptr: use n_fixed)
    uint32_t n_fixed;
    const uint32_t* overflow;    // != nullptr and *overflow != 0: skip (histograms do not describe the data)
    const uint64_t* keys_in;
    const uint32_t* vals_in;
    uint64_t* keys_out;
    uint32_t* vals_out;
    const uint32_t* hist_pass;
    uint32_t* ticket;
    uint32_t* desc;              // [tiles][256]
};
struct SortTab {
    uint32_t capacity;
    SortView v[MAX_VIEWS];
};

// match.any by ballots: lanes holding the same BITS-bit digit.  MATCH.ANY issues at ~1 per 60 cycles per SM (it was
// 58 % of a pass's stall samples); this is 4 instructions per bit (bit test, vote, select, and-xor), spelled in
// PTX because the C++ form compiles to 6-8 (second predicate, shifts, 64-bit tests folded back into the key).
template <int BITS>
__device__ __forceinline__ uint32_t ballot_match(uint32_t d) {
    uint32_t peers = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < BITS; ++b) {
        asm("{\n\t.reg .pred p;\n\t.reg .b32 t, bal;\n\t"
            "and.b32 t, %1, %2;\n\t"
            "setp.ne.u32 p, t, 0;\n\t"
            "vote.sync.ballot.b32 bal, p, 0xffffffff;\n\t"
            "selp.b32 t, 0, 0xffffffff, p;\n\t"
            "xor.b32 bal, bal, t;\n\t"
            "and.b32 %0, %0, bal;\n\t}"
            : "+r"(peers)
            : "r"(d), "r"(1u << b));
    }
    return peers;
}

template <int ITEMS, int NT = SORT_THREADS>
struct SortSmemT {
    uint64_t keys[NT * ITEMS];
    uint32_t warp_hist[NT / 32][RADIX + 1];   // +1: the same digit of different warps falls in different banks
    uint32_t local_excl[RADIX];   // exclusive offset of digit inside this tile
    uint32_t bin_offset[RADIX];   // global destination of the tile's first key of digit d, minus local_excl
    uint32_t s_h[RADIX / 32], s_l[RADIX / 32];
    uint32_t tile;
    uint32_t pad[3];
    uint32_t vals[NT * ITEMS];     // last: keys-only passes do not allocate it
};
using SortSmem = SortSmemT<16>;

// One pass: tile t of the input is ranked locally (stable), its per-digit counts are chained to the
// previous tiles by decoupled look-back, then keys/values are scattered through shared memory so that
// the global writes are coalesced per digit run.
template <bool HAS_VALS, int MINB, int SORT_ITEMS, int NT = 256>
__global__ void __launch_bounds__(NT, MINB)
onesweep_pass_kernel(int shift, int bits, const __grid_constant__ SortTab tab) {
    constexpr int SORT_THREADS = NT;
    constexpr int SORT_WARPS = NT / 32;
    constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SortSmemT<SORT_ITEMS, NT>& S = *reinterpret_cast<SortSmemT<SORT_ITEMS, NT>*>(smem_raw);
    const SortView& sv = tab.v[blockIdx.y];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (sv.overflow != nullptr && *sv.overflow != 0u) return;
    const int64_t n = sv.n_ptr ? (int64_t)min(*sv.n_ptr, tab.capacity) : (int64_t)sv.n_fixed;
    if ((int64_t)blockIdx.x * SORT_TILE >= n) return;   // launched over the capacity; tickets only order live CTAs
    if (tid == 0) S.tile = atomicAdd(sv.ticket, 1u);
    for (int i = tid; i < SORT_WARPS * (RADIX + 1); i += SORT_THREADS) (&S.warp_hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = S.tile;
    const int64_t tile_base = (int64_t)tile * SORT_TILE;
    const int valid = (int)min((int64_t)SORT_TILE, n - tile_base);
    const uint32_t mask = (1u << bits) - 1u;
    const uint64_t* __restrict__ keys_in = sv.keys_in;
    const uint32_t* __restrict__ vals_in = sv.vals_in;

    // ---- load (warp-striped: item i of lane l in warp w = w*512 + i*32 + l) ------------------
    uint64_t key[SORT_ITEMS];
    uint32_t rank[SORT_ITEMS];
    const int wbase = warp * (32 * SORT_ITEMS) + lane;
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        const int li = wbase + i * 32;
        key[i] = (li < valid) ? __ldg(keys_in + tile_base + li) : ~0ull;
    }
    // ---- warp-level stable ranking -------------------------------------------------------------
    // All 16 match.any are independent; the per-digit running count of the warp is advanced with one
    // shared-memory atomic per (item, digit group) issued by the group's first lane.  Items are issued in
    // order (the __syncwarp keeps the atomics of successive items ordered), but nothing waits for an
    // atomic's return value until all have been issued, so the latencies overlap.
    const uint32_t lt_mask = (1u << lane) - 1u;
    // fast path: the warp's digits span at most 4 consecutive values (the exponent byte of depth, the top tile
    // bits): one ballot per (value, item) gives the stable ranks with full ILP -- the generic path below would
    // chain ITEMS shared-memory atomics on the same one or two addresses
    uint32_t dmin = 0xffffffffu, dmax = 0u;
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        const uint32_t d = (uint32_t)(key[i] >> shift) & mask;
        dmin = min(dmin, d);
        dmax = max(dmax, d);
    }
    dmin = __reduce_min_sync(0xffffffffu, dmin);
    dmax = __reduce_max_sync(0xffffffffu, dmax);
    if (dmax - dmin < 4u) {
        for (uint32_t dv = dmin; dv <= dmax; ++dv) {
            uint32_t run = 0;
#pragma unroll
            for (int i = 0; i < SORT_ITEMS; ++i) {
                const bool mine = ((uint32_t)(key[i] >> shift) & mask) == dv;
                const uint32_t bal = __ballot_sync(0xffffffffu, mine);
                if (mine) rank[i] = run + __popc(bal & lt_mask);
                run += __popc(bal);
            }
            if (lane == 0) S.warp_hist[warp][dv] = run;
        }
    } else {
        // phase 1: all match.any back to back (their latency overlaps; nothing consumes a result yet)
        uint32_t info[SORT_ITEMS];   // peers mask, then leader lane | rank inside the digit group << 8
#pragma unroll
        for (int i = 0; i < SORT_ITEMS; ++i) {
            info[i] = ballot_match<RADIX_BITS>((uint32_t)(key[i] >> shift) & mask);
        }
        // phase 2: one running-count update per (item, digit group), in item order
#pragma unroll
        for (int i = 0; i < SORT_ITEMS; ++i) {
            const uint32_t d = (uint32_t)(key[i] >> shift) & mask;
            const uint32_t peers = info[i];
            const int leader = __ffs(peers) - 1;
            uint32_t pre = 0;
            // plain read-modify-write: the leaders of one item hold distinct digits and the table is private to
            // the warp (shared-memory atomics with a return value cost ~2 cycles per lane SM-wide: they were the
            // limiter of the pass); __syncwarp orders the store before the next item's load
            if (lane == leader) {
                pre = S.warp_hist[warp][d];
                S.warp_hist[warp][d] = pre + (uint32_t)__popc(peers);
            }
            rank[i] = pre;
            info[i] = (uint32_t)leader | ((uint32_t)__popc(peers & lt_mask) << 8);
            __syncwarp();
        }
#pragma unroll
        for (int i = 0; i < SORT_ITEMS; ++i)
            rank[i] = __shfl_sync(0xffffffffu, rank[i], (int)(info[i] & 31u)) + (info[i] >> 8);
    }
    __syncthreads();
    // ---- per-digit (threads 0..255): exclusive prefix over warps, tile total, look-back --------------
    uint32_t bin_total = 0, pub = 0, h = 0, hinc = 0, linc = 0;
    uint32_t* col = nullptr;
    uint32_t* my = nullptr;
    if (tid < RADIX) {
        const int d = tid;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w) {
            const uint32_t t = S.warp_hist[w][d];
            S.warp_hist[w][d] = bin_total;
            bin_total += t;
        }
        // padding keys (~0) of a partial tile landed in the top digit: not part of the data
        pub = bin_total;
        if (d == (int)mask) pub -= (uint32_t)(SORT_TILE - valid);
        col = sv.desc + d;   // descriptor of tile t for this digit: col[t * RADIX] (coalesced across d)
        my = col + (size_t)tile * RADIX;
        if (tile == 0) {
            st_volatile_u32(my, DESC_INC | pub);
        } else {
            st_volatile_u32(my, DESC_AGG | pub);
        }
        // global digit start = exclusive scan of the global histogram of this digit place
        h = (d <= (int)mask) ? sv.hist_pass[d] : 0u;
        hinc = h, linc = bin_total;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t1 = __shfl_up_sync(0xffffffffu, hinc, o);
            uint32_t t2 = __shfl_up_sync(0xffffffffu, linc, o);
            if (lane >= o) {
                hinc += t1;
                linc += t2;
            }
        }
        if (lane == 31) {
            S.s_h[warp] = hinc;
            S.s_l[warp] = linc;
        }
    }
    __syncthreads();
    if (tid < RADIX) {
        const int d = tid;
        uint32_t hoff = 0, loff = 0;
#pragma unroll
        for (int w = 0; w < RADIX / 32; ++w) {
            if (w < warp) {
                hoff += S.s_h[w];
                loff += S.s_l[w];
            }
        }
        const uint32_t digit_start = hoff + hinc - h;
        const uint32_t lexcl = loff + linc - bin_total;
        uint32_t excl = 0;
        if (tile > 0) {
            // Decoupled look-back, LOOK predecessor tiles per step: the loads of a window are independent,
            // so the walk advances LOOK tiles per L2 round trip instead of one.
            constexpr int LOOK = 8;
            int look = (int)tile - 1;
            bool found = false;
            while (!found) {
                uint32_t v[LOOK];
#pragma unroll
                for (int u = 0; u < LOOK; ++u)
                    v[u] = (look - u >= 0) ? ld_volatile_u32(col + (size_t)(look - u) * RADIX) : DESC_INC;
#pragma unroll
                for (int u = 0; u < LOOK; ++u) {
                    if (!found) {
                        uint32_t x = v[u];
                        while ((x >> 30) == 0) x = ld_volatile_u32(col + (size_t)(look - u) * RADIX);
                        excl += x & DESC_VAL;
                        found = (x >> 30) == 2;
                    }
                }
                look -= LOOK;
            }
            st_volatile_u32(my, DESC_INC | (excl + pub));
        }
        S.local_excl[d] = lexcl;
        S.bin_offset[d] = digit_start + excl - lexcl;
    }
    __syncthreads();
    // ---- scatter to shared memory at the tile-sorted position ------------------------------------
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        const uint32_t d = (uint32_t)(key[i] >> shift) & mask;
        rank[i] += S.local_excl[d] + S.warp_hist[warp][d];
        S.keys[rank[i]] = key[i];
    }
    if (HAS_VALS) {
#pragma unroll
        for (int i = 0; i < SORT_ITEMS; ++i) {
            const int li = wbase + i * 32;
            if (li < valid) S.vals[rank[i]] = __ldg(vals_in + tile_base + li);
        }
    }
    __syncthreads();
    // ---- coalesced write-out ------------------------------------------------------------------------
    uint64_t* __restrict__ keys_out = sv.keys_out;
    uint32_t* __restrict__ vals_out = sv.vals_out;
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        const int p = i * SORT_THREADS + tid;
        if (p < valid) {
            const uint64_t k = S.keys[p];
            const uint32_t d = (uint32_t)(k >> shift) & mask;
            const uint32_t dst = S.bin_offset[d] + (uint32_t)p;
            keys_out[dst] = k;
            if (HAS_VALS) vals_out[dst] = S.vals[p];
        }
    }
}

// --------------------------------------------------------------------------------------------
// K4w: stable partition of the pair words by tile id in ONE pass (tile ids of at most WIDE_BITS bits, i.e.
// up to 1024 tiles: 512 x 512 images).  Same structure as onesweep_pass_kernel with a 1024-bin digit; every
// thread owns WIDE_DPT bins (bin = tid + k * 256, so descriptor rows are read coalesced), whose look-backs run
// interleaved.  The global histogram of the digit is duplicateWithKeys' per-tile count.
// --------------------------------------------------------------------------------------------
constexpr int WIDE_DPT = WIDE_BINS / SORT_THREADS;
template <int ITEMS>
struct WideSmemT {
    uint64_t keys[SORT_THREADS * ITEMS];
    uint32_t warp_hist[SORT_WARPS][WIDE_BINS + 1];
    uint32_t a[WIDE_BINS];   // bin total of this tile -> exclusive offset of the bin inside the tile
    uint32_t b[WIDE_BINS];   // global bin count -> global bin start -> destination of the tile's first key of the bin
    uint32_t s_h[SORT_WARPS], s_l[SORT_WARPS];
    uint32_t tile;
    uint32_t pad[3];
};

template <int ITEMS>
__global__ void __launch_bounds__(SORT_THREADS, 2)
tile_partition_kernel(int n_bins, const __grid_constant__ SortTab tab) {
    constexpr int TILE = SORT_THREADS * ITEMS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WideSmemT<ITEMS>& S = *reinterpret_cast<WideSmemT<ITEMS>*>(smem_raw);
    const SortView& sv = tab.v[blockIdx.y];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (sv.overflow != nullptr && *sv.overflow != 0u) return;
    const int64_t n = sv.n_ptr ? (int64_t)min(*sv.n_ptr, tab.capacity) : (int64_t)sv.n_fixed;
    if ((int64_t)blockIdx.x * TILE >= n) return;
    if (tid == 0) S.tile = atomicAdd(sv.ticket, 1u);
    for (int i = tid; i < SORT_WARPS * (WIDE_BINS + 1); i += SORT_THREADS) (&S.warp_hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = S.tile;
    const int64_t tile_base = (int64_t)tile * TILE;
    const int valid = (int)min((int64_t)TILE, n - tile_base);
    constexpr uint32_t mask = WIDE_BINS - 1;
    const uint64_t* __restrict__ keys_in = sv.keys_in;

    uint64_t key[ITEMS];
    uint32_t rank[ITEMS];
    const int wbase = warp * (32 * ITEMS) + lane;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const int li = wbase + i * 32;
        key[i] = (li < valid) ? __ldg(keys_in + tile_base + li) : ~0ull;
    }
    // ---- warp-level stable ranking (ballot match, running counts by plain read-modify-write) --------
    const uint32_t lt_mask = (1u << lane) - 1u;
    {
        uint32_t info[ITEMS];
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            info[i] = ballot_match<WIDE_BITS>((uint32_t)(key[i] >> 32) & mask);
        }
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const uint32_t d = (uint32_t)(key[i] >> 32) & mask;
            const uint32_t peers = info[i];
            const int leader = __ffs(peers) - 1;
            uint32_t pre = 0;
            if (lane == leader) {
                pre = S.warp_hist[warp][d];
                S.warp_hist[warp][d] = pre + (uint32_t)__popc(peers);
            }
            rank[i] = pre;
            info[i] = (uint32_t)leader | ((uint32_t)__popc(peers & lt_mask) << 8);
            __syncwarp();
        }
#pragma unroll
        for (int i = 0; i < ITEMS; ++i)
            rank[i] = __shfl_sync(0xffffffffu, rank[i], (int)(info[i] & 31u)) + (info[i] >> 8);
    }
    __syncthreads();
    // ---- P1: per bin, exclusive prefix over the warps, tile total, publish the descriptor -----------
    uint32_t pub[WIDE_DPT];
#pragma unroll
    for (int k = 0; k < WIDE_DPT; ++k) {
        const int d = tid + k * SORT_THREADS;
        uint32_t bin_total = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w) {
            const uint32_t t = S.warp_hist[w][d];
            S.warp_hist[w][d] = bin_total;
            bin_total += t;
        }
        pub[k] = bin_total;
        if (d == (int)mask) pub[k] -= (uint32_t)(TILE - valid);   // padding keys (~0) of a partial tile
        st_volatile_u32(sv.desc + (size_t)tile * WIDE_BINS + d, (tile == 0 ? DESC_INC : DESC_AGG) | pub[k]);
        S.a[d] = bin_total;
        S.b[d] = d < n_bins ? sv.hist_pass[d] : 0u;
    }
    __syncthreads();
    // ---- P2: exclusive scans over the bins (thread t scans bins 4t..4t+3) ---------------------------
    {
        uint32_t la[WIDE_DPT], lb[WIDE_DPT], sa = 0, sb = 0;
#pragma unroll
        for (int k = 0; k < WIDE_DPT; ++k) {
            la[k] = sa, lb[k] = sb;
            sa += S.a[tid * WIDE_DPT + k];
            sb += S.b[tid * WIDE_DPT + k];
        }
        uint32_t ia = sa, ib = sb;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t1 = __shfl_up_sync(0xffffffffu, ia, o);
            const uint32_t t2 = __shfl_up_sync(0xffffffffu, ib, o);
            if (lane >= o) ia += t1, ib += t2;
        }
        if (lane == 31) S.s_l[warp] = ia, S.s_h[warp] = ib;
        __syncthreads();
        uint32_t oa = ia - sa, ob = ib - sb;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w)
            if (w < warp) oa += S.s_l[w], ob += S.s_h[w];
#pragma unroll
        for (int k = 0; k < WIDE_DPT; ++k) {
            S.a[tid * WIDE_DPT + k] = oa + la[k];
            S.b[tid * WIDE_DPT + k] = ob + lb[k];
        }
    }
    __syncthreads();
    // ---- P3: decoupled look-back of the thread's bins, interleaved; the loads of its first step are in flight while
    // the keys are scattered into shared memory (the scatter needs the tile-local offsets only) ---------------
    {
        uint32_t excl[WIDE_DPT];
        bool found[WIDE_DPT];
#pragma unroll
        for (int k = 0; k < WIDE_DPT; ++k) excl[k] = 0, found[k] = (tile == 0);
        constexpr int LOOK = 4;
        int look = (int)tile - 1;
        const uint32_t* col = sv.desc + tid;
        uint32_t v[WIDE_DPT][LOOK];
        auto issue = [&]() {
#pragma unroll
            for (int k = 0; k < WIDE_DPT; ++k)
#pragma unroll
                for (int u = 0; u < LOOK; ++u)
                    v[k][u] = (!found[k] && look - u >= 0)
                                  ? ld_volatile_u32(col + (size_t)(look - u) * WIDE_BINS + k * SORT_THREADS)
                                  : DESC_INC;
        };
        issue();
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const uint32_t d = (uint32_t)(key[i] >> 32) & mask;
            S.keys[rank[i] + S.a[d] + S.warp_hist[warp][d]] = key[i];
        }
        while (true) {
#pragma unroll
            for (int k = 0; k < WIDE_DPT; ++k)
#pragma unroll
                for (int u = 0; u < LOOK; ++u) {
                    if (!found[k]) {
                        uint32_t x = v[k][u];
                        while ((x >> 30) == 0)
                            x = ld_volatile_u32(col + (size_t)(look - u) * WIDE_BINS + k * SORT_THREADS);
                        excl[k] += x & DESC_VAL;
                        found[k] = (x >> 30) == 2;
                    }
                }
            if (found[0] && found[1] && found[2] && found[3]) break;
            look -= LOOK;
            issue();
        }
#pragma unroll
        for (int k = 0; k < WIDE_DPT; ++k) {
            const int d = tid + k * SORT_THREADS;
            if (tile > 0) st_volatile_u32(sv.desc + (size_t)tile * WIDE_BINS + d, DESC_INC | (excl[k] + pub[k]));
            S.b[d] = S.b[d] + excl[k] - S.a[d];
        }
    }
    __syncthreads();
    // ---- coalesced write-out ------------------------------------------------------------------------------
    uint64_t* __restrict__ keys_out = sv.keys_out;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const int p = i * SORT_THREADS + tid;
        if (p < valid) {
            const uint64_t k = S.keys[p];
            keys_out[S.b[(uint32_t)(k >> 32) & mask] + (uint32_t)p] = k;
        }
    }
}

// --------------------------------------------------------------------------------------------
// K4d: the same stable partition WITHOUT a look-back.  What a tile of the onesweep pass asks its predecessors -- how
// many words of each bin lie in front of it -- does not depend on the partition at all: scan + duplicateWithKeys emit
// the words and know where each one lands, so they count them per (partition tile, image tile) on the way
// (scan_duplicate_kernel, pt mode: the count matrix M[partition tile][1024] in the look-back descriptors' memory).
// partition_offsets_kernel turns M into exclusive prefixes along the partition tiles (PO_CHUNKS chunks of rows, the last
// CTA of a view chains the chunk totals and adds the bins' global starts; the per-tile totals it meets are
// duplicateWithKeys' tile_count), and a tile of tile_partition_direct_kernel reads its two rows of offsets and never
// waits for anybody.  In the look-back kernel 41 % of the instructions and a third of the stall samples were the
// look-back: with 296 tiles resident and 4 x 4 descriptors per thread and round trip, the first wave alone walks for
// ~25 us (profiles/r2_ncu_full_summary.md).
// --------------------------------------------------------------------------------------------
constexpr int PO_CHUNKS = 32;
struct OffsetsView {
    const uint32_t* n_ptr;       // live pair count
    const uint32_t* overflow;
    uint32_t* matrix;            // [tiles][1024] counts -> exclusive prefix inside the row's chunk
    uint32_t* chunk_base;        // [PO_CHUNKS][1024] chunk totals -> first destination of the chunk's words of a bin
    uint32_t* tile_count;        // [T] out: pairs per image tile
    uint32_t* ticket;
};
struct OffsetsTab {
    uint32_t capacity;
    uint32_t pt_words;
    OffsetsView v[MAX_VIEWS];
};
__device__ __forceinline__ void partition_rows(uint32_t n, uint32_t pt_words, int& tiles, int& rows_per_chunk) {
    tiles = (int)((n + pt_words - 1) / pt_words);
    rows_per_chunk = (tiles + PO_CHUNKS - 1) / PO_CHUNKS;
}
__global__ void __launch_bounds__(WIDE_BINS)
partition_offsets_kernel(int n_bins, const __grid_constant__ OffsetsTab tab) {
    const OffsetsView& ov = tab.v[blockIdx.y];
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_last;
    const int bin = threadIdx.x, lane = bin & 31, warp = bin >> 5;
    const bool overflow = *ov.overflow != 0u;
    const uint32_t n = overflow ? 0u : min(*ov.n_ptr, tab.capacity);
    int tiles, rpc;
    partition_rows(n, tab.pt_words, tiles, rpc);
    const int r0 = (int)blockIdx.x * rpc, r1 = min(r0 + rpc, tiles);
    uint32_t run = 0;
    constexpr int UN = 8;
    for (int r = r0; r < r1; r += UN) {
        uint32_t x[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) x[u] = (r + u < r1) ? ov.matrix[(size_t)(r + u) * WIDE_BINS + bin] : 0u;
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            if (r + u < r1) ov.matrix[(size_t)(r + u) * WIDE_BINS + bin] = run;
            run += x[u];
        }
    }
    __stcg(ov.chunk_base + (size_t)blockIdx.x * WIDE_BINS + bin, run);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(ov.ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // last CTA of the view: chunk totals -> exclusive prefix over the chunks, plus the bin's global start
    uint32_t tot[PO_CHUNKS];
    uint32_t total = 0;
#pragma unroll
    for (int c = 0; c < PO_CHUNKS; ++c) {
        tot[c] = __ldcg(ov.chunk_base + (size_t)c * WIDE_BINS + bin);
        total += tot[c];
    }
    if (bin < n_bins) ov.tile_count[bin] = total;
    uint32_t inc = total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t start = inc - total;
    for (int w = 0; w < warp; ++w) start += s_warp[w];
#pragma unroll
    for (int c = 0; c < PO_CHUNKS; ++c) {
        ov.chunk_base[(size_t)c * WIDE_BINS + bin] = start;
        start += tot[c];
    }
}

template <int ITEMS>
struct DirectSmemT {
    uint64_t keys[SORT_THREADS * ITEMS];
    uint32_t warp_hist[SORT_WARPS][WIDE_BINS + 1];
    uint32_t a[WIDE_BINS];   // bin total of this tile -> exclusive offset of the bin inside the tile
    uint32_t b[WIDE_BINS];   // destination of the tile's first word of the bin, minus a
    uint32_t s_l[SORT_WARPS];
};
template <int ITEMS>
__global__ void __launch_bounds__(SORT_THREADS, 2)
tile_partition_direct_kernel(const __grid_constant__ SortTab tab) {
    constexpr int TILE = SORT_THREADS * ITEMS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DirectSmemT<ITEMS>& S = *reinterpret_cast<DirectSmemT<ITEMS>*>(smem_raw);
    const SortView& sv = tab.v[blockIdx.y];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (*sv.overflow != 0u) return;
    const uint32_t n = min(*sv.n_ptr, tab.capacity);
    const uint32_t tile = blockIdx.x;
    if ((int64_t)tile * TILE >= (int64_t)n) return;
    int tiles, rpc;
    partition_rows(n, (uint32_t)TILE, tiles, rpc);
    const int64_t tile_base = (int64_t)tile * TILE;
    const int valid = (int)min((int64_t)TILE, (int64_t)n - tile_base);
    constexpr uint32_t mask = WIDE_BINS - 1;
    const uint64_t* __restrict__ keys_in = sv.keys_in;
    uint64_t key[ITEMS];
    uint32_t rank[ITEMS];
    const int wbase = warp * (32 * ITEMS) + lane;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const int li = wbase + i * 32;
        key[i] = (li < valid) ? __ldg(keys_in + tile_base + li) : ~0ull;
    }
    // the tile's offsets: its row of the matrix + its chunk's row (sv.desc = matrix, sv.hist_pass = chunk bases),
    // copied asynchronously into S.b / S.a by the thread that will combine them after the ranking
#pragma unroll
    for (int k = 0; k < WIDE_DPT; ++k) {
        const int d = tid + k * SORT_THREADS;
        cpa4(&S.b[d], sv.desc + (size_t)tile * WIDE_BINS + d);
        cpa4(&S.a[d], sv.hist_pass + (size_t)(tile / (uint32_t)rpc) * WIDE_BINS + d);
    }
    cpa_commit();
    for (int i = tid; i < SORT_WARPS * (WIDE_BINS + 1); i += SORT_THREADS) (&S.warp_hist[0][0])[i] = 0;
    __syncthreads();
    // ---- warp-level stable ranking (as tile_partition_kernel) ---------------------------------------
    const uint32_t lt_mask = (1u << lane) - 1u;
    {
        uint32_t info[ITEMS];
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) info[i] = ballot_match<WIDE_BITS>((uint32_t)(key[i] >> 32) & mask);
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const uint32_t d = (uint32_t)(key[i] >> 32) & mask;
            const uint32_t peers = info[i];
            const int leader = __ffs(peers) - 1;
            uint32_t pre = 0;
            if (lane == leader) {
                pre = S.warp_hist[warp][d];
                S.warp_hist[warp][d] = pre + (uint32_t)__popc(peers);
            }
            rank[i] = pre;
            info[i] = (uint32_t)leader | ((uint32_t)__popc(peers & lt_mask) << 8);
            __syncwarp();
        }
#pragma unroll
        for (int i = 0; i < ITEMS; ++i)
            rank[i] = __shfl_sync(0xffffffffu, rank[i], (int)(info[i] & 31u)) + (info[i] >> 8);
    }
    __syncthreads();
    // ---- per bin: exclusive prefix over the warps, tile total --------------------------------------
    cpa_wait<0>();   // this thread's offset copies (it reads only what it copied itself)
#pragma unroll
    for (int k = 0; k < WIDE_DPT; ++k) {
        const int d = tid + k * SORT_THREADS;
        uint32_t bin_total = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w) {
            const uint32_t t = S.warp_hist[w][d];
            S.warp_hist[w][d] = bin_total;
            bin_total += t;
        }
        S.b[d] += S.a[d];
        S.a[d] = bin_total;
    }
    __syncthreads();
    // ---- exclusive scan of the tile's bin totals (thread t scans bins 4t..4t+3) ---------------------
    {
        uint32_t la[WIDE_DPT], sa = 0;
#pragma unroll
        for (int k = 0; k < WIDE_DPT; ++k) {
            la[k] = sa;
            sa += S.a[tid * WIDE_DPT + k];
        }
        uint32_t ia = sa;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t1 = __shfl_up_sync(0xffffffffu, ia, o);
            if (lane >= o) ia += t1;
        }
        if (lane == 31) S.s_l[warp] = ia;
        __syncthreads();
        uint32_t oa = ia - sa;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w)
            if (w < warp) oa += S.s_l[w];
#pragma unroll
        for (int k = 0; k < WIDE_DPT; ++k) {
            const int d = tid * WIDE_DPT + k;
            const uint32_t ex = oa + la[k];
            S.a[d] = ex;
            S.b[d] -= ex;     // destination of local position p of bin d: b[d] + p
        }
    }
    __syncthreads();
    // ---- scatter through shared memory, coalesced write-out -----------------------------------------
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const uint32_t d = (uint32_t)(key[i] >> 32) & mask;
        S.keys[rank[i] + S.a[d] + S.warp_hist[warp][d]] = key[i];
    }
    __syncthreads();
    uint64_t* __restrict__ keys_out = sv.keys_out;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const int p = i * SORT_THREADS + tid;
        if (p < valid) {
            const uint64_t k = S.keys[p];
            keys_out[S.b[(uint32_t)(k >> 32) & mask] + (uint32_t)p] = k;
        }
    }
}

// workspace: hist [MAX_PASSES][256] u32 | tickets [64] u32 | desc [passes][tiles][256] u32
static int64_t sort_tiles_max(int64_t n) { return (n + 2047) / 2048; }
size_t sort_workspace_bytes(int64_t n) {
    const int64_t tiles = sort_tiles_max(n < 1 ? 1 : n);
    // (+ the chunk bases of the look-back-free partition behind its count matrix, which lives in the descriptors)
    return align_up((size_t)MAX_PASSES * RADIX * 4 + 256 + (size_t)MAX_PASSES * tiles * RADIX * 4 +
                    (size_t)PO_CHUNKS_WORDS * 4, 256);
}
size_t sort_workspace_zero_bytes(int64_t capacity, int end_bit) {
    const int passes = (end_bit + RADIX_BITS - 1) / RADIX_BITS;
    return (size_t)MAX_PASSES * RADIX * 4 + 256 + (size_t)passes * sort_tiles_max(capacity < 1 ? 1 : capacity) * RADIX * 4;
}
static bool wide_enabled() {
    static const bool on = [] {
        const char* e = getenv("B200SPLAT_WIDE_PARTITION");
        return !(e && e[0] == '0');
    }();
    return on;
}
int pair_sort_digit_passes(int end_bit) {
    return (end_bit <= WIDE_BITS && wide_enabled()) ? 0 : (end_bit + RADIX_BITS - 1) / RADIX_BITS;
}
int pair_sort_result_sel(int end_bit) {
    const int p = pair_sort_digit_passes(end_bit);
    return p == 0 ? 1 : (p & 1);
}
size_t pair_sort_zero_bytes(int64_t capacity, int end_bit) {
    if (pair_sort_digit_passes(end_bit) != 0) return sort_workspace_zero_bytes(capacity, end_bit);
    return (size_t)MAX_PASSES * RADIX * 4 + 256 + (size_t)sort_tiles_for(capacity < 1 ? 1 : capacity) * WIDE_BINS * 4;
}
void sort_workspace_views(void* ws, uint32_t** hist, uint32_t** tickets, uint32_t** desc) {
    uint32_t* h = reinterpret_cast<uint32_t*>(ws);
    *hist = h;
    *tickets = h + MAX_PASSES * RADIX;
    *desc = h + MAX_PASSES * RADIX + 64;
}

// keys-only passes do not touch SortSmem::vals: leave it out of the dynamic allocation (more CTAs per SM)

static cudaError_t ensure_sort_attr() {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(onesweep_pass_kernel<true, 3, 16>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SortSmem));
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(onesweep_pass_kernel<false, 3, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)sizeof(SortSmemT<16>));
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(onesweep_pass_kernel<false, 4, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)sizeof(SortSmemT<8>));
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(onesweep_pass_kernel<false, 2, 24>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)sizeof(SortSmemT<24>));
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(onesweep_pass_kernel<false, 2, 12, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)sizeof(SortSmemT<12, 512>));
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(tile_partition_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)sizeof(WideSmemT<8>));
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(tile_partition_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)sizeof(WideSmemT<16>));
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(tile_partition_kernel<24>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)sizeof(WideSmemT<24>));
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(tile_partition_direct_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)sizeof(DirectSmemT<8>));
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(tile_partition_direct_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)sizeof(DirectSmemT<16>));
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(tile_partition_direct_kernel<24>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)sizeof(DirectSmemT<24>));
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    return cudaSuccess;
}

cudaError_t launch_sort_pairs(int64_t n, int end_bit, uint64_t* keys[2], uint32_t* vals[2], void* ws, int* sel,
                              cudaStream_t st) {
    *sel = 0;
    if (n <= 0) return cudaSuccess;
    if (end_bit < 1) end_bit = 1;
    if (end_bit > 64) end_bit = 64;
    const int passes = (end_bit + RADIX_BITS - 1) / RADIX_BITS;
    const int tiles = sort_tiles_pairs(n);
    cudaError_t e = ensure_sort_attr();
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(ws, 0, sort_workspace_zero_bytes(n, end_bit), st);
    if (e != cudaSuccess) return e;
    uint32_t *hist, *tickets, *desc;
    sort_workspace_views(ws, &hist, &tickets, &desc);
    int64_t hg = (n + 255) / 256;
    int hgrid = (int)(hg < (int64_t)NUM_SMS * 8 ? hg : (int64_t)NUM_SMS * 8);
    HistTab ht;
    ht.keys[0] = keys[0];
    ht.hist[0] = hist;
    radix_histogram_kernel<0><<<dim3(hgrid, 1), 256, 0, st>>>(n, passes, end_bit, 0, ht);
    count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    int cur = 0;
    for (int p = 0; p < passes; ++p) {
        const int shift = p * RADIX_BITS;
        const int bits = (end_bit - shift) < RADIX_BITS ? (end_bit - shift) : RADIX_BITS;
        SortTab t;
        t.capacity = (uint32_t)n;
        t.v[0] = SortView{nullptr, (uint32_t)n, nullptr, keys[cur], vals[cur], keys[cur ^ 1], vals[cur ^ 1],
                          hist + p * RADIX, tickets + p, desc + (size_t)p * tiles * RADIX};
        onesweep_pass_kernel<true, 3, 16><<<dim3(tiles, 1), SORT_THREADS, sizeof(SortSmem), st>>>(shift, bits, t);
        count_launch();
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        cur ^= 1;
    }
    *sel = cur;
    return cudaSuccess;
}

// keys-only pass over all views of the batch with the configured tile size
static cudaError_t launch_keys_pass(int shift, int bits, const SortTab& t, int tiles, int V, cudaStream_t st,
                                    int items = 0) {
    const dim3 grid(tiles, V);
    switch (items ? items : sort_items()) {
        case 8:
            onesweep_pass_kernel<false, 4, 8><<<grid, SORT_THREADS, offsetof(SortSmemT<8>, vals), st>>>(shift, bits, t);
            break;
        case 16:
            onesweep_pass_kernel<false, 3, 16><<<grid, SORT_THREADS, offsetof(SortSmemT<16>, vals), st>>>(shift, bits, t);
            break;
        case 512: {
            using SM = SortSmemT<12, 512>;
            onesweep_pass_kernel<false, 2, 12, 512><<<grid, 512, offsetof(SM, vals), st>>>(shift, bits, t);
            break;
        }
        default:
            onesweep_pass_kernel<false, 2, 24><<<grid, SORT_THREADS, offsetof(SortSmemT<24>, vals), st>>>(shift, bits, t);
    }
    count_launch();
    return cudaGetLastError();
}

// Stable sort of every view's P Gaussian words (depth_bits << 32 | index) on the depth bits: 4 passes over P
// elements (instead of 6 passes over ~3.6 P pair keys).  Result in gwords[0].
cudaError_t launch_gaussian_sort(const BatchTab& tab, cudaStream_t st, bool cleared) {
    if (tab.P <= 0) return cudaSuccess;
    cudaError_t e = ensure_sort_attr();
    if (e != cudaSuccess) return e;
    const int64_t n = tab.P;
    const int items = gsort_items();
    const int tiles = (int)((n + (int64_t)SORT_THREADS * items - 1) / ((int64_t)SORT_THREADS * items));
    HistTab ht;
    for (int v = 0; v < tab.V; ++v) {
        if (!cleared) {
            e = cudaMemsetAsync(tab.v[v].ghist, 0, sort_workspace_zero_bytes(n, 32), st);
            if (e != cudaSuccess) return e;
        }
        ht.keys[v] = tab.v[v].gwords[0];
        ht.hist[v] = tab.v[v].ghist;
    }
    int64_t hg = (n + 255) / 256;
    const int hgrid = (int)(hg < (int64_t)NUM_SMS * 8 / tab.V ? hg : (int64_t)NUM_SMS * 8 / tab.V);
    radix_histogram_kernel<4><<<dim3(hgrid > 0 ? hgrid : 1, tab.V), 256, 0, st>>>(n, 4, 32, 32, ht);
    count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    int cur = 0;
    for (int p = 0; p < 4; ++p) {
        SortTab t;
        t.capacity = (uint32_t)n;
        for (int v = 0; v < tab.V; ++v) {
            const ViewTab& vt = tab.v[v];
            t.v[v] = SortView{nullptr, (uint32_t)n, nullptr, vt.gwords[cur], nullptr, vt.gwords[cur ^ 1], nullptr,
                              vt.ghist + p * RADIX, vt.gtickets + p, vt.gdesc + (size_t)p * tiles * RADIX};
        }
        e = launch_keys_pass(32 + 8 * p, 8, t, tiles, tab.V, st, items);
        if (e != cudaSuccess) return e;
        cur ^= 1;
    }
    return cudaSuccess;   // 4 passes: back in gwords[0]
}

// Look-back-free partition only (tab.pt_words > 0): the offsets of every partition tile and the per-tile pair counts
// (tile_count: what the tile ranges are scanned from) out of the count matrix of scan + duplicateWithKeys.
cudaError_t launch_partition_offsets(const BatchTab& tab, cudaStream_t st) {
    if (tab.P <= 0 || tab.capacity == 0 || tab.pt_words <= 0) return cudaSuccess;
    const int tiles = sort_tiles_for(tab.capacity);
    OffsetsTab ot;
    ot.capacity = tab.capacity;
    ot.pt_words = (uint32_t)tab.pt_words;
    for (int v = 0; v < tab.V; ++v) {
        const ViewTab& vt = tab.v[v];
        ot.v[v] = OffsetsView{vt.point_offsets + (tab.P - 1), vt.status + STATUS_OVERFLOW, vt.desc,
                              vt.desc + (size_t)tiles * WIDE_BINS, vt.tile_count, vt.tickets + 1};
    }
    partition_offsets_kernel<<<dim3(PO_CHUNKS, tab.V), WIDE_BINS, 0, st>>>(tab.grid_x * tab.grid_y, ot);
    count_launch();
    return cudaGetLastError();
}

// Stable sort of every view's pair words (tile << 32 | index) on the tile bits [32, 32 + end_bit); histograms
// come from duplicateWithKeys, the pair count from the device.  Result in keys[passes & 1].
cudaError_t launch_sort_batch(const BatchTab& tab, cudaStream_t st) {
    if (tab.P <= 0 || tab.capacity == 0) return cudaSuccess;
    const int end_bit = tab.end_bit;
    const int passes = tab.digit_passes;
    const int tiles = sort_tiles_for(tab.capacity);
    cudaError_t e = ensure_sort_attr();
    if (e != cudaSuccess) return e;
    if (passes == 0 && tab.pt_words > 0) {   // look-back-free partition (K4d): offsets from the count matrix
        SortTab t;                           // (launch_partition_offsets has run: it also produces tile_count)
        t.capacity = tab.capacity;
        for (int v = 0; v < tab.V; ++v) {
            const ViewTab& vt = tab.v[v];
            uint32_t* chunk_base = vt.desc + (size_t)tiles * WIDE_BINS;
            t.v[v] = SortView{vt.point_offsets + (tab.P - 1), 0u, vt.status + STATUS_OVERFLOW, vt.keys[0], nullptr,
                              vt.keys[1], nullptr, chunk_base, vt.tickets, vt.desc};
        }
        const dim3 grid(tiles, tab.V);
        switch (tab.pt_words / SORT_THREADS) {
            case 8:
                tile_partition_direct_kernel<8><<<grid, SORT_THREADS, sizeof(DirectSmemT<8>), st>>>(t);
                break;
            case 16:
                tile_partition_direct_kernel<16><<<grid, SORT_THREADS, sizeof(DirectSmemT<16>), st>>>(t);
                break;
            default:
                tile_partition_direct_kernel<24><<<grid, SORT_THREADS, sizeof(DirectSmemT<24>), st>>>(t);
        }
        count_launch();
        return cudaGetLastError();
    }
    if (passes == 0) {   // one wide pass: bins = tiles, global histogram = the per-tile counts
        SortTab t;
        t.capacity = tab.capacity;
        for (int v = 0; v < tab.V; ++v) {
            const ViewTab& vt = tab.v[v];
            t.v[v] = SortView{vt.point_offsets + (tab.P - 1), 0u, vt.status + STATUS_OVERFLOW, vt.keys[0], nullptr,
                              vt.keys[1], nullptr, vt.tile_count, vt.tickets, vt.desc};
        }
        const dim3 grid(tiles, tab.V);
        const int n_bins = tab.grid_x * tab.grid_y;
        switch (sort_items()) {
            case 8:
                tile_partition_kernel<8><<<grid, SORT_THREADS, sizeof(WideSmemT<8>), st>>>(n_bins, t);
                break;
            case 16:
                tile_partition_kernel<16><<<grid, SORT_THREADS, sizeof(WideSmemT<16>), st>>>(n_bins, t);
                break;
            default:
                tile_partition_kernel<24><<<grid, SORT_THREADS, sizeof(WideSmemT<24>), st>>>(n_bins, t);
        }
        count_launch();
        return cudaGetLastError();
    }
    int cur = 0;
    for (int p = 0; p < passes; ++p) {
        const int shift = p * RADIX_BITS;
        const int bits = (end_bit - shift) < RADIX_BITS ? (end_bit - shift) : RADIX_BITS;
        SortTab t;
        t.capacity = tab.capacity;
        for (int v = 0; v < tab.V; ++v) {
            const ViewTab& vt = tab.v[v];
            t.v[v] = SortView{vt.point_offsets + (tab.P - 1), 0u, vt.status + STATUS_OVERFLOW, vt.keys[cur],
                              nullptr, vt.keys[cur ^ 1], nullptr, vt.hist + p * RADIX, vt.tickets + p,
                              vt.desc + (size_t)p * tiles * RADIX};
        }
        e = launch_keys_pass(32 + shift, bits, t, tiles, tab.V, st);
        if (e != cudaSuccess) return e;
        cur ^= 1;
    }
    return cudaSuccess;
}

// ============================================================================================
// K5: tile ranges from the sorted keys, and the longest-list-first tile order of the batch
// ============================================================================================
__global__ void __launch_bounds__(256)
tile_ranges_kernel(const __grid_constant__ BatchTab tab, int sel) {
    const ViewTab& vt = tab.v[blockIdx.y];
    if (vt.status[STATUS_OVERFLOW] != 0u) return;
    const int64_t R = (int64_t)min(vt.point_offsets[tab.P - 1], tab.capacity);
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R) return;
    const uint64_t* __restrict__ keys = vt.keys[sel];
    uint32_t* __restrict__ ranges = vt.ranges;
    const int tshift = 32;
    const uint32_t t = (uint32_t)(keys[i] >> tshift);
    if (i == 0) {
        ranges[2 * t] = 0;
    } else {
        const uint32_t tp = (uint32_t)(keys[i - 1] >> tshift);
        if (tp != t) {
            ranges[2 * tp + 1] = (uint32_t)i;
            ranges[2 * t] = (uint32_t)i;
        }
    }
    if (i == R - 1) ranges[2 * t + 1] = (uint32_t)R;
}

// order[] = the batch's (view * T + tile) entries by decreasing Gaussian-list length (LPT scheduling of the
// render CTAs).  One CTA, counting sort on a monotone 11-bit key of the length (float exponent + 3 mantissa
// bits: 8 buckets per octave, common.cuh order_bucket) -- the order inside a bucket is irrelevant for load balance.
// Also zeroes the walk-length histogram that render forward fills for the block order of render backward.
__global__ void __launch_bounds__(1024)
tile_order_kernel(const __grid_constant__ BatchTab tab, int n) {
    __shared__ uint32_t s_cnt[ORDER_BUCKETS];
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_empty;      // cursor of the empty tiles (they all go to the end of the order)
    const int T = tab.grid_x * tab.grid_y;
    for (int i = threadIdx.x; i < ORDER_BUCKETS; i += blockDim.x) s_cnt[i] = 0, tab.block_hist[i] = 0;
    if (threadIdx.x == 0) s_empty = 0;
    __syncthreads();
    auto bucket_of = [&](int i) -> uint32_t {   // 0xffffffff: empty tile
        const uint32_t* r = tab.v[i / T].ranges + 2 * (i % T);
        const uint32_t len = r[1] - r[0];
        if (len == 0u) return 0xffffffffu;
        // descending: longest lists get the smallest bucket index
        return (ORDER_BUCKETS - 1) - min((uint32_t)(ORDER_BUCKETS - 2), __float_as_uint((float)len) >> 20);
    };
    // Most tiles of a sparse image are empty: they are counted per warp (one shared atomic per warp), the others
    // with one plain shared atomic each (their buckets are spread).  4 independent loads in flight per thread: a
    // single CTA is latency-bound.
    constexpr int UN = 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_round = (n + UN * 1024 - 1) / (UN * 1024) * (UN * 1024);   // whole warps take part in the votes
    uint32_t empties = 0;
    for (int i0 = threadIdx.x; i0 < n_round; i0 += UN * blockDim.x) {
        uint32_t b[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int i = i0 + u * blockDim.x;
            b[u] = i < n ? bucket_of(i) : 0xfffffffeu;   // 0xfffffffe: beyond the end
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            if (b[u] < ORDER_BUCKETS) atomicAdd(&s_cnt[b[u]], 1u);
            empties += __popc(__ballot_sync(0xffffffffu, b[u] == 0xffffffffu));
        }
    }
    if (lane == 0 && empties) atomicAdd(&s_empty, empties);
    __syncthreads();
    // exclusive scan of the bucket counts (2 buckets per thread)
    const uint32_t c0 = s_cnt[2 * threadIdx.x], c1 = s_cnt[2 * threadIdx.x + 1];
    uint32_t inc = c0 + c1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t woff = 0;
    for (int w = 0; w < warp; ++w) woff += s_warp[w];
    const uint32_t excl = woff + inc - (c0 + c1);
    __syncthreads();
    s_cnt[2 * threadIdx.x] = excl;
    s_cnt[2 * threadIdx.x + 1] = excl + c0;
    if (threadIdx.x == 0) s_empty = (uint32_t)n - s_empty;   // the empty tiles start behind all the others
    __syncthreads();
    for (int i0 = threadIdx.x; i0 < n_round; i0 += UN * blockDim.x) {
        uint32_t b[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int i = i0 + u * blockDim.x;
            b[u] = i < n ? bucket_of(i) : 0xfffffffeu;
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int i = i0 + u * blockDim.x;
            if (b[u] < ORDER_BUCKETS) tab.tile_order[atomicAdd(&s_cnt[b[u]], 1u)] = (uint32_t)i;
            const uint32_t em = __ballot_sync(0xffffffffu, b[u] == 0xffffffffu);
            if (em) {
                uint32_t base = 0;
                const int leader = __ffs(em) - 1;
                if (lane == leader) base = atomicAdd(&s_empty, (uint32_t)__popc(em));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (b[u] == 0xffffffffu) tab.tile_order[base + __popc(em & ((1u << lane) - 1u))] = (uint32_t)i;
            }
        }
    }
}

// ranges[t] = exclusive scan of the per-tile pair counts accumulated by duplicateWithKeys (== positions of the
// tile's run in the sorted array); untouched tiles stay (0,0).  One CTA per view.
__global__ void __launch_bounds__(1024)
tile_ranges_from_counts_kernel(const __grid_constant__ BatchTab tab) {
    const ViewTab& vt = tab.v[blockIdx.x];
    const int T = tab.grid_x * tab.grid_y;
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const bool overflow = vt.status[STATUS_OVERFLOW] != 0u;
    for (int base = 0; base < T; base += 1024) {
        const int t = base + threadIdx.x;
        const uint32_t c = (t < T && !overflow) ? vt.tile_count[t] : 0u;
        uint32_t inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t x = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += x;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        uint32_t woff = s_carry;
        for (int w = 0; w < warp; ++w) woff += s_warp[w];
        const uint32_t start = woff + inc - c;
        if (t < T) {
            vt.ranges[2 * t] = c ? start : 0u;
            vt.ranges[2 * t + 1] = c ? start + c : 0u;
        }
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = start + c;
        __syncthreads();
    }
}

// Work order of render backward: the batch's 8x4 pixel blocks that have anything to walk (largest n_contrib of the
// block > 0), most entries first (LPT over the SMs at the granularity the backward actually works at: a silhouette
// block walks ten times more entries than its neighbours in the same tile).  block_order[0] = count, entries follow
// at [4..).  A counting sort on the monotone 11-bit key of the walk length whose COUNTING half is done by render
// forward itself: each of its warps adds its block to block_hist[bucket] with one global atomic and keeps the returned
// rank (block_code = bucket << 20 | rank).  What is left is this scatter: every CTA scans the 2048 bucket counts and
// places its 1024 blocks at bucket start + rank.  (The single-CTA kernel that did both halves took 33 us per step
// between render forward and render backward; this one takes ~3.)
__global__ void __launch_bounds__(1024)
block_scatter_kernel(const __grid_constant__ BatchTab tab, int n) {
    __shared__ uint32_t s_start[ORDER_BUCKETS];
    __shared__ uint32_t s_warp[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    const uint32_t code = i < n ? tab.block_code[i] : BLOCK_CODE_NONE;     // in flight during the scan
    const uint32_t c0 = tab.block_hist[2 * threadIdx.x], c1 = tab.block_hist[2 * threadIdx.x + 1];
    uint32_t inc = c0 + c1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t woff = 0;
    for (int w = 0; w < warp; ++w) woff += s_warp[w];
    const uint32_t excl = woff + inc - (c0 + c1);
    if (blockIdx.x == 0 && threadIdx.x == 1023) tab.block_order[0] = excl + c0 + c1;   // number of non-empty blocks
    s_start[2 * threadIdx.x] = excl;
    s_start[2 * threadIdx.x + 1] = excl + c0;
    __syncthreads();
    if (code != BLOCK_CODE_NONE) tab.block_order[4 + s_start[code >> 20] + (code & 0xfffffu)] = (uint32_t)i;
}

cudaError_t launch_block_order(const BatchTab& tab, cudaStream_t st) {
    const int n = tab.V * tab.grid_x * tab.grid_y * (BLOCK_SIZE / 32);
    block_scatter_kernel<<<(n + 1023) / 1024, 1024, 0, st>>>(tab, n);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_tile_ranges_batch(const BatchTab& tab, int sel, cudaStream_t st) {
    const int T = tab.grid_x * tab.grid_y;
    if (tab.P > 0 && tab.capacity > 0 && T <= 8192) {
        tile_ranges_from_counts_kernel<<<tab.V, 1024, 0, st>>>(tab);
        count_launch();
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    } else {
        for (int v = 0; v < tab.V; ++v) {
            cudaError_t e = cudaMemsetAsync(tab.v[v].ranges, 0, (size_t)T * 8, st);
            if (e != cudaSuccess) return e;
        }
        if (tab.P > 0 && tab.capacity > 0) {
            const unsigned gx = (unsigned)(((int64_t)tab.capacity + 255) / 256);
            tile_ranges_kernel<<<dim3(gx, tab.V), 256, 0, st>>>(tab, sel);
            count_launch();
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) return e;
        }
    }
    const int n = tab.V * T;
    tile_order_kernel<<<1, 1024, 0, st>>>(tab, n);
    count_launch();
    return cudaGetLastError();
}

}  // namespace b200splat

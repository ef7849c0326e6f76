// LSD radix sort of (u64 key, u32 value) pairs, 8-bit digits, one HBM round trip per digit.
//
// This is the workhorse of the suffix sorter (replaces the induced-sorting loops of
// algo/string/SAIS.java:103-137 by a data-parallel sort; only the resulting order is shared).
//
// Per sort:   one histogram launch (all digits at once, 8 B/key read),
//             one tiny scan launch (digit bases),
// per digit:  one "onesweep" launch: every CTA takes the next tile by ticket, ranks its keys with
//             warp match/ballot into per-warp digit counters, learns the tile's global offsets through a
//             decoupled look-back over 64-bit status words (aggregate | inclusive-prefix flags), reorders
//             the tile in shared memory and writes digit runs out coalesced.
// Algorithmic traffic per digit pass: read 12 B + write 12 B per pair (8 + 8 for keys only).
#include "radix_sort.cuh"

#include <algorithm>

namespace gcz {

namespace {

constexpr int kRadix = 256;
constexpr int kHistThreads = 512;

constexpr unsigned long long kFlagAgg = 1ull << 62;
constexpr unsigned long long kFlagPrefix = 2ull << 62;
constexpr unsigned long long kValueMask = (1ull << 62) - 1;

// ---- histogram of every digit in one pass -------------------------------------------------------
__global__ void __launch_bounds__(kHistThreads)
radix_hist_kernel(const uint64_t* __restrict__ keys, int64_t n, int begin_bit, int npass,
                  unsigned long long* __restrict__ hist /* [npass][256] */) {
    __shared__ unsigned s_hist[8 * kRadix];
    for (int i = threadIdx.x; i < npass * kRadix; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t k = keys[i] >> begin_bit;
#pragma unroll 8
        for (int p = 0; p < npass; p++) atomicAdd(&s_hist[p * kRadix + (int)((k >> (8 * p)) & 255)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < npass * kRadix; i += blockDim.x) {
        const unsigned v = s_hist[i];
        if (v) atomicAdd(&hist[i], (unsigned long long)v);
    }
}

// exclusive scan of each pass's 256 bins, in place
__global__ void radix_scan_kernel(unsigned long long* hist, int npass) {
    __shared__ unsigned long long s_warp[8];
    for (int p = 0; p < npass; p++) {
        unsigned long long v = hist[p * kRadix + threadIdx.x];
        unsigned long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += t;
        }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
        __syncthreads();
        unsigned long long base = 0;
        for (int w = 0; w < (int)(threadIdx.x >> 5); w++) base += s_warp[w];
        hist[p * kRadix + threadIdx.x] = base + incl - v;
        __syncthreads();
    }
}

// ---- one digit pass --------------------------------------------------------------------------------
template <int THREADS, int ITEMS, bool HAS_VALS>
__global__ void __launch_bounds__(THREADS, 2)
onesweep_kernel(const uint64_t* __restrict__ keys_in, uint64_t* __restrict__ keys_out,
                const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ vals_out,
                int64_t n, int shift, const unsigned long long* __restrict__ digit_base,
                unsigned long long* __restrict__ status, unsigned* __restrict__ ticket) {
    constexpr int TILE = THREADS * ITEMS;
    constexpr int WARPS = THREADS / 32;
    static_assert(THREADS >= kRadix, "one thread per digit is needed for the look-back");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* s_keys = reinterpret_cast<uint64_t*>(smem_raw);                       // TILE
    uint32_t* s_vals = reinterpret_cast<uint32_t*>(s_keys + TILE);                  // TILE (if HAS_VALS)
    unsigned* s_warp_hist = s_vals + (HAS_VALS ? TILE : 0);                         // WARPS x 256
    long long* s_gofs = reinterpret_cast<long long*>(s_warp_hist + WARPS * kRadix); // 256: global base - tile start
    unsigned* s_digit_start = reinterpret_cast<unsigned*>(s_gofs + kRadix);         // 256
    unsigned* s_scan = s_digit_start + kRadix;                                      // 8 warp totals
    __shared__ unsigned s_tile;

    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    for (int i = threadIdx.x; i < WARPS * kRadix; i += THREADS) s_warp_hist[i] = 0;
    __syncthreads();
    const unsigned tile = s_tile;
    const int64_t tile_base = (int64_t)tile * TILE;
    const int count = (int)min((int64_t)TILE, n - tile_base);
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned lt = lanemask_lt();

    // warp-striped load: item i of lane l in warp w is element w*ITEMS*32 + i*32 + l of the tile
    uint64_t key[ITEMS];
    uint32_t val[ITEMS];
    const int warp_base = warp * ITEMS * 32;
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        const int e = warp_base + i * 32 + lane;
        key[i] = e < count ? keys_in[tile_base + e] : ~0ull;
        if (HAS_VALS) val[i] = e < count ? vals_in[tile_base + e] : 0u;
    }

    // rank inside the warp: peers with the same digit get consecutive ranks in element order
    unsigned short rank[ITEMS];
    unsigned* my_hist = s_warp_hist + warp * kRadix;
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        const unsigned d = (unsigned)(key[i] >> shift) & 255u;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const unsigned below = __popc(peers & lt);
        unsigned base = 0;
        if (below == 0) {
            base = my_hist[d];
            my_hist[d] = base + __popc(peers);
        }
        base = __shfl_sync(0xffffffffu, base, __ffs(peers) - 1);
        rank[i] = (unsigned short)(base + below);
        __syncwarp();
    }
    __syncthreads();

    // digit d (thread d): turn per-warp counts into exclusive warp offsets, get the tile total
    unsigned total = 0;
    if (threadIdx.x < kRadix) {
#pragma unroll
        for (int w = 0; w < WARPS; w++) {
            const unsigned c = s_warp_hist[w * kRadix + threadIdx.x];
            s_warp_hist[w * kRadix + threadIdx.x] = total;
            total += c;
        }
        if (threadIdx.x == kRadix - 1) total -= (unsigned)(TILE - count);     // padding keys are all digit 255
        // publish the aggregate as early as possible
        st_relaxed_u64(&status[(size_t)tile * kRadix + threadIdx.x],
                       (unsigned long long)total | (tile == 0 ? kFlagPrefix : kFlagAgg));
    }
    // exclusive scan of the 256 totals -> start of every digit inside the tile
    if (threadIdx.x < kRadix) {
        const unsigned incl = warp_incl_sum(total);
        if (lane == 31) s_scan[warp] = incl;
        s_digit_start[threadIdx.x] = incl - total;
    }
    __syncthreads();
    if (threadIdx.x < kRadix) {
        unsigned base = 0;
        for (unsigned w = 0; w < warp; w++) base += s_scan[w];
        const unsigned start = s_digit_start[threadIdx.x] + base;
        s_digit_start[threadIdx.x] = start;
        // decoupled look-back: sum aggregates of predecessor tiles until an inclusive prefix shows up
        unsigned long long excl = 0;
        if (tile > 0) {
            long long t = (long long)tile - 1;
            while (true) {
                const unsigned long long v = ld_relaxed_u64(&status[(size_t)t * kRadix + threadIdx.x]);
                const unsigned long long flag = v & ~kValueMask;
                if (flag == 0) continue;
                excl += v & kValueMask;
                if (flag == kFlagPrefix) break;
                t--;
            }
            st_relaxed_u64(&status[(size_t)tile * kRadix + threadIdx.x], (excl + total) | kFlagPrefix);
        }
        s_gofs[threadIdx.x] = (long long)(digit_base[threadIdx.x] + excl) - (long long)start;
    }
    __syncthreads();

    // reorder the tile in shared memory
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        const unsigned d = (unsigned)(key[i] >> shift) & 255u;
        const unsigned pos = s_digit_start[d] + my_hist[d] + rank[i];
        s_keys[pos] = key[i];
        if (HAS_VALS) s_vals[pos] = val[i];
    }
    __syncthreads();

    // digit runs are contiguous both in shared memory and at their destination
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        const int j = i * THREADS + threadIdx.x;
        if (j < count) {
            const uint64_t k = s_keys[j];
            const long long dst = s_gofs[(unsigned)(k >> shift) & 255u] + j;
            keys_out[dst] = k;
            if (HAS_VALS) vals_out[dst] = s_vals[j];
        }
    }
}

constexpr int kSortThreads = 384;
constexpr int kSortItems = 16;
constexpr int kSortTile = kSortThreads * kSortItems;

template <bool HAS_VALS>
constexpr size_t onesweep_smem() {
    return (size_t)kSortTile * 8 + (HAS_VALS ? (size_t)kSortTile * 4 : 0) + (size_t)(kSortThreads / 32) * kRadix * 4 +
           kRadix * 8 + kRadix * 4 + 64;
}

}  // namespace

size_t radix_sort_temp_bytes(int64_t n) {
    const int64_t tiles = (n + kSortTile - 1) / kSortTile;
    // [8][256] histogram + per-pass (status[tiles][256] + ticket)
    return 8 * kRadix * 8 + 256 + ((size_t)tiles * kRadix * 8 + 256);
}

int radix_sort_pairs(DeviceCtx* ctx, cudaStream_t st, RadixBuffers& b, int64_t n, int begin_bit, int end_bit,
                     void* temp, SortStats* stats) {
    if (n <= 0 || end_bit <= begin_bit) return GCZ_OK;
    if (end_bit - begin_bit > 64 || begin_bit < 0) return fail(GCZ_E_ARG, "radix sort bit range");
    const int npass = (end_bit - begin_bit + 7) / 8;
    const bool has_vals = b.vals[0] != nullptr;
    if (!ctx->sort_attr[has_vals ? 1 : 0]) {
        if (has_vals)
            GCZ_CUDA(cudaFuncSetAttribute(onesweep_kernel<kSortThreads, kSortItems, true>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)onesweep_smem<true>()));
        else
            GCZ_CUDA(cudaFuncSetAttribute(onesweep_kernel<kSortThreads, kSortItems, false>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)onesweep_smem<false>()));
        ctx->sort_attr[has_vals ? 1 : 0] = true;
    }
    auto* hist = static_cast<unsigned long long*>(temp);
    auto* status = hist + 8 * kRadix + 32;
    const int64_t tiles = (n + kSortTile - 1) / kSortTile;
    auto* ticket = reinterpret_cast<unsigned*>(status + (size_t)tiles * kRadix);

    GCZ_CUDA(cudaMemsetAsync(hist, 0, 8 * kRadix * 8, st));
    const int hist_grid = (int)std::min<int64_t>((n + kHistThreads * 8 - 1) / (kHistThreads * 8), (int64_t)ctx->sm_count * 4);
    GCZ_LAUNCH(ctx, radix_hist_kernel, hist_grid, kHistThreads, 0, st, b.keys[b.cur], n, begin_bit, npass, hist);
    GCZ_LAUNCH(ctx, radix_scan_kernel, 1, kRadix, 0, st, hist, npass);

    for (int p = 0; p < npass; p++) {
        GCZ_CUDA(cudaMemsetAsync(status, 0, (size_t)tiles * kRadix * 8 + 64, st));
        const int in = b.cur, out = b.cur ^ 1;
        const int shift = begin_bit + 8 * p;
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (stats) {
            GCZ_CUDA(cudaEventCreate(&e0)); GCZ_CUDA(cudaEventCreate(&e1));
            GCZ_CUDA(cudaEventRecord(e0, st));
        }
        if (has_vals) {
            GCZ_LAUNCH(ctx, (onesweep_kernel<kSortThreads, kSortItems, true>), (unsigned)tiles, kSortThreads,
                       onesweep_smem<true>(), st, b.keys[in], b.keys[out], b.vals[in], b.vals[out], n, shift,
                       hist + p * kRadix, status, ticket);
        } else {
            GCZ_LAUNCH(ctx, (onesweep_kernel<kSortThreads, kSortItems, false>), (unsigned)tiles, kSortThreads,
                       onesweep_smem<false>(), st, b.keys[in], b.keys[out], nullptr, nullptr, n, shift,
                       hist + p * kRadix, status, ticket);
        }
        b.cur = out;
        if (stats) {
            GCZ_CUDA(cudaEventRecord(e1, st));
            stats->events.push_back(e0); stats->events.push_back(e1);
            stats->passes++; stats->elements += n;
        }
    }
    return GCZ_OK;
}

void SortStats::resolve() {
    for (size_t i = 0; i + 1 < events.size(); i += 2) {
        float t = 0;
        if (cudaEventElapsedTime(&t, events[i], events[i + 1]) == cudaSuccess) ms += t;
        cudaEventDestroy(events[i]); cudaEventDestroy(events[i + 1]);
    }
    events.clear();
}

}  // namespace gcz

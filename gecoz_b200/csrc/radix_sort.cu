// LSD radix sort of (u64 key, u32 value) pairs, 8-bit digits, one HBM round trip per digit.
//
// This is the workhorse of the suffix sorter (replaces the induced-sorting loops of
// algo/string/SAIS.java:103-137 by a data-parallel sort; only the resulting order is shared).
//
// Per sort:   one histogram launch (all digits at once, 8 B/key read),
//             one tiny scan launch (digit bases),
// per digit:  one "onesweep" launch: every CTA takes tiles by ticket, ranks their keys with warp ballots into
//             per-warp digit counters, learns a tile's global offsets through a decoupled look-back over
//             64-bit status words (aggregate | inclusive-prefix flags), reorders the tile in shared memory
//             and writes digit runs out coalesced.  Pairs take the persistent kernel whose tiles arrive by
//             TMA bulk copies (onesweep_pairs_kernel); keys only and the pass that reads the text take the
//             one-tile-per-CTA kernel (onesweep_kernel).
// Algorithmic traffic per digit pass: read 12 B + write 12 B per pair (8 + 8 for keys only).
#include "radix_sort.cuh"

#include <algorithm>
#include <cstdlib>

namespace gcz {

namespace {

constexpr int RB = 8;                   // digit width
constexpr int kRadix = 1 << RB;
constexpr int kHistThreads = 512;
constexpr int kLookWindow = 8;

constexpr unsigned long long kFlagAgg = 1ull << 62;
constexpr unsigned long long kFlagPrefix = 2ull << 62;
constexpr unsigned long long kValueMask = (1ull << 62) - 1;

// ---- histogram of every digit in one pass -------------------------------------------------------
__global__ void __launch_bounds__(kHistThreads)
radix_hist_kernel(const uint64_t* __restrict__ keys, int64_t n, int begin_bit, int npass,
                  unsigned long long* __restrict__ hist /* [npass][radix] */) {
    __shared__ unsigned s_hist[8 * kRadix];
    for (int i = threadIdx.x; i < npass * kRadix; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t k = keys[i] >> begin_bit;
#pragma unroll 8
        for (int p = 0; p < npass; p++) atomicAdd(&s_hist[p * kRadix + (int)((k >> (RB * p)) & (kRadix - 1))], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < npass * kRadix; i += blockDim.x) {
        const unsigned v = s_hist[i];
        if (v) atomicAdd(&hist[i], (unsigned long long)v);
    }
}

// Same histogram, with the keys computed from the text (TextKeySource): every thread converts its own
// kTextItems consecutive symbols (one 16-byte load) to codes and slides a k-symbol window over them.
// A key equal to c * (radix^k - 1) / (radix - 1) is k copies of symbol c: that is how the first and last
// positions of runs of >= k equal symbols are found without walking the text.
constexpr int kTextThreads = 256;
constexpr int kTextItems = 16;
constexpr int kTextTile = kTextThreads * kTextItems;

__device__ __forceinline__ uint64_t slide_key(uint64_t key, uint32_t c_out, uint32_t c_in, uint32_t radix, uint64_t top) {
    key -= (uint64_t)c_out * top;                          // 8-bit x 64-bit and 64-bit x 9-bit products: two IMADs each
    return key * radix + c_in;
}

__global__ void __launch_bounds__(kTextThreads)
text_hist_kernel(TextKeySource src, int npass, unsigned long long* __restrict__ hist /* [npass][radix] */, int64_t tiles) {
    __shared__ unsigned s_hist[8 * kRadix];
    __shared__ uint8_t s_code_of[256];
    __shared__ __align__(16) uint8_t s_codes[16 + kTextTile + kMaxKeySymbols + 16];    // index 16 = first position of the tile
    for (int i = threadIdx.x; i < 8 * kRadix; i += kTextThreads) s_hist[i] = 0;
    s_code_of[threadIdx.x] = src.code_of[threadIdx.x];
    const int k = src.coder.k;
    const uint32_t radix = (uint32_t)src.coder.radix;
    const uint64_t top = src.coder.top;
    uint64_t unit = 0;                                     // k ones in base radix
    for (int j = 0; j < k; j++) unit = unit * radix + 1;
    const int64_t n = src.n;
    const bool aligned = (reinterpret_cast<uintptr_t>(src.text) & 15) == 0;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t base = tile * kTextTile;
        __syncthreads();
        {
            const int64_t p0 = base + (int64_t)threadIdx.x * kTextItems;
            uint32_t packed[4] = { 0, 0, 0, 0 };
            if (aligned && p0 + kTextItems <= n) {
                const uint4 q = *reinterpret_cast<const uint4*>(src.text + p0);
                const uint32_t w[4] = { q.x, q.y, q.z, q.w };
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    packed[j] = (uint32_t)s_code_of[w[j] & 255] | (uint32_t)s_code_of[(w[j] >> 8) & 255] << 8 |
                                (uint32_t)s_code_of[(w[j] >> 16) & 255] << 16 | (uint32_t)s_code_of[w[j] >> 24] << 24;
                }
            } else {
#pragma unroll
                for (int j = 0; j < kTextItems; j++) {
                    const int64_t p = p0 + j;
                    if (p < n) packed[j >> 2] |= (uint32_t)s_code_of[src.text[p]] << (8 * (j & 3));   // 0 = past the end
                }
            }
            *reinterpret_cast<uint4*>(s_codes + 16 + threadIdx.x * kTextItems) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
            if (threadIdx.x <= (unsigned)k) {                 // right halo: k + 1 symbols
                const int64_t p = base + kTextTile + threadIdx.x;
                s_codes[16 + kTextTile + threadIdx.x] = p < n ? s_code_of[src.text[p]] : 0;
            }
            if (threadIdx.x == kTextThreads - 1) s_codes[15] = base > 0 ? s_code_of[src.text[base - 1]] : 0;
        }
        __syncthreads();
        const int first = 16 + threadIdx.x * kTextItems;
        uint64_t key = 0;
        for (int j = 0; j < k - 1; j++) key = key * radix + s_codes[first + j];
#pragma unroll 4
        for (int i = 0; i < kTextItems; i++) {
            key = slide_key(key, i > 0 ? s_codes[first + i - 1] : 0u, s_codes[first + i + k - 1], radix, top);
            const int64_t p = base + first - 16 + i;
            if (p < n) {
#pragma unroll
                for (int d = 0; d < 8; d++) {
                    if (d < npass) atomicAdd(&s_hist[d * kRadix + (int)((key >> (RB * d)) & (kRadix - 1))], 1u);
                }
                const uint32_t c = s_codes[first + i];
                if (src.run_marks && key == (uint64_t)c * unit) {
                    if (p == 0 || s_codes[first + i - 1] != c) {
                        const unsigned at = atomicAdd(src.run_mark_count, 1u);
                        if (at < src.run_mark_cap) src.run_marks[at] = 2ull * (uint64_t)p;
                    }
                    if (s_codes[first + i + k] != c) {
                        const unsigned at = atomicAdd(src.run_mark_count, 1u);
                        if (at < src.run_mark_cap) src.run_marks[at] = 2ull * (uint64_t)(p + k - 1) + 1;
                    }
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < npass * kRadix; i += kTextThreads) {
        const unsigned v = s_hist[i];
        if (v) atomicAdd(&hist[i], (unsigned long long)v);
    }
}

// exclusive scan of each pass's bins, in place (one thread per bin)
__global__ void radix_scan_kernel(unsigned long long* hist, int npass) {
    __shared__ unsigned long long s_warp[kRadix / 32];
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

// Lanes of the warp whose 8-bit digit equals mine: one ballot per bit, each folded in with two logic ops
// (written in PTX: the compiler's own expansion of the C form spends six instructions per bit).
__device__ __forceinline__ unsigned peers_with_same_digit(unsigned d) {
    unsigned acc;
    asm(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b32 t, v, m;\n"
        "mov.b32 %0, 0xffffffff;\n"
        "and.b32 t, %1, 1;   setp.ne.u32 p, t, 0; vote.sync.ballot.b32 v, p, 0xffffffff; selp.b32 m, 0, 0xffffffff, p; lop3.b32 %0, %0, v, m, 0x60;\n"
        "and.b32 t, %1, 2;   setp.ne.u32 p, t, 0; vote.sync.ballot.b32 v, p, 0xffffffff; selp.b32 m, 0, 0xffffffff, p; lop3.b32 %0, %0, v, m, 0x60;\n"
        "and.b32 t, %1, 4;   setp.ne.u32 p, t, 0; vote.sync.ballot.b32 v, p, 0xffffffff; selp.b32 m, 0, 0xffffffff, p; lop3.b32 %0, %0, v, m, 0x60;\n"
        "and.b32 t, %1, 8;   setp.ne.u32 p, t, 0; vote.sync.ballot.b32 v, p, 0xffffffff; selp.b32 m, 0, 0xffffffff, p; lop3.b32 %0, %0, v, m, 0x60;\n"
        "and.b32 t, %1, 16;  setp.ne.u32 p, t, 0; vote.sync.ballot.b32 v, p, 0xffffffff; selp.b32 m, 0, 0xffffffff, p; lop3.b32 %0, %0, v, m, 0x60;\n"
        "and.b32 t, %1, 32;  setp.ne.u32 p, t, 0; vote.sync.ballot.b32 v, p, 0xffffffff; selp.b32 m, 0, 0xffffffff, p; lop3.b32 %0, %0, v, m, 0x60;\n"
        "and.b32 t, %1, 64;  setp.ne.u32 p, t, 0; vote.sync.ballot.b32 v, p, 0xffffffff; selp.b32 m, 0, 0xffffffff, p; lop3.b32 %0, %0, v, m, 0x60;\n"
        "and.b32 t, %1, 128; setp.ne.u32 p, t, 0; vote.sync.ballot.b32 v, p, 0xffffffff; selp.b32 m, 0, 0xffffffff, p; lop3.b32 %0, %0, v, m, 0x60;\n"
        "}\n"
        : "=&r"(acc) : "r"(d));
    return acc;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy (TMA, no tensor map): bytes a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// generic-proxy accesses to shared memory before this point are ordered before async-proxy (TMA) writes after it
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// shared-memory reduction without a return value (and without the match-based aggregation ptxas wraps atomicAdd in)
__device__ __forceinline__ void red_shared_inc(unsigned* p) {
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(smem_u32(p)) : "memory");
}

// ---- scan-ahead ---------------------------------------------------------------------------------------------------
// At 25 tiles per microsecond (249 M pairs in 1.6 ms) and ~1 us for a word to travel from one SM to another through L2,
// a decoupled look-back has ~40 unfinished predecessors to walk over, every tile, in every one of its 256 digit threads:
// 2 - 3 dependent round trips that the whole CTA waits for (profiles/onesweep_*_r02.summary.txt).  Instead, tiles
// publish their digit counts BEFORE they rank their keys (agg32: count | kAggReady), and ONE extra CTA does nothing but
// follow that stream: thread d adds up the counts of digit d tile after tile (a window of kScanWindow loads in flight)
// and writes the inclusive prefix of every tile (status: sum | kFlagPrefix).  When a tile CTA needs its global offsets,
// ~5 us after it published, the prefix of its predecessor is one load away.
constexpr unsigned kAggReady = 1u << 31;
constexpr int kScanWindow = 40;

__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u32(unsigned* p, unsigned v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __noinline__ void scan_ahead(long long tiles, unsigned long long* __restrict__ status) {
    if (threadIdx.x >= kRadix) return;
    const unsigned* agg32 = reinterpret_cast<const unsigned*>(status + (size_t)tiles * kRadix) + threadIdx.x;
    unsigned long long* prefix = status + threadIdx.x;
    unsigned long long run = 0;
    long long t0 = 0;
    while (t0 < tiles) {
        unsigned a[kScanWindow];
#pragma unroll
        for (int i = 0; i < kScanWindow; i++) a[i] = t0 + i < tiles ? ld_relaxed_u32(agg32 + (size_t)(t0 + i) * kRadix) : 0u;
        int used = 0;
#pragma unroll
        for (int i = 0; i < kScanWindow; i++) {
            if (used == i && (a[i] & kAggReady)) {
                run += a[i] & ~kAggReady;
                st_relaxed_u64(prefix + (size_t)(t0 + i) * kRadix, run | kFlagPrefix);
                used = i + 1;
            }
        }
        t0 += used;                                      // counts that were not there yet are asked for again
    }
}

// Ranks of a thread's ITEMS keys among the keys of its warp with the same digit, in element order (item, then lane),
// two to a register; the warp's digit counters (my_hist, shared memory, zero on entry) end up holding the warp's counts.
template <int ITEMS>
__device__ __forceinline__ void rank_in_warp(const uint64_t (&key)[ITEMS], int shift, unsigned* my_hist, unsigned lt,
                                             unsigned (&rank2)[ITEMS / 2]) {
    static_assert(ITEMS % 2 == 0, "ranks are kept two to a register");
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        const unsigned d = (unsigned)(key[i] >> shift) & (unsigned)(kRadix - 1);
        const unsigned peers = peers_with_same_digit(d);
        const unsigned below = __popc(peers & lt);
        unsigned base = 0;
        if (below == 0) {
            base = my_hist[d];
            my_hist[d] = base + __popc(peers);
        }
        base = __shfl_sync(0xffffffffu, base, __ffs(peers) - 1);
        if (i & 1) rank2[i >> 1] |= (base + below) << 16; else rank2[i >> 1] = base + below;
        __syncwarp();
    }
}

// ---- one digit pass, one tile per CTA (keys only, and the pass that computes its keys from the text) ------------
// Phases of one CTA (tile of THREADS x ITEMS elements):
//   1 load keys (warp-striped)
//   2 rank keys inside each warp: peers with the same digit (8 ballots) get consecutive ranks in element order;
//     per-warp digit counters live in shared memory
//   3 values are requested from global memory only now (they are not live during the ranking)
//   4 threads 0..255: per-warp counts -> exclusive warp offsets and the tile histogram; publish the tile
//     aggregate for the look-back; scan -> digit starts inside the tile
//   5 reorder keys and values in shared memory
//   6 threads 0..255: decoupled look-back over windows of predecessor tiles
//   7 write digit runs out, coalesced
//   1' (FROM_TEXT) the tile's keys are computed from the text instead: codes to shared memory, a k-symbol window
//      slid over ITEMS consecutive positions per thread, transposed to the warp-striped order through shared memory
template <int THREADS, int ITEMS, bool HAS_VALS, int MIN_BLOCKS, bool FROM_TEXT, int MODE>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
onesweep_kernel(const uint64_t* __restrict__ keys_in, uint64_t* __restrict__ keys_out,
                const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ vals_out,
                int64_t n, int shift, const unsigned long long* __restrict__ digit_base,
                unsigned long long* __restrict__ status, unsigned* __restrict__ ticket, TextKeySource src) {
    constexpr int TILE = THREADS * ITEMS;
    constexpr int WARPS = THREADS / 32;
    constexpr bool EARLY = MODE >= 1, SCAN = MODE == 2;
    static_assert(THREADS >= kRadix, "one thread per digit is needed for the look-back");
    static_assert(!FROM_TEXT || HAS_VALS, "text input produces (key, position) pairs");
    if (SCAN) {
        if (blockIdx.x == 0) {                        // the scanner CTA: aggregates -> inclusive prefixes, for the whole pass
            scan_ahead((n + TILE - 1) / TILE, status);
            return;
        }
    }
    static_assert(TILE * 4 >= TILE + kMaxKeySymbols + 8 + 256, "the value staging area holds the tile's symbol codes");

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

    // 1. warp-striped load: item i of lane l in warp w is element w*ITEMS*32 + i*32 + l of the tile
    uint64_t key[ITEMS];
    const int warp_base = warp * ITEMS * 32;
    uint8_t* s_codes = reinterpret_cast<uint8_t*>(s_vals);        // FROM_TEXT: code of position tile_base - 1 + i
    if (FROM_TEXT) {
        uint8_t* s_code_of = s_codes + TILE + kMaxKeySymbols + 8;
        if (threadIdx.x < 256) s_code_of[threadIdx.x] = src.code_of[threadIdx.x];
        __syncthreads();
        const int k = src.coder.k;
        const uint32_t radix = (uint32_t)src.coder.radix;
        // codes of positions tile_base - 1 .. tile_base + TILE + k - 2 at s_codes[0 ..]; 16-byte loads where possible
        static_assert(TILE % 16 == 0, "tile starts stay 16-byte aligned");
        const bool aligned = (reinterpret_cast<uintptr_t>(src.text) & 15) == 0;
        for (int v = threadIdx.x; v < TILE / 16; v += THREADS) {
            const int64_t p0 = tile_base + (int64_t)v * 16;
            if (aligned && p0 + 16 <= n) {
                const uint4 q = *reinterpret_cast<const uint4*>(src.text + p0);
                const uint32_t w[4] = { q.x, q.y, q.z, q.w };
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    s_codes[1 + v * 16 + 4 * j] = s_code_of[w[j] & 255];
                    s_codes[2 + v * 16 + 4 * j] = s_code_of[(w[j] >> 8) & 255];
                    s_codes[3 + v * 16 + 4 * j] = s_code_of[(w[j] >> 16) & 255];
                    s_codes[4 + v * 16 + 4 * j] = s_code_of[w[j] >> 24];
                }
            } else {
                for (int j = 0; j < 16; j++) s_codes[1 + v * 16 + j] = p0 + j < n ? s_code_of[src.text[p0 + j]] : 0;
            }
        }
        if (threadIdx.x < (unsigned)k) {
            const int64_t p = tile_base + TILE + threadIdx.x;
            s_codes[1 + TILE + threadIdx.x] = p < n ? s_code_of[src.text[p]] : 0;
        }
        if (threadIdx.x == THREADS - 1) s_codes[0] = s_code_of[src.text[tile_base > 0 ? tile_base - 1 : n - 1]];   // BWT symbol of the first suffix
        __syncthreads();
        const int first = 1 + threadIdx.x * ITEMS;
        uint64_t w = 0;
        for (int j = 0; j < k - 1; j++) w = w * radix + s_codes[first + j];
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            w = slide_key(w, i > 0 ? s_codes[first + i - 1] : 0u, s_codes[first + i + k - 1], radix, src.coder.top);
            s_keys[threadIdx.x * ITEMS + i] = w;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            const int e = warp_base + i * 32 + lane;
            key[i] = e < count ? s_keys[e] : ~0ull;
        }
    } else if (count == TILE) {               // all tiles but the last: no bounds predicates in the way of the loads
        const uint64_t* src_keys = keys_in + tile_base + warp_base + lane;
#pragma unroll
        for (int i = 0; i < ITEMS; i++) key[i] = src_keys[i * 32];
    } else {
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            const int e = warp_base + i * 32 + lane;
            key[i] = e < count ? keys_in[tile_base + e] : ~0ull;
        }
    }

    // every key is requested before the first one is used (the ranking below is short enough that the scheduler
    // would otherwise sink each load next to its use and wait for it there)
#pragma unroll
    for (int i = 0; i < ITEMS; i++) asm volatile("" : "+l"(key[i]));

    unsigned rank2[ITEMS / 2];                  // ranks of items 2j (low half) and 2j + 1 (high half)
    unsigned* my_hist = s_warp_hist + warp * kRadix;
    uint32_t val[ITEMS];
    unsigned total = 0;
    unsigned long long ahead = 0;                 // SCAN: the predecessor's inclusive prefix, requested before the reorder
    auto load_values = [&]() {
        if (FROM_TEXT) {
#pragma unroll
            for (int i = 0; i < ITEMS; i++) {
                const int e = warp_base + i * 32 + lane;
                val[i] = (uint32_t)(tile_base + e) | (src.carry_shift ? (uint32_t)s_codes[e] << src.carry_shift : 0u);
            }
        } else if (HAS_VALS) {
            if (count == TILE) {
                const uint32_t* src_vals = vals_in + tile_base + warp_base + lane;
#pragma unroll
                for (int i = 0; i < ITEMS; i++) val[i] = src_vals[i * 32];
            } else {
#pragma unroll
                for (int i = 0; i < ITEMS; i++) {
                    const int e = warp_base + i * 32 + lane;
                    val[i] = e < count ? vals_in[tile_base + e] : 0u;
                }
            }
        }
    };
    if (EARLY) {
        // 2e. digit counts of every warp first (one shared-memory reduction per key), so that the tile's aggregate is
        //     published BEFORE the ranking: by the time the tiles behind this one look back, it is microseconds old
#pragma unroll
        for (int i = 0; i < ITEMS; i++) red_shared_inc(&my_hist[(unsigned)(key[i] >> shift) & (unsigned)(kRadix - 1)]);
        __syncthreads();
        if (threadIdx.x < kRadix) {
#pragma unroll
            for (int w = 0; w < WARPS; w++) total += s_warp_hist[w * kRadix + threadIdx.x];
            const unsigned with_padding = total;
            if (threadIdx.x == kRadix - 1) total -= (unsigned)(TILE - count);     // padding keys are all the last digit
            if (SCAN) {
                unsigned* agg32 = reinterpret_cast<unsigned*>(status + (size_t)((n + TILE - 1) / TILE) * kRadix);
                st_relaxed_u32(&agg32[(size_t)tile * kRadix + threadIdx.x], total | kAggReady);
            } else {
                st_relaxed_u64(&status[(size_t)tile * kRadix + threadIdx.x],
                               (unsigned long long)total | (tile == 0 ? kFlagPrefix : kFlagAgg));
            }
            const unsigned incl = warp_incl_sum(with_padding);
            if (lane == 31) s_scan[warp] = incl;
            s_digit_start[threadIdx.x] = incl - with_padding;
        }
        __syncthreads();
        if (threadIdx.x < kRadix) {
            unsigned at = s_digit_start[threadIdx.x];
            for (unsigned w = 0; w < warp; w++) at += s_scan[w];
            s_digit_start[threadIdx.x] = at;
            // where the first key of digit d in warp w lands: digit start + keys of d in the warps before w
#pragma unroll
            for (int w = 0; w < WARPS; w++) {
                const unsigned c = s_warp_hist[w * kRadix + threadIdx.x];
                s_warp_hist[w * kRadix + threadIdx.x] = at;
                at += c;
            }
        }
        __syncthreads();
        // 3e. rank inside the warp: my_hist[d] is the next free position of digit d for this warp, so the rank IS the position
        rank_in_warp<ITEMS>(key, shift, my_hist, lt, rank2);
        load_values();
        if (FROM_TEXT) __syncthreads();           // the symbol codes share their shared memory with the reordered values
        if (SCAN && threadIdx.x < kRadix && tile > 0) ahead = ld_relaxed_u64(&status[(size_t)(tile - 1) * kRadix + threadIdx.x]);
        // 5e. reorder the tile in shared memory (padding keys are the last digit and rank last: they land at >= count)
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            const unsigned pos = (i & 1) ? rank2[i >> 1] >> 16 : rank2[i >> 1] & 0xffffu;
            s_keys[pos] = key[i];
            if (HAS_VALS) s_vals[pos] = val[i];
        }
    } else {
        // 2. rank inside the warp
        rank_in_warp<ITEMS>(key, shift, my_hist, lt, rank2);

        // 3. values
        load_values();
        __syncthreads();

        // 4. per-warp counts -> exclusive warp offsets, tile histogram, aggregate, digit starts
        if (threadIdx.x < kRadix) {
#pragma unroll
            for (int w = 0; w < WARPS; w++) {
                const unsigned c = s_warp_hist[w * kRadix + threadIdx.x];
                s_warp_hist[w * kRadix + threadIdx.x] = total;
                total += c;
            }
            if (threadIdx.x == kRadix - 1) total -= (unsigned)(TILE - count);     // padding keys are all the last digit
            st_relaxed_u64(&status[(size_t)tile * kRadix + threadIdx.x],
                           (unsigned long long)total | (tile == 0 ? kFlagPrefix : kFlagAgg));
            const unsigned incl = warp_incl_sum(total);
            if (lane == 31) s_scan[warp] = incl;
            s_digit_start[threadIdx.x] = incl - total;
        }
        __syncthreads();
        if (threadIdx.x < kRadix) {
            unsigned base = 0;
            for (unsigned w = 0; w < warp; w++) base += s_scan[w];
            s_digit_start[threadIdx.x] += base;
        }
        __syncthreads();

        // 5. reorder the tile in shared memory (padding keys are the last digit and rank last: they land at >= count)
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            const unsigned d = (unsigned)(key[i] >> shift) & (unsigned)(kRadix - 1);
            const unsigned pos = s_digit_start[d] + my_hist[d] + ((i & 1) ? rank2[i >> 1] >> 16 : rank2[i >> 1] & 0xffffu);
            s_keys[pos] = key[i];
            if (HAS_VALS) s_vals[pos] = val[i];
        }
    }

    // 6. global offsets of the tile's digit runs
    if (SCAN) {
        // the scanner CTA has turned the aggregates of all earlier tiles into the inclusive prefix of tile - 1 (requested
        // before the reorder; polled again only if it had not been written yet)
        if (threadIdx.x < kRadix) {
            while (tile > 0 && (ahead & ~kValueMask) == 0) ahead = ld_relaxed_u64(&status[(size_t)(tile - 1) * kRadix + threadIdx.x]);
            s_gofs[threadIdx.x] = (long long)(digit_base[threadIdx.x] + (ahead & kValueMask)) - (long long)s_digit_start[threadIdx.x];
        }
    } else if (threadIdx.x < kRadix) {
        // decoupled look-back, kWindow predecessors per step: their status words are loaded together so the walk back
        // to the nearest inclusive prefix costs one memory latency per window, not one per tile
        unsigned long long excl = 0;
        long long t = (long long)tile - 1;
        bool done = tile == 0;
        constexpr int kWindow = EARLY ? 16 : kLookWindow;
        while (!done) {
            unsigned long long v[kWindow];
#pragma unroll
            for (int j = 0; j < kWindow; j++) {
                v[j] = t - j >= 0 ? ld_relaxed_u64(&status[(size_t)(t - j) * kRadix + threadIdx.x]) : kFlagPrefix;
            }
            int used = 0;
#pragma unroll
            for (int j = 0; j < kWindow; j++) {
                const unsigned long long flag = v[j] & ~kValueMask;
                if (!done && used == j && flag != 0) {
                    excl += v[j] & kValueMask;
                    used = j + 1;
                    done = flag == kFlagPrefix;
                }
            }
            t -= used;                                   // a status word that was not ready yet is polled again
        }
        if (tile > 0) st_relaxed_u64(&status[(size_t)tile * kRadix + threadIdx.x], (excl + total) | kFlagPrefix);
        s_gofs[threadIdx.x] = (long long)(digit_base[threadIdx.x] + excl) - (long long)s_digit_start[threadIdx.x];
    }
    __syncthreads();

    // 7. digit runs are contiguous both in shared memory and at their destination
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        const int j = i * THREADS + threadIdx.x;
        if (j < count) {
            const uint64_t k = s_keys[j];
            const long long dst = s_gofs[(unsigned)(k >> shift) & (unsigned)(kRadix - 1)] + j;
            keys_out[dst] = k;
            if (HAS_VALS) vals_out[dst] = s_vals[j];
        }
    }
}

// ---- one digit pass over (key, value) pairs: persistent CTAs, tiles delivered by TMA -----------------------------
// Two CTAs of 512 threads per SM stay resident for the whole pass and take tiles of 6144 pairs by ticket (the ticket of
// the NEXT tile is drawn at the top of the current one, so its address is known when shared memory frees up).  The tile
// lives in ONE shared-memory image (48 KB keys + 24 KB values) that is both the landing zone of the bulk copies
// (cp.async.bulk.shared::cluster.global + mbarrier complete_tx) and the buffer the tile is reordered in:
//
//   wait(keys)   -> keys to registers; digit counts of every warp (one shared-memory reduction per key)   | barrier
//   counts -> tile histogram, PUBLISHED NOW for the look-back of the tiles behind this one; digit starts   | barrier
//   per-warp starting positions (digit start + keys of the digit in earlier warps)                         | barrier
//   rank inside each warp: the rank IS the position in the reordered tile
//   wait(values) -> values to registers                       | barrier: every element of the image is in a register
//   reorder in place (registers -> image)
//   look-back over windows of kLookWindowP predecessors                                                    | barrier
//   key image + destinations to registers                     | barrier: key image drained -> bulk copy of the next tile's keys
//   keys and values out, warp counters zeroed                 | barrier: value image drained -> bulk copy of the next values
//
// Why the histogram comes before the ranking: at 25 tiles per microsecond (249 M pairs in 1.6 ms) a tile's predecessors
// are all in flight at once, and a look-back that starts one scatter phase (~1 us) after the aggregates were published
// finds the nearest ones unpublished and polls: 2.7 round trips per tile, 10 % of all warp samples waiting behind it
// (profiles/onesweep_512x12_full_r01.*, profiles/onesweep_persistent_first_r02.*).  Published before the ranking, an
// aggregate is ~5 us old when it is needed.  The DRAM latency of a tile's keys hides behind the value write-out of the
// tile before, that of its values behind its own ranking.  The last, partial tile takes plain guarded loads.
constexpr int kPThreads = 512, kPItems = 12, kPTile = kPThreads * kPItems, kPWarps = kPThreads / 32;
constexpr int kLookWindowP = 16;

struct PairsSmem {
    uint64_t keys[kPTile];
    uint32_t vals[kPTile];
    unsigned warp_hist[kPWarps * kRadix];
    long long gofs[kRadix];
    unsigned digit_start[kRadix];
    unsigned scan[8];
    unsigned long long bar_keys, bar_vals;         // mbarriers
    unsigned next_tile;
};

__global__ void __launch_bounds__(kPThreads, 2)
onesweep_pairs_kernel(const uint64_t* __restrict__ keys_in, uint64_t* __restrict__ keys_out,
                      const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ vals_out,
                      int64_t n, int shift, const unsigned long long* __restrict__ digit_base,
                      unsigned long long* __restrict__ status, unsigned* __restrict__ ticket) {
    constexpr int TILE = kPTile, ITEMS = kPItems, THREADS = kPThreads, WARPS = kPWarps;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    PairsSmem& sm = *reinterpret_cast<PairsSmem*>(smem_raw);
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned lt = lanemask_lt();
    const long long tiles = (n + TILE - 1) / TILE;
    const int warp_base = warp * ITEMS * 32;
    unsigned* my_hist = sm.warp_hist + warp * kRadix;
    const unsigned long long my_base = threadIdx.x < kRadix ? digit_base[threadIdx.x] : 0ull;

    auto is_full = [&](long long t) { return (t + 1) * (long long)TILE <= n; };

    if (threadIdx.x == 0) {
        mbar_init(&sm.bar_keys, 1);
        mbar_init(&sm.bar_vals, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const unsigned first = atomicAdd(ticket, 1u);
        sm.next_tile = first;
        if ((long long)first < tiles && is_full(first)) {
            fence_proxy_async();
            mbar_expect_tx(&sm.bar_keys, TILE * 8);
            bulk_load(sm.keys, keys_in + (size_t)first * TILE, TILE * 8, &sm.bar_keys);
            mbar_expect_tx(&sm.bar_vals, TILE * 4);
            bulk_load(sm.vals, vals_in + (size_t)first * TILE, TILE * 4, &sm.bar_vals);
        }
    }
    for (int i = threadIdx.x; i < WARPS * kRadix; i += THREADS) sm.warp_hist[i] = 0;
    __syncthreads();
    unsigned phase = 0;                              // parity of both mbarriers: they complete once per full tile
    long long tile = sm.next_tile;

    while (tile < tiles) {
        const int64_t tile_base = (int64_t)tile * TILE;
        const int count = (int)min((int64_t)TILE, n - tile_base);
        const bool full = count == TILE;
        unsigned drawn = 0;
        if (threadIdx.x == 0) drawn = atomicAdd(ticket, 1u);          // the tile after this one (used ~10 us from now)

        // 1. keys, and the digit counts of this warp
        uint64_t key[ITEMS];
        if (full) {
            mbar_wait(&sm.bar_keys, phase);
#pragma unroll
            for (int i = 0; i < ITEMS; i++) key[i] = sm.keys[warp_base + i * 32 + lane];
        } else {
#pragma unroll
            for (int i = 0; i < ITEMS; i++) {
                const int e = warp_base + i * 32 + lane;
                key[i] = e < count ? keys_in[tile_base + e] : ~0ull;
            }
        }
#pragma unroll
        for (int i = 0; i < ITEMS; i++) red_shared_inc(&my_hist[(unsigned)(key[i] >> shift) & (unsigned)(kRadix - 1)]);
        __syncthreads();

        // 2. tile histogram -> aggregate for the look-back (published before the ranking), digit starts
        unsigned total = 0;
        if (threadIdx.x < kRadix) {
#pragma unroll
            for (int w = 0; w < WARPS; w++) total += sm.warp_hist[w * kRadix + threadIdx.x];
            const unsigned with_padding = total;
            if (threadIdx.x == kRadix - 1) total -= (unsigned)(TILE - count);     // padding keys are all the last digit
            st_relaxed_u64(&status[(size_t)tile * kRadix + threadIdx.x],
                           (unsigned long long)total | (tile == 0 ? kFlagPrefix : kFlagAgg));
            const unsigned incl = warp_incl_sum(with_padding);
            if (lane == 31) sm.scan[warp] = incl;
            sm.digit_start[threadIdx.x] = incl - with_padding;
        }
        if (threadIdx.x == 0) sm.next_tile = drawn;
        __syncthreads();
        if (threadIdx.x < kRadix) {
            unsigned at = sm.digit_start[threadIdx.x];
            for (unsigned w = 0; w < warp; w++) at += sm.scan[w];
            sm.digit_start[threadIdx.x] = at;
            // where the first key of digit d in warp w lands: digit start + keys of d in the warps before w
#pragma unroll
            for (int w = 0; w < WARPS; w++) {
                const unsigned c = sm.warp_hist[w * kRadix + threadIdx.x];
                sm.warp_hist[w * kRadix + threadIdx.x] = at;
                at += c;
            }
        }
        __syncthreads();

        // 3. rank inside the warp: my_hist[d] is the next free position of digit d for this warp, so the rank is the
        //    position of the pair in the reordered tile (padding keys are the last digit and rank last: >= count)
        unsigned rank2[ITEMS / 2];
        rank_in_warp<ITEMS>(key, shift, my_hist, lt, rank2);

        // 4. values
        uint32_t val[ITEMS];
        if (full) {
            mbar_wait(&sm.bar_vals, phase);
#pragma unroll
            for (int i = 0; i < ITEMS; i++) val[i] = sm.vals[warp_base + i * 32 + lane];
        } else {
#pragma unroll
            for (int i = 0; i < ITEMS; i++) {
                const int e = warp_base + i * 32 + lane;
                val[i] = e < count ? vals_in[tile_base + e] : 0u;
            }
        }
        if (full) phase ^= 1u;
        __syncthreads();                              // the whole image is in registers

        // 5. reorder in place
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            const unsigned pos = (i & 1) ? rank2[i >> 1] >> 16 : rank2[i >> 1] & 0xffffu;
            sm.keys[pos] = key[i];
            sm.vals[pos] = val[i];
        }

        // 6. decoupled look-back, kLookWindowP predecessors per round trip (their aggregates are microseconds old by now)
        if (threadIdx.x < kRadix) {
            unsigned long long excl = 0;
            long long t = tile - 1;
            bool done = tile == 0;
            while (!done) {
                unsigned long long v[kLookWindowP];
#pragma unroll
                for (int j = 0; j < kLookWindowP; j++) {
                    v[j] = t - j >= 0 ? ld_relaxed_u64(&status[(size_t)(t - j) * kRadix + threadIdx.x]) : kFlagPrefix;
                }
                int used = 0;
#pragma unroll
                for (int j = 0; j < kLookWindowP; j++) {
                    const unsigned long long flag = v[j] & ~kValueMask;
                    if (!done && used == j && flag != 0) {
                        excl += v[j] & kValueMask;
                        used = j + 1;
                        done = flag == kFlagPrefix;
                    }
                }
                t -= used;                                   // a status word that was not ready yet is polled again
            }
            if (tile > 0) st_relaxed_u64(&status[(size_t)tile * kRadix + threadIdx.x], (excl + total) | kFlagPrefix);
            sm.gofs[threadIdx.x] = (long long)(my_base + excl) - (long long)sm.digit_start[threadIdx.x];
        }
        __syncthreads();

        const long long next = sm.next_tile;
        const bool next_full = next < tiles && is_full(next);

        // 7a. the key image goes to registers (with the destinations, which are below 2^31): it is drained early, so that the
        //     bulk copy of the next tile's keys has the whole write-out to land
        uint64_t kout[ITEMS];
        unsigned dst[ITEMS];
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            const int j = i * THREADS + threadIdx.x;
            kout[i] = sm.keys[j];
            dst[i] = (unsigned)(sm.gofs[(unsigned)(kout[i] >> shift) & (unsigned)(kRadix - 1)] + (long long)j);
        }
        __syncthreads();                              // the key image is drained
        if (threadIdx.x == 0 && next_full) {
            fence_proxy_async();
            mbar_expect_tx(&sm.bar_keys, TILE * 8);
            bulk_load(sm.keys, keys_in + (size_t)next * TILE, TILE * 8, &sm.bar_keys);
        }

        // 7b. keys and values out; warp counters zeroed for the next tile
        if (full) {
#pragma unroll
            for (int i = 0; i < ITEMS; i++) {
                keys_out[dst[i]] = kout[i];
                vals_out[dst[i]] = sm.vals[i * THREADS + threadIdx.x];
            }
        } else {
#pragma unroll
            for (int i = 0; i < ITEMS; i++) {
                const int j = i * THREADS + threadIdx.x;
                if (j < count) {
                    keys_out[dst[i]] = kout[i];
                    vals_out[dst[i]] = sm.vals[j];
                }
            }
        }
        {
            uint4* z = reinterpret_cast<uint4*>(sm.warp_hist);
#pragma unroll
            for (int i = 0; i < WARPS * kRadix / 4 / THREADS; i++) z[i * THREADS + threadIdx.x] = make_uint4(0u, 0u, 0u, 0u);
        }
        __syncthreads();                              // the value image is drained
        if (threadIdx.x == 0 && next_full) {
            fence_proxy_async();
            mbar_expect_tx(&sm.bar_vals, TILE * 4);
            bulk_load(sm.vals, vals_in + (size_t)next * TILE, TILE * 4, &sm.bar_vals);
        }
        tile = next;
    }
}

typedef void (*OnesweepFn)(const uint64_t*, uint64_t*, const uint32_t*, uint32_t*, int64_t, int, const unsigned long long*,
                           unsigned long long*, unsigned*, TextKeySource);

// one-tile-per-CTA kernels: 512 threads x 12 elements, two CTAs per SM (the shape profiles/sort_variants_r01.md picked)
constexpr int kThreads = 512, kItems = 12, kTile = kThreads * kItems;
constexpr size_t kFixedSmem = (size_t)(kThreads / 32) * kRadix * 4 + kRadix * 8 + kRadix * 4 + 64;
constexpr size_t kSmemPairs = (size_t)kTile * 12 + kFixedSmem, kSmemKeys = (size_t)kTile * 8 + kFixedSmem;
static_assert(kPTile == kTile, "both kernels cut the array into the same tiles (one status array)");

// GCZ_SORT_PERSISTENT=0 sends pairs through the one-tile-per-CTA kernel as well (A/B timing, tools/sortbench.py)
bool use_persistent() {
    static const bool on = [] { const char* e = getenv("GCZ_SORT_PERSISTENT"); return e && e[0] == '1'; }();
    return on;
}

// GCZ_SORT_MODE (A/B timing): 0 = decoupled look-back after the ranking, 1 = aggregates published before the ranking,
// 2 = as 1 with the scan-ahead CTA instead of the look-back
int sort_mode() {
    static const int mode = [] { const char* e = getenv("GCZ_SORT_MODE"); return e ? std::max(0, std::min(2, atoi(e))) : 2; }();
    return mode;
}

}  // namespace

int radix_sort_passes(int bits) { return (bits + RB - 1) / RB; }

size_t radix_sort_temp_bytes(int64_t n) {
    const int64_t tiles = (n + kTile - 1) / kTile;
    // [8][radix] histogram + per-pass (status[tiles][radix] u64 + agg32[tiles][radix] u32 + ticket)
    return 8 * (size_t)kRadix * 8 + 256 + ((size_t)tiles * kRadix * 12 + 256);
}

int radix_sort_pairs(DeviceCtx* ctx, cudaStream_t st, RadixBuffers& b, int64_t n, int begin_bit, int end_bit,
                     void* temp, SortStats* stats, const TextKeySource* src) {
    if (n <= 0 || end_bit <= begin_bit) return GCZ_OK;
    if (end_bit - begin_bit > 64 || begin_bit < 0) return fail(GCZ_E_ARG, "radix sort bit range");
    const int npass = radix_sort_passes(end_bit - begin_bit);
    const bool has_vals = b.vals[0] != nullptr;
    if (src && (!has_vals || begin_bit != 0 || src->n != n)) return fail(GCZ_E_ARG, "radix sort from text: arguments");
    const int mode = sort_mode();
    const OnesweepFn pairs = mode == 2 ? onesweep_kernel<kThreads, kItems, true, 2, false, 2> : mode == 1 ? onesweep_kernel<kThreads, kItems, true, 2, false, 1> : onesweep_kernel<kThreads, kItems, true, 2, false, 0>;
    const OnesweepFn keys_only = mode == 2 ? onesweep_kernel<kThreads, kItems, false, 2, false, 2> : mode == 1 ? onesweep_kernel<kThreads, kItems, false, 2, false, 1> : onesweep_kernel<kThreads, kItems, false, 2, false, 0>;
    const OnesweepFn from_text = mode == 2 ? onesweep_kernel<kThreads, kItems, true, 2, true, 2> : mode == 1 ? onesweep_kernel<kThreads, kItems, true, 2, true, 1> : onesweep_kernel<kThreads, kItems, true, 2, true, 0>;
    const unsigned extra = mode == 2 ? 1u : 0u;             // the scanner CTA
    if (!ctx->sort_attr[0]) {
        GCZ_CUDA(cudaFuncSetAttribute(pairs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemPairs));
        GCZ_CUDA(cudaFuncSetAttribute(keys_only, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemKeys));
        GCZ_CUDA(cudaFuncSetAttribute(from_text, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemPairs));
        GCZ_CUDA(cudaFuncSetAttribute(onesweep_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PairsSmem)));
        ctx->sort_attr[0] = true;
    }
    auto* hist = static_cast<unsigned long long*>(temp);
    auto* status = hist + 8 * kRadix + 32;
    const int64_t tiles = (n + kTile - 1) / kTile;
    auto* ticket = reinterpret_cast<unsigned*>(reinterpret_cast<unsigned*>(status + (size_t)tiles * kRadix) + (size_t)tiles * kRadix);
    const TextKeySource none;

    GCZ_CUDA(cudaMemsetAsync(hist, 0, (size_t)8 * kRadix * 8, st));
    if (src) {
        const int64_t ttiles = (n + kTextTile - 1) / kTextTile;
        const int grid = (int)std::min<int64_t>(ttiles, (int64_t)ctx->sm_count * 8);
        GCZ_LAUNCH(ctx, text_hist_kernel, grid, kTextThreads, 0, st, *src, npass, hist, ttiles);
    } else {
        const int hist_grid = (int)std::min<int64_t>((n + kHistThreads * 8 - 1) / (kHistThreads * 8), (int64_t)ctx->sm_count * 4);
        GCZ_LAUNCH(ctx, radix_hist_kernel, hist_grid, kHistThreads, 0, st, b.keys[b.cur], n, begin_bit, npass, hist);
    }
    GCZ_LAUNCH(ctx, radix_scan_kernel, 1, kRadix, 0, st, hist, npass);

    // bulk copies need 16-byte aligned sources: arena buffers are; anything else takes the one-tile-per-CTA kernel
    const bool aligned = ((reinterpret_cast<uintptr_t>(b.keys[0]) | reinterpret_cast<uintptr_t>(b.keys[1]) |
                           reinterpret_cast<uintptr_t>(b.vals[0]) | reinterpret_cast<uintptr_t>(b.vals[1])) & 15) == 0;
    for (int p = 0; p < npass; p++) {
        GCZ_CUDA(cudaMemsetAsync(status, 0, (size_t)tiles * kRadix * 12 + 64, st));
        const int in = b.cur, out = b.cur ^ 1;
        const int shift = begin_bit + RB * p;
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (stats) {
            GCZ_CUDA(cudaEventCreate(&e0)); GCZ_CUDA(cudaEventCreate(&e1));
            GCZ_CUDA(cudaEventRecord(e0, st));
        }
        if (src && p == 0) {
            from_text<<<(unsigned)tiles + extra, kThreads, kSmemPairs, st>>>(nullptr, b.keys[out], nullptr, b.vals[out], n, shift,
                                                                     hist + p * kRadix, status, ticket, *src);
        } else if (has_vals && aligned && use_persistent()) {
            const unsigned grid = (unsigned)std::min<int64_t>(tiles, (int64_t)ctx->sm_count * 2);
            onesweep_pairs_kernel<<<grid, kPThreads, sizeof(PairsSmem), st>>>(b.keys[in], b.keys[out], b.vals[in], b.vals[out], n, shift,
                                                                              hist + p * kRadix, status, ticket);
        } else if (has_vals) {
            pairs<<<(unsigned)tiles + extra, kThreads, kSmemPairs, st>>>(b.keys[in], b.keys[out], b.vals[in], b.vals[out], n, shift,
                                                                 hist + p * kRadix, status, ticket, none);
        } else {
            keys_only<<<(unsigned)tiles + extra, kThreads, kSmemKeys, st>>>(b.keys[in], b.keys[out], nullptr, nullptr, n, shift,
                                                                    hist + p * kRadix, status, ticket, none);
        }
        ctx->launches++;
        GCZ_CUDA(cudaPeekAtLastError());
        b.cur = out;
        if (stats) {
            GCZ_CUDA(cudaEventRecord(e1, st));
            stats->events.push_back(e0); stats->events.push_back(e1);
            stats->passes++; stats->elements += n;
            stats->pass_elements.push_back((src && p == 0) ? -n : n);
        }
    }
    return GCZ_OK;
}

void SortStats::resolve() {
    for (size_t i = 0; i + 1 < events.size(); i += 2) {
        float t = 0;
        if (cudaEventElapsedTime(&t, events[i], events[i + 1]) == cudaSuccess) ms += t;
        pass_ms.push_back(t);
        cudaEventDestroy(events[i]); cudaEventDestroy(events[i + 1]);
    }
    events.clear();
}

}  // namespace gcz

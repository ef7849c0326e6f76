// LSD radix sort of (u64 key, u32 value) pairs, 8-bit digits, one HBM round trip per digit.
//
// This is the workhorse of the suffix sorter (replaces the induced-sorting loops of
// algo/string/SAIS.java:103-137 by a data-parallel sort; only the resulting order is shared).
//
// Per sort:   one histogram launch (all digits at once, 8 B/key read),
//             one tiny scan launch (digit bases),
// per digit:  one "onesweep" launch: every CTA takes tiles by ticket, ranks their keys with warp ballots into
//             per-warp digit counters, learns a tile's global offsets through a decoupled look-back over
//             status words (count + aggregate / inclusive-prefix flags; 32 bits wide below 2^30 pairs), reorders the tile
//             in shared memory and writes digit runs out coalesced.
// Algorithmic traffic per digit pass: read 12 B + write 12 B per pair (8 + 8 for keys only).
#include "radix_sort.cuh"

#include <algorithm>
#include <cstdlib>

namespace gcz {

namespace {

constexpr int RB = 8;                   // digit width
constexpr int kRadix = 1 << RB;
constexpr int kHistThreads = 512;
constexpr int kPadRows = 32;                // "prefix 0" rows in front of the status array: the widest look-back window

// Status word of (tile, digit): count in the low bits, two flag bits on top (aggregate = this tile's count, prefix = the
// count of this and all earlier tiles).  32-bit words whenever a prefix fits 30 bits (any block below 2^30 pairs): half
// the L2 traffic of the look-back and a third of its instructions.  The array starts with kPadRows rows of "prefix 0"
// (written once per sort by radix_scan_kernel), so a window of predecessors never needs a bounds test.
template <typename S> struct Status;
template <> struct Status<uint32_t> {
    static constexpr uint32_t kAgg = 1u << 30, kPrefix = 2u << 30;
    static __device__ __forceinline__ uint32_t ld(const uint32_t* p) {
        uint32_t v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
    }
    static __device__ __forceinline__ void st(uint32_t* p, uint32_t v) {
        asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    }
};
template <> struct Status<unsigned long long> {
    static constexpr unsigned long long kAgg = 1ull << 62, kPrefix = 2ull << 62;
    static __device__ __forceinline__ unsigned long long ld(const unsigned long long* p) { return ld_relaxed_u64(p); }
    static __device__ __forceinline__ void st(unsigned long long* p, unsigned long long v) { st_relaxed_u64(p, v); }
};
constexpr int64_t kNarrowStatusLimit = (int64_t)1 << 30;    // pairs a sort may have for 32-bit status words

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

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
//
// Counting: npass shared-memory reductions per position.  With one 256-bin table per digit the 32 lanes of a warp hit random
// banks (3.5 wavefronts per reduction: the first version was bound by exactly that, 245 M wavefronts per chr1-sized text).
// Here counter (digit, bin) has 16 words, one per PAIR of lanes, and each lane of the pair owns a 16-bit half of the word:
// word [digit][bin][lane / 2], bank (16 bin + lane / 2) % 32 — lanes of different pairs never meet in a bank, the two of a pair
// do when their bins have the same parity (1.5 wavefronts on average), and the increment is a per-thread constant, so a count
// is three instructions (byte of the key, address, reduction).  A half-word counts what the sixteen warps of the CTA add for one
// lane: at most 256 per tile, so the table is summed into the global histogram every kTextFlushTiles tiles.  The text of the
// next tile is requested before the current one is counted.
constexpr int kTextThreads = 512;
constexpr int kTextItems = 16;
constexpr int kTextTile = kTextThreads * kTextItems;
constexpr int kTextFlushTiles = 250;                       // 250 x 256 < 2^16
constexpr int kTextCodes = 16 + kTextTile + kMaxKeySymbols + 16;
constexpr int kTextDigitWords = kRadix * 16;               // words of one digit's table
inline size_t text_hist_smem(int npass) { return (size_t)npass * kTextDigitWords * 4 + kTextCodes + 256; }

__device__ __forceinline__ uint64_t slide_key(uint64_t key, uint32_t c_out, uint32_t c_in, uint32_t radix, uint64_t top) {
    key -= (uint64_t)c_out * top;                          // 8-bit x 64-bit and 64-bit x 9-bit products: two IMADs each
    return key * radix + c_in;
}

__device__ __forceinline__ uint32_t byte_of(const uint32_t (&w)[5], int i) { return (w[i >> 2] >> (8 * (i & 3))) & 255u; }

__device__ __forceinline__ uint32_t translate4(const uint8_t* s_code_of, uint32_t w) {
    return (uint32_t)s_code_of[w & 255] | (uint32_t)s_code_of[(w >> 8) & 255] << 8 |
           (uint32_t)s_code_of[(w >> 16) & 255] << 16 | (uint32_t)s_code_of[w >> 24] << 24;
}

template <int NPASS>
__global__ void __launch_bounds__(kTextThreads)
text_hist_kernel(TextKeySource src, unsigned long long* __restrict__ hist /* [NPASS][radix] */, int64_t tiles, int flush_tiles) {
    extern __shared__ __align__(16) unsigned char text_smem[];
    unsigned* s_cnt = reinterpret_cast<unsigned*>(text_smem);                       // [NPASS][256][16]
    uint8_t* s_codes = text_smem + (size_t)NPASS * kTextDigitWords * 4;             // index 16 = first position of the tile
    uint8_t* s_code_of = s_codes + kTextCodes;
    constexpr int cnt_words = NPASS * kTextDigitWords;
    for (int i = threadIdx.x; i < cnt_words; i += kTextThreads) s_cnt[i] = 0;
    if (threadIdx.x < 256) s_code_of[threadIdx.x] = src.code_of[threadIdx.x];
    const int k = src.coder.k;
    const uint32_t radix = (uint32_t)src.coder.radix;
    const uint64_t top = src.coder.top;
    uint64_t unit = 0;                                     // k ones in base radix
    for (int j = 0; j < k; j++) unit = unit * radix + 1;
    const int64_t n = src.n;
    const bool aligned = (reinterpret_cast<uintptr_t>(src.text) & 15) == 0;
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t my_column = smem_u32(s_cnt) + (lane >> 1) * 4;     // shared-memory address of word [0][0][lane / 2]
    const uint32_t my_one = 1u << (16 * (lane & 1));
    const int first = 16 + threadIdx.x * kTextItems;
    const int in_at = first + k - 1;                       // s_codes index of the symbol that enters the window at my first position
    const uint32_t in_sel = 0x3210u + 0x1111u * (uint32_t)(in_at & 3);

    // what the tile needs from global memory, as raw text bytes: my 16 symbols, and (threads 0 .. k, and the last) one symbol of the edges
    auto fetch = [&](int64_t tile, uint4& q, uint32_t& edge) {
        const int64_t base = tile * kTextTile;
        const int64_t p0 = base + (int64_t)threadIdx.x * kTextItems;
        q = make_uint4(0, 0, 0, 0);
        edge = 0x100u;                                     // 0x100: no symbol there (code 0 = past the end)
        if (tile >= tiles) return;
        if (aligned && p0 + kTextItems <= n) {
            q = *reinterpret_cast<const uint4*>(src.text + p0);
        } else {
            uint32_t w[4] = { 0, 0, 0, 0 };
            for (int j = 0; j < kTextItems; j++) {
                if (p0 + j < n) w[j >> 2] |= (uint32_t)src.text[p0 + j] << (8 * (j & 3));
            }
            q = make_uint4(w[0], w[1], w[2], w[3]);
        }
        if (threadIdx.x <= (unsigned)k) {                  // right halo: k + 1 symbols
            const int64_t p = base + kTextTile + threadIdx.x;
            if (p < n) edge = src.text[p];
        } else if (threadIdx.x == kTextThreads - 1 && base > 0) {
            edge = src.text[base - 1];
        }
    };
    uint4 q;
    uint32_t edge;
    fetch(blockIdx.x, q, edge);
    int since_flush = 0;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t base = tile * kTextTile;
        __syncthreads();                                   // the previous tile's codes are no longer read (first trip: tables are set)
        {
            const int64_t p0 = base + (int64_t)threadIdx.x * kTextItems;
            uint32_t packed[4] = { translate4(s_code_of, q.x), translate4(s_code_of, q.y), translate4(s_code_of, q.z), translate4(s_code_of, q.w) };
            if (p0 + kTextItems > n) {                     // positions past the end carry code 0 whatever byte 0 maps to
#pragma unroll
                for (int j = 0; j < kTextItems; j++) {
                    if (p0 + j >= n) packed[j >> 2] &= ~(255u << (8 * (j & 3)));
                }
            }
            *reinterpret_cast<uint4*>(s_codes + first) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
            if (threadIdx.x <= (unsigned)k) s_codes[16 + kTextTile + threadIdx.x] = edge < 256u ? s_code_of[edge] : 0;
            if (threadIdx.x == kTextThreads - 1) s_codes[15] = edge < 256u ? s_code_of[edge] : 0;
        }
        __syncthreads();
        fetch(tile + gridDim.x, q, edge);                  // in flight while this tile is counted

        uint32_t own[5], in[5];                            // own: codes of my positions; in: codes k - 1 positions further (17 of them)
        {
            const uint4 o = *reinterpret_cast<const uint4*>(s_codes + first);
            own[0] = o.x; own[1] = o.y; own[2] = o.z; own[3] = o.w; own[4] = 0;
            const uint32_t* w = reinterpret_cast<const uint32_t*>(s_codes + (in_at & ~3));
            uint32_t raw[6];
#pragma unroll
            for (int j = 0; j < 6; j++) raw[j] = w[j];
#pragma unroll
            for (int j = 0; j < 5; j++) in[j] = __byte_perm(raw[j], raw[j + 1], in_sel);
        }
        const uint32_t before = s_codes[first - 1];
        uint64_t key = 0;                                  // the first k - 1 symbols: my own 16 from the registers, the rest byte by byte
#pragma unroll
        for (int j = 0; j < kTextItems; j++) if (j < k - 1) key = key * radix + byte_of(own, j);
        for (int j = kTextItems; j < k - 1; j++) key = key * radix + s_codes[first + j];
        const int64_t p_first = base + first - 16;
        const bool whole = p_first + kTextItems <= n;
#pragma unroll
        for (int i = 0; i < kTextItems; i++) {
            key = slide_key(key, i > 0 ? byte_of(own, i - 1) : 0u, byte_of(in, i), radix, top);
            const int64_t p = p_first + i;
            if (whole || p < n) {
                const uint32_t klo = (uint32_t)key, khi = (uint32_t)(key >> 32);
#pragma unroll
                for (int d = 0; d < NPASS; d++) {
                    const uint32_t bin = __byte_perm(d < 4 ? klo : khi, 0, 0x4440 + (d & 3));
                    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(my_column + bin * 64 + d * (kTextDigitWords * 4)), "r"(my_one) : "memory");
                }
                const uint32_t c = byte_of(own, i);
                if (src.run_marks && key == (uint64_t)c * unit) {
                    if (p == 0 || (i > 0 ? byte_of(own, i - 1) : before) != c) {
                        const unsigned at = atomicAdd(src.run_mark_count, 1u);
                        if (at < src.run_mark_cap) src.run_marks[at] = 2ull * (uint64_t)p;
                    }
                    if (byte_of(in, i + 1) != c) {
                        const unsigned at = atomicAdd(src.run_mark_count, 1u);
                        if (at < src.run_mark_cap) src.run_marks[at] = 2ull * (uint64_t)(p + k - 1) + 1;
                    }
                }
            }
        }
        const bool last_trip = tile + gridDim.x >= tiles;
        if (++since_flush == flush_tiles || last_trip) {
            since_flush = 0;
            __syncthreads();
#pragma unroll 1
            for (int d = 0; d < NPASS && threadIdx.x < kRadix; d++) {   // thread b sums the 32 half-words of bin b (rotated: one bank per thread)
                const unsigned* row = s_cnt + d * kTextDigitWords + threadIdx.x * 16;
                unsigned sum = 0;
#pragma unroll
                for (int j = 0; j < 16; j++) { const unsigned w = row[(j + (threadIdx.x >> 1)) & 15]; sum += (w & 0xffffu) + (w >> 16); }
                if (sum) atomicAdd(&hist[d * kRadix + threadIdx.x], (unsigned long long)sum);
            }
            if (!last_trip) {
                __syncthreads();
                for (int i = threadIdx.x; i < cnt_words; i += kTextThreads) s_cnt[i] = 0;
            }
        }
    }
}

typedef void (*TextHistFn)(TextKeySource, unsigned long long*, int64_t, int);
inline TextHistFn text_hist_fn(int npass) {
    switch (npass) {
        case 1: return text_hist_kernel<1>; case 2: return text_hist_kernel<2>; case 3: return text_hist_kernel<3>;
        case 4: return text_hist_kernel<4>; case 5: return text_hist_kernel<5>; case 6: return text_hist_kernel<6>;
        case 7: return text_hist_kernel<7>; default: return text_hist_kernel<8>;
    }
}

// exclusive scan of each pass's bins, in place (one thread per bin); also writes the "prefix 0" rows in front of the status array
template <typename S>
__global__ void radix_scan_kernel(unsigned long long* hist, int npass, S* status_pad) {
    __shared__ unsigned long long s_warp[kRadix / 32];
    for (int j = 0; j < kPadRows; j++) status_pad[j * kRadix + threadIdx.x] = Status<S>::kPrefix;
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

// shared-memory reduction without a return value (and without the match-based aggregation ptxas wraps atomicAdd in)
__device__ __forceinline__ void red_shared_inc(unsigned* p) {
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(smem_u32(p)) : "memory");
}

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

// Positions of a thread's ITEMS keys: keys of the warp with the same digit get consecutive positions in element order
// (item, then lane), counted up from my_hist[digit] (shared memory: where the warp's first key of that digit goes);
// two positions to a register.
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

// ---- one digit pass ---------------------------------------------------------------------------------------------
// One CTA per tile of THREADS x ITEMS elements (tiles are taken by ticket: forward progress for the look-back):
//   0 the thread that takes the ticket starts a bulk copy (TMA) of the tile's values into shared memory (ASYNC)
//   1 load keys, warp-striped (FROM_TEXT: the tile's keys are computed from the text instead — codes to shared memory,
//     a k-symbol window slid over ITEMS consecutive positions per thread; full tiles keep those keys in registers, because
//     a first pass need not be stable, the last one is transposed through shared memory)
//   2 digit counts of every warp: one shared-memory reduction per key                                      | barrier
//   3 threads 0..255: counts -> the tile's histogram, PUBLISHED NOW as the aggregate the tiles behind this one will
//     look back at; scan -> digit starts; per-warp starting positions (digit start + keys of the digit in earlier
//     warps)                                                                                               | barrier x2
//   4 rank keys inside each warp: peers with the same digit (8 ballots) get consecutive positions in element order,
//     counted up from the warp's starting position — the rank IS the position in the reordered tile
//   5 values picked up (from the staged copy, or requested from global memory only now: they are not live during the
//     ranking); reorder in shared memory
//   6 threads 0..255: decoupled look-back over windows of LOOK predecessor tiles                           | barrier
//   7 write digit runs out, coalesced
//
// What bounds it (profiles/onesweep_experiments_r02.md, three series of measured variants): not issue slots (a third fewer
// instructions changed nothing), not the look-back (half its traffic or a variant that always ends in one round trip changed
// nothing; a wider window is slower), not DRAM (half of its cycles active) — seven barrier-separated phases per tile, each
// saturating a different unit, and only as many tiles per SM to overlap them as registers and shared memory allow.  The
// histogram comes before the ranking so that the aggregate is published early (round 1 published it after: 2.7 look-back round
// trips per tile).  Built, verified, measured and dropped: persistent CTAs fed by TMA bulk copies of keys and values (1.86 ms per
// pass against 1.69), a scan-ahead CTA that turns aggregates into prefixes (2.99 ms), 9-bit digits (+47 % per pass), a
// grouped look-back, windows of 16 / 32 rows, 512 x 12 x 2, 384 x 12 x 3, 384 x 16 x 2, 256 x 12 x 4, 256 x 18 / 20 x 3.
template <int THREADS, int ITEMS, bool HAS_VALS, int MIN_BLOCKS, bool FROM_TEXT, typename S, int LOOK, bool ASYNC>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
onesweep_kernel(const uint64_t* __restrict__ keys_in, uint64_t* __restrict__ keys_out,
                const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ vals_out,
                int64_t n, int shift, const unsigned long long* __restrict__ digit_base,
                S* __restrict__ status /* row of tile 0, kPadRows "prefix 0" rows in front */, unsigned* __restrict__ ticket,
                TextKeySource src) {
    constexpr int TILE = THREADS * ITEMS;
    constexpr int WARPS = THREADS / 32;
    static_assert(THREADS >= kRadix, "one thread per digit is needed for the look-back");
    static_assert(!FROM_TEXT || HAS_VALS, "text input produces (key, position) pairs");
    static_assert(TILE * 4 >= 16 + TILE + kMaxKeySymbols + 16 + 256, "the value staging area holds the tile's symbol codes");
    static_assert((WARPS * kRadix) % (4 * THREADS) == 0, "the warp counters are cleared with 16-byte stores");
    static_assert(LOOK <= kPadRows, "pad rows cover the look-back window");
    static_assert(!ASYNC || (HAS_VALS && !FROM_TEXT), "only (key, value) array passes stage their values");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* s_keys = reinterpret_cast<uint64_t*>(smem_raw);                       // TILE
    uint32_t* s_vals = reinterpret_cast<uint32_t*>(s_keys + TILE);                  // TILE (if HAS_VALS)
    unsigned* s_warp_hist = s_vals + (HAS_VALS ? TILE : 0);                         // WARPS x 256
    S* s_gofs = reinterpret_cast<S*>(s_warp_hist + WARPS * kRadix);                 // 256: global base - tile start (mod 2^bits)
    unsigned* s_digit_start = reinterpret_cast<unsigned*>(reinterpret_cast<unsigned long long*>(s_gofs) + kRadix);   // 256
    unsigned* s_scan = s_digit_start + kRadix;                                      // 8 warp totals
    __shared__ unsigned s_tile;
    __shared__ unsigned long long s_bar;

    // ASYNC: the tile's values are not needed before the reorder; one bulk copy (TMA) brings them to where the reordered
    // values will go, issued by the thread that learns the tile number, and they are picked up after the ranking
    const bool vals_aligned = ASYNC && (reinterpret_cast<uintptr_t>(vals_in) & 15) == 0;
    if (threadIdx.x == 0) {
        if (ASYNC) { mbar_init(&s_bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
        const unsigned t = atomicAdd(ticket, 1u);
        s_tile = t;
        if (ASYNC && vals_aligned && ((int64_t)t + 1) * TILE <= n) {
            fence_proxy_async();
            mbar_expect_tx(&s_bar, TILE * 4);
            bulk_load(s_vals, vals_in + (size_t)t * TILE, TILE * 4, &s_bar);
        }
    }
#pragma unroll
    for (int i = 0; i < WARPS * kRadix / (4 * THREADS); i++) {
        reinterpret_cast<uint4*>(s_warp_hist)[i * THREADS + threadIdx.x] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    const unsigned tile = s_tile;
    const int64_t tile_base = (int64_t)tile * TILE;
    const int count = (int)min((int64_t)TILE, n - tile_base);
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned lt = lanemask_lt();

    // 1. warp-striped load: item i of lane l in warp w is element w*ITEMS*32 + i*32 + l of the tile
    uint64_t key[ITEMS];
    const int warp_base = warp * ITEMS * 32;
    uint8_t* s_codes = reinterpret_cast<uint8_t*>(s_vals);        // FROM_TEXT: s_codes[16 + e] = code of position tile_base + e
    if (FROM_TEXT) {
        uint8_t* s_code_of = s_codes + 16 + TILE + kMaxKeySymbols + 16;
        if (threadIdx.x < 256) s_code_of[threadIdx.x] = src.code_of[threadIdx.x];
        __syncthreads();
        const int k = src.coder.k;
        const uint32_t radix = (uint32_t)src.coder.radix;
        // codes of positions tile_base - 1 .. tile_base + TILE + k - 2 at s_codes[15 ..]; 16-byte loads and stores where possible
        static_assert(TILE % 16 == 0, "tile starts stay 16-byte aligned");
        const bool aligned = (reinterpret_cast<uintptr_t>(src.text) & 15) == 0;
        for (int v = threadIdx.x; v < TILE / 16; v += THREADS) {
            const int64_t p0 = tile_base + (int64_t)v * 16;
            if (aligned && p0 + 16 <= n) {
                const uint4 q = *reinterpret_cast<const uint4*>(src.text + p0);
                *reinterpret_cast<uint4*>(s_codes + 16 + v * 16) =
                    make_uint4(translate4(s_code_of, q.x), translate4(s_code_of, q.y), translate4(s_code_of, q.z), translate4(s_code_of, q.w));
            } else {
                for (int j = 0; j < 16; j++) s_codes[16 + v * 16 + j] = p0 + j < n ? s_code_of[src.text[p0 + j]] : 0;
            }
        }
        if (threadIdx.x < (unsigned)k) {
            const int64_t p = tile_base + TILE + threadIdx.x;
            s_codes[16 + TILE + threadIdx.x] = p < n ? s_code_of[src.text[p]] : 0;
        }
        if (threadIdx.x == THREADS - 1) s_codes[15] = s_code_of[src.text[tile_base > 0 ? tile_base - 1 : n - 1]];   // BWT symbol of the first suffix
        __syncthreads();
        // every thread slides the k-symbol window over ITEMS consecutive positions
        const int first = 16 + threadIdx.x * ITEMS;
        uint64_t w = 0;
        if (ITEMS == 16 && count == TILE) {
            // A full tile keeps its keys in registers: the first pass of an LSD sort need not be stable (its input has no order to
            // keep), so the keys are ranked as (item, lane) although the thread's positions are consecutive — no trip through
            // shared memory.  The codes come as words (a thread's 16 bytes are 16 apart from its neighbour's: byte loads would
            // meet four to a bank): one 16-byte load of my own positions, six words around the window's leading edge.
            uint32_t own[5], in[5];
            {
                const uint4 o = *reinterpret_cast<const uint4*>(s_codes + first);
                own[0] = o.x; own[1] = o.y; own[2] = o.z; own[3] = o.w; own[4] = 0;
                const int in_at = first + k - 1;
                const uint32_t in_sel = 0x3210u + 0x1111u * (uint32_t)(in_at & 3);
                const uint32_t* wp = reinterpret_cast<const uint32_t*>(s_codes + (in_at & ~3));
                uint32_t raw[6];
#pragma unroll
                for (int j = 0; j < 6; j++) raw[j] = wp[j];
#pragma unroll
                for (int j = 0; j < 5; j++) in[j] = __byte_perm(raw[j], raw[j + 1], in_sel);
            }
#pragma unroll
            for (int j = 0; j < 16; j++) if (j < k - 1) w = w * radix + byte_of(own, j);
            for (int j = 16; j < k - 1; j++) w = w * radix + s_codes[first + j];
#pragma unroll
            for (int i = 0; i < ITEMS; i++) {
                w = slide_key(w, i > 0 ? byte_of(own, i - 1) : 0u, byte_of(in, i), radix, src.coder.top);
                key[i] = w;
            }
        } else if (count == TILE) {
            for (int j = 0; j < k - 1; j++) w = w * radix + s_codes[first + j];
#pragma unroll
            for (int i = 0; i < ITEMS; i++) {
                w = slide_key(w, i > 0 ? s_codes[first + i - 1] : 0u, s_codes[first + i + k - 1], radix, src.coder.top);
                key[i] = w;
            }
        } else {
            for (int j = 0; j < k - 1; j++) w = w * radix + s_codes[first + j];
            // the last tile is transposed to the warp-striped order, in which the padding keys are the last elements
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

    // every key is requested before the first one is used
#pragma unroll
    for (int i = 0; i < ITEMS; i++) asm volatile("" : "+l"(key[i]));

    // 2. digit counts of this warp (padding keys are all ones: the last digit at any shift)
    unsigned* my_hist = s_warp_hist + warp * kRadix;
#pragma unroll
    for (int i = 0; i < ITEMS; i++) red_shared_inc(&my_hist[(unsigned)(key[i] >> shift) & (unsigned)(kRadix - 1)]);
    __syncthreads();

    // 3. tile histogram -> aggregate for the look-back (published before the ranking), digit starts, warp starting positions
    unsigned total = 0;
    S* const my_status = status + (size_t)tile * kRadix + threadIdx.x;
    if (threadIdx.x < kRadix) {
#pragma unroll
        for (int w = 0; w < WARPS; w++) total += s_warp_hist[w * kRadix + threadIdx.x];
        const unsigned with_padding = total;
        if (threadIdx.x == kRadix - 1) total -= (unsigned)(TILE - count);
        Status<S>::st(my_status, (S)total | Status<S>::kAgg);
        const unsigned incl = warp_incl_sum(with_padding);
        if (lane == 31) s_scan[warp] = incl;
        s_digit_start[threadIdx.x] = incl - with_padding;
    }
    __syncthreads();
    if (threadIdx.x < kRadix) {
        unsigned at = s_digit_start[threadIdx.x];
        for (unsigned w = 0; w < warp; w++) at += s_scan[w];
        s_digit_start[threadIdx.x] = at;
#pragma unroll
        for (int w = 0; w < WARPS; w++) {
            const unsigned c = s_warp_hist[w * kRadix + threadIdx.x];
            s_warp_hist[w * kRadix + threadIdx.x] = at;           // where the first key of this digit in warp w lands
            at += c;
        }
    }
    __syncthreads();

    // 4. rank inside the warp: my_hist[d] is the next free position of digit d for this warp
    unsigned rank2[ITEMS / 2];                  // positions of items 2j (low half) and 2j + 1 (high half)
    rank_in_warp<ITEMS>(key, shift, my_hist, lt, rank2);

    // 5. values, then the reorder (padding keys are the last digit and rank last: they land at >= count)
    uint32_t val[ITEMS];
    if (FROM_TEXT) {
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            const int e = count == TILE ? (int)threadIdx.x * ITEMS + i : warp_base + i * 32 + lane;     // the element key[i] belongs to
            val[i] = (uint32_t)(tile_base + e) | (src.carry_shift ? (uint32_t)s_codes[15 + e] << src.carry_shift : 0u);
        }
        __syncthreads();                      // the symbol codes share their shared memory with the reordered values
    } else if (ASYNC && vals_aligned && count == TILE) {
        mbar_wait(&s_bar, 0);
#pragma unroll
        for (int i = 0; i < ITEMS; i++) val[i] = s_vals[warp_base + i * 32 + lane];
        __syncthreads();                      // every staged value is in a register before the reordered ones are stored
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
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        const unsigned pos = (i & 1) ? rank2[i >> 1] >> 16 : rank2[i >> 1] & 0xffffu;
        s_keys[pos] = key[i];
        if (HAS_VALS) s_vals[pos] = val[i];
    }

    // 6. decoupled look-back, LOOK predecessors per step: their status words are loaded together so the walk back to the
    //    nearest inclusive prefix costs one memory latency per window, not one per tile.  The words are summed with their flag
    //    bits on; the flags of `taken - 1` aggregates and one prefix are subtracted at the end.
    if (threadIdx.x < kRadix) {
        S raw = 0;
        unsigned taken = 0;
        const S* p = my_status - kRadix;                   // tile - 1 (rows -1 .. -kPadRows: "prefix 0")
        bool done = false;
        do {
            S v[LOOK];
#pragma unroll
            for (int j = 0; j < LOOK; j++) v[j] = Status<S>::ld(p - j * kRadix);
            bool go = true;
            unsigned used = 0;
#pragma unroll
            for (int j = 0; j < LOOK; j++) {
                go = go && v[j] >= Status<S>::kAgg;          // a status word that is not there yet ends the step: polled again
                if (go) { raw += v[j]; used = j + 1; }
                if (v[j] >= Status<S>::kPrefix) { done = done || go; go = false; }
            }
            taken += used;
            p -= used * kRadix;
        } while (!done);
        const S excl = raw - (S)(taken - 1) * Status<S>::kAgg - Status<S>::kPrefix;
        Status<S>::st(my_status, ((excl + total) & (Status<S>::kAgg - 1)) | Status<S>::kPrefix);
        s_gofs[threadIdx.x] = (S)digit_base[threadIdx.x] + excl - (S)s_digit_start[threadIdx.x];
    }
    __syncthreads();

    // 7. digit runs are contiguous both in shared memory and at their destination (destination index mod 2^32 or 2^64:
    //    base - start may wrap below zero, base - start + j does not)
    if (count == TILE) {
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            const int j = i * THREADS + threadIdx.x;
            const uint64_t k = s_keys[j];
            const S dst = s_gofs[(unsigned)(k >> shift) & (unsigned)(kRadix - 1)] + (S)j;
            keys_out[dst] = k;
            if (HAS_VALS) vals_out[dst] = s_vals[j];
        }
    } else {
#pragma unroll 1
        for (int j = threadIdx.x; j < count; j += THREADS) {
            const uint64_t k = s_keys[j];
            const S dst = s_gofs[(unsigned)(k >> shift) & (unsigned)(kRadix - 1)] + (S)j;
            keys_out[dst] = k;
            if (HAS_VALS) vals_out[dst] = s_vals[j];
        }
    }
}

// Shape of the pass: 256 threads x 16 elements, three CTAs per SM (80 registers), look-back window of 8 rows.  Round 1's sweep had
// picked 512 x 12 x 2; with the leaner kernel of round 2 fewer threads with more keys each and a third tile per SM win
// (profiles/onesweep_experiments_r02.md, third series: 1.59 ms per pass against 1.69; 384 x 16 x 2: 1.63; 256 x 18 / 20 spill).
constexpr int kThreads = 256, kItems = 16, kCtasPerSm = 3, kLook = 8;
constexpr size_t onesweep_smem(int threads, int items, bool vals) {
    return (size_t)threads * items * (vals ? 12 : 8) + (size_t)(threads / 32) * kRadix * 4 + kRadix * 8 + kRadix * 4 + 64;
}
constexpr int kMinTile = kThreads * kItems;

template <typename S, int THREADS, int ITEMS, int MINB, int LOOK, bool ASYNC>
int digit_passes(DeviceCtx* ctx, cudaStream_t st, RadixBuffers& b, int64_t n, int begin_bit, int npass, unsigned long long* hist,
                 void* status_mem, SortStats* stats, const TextKeySource* src, int attr_slot) {
    constexpr int kTile = THREADS * ITEMS;
    constexpr size_t kSmemPairs = onesweep_smem(THREADS, ITEMS, true), kSmemKeys = onesweep_smem(THREADS, ITEMS, false);
    const bool has_vals = b.vals[0] != nullptr;
    auto* pairs = onesweep_kernel<THREADS, ITEMS, true, MINB, false, S, LOOK, ASYNC>;
    auto* keys_only = onesweep_kernel<THREADS, ITEMS, false, MINB, false, S, LOOK, false>;
    auto* from_text = onesweep_kernel<THREADS, ITEMS, true, MINB, true, S, LOOK, false>;
    if (!(ctx->sort_attr_mask >> attr_slot & 1)) {
        GCZ_CUDA(cudaFuncSetAttribute(pairs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemPairs));
        GCZ_CUDA(cudaFuncSetAttribute(keys_only, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemKeys));
        GCZ_CUDA(cudaFuncSetAttribute(from_text, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemPairs));
        ctx->sort_attr_mask |= 1u << attr_slot;
    }
    const int64_t tiles = (n + kTile - 1) / kTile;
    // rows of 256 status words: kPadRows of "prefix 0", then one per tile
    S* pad = static_cast<S*>(status_mem);
    S* status = pad + (size_t)kPadRows * kRadix;
    auto* ticket = reinterpret_cast<unsigned*>(status + (size_t)tiles * kRadix);
    const TextKeySource none;
    GCZ_LAUNCH(ctx, radix_scan_kernel<S>, 1, kRadix, 0, st, hist, npass, pad);

    for (int p = 0; p < npass; p++) {
        GCZ_CUDA(cudaMemsetAsync(status, 0, (size_t)tiles * kRadix * sizeof(S) + 64, st));
        const int in = b.cur, out = b.cur ^ 1;
        const int shift = begin_bit + RB * p;
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (stats) {
            GCZ_CUDA(cudaEventCreate(&e0)); GCZ_CUDA(cudaEventCreate(&e1));
            GCZ_CUDA(cudaEventRecord(e0, st));
        }
        if (src && p == 0) {
            from_text<<<(unsigned)tiles, THREADS, kSmemPairs, st>>>(nullptr, b.keys[out], nullptr, b.vals[out], n, shift,
                                                                    hist + p * kRadix, status, ticket, *src);
        } else if (has_vals) {
            pairs<<<(unsigned)tiles, THREADS, kSmemPairs, st>>>(b.keys[in], b.keys[out], b.vals[in], b.vals[out], n, shift,
                                                                hist + p * kRadix, status, ticket, none);
        } else {
            keys_only<<<(unsigned)tiles, THREADS, kSmemKeys, st>>>(b.keys[in], b.keys[out], nullptr, nullptr, n, shift,
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

}  // namespace

int radix_sort_passes(int bits) { return (bits + RB - 1) / RB; }

size_t radix_sort_temp_bytes(int64_t n) {
    const int64_t tiles = (n + kMinTile - 1) / kMinTile;
    // [8][radix] histogram + status rows (kPadRows "prefix 0" rows, then one per tile) + ticket
    return 8 * (size_t)kRadix * 8 + 256 + ((size_t)(tiles + kPadRows) * kRadix * 8 + 256);
}

int radix_sort_pairs(DeviceCtx* ctx, cudaStream_t st, RadixBuffers& b, int64_t n, int begin_bit, int end_bit,
                     void* temp, SortStats* stats, const TextKeySource* src) {
    if (n <= 0 || end_bit <= begin_bit) return GCZ_OK;
    if (end_bit - begin_bit > 64 || begin_bit < 0) return fail(GCZ_E_ARG, "radix sort bit range");
    const int npass = radix_sort_passes(end_bit - begin_bit);
    const bool has_vals = b.vals[0] != nullptr;
    if (src && (!has_vals || begin_bit != 0 || src->n != n)) return fail(GCZ_E_ARG, "radix sort from text: arguments");
    auto* hist = static_cast<unsigned long long*>(temp);
    void* status_mem = hist + 8 * kRadix + 32;

    GCZ_CUDA(cudaMemsetAsync(hist, 0, (size_t)8 * kRadix * 8, st));
    if (src) {
        const int64_t ttiles = (n + kTextTile - 1) / kTextTile;
        const size_t smem = text_hist_smem(npass);
        const TextHistFn text_hist = text_hist_fn(npass);
        if (!(ctx->text_hist_attr >> npass & 1)) {
            GCZ_CUDA(cudaFuncSetAttribute(text_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ctx->text_hist_attr |= 1u << npass;
        }
        const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, ((size_t)220 << 10) / (smem + 1024)));
        const int grid = (int)std::min<int64_t>(ttiles, (int64_t)ctx->sm_count * per_sm);
        const char* fe = std::getenv("GCZ_TEXT_FLUSH_TILES");            // tests: the mid-run flush of the 16-bit counters
        const int flush_tiles = fe && std::atoi(fe) > 0 ? std::min(std::atoi(fe), kTextFlushTiles) : kTextFlushTiles;
        GCZ_LAUNCH(ctx, text_hist, grid, kTextThreads, smem, st, *src, hist, ttiles, flush_tiles);
    } else {
        const int hist_grid = (int)std::min<int64_t>((n + kHistThreads * 8 - 1) / (kHistThreads * 8), (int64_t)ctx->sm_count * 4);
        GCZ_LAUNCH(ctx, radix_hist_kernel, hist_grid, kHistThreads, 0, st, b.keys[b.cur], n, begin_bit, npass, hist);
    }
    const bool force_wide = std::getenv("GCZ_SORT_WIDE_STATUS") != nullptr;    // tests: the 64-bit status words at any size
    if (n >= kNarrowStatusLimit || force_wide)
        return digit_passes<unsigned long long, kThreads, kItems, kCtasPerSm, kLook, true>(ctx, st, b, n, begin_bit, npass, hist, status_mem, stats, src, 0);
    // 32-bit status words, window of 8 rows, values staged by a bulk copy: profiles/onesweep_experiments_r02.md, second series
    return digit_passes<uint32_t, kThreads, kItems, kCtasPerSm, kLook, true>(ctx, st, b, n, begin_bit, npass, hist, status_mem, stats, src, 1);
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

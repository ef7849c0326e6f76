// GPU suffix sorter: the data-parallel replacement of SAIS.suffix(ByteBuffer, int[])
// (algo/string/SAIS.java:103-137).  Only the RESULT is shared with the reference — the suffix array of the
// block text under unsigned byte order with "a proper prefix sorts first" (SURVEY.md B.1); the reference's
// induced sorting is sequential and is not followed.
//
// Algorithm (prefix doubling on packed keys, sorted-group elimination, closed form for homopolymer runs):
//   1. symbols are remapped to dense codes 1..sigma (0 = past the end).  The first k symbols of every suffix
//      are packed into one integer key in base sigma+1 (no wasted bits: 17 symbols of ACGTN+'\0' in 48 bits);
//      k is the longest prefix that fits the smallest number of 8-bit digit passes for which random k-mers
//      of this symbol distribution are expected to be nearly all distinct (collision entropy of the histogram);
//   2. one LSD radix sort of (key, position) orders all suffixes by their first k symbols;
//   3. suffixes alone in their key group are final.  The others go to a compact list
//      (SA slot, suffix, group ordinal) and are refined: round t sorts the list by
//      (group ordinal, rank[suffix + h]) with h = k * 2^t, splits the groups, updates rank[] (= group start
//      slot) and drops the suffixes that became unique.  Only the list is touched after step 2;
//   4. rank[] (the inverse suffix array) is never materialised for suffixes that were final after step 2:
//      a missing entry is found on demand by binary search of the suffix's key in the sorted key array;
//   5. suffixes that start with >= k copies of one symbol c (inside a long run: the N gaps of a genome) would
//      cost log2(run length) doubling rounds.  They are ordered in ONE sort instead: with r = copies of c left
//      and d = the symbol that ends the run, all suffixes with d < c come first by ascending r, then those with
//      d > c by descending r (the S/L-type argument of induced sorting, in closed form).  Ties (same c, side, r)
//      share r symbols and continue with the general rounds at offset h + (r - k).
#include "suffix_sort.cuh"

#include <algorithm>
#include <cmath>

namespace gcz {

namespace {

constexpr uint32_t kNoRank = 0xFFFFFFFFu;
constexpr int kMaxK = 64;                       // radix 2 (one symbol + end marker) in 64 bits

struct KeyCoder {                               // key(i) = sum_j code[i + j] * radix^(k - 1 - j), j < k
    uint64_t radix;
    uint64_t top;                               // radix^(k - 1)
    int      k;
};

// Maximal runs of one symbol that are at least k long, sorted by position.
struct Run { uint32_t start, end_side; };       // end (exclusive) in bits 0..30; bit 31: the run is followed by a LARGER symbol

// ---- 1. key packing + long-run detection ------------------------------------------------------------
constexpr int kPackThreads = 256;
constexpr int kPackItems = 8;
constexpr int kPackTile = kPackThreads * kPackItems;

__global__ void __launch_bounds__(kPackThreads)
pack_keys_kernel(const uint8_t* __restrict__ text, int64_t n, const uint8_t* __restrict__ code_of, KeyCoder kc,
                 uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, int carry_shift,
                 uint64_t* __restrict__ run_marks, unsigned* __restrict__ run_mark_count, unsigned run_mark_cap) {
    __shared__ uint8_t s_code_of[256];
    __shared__ uint8_t s_codes[kPackTile + 2 * kMaxK + 8];      // [left halo kMaxK | tile | right halo]
    __shared__ uint64_t s_keys[kPackTile];
    const int64_t base = (int64_t)blockIdx.x * kPackTile;
    const int k = kc.k;
    s_code_of[threadIdx.x] = code_of[threadIdx.x];
    __syncthreads();
    for (int i = threadIdx.x; i < kPackTile + 2 * kMaxK; i += kPackThreads) {
        const int64_t p = base - kMaxK + i;
        s_codes[i] = (p >= 0 && p < n) ? s_code_of[text[p]] : 0;          // 0 = past the end, below every symbol
    }
    __syncthreads();
    // every thread slides a k-symbol window over kPackItems consecutive positions
    const int first = kMaxK + threadIdx.x * kPackItems;
    uint64_t key = 0;
    for (int j = 0; j < k - 1; j++) key = key * kc.radix + s_codes[first + j];
#pragma unroll
    for (int i = 0; i < kPackItems; i++) {
        if (i > 0) key -= (uint64_t)s_codes[first + i - 1] * kc.top;
        key = key * kc.radix + s_codes[first + i + k - 1];
        s_keys[first - kMaxK + i] = key;
    }
    // starts and ends of runs of >= k equal symbols: mark = 2 * position (+ 1 for the last position of a run)
    if (run_marks) {
#pragma unroll 1
        for (int i = 0; i < kPackItems; i++) {
            const int64_t p = base + first - kMaxK + i;
            if (p >= n) break;
            const int li = first + i;
            const uint8_t c = s_codes[li];
            bool is_start = p == 0 || s_codes[li - 1] != c;
            bool is_end = s_codes[li + 1] != c;
            if (is_start) { for (int j = 1; j < k && is_start; j++) is_start = s_codes[li + j] == c; }
            if (is_end) {
                is_end = p >= k - 1;
                for (int j = 1; j < k && is_end; j++) is_end = s_codes[li - j] == c;
            }
            if (is_start) { const unsigned at = atomicAdd(run_mark_count, 1u); if (at < run_mark_cap) run_marks[at] = 2ull * (uint64_t)p; }
            if (is_end)   { const unsigned at = atomicAdd(run_mark_count, 1u); if (at < run_mark_cap) run_marks[at] = 2ull * (uint64_t)p + 1; }
        }
    }
    __syncthreads();
    // value = position, with the code of the PRECEDING symbol (the suffix's BWT symbol) above it when both fit
    // 32 bits: the BWT then needs no gather from the text (carry_shift = 0: not carried)
    for (int i = threadIdx.x; i < kPackTile; i += kPackThreads) {
        const int64_t p = base + i;
        if (p < n) {
            uint32_t v = (uint32_t)p;
            if (carry_shift) v |= (uint32_t)(p > 0 ? s_codes[kMaxK + i - 1] : s_code_of[text[n - 1]]) << carry_shift;
            keys[p] = s_keys[i];
            vals[p] = v;
        }
    }
}

// sorted marks (start, end, start, end, ...) -> runs
__global__ void pair_runs_kernel(const uint64_t* __restrict__ marks, int64_t n_runs, const uint8_t* __restrict__ text,
                                 int64_t n, Run* __restrict__ runs) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_runs) return;
    const uint32_t start = (uint32_t)(marks[2 * j] >> 1);
    const uint32_t end = (uint32_t)(marks[2 * j + 1] >> 1) + 1u;
    const bool larger = (int64_t)end < n && text[end] > text[start];     // dense codes keep the byte order
    Run r;
    r.start = start;
    r.end_side = end | (larger ? 0x80000000u : 0u);
    runs[j] = r;
}

// Copies of the first symbol left at suffix s if that is >= k (s is inside a long run), else 0.
__device__ __forceinline__ uint32_t run_remaining(const Run* __restrict__ runs, int n_runs, uint32_t s, int k, bool* larger) {
    int lo = 0, hi = n_runs;                    // first run with end > s
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((__ldg(&runs[mid].end_side) & 0x7FFFFFFFu) > s) hi = mid; else lo = mid + 1;
    }
    if (lo >= n_runs) return 0;
    const Run r = runs[lo];
    if (r.start > s) return 0;
    const uint32_t left = (r.end_side & 0x7FFFFFFFu) - s;
    if (left < (uint32_t)k) return 0;
    *larger = (r.end_side >> 31) != 0;
    return left;
}

// ---- 3. group bookkeeping ---------------------------------------------------------------------------
constexpr int kGrpThreads = 256;
constexpr int kGrpItems = 8;
constexpr int kGrpTile = kGrpThreads * kGrpItems;
constexpr int kAggs = 5;                        // last boundary, kept slots, kept groups, kept run slots, kept run groups

struct SlotFlags {
    unsigned valid, boundary, single, run;      // bit i = slot t0 + i
};

// index + 1 of `key` in the ascending table of all-one-symbol keys, 0 when it is none of them
__device__ __forceinline__ int allc_symbol(uint64_t key, const uint64_t* __restrict__ tab, int sigma) {
    int lo = 0, hi = sigma;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (tab[mid] < key) lo = mid + 1; else hi = mid;
    }
    return (lo < sigma && tab[lo] == key) ? lo + 1 : 0;
}

// boundary[t] = first slot of a key group; single[t] = the group has exactly one member;
// run[t] (INITIAL only) = unresolved and the key is k copies of one symbol
template <bool INITIAL>
__device__ __forceinline__ SlotFlags slot_flags(const uint64_t* __restrict__ keys, int64_t m, int64_t t0,
                                                const uint64_t* __restrict__ s_allc, int sigma) {
    uint64_t k[kGrpItems + 2];
#pragma unroll
    for (int i = 0; i < kGrpItems + 2; i++) {
        const int64_t t = t0 - 1 + i;
        k[i] = (t >= 0 && t < m) ? keys[t] : 0;
    }
    unsigned bnd = 0;                       // bits 0..kGrpItems (one extra slot to the right)
#pragma unroll
    for (int i = 0; i <= kGrpItems; i++) {
        const int64_t t = t0 + i;
        const bool b = t == 0 || t >= m || k[i + 1] != k[i];
        bnd |= (unsigned)b << i;
    }
    SlotFlags f;
    f.valid = 0;
#pragma unroll
    for (int i = 0; i < kGrpItems; i++) f.valid |= (unsigned)(t0 + i < m) << i;
    f.boundary = bnd & f.valid;
    f.single = bnd & (bnd >> 1) & f.valid;
    f.run = 0;
    if (INITIAL && sigma > 0) {
        const unsigned open = f.valid & ~f.single;
#pragma unroll
        for (int i = 0; i < kGrpItems; i++) {
            if ((open >> i) & 1) f.run |= (unsigned)(allc_symbol(k[i + 1], s_allc, sigma) != 0) << i;
        }
    }
    return f;
}

struct GroupAggs { long long* a[kAggs]; };       // per-tile aggregates / their exclusive scans

// pass A: per-tile aggregates
template <bool INITIAL>
__global__ void __launch_bounds__(kGrpThreads)
group_aggregate_kernel(const uint64_t* __restrict__ keys, int64_t m, const uint64_t* __restrict__ allc, int sigma, GroupAggs agg) {
    __shared__ long long s_last[kGrpThreads / 32];
    __shared__ unsigned s_cnt[4][kGrpThreads / 32];
    __shared__ uint64_t s_allc[256];
    if (INITIAL && sigma > 0) {
        if ((int)threadIdx.x < sigma) s_allc[threadIdx.x] = allc[threadIdx.x];
        __syncthreads();
    }
    const int64_t t0 = (int64_t)blockIdx.x * kGrpTile + (int64_t)threadIdx.x * kGrpItems;
    const SlotFlags f = slot_flags<INITIAL>(keys, m, t0, s_allc, sigma);
    long long last = f.boundary ? t0 + (31 - __clz(f.boundary)) : -1;
    unsigned c[4] = { (unsigned)__popc(f.valid & ~f.single & ~f.run), (unsigned)__popc(f.boundary & ~f.single & ~f.run),
                      (unsigned)__popc(f.run), (unsigned)__popc(f.boundary & f.run) };
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
#pragma unroll
        for (int q = 0; q < 4; q++) c[q] += __shfl_xor_sync(0xffffffffu, c[q], o);
    }
    if (lane_id() == 0) {
        s_last[threadIdx.x >> 5] = last;
        for (int q = 0; q < 4; q++) s_cnt[q][threadIdx.x >> 5] = c[q];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long l = -1, t[4] = { 0, 0, 0, 0 };
        for (int w = 0; w < kGrpThreads / 32; w++) {
            l = max(l, s_last[w]);
            for (int q = 0; q < 4; q++) t[q] += s_cnt[q][w];
        }
        agg.a[0][blockIdx.x] = l;
        for (int q = 0; q < 4; q++) agg.a[1 + q][blockIdx.x] = t[q];
    }
}

// single-CTA exclusive scans over the tile aggregates (max for `last`, sum for the others), in place;
// totals[q] = sum of aggregate 1 + q
__global__ void __launch_bounds__(1024)
group_scan_kernel(GroupAggs agg, int64_t tiles, long long* __restrict__ totals) {
    __shared__ long long s_w[kAggs][32];
    __shared__ long long s_carry[kAggs];
    if (threadIdx.x < kAggs) s_carry[threadIdx.x] = threadIdx.x == 0 ? -1 : 0;
    __syncthreads();
    for (int64_t base = 0; base < tiles; base += 1024) {
        const int64_t i = base + threadIdx.x;
        long long v[kAggs], inc[kAggs];
#pragma unroll
        for (int q = 0; q < kAggs; q++) { v[q] = i < tiles ? agg.a[q][i] : (q == 0 ? -1 : 0); inc[q] = v[q]; }
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
            for (int q = 0; q < kAggs; q++) {
                const long long t = __shfl_up_sync(0xffffffffu, inc[q], o);
                if (lane_id() >= (unsigned)o) inc[q] = q == 0 ? max(inc[q], t) : inc[q] + t;
            }
        }
        if (lane_id() == 31) {
#pragma unroll
            for (int q = 0; q < kAggs; q++) s_w[q][threadIdx.x >> 5] = inc[q];
        }
        __syncthreads();
        long long b[kAggs];
#pragma unroll
        for (int q = 0; q < kAggs; q++) b[q] = s_carry[q];
        for (unsigned w = 0; w < (threadIdx.x >> 5); w++) {
            b[0] = max(b[0], s_w[0][w]);
#pragma unroll
            for (int q = 1; q < kAggs; q++) b[q] += s_w[q][w];
        }
        // exclusive results
        const long long prev_l = __shfl_up_sync(0xffffffffu, inc[0], 1);
        if (i < tiles) {
            agg.a[0][i] = lane_id() == 0 ? b[0] : max(b[0], prev_l);
#pragma unroll
            for (int q = 1; q < kAggs; q++) agg.a[q][i] = b[q] + inc[q] - v[q];
        }
        __syncthreads();
        if (threadIdx.x == 1023) {
            s_carry[0] = max(b[0], inc[0]);
#pragma unroll
            for (int q = 1; q < kAggs; q++) s_carry[q] = b[q] + inc[q];
        }
        __syncthreads();
    }
    if (threadIdx.x < kAggs - 1) totals[threadIdx.x] = s_carry[1 + threadIdx.x];
}

struct ApplyArgs {
    const uint64_t* keys;          // sorted keys of the m slots / list entries
    const uint32_t* suf;           // suffix of every entry
    const uint32_t* pos;           // refine: SA slot of every list entry (ascending); initial: null (slot = index)
    int64_t         m;
    GroupAggs       pre;           // exclusive scans of the tile aggregates
    uint32_t*       rank;
    uint32_t*       sa;
    uint32_t*       pos_out;       // next list
    uint32_t*       suf_out;
    uint32_t*       gid_out;
    uint32_t        gid_base;
    uint32_t        pos_mask;      // suffix = value & pos_mask (the bits above carry the BWT symbol)
    // initial only: the long-run suffixes leave through a separate list, already keyed for their one sort
    const uint64_t* allc;
    int             sigma;
    const Run*      runs;
    int             n_runs;
    int             k;
    uint32_t*       run_pos_out;
    uint64_t*       run_key_out;
    uint32_t*       run_suf_out;
};

// pass B: ranks, finished suffixes, next list
template <bool INITIAL>
__global__ void __launch_bounds__(kGrpThreads)
group_apply_kernel(ApplyArgs a) {
    __shared__ long long s_last[kGrpThreads / 32];
    __shared__ unsigned s_cnt[4][kGrpThreads / 32];
    __shared__ uint64_t s_allc[256];
    if (INITIAL && a.sigma > 0) {
        if ((int)threadIdx.x < a.sigma) s_allc[threadIdx.x] = a.allc[threadIdx.x];
        __syncthreads();
    }
    const int64_t t0 = (int64_t)blockIdx.x * kGrpTile + (int64_t)threadIdx.x * kGrpItems;
    const SlotFlags f = slot_flags<INITIAL>(a.keys, a.m, t0, s_allc, a.sigma);
    const long long my_last = f.boundary ? t0 + (31 - __clz(f.boundary)) : -1;
    const unsigned mine[4] = { (unsigned)__popc(f.valid & ~f.single & ~f.run), (unsigned)__popc(f.boundary & ~f.single & ~f.run),
                               (unsigned)__popc(f.run), (unsigned)__popc(f.boundary & f.run) };
    // block-wide exclusive scans of the per-thread aggregates
    const long long il = warp_incl_max(my_last);
    unsigned inc[4];
#pragma unroll
    for (int q = 0; q < 4; q++) inc[q] = warp_incl_sum(mine[q]);
    if (lane_id() == 31) {
        s_last[threadIdx.x >> 5] = il;
#pragma unroll
        for (int q = 0; q < 4; q++) s_cnt[q][threadIdx.x >> 5] = inc[q];
    }
    __syncthreads();
    long long last = a.pre.a[0][blockIdx.x];
    long long cnt[4];
#pragma unroll
    for (int q = 0; q < 4; q++) cnt[q] = a.pre.a[1 + q][blockIdx.x];
    for (unsigned w = 0; w < (threadIdx.x >> 5); w++) {
        last = max(last, s_last[w]);
#pragma unroll
        for (int q = 0; q < 4; q++) cnt[q] += s_cnt[q][w];
    }
    const long long prev_l = __shfl_up_sync(0xffffffffu, il, 1);
    if (lane_id() > 0) last = max(last, prev_l);
#pragma unroll
    for (int q = 0; q < 4; q++) cnt[q] += inc[q] - mine[q];
    long long keep = cnt[0], groups = cnt[1], keep_run = cnt[2];

#pragma unroll
    for (int i = 0; i < kGrpItems; i++) {
        if (!((f.valid >> i) & 1)) break;
        const int64_t t = t0 + i;
        if ((f.boundary >> i) & 1) last = t;
        if ((f.single >> i) & 1) {
            // final.  Initial: already in place, and its rank is left unset (found by key search when needed)
            if (!INITIAL) { const uint32_t v = a.suf[t]; a.rank[v & a.pos_mask] = a.pos[last]; a.sa[a.pos[t]] = v; }
            continue;
        }
        const uint32_t sv = a.suf[t], s = sv & a.pos_mask;
        a.rank[s] = INITIAL ? (uint32_t)last : a.pos[last];
        if (INITIAL && ((f.run >> i) & 1)) {
            const int c = allc_symbol(a.keys[t], s_allc, a.sigma);
            bool larger = false;
            const uint32_t r = run_remaining(a.runs, a.n_runs, s, a.k, &larger);
            // symbol, then the side the run ends on, then run length left: ascending below, descending above
            const uint32_t order = larger ? (0x80000000u | (0x7FFFFFFFu - r)) : r;
            a.run_pos_out[keep_run] = (uint32_t)t;
            a.run_key_out[keep_run] = ((uint64_t)c << 32) | order;
            a.run_suf_out[keep_run] = sv;
            keep_run++;
        } else {
            if ((f.boundary >> i) & 1) groups++;
            a.pos_out[keep] = INITIAL ? (uint32_t)t : a.pos[t];
            a.suf_out[keep] = sv;
            a.gid_out[keep] = a.gid_base + (uint32_t)(groups - 1);
            keep++;
        }
    }
}

// key of the suffix at q, straight from the text (what pack_keys_kernel stored for it)
__device__ __forceinline__ uint64_t key_at(const uint8_t* __restrict__ text, int64_t n, const uint8_t* __restrict__ code_of,
                                           const KeyCoder& kc, int64_t q) {
    uint64_t key = 0;
    for (int j = 0; j < kc.k; j++) {
        const int64_t p = q + j;
        key = key * kc.radix + (p < n ? (uint64_t)code_of[text[p]] : 0ull);
    }
    return key;
}

// refinement key: (group ordinal, rank of the suffix `h` symbols further, 0 when that is past the end).
// A suffix inside a long run of r >= k symbols looks r - k symbols further than the others (header, point 5).
__global__ void __launch_bounds__(256)
refine_keys_kernel(const uint32_t* __restrict__ suf, const uint32_t* __restrict__ gid, int64_t m,
                   uint32_t* __restrict__ rank, int64_t n, int64_t h, int low_bits, uint32_t pos_mask,
                   const Run* __restrict__ runs, int n_runs, const uint8_t* __restrict__ text,
                   const uint8_t* __restrict__ code_of, KeyCoder kc, const uint64_t* __restrict__ sorted_keys,
                   uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    __shared__ uint8_t s_code_of[256];
    s_code_of[threadIdx.x] = code_of[threadIdx.x];
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < m; u += stride) {
        const uint32_t sv = suf[u], s = sv & pos_mask;
        int64_t q = (int64_t)s + h;
        if (n_runs > 0) {
            bool larger;
            const uint32_t r = run_remaining(runs, n_runs, s, kc.k, &larger);
            if (r) q += (int64_t)r - kc.k;
        }
        uint64_t low = 0;
        if (q < n) {
            uint32_t rk = rank[q];
            if (rk == kNoRank) {
                // final since the first sort: its key is unique, its slot is where the key sits
                const uint64_t want = key_at(text, n, s_code_of, kc, q);
                int64_t lo = 0, hi = n;
                while (lo < hi) {
                    const int64_t mid = (lo + hi) >> 1;
                    if (__ldg(&sorted_keys[mid]) < want) lo = mid + 1; else hi = mid;
                }
                rk = (uint32_t)lo;
                rank[q] = rk;                               // every writer stores the same value
            }
            low = (uint64_t)rk + 1;
        }
        keys[u] = ((uint64_t)gid[u] << low_bits) | low;
        vals[u] = sv;
    }
}

inline int bits_for(uint64_t max_value) {      // number of bits needed to represent values 0..max_value
    int b = 0;
    while (b < 64 && (max_value >> b) != 0) b++;
    return b == 0 ? 1 : b;
}

// How many symbols go into the first key (header, point 1).
KeyCoder choose_key(const int64_t counts[256], int64_t n, int sigma) {
    const unsigned __int128 one = 1;
    const uint64_t radix = (uint64_t)sigma + 1;
    double sum_p2 = 0;
    for (int c = 0; c < 256; c++) { const double p = (double)counts[c] / (double)n; sum_p2 += p * p; }
    const double h2 = std::max(1e-3, -std::log2(std::min(1.0, sum_p2)));
    const int k_min = (int)std::min<double>(kMaxK, std::ceil((std::log2((double)n + 1) + 6.0) / h2));
    KeyCoder kc;
    kc.radix = radix;
    for (int passes = 1; passes <= 8; passes++) {
        int k = 0;
        unsigned __int128 pw = 1;                            // radix^k
        while (k < kMaxK && pw * radix <= (one << (8 * passes))) { pw *= radix; k++; }
        kc.k = std::max(k, 1);
        if (k >= k_min) break;
    }
    kc.top = 1;
    for (int j = 1; j < kc.k; j++) kc.top *= radix;
    return kc;
}

}  // namespace

size_t suffix_sort_workspace_bytes(int64_t n) {
    // rank 4n + keys 16n + vals(other) 4n + refinement worst case (lists 24n, sort 12n, run list 4n) + run marks + sort temp
    const size_t tiles = (size_t)(n / kGrpTile + 2);
    return (size_t)n * (4 + 16 + 4 + 40 + 2) + radix_sort_temp_bytes(n) + tiles * kAggs * 8 + (1 << 20);
}

int suffix_sort(DeviceCtx* ctx, cudaStream_t st, const uint8_t* d_text, int64_t n, const int64_t counts[256],
                uint32_t* d_sa, Arena& arena, SuffixSortStats* stats, int* carry_shift) {
    if (n <= 0 || n > 0x7FFFFFFFll) return fail(GCZ_E_RANGE, "block of %lld symbols", (long long)n);

    // dense symbol codes: 1..sigma in byte order, 0 reserved for "past the end"
    uint8_t h_code[256];
    int sigma = 0;
    for (int c = 0; c < 256; c++) { if (counts[c] > 0) ++sigma; h_code[c] = counts[c] > 0 ? (uint8_t)sigma : 0; }
    // 256 distinct byte values plus the end marker would need 9-bit codes; FASTA text never gets there
    if (sigma > 255) return fail(GCZ_E_RANGE, "all 256 byte values present: not supported by the key packer");
    const KeyCoder kc = choose_key(counts, n, sigma);
    const int k = kc.k;
    uint64_t h_allc[256];                                    // key of k copies of symbol c, ascending in c
    {
        uint64_t unit = 0;
        for (int j = 0; j < k; j++) unit = unit * kc.radix + 1;
        for (int c = 1; c <= sigma; c++) h_allc[c - 1] = unit * (uint64_t)c;
    }
    const int key_bits = bits_for(h_allc[sigma - 1]);
    // the BWT symbol rides above the position when both fit 32 bits (any block up to 2^28 symbols of DNA)
    const int pos_bits = bits_for((uint64_t)n - 1);
    const int carry = (carry_shift && *carry_shift && pos_bits + bits_for((uint64_t)sigma) <= 32) ? pos_bits : 0;
    const uint32_t pos_mask = carry ? (1u << carry) - 1u : 0xFFFFFFFFu;
    if (carry_shift) *carry_shift = carry;

    const size_t mark0 = arena.mark();
    const unsigned mark_cap = (unsigned)std::min<int64_t>(n / 8 + 1024, 0x7FFFFFF0ll);
    uint8_t* d_code = arena.get<uint8_t>(256);
    uint64_t* d_allc = arena.get<uint64_t>(256);
    uint32_t* d_rank = arena.get<uint32_t>((size_t)n);
    uint64_t* d_keys0 = arena.get<uint64_t>((size_t)n);
    uint64_t* d_keys1 = arena.get<uint64_t>((size_t)n);
    uint32_t* d_vals1 = arena.get<uint32_t>((size_t)n);
    void* d_temp = arena.raw(radix_sort_temp_bytes(n));
    const int64_t tiles_n = (n + kGrpTile - 1) / kGrpTile;
    long long* d_agg = arena.get<long long>((size_t)tiles_n * kAggs + 8);
    long long* d_totals = arena.get<long long>(8);                 // [0..3] group totals, [4] run marks (as unsigned)
    uint64_t* d_marks[2] = { arena.get<uint64_t>(mark_cap), arena.get<uint64_t>(mark_cap) };
    Run* d_runs = arena.get<Run>(mark_cap / 2 + 1);
    if (!d_code || !d_allc || !d_rank || !d_keys0 || !d_keys1 || !d_vals1 || !d_temp || !d_agg || !d_totals ||
        !d_marks[0] || !d_marks[1] || !d_runs)
        return fail(GCZ_E_NOMEM, "suffix sort workspace for n=%lld", (long long)n);
    GroupAggs agg;
    for (int q = 0; q < kAggs; q++) agg.a[q] = d_agg + (size_t)q * tiles_n;
    unsigned* d_mark_count = reinterpret_cast<unsigned*>(d_totals + 4);

    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
    if (stats) {
        GCZ_CUDA(cudaEventCreate(&ev0)); GCZ_CUDA(cudaEventCreate(&ev1)); GCZ_CUDA(cudaEventCreate(&ev2));
        GCZ_CUDA(cudaEventRecord(ev0, st));
    }

    GCZ_CUDA(cudaMemcpyAsync(d_code, h_code, 256, cudaMemcpyHostToDevice, st));
    GCZ_CUDA(cudaMemcpyAsync(d_allc, h_allc, sizeof(uint64_t) * sigma, cudaMemcpyHostToDevice, st));
    GCZ_CUDA(cudaMemsetAsync(d_totals, 0, 8 * sizeof(long long), st));
    GCZ_CUDA(cudaMemsetAsync(d_rank, 0xFF, (size_t)n * 4, st));

    // full sort by the first k symbols; start in the buffer that makes the result land in d_sa
    const int npass = (key_bits + 7) / 8;
    RadixBuffers b;
    b.keys[0] = d_keys0; b.keys[1] = d_keys1;
    b.vals[0] = d_sa;    b.vals[1] = d_vals1;
    b.cur = npass & 1;
    const int pack_grid = (int)((n + kPackTile - 1) / kPackTile);
    GCZ_LAUNCH(ctx, pack_keys_kernel, pack_grid, kPackThreads, 0, st, d_text, n, d_code, kc, b.keys[b.cur], b.vals[b.cur], carry,
               d_marks[0], d_mark_count, mark_cap);
    SortStats ss;
    SortStats* ssp = stats ? &ss : nullptr;
    GCZ_TRY(radix_sort_pairs(ctx, st, b, n, 0, key_bits, d_temp, ssp));
    if (b.cur != 0) return fail(GCZ_E_INTERNAL, "initial sort landed in the wrong buffer");

    // groups of equal keys
    int sigma_runs = sigma;                 // 0 switches the long-run path off
    GCZ_LAUNCH(ctx, group_aggregate_kernel<true>, (unsigned)tiles_n, kGrpThreads, 0, st, d_keys0, n, d_allc, sigma_runs, agg);
    GCZ_LAUNCH(ctx, group_scan_kernel, 1, 1024, 0, st, agg, tiles_n, d_totals);
    long long h_totals[5] = { 0, 0, 0, 0, 0 };
    GCZ_CUDA(cudaMemcpyAsync(h_totals, d_totals, sizeof(h_totals), cudaMemcpyDeviceToHost, st));
    GCZ_CUDA(cudaStreamSynchronize(st));
    int64_t m = h_totals[0], groups = h_totals[1];
    int64_t m_run = h_totals[2];
    const int64_t n_marks = (int64_t)(uint32_t)h_totals[4];
    int n_runs = 0;
    if (m_run > 0 && (n_marks > (int64_t)mark_cap || (n_marks & 1))) {
        // more long runs than the mark buffer holds (a text made of medium runs): plain doubling, where a
        // "run" suffix is an ordinary member of its key group — regroup without the run classification
        sigma_runs = 0;
        GCZ_LAUNCH(ctx, group_aggregate_kernel<true>, (unsigned)tiles_n, kGrpThreads, 0, st, d_keys0, n, d_allc, 0, agg);
        GCZ_LAUNCH(ctx, group_scan_kernel, 1, 1024, 0, st, agg, tiles_n, d_totals);
        GCZ_CUDA(cudaMemcpyAsync(h_totals, d_totals, 4 * sizeof(long long), cudaMemcpyDeviceToHost, st));
        GCZ_CUDA(cudaStreamSynchronize(st));
        m = h_totals[0]; groups = h_totals[1]; m_run = h_totals[2];
    }
    if (m_run > 0) {
        n_runs = (int)(n_marks / 2);
        RadixBuffers rb;
        rb.keys[0] = d_marks[0]; rb.keys[1] = d_marks[1];
        GCZ_TRY(radix_sort_pairs(ctx, st, rb, n_marks, 0, bits_for(2ull * (uint64_t)n + 1), d_temp, nullptr));
        GCZ_LAUNCH(ctx, pair_runs_kernel, (unsigned)((n_runs + 255) / 256), 256, 0, st, rb.keys[rb.cur], (int64_t)n_runs, d_text, n, d_runs);
    }

    // refinement buffers: the initial sort's second key/value arrays are dead, its sorted keys stay (point 4)
    const int64_t m0 = m + m_run;
    const size_t cap = (size_t)std::max<int64_t>(m0, 1);
    uint32_t* list_pos[2] = { arena.get<uint32_t>(cap), arena.get<uint32_t>(cap) };
    uint32_t* list_suf[2] = { arena.get<uint32_t>(cap), arena.get<uint32_t>(cap) };
    uint32_t* list_gid[2] = { arena.get<uint32_t>(cap), arena.get<uint32_t>(cap) };
    uint64_t* r_keys1 = arena.get<uint64_t>(cap);
    uint32_t* r_vals1 = arena.get<uint32_t>(cap);
    uint32_t* run_pos = arena.get<uint32_t>((size_t)std::max<int64_t>(m_run, 1));
    if (!list_pos[0] || !list_pos[1] || !list_suf[0] || !list_suf[1] || !list_gid[0] || !list_gid[1] || !r_keys1 || !r_vals1 || !run_pos)
        return fail(GCZ_E_NOMEM, "suffix sort refinement lists for %lld unresolved suffixes", (long long)m0);
    uint64_t* r_keys0 = d_keys1;
    uint32_t* r_vals0 = d_vals1;

    ApplyArgs aa;
    aa.keys = d_keys0; aa.suf = d_sa; aa.pos = nullptr; aa.m = n; aa.pre = agg; aa.rank = d_rank; aa.sa = d_sa;
    aa.pos_out = list_pos[0]; aa.suf_out = list_suf[0]; aa.gid_out = list_gid[0]; aa.gid_base = 0; aa.pos_mask = pos_mask;
    aa.allc = d_allc; aa.sigma = sigma_runs; aa.runs = d_runs; aa.n_runs = n_runs; aa.k = k;
    aa.run_pos_out = run_pos; aa.run_key_out = r_keys0; aa.run_suf_out = r_vals0;
    GCZ_LAUNCH(ctx, group_apply_kernel<true>, (unsigned)tiles_n, kGrpThreads, 0, st, aa);
    if (stats) GCZ_CUDA(cudaEventRecord(ev1, st));

    int rounds = 0;
    if (m_run > 0) {
        // the long-run suffixes: one sort by (symbol, side, run length left); what stays tied joins the list
        RadixBuffers rb;
        rb.keys[0] = r_keys0; rb.keys[1] = r_keys1;
        rb.vals[0] = r_vals0; rb.vals[1] = r_vals1;
        rb.cur = 0;
        GCZ_TRY(radix_sort_pairs(ctx, st, rb, m_run, 0, 32 + bits_for((uint64_t)sigma), d_temp, ssp));
        const int64_t tiles_r = (m_run + kGrpTile - 1) / kGrpTile;
        GCZ_LAUNCH(ctx, group_aggregate_kernel<false>, (unsigned)tiles_r, kGrpThreads, 0, st, rb.keys[rb.cur], m_run, nullptr, 0, agg);
        GCZ_LAUNCH(ctx, group_scan_kernel, 1, 1024, 0, st, agg, tiles_r, d_totals);
        ApplyArgs ar = aa;
        ar.keys = rb.keys[rb.cur]; ar.suf = rb.vals[rb.cur]; ar.pos = run_pos; ar.m = m_run; ar.sigma = 0;
        ar.pos_out = list_pos[0] + m; ar.suf_out = list_suf[0] + m; ar.gid_out = list_gid[0] + m; ar.gid_base = (uint32_t)groups;
        GCZ_LAUNCH(ctx, group_apply_kernel<false>, (unsigned)tiles_r, kGrpThreads, 0, st, ar);
        GCZ_CUDA(cudaMemcpyAsync(h_totals, d_totals, 2 * sizeof(long long), cudaMemcpyDeviceToHost, st));
        GCZ_CUDA(cudaStreamSynchronize(st));
        m += h_totals[0]; groups += h_totals[1];
        rounds++;
    }

    const int low_bits = bits_for((uint64_t)n);        // rank + 1 <= n
    int cur = 0;
    int64_t h = k;
    while (m > 0) {
        if (h >= 2 * n + 64) return fail(GCZ_E_INTERNAL, "suffix sort did not converge");
        RadixBuffers rb;
        rb.keys[0] = r_keys0; rb.keys[1] = r_keys1;
        rb.vals[0] = r_vals0; rb.vals[1] = r_vals1;
        rb.cur = 0;
        const int grid = (int)std::min<int64_t>((m + 255) / 256, (int64_t)ctx->sm_count * 16);
        GCZ_LAUNCH(ctx, refine_keys_kernel, grid, 256, 0, st, list_suf[cur], list_gid[cur], m, d_rank, n, h, low_bits, pos_mask,
                   d_runs, n_runs, d_text, d_code, kc, d_keys0, rb.keys[0], rb.vals[0]);
        const int gid_bits = bits_for((uint64_t)std::max<int64_t>(groups - 1, 0));
        GCZ_TRY(radix_sort_pairs(ctx, st, rb, m, 0, low_bits + gid_bits, d_temp, ssp));
        const int64_t tiles_m = (m + kGrpTile - 1) / kGrpTile;
        GCZ_LAUNCH(ctx, group_aggregate_kernel<false>, (unsigned)tiles_m, kGrpThreads, 0, st, rb.keys[rb.cur], m, nullptr, 0, agg);
        GCZ_LAUNCH(ctx, group_scan_kernel, 1, 1024, 0, st, agg, tiles_m, d_totals);
        ApplyArgs ar = aa;
        ar.keys = rb.keys[rb.cur]; ar.suf = rb.vals[rb.cur]; ar.pos = list_pos[cur]; ar.m = m; ar.sigma = 0;
        ar.pos_out = list_pos[cur ^ 1]; ar.suf_out = list_suf[cur ^ 1]; ar.gid_out = list_gid[cur ^ 1]; ar.gid_base = 0;
        GCZ_LAUNCH(ctx, group_apply_kernel<false>, (unsigned)tiles_m, kGrpThreads, 0, st, ar);
        GCZ_CUDA(cudaMemcpyAsync(h_totals, d_totals, 2 * sizeof(long long), cudaMemcpyDeviceToHost, st));
        GCZ_CUDA(cudaStreamSynchronize(st));
        m = h_totals[0]; groups = h_totals[1];
        cur ^= 1;
        h *= 2;
        rounds++;
    }

    if (stats) {
        GCZ_CUDA(cudaEventRecord(ev2, st));
        GCZ_CUDA(cudaEventSynchronize(ev2));
        cudaEventElapsedTime(&stats->initial_ms, ev0, ev1);
        cudaEventElapsedTime(&stats->refine_ms, ev1, ev2);
        stats->rounds = rounds;
        ss.resolve();
        stats->radix_passes = ss.passes;
        stats->radix_elements = ss.elements;
        stats->radix_ms = ss.ms;
        stats->symbols_per_key = k;
        stats->unresolved_after_first_sort = m0;
        stats->long_runs = n_runs;
        cudaEventDestroy(ev0); cudaEventDestroy(ev1); cudaEventDestroy(ev2);
    }
    arena.release(mark0);
    return GCZ_OK;
}

}  // namespace gcz

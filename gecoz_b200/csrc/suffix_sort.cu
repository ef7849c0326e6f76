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
#include "row_scan.cuh"

#include <algorithm>
#include <cmath>

namespace gcz {

namespace {

constexpr uint32_t kNoRank = 0xFFFFFFFFu;
// Maximal runs of one symbol that are at least k long, sorted by position.
struct Run { uint32_t start, end_side; };       // end (exclusive) in bits 0..30; bit 31: the run is followed by a LARGER symbol

// ---- 1. long runs (the marks come from the histogram pass of the first sort, radix_sort.cu) ------------------
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
__device__ __forceinline__ uint32_t run_remaining(const Run* __restrict__ runs, int n_runs, uint32_t s, int k, bool* larger, int* index = nullptr) {
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
    if (index) *index = lo;
    return left;
}

// ---- 3. group bookkeeping ---------------------------------------------------------------------------
constexpr int kGrpThreads = 256;
constexpr int kGrpItems = 16;
constexpr int kGrpTile = kGrpThreads * kGrpItems;
constexpr int kSampleShift = 4;                 // every 16th sorted key is kept as an index for rank lookups (key_slot)
constexpr int kSampleStep = 1 << kSampleShift;
// ... and every 16th of those, and every 16th of those: five levels, a node of 16 keys (one 128-byte line) per level and lookup
constexpr int kSampleLevels = 5;
constexpr int kSampleFanShift = 4;
struct SampleIndex {
    uint64_t* level[kSampleLevels];             // level[j][i] = sorted key at slot i << (kSampleShift + j * kSampleFanShift)
    int64_t   top_count;                        // entries of the highest level
};
__host__ __device__ inline int64_t sample_level_count_dev(int64_t n, int j) { const int sh = kSampleShift + j * kSampleFanShift; return ((n - 1) >> sh) + 1; }
inline int64_t sample_level_count(int64_t n, int j) { const int sh = kSampleShift + j * kSampleFanShift; return ((n - 1) >> sh) + 1; }
inline size_t sample_index_words(int64_t n) {
    size_t w = 0;
    for (int j = 0; j < kSampleLevels; j++) w += ((size_t)sample_level_count(n, j) + 15 + 16) & ~(size_t)15;   // whole lines, one spare
    return w;
}
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

// Grouping of a sorted key sequence, three launches and no serial chain between tiles:
//   flags   every key is read once, coalesced: bit arrays "slot starts a key group" / "unresolved long-run suffix",
//           and per-tile aggregates (last group start, kept slots / groups of the general and of the long-run kind);
//   scan    exclusive scan of the tile aggregates (row_scan.cuh: chunks of a row chained by a look-back);
//   apply   works from the bit arrays only: ranks, finished suffixes and the next list.
struct GroupArgs {
    const uint64_t* keys;          // sorted keys of the m slots / list entries
    const uint32_t* suf;           // suffix of every entry
    const uint32_t* pos;           // refine: SA slot of every list entry (ascending); initial: null (slot = index)
    int64_t         m;
    unsigned*       bnd_bits;      // [m / 32 + 2] bit t: slot t starts a key group
    unsigned*       run_bits;      // [m / 32 + 2] bit t: slot t is an unresolved long-run suffix (initial only)
    unsigned*       agg;           // [kAggs][tiles] tile aggregates, then their exclusive scans
    SampleIndex     sample;        // initial only: every kSampleStep-th key and the levels above it (see key_slot)
    long long*      totals;        // [4]: kept slots, kept groups, kept long-run slots, kept long-run groups
    uint32_t*       rank;
    uint32_t*       sa;
    uint32_t*       pos_out;       // next list
    uint32_t*       suf_out;
    uint32_t*       gid_out;
    uint32_t        gid_base;
    uint32_t        pos_mask;      // suffix = value & pos_mask (the bits above carry the BWT symbol)
    // initial only: the long-run suffixes leave through a separate list (switched off when the run marks overflowed)
    const uint64_t* allc;
    int             sigma;
    const unsigned* run_mark_count;
    unsigned        run_mark_cap;
    uint32_t*       run_pos_out;
    uint32_t*       run_suf_out;
};

// Every thread owns kGrpItems CONSECUTIVE slots (eight 16-byte loads), so a boundary is one 64-bit compare against the key in
// the previous register and the 16 flag bits of a thread never leave it; the neighbours' edge keys come by shuffle.  (The first
// version staged the tile in shared memory and balloted three flags per slot: 92 instructions per key, issue-bound at a third
// of the HBM rate.)
template <bool INITIAL>
__global__ void __launch_bounds__(kGrpThreads)
group_flags_kernel(GroupArgs a) {
    static_assert(kGrpItems == 16, "two threads fill one 32-bit word of the bit arrays");
    __shared__ uint64_t s_allc[256];
    __shared__ unsigned s_agg[kAggs][kGrpThreads / 32];
    const int64_t tile_base = (int64_t)blockIdx.x * kGrpTile;
    const int64_t t0 = tile_base + (int64_t)threadIdx.x * kGrpItems;
    const int cnt = (int)max((int64_t)0, min((int64_t)kGrpItems, a.m - t0));
    uint64_t key[kGrpItems];
    if (cnt == kGrpItems && (reinterpret_cast<uintptr_t>(a.keys) & 31) == 0) {
#pragma unroll
        for (int i = 0; i < kGrpItems / 4; i++) {
            unsigned long long q[4];
            ld_nc_256(a.keys + t0 + 4 * i, q);
            key[4 * i] = q[0]; key[4 * i + 1] = q[1]; key[4 * i + 2] = q[2]; key[4 * i + 3] = q[3];
        }
    } else {
#pragma unroll
        for (int i = 0; i < kGrpItems; i++) key[i] = i < cnt ? a.keys[t0 + i] : 0;
    }
    // the small table is fetched while the keys are on their way (a CTA lives for a few microseconds: a dependent global load
    // and a barrier in front of the key loads cost a sixth of it)
    int sigma = 0;
    if (INITIAL) {
        sigma = (a.sigma > 0 && *a.run_mark_count <= a.run_mark_cap) ? a.sigma : 0;
        if ((int)threadIdx.x < sigma) s_allc[threadIdx.x] = a.allc[threadIdx.x];
        __syncthreads();
    }
    // the key before my first slot and the key after my last one
    uint64_t prev = __shfl_up_sync(0xffffffffu, key[kGrpItems - 1], 1);
    uint64_t next = __shfl_down_sync(0xffffffffu, key[0], 1);
    if (lane_id() == 0 && cnt > 0 && t0 > 0) prev = a.keys[t0 - 1];
    if (lane_id() == 31 && t0 + kGrpItems < a.m) next = a.keys[t0 + kGrpItems];
    const unsigned valid = (1u << cnt) - 1u;
    // bit i of bnd: slot t0 + i starts a key group; bit kGrpItems: so does the slot after mine (the end of the keys closes the last group)
    unsigned bnd = 0;
#pragma unroll
    for (int i = 0; i < kGrpItems; i++) bnd |= (unsigned)(key[i] != (i == 0 ? prev : key[i - 1])) << i;
    if (t0 == 0) bnd |= 1u;
    bnd &= valid;
    if (t0 + cnt >= a.m || key[kGrpItems - 1] != next) bnd |= 1u << cnt;          // cnt < kGrpItems only at the end of the keys
    const unsigned b = bnd & valid;
    const unsigned open = valid & ~(bnd & (bnd >> 1));                             // shares its key with a neighbour
    unsigned run = 0;
    if (INITIAL && sigma > 0 && open) {
        bool r = false;
#pragma unroll
        for (int i = 0; i < kGrpItems; i++) {
            if ((open >> i) & 1) {
                if (i == 0 || ((b >> i) & 1)) r = allc_symbol(key[i], s_allc, sigma) != 0;   // one search per key group
                run |= (unsigned)r << i;
            }
        }
    }
    if (INITIAL && cnt > 0 && (t0 & (kSampleStep - 1)) == 0) {
#pragma unroll
        for (int j = 0; j < kSampleLevels; j++) {
            const int sh = kSampleShift + j * kSampleFanShift;
            if ((t0 & (((int64_t)1 << sh) - 1)) == 0) a.sample.level[j][t0 >> sh] = key[0];
        }
    }
    // two threads to a word of the bit arrays
    const unsigned hi_b = __shfl_down_sync(0xffffffffu, b, 1), hi_r = __shfl_down_sync(0xffffffffu, run, 1);
    if ((lane_id() & 1) == 0 && cnt > 0) {
        a.bnd_bits[t0 >> 5] = b | hi_b << 16;
        if (INITIAL) a.run_bits[t0 >> 5] = run | hi_r << 16;
    }
    // tile aggregates: in-tile slot of the last boundary + 1 (max), kept slots / groups of the general and of the long-run kind
    unsigned v[kAggs] = { b ? threadIdx.x * kGrpItems + 32 - __clz(b) : 0u, (unsigned)__popc(open & ~run), (unsigned)__popc(b & open & ~run),
                          (unsigned)__popc(run), (unsigned)__popc(b & run) };
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int q = 0; q < kAggs; q++) {
            const unsigned t = __shfl_xor_sync(0xffffffffu, v[q], o);
            v[q] = q == 0 ? max(v[q], t) : v[q] + t;
        }
    }
    if (lane_id() == 0) {
#pragma unroll
        for (int q = 0; q < kAggs; q++) s_agg[q][threadIdx.x >> 5] = v[q];
    }
    __syncthreads();
    if (threadIdx.x < kAggs) {
        unsigned t = 0;
        for (int w = 0; w < kGrpThreads / 32; w++) t = threadIdx.x == 0 ? max(t, s_agg[0][w]) : t + s_agg[threadIdx.x][w];
        // aggregate 0 travels as a global slot + 1 (0 = no boundary in the tile)
        if (threadIdx.x == 0 && t) t += (unsigned)tile_base;
        a.agg[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = t;
    }
}

template <bool INITIAL>
__global__ void __launch_bounds__(kGrpThreads)
group_apply_kernel(GroupArgs a) {
    __shared__ unsigned s_w[kAggs][kGrpThreads / 32];
    const int64_t tile_base = (int64_t)blockIdx.x * kGrpTile;
    const int64_t t0 = tile_base + (int64_t)threadIdx.x * kGrpItems;
    // flags of my kGrpItems consecutive slots (+ the boundary bit of the slot after them)
    unsigned valid = 0, boundary = 0, single = 0, run = 0;
    if (t0 < a.m) {
        const int64_t w = t0 >> 5;
        const int sh = (int)(t0 & 31);
        const int cnt = (int)min((int64_t)kGrpItems, a.m - t0);
        valid = (1u << cnt) - 1u;
        const int64_t words = (a.m + 31) >> 5;
        const unsigned long long two = (unsigned long long)a.bnd_bits[w] | (w + 1 < words ? (unsigned long long)a.bnd_bits[w + 1] << 32 : 0ull);
        unsigned bnd = (unsigned)(two >> sh) & ((2u << kGrpItems) - 1u);          // kGrpItems + 1 bits
        if (t0 + cnt >= a.m) bnd |= 1u << cnt;                                     // the end closes the last group
        boundary = bnd & valid;
        single = bnd & (bnd >> 1) & valid;
        if (INITIAL) {
            const unsigned long long rtwo = (unsigned long long)a.run_bits[w] | (w + 1 < words ? (unsigned long long)a.run_bits[w + 1] << 32 : 0ull);
            run = (unsigned)(rtwo >> sh) & valid & ~single;
        }
    }
    const unsigned my_last1 = boundary ? (unsigned)t0 + 32 - __clz(boundary) : 0u;      // global slot + 1
    const unsigned mine[4] = { (unsigned)__popc(valid & ~single & ~run), (unsigned)__popc(boundary & ~single & ~run),
                               (unsigned)__popc(run), (unsigned)__popc(boundary & run) };
    // block-wide exclusive scans of the per-thread aggregates
    unsigned il = my_last1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, il, o); if (lane_id() >= (unsigned)o) il = max(il, t); }
    unsigned inc[4];
#pragma unroll
    for (int q = 0; q < 4; q++) inc[q] = warp_incl_sum(mine[q]);
    if (lane_id() == 31) {
        s_w[0][threadIdx.x >> 5] = il;
#pragma unroll
        for (int q = 0; q < 4; q++) s_w[1 + q][threadIdx.x >> 5] = inc[q];
    }
    unsigned last1 = a.agg[blockIdx.x];                       // (requested before the barrier: five independent loads)
    unsigned cnt[4];
#pragma unroll
    for (int q = 0; q < 4; q++) cnt[q] = a.agg[(size_t)(1 + q) * gridDim.x + blockIdx.x];
    __syncthreads();
    for (unsigned w = 0; w < (threadIdx.x >> 5); w++) {
        last1 = max(last1, s_w[0][w]);
#pragma unroll
        for (int q = 0; q < 4; q++) cnt[q] += s_w[1 + q][w];
    }
    const unsigned prev_l = __shfl_up_sync(0xffffffffu, il, 1);
    if (lane_id() > 0) last1 = max(last1, prev_l);
#pragma unroll
    for (int q = 0; q < 4; q++) cnt[q] += inc[q] - mine[q];
    if ((valid & ~single) == 0) return;                                                // nothing unresolved here
    unsigned keep = cnt[0], groups = cnt[1], keep_run = cnt[2];
    unsigned last = last1 - 1;                                                         // group start of the slot before mine

    // the suffixes of my slots, requested together (unresolved slots come in long stretches: the members of a big group)
    uint32_t sv[kGrpItems];
    if (valid == (1u << kGrpItems) - 1u && (reinterpret_cast<uintptr_t>(a.suf) & 31) == 0) {
#pragma unroll
        for (int i = 0; i < kGrpItems / 8; i++) {
            unsigned long long q[4];
            ld_nc_256(a.suf + t0 + 8 * i, q);
#pragma unroll
            for (int j = 0; j < 4; j++) { sv[8 * i + 2 * j] = (uint32_t)q[j]; sv[8 * i + 2 * j + 1] = (uint32_t)(q[j] >> 32); }
        }
    } else {
#pragma unroll
        for (int i = 0; i < kGrpItems; i++) sv[i] = ((valid & ~single) >> i) & 1 ? a.suf[t0 + i] : 0u;
    }
#pragma unroll
    for (int i = 0; i < kGrpItems; i++) {
        if (!((valid >> i) & 1)) continue;                                             // valid slots are the low bits
        const uint32_t t = (uint32_t)t0 + i;
        if ((boundary >> i) & 1) last = t;
        if ((single >> i) & 1) continue;
        a.rank[sv[i] & a.pos_mask] = INITIAL ? last : a.pos[last];
        if (INITIAL && ((run >> i) & 1)) {
            a.run_pos_out[keep_run] = t;
            a.run_suf_out[keep_run] = sv[i];
            keep_run++;
        } else {
            if ((boundary >> i) & 1) groups++;
            a.pos_out[keep] = INITIAL ? t : a.pos[t];
            a.suf_out[keep] = sv[i];
            a.gid_out[keep] = a.gid_base + groups - 1;
            keep++;
        }
    }
}

// refine rounds: a suffix alone in its group is final (the first grouping leaves those where the sort put them)
__global__ void __launch_bounds__(256)
group_finish_kernel(GroupArgs a) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t words = (a.m + 31) >> 5;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < a.m; t += stride) {
        const bool b = (a.bnd_bits[t >> 5] >> (t & 31)) & 1u;
        const int64_t t1 = t + 1;
        const bool b_next = t1 >= a.m || ((a.bnd_bits[t1 >> 5] >> (t1 & 31)) & 1u);
        (void)words;
        if (b && b_next) {
            const uint32_t v = a.suf[t], slot = a.pos[t];
            a.rank[v & a.pos_mask] = slot;
            a.sa[slot] = v;
        }
    }
}

// sort key of a long-run suffix: symbol, then the side the run ends on, then the run length left —
// ascending when the run ends below its symbol, descending when it ends above (header, point 5)
__global__ void run_keys_kernel(const uint32_t* __restrict__ run_suf, int64_t m_run, uint32_t pos_mask, const Run* __restrict__ runs,
                                int n_runs, int k, int len_bits, const uint8_t* __restrict__ text, const uint8_t* __restrict__ code_of,
                                uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= m_run) return;
    const uint32_t sv = run_suf[u], s = sv & pos_mask;
    bool larger = false;
    const uint32_t r = run_remaining(runs, n_runs, s, k, &larger);
    const uint64_t len_mask = (1ull << len_bits) - 1;                    // r <= n < 2^len_bits
    const uint64_t order = larger ? ((1ull << len_bits) | (len_mask - r)) : (uint64_t)r;
    keys[u] = ((uint64_t)code_of[text[s]] << (len_bits + 1)) | order;
    vals[u] = sv;
}

// key of the suffix at q, straight from the text (what the first sort computed for it)
__device__ __forceinline__ uint64_t key_at(const uint8_t* __restrict__ text, int64_t n, const uint8_t* __restrict__ code_of,
                                           const KeyCoder& kc, int64_t q) {
    uint64_t key = 0;
    for (int j = 0; j < kc.k; j++) {
        const int64_t p = q + j;
        key = key * kc.radix + (p < n ? (uint64_t)code_of[text[p]] : 0ull);
    }
    return key;
}

// First slot whose key is >= want.  The index levels (written by the grouping pass) narrow the answer from the top: the
// highest level is small enough for a plain binary search; below it the last entry < want of a level has its successor among
// 16 consecutive entries of the next one — one 128-byte line, searched in four steps that touch it once — and the last level
// pins the answer to kSampleStep consecutive sorted keys.  A lookup costs about five cold lines instead of the seventeen
// sectors a binary search over every 64th key touched (2.1 GB of DRAM reads per 3 M lookups, 0.8 ms, in the first version).
__device__ __forceinline__ uint32_t key_slot(const uint64_t* __restrict__ sorted_keys, int64_t n, const SampleIndex& ix, uint64_t want) {
    int64_t lo = 0, hi = ix.top_count;                                     // first top-level entry >= want
    const uint64_t* top = ix.level[kSampleLevels - 1];
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(&top[mid]) < want) lo = mid + 1; else hi = mid;
    }
    if (lo == 0) return 0;                                                 // the first sorted key is >= want already
    int64_t p = lo - 1;                                                    // last entry < want
#pragma unroll
    for (int j = kSampleLevels - 2; j >= 0; j--) {
        const uint64_t* lv = ix.level[j];
        const int64_t first = p << kSampleFanShift;                        // lv[first] is the same key as the entry above: < want
        const int64_t cnt = min((int64_t)1 << kSampleFanShift, sample_level_count_dev(n, j) - first);
        int a = 1, b = (int)cnt;                                           // first entry >= want in (first, first + cnt)
        while (a < b) {
            const int mid = (a + b) >> 1;
            if (__ldg(&lv[first + mid]) < want) a = mid + 1; else b = mid;
        }
        p = first + a - 1;
    }
    lo = (p << kSampleShift) + 1;                                          // sorted_keys[lo - 1] < want
    hi = min((p + 1) << kSampleShift, n);                                  // sorted_keys[hi] >= want (or hi == n)
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(&sorted_keys[mid]) < want) lo = mid + 1; else hi = mid;
    }
    return (uint32_t)lo;
}

// rank[q] + 1 (0 past the end), looked up by its key when the suffix at q has been final since the first sort
__device__ __forceinline__ uint32_t rank_plus_one(uint32_t* __restrict__ rank, int64_t n, int64_t q, const uint8_t* __restrict__ text,
                                                  const uint8_t* __restrict__ code_of, const KeyCoder& kc,
                                                  const uint64_t* __restrict__ sorted_keys, const SampleIndex& sample) {
    if (q >= n) return 0;
    uint32_t rk = rank[q];
    if (rk == kNoRank) {
        // final since the first sort: its key is unique, its slot is where the key sits
        rk = key_slot(sorted_keys, n, sample, key_at(text, n, code_of, kc, q));
        rank[q] = rk;                               // every writer stores the same value
    }
    return rk + 1;
}

// A suffix inside a long run of r >= k symbols looks r - k symbols further than the others (header, point 5): at
// run end + h - k, the same place for every suffix of the run.  That rank is looked up ONCE per run and round here; in the
// first version each of the 20 M run suffixes of a chr1-sized block read rank[] at one of 42 addresses, found it unset, and
// thousands of them at a time repeated the same key search (2 GB of DRAM reads, 0.74 ms).
__global__ void run_ranks_kernel(const Run* __restrict__ runs, int n_runs, int64_t h, uint32_t* __restrict__ rank, int64_t n,
                                 const uint8_t* __restrict__ text, const uint8_t* __restrict__ code_of, KeyCoder kc,
                                 const uint64_t* __restrict__ sorted_keys, SampleIndex sample, uint32_t* __restrict__ run_low) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_runs) return;
    const int64_t q = (int64_t)(runs[j].end_side & 0x7FFFFFFFu) + h - kc.k;
    run_low[j] = rank_plus_one(rank, n, q, text, code_of, kc, sorted_keys, sample);
}

// refinement key: (group ordinal, rank of the suffix `h` symbols further, 0 when that is past the end).
__global__ void __launch_bounds__(256)
refine_keys_kernel(const uint32_t* __restrict__ suf, const uint32_t* __restrict__ gid, int64_t m,
                   uint32_t* __restrict__ rank, int64_t n, int64_t h, int low_bits, uint32_t pos_mask,
                   const Run* __restrict__ runs, int n_runs, const uint8_t* __restrict__ text,
                   const uint8_t* __restrict__ code_of, KeyCoder kc, const uint64_t* __restrict__ sorted_keys,
                   SampleIndex sample, const uint32_t* __restrict__ run_low,
                   uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    __shared__ uint8_t s_code_of[256];
    s_code_of[threadIdx.x] = code_of[threadIdx.x];
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < m; u += stride) {
        const uint32_t sv = suf[u], s = sv & pos_mask;
        uint64_t low;
        int run = -1;
        if (n_runs > 0) {
            bool larger;
            if (run_remaining(runs, n_runs, s, kc.k, &larger, &run) == 0) run = -1;
        }
        if (run >= 0) low = run_low[run];
        else low = rank_plus_one(rank, n, (int64_t)s + h, text, s_code_of, kc, sorted_keys, sample);
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
    const int k_min = (int)std::min<double>(kMaxKeySymbols, std::ceil((std::log2((double)n + 1) + 6.0) / h2));
    KeyCoder kc;
    kc.radix = radix;
    for (int passes = 1; passes <= 8; passes++) {
        int k = 0;
        unsigned __int128 pw = 1;                            // radix^k
        while (k < kMaxKeySymbols && pw * radix <= (one << (8 * passes))) { pw *= radix; k++; }
        kc.k = std::max(k, 1);
        if (k >= k_min) break;
    }
    kc.top = 1;
    for (int j = 1; j < kc.k; j++) kc.top *= radix;
    return kc;
}

}  // namespace

size_t suffix_sort_workspace_bytes(int64_t n) {
    // rank 4n + keys 16n + vals(other) 4n + refinement worst case (lists 24n, sort 12n, run list 8n) + run marks + sort temp
    const size_t tiles = (size_t)(n / kGrpTile + 2);
    // run marks: two arrays of n/8 + 1024 u64 (2 x n bytes) and the run list of half as many 8-byte entries (n/2 bytes);
    // every allocation below is rounded up to 256 bytes (the 8 MB at the end covers those)
    const size_t marks = ((size_t)n / 8 + 1024) * 8 * 2 + ((size_t)n / 16 + 513) * (sizeof(Run) + 4);
    return (size_t)n * (4 + 16 + 4 + 44) + marks + radix_sort_temp_bytes(n) + tiles * kAggs * 8 + row_scan_scratch_bytes(kAggs, (int64_t)tiles) + (size_t)n / 4 + sample_index_words(n) * 8 + (8 << 20);
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
    unsigned* d_agg = arena.get<unsigned>((size_t)tiles_n * kAggs + 8);
    void* d_scan = arena.raw(row_scan_scratch_bytes(kAggs, tiles_n));
    unsigned* d_bits = arena.get<unsigned>((size_t)(n / 32 + 2) * 2);
    uint64_t* d_sample = arena.get<uint64_t>(sample_index_words(n));
    long long* d_totals = arena.get<long long>(8);                 // [0..3] group totals, [4] run marks (as unsigned)
    uint64_t* d_marks[2] = { arena.get<uint64_t>(mark_cap), arena.get<uint64_t>(mark_cap) };
    Run* d_runs = arena.get<Run>(mark_cap / 2 + 1);
    uint32_t* d_run_low = arena.get<uint32_t>(mark_cap / 2 + 1);
    // first list and the long-run list: how many suffixes stay unresolved is only known after the grouping pass
    uint32_t* list0[3] = { arena.get<uint32_t>((size_t)n), arena.get<uint32_t>((size_t)n), arena.get<uint32_t>((size_t)n) };
    uint32_t* run_pos = arena.get<uint32_t>((size_t)n);
    uint32_t* run_suf = arena.get<uint32_t>((size_t)n);
    if (!d_code || !d_allc || !d_rank || !d_keys0 || !d_keys1 || !d_vals1 || !d_temp || !d_agg || !d_scan || !d_bits || !d_sample || !d_totals ||
        !d_marks[0] || !d_marks[1] || !d_runs || !d_run_low || !list0[0] || !list0[1] || !list0[2] || !run_pos || !run_suf)
        return fail(GCZ_E_NOMEM, "suffix sort workspace for n=%lld", (long long)n);
    unsigned* d_mark_count = reinterpret_cast<unsigned*>(d_totals + 4);

    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
    if (stats) {
        GCZ_CUDA(cudaEventCreate(&ev0)); GCZ_CUDA(cudaEventCreate(&ev1)); GCZ_CUDA(cudaEventCreate(&ev2));
        GCZ_CUDA(cudaEventRecord(ev0, st));
    }

    GCZ_TRY(small_upload(ctx, st, d_code, h_code, 256));
    GCZ_TRY(small_upload(ctx, st, d_allc, h_allc, sizeof(uint64_t) * sigma));
    GCZ_CUDA(cudaMemsetAsync(d_totals, 0, 8 * sizeof(long long), st));
    GCZ_CUDA(cudaMemsetAsync(d_rank, 0xFF, (size_t)n * 4, st));

    // full sort by the first k symbols; start in the buffer that makes the result land in d_sa
    const int npass = radix_sort_passes(key_bits);
    RadixBuffers b;
    b.keys[0] = d_keys0; b.keys[1] = d_keys1;
    b.vals[0] = d_sa;    b.vals[1] = d_vals1;
    b.cur = npass & 1;
    // keys and values are never written out unsorted: the first digit pass reads the text (TextKeySource)
    TextKeySource src;
    src.text = d_text; src.n = n; src.code_of = d_code; src.coder = kc; src.carry_shift = carry;
    src.run_marks = d_marks[0]; src.run_mark_count = d_mark_count; src.run_mark_cap = mark_cap;
    SortStats ss;
    SortStats* ssp = stats ? &ss : nullptr;
    GCZ_TRY(radix_sort_pairs(ctx, st, b, n, 0, key_bits, d_temp, ssp, &src));
    if (b.cur != 0) return fail(GCZ_E_INTERNAL, "initial sort landed in the wrong buffer");

    // groups of equal keys -> ranks, first list, long-run list.  More long runs than the mark buffer holds (a
    // text made of medium runs) switches the long-run path off on the device: plain doubling, where a "run"
    // suffix is an ordinary member of its key group.
    auto launch_group = [&](bool initial, GroupArgs& ga) -> int {
        const int64_t tiles = (ga.m + kGrpTile - 1) / kGrpTile;
        ga.bnd_bits = d_bits; ga.run_bits = d_bits + (n / 32 + 2); ga.agg = d_agg; ga.totals = d_totals;
        if (initial) {
            GCZ_LAUNCH(ctx, group_flags_kernel<true>, (unsigned)tiles, kGrpThreads, 0, st, ga);
            GCZ_TRY(row_scan(ctx, st, d_agg, kAggs, tiles, true, d_scan, nullptr, d_totals));
            GCZ_LAUNCH(ctx, group_apply_kernel<true>, (unsigned)tiles, kGrpThreads, 0, st, ga);
        } else {
            GCZ_LAUNCH(ctx, group_flags_kernel<false>, (unsigned)tiles, kGrpThreads, 0, st, ga);
            GCZ_TRY(row_scan(ctx, st, d_agg, kAggs, tiles, true, d_scan, nullptr, d_totals));
            GCZ_LAUNCH(ctx, group_apply_kernel<false>, (unsigned)tiles, kGrpThreads, 0, st, ga);
            const int grid = (int)std::min<int64_t>((ga.m + 255) / 256, (int64_t)ctx->sm_count * 16);
            GCZ_LAUNCH(ctx, group_finish_kernel, grid, 256, 0, st, ga);
        }
        return GCZ_OK;
    };
    GroupArgs ga;
    ga.keys = d_keys0; ga.suf = d_sa; ga.pos = nullptr; ga.m = n; ga.rank = d_rank; ga.sa = d_sa;
    ga.pos_out = list0[0]; ga.suf_out = list0[1]; ga.gid_out = list0[2]; ga.gid_base = 0; ga.pos_mask = pos_mask;
    ga.allc = d_allc; ga.sigma = sigma; ga.run_mark_count = d_mark_count; ga.run_mark_cap = mark_cap;
    ga.run_pos_out = run_pos; ga.run_suf_out = run_suf;
    SampleIndex six;
    {
        size_t at = 0;
        for (int j = 0; j < kSampleLevels; j++) { six.level[j] = d_sample + at; at += ((size_t)sample_level_count(n, j) + 15 + 16) & ~(size_t)15; }
        six.top_count = sample_level_count(n, kSampleLevels - 1);
    }
    ga.sample = six;
    GCZ_TRY(launch_group(true, ga));
    long long h_totals[5] = { 0, 0, 0, 0, 0 };
    GCZ_TRY(small_read(ctx, st, h_totals, d_totals, sizeof(h_totals)));
    int64_t m = h_totals[0], groups = h_totals[1];
    const int64_t m_run = h_totals[2];
    const int64_t n_marks = (int64_t)(uint32_t)h_totals[4];
    if (m_run > 0 && (n_marks > (int64_t)mark_cap || (n_marks & 1))) return fail(GCZ_E_INTERNAL, "long-run bookkeeping");
    const int n_runs = m_run > 0 ? (int)(n_marks / 2) : 0;
    if (stats) GCZ_CUDA(cudaEventRecord(ev1, st));

    // refinement buffers: the initial sort's second key/value arrays are dead, its sorted keys stay (point 4)
    const int64_t m0 = m + m_run;
    const size_t cap = (size_t)std::max<int64_t>(m0, 1);
    uint32_t* list_pos[2] = { list0[0], arena.get<uint32_t>(cap) };
    uint32_t* list_suf[2] = { list0[1], arena.get<uint32_t>(cap) };
    uint32_t* list_gid[2] = { list0[2], arena.get<uint32_t>(cap) };
    uint64_t* r_keys1 = arena.get<uint64_t>(cap);
    uint32_t* r_vals1 = arena.get<uint32_t>(cap);
    if (!list_pos[1] || !list_suf[1] || !list_gid[1] || !r_keys1 || !r_vals1)
        return fail(GCZ_E_NOMEM, "suffix sort refinement lists for %lld unresolved suffixes", (long long)m0);
    uint64_t* r_keys0 = d_keys1;
    uint32_t* r_vals0 = d_vals1;

    int rounds = 0;
    if (m_run > 0) {
        // the long-run suffixes: one sort by (symbol, side, run length left); what stays tied joins the list
        RadixBuffers mb;
        mb.keys[0] = d_marks[0]; mb.keys[1] = d_marks[1];
        GCZ_TRY(radix_sort_pairs(ctx, st, mb, n_marks, 0, bits_for(2ull * (uint64_t)n + 1), d_temp, nullptr));
        GCZ_LAUNCH(ctx, pair_runs_kernel, (unsigned)((n_runs + 255) / 256), 256, 0, st, mb.keys[mb.cur], (int64_t)n_runs, d_text, n, d_runs);
        RadixBuffers rb;
        rb.keys[0] = r_keys0; rb.keys[1] = r_keys1;
        rb.vals[0] = r_vals0; rb.vals[1] = r_vals1;
        rb.cur = 0;
        const int len_bits = bits_for((uint64_t)n);
        GCZ_LAUNCH(ctx, run_keys_kernel, (unsigned)((m_run + 255) / 256), 256, 0, st, run_suf, m_run, pos_mask, d_runs, n_runs, k,
                   len_bits, d_text, d_code, rb.keys[0], rb.vals[0]);
        GCZ_TRY(radix_sort_pairs(ctx, st, rb, m_run, 0, len_bits + 1 + bits_for((uint64_t)sigma), d_temp, ssp));
        GroupArgs gr = ga;
        gr.keys = rb.keys[rb.cur]; gr.suf = rb.vals[rb.cur]; gr.pos = run_pos; gr.m = m_run;
        gr.pos_out = list_pos[0] + m; gr.suf_out = list_suf[0] + m; gr.gid_out = list_gid[0] + m; gr.gid_base = (uint32_t)groups;
        GCZ_TRY(launch_group(false, gr));
        GCZ_TRY(small_read(ctx, st, h_totals, d_totals, 2 * sizeof(long long)));
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
        if (n_runs > 0) GCZ_LAUNCH(ctx, run_ranks_kernel, (unsigned)((n_runs + 127) / 128), 128, 0, st, d_runs, n_runs, h, d_rank, n, d_text, d_code, kc,
                                   d_keys0, six, d_run_low);
        GCZ_LAUNCH(ctx, refine_keys_kernel, grid, 256, 0, st, list_suf[cur], list_gid[cur], m, d_rank, n, h, low_bits, pos_mask,
                   d_runs, n_runs, d_text, d_code, kc, d_keys0, six, d_run_low, rb.keys[0], rb.vals[0]);
        const int gid_bits = bits_for((uint64_t)std::max<int64_t>(groups - 1, 0));
        GCZ_TRY(radix_sort_pairs(ctx, st, rb, m, 0, low_bits + gid_bits, d_temp, ssp));
        GroupArgs gr = ga;
        gr.keys = rb.keys[rb.cur]; gr.suf = rb.vals[rb.cur]; gr.pos = list_pos[cur]; gr.m = m;
        gr.pos_out = list_pos[cur ^ 1]; gr.suf_out = list_suf[cur ^ 1]; gr.gid_out = list_gid[cur ^ 1]; gr.gid_base = 0;
        GCZ_TRY(launch_group(false, gr));
        GCZ_TRY(small_read(ctx, st, h_totals, d_totals, 2 * sizeof(long long)));
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
        for (size_t i = 0; i < ss.pass_ms.size() && i < ss.pass_elements.size(); i++) {
            if (ss.pass_elements[i] == n) { stats->radix_full_passes++; stats->radix_full_ms += ss.pass_ms[i]; }
            else if (ss.pass_elements[i] == -n) stats->radix_text_ms += ss.pass_ms[i];
        }
        stats->symbols_per_key = k;
        stats->unresolved_after_first_sort = m0;
        stats->long_runs = n_runs;
        cudaEventDestroy(ev0); cudaEventDestroy(ev1); cudaEventDestroy(ev2);
    }
    arena.release(mark0);
    return GCZ_OK;
}

}  // namespace gcz

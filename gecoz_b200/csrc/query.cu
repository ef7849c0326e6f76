// Query side: open a block (GSSA), batched backward search (count) and batched locate / find.
//
// Replaces algo/ssa/GSSA.java (search :187-208, locate :241-251, index :215-239, find :160-185),
// algo/tree/HuffmanShapedWaveletTree.java (occ :247-267, getRS :300-314), algo/tree/RankedWTNode.java
// (get :81-84, count :98-122), algo/ssa/GSSAIndex.java (get :171-173) and
// algo/tree/IndexWaveletTree.java (get :127-144).
//
// Device layout: the .gcz/.gcx bytes are NOT queried in place.  gcz_open_block re-lays every ranked bit
// vector out as 32-byte "rank sectors" — one uint32 = ones before the sector, then 224 data bits — so that
// bit + rank at a position cost exactly one aligned DRAM sector instead of the file format's unaligned
// 8 + 2 + up-to-64 bytes.  Ranks, and therefore every result, are identical; the counters of the file are
// what seeds the sector counts, so a corrupt file misbehaves the same way it would in the reference.
#include "query.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace gcz {

namespace {

constexpr int kSectorBits = 224;

// ---- literal rank on the file layout (only used while opening) -------------------------------------------
__device__ __forceinline__ unsigned long long load_le(const uint8_t* p, int nbytes) {
    unsigned long long v = 0;
    for (int i = 0; i < nbytes; i++) v |= (unsigned long long)p[i] << (8 * i);
    return v;
}

// RankedWTNode.count(idx): ones in [0, idx]   algo/tree/RankedWTNode.java:98-122
__device__ long long file_rank(const uint8_t* buf, long long nbytes, long long idx) {
    long long count = 0;
    const long long nlidx = idx >> 16, nsidx = (idx >> 9) & 127;
    long long lpos = 0;
    if (nlidx > 0) { lpos = nlidx * 8454; count = (long long)load_le(buf + lpos - 8, 8); }
    long long bpos = lpos + nsidx * 66;
    if (nsidx > 0) count += (long long)load_le(buf + bpos - 2, 2);
    const long long last = bpos + ((idx >> 3) & 56);
    for (; bpos < last; bpos += 8) count += __popcll(load_le(buf + bpos, 8));
    const int avail = (int)min((long long)8, nbytes - bpos);
    const unsigned long long w = load_le(buf + bpos, avail);
    return count + __popcll(w << (63 - (idx & 63)));
}

__global__ void file_rank_kernel(const uint8_t* buf, long long nbytes, long long idx, long long* out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) *out = file_rank(buf, nbytes, idx);
}

struct RelayoutDesc {
    const uint8_t* src;     // ranked vector in file layout
    int64_t  src_bytes;
    int64_t  len;           // bits
    uint64_t sector0;       // first sector of this vector in the index's sector array
    uint64_t sectors;       // ceil(len / 224)
};

__global__ void relayout_kernel(const RelayoutDesc* __restrict__ descs, int ndesc, uint64_t total_sectors,
                                uint32_t* __restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total_sectors; g += stride) {
        int v = 0;
        while (v + 1 < ndesc && descs[v + 1].sector0 <= g) v++;
        const RelayoutDesc d = descs[v];
        if (g - d.sector0 >= d.sectors) continue;             // pad sector between vectors stays zero
        const int64_t first_bit = (int64_t)(g - d.sector0) * kSectorBits;
        uint32_t w[8];
        w[0] = first_bit > 0 ? (uint32_t)file_rank(d.src, d.src_bytes, first_bit - 1) : 0u;
#pragma unroll
        for (int k = 0; k < 7; k++) {
            const int64_t b = first_bit + 32 * k;
            uint32_t word = 0;
            if (b < d.len) {
                const int valid = (int)min((int64_t)32, d.len - b);
                const int64_t off = (b >> 3) + (b >> 9) * 2 + (b >> 16) * 6;
                word = (uint32_t)load_le(d.src + off, (valid + 7) >> 3);
                if (valid < 32) word &= (1u << valid) - 1u;
            }
            w[1 + k] = word;
        }
        uint4* o = reinterpret_cast<uint4*>(out + g * 8);
        o[0] = make_uint4(w[0], w[1], w[2], w[3]);
        o[1] = make_uint4(w[4], w[5], w[6], w[7]);
    }
}

// ---- rank sector access ---------------------------------------------------------------------------------
struct BitRank { unsigned bit; unsigned rank; };       // rank = ones in [0, pos]

__device__ __forceinline__ BitRank sector_bit_rank(const uint32_t* __restrict__ sectors, uint64_t sector0, uint32_t pos) {
    const uint32_t sec = pos / kSectorBits, off = pos - sec * kSectorBits;
    const uint4* p = reinterpret_cast<const uint4*>(sectors + (sector0 + sec) * 8);
    const uint4 a = __ldg(p), b = __ldg(p + 1);
    const uint32_t w[8] = { a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w };
    const uint32_t wi = off >> 5, bi = off & 31;
    uint32_t r = w[0];
#pragma unroll
    for (int k = 0; k < 7; k++) {
        const uint32_t word = w[1 + k];
        r += (uint32_t)k < wi ? __popc(word) : 0u;
    }
    uint32_t cur = 0;
#pragma unroll
    for (int k = 0; k < 7; k++) cur = (uint32_t)k == wi ? w[1 + k] : cur;
    r += __popc(cur << (31 - bi));
    BitRank br;
    br.bit = (cur >> bi) & 1u;
    br.rank = r;
    return br;
}

// HSWT.occ(symbol, pos)  algo/tree/HuffmanShapedWaveletTree.java:247-267: rank(symbol, [0..pos]) - 1, or -1
__device__ __forceinline__ long long hswt_occ(const QueryTables* __restrict__ t, const uint32_t* __restrict__ sectors,
                                              int symbol, long long pos) {
    const int len = t->len[symbol];
    if (len == 0) return -1;
    const unsigned code = t->code[symbol];
    for (int d = 0; d < len && pos >= 0; d++) {
        const int v = t->node_of[symbol][d];
        const BitRank br = sector_bit_rank(sectors, t->node_sector0[v], (uint32_t)pos);
        pos = ((code >> d) & 1u) ? (long long)br.rank - 1 : pos - (long long)br.rank;
    }
    return pos;
}

// ones in [0, off] of one loaded rank sector
__device__ __forceinline__ uint32_t sector_rank_at(const uint32_t (&w)[8], uint32_t off) {
    const uint32_t wi = off >> 5, bi = off & 31;
    uint32_t r = w[0], cur = 0;
#pragma unroll
    for (int k = 0; k < 7; k++) {
        r += (uint32_t)k < wi ? __popc(w[1 + k]) : 0u;
        cur = (uint32_t)k == wi ? w[1 + k] : cur;
    }
    return r + __popc(cur << (31 - bi));
}

// What a batch of backward searches touches (gcz_count_stats): rank sectors of 32 bytes actually loaded, and the
// RankedWTNode.count calls the reference's loop makes for the same patterns (2 per character and code bit while the
// position is >= 0, algo/tree/HuffmanShapedWaveletTree.java:247-267) — 74 bytes each in the file layout.
struct OccStats { unsigned long long sectors = 0, ref_calls = 0, steps = 0; };

// The two occ() of one backward-search step, occ(symbol, p1) and occ(symbol, p2) with p1 <= p2, walking the
// symbol's path once: on every node the two positions usually fall into the same rank sector (always, once the
// interval has shrunk to a few suffixes), and that sector is loaded once.
template <bool STATS>
__device__ __forceinline__ void hswt_occ2(const QueryTables* __restrict__ t, const uint32_t* __restrict__ sectors,
                                          int symbol, long long& p1, long long& p2, OccStats* stats) {
    const int len = t->len[symbol];
    if (len == 0) { p1 = -1; p2 = -1; return; }
    const unsigned code = t->code[symbol];
    for (int d = 0; d < len && p2 >= 0; d++) {
        const uint64_t s0 = t->node_sector0[t->node_of[symbol][d]];
        const bool one = (code >> d) & 1u;
        const uint32_t q2 = (uint32_t)p2, sec2 = q2 / kSectorBits;
        uint32_t w[8];
        {
            const uint4* p = reinterpret_cast<const uint4*>(sectors + (s0 + sec2) * 8);
            const uint4 a = __ldg(p), b = __ldg(p + 1);
            w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
        }
        if (STATS) { stats->sectors++; stats->ref_calls += p1 >= 0 ? 2 : 1; }
        const uint32_t r2 = sector_rank_at(w, q2 - sec2 * kSectorBits);
        if (p1 >= 0) {
            const uint32_t q1 = (uint32_t)p1, sec1 = q1 / kSectorBits;
            uint32_t r1;
            if (sec1 == sec2) {
                r1 = sector_rank_at(w, q1 - sec1 * kSectorBits);
            } else {
                r1 = sector_bit_rank(sectors, s0, q1).rank;
                if (STATS) stats->sectors++;
            }
            p1 = one ? (long long)r1 - 1 : p1 - (long long)r1;
        }
        p2 = one ? (long long)r2 - 1 : p2 - (long long)r2;
    }
    if (p2 < 0) p1 = -1;                                   // p1 <= p2 throughout
}

// ---- count: backward search ---------------------------------------------------------------------------------
// Patterns differ in length (15..100 symbols) and half of a typical batch dies after ~14 steps, so a lane per pattern
// would idle most of the time.  Lanes are refilled instead: a lane that finishes its pattern takes the next one from its
// warp's reservation (64 patterns per atomic on the global counter), and every trip of the loop is one backward-search
// step for all 32 lanes.
// MODE 0: (sp, ep) per pattern.  MODE 1: totals[q] += max(0, ep - sp + 1) — one block after the other of a multi-block
// index into the same array (gcz_count_multi).  MODE 2: as 0, and the OccStats of the whole launch (gcz_count_stats).
template <int MODE>
__global__ void __launch_bounds__(256, 5)      // 48 registers: 5 CTAs per SM measured 8 % faster than the compiler's 50, 6 (40, spills) slower
count_kernel(const QueryTables* __restrict__ tables, const uint32_t* __restrict__ sectors,
             const uint8_t* __restrict__ pats, const int64_t* __restrict__ pat_off, int64_t n_pats,
             int64_t* __restrict__ sp_out, int64_t* __restrict__ ep_out, unsigned long long* __restrict__ next_pattern,
             unsigned long long* __restrict__ stats_out, int use_table) {
    __shared__ QueryTables t;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(tables);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&t);
        for (int i = threadIdx.x; i < (int)(sizeof(QueryTables) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    constexpr int kReserve = 64;
    const unsigned lane = lane_id(), lt = lanemask_lt();
    long long wnext = 0, wend = 0;                       // the warp's reservation [wnext, wend), uniform across lanes
    long long q = -1, i = 0, b = 0, sp = 0, ep = -1;     // q: my pattern, -1 = none
    bool drained = false;                                // the global counter ran past n_pats
    OccStats stats;
    while (true) {
        // hand patterns to the idle lanes
        unsigned idle = __ballot_sync(0xffffffffu, q < 0);
        while (idle && !drained) {
            if (wnext >= wend) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(next_pattern, (unsigned long long)kReserve);
                base = __shfl_sync(0xffffffffu, base, 0);
                wnext = (long long)base;
                wend = min((long long)base + kReserve, (long long)n_pats);
                if (wnext >= wend) { drained = true; break; }
            }
            const long long mine = wnext + __popc(idle & lt);
            if (q < 0 && mine < wend) {
                q = mine;
                b = pat_off[q];
                const long long e = pat_off[q + 1];
                sp = 0; ep = -1; i = b - 1;              // empty pattern / byte >= 0x80: reported as not found
                if (e > b) {
                    const int ch = pats[e - 1];
                    if (ch < 128) {                      // the reference indexes c[] with a signed byte
                        sp = t.c[ch];
                        ep = (ch < 255 ? t.c[ch + 1] : t.n) - 1;
                        i = e - 2;
                    }
                    // the last kmer_k symbols at once, when the table has a non-empty interval for them (an empty one means
                    // the search fails inside them: it is then made step by step, to report what the reference reports)
                    const int K = use_table ? t.kmer_k : 0;
                    if (K > 0 && e - b >= K) {
                        unsigned at = 0, bad = 0;
                        for (int j = K; j >= 1; j--) {
                            const unsigned c2 = t.code2[pats[e - j]];
                            bad |= c2;
                            at = at << 2 | (c2 & 3u);
                        }
                        if (bad < 4u) {
                            const uint2 iv = __ldg(&t.kmer[at]);
                            if (MODE == 2) stats.sectors++;
                            if (iv.x <= iv.y) { sp = iv.x; ep = iv.y; i = e - K - 1; }
                        }
                    }
                }
            }
            wnext = min(wend, wnext + (long long)__popc(idle));
            idle = __ballot_sync(0xffffffffu, q < 0);
        }
        if (__ballot_sync(0xffffffffu, q >= 0) == 0) break;
        if (q >= 0) {
            if (sp <= ep && i >= b) {                    // GSSA.search :193-196, one symbol
                const int ch = pats[i];
                if (ch >= 128) {
                    sp = 0; ep = -1;
                } else {
                    long long o1 = sp - 1, o2 = ep;
                    hswt_occ2<MODE == 2>(&t, sectors, ch, o1, o2, &stats);
                    if (MODE == 2) stats.steps++;
                    sp = t.c[ch] + o1 + 1;
                    ep = t.c[ch] + o2;
                }
                i--;
            } else {
                if (MODE == 1) {
                    if (ep >= sp) sp_out[q] += ep - sp + 1;
                } else {
                    sp_out[q] = sp;
                    ep_out[q] = ep;
                }
                q = -1;
            }
        }
    }
    if (MODE == 2) {
        // with the table: the sectors this kernel loads; without: the rank calls and steps of the reference's loop
        unsigned long long v[3] = { use_table ? stats.sectors : 0ull, use_table ? 0ull : stats.ref_calls, use_table ? 0ull : stats.steps };
#pragma unroll
        for (int k = 0; k < 3; k++) {
            for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
            if (lane == 0 && v[k]) atomicAdd(&stats_out[k], v[k]);
        }
    }
}

// (sp, ep) of every string of K symbols out of A, C, G, T: thread `at` searches the string whose 2-bit codes are the digits of
// `at` (first symbol highest) exactly as count_kernel would, last symbol first
__global__ void __launch_bounds__(256)
kmer_table_kernel(const QueryTables* __restrict__ tables, const uint32_t* __restrict__ sectors, int K, uint2* __restrict__ out) {
    __shared__ QueryTables t;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(tables);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&t);
        for (int i = threadIdx.x; i < (int)(sizeof(QueryTables) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const unsigned at = blockIdx.x * blockDim.x + threadIdx.x;
    if (at >= 1u << (2 * K)) return;
    const int sym[4] = { 'A', 'C', 'G', 'T' };
    int ch = sym[at & 3u];
    long long sp = t.c[ch], ep = t.c[ch + 1] - 1;
    for (int j = 1; j < K && sp <= ep; j++) {
        ch = sym[(at >> (2 * j)) & 3u];
        long long o1 = sp - 1, o2 = ep;
        hswt_occ2<false>(&t, sectors, ch, o1, o2, nullptr);
        sp = t.c[ch] + o1 + 1;
        ep = t.c[ch] + o2;
    }
    out[at] = sp <= ep ? make_uint2((unsigned)sp, (unsigned)ep) : make_uint2(1u, 0u);
}

// ---- locate ------------------------------------------------------------------------------------------------------
// IndexWaveletTree.get  algo/tree/IndexWaveletTree.java:127-144
__device__ __forceinline__ long long iwt_get(const QueryTables* __restrict__ t, const uint32_t* __restrict__ sectors, long long pos) {
    long long code = 0;
    int block = 0;
    for (int i = t->iwt_levels - 1; i >= 0; i--) {
        const BitRank br = sector_bit_rank(sectors, t->iwt_sector0[i], (uint32_t)pos);
        long long bits = br.rank;
        code = (code << 1) | br.bit;
        if (br.bit == 0) {
            bits = pos - bits - ((unsigned)block >> 1);
        } else {
            bits -= ((unsigned)block >> 1) + 1;
            block += 1 << i;
        }
        pos = block + bits;
    }
    return code;
}


// ---- select / inverse lookups (extract) -----------------------------------------------------------------------------
// Position of the n-th (1-based, counted from the start of the vector) bit equal to `want` among bits [lo, hi] of a
// sector vector, or -1.  RankedWTNode.findZero/findOne(n, lo, hi)  algo/tree/RankedWTNode.java:154-205: the
// reference steers an interpolation search with doubles; the position it returns is the exact select computed here.
__device__ long long sector_select(const uint32_t* __restrict__ sectors, uint64_t sector0, long long lo, long long hi,
                                   long long n, unsigned want) {
    if (hi < lo || n <= 0) return -1;
    long long jlo = lo / kSectorBits, jhi = hi / kSectorBits;
    auto before = [&](long long j) -> long long {
        const long long ones = __ldg(sectors + (sector0 + (uint64_t)j) * 8);
        return want ? ones : j * kSectorBits - ones;
    };
    if (before(jlo) >= n) return -1;
    while (jlo < jhi) {                                     // last sector with fewer than n matches before it
        const long long mid = (jlo + jhi + 1) >> 1;
        if (before(mid) < n) jlo = mid; else jhi = mid - 1;
    }
    long long r = n - before(jlo);
    const uint4* p = reinterpret_cast<const uint4*>(sectors + (sector0 + (uint64_t)jlo) * 8);
    const uint4 a = __ldg(p), b = __ldg(p + 1);
    const uint32_t w[7] = { a.y, a.z, a.w, b.x, b.y, b.z, b.w };
    for (int k = 0; k < 7; k++) {
        const uint32_t m = want ? w[k] : ~w[k];
        const int c = __popc(m);
        if (r <= c) {
            const long long pos = jlo * kSectorBits + 32 * k + __fns(m, 0, (int)r);
            return (pos >= lo && pos <= hi) ? pos : -1;
        }
        r -= c;
    }
    return -1;
}

// IndexWaveletTree.find  algo/tree/IndexWaveletTree.java:152-165: where the value idx sits in the sequence
__device__ long long iwt_find(const QueryTables* __restrict__ t, const uint32_t* __restrict__ sectors, long long idx) {
    long long pos = 0;
    for (int i = 0; i < t->iwt_levels; i++) {
        const unsigned bit = (unsigned)((unsigned long long)idx >> i) & 1u;
        const long long block = (long long)((unsigned long long)idx & (0xFFFFFFFFFFFFFFFEull << i));
        const long long hi = min(block + (2ll << i), (long long)t->iwt_m) - 1;
        pos = sector_select(sectors, t->iwt_sector0[i], block, hi, (long long)((unsigned long long)block >> 1) + pos + 1, bit);
        pos -= block;
    }
    return pos;
}

// GSSAIndex.find  algo/ssa/GSSAIndex.java:184-187: the SA row of a sampled text position
__device__ long long index_find(const QueryTables* __restrict__ t, const uint32_t* __restrict__ sectors, long long text_pos) {
    const long long sidx = text_pos >> t->sampling_factor;
    if (text_pos != (sidx << t->sampling_factor)) return (long long)INT_MIN;
    return sector_select(sectors, t->marker_sector0, 0, t->n - 1, iwt_find(t, sectors, sidx) + 1, 1u);
}

// One step of the LF walk: HSWT.getRS  algo/tree/HuffmanShapedWaveletTree.java:300-314 followed by
// idx = (int)(c[symbol] + rank)  algo/ssa/GSSA.java:116,123.  Returns the BWT symbol of row idx.
__device__ __forceinline__ int lf_step(const QueryTables* __restrict__ t, const uint32_t* __restrict__ sectors, long long& idx) {
    long long pos = idx;
    int v = 0;
    while (v >= 0) {
        const BitRank br = sector_bit_rank(sectors, t->node_sector0[v], (uint32_t)pos);
        pos = br.bit ? (long long)br.rank - 1 : pos - (long long)br.rank;
        v = t->child[v][br.bit];
    }
    const int symbol = ~v;
    idx = (long long)(int)(t->c[symbol] + pos);
    return symbol;
}

// GSSA.extract  algo/ssa/GSSA.java:90-126 for generalized-string positions [from, pos] into out[0 .. pos - from].
// Thread 0 does what the reference does at the top: it starts at the sampled position after `pos` (sapos, or row 0
// when that is past the text), skips down to pos and emits until the last sampled position low <= pos.  Every other
// thread owns one sampled position s in (from, low]: row by index_find, then up to 2^sf symbols backwards.  Both are
// the same walk as long as thread 0 reaches the row index_find(low) names; flag[0] tells whether it did (it does not
// when the walk crossed a separator the LF mapping gets wrong) and flag[1] keeps its row.
__global__ void __launch_bounds__(128)
extract_kernel(const QueryTables* __restrict__ tables, const uint32_t* __restrict__ sectors, long long from, long long pos,
               uint8_t* __restrict__ out, long long* __restrict__ flag) {
    __shared__ QueryTables t;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(tables);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&t);
        for (int i = threadIdx.x; i < (int)(sizeof(QueryTables) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const int sf = t.sampling_factor;
    const long long step = 1ll << sf;
    const long long low = max(from, (pos >> sf) << sf);               // thread 0 emits [low, pos]
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g == 0) {
        const long long sapos = ((pos >> sf) + 1) << sf;
        long long idx = sapos < (long long)t.n ? index_find(&t, sectors, sapos) : 0;
        long long n = min(sapos, (long long)t.n - 1) - pos;
        while (--n > 0) lf_step(&t, sectors, idx);
        for (long long p = pos; p >= low; p--) out[p - from] = (uint8_t)lf_step(&t, sectors, idx);
        // idx is now the row of position low (as the reference sees it)
        flag[1] = idx;
        flag[0] = (low > from && idx != index_find(&t, sectors, low)) ? 1 : 0;
        return;
    }
    // anchor g - 1 counts sampled positions downwards from low
    const long long s = low - (g - 1) * step;
    if (s <= from || low <= from) return;
    long long idx = index_find(&t, sectors, s);
    const long long stop = max(from, s - step);
    for (long long p = s - 1; p >= stop; p--) out[p - from] = (uint8_t)lf_step(&t, sectors, idx);
}

// the reference's own sequential walk, for the calls where it leaves the text's true LF chain
__global__ void extract_sequential_kernel(const QueryTables* __restrict__ tables, const uint32_t* __restrict__ sectors,
                                          long long from, long long low, long long row, uint8_t* __restrict__ out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    long long idx = row;
    for (long long p = low - 1; p >= from; p--) out[p - from] = (uint8_t)lf_step(tables, sectors, idx);
}

// GSSA.locate  algo/ssa/GSSA.java:241-251 with GSSAIndex.get :171-173 and HSWT.getRS :300-314
__device__ long long locate_row(const QueryTables* __restrict__ t, const uint32_t* __restrict__ sectors, long long idx) {
    long long steps = 0;
    while (true) {
        const BitRank mk = sector_bit_rank(sectors, t->marker_sector0, (uint32_t)idx);
        if (mk.bit) return (iwt_get(t, sectors, (long long)mk.rank - 1) << t->sampling_factor) + steps;
        if (steps > t->n) return -1;                     // the reference would never return here
        // LF step: walk the tree from the root reading bit + rank in every node on the way down
        long long pos = idx;
        int v = 0;
        while (v >= 0) {
            const BitRank br = sector_bit_rank(sectors, t->node_sector0[v], (uint32_t)pos);
            pos = br.bit ? (long long)br.rank - 1 : pos - (long long)br.rank;
            v = t->child[v][br.bit];
        }
        const int symbol = ~v;
        idx = (long long)(int)(t->c[symbol] + pos);
        steps++;
    }
}

__global__ void __launch_bounds__(256)
locate_kernel(const QueryTables* __restrict__ tables, const uint32_t* __restrict__ sectors,
              const int64_t* __restrict__ rows, int64_t n_rows, int64_t* __restrict__ out) {
    __shared__ QueryTables t;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(tables);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&t);
        for (int i = threadIdx.x; i < (int)(sizeof(QueryTables) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n_rows; q += stride) {
        const int64_t r = rows[q];
        out[q] = (r >= 0 && r < t.n) ? locate_row(&t, sectors, r) : -1;
    }
}

// rows of every occurrence of a chunk of patterns, tagged with the pattern, located in one launch:
// key = (pattern ordinal in chunk << 32) | text position
__global__ void __launch_bounds__(256)
locate_occurrences_kernel(const QueryTables* __restrict__ tables, const uint32_t* __restrict__ sectors,
                          const int64_t* __restrict__ sp, const int64_t* __restrict__ occ_excl /* per pattern, exclusive */,
                          int64_t first_pat, int64_t n_chunk_pats, int64_t base_occ, int64_t n_occ,
                          uint64_t* __restrict__ keys) {
    __shared__ QueryTables t;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(tables);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&t);
        for (int i = threadIdx.x; i < (int)(sizeof(QueryTables) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n_occ; o += stride) {
        // pattern p with occ_excl[p] <= base_occ + o < occ_excl[p + 1]
        const int64_t target = base_occ + o;
        int64_t lo = first_pat, hi = first_pat + n_chunk_pats;       // invariant: occ_excl[lo] <= target < occ_excl[hi]
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (occ_excl[mid] <= target) lo = mid; else hi = mid;
        }
        const int64_t row = sp[lo] + (target - occ_excl[lo]);
        const long long pos = locate_row(&t, sectors, row);
        keys[o] = ((uint64_t)(lo - first_pat) << 32) | (uint64_t)(uint32_t)pos;
    }
}

// symbol counts through occ(i, n - 1), the way GSSA.index derives C[]  algo/ssa/GSSA.java:215-226
__global__ void symbol_occ_kernel(const QueryTables* __restrict__ tables, const uint32_t* __restrict__ sectors, long long* __restrict__ out) {
    const int s = threadIdx.x;
    out[s] = hswt_occ(tables, sectors, s, tables->n - 1);
}

int launch_grid(DeviceCtx* ctx, int64_t work, int threads) {
    return (int)std::max<int64_t>(1, std::min<int64_t>((work + threads - 1) / threads, (int64_t)ctx->sm_count * 32));
}

}  // namespace

// Copies `bytes` from a host-or-device pointer into fresh arena memory.
static int to_device(cudaStream_t st, Arena& arena, const uint8_t* src, int64_t bytes, const uint8_t** out) {
    if (is_device_ptr(src)) { *out = src; return GCZ_OK; }
    uint8_t* d = arena.get<uint8_t>((size_t)bytes + 16);
    if (!d) return fail(GCZ_E_NOMEM, "device staging of %lld bytes", (long long)bytes);
    GCZ_CUDA(cudaMemcpyAsync(d, src, (size_t)bytes, cudaMemcpyHostToDevice, st));
    *out = d;
    return GCZ_OK;
}

int open_block(DeviceCtx* ctx, const uint8_t* gcz_body, int64_t body_len, int64_t text_len,
               const uint8_t* gcx_body, int64_t gcx_len, gcz_index** out) {
    if (!gcz_body || !out || body_len <= 0 || text_len <= 0 || text_len > 0x7FFFFFFFll) return fail(GCZ_E_ARG, "open_block arguments");
    if (!gcx_body || gcx_len <= 0) return fail(GCZ_E_ARG, "the .gcx index is required: the reference cannot locate without it");
    std::lock_guard<std::mutex> lock(ctx->mu);
    GCZ_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = stream_of(ctx);

    // shape table: needs the first bytes on the host
    std::vector<uint8_t> head((size_t)std::min<int64_t>(body_len, 2048));
    if (is_device_ptr(gcz_body)) GCZ_CUDA(cudaMemcpy(head.data(), gcz_body, head.size(), cudaMemcpyDeviceToHost));
    else std::memcpy(head.data(), gcz_body, head.size());
    std::unique_ptr<gcz_index> idx(new gcz_index());
    idx->ctx = ctx;
    idx->n = text_len;
    GCZ_TRY(shape_read(head.data(), (int64_t)head.size(), &idx->shape));
    gcz_shape& sh = idx->shape;
    if (sh.n_nodes <= 0) return fail(GCZ_E_FORMAT, "empty tree");

    // sampling factor: smallest f whose index fits  algo/ssa/GSSAIndex.java:62-67
    int sf = 0;
    while (gcx_len < index_size(text_len, sf)) { if (++sf > 30) return fail(GCZ_E_FORMAT, "invalid index file"); }
    idx->sampling_factor = sf;
    const int64_t m = (text_len + ((int64_t)1 << sf) - 1) >> sf;
    const int levels = 64 - __builtin_clzll((uint64_t)m);

    ctx->arena.reset();
    const size_t need = (size_t)body_len + (size_t)gcx_len + (1 << 20);
    if (ctx->arena.capacity < need) GCZ_TRY(ctx->arena.reserve(need));
    const uint8_t *d_gcz = nullptr, *d_gcx = nullptr;
    GCZ_TRY(to_device(st, ctx->arena, gcz_body, body_len, &d_gcz));
    GCZ_TRY(to_device(st, ctx->arena, gcx_body, gcx_len, &d_gcx));
    long long* d_scalar = ctx->arena.get<long long>(256);
    if (!d_scalar) return fail(GCZ_E_NOMEM, "open scratch");

    // node lengths top-down: ones = count(len - 1)  algo/tree/HuffmanShapedWaveletTree.java:197-216
    // children in file order: the node after k is its 0-child if internal; sizes propagate through a stack
    std::vector<int> child0(sh.n_nodes, -1), child1(sh.n_nodes, -1);
    {
        // rebuild child links from (depth, prefix)
        for (int k = 0; k < sh.n_nodes; k++) {
            for (int j = 0; j < sh.n_nodes; j++) {
                if (sh.node_depth[j] == sh.node_depth[k] + 1 && (sh.node_prefix[j] & ((1 << sh.node_depth[k]) - 1)) == sh.node_prefix[k]) {
                    if ((sh.node_prefix[j] >> sh.node_depth[k]) & 1) child1[k] = j; else child0[k] = j;
                }
            }
        }
    }
    sh.node_bits[0] = text_len;
    int64_t off = sh.table_bytes;
    for (int k = 0; k < sh.n_nodes; k++) {
        const int64_t len = sh.node_bits[k];
        if (len <= 0) return fail(GCZ_E_FORMAT, "wavelet node %d is empty", k);
        const int64_t nb = ranked_bytes(len);
        if (off + nb > body_len) return fail(GCZ_E_FORMAT, "wavelet nodes overrun the block body");
        sh.node_offset[k] = off;
        GCZ_LAUNCH(ctx, file_rank_kernel, 1, 32, 0, st, d_gcz + off, (long long)nb, (long long)(len - 1), d_scalar);
        long long ones = 0;
        GCZ_CUDA(cudaMemcpyAsync(&ones, d_scalar, 8, cudaMemcpyDeviceToHost, st));
        GCZ_CUDA(cudaStreamSynchronize(st));
        if (ones < 0 || ones > len) return fail(GCZ_E_FORMAT, "wavelet node %d has a corrupt counter", k);
        if (child0[k] >= 0) sh.node_bits[child0[k]] = len - ones;
        if (child1[k] >= 0) sh.node_bits[child1[k]] = ones;
        off += nb;
    }
    sh.size = off;
    sh.length = text_len;
    if (ranked_bytes(text_len) + (int64_t)levels * ranked_bytes(m) > gcx_len) return fail(GCZ_E_FORMAT, "invalid index file");

    // ---- re-layout into rank sectors -------------------------------------------------------------------------
    QueryTables qt;
    std::memset(&qt, 0, sizeof(qt));
    std::vector<RelayoutDesc> descs;
    uint64_t total_sectors = 0;
    auto add = [&](const uint8_t* src, int64_t len) {
        RelayoutDesc d;
        d.src = src; d.src_bytes = ranked_bytes(len); d.len = len; d.sector0 = total_sectors;
        d.sectors = (uint64_t)((len + kSectorBits - 1) / kSectorBits);
        total_sectors += d.sectors + 1;                      // one zero pad sector between vectors
        descs.push_back(d);
        return d.sector0;
    };
    for (int k = 0; k < sh.n_nodes; k++) qt.node_sector0[k] = add(d_gcz + sh.node_offset[k], sh.node_bits[k]);
    qt.marker_sector0 = add(d_gcx, text_len);
    for (int l = 0; l < levels; l++) {
        const int h = levels - 1 - l;                          // levels are stored highest bit first
        qt.iwt_sector0[h] = add(d_gcx + ranked_bytes(text_len) + (int64_t)l * ranked_bytes(m), m);
    }
    GCZ_CUDA(cudaMalloc(&idx->d_sectors, (size_t)total_sectors * 32));
    GCZ_CUDA(cudaMemsetAsync(idx->d_sectors, 0, (size_t)total_sectors * 32, st));
    RelayoutDesc* d_descs = ctx->arena.get<RelayoutDesc>(descs.size());
    if (!d_descs) return fail(GCZ_E_NOMEM, "open scratch");
    GCZ_CUDA(cudaMemcpyAsync(d_descs, descs.data(), sizeof(RelayoutDesc) * descs.size(), cudaMemcpyHostToDevice, st));
    GCZ_CUDA(cudaStreamSynchronize(st));
    GCZ_LAUNCH(ctx, relayout_kernel, launch_grid(ctx, (int64_t)total_sectors, 256), 256, 0, st, d_descs, (int)descs.size(),
               total_sectors, idx->d_sectors);
    idx->sector_bytes = (size_t)total_sectors * 32;

    // ---- tables ----------------------------------------------------------------------------------------------------
    qt.n = text_len;
    qt.sampling_factor = sf;
    qt.iwt_levels = levels;
    qt.iwt_m = m;
    qt.n_nodes = sh.n_nodes;
    for (int c = 0; c < 256; c++) {
        qt.len[c] = (uint8_t)sh.bit_lengths[c];
        qt.code[c] = (uint16_t)sh.codes[c];
        for (int d = 0; d < sh.bit_lengths[c]; d++) {
            int found = -1;
            for (int v = 0; v < sh.n_nodes; v++)
                if (sh.node_depth[v] == d && sh.node_prefix[v] == ((uint16_t)sh.codes[c] & ((1 << d) - 1))) { found = v; break; }
            if (found < 0) return fail(GCZ_E_FORMAT, "symbol path leaves the tree");
            qt.node_of[c][d] = (uint8_t)found;
        }
    }
    for (int k = 0; k < sh.n_nodes; k++) { qt.child[k][0] = (int16_t)child0[k]; qt.child[k][1] = (int16_t)child1[k]; }
    for (int c = 0; c < 256; c++) {
        const int len = sh.bit_lengths[c];
        if (len == 0) continue;
        const int parent = qt.node_of[c][len - 1];
        qt.child[parent][((uint16_t)sh.codes[c] >> (len - 1)) & 1] = (int16_t)~c;
    }
    GCZ_CUDA(cudaMalloc(&idx->d_tables, sizeof(QueryTables)));
    GCZ_CUDA(cudaMemcpyAsync(idx->d_tables, &qt, sizeof(qt), cudaMemcpyHostToDevice, st));

    // C[]: occ(i, n - 1) for every byte value  algo/ssa/GSSA.java:215-226
    GCZ_LAUNCH(ctx, symbol_occ_kernel, 1, 256, 0, st, idx->d_tables, idx->d_sectors, d_scalar);
    long long occ[256];
    GCZ_CUDA(cudaMemcpyAsync(occ, d_scalar, sizeof(occ), cudaMemcpyDeviceToHost, st));
    GCZ_CUDA(cudaStreamSynchronize(st));
    int64_t run = text_len;
    for (int i = 255; i >= 0; i--) {
        if (occ[i] >= 0) run -= occ[i] + 1;
        qt.c[i] = run;
    }
    if (run != 0) return fail(GCZ_E_FORMAT, "symbol counts do not add up to the text length");
    std::memcpy(idx->c, qt.c, sizeof(qt.c));
    GCZ_CUDA(cudaMemcpyAsync(idx->d_tables, &qt, sizeof(qt), cudaMemcpyHostToDevice, st));
    GCZ_CUDA(cudaStreamSynchronize(st));

    // interval table of the K-symbol strings over A, C, G, T: the largest K <= 12 whose table is no larger than the block's rank
    // sectors (it at most doubles the device footprint of a block: 134 MB for a chr1-sized one) and whose strings still occur
    // 16 times on average.  Measured on six 200 Mbp blocks, 2 M patterns of 15..100 symbols: no table 6.41 ms, K = 8 / 9 / 10 /
    // 11 / 12: 4.55 / 4.31 / 3.91 / 3.44 / 3.02 ms.
    for (int c = 0; c < 256; c++) qt.code2[c] = 0xFF;
    qt.code2['A'] = 0; qt.code2['C'] = 1; qt.code2['G'] = 2; qt.code2['T'] = 3;
    {
        int K = 0;
        while (K < 12 && ((int64_t)16 << (2 * (K + 1))) <= text_len && ((size_t)8 << (2 * (K + 1))) <= idx->sector_bytes) K++;
        if (const char* ke = std::getenv("GCZ_KMER_K")) K = std::max(0, std::min(13, std::atoi(ke)));      // tuning aid
        const bool dna = qt.len['A'] && qt.len['C'] && qt.len['G'] && qt.len['T'];
        if (dna && K >= 4 && !std::getenv("GCZ_NO_KMER_TABLE")) {
            const size_t entries = (size_t)1 << (2 * K);
            GCZ_CUDA(cudaMalloc(&idx->d_kmer, entries * sizeof(uint2)));
            GCZ_LAUNCH(ctx, kmer_table_kernel, (unsigned)((entries + 255) / 256), 256, 0, st, idx->d_tables, idx->d_sectors, K, idx->d_kmer);
            qt.kmer = idx->d_kmer;
            qt.kmer_k = K;
        }
        GCZ_CUDA(cudaMemcpyAsync(idx->d_tables, &qt, sizeof(qt), cudaMemcpyHostToDevice, st));
        GCZ_CUDA(cudaStreamSynchronize(st));
    }

    // e[]: text positions of the separators = locate(i) for i < C[1], sorted  algo/ssa/GSSA.java:232-238
    const int64_t ns = qt.c[1];
    idx->e.resize((size_t)ns);
    if (ns > 0) {
        int64_t* d_rows = ctx->arena.get<int64_t>((size_t)ns * 2);
        if (!d_rows) return fail(GCZ_E_NOMEM, "open scratch");
        std::vector<int64_t> rows((size_t)ns);
        for (int64_t i = 0; i < ns; i++) rows[(size_t)i] = i;
        GCZ_CUDA(cudaMemcpyAsync(d_rows, rows.data(), (size_t)ns * 8, cudaMemcpyHostToDevice, st));
        GCZ_LAUNCH(ctx, locate_kernel, launch_grid(ctx, ns, 256), 256, 0, st, idx->d_tables, idx->d_sectors, d_rows, ns, d_rows + ns);
        GCZ_CUDA(cudaMemcpyAsync(idx->e.data(), d_rows + ns, (size_t)ns * 8, cudaMemcpyDeviceToHost, st));
        GCZ_CUDA(cudaStreamSynchronize(st));
        std::sort(idx->e.begin(), idx->e.end());
    }
    *out = idx.release();
    return GCZ_OK;
}

void close_block(gcz_index* idx) {
    if (!idx) return;
    cudaSetDevice(idx->ctx->device);
    if (idx->d_sectors) cudaFree(idx->d_sectors);
    if (idx->d_tables) cudaFree(idx->d_tables);
    if (idx->d_kmer) cudaFree(idx->d_kmer);
    delete idx;
}

// Stages a host-or-device input array on the device (arena) and a host-or-device output array.
template <class T>
static int stage_in(cudaStream_t st, Arena& arena, const T* src, size_t count, const T** dev) {
    if (is_device_ptr(src)) { *dev = src; return GCZ_OK; }
    T* d = arena.get<T>(count + 2);
    if (!d) return fail(GCZ_E_NOMEM, "query staging");
    GCZ_CUDA(cudaMemcpyAsync(d, src, count * sizeof(T), cudaMemcpyHostToDevice, st));
    *dev = d;
    return GCZ_OK;
}

// ---- batched queries: host side ------------------------------------------------------------------------------------
namespace {

// A pattern batch on the device (uploaded once, whatever the number of blocks it is searched in).
struct DeviceBatch {
    const uint8_t* pats = nullptr;
    const int64_t* off = nullptr;
    int64_t n = 0, bytes = 0;
};

int stage_batch(cudaStream_t st, Arena& arena, const uint8_t* pats, const int64_t* pat_off, int64_t n_pats, DeviceBatch* out) {
    int64_t total_bytes = 0;
    if (is_device_ptr(pat_off)) GCZ_CUDA(cudaMemcpy(&total_bytes, pat_off + n_pats, 8, cudaMemcpyDeviceToHost));
    else total_bytes = pat_off[n_pats];
    if (total_bytes < 0) return fail(GCZ_E_ARG, "pattern offsets");
    out->n = n_pats; out->bytes = total_bytes;
    GCZ_TRY(stage_in(st, arena, pats, (size_t)total_bytes, &out->pats));
    GCZ_TRY(stage_in(st, arena, pat_off, (size_t)n_pats + 1, &out->off));
    return GCZ_OK;
}

size_t batch_bytes(const uint8_t* pats, const int64_t* pat_off, int64_t n_pats) {
    // staging space when the batch lives on the host (the byte count of a device-resident batch needs no staging)
    if (is_device_ptr(pats) && is_device_ptr(pat_off)) return 0;
    const int64_t bytes = is_device_ptr(pat_off) ? 0 : pat_off[n_pats];
    return (size_t)std::max<int64_t>(bytes, 0) + (size_t)(n_pats + 3) * 8 + 4096;
}

int count_grid(DeviceCtx* ctx, int64_t n_pats) {
    return (int)std::max<int64_t>(1, std::min<int64_t>((n_pats + 255) / 256, (int64_t)ctx->sm_count * 8));
}

int same_device(gcz_index* const* blocks, int32_t n_blocks, DeviceCtx** ctx) {
    if (!blocks || n_blocks <= 0) return fail(GCZ_E_ARG, "no blocks");
    for (int32_t b = 0; b < n_blocks; b++) {
        if (!blocks[b]) return fail(GCZ_E_ARG, "null block %d", b);
        if (blocks[b]->ctx != blocks[0]->ctx) return fail(GCZ_E_ARG, "the blocks of one call must be open on one device");
    }
    *ctx = blocks[0]->ctx;
    return GCZ_OK;
}

std::atomic<int64_t> g_find_chunk{(int64_t)1 << 26};   // occurrences located and sorted per launch (gcz_dbg_set_find_chunk)
thread_local gcz_query_stats t_query_stats = {};
thread_local size_t t_find_want = 0;            // arena bytes a find_block that ran out of workspace asks for

// exclusive prefix sums of n int64 values (out has n + 1 entries): chunk sums, one CTA over the chunk sums, chunk scans
constexpr int kScanThreads = 256, kScanItems = 8, kScanChunk = kScanThreads * kScanItems;

__global__ void __launch_bounds__(kScanThreads)
occ_chunk_sums_kernel(const int64_t* __restrict__ sp, const int64_t* __restrict__ ep, int64_t n, int64_t* __restrict__ chunk_sums) {
    __shared__ long long s_w[kScanThreads / 32];
    const int64_t base = (int64_t)blockIdx.x * kScanChunk;
    long long v = 0;
    for (int i = threadIdx.x; i < kScanChunk; i += kScanThreads) {
        const int64_t q = base + i;
        if (q < n) v += max((long long)0, (long long)(ep[q] - sp[q] + 1));
    }
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane_id() == 0) s_w[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t = 0;
        for (int w = 0; w < kScanThreads / 32; w++) t += s_w[w];
        chunk_sums[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(1024)
scan_chunk_sums_kernel(int64_t* __restrict__ chunk_sums, int64_t chunks, int64_t* __restrict__ total) {
    __shared__ long long s_w[32];
    __shared__ long long s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int64_t c0 = 0; c0 < chunks; c0 += 1024) {
        const int64_t c = c0 + threadIdx.x;
        const long long v = c < chunks ? chunk_sums[c] : 0;
        long long incl = v;
        for (int o = 1; o < 32; o <<= 1) { const long long t = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane_id() >= o) incl += t; }
        if (lane_id() == 31) s_w[threadIdx.x >> 5] = incl;
        __syncthreads();
        long long before = s_carry;
        for (int w = 0; w < (int)(threadIdx.x >> 5); w++) before += s_w[w];
        if (c < chunks) chunk_sums[c] = before + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = before + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = s_carry;
}

__global__ void __launch_bounds__(kScanThreads)
occ_scan_kernel(const int64_t* __restrict__ sp, const int64_t* __restrict__ ep, int64_t n, const int64_t* __restrict__ chunk_excl,
                const int64_t* __restrict__ total, int64_t* __restrict__ occ_excl /* n + 1 */) {
    __shared__ long long s_w[kScanThreads / 32];
    const int64_t base = (int64_t)blockIdx.x * kScanChunk + (int64_t)threadIdx.x * kScanItems;
    long long v[kScanItems], sum = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        const int64_t q = base + i;
        v[i] = q < n ? max((long long)0, (long long)(ep[q] - sp[q] + 1)) : 0;
        sum += v[i];
    }
    long long incl = sum;
    for (int o = 1; o < 32; o <<= 1) { const long long t = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane_id() >= o) incl += t; }
    if (lane_id() == 31) s_w[threadIdx.x >> 5] = incl;
    __syncthreads();
    long long before = chunk_excl[blockIdx.x] + incl - sum;
    for (int w = 0; w < (int)(threadIdx.x >> 5); w++) before += s_w[w];
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        const int64_t q = base + i;
        if (q < n) occ_excl[q] = before;
        before += v[i];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) occ_excl[n] = *total;
}

// GSSA.find :170-184 for occurrences sorted by (pattern, position): the string of an occurrence at text position x is
// the first one whose end e[i] lies above x, its position there x - start.  flags[0] is raised when a position equals a
// string end (the reference's binarySearch then FINDS the key and takes its other branch) and flags[1] when a position
// lies behind the last string end (the reference drops it): such chunks are redone by the literal loop on the host.
__global__ void __launch_bounds__(256)
split_by_string_kernel(const uint64_t* __restrict__ keys, int64_t n_occ, const int64_t* __restrict__ e, int32_t ns, int64_t first_pat,
                       int64_t* __restrict__ out_pattern, int32_t* __restrict__ out_string, int64_t* __restrict__ out_pos,
                       unsigned* __restrict__ flags) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n_occ; o += stride) {
        const long long x = (long long)(int32_t)(uint32_t)(keys[o] & 0xFFFFFFFFull);
        out_pattern[o] = first_pat + (int64_t)(keys[o] >> 32);
        int lo = 0, hi = ns;                               // first i with e[i] >= x
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (e[mid] < x) lo = mid + 1; else hi = mid;
        }
        if (lo < ns && e[lo] == x) atomicOr(&flags[0], 1u);
        if (lo >= ns) { atomicOr(&flags[1], 1u); out_string[o] = -1; out_pos[o] = x; continue; }
        out_string[o] = lo;
        out_pos[o] = x - (lo > 0 ? e[lo - 1] + 1 : 0);
    }
}

// java.util.Arrays.binarySearch(long[] a, int from, int to, long key)
int64_t java_binary_search(const int64_t* a, int64_t from, int64_t to, int64_t key) {
    int64_t low = from, high = to - 1;
    while (low <= high) {
        const int64_t mid = (int64_t)(((uint64_t)low + (uint64_t)high) >> 1);
        if (a[mid] < key) low = mid + 1;
        else if (a[mid] > key) high = mid - 1;
        else return mid;
    }
    return -(low + 1);
}

// Hits of one block in the order GSSA.find returns them: pattern after pattern, string after string, ascending.
struct BlockHits {
    std::vector<int64_t> pattern;
    std::vector<int32_t> string;
    std::vector<int64_t> pos;
};

// The whole of GSSA.find for a device-resident batch against one block.  Intervals, occurrence offsets, locate, the sort
// by (pattern, position) and the split by string ends all run on the device; what comes back is 20 bytes per hit and
// nothing per pattern.  A chunk whose split met one of the reference's corner cases is redone literally on the host.
int find_block(gcz_index* idx, cudaStream_t st, const DeviceBatch& batch, BlockHits* out) {
    DeviceCtx* ctx = idx->ctx;
    Arena& arena = ctx->arena;
    const size_t mark0 = arena.mark();
    const int64_t n_pats = batch.n;
    const int32_t ns = (int32_t)idx->e.size();
    out->pattern.clear(); out->string.clear(); out->pos.clear();
    if (n_pats == 0) return GCZ_OK;

    const int64_t chunks = (n_pats + kScanChunk - 1) / kScanChunk;
    int64_t* d_sp = arena.get<int64_t>((size_t)n_pats);
    int64_t* d_ep = arena.get<int64_t>((size_t)n_pats);
    int64_t* d_excl = arena.get<int64_t>((size_t)n_pats + 1);
    int64_t* d_chunk = arena.get<int64_t>((size_t)chunks + 1);
    int64_t* d_e = arena.get<int64_t>((size_t)ns + 1);
    unsigned long long* d_next = arena.get<unsigned long long>(4);      // [0] pattern counter, [1] total, [2] flags
    if (!d_sp || !d_ep || !d_excl || !d_chunk || !d_e || !d_next) return fail(GCZ_E_NOMEM, "find workspace");
    GCZ_CUDA(cudaMemsetAsync(d_next, 0, 32, st));
    if (ns > 0) GCZ_CUDA(cudaMemcpyAsync(d_e, idx->e.data(), (size_t)ns * 8, cudaMemcpyHostToDevice, st));
    GCZ_LAUNCH(ctx, count_kernel<0>, count_grid(ctx, n_pats), 256, 0, st, idx->d_tables, idx->d_sectors, batch.pats, batch.off, n_pats,
               d_sp, d_ep, d_next, (unsigned long long*)nullptr, 1);
    int64_t* d_total = reinterpret_cast<int64_t*>(d_next + 1);
    GCZ_LAUNCH(ctx, occ_chunk_sums_kernel, (unsigned)chunks, kScanThreads, 0, st, d_sp, d_ep, n_pats, d_chunk);
    GCZ_LAUNCH(ctx, scan_chunk_sums_kernel, 1, 1024, 0, st, d_chunk, chunks, d_total);
    GCZ_LAUNCH(ctx, occ_scan_kernel, (unsigned)chunks, kScanThreads, 0, st, d_sp, d_ep, n_pats, d_chunk, d_total, d_excl);
    int64_t total_occ = 0;
    GCZ_CUDA(cudaMemcpyAsync(&total_occ, d_total, 8, cudaMemcpyDeviceToHost, st));
    GCZ_CUDA(cudaStreamSynchronize(st));
    if (total_occ == 0) { arena.release(mark0); return GCZ_OK; }

    // pattern ranges of at most kChunk occurrences (a single pattern may exceed it); one range in the common case
    const int64_t kChunk = g_find_chunk.load();
    std::vector<int64_t> h_excl;
    std::vector<int64_t> cuts = { 0, n_pats };
    if (total_occ > kChunk) {
        h_excl.resize((size_t)n_pats + 1);
        GCZ_CUDA(cudaMemcpyAsync(h_excl.data(), d_excl, ((size_t)n_pats + 1) * 8, cudaMemcpyDeviceToHost, st));
        GCZ_CUDA(cudaStreamSynchronize(st));
        cuts.assign(1, 0);
        int64_t p0 = 0;
        while (p0 < n_pats) {
            int64_t p1 = p0 + 1;
            while (p1 < n_pats && h_excl[(size_t)p1 + 1] - h_excl[(size_t)p0] <= kChunk) p1++;
            cuts.push_back(p1);
            p0 = p1;
        }
    }
    out->pattern.resize((size_t)total_occ);
    out->string.resize((size_t)total_occ);
    out->pos.resize((size_t)total_occ);

    const size_t mark1 = arena.mark();
    std::vector<uint64_t> h_keys;
    std::vector<std::pair<int64_t, int64_t>> redo;        // occurrence ranges whose split has to be redone literally
    int64_t occ_base = 0;
    for (size_t c = 0; c + 1 < cuts.size(); c++) {
        const int64_t p0 = cuts[c], p1 = cuts[c + 1], np = p1 - p0;
        int64_t n_occ = total_occ;
        if (cuts.size() > 2) n_occ = h_excl[(size_t)p1] - h_excl[(size_t)p0];
        if (n_occ == 0) continue;
        arena.release(mark1);
        t_find_want = arena.mark() + (size_t)n_occ * 36 + radix_sort_temp_bytes(n_occ) + ((size_t)8 << 20);
        uint64_t* d_k0 = arena.get<uint64_t>((size_t)n_occ);
        uint64_t* d_k1 = arena.get<uint64_t>((size_t)n_occ);
        void* d_tmp = arena.raw(radix_sort_temp_bytes(n_occ));
        int64_t* d_pat = arena.get<int64_t>((size_t)n_occ);
        int32_t* d_str = arena.get<int32_t>((size_t)n_occ);
        int64_t* d_pos = arena.get<int64_t>((size_t)n_occ);
        if (!d_k0 || !d_k1 || !d_tmp || !d_pat || !d_str || !d_pos) return fail(GCZ_E_NOMEM, "find workspace for %lld occurrences", (long long)n_occ);
        GCZ_LAUNCH(ctx, locate_occurrences_kernel, launch_grid(ctx, n_occ, 256), 256, 0, st, idx->d_tables, idx->d_sectors,
                   d_sp, d_excl, p0, np, occ_base, n_occ, d_k0);
        RadixBuffers rb;
        rb.keys[0] = d_k0; rb.keys[1] = d_k1;
        int pat_bits = 1;
        while (((int64_t)1 << pat_bits) < np) pat_bits++;
        GCZ_TRY(radix_sort_pairs(ctx, st, rb, n_occ, 0, 32 + pat_bits, d_tmp, nullptr));
        unsigned* d_flags = reinterpret_cast<unsigned*>(d_next + 2);
        GCZ_CUDA(cudaMemsetAsync(d_flags, 0, 8, st));
        GCZ_LAUNCH(ctx, split_by_string_kernel, launch_grid(ctx, n_occ, 256), 256, 0, st, rb.keys[rb.cur], n_occ, d_e, ns, p0, d_pat, d_str, d_pos, d_flags);
        unsigned h_flags[2] = { 0, 0 };
        GCZ_CUDA(cudaMemcpyAsync(h_flags, d_flags, 8, cudaMemcpyDeviceToHost, st));
        GCZ_CUDA(cudaMemcpyAsync(out->pattern.data() + occ_base, d_pat, (size_t)n_occ * 8, cudaMemcpyDeviceToHost, st));
        GCZ_CUDA(cudaMemcpyAsync(out->string.data() + occ_base, d_str, (size_t)n_occ * 4, cudaMemcpyDeviceToHost, st));
        GCZ_CUDA(cudaMemcpyAsync(out->pos.data() + occ_base, d_pos, (size_t)n_occ * 8, cudaMemcpyDeviceToHost, st));
        GCZ_CUDA(cudaStreamSynchronize(st));
        if (h_flags[0] || h_flags[1]) {
            // keep the sorted text positions of this chunk in out->pos (sign-extended, as the reference's long[] holds them)
            h_keys.resize((size_t)n_occ);
            GCZ_CUDA(cudaMemcpyAsync(h_keys.data(), rb.keys[rb.cur], (size_t)n_occ * 8, cudaMemcpyDeviceToHost, st));
            GCZ_CUDA(cudaStreamSynchronize(st));
            for (int64_t o = 0; o < n_occ; o++) out->pos[(size_t)(occ_base + o)] = (int64_t)(int32_t)(uint32_t)(h_keys[(size_t)o] & 0xFFFFFFFFull);
            redo.emplace_back(occ_base, occ_base + n_occ);
        }
        occ_base += n_occ;
    }
    arena.release(mark0);
    if (redo.empty()) return GCZ_OK;

    // GSSA.find :170-184 as written, for the chunks that asked for it; hits may be fewer than occurrences afterwards
    BlockHits fixed;
    fixed.pattern.reserve(out->pos.size()); fixed.string.reserve(out->pos.size()); fixed.pos.reserve(out->pos.size());
    size_t r = 0;
    const int64_t total = (int64_t)out->pos.size();
    int64_t o = 0;
    while (o < total) {
        while (r < redo.size() && o >= redo[r].second) r++;
        if (!(r < redo.size() && o >= redo[r].first)) {                  // a chunk the device split completely
            const int64_t end = r < redo.size() ? redo[r].first : total;
            fixed.pattern.insert(fixed.pattern.end(), out->pattern.begin() + o, out->pattern.begin() + end);
            fixed.string.insert(fixed.string.end(), out->string.begin() + o, out->string.begin() + end);
            fixed.pos.insert(fixed.pos.end(), out->pos.begin() + o, out->pos.begin() + end);
            o = end;
            continue;
        }
        // the occurrences of one pattern: out->pos holds their sorted text positions
        const int64_t p = out->pattern[(size_t)o];
        int64_t k = 1;
        while (o + k < redo[r].second && out->pattern[(size_t)(o + k)] == p) k++;
        const int64_t* sa = out->pos.data() + o;
        int64_t idx1 = 0;
        for (int64_t i = 0; i < ns; i++) {
            const int64_t idx2 = -java_binary_search(sa, idx1, k, idx->e[(size_t)i]) - 1;
            if (idx2 > idx1) {
                const int64_t start = i > 0 ? idx->e[(size_t)i - 1] + 1 : 0;
                for (int64_t j = idx1; j < idx2; j++) { fixed.pattern.push_back(p); fixed.string.push_back((int32_t)i); fixed.pos.push_back(sa[j] - start); }
                idx1 = idx2;
            }
        }
        o += k;
    }
    *out = std::move(fixed);
    return GCZ_OK;
}

}  // namespace

int count_batch(gcz_index* idx, const uint8_t* pats, const int64_t* pat_off, int64_t n_pats, int64_t* sp, int64_t* ep) {
    if (!idx || !pats || !pat_off || !sp || !ep || n_pats < 0) return fail(GCZ_E_ARG, "count_batch arguments");
    if (n_pats == 0) return GCZ_OK;
    DeviceCtx* ctx = idx->ctx;
    std::lock_guard<std::mutex> lock(ctx->mu);
    GCZ_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = stream_of(ctx);
    ctx->arena.reset();
    const size_t need = batch_bytes(pats, pat_off, n_pats) + (size_t)n_pats * 16 + (1 << 20);
    if (ctx->arena.capacity < need) GCZ_TRY(ctx->arena.reserve(need));
    DeviceBatch batch;
    GCZ_TRY(stage_batch(st, ctx->arena, pats, pat_off, n_pats, &batch));
    const bool out_dev = is_device_ptr(sp);
    if (out_dev != is_device_ptr(ep)) return fail(GCZ_E_ARG, "sp and ep must live on the same side");
    int64_t* d_sp = out_dev ? sp : ctx->arena.get<int64_t>((size_t)n_pats);
    int64_t* d_ep = out_dev ? ep : ctx->arena.get<int64_t>((size_t)n_pats);
    unsigned long long* d_next = ctx->arena.get<unsigned long long>(1);
    if (!d_sp || !d_ep || !d_next) return fail(GCZ_E_NOMEM, "query staging");
    GCZ_CUDA(cudaMemsetAsync(d_next, 0, 8, st));
    GCZ_LAUNCH(ctx, count_kernel<0>, count_grid(ctx, n_pats), 256, 0, st, idx->d_tables, idx->d_sectors, batch.pats, batch.off, n_pats,
               d_sp, d_ep, d_next, (unsigned long long*)nullptr, 1);
    if (!out_dev) {
        GCZ_CUDA(cudaMemcpyAsync(sp, d_sp, (size_t)n_pats * 8, cudaMemcpyDeviceToHost, st));
        GCZ_CUDA(cudaMemcpyAsync(ep, d_ep, (size_t)n_pats * 8, cudaMemcpyDeviceToHost, st));
    }
    GCZ_CUDA(cudaStreamSynchronize(st));
    return GCZ_OK;
}

// The loop of GecoMatch over the blocks of a file (tools/GecoMatch.java:114-131) for a whole batch: the batch is uploaded
// once, every block is searched on the device, the occurrences of a pattern are summed there; 8 bytes per pattern come back.
int count_multi(gcz_index* const* blocks, int32_t n_blocks, const uint8_t* pats, const int64_t* pat_off, int64_t n_pats, int64_t* totals) {
    if (!pats || !pat_off || !totals || n_pats < 0) return fail(GCZ_E_ARG, "count_multi arguments");
    DeviceCtx* ctx = nullptr;
    GCZ_TRY(same_device(blocks, n_blocks, &ctx));
    if (n_pats == 0) return GCZ_OK;
    std::lock_guard<std::mutex> lock(ctx->mu);
    GCZ_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = stream_of(ctx);
    ctx->arena.reset();
    const size_t need = batch_bytes(pats, pat_off, n_pats) + (size_t)n_pats * 8 + (size_t)n_blocks * 8 + (1 << 20);
    if (ctx->arena.capacity < need) GCZ_TRY(ctx->arena.reserve(need));
    DeviceBatch batch;
    GCZ_TRY(stage_batch(st, ctx->arena, pats, pat_off, n_pats, &batch));
    const bool out_dev = is_device_ptr(totals);
    int64_t* d_tot = out_dev ? totals : ctx->arena.get<int64_t>((size_t)n_pats);
    unsigned long long* d_next = ctx->arena.get<unsigned long long>((size_t)n_blocks);
    if (!d_tot || !d_next) return fail(GCZ_E_NOMEM, "query staging");
    GCZ_CUDA(cudaMemsetAsync(d_tot, 0, (size_t)n_pats * 8, st));
    GCZ_CUDA(cudaMemsetAsync(d_next, 0, (size_t)n_blocks * 8, st));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    GCZ_CUDA(cudaEventCreate(&e0)); GCZ_CUDA(cudaEventCreate(&e1));
    GCZ_CUDA(cudaEventRecord(e0, st));
    for (int32_t b = 0; b < n_blocks; b++) {
        GCZ_LAUNCH(ctx, count_kernel<1>, count_grid(ctx, n_pats), 256, 0, st, blocks[b]->d_tables, blocks[b]->d_sectors, batch.pats, batch.off,
                   n_pats, d_tot, (int64_t*)nullptr, d_next + b, (unsigned long long*)nullptr, 1);
    }
    GCZ_CUDA(cudaEventRecord(e1, st));
    if (!out_dev) GCZ_CUDA(cudaMemcpyAsync(totals, d_tot, (size_t)n_pats * 8, cudaMemcpyDeviceToHost, st));
    GCZ_CUDA(cudaStreamSynchronize(st));
    t_query_stats = gcz_query_stats{};
    t_query_stats.patterns = n_pats;
    t_query_stats.blocks = n_blocks;
    cudaEventElapsedTime(&t_query_stats.kernel_ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return GCZ_OK;
}

// The same searches with counters: rank sectors loaded, backward-search steps, and the RankedWTNode.count calls of the
// reference's loop for these patterns.  Not a timed path (the counters cost registers); kernel_ms is of this launch.
int count_stats(gcz_index* const* blocks, int32_t n_blocks, const uint8_t* pats, const int64_t* pat_off, int64_t n_pats, gcz_query_stats* out) {
    if (!pats || !pat_off || !out || n_pats < 0) return fail(GCZ_E_ARG, "count_stats arguments");
    DeviceCtx* ctx = nullptr;
    GCZ_TRY(same_device(blocks, n_blocks, &ctx));
    std::lock_guard<std::mutex> lock(ctx->mu);
    GCZ_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = stream_of(ctx);
    ctx->arena.reset();
    const size_t need = batch_bytes(pats, pat_off, n_pats) + (size_t)n_pats * 16 + (size_t)n_blocks * 8 + (1 << 20);
    if (ctx->arena.capacity < need) GCZ_TRY(ctx->arena.reserve(need));
    DeviceBatch batch;
    GCZ_TRY(stage_batch(st, ctx->arena, pats, pat_off, n_pats, &batch));
    int64_t* d_sp = ctx->arena.get<int64_t>((size_t)n_pats + 1);
    int64_t* d_ep = ctx->arena.get<int64_t>((size_t)n_pats + 1);
    unsigned long long* d_next = ctx->arena.get<unsigned long long>((size_t)2 * n_blocks + 4);
    if (!d_sp || !d_ep || !d_next) return fail(GCZ_E_NOMEM, "query staging");
    GCZ_CUDA(cudaMemsetAsync(d_next, 0, ((size_t)2 * n_blocks + 4) * 8, st));
    unsigned long long* d_stats = d_next + 2 * n_blocks;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    GCZ_CUDA(cudaEventCreate(&e0)); GCZ_CUDA(cudaEventCreate(&e1));
    // first the search the way the reference makes it, symbol by symbol (its rank calls and steps), then the way this library
    // makes it, with the interval table for the last symbols of a pattern (the sectors it loads, and its time)
    for (int32_t b = 0; b < n_blocks && n_pats > 0; b++) {
        GCZ_LAUNCH(ctx, count_kernel<2>, count_grid(ctx, n_pats), 256, 0, st, blocks[b]->d_tables, blocks[b]->d_sectors, batch.pats, batch.off,
                   n_pats, d_sp, d_ep, d_next + b, d_stats, 0);
    }
    GCZ_CUDA(cudaEventRecord(e0, st));
    for (int32_t b = 0; b < n_blocks && n_pats > 0; b++) {
        GCZ_LAUNCH(ctx, count_kernel<2>, count_grid(ctx, n_pats), 256, 0, st, blocks[b]->d_tables, blocks[b]->d_sectors, batch.pats, batch.off,
                   n_pats, d_sp, d_ep, d_next + n_blocks + b, d_stats, 1);
    }
    GCZ_CUDA(cudaEventRecord(e1, st));
    unsigned long long h[3] = { 0, 0, 0 };
    GCZ_CUDA(cudaMemcpyAsync(h, d_stats, sizeof(h), cudaMemcpyDeviceToHost, st));
    GCZ_CUDA(cudaStreamSynchronize(st));
    *out = gcz_query_stats{};
    out->patterns = n_pats; out->blocks = n_blocks;
    out->rank_sectors = (int64_t)h[0]; out->reference_rank_calls = (int64_t)h[1]; out->steps = (int64_t)h[2];
    for (int32_t b = 0; b < n_blocks; b++) out->index_bytes += (int64_t)blocks[b]->sector_bytes;
    cudaEventElapsedTime(&out->kernel_ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return GCZ_OK;
}

void set_find_chunk(int64_t occurrences) { g_find_chunk.store(occurrences > 0 ? occurrences : (int64_t)1 << 26); }

int last_query_stats(gcz_query_stats* out) {
    if (!out) return fail(GCZ_E_ARG, "null argument");
    *out = t_query_stats;
    return GCZ_OK;
}

int locate_rows(gcz_index* idx, const int64_t* rows, int64_t n_rows, int64_t* positions) {
    if (!idx || !rows || !positions || n_rows < 0) return fail(GCZ_E_ARG, "locate_rows arguments");
    if (n_rows == 0) return GCZ_OK;
    DeviceCtx* ctx = idx->ctx;
    std::lock_guard<std::mutex> lock(ctx->mu);
    GCZ_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = stream_of(ctx);
    ctx->arena.reset();
    const size_t need = (size_t)n_rows * 16 + (1 << 20);
    if (ctx->arena.capacity < need) GCZ_TRY(ctx->arena.reserve(need));
    const int64_t* d_rows = nullptr;
    GCZ_TRY(stage_in(st, ctx->arena, rows, (size_t)n_rows, &d_rows));
    const bool out_dev = is_device_ptr(positions);
    int64_t* d_out = out_dev ? positions : ctx->arena.get<int64_t>((size_t)n_rows);
    if (!d_out) return fail(GCZ_E_NOMEM, "query staging");
    GCZ_LAUNCH(ctx, locate_kernel, launch_grid(ctx, n_rows, 256), 256, 0, st, idx->d_tables, idx->d_sectors, d_rows, n_rows, d_out);
    if (!out_dev) GCZ_CUDA(cudaMemcpyAsync(positions, d_out, (size_t)n_rows * 8, cudaMemcpyDeviceToHost, st));
    GCZ_CUDA(cudaStreamSynchronize(st));
    return GCZ_OK;
}


// GSSA.find for a batch against every block of a file (the loops of GecoMatch / SimpleGFFGenerator over the blocks): the
// batch is uploaded once; per block the hits come back as sparse records sorted by (pattern, string, position).
int find_multi(gcz_index* const* blocks, int32_t n_blocks, const uint8_t* pats, const int64_t* pat_off, int64_t n_pats, gcz_hits* out) {
    if (!pats || !pat_off || !out || n_pats < 0) return fail(GCZ_E_ARG, "find_multi arguments");
    std::memset(out, 0, sizeof(*out));
    DeviceCtx* ctx = nullptr;
    GCZ_TRY(same_device(blocks, n_blocks, &ctx));
    std::lock_guard<std::mutex> lock(ctx->mu);
    GCZ_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = stream_of(ctx);
    ctx->arena.reset();
    // the batch, the intervals and offsets of one block, and room for the occurrence keys of a typical batch up front
    const size_t need = batch_bytes(pats, pat_off, n_pats) + (size_t)n_pats * 40 + ((size_t)96 << 20);
    if (ctx->arena.capacity < need) GCZ_TRY(ctx->arena.reserve(need));
    DeviceBatch batch;
    if (n_pats > 0) GCZ_TRY(stage_batch(st, ctx->arena, pats, pat_off, n_pats, &batch));
    const size_t mark = ctx->arena.mark();

    std::vector<BlockHits> per((size_t)n_blocks);
    int64_t total = 0;
    for (int32_t b = 0; b < n_blocks; b++) {
        ctx->arena.release(mark);
        t_find_want = 0;
        int rc = find_block(blocks[b], st, batch, &per[(size_t)b]);
        if (rc == GCZ_E_NOMEM) {
            // the occurrence workspace did not fit: grow the arena (the staged batch is lost with it) and redo this block
            GCZ_CUDA(cudaStreamSynchronize(st));
            const size_t more = std::max(t_find_want, ctx->arena.capacity) + ((size_t)64 << 20);
            ctx->arena.reset();
            GCZ_TRY(ctx->arena.reserve(more));
            if (n_pats > 0) GCZ_TRY(stage_batch(st, ctx->arena, pats, pat_off, n_pats, &batch));
            if (ctx->arena.mark() != mark) return fail(GCZ_E_INTERNAL, "arena layout changed");
            rc = find_block(blocks[b], st, batch, &per[(size_t)b]);
        }
        GCZ_TRY(rc);
        total += (int64_t)per[(size_t)b].pos.size();
    }
    out->n_hits = total;
    out->block_off = static_cast<int64_t*>(std::malloc(sizeof(int64_t) * ((size_t)n_blocks + 1)));
    out->pattern = static_cast<int64_t*>(std::malloc(sizeof(int64_t) * (size_t)std::max<int64_t>(total, 1)));
    out->string = static_cast<int32_t*>(std::malloc(sizeof(int32_t) * (size_t)std::max<int64_t>(total, 1)));
    out->position = static_cast<int64_t*>(std::malloc(sizeof(int64_t) * (size_t)std::max<int64_t>(total, 1)));
    if (!out->block_off || !out->pattern || !out->string || !out->position) {
        std::free(out->block_off); std::free(out->pattern); std::free(out->string); std::free(out->position);
        std::memset(out, 0, sizeof(*out));
        return fail(GCZ_E_NOMEM, "find_multi results");
    }
    int64_t at = 0;
    for (int32_t b = 0; b < n_blocks; b++) {
        const BlockHits& h = per[(size_t)b];
        out->block_off[b] = at;
        if (!h.pos.empty()) {
            std::memcpy(out->pattern + at, h.pattern.data(), h.pattern.size() * 8);
            std::memcpy(out->string + at, h.string.data(), h.string.size() * 4);
            std::memcpy(out->position + at, h.pos.data(), h.pos.size() * 8);
        }
        at += (int64_t)h.pos.size();
    }
    out->block_off[n_blocks] = at;
    return GCZ_OK;
}

// GSSA.find  :160-185 for a batch against one block, in the dense form of the C ABI
int find_batch(gcz_index* idx, const uint8_t* pats, const int64_t* pat_off, int64_t n_pats,
               int64_t* per_string_counts, int64_t** positions, int64_t** pos_off) {
    if (!idx || !pats || !pat_off || !positions || !pos_off || n_pats < 0) return fail(GCZ_E_ARG, "find_batch arguments");
    *positions = nullptr; *pos_off = nullptr;
    gcz_hits hits;
    gcz_index* one[1] = { idx };
    GCZ_TRY(find_multi(one, 1, pats, pat_off, n_pats, &hits));
    const int64_t ns = (int64_t)idx->e.size();
    int64_t* h_off = static_cast<int64_t*>(std::malloc(sizeof(int64_t) * ((size_t)n_pats + 1)));
    if (!h_off) { std::free(hits.block_off); std::free(hits.pattern); std::free(hits.string); std::free(hits.position); return fail(GCZ_E_NOMEM, "find_batch results"); }
    if (per_string_counts) std::memset(per_string_counts, 0, sizeof(int64_t) * (size_t)(n_pats * ns));
    std::memset(h_off, 0, sizeof(int64_t) * ((size_t)n_pats + 1));
    for (int64_t j = 0; j < hits.n_hits; j++) {
        h_off[hits.pattern[j] + 1]++;
        if (per_string_counts) per_string_counts[hits.pattern[j] * ns + hits.string[j]]++;
    }
    for (int64_t p = 0; p < n_pats; p++) h_off[p + 1] += h_off[p];
    *positions = hits.position;                            // already pattern after pattern, string after string
    *pos_off = h_off;
    std::free(hits.block_off); std::free(hits.pattern); std::free(hits.string);
    return GCZ_OK;
}

// GSSA.extract(ByteBuffer buf, int nstr, long from)  algo/ssa/GSSA.java:90-126
int extract(gcz_index* idx, int32_t nstr, int64_t from, uint8_t* out, int64_t cap, int64_t* written) {
    if (!idx || !out || !written || cap < 0 || from < 0) return fail(GCZ_E_ARG, "extract arguments");
    const int64_t ns = (int64_t)idx->e.size();
    if (nstr < 0 || nstr >= ns) return fail(GCZ_E_RANGE, "String index %d is out of bound", nstr);
    if (nstr > 0) from += idx->e[(size_t)nstr - 1] + 1;
    const int64_t pos = std::min(idx->e[(size_t)nstr], from + cap) - 1;
    *written = 0;
    // the reference ends with buf.position(bpos + 1), bpos = (int)(pos - from): IllegalArgumentException below zero
    // (strings whose end the index mislocates)
    if (pos - from + 1 < 0) return fail(GCZ_E_RANGE, "newPosition < 0: the string ends before position %lld", (long long)from);
    if (pos < from) return GCZ_OK;
    const int64_t count = pos - from + 1;
    DeviceCtx* ctx = idx->ctx;
    std::lock_guard<std::mutex> lock(ctx->mu);
    GCZ_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = stream_of(ctx);
    ctx->arena.reset();
    const bool out_dev = is_device_ptr(out);
    const size_t need = (out_dev ? 0 : (size_t)count + 256) + (1 << 20);
    if (ctx->arena.capacity < need) GCZ_TRY(ctx->arena.reserve(need));
    uint8_t* d_out = out_dev ? out : ctx->arena.get<uint8_t>((size_t)count + 64);
    long long* d_flag = ctx->arena.get<long long>(2);
    if (!d_out || !d_flag) return fail(GCZ_E_NOMEM, "extract staging");
    const int sf = idx->sampling_factor;
    const int64_t low = std::max(from, (pos >> sf) << sf);
    const int64_t anchors = low > from ? ((low - from) >> sf) + 1 : 0;              // sampled positions in (from, low]
    const int64_t threads = 1 + anchors;
    GCZ_LAUNCH(ctx, extract_kernel, (unsigned)((threads + 127) / 128), 128, 0, st, idx->d_tables, idx->d_sectors,
               (long long)from, (long long)pos, d_out, d_flag);
    long long h_flag[2] = { 0, 0 };
    GCZ_CUDA(cudaMemcpyAsync(h_flag, d_flag, sizeof(h_flag), cudaMemcpyDeviceToHost, st));
    GCZ_CUDA(cudaStreamSynchronize(st));
    if (h_flag[0]) {
        GCZ_LAUNCH(ctx, extract_sequential_kernel, 1, 32, 0, st, idx->d_tables, idx->d_sectors, (long long)from, (long long)low,
                   h_flag[1], d_out);
    }
    if (!out_dev) GCZ_CUDA(cudaMemcpyAsync(out, d_out, (size_t)count, cudaMemcpyDeviceToHost, st));
    GCZ_CUDA(cudaStreamSynchronize(st));
    *written = count;
    return GCZ_OK;
}

}  // namespace gcz

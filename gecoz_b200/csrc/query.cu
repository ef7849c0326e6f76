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

// The two occ() of one backward-search step, occ(symbol, p1) and occ(symbol, p2) with p1 <= p2, walking the
// symbol's path once: on every node the two positions usually fall into the same rank sector (always, once the
// interval has shrunk to a few suffixes), and that sector is loaded once.
__device__ __forceinline__ void hswt_occ2(const QueryTables* __restrict__ t, const uint32_t* __restrict__ sectors,
                                          int symbol, long long& p1, long long& p2) {
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
        const uint32_t r2 = sector_rank_at(w, q2 - sec2 * kSectorBits);
        if (p1 >= 0) {
            const uint32_t q1 = (uint32_t)p1, sec1 = q1 / kSectorBits;
            uint32_t r1;
            if (sec1 == sec2) {
                r1 = sector_rank_at(w, q1 - sec1 * kSectorBits);
            } else {
                r1 = sector_bit_rank(sectors, s0, q1).rank;
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
// ORDERED (experimental, GCZ_COUNT_SORT=1): patterns are taken in the order of `order` (sorted by their last symbols, see
// pattern_suffix_keys_kernel), so that the lanes of a warp walk the same rows for the first steps of the backward search.
template <bool ORDERED>
__global__ void __launch_bounds__(256)
count_kernel(const QueryTables* __restrict__ tables, const uint32_t* __restrict__ sectors,
             const uint8_t* __restrict__ pats, const int64_t* __restrict__ pat_off, int64_t n_pats,
             int64_t* __restrict__ sp_out, int64_t* __restrict__ ep_out, unsigned long long* __restrict__ next_pattern,
             const uint32_t* __restrict__ order) {
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
                q = ORDERED ? (long long)order[mine] : mine;
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
                    hswt_occ2(&t, sectors, ch, o1, o2);
                    sp = t.c[ch] + o1 + 1;
                    ep = t.c[ch] + o2;
                }
                i--;
            } else {
                sp_out[q] = sp;
                ep_out[q] = ep;
                q = -1;
            }
        }
    }
}

// Sort key of a pattern for GCZ_COUNT_SORT=1: its last 12 symbols, the last one most significant, 5 bits each (the id of the
// byte among the symbols of the block, modulo 32) — the order in which the backward search consumes them.
struct SymbolIds { uint8_t id[256]; };

__global__ void pattern_suffix_keys_kernel(const uint8_t* __restrict__ pats, const int64_t* __restrict__ pat_off, int64_t n_pats,
                                           SymbolIds ids, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n_pats; q += stride) {
        const int64_t b = pat_off[q], e = pat_off[q + 1];
        uint64_t key = 0;
        for (int j = 0; j < 12; j++) {
            const int64_t at = e - 1 - j;
            key = (key << 5) | (at >= b ? (uint64_t)(ids.id[pats[at]] & 31u) : 0ull);
        }
        keys[q] = key;
        vals[q] = (uint32_t)q;
    }
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

// Experimental (GCZ_LOCATE_VARIANT=1, default off; written without hardware): the same work in two launches.
//  walk:   LF walks end after 0 .. 2^sf - 1 steps, so in locate_occurrences_kernel a warp waits for its longest one.  Here a
//          lane that has reached a marked row takes the next occurrence from a grid-wide counter (one atomic per warp and
//          refill) and the warp keeps stepping; the walk leaves (steps << 32 | sampled index) behind.
//  finish: the IndexWaveletTree descent (one node per level for every occurrence) runs with all lanes in step.
__global__ void __launch_bounds__(256)
locate_walk_refill_kernel(const QueryTables* __restrict__ tables, const uint32_t* __restrict__ sectors,
                          const int64_t* __restrict__ sp, const int64_t* __restrict__ occ_excl /* per pattern, exclusive */,
                          int64_t first_pat, int64_t n_chunk_pats, int64_t base_occ, int64_t n_occ,
                          uint64_t* __restrict__ keys, uint64_t* __restrict__ walked, unsigned long long* __restrict__ next) {
    __shared__ QueryTables t;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(tables);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&t);
        for (int i = threadIdx.x; i < (int)(sizeof(QueryTables) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const unsigned lane = threadIdx.x & 31u, lt = lanemask_lt();
    long long o = 0, idx = 0, steps = 0;
    bool have = false, exhausted = false;                    // exhausted: the counter has passed n_occ (same for the whole warp)
    while (true) {
        const unsigned need = __ballot_sync(0xffffffffu, !have);
        if (need && !exhausted) {
            const int leader = __ffs(need) - 1;
            unsigned long long base = 0;
            if ((int)lane == leader) base = atomicAdd(next, (unsigned long long)__popc(need));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (!have) {
                o = (long long)base + __popc(need & lt);
                if (o < n_occ) {
                    const int64_t target = base_occ + o;
                    int64_t lo = first_pat, hi = first_pat + n_chunk_pats;       // invariant: occ_excl[lo] <= target < occ_excl[hi]
                    while (hi - lo > 1) {
                        const int64_t mid = (lo + hi) >> 1;
                        if (occ_excl[mid] <= target) lo = mid; else hi = mid;
                    }
                    idx = sp[lo] + (target - occ_excl[lo]);
                    steps = 0;
                    have = true;
                    keys[o] = (uint64_t)(lo - first_pat) << 32;
                }
            }
            exhausted = (long long)base + __popc(need) > n_occ;
        }
        if (!__any_sync(0xffffffffu, have)) break;
        if (have) {
            const BitRank mk = sector_bit_rank(sectors, t.marker_sector0, (uint32_t)idx);
            if (mk.bit) {
                walked[o] = ((uint64_t)steps << 32) | (uint64_t)(uint32_t)(mk.rank - 1);
                have = false;
            } else if (steps > t.n) {                        // the reference would never return here
                walked[o] = ((uint64_t)steps << 32) | 0xFFFFFFFFull;
                have = false;
            } else {                                         // one LF step, as in locate_row
                long long pos = idx;
                int v = 0;
                while (v >= 0) {
                    const BitRank br = sector_bit_rank(sectors, t.node_sector0[v], (uint32_t)pos);
                    pos = br.bit ? (long long)br.rank - 1 : pos - (long long)br.rank;
                    v = t.child[v][br.bit];
                }
                idx = (long long)(int)(t.c[~v] + pos);
                steps++;
            }
        }
    }
}

__global__ void __launch_bounds__(256)
locate_finish_kernel(const QueryTables* __restrict__ tables, const uint32_t* __restrict__ sectors,
                     const uint64_t* __restrict__ walked, int64_t n_occ, uint64_t* __restrict__ keys) {
    __shared__ QueryTables t;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(tables);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&t);
        for (int i = threadIdx.x; i < (int)(sizeof(QueryTables) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n_occ; o += stride) {
        const uint64_t w = walked[o];
        const uint32_t sampled = (uint32_t)w;
        const long long steps = (long long)(w >> 32);
        const long long pos = sampled == 0xFFFFFFFFu ? -1 : (iwt_get(&t, sectors, (long long)sampled) << t.sampling_factor) + steps;
        keys[o] |= (uint64_t)(uint32_t)pos;
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

int count_batch(gcz_index* idx, const uint8_t* pats, const int64_t* pat_off, int64_t n_pats, int64_t* sp, int64_t* ep) {
    if (!idx || !pats || !pat_off || !sp || !ep || n_pats < 0) return fail(GCZ_E_ARG, "count_batch arguments");
    if (n_pats == 0) return GCZ_OK;
    DeviceCtx* ctx = idx->ctx;
    std::lock_guard<std::mutex> lock(ctx->mu);
    GCZ_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = stream_of(ctx);
    ctx->arena.reset();

    // pattern bytes: the offsets say how many
    int64_t total_bytes = 0;
    const bool off_dev = is_device_ptr(pat_off);
    if (off_dev) GCZ_CUDA(cudaMemcpy(&total_bytes, pat_off + n_pats, 8, cudaMemcpyDeviceToHost));
    else total_bytes = pat_off[n_pats];
    const char* sort_env0 = std::getenv("GCZ_COUNT_SORT");
    const size_t sort_bytes = (sort_env0 && sort_env0[0] == '1') ? (size_t)n_pats * 24 + radix_sort_temp_bytes(n_pats) + (1 << 20) : 0;
    const size_t need = (size_t)total_bytes + (size_t)n_pats * 24 + (1 << 20) + sort_bytes;
    if (ctx->arena.capacity < need) GCZ_TRY(ctx->arena.reserve(need));

    const uint8_t* d_pats = nullptr; const int64_t* d_off = nullptr;
    GCZ_TRY(stage_in(st, ctx->arena, pats, (size_t)total_bytes, &d_pats));
    GCZ_TRY(stage_in(st, ctx->arena, pat_off, (size_t)n_pats + 1, &d_off));
    const bool out_dev = is_device_ptr(sp);
    if (out_dev != is_device_ptr(ep)) return fail(GCZ_E_ARG, "sp and ep must live on the same side");
    int64_t* d_sp = out_dev ? sp : ctx->arena.get<int64_t>((size_t)n_pats);
    int64_t* d_ep = out_dev ? ep : ctx->arena.get<int64_t>((size_t)n_pats);
    if (!d_sp || !d_ep) return fail(GCZ_E_NOMEM, "query staging");

    unsigned long long* d_next = ctx->arena.get<unsigned long long>(1);
    if (!d_next) return fail(GCZ_E_NOMEM, "query staging");
    GCZ_CUDA(cudaMemsetAsync(d_next, 0, 8, st));
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n_pats + 255) / 256, (int64_t)ctx->sm_count * 8));
    const char* sort_env = std::getenv("GCZ_COUNT_SORT");
    if (sort_env && sort_env[0] == '1' && n_pats >= 1024 && n_pats < ((int64_t)1 << 31)) {
        const size_t more = (size_t)n_pats * 24 + radix_sort_temp_bytes(n_pats) + (1 << 20);
        RadixBuffers rb;
        rb.keys[0] = ctx->arena.get<uint64_t>((size_t)n_pats); rb.keys[1] = ctx->arena.get<uint64_t>((size_t)n_pats);
        rb.vals[0] = ctx->arena.get<uint32_t>((size_t)n_pats); rb.vals[1] = ctx->arena.get<uint32_t>((size_t)n_pats);
        void* d_tmp = ctx->arena.raw(radix_sort_temp_bytes(n_pats));
        if (!rb.keys[0] || !rb.keys[1] || !rb.vals[0] || !rb.vals[1] || !d_tmp) return fail(GCZ_E_NOMEM, "query staging (%zu more bytes)", more);
        SymbolIds ids;
        int next_id = 1;
        for (int ch = 0; ch < 256; ch++) {
            const int64_t here = (ch < 255 ? idx->c[ch + 1] : idx->n) - idx->c[ch];
            ids.id[ch] = here > 0 ? (uint8_t)(next_id++ & 31) : 0;
        }
        GCZ_LAUNCH(ctx, pattern_suffix_keys_kernel, launch_grid(ctx, n_pats, 256), 256, 0, st, d_pats, d_off, n_pats, ids, rb.keys[0], rb.vals[0]);
        GCZ_TRY(radix_sort_pairs(ctx, st, rb, n_pats, 0, 60, d_tmp, nullptr));
        GCZ_LAUNCH(ctx, count_kernel<true>, grid, 256, 0, st, idx->d_tables, idx->d_sectors, d_pats, d_off, n_pats, d_sp, d_ep, d_next,
                   (const uint32_t*)rb.vals[rb.cur]);
    } else {
        GCZ_LAUNCH(ctx, count_kernel<false>, grid, 256, 0, st, idx->d_tables, idx->d_sectors, d_pats, d_off, n_pats, d_sp, d_ep, d_next,
                   (const uint32_t*)nullptr);
    }
    if (!out_dev) {
        GCZ_CUDA(cudaMemcpyAsync(sp, d_sp, (size_t)n_pats * 8, cudaMemcpyDeviceToHost, st));
        GCZ_CUDA(cudaMemcpyAsync(ep, d_ep, (size_t)n_pats * 8, cudaMemcpyDeviceToHost, st));
    }
    GCZ_CUDA(cudaStreamSynchronize(st));
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

// java.util.Arrays.binarySearch(long[] a, int from, int to, long key)
static int64_t java_binary_search(const int64_t* a, int64_t from, int64_t to, int64_t key) {
    int64_t low = from, high = to - 1;
    while (low <= high) {
        const int64_t mid = (int64_t)(((uint64_t)low + (uint64_t)high) >> 1);
        if (a[mid] < key) low = mid + 1;
        else if (a[mid] > key) high = mid - 1;
        else return mid;
    }
    return -(low + 1);
}

int find_batch(gcz_index* idx, const uint8_t* pats, const int64_t* pat_off, int64_t n_pats,
               int64_t* per_string_counts, int64_t** positions, int64_t** pos_off) {
    if (!idx || !pats || !pat_off || !positions || !pos_off || n_pats < 0) return fail(GCZ_E_ARG, "find_batch arguments");
    if (is_device_ptr(pats) || is_device_ptr(pat_off)) return fail(GCZ_E_ARG, "find_batch takes host pattern buffers");
    DeviceCtx* ctx = idx->ctx;
    const int64_t ns = (int64_t)idx->e.size();
    *positions = nullptr; *pos_off = nullptr;

    // 1. intervals
    std::vector<int64_t> sp((size_t)n_pats), ep((size_t)n_pats);
    if (n_pats > 0) GCZ_TRY(count_batch(idx, pats, pat_off, n_pats, sp.data(), ep.data()));
    std::vector<int64_t> occ_excl((size_t)n_pats + 1, 0);
    for (int64_t i = 0; i < n_pats; i++) occ_excl[(size_t)i + 1] = occ_excl[(size_t)i] + std::max<int64_t>(0, ep[(size_t)i] - sp[(size_t)i] + 1);
    const int64_t total_occ = occ_excl[(size_t)n_pats];

    int64_t* h_pos = static_cast<int64_t*>(std::malloc(sizeof(int64_t) * (size_t)std::max<int64_t>(total_occ, 1)));
    int64_t* h_off = static_cast<int64_t*>(std::malloc(sizeof(int64_t) * ((size_t)n_pats + 1)));
    if (!h_pos || !h_off) { std::free(h_pos); std::free(h_off); return fail(GCZ_E_NOMEM, "find_batch results"); }
    if (per_string_counts) std::memset(per_string_counts, 0, sizeof(int64_t) * (size_t)(n_pats * ns));

    std::lock_guard<std::mutex> lock(ctx->mu);
    GCZ_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = stream_of(ctx);

    // 2. locate + sort in chunks of at most kChunk occurrences (a single pattern may exceed it)
    const int64_t kChunk = (int64_t)1 << 26;
    std::vector<uint64_t> h_keys;
    int64_t written = 0;
    h_off[0] = 0;
    int64_t p0 = 0;
    while (p0 < n_pats) {
        int64_t p1 = p0 + 1;
        while (p1 < n_pats && occ_excl[(size_t)p1 + 1] - occ_excl[(size_t)p0] <= kChunk) p1++;
        const int64_t n_occ = occ_excl[(size_t)p1] - occ_excl[(size_t)p0];
        const int64_t np = p1 - p0;
        if (n_occ > 0) {
            ctx->arena.reset();
            const size_t need = (size_t)n_occ * 16 + (size_t)np * 16 + radix_sort_temp_bytes(n_occ) + (1 << 20);
            if (ctx->arena.capacity < need) { int rc = ctx->arena.reserve(need); if (rc) { std::free(h_pos); std::free(h_off); return rc; } }
            int64_t* d_sp = ctx->arena.get<int64_t>((size_t)np);
            int64_t* d_ex = ctx->arena.get<int64_t>((size_t)np + 1);
            uint64_t* d_k0 = ctx->arena.get<uint64_t>((size_t)n_occ);
            uint64_t* d_k1 = ctx->arena.get<uint64_t>((size_t)n_occ);
            void* d_tmp = ctx->arena.raw(radix_sort_temp_bytes(n_occ));
            if (!d_sp || !d_ex || !d_k0 || !d_k1 || !d_tmp) { std::free(h_pos); std::free(h_off); return fail(GCZ_E_NOMEM, "find_batch workspace"); }
            GCZ_CUDA(cudaMemcpyAsync(d_sp, sp.data() + p0, (size_t)np * 8, cudaMemcpyHostToDevice, st));
            GCZ_CUDA(cudaMemcpyAsync(d_ex, occ_excl.data() + p0, ((size_t)np + 1) * 8, cudaMemcpyHostToDevice, st));
            const char* locate_env = std::getenv("GCZ_LOCATE_VARIANT");
            if (locate_env && locate_env[0] == '1') {
                unsigned long long* d_next = ctx->arena.get<unsigned long long>(2);
                if (!d_next) { std::free(h_pos); std::free(h_off); return fail(GCZ_E_NOMEM, "find_batch workspace"); }
                GCZ_CUDA(cudaMemsetAsync(d_next, 0, 16, st));
                const int walk_grid = (int)std::max<int64_t>(1, std::min<int64_t>((n_occ + 255) / 256, (int64_t)ctx->sm_count * 8));
                GCZ_LAUNCH(ctx, locate_walk_refill_kernel, walk_grid, 256, 0, st, idx->d_tables, idx->d_sectors, d_sp, d_ex, (int64_t)0, np,
                           occ_excl[(size_t)p0], n_occ, d_k0, d_k1, d_next);
                GCZ_LAUNCH(ctx, locate_finish_kernel, launch_grid(ctx, n_occ, 256), 256, 0, st, idx->d_tables, idx->d_sectors, d_k1, n_occ, d_k0);
            } else {
                GCZ_LAUNCH(ctx, locate_occurrences_kernel, launch_grid(ctx, n_occ, 256), 256, 0, st, idx->d_tables, idx->d_sectors,
                           d_sp, d_ex, (int64_t)0, np, occ_excl[(size_t)p0], n_occ, d_k0);
            }
            RadixBuffers rb;
            rb.keys[0] = d_k0; rb.keys[1] = d_k1;
            int pat_bits = 1;
            while (((int64_t)1 << pat_bits) < np) pat_bits++;
            int rc = radix_sort_pairs(ctx, st, rb, n_occ, 0, 32 + pat_bits, d_tmp, nullptr);
            if (rc) { std::free(h_pos); std::free(h_off); return rc; }
            h_keys.resize((size_t)n_occ);
            GCZ_CUDA(cudaMemcpyAsync(h_keys.data(), rb.keys[rb.cur], (size_t)n_occ * 8, cudaMemcpyDeviceToHost, st));
            GCZ_CUDA(cudaStreamSynchronize(st));
        }
        // 3. GSSA.find :170-184 per pattern: split the ascending positions by the string ends
        std::vector<int64_t> sa;
        for (int64_t p = p0; p < p1; p++) {
            const int64_t k = occ_excl[(size_t)p + 1] - occ_excl[(size_t)p];
            const int64_t first = occ_excl[(size_t)p] - occ_excl[(size_t)p0];
            sa.resize((size_t)k);
            for (int64_t j = 0; j < k; j++) sa[(size_t)j] = (int64_t)(int32_t)(uint32_t)(h_keys[(size_t)(first + j)] & 0xFFFFFFFFull);
            int64_t idx1 = 0;
            for (int64_t i = 0; i < ns && k > 0; i++) {
                const int64_t idx2 = -java_binary_search(sa.data(), idx1, k, idx->e[(size_t)i]) - 1;
                if (idx2 > idx1) {
                    const int64_t start = i > 0 ? idx->e[(size_t)i - 1] + 1 : 0;
                    if (per_string_counts) per_string_counts[p * ns + i] = idx2 - idx1;
                    for (int64_t j = idx1; j < idx2; j++) h_pos[written++] = sa[(size_t)j] - start;
                    idx1 = idx2;
                }
            }
            h_off[p + 1] = written;
        }
        p0 = p1;
    }
    *positions = h_pos;
    *pos_off = h_off;
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

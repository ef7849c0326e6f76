// BWT, Huffman-shaped wavelet tree, sampled-SA index and the ranked bit-vector layout, built on the GPU.
//
// Replaces, per block (fmt/GecozFileWriter.java:256-284):
//   BWTDataSource.get                    fmt/GecozFileWriter.java:301-303   -> bwt_count_kernel
//   HuffmanShapedWaveletTree.fill        algo/tree/HuffmanShapedWaveletTree.java:127-146 -> hswt_emit_kernel
//   RankedWTNode.putLong (counters)      algo/tree/RankedWTNode.java:228-245 -> ranked_layout_kernel
//   GSSAIndex write ctor                 algo/ssa/GSSAIndex.java:129-150    -> marker bits + sample_kernel
//   IndexWaveletTree build ctor          algo/tree/IndexWaveletTree.java:83-112 -> iwt_* kernels
// The reference streams one symbol at a time into bit writers; here every structure is produced by
// prefix sums over warp tiles (a bit's position in a node = number of earlier symbols routed to that node).
#include "wavelet_build.cuh"
#include "row_scan.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace gcz {

namespace {

constexpr int kWarpTile = 1024;             // BWT positions per warp
constexpr int kWtThreads = 256;             // 8 warps
constexpr int kNoNode = 0xFFFF;

// symbol tables of one block, resident in global memory for the duration of the build
struct SymbolTables {
    uint8_t  dense[256];        // byte -> dense symbol index, 0xFF when absent
    uint8_t  byte_of[256];      // dense symbol index -> byte
    uint16_t code[256];         // dense -> code bits (bit d = branch at depth d)
    uint8_t  len[256];          // dense -> code length
    uint8_t  node_of[256][16];  // dense, depth -> node (file order) on the symbol's path
    uint32_t path_mask[256];    // dense -> nodes on the symbol's path, one bit per node (trees of <= 32 nodes)
    uint32_t bit_mask[256];     // dense -> those of them where the symbol branches to the 1-child
    uint16_t node_prefix[256];  // node -> path bits
    uint8_t  node_depth[256];   // node -> depth
    int32_t  sigma;             // symbols present
    int32_t  n_nodes;
    int32_t  max_len;
};

// ---- BWT gather + per-tile symbol counts + marker bits -----------------------------------------------
// carry_shift != 0: SA entries hold the dense code (1..sigma) of the preceding text symbol above bit carry_shift
// (suffix_sort.cuh) — the BWT is read off the entries and, when sa_clean is set, plain positions are written back.
// PACKED (alphabets of up to 8 symbols): every lane counts its own symbols in eight 8-bit
// fields of one register (32 symbols per lane and tile, so no field overflows) and the warp adds the fields up once per
// tile, instead of one ballot round per distinct symbol of every 32-symbol chunk.
template <bool PACKED>
__global__ void __launch_bounds__(kWtThreads)
bwt_count_kernel(const uint8_t* __restrict__ text, const uint32_t* __restrict__ sa, int64_t n,
                 const SymbolTables* __restrict__ tab, uint32_t sample_mask, int carry_shift, uint32_t* __restrict__ sa_clean,
                 uint8_t* __restrict__ bwt, uint32_t* __restrict__ marker_raw,
                 uint32_t* __restrict__ tile_counts /* [(sigma + 1)][tiles] */, int64_t tiles) {
    __shared__ uint32_t s_cnt[kWtThreads / 32][260];
    __shared__ uint8_t s_dense[256], s_byte_of[256];
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    s_dense[threadIdx.x] = tab->dense[threadIdx.x];
    s_byte_of[threadIdx.x] = tab->byte_of[threadIdx.x];
    const int sigma = tab->sigma;
    for (int i = lane; i <= sigma; i += 32) s_cnt[warp][i] = 0;
    __syncthreads();
    const int64_t tile = (int64_t)blockIdx.x * (kWtThreads / 32) + warp;
    if (tile >= tiles) return;
    const int64_t base = tile * kWarpTile;
    const uint32_t pos_mask = carry_shift ? (1u << carry_shift) - 1u : 0xFFFFFFFFu;
    uint32_t* cnt = s_cnt[warp];
    unsigned marks = 0;
    unsigned long long packed = 0;
#pragma unroll 1
    for (int c0 = 0; c0 < kWarpTile / 32; c0 += 4) {
        uint32_t s[4];
        int d[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int64_t j = base + (c0 + u) * 32 + lane;
            s[u] = j < n ? sa[j] : 1u;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int64_t j = base + (c0 + u) * 32 + lane;
            if (carry_shift) {
                d[u] = j < n ? (int)(s[u] >> carry_shift) - 1 : -1;
                s[u] &= pos_mask;
            } else {
                d[u] = j < n ? (int)s_dense[text[s[u] ? (int64_t)s[u] - 1 : n - 1]] : -1;
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int64_t j = base + (c0 + u) * 32 + lane;
            const bool valid = j < n;
            if (valid) bwt[j] = s_byte_of[d[u]];
            if (valid && sa_clean) sa_clean[j] = s[u];
            const unsigned mk = __ballot_sync(0xffffffffu, valid && (s[u] & sample_mask) == 0);
            if (lane == 0 && base + (c0 + u) * 32 < n) marker_raw[(base >> 5) + c0 + u] = mk;
            marks += (lane == 0) ? __popc(mk) : 0;
            if (PACKED) {
                if (valid) packed += 1ull << (8 * d[u]);
                continue;
            }
            // count every distinct symbol of this chunk once
            unsigned todo = __ballot_sync(0xffffffffu, valid);
            while (todo) {
                const int leader = __ffs(todo) - 1;
                const int v = __shfl_sync(0xffffffffu, d[u], leader);
                const unsigned same = __ballot_sync(0xffffffffu, d[u] == v);
                if ((int)lane == leader) cnt[v] += __popc(same);
                todo &= ~same;
            }
            __syncwarp();
        }
    }
    if (PACKED) {
        // symbols 0, 2, 4, 6 and 1, 3, 5, 7 in 16-bit fields: a tile holds 1024 symbols
        unsigned long long even = packed & 0x00FF00FF00FF00FFull, odd = (packed >> 8) & 0x00FF00FF00FF00FFull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            even += __shfl_xor_sync(0xffffffffu, even, o);
            odd += __shfl_xor_sync(0xffffffffu, odd, o);
        }
        if ((int)lane < sigma) cnt[lane] = (uint32_t)(((lane & 1u) ? odd : even) >> (16 * (lane >> 1))) & 0xFFFFu;
    }
    __syncwarp();
    if (lane == 0) cnt[sigma] = marks;
    __syncwarp();
    for (int i = lane; i <= sigma; i += 32) tile_counts[(size_t)i * tiles + tile] = cnt[i];
}

// ---- node bit emission ------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWtThreads)
hswt_emit_kernel(const uint8_t* __restrict__ bwt, int64_t n, const SymbolTables* __restrict__ tab,
                 const uint32_t* __restrict__ tile_prefix /* [sigma][tiles], exclusive */, int64_t tiles,
                 const uint64_t* __restrict__ node_raw_word /* node -> first u32 word of its raw vector */,
                 uint32_t* __restrict__ raw) {
    __shared__ SymbolTables s_tab;
    __shared__ unsigned long long s_acc[kWtThreads / 32][256];
    __shared__ uint32_t s_word[kWtThreads / 32][256];
    __shared__ uint8_t s_fill[kWtThreads / 32][256];
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(tab);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&s_tab);
        for (int i = threadIdx.x; i < (int)(sizeof(SymbolTables) / 4); i += kWtThreads) dst[i] = src[i];
    }
    __syncthreads();
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    const unsigned lt = lanemask_lt();
    const int64_t tile = (int64_t)blockIdx.x * (kWtThreads / 32) + warp;
    if (tile >= tiles) return;
    const int sigma = s_tab.sigma, n_nodes = s_tab.n_nodes, max_len = s_tab.max_len;

    // where this tile's bits start in every node: symbols routed through the node, summed over earlier tiles
    for (int v = lane; v < n_nodes; v += 32) {
        const int depth = s_tab.node_depth[v];
        const unsigned prefix = s_tab.node_prefix[v], pmask = (1u << depth) - 1;
        unsigned long long bit0 = 0;
        for (int s = 0; s < sigma; s++) {
            if (s_tab.len[s] > depth && (s_tab.code[s] & pmask) == prefix) bit0 += tile_prefix[(size_t)s * tiles + tile];
        }
        bit0 += node_raw_word[v] * 32ull;
        s_acc[warp][v] = 0;
        s_word[warp][v] = (uint32_t)(bit0 >> 5);     // raw areas are far below 2^32 words
        s_fill[warp][v] = (uint8_t)(bit0 & 31);
    }
    __syncwarp();

    const int64_t base = tile * kWarpTile;
#pragma unroll 1
    for (int c = 0; c < kWarpTile / 32; c++) {
        const int64_t j = base + c * 32 + lane;
        if (base + c * 32 >= n) break;
        const int d = j < n ? (int)s_tab.dense[bwt[j]] : 255;
        const int len = j < n ? (int)s_tab.len[d] : 0;
        const unsigned code = j < n ? s_tab.code[d] : 0;
        for (int depth = 0; depth < max_len; depth++) {
            const int nid = depth < len ? (int)s_tab.node_of[d][depth] : kNoNode;
            const unsigned bit = (code >> depth) & 1u;
            unsigned todo = __ballot_sync(0xffffffffu, nid != kNoNode);
            while (todo) {
                const int leader = __ffs(todo) - 1;
                const int v = __shfl_sync(0xffffffffu, nid, leader);
                const bool member = nid == v;
                const unsigned members = __ballot_sync(0xffffffffu, member);
                const unsigned packed = __reduce_or_sync(0xffffffffu, member ? (bit << __popc(members & lt)) : 0u);
                if ((int)lane == leader) {
                    unsigned long long acc = s_acc[warp][v] | ((unsigned long long)packed << s_fill[warp][v]);
                    unsigned fill = s_fill[warp][v] + __popc(members);
                    if (fill >= 32) {
                        atomicOr(&raw[s_word[warp][v]], (uint32_t)acc);
                        s_word[warp][v]++;
                        acc >>= 32;
                        fill -= 32;
                    }
                    s_acc[warp][v] = acc;
                    s_fill[warp][v] = (uint8_t)fill;
                }
                todo &= ~members;
                __syncwarp();
            }
        }
    }
    __syncwarp();
    for (int v = lane; v < n_nodes; v += 32) {
        if (s_fill[warp][v] > 0 && (uint32_t)s_acc[warp][v] != 0) atomicOr(&raw[s_word[warp][v]], (uint32_t)s_acc[warp][v]);
    }
}

// Trees of at most 32 internal nodes (any DNA/IUPAC text): lane v of the warp owns node v and keeps its bit
// accumulator in registers.  Per 32 BWT symbols and node: the lanes routed through the node (one ballot), their
// branch bits compacted to the low end (`pext` spelled as shift + warp OR-reduction), appended by the owning lane.
__global__ void __launch_bounds__(kWtThreads)
hswt_emit_small_kernel(const uint8_t* __restrict__ bwt, int64_t n, const SymbolTables* __restrict__ tab,
                       const uint32_t* __restrict__ tile_prefix /* [sigma][tiles], exclusive */, int64_t tiles,
                       const uint64_t* __restrict__ node_raw_word /* node -> first u32 word of its raw vector */,
                       uint32_t* __restrict__ raw) {
    __shared__ uint32_t s_path[256], s_bit[256];
    __shared__ uint8_t s_dense[256];
    s_path[threadIdx.x] = tab->path_mask[threadIdx.x];
    s_bit[threadIdx.x] = tab->bit_mask[threadIdx.x];
    s_dense[threadIdx.x] = tab->dense[threadIdx.x];
    __syncthreads();
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    const unsigned lt = lanemask_lt();
    const int64_t tile = (int64_t)blockIdx.x * (kWtThreads / 32) + warp;
    if (tile >= tiles) return;
    const int sigma = tab->sigma, n_nodes = tab->n_nodes;

    // where this tile's bits start in "my" node: symbols routed through it, summed over earlier tiles
    unsigned long long acc = 0;
    uint32_t word = 0;
    unsigned fill = 0;
    if ((int)lane < n_nodes) {
        unsigned long long bit0 = node_raw_word[lane] * 32ull;
        for (int s = 0; s < sigma; s++) {
            if ((s_path[s] >> lane) & 1u) bit0 += tile_prefix[(size_t)s * tiles + tile];
        }
        word = (uint32_t)(bit0 >> 5);                    // raw areas are far below 2^32 words
        fill = (unsigned)(bit0 & 31);
    }
    const int64_t base = tile * kWarpTile;
#pragma unroll 1
    for (int c = 0; c < kWarpTile / 32; c++) {
        if (base + c * 32 >= n) break;
        const int64_t j = base + c * 32 + lane;
        const int d = j < n ? (int)s_dense[bwt[j]] : -1;
        const uint32_t pm = d >= 0 ? s_path[d] : 0u;
        const uint32_t bm = d >= 0 ? (s_bit[d] & pm) : 0u;
        unsigned my_packed = 0, my_count = 0;
        for (int v = 0; v < n_nodes; v++) {
            const unsigned members = __ballot_sync(0xffffffffu, (pm >> v) & 1u);
            if (members == 0) continue;
            const unsigned packed = __reduce_or_sync(0xffffffffu, ((bm >> v) & 1u) << __popc(members & lt));
            if ((int)lane == v) { my_packed = packed; my_count = __popc(members); }
        }
        acc |= (unsigned long long)my_packed << fill;
        fill += my_count;
        if (fill >= 32) {
            atomicOr(&raw[word], (uint32_t)acc);
            word++;
            acc >>= 32;
            fill -= 32;
        }
    }
    if (fill > 0 && (uint32_t)acc != 0) atomicOr(&raw[word], (uint32_t)acc);
}

// Trees of at most kLutNodes nodes (alphabets of up to 8 symbols: the path every DNA block takes),
// i.e. alphabets of up to 8 symbols.  Every lane takes 32 consecutive BWT symbols and appends to its own per-node
// accumulators, four symbols at a time through a table indexed by the four dense codes (entry: for every node one byte,
// the branch bits of those of the four symbols that pass the node, low bit first, and how many they are above bit 4).
// A warp-wide prefix over the per-lane bit counts places each lane's fragment; fragments meet in shared memory and the
// tile's words go out with plain stores, except the first and last word of a node, which neighbouring tiles share.
// About 12 instructions per symbol and lane instead of one ballot + reduction per node and 32 symbols.
constexpr int kLutNodes = 8;
constexpr int kLutStageWords = 34;                      // 31 + 1024 bits of one tile in one node, + 1 for the carry word

template <int NODES>
__global__ void __launch_bounds__(kWtThreads)
hswt_emit_lut_kernel(const uint8_t* __restrict__ bwt, int64_t n, const SymbolTables* __restrict__ tab,
                     const uint2* __restrict__ lut, int lut_entries,
                     const uint32_t* __restrict__ tile_prefix /* [sigma][tiles], exclusive */, int64_t tiles,
                     const uint64_t* __restrict__ node_raw_word, uint32_t* __restrict__ raw) {
    extern __shared__ __align__(16) unsigned char s_dyn[];
    uint2* s_lut = reinterpret_cast<uint2*>(s_dyn);
    uint32_t* s_stage = reinterpret_cast<uint32_t*>(s_lut + lut_entries);          // [warps][NODES][kLutStageWords]
    __shared__ uint32_t s_path[256], s_bit[256];
    __shared__ uint8_t s_dense[256];
    for (int i = threadIdx.x; i < lut_entries; i += kWtThreads) s_lut[i] = lut[i];
    s_path[threadIdx.x] = tab->path_mask[threadIdx.x];
    s_bit[threadIdx.x] = tab->bit_mask[threadIdx.x];
    s_dense[threadIdx.x] = tab->dense[threadIdx.x];
    __syncthreads();
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    const unsigned sigma = (unsigned)tab->sigma;
    uint32_t* stage = s_stage + warp * (NODES * kLutStageWords);
    const bool aligned = (reinterpret_cast<uintptr_t>(bwt) & 15) == 0;

    for (int64_t tile = (int64_t)blockIdx.x * (kWtThreads / 32) + warp; tile < tiles; tile += (int64_t)gridDim.x * (kWtThreads / 32)) {
        // where this tile's bits start in "my" node: symbols routed through it, summed over earlier tiles
        uint32_t my_word = 0, my_fill = 0;
        if ((int)lane < NODES) {
            unsigned long long bit0 = node_raw_word[lane] * 32ull;
            for (unsigned s = 0; s < sigma; s++) {
                if ((s_path[s] >> lane) & 1u) bit0 += tile_prefix[(size_t)s * tiles + tile];
            }
            my_word = (uint32_t)(bit0 >> 5);                   // raw areas are far below 2^32 words
            my_fill = (uint32_t)(bit0 & 31);
        }
        for (int i = lane; i < NODES * kLutStageWords; i += 32) stage[i] = 0;
        __syncwarp();

        uint32_t acc[NODES], cnt[NODES];
#pragma unroll
        for (int v = 0; v < NODES; v++) { acc[v] = 0; cnt[v] = 0; }
        const int64_t p0 = tile * kWarpTile + (int64_t)lane * 32;
        if (aligned && p0 + 32 <= n) {
            const uint4 q0 = *reinterpret_cast<const uint4*>(bwt + p0), q1 = *reinterpret_cast<const uint4*>(bwt + p0 + 16);
            const uint32_t w[8] = { q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w };
#pragma unroll
            for (int g = 0; g < 8; g++) {
                const uint32_t x = w[g];                       // four symbols, the first one in the low byte
                const unsigned idx = (((unsigned)s_dense[x & 255] * sigma + s_dense[(x >> 8) & 255]) * sigma + s_dense[(x >> 16) & 255]) * sigma
                                     + s_dense[x >> 24];
                const uint2 e = s_lut[idx];
#pragma unroll
                for (int v = 0; v < NODES; v++) {
                    const uint32_t b = ((v < 4 ? e.x : e.y) >> (8 * (v & 3))) & 255u;
                    acc[v] |= (b & 15u) << cnt[v];             // never shifted out: a lane holds 32 symbols
                    cnt[v] += b >> 4;
                }
            }
        } else {
            for (int i = 0; i < 32 && p0 + i < n; i++) {
                const unsigned d = s_dense[bwt[p0 + i]];
                const uint32_t pm = s_path[d], bm = s_bit[d] & pm;
#pragma unroll
                for (int v = 0; v < NODES; v++) {
                    acc[v] |= (((pm >> v) & 1u) ? ((bm >> v) & 1u) : 0u) << cnt[v];
                    cnt[v] += (pm >> v) & 1u;
                }
            }
        }

#pragma unroll
        for (int v = 0; v < NODES; v++) {
            unsigned incl = cnt[v];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
                if ((int)lane >= o) incl += t;
            }
            const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
            const unsigned fill0 = __shfl_sync(0xffffffffu, my_fill, v), word0 = __shfl_sync(0xffffffffu, my_word, v);
            uint32_t* node_stage = stage + v * kLutStageWords;
            if (cnt[v]) {
                const unsigned o = fill0 + (incl - cnt[v]), wi = o >> 5, sh = o & 31;
                atomicOr(&node_stage[wi], acc[v] << sh);
                if (sh && sh + cnt[v] > 32) atomicOr(&node_stage[wi + 1], acc[v] >> (32 - sh));
            }
            __syncwarp();
            const unsigned bits = fill0 + total, words = (bits + 31) >> 5;
            for (unsigned wd = lane; wd < words; wd += 32) {
                const uint32_t val = node_stage[wd];
                const bool shared_word = (wd == 0 && fill0 != 0) || (wd == words - 1 && (bits & 31) != 0);
                if (!shared_word) raw[word0 + wd] = val;       // this tile owns every bit of the word
                else if (val) atomicOr(&raw[word0 + wd], val);
            }
        }
        __syncwarp();
    }
}

// ---- sampled suffix array values in SA order -----------------------------------------------------------
__global__ void __launch_bounds__(kWtThreads)
sample_kernel(const uint32_t* __restrict__ sa, int64_t n, uint32_t pos_mask, uint32_t sample_mask, int sample_shift,
              const uint32_t* __restrict__ marker_prefix /* [tiles] exclusive */, int64_t tiles,
              uint32_t* __restrict__ ssa) {
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    const int64_t tile = (int64_t)blockIdx.x * (kWtThreads / 32) + warp;
    if (tile >= tiles) return;
    const int64_t base = tile * kWarpTile;
    uint32_t out = marker_prefix[tile];
#pragma unroll 4
    for (int c = 0; c < kWarpTile / 32; c++) {
        const int64_t j = base + c * 32 + lane;
        const uint32_t s = j < n ? sa[j] & pos_mask : 1u;
        const bool mk = j < n && (s & sample_mask) == 0;
        const unsigned m = __ballot_sync(0xffffffffu, mk);
        if (mk) ssa[out + __popc(m & lanemask_lt())] = s >> sample_shift;
        out += __popc(m);
    }
}

// ---- IndexWaveletTree levels ---------------------------------------------------------------------------
// The values are a permutation of 0 .. m - 1 (sampled suffix array entries >> sampling factor).  Level h of the tree shows
// bit h of the values in the order "grouped (stably) by v >> (h + 1)", group g at slot g << (h + 1).
//
// High levels (groups longer than a CTA's block): up to kTopBits levels per pass.  A pass over bits H .. H - L + 1 starts from
// the order grouped by v >> (H + 1) and ends in the order grouped by v >> (H - L + 1):
//   count   per block of kIwtBlock values, how many carry each value of the L-bit digit;
//   scan    the counts over the blocks (row_scan.cuh); a block's parent group starts on a block boundary, so the number of
//           earlier values of the same parent group with digit b is scanned[b][block] - scanned[b][first block of the parent];
//   emit    the block runs the L levels in shared memory like the low-level kernel below: at sub-level l its values are
//           grouped by the first l bits of the digit, group q's bits are consecutive in the block AND at their destination —
//           after the values of q in earlier blocks, whose number the scan gives — so they go out as one bit run per group
//           (whole words stored, edge words by atomicOr), then every group is split by the bit; after the last split the block
//           is sorted by digit and each digit's run is copied to its place in the next order.
// The first version ran three launches per level over global memory (74 us per level of 7.8 M values, ten levels per chr1 block).
constexpr int kIwtLowBits = 13;
constexpr int kIwtBlock = 1 << kIwtLowBits;              // values per CTA
constexpr int kIwtThreads = 1024;
constexpr int kIwtPer = kIwtBlock / kIwtThreads;         // consecutive values per thread
static_assert(kIwtPer == 8, "one byte of level bits per thread");
constexpr int kTopBits = 5;
constexpr int kTopBins = 1 << kTopBits;

__global__ void __launch_bounds__(kIwtThreads)
iwt_top_count_kernel(const uint32_t* __restrict__ vals, int64_t m, int low_bit, int L, uint32_t* __restrict__ counts /* [1 << L][blocks] */,
                     int64_t blocks) {
    __shared__ unsigned s_c[kTopBins];
    if (threadIdx.x < kTopBins) s_c[threadIdx.x] = 0;
    __syncthreads();
    const int64_t block_start = (int64_t)blockIdx.x * kIwtBlock;
    const unsigned mask = (1u << L) - 1u;
#pragma unroll
    for (int i = 0; i < kIwtPer; i++) {
        const int64_t p = block_start + i * kIwtThreads + threadIdx.x;
        if (p < m) atomicAdd(&s_c[(vals[p] >> low_bit) & mask], 1u);
    }
    __syncthreads();
    if (threadIdx.x <= mask) counts[(size_t)threadIdx.x * blocks + blockIdx.x] = s_c[threadIdx.x];
}

constexpr size_t kIwtTopSmem = (size_t)(2 * kIwtBlock + kIwtThreads + 32 + kIwtBlock / 32 + 8 + kIwtThreads / 4) * 4;

__global__ void __launch_bounds__(kIwtThreads, 2)
iwt_top_emit_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int64_t m64, int H, int L, int levels,
                    uint32_t* __restrict__ raw, const uint64_t* __restrict__ level_raw_word,
                    const uint32_t* __restrict__ scanned /* [1 << L][blocks], exclusive over the blocks */, int64_t blocks) {
    extern __shared__ __align__(16) uint32_t s_iwt[];
    uint32_t* cur = s_iwt;
    uint32_t* nxt = s_iwt + kIwtBlock;
    uint32_t* s_zc = s_iwt + 2 * kIwtBlock;              // zeros before each thread's first value
    uint32_t* s_wsum = s_zc + kIwtThreads;               // 32
    uint32_t* s_bits = s_wsum + 32;                      // kIwtBlock / 32 words of level bits (+ 8 words of padding)
    uint8_t* s_zm = reinterpret_cast<uint8_t*>(s_bits + kIwtBlock / 32 + 8);   // zero masks, one byte per thread
    __shared__ unsigned s_c[kTopBins];                   // values of the block per digit
    __shared__ unsigned s_pc[kTopBins + 1];              // ... with a smaller digit: where a digit's (or a group's) values start in the block
    __shared__ unsigned s_pb[kTopBins + 1];              // values of the parent group with a smaller digit in earlier blocks

    // slots and values are below m <= 2^31: 32-bit arithmetic throughout
    const unsigned m = (unsigned)m64;
    const unsigned block_start = blockIdx.x * (unsigned)kIwtBlock;
    const int cnt = (int)min((unsigned)kIwtBlock, m - block_start);
    const int t = threadIdx.x;
    const int low_bit = H - L + 1;
    const unsigned dmask = (1u << L) - 1u;
    const unsigned parent = H + 1 >= 32 ? 0u : (block_start >> (H + 1)) << (H + 1);
    const unsigned parent_len = H + 1 >= 31 ? m - parent : min(1u << (H + 1), m - parent);
    if (t < kTopBins) s_c[t] = 0;
    if (t < kIwtBlock / 32 + 8) s_bits[t] = 0;
    __syncthreads();
    for (int i = 0; i < kIwtPer; i++) {
        const int e = i * kIwtThreads + t;
        const uint32_t v = e < cnt ? in[block_start + e] : 0xFFFFFFFFu;
        cur[e] = v;
        if (e < cnt) atomicAdd(&s_c[(v >> low_bit) & dmask], 1u);
    }
    __syncthreads();
    if (t < 32) {                                        // prefix sums over the digits (warp 0)
        const unsigned c = t <= (int)dmask ? s_c[t] : 0u;
        const unsigned before = t <= (int)dmask ? scanned[(size_t)t * blocks + blockIdx.x] - scanned[(size_t)t * blocks + (parent >> kIwtLowBits)] : 0u;
        const unsigned ic = warp_incl_sum(c), ib = warp_incl_sum(before);
        s_pc[t + 1] = ic; s_pb[t + 1] = ib;
        if (t == 0) { s_pc[0] = 0; s_pb[0] = 0; }
    }
    __syncthreads();

    const int e0 = t * kIwtPer;
    const unsigned valid = e0 >= cnt ? 0u : (e0 + kIwtPer <= cnt ? 0xFFu : (1u << (cnt - e0)) - 1u);
    for (int l = 0; l < L; l++) {
        const int h = H - l;
        const int groups = 1 << l, span_shift = L - l;                        // 1 << span_shift digits per group
        const uint4 q0 = reinterpret_cast<const uint4*>(cur)[2 * t], q1 = reinterpret_cast<const uint4*>(cur)[2 * t + 1];
        const uint32_t v[kIwtPer] = { q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w };
        unsigned bits = 0;
#pragma unroll
        for (int i = 0; i < kIwtPer; i++) bits |= ((v[i] >> h) & 1u) << i;
        bits &= valid;
        const unsigned zmask = ~bits & valid;
        reinterpret_cast<uint8_t*>(s_bits)[t] = (uint8_t)bits;
        s_zm[t] = (uint8_t)zmask;
        const unsigned z = __popc(zmask);
        const unsigned incl = warp_incl_sum(z);
        if (lane_id() == 31) s_wsum[t >> 5] = incl;
        __syncthreads();
        const unsigned wincl = warp_incl_sum(s_wsum[lane_id()]);                         // every warp scans the warp totals
        const unsigned wbase = __shfl_sync(0xffffffffu, wincl - s_wsum[lane_id()], t >> 5);
        const unsigned zexcl = wbase + incl - z;
        s_zc[t] = zexcl;
        // the level's bits: one run per group (group q = digits q << span_shift .. : consecutive in the block and at the destination)
        uint32_t* lraw = raw + level_raw_word[levels - 1 - h];
        for (int q = 0; q < groups; q++) {
            const unsigned A = s_pc[q << span_shift], C = s_pc[(q + 1) << span_shift] - A;
            if (C == 0) continue;
            const unsigned off = h + 1 >= 32 ? 0u : (unsigned)q << (h + 1);                      // the group's first slot inside the parent group
            const unsigned D = parent + min(off, parent_len) + (s_pb[(q + 1) << span_shift] - s_pb[q << span_shift]);
            const unsigned w0 = D >> 5;
            const unsigned nwords = ((D & 31u) + C + 31u) >> 5;
            for (unsigned k = t; k < nwords; k += kIwtThreads) {
                const unsigned lo = max(D, (w0 + k) << 5), hi = min(D + C, (w0 + k + 1) << 5);
                const unsigned src = A + (lo - D), len = hi - lo;
                const unsigned long long x = (unsigned long long)s_bits[src >> 5] | (unsigned long long)s_bits[(src >> 5) + 1] << 32;
                const unsigned val = (unsigned)(x >> (src & 31u)) & (len == 32u ? 0xFFFFFFFFu : (1u << len) - 1u);
                if (len == 32u) lraw[w0 + k] = val; else atomicOr(&lraw[w0 + k], val << (lo & 31u));
            }
        }
        __syncthreads();                                                       // s_zc is complete
        int q = -1;
        unsigned A = 0, next_A = 0, zb = 0, zeros = 0;
#pragma unroll
        for (int i = 0; i < kIwtPer; i++) {
            if (!((valid >> i) & 1u)) break;
            const unsigned e = e0 + i;
            if (q < 0 || e >= next_A) {                                            // first value, or the next group begins
                do { q++; next_A = s_pc[(q + 1) << span_shift]; } while (e >= next_A);
                A = s_pc[q << span_shift];
                zeros = s_pc[(q << span_shift) + (1 << (span_shift - 1))] - A;       // values of the group whose bit h is 0
                zb = s_zc[A >> 3] + __popc((unsigned)s_zm[A >> 3] & ((1u << (A & 7u)) - 1u));   // zeros before the group
            }
            const unsigned zp = zexcl + __popc(zmask & ((1u << i) - 1u));
            const unsigned np = ((bits >> i) & 1u) ? A + zeros + ((e - zp) - (A - zb)) : A + (zp - zb);
            nxt[np] = v[i];
        }
        __syncthreads();
        uint32_t* tmp = cur; cur = nxt; nxt = tmp;
    }
    // the block is sorted by digit: every digit's run goes to its place among the values of the parent group with that digit
    for (int i = 0; i < kIwtPer; i++) {
        const int e = i * kIwtThreads + t;
        if (e < cnt) {
            const uint32_t v = cur[e];
            const unsigned b = (v >> low_bit) & dmask;
            out[parent + min(b << low_bit, parent_len) + (s_pb[b + 1] - s_pb[b]) + ((unsigned)e - s_pc[b])] = v;
        }
    }
}

// The low levels: once a group (values sharing v >> (h + 1)) is no longer than kIwtBlock, one CTA keeps its
// block of values in shared memory and runs every remaining level there — bit h of the values to the level's raw
// vector, then the stable split inside each group — instead of three launches per level over global memory.
__global__ void __launch_bounds__(kIwtThreads)
iwt_low_levels_kernel(const uint32_t* __restrict__ vals, int64_t m, int top_h, int levels, uint32_t* __restrict__ raw,
                      const uint64_t* __restrict__ level_raw_word /* level vector (highest bit first) -> first raw word */) {
    extern __shared__ __align__(16) uint32_t s_iwt[];
    uint32_t* cur = s_iwt;
    uint32_t* nxt = s_iwt + kIwtBlock;
    uint32_t* s_zc = s_iwt + 2 * kIwtBlock;              // zeros before each thread's first value
    uint32_t* s_wsum = s_zc + kIwtThreads;
    const int64_t block_start = (int64_t)blockIdx.x * kIwtBlock;
    const int cnt = (int)min((int64_t)kIwtBlock, m - block_start);
    const int t = threadIdx.x;
    for (int i = 0; i < kIwtPer; i++) {
        const int e = i * kIwtThreads + t;
        cur[e] = e < cnt ? vals[block_start + e] : 0xFFFFFFFFu;
    }
    __syncthreads();
    for (int h = top_h; h >= 0; h--) {
        const uint4 q0 = reinterpret_cast<const uint4*>(cur)[2 * t], q1 = reinterpret_cast<const uint4*>(cur)[2 * t + 1];
        const uint32_t v[kIwtPer] = { q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w };
        const int e0 = t * kIwtPer;
        const unsigned valid = e0 >= cnt ? 0u : (e0 + kIwtPer <= cnt ? 0xFFu : (1u << (cnt - e0)) - 1u);
        unsigned bits = 0;
#pragma unroll
        for (int i = 0; i < kIwtPer; i++) bits |= ((v[i] >> h) & 1u) << i;
        bits &= valid;
        const unsigned zmask = ~bits & valid;
        if (valid) reinterpret_cast<uint8_t*>(raw + level_raw_word[levels - 1 - h])[(block_start >> 3) + t] = (uint8_t)bits;
        if (h == 0) break;
        const unsigned z = __popc(zmask);
        const unsigned incl = warp_incl_sum(z);
        if (lane_id() == 31) s_wsum[t >> 5] = incl;
        __syncthreads();
        const unsigned wincl = warp_incl_sum(s_wsum[lane_id()]);                         // every warp scans the warp totals
        const unsigned wbase = __shfl_sync(0xffffffffu, wincl - s_wsum[lane_id()], t >> 5);
        const unsigned zexcl = wbase + incl - z;
        s_zc[t] = zexcl;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < kIwtPer; i++) {
            if (!((valid >> i) & 1u)) break;
            const int e = e0 + i;
            const int bs = (e >> (h + 1)) << (h + 1);
            const unsigned zp = zexcl + __popc(zmask & ((1u << i) - 1u));
            const unsigned zb = (h + 1 >= 3) ? s_zc[bs >> 3] : zexcl + __popc(zmask & ((1u << (bs - e0)) - 1u));
            const int zeros_in_group = min(1 << h, cnt - bs);
            const int np = ((bits >> i) & 1u) ? bs + zeros_in_group + ((e - (int)zp) - (bs - (int)zb)) : bs + (int)(zp - zb);
            nxt[np] = v[i];
        }
        __syncthreads();
        uint32_t* tmp = cur; cur = nxt; nxt = tmp;
    }
}
constexpr size_t kIwtLowSmem = (size_t)(2 * kIwtBlock + kIwtThreads + 32) * 4;

// ---- ranked layout: raw bits -> RankedWTNode bytes -------------------------------------------------------
struct VectorDesc {
    uint64_t raw_word;      // first u32 word of the raw (contiguous) bits; padded with zeros to whole superblocks
    int64_t  len;           // bits
    uint8_t* out;           // where the ranked bytes go (device)
    int64_t  sb_first;      // first entry of this vector in the global superblock arrays
    int64_t  sb_count;
};

// ones per 65536-bit superblock (one warp each)
__global__ void superblock_popcount_kernel(const uint32_t* __restrict__ raw, const VectorDesc* __restrict__ vecs, int nvec,
                                           int64_t sb_begin, int64_t sb_end, uint32_t* __restrict__ sb_ones) {
    const int64_t g = sb_begin + (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (g >= sb_end) return;
    int v = 0;
    while (v + 1 < nvec && vecs[v + 1].sb_first <= g) v++;
    const uint4* p = reinterpret_cast<const uint4*>(raw + vecs[v].raw_word + (uint64_t)(g - vecs[v].sb_first) * 2048u);
    unsigned ones = 0;
#pragma unroll 4
    for (int i = lane_id(); i < 512; i += 32) {
        const uint4 q = p[i];
        ones += __popc(q.x) + __popc(q.y) + __popc(q.z) + __popc(q.w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ones += __shfl_xor_sync(0xffffffffu, ones, o);
    if (lane_id() == 0) sb_ones[g] = ones;
}

// exclusive scan of the superblock counts of every vector separately (one CTA per vector)
__global__ void __launch_bounds__(1024)
superblock_scan_kernel(uint32_t* __restrict__ sb_ones, const VectorDesc* __restrict__ vecs) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const int64_t first = vecs[blockIdx.x].sb_first, len = vecs[blockIdx.x].sb_count;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < len; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const uint32_t v = i < len ? sb_ones[first + i] : 0;
        const uint32_t incl = warp_incl_sum(v);
        if (lane_id() == 31) s_warp[threadIdx.x >> 5] = incl;
        __syncthreads();
        uint32_t b = s_carry;
        for (unsigned w = 0; w < (threadIdx.x >> 5); w++) b += s_warp[w];
        if (i < len) sb_ones[first + i] = b + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = b + incl;
        __syncthreads();
    }
}

// One CTA per superblock, one thread per 512-bit chunk: 64 data bytes, then a uint16 running count inside the
// superblock, or a uint64 absolute count after the 128th chunk — only when more data follows
// (algo/tree/RankedWTNode.java:228-245, layout in SURVEY.md A.3).
__global__ void __launch_bounds__(128)
ranked_layout_kernel(const uint32_t* __restrict__ raw, const VectorDesc* __restrict__ vecs, int nvec, int64_t sb_begin,
                     const uint32_t* __restrict__ sb_excl) {
    __shared__ __align__(16) uint16_t s_out[4232];
    __shared__ uint32_t s_warp[4];
    const int64_t g = sb_begin + blockIdx.x;
    int v = 0;
    while (v + 1 < nvec && vecs[v + 1].sb_first <= g) v++;
    const VectorDesc vd = vecs[v];
    const int64_t sb = g - vd.sb_first;
    const int c = threadIdx.x;
    const uint4* p = reinterpret_cast<const uint4*>(raw + vd.raw_word + (uint64_t)sb * 2048u + (uint64_t)c * 16u);
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint4 q = p[i];
        w[4 * i] = q.x; w[4 * i + 1] = q.y; w[4 * i + 2] = q.z; w[4 * i + 3] = q.w;
    }
    unsigned ones = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) ones += __popc(w[i]);
    const unsigned incl_w = warp_incl_sum(ones);
    if (lane_id() == 31) s_warp[c >> 5] = incl_w;
    __syncthreads();
    unsigned incl = incl_w;
    for (int k = 0; k < (c >> 5); k++) incl += s_warp[k];

    uint16_t* o = s_out + c * 33;
#pragma unroll
    for (int i = 0; i < 16; i++) { o[2 * i] = (uint16_t)w[i]; o[2 * i + 1] = (uint16_t)(w[i] >> 16); }
    const int64_t chunk = sb * 128 + c;                       // chunk index inside the vector
    const bool counter_follows = (chunk + 1) * 512 < vd.len;
    if (counter_follows) {
        if (c < 127) {
            o[32] = (uint16_t)incl;
        } else {
            const unsigned long long abs_ones = (unsigned long long)sb_excl[g] + incl;
            o[32] = (uint16_t)abs_ones; o[33] = (uint16_t)(abs_ones >> 16);
            o[34] = (uint16_t)(abs_ones >> 32); o[35] = (uint16_t)(abs_ones >> 48);
        }
    }
    __syncthreads();
    const int64_t total = (int64_t)(((uint64_t)(vd.len - 1) >> 16) * 6 + ((uint64_t)(vd.len - 1) >> 9) * 2 + ((uint64_t)(vd.len + 7) >> 3));
    const int64_t start = sb * 8454;
    const int nbytes = (int)min((int64_t)8454, total - start);
    const uint8_t* sbytes = reinterpret_cast<const uint8_t*>(s_out);
    uint8_t* dst = vd.out + start;
    for (int i = threadIdx.x; i < nbytes; i += 128) dst[i] = sbytes[i];
}

inline int64_t superblocks(int64_t len) { return (len + 65535) >> 16; }

// counters + final byte layout of vectors [v0, v1) of `vecs` (their superblocks are [sb0, sb1))
int layout_vectors(DeviceCtx* ctx, cudaStream_t st, const uint32_t* d_raw, const VectorDesc* d_vecs, int nvec, int v0, int v1,
                   int64_t sb0, int64_t sb1, uint32_t* d_sb) {
    if (v1 <= v0 || sb1 <= sb0) return GCZ_OK;
    GCZ_LAUNCH(ctx, superblock_popcount_kernel, (unsigned)(((sb1 - sb0) * 32 + 255) / 256), 256, 0, st, d_raw, d_vecs, nvec, sb0, sb1, d_sb);
    GCZ_LAUNCH(ctx, superblock_scan_kernel, (unsigned)(v1 - v0), 1024, 0, st, d_sb, d_vecs + v0);
    GCZ_LAUNCH(ctx, ranked_layout_kernel, (unsigned)(sb1 - sb0), 128, 0, st, d_raw, d_vecs, nvec, sb0, d_sb);
    return GCZ_OK;
}

// All levels of an IndexWaveletTree over the m values in d_ssa[0] (d_ssa[1]: scratch of the same size): the high levels in
// passes of up to kTopBits levels (count, scan, emit), the low ones in one launch (iwt_low_levels_kernel).
inline int64_t iwt_blocks(int64_t m) { return (m + kIwtBlock - 1) / kIwtBlock; }
inline size_t iwt_scratch_bytes(int64_t m) {
    return (size_t)kTopBins * (size_t)iwt_blocks(m) * 4 + 256 + row_scan_scratch_bytes(kTopBins, iwt_blocks(m)) + 256;
}

int iwt_levels(DeviceCtx* ctx, cudaStream_t st, uint32_t* const d_ssa[2], int64_t m, int levels, uint32_t* d_raw,
               const std::vector<VectorDesc>& vecs, int level_vec0, void* d_scratch, uint64_t* d_level_raw) {
    std::vector<uint64_t> h_level_raw((size_t)levels);
    for (int l = 0; l < levels; l++) h_level_raw[(size_t)l] = vecs[(size_t)(level_vec0 + l)].raw_word;
    GCZ_TRY(small_upload(ctx, st, d_level_raw, h_level_raw.data(), sizeof(uint64_t) * levels));
    const int64_t blocks = iwt_blocks(m);
    uint32_t* d_counts = static_cast<uint32_t*>(d_scratch);
    void* d_scan = reinterpret_cast<char*>(d_scratch) + (((size_t)kTopBins * (size_t)blocks * 4 + 255) & ~(size_t)255);
    if (!ctx->iwt_attr) {
        GCZ_CUDA(cudaFuncSetAttribute(iwt_low_levels_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kIwtLowSmem));
        GCZ_CUDA(cudaFuncSetAttribute(iwt_top_emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kIwtTopSmem));
        ctx->iwt_attr = true;
    }
    int cur = 0;
    int H = levels - 1;                                  // highest bit not yet done
    while (H >= kIwtLowBits) {
        const int L = std::min(kTopBits, H - kIwtLowBits + 1);
        GCZ_LAUNCH(ctx, iwt_top_count_kernel, (unsigned)blocks, kIwtThreads, 0, st, d_ssa[cur], m, H - L + 1, L, d_counts, blocks);
        GCZ_TRY(row_scan(ctx, st, d_counts, 1 << L, blocks, false, d_scan, nullptr, nullptr));
        GCZ_LAUNCH(ctx, iwt_top_emit_kernel, (unsigned)blocks, kIwtThreads, kIwtTopSmem, st, d_ssa[cur], d_ssa[cur ^ 1], m, H, L, levels,
                   d_raw, d_level_raw, d_counts, blocks);
        cur ^= 1;
        H -= L;
    }
    GCZ_LAUNCH(ctx, iwt_low_levels_kernel, (unsigned)blocks, kIwtThreads, kIwtLowSmem, st, d_ssa[cur], m, H, levels, d_raw, d_level_raw);
    return GCZ_OK;
}

}  // namespace

size_t wavelet_workspace_bytes(int64_t n, int sampling_factor) {
    const int64_t m = (n + ((int64_t)1 << sampling_factor) - 1) >> sampling_factor;
    const int levels = 64 - __builtin_clzll((uint64_t)std::max<int64_t>(m, 1));
    const size_t tiles = (size_t)((n + kWarpTile - 1) / kWarpTile);
    size_t raw_words = (size_t)superblocks(n) * 2048 * 15 + 256 * 2048;   // nodes: <= 15 levels of n bits + padding
    raw_words += (size_t)superblocks(n) * 2048;                   // marker
    raw_words += (size_t)superblocks(m) * 2048 * levels;          // IWT levels
    return raw_words * 4 + tiles * 4 * 258 + (size_t)m * 8 + (size_t)((m + 31) / 32) * 4 + row_scan_scratch_bytes(257, (int64_t)tiles) + (4 << 20);
}

int build_wavelet_structures(DeviceCtx* ctx, cudaStream_t st, const uint8_t* d_text, uint32_t* d_sa, int carry_shift, bool clean_sa,
                             int64_t n, const gcz_shape* shape, int sampling_factor, uint8_t* d_bwt,
                             uint8_t* d_gcz_body, uint8_t* d_gcx_body, Arena& arena, WaveletStats* stats,
                             uint8_t* h_gcz_out, cudaStream_t copy_stream, cudaEvent_t gcz_copied) {
    const size_t mark0 = arena.mark();
    // ---- host tables -------------------------------------------------------------------------------
    SymbolTables h_tab;
    std::memset(&h_tab, 0, sizeof(h_tab));
    std::memset(h_tab.dense, 0xFF, sizeof(h_tab.dense));
    int sigma = 0, max_len = 0;
    for (int c = 0; c < 256; c++) {
        if (shape->bit_lengths[c] > 0) {
            h_tab.dense[c] = (uint8_t)sigma;
            h_tab.byte_of[sigma] = (uint8_t)c;
            h_tab.code[sigma] = (uint16_t)shape->codes[c];
            h_tab.len[sigma] = (uint8_t)shape->bit_lengths[c];
            max_len = std::max(max_len, (int)shape->bit_lengths[c]);
            sigma++;
        }
    }
    const int n_nodes = shape->n_nodes;
    if (n_nodes <= 0 || n_nodes > 255 || max_len > 15) return fail(GCZ_E_RANGE, "unsupported tree shape");
    for (int v = 0; v < n_nodes; v++) {
        h_tab.node_prefix[v] = (uint16_t)shape->node_prefix[v];
        h_tab.node_depth[v] = (uint8_t)shape->node_depth[v];
    }
    for (int s = 0; s < sigma; s++) {
        for (int d = 0; d < h_tab.len[s]; d++) {
            int found = -1;
            for (int v = 0; v < n_nodes; v++) {
                if (h_tab.node_depth[v] == d && h_tab.node_prefix[v] == (h_tab.code[s] & ((1u << d) - 1))) { found = v; break; }
            }
            if (found < 0) return fail(GCZ_E_INTERNAL, "symbol path leaves the tree");
            h_tab.node_of[s][d] = (uint8_t)found;
            if (found < 32) {
                h_tab.path_mask[s] |= 1u << found;
                if ((h_tab.code[s] >> d) & 1u) h_tab.bit_mask[s] |= 1u << found;
            }
        }
    }
    h_tab.sigma = sigma; h_tab.n_nodes = n_nodes; h_tab.max_len = max_len;

    // ---- vectors: HSWT nodes, marker, IWT levels ------------------------------------------------------
    const int64_t m = (n + ((int64_t)1 << sampling_factor) - 1) >> sampling_factor;
    const int levels = 64 - __builtin_clzll((uint64_t)m);
    std::vector<VectorDesc> vecs;
    uint64_t raw_words = 0;
    int64_t total_sb = 0;
    auto add_vec = [&](int64_t len, uint8_t* out) {
        VectorDesc d;
        d.raw_word = raw_words; d.len = len; d.out = out; d.sb_first = total_sb; d.sb_count = superblocks(len);
        raw_words += (uint64_t)d.sb_count * 2048u;
        total_sb += d.sb_count;
        vecs.push_back(d);
    };
    std::vector<uint64_t> h_node_raw(n_nodes);
    for (int v = 0; v < n_nodes; v++) {
        if (shape->node_bits[v] <= 0) return fail(GCZ_E_ARG, "shape has an empty node (was it built from counts?)");
        h_node_raw[v] = raw_words;
        add_vec(shape->node_bits[v], d_gcz_body + shape->node_offset[v]);
    }
    const int marker_vec = (int)vecs.size();
    add_vec(n, d_gcx_body);
    const int64_t rank_bytes = ranked_bytes(n), level_bytes = ranked_bytes(m);
    const int level_vec0 = (int)vecs.size();
    for (int l = 0; l < levels; l++) add_vec(m, d_gcx_body + rank_bytes + (int64_t)l * level_bytes);   // highest bit first
    if (raw_words >= (1ull << 32)) return fail(GCZ_E_RANGE, "raw bit area above 2^32 words");

    const int64_t tiles = (n + kWarpTile - 1) / kWarpTile;
    SymbolTables* d_tab = arena.get<SymbolTables>(1);
    uint32_t* d_raw = arena.get<uint32_t>((size_t)raw_words + 16);
    uint32_t* d_tile_counts = arena.get<uint32_t>((size_t)tiles * (sigma + 1));
    uint64_t* d_node_raw = arena.get<uint64_t>((size_t)n_nodes);
    VectorDesc* d_vecs = arena.get<VectorDesc>(vecs.size());
    uint32_t* d_sb = arena.get<uint32_t>((size_t)total_sb + 1);
    uint32_t* d_ssa[2] = { arena.get<uint32_t>((size_t)m), arena.get<uint32_t>((size_t)m) };
    void* d_iwt_scratch = arena.raw(iwt_scratch_bytes(m));
    uint64_t* d_level_raw = arena.get<uint64_t>(64);
    void* d_scan = arena.raw(row_scan_scratch_bytes(sigma + 1, tiles));
    if (!d_iwt_scratch || !d_level_raw || !d_scan) return fail(GCZ_E_NOMEM, "wavelet workspace for n=%lld", (long long)n);
    if (!d_tab || !d_raw || !d_tile_counts || !d_node_raw || !d_vecs || !d_sb || !d_ssa[0] || !d_ssa[1])
        return fail(GCZ_E_NOMEM, "wavelet workspace for n=%lld", (long long)n);

    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
    if (stats) {
        GCZ_CUDA(cudaEventCreate(&ev0)); GCZ_CUDA(cudaEventCreate(&ev1)); GCZ_CUDA(cudaEventCreate(&ev2));
        GCZ_CUDA(cudaEventRecord(ev0, st));
    }
    GCZ_TRY(small_upload(ctx, st, d_tab, &h_tab, sizeof(h_tab)));
    GCZ_TRY(small_upload(ctx, st, d_node_raw, h_node_raw.data(), sizeof(uint64_t) * n_nodes));
    GCZ_TRY(small_upload(ctx, st, d_vecs, vecs.data(), sizeof(VectorDesc) * vecs.size()));
    GCZ_CUDA(cudaMemsetAsync(d_raw, 0, ((size_t)raw_words + 16) * 4, st));

    // table-driven node emission for alphabets of up to 8 symbols (every DNA block): 1.6 ms -> 0.55 ms on the chr1-shaped block
    // (profiles/variants_r02.md); larger alphabets take the ballot kernels below
    const bool emit_lut = sigma <= 8 && n_nodes <= kLutNodes;
    std::vector<uint2> h_lut;
    uint2* d_lut = nullptr;
    if (emit_lut) {
        const int entries = sigma * sigma * sigma * sigma;
        h_lut.assign((size_t)entries, make_uint2(0u, 0u));
        for (int g = 0; g < entries; g++) {
            const int d[4] = { g / (sigma * sigma * sigma), (g / (sigma * sigma)) % sigma, (g / sigma) % sigma, g % sigma };   // text order
            uint8_t bytes[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
            for (int v = 0; v < n_nodes; v++) {
                unsigned p = 0, c = 0;
                for (int i = 0; i < 4; i++) {
                    if ((h_tab.path_mask[d[i]] >> v) & 1u) { p |= ((h_tab.bit_mask[d[i]] >> v) & 1u) << c; c++; }
                }
                bytes[v] = (uint8_t)(p | (c << 4));
            }
            std::memcpy(&h_lut[(size_t)g], bytes, 8);
        }
        d_lut = arena.get<uint2>((size_t)entries);
        if (!d_lut) return fail(GCZ_E_NOMEM, "wavelet workspace for n=%lld", (long long)n);
        GCZ_TRY(small_upload(ctx, st, d_lut, h_lut.data(), sizeof(uint2) * (size_t)entries));
    }

    // shape table at the head of the body
    {
        std::vector<uint8_t> tbl((size_t)shape->table_bytes + 8);
        const int64_t w = shape_write(shape, tbl.data(), (int64_t)tbl.size());
        if (w != shape->table_bytes) return w < 0 ? (int)w : fail(GCZ_E_INTERNAL, "shape table size changed");
        GCZ_TRY(small_upload(ctx, st, d_gcz_body, tbl.data(), (size_t)w));     // (the host bytes are taken at once: no wait for the stream)
    }

    // ---- BWT, counts, marker bits ------------------------------------------------------------------------
    const unsigned wt_grid = (unsigned)((tiles + kWtThreads / 32 - 1) / (kWtThreads / 32));
    const uint32_t sample_mask = (1u << sampling_factor) - 1u;
    uint32_t* d_marker_raw = d_raw + vecs[marker_vec].raw_word;
    if (sigma <= 8) {                      // packed per-lane counters (0.95 ms -> 0.47 ms on the chr1-shaped block)
        GCZ_LAUNCH(ctx, bwt_count_kernel<true>, wt_grid, kWtThreads, 0, st, d_text, d_sa, n, d_tab, sample_mask, carry_shift,
                   (carry_shift && clean_sa) ? d_sa : (uint32_t*)nullptr, d_bwt, d_marker_raw, d_tile_counts, tiles);
    } else {
        GCZ_LAUNCH(ctx, bwt_count_kernel<false>, wt_grid, kWtThreads, 0, st, d_text, d_sa, n, d_tab, sample_mask, carry_shift,
                   (carry_shift && clean_sa) ? d_sa : (uint32_t*)nullptr, d_bwt, d_marker_raw, d_tile_counts, tiles);
    }
    GCZ_TRY(row_scan(ctx, st, d_tile_counts, sigma + 1, tiles, false, d_scan, nullptr, nullptr));
    if (emit_lut) {
        const int entries = (int)h_lut.size();
        const size_t smem = sizeof(uint2) * (size_t)entries + (size_t)(kWtThreads / 32) * n_nodes * kLutStageWords * 4;
        const unsigned grid = (unsigned)std::min<int64_t>((int64_t)wt_grid, (int64_t)ctx->sm_count * 8);
        switch (n_nodes) {
#define GCZ_EMIT_LUT(N) case N: GCZ_LAUNCH(ctx, hswt_emit_lut_kernel<N>, grid, kWtThreads, smem, st, d_bwt, n, d_tab, d_lut, entries, \
                                           d_tile_counts, tiles, d_node_raw, d_raw); break;
            GCZ_EMIT_LUT(1) GCZ_EMIT_LUT(2) GCZ_EMIT_LUT(3) GCZ_EMIT_LUT(4) GCZ_EMIT_LUT(5) GCZ_EMIT_LUT(6) GCZ_EMIT_LUT(7) GCZ_EMIT_LUT(8)
#undef GCZ_EMIT_LUT
        }
    } else if (n_nodes <= 32) {
        GCZ_LAUNCH(ctx, hswt_emit_small_kernel, wt_grid, kWtThreads, 0, st, d_bwt, n, d_tab, d_tile_counts, tiles, d_node_raw, d_raw);
    } else {
        GCZ_LAUNCH(ctx, hswt_emit_kernel, wt_grid, kWtThreads, 0, st, d_bwt, n, d_tab, d_tile_counts, tiles, d_node_raw, d_raw);
    }
    // the .gcz body is complete once its nodes are laid out: its copy to the host overlaps the index build
    const int64_t sb_nodes = vecs[marker_vec].sb_first;
    GCZ_TRY(layout_vectors(ctx, st, d_raw, d_vecs, (int)vecs.size(), 0, marker_vec, 0, sb_nodes, d_sb));
    if (h_gcz_out) {
        GCZ_CUDA(cudaEventRecord(gcz_copied, st));
        GCZ_CUDA(cudaStreamWaitEvent(copy_stream, gcz_copied, 0));
        GCZ_CUDA(cudaMemcpyAsync(h_gcz_out, d_gcz_body, (size_t)shape->size, cudaMemcpyDeviceToHost, copy_stream));
        GCZ_CUDA(cudaEventRecord(gcz_copied, copy_stream));
    }
    const int tail_vec0 = marker_vec;
    const int64_t tail_sb0 = sb_nodes;
    if (stats) GCZ_CUDA(cudaEventRecord(ev1, st));

    // ---- sampled SA + IndexWaveletTree ---------------------------------------------------------------------
    GCZ_LAUNCH(ctx, sample_kernel, wt_grid, kWtThreads, 0, st, d_sa, n,
               carry_shift ? (1u << carry_shift) - 1u : 0xFFFFFFFFu, sample_mask, sampling_factor,
               d_tile_counts + (size_t)sigma * tiles, tiles, d_ssa[0]);
    GCZ_TRY(iwt_levels(ctx, st, d_ssa, m, levels, d_raw, vecs, level_vec0, d_iwt_scratch, d_level_raw));

    // ---- counters + final byte layout of every vector ---------------------------------------------------------
    GCZ_TRY(layout_vectors(ctx, st, d_raw, d_vecs, (int)vecs.size(), tail_vec0, (int)vecs.size(), tail_sb0, total_sb, d_sb));

    if (stats) {
        GCZ_CUDA(cudaEventRecord(ev2, st));
        GCZ_CUDA(cudaEventSynchronize(ev2));
        cudaEventElapsedTime(&stats->bwt_hswt_ms, ev0, ev1);
        cudaEventElapsedTime(&stats->ssa_ms, ev1, ev2);
        cudaEventDestroy(ev0); cudaEventDestroy(ev1); cudaEventDestroy(ev2);
    } else {
        GCZ_CUDA(cudaStreamSynchronize(st));
    }
    arena.release(mark0);
    return GCZ_OK;
}

// Stage hooks for the parity tests ---------------------------------------------------------------------------

namespace {
__global__ void pack_bits_kernel(const uint8_t* __restrict__ bits, int64_t len, uint32_t* __restrict__ raw) {
    const int64_t words = (len + 31) >> 5;
    const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t w = gw; w < words; w += nwarps) {
        const int64_t p = w * 32 + lane_id();
        const unsigned word = __ballot_sync(0xffffffffu, p < len && (bits[p] & 1));
        if (lane_id() == 0) raw[w] = word;
    }
}
}  // namespace

int ranked_vector_from_bits(DeviceCtx* ctx, cudaStream_t st, const uint8_t* d_bits, int64_t len, uint8_t* d_out, Arena& arena) {
    const size_t mark0 = arena.mark();
    VectorDesc d;
    d.raw_word = 0; d.len = len; d.out = d_out; d.sb_first = 0; d.sb_count = superblocks(len);
    uint32_t* d_raw = arena.get<uint32_t>((size_t)d.sb_count * 2048 + 16);
    VectorDesc* d_vec = arena.get<VectorDesc>(1);
    uint32_t* d_sb = arena.get<uint32_t>((size_t)d.sb_count + 1);
    if (!d_raw || !d_vec || !d_sb) return fail(GCZ_E_NOMEM, "ranked vector workspace");
    GCZ_CUDA(cudaMemsetAsync(d_raw, 0, ((size_t)d.sb_count * 2048 + 16) * 4, st));
    GCZ_CUDA(cudaMemcpyAsync(d_vec, &d, sizeof(d), cudaMemcpyHostToDevice, st));
    GCZ_CUDA(cudaStreamSynchronize(st));
    GCZ_LAUNCH(ctx, pack_bits_kernel, ctx->sm_count * 4, 256, 0, st, d_bits, len, d_raw);
    GCZ_TRY(layout_vectors(ctx, st, d_raw, d_vec, 1, 0, 1, 0, d.sb_count, d_sb));
    GCZ_CUDA(cudaStreamSynchronize(st));
    arena.release(mark0);
    return GCZ_OK;
}

int index_wavelet_tree_from_values(DeviceCtx* ctx, cudaStream_t st, const uint32_t* d_vals, int64_t m, uint8_t* d_out, Arena& arena) {
    const size_t mark0 = arena.mark();
    const int levels = 64 - __builtin_clzll((uint64_t)m);
    const int64_t level_bytes = ranked_bytes(m);
    std::vector<VectorDesc> vecs;
    uint64_t raw_words = 0; int64_t total_sb = 0;
    for (int l = 0; l < levels; l++) {
        VectorDesc d;
        d.raw_word = raw_words; d.len = m; d.out = d_out + (int64_t)l * level_bytes; d.sb_first = total_sb; d.sb_count = superblocks(m);
        raw_words += (uint64_t)d.sb_count * 2048u; total_sb += d.sb_count;
        vecs.push_back(d);
    }
    uint32_t* d_raw = arena.get<uint32_t>((size_t)raw_words + 16);
    VectorDesc* d_vecs = arena.get<VectorDesc>(vecs.size());
    uint32_t* d_sb = arena.get<uint32_t>((size_t)total_sb + 1);
    uint32_t* d_ssa[2] = { arena.get<uint32_t>((size_t)m), arena.get<uint32_t>((size_t)m) };
    void* d_iwt_scratch = arena.raw(iwt_scratch_bytes(m));
    uint64_t* d_level_raw = arena.get<uint64_t>(64);
    if (!d_raw || !d_vecs || !d_sb || !d_ssa[0] || !d_ssa[1] || !d_iwt_scratch || !d_level_raw) return fail(GCZ_E_NOMEM, "IWT workspace");
    GCZ_CUDA(cudaMemsetAsync(d_raw, 0, ((size_t)raw_words + 16) * 4, st));
    GCZ_CUDA(cudaMemcpyAsync(d_vecs, vecs.data(), sizeof(VectorDesc) * vecs.size(), cudaMemcpyHostToDevice, st));
    GCZ_CUDA(cudaMemcpyAsync(d_ssa[0], d_vals, (size_t)m * 4, cudaMemcpyDeviceToDevice, st));
    GCZ_CUDA(cudaStreamSynchronize(st));
    GCZ_TRY(iwt_levels(ctx, st, d_ssa, m, levels, d_raw, vecs, 0, d_iwt_scratch, d_level_raw));
    GCZ_TRY(layout_vectors(ctx, st, d_raw, d_vecs, (int)vecs.size(), 0, (int)vecs.size(), 0, total_sb, d_sb));
    GCZ_CUDA(cudaStreamSynchronize(st));
    arena.release(mark0);
    return GCZ_OK;
}

}  // namespace gcz

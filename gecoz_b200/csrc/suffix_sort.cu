// GPU suffix sorter: the data-parallel replacement of SAIS.suffix(ByteBuffer, int[])
// (algo/string/SAIS.java:103-137).  Only the RESULT is shared with the reference — the suffix array of the
// block text under unsigned byte order with "a proper prefix sorts first" (SURVEY.md B.1); the reference's
// induced sorting is sequential and is not followed.
//
// Algorithm (prefix doubling on packed keys, with sorted-group elimination):
//   1. symbols are remapped to dense codes 1..sigma (0 = past the end) of `bits` bits; the first
//      k = 64 / bits symbols of every suffix are packed into one 64-bit key (k = 21 for ACGTN + '\0');
//   2. one full radix sort of (key, position) orders all suffixes by their first k symbols;
//   3. suffixes that are alone in their key group are final.  The others are kept in a compact list
//      (slot position in SA, suffix, group ordinal) and refined: round r sorts the list by
//      (group ordinal, rank[suffix + h]) with h = k * 2^r, splits the groups, updates rank[] (= group
//      start slot) and drops the suffixes that became unique.  Only the list is touched after step 2.
#include "suffix_sort.cuh"

#include <algorithm>

namespace gcz {

namespace {

// ---- 1. key packing -------------------------------------------------------------------------------
constexpr int kPackThreads = 256;
constexpr int kPackItems = 8;
constexpr int kPackTile = kPackThreads * kPackItems;

__global__ void __launch_bounds__(kPackThreads)
pack_keys_kernel(const uint8_t* __restrict__ text, int64_t n, const uint8_t* __restrict__ code_of,
                 int bits, int k, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    __shared__ uint8_t s_code_of[256];
    __shared__ uint8_t s_codes[kPackTile + 72];
    __shared__ uint64_t s_keys[kPackTile];
    const int64_t base = (int64_t)blockIdx.x * kPackTile;
    s_code_of[threadIdx.x] = code_of[threadIdx.x];
    __syncthreads();
    for (int i = threadIdx.x; i < kPackTile + k; i += kPackThreads) {
        const int64_t p = base + i;
        s_codes[i] = p < n ? s_code_of[text[p]] : 0;          // 0 = past the end, below every symbol
    }
    __syncthreads();
    // every thread slides a k-symbol window over kPackItems consecutive positions
    const int first = threadIdx.x * kPackItems;
    const uint64_t mask = (k * bits == 64) ? ~0ull : ((1ull << (k * bits)) - 1);
    uint64_t key = 0;
    for (int j = 0; j < k - 1; j++) key = (key << bits) | s_codes[first + j];
#pragma unroll
    for (int i = 0; i < kPackItems; i++) {
        key = ((key << bits) | s_codes[first + i + k - 1]) & mask;
        s_keys[first + i] = key;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kPackTile; i += kPackThreads) {
        const int64_t p = base + i;
        if (p < n) { keys[p] = s_keys[i]; vals[p] = (uint32_t)p; }
    }
}

// ---- 3. group bookkeeping ---------------------------------------------------------------------------
constexpr int kGrpThreads = 256;
constexpr int kGrpItems = 8;
constexpr int kGrpTile = kGrpThreads * kGrpItems;

struct SlotFlags {
    unsigned valid, boundary, single;      // bit i = slot t0 + i
};

// boundary[t] = first slot of a key group; single[t] = the group has exactly one member
__device__ __forceinline__ SlotFlags slot_flags(const uint64_t* __restrict__ keys, int64_t m, int64_t t0) {
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
    return f;
}

// pass A: per-tile aggregates {last boundary slot, kept (non-single) slots, kept groups}
__global__ void __launch_bounds__(kGrpThreads)
group_aggregate_kernel(const uint64_t* __restrict__ keys, int64_t m, long long* __restrict__ agg_last,
                       long long* __restrict__ agg_keep, long long* __restrict__ agg_groups) {
    __shared__ long long s_last[kGrpThreads / 32];
    __shared__ unsigned s_keep[kGrpThreads / 32], s_groups[kGrpThreads / 32];
    const int64_t t0 = (int64_t)blockIdx.x * kGrpTile + (int64_t)threadIdx.x * kGrpItems;
    const SlotFlags f = slot_flags(keys, m, t0);
    long long last = f.boundary ? t0 + (31 - __clz(f.boundary)) : -1;
    unsigned keep = __popc(f.valid & ~f.single);
    unsigned groups = __popc(f.boundary & ~f.single);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
        keep += __shfl_xor_sync(0xffffffffu, keep, o);
        groups += __shfl_xor_sync(0xffffffffu, groups, o);
    }
    if (lane_id() == 0) { s_last[threadIdx.x >> 5] = last; s_keep[threadIdx.x >> 5] = keep; s_groups[threadIdx.x >> 5] = groups; }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long l = -1, kk = 0, g = 0;
        for (int w = 0; w < kGrpThreads / 32; w++) { l = max(l, s_last[w]); kk += s_keep[w]; g += s_groups[w]; }
        agg_last[blockIdx.x] = l; agg_keep[blockIdx.x] = kk; agg_groups[blockIdx.x] = g;
    }
}

// single-CTA exclusive scans over the tile aggregates (max for `last`, sum for the others), in place;
// totals[0] = kept slots, totals[1] = kept groups
__global__ void __launch_bounds__(1024)
group_scan_kernel(long long* __restrict__ agg_last, long long* __restrict__ agg_keep, long long* __restrict__ agg_groups,
                  int64_t tiles, long long* __restrict__ totals) {
    __shared__ long long s_l[32], s_k[32], s_g[32];
    __shared__ long long s_carry[3];
    if (threadIdx.x == 0) { s_carry[0] = -1; s_carry[1] = 0; s_carry[2] = 0; }
    __syncthreads();
    for (int64_t base = 0; base < tiles; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const long long l = i < tiles ? agg_last[i] : -1;
        const long long k = i < tiles ? agg_keep[i] : 0;
        const long long g = i < tiles ? agg_groups[i] : 0;
        long long il = l, ik = k, ig = g;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long tl = __shfl_up_sync(0xffffffffu, il, o);
            const long long tk = __shfl_up_sync(0xffffffffu, ik, o);
            const long long tg = __shfl_up_sync(0xffffffffu, ig, o);
            if (lane_id() >= (unsigned)o) { il = max(il, tl); ik += tk; ig += tg; }
        }
        if (lane_id() == 31) { s_l[threadIdx.x >> 5] = il; s_k[threadIdx.x >> 5] = ik; s_g[threadIdx.x >> 5] = ig; }
        __syncthreads();
        long long bl = s_carry[0], bk = s_carry[1], bg = s_carry[2];
        for (unsigned w = 0; w < (threadIdx.x >> 5); w++) { bl = max(bl, s_l[w]); bk += s_k[w]; bg += s_g[w]; }
        // exclusive results
        const long long el = max(bl, __shfl_up_sync(0xffffffffu, il, 1));
        const long long ek = bk + ik - k, eg = bg + ig - g;
        if (i < tiles) {
            agg_last[i] = lane_id() == 0 ? bl : el;
            agg_keep[i] = ek;
            agg_groups[i] = eg;
        }
        __syncthreads();
        if (threadIdx.x == 1023) { s_carry[0] = max(bl, il); s_carry[1] = bk + ik; s_carry[2] = bg + ig; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { totals[0] = s_carry[1]; totals[1] = s_carry[2]; }
}

// pass B: ranks, finished suffixes, next list
template <bool INITIAL>
__global__ void __launch_bounds__(kGrpThreads)
group_apply_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ suf, const uint32_t* __restrict__ pos,
                   int64_t m, const long long* __restrict__ pre_last, const long long* __restrict__ pre_keep,
                   const long long* __restrict__ pre_groups, uint32_t* __restrict__ rank, uint32_t* __restrict__ sa,
                   uint32_t* __restrict__ pos_out, uint32_t* __restrict__ suf_out, uint32_t* __restrict__ gid_out) {
    __shared__ long long s_last[kGrpThreads / 32];
    __shared__ unsigned s_keep[kGrpThreads / 32], s_groups[kGrpThreads / 32];
    const int64_t t0 = (int64_t)blockIdx.x * kGrpTile + (int64_t)threadIdx.x * kGrpItems;
    const SlotFlags f = slot_flags(keys, m, t0);
    const long long my_last = f.boundary ? t0 + (31 - __clz(f.boundary)) : -1;
    const unsigned my_keep = __popc(f.valid & ~f.single);
    const unsigned my_groups = __popc(f.boundary & ~f.single);
    // block-wide exclusive scans of the per-thread aggregates
    long long il = warp_incl_max(my_last);
    unsigned ik = warp_incl_sum(my_keep), ig = warp_incl_sum(my_groups);
    if (lane_id() == 31) { s_last[threadIdx.x >> 5] = il; s_keep[threadIdx.x >> 5] = ik; s_groups[threadIdx.x >> 5] = ig; }
    __syncthreads();
    long long last = pre_last[blockIdx.x];
    long long keep = pre_keep[blockIdx.x], groups = pre_groups[blockIdx.x];
    for (unsigned w = 0; w < (threadIdx.x >> 5); w++) { last = max(last, s_last[w]); keep += s_keep[w]; groups += s_groups[w]; }
    const long long prev_l = __shfl_up_sync(0xffffffffu, il, 1);
    if (lane_id() > 0) last = max(last, prev_l);
    keep += ik - my_keep;
    groups += ig - my_groups;

#pragma unroll
    for (int i = 0; i < kGrpItems; i++) {
        if (!((f.valid >> i) & 1)) break;
        const int64_t t = t0 + i;
        if ((f.boundary >> i) & 1) last = t;
        const uint32_t s = suf[t];
        const uint32_t start = INITIAL ? (uint32_t)last : pos[last];
        rank[s] = start;
        if ((f.single >> i) & 1) {
            if (!INITIAL) sa[pos[t]] = s;
        } else {
            if ((f.boundary >> i) & 1) groups++;
            pos_out[keep] = INITIAL ? (uint32_t)t : pos[t];
            suf_out[keep] = s;
            gid_out[keep] = (uint32_t)(groups - 1);
            keep++;
        }
    }
}

// refinement key: (group ordinal, rank of the suffix h symbols further, 0 when that is past the end)
__global__ void refine_keys_kernel(const uint32_t* __restrict__ suf, const uint32_t* __restrict__ gid, int64_t m,
                                   const uint32_t* __restrict__ rank, int64_t n, int64_t h, int low_bits,
                                   uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < m; u += stride) {
        const uint32_t s = suf[u];
        const int64_t q = (int64_t)s + h;
        const uint64_t low = q < n ? (uint64_t)rank[q] + 1 : 0;
        keys[u] = ((uint64_t)gid[u] << low_bits) | low;
        vals[u] = s;
    }
}

inline int bits_for(uint64_t max_value) {      // number of bits needed to represent values 0..max_value
    int b = 0;
    while (b < 64 && (max_value >> b) != 0) b++;
    return b == 0 ? 1 : b;
}

}  // namespace

size_t suffix_sort_workspace_bytes(int64_t n) {
    // rank 4n + keys 16n + vals(other) 4n + refinement worst case 28n + sort temp + aggregates
    const size_t tiles = (size_t)(n / kGrpTile + 2);
    return (size_t)n * (4 + 16 + 4 + 28) + radix_sort_temp_bytes(n) + tiles * 3 * 8 + (1 << 20);
}

int suffix_sort(DeviceCtx* ctx, cudaStream_t st, const uint8_t* d_text, int64_t n, const int64_t counts[256],
                uint32_t* d_sa, Arena& arena, SuffixSortStats* stats) {
    if (n <= 0 || n > 0x7FFFFFFFll) return fail(GCZ_E_RANGE, "block of %lld symbols", (long long)n);

    // dense symbol codes: 1..sigma in byte order, 0 reserved for "past the end"
    uint8_t h_code[256];
    int sigma = 0;
    for (int c = 0; c < 256; c++) { if (counts[c] > 0) ++sigma; h_code[c] = counts[c] > 0 ? (uint8_t)sigma : 0; }
    // 256 distinct byte values plus the end marker would need 9-bit codes; FASTA text never gets there
    if (sigma > 255) return fail(GCZ_E_RANGE, "all 256 byte values present: not supported by the key packer");
    const int bits = bits_for((uint64_t)sigma);
    const int k = 64 / bits;
    const int key_bits = k * bits;

    const size_t mark0 = arena.mark();
    uint8_t* d_code = arena.get<uint8_t>(256);
    uint32_t* d_rank = arena.get<uint32_t>((size_t)n);
    uint64_t* d_keys0 = arena.get<uint64_t>((size_t)n);
    uint64_t* d_keys1 = arena.get<uint64_t>((size_t)n);
    uint32_t* d_vals1 = arena.get<uint32_t>((size_t)n);
    void* d_temp = arena.raw(radix_sort_temp_bytes(n));
    const int64_t tiles_n = (n + kGrpTile - 1) / kGrpTile;
    long long* d_agg = arena.get<long long>((size_t)tiles_n * 3 + 8);
    long long* d_totals = arena.get<long long>(8);
    if (!d_code || !d_rank || !d_keys0 || !d_keys1 || !d_vals1 || !d_temp || !d_agg || !d_totals)
        return fail(GCZ_E_NOMEM, "suffix sort workspace for n=%lld", (long long)n);
    long long* agg_last = d_agg;
    long long* agg_keep = d_agg + tiles_n;
    long long* agg_groups = d_agg + 2 * tiles_n;

    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
    if (stats) {
        GCZ_CUDA(cudaEventCreate(&ev0)); GCZ_CUDA(cudaEventCreate(&ev1)); GCZ_CUDA(cudaEventCreate(&ev2));
        GCZ_CUDA(cudaEventRecord(ev0, st));
    }

    GCZ_CUDA(cudaMemcpyAsync(d_code, h_code, 256, cudaMemcpyHostToDevice, st));

    // full sort by the first k symbols; start in the buffer that makes the result land in d_sa
    const int npass = (key_bits + 7) / 8;
    RadixBuffers b;
    b.keys[0] = d_keys0; b.keys[1] = d_keys1;
    b.vals[0] = d_sa;    b.vals[1] = d_vals1;
    b.cur = npass & 1;
    const int pack_grid = (int)((n + kPackTile - 1) / kPackTile);
    GCZ_LAUNCH(ctx, pack_keys_kernel, pack_grid, kPackThreads, 0, st, d_text, n, d_code, bits, k, b.keys[b.cur], b.vals[b.cur]);
    SortStats ss;
    SortStats* ssp = stats ? &ss : nullptr;
    GCZ_TRY(radix_sort_pairs(ctx, st, b, n, 0, key_bits, d_temp, ssp));
    if (b.cur != 0) return fail(GCZ_E_INTERNAL, "initial sort landed in the wrong buffer");

    // groups of equal keys
    GCZ_LAUNCH(ctx, group_aggregate_kernel, (unsigned)tiles_n, kGrpThreads, 0, st, b.keys[0], n, agg_last, agg_keep, agg_groups);
    GCZ_LAUNCH(ctx, group_scan_kernel, 1, 1024, 0, st, agg_last, agg_keep, agg_groups, tiles_n, d_totals);
    long long h_totals[2] = { 0, 0 };
    GCZ_CUDA(cudaMemcpyAsync(h_totals, d_totals, sizeof(h_totals), cudaMemcpyDeviceToHost, st));
    GCZ_CUDA(cudaStreamSynchronize(st));
    int64_t m = h_totals[0], groups = h_totals[1];

    // refinement buffers: the list ping-pong is fresh memory, the sort buffers reuse the dead key arrays
    const size_t cap = (size_t)std::max<int64_t>(m, 1);
    uint32_t* list_pos[2] = { arena.get<uint32_t>(cap), arena.get<uint32_t>(cap) };
    uint32_t* list_suf[2] = { arena.get<uint32_t>(cap), arena.get<uint32_t>(cap) };
    uint32_t* list_gid[2] = { arena.get<uint32_t>(cap), arena.get<uint32_t>(cap) };
    uint32_t* r_vals1 = arena.get<uint32_t>(cap);
    if (!list_pos[0] || !list_pos[1] || !list_suf[0] || !list_suf[1] || !list_gid[0] || !list_gid[1] || !r_vals1)
        return fail(GCZ_E_NOMEM, "suffix sort refinement lists for %lld unresolved suffixes", (long long)m);

    GCZ_LAUNCH(ctx, (group_apply_kernel<true>), (unsigned)tiles_n, kGrpThreads, 0, st, b.keys[0], d_sa, nullptr, n,
               agg_last, agg_keep, agg_groups, d_rank, d_sa, list_pos[0], list_suf[0], list_gid[0]);
    if (stats) GCZ_CUDA(cudaEventRecord(ev1, st));

    const int low_bits = bits_for((uint64_t)n);        // rank + 1 <= n
    int cur = 0, rounds = 0;
    int64_t h = k;
    while (m > 0) {
        if (h >= 2 * n + 64) return fail(GCZ_E_INTERNAL, "suffix sort did not converge");
        RadixBuffers rb;
        rb.keys[0] = d_keys0; rb.keys[1] = d_keys1;
        rb.vals[0] = d_vals1; rb.vals[1] = r_vals1;
        rb.cur = 0;
        const int grid = (int)std::min<int64_t>((m + 255) / 256, (int64_t)ctx->sm_count * 16);
        GCZ_LAUNCH(ctx, refine_keys_kernel, grid, 256, 0, st, list_suf[cur], list_gid[cur], m, d_rank, n, h, low_bits,
                   rb.keys[0], rb.vals[0]);
        const int gid_bits = bits_for((uint64_t)std::max<int64_t>(groups - 1, 0));
        GCZ_TRY(radix_sort_pairs(ctx, st, rb, m, 0, low_bits + gid_bits, d_temp, ssp));
        const int64_t tiles_m = (m + kGrpTile - 1) / kGrpTile;
        GCZ_LAUNCH(ctx, group_aggregate_kernel, (unsigned)tiles_m, kGrpThreads, 0, st, rb.keys[rb.cur], m, agg_last, agg_keep, agg_groups);
        GCZ_LAUNCH(ctx, group_scan_kernel, 1, 1024, 0, st, agg_last, agg_keep, agg_groups, tiles_m, d_totals);
        GCZ_LAUNCH(ctx, (group_apply_kernel<false>), (unsigned)tiles_m, kGrpThreads, 0, st, rb.keys[rb.cur], rb.vals[rb.cur],
                   list_pos[cur], m, agg_last, agg_keep, agg_groups, d_rank, d_sa, list_pos[cur ^ 1], list_suf[cur ^ 1], list_gid[cur ^ 1]);
        GCZ_CUDA(cudaMemcpyAsync(h_totals, d_totals, sizeof(h_totals), cudaMemcpyDeviceToHost, st));
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
        cudaEventDestroy(ev0); cudaEventDestroy(ev1); cudaEventDestroy(ev2);
    }
    arena.release(mark0);
    return GCZ_OK;
}

}  // namespace gcz

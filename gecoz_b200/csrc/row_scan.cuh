// Exclusive scans of the rows of a small u32 matrix (per-tile counts: one row per symbol, one column per tile), one launch.
//
// The first version gave a row to ONE CTA that walked it 1024 values at a time — three barriers per step, 259 us for the six
// rows of 243 000 tile counts of a chr1-sized block, 142 us for the five rows of group aggregates of the suffix sorter.
// Here a row is cut into chunks of 4096 values, every CTA takes a (row, chunk) by ticket, scans it, publishes the chunk total
// and learns its offset by a decoupled look-back over the totals of the chunks before it in the row (a warp reads 32 of them
// at a time; rows have tens of chunks, so the chain is short).  Row 0 may be a running maximum instead of a sum (the suffix
// sorter's "last group boundary so far").
#pragma once

#include "device.cuh"

namespace gcz {

constexpr int kScanThreads = 1024;
constexpr int kScanPer = 4;
constexpr int kScanChunk = kScanThreads * kScanPer;

inline int64_t row_scan_chunks(int64_t len) { return (len + kScanChunk - 1) / kScanChunk; }
// scratch: one status word per (row, chunk) and the ticket
inline size_t row_scan_scratch_bytes(int rows, int64_t len) { return ((size_t)rows * (size_t)row_scan_chunks(len) + 2) * 8; }

#ifdef __CUDACC__
template <bool MAX0>
__global__ void __launch_bounds__(kScanThreads)
row_scan_chained_kernel(uint32_t* __restrict__ rows, int64_t len, int64_t chunks, unsigned long long* __restrict__ status,
                        unsigned* __restrict__ ticket, uint32_t* __restrict__ totals32, long long* __restrict__ totals64) {
    constexpr unsigned long long kAgg = 1ull << 62, kPrefix = 2ull << 62, kMask = kAgg - 1;
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_excl;
    __shared__ unsigned s_ticket;
    if (threadIdx.x == 0) s_ticket = atomicAdd(ticket, 1u);
    __syncthreads();
    const int64_t t = s_ticket;
    const int64_t r = t / chunks, chunk = t % chunks;
    const bool is_max = MAX0 && r == 0;
    uint32_t* row = rows + (size_t)r * len;
    const int64_t i0 = chunk * kScanChunk + (int64_t)threadIdx.x * kScanPer;
    uint32_t v[kScanPer];
#pragma unroll
    for (int j = 0; j < kScanPer; j++) v[j] = i0 + j < len ? row[i0 + j] : 0u;
    uint32_t mine = v[0];
#pragma unroll
    for (int j = 1; j < kScanPer; j++) mine = is_max ? max(mine, v[j]) : mine + v[j];
    // inclusive scan over the threads of the CTA
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane_id() >= (unsigned)o) incl = is_max ? max(incl, u) : incl + u;
    }
    if (lane_id() == 31) s_warp[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t before = 0;                                             // all warps before mine
    for (unsigned w = 0; w < (threadIdx.x >> 5); w++) before = is_max ? max(before, s_warp[w]) : before + s_warp[w];
    if (threadIdx.x < 32) {
        uint32_t total = 0;
        for (int w = 0; w < 32; w++) total = is_max ? max(total, s_warp[w]) : total + s_warp[w];
        unsigned long long* st = status + (size_t)r * chunks;
        uint32_t excl = 0;
        if (chunk == 0) {
            if (threadIdx.x == 0) st_relaxed_u64(&st[0], kPrefix | total);
        } else {
            if (threadIdx.x == 0) st_relaxed_u64(&st[chunk], kAgg | total);
            int64_t look = chunk - 1;
            while (true) {
                const int64_t idx = look - (int64_t)threadIdx.x;
                const unsigned long long w = idx >= 0 ? ld_relaxed_u64(&st[idx]) : kPrefix;      // before the row: prefix 0
                const unsigned flag = (unsigned)(w >> 62);
                const unsigned waiting = __ballot_sync(0xffffffffu, flag == 0);
                const unsigned prefixes = __ballot_sync(0xffffffffu, flag == 2);
                const unsigned upto = prefixes ? (unsigned)__ffs(prefixes) - 1u : 31u;          // lanes 0 .. upto are summed
                if (waiting & (0xffffffffu >> (31u - upto))) continue;                         // one of them is not there yet
                uint32_t part = threadIdx.x <= upto ? (uint32_t)(w & kMask) : 0u;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const uint32_t u = __shfl_xor_sync(0xffffffffu, part, o);
                    part = is_max ? max(part, u) : part + u;
                }
                excl = is_max ? max(excl, part) : excl + part;
                if (prefixes) break;
                look -= 32;
            }
            if (threadIdx.x == 0) st_relaxed_u64(&st[chunk], kPrefix | (is_max ? max(excl, total) : excl + total));
        }
        if (threadIdx.x == 0) {
            s_excl = excl;
            if (chunk == chunks - 1) {
                const uint32_t all = is_max ? max(excl, total) : excl + total;
                if (totals32) totals32[r] = all;
                if (totals64 && !(MAX0 && r == 0)) totals64[MAX0 ? r - 1 : r] = (long long)all;
            }
        }
    }
    __syncthreads();
    uint32_t run = is_max ? max(s_excl, before) : s_excl + before;
    const uint32_t up = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane_id() > 0) run = is_max ? max(run, up) : run + up;
#pragma unroll
    for (int j = 0; j < kScanPer; j++) {
        if (i0 + j < len) row[i0 + j] = run;
        run = is_max ? max(run, v[j]) : run + v[j];
    }
}

// rows[r][0 .. len) -> exclusive scans in place.  max0: row 0 is a running maximum and totals64 (if given) takes the totals
// of rows 1 .. as totals64[r - 1]; otherwise totals64[r].  `scratch` holds row_scan_scratch_bytes(rows, len).
inline int row_scan(DeviceCtx* ctx, cudaStream_t st, uint32_t* rows, int n_rows, int64_t len, bool max0, void* scratch,
                    uint32_t* totals32, long long* totals64) {
    if (n_rows <= 0 || len <= 0) return GCZ_OK;
    const int64_t chunks = row_scan_chunks(len);
    auto* status = static_cast<unsigned long long*>(scratch);
    auto* ticket = reinterpret_cast<unsigned*>(status + (size_t)n_rows * chunks);
    GCZ_CUDA(cudaMemsetAsync(scratch, 0, row_scan_scratch_bytes(n_rows, len), st));
    const unsigned grid = (unsigned)(chunks * n_rows);
    if (max0) GCZ_LAUNCH(ctx, row_scan_chained_kernel<true>, grid, kScanThreads, 0, st, rows, len, chunks, status, ticket, totals32, totals64);
    else      GCZ_LAUNCH(ctx, row_scan_chained_kernel<false>, grid, kScanThreads, 0, st, rows, len, chunks, status, ticket, totals32, totals64);
    return GCZ_OK;
}
#endif

}  // namespace gcz

// Device context, workspace arena and small device helpers shared by the CUDA translation units.
#pragma once

#include "gcz_host.h"

#include <cuda_runtime.h>

#include <atomic>
#include <condition_variable>
#include <mutex>

namespace gcz {

#define GCZ_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            return ::gcz::fail(e__ == cudaErrorMemoryAllocation ? GCZ_E_NOMEM : GCZ_E_CUDA,         \
                               "%s -> %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
        }                                                                                           \
    } while (0)

#define GCZ_TRY(expr)                       \
    do {                                    \
        int rc__ = (expr);                  \
        if (rc__ != GCZ_OK) return rc__;    \
    } while (0)

// Grow-only device arena: one cudaMalloc per device, bump-allocated per call.  A build of n symbols
// needs tens of n bytes; asking the driver for that on every block would cost more than the sort.
struct Arena {
    char*  base = nullptr;
    size_t capacity = 0;
    size_t top = 0;
    int reserve(size_t bytes);                 // ensures capacity >= bytes (frees + reallocates; contents lost)
    void reset() { top = 0; }
    size_t mark() const { return top; }
    void release(size_t m) { top = m; }
    void* raw(size_t bytes) {
        const size_t aligned = (top + 255) & ~size_t(255);
        if (aligned + bytes > capacity) return nullptr;
        top = aligned + bytes;
        return base + aligned;
    }
    template <class T> T* get(size_t count) { return static_cast<T*>(raw(count * sizeof(T))); }
    void destroy();
};

// A staged text is only used for a host buffer that still shows the same bytes at kProbeBytes evenly spaced places (a buffer
// freed and reallocated at the same address must not find the old text).  The contract of gcz_count_symbols is that the
// buffer does not change before gcz_build_block; the probe catches the accidents, not an adversary.
constexpr int kProbeBytes = 2048;

struct DeviceCtx {
    int          device = -1;
    int          sm_count = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t copy_stream = nullptr;       // device-to-host copies that overlap the build
    cudaEvent_t  copy_event = nullptr;
    // Texts left on the device by gcz_count_symbols for the gcz_build_block that follows on the same host buffer.
    // Two slots and a stream and lock of their own: the upload + histogram of block k + 1 runs while block k is
    // being built (the build holds `mu`, the staging only `stage_mu`).
    struct StagedText {
        const void* host = nullptr;           // key: host buffer and length
        int64_t     n = 0;
        uint8_t*    dev = nullptr;
        size_t      cap = 0;
        int         state = 0;                // 0 free, 1 staged, 2 in use by a build, 3 being filled
        uint64_t    stamp = 0;                // staging order (the older staged text is overwritten first)
        uint8_t     probe[kProbeBytes] = {};  // bytes of the host buffer at fixed places, compared again at build time
        int64_t     counts[256] = {};         // the histogram gcz_count_symbols returned for this text (the build reuses it)
    };
    std::mutex   stage_mu;                   // slot bookkeeping (short critical sections only)
    std::mutex   stage_io_mu;                // one staging at a time
    cudaStream_t stage_stream = nullptr;
    StagedText   staged[2];
    uint64_t     stage_clock = 0;
    unsigned long long* stage_counts = nullptr;   // device, 256 counters
    // Device copies of the two bodies of a block whose outputs are host buffers.  They live outside the arena, two of
    // them: the call that built them releases the device (`mu`) as soon as its kernels are done and waits for its copies
    // on its own, so that the next block's kernels run while the previous block's bodies are still on their way to the
    // host — what a writer with two blocks in flight per device (fmt/GecozFileWriter.java:174-227) needs.
    struct OutSlot {
        uint8_t*    dev = nullptr;
        size_t      cap = 0;
        cudaEvent_t done = nullptr;            // recorded on copy_stream after the last copy of the block
        bool        busy = false;
    };
    std::mutex              out_mu;
    std::condition_variable out_cv;
    OutSlot                 out_slot[2];
    std::mutex   mu;                          // one build / query batch at a time per device (shared arena)
    Arena        arena;
    // Small transfers of a build (symbol tables, vector descriptors, the totals the host reads between refinement rounds) do
    // not go through the copy engines: those are busy with the next block's text and the previous block's bodies, and a
    // 256-byte copy queued behind a 77 MB one waits a millisecond.  They go through this mapped pinned buffer instead, moved
    // by a one-CTA kernel on the build's stream (small_upload / small_read, api.cu).  Used under `mu` only.
    uint8_t*     pinned = nullptr;             // host address
    uint8_t*     pinned_dev = nullptr;         // the same memory as the device sees it
    size_t       pinned_bytes = 0;
    size_t       pinned_top = 0;
    std::atomic<int64_t> launches{0};        // kernels launched by this library on this device
    bool         iwt_attr = false;             // dynamic-smem opt-in done for iwt_low_levels_kernel
    unsigned     sort_attr_mask = 0;           // bit 0 / 1: dynamic-smem opt-in done for the onesweep kernels with 64- / 32-bit status words
    unsigned     text_hist_attr = 0;           // bit p: the same for text_hist_kernel<p>
};

int          get_ctx(int device, DeviceCtx** out);   // creates on first use; fails with GCZ_E_NODEVICE
cudaStream_t stream_of(DeviceCtx* ctx);              // thread-local override from gcz_set_stream, else own
void         destroy_all_ctx();

// stream-ordered host -> device copy of a small object (the host bytes are taken before the call returns)
int small_upload(DeviceCtx* ctx, cudaStream_t st, void* d_dst, const void* h_src, size_t bytes);
// device -> host copy of a small object; returns after the stream has been synchronised
int small_read(DeviceCtx* ctx, cudaStream_t st, void* h_dst, const void* d_src, size_t bytes);

inline bool is_device_ptr(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

#define GCZ_LAUNCH(ctx, kernel, grid, block, smem, stream, ...)                                    \
    do {                                                                                            \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                                 \
        (ctx)->launches++;                                                                          \
        GCZ_CUDA(cudaPeekAtLastError());                                                            \
    } while (0)

// ---- device helpers -----------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt() { unsigned m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// 32 bytes per thread and instruction (sm_100: LDG.E.256): a thread that owns consecutive elements fetches whole 32-byte sectors,
// where two 16-byte loads ask for each sector twice.  The address is 32-byte aligned; the data is not written by the kernel.
__device__ __forceinline__ void ld_nc_256(const void* p, unsigned long long (&v)[4]) {
    asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3]) : "l"(p));
}

// warp inclusive scans
__device__ __forceinline__ unsigned warp_incl_sum(unsigned v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { unsigned t = __shfl_up_sync(0xffffffffu, v, o); if (lane_id() >= (unsigned)o) v += t; }
    return v;
}
__device__ __forceinline__ long long warp_incl_max(long long v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { long long t = __shfl_up_sync(0xffffffffu, v, o); if (lane_id() >= (unsigned)o && t > v) v = t; }
    return v;
}
#endif

}  // namespace gcz

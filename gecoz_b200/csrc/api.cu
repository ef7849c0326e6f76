// C ABI (include/gcz.h): device contexts, the per-block build driver and thin wrappers over the query side.
#include "query.cuh"
#include "suffix_sort.cuh"
#include "wavelet_build.cuh"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>

namespace gcz {

// ---- errors ---------------------------------------------------------------------------------------------
static thread_local char t_error[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_error, sizeof(t_error), fmt, ap);
    va_end(ap);
    return code;
}
void clear_error() { t_error[0] = 0; }

// ---- arena -------------------------------------------------------------------------------------------------
int Arena::reserve(size_t bytes) {
    if (bytes <= capacity) return GCZ_OK;
    if (base) { cudaFree(base); base = nullptr; capacity = 0; }
    const size_t want = (bytes + ((size_t)1 << 26)) & ~(((size_t)1 << 20) - 1);
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&base), want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        base = nullptr;
        return fail(GCZ_E_NOMEM, "device workspace of %zu bytes: %s", want, cudaGetErrorString(e));
    }
    capacity = want;
    top = 0;
    return GCZ_OK;
}
void Arena::destroy() {
    if (base) cudaFree(base);
    base = nullptr; capacity = 0; top = 0;
}

// ---- device contexts ----------------------------------------------------------------------------------------
static std::mutex g_ctx_mu;
static std::map<int, std::unique_ptr<DeviceCtx>> g_ctx;
static thread_local std::map<int, cudaStream_t> t_stream_override;

int get_ctx(int device, DeviceCtx** out) {
    std::lock_guard<std::mutex> lock(g_ctx_mu);
    auto it = g_ctx.find(device);
    if (it != g_ctx.end()) { *out = it->second.get(); return GCZ_OK; }
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
        cudaGetLastError();
        return fail(GCZ_E_NODEVICE, "no CUDA device is visible; this library has no CPU path");
    }
    if (device < 0 || device >= count) return fail(GCZ_E_ARG, "device %d out of range (%d visible)", device, count);
    GCZ_CUDA(cudaSetDevice(device));
    std::unique_ptr<DeviceCtx> ctx(new DeviceCtx());
    ctx->device = device;
    cudaDeviceProp prop;
    GCZ_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    GCZ_CUDA(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
    GCZ_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    GCZ_CUDA(cudaStreamCreateWithFlags(&ctx->stage_stream, cudaStreamNonBlocking));
    GCZ_CUDA(cudaMalloc(reinterpret_cast<void**>(&ctx->stage_counts), 256 * 8));
    GCZ_CUDA(cudaEventCreateWithFlags(&ctx->copy_event, cudaEventDisableTiming));
    {
        void* h = nullptr; void* d = nullptr;
        const size_t bytes = (size_t)256 << 10;
        if (cudaHostAlloc(&h, bytes, cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess &&
            cudaHostGetDevicePointer(&d, h, 0) == cudaSuccess) {
            ctx->pinned = static_cast<uint8_t*>(h); ctx->pinned_dev = static_cast<uint8_t*>(d); ctx->pinned_bytes = bytes;
        } else {
            cudaGetLastError();                           // without it small transfers take the copy engines
            if (h) cudaFreeHost(h);
        }
    }
    *out = ctx.get();
    g_ctx[device] = std::move(ctx);
    return GCZ_OK;
}

__global__ void small_copy_kernel(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, size_t bytes) {
    if (((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src) | bytes) & 3) == 0) {
        for (size_t i = threadIdx.x; i < bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(dst)[i] = reinterpret_cast<const uint32_t*>(src)[i];
    } else {
        for (size_t i = threadIdx.x; i < bytes; i += blockDim.x) dst[i] = src[i];
    }
}

// a slot of the mapped buffer (16-byte aligned); when the buffer is full the stream is drained and it starts over
static int small_slot(DeviceCtx* ctx, cudaStream_t st, size_t bytes, size_t* at) {
    const size_t need = (bytes + 15) & ~(size_t)15;
    if (ctx->pinned_top + need > ctx->pinned_bytes) {
        GCZ_CUDA(cudaStreamSynchronize(st));
        ctx->pinned_top = 0;
    }
    *at = ctx->pinned_top;
    ctx->pinned_top += need;
    return GCZ_OK;
}

int small_upload(DeviceCtx* ctx, cudaStream_t st, void* d_dst, const void* h_src, size_t bytes) {
    if (bytes == 0) return GCZ_OK;
    if (!ctx->pinned || bytes > ctx->pinned_bytes / 2) {
        GCZ_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, st));
        GCZ_CUDA(cudaStreamSynchronize(st));              // the caller's buffer may go away
        return GCZ_OK;
    }
    size_t at = 0;
    GCZ_TRY(small_slot(ctx, st, bytes, &at));
    std::memcpy(ctx->pinned + at, h_src, bytes);
    GCZ_LAUNCH(ctx, small_copy_kernel, 1, 256, 0, st, static_cast<uint8_t*>(d_dst), ctx->pinned_dev + at, bytes);
    return GCZ_OK;
}

int small_read(DeviceCtx* ctx, cudaStream_t st, void* h_dst, const void* d_src, size_t bytes) {
    if (bytes == 0) return GCZ_OK;
    if (!ctx->pinned || bytes > ctx->pinned_bytes / 2) {
        GCZ_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, st));
        GCZ_CUDA(cudaStreamSynchronize(st));
        return GCZ_OK;
    }
    size_t at = 0;
    GCZ_TRY(small_slot(ctx, st, bytes, &at));
    GCZ_LAUNCH(ctx, small_copy_kernel, 1, 256, 0, st, ctx->pinned_dev + at, static_cast<const uint8_t*>(d_src), bytes);
    GCZ_CUDA(cudaStreamSynchronize(st));
    std::memcpy(h_dst, ctx->pinned + at, bytes);
    return GCZ_OK;
}

cudaStream_t stream_of(DeviceCtx* ctx) {
    auto it = t_stream_override.find(ctx->device);
    return it != t_stream_override.end() ? it->second : ctx->own_stream;
}

void destroy_all_ctx() {
    std::lock_guard<std::mutex> lock(g_ctx_mu);
    for (auto& kv : g_ctx) {
        cudaSetDevice(kv.first);
        kv.second->arena.destroy();
        if (kv.second->own_stream) cudaStreamDestroy(kv.second->own_stream);
        if (kv.second->copy_stream) cudaStreamDestroy(kv.second->copy_stream);
        if (kv.second->copy_event) cudaEventDestroy(kv.second->copy_event);
        if (kv.second->stage_stream) cudaStreamDestroy(kv.second->stage_stream);
        if (kv.second->stage_counts) cudaFree(kv.second->stage_counts);
        if (kv.second->pinned) cudaFreeHost(kv.second->pinned);
        for (auto& slot : kv.second->staged) if (slot.dev) cudaFree(slot.dev);
        for (auto& slot : kv.second->out_slot) { if (slot.dev) cudaFree(slot.dev); if (slot.done) cudaEventDestroy(slot.done); }
    }
    g_ctx.clear();
}

static thread_local gcz_build_timing t_timing;

// ---- histogram -------------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(512)
symbol_histogram_kernel(const uint8_t* __restrict__ text, int64_t n, unsigned long long* __restrict__ counts) {
    __shared__ unsigned s_cnt[8][256];          // one copy per pair of warps to spread same-symbol atomics
    for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) (&s_cnt[0][0])[i] = 0;
    __syncthreads();
    unsigned* mine = s_cnt[(threadIdx.x >> 6) & 7];
    const int64_t n16 = n >> 4;
    const uint4* t16 = reinterpret_cast<const uint4*>(text);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
        const uint4 q = t16[i];
        const uint32_t w[4] = { q.x, q.y, q.z, q.w };
#pragma unroll
        for (int k = 0; k < 4; k++) {
            atomicAdd(&mine[w[k] & 255], 1u);
            atomicAdd(&mine[(w[k] >> 8) & 255], 1u);
            atomicAdd(&mine[(w[k] >> 16) & 255], 1u);
            atomicAdd(&mine[w[k] >> 24], 1u);
        }
    }
    if (blockIdx.x == 0) {
        for (int64_t i = (n16 << 4) + threadIdx.x; i < n; i += blockDim.x) atomicAdd(&mine[text[i]], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        unsigned long long v = 0;
        for (int k = 0; k < 8; k++) v += s_cnt[k][i];
        if (v) atomicAdd(&counts[i], v);
    }
}
}  // namespace

static int histogram_device(DeviceCtx* ctx, cudaStream_t st, const uint8_t* d_text, int64_t n, unsigned long long* d_counts) {
    GCZ_CUDA(cudaMemsetAsync(d_counts, 0, 256 * 8, st));
    if ((reinterpret_cast<uintptr_t>(d_text) & 15) != 0) return fail(GCZ_E_ARG, "device text must be 16-byte aligned");
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n / 16 + 511) / 512, (int64_t)ctx->sm_count * 4));
    GCZ_LAUNCH(ctx, symbol_histogram_kernel, grid, 512, 0, st, d_text, n, d_counts);
    return GCZ_OK;
}

constexpr int64_t kStageChunk = 4 << 20;

static void text_probe(const uint8_t* text, int64_t n, uint8_t out[kProbeBytes]) {
    for (int i = 0; i < kProbeBytes; i++) out[i] = text[(int64_t)((__int128)(n - 1) * i / (kProbeBytes - 1))];
}

static int count_symbols(int device, const uint8_t* text, int64_t n, int64_t counts[256]) {
    if (!text || !counts || n <= 0) return fail(GCZ_E_ARG, "count_symbols arguments");
    DeviceCtx* ctx = nullptr;
    GCZ_TRY(get_ctx(device, &ctx));
    if (is_device_ptr(text)) {
        std::lock_guard<std::mutex> lock(ctx->mu);
        GCZ_CUDA(cudaSetDevice(ctx->device));
        cudaStream_t st = stream_of(ctx);
        ctx->arena.reset();
        if (ctx->arena.capacity < (1 << 20)) GCZ_TRY(ctx->arena.reserve(1 << 20));
        unsigned long long* d_counts = ctx->arena.get<unsigned long long>(256);
        if (!d_counts) return fail(GCZ_E_NOMEM, "histogram scratch");
        GCZ_TRY(histogram_device(ctx, st, text, n, d_counts));
        GCZ_CUDA(cudaMemcpyAsync(counts, d_counts, 256 * 8, cudaMemcpyDeviceToHost, st));
        GCZ_CUDA(cudaStreamSynchronize(st));
        return GCZ_OK;
    }
    // host text: the upload is kept in one of two slots, so that the gcz_build_block that follows on the same host
    // buffer finds the text on the device — and this call may overlap the build of the previous block.
    // stage_io_mu: one staging at a time (stream, counters); stage_mu: slot bookkeeping only, never held across a copy.
    std::lock_guard<std::mutex> io(ctx->stage_io_mu);
    GCZ_CUDA(cudaSetDevice(ctx->device));
    struct Filling {
        DeviceCtx* ctx; DeviceCtx::StagedText* slot = nullptr; bool done = false;
        ~Filling() { if (slot && !done) { std::lock_guard<std::mutex> l(ctx->stage_mu); slot->state = 0; slot->host = nullptr; } }
    } fill{ctx};
    {
        std::lock_guard<std::mutex> l(ctx->stage_mu);
        DeviceCtx::StagedText* slot = nullptr;
        for (auto& c : ctx->staged) if (c.state == 0) { slot = &c; break; }
        if (!slot) for (auto& c : ctx->staged) if (c.state == 1 && (!slot || c.stamp < slot->stamp)) slot = &c;
        if (!slot) return fail(GCZ_E_INTERNAL, "both text slots are in use by builds");
        slot->state = 3;                                      // being filled: nobody else looks at it
        slot->host = nullptr;
        fill.slot = slot;
    }
    DeviceCtx::StagedText* slot = fill.slot;
    if (slot->cap < (size_t)n + 64) {
        if (slot->dev) cudaFree(slot->dev);
        slot->dev = nullptr; slot->cap = 0;
        const size_t want = ((size_t)n + 64 + ((size_t)1 << 20)) & ~(((size_t)1 << 20) - 1);
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&slot->dev), want);
        if (e != cudaSuccess) { cudaGetLastError(); slot->dev = nullptr; return fail(GCZ_E_NOMEM, "text staging of %zu bytes", want); }
        slot->cap = want;
    }
    cudaStream_t st = ctx->stage_stream;
    // In pieces, and never more than two of them queued: the small uploads of a build running on the other stream share
    // the copy engine, which serves its queue in order — with the whole text queued at once, a 256-byte table of the build
    // waited 4.5 ms for it (every other build of a pipelined writer: GCZ_BUILD_TRACE=1, profiles/e2e_staging_r02.md).
    {
        cudaEvent_t piece[2] = { nullptr, nullptr };
        GCZ_CUDA(cudaEventCreateWithFlags(&piece[0], cudaEventDisableTiming));
        GCZ_CUDA(cudaEventCreateWithFlags(&piece[1], cudaEventDisableTiming));
        int k = 0;
        cudaError_t err = cudaSuccess;
        for (int64_t off = 0; off < n && err == cudaSuccess; off += kStageChunk, k++) {
            if (k >= 2) err = cudaEventSynchronize(piece[k & 1]);          // the piece before the previous one has landed
            if (err == cudaSuccess) err = cudaMemcpyAsync(slot->dev + off, text + off, (size_t)std::min<int64_t>(kStageChunk, n - off), cudaMemcpyHostToDevice, st);
            if (err == cudaSuccess) err = cudaEventRecord(piece[k & 1], st);
        }
        cudaEventDestroy(piece[0]); cudaEventDestroy(piece[1]);
        GCZ_CUDA(err);
    }
    GCZ_TRY(histogram_device(ctx, st, slot->dev, n, ctx->stage_counts));
    GCZ_CUDA(cudaMemcpyAsync(counts, ctx->stage_counts, 256 * 8, cudaMemcpyDeviceToHost, st));
    GCZ_CUDA(cudaStreamSynchronize(st));
    {
        std::lock_guard<std::mutex> l(ctx->stage_mu);
        slot->host = text; slot->n = n; slot->state = 1; slot->stamp = ++ctx->stage_clock;
        text_probe(text, n, slot->probe);
        std::memcpy(slot->counts, counts, sizeof(slot->counts));
        fill.done = true;
    }
    return GCZ_OK;
}

// ---- one block -------------------------------------------------------------------------------------------------
static int build_block(int device, const uint8_t* text, int64_t n, int32_t sampling_rate, const gcz_shape* shape,
                       uint8_t* gcz_body, int64_t gcz_body_len, uint8_t* gcx_body, int64_t gcx_body_len,
                       int32_t* sa_out, uint8_t* bwt_out) {
    if (!text || !shape || !gcz_body || !gcx_body) return fail(GCZ_E_ARG, "build_block: null buffer");
    if (n <= 0 || n > 0x7FFFFFFFll) return fail(GCZ_E_RANGE, "block of %lld symbols (1 .. 2^31-1 supported)", (long long)n);
    if (sampling_rate <= 0 || (sampling_rate & (sampling_rate - 1)) != 0) return fail(GCZ_E_ARG, "sampling rate must be a power of two");
    const int sf = 31 - __builtin_clz((unsigned)sampling_rate);
    if (shape->length != n) return fail(GCZ_E_ARG, "shape was built for %lld symbols, text has %lld", (long long)shape->length, (long long)n);
    if (gcz_body_len != shape->size) return fail(GCZ_E_ARG, "gcz body must be %lld bytes", (long long)shape->size);
    if (gcx_body_len != index_size(n, sf)) return fail(GCZ_E_ARG, "gcx body must be %lld bytes", (long long)index_size(n, sf));

    DeviceCtx* ctx = nullptr;
    GCZ_TRY(get_ctx(device, &ctx));
    GCZ_CUDA(cudaSetDevice(ctx->device));
    // both bodies go to host buffers and no parity artefact is asked for: they are built in an output slot outside the
    // arena, and the device is released before their copies have landed (see DeviceCtx::OutSlot)
    const bool tail_overlap = !is_device_ptr(gcz_body) && !is_device_ptr(gcx_body) && !sa_out && !bwt_out;
    struct SlotHold {
        DeviceCtx* ctx; DeviceCtx::OutSlot* slot = nullptr;
        ~SlotHold() { if (slot) { { std::lock_guard<std::mutex> l(ctx->out_mu); slot->busy = false; } ctx->out_cv.notify_all(); } }
    } hold{ctx};
    if (tail_overlap) {
        std::unique_lock<std::mutex> l(ctx->out_mu);
        ctx->out_cv.wait(l, [&] { return !ctx->out_slot[0].busy || !ctx->out_slot[1].busy; });
        hold.slot = !ctx->out_slot[0].busy ? &ctx->out_slot[0] : &ctx->out_slot[1];
        hold.slot->busy = true;
        l.unlock();
        const size_t want = (((size_t)gcz_body_len + 511) & ~size_t(255)) + (size_t)gcx_body_len + 512;
        if (hold.slot->cap < want) {
            if (hold.slot->dev) cudaFree(hold.slot->dev);
            hold.slot->dev = nullptr; hold.slot->cap = 0;
            const size_t rounded = (want + ((size_t)8 << 20)) & ~(((size_t)1 << 20) - 1);
            cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&hold.slot->dev), rounded);
            if (e != cudaSuccess) { cudaGetLastError(); hold.slot->dev = nullptr; return fail(GCZ_E_NOMEM, "output staging of %zu bytes", rounded); }
            hold.slot->cap = rounded;
        }
        if (!hold.slot->done) GCZ_CUDA(cudaEventCreate(&hold.slot->done));
    }
    std::unique_lock<std::mutex> lock(ctx->mu);
    cudaStream_t st = stream_of(ctx);
    const int64_t launches0 = ctx->launches;
    std::memset(&t_timing, 0, sizeof(t_timing));

    // a text staged by gcz_count_symbols for this host buffer is used once (the buffer may change afterwards)
    struct StageClaim {
        DeviceCtx* ctx; DeviceCtx::StagedText* slot = nullptr;
        ~StageClaim() { if (slot) { std::lock_guard<std::mutex> l(ctx->stage_mu); slot->state = 0; slot->host = nullptr; } }
    } claim{ctx};
    {
        std::lock_guard<std::mutex> l(ctx->stage_mu);
        uint8_t probe[kProbeBytes];
        bool probed = false;
        for (auto& c : ctx->staged) {
            if (c.state != 1 || c.host != text || c.n != n) continue;
            if (!probed) { text_probe(text, n, probe); probed = true; }
            if (std::memcmp(probe, c.probe, kProbeBytes) != 0) { c.state = 0; c.host = nullptr; continue; }     // stale
            if (!claim.slot || c.stamp < claim.slot->stamp) claim.slot = &c;
        }
        if (claim.slot) claim.slot->state = 2;
    }
    const bool text_staged = claim.slot != nullptr;
    const bool text_dev = text_staged || is_device_ptr(text), gcz_dev = is_device_ptr(gcz_body), gcx_dev = is_device_ptr(gcx_body);
    const bool sa_dev = sa_out && is_device_ptr(sa_out), bwt_dev = bwt_out && is_device_ptr(bwt_out);

    const size_t fixed = (text_dev ? 0 : (size_t)n + 256) + (size_t)n * 4 + (size_t)n + 256 +
                         ((gcz_dev || tail_overlap) ? 0 : (size_t)gcz_body_len + 256) + ((gcx_dev || tail_overlap) ? 0 : (size_t)gcx_body_len + 256) + 4096;
    const size_t need = fixed + std::max(suffix_sort_workspace_bytes(n), wavelet_workspace_bytes(n, sf)) + ((size_t)8 << 20);
    ctx->arena.reset();
    if (ctx->arena.capacity < need) GCZ_TRY(ctx->arena.reserve(need));
    Arena& arena = ctx->arena;

    cudaEvent_t ev[4];
    for (auto& e : ev) GCZ_CUDA(cudaEventCreate(&e));
    GCZ_CUDA(cudaEventRecord(ev[0], st));
    // GCZ_BUILD_TRACE=1: host wall clock at the stage boundaries of every call, one line on stderr
    static const bool trace = [] { const char* e = std::getenv("GCZ_BUILD_TRACE"); return e && e[0] == '1'; }();
    const auto w0 = std::chrono::steady_clock::now();
    auto since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count(); };
    double w_hist = 0, w_sort = 0, w_wave = 0;

    const uint8_t* d_text = text_staged ? claim.slot->dev : text;
    if (!text_dev) {
        uint8_t* d = arena.get<uint8_t>((size_t)n + 64);
        if (!d) return fail(GCZ_E_NOMEM, "text staging");
        GCZ_CUDA(cudaMemcpyAsync(d, text, (size_t)n, cudaMemcpyHostToDevice, st));
        d_text = d;
    }
    uint32_t* d_sa = (sa_dev) ? reinterpret_cast<uint32_t*>(sa_out) : arena.get<uint32_t>((size_t)n);
    uint8_t* d_bwt = (bwt_dev) ? bwt_out : arena.get<uint8_t>((size_t)n + 64);
    uint8_t* d_gcz = gcz_dev ? gcz_body : tail_overlap ? hold.slot->dev : arena.get<uint8_t>((size_t)gcz_body_len + 64);
    uint8_t* d_gcx = gcx_dev ? gcx_body : tail_overlap ? hold.slot->dev + (((size_t)gcz_body_len + 511) & ~size_t(255))
                                                       : arena.get<uint8_t>((size_t)gcx_body_len + 64);
    unsigned long long* d_counts = arena.get<unsigned long long>(256);
    if (!d_sa || !d_bwt || !d_gcz || !d_gcx || !d_counts) return fail(GCZ_E_NOMEM, "block buffers for n=%lld", (long long)n);
    GCZ_CUDA(cudaEventRecord(ev[1], st));

    // the histogram drives the key packer and guards against a shape that belongs to another text (the reference trusts
    // its caller; a mismatch there corrupts the file).  A staged text brings the histogram gcz_count_symbols computed from
    // the same device copy; anything else is counted here.
    int64_t counts[256];
    if (text_staged) {
        std::memcpy(counts, claim.slot->counts, sizeof(counts));
    } else {
        GCZ_TRY(histogram_device(ctx, st, d_text, n, d_counts));
        GCZ_CUDA(cudaMemcpyAsync(counts, d_counts, sizeof(counts), cudaMemcpyDeviceToHost, st));
        GCZ_CUDA(cudaStreamSynchronize(st));
    }
    for (int c = 0; c < 256; c++) {
        if ((counts[c] > 0) != (shape->bit_lengths[c] > 0)) return fail(GCZ_E_ARG, "shape does not match the text (symbol %d)", c);
    }
    {
        // the exact test: the shape this text's histogram gives must be the one the caller sized the file slices with
        gcz_shape mine;
        GCZ_TRY(shape_from_counts(counts, &mine));
        if (mine.size != shape->size || std::memcmp(mine.bit_lengths, shape->bit_lengths, sizeof(mine.bit_lengths)) != 0 ||
            mine.n_nodes != shape->n_nodes || std::memcmp(mine.node_bits, shape->node_bits, sizeof(int64_t) * (size_t)mine.n_nodes) != 0)
            return fail(GCZ_E_ARG, "shape does not match the text (it was built from another histogram)");
    }
    w_hist = since(w0);

    SuffixSortStats ss;
    // SA entries may carry their BWT symbol (no text gather afterwards); GCZ_BWT_GATHER=1 forces the gather path
    // that blocks too large for the carry take (tests)
    int carry_shift = std::getenv("GCZ_BWT_GATHER") ? 0 : 1;
    GCZ_TRY(suffix_sort(ctx, st, d_text, n, counts, d_sa, arena, &ss, &carry_shift));
    w_sort = since(w0);
    WaveletStats ws;
    GCZ_TRY(build_wavelet_structures(ctx, st, d_text, d_sa, carry_shift, sa_out != nullptr, n, shape, sf, d_bwt, d_gcz, d_gcx, arena, &ws,
                                     gcz_dev ? nullptr : gcz_body, ctx->copy_stream, ctx->copy_event));
    w_wave = since(w0);
    GCZ_CUDA(cudaEventRecord(ev[2], st));

    const int64_t launches = ctx->launches - launches0;
    if (tail_overlap) {
        // the kernels are done (build_wavelet_structures waits for them to read its stage times): the .gcx body follows the
        // .gcz body on the copy stream, the device goes to the next block, and this call waits for its own copies only
        GCZ_CUDA(cudaMemcpyAsync(gcx_body, d_gcx, (size_t)gcx_body_len, cudaMemcpyDeviceToHost, ctx->copy_stream));
        GCZ_CUDA(cudaEventRecord(ev[3], ctx->copy_stream));
        GCZ_CUDA(cudaEventRecord(hold.slot->done, ctx->copy_stream));
        lock.unlock();
        GCZ_CUDA(cudaEventSynchronize(hold.slot->done));
    } else {
        if (!gcz_dev) GCZ_CUDA(cudaStreamWaitEvent(st, ctx->copy_event, 0));         // the .gcz body went out while the index was built
        if (!gcx_dev) GCZ_CUDA(cudaMemcpyAsync(gcx_body, d_gcx, (size_t)gcx_body_len, cudaMemcpyDeviceToHost, st));
        if (sa_out && !sa_dev) GCZ_CUDA(cudaMemcpyAsync(sa_out, d_sa, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        if (bwt_out && !bwt_dev) GCZ_CUDA(cudaMemcpyAsync(bwt_out, d_bwt, (size_t)n, cudaMemcpyDeviceToHost, st));
        GCZ_CUDA(cudaEventRecord(ev[3], st));
        GCZ_CUDA(cudaEventSynchronize(ev[3]));
    }

    if (trace) {
        std::fprintf(stderr, "[gcz build] n=%lld host wall ms: histogram %.3f, suffix sort %.3f (device: initial %.3f refine %.3f), wavelet %.3f "
                             "(device: bwt+hswt %.3f ssa %.3f), copies out %.3f; text %s\n", (long long)n, w_hist, w_sort - w_hist, ss.initial_ms,
                     ss.refine_ms, w_wave - w_sort, ws.bwt_hswt_ms, ws.ssa_ms, since(w0) - w_wave, text_staged ? "staged" : text_dev ? "device" : "host");
    }
    cudaEventElapsedTime(&t_timing.h2d_ms, ev[0], ev[1]);
    cudaEventElapsedTime(&t_timing.d2h_ms, ev[2], ev[3]);
    cudaEventElapsedTime(&t_timing.total_ms, ev[0], ev[3]);
    t_timing.sort_initial_ms = ss.initial_ms;
    t_timing.sort_refine_ms = ss.refine_ms;
    t_timing.bwt_hswt_ms = ws.bwt_hswt_ms;
    t_timing.ssa_ms = ws.ssa_ms;
    t_timing.refine_rounds = ss.rounds;
    t_timing.radix_launches = ss.radix_passes;
    t_timing.radix_elements = ss.radix_elements;
    t_timing.radix_ms = ss.radix_ms;
    t_timing.kernel_launches = launches;
    t_timing.symbols_per_key = ss.symbols_per_key;
    t_timing.long_runs = ss.long_runs;
    t_timing.unresolved_after_first_sort = ss.unresolved_after_first_sort;
    t_timing.radix_full_launches = ss.radix_full_passes;
    t_timing.radix_full_ms = ss.radix_full_ms;
    t_timing.radix_text_ms = ss.radix_text_ms;
    for (auto& e : ev) cudaEventDestroy(e);
    return GCZ_OK;
}

}  // namespace gcz

// =====================================================================================================================
using namespace gcz;

extern "C" {

int gcz_init(int n_devices, const int* device_ids) {
    clear_error();
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
        cudaGetLastError();
        return fail(GCZ_E_NODEVICE, "no CUDA device is visible; this library has no CPU path");
    }
    if (n_devices <= 0) n_devices = count;
    for (int i = 0; i < n_devices; i++) {
        DeviceCtx* ctx = nullptr;
        GCZ_TRY(get_ctx(device_ids ? device_ids[i] : i, &ctx));
    }
    return GCZ_OK;
}

void gcz_shutdown(void) { destroy_all_ctx(); }
const char* gcz_last_error(void) { return t_error; }
const char* gcz_version(void) { return "gecoz_b200 0.1 (sm_100a)"; }

int gcz_device_count(void) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) { cudaGetLastError(); return 0; }
    return count;
}

int gcz_set_stream(int device, void* cuda_stream) {
    clear_error();
    DeviceCtx* ctx = nullptr;
    GCZ_TRY(get_ctx(device, &ctx));
    if (cuda_stream) t_stream_override[device] = static_cast<cudaStream_t>(cuda_stream);
    else t_stream_override.erase(device);
    return GCZ_OK;
}

int gcz_shape_from_counts(const int64_t counts[256], gcz_shape* out) {
    clear_error();
    if (!counts || !out) return fail(GCZ_E_ARG, "null argument");
    return guarded([&] { return shape_from_counts(counts, out); });
}
int64_t gcz_shape_write(const gcz_shape* shape, uint8_t* out, int64_t cap) {
    clear_error();
    if (!shape || !out) return fail(GCZ_E_ARG, "null argument");
    return shape_write(shape, out, cap);
}
int gcz_shape_read(const uint8_t* body, int64_t body_len, gcz_shape* out) {
    clear_error();
    if (!body || !out || body_len <= 0) return fail(GCZ_E_ARG, "null argument");
    return shape_read(body, body_len, out);
}
int64_t gcz_ranked_bytes(int64_t len_bits) { return ranked_bytes(len_bits); }
int64_t gcz_index_size(int64_t n, int32_t sampling_factor) { return index_size(n, sampling_factor); }

int gcz_count_symbols(int device, const uint8_t* text, int64_t n, int64_t counts[256]) {
    clear_error();
    return guarded([&] { return count_symbols(device, text, n, counts); });
}

int gcz_build_block(int device, const uint8_t* text, int64_t n, int32_t sampling_rate, const gcz_shape* shape,
                    uint8_t* gcz_body, int64_t gcz_body_len, uint8_t* gcx_body, int64_t gcx_body_len,
                    int32_t* sa_out, uint8_t* bwt_out) {
    clear_error();
    return guarded([&] { return build_block(device, text, n, sampling_rate, shape, gcz_body, gcz_body_len, gcx_body, gcx_body_len, sa_out, bwt_out); });
}

int gcz_last_build_timing(gcz_build_timing* out) {
    if (!out) return fail(GCZ_E_ARG, "null argument");
    *out = t_timing;
    return GCZ_OK;
}

int gcz_open_block(int device, const uint8_t* gcz_body, int64_t body_len, int64_t text_len,
                   const uint8_t* gcx_body, int64_t gcx_len, gcz_index** out) {
    clear_error();
    DeviceCtx* ctx = nullptr;
    GCZ_TRY(get_ctx(device, &ctx));
    return guarded([&] { return open_block(ctx, gcz_body, body_len, text_len, gcx_body, gcx_len, out); });
}
void gcz_close_block(gcz_index* idx) { close_block(idx); }

int gcz_text_length(const gcz_index* idx, int64_t* out) { if (!idx || !out) return fail(GCZ_E_ARG, "null argument"); *out = idx->n; return GCZ_OK; }
int gcz_sampling_factor(const gcz_index* idx, int32_t* out) { if (!idx || !out) return fail(GCZ_E_ARG, "null argument"); *out = idx->sampling_factor; return GCZ_OK; }
int gcz_num_strings(const gcz_index* idx, int32_t* out) { if (!idx || !out) return fail(GCZ_E_ARG, "null argument"); *out = (int32_t)idx->e.size(); return GCZ_OK; }
int gcz_string_ends(const gcz_index* idx, int64_t* e) {
    if (!idx || !e) return fail(GCZ_E_ARG, "null argument");
    std::memcpy(e, idx->e.data(), idx->e.size() * sizeof(int64_t));
    return GCZ_OK;
}
int gcz_c_array(const gcz_index* idx, int64_t c[256]) {
    if (!idx || !c) return fail(GCZ_E_ARG, "null argument");
    std::memcpy(c, idx->c, sizeof(idx->c));
    return GCZ_OK;
}

int gcz_count_batch(gcz_index* idx, const uint8_t* pats, const int64_t* pat_off, int64_t n_pats, int64_t* sp, int64_t* ep) {
    clear_error();
    return guarded([&] { return count_batch(idx, pats, pat_off, n_pats, sp, ep); });
}
int gcz_locate_rows(gcz_index* idx, const int64_t* rows, int64_t n_rows, int64_t* positions) {
    clear_error();
    return guarded([&] { return locate_rows(idx, rows, n_rows, positions); });
}
int gcz_find_batch(gcz_index* idx, const uint8_t* pats, const int64_t* pat_off, int64_t n_pats,
                   int64_t* per_string_counts, int64_t** positions, int64_t** pos_off) {
    clear_error();
    return guarded([&] { return find_batch(idx, pats, pat_off, n_pats, per_string_counts, positions, pos_off); });
}
int gcz_extract(gcz_index* idx, int32_t nstr, int64_t from, uint8_t* out, int64_t cap, int64_t* written) {
    clear_error();
    return guarded([&] { return extract(idx, nstr, from, out, cap, written); });
}
void gcz_free(void* p) { std::free(p); }

int gcz_count_multi(gcz_index* const* blocks, int32_t n_blocks, const uint8_t* pats, const int64_t* pat_off, int64_t n_pats, int64_t* totals) {
    clear_error();
    return guarded([&] { return count_multi(blocks, n_blocks, pats, pat_off, n_pats, totals); });
}
int gcz_count_stats(gcz_index* const* blocks, int32_t n_blocks, const uint8_t* pats, const int64_t* pat_off, int64_t n_pats, gcz_query_stats* out) {
    clear_error();
    return guarded([&] { return count_stats(blocks, n_blocks, pats, pat_off, n_pats, out); });
}
int gcz_last_query_stats(gcz_query_stats* out) { return last_query_stats(out); }
int gcz_find_multi(gcz_index* const* blocks, int32_t n_blocks, const uint8_t* pats, const int64_t* pat_off, int64_t n_pats, gcz_hits* out) {
    clear_error();
    return guarded([&] { return find_multi(blocks, n_blocks, pats, pat_off, n_pats, out); });
}
void gcz_hits_free(gcz_hits* hits) {
    if (!hits) return;
    std::free(hits->block_off); std::free(hits->pattern); std::free(hits->string); std::free(hits->position);
    std::memset(hits, 0, sizeof(*hits));
}

// ---- stage hooks ---------------------------------------------------------------------------------------------------
int gcz_dbg_set_find_chunk(int64_t occurrences) { set_find_chunk(occurrences); return GCZ_OK; }
int gcz_dbg_sort_pairs(int device, uint64_t* keys, uint32_t* vals, int64_t n, int32_t begin_bit, int32_t end_bit) {
    clear_error();
    DeviceCtx* ctx = nullptr;
    GCZ_TRY(get_ctx(device, &ctx));
    std::lock_guard<std::mutex> lock(ctx->mu);
    GCZ_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = stream_of(ctx);
    ctx->arena.reset();
    const size_t need = (size_t)n * 24 + radix_sort_temp_bytes(n) + (1 << 20);
    if (ctx->arena.capacity < need) GCZ_TRY(ctx->arena.reserve(need));
    RadixBuffers b;
    b.keys[0] = ctx->arena.get<uint64_t>((size_t)n); b.keys[1] = ctx->arena.get<uint64_t>((size_t)n);
    if (vals) { b.vals[0] = ctx->arena.get<uint32_t>((size_t)n); b.vals[1] = ctx->arena.get<uint32_t>((size_t)n); }
    void* tmp = ctx->arena.raw(radix_sort_temp_bytes(n));
    if (!b.keys[0] || !b.keys[1] || (vals && (!b.vals[0] || !b.vals[1])) || !tmp) return fail(GCZ_E_NOMEM, "sort workspace");
    GCZ_CUDA(cudaMemcpyAsync(b.keys[0], keys, (size_t)n * 8, cudaMemcpyDefault, st));
    if (vals) GCZ_CUDA(cudaMemcpyAsync(b.vals[0], vals, (size_t)n * 4, cudaMemcpyDefault, st));
    SortStats ss;
    GCZ_TRY(radix_sort_pairs(ctx, st, b, n, begin_bit, end_bit, tmp, &ss));
    GCZ_CUDA(cudaMemcpyAsync(keys, b.keys[b.cur], (size_t)n * 8, cudaMemcpyDefault, st));
    if (vals) GCZ_CUDA(cudaMemcpyAsync(vals, b.vals[b.cur], (size_t)n * 4, cudaMemcpyDefault, st));
    GCZ_CUDA(cudaStreamSynchronize(st));
    ss.resolve();                                   // readable through gcz_last_build_timing (tuning aid)
    std::memset(&t_timing, 0, sizeof(t_timing));
    t_timing.radix_ms = ss.ms; t_timing.radix_launches = ss.passes; t_timing.radix_elements = ss.elements;
    return GCZ_OK;
}

int gcz_dbg_suffix_array(int device, const uint8_t* text, int64_t n, int32_t* sa) {
    clear_error();
    if (!text || !sa || n <= 0) return fail(GCZ_E_ARG, "null argument");
    int64_t counts[256];
    GCZ_TRY(count_symbols(device, text, n, counts));
    DeviceCtx* ctx = nullptr;
    GCZ_TRY(get_ctx(device, &ctx));
    std::lock_guard<std::mutex> lock(ctx->mu);
    GCZ_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = stream_of(ctx);
    ctx->arena.reset();
    const size_t need = (size_t)n * 5 + suffix_sort_workspace_bytes(n) + (1 << 20);
    if (ctx->arena.capacity < need) GCZ_TRY(ctx->arena.reserve(need));
    uint8_t* d_text = ctx->arena.get<uint8_t>((size_t)n + 64);
    uint32_t* d_sa = ctx->arena.get<uint32_t>((size_t)n);
    if (!d_text || !d_sa) return fail(GCZ_E_NOMEM, "suffix array workspace");
    GCZ_CUDA(cudaMemcpyAsync(d_text, text, (size_t)n, cudaMemcpyDefault, st));
    GCZ_TRY(suffix_sort(ctx, st, d_text, n, counts, d_sa, ctx->arena, nullptr));
    GCZ_CUDA(cudaMemcpyAsync(sa, d_sa, (size_t)n * 4, cudaMemcpyDefault, st));
    GCZ_CUDA(cudaStreamSynchronize(st));
    return GCZ_OK;
}

int gcz_dbg_ranked_vector(int device, const uint8_t* bits, int64_t len, uint8_t* out) {
    clear_error();
    if (!bits || !out || len <= 0) return fail(GCZ_E_ARG, "null argument");
    DeviceCtx* ctx = nullptr;
    GCZ_TRY(get_ctx(device, &ctx));
    std::lock_guard<std::mutex> lock(ctx->mu);
    GCZ_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = stream_of(ctx);
    ctx->arena.reset();
    const size_t need = (size_t)len * 2 + ((size_t)len / 65536 + 2) * 8200 + (1 << 20);
    if (ctx->arena.capacity < need) GCZ_TRY(ctx->arena.reserve(need));
    const int64_t nb = ranked_bytes(len);
    uint8_t* d_bits = ctx->arena.get<uint8_t>((size_t)len + 64);
    uint8_t* d_out = ctx->arena.get<uint8_t>((size_t)nb + 64);
    if (!d_bits || !d_out) return fail(GCZ_E_NOMEM, "ranked vector workspace");
    GCZ_CUDA(cudaMemcpyAsync(d_bits, bits, (size_t)len, cudaMemcpyDefault, st));
    GCZ_TRY(ranked_vector_from_bits(ctx, st, d_bits, len, d_out, ctx->arena));
    GCZ_CUDA(cudaMemcpyAsync(out, d_out, (size_t)nb, cudaMemcpyDefault, st));
    GCZ_CUDA(cudaStreamSynchronize(st));
    return GCZ_OK;
}

int gcz_dbg_index_wavelet_tree(int device, const int32_t* vals, int64_t m, uint8_t* out) {
    clear_error();
    if (!vals || !out || m <= 0) return fail(GCZ_E_ARG, "null argument");
    DeviceCtx* ctx = nullptr;
    GCZ_TRY(get_ctx(device, &ctx));
    std::lock_guard<std::mutex> lock(ctx->mu);
    GCZ_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = stream_of(ctx);
    ctx->arena.reset();
    const int levels = 64 - __builtin_clzll((uint64_t)m);
    const int64_t nb = ranked_bytes(m) * levels;
    const size_t need = (size_t)m * 16 + (size_t)nb + ((size_t)m / 65536 + 2) * 8200 * levels + (1 << 20);
    if (ctx->arena.capacity < need) GCZ_TRY(ctx->arena.reserve(need));
    uint32_t* d_vals = ctx->arena.get<uint32_t>((size_t)m);
    uint8_t* d_out = ctx->arena.get<uint8_t>((size_t)nb + 64);
    if (!d_vals || !d_out) return fail(GCZ_E_NOMEM, "IWT workspace");
    GCZ_CUDA(cudaMemcpyAsync(d_vals, vals, (size_t)m * 4, cudaMemcpyDefault, st));
    GCZ_TRY(index_wavelet_tree_from_values(ctx, st, d_vals, m, d_out, ctx->arena));
    GCZ_CUDA(cudaMemcpyAsync(out, d_out, (size_t)nb, cudaMemcpyDefault, st));
    GCZ_CUDA(cudaStreamSynchronize(st));
    return GCZ_OK;
}

}  // extern "C"

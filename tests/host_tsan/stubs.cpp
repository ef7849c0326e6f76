// stand-ins for the CUDA runtime and the CUDA entry points: the host layer alone, under ThreadSanitizer
#include "gcz_host.h"
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <chrono>
#include <thread>
static thread_local char t_error[512];
namespace gcz {
int fail(int code, const char* fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(t_error, sizeof t_error, fmt, ap); va_end(ap); return code; }
void clear_error() { t_error[0] = 0; }
}
extern "C" {
const char* gcz_last_error(void) { return t_error; }
cudaError_t cudaHostAlloc(void** p, size_t, unsigned) { *p = nullptr; return cudaErrorNoDevice; }
cudaError_t cudaFreeHost(void*) { return cudaSuccess; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaErrorNoDevice; }
int gcz_shape_from_counts(const int64_t counts[256], gcz_shape* out) { return gcz::shape_from_counts(counts, out); }
int gcz_count_symbols(int, const uint8_t* text, int64_t n, int64_t counts[256]) {
    memset(counts, 0, 256 * 8); for (int64_t i = 0; i < n; i++) counts[text[i]]++; return 0; }
int gcz_build_block(int, const uint8_t*, int64_t n, int32_t, const gcz_shape*, uint8_t* a, int64_t al, uint8_t* b, int64_t bl, int32_t*, uint8_t*) {
    std::this_thread::sleep_for(std::chrono::microseconds(200 + n % 300));
    memset(a, (int)(n % 251), (size_t)al); memset(b, (int)(n % 241), (size_t)bl); return 0; }
int gcz_open_block(int, const uint8_t*, int64_t, int64_t, const uint8_t*, int64_t, gcz_index**) { return -1; }
void gcz_close_block(gcz_index*) {}
int gcz_num_strings(const gcz_index*, int32_t*) { return -1; }
int gcz_string_ends(const gcz_index*, int64_t*) { return -1; }
int gcz_find_batch(gcz_index*, const uint8_t*, const int64_t*, int64_t, int64_t*, int64_t**, int64_t**) { return -1; }
int gcz_extract(gcz_index*, int32_t, int64_t, uint8_t*, int64_t, int64_t*) { return -1; }
void gcz_free(void* p) { free(p); }
}

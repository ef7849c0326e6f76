// Internal host-side declarations shared by the translation units of libgcz_b200.so.
#pragma once

#include "../../include/gcz.h"

#include <cstdint>
#include <exception>
#include <new>
#include <string>
#include <vector>

namespace gcz {

// error reporting: sets the thread-local message behind gcz_last_error() and returns `code`
int fail(int code, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
void clear_error();

// no C++ exception leaves the library: memory exhaustion and anything else unexpected become status codes
template <class F>
auto guarded(F&& body) -> decltype(body()) {
    try {
        return body();
    } catch (const std::bad_alloc&) {
        return fail(GCZ_E_NOMEM, "out of host memory");
    } catch (const std::exception& ex) {
        return fail(GCZ_E_INTERNAL, "unexpected: %s", ex.what());
    }
}

// shape.cpp — byte-defining host code (no CUDA)
bool    deflate_code(const std::vector<int64_t>& weights, int max_bits, std::vector<int>& len, std::vector<int>& code);
int     shape_from_counts(const int64_t counts[256], gcz_shape* s);
int64_t shape_write(const gcz_shape* s, uint8_t* out, int64_t cap);
int     shape_read(const uint8_t* body, int64_t body_len, gcz_shape* s);
int64_t ranked_bytes(int64_t len_bits);
int64_t index_size(int64_t n, int sampling_factor);

}  // namespace gcz

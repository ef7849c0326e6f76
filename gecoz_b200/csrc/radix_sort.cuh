// Device radix sort used by the suffix sorter and by the locate post-processing.
#pragma once

#include "device.cuh"

#include <vector>

namespace gcz {

struct RadixBuffers {
    uint64_t* keys[2] = { nullptr, nullptr };
    uint32_t* vals[2] = { nullptr, nullptr };   // both null: keys only
    int cur = 0;                                  // which buffer holds the data (flips every digit pass)
};

// Packed prefix keys of a text: key(i) = sum_j code[i + j] * radix^(k - 1 - j), j < k; code 0 = past the end.
struct KeyCoder {
    uint64_t radix;
    uint64_t top;                               // radix^(k - 1)
    int      k;
};
constexpr int kMaxKeySymbols = 64;              // radix 2 (one symbol + end marker) in 64 bits

// When given to radix_sort_pairs, the (key, value) input arrays are never materialised: the histogram pass and
// the first digit pass compute key(i) from the text, and value(i) = i | code[i - 1] << carry_shift
// (code[-1] = code[n - 1]; carry_shift 0 = plain positions).  The histogram pass also reports the first and
// last positions of every run of >= k equal symbols (mark = 2 * position, + 1 for a last position).
struct TextKeySource {
    const uint8_t* text = nullptr;
    int64_t        n = 0;
    const uint8_t* code_of = nullptr;           // device, 256 entries
    KeyCoder       coder{};
    int            carry_shift = 0;
    uint64_t*      run_marks = nullptr;         // device, run_mark_cap entries (optional)
    unsigned*      run_mark_count = nullptr;
    unsigned       run_mark_cap = 0;
};

struct SortStats {
    int64_t passes = 0;
    int64_t elements = 0;
    float   ms = 0;                       // device time inside the digit passes (after resolve())
    std::vector<cudaEvent_t> events;      // start/stop pairs, one per digit pass
    std::vector<int64_t> pass_elements;   // per digit pass: pairs moved (negative: the pass read the text, not arrays)
    std::vector<float>   pass_ms;         // per digit pass: device time (after resolve())
    void resolve();                       // call after the stream has been synchronised
};

size_t radix_sort_temp_bytes(int64_t n);
int    radix_sort_passes(int bits);         // digit passes a sort of `bits` key bits takes (8-bit digits unless GCZ_SORT_VARIANT says otherwise)

// Sorts bits [begin_bit, end_bit) of the keys, stable, ascending.  Result is in b.keys[b.cur] / b.vals[b.cur].
// With `src` the input is the text (see TextKeySource; begin_bit must be 0 and values are required).
int radix_sort_pairs(DeviceCtx* ctx, cudaStream_t st, RadixBuffers& b, int64_t n, int begin_bit, int end_bit,
                     void* temp, SortStats* stats, const TextKeySource* src = nullptr);

}  // namespace gcz

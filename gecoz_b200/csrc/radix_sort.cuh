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

struct SortStats {
    int64_t passes = 0;
    int64_t elements = 0;
    float   ms = 0;                       // device time inside the digit passes (after resolve())
    std::vector<cudaEvent_t> events;      // start/stop pairs, one per digit pass
    void resolve();                       // call after the stream has been synchronised
};

size_t radix_sort_temp_bytes(int64_t n);

// Sorts bits [begin_bit, end_bit) of the keys, stable, ascending.  Result is in b.keys[b.cur] / b.vals[b.cur].
int radix_sort_pairs(DeviceCtx* ctx, cudaStream_t st, RadixBuffers& b, int64_t n, int begin_bit, int end_bit,
                     void* temp, SortStats* stats);

}  // namespace gcz

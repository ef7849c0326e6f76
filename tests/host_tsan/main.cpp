// Drives the native host layer (FASTA scan, threaded assembly, gcz_index_fasta with pooled buffers and worker threads)
// with the stand-in engine of stubs.cpp; built with -fsanitize=thread by tests/test_host_cpu.py.
#include "gcz_file.h"

#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

int main(int argc, char** argv) {
    const char* out = argc > 1 ? argv[1] : "/tmp/gcz_tsan_out.gcz";
    std::string fa;
    unsigned s = 12345;
    const int big = 20000000;
    for (int r = 0; r < 40; r++) {
        fa += ">seq" + std::to_string(r) + "\n";
        const int len = r == 0 ? big : 300000 - 977 * r;
        for (int i = 0; i < len; i++) {
            s = s * 1103515245u + 12345u;
            fa.push_back("ACGT"[(s >> 16) & 3]);
            if (i % 60 == 59) fa.push_back('\n');
        }
        fa.push_back('\n');
    }
    setenv("GCZ_FASTA_SLICE", "4096", 1);                      // threaded scan steps on small records too
    setenv("GCZ_HOST_THREADS", "4", 1);
    for (int rep = 0; rep < 2; rep++) {
        gcz_fasta* f = nullptr;
        if (gcz_fasta_open_buffer(reinterpret_cast<const uint8_t*>(fa.data()), (int64_t)fa.size(), &f) != 0) return 1;
        if (gcz_fasta_count(f) != 40) return 3;
        int devs[4] = { 0, 1, 2, 3 };
        gcz_index_report report;
        const int rc = gcz_index_fasta(f, out, nullptr, 32, rep == 0 ? 1 : 4, devs, nullptr, &report);
        std::printf("rc %d blocks %lld\n", rc, (long long)report.blocks);
        std::vector<uint8_t> seq((size_t)big + 1);
        if (gcz_fasta_read(f, 0, seq.data(), big + 1) != 0) return 4;
        gcz_fasta_close(f);
        if (rc != 0) return 2;
    }
    return 0;
}

// Native host layer above the C ABI (include/gcz_file.h): FASTA records, block planning, the .gcz/.gcx writer and
// reader.  C++ counterpart of the reference's host classes for this path — every function cites the Java lines it
// follows.  No CUDA kernels here; the per-block device work goes through a gcz_engine (default: this library's
// gcz_count_symbols / gcz_build_block).
#include "../../include/gcz_file.h"
#include "gcz_host.h"

#include <cuda_runtime.h>

#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include <dlfcn.h>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <exception>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#define GCZ_TRY_HOST(expr)                  \
    do {                                    \
        int rc__ = (expr);                  \
        if (rc__ != GCZ_OK) return rc__;    \
    } while (0)

namespace gcz {
namespace {

// ---- read-only file mapping ---------------------------------------------------------------------------------------
struct MappedFile {
    const uint8_t* data = nullptr;
    int64_t size = 0;
    int fd = -1;
    int open_ro(const char* path) {
        fd = ::open(path, O_RDONLY);
        if (fd < 0) return fail(GCZ_E_ARG, "cannot open %s", path);
        struct stat st;
        if (fstat(fd, &st) != 0) return fail(GCZ_E_ARG, "cannot stat %s", path);
        size = (int64_t)st.st_size;
        if (size > 0) {
            void* p = mmap(nullptr, (size_t)size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (p == MAP_FAILED) return fail(GCZ_E_NOMEM, "cannot map %s", path);
            data = static_cast<const uint8_t*>(p);
        }
        return GCZ_OK;
    }
    ~MappedFile() {
        if (data) munmap(const_cast<uint8_t*>(data), (size_t)size);
        if (fd >= 0) ::close(fd);
    }
};

// ---- FastaIterator (lazy) ------------------------------------------------------------------------------------------
struct FastaRecord {
    std::string header;
    int64_t position = 0, length = 0;
    int64_t span_end = 0;        // the sequence's characters are the non-CR/LF bytes of [position, span_end)
    bool multiline = false;
    // long sequences: (file offset, CR/LF bytes of [position, offset)) at the slice starts of the threaded scan, ascending —
    // the assembly threads start there without counting again
    std::vector<std::pair<int64_t, int64_t>> marks;
};

// first byte of [p, end) that is one of the two, or end
template <char A, char B>
inline const uint8_t* find_either(const uint8_t* p, const uint8_t* end) {
#if defined(__SSE2__)
    const __m128i a = _mm_set1_epi8(A), b = _mm_set1_epi8(B);
    while (p + 16 <= end) {
        const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i*>(p));
        const int m = _mm_movemask_epi8(_mm_or_si128(_mm_cmpeq_epi8(v, a), _mm_cmpeq_epi8(v, b)));
        if (m) return p + __builtin_ctz((unsigned)m);
        p += 16;
    }
#endif
    while (p < end && *p != (uint8_t)A && *p != (uint8_t)B) p++;
    return p;
}
inline const uint8_t* find_eol(const uint8_t* p, const uint8_t* end) { return find_either<'\r', '\n'>(p, end); }

// number of CR and LF bytes in [p, end)
inline int64_t count_eol(const uint8_t* p, const uint8_t* end) {
    int64_t c = 0;
#if defined(__SSE2__)
    const __m128i a = _mm_set1_epi8('\r'), b = _mm_set1_epi8('\n');
    while (p + 16 <= end) {
        const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i*>(p));
        c += __builtin_popcount((unsigned)_mm_movemask_epi8(_mm_or_si128(_mm_cmpeq_epi8(v, a), _mm_cmpeq_epi8(v, b))));
        p += 16;
    }
#endif
    for (; p < end; p++) c += (*p == '\r') | (*p == '\n');
    return c;
}

// The character-level state machine of fasta/FastaIterator.java:39-127 (hasNext :40-69, next :72-127): which byte opens a
// record, what counts as sequence, how FASTQ qualities are skipped — all of it decides which bytes end up in the index.
// scan_fasta_literal reads one character at a time exactly like the Java code; scan_fasta is the same machine with its
// three run loops (skip to the next record, one sequence line, one quality line) replaced by vector searches for the byte
// that ends the run.  GCZ_FASTA_LITERAL=1 selects the literal one (the tests compare the two on adversarial inputs).
void scan_fasta_literal(const uint8_t* buf, int64_t size, std::vector<FastaRecord>& out) {
    int64_t p = 0;
    auto rd = [&]() -> int { return p < size ? (int)buf[p++] : -1; };
    int ch = '\r';                                          // FastaIterator(InputStream, boolean) :33
    int64_t position = 0;
    while (true) {
        while (ch >= 0 && ch != '>' && ch != '@') { ch = rd(); position++; }          // hasNext :46-49
        if (ch < 0) break;
        FastaRecord rec;
        while ((ch = rd()) >= 0 && ch != '\n') {                                       // :55-61
            position++;
            if (ch != '\r') rec.header.push_back((char)ch);
        }
        position++;
        int lines = 0;                                                                 // next :81-96
        int64_t length = 0, posnew = position;
        do {
            if (ch >= 0 && ch != '\r' && ch != '\n') {
                lines++;
                do { posnew++; length++; } while ((ch = rd()) >= 0 && ch != '\r' && ch != '\n');
            }
            posnew++;
        } while ((ch = rd()) >= 0 && ch != '>' && ch != '@' && ch != '+');
        rec.span_end = ch >= 0 ? p - 1 : size;
        if (ch == '+') {                                                               // skip qualities :98-113
            int qlines = -1;
            int64_t qlength = 0;
            do {
                while ((ch = rd()) >= 0 && ch != '\r' && ch != '\n') { qlength++; posnew++; }
                posnew++;
                qlines++;
            } while (qlength < length && qlines < lines);
        }
        rec.position = position;
        rec.length = length;
        rec.multiline = lines > 1;
        out.push_back(std::move(rec));
        position = posnew;
    }
}

// One sequence, as the machine's loop :81-96 sees it, without going line by line.  Every test for the character that ends
// a sequence ('>', '@', '+') is made on the byte that follows a CR or LF (the first one follows the header's LF), so the
// sequence region is [from, k) with k the first such byte whose predecessor is CR/LF, or the end of the input; `length` is
// the number of non-CR/LF bytes in the region and `lines` the number of maximal runs of them.
struct SliceScan { int64_t hit = -1, eol = 0, runs = 0; };      // hit: index of the ending byte in the slice, counts before it

int host_threads();

// [a, b) with a >= 1 (buf[a - 1] is read)
SliceScan scan_slice(const uint8_t* buf, int64_t a, int64_t b) {
    SliceScan r;
    int64_t i = a;
#if defined(__SSE2__)
    const __m128i cr = _mm_set1_epi8('\r'), lf = _mm_set1_epi8('\n'), gt = _mm_set1_epi8('>'), at = _mm_set1_epi8('@'), plus = _mm_set1_epi8('+');
    const __m128i zero = _mm_setzero_si128();
    __m128i eol_acc = zero, run_acc = zero;
    for (; i + 16 <= b; i += 16) {
        const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i*>(buf + i));
        const __m128i w = _mm_loadu_si128(reinterpret_cast<const __m128i*>(buf + i - 1));          // the predecessors
        const __m128i eol_v = _mm_or_si128(_mm_cmpeq_epi8(v, cr), _mm_cmpeq_epi8(v, lf));
        const __m128i eol_w = _mm_or_si128(_mm_cmpeq_epi8(w, cr), _mm_cmpeq_epi8(w, lf));
        const __m128i special = _mm_or_si128(_mm_or_si128(_mm_cmpeq_epi8(v, gt), _mm_cmpeq_epi8(v, at)), _mm_cmpeq_epi8(v, plus));
        if (_mm_movemask_epi8(_mm_and_si128(special, eol_w))) break;                               // finish this block one byte at a time
        eol_acc = _mm_add_epi64(eol_acc, _mm_sad_epu8(_mm_sub_epi8(zero, eol_v), zero));           // 0xFF -> 1 per byte, summed
        run_acc = _mm_add_epi64(run_acc, _mm_sad_epu8(_mm_sub_epi8(zero, _mm_andnot_si128(eol_v, eol_w)), zero));
    }
    uint64_t lanes[2];
    _mm_storeu_si128(reinterpret_cast<__m128i*>(lanes), eol_acc);
    r.eol = (int64_t)(lanes[0] + lanes[1]);
    _mm_storeu_si128(reinterpret_cast<__m128i*>(lanes), run_acc);
    r.runs = (int64_t)(lanes[0] + lanes[1]);
#endif
    for (; i < b; i++) {
        const uint8_t c = buf[i], q = buf[i - 1];
        const bool eol_c = c == '\r' || c == '\n', eol_q = q == '\r' || q == '\n';
        if (eol_q && (c == '>' || c == '@' || c == '+')) { r.hit = i; return r; }
        r.eol += eol_c;
        r.runs += !eol_c && eol_q;
    }
    return r;
}

struct Region { int64_t end, length, lines; std::vector<std::pair<int64_t, int64_t>> marks; };

// slice: bytes per thread and step; the first `slice` bytes of a region are scanned on the calling thread, so that reads and
// contigs (a file may hold millions) never start a thread
Region scan_region(const uint8_t* buf, int64_t from, int64_t size, int64_t slice, int threads) {
    int64_t eol = 0, runs = 0, at = from, hit = -1;
    std::vector<std::pair<int64_t, int64_t>> marks;
    while (at < size && hit < 0) {
        const bool first = at - from < slice;
        const int64_t step = first || threads <= 1 ? std::min<int64_t>(slice / 8 + 1, size - at) : std::min<int64_t>(slice * threads, size - at);
        if (first || threads <= 1) {
            const SliceScan r = scan_slice(buf, at, at + step);
            eol += r.eol; runs += r.runs; hit = r.hit;
        } else {
            std::vector<SliceScan> part((size_t)threads);
            std::vector<std::thread> pool;
            for (int t = 0; t < threads; t++)
                pool.emplace_back([&, t] { part[(size_t)t] = scan_slice(buf, at + step * t / threads, at + step * (t + 1) / threads); });
            for (std::thread& th : pool) th.join();
            for (int t = 0; t < threads && hit < 0; t++) {
                marks.emplace_back(at + step * t / threads, eol);
                eol += part[(size_t)t].eol; runs += part[(size_t)t].runs; hit = part[(size_t)t].hit;
            }
        }
        at += step;
    }
    const int64_t end = hit >= 0 ? hit : size;
    while (!marks.empty() && marks.back().first >= end) marks.pop_back();
    return Region{ end, (end - from) - eol, runs, std::move(marks) };
}

void scan_fasta(const uint8_t* buf, int64_t size, std::vector<FastaRecord>& out) {
    const char* lit = std::getenv("GCZ_FASTA_LITERAL");
    if (lit && lit[0] == '1') { scan_fasta_literal(buf, size, out); return; }
    const uint8_t* const end = buf + size;
    const char* slice_env = std::getenv("GCZ_FASTA_SLICE");               // the tests shrink it to reach the threaded steps
    const int64_t slice = slice_env && std::atoll(slice_env) > 0 ? (int64_t)std::atoll(slice_env) : (int64_t)8 << 20;
    const int threads = host_threads();
    int64_t p = 0;
    auto rd = [&]() -> int { return p < size ? (int)buf[p++] : -1; };
    // the reads of one run: everything up to the byte that ends it (consumed too), or up to the end of the input (where the
    // machine's last read returns -1 and consumes nothing); returns how many bytes the run itself holds
    auto run_to = [&](const uint8_t* stop, int& ch) -> int64_t {
        const int64_t k = stop - buf, n = k - p;
        if (k < size) { ch = buf[k]; p = k + 1; } else { ch = -1; p = size; }
        return n;
    };
    int ch = '\r';
    int64_t position = 0;
    while (true) {
        if (ch >= 0 && ch != '>' && ch != '@') position += run_to(find_either<'>', '@'>(buf + p, end), ch) + 1;    // hasNext :46-49
        if (ch < 0) break;
        FastaRecord rec;
        while ((ch = rd()) >= 0 && ch != '\n') {                                       // :55-61
            position++;
            if (ch != '\r') rec.header.push_back((char)ch);
        }
        position++;
        rec.position = position;
        if (ch < 0) {                                                                  // the input ends inside the header line
            rec.span_end = size;
            out.push_back(std::move(rec));
            break;
        }
        const Region seq = scan_region(buf, position, size, slice, threads);                           // next :81-96 (p == position here)
        rec.span_end = seq.end;
        rec.length = seq.length;
        rec.multiline = seq.lines > 1;
        rec.marks = seq.marks;
        int64_t posnew = seq.end + 1;                                                  // == p after the ending byte was read
        run_to(buf + seq.end, ch);
        if (ch == '+') {                                                               // skip qualities :98-113
            int64_t qlines = -1, qlength = 0;
            do {
                const int64_t n = run_to(find_eol(buf + p, end), ch);
                qlength += n;
                posnew += n + 1;
                qlines++;
            } while (qlength < seq.length && qlines < seq.lines);
        }
        out.push_back(std::move(rec));
        position = posnew;
    }
}

}  // namespace
}  // namespace gcz

struct gcz_fasta {
    std::unique_ptr<gcz::MappedFile> file;
    std::vector<uint8_t> inflated;                 // gzipped input: the decompressed bytes
    const uint8_t* data = nullptr;
    int64_t size = 0;
    std::vector<gcz::FastaRecord> records;
};

namespace gcz {
namespace {

// how many host threads the bulk loops (record scan, sequence assembly) may use: GCZ_HOST_THREADS, else the machine's, at most 16
int host_threads() {
    if (const char* e = std::getenv("GCZ_HOST_THREADS")) {
        const int v = std::atoi(e);
        if (v > 0) return std::min(v, 64);
    }
    const unsigned hw = std::thread::hardware_concurrency();
    return (int)std::min(16u, std::max(1u, hw));
}

// the non-CR/LF bytes of [src, end) -> out, at most `want` of them; returns how many
int64_t strip_eol(const uint8_t* src, const uint8_t* end, uint8_t* out, int64_t want) {
    int64_t i = 0;
    while (src < end && i < want) {
        const uint8_t* e = find_eol(src, end);
        const int64_t n = std::min<int64_t>(e - src, want - i);
        std::memcpy(out + i, src, (size_t)n);
        i += n;
        src = e + 1;
    }
    return i;
}

// the same over `threads` pieces of the source: count per piece, prefix, copy per piece
int64_t strip_eol_parallel(const uint8_t* src, const uint8_t* end, uint8_t* out, int64_t want, int threads) {
    const int64_t bytes = end - src;
    threads = (int)std::min<int64_t>(threads, bytes >> 22);                  // >= 4 MiB of source per thread
    if (threads <= 1) return strip_eol(src, end, out, want);
    std::vector<const uint8_t*> cut((size_t)threads + 1);
    for (int t = 0; t <= threads; t++) cut[(size_t)t] = src + bytes * t / threads;
    std::vector<int64_t> first((size_t)threads + 1, 0);
    {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; t++)
            pool.emplace_back([&, t] { first[(size_t)t + 1] = (cut[(size_t)t + 1] - cut[(size_t)t]) - count_eol(cut[(size_t)t], cut[(size_t)t + 1]); });
        for (std::thread& th : pool) th.join();
    }
    for (int t = 0; t < threads; t++) first[(size_t)t + 1] += first[(size_t)t];
    if (first[(size_t)threads] > want) return strip_eol(src, end, out, want);  // a buffer shorter than the sequence: in order
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++)
        pool.emplace_back([&, t] { strip_eol(cut[(size_t)t], cut[(size_t)t + 1], out + first[(size_t)t], first[(size_t)t + 1] - first[(size_t)t]); });
    for (std::thread& th : pool) th.join();
    return first[(size_t)threads];
}

// FastaFileReader.read  fasta/FastaFileReader.java:109-160 (plain file): one-line sequences are `length` raw bytes
// at `position`; multi-line ones are read from there with CR/LF dropped until `length` bytes are in — which are the
// non-CR/LF bytes of [position, span_end) (scan_fasta)
int64_t read_sequence(const gcz_fasta* f, const FastaRecord& r, uint8_t* out, int64_t cap) {
    const int64_t want = std::min(r.length, cap);
    if (!r.multiline) {
        const int64_t avail = std::max<int64_t>(0, std::min(want, f->size - r.position));
        std::memcpy(out, f->data + r.position, (size_t)avail);
        return avail;
    }
    const int threads = host_threads();
    if (want == r.length && threads > 1 && r.marks.size() >= 2) {
        // pieces: [position, marks[0]) and one per mark; piece i starts at out[(offset - position) - CR/LF bytes before it]
        const int64_t pieces = (int64_t)r.marks.size() + 1;
        auto piece = [&](int64_t i, int64_t* begin, int64_t* end, int64_t* out_at) {
            *begin = i == 0 ? r.position : r.marks[(size_t)i - 1].first;
            *end = i + 1 < pieces ? r.marks[(size_t)i].first : r.span_end;
            *out_at = i == 0 ? 0 : (*begin - r.position) - r.marks[(size_t)i - 1].second;
        };
        std::vector<std::thread> pool;
        const int workers = (int)std::min<int64_t>(threads, pieces);
        for (int t = 0; t < workers; t++)
            pool.emplace_back([&, t] {
                for (int64_t i = t; i < pieces; i += workers) {
                    int64_t b, e, o, nb, ne, no;
                    piece(i, &b, &e, &o);
                    int64_t stop = r.length;
                    if (i + 1 < pieces) { piece(i + 1, &nb, &ne, &no); stop = no; }
                    strip_eol(f->data + b, f->data + e, out + o, stop - o);
                }
            });
        for (std::thread& th : pool) th.join();
        return want;
    }
    return strip_eol_parallel(f->data + r.position, f->data + r.span_end, out, want, threads);
}

// ---- GecoIndex: one block per sequence, greedy merge, file order  tools/GecoIndex.java:57-98 ------------------------
inline int32_t java_int(int64_t v) { return (int32_t)(uint32_t)(uint64_t)v; }   // int arithmetic wraps

struct Seq { int64_t length; const char* header; int64_t id; };

// TFastaSequence.compareTo  fasta/TFastaSequence.java:46-52: longer first, then header (String.compareTo)
int cmp_seq(const Seq& a, const Seq& b) {
    if (a.length != b.length) return a.length > b.length ? -1 : 1;
    const int c = std::strcmp(a.header, b.header);          // bytes < 0x80: same order as UTF-16 code units
    return c < 0 ? -1 : (c > 0 ? 1 : 0);
}

struct Block {                                               // fmt/GecozRefBlock.java:38-71
    std::vector<Seq> sequences;                              // TreeSet<TFastaSequence>
    int32_t size = 0;
    explicit Block(const Seq& s) : sequences{ s }, size(java_int(s.length + 1)) {}
    void add(const Seq& s) {                                 // :45-48
        size_t lo = 0, hi = sequences.size();
        bool dup = false;
        while (lo < hi) {
            const size_t mid = (lo + hi) / 2;
            const int c = cmp_seq(s, sequences[mid]);
            if (c == 0) { dup = true; break; }
            if (c < 0) hi = mid; else lo = mid + 1;
        }
        if (!dup) sequences.insert(sequences.begin() + (long)lo, s);
        size = java_int((int64_t)size + s.length + 1);
    }
};

int cmp_block(const Block& a, const Block& b) {              // compareTo :62-69
    if (a.size != b.size) return a.size > b.size ? 1 : -1;
    return cmp_seq(a.sequences[0], b.sequences[0]);
}

template <class Cmp>
bool tree_add(std::vector<Block>& set, Block&& b, Cmp cmp) {     // TreeSet.add: equal elements are rejected
    size_t lo = 0, hi = set.size();
    while (lo < hi) {
        const size_t mid = (lo + hi) / 2;
        const int c = cmp(b, set[mid]);
        if (c == 0) return false;
        if (c < 0) hi = mid; else lo = mid + 1;
    }
    set.insert(set.begin() + (long)lo, std::move(b));
    return true;
}

std::vector<Block> plan_blocks(const std::vector<Seq>& seqs) {
    std::vector<Block> blocks;
    for (const Seq& s : seqs) tree_add(blocks, Block(s), cmp_block);
    if (blocks.empty()) return blocks;
    const int32_t max_size = blocks.back().size;                                   // :72
    while (blocks.size() > 1) {                                                     // :73-85
        Block first = std::move(blocks[0]), second = std::move(blocks[1]);
        blocks.erase(blocks.begin(), blocks.begin() + 2);
        const int32_t size = java_int((int64_t)first.size + second.size);
        if (size > 0 && size <= max_size) {
            for (const Seq& s : second.sequences) first.add(s);
            tree_add(blocks, std::move(first), cmp_block);
        } else {
            tree_add(blocks, std::move(first), cmp_block);
            tree_add(blocks, std::move(second), cmp_block);
            break;
        }
    }
    auto by_longest = [](const Block& a, const Block& b) {                          // :88-96
        if (a.sequences[0].length != b.sequences[0].length) return a.sequences[0].length > b.sequences[0].length ? -1 : 1;
        return cmp_block(a, b);
    };
    std::vector<Block> sorted;
    for (Block& b : blocks) tree_add(sorted, std::move(b), by_longest);
    return sorted;
}

// ---- headers ------------------------------------------------------------------------------------------------------------
void put_le64(uint8_t* p, int64_t v) { for (int i = 0; i < 8; i++) p[i] = (uint8_t)((uint64_t)v >> (8 * i)); }
int64_t get_le64(const uint8_t* p) { uint64_t v = 0; for (int i = 0; i < 8; i++) v |= (uint64_t)p[i] << (8 * i); return (int64_t)v; }

int64_t ref_header_length(const char* const* headers, int32_t n) {                  // getBlockHeaderLength :130-136
    int64_t len = 26;
    for (int32_t i = 0; i < n; i++) len += (int64_t)std::strlen(headers[i]) + 1;
    return len;
}

int64_t header_hash(const char* const* headers, int32_t n) {                        // getBlockHeaderHash :120-128
    uint64_t h = 1125899906842597ull;
    for (int32_t i = 0; i < n; i++) {
        for (const unsigned char* c = reinterpret_cast<const unsigned char*>(headers[i]); *c; c++) h = 31 * h + *c;
    }
    return (int64_t)h;
}

std::string ssa_path_for(const std::string& ref) {                                  // fmt/GecozFileWriter.java:97-104
    const size_t slash = ref.find_last_of('/');
    std::string dir = slash == std::string::npos ? "" : ref.substr(0, slash + 1);
    std::string name = slash == std::string::npos ? ref : ref.substr(slash + 1);
    if (name.size() >= 4 && name.compare(name.size() - 4, 4, ".gcz") == 0) name.resize(name.size() - 3);
    return dir + name + "gcx";
}

// host buffer: pinned when a CUDA device is there (full-speed DMA both ways), plain otherwise
struct HostBuffer {
    uint8_t* data = nullptr;
    size_t cap = 0;
    bool pinned = false;
    // device: the one whose context the pinning goes through (a thread that never chose one would bring up device 0's)
    HostBuffer(size_t bytes, int device) : cap(bytes ? bytes : 1) {
        if (device >= 0 && cudaSetDevice(device) != cudaSuccess) cudaGetLastError();
        if (cudaHostAlloc(reinterpret_cast<void**>(&data), cap, cudaHostAllocDefault) == cudaSuccess) { pinned = true; return; }
        cudaGetLastError();
        data = static_cast<uint8_t*>(std::malloc(cap));
    }
    ~HostBuffer() { if (pinned) cudaFreeHost(data); else std::free(data); }
    HostBuffer(const HostBuffer&) = delete;
    HostBuffer& operator=(const HostBuffer&) = delete;
};

// Buffers of one kind, reused from block to block: pinning a quarter of a gigabyte costs more than building the block.
// At most `limit` are out at a time (take() waits); returned buffers that are too small for a request are dropped.
// prefill() allocates in the background, so that the second text buffer and the first body buffers are being pinned
// while the first block is assembled.
class BufferPool {
public:
    BufferPool(size_t limit, int device) : limit_(limit), device_(device) {}
    ~BufferPool() { if (filler_.joinable()) filler_.join(); }
    void prefill(size_t count, size_t bytes) {
        {
            std::lock_guard<std::mutex> l(mu_);
            largest_ = std::max(largest_, bytes);
            pending_ += count;
        }
        filler_ = std::thread([this, count] {
            for (size_t i = 0; i < count; i++) {
                size_t want;
                { std::lock_guard<std::mutex> l(mu_); want = largest_; }
                std::unique_ptr<HostBuffer> b(new HostBuffer(want, device_));
                std::lock_guard<std::mutex> l(mu_);
                pending_--;
                if (b->data) free_.push_back(std::move(b));
                cv_.notify_all();
            }
        });
    }
    std::unique_ptr<HostBuffer> take(size_t bytes) {
        std::unique_lock<std::mutex> l(mu_);
        size_t best = 0;
        auto fits = [&] {
            best = free_.size();
            for (size_t i = 0; i < free_.size(); i++)
                if (free_[i]->cap >= bytes && (best == free_.size() || free_[i]->cap < free_[best]->cap)) best = i;
            return best < free_.size();
        };
        // a buffer the filler is about to deliver is worth waiting for only if it will be large enough
        cv_.wait(l, [&] { return out_ < limit_ && (fits() || pending_ == 0 || largest_ < bytes); });
        out_++;
        if (fits()) {
            std::unique_ptr<HostBuffer> b = std::move(free_[best]);
            free_.erase(free_.begin() + (long)best);
            return b;
        }
        free_.clear();                                      // all too small: their memory goes first
        largest_ = std::max(largest_, bytes);
        const size_t want = largest_;
        l.unlock();
        return std::unique_ptr<HostBuffer>(new HostBuffer(want, device_));
    }
    void give(std::unique_ptr<HostBuffer> b) {
        std::lock_guard<std::mutex> l(mu_);
        if (b && b->data) free_.push_back(std::move(b));
        out_--;
        cv_.notify_all();
    }
private:
    std::mutex mu_;
    std::condition_variable cv_;
    std::vector<std::unique_ptr<HostBuffer>> free_;
    std::thread filler_;
    size_t limit_, out_ = 0, largest_ = 0, pending_ = 0;
    int device_;
};

// all of buf[0, len) at file offset off
bool write_fully(int fd, const uint8_t* buf, int64_t len, int64_t off) {
    while (len > 0) {
        const ssize_t w = pwrite(fd, buf, (size_t)std::min<int64_t>(len, (int64_t)1 << 30), off);
        if (w < 0 && errno == EINTR) continue;
        if (w <= 0) return false;
        buf += w; off += w; len -= w;
    }
    return true;
}

}  // namespace
}  // namespace gcz

using namespace gcz;

namespace {
bool host_trace() { static const bool on = [] { const char* e = std::getenv("GCZ_HOST_TRACE"); return e && e[0] == '1'; }(); return on; }
double seconds_since(std::chrono::steady_clock::time_point t) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t).count(); }
}  // namespace

// =====================================================================================================================
extern "C" {

static int fasta_open_impl(const char* path, gcz_fasta** out) {
    clear_error();
    if (!path || !out) return fail(GCZ_E_ARG, "null argument");
    std::unique_ptr<gcz_fasta> f(new gcz_fasta());
    f->file.reset(new MappedFile());
    GCZ_TRY_HOST(f->file->open_ro(path));
    f->data = f->file->data;
    f->size = f->file->size;
    if (f->size >= 2 && f->data[0] == 0x1f && f->data[1] == 0x8b) {
        // FastaFileReader(Path) probes for GZIP (fasta/FastaFileReader.java:71-81) and then reads the decompressed stream
        // (lazy loading is off for gzipped files).  nova-gzip's job in the reference; zlib here — host I/O either way.
        // zlib is looked up at run time: the library itself must load on a machine without it
        void* z = dlopen("libz.so.1", RTLD_NOW | RTLD_LOCAL);
        if (!z) return fail(GCZ_E_FORMAT, "%s is gzipped and libz.so.1 is not available: decompress it and use gcz_fasta_open_buffer", path);
        auto z_open = reinterpret_cast<void* (*)(const char*, const char*)>(dlsym(z, "gzopen"));
        auto z_read = reinterpret_cast<int (*)(void*, void*, unsigned)>(dlsym(z, "gzread"));
        auto z_close = reinterpret_cast<int (*)(void*)>(dlsym(z, "gzclose"));
        if (!z_open || !z_read || !z_close) { dlclose(z); return fail(GCZ_E_FORMAT, "libz.so.1 lacks the gz* functions"); }
        void* gz = z_open(path, "rb");
        if (!gz) { dlclose(z); return fail(GCZ_E_ARG, "cannot open %s", path); }
        std::vector<uint8_t> chunk((size_t)8 << 20);
        int got;
        while ((got = z_read(gz, chunk.data(), (unsigned)chunk.size())) > 0) f->inflated.insert(f->inflated.end(), chunk.begin(), chunk.begin() + got);
        const bool bad = z_close(gz) != 0 || got < 0;           // gzclose reports a stream that ended early
        dlclose(z);
        if (bad) return fail(GCZ_E_FORMAT, "%s: corrupt gzip stream", path);
        f->file.reset();
        f->data = f->inflated.data();
        f->size = (int64_t)f->inflated.size();
    }
    scan_fasta(f->data, f->size, f->records);
    *out = f.release();
    return GCZ_OK;
}

static int fasta_open_buffer_impl(const uint8_t* data, int64_t size, gcz_fasta** out) {
    clear_error();
    if ((!data && size > 0) || size < 0 || !out) return fail(GCZ_E_ARG, "null argument");
    std::unique_ptr<gcz_fasta> f(new gcz_fasta());
    f->data = data;
    f->size = size;
    scan_fasta(data, size, f->records);
    *out = f.release();
    return GCZ_OK;
}

int64_t gcz_fasta_count(const gcz_fasta* f) { return f ? (int64_t)f->records.size() : 0; }

int gcz_fasta_record(const gcz_fasta* f, int64_t i, const char** header, int64_t* position, int64_t* length, int32_t* multiline) {
    if (!f || i < 0 || i >= (int64_t)f->records.size()) return fail(GCZ_E_ARG, "record index");
    const FastaRecord& r = f->records[(size_t)i];
    if (header) *header = r.header.c_str();
    if (position) *position = r.position;
    if (length) *length = r.length;
    if (multiline) *multiline = r.multiline ? 1 : 0;
    return GCZ_OK;
}

int gcz_fasta_read(const gcz_fasta* f, int64_t i, uint8_t* out, int64_t cap) {
    if (!f || !out || i < 0 || i >= (int64_t)f->records.size()) return fail(GCZ_E_ARG, "record index");
    const FastaRecord& r = f->records[(size_t)i];
    if (cap < r.length) return fail(GCZ_E_ARG, "buffer of %lld bytes for a sequence of %lld", (long long)cap, (long long)r.length);
    read_sequence(f, r, out, cap);
    return GCZ_OK;
}

void gcz_fasta_close(gcz_fasta* f) { delete f; }

int64_t gcz_plan_blocks(const int64_t* lengths, const char* const* headers, int64_t n, int64_t* block_of, int64_t* order_in_block) {
    clear_error();
    if (n < 0 || (n > 0 && (!lengths || !headers || !block_of || !order_in_block))) return fail(GCZ_E_ARG, "null argument");
    std::vector<Seq> seqs((size_t)n);
    for (int64_t i = 0; i < n; i++) { seqs[(size_t)i] = Seq{ lengths[i], headers[i], i }; block_of[i] = -1; order_in_block[i] = -1; }
    const std::vector<Block> blocks = plan_blocks(seqs);
    for (size_t b = 0; b < blocks.size(); b++) {
        for (size_t k = 0; k < blocks[b].sequences.size(); k++) {
            block_of[blocks[b].sequences[k].id] = (int64_t)b;
            order_in_block[blocks[b].sequences[k].id] = (int64_t)k;
        }
    }
    return (int64_t)blocks.size();
}

int64_t gcz_ref_header_length(const char* const* headers, int32_t n) { return ref_header_length(headers, n); }
int64_t gcz_header_hash(const char* const* headers, int32_t n) { return header_hash(headers, n); }

int64_t gcz_ref_header_write(const char* const* headers, int32_t n, int64_t block_size, int64_t text_len, uint8_t* out, int64_t cap) {
    clear_error();                                                                   // write(ByteBuffer) :90-101
    const int64_t len = ref_header_length(headers, n);
    if (!out || cap < len) return fail(GCZ_E_ARG, "header buffer of %lld bytes, %lld needed", (long long)cap, (long long)len);
    std::memcpy(out, "GecozBWT", 8);
    out[8] = 1;
    put_le64(out + 9, block_size);
    put_le64(out + 17, text_len);
    int64_t p = 25;
    for (int32_t i = 0; i < n; i++) {
        const size_t l = std::strlen(headers[i]);
        std::memcpy(out + p, headers[i], l);
        p += (int64_t)l;
        out[p++] = 0;
    }
    out[p++] = 0;
    return p;
}

int64_t gcz_ssa_header_write(const char* const* headers, int32_t n, int64_t index_len, uint8_t out[25]) {
    std::memcpy(out, "GecozSSA", 8);                                                 // fmt/GecozSSABlockHeader.java:69-74
    out[8] = 1;
    put_le64(out + 9, index_len);
    put_le64(out + 17, header_hash(headers, n));
    return 25;
}

// ---- writer -------------------------------------------------------------------------------------------------------------
static int index_fasta_impl(const gcz_fasta* fasta, const char* gcz_path, const char* gcx_path, int32_t sampling_rate,
                    int32_t n_devices, const int* devices, const gcz_engine* engine, gcz_index_report* report) {
    clear_error();
    if (!fasta || !gcz_path) return fail(GCZ_E_ARG, "null argument");
    if (sampling_rate <= 0 || (sampling_rate & (sampling_rate - 1)) != 0) return fail(GCZ_E_ARG, "sampling rate must be a power of two");
    const int sf = 31 - __builtin_clz((unsigned)sampling_rate);
    gcz_engine eng;
    eng.count_symbols = engine && engine->count_symbols ? engine->count_symbols : gcz_count_symbols;
    eng.build_block = engine && engine->build_block ? engine->build_block : gcz_build_block;
    std::vector<int> devs;
    for (int32_t i = 0; i < n_devices; i++) devs.push_back(devices ? devices[i] : i);
    if (devs.empty()) devs.push_back(0);
    const auto t_start = std::chrono::steady_clock::now();

    // blocks in file order (tools/GecoIndex.java:57-98)
    std::vector<Seq> seqs;
    for (size_t i = 0; i < fasta->records.size(); i++) seqs.push_back(Seq{ fasta->records[i].length, fasta->records[i].header.c_str(), (int64_t)i });
    const std::vector<Block> blocks = plan_blocks(seqs);
    if (blocks.empty()) return fail(GCZ_E_ARG, "no data found");

    const std::string ref_path = gcz_path, ssa_path = gcx_path ? std::string(gcx_path) : ssa_path_for(ref_path);
    const int ref_fd = ::open(ref_path.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644);
    const int ssa_fd = ::open(ssa_path.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644);
    if (ref_fd < 0 || ssa_fd < 0) {
        if (ref_fd >= 0) ::close(ref_fd);
        if (ssa_fd >= 0) ::close(ssa_fd);
        return fail(GCZ_E_ARG, "cannot create %s / %s", ref_path.c_str(), ssa_path.c_str());
    }

    // two tokens per device: one block being built, the next one being counted (= uploaded)
    std::mutex mu;
    std::condition_variable cv;
    std::vector<int> free_tokens;
    for (int rep = 0; rep < 2; rep++) for (int d : devs) free_tokens.push_back(d);
    const size_t all_tokens = free_tokens.size();
    int first_error = GCZ_OK;
    std::string first_message;
    std::deque<std::thread> workers;                 // oldest first; joined as new ones start, so that a file of many small
    int64_t n_blocks = 0;                            // blocks does not leave a finished thread per block behind
    int64_t ref_pos = 0, ssa_pos = 0, symbols = 0, sequences = 0;

    // One text buffer per token, sized for the largest block.  The bodies come back into buffers of a second pool and go
    // to the files with pwrite AFTER the device token is released: writing into the page cache is the slow part of the host
    // side (about 2 GB/s per thread) and must not sit between two builds.
    int64_t largest = 0;
    for (const Block& block : blocks) {
        int64_t n = 0;
        for (const Seq& s : block.sequences) n += s.length + 1;
        largest = std::max(largest, n);
    }
    BufferPool texts(all_tokens, devs[0]), bodies(all_tokens + 2, devs[0]);
    texts.prefill(std::min(all_tokens, blocks.size()), (size_t)largest);
    // the bodies of a DNA block take about 0.3 n + n / 8 + 23 levels of n / 256 bytes; other alphabets grow the pool on demand
    bodies.prefill(std::min<size_t>(2, blocks.size()), (size_t)(largest * 0.32) + (size_t)index_size(largest, sf) + 3 * 4096);

    auto release = [&](int d) { std::lock_guard<std::mutex> l(mu); free_tokens.push_back(d); cv.notify_all(); };
    auto record_error = [&](int rc) {
        std::lock_guard<std::mutex> l(mu);
        if (first_error == GCZ_OK) { first_error = rc; first_message = gcz_last_error(); }
    };

    bool aborting = false;                           // set under mu when the submitting loop dies of an exception
    try {
    for (const Block& block : blocks) {
        {
            std::lock_guard<std::mutex> l(mu);
            if (first_error != GCZ_OK) break;
        }
        int device;
        {
            std::unique_lock<std::mutex> l(mu);
            cv.wait(l, [&] { return !free_tokens.empty(); });
            device = free_tokens.front();
            free_tokens.erase(free_tokens.begin());
        }
        // writeBlock (tools/GecoIndex.java:119-146): the member sequences in block order, each followed by '\0'
        int64_t n = 0;
        for (const Seq& s : block.sequences) n += s.length + 1;
        auto t0 = std::chrono::steady_clock::now();
        std::shared_ptr<HostBuffer> text(texts.take((size_t)n).release(), [&texts](HostBuffer* b) { texts.give(std::unique_ptr<HostBuffer>(b)); });
        const double t_alloc = seconds_since(t0);
        if (!text->data) { release(device); record_error(fail(GCZ_E_NOMEM, "host buffer of %lld bytes", (long long)n)); break; }
        std::vector<const char*> hdrs;
        int64_t p = 0;
        for (const Seq& s : block.sequences) {
            read_sequence(fasta, fasta->records[(size_t)s.id], text->data + p, s.length);
            p += s.length;
            text->data[p++] = 0;
            hdrs.push_back(s.header);
        }
        const double t_read = seconds_since(t0) - t_alloc;
        // GecozFileWriter.write (fmt/GecozFileWriter.java:124-159): counts, shape, slices, headers, queue the block
        int64_t counts[256];
        std::shared_ptr<gcz_shape> shape(new gcz_shape());
        int rc = eng.count_symbols(device, text->data, n, counts);
        const double t_count = seconds_since(t0) - t_alloc - t_read;
        if (rc == GCZ_OK) rc = gcz_shape_from_counts(counts, shape.get());
        if (rc != GCZ_OK) { release(device); record_error(rc); break; }
        const int64_t hlen = ref_header_length(hdrs.data(), (int32_t)hdrs.size());
        const int64_t idx_size = index_size(n, sf);
        const int64_t my_ref = ref_pos, my_ssa = ssa_pos;
        ref_pos += hlen + shape->size;
        ssa_pos += 25 + idx_size;
        std::vector<uint8_t> hb((size_t)hlen), sb(25);
        gcz_ref_header_write(hdrs.data(), (int32_t)hdrs.size(), hlen + shape->size, n, hb.data(), hlen);
        gcz_ssa_header_write(hdrs.data(), (int32_t)hdrs.size(), idx_size, sb.data());
        symbols += n;
        sequences += (int64_t)block.sequences.size();
        if (host_trace())
            std::fprintf(stderr, "[gcz host] block n=%lld: buffer %.1f ms, assemble %.1f ms, count %.1f ms\n", (long long)n,
                         t_alloc * 1e3, t_read * 1e3, t_count * 1e3);
        // BlockWriter.run (:256-284) on its own thread; GCZ_E_NOMEM: once more when nothing else is in flight
        // (WriterPoolExecutor.afterExecute :203-226)
        n_blocks++;
        while (workers.size() >= 4 * all_tokens) { workers.front().join(); workers.pop_front(); }
        try {
            workers.emplace_back([&, device, text, shape, n, my_ref, my_ssa, hlen, idx_size, hb, sb]() mutable {
                const auto b0 = std::chrono::steady_clock::now();
                // one buffer: [.. ref header | .gcz body .. ssa header | .gcx body], both bodies on a 4 KiB boundary, so that each
                // file gets its header and body with one write
                const int64_t ref_at = (hlen + 4095) & ~(int64_t)4095;
                const int64_t ssa_at = ((ref_at + shape->size + 4095) & ~(int64_t)4095) + 4096;
                std::unique_ptr<HostBuffer> out = bodies.take((size_t)(ssa_at + idx_size));
                int rc2 = out->data ? GCZ_OK : fail(GCZ_E_NOMEM, "host buffer of %lld bytes", (long long)(ssa_at + idx_size));
                for (int attempts = 0; rc2 == GCZ_OK; ) {
                    rc2 = eng.build_block(device, text->data, n, sampling_rate, shape.get(), out->data + ref_at, shape->size, out->data + ssa_at,
                                          idx_size, nullptr, nullptr);
                    if (rc2 != GCZ_E_NOMEM || attempts++ > 0) break;
                    std::unique_lock<std::mutex> l(mu);
                    cv.wait(l, [&] { return aborting || free_tokens.size() == all_tokens - 1; });
                    if (aborting) break;
                    rc2 = GCZ_OK;
                }
                if (rc2 != GCZ_OK) record_error(rc2);
                text.reset();                                                            // the text buffer goes back with the token
                release(device);
                const double t_build = seconds_since(b0);
                if (rc2 == GCZ_OK) {
                    std::memcpy(out->data + ref_at - hlen, hb.data(), (size_t)hlen);
                    std::memcpy(out->data + ssa_at - 25, sb.data(), 25);
                    if (!write_fully(ref_fd, out->data + ref_at - hlen, hlen + shape->size, my_ref) ||
                        !write_fully(ssa_fd, out->data + ssa_at - 25, 25 + idx_size, my_ssa))
                        record_error(fail(GCZ_E_ARG, "cannot write %s / %s", ref_path.c_str(), ssa_path.c_str()));
                }
                bodies.give(std::move(out));
                if (host_trace()) std::fprintf(stderr, "[gcz host] block n=%lld: build %.1f ms, write %.1f ms\n", (long long)n, t_build * 1e3,
                                               (seconds_since(b0) - t_build) * 1e3);
            });
        } catch (const std::exception& ex) {                  // no thread could be started
            release(device);
            record_error(fail(GCZ_E_NOMEM, "cannot start a block worker: %s", ex.what()));
            break;
        }
    }
    } catch (const std::exception& ex) {                  // the workers that are running must still be joined
        record_error(fail(dynamic_cast<const std::bad_alloc*>(&ex) ? GCZ_E_NOMEM : GCZ_E_INTERNAL, "index writer: %s", ex.what()));
        std::lock_guard<std::mutex> l(mu);
        aborting = true;
        cv.notify_all();
    }
    for (std::thread& t : workers) t.join();
    ::close(ref_fd);
    ::close(ssa_fd);
    if (report) {
        report->blocks = n_blocks;
        report->sequences = sequences;
        report->symbols = symbols;
        report->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
    }
    if (first_error != GCZ_OK) return fail(first_error, "%s", first_message.c_str());
    return GCZ_OK;
}

}  // extern "C"

// ---- reader -------------------------------------------------------------------------------------------------------------
struct gcz_reader {
    gcz::MappedFile ref, ssa;
    bool has_ssa = false;
    struct BlockInfo {
        int64_t position = 0, size = 0, len = 0, header_len = 0;
        std::vector<std::string> headers;
    };
    std::vector<BlockInfo> blocks;
    int32_t sampling_factor = -1;
};

extern "C" {

static int reader_open_impl(const char* gcz_path, gcz_reader** out) {
    clear_error();
    if (!gcz_path || !out) return fail(GCZ_E_ARG, "null argument");
    std::unique_ptr<gcz_reader> r(new gcz_reader());
    GCZ_TRY_HOST(r->ref.open_ro(gcz_path));
    if (r->ref.size < 26) return fail(GCZ_E_FORMAT, "%s is too short for a gecoz file", gcz_path);
    // GecozFileReader(Path) :65-91: walk the block headers by `size` (do / while)
    int64_t position = 0;
    const int64_t total = r->ref.size;
    do {
        if (position + 26 > total) return fail(GCZ_E_FORMAT, "truncated block header at %lld", (long long)position);
        gcz_reader::BlockInfo b;
        const uint8_t* h = r->ref.data + position;
        b.position = position;
        b.size = get_le64(h + 9);
        b.len = get_le64(h + 17);
        int64_t p = position + 25;
        while (p < total && r->ref.data[p] > 0) {                                   // GecozRefBlockHeader(InputStream) :59-82
            const void* q = std::memchr(r->ref.data + p, 0, (size_t)(total - p));
            if (!q) return fail(GCZ_E_FORMAT, "unterminated header at %lld", (long long)p);
            const int64_t e = static_cast<const uint8_t*>(q) - r->ref.data;
            b.headers.emplace_back(reinterpret_cast<const char*>(r->ref.data + p), (size_t)(e - p));
            p = e + 1;
        }
        b.header_len = 26;
        for (const std::string& s : b.headers) b.header_len += (int64_t)s.size() + 1;
        const int64_t size = b.size;
        r->blocks.push_back(std::move(b));
        position += size;
        if (size <= 0) break;
    } while (position < total);

    const std::string ssa_path = ssa_path_for(gcz_path);
    if (access(ssa_path.c_str(), R_OK) == 0 && r->ssa.open_ro(ssa_path.c_str()) == GCZ_OK) {
        r->has_ssa = true;
        // the sampling factor is not stored: the smallest one whose index fits the file  :134-149
        const int64_t data_len = r->ssa.size - (int64_t)r->blocks.size() * 25;
        int sf = -1;
        while (true) {
            if (++sf > 30) return fail(GCZ_E_FORMAT, "invalid index file");
            int64_t need = 0;
            for (const auto& b : r->blocks) need += index_size(b.len, sf);
            if (data_len >= need) break;
        }
        r->sampling_factor = sf;
    }
    clear_error();
    *out = r.release();
    return GCZ_OK;
}

int32_t gcz_reader_num_blocks(const gcz_reader* r) { return r ? (int32_t)r->blocks.size() : 0; }

int gcz_reader_block(const gcz_reader* r, int32_t block, int64_t* text_len, int64_t* block_size, int32_t* n_headers) {
    if (!r || block < 0 || block >= (int32_t)r->blocks.size()) return fail(GCZ_E_ARG, "block index");
    const auto& b = r->blocks[(size_t)block];
    if (text_len) *text_len = b.len;
    if (block_size) *block_size = b.size;
    if (n_headers) *n_headers = (int32_t)b.headers.size();
    return GCZ_OK;
}

const char* gcz_reader_header(const gcz_reader* r, int32_t block, int32_t i) {
    if (!r || block < 0 || block >= (int32_t)r->blocks.size()) return nullptr;
    const auto& b = r->blocks[(size_t)block];
    return (i >= 0 && i < (int32_t)b.headers.size()) ? b.headers[(size_t)i].c_str() : nullptr;
}

int gcz_reader_find(const gcz_reader* r, const char* header, int32_t* block, int32_t* nstr) {
    if (!r || !header) return fail(GCZ_E_ARG, "null argument");
    for (size_t b = 0; b < r->blocks.size(); b++) {                                  // findBlockHeader :93-101 + findHeader
        for (size_t i = 0; i < r->blocks[b].headers.size(); i++) {
            if (r->blocks[b].headers[i] == header) {
                if (block) *block = (int32_t)b;
                if (nstr) *nstr = (int32_t)i;
                return GCZ_OK;
            }
        }
    }
    return fail(GCZ_E_ARG, "no sequence found: %s", header);
}

int32_t gcz_reader_sampling_factor(const gcz_reader* r) { return r ? r->sampling_factor : -1; }

// read(header) :115-177 up to the point where the GSSA is made: the two slices of a block, checked
static int block_slices(const gcz_reader* r, int32_t block, const uint8_t** body, int64_t* body_len, int64_t* text_len,
                        const uint8_t** ssa_body, int64_t* ssa_len) {
    if (!r || block < 0 || block >= (int32_t)r->blocks.size()) return fail(GCZ_E_ARG, "block index");
    if (!r->has_ssa) return fail(GCZ_E_ARG, "the .gcx index is missing: queries need it");
    const auto& b = r->blocks[(size_t)block];
    const int sf = r->sampling_factor;
    int64_t ssa_pos = 0;
    for (int32_t i = 0; i < block; i++) ssa_pos += 25 + index_size(r->blocks[(size_t)i].len, sf);
    const int64_t ssa_size = index_size(b.len, sf);
    if (ssa_pos + 25 + ssa_size > r->ssa.size || b.position + b.size > r->ref.size || b.size < b.header_len)
        return fail(GCZ_E_FORMAT, "invalid index file");
    std::vector<const char*> hdrs;
    for (const std::string& s : b.headers) hdrs.push_back(s.c_str());
    const uint8_t* sh = r->ssa.data + ssa_pos;
    if (get_le64(sh + 17) != header_hash(hdrs.data(), (int32_t)hdrs.size()) || get_le64(sh + 9) != ssa_size)
        return fail(GCZ_E_FORMAT, "invalid index file");                             // :165-172
    *body = r->ref.data + b.position + b.header_len;
    *body_len = b.size - b.header_len;
    *text_len = b.len;
    *ssa_body = sh + 25;
    *ssa_len = ssa_size;
    return GCZ_OK;
}

int gcz_reader_open_block(const gcz_reader* r, int32_t block, int device, gcz_index** out) {
    clear_error();
    if (!out) return fail(GCZ_E_ARG, "null argument");
    const uint8_t *body = nullptr, *ssa = nullptr;
    int64_t body_len = 0, text_len = 0, ssa_len = 0;
    GCZ_TRY_HOST(block_slices(r, block, &body, &body_len, &text_len, &ssa, &ssa_len));
    return gcz_open_block(device, body, body_len, text_len, ssa, ssa_len, out);
}

void gcz_reader_close(gcz_reader* r) { delete r; }

// ---- callers ---------------------------------------------------------------------------------------------------------------
namespace {

// this library's CUDA entry points behind the engine signature
int  def_open(int device, const uint8_t* a, int64_t al, int64_t tl, const uint8_t* b, int64_t bl, void** out) {
    gcz_index* idx = nullptr;
    const int rc = gcz_open_block(device, a, al, tl, b, bl, &idx);
    *out = idx;
    return rc;
}
void def_close(void* idx) { gcz_close_block(static_cast<gcz_index*>(idx)); }
int  def_num_strings(const void* idx, int32_t* out) { return gcz_num_strings(static_cast<const gcz_index*>(idx), out); }
int  def_string_ends(const void* idx, int64_t* e) { return gcz_string_ends(static_cast<const gcz_index*>(idx), e); }
int  def_find(void* idx, const uint8_t* p, const int64_t* o, int64_t n, int64_t* per, int64_t** pos, int64_t** off) {
    return gcz_find_batch(static_cast<gcz_index*>(idx), p, o, n, per, pos, off);
}
int  def_extract(void* idx, int32_t nstr, int64_t from, uint8_t* out, int64_t cap, int64_t* w) {
    return gcz_extract(static_cast<gcz_index*>(idx), nstr, from, out, cap, w);
}
void def_release(void* p) { gcz_free(p); }

gcz_query_engine resolve(const gcz_query_engine* e) {
    gcz_query_engine q;
    q.open_block = e && e->open_block ? e->open_block : def_open;
    q.close_block = e && e->close_block ? e->close_block : def_close;
    q.num_strings = e && e->num_strings ? e->num_strings : def_num_strings;
    q.string_ends = e && e->string_ends ? e->string_ends : def_string_ends;
    q.find_batch = e && e->find_batch ? e->find_batch : def_find;
    q.extract = e && e->extract ? e->extract : def_extract;
    q.release = e && e->release ? e->release : def_release;
    return q;
}

// one open block, closed on scope exit
struct OpenBlock {
    const gcz_query_engine& q;
    void* idx = nullptr;
    int32_t n_strings = 0;
    explicit OpenBlock(const gcz_query_engine& engine) : q(engine) {}
    int open(const gcz_reader* r, int32_t block, int device) {
        const uint8_t *body = nullptr, *ssa = nullptr;
        int64_t body_len = 0, text_len = 0, ssa_len = 0;
        GCZ_TRY_HOST(block_slices(r, block, &body, &body_len, &text_len, &ssa, &ssa_len));
        GCZ_TRY_HOST(q.open_block(device, body, body_len, text_len, ssa, ssa_len, &idx));
        return q.num_strings(idx, &n_strings);
    }
    ~OpenBlock() { if (idx) q.close_block(idx); }
};

// GSSA.find of a batch against one block: per-string counts [n_pats x n_strings], positions, offsets
struct Found {
    const gcz_query_engine& q;
    std::vector<int64_t> per;
    int64_t* pos = nullptr;
    int64_t* off = nullptr;
    explicit Found(const gcz_query_engine& engine) : q(engine) {}
    int run(OpenBlock& b, const std::vector<uint8_t>& data, const std::vector<int64_t>& offsets) {
        const int64_t n = (int64_t)offsets.size() - 1;
        per.assign((size_t)std::max<int64_t>(1, n * b.n_strings), 0);
        static const uint8_t none = 0;
        return q.find_batch(b.idx, data.empty() ? &none : data.data(), offsets.data(), n, per.data(), &pos, &off);
    }
    ~Found() { if (pos) q.release(pos); if (off) q.release(off); }
};

int give_text(const std::string& text, char** out_text, int64_t* out_len) {
    char* p = static_cast<char*>(std::malloc(text.size() + 1));
    if (!p) return fail(GCZ_E_NOMEM, "result text of %zu bytes", text.size());
    std::memcpy(p, text.data(), text.size());
    p[text.size()] = 0;
    *out_text = p;
    if (out_len) *out_len = (int64_t)text.size();
    return GCZ_OK;
}

// GecoMatch.print :143-157 for one block
void print_found(const std::vector<std::string>& headers, const int64_t* per, const int64_t* pos, bool with_positions, std::string& out) {
    int64_t o = 0;
    for (size_t i = 0; i < headers.size(); i++) {
        const int64_t k = per[i];
        if (k > 0) {
            out += ">" + headers[i] + " found : " + std::to_string(k) + "\n";
            if (with_positions) for (int64_t j = 0; j < k; j++) out += std::to_string(pos[o + j]) + "\n";
        }
        o += k;
    }
}

// String.split("\\|") + the ID= / ;Note= column of SimpleGFFGenerator :146-154 (trailing empty strings are dropped)
std::string gff_attributes(const std::string& header) {
    std::vector<std::string> parts;
    size_t start = 0;
    while (true) {
        const size_t bar = header.find('|', start);
        parts.push_back(header.substr(start, bar == std::string::npos ? std::string::npos : bar - start));
        if (bar == std::string::npos) break;
        start = bar + 1;
    }
    while (!parts.empty() && parts.back().empty()) parts.pop_back();
    if (parts.empty() && header.find('|') == std::string::npos) parts.push_back(header);     // "" -> [""]
    std::string s;
    if (!parts.empty()) s += "ID=" + parts[0];
    for (size_t i = 1; i < parts.size(); i++) s += ";Note=" + parts[i];
    return s;
}

struct PatternRecord { std::string header; std::vector<uint8_t> seq; };

// the record loop of SimpleGFFGenerator.search :59-86 (BufferedReader.readLine: \n, \r or \r\n end a line)
void read_pattern_records(const uint8_t* data, int64_t size, std::vector<PatternRecord>& out) {
    bool open = false;
    PatternRecord cur;
    auto flush = [&] { if (open && !cur.seq.empty()) out.push_back(cur); };
    int64_t p = 0;
    while (p < size) {
        int64_t e = p;
        while (e < size && data[e] != '\n' && data[e] != '\r') e++;
        const uint8_t* line = data + p;
        const int64_t len = e - p;
        p = e;
        if (p < size) { if (data[p] == '\r' && p + 1 < size && data[p + 1] == '\n') p += 2; else p += 1; }
        if (len > 0 && (line[0] == '>' || line[0] == '@')) {
            flush();
            cur.header.assign(reinterpret_cast<const char*>(line + 1), (size_t)(len - 1));
            cur.seq.clear();
            open = true;
        } else if (len > 0 && line[0] == '+') {
            flush();
            open = false;
            cur.seq.clear();
        } else if (open) {
            cur.seq.insert(cur.seq.end(), line, line + len);
        }
    }
    flush();
}

uint8_t complement(uint8_t b) {                                                       // reverse(byte) :110-118
    switch (b) { case 'A': return 'T'; case 'T': return 'A'; case 'C': return 'G'; case 'G': return 'C'; default: return b; }
}

}  // namespace

static int match_impl(const gcz_reader* r, int device, const char* header, const uint8_t* pattern, int64_t pattern_len,
              int32_t with_positions, const gcz_query_engine* engine, char** out_text, int64_t* out_len) {
    clear_error();
    if (!r || !pattern || pattern_len <= 0 || !out_text) return fail(GCZ_E_ARG, "match arguments");
    const gcz_query_engine q = resolve(engine);
    const std::vector<uint8_t> data(pattern, pattern + pattern_len);
    const std::vector<int64_t> off = { 0, pattern_len };
    std::string text;
    if (header) {                                                                     // :79-108
        int32_t block = -1, nstr = -1;
        GCZ_TRY_HOST(gcz_reader_find(r, header, &block, &nstr));
        OpenBlock b(q);
        GCZ_TRY_HOST(b.open(r, block, device));
        Found f(q);
        GCZ_TRY_HOST(f.run(b, data, off));
        if (nstr < b.n_strings && f.per[(size_t)nstr] > 0) {
            int64_t o = 0;
            for (int32_t i = 0; i < nstr; i++) o += f.per[(size_t)i];
            text += ">" + std::string(header) + " found : " + std::to_string(f.per[(size_t)nstr]) + "\n";
            if (with_positions) for (int64_t j = 0; j < f.per[(size_t)nstr]; j++) text += std::to_string(f.pos[o + j]) + "\n";
        }
    } else {                                                                          // :109-134
        for (int32_t block = 0; block < (int32_t)r->blocks.size(); block++) {
            OpenBlock b(q);
            GCZ_TRY_HOST(b.open(r, block, device));
            Found f(q);
            GCZ_TRY_HOST(f.run(b, data, off));
            std::vector<std::string> headers = r->blocks[(size_t)block].headers;
            headers.resize((size_t)b.n_strings);
            print_found(headers, f.per.data(), f.pos, with_positions != 0, text);
        }
    }
    return give_text(text, out_text, out_len);
}

static int gff_search_impl(const gcz_reader* r, int device, const uint8_t* patterns, int64_t patterns_len,
                   const gcz_query_engine* engine, char** out_text, int64_t* out_len) {
    clear_error();
    if (!r || (!patterns && patterns_len > 0) || patterns_len < 0 || !out_text) return fail(GCZ_E_ARG, "gff arguments");
    const gcz_query_engine q = resolve(engine);
    std::vector<PatternRecord> recs;
    read_pattern_records(patterns, patterns_len, recs);
    const int64_t nrec = (int64_t)recs.size();
    // forward (U -> T :96-100) then reverse complement (:104-108) of every record: one batch of 2 x nrec patterns
    std::vector<uint8_t> data;
    std::vector<int64_t> off(1, 0);
    for (PatternRecord& rec : recs) {
        for (uint8_t& c : rec.seq) if (c == 'U') c = 'T';
        data.insert(data.end(), rec.seq.begin(), rec.seq.end());
        off.push_back((int64_t)data.size());
    }
    for (const PatternRecord& rec : recs) {
        for (size_t i = rec.seq.size(); i-- > 0;) data.push_back(complement(rec.seq[i]));
        off.push_back((int64_t)data.size());
    }
    const int32_t nblocks = (int32_t)r->blocks.size();
    std::vector<std::unique_ptr<OpenBlock>> blocks;
    std::vector<std::unique_ptr<Found>> found;
    for (int32_t b = 0; b < nblocks && nrec > 0; b++) {                                // :52-56: every block is opened
        blocks.emplace_back(new OpenBlock(q));
        GCZ_TRY_HOST(blocks.back()->open(r, b, device));
        found.emplace_back(new Found(q));
        GCZ_TRY_HOST(found.back()->run(*blocks.back(), data, off));
    }
    std::string text;
    for (int64_t rec = 0; rec < nrec; rec++) {
        const std::string attrs = gff_attributes(recs[(size_t)rec].header);
        const int64_t length = (int64_t)recs[(size_t)rec].seq.size();
        for (int strand = 0; strand < 2; strand++) {
            const int64_t qi = strand == 0 ? rec : nrec + rec;
            for (int32_t b = 0; b < nblocks; b++) {
                const Found& f = *found[(size_t)b];
                const int32_t ns = blocks[(size_t)b]->n_strings;
                int64_t o = f.off[qi];
                for (int32_t j = 0; j < ns; j++) {
                    const int64_t k = f.per[(size_t)(qi * ns + j)];
                    const std::string& name = j < (int32_t)r->blocks[(size_t)b].headers.size() ? r->blocks[(size_t)b].headers[(size_t)j] : std::string();
                    for (int64_t x = 0; x < k; x++) {                                  // :134-156
                        const int64_t p = f.pos[o + x];
                        text += name + "\tgecotools\tdna\t" + std::to_string(p + 1) + "\t" + std::to_string(p + length) + "\t1.000\t" +
                                (strand ? "-" : "+") + "\t.\t" + attrs + "\n";
                    }
                    o += k;
                }
            }
        }
    }
    return give_text(text, out_text, out_len);
}

static int extract_fasta_impl(const gcz_reader* r, int device, const char* fasta_path, const gcz_query_engine* engine, int64_t* n_sequences) {
    clear_error();
    if (!r || !fasta_path) return fail(GCZ_E_ARG, "extract arguments");
    const gcz_query_engine q = resolve(engine);
    const int fd = ::open(fasta_path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) return fail(GCZ_E_ARG, "cannot create %s", fasta_path);
    struct Closer { int fd; ~Closer() { ::close(fd); } } closer{ fd };
    constexpr int64_t kBuffer = 1024 * 1024 * 4;                                      // tools/GecoRead.java:163
    constexpr int64_t kLine = 50;                                                     // fasta/FastaFileWriter.java:32
    std::vector<uint8_t> buf((size_t)kBuffer);
    int64_t count = 0;
    for (int32_t block = 0; block < (int32_t)r->blocks.size(); block++) {
        OpenBlock b(q);
        GCZ_TRY_HOST(b.open(r, block, device));
        std::vector<int64_t> e((size_t)std::max(1, b.n_strings));
        GCZ_TRY_HOST(q.string_ends(b.idx, e.data()));
        const auto& headers = r->blocks[(size_t)block].headers;
        for (const std::string& header : headers) {
            int32_t nstr = -1;                                                        // findHeader: the FIRST string of that name
            for (size_t i = 0; i < headers.size(); i++) if (headers[i] == header) { nstr = (int32_t)i; break; }
            if (nstr < 0 || nstr >= b.n_strings) continue;
            const int64_t len = nstr == 0 ? e[0] : e[(size_t)nstr] - e[(size_t)nstr - 1] - 1;   // GSSA.getLength :71-88
            std::vector<uint8_t> seq((size_t)std::max<int64_t>(len, 0));
            int64_t from = 0;
            do {                                                                      // SequenceExtractor.run :155-174
                int64_t w = 0;
                GCZ_TRY_HOST(q.extract(b.idx, nstr, from, buf.data(), kBuffer, &w));
                if (w <= 0) break;
                const int64_t take = std::min(w, std::max<int64_t>(len - from, 0));
                std::memcpy(seq.data() + from, buf.data(), (size_t)take);
                from += w;
            } while (from < len);
            // FastaFileWriter.write(TFastaSequence) + FastaSequenceWriter.run: a break after every 50 symbols, one more at the end
            std::string rec = ">" + header + "\n";
            rec.reserve(rec.size() + (size_t)(len + len / kLine + 1));
            for (int64_t i = 0; i < len; i += kLine) {
                rec.append(reinterpret_cast<const char*>(seq.data() + i), (size_t)std::min(kLine, len - i));
                rec.push_back('\n');
            }
            if (len % kLine == 0) rec.push_back('\n');
            for (size_t done = 0; done < rec.size();) {
                const ssize_t wr = ::write(fd, rec.data() + done, rec.size() - done);
                if (wr <= 0) return fail(GCZ_E_ARG, "cannot write %s", fasta_path);
                done += (size_t)wr;
            }
            count++;
        }
    }
    if (n_sequences) *n_sequences = count;
    return GCZ_OK;
}

// GecoRead.sequence  tools/GecoRead.java:33-81: the raw symbols [from, min(to, length)) of one sequence into `path`, one
// GSSA.extract call (ssa.extract(buf, nstr, from) :72 with a buffer of to - from bytes)
static int extract_sequence_impl(const gcz_reader* r, int device, const char* header, int64_t from, int64_t to, const char* path,
                         const gcz_query_engine* engine, int64_t* written) {
    clear_error();
    if (!r || !header || !path) return fail(GCZ_E_ARG, "extract arguments");
    const gcz_query_engine q = resolve(engine);
    int32_t block = -1, nstr = -1;
    GCZ_TRY_HOST(gcz_reader_find(r, header, &block, &nstr));                              // "no sequence found"
    OpenBlock b(q);
    GCZ_TRY_HOST(b.open(r, block, device));
    if (nstr >= b.n_strings) return fail(GCZ_E_ARG, "no sequence found: %s", header);
    std::vector<int64_t> e((size_t)std::max(1, b.n_strings));
    GCZ_TRY_HOST(q.string_ends(b.idx, e.data()));
    const int64_t len = nstr == 0 ? e[0] : e[(size_t)nstr] - e[(size_t)nstr - 1] - 1;     // GSSA.getLength :71-88
    to = std::min(to, len);                                                               // :66
    if (from < 0 || to < from) return fail(GCZ_E_RANGE, "range [%lld, %lld) of a sequence of %lld symbols", (long long)from, (long long)to, (long long)len);
    std::vector<uint8_t> buf((size_t)std::max<int64_t>(to - from, 1));
    int64_t w = 0;
    if (to > from) GCZ_TRY_HOST(q.extract(b.idx, nstr, from, buf.data(), to - from, &w));
    const int fd = ::open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) return fail(GCZ_E_ARG, "cannot create %s", path);
    const bool ok = write_fully(fd, buf.data(), w, 0);
    ::close(fd);
    if (!ok) return fail(GCZ_E_ARG, "cannot write %s", path);
    if (written) *written = w;
    return GCZ_OK;
}

// ---- the entry points above, behind the exception guard -----------------------------------------------------------------------
int gcz_fasta_open(const char* path, gcz_fasta** out) {
    return guarded([&] { return fasta_open_impl(path, out); });
}

int gcz_fasta_open_buffer(const uint8_t* data, int64_t size, gcz_fasta** out) {
    return guarded([&] { return fasta_open_buffer_impl(data, size, out); });
}

int gcz_index_fasta(const gcz_fasta* fasta, const char* gcz_path, const char* gcx_path, int32_t sampling_rate, int32_t n_devices, const int* devices, const gcz_engine* engine, gcz_index_report* report) {
    return guarded([&] { return index_fasta_impl(fasta, gcz_path, gcx_path, sampling_rate, n_devices, devices, engine, report); });
}

int gcz_reader_open(const char* gcz_path, gcz_reader** out) {
    return guarded([&] { return reader_open_impl(gcz_path, out); });
}

int gcz_match(const gcz_reader* r, int device, const char* header, const uint8_t* pattern, int64_t pattern_len, int32_t with_positions, const gcz_query_engine* engine, char** out_text, int64_t* out_len) {
    return guarded([&] { return match_impl(r, device, header, pattern, pattern_len, with_positions, engine, out_text, out_len); });
}

int gcz_gff_search(const gcz_reader* r, int device, const uint8_t* patterns, int64_t patterns_len, const gcz_query_engine* engine, char** out_text, int64_t* out_len) {
    return guarded([&] { return gff_search_impl(r, device, patterns, patterns_len, engine, out_text, out_len); });
}

int gcz_extract_fasta(const gcz_reader* r, int device, const char* fasta_path, const gcz_query_engine* engine, int64_t* n_sequences) {
    return guarded([&] { return extract_fasta_impl(r, device, fasta_path, engine, n_sequences); });
}

int gcz_extract_sequence(const gcz_reader* r, int device, const char* header, int64_t from, int64_t to, const char* path, const gcz_query_engine* engine, int64_t* written) {
    return guarded([&] { return extract_sequence_impl(r, device, header, from, to, path, engine, written); });
}

}  // extern "C"

/*
 * gcz.h — C ABI of the B200-native FM-index engine that replaces gecoz's nova-algo hot path.
 *
 * The reference (redmitry/gecoz) is pure Java and has no FFI today; the seam is nova-algo's public
 * Java API as used by nova-formats / nova-gecoz.  Every entry point below names the reference
 * interface it replaces (paths relative to /root/reference/java, see INTEGRATION.md for the JNI /
 * Panama-FFM stubs a maintainer would add on the Java side):
 *   algo/  = nova-algo/src/main/java/es/elixir/bsc/ngs/nova/algo/
 *   fmt/   = nova-formats/src/main/java/es/elixir/bsc/ngs/nova/gecoz/
 *
 * Conventions: plain pointers and sizes only; no exceptions cross the boundary; every function
 * returns 0 on success or a negative gcz_status; gcz_last_error() gives a thread-local message.
 * Buffers are caller-owned unless stated.  `text`, `gcz_body`, `gcx_body`, pattern and result
 * buffers may be host pointers (pageable, pinned, or mmap'd file slices) or device pointers of
 * the selected device; the library detects which.  There is NO CPU fallback: without a CUDA
 * device every compute entry point fails with GCZ_E_NODEVICE.
 */
#ifndef GCZ_H
#define GCZ_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum gcz_status {
    GCZ_OK          =  0,
    GCZ_E_ARG       = -1,   /* bad argument (null, size mismatch, n out of range)                      */
    GCZ_E_NOMEM     = -2,   /* device/host memory exhausted: the caller may re-queue the block, like
                               WriterPoolExecutor.afterExecute  fmt/GecozFileWriter.java:203-226          */
    GCZ_E_CUDA      = -3,   /* CUDA runtime error (message in gcz_last_error)                            */
    GCZ_E_FORMAT    = -4,   /* malformed .gcz/.gcx bytes: DataFormatException("invalid index file")
                               fmt/GecozFileReader.java:165-172                                          */
    GCZ_E_NODEVICE  = -5,   /* no CUDA device / library built without one: there is no CPU path          */
    GCZ_E_RANGE     = -6,   /* code length > 15, block > 2^31-1 symbols, pattern byte >= 0x80 ...        */
    GCZ_E_INTERNAL  = -7
} gcz_status;

/* ---- library --------------------------------------------------------------------------------- */
int         gcz_init(int n_devices, const int* device_ids);   /* NULL ids = devices 0..n-1; 0 = all visible */
void        gcz_shutdown(void);                                /* releases cached device workspaces        */
const char* gcz_last_error(void);                              /* thread-local                             */
const char* gcz_version(void);
int         gcz_device_count(void);
/* Run every launch of the calling thread on `cuda_stream` (a cudaStream_t; NULL = the library's own
 * per-device stream).  Lets a harness time the kernels with events on its own stream. */
int         gcz_set_stream(int device, void* cuda_stream);

/* ---- shape: HSWTShape(long[] counts)   algo/tree/HSWTShape.java:55-87 ------------------------- */
typedef struct gcz_shape {
    int8_t   bit_lengths[256];  /* DeflateEncodeTable.bit_lengths   algo/deflate/DeflateEncodeTable.java:52-56 */
    int16_t  codes[256];        /* DeflateEncodeTable.table: bit j = branch taken at depth j (:150-173)        */
    int32_t  n_nodes;           /* internal nodes of the Huffman tree                                            */
    int32_t  node_name[256];    /* file (pre-)order -> name = left-most leaf of the 1-subtree
                                   algo/tree/HuffmanShapedWaveletTree.java:106-108,165-182                       */
    int32_t  node_depth[256];   /* file order -> depth of the node                                               */
    int32_t  node_prefix[256];  /* file order -> path bits (LSB = root branch)                                   */
    int64_t  node_bits[256];    /* file order -> length of the node's bit vector                                 */
    int64_t  node_offset[256];  /* file order -> byte offset inside the block body (after the shape table)       */
    int64_t  table_bytes;       /* serialized RFC1951-style length table  algo/deflate/DeflateLengthsTable.java:136-171 */
    int64_t  length;            /* HSWTShape.length = sum(counts)                                                */
    int64_t  size;              /* HSWTShape.size   = table_bytes + sum RankedWTNode.bytes(node_bits)            */
} gcz_shape;

int     gcz_shape_from_counts(const int64_t counts[256], gcz_shape* out);
/* HSWTShape.write(ByteBuffer)  :111-115; writes table_bytes bytes, returns them (or <0) */
int64_t gcz_shape_write(const gcz_shape* shape, uint8_t* out, int64_t cap);
/* HSWTShape.read(ByteBuffer, long)  :89-109; fills bit_lengths/codes/table_bytes/tree, node_bits = 0 */
int     gcz_shape_read(const uint8_t* body, int64_t body_len, gcz_shape* out);

int64_t gcz_ranked_bytes(int64_t len_bits);                    /* RankedWTNode.bytes  algo/tree/RankedWTNode.java:60-67 */
int64_t gcz_index_size(int64_t n, int32_t sampling_factor);    /* GSSAIndex.getIndexSize  algo/ssa/GSSAIndex.java:200-205 */

/* ---- build ------------------------------------------------------------------------------------ */
/* The counting loop of GecozFileWriter.write  fmt/GecozFileWriter.java:127-130 (GPU histogram).
 * When `text` is a host buffer its upload stays on the device: the gcz_build_block that follows for the SAME
 * buffer and length (what GecozFileWriter.write does: count, reserve the file slices, queue the block) finds the
 * text there and does not copy it a second time.  The buffer must not change between the two calls; any other
 * gcz_build_block / gcz_count_symbols on the device drops the staged copy. */
int gcz_count_symbols(int device, const uint8_t* text, int64_t n, int64_t counts[256]);

/* One call == BlockWriter.run  fmt/GecozFileWriter.java:256-284:
 *   SAIS.suffix(in, sa) :262, shape.write(out) :267, HuffmanShapedWaveletTree.write(shape, BWT, out) :268,
 *   GSSAIndex.write(sa, rate, idx) :274.
 * gcz_body (shape->size bytes) and gcx_body (gcz_index_size(n, log2 rate) bytes) are the mapped file slices
 * positioned just after their headers; every byte of both is written.  sa_out (n int32) and bwt_out
 * (n bytes) are optional parity artefacts.  Thread-safe for different blocks; blocks on the same device
 * serialize on that device's workspace. */
int gcz_build_block(int device, const uint8_t* text, int64_t n, int32_t sampling_rate,
                    const gcz_shape* shape,
                    uint8_t* gcz_body, int64_t gcz_body_len,
                    uint8_t* gcx_body, int64_t gcx_body_len,
                    int32_t* sa_out, uint8_t* bwt_out);

/* Device-time breakdown (milliseconds, CUDA events) of the calling thread's last gcz_build_block. */
typedef struct gcz_build_timing {
    float h2d_ms, sort_initial_ms, sort_refine_ms, bwt_hswt_ms, ssa_ms, d2h_ms, total_ms;
    int32_t refine_rounds;
    int64_t radix_launches, radix_elements;     /* onesweep passes launched, elements they moved     */
    float   radix_ms;                            /* device time inside those passes                   */
    int64_t kernel_launches;                     /* all kernels of this library launched by the call  */
    int32_t symbols_per_key;                     /* k: symbols packed into the first sort key         */
    int32_t long_runs;                           /* runs of >= k equal symbols (ordered in closed form) */
    int64_t unresolved_after_first_sort;         /* suffixes not unique in their first k symbols      */
    int64_t radix_full_launches;                 /* onesweep passes over all n pairs, array input     */
    float   radix_full_ms;                       /* device time inside those                          */
    float   radix_text_ms;                       /* device time of the first pass (reads the text)    */
} gcz_build_timing;
int gcz_last_build_timing(gcz_build_timing* out);

/* ---- query: handle == GSSA  algo/ssa/GSSA.java ------------------------------------------------- */
typedef struct gcz_index gcz_index;

/* GecozFileReader.read(header)  fmt/GecozFileReader.java:115-177: gcz_body starts at the shape table
 * (just after the block header), gcx_body just after the 25-byte GecozSSA header.  The sampling factor is
 * recovered from gcx_len like GSSAIndex(ByteBuffer,long)  algo/ssa/GSSAIndex.java:57-71.  The .gcx is
 * mandatory (the reference cannot locate without it): NULL fails fast with GCZ_E_ARG.
 * Device memory of an open block: the two bodies re-laid out as 32-byte rank sectors (1.14 x their size) and, for DNA
 * blocks, a table of the backward-search intervals of all strings of K <= 12 symbols over A, C, G, T (never larger than
 * the sectors), which the searches use for the last K symbols of a pattern; results are those of GSSA.search. */
int  gcz_open_block(int device, const uint8_t* gcz_body, int64_t body_len, int64_t text_len,
                    const uint8_t* gcx_body, int64_t gcx_len, gcz_index** out);
void gcz_close_block(gcz_index* idx);

int  gcz_text_length(const gcz_index* idx, int64_t* out);          /* GSSA.getLength()        :67-69   */
int  gcz_sampling_factor(const gcz_index* idx, int32_t* out);
int  gcz_num_strings(const gcz_index* idx, int32_t* out);          /* e.length, GSSA.index    :232-238 */
int  gcz_string_ends(const gcz_index* idx, int64_t* e);            /* sorted '\0' positions            */
int  gcz_c_array(const gcz_index* idx, int64_t c[256]);            /* GSSA.index              :215-226 */

/* Backward search, the interval part of GSSA.search  :187-197 with HSWT.occ  algo/tree/
 * HuffmanShapedWaveletTree.java:247-267.  Pattern i is pats[pat_off[i] .. pat_off[i+1]).  sp/ep are the
 * values the Java loop holds when it exits (ep < sp  <=>  not found). */
int  gcz_count_batch(gcz_index* idx, const uint8_t* pats, const int64_t* pat_off, int64_t n_pats,
                     int64_t* sp, int64_t* ep);

/* GSSA.locate  :241-251 for explicit SA rows (LF-walk to a marked row + IndexWaveletTree.get). */
int  gcz_locate_rows(gcz_index* idx, const int64_t* rows, int64_t n_rows, int64_t* positions);

/* GSSA.find  :160-185 for a batch.  per_string_counts is n_pats x n_strings (row-major) = GSSA.count :136-148.
 * *positions receives, pattern after pattern and string after string, the ascending 0-based positions
 * relative to the string start; (*pos_off)[i] .. (*pos_off)[i+1] delimits pattern i.  Both arrays are
 * callee-allocated host memory, released with gcz_free. */
int  gcz_find_batch(gcz_index* idx, const uint8_t* pats, const int64_t* pat_off, int64_t n_pats,
                    int64_t* per_string_counts, int64_t** positions, int64_t** pos_off);
/* GSSA.extract(ByteBuffer buf, int nstr, long from)  :90-126: the symbols of string nstr from position `from`
 * (0-based inside the string), at most `cap` of them and never past the string's end, into out[0 ..]; *written =
 * the buffer position the reference leaves behind.  One thread per sampled text position walks 2^sampling_factor
 * LF steps (GSSAIndex.find :184-187, IndexWaveletTree.find, RankedWTNode.findZero/findOne for the anchors); a call
 * whose reference walk leaves the true LF chain (merged blocks) is replayed sequentially so that
 * the bytes stay the reference's. */
int  gcz_extract(gcz_index* idx, int32_t nstr, int64_t from, uint8_t* out, int64_t cap, int64_t* written);
void gcz_free(void* p);

/* ---- a batch against every block of a file -------------------------------------------------------
 * The callers of the reference loop over the blocks for every pattern (GecoMatch.match  tools/GecoMatch.java:114-131,
 * SimpleGFFGenerator.search  tools/SimpleGFFGenerator.java:52-56,128).  For a batch the loops are swapped: the batch is
 * uploaded ONCE (or already lives on the device: `pats` / `pat_off` may be device pointers), every block is searched
 * there, and only results come back.  All blocks of one call must be open on the same device. */

/* "total found" of GecoMatch for every pattern: totals[i] = sum over the blocks of max(0, ep - sp + 1) of pattern i
 * (host or device array of n_pats). */
int  gcz_count_multi(gcz_index* const* blocks, int32_t n_blocks, const uint8_t* pats, const int64_t* pat_off, int64_t n_pats,
                     int64_t* totals);

/* GSSA.find :160-185 of every pattern in every block, as sparse records: the hits of block b are
 * [block_off[b], block_off[b + 1]), sorted by (pattern, string, position) — for one pattern that is the order in which
 * find returns them (string after string, positions ascending, 0-based inside the string).  Intervals, locate
 * (:241-251), the sort and the split by string ends all run on the device.  The arrays are callee-allocated host
 * memory: release them with gcz_hits_free. */
typedef struct gcz_hits {
    int64_t  n_hits;
    int64_t* block_off;    /* n_blocks + 1 */
    int64_t* pattern;      /* index of the pattern in the batch */
    int32_t* string;       /* string ordinal inside its block (= index into the block header's list) */
    int64_t* position;
} gcz_hits;
int  gcz_find_multi(gcz_index* const* blocks, int32_t n_blocks, const uint8_t* pats, const int64_t* pat_off, int64_t n_pats,
                    gcz_hits* out);
void gcz_hits_free(gcz_hits* hits);

/* What a batch of backward searches reads (measurement, not a timed path).  rank_sectors: 32-byte sectors the search kernel
 * loads (a lookup in the interval table counts as one); kernel_ms: its device time.  steps, reference_rank_calls: the
 * backward-search steps and the RankedWTNode.count calls (algo/tree/RankedWTNode.java:98-122) of the reference's own loop
 * for the same patterns — 2 per character and code bit while the position is >= 0
 * (algo/tree/HuffmanShapedWaveletTree.java:247-267), 74 bytes each in the file layout — counted by running
 * the search once more symbol by symbol, without the table. */
typedef struct gcz_query_stats {
    int64_t patterns, blocks, steps, rank_sectors, reference_rank_calls, index_bytes;
    float   kernel_ms;       /* device time of the search kernels of the call */
} gcz_query_stats;
int  gcz_count_stats(gcz_index* const* blocks, int32_t n_blocks, const uint8_t* pats, const int64_t* pat_off, int64_t n_pats,
                     gcz_query_stats* out);
int  gcz_last_query_stats(gcz_query_stats* out);   /* of the calling thread's last gcz_count_multi: patterns, blocks, kernel_ms */

/* ---- stage-level hooks used by the parity tests (each one is a stage of gcz_build_block) ------- */
int gcz_dbg_set_find_chunk(int64_t occurrences);   /* occurrences located + sorted per launch inside find (default 2^26; <= 0 restores it) */
int gcz_dbg_sort_pairs(int device, uint64_t* keys, uint32_t* vals, int64_t n, int32_t begin_bit, int32_t end_bit);
int gcz_dbg_suffix_array(int device, const uint8_t* text, int64_t n, int32_t* sa);
int gcz_dbg_ranked_vector(int device, const uint8_t* bits, int64_t len, uint8_t* out);   /* one bit per byte in */
int gcz_dbg_index_wavelet_tree(int device, const int32_t* vals, int64_t m, uint8_t* out);

#ifdef __cplusplus
}
#endif
#endif /* GCZ_H */

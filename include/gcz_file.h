/*
 * gcz_file.h — native host layer above gcz.h: FASTA records, block planning, the .gcz/.gcx writer and reader.
 *
 * C ABI counterpart of the reference's host classes for the FM-index path (paths relative to /root/reference/java):
 *   fasta/ = nova-formats/src/main/java/es/elixir/bsc/ngs/nova/fasta/
 *   fmt/   = nova-formats/src/main/java/es/elixir/bsc/ngs/nova/gecoz/
 *   tools/ = nova-gecoz/src/main/java/es/elixir/bsc/ngs/nova/gecoz/tools/
 * Same conventions as gcz.h (0 / negative gcz_status, gcz_last_error()).  The per-block device work goes through a
 * gcz_engine; NULL selects this library's CUDA entry points (gcz_count_symbols, gcz_build_block) — the tests on a
 * machine without a GPU pass an engine of their own to exercise the host logic.  Gzipped FASTA (nova-gzip in the
 * reference; host I/O either way) is inflated with zlib, looked up at run time; without libz.so.1 gcz_fasta_open
 * fails with GCZ_E_FORMAT and the caller hands the decompressed bytes to gcz_fasta_open_buffer.
 */
#ifndef GCZ_FILE_H
#define GCZ_FILE_H

#include "gcz.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- FASTA / FASTQ records: FastaIterator (fasta/FastaIterator.java:39-127), lazy = true --------------------- */
typedef struct gcz_fasta gcz_fasta;

int     gcz_fasta_open(const char* path, gcz_fasta** out);                       /* maps the file                    */
int     gcz_fasta_open_buffer(const uint8_t* data, int64_t size, gcz_fasta** out); /* borrows `data` until close       */
int64_t gcz_fasta_count(const gcz_fasta* f);
/* header: the line after '>' / '@' without CR; position: offset of the first sequence byte; length: sequence bytes
 * (CR/LF not counted); multiline: the sequence spans more than one line (FastaSequence fields). */
int     gcz_fasta_record(const gcz_fasta* f, int64_t i, const char** header, int64_t* position, int64_t* length,
                         int32_t* multiline);
/* FastaFileReader.read(buf, seq)  fasta/FastaFileReader.java:109-160: the sequence bytes, verbatim, CR/LF removed */
int     gcz_fasta_read(const gcz_fasta* f, int64_t i, uint8_t* out, int64_t cap);
void    gcz_fasta_close(gcz_fasta* f);

/* ---- which sequences share a block, and in which order: GecoIndex.index  tools/GecoIndex.java:57-98 ------------
 * block_of[i] = block (file order) of sequence i, or -1 when the reference's TreeSets drop it (same length and
 * header as an earlier one); order_in_block[i] = its rank inside the block (fmt/GecozRefBlock.java:38-71,
 * fasta/TFastaSequence.java:46-52).  Returns the number of blocks. */
int64_t gcz_plan_blocks(const int64_t* lengths, const char* const* headers, int64_t n_sequences,
                        int64_t* block_of, int64_t* order_in_block);

/* ---- headers: fmt/GecozRefBlockHeader.java:90-136, fmt/GecozSSABlockHeader.java:69-74 ------------------------- */
int64_t gcz_ref_header_length(const char* const* headers, int32_t n_headers);    /* getBlockHeaderLength :130-136 */
int64_t gcz_header_hash(const char* const* headers, int32_t n_headers);          /* getBlockHeaderHash   :120-128 */
int64_t gcz_ref_header_write(const char* const* headers, int32_t n_headers, int64_t block_size, int64_t text_len,
                             uint8_t* out, int64_t cap);                          /* bytes written or < 0          */
int64_t gcz_ssa_header_write(const char* const* headers, int32_t n_headers, int64_t index_len, uint8_t out[25]);

/* ---- the per-block device work ---------------------------------------------------------------------------------- */
typedef struct gcz_engine {
    int (*count_symbols)(int device, const uint8_t* text, int64_t n, int64_t counts[256]);
    int (*build_block)(int device, const uint8_t* text, int64_t n, int32_t sampling_rate, const gcz_shape* shape,
                       uint8_t* gcz_body, int64_t gcz_body_len, uint8_t* gcx_body, int64_t gcx_body_len,
                       int32_t* sa_out, uint8_t* bwt_out);
} gcz_engine;

/* ---- writer: GecoIndex.index + GecozFileWriter  tools/GecoIndex.java:51-146, fmt/GecozFileWriter.java:60-310 ------
 * Blocks in the reference's order; offsets fixed before a block is built; `devices` replaces `threads`: two blocks in
 * flight per device (one being built, the next one counted = uploaded).  A block that fails with GCZ_E_NOMEM is
 * retried once when nothing else is in flight (WriterPoolExecutor.afterExecute :203-226).  gcx_path NULL: the
 * reference's rule (x.gcz -> x.gcx, else name + "gcx").  report (optional): blocks written, symbols, seconds. */
typedef struct gcz_index_report {
    int64_t blocks, sequences, symbols;
    double  seconds;
} gcz_index_report;

int gcz_index_fasta(const gcz_fasta* fasta, const char* gcz_path, const char* gcx_path, int32_t sampling_rate,
                    int32_t n_devices, const int* devices, const gcz_engine* engine, gcz_index_report* report);

/* ---- reader: GecozFileReader  fmt/GecozFileReader.java:57-200 ---------------------------------------------------- */
typedef struct gcz_reader gcz_reader;

int     gcz_reader_open(const char* gcz_path, gcz_reader** out);                 /* walks the block headers :65-91  */
int32_t gcz_reader_num_blocks(const gcz_reader* r);
int     gcz_reader_block(const gcz_reader* r, int32_t block, int64_t* text_len, int64_t* block_size, int32_t* n_headers);
const char* gcz_reader_header(const gcz_reader* r, int32_t block, int32_t i);
/* findBlockHeader + findHeader :93-113: the block that holds `header` and the string's index in it, or GCZ_E_ARG */
int     gcz_reader_find(const gcz_reader* r, const char* header, int32_t* block, int32_t* nstr);
int32_t gcz_reader_sampling_factor(const gcz_reader* r);                         /* recovered from the .gcx size :134-149 */
/* read(header) :115-177: checks the .gcx block header (hash, length) and opens the block on `device` */
int     gcz_reader_open_block(const gcz_reader* r, int32_t block, int device, gcz_index** out);
void    gcz_reader_close(gcz_reader* r);

/* ---- the callers: gecotools -c / -s / -s file / -o  --------------------------------------------------------------
 * The query-side device work goes through a gcz_query_engine; NULL (or a NULL member) selects this library's CUDA
 * entry points (gcz_open_block, gcz_find_batch, gcz_extract, ...).  Handles are opaque to the host layer. */
typedef struct gcz_query_engine {
    int  (*open_block)(int device, const uint8_t* gcz_body, int64_t body_len, int64_t text_len,
                       const uint8_t* gcx_body, int64_t gcx_len, void** out);
    void (*close_block)(void* idx);
    int  (*num_strings)(const void* idx, int32_t* out);
    int  (*string_ends)(const void* idx, int64_t* e);
    int  (*find_batch)(void* idx, const uint8_t* pats, const int64_t* pat_off, int64_t n_pats,
                       int64_t* per_string_counts, int64_t** positions, int64_t** pos_off);
    int  (*extract)(void* idx, int32_t nstr, int64_t from, uint8_t* out, int64_t cap, int64_t* written);
    void (*release)(void* p);                    /* frees what find_batch allocated */
} gcz_query_engine;

/* GecoMatch.match / count  tools/GecoMatch.java:51-157: the lines the tool prints (">hdr found : k", then the 0-based
 * positions when with_positions), '\n'-terminated, in *out_text (malloc'd, release with gcz_free).  header NULL:
 * every block; else only the block that holds it and only that string (GCZ_E_ARG when there is none). */
int gcz_match(const gcz_reader* r, int device, const char* header, const uint8_t* pattern, int64_t pattern_len,
              int32_t with_positions, const gcz_query_engine* engine, char** out_text, int64_t* out_len);

/* SimpleGFFGenerator.search  tools/SimpleGFFGenerator.java:45-163: every record of a FASTA/FASTQ pattern file (bytes in
 * `patterns`), as given (U -> T) and reverse-complemented, against every block; one GFF line per occurrence in the
 * reference's order.  The searches are batched: one find_batch call per block. */
int gcz_gff_search(const gcz_reader* r, int device, const uint8_t* patterns, int64_t patterns_len,
                   const gcz_query_engine* engine, char** out_text, int64_t* out_len);

/* GecoRead.fasta  tools/GecoRead.java:83-175: every sequence of every block into a FASTA file — 4 MiB extract calls
 * (SequenceExtractor :155-174), 50 symbols per line and one more line break at the end of a record
 * (fasta/FastaFileWriter.java:132-215). */
int gcz_extract_fasta(const gcz_reader* r, int device, const char* fasta_path, const gcz_query_engine* engine,
                      int64_t* n_sequences);

/* GecoRead.sequence  tools/GecoRead.java:33-81 (`gecotools -i x.gcz -o out.seq header [from] [to]`): the raw symbols
 * [from, min(to, length)) of the sequence `header` into `path` with one extract call.  GCZ_E_ARG: no such sequence;
 * GCZ_E_RANGE: from < 0 or from beyond the end (where the reference fails mapping a negative size). */
int gcz_extract_sequence(const gcz_reader* r, int device, const char* header, int64_t from, int64_t to, const char* path,
                         const gcz_query_engine* engine, int64_t* written);

#ifdef __cplusplus
}
#endif
#endif /* GCZ_FILE_H */

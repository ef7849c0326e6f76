/*
 * gcz_file.h — native host layer above gcz.h: FASTA records, block planning, the .gcz/.gcx writer and reader.
 *
 * C ABI counterpart of the reference's host classes for the FM-index path (paths relative to /root/reference/java):
 *   fasta/ = nova-formats/src/main/java/es/elixir/bsc/ngs/nova/fasta/
 *   fmt/   = nova-formats/src/main/java/es/elixir/bsc/ngs/nova/gecoz/
 *   tools/ = nova-gecoz/src/main/java/es/elixir/bsc/ngs/nova/gecoz/tools/
 * Same conventions as gcz.h (0 / negative gcz_status, gcz_last_error()).  The per-block device work goes through a
 * gcz_engine; NULL selects this library's CUDA entry points (gcz_count_symbols, gcz_build_block) — the tests on a
 * machine without a GPU pass an engine of their own to exercise the host logic.  Gzipped FASTA is not read here
 * (nova-gzip stays on the host side of the caller): hand the decompressed bytes to gcz_fasta_open_buffer.
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

#ifdef __cplusplus
}
#endif
#endif /* GCZ_FILE_H */

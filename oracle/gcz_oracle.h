/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see orc_internal.h for the full statement).
 *
 * Public surface of the CPU restatement of gecoz's FM-index path, loaded through ctypes by
 * oracle/gcz_oracle.py.  Every function cites the Java lines it restates; paths are relative
 * to /root/reference/java and abbreviated:
 *   io/    = nova-io/src/main/java/es/elixir/bsc/ngs/nova/io/
 *   algo/  = nova-algo/src/main/java/es/elixir/bsc/ngs/nova/algo/
 *   fmt/   = nova-formats/src/main/java/es/elixir/bsc/ngs/nova/gecoz/
 *   tools/ = nova-gecoz/src/main/java/es/elixir/bsc/ngs/nova/gecoz/tools/
 */
#ifndef GCZ_ORACLE_H
#define GCZ_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- io/AbstractBitStream.java + io/BitBuffer.java ------------------------------------ */
/* Emulates `new BitBuffer(ByteBuffer.allocate(cap))`, a sequence of writeBits(vals[i],
 * nbits[i]) and a final flush(); returns the buffer position after the flush. */
int64_t orc_bitbuffer_write(const int64_t* vals, const int32_t* nbits, int32_t count,
                            uint8_t* out, int64_t cap);
/* writeBits..., rewind(), then readBits(rn[i]) for i < rcount (raw return values). */
void orc_bitbuffer_write_read(const int64_t* vals, const int32_t* nbits, int32_t count,
                              int64_t cap, const int32_t* rn, int32_t rcount, int64_t* rout);

/* ---- algo/tree/RankedWTNode.java -------------------------------------------------------- */
int64_t orc_ranked_bytes(int64_t len);                                   /* :60-67  */
/* put() every bit of bits[0..len) (one byte per bit) then flush(); out has bytes(len). :228-245 */
int64_t orc_ranked_write(const uint8_t* bits, int64_t len, uint8_t* out);
int32_t orc_ranked_get(const uint8_t* buf, int64_t len, int64_t idx);    /* :81-84  */
int64_t orc_ranked_count(const uint8_t* buf, int64_t len, int64_t idx);  /* :98-122 */
int64_t orc_ranked_find_one(const uint8_t* buf, int64_t len, int64_t n);  /* :150-152,181-205 */
int64_t orc_ranked_find_zero(const uint8_t* buf, int64_t len, int64_t n); /* :140-142,154-179 */

/* ---- algo/huffman + algo/deflate tables -------------------------------------------------- */
/* new DeflateEncodeTable(counts, max_bits): lengths + reversed canonical codes.
 * algo/huffman/HuffmanEncodeTable.java:48-111, algo/deflate/DeflateEncodeTable.java:52-173 */
int32_t orc_deflate_encode_table(const int64_t* counts, int32_t n, int32_t max_bits,
                                 int8_t* bit_lengths, int16_t* table);
/* new DeflateLookupTable(bit_lengths).getSymbol(code)  algo/deflate/DeflateLookupTable.java:40-153 */
int32_t orc_lookup_get_symbol(const int8_t* bit_lengths, int32_t n, int32_t code);
int32_t orc_lookup_get_symbol_nbits(const int8_t* bit_lengths, int32_t n, int32_t code, int32_t nbits);
/* putSymbol every data byte into a BitBuffer(cap), rewind, getSymbol(in) n times: returns
 * number of mismatches (DeflateTablesTest.test / stress_test2). */
int64_t orc_deflate_stream_roundtrip(const uint8_t* data, int64_t n, int64_t cap);
int32_t orc_deflate_lengths_bits(const int8_t* bit_lengths, int32_t n);  /* DeflateLengthsTable.length :136-171 */

/* ---- algo/tree/HSWTShape.java:55-87 -------------------------------------------------------- */
typedef struct {
    int8_t  bit_lengths[256];   /* encode.bit_lengths */
    int16_t table[256];         /* encode.table (bit j = branch at depth j) */
    int32_t node_bits[256];     /* bit-vector length of the internal node NAMED by that symbol */
    int64_t length;             /* sum of counts */
    int64_t size;               /* serialized size in bytes: shape table + all ranked nodes */
    int64_t table_bytes;        /* bytes of the RFC1951-style length table alone */
} orc_shape;
int32_t orc_shape_from_counts(const int64_t counts[256], orc_shape* out);
int64_t orc_shape_write(const orc_shape* s, uint8_t* out, int64_t cap);  /* HSWTShape.write :111-115 */

/* ---- algo/string/SAIS.java:103-137 (RESULT ONLY: unsigned bytes, shorter-is-smaller) ------- */
int32_t orc_suffix_array(const uint8_t* text, int64_t n, int32_t* sa);
int32_t orc_suffix_array_naive(const uint8_t* text, int64_t n, int32_t* sa);

/* ---- fmt/GecozFileWriter.java:240-303 BlockWriter.run == one block ------------------------ */
int64_t orc_index_size(int64_t n, int32_t sampling_factor);              /* GSSAIndex.getIndexSize :200-205 */
/* body = shape table + HSWT nodes, index = marker vector + IWT levels.  sa_out/bwt_out nullable.
 * threads: 1 = sequential, 2 = HSWT on a side thread like the reference (:264-277). */
int32_t orc_build_block(const uint8_t* text, int64_t n, int32_t sampling_rate,
                        uint8_t* gcz_body, int64_t gcz_body_len,
                        uint8_t* gcx_body, int64_t gcx_body_len,
                        int32_t* sa_out, uint8_t* bwt_out, int32_t threads);
/* pieces, for tests */
int32_t orc_hswt_write(const orc_shape* s, const uint8_t* text, const int32_t* sa, int64_t n,
                       uint8_t* out, int64_t cap);       /* algo/tree/HuffmanShapedWaveletTree.java:95-146,165-182 */
int32_t orc_gssa_index_write(const int32_t* sa, int64_t n, int32_t sampling_rate,
                             uint8_t* out, int64_t cap); /* algo/ssa/GSSAIndex.java:129-150 */
int32_t orc_iwt_write(const int32_t* vals, int64_t m, uint8_t* out, int64_t cap); /* algo/tree/IndexWaveletTree.java:83-112 */
int64_t orc_iwt_get(const uint8_t* buf, int64_t m, int64_t pos);                  /* :127-144 */
int64_t orc_iwt_find(const uint8_t* buf, int64_t m, int64_t idx);                 /* :152-165 */

/* ---- fmt/GecozFileReader.read :115-177 + algo/ssa/GSSA.java ----------------------------- */
typedef struct orc_gssa orc_gssa;
/* gcz_body starts at the shape table (just after the block header).  Returns NULL on error. */
orc_gssa* orc_open(const uint8_t* gcz_body, int64_t body_len, int64_t text_len,
                   const uint8_t* gcx_body, int64_t gcx_len);
void      orc_close(orc_gssa*);
int32_t   orc_sampling_factor(const orc_gssa*);
int32_t   orc_num_strings(orc_gssa*);                       /* e.length, GSSA.index :232-238 */
void      orc_string_ends(orc_gssa*, int64_t* e);
void      orc_c_array(orc_gssa*, int64_t* c256);            /* GSSA.index :215-226 */
int32_t   orc_num_nodes(const orc_gssa*);                   /* mapped internal nodes */
void      orc_node_info(const orc_gssa*, int32_t* names, int64_t* lens, int64_t* offsets); /* file order */
int64_t   orc_occ(const orc_gssa*, int32_t symbol, int64_t pos);  /* HSWT.occ :247-267 */
int64_t   orc_get_rs(const orc_gssa*, int64_t pos);                /* HSWT.getRS :300-314 */
/* backward search GSSA.search :187-197; returns number of RankedWTNode.count calls made */
int64_t   orc_search(orc_gssa*, const uint8_t* pat, int64_t len, int64_t* sp, int64_t* ep);
int64_t   orc_locate(orc_gssa*, int64_t row);               /* GSSA.locate :241-251 */
int64_t   orc_index_find(orc_gssa*, int64_t pos);           /* GSSAIndex.find  algo/ssa/GSSAIndex.java:184-187 */
int64_t   orc_extract(orc_gssa*, int32_t nstr, int64_t from, uint8_t* out, int64_t cap);   /* GSSA.extract :90-126 */
/* GSSA.find :160-185.  Returns total hits k (0 == Java null).  positions (cap ints) receives the
 * per-string relative positions concatenated in string order; per_string[ns] the counts. */
int64_t   orc_find(orc_gssa*, const uint8_t* pat, int64_t len,
                   int64_t* per_string, int64_t* positions, int64_t cap);
/* batch forms used for CPU baselines */
int64_t   orc_search_batch(orc_gssa*, const uint8_t* pats, const int64_t* off, int64_t np,
                           int64_t* sp, int64_t* ep);
uint64_t  orc_rank_calls(void);      /* instrumentation: RankedWTNode.count calls so far (this thread) */
void      orc_rank_calls_reset(void);

/* ---- headers: fmt/GecozRefBlockHeader.java:90-128, fmt/GecozSSABlockHeader.java:69-74 ---- */
int64_t orc_header_hash(const char* const* headers, int32_t nh);
int32_t orc_ref_header_len(const char* const* headers, int32_t nh);
int32_t orc_ref_header_write(const char* const* headers, int32_t nh, int64_t size, int64_t len, uint8_t* out);
int32_t orc_ssa_header_write(const char* const* headers, int32_t nh, int64_t len, uint8_t* out);

/* ---- tools/GecoIndex.java:72-98 block merge + order ------------------------------------- */
/* in: nseq (length, header); out: block_of[i] = block ordinal in FILE order, pos_in_block[i] =
 * ordinal of sequence i inside its block.  Returns the number of blocks. */
int32_t orc_merge_blocks(const int32_t* lengths, const char* const* headers, int32_t nseq,
                         int32_t* block_of, int32_t* pos_in_block);

#ifdef __cplusplus
}
#endif
#endif

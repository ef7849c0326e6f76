/* ORACLE — TEST INFRASTRUCTURE ONLY (see orc_internal.h).  Table helpers shared by the oracle TUs. */
#ifndef ORC_TABLES_H
#define ORC_TABLES_H
#include "orc_internal.h"
#include "orc_bits.h"

typedef struct { jshort* table; jint size; } orc_lookup;      /* DeflateLookupTable.table */

int  orc_encode_table(const jlong* counts, int n, int max_bits, jbyte* bit_lengths, jshort* table);
int  orc_encode_table_from_lengths(int n, const jbyte* bit_lengths, jshort* table);
int  orc_lookup_build(orc_lookup* t, const jbyte* bit_lengths, int n);
void orc_lookup_free(orc_lookup* t);
int  orc_lookup_symbol(const orc_lookup* t, jint code);
int  orc_lookup_symbol_nbits(const orc_lookup* t, jint code, jint nbits);
int  orc_lookup_symbol_stream(const orc_lookup* t, orc_bits* in);
int  orc_lengths_table_bits(const jbyte* bit_lengths, int len_n);
void orc_lengths_table_write(const jbyte* d_tree, int len_n, orc_bits* out);
int  orc_lengths_table_read(orc_bits* in, jbyte* d_tree, int len_n);

#endif

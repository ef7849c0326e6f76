/* ORACLE — TEST INFRASTRUCTURE ONLY (see orc_internal.h).  Bit stream state shared by the oracle TUs. */
#ifndef ORC_BITS_H
#define ORC_BITS_H
#include "orc_internal.h"

/* io/AbstractBitStream.java fields (:38-49) over a ByteBuffer view [0, limit) with a position */
typedef struct {
    uint8_t* buf;
    int64_t  pos, limit;
    jlong    value;
    int      bits_left;     /* Java `byte` */
    int      ranked;        /* 1: RankedWTNode.putLong override is active */
    int64_t  size;          /* bits */
} orc_bits;

void  orc_bits_init(orc_bits* b, uint8_t* buf, int64_t limit, int64_t size_bits, int ranked);
void  orc_bits_write(orc_bits* b, jlong bits, int nbits);
void  orc_bits_flush(orc_bits* b);
void  orc_bits_rewind(orc_bits* b);
jlong orc_bits_peek(orc_bits* b, int nbits);
int   orc_bits_skip(orc_bits* b, int nbits);
jlong orc_bits_read(orc_bits* b, int nbits);
void  orc_bits_align(orc_bits* b);

int64_t orc_ranked_count_raw(const uint8_t* buf, int64_t limit, int64_t idx);
int     orc_ranked_get_raw(const uint8_t* buf, int64_t idx);
int64_t orc_ranked_find_zero_range(const uint8_t* buf, int64_t limit, int64_t n, int64_t lo, int64_t hi);
int64_t orc_ranked_find_one_range(const uint8_t* buf, int64_t limit, int64_t n, int64_t lo, int64_t hi);

#endif

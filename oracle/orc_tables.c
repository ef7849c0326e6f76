/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see orc_internal.h).
 *
 * Literal restatement of the byte-defining table code:
 *   algo/huffman/HuffmanEncodeTable.java      (code LENGTHS, tie-breaking matters)
 *   algo/deflate/DeflateEncodeTable.java      (length limiting + canonical reversed codes)
 *   algo/deflate/DeflateLookupTable.java      (decode table; also the HSWT "node naming" function)
 *   algo/deflate/DeflateLengthsTable.java     (RFC 1951 3.2.7 style serialisation of the lengths)
 *   algo/tree/HSWTShape.java                  (node sizes, serialized size)
 */
#include "orc_tables.h"
#include "gcz_oracle.h"
#include <limits.h>

/* ---- algo/huffman/HuffmanEncodeTable.java:48-111 ------------------------------------------ */
static void huffman_encode_table(jlong* counts /* clobbered, like the Java copy */, int n,
                                 jbyte* bit_lengths, jshort* table) {
    jint* bt = (jint*)calloc((size_t)n, sizeof(jint));   /* tree circular pointers */
    memset(bit_lengths, 0, (size_t)n);
    memset(table, 0, (size_t)n * sizeof(jshort));

    for (int i = 1; i < n; i++) {
        int idx1 = 0, idx2 = 0;
        jlong min1 = INT64_MAX, min2 = INT64_MAX;
        for (int j = 0; j < n; j++) {
            const jlong fq = counts[j];
            if (fq > 0) {
                if (fq < min1) {
                    idx2 = idx1; min2 = min1;
                    idx1 = j;    min1 = fq;
                } else if (fq < min2) {
                    idx2 = j;    min2 = fq;
                }
            }
        }
        if (min2 == INT64_MAX) {
            if (i == 1) {
                /* all characters are the same, but one bit is still needed */
                bit_lengths[idx1] = 1;
                table[idx1] = 1;
            }
            break;
        }
        counts[idx1] = INT64_MIN;
        counts[idx2] = min1 + min2;

        if (bt[idx1] == 0) {
            bt[idx1] = bt[idx2] < 0 ? bt[idx2] : ~idx2;
            bt[idx2] = ~idx1;
        } else if (bt[idx2] == 0) {
            bt[idx2] = bt[idx1];
            bt[idx1] = ~idx2;
        } else {
            const jint idx = bt[idx1];
            bt[idx1] = bt[idx2];
            bt[idx2] = idx;
        }

        jint idx = idx1;
        do {
            idx = ~bt[idx];
            table[idx] = (jshort)j_shl(table[idx], 1);
            bit_lengths[idx]++;
        } while (idx != idx2);
        do {
            idx = ~bt[idx];
            table[idx] = (jshort)(j_shl(table[idx], 1) | 1);
            bit_lengths[idx]++;
        } while (idx != idx1);
    }
    free(bt);
}

/* DeflateEncodeTable.reverse :175-180 and DeflateLookupTable.reverse :175-180 (identical) */
static jint reverse16(jint i) {
    i = j_shl(i & 0x00005555, 1) | (j_ushr(i, 1) & 0x00005555);
    i = j_shl(i & 0x00003333, 2) | (j_ushr(i, 2) & 0x00003333);
    i = j_shl(i & 0x000000F0F, 4) | (j_ushr(i, 4) & 0x00000F0F);
    return j_ushr(i, 8) | (j_shl(i, 8) & 0xFFFF);
}

static int cmp_jlong(const void* a, const void* b) {
    const jlong x = *(const jlong*)a, y = *(const jlong*)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}

/* algo/deflate/DeflateEncodeTable.java:63-148 */
static void restrict_lengths(const jlong* counts, int n, jbyte* bit_lengths, const jbyte max_bits) {
    jlong count = 0;
    jint* bl_count = (jint*)calloc((size_t)(n > 64 ? n : 64), sizeof(jint));
    for (int i = 0; i < n; i++) {
        const int bl = bit_lengths[i];
        if (bl > 0) {
            count += bl;
            bl_count[bl]++;       /* Java: index < bit_lengths.length; lengths never reach n here */
        }
    }
    if (count <= 1) { free(bl_count); return; }

    jint nodes = 1;
    for (int i = 1; i <= max_bits && nodes > 0; i++) {
        nodes = j_shl(nodes, 1);
        if (bl_count[i] != 0) nodes -= bl_count[i];
    }
    free(bl_count);

    if (nodes > 0) {
        nodes = -nodes;
        for (int i = 0; i < n; i++) {
            if (bit_lengths[i] > max_bits) {
                bit_lengths[i] = max_bits;
                nodes++;
            }
        }
        /* 0xLL_CCCCCCCC_IIII (I - index, C - count, L - bit length) */
        jlong* list = (jlong*)malloc((size_t)n * sizeof(jlong));
        for (int i = 0; i < n; i++) {
            list[i] = j_lshl((jlong)bit_lengths[i], 48) | j_lshl((jlong)counts[i], 16) | (jlong)i;
        }
        qsort(list, (size_t)n, sizeof(jlong), cmp_jlong);   /* Arrays.sort(long[]): values are distinct */

        do {
            int done = 0;
            for (int i = max_bits - 1; i > 0 && !done; i--) {
                for (int level = i; level < max_bits && !done; level++) {
                    const jlong bit_length = j_lshl((jlong)(level + 1), 48);
                    for (int j = 0; j < n; j++) {
                        const jbyte bl = (jbyte)j_lushr(list[j], 48);
                        if (bl == level) {
                            list[j] = (list[j] & (jlong)0xFF00FFFFFFFFFFFFULL) | bit_length;
                            nodes -= j_shl(1, max_bits - 1 - level);
                            if (nodes <= 0) { done = 1; break; }     /* break loop; */
                        }
                    }
                }
            }
            for (int level = max_bits; nodes < 0 && level > 0; level--) {
                const jlong bit_length = j_lshl((jlong)(level - 1), 48);
                for (int i = n - 1; nodes < 0 && i >= 0; i--) {
                    const jbyte bl = (jbyte)j_lushr(list[i], 48);
                    if (bl == level) {
                        list[i] = (list[i] & (jlong)0xFF00FFFFFFFFFFFFULL) | bit_length;
                        nodes += j_shl(1, max_bits - level);
                    }
                }
            }
        } while (nodes != 0);

        for (int i = 0; i < n; i++) {
            bit_lengths[(jint)(list[i] & 0xFFFF)] = (jbyte)j_lushr(list[i], 48);
        }
        free(list);
    }
}

/* algo/deflate/DeflateEncodeTable.java:150-173 */
static int remap_codes(int n, const jbyte* bit_lengths, jshort* table, const jbyte max_bits) {
    jint bl_count[17] = {0}, next_code[17] = {0};
    for (int i = 0; i < n; i++) {
        const int bl = bit_lengths[i];
        if (bl > 0) {
            if (bl > max_bits) return -1;      /* Java: ArrayIndexOutOfBoundsException */
            bl_count[bl]++;
        }
    }
    for (jint bits = 1, code = 0; bits <= max_bits; bits++) {
        code = j_shl(code + bl_count[bits - 1], 1);
        next_code[bits] = code;
    }
    for (int i = 0; i < n; i++) {
        const int len = bit_lengths[i];
        if (len != 0) {
            table[i] = (jshort)j_shr(reverse16(next_code[len]), 16 - len);
            next_code[len]++;
        }
    }
    return 0;
}

/* new DeflateEncodeTable(counts, max_bits)  :52-56 */
int orc_encode_table(const jlong* counts, int n, int max_bits, jbyte* bit_lengths, jshort* table) {
    jlong* copy = (jlong*)malloc((size_t)n * sizeof(jlong));
    memcpy(copy, counts, (size_t)n * sizeof(jlong));        /* Arrays.copyOf */
    huffman_encode_table(copy, n, bit_lengths, table);
    free(copy);
    restrict_lengths(counts, n, bit_lengths, (jbyte)max_bits);
    return remap_codes(n, bit_lengths, table, (jbyte)max_bits);
}

/* new DeflateEncodeTable(bit_lengths)  :47-50 */
int orc_encode_table_from_lengths(int n, const jbyte* bit_lengths, jshort* table) {
    memset(table, 0, (size_t)n * sizeof(jshort));
    return remap_codes(n, bit_lengths, table, 15);
}

int32_t orc_deflate_encode_table(const int64_t* counts, int32_t n, int32_t max_bits,
                                 int8_t* bit_lengths, int16_t* table) {
    return orc_encode_table(counts, n, max_bits, bit_lengths, table);
}

/* ---- algo/deflate/DeflateLookupTable.java:40-115 ------------------------------------------- */
int orc_lookup_build(orc_lookup* t, const jbyte* bit_lengths, int n) {
    enum { MAX_BITS = 15 };
    jint bl_count[MAX_BITS + 1] = {0};
    for (int i = 0; i < n; i++) {
        const int bl = bit_lengths[i];
        if (bl > 0) bl_count[bl]++;
    }
    jint next_code[MAX_BITS + 1] = {0};

    jint tree_size = 512;
    for (jint bits = 1, code = 0; bits <= MAX_BITS; bits++) {
        const jint count = bl_count[bits];
        const jint tail = j_ushr(j_shl(code, 25), 42 - bits);
        code += j_shl(count, 16 - bits);
        if (bits > 9) {
            tree_size += j_ushr(j_shl(code - next_code[bits - 1], bits - 9), 7) + tail;
        }
        next_code[bits] = code;
    }
    if (tree_size < 512 || tree_size > (1 << 20)) return -1;
    t->size = tree_size;
    t->table = (jshort*)calloc((size_t)tree_size, sizeof(jshort));

    if (tree_size > 512) {
        for (jint bits = 10, ptr = 512, step = 64; bits <= MAX_BITS; bits++, step = j_ushr(step, 1)) {
            const jint ext = bits - 9;
            const jint code = next_code[bits - 1];
            const jint next = next_code[bits];
            const jint tail = j_ushr(j_shl(code, 25), 42 - bits);
            if (tail > 0) {
                ptr += tail;
                const jint idx = reverse16(code) & 511;
                t->table[idx] = (jshort)((t->table[idx] & (jint)0xFFFFFFF0) | ext);
            }
            for (jint i = code; i < next; i += step, ptr++) {
                const jint idx = reverse16(i) & 511;
                if (t->table[idx] == 0) {
                    if (ptr >= tree_size) { /* Java would throw later on access; keep going */ }
                    t->table[idx] = (jshort)(j_shl(j_neg(ptr), 4) | ext);
                }
            }
        }
    }

    for (int i = 0; i < n; i++) {
        const jint bits = bit_lengths[i];
        if (bits > 0) {
            const jint code = next_code[bits - 1];
            jint revcode = reverse16(code);
            if (bits < 9) {
                for (jint j = revcode, l = j_shl(1, bits); j < 512; j += l) {
                    t->table[j] = (jshort)(j_shl(i, 4) | bits);
                }
            } else if (bits == 9) {
                t->table[revcode] = (jshort)(j_shl(i, 4) | bits);
            } else {
                const jshort ptr = t->table[revcode & 0x1FF];
                const jint ext = ptr & 15;
                jint idx = j_ushr(j_shl(code, 25), 32 - ext) - j_shr((jint)ptr, 4);
                for (jint k = idx + j_shl(1, ext + 9 - bits); idx < k; idx++) {
                    if (idx < 0 || idx >= tree_size) return -2;     /* Java: AIOOBE */
                    t->table[idx] = (jshort)(j_shl(i, 4) | (bits - 9));
                }
            }
            next_code[bits - 1] = code + j_shl(1, 16 - bits);
        }
    }
    return 0;
}

void orc_lookup_free(orc_lookup* t) { free(t->table); t->table = NULL; t->size = 0; }

/* getSymbol(int code) :145-153 */
int orc_lookup_symbol(const orc_lookup* t, jint code) {
    const jint peek = code & 511;
    jint symbol = t->table[peek];
    if (symbol < 0) {
        const jint idx = j_ushr(reverse16(j_ushr(code, 9)), 16 - (symbol & 15)) - j_shr(symbol, 4);
        if (idx < 0 || idx >= t->size) return INT32_MIN;          /* Java: AIOOBE */
        symbol = t->table[idx];
    }
    return j_ushr(symbol, 4);
}

/* getSymbol(int code, int nbits) :162-173 */
int orc_lookup_symbol_nbits(const orc_lookup* t, jint code, jint nbits) {
    const jint peek = code & 511;
    jint symbol = t->table[peek];
    jint len = symbol & 15;
    if (symbol < 0) {
        len += 9;
        const jint idx = j_ushr(reverse16(j_ushr(code, 9)), 16 - len) - j_shr(symbol, 4);
        if (idx < 0 || idx >= t->size) return INT32_MIN;
        symbol = t->table[idx];
    }
    return nbits >= len ? j_ushr(symbol, 4) : INT32_MIN;
}

/* getSymbol(BitInputStream in) :124-137 */
int orc_lookup_symbol_stream(const orc_lookup* t, orc_bits* in) {
    const jint peek = (jint)(orc_bits_peek(in, 9) & 511);
    jint symbol = t->table[peek];
    const jint bits = symbol & 15;
    if (symbol >= 0) {
        orc_bits_skip(in, bits);
    } else {
        orc_bits_skip(in, 9);
        const jint idx = j_ushr(reverse16((jint)orc_bits_peek(in, bits)), 16 - bits) - j_shr(symbol, 4);
        if (idx < 0 || idx >= t->size) return INT32_MIN;
        symbol = t->table[idx];
        orc_bits_skip(in, symbol & 15);
    }
    return j_ushr(symbol, 4);
}

int32_t orc_lookup_get_symbol(const int8_t* bit_lengths, int32_t n, int32_t code) {
    orc_lookup t;
    if (orc_lookup_build(&t, bit_lengths, n) != 0) return INT32_MIN;
    const int r = orc_lookup_symbol(&t, code);
    orc_lookup_free(&t);
    return r;
}

int32_t orc_lookup_get_symbol_nbits(const int8_t* bit_lengths, int32_t n, int32_t code, int32_t nbits) {
    orc_lookup t;
    if (orc_lookup_build(&t, bit_lengths, n) != 0) return INT32_MIN;
    const int r = orc_lookup_symbol_nbits(&t, code, nbits);
    orc_lookup_free(&t);
    return r;
}

/* DeflateTablesTest.test / stress_test2: putSymbol ... rewind ... getSymbol(in) */
int64_t orc_deflate_stream_roundtrip(const uint8_t* data, int64_t n, int64_t cap) {
    jlong counts[256] = {0};
    for (int64_t i = 0; i < n; i++) counts[data[i]]++;
    jbyte bl[256]; jshort tb[256];
    if (orc_encode_table(counts, 256, 15, bl, tb) != 0) return -1;
    uint8_t* buf = (uint8_t*)calloc((size_t)cap + 8, 1);
    orc_bits b;
    orc_bits_init(&b, buf, cap, cap, 0);
    for (int64_t i = 0; i < n; i++) orc_bits_write(&b, tb[data[i]], bl[data[i]]);   /* putSymbol :113-115 */
    orc_bits_rewind(&b);
    orc_lookup t;
    if (orc_lookup_build(&t, bl, 256) != 0) { free(buf); return -2; }
    int64_t bad = 0;
    for (int64_t i = 0; i < n; i++) {
        if (orc_lookup_symbol_stream(&t, &b) != data[i]) bad++;
    }
    orc_lookup_free(&t);
    free(buf);
    return bad;
}

/* ---- algo/deflate/DeflateLengthsTable.java -------------------------------------------------- */
static const jbyte CL_ORDER[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };

/* counts() :173-207 */
static int lengths_counts(const jbyte* bit_lengths, int len_n, jlong* counts) {
    for (int i = 0, len = 0, count = 0, n = len_n - 1; i <= n; i++) {
        if (len != bit_lengths[i] || i == n) {
            while (count >= 3) {
                if (len != 0)         { counts[16]++; count -= 6; }
                else if (count <= 10) { counts[17]++; count -= 10; }
                else                  { counts[18]++; count -= 138; }
            }
            while (count-- > 0) counts[len]++;
            len = bit_lengths[i];
            counts[len]++;
            count = 0;
        } else {
            count++;
        }
    }
    int hclen = 18;
    do {
        if (counts[CL_ORDER[hclen]] > 0) break;
    } while (--hclen >= 0);
    return hclen;
}

/* length() :136-171 */
int orc_lengths_table_bits(const jbyte* bit_lengths, int len_n) {
    jlong counts[19] = {0};
    const int hclen = lengths_counts(bit_lengths, len_n, counts);
    int bits = 7 + hclen * 3;
    jbyte tbl[19]; jshort ttb[19];
    orc_encode_table(counts, 19, 15, tbl, ttb);
    for (int i = 0, len = 0, count = 0, n = len_n - 1; i <= n; i++) {
        if (len != bit_lengths[i] || i == n) {
            while (count >= 3) {
                if (len != 0)         { bits += tbl[16] + 2; count -= 6; }
                else if (count <= 10) { bits += tbl[17] + 3; count -= 10; }
                else                  { bits += tbl[18] + 7; count -= 138; }
            }
            while (count-- > 0) bits += tbl[len];
            len = bit_lengths[i];
            bits += tbl[len];
            count = 0;
        } else {
            count++;
        }
    }
    return bits;
}

int32_t orc_deflate_lengths_bits(const int8_t* bit_lengths, int32_t n) {
    return orc_lengths_table_bits(bit_lengths, n);
}

static inline int imin(int a, int b) { return a < b ? a : b; }

/* write() :82-125 */
void orc_lengths_table_write(const jbyte* d_tree, int len_n, orc_bits* out) {
    jlong counts[19] = {0};
    const int hclen = lengths_counts(d_tree, len_n, counts);
    orc_bits_write(out, hclen - 3, 4);
    jbyte tbl[19]; jshort ttb[19];
    orc_encode_table(counts, 19, 7, tbl, ttb);
    for (int i = 0; i <= hclen; i++) {
        orc_bits_write(out, tbl[CL_ORDER[i]], 3);
    }
#define PUT_SYMBOL(s) orc_bits_write(out, ttb[(s)], tbl[(s)])
    for (int i = 0, len = 0, count = 0, n = len_n - 1; i <= n; i++) {
        if (len != d_tree[i] || i == n) {
            while (count >= 3) {
                if (len != 0) {
                    PUT_SYMBOL(16);
                    count -= 3;
                    orc_bits_write(out, imin(count, 3), 2);
                    count -= 3;
                } else if (count <= 10) {
                    PUT_SYMBOL(17);
                    count -= 3;
                    orc_bits_write(out, imin(count, 7), 3);
                    count -= 7;
                } else {
                    PUT_SYMBOL(18);
                    count -= 11;
                    orc_bits_write(out, imin(count, 127), 7);
                    count -= 127;
                }
            }
            while (count-- > 0) PUT_SYMBOL(len);
            len = d_tree[i];
            PUT_SYMBOL(len);
            count = 0;
        } else {
            count++;
        }
    }
#undef PUT_SYMBOL
}

/* read ctor :47-80 */
int orc_lengths_table_read(orc_bits* in, jbyte* d_tree, int len_n) {
    memset(d_tree, 0, (size_t)len_n);
    const int hclen = (int)((orc_bits_read(in, 4) & 15) + 4);
    jbyte l_tree[19] = {0};
    for (int i = 0; i < hclen; i++) {
        l_tree[CL_ORDER[i]] = (jbyte)(orc_bits_read(in, 3) & 7);
    }
    orc_lookup table;
    if (orc_lookup_build(&table, l_tree, 19) != 0) return -1;
    jbyte symbol = 0;
    int rc = 0;
    for (int i = 0, n = len_n; i < n;) {
        const int sym = orc_lookup_symbol_stream(&table, in);
        if (sym == INT32_MIN) { rc = -2; break; }
        const jbyte code = (jbyte)sym;
        if (code <= 15) {
            d_tree[i++] = symbol = code;
        } else if (code == 16) {
            const int rep = (int)((orc_bits_read(in, 2) & 3) + 3);
            for (int j = 0; j < rep; j++, i++) {
                if (i >= n) { rc = -3; break; }          /* Java: AIOOBE */
                d_tree[i] = symbol;
            }
            if (rc) break;
        } else if (code == 17) {
            i += (int)((orc_bits_read(in, 3) & 7) + 3);
        } else if (code == 18) {
            i += (int)((orc_bits_read(in, 7) & 127) + 11);
        }
    }
    orc_lookup_free(&table);
    return rc;
}

/* ---- algo/tree/HSWTShape.java:55-87 ---------------------------------------------------------- */
int32_t orc_shape_from_counts(const int64_t counts[256], orc_shape* out) {
    memset(out, 0, sizeof(*out));
    if (orc_encode_table(counts, 256, 15, out->bit_lengths, out->table) != 0) return -1;
    orc_lookup decode;
    if (orc_lookup_build(&decode, out->bit_lengths, 256) != 0) return -2;

    jlong len = 0;
    for (int i = 0; i < 256; i++) {
        if (counts[i] > 0) {
            len += counts[i];
            const jint code = out->table[i];
            for (int j = 0, n = out->bit_lengths[i]; j < n; j++) {
                jint idx = code & j_ushr(0x0000FFFF, 16 - j);
                idx |= j_ushr(0x8000, 15 - j);
                idx = orc_lookup_symbol(&decode, idx);
                if (idx < 0 || idx > 255) { orc_lookup_free(&decode); return -3; }
                out->node_bits[idx] = (jint)((jlong)out->node_bits[idx] + counts[i]);   /* int += long */
            }
        }
    }
    orc_lookup_free(&decode);

    out->table_bytes = j_ushr(orc_lengths_table_bits(out->bit_lengths, 256) + 7, 3);
    jlong sz = out->table_bytes;
    for (int i = 0; i < 256; i++) {
        if (out->node_bits[i] > 0) sz += orc_ranked_bytes(out->node_bits[i]);
    }
    out->length = len;
    out->size = sz;
    return 0;
}

/* HSWTShape.write :111-115 */
int64_t orc_shape_write(const orc_shape* s, uint8_t* out, int64_t cap) {
    orc_bits b;
    orc_bits_init(&b, out, cap, cap, 0);
    orc_lengths_table_write(s->bit_lengths, 256, &b);
    orc_bits_flush(&b);
    return b.pos;
}

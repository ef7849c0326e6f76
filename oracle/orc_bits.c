/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see orc_internal.h).
 *
 * Literal restatement of
 *   io/AbstractBitStream.java   (LSB-first 64-bit little-endian bit writer/reader)
 *   io/BitBuffer.java
 *   algo/tree/RankedWTNode.java (bit vector with interleaved 16-bit / 64-bit rank counters)
 * The byte behaviour of these three classes defines the .gcz/.gcx formats.
 */
#include "orc_bits.h"
#include "gcz_oracle.h"

static __thread uint64_t g_rank_calls = 0;
uint64_t orc_rank_calls(void) { return g_rank_calls; }
void orc_rank_calls_reset(void) { g_rank_calls = 0; }

void orc_bits_init(orc_bits* b, uint8_t* buf, int64_t limit, int64_t size_bits, int ranked) {
    b->buf = buf; b->pos = 0; b->limit = limit;
    b->value = 0; b->bits_left = 0; b->ranked = ranked; b->size = size_bits;
}

/* io/AbstractBitStream.java:172-182 getLong(long index): bytewise when < 8 bytes remain */
static jlong abs_get_long_at(const uint8_t* buf, int64_t limit, int64_t index) {
    if (limit - index < 8) {
        jlong l = 0;
        for (int i = 0; index < limit; i += 8, index++) {
            l |= j_lshl((jlong)(buf[index] & 0xFF), i);
        }
        return l;
    }
    return le_get64(buf + index);
}

/* io/AbstractBitStream.java:160-170 getLong(): sequential */
static jlong abs_get_long(orc_bits* b) {
    if (b->limit - b->pos < 8) {
        jlong l = b->buf[b->pos++] & 0xFF;       /* Java would throw on an empty buffer */
        for (int i = 8; b->pos < b->limit; i += 8) {
            l |= j_lshl((jlong)(b->buf[b->pos++] & 0xFF), i);
        }
        return l;
    }
    jlong v = le_get64(b->buf + b->pos);
    b->pos += 8;
    return v;
}

/* io/AbstractBitStream.java:184-193 putLong */
static void abs_put_long(orc_bits* b, jlong value) {
    if (b->limit - b->pos < 8) {
        for (int i = 0, n = 64 - b->bits_left; b->pos < b->limit && i <= n; i += 8) {
            b->buf[b->pos++] = (uint8_t)(j_lushr(value, i) & 0xFF);
        }
    } else {
        le_put64(b->buf + b->pos, value);
        b->pos += 8;
    }
}

/* algo/tree/RankedWTNode.java:98-122 count(idx): ones in [0, idx] */
int64_t orc_ranked_count_raw(const uint8_t* buf, int64_t limit, int64_t idx) {
    g_rank_calls++;
    jlong count = 0;
    const jlong nlidx = j_lushr(idx, 16);
    const jlong nsidx = j_lushr(idx, 9) & 127;
    const jlong spos = nsidx * 66;
    jlong lpos = 0;
    if (nlidx > 0) {
        lpos = nlidx * 8454;
        count = le_get64(buf + (jint)(lpos - 8));
    }
    jlong bpos = lpos + spos;
    if (nsidx > 0) {
        count += le_get16u(buf + (jint)(bpos - 2)) & 0xFFFF;
    }
    for (jint n = (jint)(bpos + (j_lushr(idx, 3) & 56)); bpos < n; bpos += 8) {
        count += j_bitcount64(abs_get_long_at(buf, limit, bpos));
    }
    return count + j_bitcount64(j_lshl(abs_get_long_at(buf, limit, bpos), (int)(63 - (idx & 63))));
}

/* algo/tree/RankedWTNode.java:81-84 */
int orc_ranked_get_raw(const uint8_t* buf, int64_t idx) {
    const jlong pos = j_lushr(idx, 3) + j_lushr(idx, 9) * 2 + j_lushr(idx, 16) * 6;
    /* (byte)((buf.get(pos) >>> (idx & 7)) & 1): buf.get() is a signed byte promoted to int, the
     * masked result is the same as for the unsigned byte because idx&7 <= 7 */
    return (int)(((jint)(jbyte)buf[(jint)pos] >> (idx & 7)) & 1);
}

/* algo/tree/RankedWTNode.java:228-245 putLong override, falling through to the base putLong */
static void ranked_put_long(orc_bits* b, jlong value) {
    const jint pos = (jint)b->pos;
    jlong nlong = pos - ((pos / 8454) * 6);
    nlong -= ((nlong / 66) << 1);
    if ((nlong & 0x1FFF) == 0 && nlong > 0) {
        le_put64(b->buf + b->pos, orc_ranked_count_raw(b->buf, b->limit, (nlong << 3) - 1));
        g_rank_calls--;                       /* construction-time count() is not a query */
        b->pos += 8;
    } else if ((nlong & 63) == 0 && nlong > 0) {
        jint count = (nlong & 0x1FFF) > 64 ? (le_get16u(b->buf + pos - 66) & 0xFFFF) : 0;
        for (jint i = pos - 64; i < pos; i += 8) {
            count += j_bitcount64(le_get64(b->buf + i));
        }
        le_put16(b->buf + b->pos, (jshort)count);
        b->pos += 2;
    }
    abs_put_long(b, value);
}

static void put_long(orc_bits* b, jlong value) {
    if (b->ranked) ranked_put_long(b, value); else abs_put_long(b, value);
}

/* io/AbstractBitStream.java:125-141 */
void orc_bits_write(orc_bits* b, jlong bits, int nbits) {
    if (b->bits_left > nbits) {
        b->value |= j_lshl(bits, 64 - b->bits_left);
        b->bits_left -= nbits;
    } else if (b->bits_left == 0) {
        b->value = bits;
        b->bits_left = (jbyte)(64 - nbits);
    } else if (b->bits_left < nbits) {
        put_long(b, b->value | j_lshl(bits, 64 - b->bits_left));
        b->value = j_lshr(bits, b->bits_left);
        b->bits_left += (64 - nbits);
    } else {
        put_long(b, b->value | j_lshl(bits, 64 - b->bits_left));
        b->bits_left = 0;
    }
}

/* io/AbstractBitStream.java:150-158 */
void orc_bits_flush(orc_bits* b) {
    if (b->bits_left > 0) {
        const int64_t pos = b->pos;
        const int len = (71 - b->bits_left) >> 3;
        put_long(b, b->value);
        b->pos = pos + len;
        b->bits_left = 0;
    }
}

/* io/AbstractBitStream.java:116-122 */
void orc_bits_rewind(orc_bits* b) {
    if (b->bits_left > 0) {
        put_long(b, b->value);
        b->bits_left = 0;
    }
    b->pos = 0;
}

/* io/AbstractBitStream.java:64-79 */
jlong orc_bits_peek(orc_bits* b, int nbits) {
    if (b->bits_left == 0) {
        b->value = abs_get_long(b);
        b->bits_left = 64;
        return b->value;
    }
    const jlong last = j_lushr(b->value, 64 - b->bits_left);
    if (b->bits_left >= nbits) return last;
    const jlong next = abs_get_long_at(b->buf, b->limit, b->pos);
    return last | j_lshl(next, b->bits_left);
}

/* io/AbstractBitStream.java:82-92; returns -1 where Java throws EOFException */
int orc_bits_skip(orc_bits* b, int nbits) {
    b->bits_left -= nbits;
    if (b->bits_left < 0) {
        if (b->pos < b->limit) {
            b->bits_left += 64;
            b->value = abs_get_long(b);
        } else {
            return -1;
        }
    }
    return 0;
}

/* io/AbstractBitStream.java:95-113 */
jlong orc_bits_read(orc_bits* b, int nbits) {
    if (b->bits_left == 0) {
        b->value = abs_get_long(b);
        b->bits_left = (jbyte)(64 - nbits);
        return b->value;
    } else {
        jlong bits = j_lushr(b->value, 64 - b->bits_left);
        if (b->bits_left < nbits) {
            b->value = abs_get_long(b);
            bits |= j_lshl(b->value, b->bits_left);
            b->bits_left += (64 - nbits);
        } else {
            b->bits_left -= nbits;
        }
        return bits;
    }
}

/* io/BitBuffer.java:44-47 */
void orc_bits_align(orc_bits* b) {
    b->pos = b->pos - (b->bits_left >> 3);
    b->bits_left = 0;
}

/* ---------------------------------------------------------------------------------------------
 * exported test entry points
 * ------------------------------------------------------------------------------------------- */
int64_t orc_bitbuffer_write(const int64_t* vals, const int32_t* nbits, int32_t count,
                            uint8_t* out, int64_t cap) {
    orc_bits b;
    orc_bits_init(&b, out, cap, cap, 0);
    for (int i = 0; i < count; i++) orc_bits_write(&b, vals[i], nbits[i]);
    orc_bits_flush(&b);
    return b.pos;
}

void orc_bitbuffer_write_read(const int64_t* vals, const int32_t* nbits, int32_t count,
                              int64_t cap, const int32_t* rn, int32_t rcount, int64_t* rout) {
    uint8_t* buf = (uint8_t*)calloc((size_t)cap + 8, 1);
    orc_bits b;
    orc_bits_init(&b, buf, cap, cap, 0);
    for (int i = 0; i < count; i++) orc_bits_write(&b, vals[i], nbits[i]);
    orc_bits_rewind(&b);
    for (int i = 0; i < rcount; i++) rout[i] = orc_bits_read(&b, rn[i]);
    free(buf);
}

/* algo/tree/RankedWTNode.java:60-67 */
int64_t orc_ranked_bytes(int64_t len) {
    return j_lushr(len - 1, 16) * 6 + j_lushr(len - 1, 9) * 2 + j_lushr(len + 7, 3);
}

int64_t orc_ranked_write(const uint8_t* bits, int64_t len, uint8_t* out) {
    orc_bits b;
    orc_bits_init(&b, out, orc_ranked_bytes(len), len, 1);
    for (int64_t i = 0; i < len; i++) orc_bits_write(&b, bits[i] & 1, 1);
    orc_bits_flush(&b);
    return b.limit;
}

int32_t orc_ranked_get(const uint8_t* buf, int64_t len, int64_t idx) {
    (void)len;
    return orc_ranked_get_raw(buf, idx);
}

int64_t orc_ranked_count(const uint8_t* buf, int64_t len, int64_t idx) {
    return orc_ranked_count_raw(buf, orc_ranked_bytes(len), idx);
}

/* algo/tree/RankedWTNode.java:154-179 */
int64_t orc_ranked_find_zero_range(const uint8_t* buf, int64_t limit, int64_t n, int64_t lo, int64_t hi) {
    while (lo < hi) {
        const jlong clo = lo - orc_ranked_count_raw(buf, limit, lo) + 1;
        const jlong chi = hi - orc_ranked_count_raw(buf, limit, hi) + 1;
        if (clo >= chi) {
            return n == clo && orc_ranked_get_raw(buf, lo) == 0 ? lo : -1;
        }
        const jlong mid = lo + (jlong)(((hi - lo) * (double)(n - clo)) / (chi - clo));
        const jlong cmid = mid - orc_ranked_count_raw(buf, limit, mid) + 1;
        if (n < cmid) hi = mid - 1;
        else if (n > cmid) lo = mid + 1;
        else if (orc_ranked_get_raw(buf, mid) == 0) return mid;
        else hi = mid - 1;
    }
    if (lo == hi && orc_ranked_get_raw(buf, lo) == 0 && n == lo - orc_ranked_count_raw(buf, limit, lo) + 1) {
        return lo;
    }
    return -1;
}

/* algo/tree/RankedWTNode.java:181-205 */
int64_t orc_ranked_find_one_range(const uint8_t* buf, int64_t limit, int64_t n, int64_t lo, int64_t hi) {
    while (lo < hi) {
        const jlong clo = orc_ranked_count_raw(buf, limit, lo);
        const jlong chi = orc_ranked_count_raw(buf, limit, hi);
        if (clo >= chi) {
            return n == clo && orc_ranked_get_raw(buf, lo) > 0 ? lo : -1;
        }
        const jlong mid = lo + (jlong)(((hi - lo) * (double)(n - clo)) / (chi - clo));
        const jlong cmid = orc_ranked_count_raw(buf, limit, mid);
        if (n < cmid) hi = mid - 1;
        else if (n > cmid) lo = mid + 1;
        else if (orc_ranked_get_raw(buf, mid) > 0) return mid;
        else hi = mid - 1;
    }
    if (lo == hi && orc_ranked_get_raw(buf, lo) > 0 && n == orc_ranked_count_raw(buf, limit, lo)) {
        return lo;
    }
    return -1;
}

int64_t orc_ranked_find_zero(const uint8_t* buf, int64_t len, int64_t n) {   /* :140-142 */
    return orc_ranked_find_zero_range(buf, orc_ranked_bytes(len), n, n - 1, len - 1);
}
int64_t orc_ranked_find_one(const uint8_t* buf, int64_t len, int64_t n) {    /* :150-152 */
    return orc_ranked_find_one_range(buf, orc_ranked_bytes(len), n, n - 1, len - 1);
}

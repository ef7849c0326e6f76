/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C99) of the FM-index hot path of redmitry/gecoz.  It exists to
 * CHECK the CUDA product in gecoz_b200/, never to serve as (or behind) it.  Only tests/,
 * __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of bench.py may
 * link, load or call anything in this directory.
 *
 * Pinning status: the reference is Java and there is no JVM in the build image or on the
 * GPU box, so the reference itself cannot be executed.  The restatement is pinned by the
 * reference's own known-answer material (tests/test_oracle_kat.py):
 *   - java/nova-io/src/test/.../BitBufferTest.java:18-64  (bit packing + short-buffer flush)
 *   - java/nova-algo/src/test/.../DeflateTablesTest.java:57-198 (code-gen / lookup properties)
 *   - doc/GECOZ.pdf Table 3 (IndexWaveletTree level bits), Table 1/2 + Fig. 1 (layout)
 * Everything else (Huffman tie-breaks, shape header bytes, HSWT node order, GSSA results)
 * is "parity unpinned by reference tests": it is a literal transliteration of the cited
 * Java lines, cross-checked by invariants (sizes reserved == bytes written, re-open
 * recovers node lengths, find == naive search).
 *
 * This header: Java integer semantics (wrap-around, masked shift counts) used by the
 * literal transliterations.
 */
#ifndef ORC_INTERNAL_H
#define ORC_INTERNAL_H

#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>

typedef int8_t   jbyte;
typedef int16_t  jshort;
typedef int32_t  jint;
typedef int64_t  jlong;

/* Java `int` shifts use only the low 5 bits of the count, `long` shifts the low 6. */
static inline jint  j_shl (jint x, int s)  { return (jint)((uint32_t)x << (s & 31)); }
static inline jint  j_shr (jint x, int s)  { return x >> (s & 31); }               /* arithmetic */
static inline jint  j_ushr(jint x, int s)  { return (jint)((uint32_t)x >> (s & 31)); }
static inline jlong j_lshl (jlong x, int s) { return (jlong)((uint64_t)x << (s & 63)); }
static inline jlong j_lshr (jlong x, int s) { return x >> (s & 63); }
static inline jlong j_lushr(jlong x, int s) { return (jlong)((uint64_t)x >> (s & 63)); }
static inline jint  j_add(jint a, jint b)  { return (jint)((uint32_t)a + (uint32_t)b); }
static inline jint  j_sub(jint a, jint b)  { return (jint)((uint32_t)a - (uint32_t)b); }
static inline jint  j_neg(jint a)          { return (jint)(0u - (uint32_t)a); }

static inline int j_nlz32(jint x)  { return x == 0 ? 32 : __builtin_clz((uint32_t)x); }
static inline int j_nlz64(jlong x) { return x == 0 ? 64 : __builtin_clzll((uint64_t)x); }
static inline int j_bitcount64(jlong x) { return __builtin_popcountll((uint64_t)x); }

/* little-endian loads/stores on byte buffers (ByteBuffer.order(LITTLE_ENDIAN)) */
static inline jlong le_get64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return (jlong)v; }
static inline void  le_put64(uint8_t* p, jlong v) { memcpy(p, &v, 8); }
static inline jint  le_get16u(const uint8_t* p) { return (jint)p[0] | ((jint)p[1] << 8); }
static inline void  le_put16(uint8_t* p, jint v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }

#endif

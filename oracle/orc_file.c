/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see orc_internal.h).
 *
 * Literal restatement of the container-level, byte-defining host code:
 *   fmt/GecozRefBlockHeader.java:90-136   (.gcz block header, header hash)
 *   fmt/GecozSSABlockHeader.java:69-78    (.gcx block header)
 *   fmt/GecozRefBlock.java:38-71, fasta/TFastaSequence.java:46-52, tools/GecoIndex.java:72-98
 *                                         (which sequences share a block, and every order)
 */
#include "orc_internal.h"
#include "gcz_oracle.h"

/* getBlockHeaderHash :120-128 */
int64_t orc_header_hash(const char* const* headers, int32_t nh) {
    jlong hash = 1125899906842597LL;
    for (int h = 0; h < nh; h++) {
        for (const unsigned char* p = (const unsigned char*)headers[h]; *p; p++) {
            hash = (jlong)(((uint64_t)hash << 5) - (uint64_t)hash + (uint64_t)*p);
        }
    }
    return hash;
}

/* getBlockHeaderLength :130-136 */
int32_t orc_ref_header_len(const char* const* headers, int32_t nh) {
    int len = 26;
    for (int h = 0; h < nh; h++) len += (int)strlen(headers[h]) + 1;
    return len;
}

/* GecozRefBlockHeader.write :90-101 (buffer is LITTLE_ENDIAN, fmt/GecozFileWriter.java:139) */
int32_t orc_ref_header_write(const char* const* headers, int32_t nh, int64_t size, int64_t len, uint8_t* out) {
    uint8_t* p = out;
    memcpy(p, "GecozBWT", 8); p += 8;
    *p++ = 1;
    le_put64(p, size); p += 8;
    le_put64(p, len);  p += 8;
    for (int h = 0; h < nh; h++) {
        const size_t l = strlen(headers[h]);
        memcpy(p, headers[h], l); p += l;
        *p++ = 0;
    }
    *p++ = 0;
    return (int32_t)(p - out);
}

/* GecozSSABlockHeader.write :69-74 */
int32_t orc_ssa_header_write(const char* const* headers, int32_t nh, int64_t len, uint8_t* out) {
    uint8_t* p = out;
    memcpy(p, "GecozSSA", 8); p += 8;
    *p++ = 1;
    le_put64(p, len); p += 8;
    le_put64(p, orc_header_hash(headers, nh)); p += 8;
    return 25;
}

/* ---- block merge ---------------------------------------------------------------------------- */
typedef struct { int32_t length; const char* header; int32_t id; } seq_t;

/* TFastaSequence.compareTo :46-52 */
static int seq_cmp(const seq_t* a, const seq_t* b) {
    if (a->length != b->length) return a->length > b->length ? -1 : 1;
    /* String.compareTo on ASCII headers */
    const unsigned char* x = (const unsigned char*)a->header;
    const unsigned char* y = (const unsigned char*)b->header;
    for (; *x && *y; x++, y++) {
        if (*x != *y) return (int)*x - (int)*y;
    }
    return (int)strlen((const char*)x) - (int)strlen((const char*)y);
}

typedef struct { jint size; seq_t* seqs; int nseq, cap; } block_t;   /* GecozRefBlock: TreeSet + size */

static void block_add(block_t* b, const seq_t* s) {                   /* add(FastaSequence) :45-48 */
    int pos = 0, dup = 0;
    for (; pos < b->nseq; pos++) {
        const int c = seq_cmp(s, &b->seqs[pos]);
        if (c == 0) { dup = 1; break; }
        if (c < 0) break;
    }
    if (!dup) {                                                      /* TreeSet drops equal elements */
        if (b->nseq == b->cap) { b->cap = b->cap ? b->cap * 2 : 4; b->seqs = (seq_t*)realloc(b->seqs, (size_t)b->cap * sizeof(seq_t)); }
        memmove(&b->seqs[pos + 1], &b->seqs[pos], (size_t)(b->nseq - pos) * sizeof(seq_t));
        b->seqs[pos] = *s;
        b->nseq++;
    }
    b->size = j_add(b->size, s->length + 1);
}

/* GecozRefBlock.compareTo :62-69 */
static int block_cmp(const block_t* a, const block_t* b) {
    if (a->size != b->size) return a->size > b->size ? 1 : -1;
    return seq_cmp(&a->seqs[0], &b->seqs[0]);
}

/* comparator of the `sorted` TreeSet, tools/GecoIndex.java:88-96 */
static int block_cmp_file(const block_t* a, const block_t* b) {
    if (a->seqs[0].length != b->seqs[0].length) return a->seqs[0].length > b->seqs[0].length ? -1 : 1;
    return block_cmp(a, b);
}

typedef int (*bcmp_fn)(const block_t*, const block_t*);

/* TreeSet.add: ordered insert, equal elements are dropped (returns 0 when dropped) */
static int set_add(block_t** set, int* n, block_t* b, bcmp_fn cmp) {
    int pos = 0;
    for (; pos < *n; pos++) {
        const int c = cmp(b, set[pos]);
        if (c == 0) return 0;
        if (c < 0) break;
    }
    memmove(&set[pos + 1], &set[pos], (size_t)(*n - pos) * sizeof(block_t*));
    set[pos] = b;
    (*n)++;
    return 1;
}

int32_t orc_merge_blocks(const int32_t* lengths, const char* const* headers, int32_t nseq,
                         int32_t* block_of, int32_t* pos_in_block) {
    block_t* store = (block_t*)calloc((size_t)nseq + 1, sizeof(block_t));
    block_t** blocks = (block_t**)calloc((size_t)nseq + 2, sizeof(block_t*));
    block_t** sorted = (block_t**)calloc((size_t)nseq + 2, sizeof(block_t*));
    int nb = 0;
    for (int i = 0; i < nseq; i++) {
        block_of[i] = -1; pos_in_block[i] = -1;
        seq_t s = { lengths[i], headers[i], i };
        block_add(&store[i], &s);                                   /* new GecozRefBlock(seq) :43-46 */
        set_add(blocks, &nb, &store[i], block_cmp);                 /* GecoIndex.java:62-64 */
    }
    if (nb > 0) {
        const jint max_size = blocks[nb - 1]->size;                 /* :72 */
        while (nb > 1) {                                            /* :73-85 */
            block_t* first = blocks[0];
            block_t* second = blocks[1];
            memmove(&blocks[0], &blocks[2], (size_t)(nb - 2) * sizeof(block_t*));
            nb -= 2;
            const jint size = j_add(first->size, second->size);
            if (size > 0 && size <= max_size) {
                for (int k = 0; k < second->nseq; k++) block_add(first, &second->seqs[k]);
                set_add(blocks, &nb, first, block_cmp);
            } else {
                set_add(blocks, &nb, first, block_cmp);
                set_add(blocks, &nb, second, block_cmp);
                break;
            }
        }
    }
    int ns = 0;
    for (int i = 0; i < nb; i++) set_add(sorted, &ns, blocks[i], block_cmp_file);   /* :88-98 */
    for (int b = 0; b < ns; b++) {
        for (int k = 0; k < sorted[b]->nseq; k++) {
            block_of[sorted[b]->seqs[k].id] = b;
            pos_in_block[sorted[b]->seqs[k].id] = k;
        }
    }
    for (int i = 0; i < nseq; i++) free(store[i].seqs);
    free(store); free(blocks); free(sorted);
    return ns;
}

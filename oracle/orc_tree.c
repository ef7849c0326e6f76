/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see orc_internal.h).
 *
 * Literal restatement of
 *   algo/tree/HuffmanShapedWaveletTree.java   (build, mapNodes, occ, getRS)
 *   algo/tree/IndexWaveletTree.java           (build, get, find)
 *   algo/ssa/GSSAIndex.java                   (write ctor, read ctor, get)
 *   algo/ssa/GSSA.java                        (index, search, locate, find)
 *   fmt/GecozFileWriter.java:240-303          (BlockWriter.run, BWTDataSource)
 *   fmt/GecozFileReader.java:115-177          (read: shape, tree, sampling-factor recovery)
 */
#include "orc_tables.h"
#include "gcz_oracle.h"
#include <pthread.h>

/* ============================================================================================
 * HuffmanShapedWaveletTree
 * ========================================================================================== */
typedef struct {
    jbyte   bit_lengths[256];
    jshort  table[256];
    orc_lookup decode;
    /* nodes[256] indexed by node NAME (HuffmanShapedWaveletTree.java:41) */
    int      present[256];
    orc_bits nodes[256];
    int64_t  node_off[256];      /* byte offset of the node inside the node region */
    int      order[256];         /* names in file (pre-) order */
    int      n_nodes;
    int64_t  length;
    /* (symbol, depth) -> node name, a cache of decode.getSymbol(prefix | 1 << depth) */
    jshort   path_name[256][16];
} hswt;

static void hswt_cache_paths(hswt* w) {
    for (int s = 0; s < 256; s++) {
        const jint code = w->table[s];
        for (int j = 0, n = w->bit_lengths[s]; j < n; j++) {
            jint idx = code & j_ushr(0x0000FFFF, 16 - j);
            idx |= j_ushr(0x8000, 15 - j);
            w->path_name[s][j] = (jshort)orc_lookup_symbol(&w->decode, idx);
        }
    }
}

/* mapNodes(ByteBuffer in, int[] lengths, int code)  :165-182 */
static int map_nodes_lengths(hswt* w, uint8_t* base, int64_t* pos, int64_t cap, const jint* lengths, jint code) {
    const int idx = orc_lookup_symbol(&w->decode, code);
    if (idx < 0 || idx > 255) return -1;
    const int level = j_nlz32(code) - 1;
    if (31 - level > w->bit_lengths[idx]) return 0;          /* the leaf */
    code |= j_ushr(INT32_MIN, level);
    if (!w->present[idx]) {
        const int64_t nb = orc_ranked_bytes(lengths[idx]);
        if (*pos + nb > cap) return -2;
        w->present[idx] = 1;
        w->node_off[idx] = *pos;
        w->order[w->n_nodes++] = idx;
        orc_bits_init(&w->nodes[idx], base + *pos, nb, lengths[idx], 1);
        *pos += nb;
        int rc = map_nodes_lengths(w, base, pos, cap, lengths, code & j_shr((jint)0xBFFFFFFF, level));
        if (rc) return rc;
        rc = map_nodes_lengths(w, base, pos, cap, lengths, code | j_shr(0x40000000, level));
        if (rc) return rc;
    }
    return 0;
}

/* mapNodes(ByteBuffer in, long length, int code)  :197-216 */
static int map_nodes_counts(hswt* w, uint8_t* base, int64_t* pos, int64_t cap, int64_t length, jint code) {
    const int idx = orc_lookup_symbol(&w->decode, code);
    if (idx < 0 || idx > 255) return -1;
    const int level = j_nlz32(code) - 1;
    if (31 - level > w->bit_lengths[idx]) return 0;
    code |= j_ushr(INT32_MIN, level);
    if (!w->present[idx]) {
        const int64_t nb = orc_ranked_bytes(length);
        if (length <= 0 || *pos + nb > cap) return -2;
        w->present[idx] = 1;
        w->node_off[idx] = *pos;
        w->order[w->n_nodes++] = idx;
        orc_bits_init(&w->nodes[idx], base + *pos, nb, length, 1);
        *pos += nb;
        const int64_t bits = orc_ranked_count_raw(w->nodes[idx].buf, w->nodes[idx].limit, length - 1);
        int rc = map_nodes_counts(w, base, pos, cap, length - bits, code & j_shr((jint)0xBFFFFFFF, level));
        if (rc) return rc;
        rc = map_nodes_counts(w, base, pos, cap, bits, code | j_shr(0x40000000, level));
        if (rc) return rc;
    }
    return 0;
}

/* build ctor :95-125 + fill :127-146, data source = BWTDataSource (fmt/GecozFileWriter.java:289-309) */
int32_t orc_hswt_write(const orc_shape* s, const uint8_t* text, const int32_t* sa, int64_t n,
                       uint8_t* out, int64_t cap) {
    hswt* w = (hswt*)calloc(1, sizeof(hswt));
    memcpy(w->bit_lengths, s->bit_lengths, 256);
    memcpy(w->table, s->table, sizeof(w->table));
    if (orc_lookup_build(&w->decode, w->bit_lengths, 256) != 0) { free(w); return -1; }
    hswt_cache_paths(w);
    int64_t pos = 0;
    int rc = map_nodes_lengths(w, out, &pos, cap, s->node_bits, 1);
    if (rc == 0) {
        for (int64_t i = 0; i < n; i++) {
            const int symbol = sa[i] == 0 ? text[n - 1] : text[sa[i] - 1];       /* BWTDataSource.get :301-303 */
            const jint code = w->table[symbol];
            for (int j = 0, m = w->bit_lengths[symbol]; j < m; j++) {
                const int idx = w->path_name[symbol][j];
                if (idx < 0 || !w->present[idx]) { rc = -3; break; }             /* Java: NullPointerException */
                orc_bits_write(&w->nodes[idx], j_ushr(code, j) & 1, 1);
            }
            if (rc) break;
        }
        for (int i = 0; i < 256 && rc == 0; i++) {
            if (w->present[i]) orc_bits_flush(&w->nodes[i]);
        }
    }
    orc_lookup_free(&w->decode);
    free(w);
    return rc;
}

/* occ :247-267 */
static int64_t hswt_occ(const hswt* w, int symbol, int64_t pos) {
    if (w->bit_lengths[symbol] == 0) return -1;
    const jint code = w->table[symbol];
    for (int i = 0, n = w->bit_lengths[symbol]; i < n && pos >= 0; i++) {
        const int idx = w->path_name[symbol][i];
        const int64_t bits = orc_ranked_count_raw(w->nodes[idx].buf, w->nodes[idx].limit, pos);
        if ((j_ushr(code, i) & 1) == 0) pos -= bits; else pos = bits - 1;
    }
    return pos;
}

/* getRS :300-314 */
static int64_t hswt_get_rs(const hswt* w, int64_t pos) {
    int idx = orc_lookup_symbol(&w->decode, 1);
    for (jint i = 0, code = 0; i < w->bit_lengths[idx]; i++) {
        const int bit = orc_ranked_get_raw(w->nodes[idx].buf, pos);
        const int64_t bits = orc_ranked_count_raw(w->nodes[idx].buf, w->nodes[idx].limit, pos);
        pos = bit == 0 ? pos - bits : bits - 1;
        code |= j_shl(bit, i);
        idx = orc_lookup_symbol(&w->decode, code | j_ushr(0x8000, 14 - i));
    }
    return j_lshl(pos, 32) | idx;
}

/* ============================================================================================
 * IndexWaveletTree
 * ========================================================================================== */
/* build ctor :83-112 (the in-place block counters are kept as written) */
int32_t orc_iwt_write(const int32_t* vals, int64_t m, uint8_t* out, int64_t cap) {
    const jint len = (jint)m;
    jint* sa = (jint*)malloc((size_t)(len > 0 ? len : 1) * sizeof(jint));
    jint* _ssa = (jint*)calloc((size_t)(len > 0 ? len : 1), sizeof(jint));
    memcpy(sa, vals, (size_t)len * sizeof(jint));
    int hibit = 32 - j_nlz32(len);
    const int64_t nb = orc_ranked_bytes(len);
    int64_t pos = 0;
    int rc = 0;
    while (hibit-- > 0) {
        if (pos + nb > cap) { rc = -1; break; }
        orc_bits node;
        orc_bits_init(&node, out + pos, nb, len, 1);
        pos += nb;
        const jint mask = j_shl((jint)0xFFFFFFFF, hibit);
        for (jint i = 0, n = len; i < n; i++) {
            const jint idx = sa[i];
            const jint block = idx & mask;
            const jint lim = j_add(block, j_shl(1, hibit));
            const jint c = (lim < len ? lim : len) - 1;
            jint ptr = _ssa[c];
            if (ptr >= 0) {
                _ssa[c] = ~block;
                _ssa[block] = idx;
            } else {
                _ssa[c] = --ptr;
                _ssa[~ptr] = idx;
            }
            orc_bits_write(&node, (jbyte)(j_shr(idx, hibit) & 1), 1);
        }
        orc_bits_flush(&node);
        jint* tmp = sa; sa = _ssa; _ssa = tmp;
    }
    free(sa); free(_ssa);
    return rc;
}

typedef struct { int levels; int64_t size; const uint8_t* node[64]; int64_t nb; } iwt;

/* read ctor :67-74 */
static void iwt_map(iwt* w, const uint8_t* in, int64_t size) {
    int hibit = 64 - j_nlz64(size);
    w->levels = hibit; w->size = size; w->nb = orc_ranked_bytes(size);
    int64_t pos = 0;
    while (hibit-- > 0) { w->node[hibit] = in + pos; pos += w->nb; }
}

/* get :127-144 */
static int64_t iwt_get(const iwt* w, int64_t pos) {
    jlong code = 0;
    jint block = 0;
    for (int i = 63 - j_nlz64(w->size); i >= 0; i--) {
        const int bit = orc_ranked_get_raw(w->node[i], pos);
        jlong bits = orc_ranked_count_raw(w->node[i], w->nb, pos);
        code = j_lshl(code, 1) | bit;
        if (bit == 0) {
            bits = pos - bits - j_ushr(block, 1);
        } else {
            bits -= j_ushr(block, 1) + 1;
            block += j_shl(1, i);
        }
        pos = block + bits;
    }
    return code;
}

/* find :152-165 */
static int64_t iwt_find(const iwt* w, int64_t idx) {
    jlong pos = 0;
    for (int i = 0, n = 64 - j_nlz64(w->size); i < n; i++) {
        const jlong bit = j_lushr(idx, i) & 1;
        const jlong block = idx & j_lshl((jlong)0xFFFFFFFFFFFFFFFEULL, i);
        jlong hi = block + j_shl(2, i);                    /* (2 << i) is an int expression in Java */
        if (hi > w->size) hi = w->size;
        hi -= 1;
        pos = bit == 0 ? orc_ranked_find_zero_range(w->node[i], w->nb, j_lushr(block, 1) + pos + 1, block, hi)
                       : orc_ranked_find_one_range (w->node[i], w->nb, j_lushr(block, 1) + pos + 1, block, hi);
        pos -= block;
    }
    return pos;
}

int64_t orc_iwt_get(const uint8_t* buf, int64_t m, int64_t pos) { iwt w; iwt_map(&w, buf, m); return iwt_get(&w, pos); }
int64_t orc_iwt_find(const uint8_t* buf, int64_t m, int64_t idx) { iwt w; iwt_map(&w, buf, m); return iwt_find(&w, idx); }

/* ============================================================================================
 * GSSAIndex
 * ========================================================================================== */
/* getIndexSize :200-205 with IndexWaveletTree.size :173-175 */
int64_t orc_index_size(int64_t size, int32_t sampling_factor) {
    const jlong ssa_len = j_lshr(size + j_shl(1, sampling_factor) - 1, sampling_factor);
    const jlong ssa_size = (jlong)(jint)orc_ranked_bytes(ssa_len) * (64LL - j_nlz64(ssa_len));
    const jlong rnk_size = (jint)orc_ranked_bytes(size);
    return ssa_size + rnk_size;
}

/* private GSSAIndex(int[] sa, int sampling_rate, ByteBuffer out)  :129-150 */
int32_t orc_gssa_index_write(const int32_t* sa, int64_t n, int32_t sampling_rate, uint8_t* out, int64_t cap) {
    const int sampling_factor = 31 - j_nlz32(sampling_rate);
    const jint len = (jint)n;
    const jint m = j_shr(len + j_shl(1, sampling_factor) - 1, sampling_factor);
    jint* ssa = (jint*)calloc((size_t)(m > 0 ? m : 1), sizeof(jint));
    const jint mask = j_ushr((jint)0xFFFFFFFF, 32 - sampling_factor);
    const int64_t nb = orc_ranked_bytes(len);
    if (nb > cap) { free(ssa); return -1; }
    orc_bits rank;
    orc_bits_init(&rank, out, nb, len, 1);
    for (jint i = 0, j = 0; i < len; i++) {
        const jint pos = sa[i];
        if ((pos & mask) == 0) {
            ssa[j++] = j_shr(pos, sampling_factor);
            orc_bits_write(&rank, 1, 1);
        } else {
            orc_bits_write(&rank, 0, 1);
        }
    }
    orc_bits_flush(&rank);
    const int rc = orc_iwt_write(ssa, m, out + nb, cap - nb);
    free(ssa);
    return rc;
}

/* ============================================================================================
 * BlockWriter.run  fmt/GecozFileWriter.java:256-284
 * ========================================================================================== */
typedef struct { const orc_shape* shape; const uint8_t* text; const int32_t* sa; int64_t n;
                 uint8_t* out; int64_t cap; int rc; } hswt_job;

static void* hswt_job_run(void* p) {
    hswt_job* j = (hswt_job*)p;
    const int64_t tb = orc_shape_write(j->shape, j->out, j->cap);               /* shape.write(out) :267 */
    j->rc = orc_hswt_write(j->shape, j->text, j->sa, j->n, j->out + tb, j->cap - tb);   /* :268 */
    return NULL;
}

int32_t orc_build_block(const uint8_t* text, int64_t n, int32_t sampling_rate,
                        uint8_t* gcz_body, int64_t gcz_body_len,
                        uint8_t* gcx_body, int64_t gcx_body_len,
                        int32_t* sa_out, uint8_t* bwt_out, int32_t threads) {
    if (n <= 0 || n > INT32_MAX) return -1;
    /* GecozFileWriter.write :127-132 */
    jlong counts[256] = {0};
    for (int64_t i = 0; i < n; i++) counts[text[i]]++;
    orc_shape shape;
    if (orc_shape_from_counts(counts, &shape) != 0) return -2;
    if (shape.size != gcz_body_len) return -3;
    if (orc_index_size(n, 31 - j_nlz32(sampling_rate)) != gcx_body_len) return -4;

    int32_t* sa = sa_out ? sa_out : (int32_t*)malloc((size_t)n * sizeof(int32_t));
    if (!sa) return -5;
    int rc = orc_suffix_array(text, n, sa);                                       /* SAIS.suffix :262 */
    if (rc == 0) {
        memset(gcz_body, 0, (size_t)gcz_body_len);                                 /* fresh mmap slices are zero */
        memset(gcx_body, 0, (size_t)gcx_body_len);
        hswt_job job = { &shape, text, sa, n, gcz_body, gcz_body_len, 0 };
        pthread_t th;
        int threaded = threads > 1 && pthread_create(&th, NULL, hswt_job_run, &job) == 0;
        if (!threaded) hswt_job_run(&job);
        rc = orc_gssa_index_write(sa, n, sampling_rate, gcx_body, gcx_body_len);  /* :274 */
        if (threaded) pthread_join(th, NULL);
        if (rc == 0) rc = job.rc;
        if (bwt_out) {
            for (int64_t i = 0; i < n; i++) bwt_out[i] = sa[i] == 0 ? text[n - 1] : text[sa[i] - 1];
        }
    }
    if (!sa_out) free(sa);
    return rc;
}

/* ============================================================================================
 * GecozFileReader.read + GSSA
 * ========================================================================================== */
struct orc_gssa {
    hswt    tree;
    int64_t table_bytes;
    /* GSSAIndex */
    const uint8_t* rank; int64_t rank_nb;
    iwt     wsa;
    int     sampling_factor;
    /* GSSA */
    jlong   c[256];
    jlong*  e; int ne;
};

/* GSSAIndex.get :171-173 */
static int64_t gssa_index_get(const orc_gssa* g, int64_t pos) {
    return orc_ranked_get_raw(g->rank, pos) == 0
        ? (int64_t)INT32_MIN
        : j_lshl(iwt_get(&g->wsa, orc_ranked_count_raw(g->rank, g->rank_nb, pos) - 1), g->sampling_factor);
}

/* GSSA.locate :241-251 */
static int64_t gssa_locate(const orc_gssa* g, int64_t idx) {
    jlong len = 0;
    jlong sa = gssa_index_get(g, idx);
    while (sa < 0) {
        len++;
        const jlong rs = hswt_get_rs(&g->tree, idx);
        idx = (jint)(g->c[(jint)rs] + j_lushr(rs, 32));
        sa = gssa_index_get(g, idx);
    }
    return sa + len;
}

/* GSSAIndex.find  algo/ssa/GSSAIndex.java:184-187: the suffix-array row that holds text position idx (sampled ones only) */
static int64_t gssa_index_find(const orc_gssa* g, int64_t idx) {
    const jlong sidx = j_lshr(idx, g->sampling_factor);
    return idx == j_lshl(sidx, g->sampling_factor)
        ? orc_ranked_find_one(g->rank, g->tree.length, iwt_find(&g->wsa, sidx) + 1)
        : (int64_t)INT32_MIN;
}

static int cmp_jlong2(const void* a, const void* b) {
    const jlong x = *(const jlong*)a, y = *(const jlong*)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}

/* GSSA.index :215-239 */
static void gssa_index(orc_gssa* g) {
    jlong idx = g->tree.length;
    for (int i = 255; i >= 0; i--) {
        const jlong rank = hswt_occ(&g->tree, i, g->tree.length - 1);
        if (rank >= 0) idx -= rank + 1;
        g->c[i] = idx;
    }
    g->ne = (jint)g->c[1];
    g->e = (jlong*)malloc((size_t)(g->ne > 0 ? g->ne : 1) * sizeof(jlong));
    for (int i = 0; i < g->ne; i++) g->e[i] = gssa_locate(g, i);
    qsort(g->e, (size_t)g->ne, sizeof(jlong), cmp_jlong2);
}

orc_gssa* orc_open(const uint8_t* gcz_body, int64_t body_len, int64_t text_len,
                   const uint8_t* gcx_body, int64_t gcx_len) {
    if (!gcz_body || !gcx_body || text_len <= 0) return NULL;    /* the reference needs the .gcx */
    orc_gssa* g = (orc_gssa*)calloc(1, sizeof(orc_gssa));
    hswt* w = &g->tree;
    /* HSWTShape(ByteBuffer in, long length)  algo/tree/HSWTShape.java:89-109 */
    orc_bits buf;
    orc_bits_init(&buf, (uint8_t*)gcz_body, body_len, body_len, 0);
    if (orc_lengths_table_read(&buf, w->bit_lengths, 256) != 0) { free(g); return NULL; }
    orc_bits_align(&buf);
    g->table_bytes = buf.pos;
    if (orc_encode_table_from_lengths(256, w->bit_lengths, w->table) != 0) { free(g); return NULL; }
    if (orc_lookup_build(&w->decode, w->bit_lengths, 256) != 0) { free(g); return NULL; }
    hswt_cache_paths(w);
    w->length = text_len;
    /* HuffmanShapedWaveletTree(HSWTShape, ByteBuffer in)  :157-163 */
    int64_t pos = 0;
    if (map_nodes_counts(w, (uint8_t*)gcz_body + g->table_bytes, &pos, body_len - g->table_bytes, text_len, 1) != 0) {
        orc_lookup_free(&w->decode); free(g); return NULL;
    }
    /* GSSAIndex(ByteBuffer in, long len)  algo/ssa/GSSAIndex.java:57-71 */
    int sf = -1;
    do {
        sf++;
        if (sf > 30) { orc_lookup_free(&w->decode); free(g); return NULL; }
    } while (gcx_len < orc_index_size(text_len, sf));
    g->sampling_factor = sf;
    g->rank = gcx_body;
    g->rank_nb = orc_ranked_bytes(text_len);
    iwt_map(&g->wsa, gcx_body + g->rank_nb, j_lshr(text_len + j_shl(1, sf) - 1, sf));
    gssa_index(g);
    return g;
}

void orc_close(orc_gssa* g) {
    if (!g) return;
    orc_lookup_free(&g->tree.decode);
    free(g->e);
    free(g);
}

int32_t orc_sampling_factor(const orc_gssa* g) { return g->sampling_factor; }
int32_t orc_num_strings(orc_gssa* g) { return g->ne; }
void orc_string_ends(orc_gssa* g, int64_t* e) { memcpy(e, g->e, (size_t)g->ne * sizeof(int64_t)); }
void orc_c_array(orc_gssa* g, int64_t* c256) { memcpy(c256, g->c, sizeof(g->c)); }
int32_t orc_num_nodes(const orc_gssa* g) { return g->tree.n_nodes; }
void orc_node_info(const orc_gssa* g, int32_t* names, int64_t* lens, int64_t* offsets) {
    for (int i = 0; i < g->tree.n_nodes; i++) {
        const int nm = g->tree.order[i];
        names[i] = nm; lens[i] = g->tree.nodes[nm].size; offsets[i] = g->table_bytes + g->tree.node_off[nm];
    }
}
int64_t orc_occ(const orc_gssa* g, int32_t symbol, int64_t pos) { return hswt_occ(&g->tree, symbol & 255, pos); }
int64_t orc_get_rs(const orc_gssa* g, int64_t pos) { return hswt_get_rs(&g->tree, pos); }
int64_t orc_locate(orc_gssa* g, int64_t row) { return gssa_locate(g, row); }

/* GSSA.search :187-197 (interval part) */
int64_t orc_search(orc_gssa* g, const uint8_t* str, int64_t len, int64_t* sp_out, int64_t* ep_out) {
    const uint64_t calls0 = orc_rank_calls();
    jint ch = (jbyte)(str[len - 1] & 0xFF);
    if (ch < 0) { *sp_out = 0; *ep_out = -1; return 0; }     /* Java: AIOOBE for bytes >= 0x80 */
    jlong sp = g->c[ch];
    jlong ep = ch < 255 ? g->c[ch + 1] - 1 : g->tree.length - 1;
    for (int64_t i = len - 2; sp <= ep && i >= 0; i--) {
        ch = (jbyte)(str[i] & 0xFF);
        if (ch < 0) { sp = 0; ep = -1; break; }
        sp = g->c[ch] + hswt_occ(&g->tree, ch, sp - 1) + 1;
        ep = g->c[ch] + hswt_occ(&g->tree, ch, ep);
    }
    *sp_out = sp; *ep_out = ep;
    return (int64_t)(orc_rank_calls() - calls0);
}

int64_t orc_search_batch(orc_gssa* g, const uint8_t* pats, const int64_t* off, int64_t np,
                         int64_t* sp, int64_t* ep) {
    int64_t calls = 0;
    for (int64_t i = 0; i < np; i++) {
        calls += orc_search(g, pats + off[i], off[i + 1] - off[i], &sp[i], &ep[i]);
    }
    return calls;
}

/* java.util.Arrays.binarySearch(long[] a, int from, int to, long key) */
static jint java_binary_search(const jlong* a, jint from, jint to, jlong key) {
    jint low = from, high = to - 1;
    while (low <= high) {
        const jint mid = (jint)(((uint32_t)low + (uint32_t)high) >> 1);
        const jlong v = a[mid];
        if (v < key) low = mid + 1;
        else if (v > key) high = mid - 1;
        else return mid;
    }
    return -(low + 1);
}

/* GSSA.find :160-185 with the locate loop of search :203-207 */
int64_t orc_find(orc_gssa* g, const uint8_t* pat, int64_t len,
                 int64_t* per_string, int64_t* positions, int64_t cap) {
    for (int i = 0; i < g->ne; i++) per_string[i] = 0;
    int64_t sp, ep;
    orc_search(g, pat, len, &sp, &ep);
    if (ep < sp) return 0;
    const jint k = (jint)(ep - sp + 1);
    jlong* sa = (jlong*)malloc((size_t)k * sizeof(jlong));
    for (jint i = 0; i < k; i++) sa[i] = gssa_locate(g, sp++);
    qsort(sa, (size_t)k, sizeof(jlong), cmp_jlong2);
    int64_t w = 0;
    for (jint i = 0, idx1 = 0; i < g->ne; i++) {
        const jint idx2 = -java_binary_search(sa, idx1, k, g->e[i]) - 1;
        if (idx2 > idx1) {
            const jlong pos = i > 0 ? g->e[i - 1] + 1 : 0;
            per_string[i] = idx2 - idx1;
            for (jint j = 0; j < idx2 - idx1; j++) {
                if (w < cap) positions[w] = sa[idx1 + j] - pos;
                w++;
            }
            idx1 = idx2;
        }
    }
    free(sa);
    return w;
}

/* GSSA.extract(ByteBuffer buf, int nstr, long from)  algo/ssa/GSSA.java:90-126 with buf.position() == 0 and
 * buf.remaining() == cap.  Returns the new buffer position (= bytes written; negative where Java's
 * buf.position(bpos + 1) throws IllegalArgumentException), or INT64_MIN for a bad string index. */
int64_t orc_extract(orc_gssa* g, int32_t nstr, int64_t from, uint8_t* out, int64_t cap) {
    if (nstr < 0 || nstr >= g->ne) return INT64_MIN;
    if (nstr > 0) from += g->e[nstr - 1] + 1;                                 /* :97-100 */
    const jlong lim = from + cap;
    jlong pos = (g->e[nstr] < lim ? g->e[nstr] : lim) - 1;                     /* :104 */
    const jlong sapos = j_lshl(j_lshr(pos, g->sampling_factor) + 1, g->sampling_factor);   /* :107 */
    jlong idx = sapos < g->tree.length ? gssa_index_find(g, sapos) : 0;       /* :110 */
    jlong n = (sapos < g->tree.length - 1 ? sapos : g->tree.length - 1) - pos; /* :113 */
    while (--n > 0) {
        const jlong rs = hswt_get_rs(&g->tree, idx);
        idx = (jint)(g->c[(jint)rs] + j_lushr(rs, 32));
    }
    const jint bpos = (jint)(pos - from);                                     /* :119 */
    for (jint i = bpos; i >= 0; i--) {
        const jlong rs = hswt_get_rs(&g->tree, idx);
        out[i] = (uint8_t)(rs & 0xFF);
        idx = (jint)(g->c[(jint)rs] + j_lushr(rs, 32));
    }
    return (int64_t)bpos + 1;
}

int64_t orc_index_find(orc_gssa* g, int64_t pos) { return gssa_index_find(g, pos); }

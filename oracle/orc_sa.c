/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see orc_internal.h).
 *
 * Suffix array of a byte string, RESULT-ONLY restatement of algo/string/SAIS.java:103-137
 * (`SAIS.suffix(ByteBuffer, int[])`): bytes compare unsigned, a suffix that is a prefix of
 * another one is the smaller (the reference seeds the last suffix as L-type in
 * sortLMS/induceLMS, SAIS.java:547-550,593-596, i.e. a virtual end marker below every symbol).
 * The suffix array of a string is unique, so the algorithm is free: this file is a fresh
 * textbook SA-IS (Nong, Zhang, Chan 2009: classify S/L, induce-sort LMS substrings, name,
 * recurse, induce) and a naive comparison sorter used to pin it at small sizes.  It does
 * not follow the reference's SACA-K variant.
 */
#include "orc_internal.h"
#include "gcz_oracle.h"

#define CHR(T, cs, i) ((cs) == 1 ? (int32_t)((const uint8_t*)(T))[(i)] : ((const int32_t*)(T))[(i)])
#define T_GET(t, i)   (((t)[(i) >> 3] >> ((i) & 7)) & 1)          /* 1 = S-type */
#define T_SET(t, i)   ((t)[(i) >> 3] |= (uint8_t)(1u << ((i) & 7)))
#define IS_LMS(t, i)  ((i) > 0 && T_GET(t, i) && !T_GET(t, (i) - 1))

static void bucket_bounds(const int32_t* C, int32_t* B, int32_t k, int ends) {
    int32_t sum = 0;
    for (int32_t c = 0; c < k; c++) {
        sum += C[c];
        B[c] = ends ? sum : sum - C[c];
    }
}

static void induce(const void* T, int cs, int32_t* SA, int32_t n, int32_t k,
                   const uint8_t* t, const int32_t* C, int32_t* B) {
    /* L-types, left to right; the virtual end marker sits "before" SA[0] and induces n-1 */
    bucket_bounds(C, B, k, 0);
    SA[B[CHR(T, cs, n - 1)]++] = n - 1;
    for (int32_t i = 0; i < n; i++) {
        const int32_t j = SA[i] - 1;
        if (SA[i] > 0 && !T_GET(t, j)) SA[B[CHR(T, cs, j)]++] = j;
    }
    /* S-types, right to left */
    bucket_bounds(C, B, k, 1);
    for (int32_t i = n - 1; i >= 0; i--) {
        const int32_t j = SA[i] - 1;
        if (SA[i] > 0 && T_GET(t, j)) SA[--B[CHR(T, cs, j)]] = j;
    }
}

static int lms_substr_equal(const void* T, int cs, const uint8_t* t, int32_t n, int32_t p, int32_t q) {
    for (int32_t d = 0;; d++) {
        if (p + d >= n || q + d >= n) return 0;        /* one ran into the end marker first */
        if (CHR(T, cs, p + d) != CHR(T, cs, q + d) || T_GET(t, p + d) != T_GET(t, q + d)) return 0;
        if (d > 0) {
            const int lp = IS_LMS(t, p + d), lq = IS_LMS(t, q + d);
            if (lp || lq) return lp && lq;
        }
    }
}

static int sais(const void* T, int32_t* SA, int32_t n, int32_t k, int cs) {
    if (n == 0) return 0;
    if (n == 1) { SA[0] = 0; return 0; }

    uint8_t* t = (uint8_t*)calloc((size_t)(n >> 3) + 1, 1);
    int32_t* C = (int32_t*)calloc((size_t)k, sizeof(int32_t));
    int32_t* B = (int32_t*)malloc((size_t)k * sizeof(int32_t));
    if (!t || !C || !B) { free(t); free(C); free(B); return -1; }

    for (int32_t i = 0; i < n; i++) C[CHR(T, cs, i)]++;
    /* S/L classification; position n-1 is L (end marker is smaller than everything) */
    for (int32_t i = n - 2; i >= 0; i--) {
        const int32_t a = CHR(T, cs, i), b = CHR(T, cs, i + 1);
        if (a < b || (a == b && T_GET(t, i + 1))) T_SET(t, i);
    }

    /* stage 1: sort LMS substrings */
    for (int32_t i = 0; i < n; i++) SA[i] = -1;
    bucket_bounds(C, B, k, 1);
    int32_t m = 0;
    for (int32_t i = 1; i < n; i++) {
        if (IS_LMS(t, i)) { SA[--B[CHR(T, cs, i)]] = i; m++; }
    }
    int rc = 0;
    if (m > 0) {
        induce(T, cs, SA, n, k, t, C, B);

        /* compact sorted LMS substrings, name them */
        int32_t w = 0;
        for (int32_t i = 0; i < n; i++) {
            if (IS_LMS(t, SA[i])) SA[w++] = SA[i];
        }
        for (int32_t i = m; i < n; i++) SA[i] = -1;
        int32_t names = 0, prev = -1;
        for (int32_t i = 0; i < m; i++) {
            const int32_t p = SA[i];
            if (prev < 0 || !lms_substr_equal(T, cs, t, n, prev, p)) names++;
            prev = p;
            SA[m + (p >> 1)] = names - 1;
        }
        int32_t* s1 = SA + n - m;
        for (int32_t i = n - 1, j = n - 1; i >= m; i--) {
            if (SA[i] >= 0) SA[j--] = SA[i];
        }
        int32_t* SA1 = SA;
        if (names < m) {
            rc = sais(s1, SA1, m, names, 4);
        } else {
            for (int32_t i = 0; i < m; i++) SA1[s1[i]] = i;
        }
        if (rc == 0) {
            /* map reduced suffixes back to text positions */
            for (int32_t i = 1, j = 0; i < n; i++) {
                if (IS_LMS(t, i)) s1[j++] = i;
            }
            for (int32_t i = 0; i < m; i++) SA1[i] = s1[SA1[i]];
            /* stage 3: seed sorted LMS suffixes at their bucket ends */
            for (int32_t i = m; i < n; i++) SA[i] = -1;
            bucket_bounds(C, B, k, 1);
            for (int32_t i = m - 1; i >= 0; i--) {
                const int32_t j = SA[i];
                SA[i] = -1;
                SA[--B[CHR(T, cs, j)]] = j;
            }
        }
    }
    if (rc == 0) induce(T, cs, SA, n, k, t, C, B);
    free(t); free(C); free(B);
    return rc;
}

int32_t orc_suffix_array(const uint8_t* text, int64_t n, int32_t* sa) {
    if (n < 0 || n > INT32_MAX) return -1;
    return sais(text, sa, (int32_t)n, 256, 1);
}

/* pinning reference for small inputs: plain comparison sort */
static const uint8_t* g_text; static int64_t g_n;
static int cmp_suffix(const void* a, const void* b) {
    const int32_t p = *(const int32_t*)a, q = *(const int32_t*)b;
    const int64_t lp = g_n - p, lq = g_n - q, l = lp < lq ? lp : lq;
    const int c = memcmp(g_text + p, g_text + q, (size_t)l);
    if (c) return c;
    return lp < lq ? -1 : (lp > lq ? 1 : 0);
}
int32_t orc_suffix_array_naive(const uint8_t* text, int64_t n, int32_t* sa) {
    if (n < 0 || n > INT32_MAX) return -1;
    for (int32_t i = 0; i < n; i++) sa[i] = i;
    g_text = text; g_n = n;
    qsort(sa, (size_t)n, sizeof(int32_t), cmp_suffix);
    return 0;
}

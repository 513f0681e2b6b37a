/*
 * gotoh2_oracle.c - CPU restatement of MiCall-Lite's live aligner `_gotoh2.align`
 * ("next" row #1: the aligner bin/micall really calls).
 *
 * TEST INFRASTRUCTURE ONLY - never linked into the product library.
 *
 * Parity status: PINNED.  oracle/_ref/_gotoh2*.so is the reference's own
 * /root/reference/micall/alignment/src/_gotoh2.c compiled unmodified (oracle/Makefile);
 * tests/test_oracle.py checks this restatement against it on the reference's own unit-test
 * vectors (micall/alignment/tests/test.py:174-298, captured into tests/golden/gotoh2.json)
 * and on a seeded fuzz.
 *
 * Reference lines restated (all in micall/alignment/src/_gotoh2.c):
 *   initialize :93-134, cost_assignment :137-201, edge_assignment :205-312 (Altschul-Erickson
 *   steps 8-11), traceback :316-437, align :442-541.
 * Differences in form, not in result: each cell records how ITS OWN p and q were formed
 * (the reference stores those bits on the neighbouring cell, :157-176); the reverse pass only
 * keeps the final a/b/c bits because nothing ever reads the d/e/f/g bits it rewrites
 * (:258-308 write them after their last read); +infinity is a 64-bit constant instead of
 * INT_MAX arithmetic that overflows (:96-109,157-167).
 */
#include "gotoh2_oracle.h"

#include <stdlib.h>
#include <string.h>

typedef long long i64;
#define INF ((i64)1 << 60)

enum { BA = 1, BB = 2, BC = 4, BD = 8, BE = 16, BF = 32, BG = 64 };

static i64 min2(i64 a, i64 b) { return a <= b ? a : b; }

int gotoh2_oracle_align(const char* seq1, long l1, const char* seq2, long l2, int gop, int gep,
                        int is_global, const char* alphabet, const int* d, char* out1, char* out2,
                        int* out_len, int* out_score) {
    const i64 v = gop, u = gep;                                  /* :563-564 */
    const int l = (int)strlen(alphabet);
    const long nrows = l1 + 1, ncols = l2 + 1;
    int map[256];
    long i, j, k, alen = 0, init_i, init_j;
    i64 *R, *p, *q, best;
    unsigned char* bits;   /* own-cell convention: a,b,c + D,E (how p[i][j] was formed) + F,G (how q[i][j] was formed) */
    unsigned char* fin;    /* final a,b,c after the reverse pass; (nrows+1) x (ncols+1) incl. the sentinel border */
    int *s1, *s2;
    char *r1, *r2;

    if (l1 <= 0 || l2 <= 0) return GOTOH2_ORACLE_EEMPTY;         /* gotoh2.py:84-85 asserts non-empty */
    for (i = 0; i < 256; ++i) map[i] = -1;                        /* map_ascii_to_alphabet :68-77 */
    for (i = 0; i < l; ++i) map[(unsigned char)alphabet[i]] = (int)i;
    s1 = (int*)malloc(sizeof(int) * (size_t)(l1 + l2));
    R = (i64*)malloc(sizeof(i64) * (size_t)nrows * (size_t)ncols * 3);
    bits = (unsigned char*)calloc((size_t)nrows * (size_t)ncols, 1);
    fin = (unsigned char*)calloc((size_t)(nrows + 1) * (size_t)(ncols + 1), 1);
    r1 = (char*)malloc((size_t)(l1 + l2) * 2 + 2);
    if (!s1 || !R || !bits || !fin || !r1) { free(s1); free(R); free(bits); free(fin); free(r1); return GOTOH2_ORACLE_ENOMEM; }
    s2 = s1 + l1;
    p = R + (size_t)nrows * (size_t)ncols;
    q = p + (size_t)nrows * (size_t)ncols;
    r2 = r1 + (l1 + l2 + 1);
    for (i = 0; i < l1; ++i) { s1[i] = map[(unsigned char)seq1[i]]; if (s1[i] < 0) goto domain; }
    for (j = 0; j < l2; ++j) { s2[j] = map[(unsigned char)seq2[j]]; if (s2[j] < 0) goto domain; }

    /* forward: initialize + cost_assignment (:93-201) */
    for (i = 0; i < nrows; ++i)
        for (j = 0; j < ncols; ++j) {
            const size_t here = (size_t)i * (size_t)ncols + (size_t)j;
            unsigned char b = 0;
            i64 pv = INF, qv = INF, rv, dg = INF;
            if (i > 0) {
                const size_t up = here - (size_t)ncols;
                pv = u + min2(p[up], R[up] + v);                 /* :156 */
                if (p[up] < INF && pv == p[up] + u) b |= BD;     /* :157-159 */
                if (pv == R[up] + v + u) b |= BE;                /* :160-162 */
            }
            if (j > 0) {
                qv = u + min2(q[here - 1], R[here - 1] + v);     /* :166 */
                if (q[here - 1] < INF && qv == q[here - 1] + u) b |= BF;
                if (qv == R[here - 1] + v + u) b |= BG;
            }
            if (i == 0 || j == 0) {                              /* :175-183 */
                if (i == 0 && j == 0) rv = 0;
                else rv = is_global ? min2(pv, qv) : 0;
            } else {
                dg = R[here - (size_t)ncols - 1] - d[s1[i - 1] * l + s2[j - 1]];
                rv = min2(min2(dg, pv), qv);                     /* :185-187 */
            }
            if (rv == pv) b |= BA;                               /* :190-198 */
            if (rv == qv) b |= BB;
            if (i > 0 && j > 0 && rv == dg) b |= BC;
            p[here] = pv; q[here] = qv; R[here] = rv; bits[here] = b;
        }

    /* reverse: edge_assignment steps 8-11 (:205-312) on the (nrows+1) x (ncols+1) bit grid whose extra
     * row/column is c=1 everywhere (local) or only in the corner (global) (:118-133) */
    for (i = 0; i <= nrows; ++i)
        for (j = 0; j <= ncols; ++j) {
            const size_t x = (size_t)i * (size_t)(ncols + 1) + (size_t)j;
            if (i == nrows || j == ncols) fin[x] = (unsigned char)((!is_global || (i == nrows && j == ncols)) ? BC : 0);
        }
    for (i = nrows - 1; i >= 0; --i)
        for (j = ncols - 1; j >= 0; --j) {
            const size_t x = (size_t)i * (size_t)(ncols + 1) + (size_t)j;
            const size_t here = (size_t)i * (size_t)ncols + (size_t)j;
            const int A1 = fin[x + (size_t)(ncols + 1)] & BA;            /* a[i+1,j] */
            const int B1 = fin[x + 1] & BB;                              /* b[i,j+1] */
            const int C1 = fin[x + (size_t)(ncols + 1) + 1] & BC;        /* c[i+1,j+1] */
            const int below = (i + 1 < nrows) ? bits[here + (size_t)ncols] : 0;
            const int right = (j + 1 < ncols) ? bits[here + 1] : 0;
            const int d0 = below & BD, e0 = below & BE, f0 = right & BF, g0 = right & BG;
            unsigned char abc = bits[here] & (BA | BB | BC);
            if ((!A1 || !e0) && (!B1 || !g0) && !C1) abc = 0;            /* step 8 :237-242 */
            if (A1 || B1 || C1) {                                        /* step 9 :245 */
                if (A1 && d0) abc |= BA;                                 /* step 10 :251-270 */
                if (B1 && f0) abc |= BB;                                 /* step 11 :283-300 */
            }
            fin[x] = abc;
        }

    /* traceback (:316-437) */
    init_i = nrows - 1; init_j = ncols - 1;
    best = R[(size_t)init_i * (size_t)ncols + (size_t)init_j];
    if (!is_global) {
        for (i = 0; i < nrows; ++i)                                      /* right-most column :330-339 */
            if (R[(size_t)i * (size_t)ncols + (size_t)(ncols - 1)] < best) { best = R[(size_t)i * (size_t)ncols + (size_t)(ncols - 1)]; init_i = i; init_j = ncols - 1; }
        for (j = 0; j < ncols; ++j)                                      /* bottom row :341-350 */
            if (R[(size_t)(nrows - 1) * (size_t)ncols + (size_t)j] < best) { best = R[(size_t)(nrows - 1) * (size_t)ncols + (size_t)j]; init_i = nrows - 1; init_j = j; }
    }
    i = init_i; j = init_j;
    for (k = nrows - 1; k > i; --k) { r1[alen] = seq1[k - 1]; r2[alen] = '-'; ++alen; }   /* :361-372 */
    for (k = ncols - 1; k > j; --k) { r1[alen] = '-'; r2[alen] = seq2[k - 1]; ++alen; }
    while (i > 0 && j > 0) {                                             /* :374-405 */
        const int b = fin[(size_t)i * (size_t)(ncols + 1) + (size_t)j];
        if (b & BA) { r1[alen] = seq1[i - 1]; r2[alen] = '-'; --i; }
        else if (b & BB) { r1[alen] = '-'; r2[alen] = seq2[j - 1]; --j; }
        else if (b & BC) { r1[alen] = seq1[i - 1]; r2[alen] = seq2[j - 1]; --i; --j; }
        else { free(s1); free(R); free(bits); free(fin); free(r1); return GOTOH2_ORACLE_ETRACEBACK; }  /* :403-407 */
        ++alen;
    }
    while (i > 0) { r1[alen] = seq1[i - 1]; r2[alen] = '-'; --i; ++alen; }                /* :411-422 */
    while (j > 0) { r1[alen] = '-'; r2[alen] = seq2[j - 1]; --j; ++alen; }
    for (k = 0; k < alen; ++k) { out1[k] = r1[alen - 1 - k]; out2[k] = r2[alen - 1 - k]; }
    *out_len = (int)alen;
    *out_score = (int)(-best);                                           /* :437 */
    free(s1); free(R); free(bits); free(fin); free(r1);
    return 0;
domain:
    free(s1); free(R); free(bits); free(fin); free(r1);
    return GOTOH2_ORACLE_EDOMAIN;
}

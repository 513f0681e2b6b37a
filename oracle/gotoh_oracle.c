/*
 * gotoh_oracle.c - CPU restatement of MiCall-Lite's Gotoh aligner hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA path
 * (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg).  It is never
 * linked into, imported by, or called from the product library
 * (micall-lite_b200/csrc -> libgotoh_b200.so), which has no CPU fallback.
 *
 * Parity status: PINNED against the reference itself - oracle/_ref/libgotoh_ref.so
 * is the reference's untouched gotoh.cpp compiled here (oracle/Makefile), and
 * tests/test_oracle.py checks this restatement against it on the SURVEY
 * Appendix-B vectors, the tests/golden JSON vectors and a seeded fuzz.  The reference's own
 * test-suite holds no vectors for align_it/align_it_aa.
 *
 * Every function cites the reference lines it restates
 * (/root/reference/micall/alignment/gotoh.cpp).  It is written from the
 * semantics, not transliterated: one byte of direction
 * per cell instead of two int pointer matrices, rolling rows, explicit overhangs.
 */
#include "gotoh_oracle.h"

#include <stdlib.h>
#include <string.h>

#define TBL 128
#define SENTINEL (-100000) /* gotoh.cpp:284-285 */

/* ---- score tables (gotoh.cpp:25-213) ------------------------------------ */

static void set2(int* t, int a, int b, int v) { t[a * TBL + b] = v; t[b * TBL + a] = v; }

/* init_pairscore(match, mismatch): gotoh.cpp:26-131.  Assignment ORDER matters. */
static void table_nt(int* t, int match, int mismatch) {
    int i, j;
    const char* p;
    for (i = 0; i < TBL; ++i)
        for (j = 0; j < TBL; ++j) t[i * TBL + j] = (i == j) ? match : -mismatch; /* :28-45 */
    /* case-insensitive ACGT/U (:48-53).  Note the reference sets ('T','u') but
     * not ('u','T'), and ('U','t'),('U','T'),('u','t'),('t','T') but not
     * ('t','U')'s mirror beyond what is listed - restated entry by entry. */
    set2(t, 'a', 'A', match); set2(t, 'c', 'C', match); set2(t, 'g', 'G', match);
    set2(t, 't', 'T', match); set2(t, 'u', 'U', match);              /* :51 */
    t['t' * TBL + 'u'] = match; t['t' * TBL + 'U'] = match;          /* :52 */
    t['T' * TBL + 'u'] = match; t['T' * TBL + 'U'] = match;
    t['u' * TBL + 't'] = match; t['t' * TBL + 'T'] = match;          /* :53 */
    t['U' * TBL + 't'] = match; t['U' * TBL + 'T'] = match;
    t['N' * TBL + 'N'] = 0; t['n' * TBL + 'N'] = 0; t['N' * TBL + 'n'] = 0; /* :54 */
    /* IUPAC bi- and tri-mixtures, upper case only (:58-91) */
    {
        static const char* mix[] = {"RAG", "YCT", "KGT", "MCA", "SCG", "WTA",
                                    "BCGT", "DAGT", "HACT", "VACG", 0};
        int m;
        for (m = 0; mix[m]; ++m)
            for (p = mix[m] + 1; *p; ++p) set2(t, mix[m][0], *p, match);
    }
    /* '*' wildcard (:94-97), '$' (:99), '.' (:105-108), 'N' (:110-113) */
    for (p = "ACTG"; *p; ++p) {
        set2(t, '*', *p, match); set2(t, '*', *p + 32, match);
    }
    t['$' * TBL + '$'] = 50;
    for (p = "ACTG"; *p; ++p) { set2(t, '.', *p, -20); set2(t, '.', *p + 32, -20); }
    for (p = "ACTG"; *p; ++p) { set2(t, 'N', *p, -3); set2(t, 'N', *p + 32, -3); }
    /* 'X' against bases and IUPAC codes, both cases (:116-129) */
    for (p = "ACTGRYKMSWBDHV"; *p; ++p) { set2(t, 'X', *p, -6); set2(t, 'X', *p + 32, -6); }
    t['X' * TBL + '-'] = 3; /* :130 - one-sided in the reference */
}

/* init_pairscore_aa(match, mismatchPenalty): gotoh.cpp:134-157 */
static void table_aa_rb(int* t, int match, int mismatch) {
    int i, j;
    for (i = 0; i < TBL; ++i)
        for (j = 0; j < TBL; ++j) {
            int v = (i == j) ? match : -mismatch;
            if (i != j && (i == 'X' || j == 'X')) v = -4;
            t[i * TBL + j] = v;
        }
    t['Z' * TBL + 'Z'] = 0; t['z' * TBL + 'Z'] = 0; t['Z' * TBL + 'z'] = 0;
    t['X' * TBL + '-'] = match; t['-' * TBL + 'X'] = match;
}

/* empirical_hiv25: gotoh.cpp:165-189 (Nickle et al. 2007, PLoS One 2(6):e503).
 * Pure data; alphabet order ARNDCQEGHILKMFPSTWYVBZ?* (gotoh.cpp:192-194). */
static const signed char HIV25[24][24] = {
    {7,-7,-7,-4,-10,-11,-4,-3,-10,-6,-9,-9,-7,-13,-3,-2,1,-16,-15,0,-5,-5,-3,-17},
    {-7,7,-5,-11,-8,-2,-7,-2,0,-6,-6,2,-3,-12,-4,-2,-2,-5,-9,-10,-7,-3,-3,-17},
    {-7,-5,8,2,-9,-6,-6,-7,0,-6,-12,0,-10,-12,-9,1,0,-17,-3,-10,6,-6,-3,-17},
    {-4,-11,2,8,-14,-10,0,-2,-3,-11,-15,-7,-13,-15,-13,-5,-6,-16,-6,-5,7,0,-3,-17},
    {-10,-8,-9,-14,11,-16,-15,-5,-7,-11,-9,-13,-14,0,-12,-1,-6,-2,0,-8,-10,-16,-5,-17},
    {-11,-2,-6,-10,-16,8,-2,-10,0,-12,-4,0,-8,-12,-1,-9,-8,-14,-9,-13,-7,6,-4,-17},
    {-4,-7,-6,0,-15,-2,7,-1,-9,-12,-15,-1,-10,-17,-13,-11,-8,-15,-12,-5,0,6,-4,-17},
    {-3,-2,-7,-2,-5,-10,-1,7,-10,-11,-14,-6,-12,-9,-11,-1,-7,-5,-14,-5,-4,-3,-4,-17},
    {-10,0,0,-3,-7,0,-9,-10,10,-10,-4,-5,-10,-6,-3,-6,-6,-11,2,-14,-1,-2,-3,-17},
    {-6,-6,-6,-11,-11,-12,-12,-11,-10,7,0,-7,0,-2,-10,-4,0,-14,-9,2,-7,-12,-2,-17},
    {-9,-6,-12,-15,-9,-4,-15,-14,-4,0,6,-10,0,0,-3,-5,-8,-6,-8,-4,-13,-6,-4,-17},
    {-9,2,0,-7,-13,0,-1,-6,-5,-7,-10,7,-4,-14,-9,-5,-1,-12,-13,-9,-1,-1,-2,-17},
    {-7,-3,-10,-13,-14,-8,-10,-12,-10,0,0,-4,10,-7,-11,-9,-1,-11,-15,0,-11,-9,-3,-17},
    {-13,-12,-12,-15,0,-12,-17,-9,-6,-2,0,-14,-7,10,-11,-5,-10,-5,1,-5,-13,-14,-3,-17},
    {-3,-4,-9,-13,-12,-1,-13,-11,-3,-10,-3,-9,-11,-11,8,-1,-3,-13,-11,-12,-10,-3,-5,-17},
    {-2,-2,1,-5,-1,-9,-11,-1,-6,-4,-5,-5,-9,-5,-1,8,0,-12,-6,-9,0,-10,-3,-17},
    {1,-2,0,-6,-6,-8,-8,-7,-6,0,-8,-1,-1,-10,-3,0,7,-16,-10,-4,-2,-8,-2,-17},
    {-16,-5,-17,-16,-2,-14,-15,-5,-11,-14,-6,-12,-11,-5,-13,-12,-16,10,-4,-16,-16,-14,-8,-17},
    {-15,-9,-3,-6,0,-9,-12,-14,2,-9,-8,-13,-15,1,-11,-6,-10,-4,10,-12,-4,-10,-4,-17},
    {0,-10,-10,-5,-8,-13,-5,-5,-14,2,-4,-9,0,-5,-12,-9,-4,-16,-12,7,-7,-7,-3,-17},
    {-5,-7,6,7,-10,-7,0,-4,-1,-7,-13,-1,-11,-13,-10,0,-2,-16,-4,-7,7,-2,-4,-17},
    {-5,-3,-6,0,-16,6,6,-3,-2,-12,-6,-1,-9,-14,-3,-10,-8,-14,-10,-7,-2,6,-4,-17},
    {-3,-3,-3,-3,-5,-4,-4,-4,-3,-2,-4,-2,-3,-3,-5,-3,-2,-8,-4,-3,-4,-4,-3,-17},
    {-17,-17,-17,-17,-17,-17,-17,-17,-17,-17,-17,-17,-17,-17,-17,-17,-17,-17,-17,-17,-17,-17,-17,1}};

/* init_pairscore_hiv25(): gotoh.cpp:191-213.  The "+32" lower-case aliasing is
 * applied to every alphabet entry, including '?' (63 -> 95 '_') and '*' (42 -> 74 'J'). */
static void table_hiv25(int* t) {
    static const char alpha[] = "ARNDCQEGHILKMFPSTWYVBZ?*";
    int i, j;
    memset(t, 0, sizeof(int) * TBL * TBL);
    for (i = 0; i < 24; ++i)
        for (j = 0; j < 24; ++j) {
            int a = alpha[i], b = alpha[j], v = HIV25[i][j];
            t[a * TBL + b] = v; t[(a + 32) * TBL + b] = v;
            t[a * TBL + (b + 32)] = v; t[(a + 32) * TBL + (b + 32)] = v;
        }
}

void gotoh_oracle_table(int matrix_id, int* out127) {
    int t[TBL * TBL];
    int a, b;
    if (matrix_id == GOTOH_ORACLE_NT) table_nt(t, 5, 4);          /* gotoh.cpp:637 */
    else if (matrix_id == GOTOH_ORACLE_HIV25) table_hiv25(t);     /* gotoh.cpp:673 */
    else table_aa_rb(t, 4, -2);                                   /* gotoh.cpp:707 */
    for (a = 0; a < 127; ++a)
        for (b = 0; b < 127; ++b) out127[a * 127 + b] = t[a * TBL + b];
}

/* ---- string helpers ------------------------------------------------------ */

/* trim(): gotoh.cpp:545-559 - strips leading/trailing " \t\n\r". */
static void trim_span(const char* s, long n, long* lo, long* hi) {
    long a = 0, b = n;
    while (a < b && (s[a] == ' ' || s[a] == '\t' || s[a] == '\n' || s[a] == '\r')) ++a;
    while (b > a && (s[b - 1] == ' ' || s[b - 1] == '\t' || s[b - 1] == '\n' || s[b - 1] == '\r')) --b;
    *lo = a; *hi = b;
}

/* degap(): gotoh.cpp:529-543 - removes every '-'. Returns new length. */
static long degap_copy(const char* s, long n, char* dst) {
    long i, m = 0;
    for (i = 0; i < n; ++i)
        if (s[i] != '-') dst[m++] = s[i];
    return m;
}

/* ---- the aligner (gotoh.cpp:233-527) -------------------------------------- */

enum { DIR_DIAG = 0, DIR_UP = 1, DIR_LEFT = 2 };

/* Reads past the end see the std::string terminator (NUL) and short-circuit in
 * the reference (:324-344), so out-of-range positions simply compare unequal. */
static char at(const char* s, long n, long p) { return (p >= 0 && p < n) ? s[p] : 0; }

static int is_stop3(const char* b, long N, long p) { /* b[p..p+2] in {TAG,TAA,TGA} */
    char x = at(b, N, p), y = at(b, N, p + 1), z = at(b, N, p + 2);
    return x == 'T' && ((y == 'A' && (z == 'G' || z == 'A')) || (y == 'G' && z == 'A'));
}

static int is_dollar3(const char* a, long M, long p) { /* a[p..p+2] == "$$$" */
    return at(a, M, p) == '$' && at(a, M, p + 1) == '$' && at(a, M, p + 2) == '$';
}

/* Core align() on already trimmed/degapped byte strings.  Returns 0 or a negative
 * GOTOH_ORACLE_E* code; outputs have length *out_len <= M+N (no terminator added). */
static int align_core(const int* T, const char* a, long M, const char* b, long N,
                      int gip, int gep, int term, char* out_a, char* out_b,
                      int* out_len, int* out_score) {
    const int u = -gip, v = -gep;                    /* :259-260 */
    int *S, *P;
    unsigned char* dir;
    long i, j, L = 0, k;
    int maxiS = SENTINEL, maxjS = SENTINEL;          /* :284-285 */
    long maxij = -1, maxji = -1;
    int score, has_dollar = 0;
    char *ra, *rb;

    if (M <= 0 || N <= 0) return GOTOH_ORACLE_EEMPTY; /* reference is UB here (Appendix A.7) */
    for (i = 0; i < M; ++i) {
        if ((unsigned char)a[i] < 1 || (unsigned char)a[i] > 126) return GOTOH_ORACLE_EDOMAIN;
        if (a[i] == '$') has_dollar = 1;
    }
    for (j = 0; j < N; ++j)
        if ((unsigned char)b[j] < 1 || (unsigned char)b[j] > 126) return GOTOH_ORACLE_EDOMAIN;

    S = (int*)calloc((size_t)(N + 1) * 2, sizeof(int));   /* S_0[j] = 0, P_0[j] = 0 (:267-272) */
    dir = (unsigned char*)malloc((size_t)M * (size_t)N);
    ra = (char*)malloc((size_t)(M + N) * 2 + 2);
    if (!S || !dir || !ra) { free(S); free(dir); free(ra); return GOTOH_ORACLE_ENOMEM; }
    P = S + (N + 1);
    rb = ra + (M + N + 1);

    for (i = 1; i <= M; ++i) {
        const int t = u + (int)i * v;                /* :290 (t starts at u, += v per row) */
        int s = t, q = t + u;                        /* :291,293 */
        int diag = S[0];                             /* oldSS[0]; SS[0] is kept at 0 (:292) */
        const int* Ta = T + (int)(unsigned char)a[i - 1] * TBL;
        unsigned char* drow = dir + (size_t)(i - 1) * (size_t)N;
        for (j = 1; j <= N; ++j) {
            int up = S[j], p, d, best, dd;
            q = (q >= s + u ? q : s + u) + v;                       /* :305-308 */
            p = (up + u > P[j] ? up + u : P[j]) + v;                /* :311-314 */
            P[j] = p;
            d = diag + Ta[(unsigned char)b[j - 1]];                 /* :319 */
            if (has_dollar && i >= 3 && j >= 3) {                   /* :324-344 */
                if (is_dollar3(a, M, i - 3) && is_stop3(b, N, j - 3)) d += 6;
                if (is_dollar3(a, M, i - 2) && is_stop3(b, N, j - 2)) d += 6;
                if (is_dollar3(a, M, i - 1) && is_stop3(b, N, j - 1)) d += 6;
            }
            /* :362-395  LEFT if Q >= max(P,D); else UP if P >= D; else DIAG */
            if (p >= d) { if (p > q) { best = p; dd = DIR_UP; } else { best = q; dd = DIR_LEFT; } }
            else        { if (d > q) { best = d; dd = DIR_DIAG; } else { best = q; dd = DIR_LEFT; } }
            diag = up;
            S[j] = s = best;
            drow[j - 1] = (unsigned char)dd;
            if (i == M && s >= maxiS) { maxiS = s; maxij = j; }     /* :399-403 */
        }
        if (S[N] >= maxjS) { maxjS = S[N]; maxji = i; }             /* :406-410 */
    }
    if (maxij < 0 || maxji < 0) { /* every boundary score below the sentinel: reference reads
                                   * uninitialised indices (Appendix A.7) */
        free(S); free(dir); free(ra);
        return GOTOH_ORACLE_ESENTINEL;
    }

    /* end cell + right overhang (:429-450) */
    if (maxiS > maxjS) {
        score = maxiS; i = M; j = maxij;
        for (k = N; k > maxij; --k) { ra[L] = '-'; rb[L] = b[k - 1]; ++L; }
    } else {
        score = maxjS; i = maxji; j = N;
        for (k = M; k > maxji; --k) { ra[L] = a[k - 1]; rb[L] = '-'; ++L; }
    }
    /* traceback (:455-487) */
    while (i >= 1 && j >= 1) {
        int dd = dir[(size_t)(i - 1) * (size_t)N + (size_t)(j - 1)];
        if (dd == DIR_DIAG) { ra[L] = a[i - 1]; rb[L] = b[j - 1]; --i; --j; }
        else if (dd == DIR_UP) { ra[L] = a[i - 1]; rb[L] = '-'; --i; }
        else { ra[L] = '-'; rb[L] = b[j - 1]; --j; }
        ++L;
    }
    /* left overhang + terminal-gap add-back (:491-510) */
    if (i < j) {
        for (k = j; k >= 1; --k) { ra[L] = '-'; rb[L] = b[k - 1]; ++L; if (term == 0) score += gep; }
        if (term == 0) score += gip;
    } else if (i > j) {
        for (k = i; k >= 1; --k) { ra[L] = a[k - 1]; rb[L] = '-'; ++L; if (term == 0) score += gep; }
        if (term == 0) score += gip;
    }
    /* reverse (:512-513) */
    for (k = 0; k < L; ++k) { out_a[k] = ra[L - 1 - k]; out_b[k] = rb[L - 1 - k]; }
    *out_len = (int)L;
    *out_score = score;
    free(S); free(dir); free(ra);
    return 0;
}

int gotoh_oracle_align(int matrix_id, const char* a, long a_len, const char* b, long b_len,
                       int gip, int gep, int term, char* out_a, char* out_b,
                       int* out_len, int* out_score) {
    static int T[3][TBL * TBL];
    static int ready[3];
    long alo, ahi, blo, bhi;
    int rc;
    if (matrix_id < 0 || matrix_id > 2) return GOTOH_ORACLE_EDOMAIN;
    if (!ready[matrix_id]) { /* the reference rewrites its table per call (:637,673,707); values are identical */
        if (matrix_id == GOTOH_ORACLE_NT) table_nt(T[0], 5, 4);
        else if (matrix_id == GOTOH_ORACLE_HIV25) table_hiv25(T[1]);
        else table_aa_rb(T[2], 4, -2);
        ready[matrix_id] = 1;
    }
    trim_span(a, a_len, &alo, &ahi);                 /* :641-642 */
    trim_span(b, b_len, &blo, &bhi);
    if (matrix_id == GOTOH_ORACLE_AA_RB) {           /* :713-718: degap both, term forced to 0 */
        char* da = (char*)malloc((size_t)(ahi - alo) + 1);
        char* db = (char*)malloc((size_t)(bhi - blo) + 1);
        long m, n;
        if (!da || !db) { free(da); free(db); return GOTOH_ORACLE_ENOMEM; }
        m = degap_copy(a + alo, ahi - alo, da);
        n = degap_copy(b + blo, bhi - blo, db);
        rc = align_core(T[2], da, m, db, n, gip, gep, 0, out_a, out_b, out_len, out_score);
        free(da); free(db);
        return rc;
    }
    return align_core(T[matrix_id], a + alo, ahi - alo, b + blo, bhi - blo, gip, gep, term,
                      out_a, out_b, out_len, out_score);
}

int gotoh_oracle_align_batch(int matrix_id, const char* ref_bytes, const long long* ref_off,
                             const int* ref_idx, const char* qry_bytes, const long long* qry_off,
                             long long first, long long last, int gip, int gep, int term,
                             char* out_a, char* out_b, const long long* out_off,
                             int* out_len, int* out_score) {
    long long k;
    for (k = first; k < last; ++k) {
        long long r = ref_idx ? ref_idx[k] : k;
        int rc = gotoh_oracle_align(matrix_id, ref_bytes + ref_off[r], (long)(ref_off[r + 1] - ref_off[r]),
                                    qry_bytes + qry_off[k], (long)(qry_off[k + 1] - qry_off[k]),
                                    gip, gep, term, out_a + out_off[k], out_b + out_off[k],
                                    out_len + k, out_score + k);
        if (rc) return rc;
    }
    return 0;
}

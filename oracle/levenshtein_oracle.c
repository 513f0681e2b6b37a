/*
 * levenshtein_oracle.c - CPU oracle for `Levenshtein.distance(a, b)` as used by the seed filter of
 * remap.sam_to_conseqs (/root/reference/micall/core/remap.py:250).
 *
 * TEST INFRASTRUCTURE ONLY - never linked into the product library.
 *
 * The arithmetic lives in a third-party dependency that is NOT in /root/reference: python-Levenshtein
 * (INSTALL.md:8,22 "sudo apt install python3-levenshtein"; no version is pinned anywhere in the tree) and is not
 * installed in this image.  This file restates the published algorithm its `distance()` documents - the classic
 * Wagner-Fischer / Levenshtein unit-cost edit distance (insert, delete, substitute all cost 1) - with two rolling
 * rows.  Parity status: PINNED by the reference's own expectations for this call site: driving the reference's
 * unmodified remap.sam_to_conseqs with this function as `Levenshtein.distance` reproduces every distance and every
 * kept/dropped consensus its tests assert (micall/tests/remap_test.py:415-545, incl. seed_dist=2/other_dist=5/...);
 * tests/golden/make_golden_callers.py does exactly that and commits the vectors.
 */
#include <stdlib.h>

long levenshtein_oracle(const unsigned char* a, long la, const unsigned char* b, long lb) {
    long *prev, *cur, *tmp, i, j, r;
    if (la == 0) return lb;
    if (lb == 0) return la;
    prev = (long*)malloc(sizeof(long) * (size_t)(lb + 1) * 2);
    if (!prev) return -1;
    cur = prev + (lb + 1);
    for (j = 0; j <= lb; ++j) prev[j] = j;
    for (i = 1; i <= la; ++i) {
        cur[0] = i;
        for (j = 1; j <= lb; ++j) {
            long best = prev[j - 1] + (a[i - 1] != b[j - 1]);      /* substitute / match */
            if (prev[j] + 1 < best) best = prev[j] + 1;             /* delete from a */
            if (cur[j - 1] + 1 < best) best = cur[j - 1] + 1;       /* insert into a */
            cur[j] = best;
        }
        tmp = prev; prev = cur; cur = tmp;
    }
    r = prev[lb];
    free(prev < cur ? prev : cur);
    return r;
}

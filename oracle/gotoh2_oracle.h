/* gotoh2_oracle.h - CPU restatement of MiCall-Lite's _gotoh2.align (TEST INFRASTRUCTURE ONLY). */
#ifndef GOTOH2_ORACLE_H
#define GOTOH2_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif
enum {
    GOTOH2_ORACLE_EEMPTY = -1,      /* gotoh2.py:84-85 asserts both sequences non-empty */
    GOTOH2_ORACLE_EDOMAIN = -2,     /* character outside the alphabet: map[] = -1 indexes d[] out of bounds (_gotoh2.c:185) */
    GOTOH2_ORACLE_ENOMEM = -3,
    GOTOH2_ORACLE_ETRACEBACK = -4   /* "Traceback failed, try local alignment" (_gotoh2.c:403-407,601-603) */
};
/* seq1/seq2: already cleaned (gotoh2.py:70-72); d: l*l substitution scores, row-major in alphabet order.
 * out1/out2 need l1+l2 bytes. */
int gotoh2_oracle_align(const char* seq1, long l1, const char* seq2, long l2, int gop, int gep,
                        int is_global, const char* alphabet, const int* d, char* out1, char* out2,
                        int* out_len, int* out_score);
#ifdef __cplusplus
}
#endif
#endif

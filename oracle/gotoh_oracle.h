/* gotoh_oracle.h - CPU restatement of MiCall-Lite's align_it/align_it_aa path.
 * TEST INFRASTRUCTURE ONLY (see gotoh_oracle.c). */
#ifndef GOTOH_ORACLE_H
#define GOTOH_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

enum { GOTOH_ORACLE_NT = 0, GOTOH_ORACLE_HIV25 = 1, GOTOH_ORACLE_AA_RB = 2 };
enum {
    GOTOH_ORACLE_EEMPTY = -1,    /* empty after trim: UB in the reference */
    GOTOH_ORACLE_EDOMAIN = -2,   /* byte outside 1..126: out-of-bounds table read in the reference */
    GOTOH_ORACLE_ENOMEM = -3,
    GOTOH_ORACLE_ESENTINEL = -4  /* all boundary scores < -100000: uninitialised read in the reference */
};

/* pairscore() for all 127x127 byte pairs after init_pairscore(5,4) / _hiv25() / _aa(4,-2). */
void gotoh_oracle_table(int matrix_id, int* out127x127);

/* One alignment with the wrapper semantics of gotoh.cpp:624-727 (trim, degap for AA_RB).
 * out_a/out_b need a_len+b_len bytes; no terminator is written. */
int gotoh_oracle_align(int matrix_id, const char* a, long a_len, const char* b, long b_len,
                       int gip, int gep, int term, char* out_a, char* out_b,
                       int* out_len, int* out_score);

/* Packed batch form, pairs [first,last). ref_idx may be NULL (pair k uses ref k). */
int gotoh_oracle_align_batch(int matrix_id, const char* ref_bytes, const long long* ref_off,
                             const int* ref_idx, const char* qry_bytes, const long long* qry_off,
                             long long first, long long last, int gip, int gep, int term,
                             char* out_a, char* out_b, const long long* out_off,
                             int* out_len, int* out_score);
#ifdef __cplusplus
}
#endif
#endif

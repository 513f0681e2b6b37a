// gotoh_tables.h - K0: the three substitution tables of the reference, rebuilt from rules.
//
// Reference: /root/reference/micall/alignment/gotoh.cpp
//   init_pairscore(5,4)       :26-131   (align_it)
//   init_pairscore_hiv25()    :165-213  (align_it_aa)
//   init_pairscore_aa(4,-2)   :134-157  (align_it_aa_rb)
// The reference refills a process-global int[127][127] on every call; the values never
// change, so here each table is built once on the host as a list of override rules applied
// in the reference's order (later rules win), then uploaded per plan as a compact
// [class][128] slice.  tests/test_host.py diffs all 3 x 127 x 127 entries against the
// table dumped from the compiled reference (tests/golden/pairscore_tables.json).
#pragma once
#include <stdint.h>
#include <string.h>

namespace gotoh {

struct ScoreTable {
    int v[128][128];
    int at(int a, int b) const { return v[a][b]; }
};

namespace detail {

inline void sym(ScoreTable& t, const char* xs, const char* ys, int val, bool both_cases_of_y) {
    for (const char* x = xs; *x; ++x)
        for (const char* y = ys; *y; ++y) {
            const int a = (unsigned char)*x, b = (unsigned char)*y;
            t.v[a][b] = t.v[b][a] = val;
            if (both_cases_of_y) t.v[a][b + 32] = t.v[b + 32][a] = val;
        }
}

inline void build_nt(ScoreTable& t, int match, int mismatch) {
    for (int a = 0; a < 128; ++a)
        for (int b = 0; b < 128; ++b) t.v[a][b] = (a == b) ? match : -mismatch;
    // upper/lower identity for A C G T U (gotoh.cpp:48-51)
    for (const char* c = "ACGTU"; *c; ++c) t.v[(int)*c][*c + 32] = t.v[*c + 32][(int)*c] = match;
    // T ~ U in the directions the reference lists (gotoh.cpp:52-53): every (t|T , u|U) and
    // (u|U , t|T) ordered pair EXCEPT ('u','T'), which the reference never assigns.
    const char tt[2] = {'t', 'T'}, uu[2] = {'u', 'U'};
    for (int x = 0; x < 2; ++x)
        for (int y = 0; y < 2; ++y) {
            t.v[(int)tt[x]][(int)uu[y]] = match;
            if (!(uu[y] == 'u' && tt[x] == 'T')) t.v[(int)uu[y]][(int)tt[x]] = match;
        }
    t.v['N']['N'] = t.v['n']['N'] = t.v['N']['n'] = 0;   // :54
    // IUPAC mixtures match their members, upper case only (:58-91)
    sym(t, "R", "AG", match, false); sym(t, "Y", "CT", match, false); sym(t, "K", "GT", match, false);
    sym(t, "M", "CA", match, false); sym(t, "S", "CG", match, false); sym(t, "W", "TA", match, false);
    sym(t, "B", "CGT", match, false); sym(t, "D", "AGT", match, false);
    sym(t, "H", "ACT", match, false); sym(t, "V", "ACG", match, false);
    sym(t, "*", "ACTG", match, true);                    // wildcard :94-97
    t.v['$']['$'] = 50;                                  // :99
    sym(t, ".", "ACTG", -20, true);                      // :105-108
    sym(t, "N", "ACTG", -3, true);                       // :110-113
    sym(t, "X", "ACTGRYKMSWBDHV", -6, true);             // :116-129
    t.v['X']['-'] = 3;                                   // :130 (one direction only)
}

inline void build_aa_rb(ScoreTable& t, int match, int mismatch_penalty) {
    for (int a = 0; a < 128; ++a)
        for (int b = 0; b < 128; ++b) {
            int s = (a == b) ? match : -mismatch_penalty;
            if (a != b && (a == 'X' || b == 'X')) s = -4;     // :147-150
            t.v[a][b] = s;
        }
    t.v['Z']['Z'] = t.v['z']['Z'] = t.v['Z']['z'] = 0;        // :155
    t.v['X']['-'] = t.v['-']['X'] = match;                    // :156
}

// Nickle et al. 2007 (PLoS One 2(6):e503) HIV-specific 25%-divergence matrix as shipped by
// the reference (gotoh.cpp:165-189); rows/cols in the order of kHiv25Alphabet.  Stored as the
// upper triangle incl. diagonal - the published matrix is symmetric (asserted at build).
static const char kHiv25Alphabet[] = "ARNDCQEGHILKMFPSTWYVBZ?*";
static const signed char kHiv25Upper[] = {
    7, -7, -7, -4, -10, -11, -4, -3, -10, -6, -9, -9, -7, -13, -3, -2, 1, -16, -15, 0, -5, -5, -3, -17,
    7, -5, -11, -8, -2, -7, -2, 0, -6, -6, 2, -3, -12, -4, -2, -2, -5, -9, -10, -7, -3, -3, -17,
    8, 2, -9, -6, -6, -7, 0, -6, -12, 0, -10, -12, -9, 1, 0, -17, -3, -10, 6, -6, -3, -17,
    8, -14, -10, 0, -2, -3, -11, -15, -7, -13, -15, -13, -5, -6, -16, -6, -5, 7, 0, -3, -17,
    11, -16, -15, -5, -7, -11, -9, -13, -14, 0, -12, -1, -6, -2, 0, -8, -10, -16, -5, -17,
    8, -2, -10, 0, -12, -4, 0, -8, -12, -1, -9, -8, -14, -9, -13, -7, 6, -4, -17,
    7, -1, -9, -12, -15, -1, -10, -17, -13, -11, -8, -15, -12, -5, 0, 6, -4, -17,
    7, -10, -11, -14, -6, -12, -9, -11, -1, -7, -5, -14, -5, -4, -3, -4, -17,
    10, -10, -4, -5, -10, -6, -3, -6, -6, -11, 2, -14, -1, -2, -3, -17,
    7, 0, -7, 0, -2, -10, -4, 0, -14, -9, 2, -7, -12, -2, -17,
    6, -10, 0, 0, -3, -5, -8, -6, -8, -4, -13, -6, -4, -17,
    7, -4, -14, -9, -5, -1, -12, -13, -9, -1, -1, -2, -17,
    10, -7, -11, -9, -1, -11, -15, 0, -11, -9, -3, -17,
    10, -11, -5, -10, -5, 1, -5, -13, -14, -3, -17,
    8, -1, -3, -13, -11, -12, -10, -3, -5, -17,
    8, 0, -12, -6, -9, 0, -10, -3, -17,
    7, -16, -10, -4, -2, -8, -2, -17,
    10, -4, -16, -16, -14, -8, -17,
    10, -12, -4, -10, -4, -17,
    7, -7, -7, -3, -17,
    7, -2, -4, -17,
    6, -4, -17,
    -3, -17,
    1};

inline void build_hiv25(ScoreTable& t) {
    memset(t.v, 0, sizeof(t.v));                              // :198-202
    int idx = 0;
    for (int i = 0; i < 24; ++i)
        for (int j = i; j < 24; ++j) {
            const int s = kHiv25Upper[idx++];
            const int a = kHiv25Alphabet[i], b = kHiv25Alphabet[j];
            // "+32" aliasing applied to every letter, also '?'->'_' and '*'->'J' (:210)
            for (int ca = 0; ca < 2; ++ca)
                for (int cb = 0; cb < 2; ++cb) {
                    t.v[a + 32 * ca][b + 32 * cb] = s;
                    t.v[b + 32 * cb][a + 32 * ca] = s;
                }
        }
}

}  // namespace detail

// matrix_id as in gotoh_b200.h.  Built once (thread-safe function-local static), immutable after.
struct ScoreTables {
    ScoreTable t[3];
    ScoreTables() {
        detail::build_nt(t[0], 5, 4);        // gotoh.cpp:637
        detail::build_hiv25(t[1]);           // gotoh.cpp:673
        detail::build_aa_rb(t[2], 4, -2);    // gotoh.cpp:707
    }
};
inline const ScoreTable& score_table(int matrix_id) {
    static const ScoreTables tabs;
    return tabs.t[matrix_id];
}

}  // namespace gotoh

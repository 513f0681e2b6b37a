// TEST INFRASTRUCTURE ONLY (parity oracle / CPU baseline) - never linked into the product.
//
// Compiles the reference's own, untouched aligner by #including it from where it
// lies (/root/reference/micall/alignment/gotoh.cpp; path supplied by the Makefile as
// -DGOTOH_REF_SOURCE=...), with oracle/stub/ruby.h on the include path so the
// non-Python branch (gotoh.cpp:5-6,741-797) compiles.  No reference source is copied.
//
// The exported entry points do exactly what the reference's Python wrappers do
// (gotoh.cpp:624-727): table init -> trim (-> degap) -> align -> copy the strings out.
#include <cstdarg>
#include <cstring>
#include GOTOH_REF_SOURCE

extern "C" {
// Ruby runtime symbols named by the reference's (dead) Ruby binding; never called.
VALUE rb_ary_new3(long, ...) { return 0; }
VALUE rb_str_new2(const char*) { return 0; }
void rb_define_global_function(const char*, VALUE (*)(...), int) {}

// mode: 0 = align_it (init_pairscore(5,4), gotoh.cpp:637)
//       1 = align_it_aa (init_pairscore_hiv25, gotoh.cpp:673)
//       2 = align_it_aa_rb (init_pairscore_aa(4,-2) + degap, term=0, gotoh.cpp:707-718)
static void ref_init_table(int mode) {
    if (mode == 0) init_pairscore(5, 4);
    else if (mode == 1) init_pairscore_hiv25();
    else init_pairscore_aa(4, -2);
}

// out_a/out_b must hold strlen(a)+strlen(b)+1 bytes.  Returns the score.
int ref_align(int mode, const char* a, const char* b, int gip, int gep, int term,
              char* out_a, char* out_b) {
    ref_init_table(mode);
    std::string sa(a), sb(b), na, nb;
    trim(&sa);
    trim(&sb);
    if (mode == 2) { degap(&sa); degap(&sb); term = 0; }
    int score = align(&sa, &sb, &na, &nb, gip, gep, term);
    std::memcpy(out_a, na.c_str(), na.size() + 1);
    std::memcpy(out_b, nb.c_str(), nb.size() + 1);
    return score;
}

// Batched form for the CPU baseline: packed bytes + offsets, same table init per
// pair as the wrappers do.  Output stride per pair = out_off[k+1]-out_off[k].
void ref_align_batch(int mode, const char* ref_bytes, const long long* ref_off,
                     const int* ref_idx, const char* qry_bytes, const long long* qry_off,
                     long long first, long long last, int gip, int gep, int term,
                     char* out_a, char* out_b, const long long* out_off, int* out_len,
                     int* out_score) {
    for (long long k = first; k < last; ++k) {
        long long r = ref_idx ? ref_idx[k] : k;
        std::string sa(ref_bytes + ref_off[r], ref_bytes + ref_off[r + 1]);
        std::string sb(qry_bytes + qry_off[k], qry_bytes + qry_off[k + 1]);
        std::string na, nb;
        ref_init_table(mode);
        trim(&sa);
        trim(&sb);
        int t = term;
        if (mode == 2) { degap(&sa); degap(&sb); t = 0; }
        out_score[k] = align(&sa, &sb, &na, &nb, gip, gep, t);
        out_len[k] = (int)na.size();
        std::memcpy(out_a + out_off[k], na.data(), na.size());
        std::memcpy(out_b + out_off[k], nb.data(), nb.size());
    }
}

// Dump pairscore() for every (a,b) in 0..126 after the reference's own init.
void ref_pairscore_table(int mode, int* out /* 127*127 */) {
    ref_init_table(mode);
    for (int a = 0; a < 127; ++a)
        for (int b = 0; b < 127; ++b) out[a * 127 + b] = pairscore((char)a, (char)b);
}
}

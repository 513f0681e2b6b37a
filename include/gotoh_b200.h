/*
 * gotoh_b200.h - C ABI of libgotoh_b200.so, the B200-native (sm_100a CUDA) replacement
 * for MiCall-Lite's Gotoh aligner hot path.
 *
 * What it replaces.  The reference exposes this path only as a CPython extension
 * ("gotoh", /root/reference/micall/alignment/gotoh.cpp:729-739) with three callables:
 *     align_it      (gotoh.cpp:624-658)   nt,  init_pairscore(5,4)
 *     align_it_aa   (gotoh.cpp:660-693)   aa,  init_pairscore_hiv25()
 *     align_it_aa_rb(gotoh.cpp:695-727)   aa,  init_pairscore_aa(4,-2), degap, term=0
 * all three being: score-table init -> trim -> align() (gotoh.cpp:233-527) -> two strings
 * and a score.  There is no C ABI in the reference; this header is the FFI a maintainer
 * would bind instead (ctypes stub shown in INTEGRATION.md).  Plain pointers and sizes
 * only; the caller owns every input and output buffer; nothing returned is owned by the
 * library except opaque plan handles.
 *
 * No CPU fallback exists: every compute entry point fails with GOTOH_B200_ENODEVICE when
 * no CUDA device is usable.
 */
#ifndef GOTOH_B200_H
#define GOTOH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GOTOH_B200_VERSION 200 /* 0.2.0 */

/* matrix_id: which of the reference's table initialisers applies (gotoh.cpp:26-213). */
enum {
    GOTOH_B200_NT = 0,    /* align_it:       init_pairscore(5,4)          gotoh.cpp:637 */
    GOTOH_B200_HIV25 = 1, /* align_it_aa:    init_pairscore_hiv25()       gotoh.cpp:673 */
    GOTOH_B200_AA_RB = 2  /* align_it_aa_rb: init_pairscore_aa(4,-2), inputs degapped,
                             use_terminal forced to 0          gotoh.cpp:707-718 */
};

/* Return codes (0 = success).  The reference has no error path after argument parsing
 * (gotoh.cpp:633-635); inputs outside its defined domain are undefined behaviour there
 *  and are rejected here instead. */
enum {
    GOTOH_B200_OK = 0,
    GOTOH_B200_EINVAL = -1,    /* bad argument (NULL pointer, negative count, bad matrix_id ...) */
    GOTOH_B200_EEMPTY = -2,    /* a sequence is empty after trim/degap: trim() reads seq[size-1], gotoh.cpp:555 */
    GOTOH_B200_EDOMAIN = -3,   /* a byte outside 1..126: pairscore() index, gotoh.cpp:216-219 */
    GOTOH_B200_ESENTINEL = -4, /* 2*gip+(max(M,N)+1)*gep >= 100000: boundary scores could fall below the
                                  -100000 sentinel, gotoh.cpp:284-286 (uninitialised maxij/maxji) */
    GOTOH_B200_ERANGE = -5,    /* output stride < M+N, lengths/penalties outside int32-safe range */
    GOTOH_B200_ENODEVICE = -6, /* no usable CUDA device / device index not present */
    GOTOH_B200_ECUDA = -7,     /* CUDA runtime error (message in gotoh_b200_last_error) */
    GOTOH_B200_ENOMEM = -8,    /* host or device allocation failed */
    GOTOH_B200_ETRACEBACK = -9, /* gotoh2 only: no a/b/c bit set on the path, "Traceback failed, try local
                                  alignment" (_gotoh2.c:403-407,601-603); that pair's score is INT32_MIN */
    GOTOH_B200_ECAPACITY = -10 /* tight / compact result forms: the caller's output buffer is too small for the
                                  results; gotoh_b200_last_error() names the size that is needed */
};

/* Library / device info. */
int32_t gotoh_b200_version(void);
/* Thread-local, NUL-terminated description of the last failure on this thread. */
const char* gotoh_b200_last_error(void);
/* Number of visible CUDA devices (0 if none / driver missing). */
int32_t gotoh_b200_device_count(void);

/* The score table exactly as the reference's pairscore() (gotoh.cpp:216-219) returns it
 * after init_pairscore(5,4) / init_pairscore_hiv25() / init_pairscore_aa(4,-2):
 * out[a*127+b] for a,b in 0..126.  Host-side builder, the same one the device tables are
 * uploaded from. */
int32_t gotoh_b200_pairscore_table(int32_t matrix_id, int32_t* out_127x127);

/*
 * One-shot batched alignment: host buffers in, host buffers out.  This is the batched
 * form of align_it / align_it_aa / align_it_aa_rb: pair k aligns
 *     standard = ref_bytes[ref_off[r] .. ref_off[r+1])   with r = ref_idx ? ref_idx[k] : k
 *     seq      = qry_bytes[qry_off[k] .. qry_off[k+1])
 * with the wrapper semantics of gotoh.cpp:624-727 (trim of " \t\n\r" on both, degap for
 * AA_RB).  Results for pair k:
 *     out_ref[out_off[k] .. +out_len[k])  aligned standard   (gotoh.cpp:650 1st string)
 *     out_qry[out_off[k] .. +out_len[k])  aligned seq        (2nd string)
 *     out_score[k]                        alignment score    (3rd value)
 * out_off[k+1]-out_off[k] must be >= M_k+N_k (trimmed lengths); the bytes between
 * out_len[k] and the pair's stride are set to 0.  No NUL terminators are written.
 *
 * n_refs    number of references (ref_off has n_refs+1 entries)
 * ref_idx   per-pair reference index, or NULL when n_refs == n_pairs and pair k uses ref k
 * device_mask  bit d set => shard onto CUDA device d (0 => device 0 only).  Pairs are
 *           statically partitioned by cell count; there is no inter-device traffic.
 */
int32_t gotoh_b200_align_batch(const uint8_t* ref_bytes, const int64_t* ref_off, int64_t n_refs,
                               const int32_t* ref_idx,
                               const uint8_t* qry_bytes, const int64_t* qry_off, int64_t n_pairs,
                               int32_t gip, int32_t gep, int32_t use_terminal, int32_t matrix_id,
                               uint8_t* out_ref, uint8_t* out_qry, const int64_t* out_off,
                               int32_t* out_len, int32_t* out_score, uint32_t device_mask);

/*
 * On a non-zero return of any batched entry point the contents of the caller's output buffers are unspecified
 * (slabs that were finished before the failure was detected may already have been copied back).
 *
 * Tight form of gotoh_b200_align_batch: the same results, but the library chooses the layout - pair k's two aligned
 * strings start at out_off[k] (an OUTPUT, n_pairs entries) and are out_len[k] bytes long, pairs follow each other
 * without stride tails, so only bytes that carry results cross PCIe (the strided form ships the zero tail between
 * out_len and the M+N stride: 7.6 % of the bytes on 251-nt reads).  out_cap = bytes available in EACH of out_ref /
 * out_qry; sum of (len(standard)+len(seq)) over the pairs always suffices (GOTOH_B200_ECAPACITY otherwise).  With
 * several devices in device_mask each device fills its own slice of the capacity, so out_off is increasing but not
 * gap-free across device boundaries.
 */
int32_t gotoh_b200_align_batch_tight(const uint8_t* ref_bytes, const int64_t* ref_off, int64_t n_refs,
                                     const int32_t* ref_idx,
                                     const uint8_t* qry_bytes, const int64_t* qry_off, int64_t n_pairs,
                                     int32_t gip, int32_t gep, int32_t use_terminal, int32_t matrix_id,
                                     uint8_t* out_ref, uint8_t* out_qry, int64_t out_cap, int64_t* out_off,
                                     int32_t* out_len, int32_t* out_score, uint32_t device_mask);

/*
 * Compact form: the alignment itself instead of its rendering.  The reference returns two strings of length ~M+N per
 * pair (gotoh.cpp:436-513, Py_BuildValue("ssi") :650) - for a 251-nt read against a 3039-nt standard 6.6 KB of which
 * 92 % is the standard's overhang against '-'.  This entry point returns, per pair k, one record of 8 int32
 *     out_rec[8k + GOTOH_B200_REC_*]:  SCORE (3rd value of align_it), OUT_LEN (length of either aligned string),
 *         I0, J0   cell where the traceback stopped (gotoh.cpp:452 loop exit; one of them is 0): the left overhang is
 *                  standard[0..I0) against '-' or seq[0..J0) against '-'                       (gotoh.cpp:489-496)
 *         END_I, END_J  end cell chosen at gotoh.cpp:418-450: the right overhang is seq[END_J..N) against '-' when
 *                  END_I == M and END_J < N, else standard[END_I..M) against '-'
 *         N_OPS    traceback steps between them, M_N = (M << 16 | N) when both trimmed lengths are < 65536, else -1
 * and the op script: op t (t = 0 .. N_OPS-1, counted from the END cell backwards, exactly the order of the reference's
 * traceback loop gotoh.cpp:452-487) is bits [2(t&15)+1 : 2(t&15)] of word out_ops[out_ops_off[k] + (t>>4)]:
 *     0 diagonal (one character of each), 1 up (a standard character against '-'), 2 left ('-' against a seq character).
 * (standard, seq are the TRIMMED - AA_RB: and degapped - inputs.)  The two strings of align_it are a pure function of
 * the inputs and this record (gotoh_b200/compact.py expands them on demand; tests compare with the string entry point
 * and the oracle byte for byte).  out_ops_cap = capacity of out_ops in 32-bit words; sum of ceil((M+N)/16) always
 * suffices, real alignments need ~ceil(min(M,N)/16)+1 per pair (GOTOH_B200_ECAPACITY names the need).  out_ops_off has
 * n_pairs entries (OUTPUT); the scripts of different pairs follow each other in no particular order.  ~100 B per 251-nt read cross PCIe instead of 6.6 KB.
 */
enum { GOTOH_B200_REC_SCORE = 0, GOTOH_B200_REC_OUT_LEN = 1, GOTOH_B200_REC_I0 = 2, GOTOH_B200_REC_J0 = 3,
       GOTOH_B200_REC_END_I = 4, GOTOH_B200_REC_END_J = 5, GOTOH_B200_REC_N_OPS = 6, GOTOH_B200_REC_M_N = 7,
       GOTOH_B200_REC_WORDS = 8 };
int32_t gotoh_b200_align_batch_compact(const uint8_t* ref_bytes, const int64_t* ref_off, int64_t n_refs,
                                       const int32_t* ref_idx,
                                       const uint8_t* qry_bytes, const int64_t* qry_off, int64_t n_pairs,
                                       int32_t gip, int32_t gep, int32_t use_terminal, int32_t matrix_id,
                                       int32_t* out_rec, uint32_t* out_ops, int64_t out_ops_cap,
                                       int64_t* out_ops_off, uint32_t device_mask);

/*
 * Host-ceiling probe for the result copy (bench.py: e2e.host_ceiling): `reps` plain cudaMemcpyAsync device-to-host
 * copies of `bytes` from a scratch device buffer on `device` into host_buf (pinned memory from gotoh_b200_host_alloc
 * for the full rate), timed with CUDA events; *seconds = time of all reps.  Run on every rank at once it measures what
 * the box can absorb, which is what bounds the string forms on multi-GPU hosts.
 */
int32_t gotoh_b200_d2h_probe(int32_t device, void* host_buf, int64_t bytes, int32_t reps, double* seconds);

/*
 * Staged form of the same call, for callers that keep data resident in HBM or want the
 * device time of each phase:
 *   plan_create  validate + trim + bucket + pack on the host, allocate HBM, copy inputs H2D
 *   plan_run     forward DP + traceback + string emit on the device; results stay in HBM.
 *                Re-runnable; *device_ms (optional) = CUDA-event time of this run on the
 *                plan's stream, *forward_ms (optional) = the forward-DP kernels alone.
 *   plan_fetch   copy results D2H into caller buffers (same meaning as align_batch)
 *   plan_destroy release everything
 * A plan is bound to one device and must be used by one host thread at a time.
 */
typedef struct gotoh_b200_plan gotoh_b200_plan;

int32_t gotoh_b200_plan_create(int32_t device,
                               const uint8_t* ref_bytes, const int64_t* ref_off, int64_t n_refs,
                               const int32_t* ref_idx,
                               const uint8_t* qry_bytes, const int64_t* qry_off, int64_t n_pairs,
                               int32_t gip, int32_t gep, int32_t use_terminal, int32_t matrix_id,
                               const int64_t* out_off, gotoh_b200_plan** plan_out);
int32_t gotoh_b200_plan_run(gotoh_b200_plan* plan, float* device_ms, float* forward_ms);
int32_t gotoh_b200_plan_fetch(gotoh_b200_plan* plan, uint8_t* out_ref, uint8_t* out_qry,
                              int32_t* out_len, int32_t* out_score);
void gotoh_b200_plan_destroy(gotoh_b200_plan* plan);

/* Introspection of a plan (for benchmarks and tests).  what:
 *   0 total DP cells (sum M*N)          1 kernel launches per plan_run
 *   2 bytes copied H2D by plan_create   3 bytes copied D2H by plan_fetch
 *   4 direction-arena bytes in HBM      5 pairs on the 16-bit x2 path
 *   6 pairs on the 32-bit path          7 number of arena chunks per run
 *   8 1 if the plan was laid out by the device-side builder (csrc/gotoh_prep.cuh), 0 if by the host builder */
int64_t gotoh_b200_plan_stat(const gotoh_b200_plan* plan, int32_t what);

/* gotoh_b200_align_batch keeps three workspaces (device buffers, pinned staging, a stream) per
 * device alive between calls so that small calls do not pay for cudaMalloc; this frees them. */
void gotoh_b200_release_cache(void);

/* Pinned (page-locked) host memory for callers that want full-rate H2D/D2H copies. */
void* gotoh_b200_host_alloc(int64_t bytes);
void gotoh_b200_host_free(void* p);

/*
 * NEXT #1: the aligner the live pipeline calls, gotoh2.Aligner.align ->
 * _gotoh2.align (micall/alignment/gotoh2.py:74-96, src/_gotoh2.c:442-607): Altschul-Erickson
 * min-cost affine alignment, global or "local" (free end gaps), alphabet-indexed substitution
 * matrix.  Pair k aligns seq1 = s1_bytes[s1_off[r]..s1_off[r+1]) (r = s1_idx ? s1_idx[k] : k)
 * with seq2 = s2_bytes[s2_off[k]..s2_off[k+1]).  Bytes are cleaned like gotoh2.py:70-72 (ASCII
 * upper-case, non-alphabet -> '?'); no trimming.  alphabet: NUL-terminated, <= 32 letters;
 * matrix: l*l scores, row-major in alphabet order (gotoh2.py:47-64).  Outputs as in
 * gotoh_b200_align_batch (the aligned strings contain the CLEANED characters, like the reference).
 * Returns GOTOH_B200_ETRACEBACK if some pair's traceback fails (its out_score is INT32_MIN, its
 * out_len 0; all other pairs are valid).
 */
int32_t gotoh_b200_gotoh2_align_batch(const uint8_t* s1_bytes, const int64_t* s1_off, int64_t n_s1,
                                      const int32_t* s1_idx,
                                      const uint8_t* s2_bytes, const int64_t* s2_off, int64_t n_pairs,
                                      int32_t gop, int32_t gep, int32_t is_global,
                                      const char* alphabet, const int32_t* matrix,
                                      uint8_t* out1, uint8_t* out2, const int64_t* out_off,
                                      int32_t* out_len, int32_t* out_score, int32_t device);

/*
 * NEXT #3: the edit distance of the seed filter in remap.sam_to_conseqs,
 * `Levenshtein.distance(relevant_seed, relevant_conseq)` (micall/core/remap.py:250; third-party python-Levenshtein,
 * INSTALL.md:8,22).  out_dist[k] = unit-cost insert/delete/substitute distance between
 * a_bytes[a_off[k]..a_off[k+1]) and b_bytes[b_off[k]..b_off[k+1]), compared byte by byte (no cleaning); empty
 * strings are allowed.  Runs the score-only forward kernel (no traceback arena).  At most 30 distinct bytes may
 * occur on BOTH sides of the batch (GOTOH_B200_ERANGE otherwise).
 */
int32_t gotoh_b200_edit_distance_batch(const uint8_t* a_bytes, const int64_t* a_off,
                                       const uint8_t* b_bytes, const int64_t* b_off, int64_t n_pairs,
                                       int32_t* out_dist, int32_t device);

/* Statistics of the calling thread's last gotoh_b200_gotoh2_align_batch or gotoh_b200_edit_distance_batch (benchmarks, tests): fills up to
 * n <= 11 doubles and returns how many: 0 grid cells sum (l1+1)(l2+1), 1 device ms of all kernels (CUDA
 * events), 2 forward ms, 3 reverse-sweep ms, 4 walk+emit ms, 5 kernel launches, 6 tie-bit arena bytes,
 * 7 arena chunks, 8 bytes copied H2D, 9 bytes copied D2H, 10 forward warp tasks that ran in int16x2 (two pairs sharing seq1
 * per warp; 0 when the 16-bit range proof failed or GOTOH_B200_GOTOH2=x1 pinned the int32 forward kernel). */
int32_t gotoh_b200_gotoh2_last_stats(double* out, int32_t n);

/* Integer-issue microbenchmark used for the roofline denominator ("peak
 * INT32 issue must be measured").  Runs `which` (0 IADD3, 1 VIMNMX, 2 VIADDMNMX,
 * 3 VIADDMNMX.S16x2, 4 VIMNMX3, 5 IMAD, 6 LOP3, 7 mixed ALU+IMAD, 8 forward-DP cell mix)
 * on `device` and returns giga warp-instructions... see DESIGN.md; result in
 * *ginstr_per_s (thread-level instructions per second / 1e9). */
int32_t gotoh_b200_int_peak(int32_t device, int32_t which, double* ginstr_per_s);

#ifdef __cplusplus
}
#endif
#endif /* GOTOH_B200_H */

// gotoh_kernels.cuh - sm_100a kernels for MiCall-Lite's Gotoh aligner hot path
// (reference: /root/reference/micall/alignment/gotoh.cpp:233-527).
//
//   k_forward<V,K,MULTI>  K1/K2/K3: affine-gap forward DP, one warp per task, anti-diagonal
//                   wavefront across the 32 lanes (lane l owns K query columns and is one
//                   reference row behind lane l-1), register-resident S/P/Q, shuffle
//                   hand-off, 2-bit directions packed and stored coalesced, end-cell
//                   selection fused into the epilogue.   (gotoh.cpp:288-450)
//   k_walk          K4a: pointer-chase of the packed directions, one thread per pair,
//                   emits a 2-bit op script + the terminal-gap score fix-up. (gotoh.cpp:452-510)
//   k_emit          K4b: expands the op script into the two aligned strings with
//                   coalesced stores, one warp per pair.                    (gotoh.cpp:436-513)
//
// Arithmetic is exact integer arithmetic; the two "vector" policies below only differ in
// how many alignments share a 32-bit register:
//   Vec32  one alignment per warp, int32 cells      (any length, any score range)
//   Vec16  two alignments per warp, int16x2 cells    (VIADDMNMX.S16x2 / VIMNMX3.S16x2; host
//          proves the score range fits before choosing it - see plan.cpp "range proof")
//
// Frames: with g = gep, u = -gip every stored value is
//   X^(i,j) = 4 * ( X(i,j) + (i - base(i) + j) * g ) + tag,   tag: D=0, P=1, Q=2
// The (i+j)*g shift removes the "+v" from both gap recurrences (gotoh.cpp:305-314); the
// factor 4 and the tags make max3() of the three candidates carry the reference's
// tie-break LEFT > UP > DIAG (gotoh.cpp:362-395) in its two low bits, which ARE the
// direction code.  base(i) = R*floor(i/R) keeps int16 in range (Vec16 only).
#pragma once

#include <stdint.h>

#ifndef GOTOH_SIMT_EMU
#include <cuda_runtime.h>
#define GOTOH_LAUNCH(kern, grid, block, smem, stream, ...)            \
    do {                                                              \
        auto _gotoh_k = kern;                                         \
        _gotoh_k<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__); \
    } while (0)
#define GOTOH_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

namespace gotoh {

enum { DIR_DIAG = 0, DIR_UP = 1, DIR_LEFT = 2 };
enum { FWD_WARPS = 4 };          // warps per CTA in k_forward
#ifndef GOTOH_MIN_CTAS
#define GOTOH_MIN_CTAS 4         // resident CTAs per SM the forward kernel is compiled for (register cap 65536/(128*N))
#endif
enum { REF_PAD = 64 };           // class bytes of padding on both sides of every reference
enum { PAD_CLASS = 0 };

// One alignment (pair), in plan order.  64 bytes.
struct PairInfo {
    int64_t ref_pos;   // index of row 1's byte in d_ref_raw / d_ref_cls (both carry REF_PAD padding)
    int64_t qry_pos;   // index of column 1's byte in d_qry
    int64_t dir_off;   // first uint4 of this pair's (task's) direction arena
    int64_t out_off;   // first output byte (caller's out_off[orig])
    int32_t M, N;      // trimmed lengths: rows (standard) and columns (seq)
    int32_t nblk;      // step blocks per strip in the arena
    int32_t ops_off;   // first uint32 of this pair's op script
    int16_t K;         // query columns per lane
    int8_t x2;         // 1: Vec16 arena (32 dir bits per lane-step, two alignments)
    int8_t half;       // bit 0: which half of the Vec16 word belongs to this pair; bit 1: second couple of a HALF-kernel warp (lanes 16-31)
    int32_t orig;      // caller's pair index (relative to the plan's first pair)
    int32_t out_cap;   // caller's output stride for this pair (>= M+N); the tail past out_len is zeroed
    int32_t pad1;
};

struct Task {          // one warp's work item
    int32_t pair_a;
    int32_t pair_b;    // -1: none (Vec32, or odd leftover in Vec16)
};

struct FwdParams {
    const PairInfo* pairs;
    const Task* tasks;
    int32_t task_first, task_count;
    const uint8_t* ref_cls;     // class index per reference position (0 = padding class)
    const uint8_t* qry;
    const int32_t* table4;      // [ncls][128] : 4*(T[rep(c)][b] + 2*gep), row 0 unused
    const int32_t* bonus4;      // [ncls][8] : 4*6*popcount(rmask(c) & cmask) stop-codon bonuses (gotoh.cpp:324-344)
    int32_t has_dollar;         // some reference contains "$$$"
    int32_t ncls;               // number of classes incl. the padding class 0
    int32_t gip, gep;
    int32_t rebase_mask;        // Vec16: R-1 (R power of two); Vec32: unused
    int32_t smin_m1;            // lower bound of any true score, minus 1 (last-column tracking seed)
    int32_t zshift;             // Vec16: constant (multiple of 4) added to every stored value so that none is negative
    uint32_t four;              // always 4; passed at run time so that acc*four+x compiles to IMAD
    uint4* dir;                 // direction arena
    int2* bnd;                  // Vec32 multi-strip boundary columns: [task_slot][2][bnd_stride]
    int64_t bnd_stride;
    int32_t* score;             // per pair: S at the chosen end cell (before terminal fix-up)
    int32_t* end_i;
    int32_t* end_j;
    uint32_t* work_counter;     // dynamic task scheduler
    // strip dataflow (k_forward_flow): tasks are (pair_a = pair, pair_b = strip); slot = pairs[pair].pad1 + strip
    int32_t stab_bytes;         // half-warp kernels: bytes of the CTA-wide int16 copy of table4 at the start of shared memory (0: none)
    int32_t task_limit;         // 0: persistent warps drain the queue; n > 0: a warp retires after n tasks (see plan_run)
    int32_t* prog;              // rows published per slot
    int32_t* part_best;         // last-row partial maximum per slot
    int32_t* part_j;
};

// ------------------------------------------------------------------------------------
// Vector policies
// ------------------------------------------------------------------------------------
struct Vec32 {
    typedef int T;
    enum { NPAIR = 1, STEPS = 8 };  // STEPS lane-steps fill one uint4 of directions
    static __device__ __forceinline__ T addmax(T a, T b, T c) { return __viaddmax_s32(a, b, c); }
    static __device__ __forceinline__ T max3(T a, T b, T c) { return __vimax3_s32(a, b, c); }
    static __device__ __forceinline__ T add(T a, T b) { return a + b; }
    // written as a multiply-add with a run-time 1 so that it issues as IMAD on the FMA pipe and the three-way maximum
    // stays ONE VIMNMX3 (ptxas otherwise fuses the add into a VIADDMNMX + VIMNMX pair on the saturated ALU pipe)
    static __device__ __forceinline__ T addlin(T a, unsigned one, T e) { return (T)((unsigned)a * one + (unsigned)e); }
    static __device__ __forceinline__ unsigned lin(int lo, int) { return (unsigned)lo; }
    static __device__ __forceinline__ T clr(T c) { return c & ~3; }
    static __device__ __forceinline__ T both(int x) { return x; }
    static __device__ __forceinline__ T pack(int lo, int) { return lo; }
    static __device__ __forceinline__ int lo(T v) { return v; }
    static __device__ __forceinline__ int hi(T v) { return v; }
    static __device__ __forceinline__ unsigned raw(T v) { return (unsigned)v; }
    // max(a,b) and the predicate a >= b (ties: the new candidate a wins)
    static __device__ __forceinline__ T bmax(T a, T b, bool* pa, bool* pb) { *pb = false; return __vibmax_s32(a, b, pa); }
};

struct Vec16 {
    typedef unsigned T;
    enum { NPAIR = 2, STEPS = 4 };
    static __device__ __forceinline__ T addmax(T a, T b, T c) { return __viaddmax_s16x2(a, b, c); }
    static __device__ __forceinline__ T max3(T a, T b, T c) { return __vimax3_s16x2(a, b, c); }
    static __device__ __forceinline__ T add(T a, T b) { return __vadd2(a, b); }
    // a + e per half as ONE 32-bit multiply-add on the FMA pipe: valid because the stored frame is shifted so that the
    // low halves of a and of a + e are never negative (packed == lo + 65536*hi then, and such "linear" words add
    // component-wise); e is stored linear (lin), `one` is a run-time 1 so that ptxas keeps the IMAD.
    static __device__ __forceinline__ T addlin(T a, unsigned one, T e) { return a * one + e; }
    static __device__ __forceinline__ unsigned lin(int lo, int hi) { return (unsigned)lo + ((unsigned)hi << 16); }
    static __device__ __forceinline__ T clr(T c) { return c & 0xfffcfffcu; }
    static __device__ __forceinline__ T pack(int lo, int hi) { return ((unsigned)lo & 0xffffu) | ((unsigned)hi << 16); }
    static __device__ __forceinline__ T both(int x) { return pack(x, x); }
    static __device__ __forceinline__ int lo(T v) { return (int)(short)(v & 0xffffu); }
    static __device__ __forceinline__ int hi(T v) { return (int)(short)(v >> 16); }
    static __device__ __forceinline__ unsigned raw(T v) { return v; }
    // per-half max and predicates a >= b: one VIMNMX.S16x2 with two predicate outputs (pa: low half)
    static __device__ __forceinline__ T bmax(T a, T b, bool* pa, bool* pb) { return __vibmax_s16x2(a, b, pb, pa); }
};

// Location of the 2-bit direction of cell (i,j) (1-based) inside a pair's arena.
// Arena layout: [strip][step block][lane] uint4; one uint4 = STEPS lane-steps.
//   Vec32: 16 bits per lane-step, two steps per 32-bit word (even step in the low half)
//   Vec16: 32 bits per lane-step, low half = pair_a, high half = pair_b
// Inside a 16-bit field, column k of the lane sits at bits [2(K-1-k)+1 : 2(K-1-k)].
struct DirAddr {
    int64_t word;   // index in uint32 units from the arena base
    int shift;
};
__device__ __forceinline__ DirAddr dir_addr(const PairInfo& p, int i, int j) {
    const int K = p.K;
    const int jj = j - 1;
    const int strip = jj / (32 * K);
    const int r = jj - strip * 32 * K;
    const int lane = r / K;
    const int k = r - lane * K;
    const int tt = i + lane - 1;  // step index, 0-based
    DirAddr a;
    if (p.x2) {
        // half bit 0: which 16-bit half of the word; bit 1 (HALF kernels): this pair's couple ran on lanes 16-31
        const int tb = tt >> 2, s = tt & 3;
        a.word = ((p.dir_off + ((int64_t)strip * p.nblk + tb) * 32 + lane + ((p.half & 2) ? 16 : 0)) << 2) + s;
        a.shift = 16 * (p.half & 1) + 2 * (K - 1 - k);
    } else {
        const int tb = tt >> 3, s = tt & 7;
        a.word = ((p.dir_off + ((int64_t)strip * p.nblk + tb) * 32 + lane) << 2) + (s >> 1);
        a.shift = 16 * (s & 1) + 2 * (K - 1 - k);
    }
    return a;
}

// ------------------------------------------------------------------------------------
// K1/K2/K3  forward DP
// ------------------------------------------------------------------------------------
// Which of the three stop-codon windows around column j (1-based) of query b[0..N) hold TAG/TAA/TGA.
__device__ __forceinline__ bool is_stop3(const uint8_t* b, int N, int p0) {
    if (p0 < 0 || p0 + 2 > N - 1) return false;   // reads past the end hit the terminator in the reference
    const uint8_t x = b[p0], y = b[p0 + 1], z = b[p0 + 2];
    return x == 'T' && ((y == 'A' && (z == 'G' || z == 'A')) || (y == 'G' && z == 'A'));
}
__device__ __forceinline__ int stop_mask(const uint8_t* b, int N, int j) {
    if (j < 3 || j > N) return 0;
    return (is_stop3(b, N, j - 3) ? 1 : 0) | (is_stop3(b, N, j - 2) ? 2 : 0) | (is_stop3(b, N, j - 1) ? 4 : 0);
}

template <class V, int K, bool LEAN = false>
struct FwdSmem {
    enum { K4 = (K + 3) / 4, USE2 = (K == 2 || K == 6), E2 = (K + 1) / 2, ROW_BYTES = USE2 ? E2 * 8 : K4 * 16 };
    // profile: [class][K4][lane] of 4 packed entries (K = 2, 6: [class][K/2][lane] of 2 entries - no padding words, so
    // a 21-class amino-acid profile at K = 6 leaves room for 3 CTAs per SM instead of 2);
    // + a 2x32 int2 ring for multi-strip boundaries
    // LEAN (half-warp wavefronts): + a slot of K words per lane for the lane's last-row values (kept at row M, reduced once,
    // uniformly, after the strip instead of inside a divergent branch)
    enum { SLOT_BYTES = LEAN ? K * 32 * 4 : 0 };
    static __host__ __device__ size_t prof_bytes(int ncls) { return (size_t)ncls * ROW_BYTES * 32; }
    static __host__ __device__ size_t per_warp(int ncls) { return prof_bytes(ncls) + 2 * 32 * sizeof(int2) + SLOT_BYTES; }
};

// State of one warp sweeping one strip (32*K query columns) down the reference.
// Every member is a register after inlining; all indexing is static.
__device__ __forceinline__ void gotoh_pause() {
#ifdef GOTOH_SIMT_EMU
    simt::yield();
#else
    __nanosleep(40);
#endif
}

// Cross-warp hand-over through global memory (strip dataflow): flags are polled with volatile loads, the data
// behind a flag is read with ld.global.cg (L2; L1 is not coherent between SMs).
__device__ __forceinline__ int ld_volatile(const int32_t* p) { return *reinterpret_cast<const volatile int32_t*>(p); }
__device__ __forceinline__ void st_volatile(int32_t* p, int v) { *reinterpret_cast<volatile int32_t*>(p) = v; }
#ifdef GOTOH_SIMT_EMU
template <class T> __device__ __forceinline__ T ld_cg(const T* p) { return *p; }
#else
template <class T> __device__ __forceinline__ T ld_cg(const T* p) { return __ldcg(p); }
#endif

// MODE 0: the query fits one strip.  MODE 1: one warp walks the strips one after another, boundary
// columns through global memory.  MODE 2 (K2, long pairs): the 4 warps of a CTA work on adjacent strips
// of ONE pair at the same time, each ~64 rows behind its left neighbour - an anti-diagonal wavefront
// across the CTA - handing boundary columns over through shared-memory rings.  MODE 3 (strip dataflow): every
// (pair, strip) is its own warp task; strip s publishes its last column to global memory every 32 rows and strip
// s+1 - on any warp of any CTA - follows ~64 rows behind.  Tasks are claimed from an atomic counter in (pair, strip)
// order, so a strip's producer is always claimed earlier and is running: the spin-waits cannot deadlock.
// LEAN (single strip, short tasks: the half-warp kernels, where fill and drain are 10-30 % of a task): rows above the
// reference run on the padding class and leave the row-0 state untouched (no re-initialisation), the last-row values are
// kept in shared memory and reduced after the strip.  The other instantiations keep the plain form (re-initialise at row
// 0, reduce the last row in place): their tasks are hundreds of blocks long and ptxas schedules their FAST body - which
// runs at the 128-register cap - measurably better without the extra code (A/B on B200, C2: 142.5 vs 147.4 ms).
template <class V, int K, int MODE, bool LEAN = false>
struct Wave {
    typedef typename V::T T;
    enum { K4 = (K + 3) / 4, USE2 = (K == 2 || K == 6), E2 = (K + 1) / 2, STEPS = V::STEPS, NP = V::NPAIR, MULTI = (MODE != 0), CTA = (MODE == 2), FLOW = (MODE == 3), XR = 256 };
    static_assert(!LEAN || MODE == 0, "LEAN is a single-strip form");

    // ---- per-task / per-strip constants --------------------------------------------------
    int lane, M, Na, Nb, j0, strip, gep, g4, rebase_mask, smin_m1, z4;
    bool last_strip;
    unsigned four, one;                 // == 4 at run time; opaque to ptxas so acc*four+x stays an IMAD (FMA pipe)
    const uint4* prof_lane;        // prof + lane
    const uint8_t* cls;            // cls[i-1] = class of reference row i
    const uint8_t* clsl;           // cls - lane: clsl[t] is the class of row t - lane + 1 (one add per block, not per step)
    int2* ring;
    int slot_off, warp_bytes;      // LEAN, CTA-uniform: byte offset of warp 0's slot area in dynamic shared memory, bytes per warp (the
                                   // slot address is recomputed where it is needed instead of living in a register)
    const int2* bnd_in;
    int2* bnd_out;
    // MODE 2: rings of XR rows in shared memory + published/consumed row sequence numbers
    const int2* xin;
    int2* xout;
    volatile long long* pub_in;    // producer of my left boundary: rows published
    volatile long long* cons_in;   //   ... and where I tell it how far I have consumed
    volatile long long* pub_out;   // my own ring
    volatile long long* cons_out;
    long long seq_in, seq_out;     // round * M of the producer's / my current strip
    // The hand-over from the LAST warp of a round to warp 0 of the next round cannot be a ring: warp 0
    // only starts its next strip after finishing the current one, a whole column later.  That link is a
    // full-length column in global memory (pub only, no back-pressure: warp 0 is always ahead of it).
    int2* col;
    bool in_col, out_col;
    // MODE 3: the boundary column is self-validating - the host fills it with 0xff bytes (x = -1 is no stored S^, those
    // are multiples of 4), the producer overwrites a row with ONE 64-bit store, the consumer re-reads a row until it is
    // real.  No progress flag, no fence, and the 32 rows of the NEXT window are prefetched a window ahead, so the L2
    // round trip is off the critical path (it used to be paid twice every 32 steps: flag poll, then data).
    int2 nextb;
    T c_up, c_sl0, c_q0, c_g4, c_g4_lane0;
    unsigned keep, inj_s, inj_q;   // lane-0 injection of column 0
    T Uq[K];

    // ---- lane state ------------------------------------------------------------------------
    T S[K], P[K];
    T sendS, sendQ;                // what lane+1 consumes next step: S^(i, j0+K), Q^(i, j0+K)
    T Sd_in;                       // S^(i-1, j0): diagonal input of column j0+1
    T diag0;                       // lane 0 / strip 0: "S(i-1,0) = 0" (gotoh.cpp:292) in the stored frame
    T best;                        // last column: running max in the stored frame of the current row
    int best_i_a, best_i_b;        //              and the STEP t of its row i = t - lane (largest i wins ties, gotoh.cpp:406-410)
    int lr_best_a, lr_j_a, lr_best_b, lr_j_b;   // last row (largest j wins ties, gotoh.cpp:399-403)
    int next_cls;
    uint4 dwords;

    // column-0 injection constants of this lane for the strip it is about to sweep (multi-strip forms call it per strip)
    __device__ __forceinline__ void set_injection(bool first_strip) {
        const bool inj = (lane == 0) && first_strip;
        c_g4_lane0 = inj ? c_g4 : V::both(0);
        keep = one - (inj ? 1u : 0u);                 // opaque (one is a run-time 1) so x*keep stays an IMAD
        inj_s = inj ? V::raw(c_sl0) : 0u;
        inj_q = inj ? V::raw(c_q0) : 0u;
    }
    // this lane's first slot word (stride 32 words): see FwdSmem::SLOT_BYTES
    __device__ __forceinline__ unsigned* slot_ptr() const {
        GOTOH_DYN_SMEM(smem_all);
        return reinterpret_cast<unsigned*>(smem_all + slot_off + (int)(threadIdx.x >> 5) * warp_bytes) + (threadIdx.x & 31);
    }
    // the last-column tracker's seed (row 0)
    __device__ __forceinline__ T best_seed() const {
        // (Vec16: every stored value of an admitted pair is >= 0 after the shift, so a stored 0 carried to column N is below
        // all of them - real values win ties - and, unlike the plan-wide bound smin_m1, always fits 16 bits)
        if (NP == 2) return V::pack(Na * g4, Nb * g4);
        return V::pack(4 * smin_m1 + Na * g4, 4 * smin_m1 + Nb * g4);
    }

    // S(0,j) = 0 and P(0,j) = 0 (gotoh.cpp:267-272) in the frame: 4*j*g (+1 tag for P); padding
    // columns (j > N) clone column N so that S[K-1] of the owner lane always reads S^(i,N).
    __device__ __forceinline__ void row0_init() {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int ja = min(j0 + k + 1, Na), jb = min(j0 + k + 1, Nb);
            S[k] = V::pack(ja * g4 + z4, jb * g4 + z4);
            P[k] = V::pack(ja * g4 + 1 + z4, jb * g4 + 1 + z4);
        }
        Sd_in = V::pack(min(j0, Na) * g4 + z4, min(j0, Nb) * g4 + z4);   // S^(0, j0)
        diag0 = lane == 0 ? V::both(z4) : V::both(0);           // 4*(i-1)*g at i = 1 (stays 0 outside lane 0)
        best = best_seed();
        best_i_a = best_i_b = lane;                              // row 0
    }

    // One lane-step: K cells of row i = t - lane.  SLOW adds the rarely needed per-lane checks
    // (row 0 re-init, int16 rebase, row M capture, i <= M guard); FAST blocks are proven by the
    // caller to need none of them.
    template <bool SLOW>
    __device__ __forceinline__ void step(const int t, const int s) {
        const int i = t - lane;
        const int my_cls = next_cls;
        next_cls = clsl[t];                                 // class of row i+1 (padded on both sides)

        // ---- hand-off from the left neighbour (its row i, produced last step) ---------------
        T Sl = __shfl_up_sync(0xffffffffu, sendS, 1);
        T Ql = __shfl_up_sync(0xffffffffu, sendQ, 1);
        T sdiag = Sd_in;
        if (CTA && strip > 0) {
            // lane 0 is at row t; the producer warp published it (block() waited for that)
            if (lane == 0) {
                const long long raw = in_col ? *reinterpret_cast<const volatile long long*>(&col[min(t, M)])
                                             : *reinterpret_cast<const volatile long long*>(&xin[t & (XR - 1)]);
                Sl = (T)(int)(raw & 0xffffffffLL); Ql = (T)(int)(raw >> 32);
            }
        } else {
            if (MULTI && !CTA) {
                // boundary column written by the previous strip; staged 32 rows at a time
                if (strip > 0 && ((t - 1) & 31) == 0) {
                    __syncwarp();
                    const int row = t + lane;                   // lane 0 is at row t+l at step t+l
                    int2 b = make_int2(0, 0);
                    if (FLOW) {
                        b = nextb;
                        const bool real_row = (row >= 1 && row <= M);
                        for (;;) {
                            const bool miss = real_row && b.x == -1;
                            if (!__any_sync(0xffffffffu, miss)) break;
                            if (miss) b = ld_cg(&bnd_in[row]);
                            gotoh_pause();
                        }
                        const int row2 = row + 32;              // next window, consumed 32 steps from now
                        nextb = (row2 >= 1 && row2 <= M) ? ld_cg(&bnd_in[row2]) : make_int2(0, 0);
                    } else if (row >= 1 && row <= M) b = bnd_in[row];
                    ring[(((t - 1) >> 5) & 1) * 32 + lane] = b;
                    __syncwarp();
                }
                // lane 0 takes its left neighbour from the ring - in every strip, so that the steps carry no branch on the
                // strip number: in strip 0 what it reads is overridden by the column-0 injection below (keep = 0)
                if (lane == 0) {
                    const int2 b = ring[(((t - 1) >> 5) & 1) * 32 + ((t - 1) & 31)];
                    Sl = (T)b.x; Ql = (T)b.y;
                }
            }
            // column 0 (gotoh.cpp:290-293) is injected into lane 0 of strip 0.  Written as x*keep + inj (keep = 0 for that
            // lane, 1 elsewhere; inj = 0 elsewhere) so that it issues as IMAD on the FMA pipe instead of three selects on
            // the saturated ALU pipe; in the later strips of a multi-strip pair keep = 1 and inj = 0 for every lane.
#ifdef GOTOH_LANE0_SEL
            if (lane == 0 && strip == 0) { Sl = c_sl0; Ql = c_q0; sdiag = diag0; }
#else
            Sl = (T)(V::raw(Sl) * keep + inj_s);
            Ql = (T)(V::raw(Ql) * keep + inj_q);
            sdiag = (T)(V::raw(sdiag) * keep + V::raw(diag0));
#endif
        }

        // ---- int16 range control: rebase every R rows (Vec16 only) --------------------------
        if (SLOW && NP == 2) {
            if (i > 0 && (i & rebase_mask) == 0) {
                const T d = V::both(-(rebase_mask + 1) * g4);
#pragma unroll
                for (int k = 0; k < K; ++k) { S[k] = V::add(S[k], d); P[k] = V::add(P[k], d); }
                sdiag = V::add(sdiag, d);
                best = V::add(best, d);
                if (lane == 0) {
                    diag0 = V::add(diag0, d);
                    // Column 0 as the Q recurrence sees it (gotoh.cpp:290-293: s = t_i = u + i*v, q = t_i + u) is
                    // 4(u - base*g) resp. 4(2u - base*g) + 2 in this frame: it moves with the base like every other stored
                    // value.  (Round 1 kept the injected constants fixed, which is only right until the first rebase row: with a
                    // small gip the stale, too high seed could win column 1 against a mismatch right after a rebase row - found
                    // by tools/fuzz_emu.py, gip = 0 / gep = 10.)  The seeds are floored where the add of the next cell cannot
                    // leave 16 bits; a candidate that low never wins: every real stored value is >= 0 in the shifted frame.
                    // (32-bit arithmetic per half: a seed that sits at its floor must not wrap when the next delta is added)
                    const int u4 = V::lo(c_sl0) - z4, dv = -(rebase_mask + 1) * g4;
                    const int fs = -32000 - u4, fq = -32000;
                    const T ns = V::pack(max(V::lo((T)inj_s) + dv, fs), max(V::hi((T)inj_s) + dv, fs));
                    const T nq = V::pack(max(V::lo((T)inj_q) + dv, fq), max(V::hi((T)inj_q) + dv, fq));
                    inj_s = V::raw(ns); inj_q = V::raw(nq);
                    Sl = ns; Ql = nq;        // this row's injection was made in the old frame, a few lines up
                }
            }
        }
        Sd_in = Sl;

        // ---- K cells of row i ------------------------------------------------------------------
        const uint4* prow = prof_lane + my_cls * (K4 * 32);
        const uint2* prow2 = reinterpret_cast<const uint2*>(prof_lane) + my_cls * (E2 * 32);   // K = 2, 6: rows of uint2
        T sleft = Sl, q = Ql;
        unsigned accC = 0, accS = 0;
#pragma unroll
        for (int kq = 0; kq < K4; ++kq) {
            unsigned ev[4] = {0u, 0u, 0u, 0u};
            if (USE2) {
                const uint2 a2 = prow2[(2 * kq) * 32];
                ev[0] = a2.x; ev[1] = a2.y;
                if (2 * kq + 1 < E2) { const uint2 b2 = prow2[(2 * kq + 1) * 32]; ev[2] = b2.x; ev[3] = b2.y; }
            } else {
                const uint4 e4 = prow[kq * 32];
                ev[0] = e4.x; ev[1] = e4.y; ev[2] = e4.z; ev[3] = e4.w;
            }
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const int k = kq * 4 + kk;
                if (k < K) {
                    q = V::addmax(sleft, Uq[k], q);            // Q^ (gotoh.cpp:305-308)
                    const T pp = V::addmax(S[k], c_up, P[k]);  // P^ (gotoh.cpp:311-314)
                    const T d = V::addlin(sdiag, one, (T)ev[kk]);   // D^ (gotoh.cpp:319)
                    const T C = V::max3(d, pp, q);             // select + tie-break (gotoh.cpp:362-395)
                    sdiag = S[k];
                    P[k] = pp;
                    S[k] = V::clr(C);
                    sleft = S[k];
                    accC = accC * four + V::raw(C);
                    accS = accS * four + V::raw(S[k]);
                }
            }
        }
        sendS = S[K - 1];
        sendQ = q;
        diag0 = V::addlin(diag0, one, c_g4_lane0);   // >= 0 in the shifted frame: an IMAD (c_g4_lane0 = 0 outside lane 0 of strip 0)

        // ---- directions: 2K bits per alignment per lane-step -------------------------------------
        const unsigned dstep = accC - accS;
        if (NP == 2) {
            if (s == 0) dwords.x = dstep; else if (s == 1) dwords.y = dstep;
            else if (s == 2) dwords.z = dstep; else dwords.w = dstep;
        } else {
            const unsigned sh = dstep << (16 * (s & 1));
            if ((s >> 1) == 0) dwords.x = (s & 1) ? (dwords.x | sh) : sh;
            else if ((s >> 1) == 1) dwords.y = (s & 1) ? (dwords.y | sh) : sh;
            else if ((s >> 1) == 2) dwords.z = (s & 1) ? (dwords.z | sh) : sh;
            else dwords.w = (s & 1) ? (dwords.w | sh) : sh;
        }

        // ---- last column: running max, ties -> larger i (gotoh.cpp:406-410) -----------------------
        // Tracked in the stored frame of the current row: carrying a score one row down adds 4g.
        if (!MULTI || last_strip) {
            best = V::addlin(best, one, c_g4);                 // best >= 0 (shifted frame): FMA pipe instead of VIADD.16x2
            bool pa, pb;
            const T nb = V::bmax(S[K - 1], best, &pa, &pb);
            if (!SLOW || i <= M) {
                best = nb;
                if (pa) best_i_a = t;
                if (NP == 2) { if (pb) best_i_b = t; }
            }
        }
        // ---- boundary column for the next strip --------------------------------------------------
        // (FAST blocks: every lane is on a real row, no range test)
        if (MULTI && !last_strip && lane == 31 && (!SLOW || (i >= 1 && i <= M))) {
            if (CTA) {
                const long long v = (long long)(((unsigned long long)V::raw(q) << 32) | V::raw(S[K - 1]));
                if (out_col) *reinterpret_cast<volatile long long*>(&col[i]) = v;
                else *reinterpret_cast<volatile long long*>(&xout[i & (XR - 1)]) = v;
            }
            else bnd_out[i] = make_int2((int)V::raw(S[K - 1]), (int)V::raw(q));
        }

        if (SLOW) {
            if (LEAN) {
                // Row 0.  While the wavefront fills, a lane runs on the rows above the reference: padding class 0, whose profile
                // entries are 0 (real columns) / 4u (padding columns).  That leaves S^ and P^ exactly at their row-0 values:
                // P^' = max(S^ + 4u+1, P^) = P^ (P^ = S^+1), D^ = S^(j-1) + 0 and Q^ <= S^(j-1) + 4u+2 are both below P^ (or,
                // with gip = gep = 0, clear to the same S^), so C = P^ and S^' = clr(C) = S^; the left neighbour is above the
                // reference too and hands over the same constants (its first hand-over is preset to them).  Only the
                // last-column tracker ran ahead: it gets its seed back.
                const bool z = (i == 0);
                best = z ? best_seed() : best;
                best_i_a = z ? t : best_i_a;
                if (NP == 2) best_i_b = z ? t : best_i_b;
                // row M: keep the lane's last-row values; lastrow_finish() reduces them after the strip (gotoh.cpp:399-403)
                if (i == M) {
                    unsigned* slot = slot_ptr();
#pragma unroll
                    for (int k = 0; k < K; ++k) slot[k * 32] = V::raw(S[k]);
                }
            } else {
                // row 0: re-initialise the lane just before its first real row
                if (i == 0) row0_init();
                // row M: last-row maximum over this lane's real columns
                if (i == M) {
                    const int roff = (NP == 2) ? (M & rebase_mask) : M;   // i - base(i)
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const int j = j0 + k + 1;
                        const int sa = ((V::lo(S[k]) - z4) >> 2) - (roff + j) * gep;
                        if (j <= Na && sa >= lr_best_a) { lr_best_a = sa; lr_j_a = j; }
                        if (NP == 2) {
                            const int sb = ((V::hi(S[k]) - z4) >> 2) - (roff + j) * gep;
                            if (j <= Nb && sb >= lr_best_b) { lr_best_b = sb; lr_j_b = j; }
                        }
                    }
                }
            }
        }
    }

    // Last-row maximum over this lane's real columns of the strip, from the values kept at row M (larger j wins ties).
    // Uniform code, once per strip (it used to run inside the divergent `i == M` branch of 16-32 consecutive steps).
    __device__ __forceinline__ void lastrow_finish() {
        if (!LEAN) return;
        const int roff = (NP == 2) ? (M & rebase_mask) : M;   // i - base(i) at i = M
        const unsigned* slot = slot_ptr();
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const T sk = (T)slot[k * 32];
            const int j = j0 + k + 1;
            const int sa = ((V::lo(sk) - z4) >> 2) - (roff + j) * gep;
            if (j <= Na && sa >= lr_best_a) { lr_best_a = sa; lr_j_a = j; }
            if (NP == 2) {
                const int sb = ((V::hi(sk) - z4) >> 2) - (roff + j) * gep;
                if (j <= Nb && sb >= lr_best_b) { lr_best_b = sb; lr_j_b = j; }
            }
        }
    }

    template <bool SLOW>
    __device__ __forceinline__ void block(const int tb, uint4* dst) {
        const int t0 = tb * STEPS + 1, hi = t0 + STEPS - 1;
        if (CTA) {
            if (strip > 0 && ((t0 - 1) & 31) == 0) {
                // lane 0 consumes rows t0 .. t0+31 during the next 32 steps: wait until they are published,
                // and tell the producer that everything before t0 has been consumed
                if (lane == 0) {
                    const long long need = seq_in + min(t0 + 31, M);
                    while (*pub_in < need) gotoh_pause();
                    *cons_in = seq_in + (t0 - 1);
                }
                __syncwarp();
            }
            if (!last_strip && !out_col && hi - 31 >= 1) {
                // lane 31 writes rows up to hi-31 in this block; their ring slots held rows XR earlier
                // (or the previous round's strip, which must be consumed completely)
                if (lane == 31) {
                    const long long must = seq_out + max(0, min(hi - 31, M) - XR);
                    while (*cons_out < must) gotoh_pause();
                }
                __syncwarp();
            }
        }
#pragma unroll
        for (int s = 0; s < STEPS; ++s) step<SLOW>(tb * STEPS + s + 1, s);
        *dst = dwords;   // one coalesced 512-byte store per warp per STEPS lane-steps
        if (CTA && !last_strip) {
            __syncwarp();
            if (lane == 31) {
                __threadfence_block();
                *pub_out = seq_out + min(max(hi - 31, 0), M);
            }
        }
    }
};

// HALF (Vec16, single strip, queries of at most 16*K columns): the warp carries TWO independent wavefronts of 16 lanes,
// lanes 0-15 one couple of alignments and lanes 16-31 another couple against the same reference (same M, so the block
// loop stays warp-uniform).  Short queries (84-aa windows, K = 6) then pay half the wavefront fill/drain (15 rows, not 31)
// and amortise the per-step overhead over twice the columns per lane.  Tasks come in twos: entry 2*tsk + (lane >> 4).
// Both couples share one arena slab: physical lane = wave lane + 16 * (second couple), see dir_addr().  The hand-over
// shuffle needs no width: wave lane 0 of either half overrides what it receives with the column-0 injection.
template <class V, int K, bool MULTI, bool HALF = false>
__global__ void __launch_bounds__(FWD_WARPS * 32, GOTOH_MIN_CTAS) k_forward(const FwdParams p) {
    static_assert(!HALF || (V::NPAIR == 2 && !MULTI), "HALF is a single-strip int16x2 mode");
    typedef typename V::T T;
    typedef Wave<V, K, MULTI ? 1 : 0, HALF> W;
    typedef FwdSmem<V, K, HALF> SM;
    enum { K4 = W::K4, STEPS = V::STEPS, NP = V::NPAIR, FILL = HALF ? 15 : 31 };
    GOTOH_DYN_SMEM(smem_raw);
    const int warp = threadIdx.x >> 5, plane = threadIdx.x & 31;
    const int lane = HALF ? (plane & 15) : plane;       // lane within its wavefront
    const int sub = HALF ? (plane >> 4) : 0;
    unsigned char* my_smem = smem_raw + p.stab_bytes + (size_t)warp * SM::per_warp(p.ncls);
    uint4* prof = reinterpret_cast<uint4*>(my_smem);
    // half-warp kernels rebuild the query profile every 100-440 rows: the score table comes from a CTA-wide int16 copy in
    // shared memory (32-bit addresses with immediate offsets instead of a 64-bit address per global load)
    const short* stab = reinterpret_cast<const short*>(smem_raw);
    if (HALF && p.stab_bytes) {
        short* st = reinterpret_cast<short*>(smem_raw);
        for (int x = threadIdx.x; x < p.ncls * 128; x += blockDim.x) st[x] = (short)p.table4[x];
        __syncthreads();
    }

    W w;
    w.lane = lane;
    w.gep = p.gep;
    w.g4 = 4 * p.gep;
    w.rebase_mask = p.rebase_mask;
    w.smin_m1 = p.smin_m1;
    w.four = p.four;
    w.one = p.four >> 2;
    w.z4 = (NP == 2) ? p.zshift : 0;
    w.prof_lane = W::USE2 ? reinterpret_cast<const uint4*>(reinterpret_cast<const uint2*>(prof) + plane) : prof + plane;
    w.ring = reinterpret_cast<int2*>(my_smem + SM::prof_bytes(p.ncls));  // [2][32]
    w.slot_off = p.stab_bytes + (int)(SM::prof_bytes(p.ncls) + 2 * 32 * sizeof(int2));
    w.warp_bytes = (int)SM::per_warp(p.ncls);
    const int u4 = -4 * p.gip;
    w.c_up = V::both(u4 + 1);          // P^ = max(S^up + 4u+1, P^up)
    w.c_sl0 = V::both(u4 + w.z4);          // column 0 seen by the Q recurrence: s~ = u   (gotoh.cpp:291)
    w.c_q0 = V::both(2 * u4 + 2 + w.z4);   //                                     q~ = 2u  (gotoh.cpp:293)
    w.c_g4 = V::both(w.g4);
    w.c_g4_lane0 = lane == 0 ? w.c_g4 : V::both(0);
    w.keep = (p.four >> 2) - (lane == 0 ? 1u : 0u);   // 0 for lane 0, 1 elsewhere; opaque so x*keep stays an IMAD
    w.inj_s = lane == 0 ? V::raw(w.c_sl0) : 0u;
    w.inj_q = lane == 0 ? V::raw(w.c_q0) : 0u;

    for (int done = 0;; ++done) {
        // ---- dynamic task fetch (one atomic per warp) ---------------------------------------
        if (p.task_limit > 0 && done >= p.task_limit) break;
        unsigned tsk = 0;
        if (plane == 0) tsk = atomicAdd(p.work_counter, 1u);
        tsk = __shfl_sync(0xffffffffu, tsk, 0);
        if (tsk >= (unsigned)p.task_count) break;
        const Task task = p.tasks[p.task_first + (HALF ? 2 * tsk + sub : tsk)];
        const PairInfo pa = p.pairs[task.pair_a];
        const PairInfo pb = p.pairs[task.pair_b >= 0 ? task.pair_b : task.pair_a];
        const int M = pa.M;                       // both halves share the reference
        const int Na = pa.N, Nb = (NP == 2) ? pb.N : pa.N;
        const int Nmax = Na > Nb ? Na : Nb;
        const int nstrips = MULTI ? (Nmax + 32 * K - 1) / (32 * K) : 1;
        const int nblk = pa.nblk;
        const uint8_t* qa = p.qry + pa.qry_pos;
        const uint8_t* qb = p.qry + pb.qry_pos;
        int2* bnd0 = MULTI ? p.bnd + ((int64_t)(blockIdx.x * (blockDim.x >> 5) + warp) * 2) * p.bnd_stride : nullptr;
        w.M = M; w.Na = Na; w.Nb = Nb;
        w.cls = p.ref_cls + pa.ref_pos;
        w.clsl = w.cls - lane;
        w.lr_best_a = w.lr_best_b = -2147483647;
        w.lr_j_a = w.lr_j_b = 0;

        for (int strip = 0; strip < nstrips; ++strip) {
            const int j0 = (strip * 32 + lane) * K;           // this lane owns columns j0+1 .. j0+K (HALF: strip = 0)
            w.strip = strip;
            w.j0 = j0;
            w.last_strip = (strip == nstrips - 1);
            w.set_injection(strip == 0);      // (per task: the int16x2 rebase rows move the injected column-0 seeds)
            if (MULTI) {
                w.bnd_in = bnd0 + (int64_t)((strip + 1) & 1) * p.bnd_stride;
                w.bnd_out = bnd0 + (int64_t)(strip & 1) * p.bnd_stride;
            }

            // ---- query profile for this strip: prof[c][k4][lane].{x,y,z,w} ----------------------
            __syncwarp();
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const bool real_a = (j0 + k) < Na, real_b = (j0 + k) < Nb;
                w.Uq[k] = V::pack(real_a ? u4 + 2 : 3, real_b ? u4 + 2 : 3);
            }
            // stop-codon masks of this lane's columns (only when some reference holds "$$$"):
            // bit0 b[j-3..j-1], bit1 b[j-2..j], bit2 b[j-1..j+1] in {TAG,TAA,TGA}, j >= 3 (gotoh.cpp:324-344)
            int cm_a[K], cm_b[K];
#pragma unroll
            for (int k = 0; k < K; ++k) { cm_a[k] = 0; cm_b[k] = 0; }
            if (p.has_dollar) {
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    cm_a[k] = stop_mask(qa, Na, j0 + k + 1);
                    if (NP == 2) cm_b[k] = stop_mask(qb, Nb, j0 + k + 1);
                }
            }
            // the query bytes of this lane's columns, read once (-1: padding column)
            int qca[K], qcb[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int ja = j0 + k;
                qca[k] = ja < Na ? (int)qa[ja] : -1;
                qcb[k] = (NP == 2 && ja < Nb) ? (int)qb[ja] : -1;
            }
            if (HALF && p.stab_bytes) {
                // Column-major build from the shared-memory table (short queries rebuild the profile every 100-440 rows, so
                // its cost counts: it was 8 % of the C3 kernel's instructions): per column two 32-bit table offsets, per
                // class two 16-bit loads, one multiply-add and one 32-bit store - no per-entry conditions, because padding
                // columns read entry 0 of the class row, which the host sets to 4u (the value that makes them clone
                // column N).
                unsigned* profw = reinterpret_cast<unsigned*>(prof);
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const short* ta = stab + (qca[k] >= 0 ? qca[k] : 0);
                    const short* tb = stab + (qcb[k] >= 0 ? qcb[k] : 0);
                    unsigned* dw = profw + (W::USE2 ? ((k >> 1) * 32 + plane) * 2 + (k & 1) : ((k >> 2) * 32 + plane) * 4 + (k & 3));
                    const int cstride = W::USE2 ? W::E2 * 64 : K4 * 128;        // 32-bit words per class row
#pragma unroll 7
                    for (int c = 0; c < p.ncls; ++c) dw[c * cstride] = V::lin((int)ta[c * 128], (int)tb[c * 128]);
                }
            } else
            for (int c = 0; c < p.ncls; ++c) {
                const int32_t* trow = p.table4 + c * 128;
                const int32_t* brow = p.bonus4 + c * 8;
                unsigned e[K4 * 4];
#pragma unroll
                for (int k = 0; k < K4 * 4; ++k) {
                    // padding columns (beyond N) use E = 4u so that they clone column N
                    int ea = u4, eb = u4;
                    if (k < K) {
                        const int ja = j0 + k;
                        if (HALF) {
                            if (qca[k] >= 0) ea = trow[qca[k]] + (p.has_dollar ? brow[cm_a[k]] : 0);
                            if (NP == 2 && qcb[k] >= 0) eb = trow[qcb[k]] + (p.has_dollar ? brow[cm_b[k]] : 0);
                        } else {
                            if (ja < Na) ea = trow[qa[ja]] + (p.has_dollar ? brow[cm_a[k]] : 0);
                            if (NP == 2 && ja < Nb) eb = trow[qb[ja]] + (p.has_dollar ? brow[cm_b[k]] : 0);
                        }
                    }
                    e[k] = V::lin(ea, NP == 2 ? eb : 0);
                }
                if (W::USE2) {
                    uint2* prof2 = reinterpret_cast<uint2*>(prof);
#pragma unroll
                    for (int kp = 0; kp < W::E2; ++kp) prof2[(c * W::E2 + kp) * 32 + plane] = make_uint2(e[2 * kp], e[2 * kp + 1]);
                } else {
#pragma unroll
                    for (int kq = 0; kq < K4; ++kq) prof[(c * K4 + kq) * 32 + plane] = make_uint4(e[4 * kq], e[4 * kq + 1], e[4 * kq + 2], e[4 * kq + 3]);
                }
            }
            __syncwarp();

            w.dwords = make_uint4(0, 0, 0, 0);
            w.row0_init();                              // lane 0 starts at row 1 in the very first step
            // what the right neighbour consumes in the first step: it is above the reference then and must see this lane's
            // row-0 constants (see step(): rows above the reference leave the row-0 state untouched)
            w.sendS = HALF ? w.S[K - 1] : V::both(0);
            w.sendQ = HALF ? V::add(w.S[K - 1], V::both(u4 + 2)) : V::both(0);
            w.next_cls = w.cls[-lane];                  // class of row i = 1 - lane (padding when i < 1)

            // Which per-lane events can occur in the steps t0 .. hi of a block (rows t-FILL .. t)?  Row 0 (re-init) in the
            // blocks with t0 <= FILL, row M (capture, i <= M guard) in those with hi >= M, a rebase row (int16x2) in those
            // that meet [m, m+FILL] for a multiple m of R: these run the SLOW body.
            uint4* dst = p.dir + pa.dir_off + (int64_t)strip * nblk * 32 + plane;
            int tb = 0;
            if (!HALF) {
                // 32-lane wavefronts: one loop with a per-block test (their tasks are hundreds of blocks long)
                for (; tb < nblk; ++tb, dst += 32) {
                    const int t0 = tb * STEPS + 1, hi = t0 + STEPS - 1;
                    bool slow = (t0 <= FILL) || (hi >= M);                    // row 0 re-init / row M capture, i <= M
                    if (NP == 2) {
                        const int m = hi & ~p.rebase_mask;                  // largest multiple of R that is <= hi
                        slow = slow || (m > 0 && m >= t0 - FILL);             // some lane crosses a rebase row
                    }
                    if (slow) w.template block<true>(tb, dst);
                    else w.template block<false>(tb, dst);
                }
            } else {
                // half-warp wavefronts (tasks of 25-115 blocks, 9 of them SLOW): the loop is cut into phases so that the FAST
                // blocks in between carry no classification at all
                const int head = min(nblk, (FILL - 1) / STEPS + 1);                 // blocks with t0 = tb*STEPS+1 <= FILL
                const int tail = max(head, min(nblk, (M + STEPS - 1) / STEPS - 1));  // first block with hi = (tb+1)*STEPS >= M
                for (; tb < head; ++tb, dst += 32) w.template block<true>(tb, dst);
                const int R = p.rebase_mask + 1;
                for (int m = R; tb < tail; m += R) {
                    const int rb0 = min(tail, max(tb, (m - 1) / STEPS));          // block that holds step m
                    const int rb1 = min(tail, max(rb0, (m + FILL - 1) / STEPS + 1));  // one past the block that holds step m+FILL
                    for (; tb < rb0; ++tb, dst += 32) w.template block<false>(tb, dst);
                    for (; tb < rb1; ++tb, dst += 32) w.template block<true>(tb, dst);
                }
                for (; tb < nblk; ++tb, dst += 32) w.template block<true>(tb, dst);
            }
            w.lastrow_finish();
        }

        // ---- K3 epilogue: end-cell choice (gotoh.cpp:418-450) -----------------------------
        // last row: reduce (score, j) with larger j winning ties
        int lr_best_a = w.lr_best_a, lr_j_a = w.lr_j_a, lr_best_b = w.lr_best_b, lr_j_b = w.lr_j_b;
#pragma unroll
        for (int off = HALF ? 8 : 16; off > 0; off >>= 1) {
            const int os = __shfl_xor_sync(0xffffffffu, lr_best_a, off);
            const int oj = __shfl_xor_sync(0xffffffffu, lr_j_a, off);
            if (os > lr_best_a || (os == lr_best_a && oj > lr_j_a)) { lr_best_a = os; lr_j_a = oj; }
            if (NP == 2) {
                const int os2 = __shfl_xor_sync(0xffffffffu, lr_best_b, off);
                const int oj2 = __shfl_xor_sync(0xffffffffu, lr_j_b, off);
                if (os2 > lr_best_b || (os2 == lr_best_b && oj2 > lr_j_b)) { lr_best_b = os2; lr_j_b = oj2; }
            }
        }
        // last column: owner lane's tracker, converted from the stored frame of the lane's final row
        {
            const int i_fin = nblk * STEPS - lane;          // > 0: nblk*STEPS >= M + 31
            const int roff = (NP == 2) ? (i_fin & p.rebase_mask) : i_fin;   // i_fin - base(i_fin)
            const int la = (Na - 1) / K - (nstrips - 1) * 32 + 16 * sub;   // owner lane within the last strip
            const int lb = (Nb - 1) / K + 16 * sub;
            int best_a = ((V::lo(w.best) - w.z4) >> 2) - (roff + Na) * p.gep;
            int best_b = ((V::hi(w.best) - w.z4) >> 2) - (roff + Nb) * p.gep;
            best_a = __shfl_sync(0xffffffffu, best_a, la);
            const int bi_a = __shfl_sync(0xffffffffu, w.best_i_a - lane, la);
            best_b = __shfl_sync(0xffffffffu, best_b, lb & 31);
            const int bi_b = __shfl_sync(0xffffffffu, w.best_i_b - lane, lb & 31);
            if (lane == 0) {
                // strict '>' : the last row only wins when it is strictly better (gotoh.cpp:429)
                if (lr_best_a > best_a) { p.score[task.pair_a] = lr_best_a; p.end_i[task.pair_a] = M; p.end_j[task.pair_a] = lr_j_a; }
                else { p.score[task.pair_a] = best_a; p.end_i[task.pair_a] = bi_a; p.end_j[task.pair_a] = Na; }
                if (NP == 2 && task.pair_b >= 0) {
                    if (lr_best_b > best_b) { p.score[task.pair_b] = lr_best_b; p.end_i[task.pair_b] = M; p.end_j[task.pair_b] = lr_j_b; }
                    else { p.score[task.pair_b] = best_b; p.end_i[task.pair_b] = bi_b; p.end_j[task.pair_b] = Nb; }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------
// K2  long pairs as strip dataflow: one warp per (pair, strip), strips of a pair pipelined across the whole grid
// ------------------------------------------------------------------------------------
enum { PART_EMPTY = (int)0x80808080 };   // cudaMemset(0x80) pattern: no alignment score (all > -100000, gotoh.cpp:284)

template <int K>
__global__ void __launch_bounds__(FWD_WARPS * 32, GOTOH_MIN_CTAS) k_forward_flow(const FwdParams p) {
    typedef Vec32 V;
    typedef Wave<V, K, 3> W;
    enum { K4 = W::K4, STEPS = V::STEPS };
    GOTOH_DYN_SMEM(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* my_smem = smem_raw + (size_t)warp * FwdSmem<V, K>::per_warp(p.ncls);
    uint4* prof = reinterpret_cast<uint4*>(my_smem);
    W w;
    w.lane = lane;
    w.gep = p.gep;
    w.g4 = 4 * p.gep;
    w.rebase_mask = p.rebase_mask;
    w.smin_m1 = p.smin_m1;
    w.four = p.four;
    w.one = p.four >> 2;
    w.z4 = 0;
    w.prof_lane = prof + lane;
    w.ring = reinterpret_cast<int2*>(my_smem + (size_t)p.ncls * K4 * 32 * 16);
    w.slot_off = 0; w.warp_bytes = 0;
    const int u4 = -4 * p.gip;
    w.c_up = V::both(u4 + 1);
    w.c_sl0 = V::both(u4);
    w.c_q0 = V::both(2 * u4 + 2);
    w.c_g4 = V::both(w.g4);
    w.c_g4_lane0 = lane == 0 ? w.c_g4 : V::both(0);
    w.keep = (p.four >> 2) - (lane == 0 ? 1u : 0u);
    w.inj_s = lane == 0 ? V::raw(w.c_sl0) : 0u;
    w.inj_q = lane == 0 ? V::raw(w.c_q0) : 0u;

    for (;;) {
        unsigned tsk = 0;
        if (lane == 0) tsk = atomicAdd(p.work_counter, 1u);
        tsk = __shfl_sync(0xffffffffu, tsk, 0);
        if (tsk >= (unsigned)p.task_count) break;
        const Task task = p.tasks[p.task_first + tsk];
        const PairInfo pa = p.pairs[task.pair_a];
        const int strip = task.pair_b;
        const int M = pa.M, Na = pa.N;
        const int nstrips = (Na + 32 * K - 1) / (32 * K);
        const int nblk = pa.nblk;
        const int slot = pa.pad1 + strip;
        const uint8_t* qa = p.qry + pa.qry_pos;
        const int j0 = (strip * 32 + lane) * K;
        w.M = M; w.Na = Na; w.Nb = Na;
        w.cls = p.ref_cls + pa.ref_pos;
        w.clsl = w.cls - lane;
        w.lr_best_a = w.lr_best_b = -2147483647;
        w.lr_j_a = w.lr_j_b = 0;
        w.strip = strip;
        w.j0 = j0;
        w.last_strip = (strip == nstrips - 1);
        w.bnd_in = p.bnd + (int64_t)(slot - 1) * p.bnd_stride;
        w.bnd_out = p.bnd + (int64_t)slot * p.bnd_stride;
        w.nextb = make_int2(-1, -1);                  // nothing prefetched yet: the first window is fetched on demand
        w.set_injection(strip == 0);

        __syncwarp();
#pragma unroll
        for (int k = 0; k < K; ++k) w.Uq[k] = V::both((j0 + k) < Na ? u4 + 2 : 3);
        int cm_a[K];
#pragma unroll
        for (int k = 0; k < K; ++k) cm_a[k] = p.has_dollar ? stop_mask(qa, Na, j0 + k + 1) : 0;
        for (int c = 0; c < p.ncls; ++c) {
            const int32_t* trow = p.table4 + c * 128;
            const int32_t* brow = p.bonus4 + c * 8;
#pragma unroll
            for (int kq = 0; kq < K4; ++kq) {
                unsigned e[4];
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const int k = kq * 4 + kk;
                    const int ja = j0 + k;
                    int ea = u4;
                    if (k < K && ja < Na) ea = trow[qa[ja]] + (p.has_dollar ? brow[cm_a[k]] : 0);
                    e[kk] = (unsigned)ea;
                }
                prof[(c * K4 + kq) * 32 + lane] = make_uint4(e[0], e[1], e[2], e[3]);
            }
        }
        __syncwarp();
        w.sendS = V::both(0);
        w.sendQ = V::both(0);
        w.dwords = make_uint4(0, 0, 0, 0);
        w.row0_init();
        w.next_cls = w.cls[-lane];

        uint4* dst = p.dir + pa.dir_off + (int64_t)strip * nblk * 32 + lane;
        for (int tb = 0; tb < nblk; ++tb, dst += 32) {
            const int t0 = tb * STEPS + 1, hi = t0 + STEPS - 1;
            const bool slow = (t0 <= 31) || (hi >= M);
            if (slow) w.template block<true>(tb, dst);
            else w.template block<false>(tb, dst);
        }
        w.lastrow_finish();
        // last row of this strip: (score, j) with larger j winning ties (gotoh.cpp:399-403)
        int lr_best = w.lr_best_a, lr_j = w.lr_j_a;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const int os = __shfl_xor_sync(0xffffffffu, lr_best, off);
            const int oj = __shfl_xor_sync(0xffffffffu, lr_j, off);
            if (os > lr_best || (os == lr_best && oj > lr_j)) { lr_best = os; lr_j = oj; }
        }
        if (!w.last_strip) {
            // partial of this strip: column first, then (fenced) the score, which the host preset to PART_EMPTY
            if (lane == 0) { p.part_j[slot] = lr_j; __threadfence(); st_volatile(&p.part_best[slot], lr_best); }
            continue;
        }
        // the last strip finishes last: fold the earlier strips' partials in (they lie to the left, so on ties
        // the later strip wins, as `>=` does in the reference's left-to-right scan)
        if (lane == 0) {
            int best = -2147483647, bj = 0;
            for (int s = 0; s < nstrips - 1; ++s) {
                // an earlier strip has handed over its last boundary row but may still be writing its partial
                int os;
                while ((os = ld_volatile(&p.part_best[pa.pad1 + s])) == PART_EMPTY) gotoh_pause();
                __threadfence();
                const int oj = ld_cg(&p.part_j[pa.pad1 + s]);
                if (os >= best) { best = os; bj = oj; }
            }
            if (lr_best >= best) { best = lr_best; bj = lr_j; }
            lr_best = best; lr_j = bj;
        }
        {
            const int i_fin = nblk * STEPS - lane;
            const int la = (Na - 1) / K - (nstrips - 1) * 32;
            int best_a = (V::lo(w.best) >> 2) - (i_fin + Na) * p.gep;
            best_a = __shfl_sync(0xffffffffu, best_a, la);
            const int bi_a = __shfl_sync(0xffffffffu, w.best_i_a - lane, la);
            if (lane == 0) {
                if (lr_best > best_a) { p.score[task.pair_a] = lr_best; p.end_i[task.pair_a] = M; p.end_j[task.pair_a] = lr_j; }
                else { p.score[task.pair_a] = best_a; p.end_i[task.pair_a] = bi_a; p.end_j[task.pair_a] = Na; }
            }
        }
    }
}

// ------------------------------------------------------------------------------------
// K2  long pairs: one CTA per pair, anti-diagonal wavefront across its 4 warps (int32 cells)
// ------------------------------------------------------------------------------------
// Warp w takes strips w, w+4, w+8, ... of 256 query columns.  Strip s+1 can start as soon as strip s
// has published its first 32 boundary rows, so the four warps run ~64 rows apart; boundary columns
// never touch global memory.  Same cell code, same direction layout, same epilogue as k_forward.
template <int K>
__global__ void __launch_bounds__(FWD_WARPS * 32, GOTOH_MIN_CTAS) k_forward_cta(const FwdParams p) {
    typedef Vec32 V;
    typedef Wave<V, K, 2> W;
    enum { K4 = W::K4, STEPS = V::STEPS, XR = W::XR };
    GOTOH_DYN_SMEM(smem_raw);
    __shared__ long long sh_pub[FWD_WARPS], sh_cons[FWD_WARPS];
    __shared__ int sh_task;
    __shared__ int sh_lr_best[FWD_WARPS], sh_lr_j[FWD_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t prof_bytes = (size_t)p.ncls * K4 * 32 * 16;
    const size_t warp_bytes = prof_bytes + XR * sizeof(int2);
    unsigned char* my_smem = smem_raw + (size_t)warp * warp_bytes;
    uint4* prof = reinterpret_cast<uint4*>(my_smem);

    W w;
    w.lane = lane;
    w.gep = p.gep;
    w.g4 = 4 * p.gep;
    w.rebase_mask = p.rebase_mask;
    w.smin_m1 = p.smin_m1;
    w.four = p.four;
    w.one = p.four >> 2;
    w.z4 = 0;
    w.prof_lane = prof + lane;
    w.ring = nullptr;
    w.bnd_in = nullptr;
    w.bnd_out = nullptr;
    const int u4 = -4 * p.gip;
    w.c_up = V::both(u4 + 1);
    w.c_sl0 = V::both(u4);
    w.c_q0 = V::both(2 * u4 + 2);
    w.c_g4 = V::both(w.g4);
    w.c_g4_lane0 = lane == 0 ? w.c_g4 : V::both(0);
    w.keep = (p.four >> 2) - (lane == 0 ? 1u : 0u);
    w.inj_s = lane == 0 ? V::raw(w.c_sl0) : 0u;
    w.inj_q = lane == 0 ? V::raw(w.c_q0) : 0u;
    const int pw = (warp + FWD_WARPS - 1) % FWD_WARPS;      // the warp that produces my left boundary
    w.xout = reinterpret_cast<int2*>(my_smem + prof_bytes);
    w.xin = reinterpret_cast<const int2*>(smem_raw + (size_t)pw * warp_bytes + prof_bytes);
    w.slot_off = 0; w.warp_bytes = 0;
    w.pub_out = &sh_pub[warp];
    w.cons_out = &sh_cons[warp];
    w.pub_in = &sh_pub[pw];
    w.cons_in = &sh_cons[pw];
    w.col = p.bnd + (int64_t)blockIdx.x * 2 * p.bnd_stride;
    w.in_col = (warp == 0);                      // strips 4r (r > 0) read the column written by strip 4r-1
    w.out_col = (warp == FWD_WARPS - 1);         // strips 4r+3 write it

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) sh_task = (int)atomicAdd(p.work_counter, 1u);
        if (threadIdx.x < FWD_WARPS) { sh_pub[threadIdx.x] = 0; sh_cons[threadIdx.x] = 0; }
        __syncthreads();
        const int tsk = sh_task;
        if (tsk >= p.task_count) break;
        const Task task = p.tasks[p.task_first + tsk];
        const PairInfo pa = p.pairs[task.pair_a];
        const int M = pa.M, Na = pa.N;
        const int nstrips = (Na + 32 * K - 1) / (32 * K);
        const int nblk = pa.nblk;
        const uint8_t* qa = p.qry + pa.qry_pos;
        w.M = M; w.Na = Na; w.Nb = Na;
        w.cls = p.ref_cls + pa.ref_pos;
        w.clsl = w.cls - lane;
        w.lr_best_a = w.lr_best_b = -2147483647;
        w.lr_j_a = w.lr_j_b = 0;
        w.best = V::both(0);
        w.best_i_a = w.best_i_b = lane;

        for (int strip = warp; strip < nstrips; strip += FWD_WARPS) {
            const int j0 = (strip * 32 + lane) * K;
            w.strip = strip;
            w.j0 = j0;
            w.last_strip = (strip == nstrips - 1);
            w.seq_out = (long long)(strip / FWD_WARPS) * M;
            w.seq_in = (long long)((strip - 1) / FWD_WARPS) * M;      // unused for strip 0

            __syncwarp();
#pragma unroll
            for (int k = 0; k < K; ++k) w.Uq[k] = V::both((j0 + k) < Na ? u4 + 2 : 3);
            int cm_a[K];
#pragma unroll
            for (int k = 0; k < K; ++k) cm_a[k] = p.has_dollar ? stop_mask(qa, Na, j0 + k + 1) : 0;
            for (int c = 0; c < p.ncls; ++c) {
                const int32_t* trow = p.table4 + c * 128;
                const int32_t* brow = p.bonus4 + c * 8;
#pragma unroll
                for (int kq = 0; kq < K4; ++kq) {
                    unsigned e[4];
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const int k = kq * 4 + kk;
                        const int ja = j0 + k;
                        int ea = u4;
                        if (k < K && ja < Na) ea = trow[qa[ja]] + (p.has_dollar ? brow[cm_a[k]] : 0);
                        e[kk] = (unsigned)ea;
                    }
                    prof[(c * K4 + kq) * 32 + lane] = make_uint4(e[0], e[1], e[2], e[3]);
                }
            }
            __syncwarp();

            w.sendS = V::both(0);
            w.sendQ = V::both(0);
            w.dwords = make_uint4(0, 0, 0, 0);
            w.row0_init();
            w.next_cls = w.cls[-lane];

            uint4* dst = p.dir + pa.dir_off + (int64_t)strip * nblk * 32 + lane;
            for (int tb = 0; tb < nblk; ++tb, dst += 32) {
                const int t0 = tb * STEPS + 1, hi = t0 + STEPS - 1;
                const bool slow = (t0 <= 31) || (hi >= M);
                if (slow) w.template block<true>(tb, dst);
                else w.template block<false>(tb, dst);
            }
            w.lastrow_finish();
            // everything of the left boundary has been consumed
            if (strip > 0 && lane == 0) *w.cons_in = w.seq_in + M;
            __syncwarp();
        }

        // ---- epilogue: last row over all strips (every warp contributes), last column from the
        //      warp that ran the last strip ------------------------------------------------------
        int lr_best = w.lr_best_a, lr_j = w.lr_j_a;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const int os = __shfl_xor_sync(0xffffffffu, lr_best, off);
            const int oj = __shfl_xor_sync(0xffffffffu, lr_j, off);
            if (os > lr_best || (os == lr_best && oj > lr_j)) { lr_best = os; lr_j = oj; }
        }
        if (lane == 0) { sh_lr_best[warp] = lr_best; sh_lr_j[warp] = lr_j; }
        __syncthreads();
        const int last_warp = (nstrips - 1) % FWD_WARPS;
        if (warp == last_warp) {
            for (int x = 0; x < FWD_WARPS; ++x) {
                const int os = sh_lr_best[x], oj = sh_lr_j[x];
                if (os > lr_best || (os == lr_best && oj > lr_j)) { lr_best = os; lr_j = oj; }
            }
            const int i_fin = nblk * STEPS - lane;
            const int la = (Na - 1) / K - (nstrips - 1) * 32;
            int best_a = (V::lo(w.best) >> 2) - (i_fin + Na) * p.gep;
            best_a = __shfl_sync(0xffffffffu, best_a, la);
            const int bi_a = __shfl_sync(0xffffffffu, w.best_i_a - lane, la);
            if (lane == 0) {
                if (lr_best > best_a) { p.score[task.pair_a] = lr_best; p.end_i[task.pair_a] = M; p.end_j[task.pair_a] = lr_j; }
                else { p.score[task.pair_a] = best_a; p.end_i[task.pair_a] = bi_a; p.end_j[task.pair_a] = Na; }
            }
        }
    }
}

// ------------------------------------------------------------------------------------
// K4a  traceback walk: one thread per pair (gotoh.cpp:452-510)
// ------------------------------------------------------------------------------------
struct WalkParams {
    const PairInfo* pairs;
    int32_t pair_first, pair_count;
    const uint32_t* dir;        // arena viewed as uint32
    const int32_t* end_i;
    const int32_t* end_j;
    int32_t* score;             // in: S at end cell; out: final score with terminal fix-up
    uint32_t* ops;              // op scripts, 2 bits per op, 16 per word, in traceback order
    int32_t* nops;
    int32_t* i0;
    int32_t* j0;
    int32_t* out_len;           // plan order
    int32_t gip, gep, term;
    // tight / compact result forms: what the caller-order prefix sum (k_scan) runs over - the pair's aligned length in
    // bytes (scan_words = 0) or its op-script length in 32-bit words (scan_words = 1); NULL for the strided form
    int32_t* scan_in;
    int32_t scan_words;
};

__global__ void __launch_bounds__(128) k_walk(const WalkParams p) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p.pair_count) return;
    const int pi = p.pair_first + idx;
    const PairInfo pr = p.pairs[pi];
    int i = p.end_i[pi], j = p.end_j[pi];
    const int right = (i == pr.M && j < pr.N) ? (pr.N - j) : (pr.M - i);  // right overhang length
    uint32_t* ops = p.ops + pr.ops_off;
    uint32_t cur = 0;
    int n = 0;
    int64_t cached_word = -1;
    uint32_t cached = 0;
    while (i >= 1 && j >= 1) {
        const DirAddr a = dir_addr(pr, i, j);
        if (a.word != cached_word) { cached = p.dir[a.word]; cached_word = a.word; }
        const uint32_t d = (cached >> a.shift) & 3u;
        cur |= d << (2 * (n & 15));
        if ((n & 15) == 15) { ops[n >> 4] = cur; cur = 0; }
        ++n;
        if (d == DIR_DIAG) { --i; --j; }
        else if (d == DIR_UP) { --i; }
        else { --j; }
    }
    if (n & 15) ops[n >> 4] = cur;
    // left overhang and terminal-gap add-back (gotoh.cpp:491-510): k leftover characters of
    // exactly one sequence; term==0 ADDS gep per character and gip once.
    const int k = i > j ? i : j;   // one of them is 0
    int sc = p.score[pi];
    if (p.term == 0 && k > 0) sc += k * p.gep + p.gip;
    p.score[pi] = sc;
    p.nops[pi] = n;
    p.i0[pi] = i;
    p.j0[pi] = j;
    p.out_len[pi] = k + n + right;
    if (p.scan_in) p.scan_in[pr.orig] = p.scan_words ? ((n + 15) >> 4) : (k + n + right);
}

// ------------------------------------------------------------------------------------
// Caller-order exclusive prefix sum of the per-pair result sizes (tight / compact result forms): out[k] = sum of
// in[0..k), out[n] = total.  One SMALL CTA (4 warps), tiles of 128 x 16 elements, warp-shuffle scans: in the one-shot
// pipeline it has to find room on an SM next to the persistent forward CTAs of the following slabs (a 1024-thread CTA
// needs most of an SM's registers and waited for a forward kernel to drain - measured on B200), and the input is at
// most a slab's pairs (L2-resident).
// ------------------------------------------------------------------------------------
enum { SCAN_THREADS = 128, SCAN_ITEMS = 16 };
__global__ void __launch_bounds__(SCAN_THREADS) k_scan(const int32_t* in, int64_t* out, int n, int64_t* total_host) {
    enum { NW = SCAN_THREADS / 32 };
    __shared__ long long warp_sum[NW];
    __shared__ long long carry_sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_sh = 0;
    __syncthreads();
    for (int base = 0; base < n; base += SCAN_THREADS * SCAN_ITEMS) {
        const int first = base + tid * SCAN_ITEMS;
        long long v[SCAN_ITEMS], sum = 0;
#pragma unroll
        for (int x = 0; x < SCAN_ITEMS; ++x) { v[x] = (first + x < n) ? (long long)in[first + x] : 0; sum += v[x]; }
        long long inc = sum;                                   // inclusive scan of the thread sums within the warp
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { const long long o = __shfl_up_sync(0xffffffffu, inc, off); if (lane >= off) inc += o; }
        if (lane == 31) warp_sum[warp] = inc;
        __syncthreads();
        if (tid == 0) {                                        // exclusive prefix of the NW warp totals
            long long run = 0;
#pragma unroll
            for (int w = 0; w < NW; ++w) { const long long t = warp_sum[w]; warp_sum[w] = run; run += t; }
        }
        __syncthreads();
        const long long carry = carry_sh;
        long long run = carry + warp_sum[warp] + (inc - sum);
#pragma unroll
        for (int x = 0; x < SCAN_ITEMS; ++x) { if (first + x < n) out[first + x] = run; run += v[x]; }
        __syncthreads();
        if (tid == SCAN_THREADS - 1) carry_sh = run;           // last thread's running sum = total so far
        __syncthreads();
    }
    if (tid == 0) { out[n] = carry_sh; *total_host = carry_sh; }     // total_host: mapped pinned memory, see k_publish
}

// The slab's total result size goes to the host through mapped pinned memory, written by a kernel: a D2H copy issued
// at enqueue time would sit in the copy engine's FIFO in front of the result copies of EARLIER slabs (which the host
// can only issue once it knows their totals) and block them until this slab's kernels are done - measured on B200.
__global__ void k_publish(const unsigned long long* counter, int64_t* total_host) { *total_host = (int64_t)*counter; }

// ------------------------------------------------------------------------------------
// Compact result form: per pair one 8-word record and its op script.  The scripts are packed in the order the warps
// get there (one atomicAdd on a word counter per pair; the per-pair offset is an explicit output of the call, so no
// caller-order prefix sum - and no extra kernel between traceback and copy - is needed).  One warp per pair.
// (include/gotoh_b200.h: GOTOH_B200_REC_*)
// ------------------------------------------------------------------------------------
struct PackParams {
    const PairInfo* pairs;
    int32_t pair_count;
    const uint32_t* ops;
    const int32_t* nops;
    const int32_t* i0;
    const int32_t* j0;
    const int32_t* end_i;
    const int32_t* end_j;
    const int32_t* out_len_plan;
    const int32_t* score_plan;
    int64_t* off;                  // out, caller order: first word of the pair's script in cops
    unsigned long long* counter;   // words handed out so far (zeroed per run; its final value is the slab's total)
    int32_t* rec;                  // caller order, 8 words per pair
    uint32_t* cops;
};

__global__ void __launch_bounds__(128) k_pack_ops(const PackParams p) {
    const int lane = threadIdx.x & 31;
    const int pi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (pi >= p.pair_count) return;
    const PairInfo pr = p.pairs[pi];
    const int n = p.nops[pi];
    const uint32_t* src = p.ops + pr.ops_off;
    const int nw = (n + 15) >> 4;
    unsigned long long at = 0;
    if (lane == 0) { at = atomicAdd(p.counter, (unsigned long long)nw); p.off[pr.orig] = (int64_t)at; }
    at = __shfl_sync(0xffffffffu, at, 0);
    uint32_t* dst = p.cops + at;
    for (int x = lane; x < nw; x += 32) dst[x] = src[x];
    if (lane < 8) {
        int v;
        switch (lane) {
            case 0: v = p.score_plan[pi]; break;
            case 1: v = p.out_len_plan[pi]; break;
            case 2: v = p.i0[pi]; break;
            case 3: v = p.j0[pi]; break;
            case 4: v = p.end_i[pi]; break;
            case 5: v = p.end_j[pi]; break;
            case 6: v = n; break;
            default: v = (pr.M < 65536 && pr.N < 65536) ? ((pr.M << 16) | pr.N) : -1; break;
        }
        p.rec[(int64_t)pr.orig * 8 + lane] = v;
    }
}

// ------------------------------------------------------------------------------------
// K4b  string emit: one warp per pair (gotoh.cpp:436-449, 461-476, 493-513)
// ------------------------------------------------------------------------------------
struct EmitParams {
    const PairInfo* pairs;
    int32_t pair_first, pair_count;
    const uint8_t* ref_raw;
    const uint8_t* qry;
    const uint32_t* ops;
    const int32_t* nops;
    const int32_t* i0;
    const int32_t* j0;
    const int32_t* end_i;
    const int32_t* end_j;
    const int32_t* out_len_plan;   // plan order
    const int32_t* score_plan;
    uint8_t* out_ref;
    uint8_t* out_qry;
    int32_t* out_len;              // caller order
    int32_t* out_score;            // caller order
    const int64_t* tight_off;      // tight result form: caller-order byte offsets from k_scan (no stride tail); else NULL
};

__global__ void __launch_bounds__(128) k_emit(const EmitParams p) {
    const int lane = threadIdx.x & 31;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= p.pair_count) return;
    const int pi = p.pair_first + w;
    const PairInfo pr = p.pairs[pi];
    const uint8_t* a = p.ref_raw + pr.ref_pos;
    const uint8_t* b = p.qry + pr.qry_pos;
    const int64_t ooff = p.tight_off ? p.tight_off[pr.orig] : pr.out_off;
    uint8_t* oa = p.out_ref + ooff;
    uint8_t* ob = p.out_qry + ooff;
    const int i0 = p.i0[pi], j0 = p.j0[pi], n = p.nops[pi];
    const int ei = p.end_i[pi], ej = p.end_j[pi];
    const uint32_t* ops = p.ops + pr.ops_off;
    if (lane == 0) { p.out_len[pr.orig] = p.out_len_plan[pi]; p.out_score[pr.orig] = p.score_plan[pi]; }

    // left overhang: the leftover prefix of exactly one sequence against '-'
    const int lo = i0 > j0 ? i0 : j0;
    if (i0 > j0) { for (int x = lane; x < lo; x += 32) { oa[x] = a[x]; ob[x] = '-'; } }
    else { for (int x = lane; x < lo; x += 32) { oa[x] = '-'; ob[x] = b[x]; } }

    // middle: op script reversed; cursors by ballot prefix sums
    int ia = i0, jb = j0;
    for (int base = 0; base < n; base += 32) {
        const int x = base + lane;
        uint32_t op = 3;
        if (x < n) { const int f = n - 1 - x; op = (ops[f >> 4] >> (2 * (f & 15))) & 3u; }
        const bool ca = (x < n) && (op != DIR_LEFT);   // consumes a reference character
        const bool cb = (x < n) && (op != DIR_UP);     // consumes a query character
        const unsigned ma = __ballot_sync(0xffffffffu, ca), mb = __ballot_sync(0xffffffffu, cb);
        const unsigned lt = (1u << lane) - 1u;
        if (x < n) {
            oa[lo + x] = ca ? a[ia + __popc(ma & lt)] : (uint8_t)'-';
            ob[lo + x] = cb ? b[jb + __popc(mb & lt)] : (uint8_t)'-';
        }
        ia += __popc(ma);
        jb += __popc(mb);
    }
    // right overhang (gotoh.cpp:434-449)
    const int ro_base = lo + n;
    if (ei == pr.M && ej < pr.N) { for (int x = lane; x < pr.N - ej; x += 32) { oa[ro_base + x] = '-'; ob[ro_base + x] = b[ej + x]; } }
    else { for (int x = lane; x < pr.M - ei; x += 32) { oa[ro_base + x] = a[ei + x]; ob[ro_base + x] = '-'; } }
    // bytes between out_len and the caller's stride are defined as 0
    if (!p.tight_off) for (int x = p.out_len_plan[pi] + lane; x < pr.out_cap; x += 32) { oa[x] = 0; ob[x] = 0; }
}

}  // namespace gotoh

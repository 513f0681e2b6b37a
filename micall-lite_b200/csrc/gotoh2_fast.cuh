// gotoh2_fast.cuh - tuned sm_100a kernels for MiCall-Lite's LIVE aligner `_gotoh2.align`
// (SURVEY.md 8f "next" #1; reference micall/alignment/src/_gotoh2.c, driven by gotoh2.py:74-96 from
// core/remap.py:248 and core/aln2counts.py:187).  Used whenever gop >= 0 and gep >= 0 (every caller in the
// reference); gotoh2_kernels.cuh keeps the general kernels for exotic penalties.
//
//   k2f<K,MULTI,BITS>  cost_assignment (_gotoh2.c:137-201).  One warp per STRIP TASK (pair, strip of 32*K
//        columns); lane l owns K columns and is one row behind lane l-1.  Costs live in the frame
//        X~(i,j) = X(i,j) - (i+j)*u, which removes the "+u" of both gap recurrences (:156,:166):
//            q~ = min(q~left, R~left + v)   p~ = min(p~up, R~up + v)   R~ = min(R~diag - d - 2u, p~, q~)
//        = 2 VIADDMNMX + 1 add + 1 VIMNMX3.  The seven tie bits of a cell (:157-176 d,e,f,g; :190-198 a,b,c)
//        come from saturating differences instead of compare+select:
//            code_DE = clamp(p~up - (R~up+v) + 1, 0, 2)     one VIADDMNMX.RELU   (0: d  1: d,e  2: e)
//            code_FG = clamp(q~left - (R~left+v) + 1, 0, 2) one VIADDMNMX.RELU
//            not_a   = min(p~ - R~, 1), not_b, not_c        one VIADDMNMX each
//        and are packed by multiply-add chains on the FMA pipe.  BITS=false is the score-only form (no tie bits,
//        no arena) used for the edit distance of remap.py:250.
//        Strips of one pair run CONCURRENTLY on different warps (any CTA): strip s publishes its last column
//        (R~, q~) to global memory every 32 rows and strip s+1 follows it ~64 rows behind - an anti-diagonal
//        wavefront across the whole grid.  Tasks are claimed from an atomic counter in (pair, strip) order, so
//        a strip's producer is always claimed earlier and is running: the spin-waits cannot deadlock.
//   k2r<K,MULTI>       edge_assignment, Altschul-Erickson steps 8-11 (_gotoh2.c:205-312), the exact time
//        reversal of k2f, bit-parallel over the K columns of a lane: the b-chain along a row
//        fin_b[j] = G[j] | (P[j] & fin_b[j+1]) is a carry chain and is resolved by ONE integer addition.
//   k2f_walk           traceback with priority a > b > c (_gotoh2.c:374-408) into the op script k_emit consumes.
#pragma once

#include "gotoh_kernels.cuh"

namespace gotoh {
namespace g2f {

enum { INF = 1 << 29, BIGE = 1 << 28, FSTEPS = 4, PUB = 32, BND_EMPTY = (int)0x80808080, PART_EMPTY2 = (int)0x80808080 };

struct StripTask {
    int32_t pair;
    int32_t strip;
};

struct Extra {              // per pair, beside PairInfo
    int64_t bnd_off;        // first boundary-column entry of this pair: [(nstrips-1)][M+2]
    int32_t slot0;          // first (pair, strip) slot: progress counters, last-row partials
    int32_t nstrips;
};

struct Params {
    const PairInfo* pairs;      // M = l1, N = l2, dir_off = first uint4 of the pair in BOTH planes, nblk = blocks per strip
    const Extra* extra;
    const StripTask* tasks;     // (pair, strip) in pair-major, strip-ascending order
    int32_t task_count;
    const uint8_t* s1_idx;      // alphabet indices of seq1, REF_PAD zero bytes on both sides of every sequence
    const uint8_t* s2_idx;
    const int32_t* dmat;        // l*l substitution scores (gotoh2.py:47-64)
    int32_t l, v, u, is_global; // v = gap open, u = gap extend (_gotoh2.c:31-32)
    uint32_t two, four;         // == 2, 4 at run time; opaque so acc*two+x stays an IMAD (FMA pipe)
    uint32_t neg1;              // == 0xffffffff at run time: c - r as r*neg1 + c on the FMA pipe
    int32_t inf16, shift16;     // k2f_x2: +infinity (already shifted) and the frame shift S of the 16-bit frame (host: g2_fits_int16)
    uint4* lo;                  // plane 0: code_DE | code_FG << 16, one word per lane-step
    uint4* hi;                  // plane 1: a | b << 8 | c << 16 (k2f: forward bits, k2r: final bits)
    int2* bnd;                  // forward strip boundaries (R~, q~) per row
    uint8_t* rbnd;              // reverse strip boundaries: F | G<<1 | fin_b<<2 | fin_c<<3 of a strip's first column
    int32_t* prog_f;            // rows published per (pair, strip): forward counts up from 0 ...
    int32_t* prog_r;            // ... reverse counts DOWN from M+1 (stored as rows still missing)
    int32_t* part_min;          // last-row partial minimum per (pair, strip)
    int32_t* part_j;
    int32_t* best;              // R at the start cell (score = -best)
    int32_t* start_i;
    int32_t* start_j;
    uint32_t* counter_f;        // task schedulers
    uint32_t* counter_r;
};

template <int K>
struct Smem {
    enum { K4 = (K + 3) / 4 };
    static __host__ __device__ size_t per_warp(int l) { return (size_t)l * K4 * 32 * 16 + 2 * 32 * sizeof(int2); }
};

// -------------------------------------------------------------------------------------------------------
// forward
// -------------------------------------------------------------------------------------------------------
template <int K, bool MULTI, bool BITS>
struct Fwd {
    enum { K4 = (K + 3) / 4, KMASK = (1 << K) - 1 };
    int lane, M, N, j0, strip, u, v, c1v, c2v;
    bool last_strip;
    unsigned two, four, keep, neg1;
    int injq, c0run, c0step;
    unsigned CA3;                  // (KMASK + off_K) * 0x010101
    const uint4* prof_lane;
    const uint8_t* cls;
    int2* ring;
    const int2* bnd_in;
    int2* bnd_out;
    int2 nextb;                    // MULTI: the next window of the left boundary column, prefetched
    int vq[K];
    int R[K], P[K], nRu[K];
    int sendR, sendQ, Rd_in, next_cls;
    int bf, best_i;                // last column: running minimum in the frame of the current row, and its row
    int row_min, row_j, r_ll;      // last row candidates of this lane (unframed)
    uint4 wlo, whi;

    // row 0: R(0,j) = v + j*u (global) or 0 (local), p(0,j) = +inf (_gotoh2.c:96-116); padding columns clone column N
    __device__ __forceinline__ int row0(int j, int is_global) const { return j == 0 ? 0 : (is_global ? v : -j * u); }
    __device__ __forceinline__ void row0_init(int is_global) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            R[k] = row0(min(j0 + k + 1, N), is_global);
            P[k] = INF;
            nRu[k] = c1v - R[k];
        }
        Rd_in = row0(min(j0, N), is_global);
        bf = row0(N, is_global);
        best_i = 0;
    }

    template <bool SLOW>
    __device__ __forceinline__ void step(const int t, const int s, const int is_global) {
        const int i = t - lane;
        const int my_cls = next_cls;
        next_cls = cls[i];
        int Rl = __shfl_up_sync(0xffffffffu, sendR, 1);
        int Ql = __shfl_up_sync(0xffffffffu, sendQ, 1);
        int rdiag = Rd_in;
        if (MULTI && strip > 0) {
            if (((t - 1) & 31) == 0) {
                // self-validating boundary column (host preset: 0x80 bytes, BND_EMPTY is no reachable cost): re-read a row
                // until the left strip has written it, and prefetch the next window a window ahead
                __syncwarp();
                const int row = t + lane;
                const bool real_row = (row >= 1 && row <= M);
                int2 b = nextb;
                for (;;) {
                    const bool miss = real_row && b.x == BND_EMPTY;
                    if (!__any_sync(0xffffffffu, miss)) break;
                    if (miss) b = ld_cg(&bnd_in[row]);
                    gotoh_pause();
                }
                const int row2 = row + 32;
                nextb = (row2 >= 1 && row2 <= M) ? ld_cg(&bnd_in[row2]) : make_int2(0, 0);
                ring[(((t - 1) >> 5) & 1) * 32 + lane] = b;
                __syncwarp();
            }
            if (lane == 0) {
                const int2 b = ring[(((t - 1) >> 5) & 1) * 32 + ((t - 1) & 31)];
                Rl = b.x; Ql = b.y;
            }
        } else {
            // column 0 (_gotoh2.c:101-116): R~(i,0) = v (global) or -i*u (local), q(i,0) = +inf, injected into lane 0
            c0run += c0step;
            Rl = (int)((unsigned)Rl * keep + (unsigned)c0run);
            Ql = (int)((unsigned)Ql * keep + (unsigned)injq);
        }
        Rd_in = Rl;

        const uint4* prow = prof_lane + my_cls * (K4 * 32);
        int Rleft = Rl, nRleft = c1v - Rl, q = Ql;
        unsigned accDE = 0, accFG = 0, accA = 0, accB = 0, accC = 0;
#pragma unroll
        for (int kq = 0; kq < K4; ++kq) {
            const uint4 e4 = prow[kq * 32];
            const int ev[4] = {(int)e4.x, (int)e4.y, (int)e4.z, (int)e4.w};
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const int k = kq * 4 + kk;
                if (k < K) {
                    if (BITS) {
                        const int cFG = __viaddmin_s32_relu(q, nRleft, 2);       // f,g of this cell (_gotoh2.c:167-173)
                        const int cDE = __viaddmin_s32_relu(P[k], nRu[k], 2);    // d,e of this cell (_gotoh2.c:157-163)
                        accFG = accFG * four + (unsigned)cFG;
                        accDE = accDE * four + (unsigned)cDE;
                    }
                    q = __viaddmin_s32(Rleft, vq[k], q);                          // _gotoh2.c:166
                    const int p = __viaddmin_s32(R[k], v, P[k]);                  // _gotoh2.c:156
                    const int dg = rdiag + ev[kk];                                // _gotoh2.c:185
                    const int r = __vimin3_s32(dg, p, q);                         // _gotoh2.c:186-187
                    const int nr = (int)((unsigned)r * neg1 + (unsigned)c1v);    // (1 - v) - r, on the FMA pipe
                    if (BITS) {
                        accA = accA * two + (unsigned)__viaddmin_s32(p, nr, c2v);   // (1-v) + [R != p]   (_gotoh2.c:190-192)
                        accB = accB * two + (unsigned)__viaddmin_s32(q, nr, c2v);   // (1-v) + [R != q]   (:193-195)
                        accC = accC * two + (unsigned)__viaddmin_s32(dg, nr, c2v);  // (1-v) + [R != diag] (:196-198)
                    }
                    rdiag = R[k];
                    R[k] = r; P[k] = p; nRu[k] = nr;
                    Rleft = r; nRleft = nr;
                }
            }
        }
        sendR = R[K - 1];
        sendQ = q;
        if (BITS) {
            const unsigned wl = accFG * 65536u + accDE;
            const unsigned wh = CA3 - (accA + accB * 256u + accC * 65536u);
            if (s == 0) { wlo.x = wl; whi.x = wh; } else if (s == 1) { wlo.y = wl; whi.y = wh; }
            else if (s == 2) { wlo.z = wl; whi.z = wh; } else { wlo.w = wl; whi.w = wh; }
        }
        // last column: first strict minimum scanning top-down (_gotoh2.c:330-339), in the frame of the current row
        if (!MULTI || last_strip) {
            bf -= u;
            if (!SLOW || (i >= 1 && i <= M)) {
                if (R[K - 1] < bf) { bf = R[K - 1]; best_i = i; }
            }
        }
        if (MULTI && !last_strip && lane == 31 && i >= 1 && i <= M) bnd_out[i] = make_int2(R[K - 1], q);
        if (SLOW) {
            if (i == 0) row0_init(is_global);
            if (i == M) {
                // bottom row: first strict minimum left to right (_gotoh2.c:341-350), unframed
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int j = j0 + k + 1;
                    const int val = R[k] + (M + j) * u;
                    if (j <= N && val < row_min) { row_min = val; row_j = j; }
                }
                r_ll = R[K - 1] + (M + N) * u;          // owner lane of column N: R(l1, l2)
            }
        }
    }
};

template <int K, bool MULTI, bool BITS>
__global__ void __launch_bounds__(128, 4) k2f(const Params p) {
    typedef Fwd<K, MULTI, BITS> W;
    enum { K4 = W::K4 };
    GOTOH_DYN_SMEM(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* my_smem = smem_raw + (size_t)warp * Smem<K>::per_warp(p.l);
    uint4* prof = reinterpret_cast<uint4*>(my_smem);
    W w;
    w.lane = lane;
    w.u = p.u; w.v = p.v;
    w.c1v = 1 - p.v; w.c2v = 2 - p.v;
    w.two = p.two; w.four = p.four; w.neg1 = p.neg1;
    w.prof_lane = prof + lane;
    w.ring = reinterpret_cast<int2*>(my_smem + (size_t)p.l * K4 * 32 * 16);
    w.CA3 = (unsigned)(W::KMASK + (1 - p.v) * W::KMASK) * 0x010101u;
    const int e_pad = BIGE;

    for (;;) {
        unsigned tsk = 0;
        if (lane == 0) tsk = atomicAdd(p.counter_f, 1u);
        tsk = __shfl_sync(0xffffffffu, tsk, 0);
        if (tsk >= (unsigned)p.task_count) break;
        const StripTask task = p.tasks[tsk];
        const PairInfo pr = p.pairs[task.pair];
        const Extra ex = p.extra[task.pair];
        const int M = pr.M, N = pr.N, strip = task.strip, nstrips = ex.nstrips, nblk = pr.nblk;
        const uint8_t* s2 = p.s2_idx + pr.qry_pos;
        const int j0 = (strip * 32 + lane) * K;
        w.M = M; w.N = N; w.strip = strip; w.j0 = j0;
        w.last_strip = (strip == nstrips - 1);
        w.cls = p.s1_idx + pr.ref_pos;
        w.row_min = 2147483647; w.row_j = 0; w.r_ll = 0;
        if (MULTI) {
            w.bnd_in = p.bnd + ex.bnd_off + (int64_t)(strip - 1) * (M + 2);
            w.bnd_out = p.bnd + ex.bnd_off + (int64_t)strip * (M + 2);
        }
        const bool col0 = (!MULTI || strip == 0);
        w.keep = (p.two >> 1) - ((lane == 0 && col0) ? 1u : 0u);
        w.injq = (lane == 0 && col0) ? INF : 0;
        w.c0run = (lane == 0 && col0 && p.is_global) ? p.v : 0;
        w.c0step = (lane == 0 && col0 && !p.is_global) ? -p.u : 0;

        // ---- query profile of this strip: prof[class][k/4][lane] = -d[class][b_j] - 2u; padding columns never win
        __syncwarp();
#pragma unroll
        for (int k = 0; k < K; ++k) w.vq[k] = (j0 + k) < N ? p.v : 0;
        for (int c = 0; c < p.l; ++c) {
            const int32_t* drow = p.dmat + c * p.l;
#pragma unroll
            for (int kq = 0; kq < K4; ++kq) {
                unsigned e[4];
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const int k = kq * 4 + kk, ja = j0 + k;
                    int ea = e_pad;
                    if (k < K && ja < N) ea = -drow[s2[ja]] - 2 * p.u;
                    e[kk] = (unsigned)ea;
                }
                prof[(c * K4 + kq) * 32 + lane] = make_uint4(e[0], e[1], e[2], e[3]);
            }
        }
        __syncwarp();
        w.sendR = 0; w.sendQ = 0;
        w.wlo = make_uint4(0, 0, 0, 0); w.whi = make_uint4(0, 0, 0, 0);
        w.row0_init(p.is_global);
        w.next_cls = w.cls[-lane];

        w.nextb = make_int2(BND_EMPTY, BND_EMPTY);
        uint4* dlo = BITS ? p.lo + pr.dir_off + (int64_t)strip * nblk * 32 + lane : nullptr;
        uint4* dhi = BITS ? p.hi + pr.dir_off + (int64_t)strip * nblk * 32 + lane : nullptr;
        for (int tb = 0; tb < nblk; ++tb) {
            const int t0 = tb * FSTEPS + 1, hi = t0 + FSTEPS - 1;
            const bool slow = (t0 <= 31) || (hi >= M);
            if (slow) {
#pragma unroll
                for (int s = 0; s < FSTEPS; ++s) w.template step<true>(t0 + s, s, p.is_global);
            } else {
#pragma unroll
                for (int s = 0; s < FSTEPS; ++s) w.template step<false>(t0 + s, s, p.is_global);
            }
            if (BITS) { dlo[(int64_t)tb * 32] = w.wlo; dhi[(int64_t)tb * 32] = w.whi; }
        }

        // ---- bottom-row partial of this strip (smallest j wins ties), then the final publish ------------
        int row_min = w.row_min, row_j = w.row_j;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const int om = __shfl_xor_sync(0xffffffffu, row_min, off);
            const int oj = __shfl_xor_sync(0xffffffffu, row_j, off);
            if (om < row_min || (om == row_min && oj < row_j)) { row_min = om; row_j = oj; }
        }
        if (MULTI && !w.last_strip) {
            // partial of this strip: column first, then (fenced) the minimum, which the host preset to PART_EMPTY2
            if (lane == 0) { p.part_j[ex.slot0 + strip] = row_j; __threadfence(); st_volatile(&p.part_min[ex.slot0 + strip], row_min); }
            continue;
        }
        // ---- last strip: start cell (_gotoh2.c:327-352) ---------------------------------------------------
        const int owner = ((N - 1) / K) & 31;
        const int i_fin = nblk * FSTEPS - lane;
        int col_min = w.bf + (i_fin + N) * p.u;            // unframed
        col_min = __shfl_sync(0xffffffffu, col_min, owner);
        const int col_i = __shfl_sync(0xffffffffu, w.best_i, owner);
        const int r_ll = __shfl_sync(0xffffffffu, w.r_ll, owner);
        if (lane == 0) {
            int best = r_ll, bi = M, bj = N;
            if (!p.is_global) {
                if (col_min < best) { best = col_min; bi = col_i; bj = N; }
                int rm = 0, rj = 0;                          // R(l1, 0) = 0 in local mode (_gotoh2.c:111)
                if (MULTI) {
                    for (int s = 0; s < nstrips - 1; ++s) {
                        int pm;                                   // the strip may still be writing its partial
                        while ((pm = ld_volatile(&p.part_min[ex.slot0 + s])) == PART_EMPTY2) gotoh_pause();
                        __threadfence();
                        const int pj = ld_cg(&p.part_j[ex.slot0 + s]);
                        if (pm < rm) { rm = pm; rj = pj; }
                    }
                }
                if (row_min < rm) { rm = row_min; rj = row_j; }
                if (rm < best) { best = rm; bi = M; bj = rj; }
            }
            p.best[task.pair] = best;
            p.start_i[task.pair] = bi;
            p.start_j[task.pair] = bj;
        }
    }
}

// -------------------------------------------------------------------------------------------------------
// forward, int16x2: two pairs that share seq1 in the halves of one register (single strip only)
// -------------------------------------------------------------------------------------------------------
// In the frame X~ = X - (i+j)*u every value of a pair lies in [-(dmax*min(l1,l2) + (l1+l2)*u), 3v], so short second
// sequences fit 16 bits whatever the length of the first (the host proves it per pair, g2_fits_int16).  The cell is
// the int32 one with VIADDMNMX.S16x2 / VIMNMX3.S16x2 (7 + 1 ALU-pipe instructions per cell COUPLE).  Everything else
// is moved to the FMA pipe by a second shift of the frame, X' = X~ - S with S = p.shift16 > 4v - dmin:
//   * every R' and every diagonal candidate dg' is then strictly NEGATIVE and every negated value (1-v) - R', -R' is
//     strictly POSITIVE in both halves;
//   * a packed word whose low half is negative equals the "linear" word lo + 65536*hi plus 65536, a packed word whose
//     low half is non-negative equals it exactly, and linear words add and negate component-wise in plain 32-bit
//     arithmetic.  Hence   dg' = rdiag'*1 + lin(e)          (profile entries are stored linear, not packed)
//                          nr  = R'*(-1) + lin(1-v) + 65536  ((1-v) - R' per half)
//                          nn  = R'*(-1) + 65536             (-R' per half)
//     are single IMADs (multipliers opaque run-time values so ptxas keeps them on the FMA pipe) instead of the
//     LOP3 + 2 VIADD.16x2 a per-half subtraction costs.
// S cancels in every recurrence and tie code (they are differences); it enters row 0 / column 0 and is added back when
// the start-cell candidates are unframed.  Each pair still gets its own two planes in the int32 layout, so k2r and the
// traceback kernels are unchanged.
__device__ __forceinline__ unsigned pk2(int lo, int hi) { return ((unsigned)lo & 0xffffu) | ((unsigned)hi << 16); }
__device__ __forceinline__ unsigned lin2(int lo, int hi) { return (unsigned)lo + ((unsigned)hi << 16); }
__device__ __forceinline__ int lo16(unsigned v) { return (int)(short)(v & 0xffffu); }
__device__ __forceinline__ int hi16(unsigned v) { return (int)(short)(v >> 16); }

template <int K>
struct FwdX2 {
    enum { K4 = (K + 3) / 4, KMASK = (1 << K) - 1 };
    int lane, M, Na, Nb, j0, u, v, inf16, S;
    unsigned one, two, four, neg1, keep;
    unsigned v2, negu2l, c_nr, c_nn, injq, c0run, c0step;   // negu2l, c0step: LINEAR words (lo + 65536*hi), added by IMAD
    const uint4* prof_lane;
    const uint8_t* cls;
    const uint8_t* clsl;           // cls - lane: one pointer add per block instead of per step
    unsigned vq[K];
    unsigned R[K], P[K], nRu[K];
    unsigned sendR, sendQ, Rd_in;
    int next_cls;
    unsigned bf;
    int best_i_a, best_i_b;
    int row_min_a, row_j_a, r_ll_a, row_min_b, row_j_b, r_ll_b;
    uint4 wlo_a, whi_a, wlo_b, whi_b;

    // row 0 in the shifted frame (_gotoh2.c:96-116): R'(0,j) = v - S (global) or -j*u - S (local), p(0,j) = +inf
    __device__ __forceinline__ int row0(int j, int is_global) const { return (j == 0 ? 0 : (is_global ? v : -j * u)) - S; }
    __device__ __forceinline__ void row0_init(int is_global) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int ra = row0(min(j0 + k + 1, Na), is_global), rb = row0(min(j0 + k + 1, Nb), is_global);
            R[k] = pk2(ra, rb);
            P[k] = pk2(inf16, inf16);
            nRu[k] = pk2(1 - v - ra, 1 - v - rb);
        }
        Rd_in = pk2(row0(min(j0, Na), is_global), row0(min(j0, Nb), is_global));
        bf = pk2(row0(Na, is_global), row0(Nb, is_global));
        best_i_a = best_i_b = lane;                 // the STEP of row 0 (row = step - lane)
    }

    template <bool SLOW>
    __device__ __forceinline__ void step(const int t, const int s, const int is_global) {
        const int i = t - lane;
        const int my_cls = next_cls;
        next_cls = clsl[t];
        unsigned Rl = __shfl_up_sync(0xffffffffu, sendR, 1);
        unsigned Ql = __shfl_up_sync(0xffffffffu, sendQ, 1);
        unsigned rdiag = Rd_in;
        c0run = c0run * one + c0step;                   // column 0 (_gotoh2.c:101-116); both operands keep a negative low half (or are 0)
        Rl = Rl * keep + c0run;
        Ql = Ql * keep + injq;
        Rd_in = Rl;

        const uint4* prow = prof_lane + my_cls * (K4 * 32);
        unsigned Rleft = Rl, nRleft = Rl * neg1 + c_nr, q = Ql;
        unsigned accDE = 0, accFG = 0, accA = 0, accB = 0, accC = 0;
#pragma unroll
        for (int kq = 0; kq < K4; ++kq) {
            const uint4 e4 = prow[kq * 32];
            const unsigned ev[4] = {e4.x, e4.y, e4.z, e4.w};
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const int k = kq * 4 + kk;
                if (k < K) {
                    const unsigned cFG = __viaddmin_s16x2_relu(q, nRleft, 0x00020002u);      // _gotoh2.c:167-173
                    const unsigned cDE = __viaddmin_s16x2_relu(P[k], nRu[k], 0x00020002u);   // _gotoh2.c:157-163
                    accFG = accFG * four + cFG;
                    accDE = accDE * four + cDE;
                    q = __viaddmin_s16x2(Rleft, vq[k], q);                                    // _gotoh2.c:166
                    const unsigned p = __viaddmin_s16x2(R[k], v2, P[k]);                      // _gotoh2.c:156
                    const unsigned dg = rdiag * one + ev[kk];                                 // _gotoh2.c:185 (linear add)
                    const unsigned r = __vimin3_s16x2(dg, p, q);                              // _gotoh2.c:186-187
                    const unsigned nn = r * neg1 + c_nn;                                      // -R' per half
                    const unsigned nr = r * neg1 + c_nr;                                      // (1-v) - R' per half
                    accA = accA * two + __viaddmin_s16x2(p, nn, 0x00010001u);                 // [R != p]    (:190-192)
                    accB = accB * two + __viaddmin_s16x2(q, nn, 0x00010001u);                 // [R != q]    (:193-195)
                    accC = accC * two + __viaddmin_s16x2(dg, nn, 0x00010001u);                // [R != diag] (:196-198)
                    rdiag = R[k];
                    R[k] = r; P[k] = p; nRu[k] = nr;
                    Rleft = r; nRleft = nr;
                }
            }
        }
        sendR = R[K - 1];
        sendQ = q;
        {
            // per pair: lo = code_DE | code_FG << 16, hi = a | b << 8 | c << 16 (bits are "equal" = NOT of the accumulated)
            const unsigned nA = accA ^ (KMASK * 0x00010001u), nB = accB ^ (KMASK * 0x00010001u), nC = accC ^ (KMASK * 0x00010001u);
            const unsigned la = __byte_perm(accDE, accFG, 0x5410), lb = __byte_perm(accDE, accFG, 0x7632);
            const unsigned ha = __byte_perm(__byte_perm(nA, nB, 0x4440), nC, 0x7410);
            const unsigned hb = __byte_perm(__byte_perm(nA, nB, 0x6662), nC, 0x7610);
            if (s == 0) { wlo_a.x = la; whi_a.x = ha; wlo_b.x = lb; whi_b.x = hb; }
            else if (s == 1) { wlo_a.y = la; whi_a.y = ha; wlo_b.y = lb; whi_b.y = hb; }
            else if (s == 2) { wlo_a.z = la; whi_a.z = ha; wlo_b.z = lb; whi_b.z = hb; }
            else { wlo_a.w = la; whi_a.w = ha; wlo_b.w = lb; whi_b.w = hb; }
        }
        // last column: first strict minimum scanning top-down (_gotoh2.c:330-339), per half
        bf = bf * one + negu2l;                         // bf < 0 in both halves, always: a linear add on the FMA pipe
        {
            bool keep_hi, keep_lo;                        // bf <= R: no new minimum
            const unsigned nb = __vibmin_s16x2(bf, R[K - 1], &keep_hi, &keep_lo);
            if (!SLOW || (i >= 1 && i <= M)) {
                bf = nb;
                if (!keep_lo) best_i_a = t;
                if (!keep_hi) best_i_b = t;
            }
        }
        if (SLOW) {
            if (i == 0) row0_init(is_global);
            if (i == M) {
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int j = j0 + k + 1;
                    const int va = lo16(R[k]) + (M + j) * u + S, vb = hi16(R[k]) + (M + j) * u + S;
                    if (j <= Na && va < row_min_a) { row_min_a = va; row_j_a = j; }
                    if (j <= Nb && vb < row_min_b) { row_min_b = vb; row_j_b = j; }
                }
                r_ll_a = lo16(R[K - 1]) + (M + Na) * u + S;
                r_ll_b = hi16(R[K - 1]) + (M + Nb) * u + S;
            }
        }
    }
};

// tasks: (pair = first pair, strip = second pair or -1)
template <int K>
__global__ void __launch_bounds__(128, 4) k2f_x2(const Params p) {
    typedef FwdX2<K> W;
    enum { K4 = W::K4 };
    GOTOH_DYN_SMEM(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* my_smem = smem_raw + (size_t)warp * Smem<K>::per_warp(p.l);
    uint4* prof = reinterpret_cast<uint4*>(my_smem);
    W w;
    w.lane = lane;
    w.u = p.u; w.v = p.v;
    w.inf16 = p.inf16; w.S = p.shift16;
    w.v2 = pk2(p.v, p.v); w.negu2l = lin2(-p.u, -p.u);
    w.c_nn = 65536u;
    w.c_nr = lin2(1 - p.v, 1 - p.v) + 65536u;
    w.one = p.two >> 1; w.two = p.two; w.four = p.four; w.neg1 = p.neg1;
    w.prof_lane = prof + lane;
    w.keep = (p.two >> 1) - (lane == 0 ? 1u : 0u);
    w.injq = lane == 0 ? pk2(p.inf16, p.inf16) : 0u;
    const int e_pad = p.v + 1;                            // padding columns: diag' = R'(i-1,N) + v + 1 > R'(i,N), never wins

    for (;;) {
        unsigned tsk = 0;
        if (lane == 0) tsk = atomicAdd(p.counter_f, 1u);
        tsk = __shfl_sync(0xffffffffu, tsk, 0);
        if (tsk >= (unsigned)p.task_count) break;
        const StripTask task = p.tasks[tsk];
        const int ia = task.pair, ib = task.strip >= 0 ? task.strip : task.pair;
        const PairInfo pa = p.pairs[ia];
        const PairInfo pb = p.pairs[ib];
        const int M = pa.M, Na = pa.N, Nb = pb.N, nblk = pa.nblk;
        const uint8_t* sa = p.s2_idx + pa.qry_pos;
        const uint8_t* sb = p.s2_idx + pb.qry_pos;
        const int j0 = lane * K;
        w.M = M; w.Na = Na; w.Nb = Nb; w.j0 = j0;
        w.cls = p.s1_idx + pa.ref_pos;
        w.clsl = w.cls - lane;
        w.row_min_a = w.row_min_b = 2147483647; w.row_j_a = w.row_j_b = 0; w.r_ll_a = w.r_ll_b = 0;
        // column 0 in the shifted frame: R'(i,0) = v - S (global) or -i*u - S (local)
        w.c0run = lane == 0 ? (p.is_global ? pk2(p.v - p.shift16, p.v - p.shift16) : pk2(-p.shift16, -p.shift16)) : 0u;
        w.c0step = (lane == 0 && !p.is_global) ? lin2(-p.u, -p.u) : 0u;

        __syncwarp();
#pragma unroll
        for (int k = 0; k < K; ++k) w.vq[k] = pk2((j0 + k) < Na ? p.v : 0, (j0 + k) < Nb ? p.v : 0);
        // alphabet indices of this lane's columns, read once (-1: padding column), then one table load per class
        int xa[K], xb[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            xa[k] = (j0 + k) < Na ? (int)sa[j0 + k] : -1;
            xb[k] = (j0 + k) < Nb ? (int)sb[j0 + k] : -1;
        }
        for (int c = 0; c < p.l; ++c) {
            const int32_t* drow = p.dmat + c * p.l;
#pragma unroll
            for (int kq = 0; kq < K4; ++kq) {
                unsigned e[4];
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const int k = kq * 4 + kk;
                    int ea = e_pad, eb = e_pad;
                    if (k < K) {
                        if (xa[k] >= 0) ea = -drow[xa[k]] - 2 * p.u;
                        if (xb[k] >= 0) eb = -drow[xb[k]] - 2 * p.u;
                    }
                    e[kk] = lin2(ea, eb);
                }
                prof[(c * K4 + kq) * 32 + lane] = make_uint4(e[0], e[1], e[2], e[3]);
            }
        }
        __syncwarp();
        w.sendR = 0; w.sendQ = 0;
        w.wlo_a = w.whi_a = w.wlo_b = w.whi_b = make_uint4(0, 0, 0, 0);
        w.row0_init(p.is_global);
        w.next_cls = w.cls[-lane];

        uint4* dlo_a = p.lo + pa.dir_off + lane;
        uint4* dhi_a = p.hi + pa.dir_off + lane;
        uint4* dlo_b = p.lo + pb.dir_off + lane;
        uint4* dhi_b = p.hi + pb.dir_off + lane;
        const bool two_pairs = task.strip >= 0;
        for (int tb = 0; tb < nblk; ++tb) {
            const int t0 = tb * FSTEPS + 1, hi = t0 + FSTEPS - 1;
            const bool slow = (t0 <= 31) || (hi >= M);
            if (slow) {
#pragma unroll
                for (int s = 0; s < FSTEPS; ++s) w.template step<true>(t0 + s, s, p.is_global);
            } else {
#pragma unroll
                for (int s = 0; s < FSTEPS; ++s) w.template step<false>(t0 + s, s, p.is_global);
            }
            dlo_a[(int64_t)tb * 32] = w.wlo_a; dhi_a[(int64_t)tb * 32] = w.whi_a;
            if (two_pairs) { dlo_b[(int64_t)tb * 32] = w.wlo_b; dhi_b[(int64_t)tb * 32] = w.whi_b; }
        }

        // ---- start cells (_gotoh2.c:327-352), per half ------------------------------------------------------
        const int i_fin = nblk * FSTEPS - lane;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            if (half == 1 && !two_pairs) break;
            const int N = half ? Nb : Na;
            int row_min = half ? w.row_min_b : w.row_min_a, row_j = half ? w.row_j_b : w.row_j_a;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const int om = __shfl_xor_sync(0xffffffffu, row_min, off);
                const int oj = __shfl_xor_sync(0xffffffffu, row_j, off);
                if (om < row_min || (om == row_min && oj < row_j)) { row_min = om; row_j = oj; }
            }
            const int owner = ((N - 1) / K) & 31;
            int col_min = (half ? hi16(w.bf) : lo16(w.bf)) + (i_fin + N) * p.u + p.shift16;
            col_min = __shfl_sync(0xffffffffu, col_min, owner);
            const int col_i = __shfl_sync(0xffffffffu, (half ? w.best_i_b : w.best_i_a) - lane, owner);
            const int r_ll = __shfl_sync(0xffffffffu, half ? w.r_ll_b : w.r_ll_a, owner);
            if (lane == 0) {
                int best = r_ll, bi = M, bj = N;
                if (!p.is_global) {
                    if (col_min < best) { best = col_min; bi = col_i; bj = N; }
                    int rm = 0, rj = 0;                          // R(l1, 0) = 0 in local mode (_gotoh2.c:111)
                    if (row_min < rm) { rm = row_min; rj = row_j; }
                    if (rm < best) { best = rm; bi = M; bj = rj; }
                }
                const int pi = half ? ib : ia;
                p.best[pi] = best;
                p.start_i[pi] = bi;
                p.start_j[pi] = bj;
            }
        }
    }
}

// -------------------------------------------------------------------------------------------------------
// reverse sweep
// -------------------------------------------------------------------------------------------------------
// even bits of each 16-bit half -> the low 8 bits of that half
__device__ __forceinline__ unsigned compress_even(unsigned x) {
    x = (x | (x >> 1)) & 0x33333333u;
    x = (x | (x >> 2)) & 0x0f0f0f0fu;
    x = (x | (x >> 4)) & 0x00ff00ffu;
    return x;
}

// One warp sweeping one strip bottom-up.  All planes are K-bit masks (column k at bit K-1-k); pairs of planes travel
// byte-packed in one register so that masks, shifts and the hand-over to the left lane act on several at once.
template <int K, bool MULTI>
struct Rev {
    enum { KMASK = (1 << K) - 1 };
    int lane, M;
    bool last_strip, first_strip;
    unsigned colmask, sent_last, sent_rows;
    unsigned finA_dn, finC_dn, DE_dn;   // row below: final a, final c, forward d | e << 8
    unsigned send;                      // top bits (my first column) of F, G, fin_b of the row just done and fin_c of the row below it
    unsigned in_c_prev;                 // MULTI lane 31: fin_c of the right strip's first column, one row below
    const uint8_t* ring;
    uint8_t* rb_out;

    template <bool SLOW>
    __device__ __forceinline__ unsigned step(const int t, const unsigned wl, const unsigned wh) {
        const int i = t - lane;
        const bool valid = !SLOW || (i >= 1 && i <= M);
        const unsigned m8 = valid ? colmask : 0u;
        const unsigned sent = (SLOW && i == M) ? sent_last : sent_rows;
        // forward bits of this row: 2-bit codes -> planes
        const unsigned even = wl & 0x55555555u, odd = (wl >> 1) & 0x55555555u;
        const unsigned eg = compress_even(even | odd);            // e | g << 16
        const unsigned df = compress_even(odd ^ 0x55555555u);     // d | f << 16
        const unsigned DE = __byte_perm(df, eg, 0x7740) & (m8 * 0x0101u);     // d | e << 8
        const unsigned FG = __byte_perm(df, eg, 0x7762) & (m8 * 0x0101u);     // f | g << 8
        const unsigned own = wh & (m8 * 0x010101u);                           // a | b << 8 | c << 16
        // right neighbour column: lane+1 processed this row in its previous step
        unsigned in = __shfl_down_sync(0xffffffffu, send, 1);
        if (lane == 31) {
            in = 0;
            if (MULTI && !last_strip && valid) {
                const int wdw = (t - 1) >> 5;
                const unsigned x = ring[(wdw & 1) * 32 + (i - (32 * wdw - 30))];
                in = (x & 1u) | ((x & 2u) << 7) | ((x & 4u) << 14) | (in_c_prev << 24);
                in_c_prev = (x >> 3) & 1u;
            }
        }
        const unsigned FGr = (((FG << 1) & 0xfefeu) | (in & 0x0101u)) & (KMASK * 0x0101u);   // f, g of the column to the right
        const unsigned Fr = FGr & 0xffu, Gr = FGr >> 8;
        const unsigned cin = (in >> 16) & 1u;
        const unsigned C1 = ((((finC_dn << 1) | (in >> 24)) & KMASK) | sent) & m8;
        const unsigned A1 = finA_dn, D_dn = DE_dn & 0xffu, E_dn = DE_dn >> 8;
        const unsigned ownA = own & 0xffu, ownB = (own >> 8) & 0xffu, ownC = own >> 16;
        const unsigned K0 = (A1 & E_dn) | C1;
        const unsigned Gg = ownB & K0, Pp = (ownB & Gr) | Fr;
        const unsigned x = Gg | Pp;
        const unsigned cvec = (x + Gg + cin) ^ x ^ Gg;            // bit b: fin_b of the column right of bit b
        const unsigned finB = (cvec >> 1) & m8;
        const unsigned keepm = K0 | (cvec & Gr);                  // step 8 (_gotoh2.c:237-242)
        const unsigned finA = (ownA & keepm) | (A1 & D_dn);       // step 10 (:251-270); A1 and own bits are already masked
        const unsigned finC = ownC & keepm;
        // hand-over to the left lane: top bits of f, g, fin_b of this row and of fin_c one row below
        send = ((FG | (finB << 16) | (finC_dn << 24)) >> (K - 1)) & 0x01010101u;
        if (MULTI && !first_strip && lane == 0 && valid) {
            const unsigned y = ((FG | (finB << 16) | (finC << 24)) >> (K - 1)) & 0x01010101u;
            rb_out[i] = (uint8_t)((y * 0x10204080u) >> 28);       // f | g << 1 | fin_b << 2 | fin_c << 3
        }
        finA_dn = finA; finC_dn = finC; DE_dn = DE;
        return finA | (finB << 8) | (finC << 16);
    }
};

template <int K, bool MULTI>
__global__ void __launch_bounds__(128, 8) k2r(const Params p) {
    typedef Rev<K, MULTI> R;
    enum { KMASK = R::KMASK };
    __shared__ uint8_t s_ring[4][2][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned local = p.is_global ? 0u : 1u;
    R r;
    r.lane = lane;
    r.ring = &s_ring[warp][0][0];

    for (;;) {
        unsigned tsk = 0;
        if (lane == 0) tsk = atomicAdd(p.counter_r, 1u);
        tsk = __shfl_sync(0xffffffffu, tsk, 0);
        if (tsk >= (unsigned)p.task_count) break;
        const StripTask task = p.tasks[p.task_count - 1 - (int)tsk];      // reversed claim order: a strip's right neighbour first
        const PairInfo pr = p.pairs[task.pair];
        const Extra ex = p.extra[task.pair];
        const int M = pr.M, N = pr.N, strip = task.strip, nstrips = ex.nstrips, nblk = pr.nblk;
        const int j0 = (strip * 32 + lane) * K;
        const bool last_strip = (strip == nstrips - 1);
        // columns j0+1 .. j0+K, column k at bit K-1-k; real columns are j <= N
        const int nreal = max(0, min(K, N - j0));
        const unsigned colmask = (unsigned)(KMASK & ~((1 << (K - nreal)) - 1));
        const unsigned bitN = (N > j0 && N <= j0 + K) ? (1u << (K - 1 - (N - j0 - 1))) : 0u;
        r.M = M;
        r.last_strip = last_strip;
        r.first_strip = (strip == 0);
        r.colmask = colmask;
        r.sent_last = local ? colmask : bitN;       // c of the sentinel row below row M (_gotoh2.c:118-133)
        r.sent_rows = local ? bitN : 0u;            // c of the sentinel column right of column N
        r.finA_dn = 0; r.finC_dn = 0; r.DE_dn = 0; r.send = 0; r.in_c_prev = 0;
        r.rb_out = p.rbnd + ex.bnd_off + (int64_t)(strip - 1) * (M + 2);                // read by strip-1
        const uint8_t* rb_in = p.rbnd + ex.bnd_off + (int64_t)strip * (M + 2);          // written by strip+1
        unsigned nextrb = 0xffu;                    // nothing prefetched yet
        uint8_t* ring = &s_ring[warp][0][0];

        const uint4* slo = p.lo + pr.dir_off + (int64_t)strip * nblk * 32 + lane;
        uint4* shi = p.hi + pr.dir_off + (int64_t)strip * nblk * 32 + lane;
        __syncwarp();
        uint4 nl4 = slo[(int64_t)(nblk - 1) * 32], nh4 = shi[(int64_t)(nblk - 1) * 32];
        for (int tb = nblk - 1; tb >= 0; --tb) {
            const int t0 = tb * FSTEPS + 1, thi = t0 + FSTEPS - 1;
            if (MULTI && !last_strip && ((thi & 31) == 0 || tb == nblk - 1)) {
                // lane 31 consumes the right strip's first column: window wdw covers steps 32*wdw+1 .. 32*wdw+32,
                // i.e. lane-31 rows 32*wdw-30 .. 32*wdw+1.  The column is self-validating (host preset 0xff, real bytes
                // are < 16): re-read a row until the right strip has written it; the next (lower) window is prefetched.
                const int wdw = (thi - 1) >> 5;
                const int row = 32 * wdw - 30 + lane;
                const bool real_row = (row >= 1 && row <= M);
                unsigned b = nextrb;
                for (;;) {
                    const bool miss = real_row && b == 0xffu;
                    if (!__any_sync(0xffffffffu, miss)) break;
                    if (miss) b = ld_cg(&rb_in[row]);
                    gotoh_pause();
                }
                if (!real_row) b = 0;
                const int row2 = row - 32;
                nextrb = (row2 >= 1 && row2 <= M) ? (unsigned)ld_cg(&rb_in[row2]) : 0xffu;
                __syncwarp();
                ring[(wdw & 1) * 32 + lane] = (uint8_t)b;
                __syncwarp();
            }
            const uint4 l4 = nl4;
            uint4 h4 = nh4;
            if (tb > 0) { nl4 = slo[(int64_t)(tb - 1) * 32]; nh4 = shi[(int64_t)(tb - 1) * 32]; }   // prefetch the next block
            // FAST: every lane's rows of this block are inside 1 .. M-1
            if (t0 >= 32 && thi < M) {
                h4.w = r.template step<false>(t0 + 3, l4.w, h4.w);
                h4.z = r.template step<false>(t0 + 2, l4.z, h4.z);
                h4.y = r.template step<false>(t0 + 1, l4.y, h4.y);
                h4.x = r.template step<false>(t0 + 0, l4.x, h4.x);
            } else {
                h4.w = r.template step<true>(t0 + 3, l4.w, h4.w);
                h4.z = r.template step<true>(t0 + 2, l4.z, h4.z);
                h4.y = r.template step<true>(t0 + 1, l4.y, h4.y);
                h4.x = r.template step<true>(t0 + 0, l4.x, h4.x);
            }
            shi[(int64_t)tb * 32] = h4;
        }
    }
}

// -------------------------------------------------------------------------------------------------------
// reverse sweep, two pairs per warp (K <= 3, single strip, pairs that share seq1)
// -------------------------------------------------------------------------------------------------------
// The reverse sweep costs ~66 instructions per lane-step whatever K is - it is bit-parallel over a lane's columns - so at
// K = 3 (84-column amino-acid windows) a cell costs 22 instructions.  Every plane of a lane then uses 3 bits of a byte:
// a second pair fits the other nibble (column k of pair B at bit 4 + K-1-k), every operation of Rev::step is bitwise except
// the carry-chain addition, and two K-bit operands plus a carry-in sum to at most 15, so nothing crosses from one nibble
// into the next.  The couples are the ones the int16x2 forward kernel formed (same seq1, same K, hence the same rows and
// blocks); a single pair rides along with a copy of itself.
template <int K>
struct Rev2 {
    static_assert(K <= 3, "two K-bit operands + carry must fit a nibble");
    enum { KMASK = (1 << K) - 1, KM2 = KMASK | (KMASK << 4) };
    int lane, M;
    unsigned colmask, sent_last, sent_rows;          // nibble-packed: pair A | pair B << 4
    unsigned finA_dn, finC_dn, DE_dn;
    unsigned send;

    template <bool SLOW>
    __device__ __forceinline__ unsigned step(const int t, const unsigned wlA, const unsigned whA, const unsigned wlB, const unsigned whB) {
        const int i = t - lane;
        const bool valid = !SLOW || (i >= 1 && i <= M);
        const unsigned m8 = valid ? colmask : 0u;
        const unsigned sent = (SLOW && i == M) ? sent_last : sent_rows;
        // pair B's 2-bit codes move up by 8 bit positions: compress_even then drops them into the high nibble
        const unsigned wl = wlA | (wlB << 8);
        const unsigned even = wl & 0x55555555u, odd = (wl >> 1) & 0x55555555u;
        const unsigned eg = compress_even(even | odd);            // e | g << 16
        const unsigned df = compress_even(odd ^ 0x55555555u);     // d | f << 16
        const unsigned DE = __byte_perm(df, eg, 0x7740) & (m8 * 0x0101u);     // d | e << 8
        const unsigned FG = __byte_perm(df, eg, 0x7762) & (m8 * 0x0101u);     // f | g << 8
        const unsigned own = (whA | (whB << 4)) & (m8 * 0x010101u);           // a | b << 8 | c << 16
        unsigned in = __shfl_down_sync(0xffffffffu, send, 1);
        if (lane == 31) in = 0;
        const unsigned FGr = (((FG << 1) & 0xeeeeu) | (in & 0x1111u)) & (KM2 * 0x0101u);
        const unsigned Fr = FGr & 0xffu, Gr = FGr >> 8;
        const unsigned cin = (in >> 16) & 0x11u;
        const unsigned C1 = ((((finC_dn << 1) | ((in >> 24) & 0x11u)) & KM2) | sent) & m8;
        const unsigned A1 = finA_dn, D_dn = DE_dn & 0xffu, E_dn = DE_dn >> 8;
        const unsigned ownA = own & 0xffu, ownB = (own >> 8) & 0xffu, ownC = own >> 16;
        const unsigned K0 = (A1 & E_dn) | C1;
        const unsigned Gg = ownB & K0, Pp = (ownB & Gr) | Fr;
        const unsigned x = Gg | Pp;
        const unsigned cvec = (x + Gg + cin) ^ x ^ Gg;
        const unsigned finB = (cvec >> 1) & m8;
        const unsigned keepm = K0 | (cvec & Gr);
        const unsigned finA = (ownA & keepm) | (A1 & D_dn);
        const unsigned finC = ownC & keepm;
        send = ((FG | (finB << 16) | (finC_dn << 24)) >> (K - 1)) & 0x11111111u;
        finA_dn = finA; finC_dn = finC; DE_dn = DE;
        return finA | (finB << 8) | (finC << 16);
    }
};

// tasks: (pair = first pair, strip = second pair or -1), the list k2f_x2 consumed
template <int K>
__global__ void __launch_bounds__(128, 6) k2r_x2(const Params p) {
    typedef Rev2<K> R;
    enum { KMASK = R::KMASK };
    const int lane = threadIdx.x & 31;
    const unsigned local = p.is_global ? 0u : 1u;
    R r;
    r.lane = lane;
    for (;;) {
        unsigned tsk = 0;
        if (lane == 0) tsk = atomicAdd(p.counter_r, 1u);
        tsk = __shfl_sync(0xffffffffu, tsk, 0);
        if (tsk >= (unsigned)p.task_count) break;
        const StripTask task = p.tasks[tsk];
        const PairInfo pa = p.pairs[task.pair];
        const PairInfo pb = p.pairs[task.strip >= 0 ? task.strip : task.pair];
        const int M = pa.M, nblk = pa.nblk;
        const int j0 = lane * K;
        unsigned colmask = 0, bitN = 0;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int N = half ? pb.N : pa.N;
            const int nreal = max(0, min(K, N - j0));
            const unsigned cm = (unsigned)(KMASK & ~((1 << (K - nreal)) - 1));
            const unsigned bn = (N > j0 && N <= j0 + K) ? (1u << (K - 1 - (N - j0 - 1))) : 0u;
            colmask |= cm << (4 * half);
            bitN |= bn << (4 * half);
        }
        r.M = M;
        r.colmask = colmask;
        r.sent_last = local ? colmask : bitN;
        r.sent_rows = local ? bitN : 0u;
        r.finA_dn = 0; r.finC_dn = 0; r.DE_dn = 0; r.send = 0;
        const uint4* sloA = p.lo + pa.dir_off + lane;
        uint4* shiA = p.hi + pa.dir_off + lane;
        const uint4* sloB = p.lo + pb.dir_off + lane;
        uint4* shiB = p.hi + pb.dir_off + lane;
        const unsigned split = KMASK * 0x010101u;
        uint4 nlA = sloA[(int64_t)(nblk - 1) * 32], nhA = shiA[(int64_t)(nblk - 1) * 32];
        uint4 nlB = sloB[(int64_t)(nblk - 1) * 32], nhB = shiB[(int64_t)(nblk - 1) * 32];
        for (int tb = nblk - 1; tb >= 0; --tb) {
            const int t0 = tb * FSTEPS + 1, thi = t0 + FSTEPS - 1;
            const uint4 lA = nlA, hA = nhA, lB = nlB, hB = nhB;
            if (tb > 0) {
                nlA = sloA[(int64_t)(tb - 1) * 32]; nhA = shiA[(int64_t)(tb - 1) * 32];
                nlB = sloB[(int64_t)(tb - 1) * 32]; nhB = shiB[(int64_t)(tb - 1) * 32];
            }
            uint4 f;
            if (t0 >= 32 && thi < M) {
                f.w = r.template step<false>(t0 + 3, lA.w, hA.w, lB.w, hB.w);
                f.z = r.template step<false>(t0 + 2, lA.z, hA.z, lB.z, hB.z);
                f.y = r.template step<false>(t0 + 1, lA.y, hA.y, lB.y, hB.y);
                f.x = r.template step<false>(t0 + 0, lA.x, hA.x, lB.x, hB.x);
            } else {
                f.w = r.template step<true>(t0 + 3, lA.w, hA.w, lB.w, hB.w);
                f.z = r.template step<true>(t0 + 2, lA.z, hA.z, lB.z, hB.z);
                f.y = r.template step<true>(t0 + 1, lA.y, hA.y, lB.y, hB.y);
                f.x = r.template step<true>(t0 + 0, lA.x, hA.x, lB.x, hB.x);
            }
            shiA[(int64_t)tb * 32] = make_uint4(f.x & split, f.y & split, f.z & split, f.w & split);
            if (task.strip >= 0)
                shiB[(int64_t)tb * 32] = make_uint4((f.x >> 4) & split, (f.y >> 4) & split, (f.z >> 4) & split, (f.w >> 4) & split);
        }
    }
}

// -------------------------------------------------------------------------------------------------------
// traceback
// -------------------------------------------------------------------------------------------------------
struct WalkParams {
    const PairInfo* pairs;
    int32_t pair_first, pair_count;
    const uint32_t* hi;         // final a/b/c plane viewed as uint32
    const int32_t* best;
    const int32_t* start_i;
    const int32_t* start_j;
    uint32_t* ops;
    int32_t* nops;
    int32_t* i0;
    int32_t* j0;
    int32_t* out_len;
    int32_t* score;             // -best, or INT_MIN when the traceback failed (_gotoh2.c:403-407)
};

__global__ void __launch_bounds__(128) k2f_walk(const WalkParams p) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p.pair_count) return;
    const int pi = p.pair_first + idx;
    const PairInfo pr = p.pairs[pi];
    const int K = pr.K, nblk = pr.nblk;
    int i = p.start_i[pi], j = p.start_j[pi];
    const int right = (i == pr.M && j < pr.N) ? (pr.N - j) : (pr.M - i);
    uint32_t* ops = p.ops + pr.ops_off;
    uint32_t cur = 0;
    int n = 0;
    bool failed = false;
    const uint4* hi4 = reinterpret_cast<const uint4*>(p.hi);
    int64_t cached_at = -1;
    uint4 cached = make_uint4(0, 0, 0, 0);
    while (i > 0 && j > 0) {
        const int jj = j - 1, strip = jj / (32 * K), rem = jj - strip * 32 * K, lane = rem / K, k = rem - lane * K;
        const int tt = i + lane - 1, tb = tt >> 2, s = tt & 3;
        // one uint4 holds 4 consecutive rows of this lane's K columns: a diagonal run stays inside it for ~4 cells
        const int64_t at = pr.dir_off + ((int64_t)strip * nblk + tb) * 32 + lane;
        if (at != cached_at) { cached = hi4[at]; cached_at = at; }
        const uint32_t w = s == 0 ? cached.x : s == 1 ? cached.y : s == 2 ? cached.z : cached.w;
        const int bit = K - 1 - k;
        uint32_t d;
        if ((w >> bit) & 1u) { d = DIR_UP; --i; }                         // vertical first (_gotoh2.c:381-388)
        else if ((w >> (8 + bit)) & 1u) { d = DIR_LEFT; --j; }            // then horizontal (:389-395)
        else if ((w >> (16 + bit)) & 1u) { d = DIR_DIAG; --i; --j; }      // then diagonal (:396-402)
        else { failed = true; break; }                                    // "traceback failed" (:403-407)
        cur |= d << (2 * (n & 15));
        if ((n & 15) == 15) { ops[n >> 4] = cur; cur = 0; }
        ++n;
    }
    if (n & 15) ops[n >> 4] = cur;
    const int k = i > j ? i : j;
    p.nops[pi] = n;
    p.i0[pi] = failed ? 0 : i;
    p.j0[pi] = failed ? 0 : j;
    p.out_len[pi] = failed ? 0 : k + n + right;
    p.score[pi] = failed ? (int)0x80000000 : -p.best[pi];
}

// Warp-per-pair traceback for long paths: a dependent pointer chase pays one memory latency per step when every
// step loads its own word (8 ms for a 9.6 kb x 9.6 kb pair, whatever the batch size).  Here the 32 lanes prefetch a
// tile of the final a/b/c plane around the current cell - 4 column groups (lanes of the forward wavefront) x 8 blocks
// of 4 rows - and all lanes walk the path redundantly, fetching each step's word from the tile by shuffle; the tile is
// refilled only when the path leaves it (~every 25-30 steps on a diagonal).
template <int K>
__device__ __forceinline__ void walk_warp(const WalkParams& p, const int pi, const PairInfo& pr) {
    const int lane = threadIdx.x & 31;
    const int nblk = pr.nblk;
    int i = p.start_i[pi], j = p.start_j[pi];
    const int right = (i == pr.M && j < pr.N) ? (pr.N - j) : (pr.M - i);
    uint32_t* ops = p.ops + pr.ops_off;
    const uint4* hi4 = reinterpret_cast<const uint4*>(p.hi) + pr.dir_off;
    uint32_t cur = 0;
    int n = 0;
    bool failed = false;
    const int dl = lane & 3, db = lane >> 2;      // my tile element: column group G0 - dl, block (top of that group) - db
    int G0 = -1000000, i0 = 0;                    // tile anchor
    uint4 mine = make_uint4(0, 0, 0, 0);
    while (i > 0 && j > 0) {
        const int G = (j - 1) / K, k = (j - 1) - G * K;
        const int L = G & 31, tt = i + L - 1, tb = tt >> 2, s = tt & 3;
        int ddl = G0 - G;
        int ddb = ((i0 + L - 1) >> 2) - tb;
        if (ddl < 0 || ddl > 3 || ddb < 0 || ddb > 7) {
            // refill: anchor the tile at the current cell
            G0 = G; i0 = i;
            const int Gm = G0 - dl;
            if (Gm >= 0) {
                const int Lm = Gm & 31, sm = Gm >> 5;
                const int tbm = ((i0 + Lm - 1) >> 2) - db;
                if (tbm >= 0) mine = hi4[((int64_t)sm * nblk + tbm) * 32 + Lm];
            }
            ddl = 0; ddb = 0;
        }
        const int src = ddb * 4 + ddl;
        const unsigned wx = __shfl_sync(0xffffffffu, mine.x, src), wy = __shfl_sync(0xffffffffu, mine.y, src);
        const unsigned wz = __shfl_sync(0xffffffffu, mine.z, src), ww = __shfl_sync(0xffffffffu, mine.w, src);
        const uint32_t w = s == 0 ? wx : s == 1 ? wy : s == 2 ? wz : ww;
        const int bit = K - 1 - k;
        uint32_t d;
        if ((w >> bit) & 1u) { d = DIR_UP; --i; }                         // vertical first (_gotoh2.c:381-388)
        else if ((w >> (8 + bit)) & 1u) { d = DIR_LEFT; --j; }            // then horizontal (:389-395)
        else if ((w >> (16 + bit)) & 1u) { d = DIR_DIAG; --i; --j; }      // then diagonal (:396-402)
        else { failed = true; break; }                                    // "traceback failed" (:403-407)
        cur |= d << (2 * (n & 15));
        if ((n & 15) == 15) { if (lane == 0) ops[n >> 4] = cur; cur = 0; }
        ++n;
    }
    if (lane == 0) {
        if (n & 15) ops[n >> 4] = cur;
        const int kk = i > j ? i : j;
        p.nops[pi] = n;
        p.i0[pi] = failed ? 0 : i;
        p.j0[pi] = failed ? 0 : j;
        p.out_len[pi] = failed ? 0 : kk + n + right;
        p.score[pi] = failed ? (int)0x80000000 : -p.best[pi];
    }
}

__global__ void __launch_bounds__(128) k2f_walk_warp(const WalkParams p) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= p.pair_count) return;
    const int pi = p.pair_first + w;
    const PairInfo pr = p.pairs[pi];
    switch (pr.K) {
        case 2: walk_warp<2>(p, pi, pr); break;
        case 3: walk_warp<3>(p, pi, pr); break;
        case 4: walk_warp<4>(p, pi, pr); break;
        case 6: walk_warp<6>(p, pi, pr); break;
        default: walk_warp<8>(p, pi, pr); break;
    }
}

}  // namespace g2f
}  // namespace gotoh

// gotoh2_kernels.cuh - sm_100a kernels for MiCall-Lite's LIVE aligner, `_gotoh2.align`
// ("next" row #1; reference: micall/alignment/src/_gotoh2.c, driven by
// micall/alignment/gotoh2.py:74-96 from core/remap.py:248 and core/aln2counts.py:187).
//
// Altschul-Erickson min-cost affine alignment with explicit tie bits:
//   k2_forward   cost_assignment (_gotoh2.c:137-201): R, p, q over the (l1+1) x (l2+1) grid, one byte of
//                tie bits per cell (a,b,c: which of p,q,diag attain R; D,E / F,G: how this cell's own p / q
//                were formed); start-cell search of the local mode (_gotoh2.c:327-352) fused in.
//   k2_reverse   edge_assignment, steps 8-11 (_gotoh2.c:205-312): the exact time reversal of k2_forward
//                over the same [strip][step][lane] words, rewriting a,b,c in place.
//   k2_walk      traceback with priority a > b > c (_gotoh2.c:374-408) into the op script k_emit consumes.
// One warp per pair; lane l owns 8 grid columns and runs one row behind lane l-1 (forward) / lane l+1
// (reverse); int32 costs; wider grids are cut into 256-column strips handed over through global memory.
// Correctness first: this path is measured but not yet tuned.
#pragma once

#include "gotoh_kernels.cuh"

namespace gotoh {
namespace g2 {

enum { G2K = 8, G2_INF = 1 << 29 };
enum { BA = 1, BB = 2, BC = 4, BD = 8, BE = 16, BF = 32, BG = 64 };

struct Params {
    const PairInfo* pairs;      // M = l1, N = l2, dir_off = first uint2 of the pair's arena, nblk = T = l1 + 1 + 31
    int32_t pair_first, pair_count;
    const uint8_t* s1_idx;      // alphabet indices of the cleaned sequences (same positions as the raw bytes)
    const uint8_t* s2_idx;
    const int32_t* dmat;        // l*l substitution scores (gotoh2.py:47-64)
    int32_t l, v, u, is_global; // v = gap open, u = gap extend (_gotoh2.c:31-32)
    uint2* arena;               // 8 tie-bit bytes per lane-step
    int2* fbnd;                 // forward strip boundary: (R, q) per row, per resident warp
    uint8_t* rbnd;              // reverse strip boundary: final abc | F,G of a strip's first column, per row
    int64_t bnd_stride;
    int32_t* best;              // R at the start cell (score = -best)
    int32_t* start_i;
    int32_t* start_j;
};

// ---------------------------------------------------------------------------------------------------
template <int DUMMY>
__global__ void __launch_bounds__(128) k2_forward(const Params p) {
    __shared__ int s_d[32 * 32];
    for (int x = threadIdx.x; x < p.l * p.l; x += blockDim.x) s_d[x] = p.dmat[x];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int wglobal = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int v = p.v, u = p.u, l = p.l;
    int2* bnd = p.fbnd + (int64_t)wglobal * p.bnd_stride;

    for (int idx = wglobal; idx < p.pair_count; idx += nwarps) {
        const PairInfo pr = p.pairs[p.pair_first + idx];
        const int l1 = pr.M, l2 = pr.N, nrows = l1 + 1, ncols = l2 + 1, T = pr.nblk;
        const int nstrips = (ncols + 32 * G2K - 1) / (32 * G2K);
        const uint8_t* s1 = p.s1_idx + pr.ref_pos;
        const uint8_t* s2 = p.s2_idx + pr.qry_pos;
        // start-cell search state (local mode): first strict minimum of the last column, then of the last row
        int col_min = 2147483647, col_i = 0, row_min = 2147483647, row_j = 0, r_ll = 0;

        for (int strip = 0; strip < nstrips; ++strip) {
            const int c0 = (strip * 32 + lane) * G2K;
            const bool last_strip = (strip == nstrips - 1);
            int bq[G2K], Rup[G2K], Pup[G2K];
#pragma unroll
            for (int k = 0; k < G2K; ++k) {
                const int c = c0 + k;
                bq[k] = (c >= 1 && c <= l2) ? (int)s2[c - 1] : 0;
                Rup[k] = 0; Pup[k] = G2_INF;
            }
            int sendR = 0, sendQ = G2_INF, Rd_in = 0;
            __syncwarp();
            for (int t = 0; t < T; ++t) {
                const int r = t - lane;
                int Rl = __shfl_up_sync(0xffffffffu, sendR, 1);
                int Ql = __shfl_up_sync(0xffffffffu, sendQ, 1);
                if (lane == 0 && strip > 0) {
                    const int rr = min(max(r, 0), nrows - 1);
                    const int2 b = bnd[rr];
                    Rl = b.x; Ql = b.y;
                }
                const bool top = (r == 0);
                const int arow = (r >= 1 && r <= l1) ? (int)s1[r - 1] * l : 0;
                int Rdiag = Rd_in, Rleft = Rl, qleft = Ql;
                Rd_in = Rl;
                unsigned lo = 0, hi = 0;
#pragma unroll
                for (int k = 0; k < G2K; ++k) {
                    const bool leftmost = (c0 + k == 0);
                    int pv = G2_INF, qv = G2_INF, rv, dg = G2_INF;
                    unsigned b = 0;
                    if (!top) {                                          // _gotoh2.c:155-163
                        pv = u + min(Pup[k], Rup[k] + v);
                        if (Pup[k] < G2_INF / 2 && pv == Pup[k] + u) b |= BD;
                        if (pv == Rup[k] + v + u) b |= BE;
                    }
                    if (!leftmost) {                                     // _gotoh2.c:165-173
                        qv = u + min(qleft, Rleft + v);
                        if (qleft < G2_INF / 2 && qv == qleft + u) b |= BF;
                        if (qv == Rleft + v + u) b |= BG;
                    }
                    if (top || leftmost) {                               // _gotoh2.c:175-183
                        rv = (top && leftmost) ? 0 : (p.is_global ? min(pv, qv) : 0);
                    } else {                                             // _gotoh2.c:185-187
                        dg = Rdiag - s_d[arow + bq[k]];
                        rv = min(min(dg, pv), qv);
                        if (rv == dg) b |= BC;
                    }
                    if (rv == pv) b |= BA;                               // _gotoh2.c:190-195
                    if (rv == qv) b |= BB;
                    Rdiag = Rup[k];
                    Rup[k] = rv; Pup[k] = pv; Rleft = rv; qleft = qv;
                    if (k < 4) lo |= b << (8 * k); else hi |= b << (8 * (k - 4));
                }
                sendR = Rleft; sendQ = qleft;
                p.arena[pr.dir_off + ((int64_t)strip * T + t) * 32 + lane] = make_uint2(lo, hi);
                const bool valid = (r >= 0 && r < nrows);
                if (!last_strip && lane == 31 && valid) bnd[r] = make_int2(Rleft, qleft);
                if (last_strip && valid) {
                    // the grid's last column l2 sits at (lane_n, k_n) of this strip
                    const int cl = l2 - c0;
                    if (cl >= 0 && cl < G2K) {
                        int val = Rup[0];
#pragma unroll
                        for (int k = 1; k < G2K; ++k) if (k == cl) val = Rup[k];
                        if (val < col_min) { col_min = val; col_i = r; }      // first strict minimum, _gotoh2.c:330-339
                        if (r == l1) r_ll = val;
                    }
                }
                if (valid && r == l1) {
#pragma unroll
                    for (int k = 0; k < G2K; ++k)
                        if (c0 + k <= l2 && Rup[k] < row_min) { row_min = Rup[k]; row_j = c0 + k; }   // _gotoh2.c:341-350
                }
            }
        }
        // combine: candidates are visited in the order (l1,l2), column top-down, row left-to-right with '<'
        const int owner = ((l2) % (32 * G2K)) / G2K;
        col_min = __shfl_sync(0xffffffffu, col_min, owner);
        col_i = __shfl_sync(0xffffffffu, col_i, owner);
        r_ll = __shfl_sync(0xffffffffu, r_ll, owner);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const int om = __shfl_xor_sync(0xffffffffu, row_min, off);
            const int oj = __shfl_xor_sync(0xffffffffu, row_j, off);
            if (om < row_min || (om == row_min && oj < row_j)) { row_min = om; row_j = oj; }
        }
        if (lane == 0) {
            int best = r_ll, bi = l1, bj = l2;
            if (!p.is_global) {
                if (col_min < best) { best = col_min; bi = col_i; bj = l2; }
                if (row_min < best) { best = row_min; bi = l1; bj = row_j; }
            }
            p.best[p.pair_first + idx] = best;
            p.start_i[p.pair_first + idx] = bi;
            p.start_j[p.pair_first + idx] = bj;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// edge_assignment (_gotoh2.c:205-312).  Only the final a,b,c bits are kept: the d,e,f,g bits that steps
// 10/11 rewrite are never read again (they are written after their last read, :258-308).
__device__ __forceinline__ unsigned edge_abc(unsigned own, unsigned A1, unsigned B1, unsigned C1,
                                             unsigned below, unsigned right) {
    unsigned abc = own & (BA | BB | BC);
    const bool d0 = below & BD, e0 = below & BE, f0 = right & BF, g0 = right & BG;
    if ((!A1 || !e0) && (!B1 || !g0) && !C1) abc = 0;       // step 8  :237-242
    if (A1 || B1 || C1) {                                   // step 9  :245
        if (A1 && d0) abc |= BA;                            // step 10 :251-270
        if (B1 && f0) abc |= BB;                            // step 11 :283-300
    }
    return abc;
}

template <int DUMMY>
__global__ void __launch_bounds__(128) k2_reverse(const Params p) {
    const int lane = threadIdx.x & 31;
    const int wglobal = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    uint8_t* rb = p.rbnd + (int64_t)wglobal * p.bnd_stride;

    for (int idx = wglobal; idx < p.pair_count; idx += nwarps) {
        const PairInfo pr = p.pairs[p.pair_first + idx];
        const int l1 = pr.M, l2 = pr.N, nrows = l1 + 1, ncols = l2 + 1, T = pr.nblk;
        const int nstrips = (ncols + 32 * G2K - 1) / (32 * G2K);
        const unsigned local = p.is_global ? 0u : 1u;

        for (int strip = nstrips - 1; strip >= 0; --strip) {
            const int c0 = (strip * 32 + lane) * G2K;
            unsigned fin_dn[G2K], byte_dn[G2K];    // final abc and forward byte of the cell below (row r+1)
#pragma unroll
            for (int k = 0; k < G2K; ++k) { fin_dn[k] = 0; byte_dn[k] = 0; }
            unsigned send_fin = 0, send_byte = 0, send_cdn = 0;   // what lane-1 needs next step
            __syncwarp();
            for (int t = T - 1; t >= 0; --t) {
                const int r = t - lane;
                const bool valid = (r >= 0 && r < nrows);
                uint2* wp = p.arena + pr.dir_off + ((int64_t)strip * T + t) * 32 + lane;
                const uint2 w = *wp;
                unsigned byt[G2K];
#pragma unroll
                for (int k = 0; k < G2K; ++k) byt[k] = ((k < 4 ? w.x : w.y) >> (8 * (k & 3))) & 0xffu;
                // right neighbour column c0+8 at row r (final abc, forward byte) and at row r+1 (final c)
                unsigned rfin = __shfl_down_sync(0xffffffffu, send_fin, 1);
                unsigned rbyte = __shfl_down_sync(0xffffffffu, send_byte, 1);
                unsigned rcdn = __shfl_down_sync(0xffffffffu, send_cdn, 1);
                if (lane == 31) {
                    rfin = 0; rbyte = 0; rcdn = 0;
                    if (strip < nstrips - 1 && valid) {
                        const unsigned x = rb[r];
                        rfin = x & 7u; rbyte = x & (BF | BG);
                        rcdn = (r + 1 < nrows) ? (rb[r + 1] & BC) : 0u;
                    }
                }
                unsigned fin[G2K];
                const unsigned my_cdn0 = fin_dn[0] & BC;       // final c of (r+1, c0): lane-1's diagonal next step
#pragma unroll
                for (int k = G2K - 1; k >= 0; --k) {
                    const int c = c0 + k;
                    fin[k] = 0;
                    if (valid && c <= l2) {
                        const bool last_r = (r == l1), last_c = (c == l2);
                        const unsigned A1 = last_r ? 0u : (fin_dn[k] & BA);
                        const unsigned below = last_r ? 0u : byte_dn[k];
                        unsigned B1, right, C1;
                        if (last_c) { B1 = 0; right = 0; }
                        else if (k < G2K - 1) { B1 = fin[k + 1] & BB; right = byt[k + 1]; }
                        else { B1 = rfin & BB; right = rbyte; }
                        if (last_r || last_c) C1 = (local || (last_r && last_c)) ? 1u : 0u;   // sentinel border :118-133
                        else if (k < G2K - 1) C1 = fin_dn[k + 1] & BC;
                        else C1 = rcdn;
                        fin[k] = edge_abc(byt[k], A1, B1, C1, below, right);
                    }
                }
                unsigned lo = 0, hi = 0;
#pragma unroll
                for (int k = 0; k < G2K; ++k) {
                    const unsigned nb = (byt[k] & ~7u) | fin[k];
                    if (k < 4) lo |= nb << (8 * k); else hi |= nb << (8 * (k - 4));
                    fin_dn[k] = fin[k];
                    byte_dn[k] = byt[k];
                }
                *wp = make_uint2(lo, hi);
                send_fin = fin[0];
                send_byte = byt[0] & (BF | BG);
                send_cdn = my_cdn0;
                if (strip > 0 && lane == 0 && valid) rb[r] = (uint8_t)(fin[0] | (byt[0] & (BF | BG)));
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
struct WalkParams2 {
    const PairInfo* pairs;
    int32_t pair_first, pair_count;
    const uint8_t* arena;       // byte view
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

__global__ void __launch_bounds__(128) k2_walk(const WalkParams2 p) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p.pair_count) return;
    const int pi = p.pair_first + idx;
    const PairInfo pr = p.pairs[pi];
    const int T = pr.nblk;
    int i = p.start_i[pi], j = p.start_j[pi];
    const int right = (i == pr.M && j < pr.N) ? (pr.N - j) : (pr.M - i);
    uint32_t* ops = p.ops + pr.ops_off;
    uint32_t cur = 0;
    int n = 0;
    bool failed = false;
    while (i > 0 && j > 0) {
        const int strip = j / (32 * G2K), rem = j - strip * 32 * G2K, lane = rem / G2K, k = rem - lane * G2K;
        const int64_t byte_idx = ((pr.dir_off + ((int64_t)strip * T + (i + lane)) * 32 + lane) << 3) + k;
        const unsigned b = p.arena[byte_idx];
        uint32_t d;
        if (b & BA) { d = DIR_UP; --i; }                    // vertical first (_gotoh2.c:381-388)
        else if (b & BB) { d = DIR_LEFT; --j; }             // then horizontal (:389-395)
        else if (b & BC) { d = DIR_DIAG; --i; --j; }        // then diagonal (:396-402)
        else { failed = true; break; }                      // "traceback failed" (:403-407)
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

}  // namespace g2
}  // namespace gotoh

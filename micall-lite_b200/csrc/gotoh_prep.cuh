// gotoh_prep.cuh - device-side plan builder for the common batch shape (many short queries against a few shared
// references): trim + validation (gotoh.cpp:545-559, :216-219 domain), int16x2 admission (the same range proof the
// host applies, gotoh_plan_math.h), grouping into warp tasks and the HBM layout - everything plan_build() does per pair,
// as five small kernels, so that the host's work per call is O(#references) and 1 M 84-aa windows cost H2D + kernels +
// D2H instead of ~0.3 us of one host core each.
//
//   k_prep_scan    thread per pair: trim span, byte domain, byte-presence mask, M, N
//   k_prep_range   one CTA: score range (minT, maxT) of the bytes present against the classes of the references
//   k_prep_bins    thread per pair: frame-shift need, worst min(M,N) per K, histogram over (kernel, reference, N) bins
//   k_prep_layout  one CTA: frame shift z4, rebase period R, admission of the worst pairs, prefix sums over the bins,
//                  groups -> tasks / arena / op-script bases, launch list; summary to mapped pinned memory
//   k_prep_scatter thread per pair: position inside its bin (atomic), PairInfo, task entries (couples, half-warp quads,
//                  fillers)
// Pairs of one bin have identical shapes, so the order inside a bin is irrelevant (results are per pair); the bins are
// laid out in the order of the host builder's sort key (kernel, M descending, reference, N descending).
// Anything this path does not cover - a byte outside the domain, an empty or too long query, a pair that fails the
// int16 proof, degapping, more than PREP_MAX_REFS references - raises `fallback` and the host builder (plan_build)
// runs instead and reports errors exactly as before.
#pragma once

#include "gotoh_kernels.cuh"
#include "gotoh_plan_math.h"

namespace gotoh {

enum { PREP_MAX_REFS = 64, PREP_NKH = 10, PREP_NSLOT = 256, PREP_MAX_LAUNCH = PREP_NKH };

struct PrepLaunch { int32_t K, hw, task_first, task_count; };

// Written by k_prep_layout into mapped pinned memory, read by the host after one event synchronisation.
struct PrepSummary {
    int32_t fallback;          // != 0: the host builder must take this slab (reason code, diagnostics only)
    int32_t z4, R;
    int32_t minT, maxT;
    int32_t n_launch;
    int64_t n_tasks, arena_u4, ops_words, cells, sum_mn;
    PrepLaunch launch[PREP_MAX_LAUNCH];
};

struct PrepParams {
    // inputs
    const uint8_t* qry;            // raw query bytes of the slab (untrimmed)
    const int64_t* qry_off;        // n+1 offsets (caller's, absolute); qry holds bytes from qry_off[0]
    const int32_t* ref_idx;        // n caller reference indices
    const int64_t* out_off;        // strided form: n+1 caller offsets (absolute); else NULL
    int32_t n;
    int32_t n_refs;                // size of the per-reference tables below (caller indices)
    const int32_t* ref_M;          // trimmed length per caller reference (0: unused)
    const int64_t* ref_pos;        // position of row 1 in d_ref_raw / d_ref_cls
    const int32_t* ref_rank;       // rank among the used references in (M descending, index) order, -1: unused
    int32_t n_used;                // used references (<= PREP_MAX_REFS)
    const int32_t* rank_M;         // M by rank
    const int32_t* table4;         // [ncls][128] of 4*(T + 2*gep) (row 0 unused), as uploaded for the forward kernels
    int32_t ncls, gip, gep, has_dollar, half_off;
    // scratch
    int32_t* pairN;                // n: trimmed length
    int32_t* pairLo;               // n: first byte of the trimmed span (relative to the pair's start)
    uint32_t* present;             // 4 words: bytes 0..127 seen in some query
    int32_t* scal;                 // [0] fallback reason, [1] max need, [2..2+8] 1 + worst min(M,N) per K, [16] minT, [17] maxT
    unsigned long long* acc;       // [0] cells, [1] sum of M+N
    uint32_t* hist;                // PREP_NKH * n_used * PREP_NSLOT bins
    uint32_t* cursor;              // same shape: next rank inside the bin (k_prep_scatter)
    int64_t* bin_pair;             // bins+1: exclusive prefix of the pair counts
    int64_t* bin_ops;              // bins+1: exclusive prefix of count * words(M, N)
    int64_t* grp;                  // per group (kh, ref rank): [0] task base, [1] arena base (uint4), [2] pairs in the group
    // outputs
    PairInfo* pairs;
    Task* tasks;
    PrepSummary* summary;          // mapped pinned
};

__device__ __forceinline__ void prep_fallback(const PrepParams& p, int why) { atomicMax(&p.scal[0], why); }

// ---- thread per pair: trim (gotoh.cpp:545-559), domain 1..126 (gotoh.cpp:216-219), presence ------------------------
__global__ void __launch_bounds__(256) k_prep_scan(const PrepParams p) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t m0 = 0, m1 = 0, m2 = 0, m3 = 0;
    if (k < p.n) {
        const int64_t a = p.qry_off[k] - p.qry_off[0], b = p.qry_off[k + 1] - p.qry_off[0];
        const int r = p.ref_idx[k];
        int why = 0;
        if (b < a || b - a > 4096 || r < 0 || r >= p.n_refs) why = 1;
        int lo = 0, hi = why ? 0 : (int)(b - a);
        const uint8_t* s = p.qry + a;
        if (!why) {
            while (lo < hi && plan_is_ws(s[lo])) ++lo;
            while (hi > lo && plan_is_ws(s[hi - 1])) --hi;
            for (int x = lo; x < hi; ++x) {
                const unsigned c = s[x];
                if ((uint8_t)(c - 1) > 125) { why = 2; break; }
                if (c < 32) m0 |= 1u << c; else if (c < 64) m1 |= 1u << (c - 32); else if (c < 96) m2 |= 1u << (c - 64); else m3 |= 1u << (c - 96);
            }
            const int N = hi - lo;
            if (N < 1 || N > 32 * PLAN_MAX_K) why = why ? why : 3;
            const int M = why ? 0 : p.ref_M[r];
            if (!why && (M < 1 || p.ref_rank[r] < 0)) why = 4;
            // the -100000 sentinel domain (gotoh.cpp:284-286)
            if (!why && 2LL * p.gip + ((long long)(M > N ? M : N) + 1) * p.gep >= 100000) why = 5;
            if (!why && p.out_off) {
                const int64_t st = p.out_off[k + 1] - p.out_off[k];
                if (st < (int64_t)M + N || st > 0x7fffffffLL) why = 6;
            }
        }
        p.pairN[k] = why ? 0 : hi - lo;
        p.pairLo[k] = lo;
        if (why) prep_fallback(p, why);
    }
    // presence: one atomicOr per warp and word
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        m0 |= __shfl_xor_sync(0xffffffffu, m0, off); m1 |= __shfl_xor_sync(0xffffffffu, m1, off);
        m2 |= __shfl_xor_sync(0xffffffffu, m2, off); m3 |= __shfl_xor_sync(0xffffffffu, m3, off);
    }
    if ((threadIdx.x & 31) == 0) {
        if (m0) atomicOr(&p.present[0], m0);
        if (m1) atomicOr(&p.present[1], m1);
        if (m2) atomicOr(&p.present[2], m2);
        if (m3) atomicOr(&p.present[3], m3);
    }
}

// ---- one CTA: score range of the bytes present (what plan_build derives from qry_present and the class table) --------
__global__ void __launch_bounds__(256) k_prep_range(const PrepParams p) {
    __shared__ int s_min[256], s_max[256];
    int mn = 0, mx = 0;
    for (int x = threadIdx.x; x < (p.ncls - 1) * 126; x += blockDim.x) {
        const int c = 1 + x / 126, b = 1 + x % 126;
        if ((p.present[b >> 5] >> (b & 31)) & 1u) {
            const int t = p.table4[c * 128 + b] / 4 - 2 * p.gep;       // the entry is 4*(T + 2*gep)
            mn = t < mn ? t : mn;
            mx = t > mx ? t : mx;
        }
    }
    s_min[threadIdx.x] = mn; s_max[threadIdx.x] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int x = 1; x < (int)blockDim.x; ++x) { mn = s_min[x] < mn ? s_min[x] : mn; mx = s_max[x] > mx ? s_max[x] : mx; }
        if (p.has_dollar) mx += 18;                                     // up to three +6 stop-codon bonuses on one cell
        p.scal[16] = mn; p.scal[17] = mx;
    }
}

__device__ __forceinline__ int prep_kh_rank(int KH) {
    // ascending order of the host builder's kernel selector: K (32-lane wavefronts), then 16|K (half-warp wavefronts)
    switch (KH) {
        case 2: return 0; case 3: return 1; case 4: return 2; case 6: return 3; case 8: return 4;
        case 18: return 5; case 19: return 6; case 20: return 7; case 22: return 8; default: return 9;
    }
}
__device__ __forceinline__ int prep_bin(const PrepParams& p, int KH, int rank, int N) {
    return (prep_kh_rank(KH) * p.n_used + rank) * PREP_NSLOT + (PREP_NSLOT - N);      // N descending inside the group
}

// Warp-aggregated counter increment: the lanes of a warp that hit the same counter elect a leader, which adds their
// number once; every lane gets its own rank.  (One atomic per thread on a handful of hot bins - a batch has a few
// references and query lengths - serialised 300 k threads: k_prep_bins took 570 us, a quarter of the C3 forward kernel.)
__device__ __forceinline__ unsigned prep_count_in(uint32_t* counters, int bin, bool active) {
    const unsigned lane = threadIdx.x & 31;
    const unsigned peers = __match_any_sync(0xffffffffu, active ? bin : -1);
    unsigned base = 0;
    const int leader = __ffs(peers) - 1;
    if (active && (int)lane == leader) base = atomicAdd(&counters[bin], (uint32_t)__popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + __popc(peers & ((1u << lane) - 1u));
}

// ---- thread per pair: frame-shift need, worst pair per K, histogram ------------------------------------------------------
__global__ void __launch_bounds__(256) k_prep_bins(const PrepParams p) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int N = k < p.n ? p.pairN[k] : 0;
    bool active = N >= 1;                                               // (inactive: beyond n, or already flagged)
    int r = 0, M = 0, K = 0, bin = 0;
    long long need = 0;
    if (active) {
        r = p.ref_idx[k];
        M = p.ref_M[r];
        need = plan_int16_low_need(M, N, p.gip, p.gep, p.scal[16]);
        if (need > 32000) { prep_fallback(p, 7); active = false; }
    }
    if (active) {
        K = plan_pick_K(N);
        bin = prep_bin(p, plan_pick_KH(N, p.half_off != 0), p.ref_rank[r], N);
    }
    (void)prep_count_in(p.hist, bin, active);
    // per-warp reductions, one atomic per warp and quantity
    int v_need = active ? (int)need : 0;
    unsigned long long cells = active ? (unsigned long long)M * (unsigned long long)N : 0ull, mn_sum = active ? (unsigned long long)(M + N) : 0ull;
    int worst[5];
    const int KS[5] = {2, 3, 4, 6, 8};
#pragma unroll
    for (int x = 0; x < 5; ++x) worst[x] = (active && K == KS[x]) ? (M < N ? M : N) + 1 : 0;    // + 1: the scratch block starts zeroed, 0 = no pair with this K
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        v_need = max(v_need, __shfl_xor_sync(0xffffffffu, v_need, off));
        cells += __shfl_xor_sync(0xffffffffu, cells, off);
        mn_sum += __shfl_xor_sync(0xffffffffu, mn_sum, off);
#pragma unroll
        for (int x = 0; x < 5; ++x) worst[x] = max(worst[x], __shfl_xor_sync(0xffffffffu, worst[x], off));
    }
    if (lane == 0) {
        if (v_need) atomicMax(&p.scal[1], v_need);
#pragma unroll
        for (int x = 0; x < 5; ++x) if (worst[x]) atomicMax(&p.scal[2 + KS[x]], worst[x]);
        if (cells) { atomicAdd(&p.acc[0], cells); atomicAdd(&p.acc[1], mn_sum); }
    }
}

// ---- one CTA: z4, R, admission, prefix sums over the bins, groups, launches ----------------------------------------------
__global__ void __launch_bounds__(256) k_prep_layout(const PrepParams p) {
    __shared__ long long s_cnt[256], s_ops[256];
    __shared__ int s_R, s_fb;
    const int nbins = PREP_NKH * p.n_used * PREP_NSLOT;
    const int tid = threadIdx.x;
    if (tid == 0) {
        const int minT = p.scal[16], maxT = p.scal[17];
        const long long z4 = p.scal[1];
        int fb = p.scal[0];
        // every pair must pass the proof at the smallest rebase period with the plan-wide shift (the host tests each pair;
        // the proof depends on a pair through min(M,N) and K only and is monotone in min(M,N))
        for (int K = 0; K <= PLAN_MAX_K && !fb; ++K) {
            const int w = p.scal[2 + K] - 1;
            if (w >= 0 && !plan_fits_int16(w, w, K, 32, p.gip, p.gep, minT, maxT, z4)) fb = 8;
        }
        int R = 32;
        for (int cand = 4096; cand >= 32 && !fb; cand >>= 1) {
            bool all = true;
            for (int K = 0; K <= PLAN_MAX_K && all; ++K) {
                const int w = p.scal[2 + K] - 1;
                if (w >= 0 && !plan_fits_int16(w, w, K, cand, p.gip, p.gep, minT, maxT, z4)) all = false;
            }
            if (all) { R = cand; break; }
        }
        s_R = R; s_fb = fb;
    }
    __syncthreads();
    // exclusive prefix sums over the bins: pairs, and op-script words (words of a bin's pairs: (M + N + 15) / 16)
    const int per = (nbins + 255) / 256;
    const int b0 = tid * per, b1 = min(nbins, b0 + per);
    long long c = 0, o = 0;
    for (int b = b0; b < b1; ++b) {
        const long long h = p.hist[b];
        if (h) {
            const int rank = (b / PREP_NSLOT) % p.n_used, N = PREP_NSLOT - (b % PREP_NSLOT);
            c += h; o += h * ((p.rank_M[rank] + N + 15) >> 4);
        }
    }
    s_cnt[tid] = c; s_ops[tid] = o;
    __syncthreads();
    if (tid == 0) {
        long long rc = 0, ro = 0;
        for (int x = 0; x < 256; ++x) { const long long tc = s_cnt[x], to = s_ops[x]; s_cnt[x] = rc; s_ops[x] = ro; rc += tc; ro += to; }
        p.bin_pair[nbins] = rc; p.bin_ops[nbins] = ro;
    }
    __syncthreads();
    c = s_cnt[tid]; o = s_ops[tid];
    for (int b = b0; b < b1; ++b) {
        p.bin_pair[b] = c; p.bin_ops[b] = o;
        const long long h = p.hist[b];
        if (h) {
            const int rank = (b / PREP_NSLOT) % p.n_used, N = PREP_NSLOT - (b % PREP_NSLOT);
            c += h; o += h * ((p.rank_M[rank] + N + 15) >> 4);
        }
    }
    __syncthreads();
    if (tid != 0) return;
    // groups in key order -> task and arena bases, one launch per kernel selector with work
    const int KH_OF[PREP_NKH] = {2, 3, 4, 6, 8, 18, 19, 20, 22, 24};
    PrepSummary* s = p.summary;
    long long tasks = 0, arena = 0;
    int nl = 0;
    for (int kh = 0; kh < PREP_NKH; ++kh) {
        const int K = KH_OF[kh] & 15, hw = KH_OF[kh] >> 4;
        const long long t_first = tasks;
        for (int rank = 0; rank < p.n_used; ++rank) {
            const int g = kh * p.n_used + rank;
            const long long cnt = p.bin_pair[(g + 1) * PREP_NSLOT] - p.bin_pair[g * PREP_NSLOT];
            p.grp[3 * g + 0] = tasks; p.grp[3 * g + 1] = arena; p.grp[3 * g + 2] = cnt;
            if (!cnt) continue;
            const int M = p.rank_M[rank];
            const long long nblk = (M + (hw ? 15 : 31) + 3) / 4;          // wavefront fill: 31 rows, 15 on half-warp wavefronts
            const long long couples = (cnt + 1) >> 1;
            const long long warps = hw ? (couples + 1) >> 1 : couples;
            tasks += hw ? 2 * warps : couples;
            arena += warps * nblk * 32;
        }
        if (tasks > t_first && nl < PREP_MAX_LAUNCH) {
            s->launch[nl].K = K; s->launch[nl].hw = hw;
            s->launch[nl].task_first = (int)t_first;
            s->launch[nl].task_count = (int)(tasks - t_first);
            ++nl;
        }
    }
    int fb = s_fb;
    if (!fb && (tasks > 0x3fffffffLL || p.bin_ops[nbins] > 0x7fffffffLL)) fb = 9;
    s->z4 = p.scal[1]; s->R = s_R; s->minT = p.scal[16]; s->maxT = p.scal[17];
    s->n_launch = nl; s->n_tasks = tasks; s->arena_u4 = arena; s->ops_words = p.bin_ops[nbins];
    s->cells = (int64_t)p.acc[0]; s->sum_mn = (int64_t)p.acc[1];
    s->fallback = fb;
}

// ---- thread per pair: PairInfo and task entries ------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_prep_scatter(const PrepParams p) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = k < p.n;                                          // (every lane takes part in the warp-wide count)
    const int N = active ? p.pairN[k] : 1;
    const int r = active ? p.ref_idx[k] : 0;
    const int M = active ? p.ref_M[r] : 1, rank = active ? p.ref_rank[r] : 0;
    const int KH = plan_pick_KH(N, p.half_off != 0), K = KH & 15, hw = KH >> 4;
    const int b = active ? prep_bin(p, KH, rank, N) : 0;
    const unsigned rk = prep_count_in(p.cursor, b, active);
    if (!active) return;
    const int g = b / PREP_NSLOT;
    const long long pos = p.bin_pair[b] + rk;                              // plan-order index of this pair
    const long long q = pos - p.bin_pair[g * PREP_NSLOT];                  // index inside the group
    const long long cnt = p.grp[3 * g + 2];
    const long long couple = q >> 1;
    const int member = (int)(q & 1);
    const int nblk = (M + (hw ? 15 : 31) + 3) / 4;
    long long t, arena;
    int sub = 0;
    if (hw) {
        const long long w = couple >> 1;
        sub = (int)(couple & 1);
        t = p.grp[3 * g + 0] + 2 * w + sub;
        arena = p.grp[3 * g + 1] + w * nblk * 32;
    } else {
        t = p.grp[3 * g + 0] + couple;
        arena = p.grp[3 * g + 1] + couple * nblk * 32;
    }
    PairInfo pi;
    pi.ref_pos = p.ref_pos[r];
    pi.qry_pos = (p.qry_off[k] - p.qry_off[0]) + p.pairLo[k];
    pi.dir_off = arena;
    pi.out_off = p.out_off ? p.out_off[k] - p.out_off[0] : 0;
    pi.M = M; pi.N = N;
    pi.nblk = nblk;
    pi.ops_off = (int32_t)(p.bin_ops[b] + (long long)rk * ((M + N + 15) >> 4));
    pi.K = (int16_t)K; pi.x2 = 1;
    pi.half = (int8_t)(member | (sub << 1));
    pi.orig = k;
    pi.out_cap = p.out_off ? (int32_t)(p.out_off[k + 1] - p.out_off[k]) : M + N;
    pi.pad1 = 0;
    p.pairs[pos] = pi;
    if (member == 0) {
        Task tk;
        tk.pair_a = (int32_t)pos;
        tk.pair_b = (q + 1 < cnt) ? (int32_t)(pos + 1) : -1;
        p.tasks[t] = tk;
        // a warp of the half-warp kernels takes two entries: a lone last couple is doubled (the copy recomputes the same
        // pairs on lanes 16-31 and writes identical results)
        const long long couples = (cnt + 1) >> 1;
        if (hw && sub == 0 && couple == couples - 1) p.tasks[t + 1] = tk;
    }
}

}  // namespace gotoh

// gotoh2_host.cuh - host side of gotoh_b200_gotoh2_align_batch (the live aligner).
// Included by gotoh_b200.cu after its helpers (fail(), CU(), DevBuf, trim-free: gotoh2.py does not trim).
#pragma once

#include "gotoh2_kernels.cuh"
#include "gotoh2_fast.cuh"

namespace {

// stats of this thread's last gotoh2 call (gotoh_b200_gotoh2_last_stats)
thread_local double g2_stats[11] = {0};

struct G2Events {
    cudaEvent_t e[5] = {0, 0, 0, 0, 0};
    bool ok = false;
    bool create() {
        if (ok) return true;
        for (auto& x : e) if (cudaEventCreate(&x) != cudaSuccess) return false;
        return ok = true;
    }
    void destroy() { for (auto& x : e) if (x) { cudaEventDestroy(x); x = 0; } ok = false; }
};

struct G2Bufs {
    DevBuf<uint8_t> raw1, idx1, raw2, idx2, o1, o2, rbnd;
    DevBuf<PairInfo> pairs;
    DevBuf<int32_t> dmat, best, si, sj, nops, i0, j0, lenp, score, olen, oscore, prog, part;
    DevBuf<uint32_t> ops, counters;
    DevBuf<uint2> arena;
    DevBuf<uint4> lo, hi;
    DevBuf<int2> fbnd;
    DevBuf<gotoh::g2f::Extra> extra;
    DevBuf<gotoh::g2f::StripTask> tasks;
    ~G2Bufs() { release_all(); }
    void release_all() {
        raw1.release(); idx1.release(); raw2.release(); idx2.release(); o1.release(); o2.release(); rbnd.release();
        pairs.release(); dmat.release(); best.release(); si.release(); sj.release(); nops.release(); i0.release();
        j0.release(); lenp.release(); score.release(); olen.release(); oscore.release(); prog.release(); part.release();
        ops.release(); counters.release(); arena.release(); lo.release(); hi.release(); fbnd.release();
        extra.release(); tasks.release();
    }
};

// Per device, kept between calls: the grow-only buffers, the timing events and the SM count.  A single
// Aligner.align() call is a batch of one: it must not pay for cudaGetDeviceProperties (milliseconds), cudaMemGetInfo
// (occasionally tens of ms), event creation or cudaMalloc - measured on B200: 17.6 ms -> well under 1 ms per call.
struct G2Cache {
    std::mutex mu;
    G2Bufs bufs;
    G2Events ev;
    int sm_count = 0;
};
std::mutex g2_cache_mu;
G2Cache* g2_cache[64] = {nullptr};
G2Cache* g2_cache_for(int dev) {
    std::lock_guard<std::mutex> lk(g2_cache_mu);
    if (!g2_cache[dev]) g2_cache[dev] = new (std::nothrow) G2Cache();
    return g2_cache[dev];
}
void g2_release_cache() {
    std::lock_guard<std::mutex> lk(g2_cache_mu);
    for (int d = 0; d < 64; ++d)
        if (g2_cache[d]) { std::lock_guard<std::mutex> lk2(g2_cache[d]->mu); g2_cache[d]->bufs.release_all(); g2_cache[d]->ev.destroy(); }
}

struct G2Run {
    int64_t n = 0;
    int l = 0, gop = 0, gep = 0, is_global = 0, sm_count = 1;
    bool score_only = false;
    int64_t max_rows = 0;
    G2Events ev;
    double ms_f = 0, ms_r = 0, ms_w = 0;
    int launches = 0, chunks = 0;
    int64_t arena_bytes = 0;
    int64_t pairs_x2 = 0;           // forward tasks that ran as int16x2 couples
    const int32_t* h_dmat = nullptr;
};

// k_emit + the per-chunk timing bookkeeping shared by both kernel families
int g2_emit(G2Bufs& b, G2Run& run, int first, int count) {
    EmitParams ep;
    memset(&ep, 0, sizeof(ep));
    ep.pairs = b.pairs.p; ep.pair_first = first; ep.pair_count = count;
    ep.ref_raw = b.raw1.p; ep.qry = b.raw2.p; ep.ops = b.ops.p; ep.nops = b.nops.p;
    ep.i0 = b.i0.p; ep.j0 = b.j0.p; ep.end_i = b.si.p; ep.end_j = b.sj.p;
    ep.out_len_plan = b.lenp.p; ep.score_plan = b.score.p;
    ep.out_ref = b.o1.p; ep.out_qry = b.o2.p; ep.out_len = b.olen.p; ep.out_score = b.oscore.p;
    GOTOH_LAUNCH(k_emit, dim3((count + 3) / 4), dim3(128), 0, (cudaStream_t)0, ep);
    CU(cudaGetLastError());
    CU(cudaEventRecord(run.ev.e[3], 0));
    CU(cudaEventSynchronize(run.ev.e[3]));
    float a = 0, bb = 0, c = 0;
    CU(cudaEventElapsedTime(&a, run.ev.e[0], run.ev.e[1]));
    CU(cudaEventElapsedTime(&bb, run.ev.e[1], run.ev.e[2]));
    CU(cudaEventElapsedTime(&c, run.ev.e[2], run.ev.e[3]));
    run.ms_f += a; run.ms_r += bb; run.ms_w += c;
    run.launches += 2;
    return 0;
}

// ---- general kernels (gotoh2_kernels.cuh): any penalties, one warp per pair -------------------------------
int g2_run_general(G2Bufs& b, G2Run& run, std::vector<PairInfo>& pairs) {
    using namespace gotoh::g2;
    const size_t n = (size_t)run.n;
    const int grid = run.sm_count * 4, nwarps = grid * 4;
    const int64_t bstride = (run.max_rows + 2 + 15) & ~15LL;
    CU(b.fbnd.ensure((size_t)(nwarps * bstride)));
    CU(b.rbnd.ensure((size_t)(nwarps * bstride)));
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    int64_t budget = (int64_t)(free_b * 0.7) / 8;            // in uint2
    if (getenv("GOTOH_B200_ARENA_MB")) budget = ((int64_t)atoll(getenv("GOTOH_B200_ARENA_MB")) << 20) / 8;
    std::vector<std::pair<int, int>> chunks;               // (first, count)
    int64_t used_u2 = 0, arena_max = 0;
    int first = 0;
    for (size_t k = 0; k < n; ++k) {
        PairInfo& pi = pairs[k];
        pi.K = G2K;
        pi.nblk = pi.M + 1 + 31;                     // T: steps of the wavefront over the (l1+1)-row grid
        const int64_t nstrips = ((int64_t)pi.N + 1 + 32 * G2K - 1) / (32 * G2K);
        const int64_t need = nstrips * pi.nblk * 32;
        if (need > (int64_t)(free_b * 0.9) / 8) return fail(GOTOH_B200_ENOMEM, "pair %lld needs %lld bytes of tie-bit arena", (long long)k, (long long)need * 8);
        if (used_u2 > 0 && used_u2 + need > budget) { chunks.push_back({first, (int)(k - first)}); first = (int)k; used_u2 = 0; }
        pi.dir_off = used_u2;
        used_u2 += need;
        arena_max = std::max(arena_max, used_u2);
    }
    chunks.push_back({first, (int)(n - first)});
    CU(b.arena.ensure((size_t)arena_max));
    CU(cudaMemcpy(b.pairs.p, pairs.data(), n * sizeof(PairInfo), cudaMemcpyHostToDevice));
    run.arena_bytes = arena_max * 8;
    run.chunks = (int)chunks.size();
    for (const auto& ch : chunks) {
        Params p;
        memset(&p, 0, sizeof(p));
        p.pairs = b.pairs.p; p.pair_first = ch.first; p.pair_count = ch.second;
        p.s1_idx = b.idx1.p; p.s2_idx = b.idx2.p; p.dmat = b.dmat.p;
        p.l = run.l; p.v = run.gop; p.u = run.gep; p.is_global = run.is_global;
        p.arena = b.arena.p; p.fbnd = b.fbnd.p; p.rbnd = b.rbnd.p; p.bnd_stride = bstride;
        p.best = b.best.p; p.start_i = b.si.p; p.start_j = b.sj.p;
        const int g = std::max(1, std::min(grid, (ch.second + 3) / 4));
        CU(cudaEventRecord(run.ev.e[0], 0));
        GOTOH_LAUNCH((k2_forward<0>), dim3(g), dim3(128), 0, (cudaStream_t)0, p);
        CU(cudaGetLastError());
        CU(cudaEventRecord(run.ev.e[1], 0));
        GOTOH_LAUNCH((k2_reverse<0>), dim3(g), dim3(128), 0, (cudaStream_t)0, p);
        CU(cudaGetLastError());
        CU(cudaEventRecord(run.ev.e[2], 0));
        WalkParams2 wp;
        memset(&wp, 0, sizeof(wp));
        wp.pairs = b.pairs.p; wp.pair_first = ch.first; wp.pair_count = ch.second;
        wp.arena = reinterpret_cast<const uint8_t*>(b.arena.p);
        wp.best = b.best.p; wp.start_i = b.si.p; wp.start_j = b.sj.p;
        wp.ops = b.ops.p; wp.nops = b.nops.p; wp.i0 = b.i0.p; wp.j0 = b.j0.p; wp.out_len = b.lenp.p; wp.score = b.score.p;
        GOTOH_LAUNCH(k2_walk, dim3((ch.second + 127) / 128), dim3(128), 0, (cudaStream_t)0, wp);
        CU(cudaGetLastError());
        const int rc = g2_emit(b, run, ch.first, ch.second);
        if (rc) return rc;
        run.launches += 2;
    }
    return 0;
}

// ---- tuned kernels (gotoh2_fast.cuh): strip tasks, bit-plane tie codes --------------------------------------
template <int K, bool MULTI, bool BITS>
int g2_launch_fwd(G2Run& run, const gotoh::g2f::Params& p, int ntasks) {
    using namespace gotoh::g2f;
    const size_t per_warp = Smem<K>::per_warp(run.l);
    const size_t smem = per_warp * 4;
    if (smem > 220 * 1024) return fail(GOTOH_B200_ERANGE, "profile needs %zu bytes of shared memory", smem);
    CU(cudaFuncSetAttribute(k2f<K, MULTI, BITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (220 * 1024) / smem));
    const int grid = std::max(1, std::min((ntasks + 3) / 4, run.sm_count * ctas_per_sm));
    GOTOH_LAUNCH((k2f<K, MULTI, BITS>), dim3(grid), dim3(128), smem, (cudaStream_t)0, p);
    CU(cudaGetLastError());
    return 0;
}
template <int K>
int g2_launch_fwd_x2(G2Run& run, const gotoh::g2f::Params& p, int ntasks) {
    using namespace gotoh::g2f;
    const size_t per_warp = Smem<K>::per_warp(run.l);
    const size_t smem = per_warp * 4;
    if (smem > 220 * 1024) return fail(GOTOH_B200_ERANGE, "profile needs %zu bytes of shared memory", smem);
    CU(cudaFuncSetAttribute(k2f_x2<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (220 * 1024) / smem));
    const int grid = std::max(1, std::min((ntasks + 3) / 4, run.sm_count * ctas_per_sm));
    GOTOH_LAUNCH((k2f_x2<K>), dim3(grid), dim3(128), smem, (cudaStream_t)0, p);
    CU(cudaGetLastError());
    return 0;
}
template <int K, bool MULTI>
int g2_launch_rev(G2Run& run, const gotoh::g2f::Params& p, int ntasks) {
    using namespace gotoh::g2f;
    const int grid = std::max(1, std::min((ntasks + 3) / 4, run.sm_count * 8));
    GOTOH_LAUNCH((k2r<K, MULTI>), dim3(grid), dim3(128), 0, (cudaStream_t)0, p);
    CU(cudaGetLastError());
    return 0;
}
template <int K>
int g2_launch_rev_x2(G2Run& run, const gotoh::g2f::Params& p, int ntasks) {
    using namespace gotoh::g2f;
    const int grid = std::max(1, std::min((ntasks + 3) / 4, run.sm_count * 6));
    GOTOH_LAUNCH((k2r_x2<K>), dim3(grid), dim3(128), 0, (cudaStream_t)0, p);
    CU(cudaGetLastError());
    return 0;
}
template <int K>
int g2_launch_group(G2Run& run, const gotoh::g2f::Params& p, int ntasks, bool multi, int phase) {
    if (phase == 2) return g2_launch_fwd_x2<K>(run, p, ntasks);
    if (phase == 0) {
        if (run.score_only) return multi ? g2_launch_fwd<K, true, false>(run, p, ntasks) : g2_launch_fwd<K, false, false>(run, p, ntasks);
        return multi ? g2_launch_fwd<K, true, true>(run, p, ntasks) : g2_launch_fwd<K, false, true>(run, p, ntasks);
    }
    return multi ? g2_launch_rev<K, true>(run, p, ntasks) : g2_launch_rev<K, false>(run, p, ntasks);
}

// k2f_x2 admits a pair when every value of its grid in the frame X' = X - (i+j)*u - S, every difference the tie codes
// take and the "+infinity" stay inside 16 bits without wrapping (the adds of VIADDMNMX.S16x2 wrap).  With
// dmax = max(0, max d), gap costs >= 0 and at most min(l1,l2) diagonal steps: X~ >= -(dmax*min(l1,l2) + (l1+l2+40)*u)
// =: -L (40 rows of wavefront fill/drain are swept too), R~ <= 2v, p~, q~ <= 3v.  S = 4v + max(0,-dmin) + 8 makes every
// R' and every diagonal candidate negative and every negated value positive (see the kernel); "+infinity" is
// 3v + 16 - S.  Largest sum formed: a diagonal candidate minus the smallest R' < L + 4v - dmin; smallest value:
// -L - S - dmax - 2u.
inline int g2_shift16(int gop, int dmin) { return 4 * gop + std::max(0, -dmin) + 8; }
inline int g2_inf16(int gop, int dmin) { return 3 * gop + 16 - g2_shift16(gop, dmin); }
inline bool g2_fits_int16(int M, int N, int gop, int gep, int dmax, int dmin) {
    const long long L = (long long)std::max(0, dmax) * std::min(M, N) + (long long)(M + N + 40) * gep;
    return L + 5LL * gop + 2LL * std::max(0, -dmin) + std::max(0, dmax) + 2LL * gep + 64 <= 32000;
}

int g2_run_fast(G2Bufs& b, G2Run& run, std::vector<PairInfo>& pairs) {
    using namespace gotoh::g2f;
    const size_t n = (size_t)run.n;
    static const int kK[] = {2, 3, 4, 6, 8};
    // forward in int16x2 (two pairs that share seq1 per warp) wherever the range proof holds; GOTOH_B200_GOTOH2=x1
    // pins the int32 forward kernel for tests (still a GPU path)
    const char* g2sel = getenv("GOTOH_B200_GOTOH2");
    const bool use_x2 = !run.score_only && !(g2sel && !strcmp(g2sel, "x1"));
    // ---- geometry per pair: K columns per lane, strips, blocks; chunks by arena budget ----------------------
    // the arena budget comes from cudaMemGetInfo - which now and then takes tens of ms - only when the planes cached from
    // earlier calls cannot hold this batch in one chunk
    int64_t total_need = 0;
    for (size_t k = 0; k < n; ++k) {
        int K = 8;
        for (int x : kK) if (32 * x >= pairs[k].N) { K = x; break; }
        total_need += (((int64_t)pairs[k].N + 32 * K - 1) / (32 * K)) * ((pairs[k].M + 31 + FSTEPS - 1) / FSTEPS) * 32;
    }
    size_t free_b = 0, total_b = 0;
    int64_t budget;
    const int64_t cached = (int64_t)std::min(b.lo.cap, b.hi.cap);
    if (run.score_only) { budget = (int64_t)1 << 60; free_b = (size_t)1 << 60; }
    else if (!getenv("GOTOH_B200_ARENA_MB") && total_need <= cached) { budget = cached; free_b = (size_t)1 << 60; }
    else {
        CU(cudaMemGetInfo(&free_b, &total_b));
        free_b += (size_t)cached * 32;                       // what the cached planes hold is available to this call too
        budget = (int64_t)(free_b * 0.7) / 32;               // in uint4 per plane pair (two planes)
        if (getenv("GOTOH_B200_ARENA_MB")) budget = ((int64_t)atoll(getenv("GOTOH_B200_ARENA_MB")) << 20) / 32;
    }
    std::vector<Extra> extra(n);
    struct Chunk { int first, count; int64_t slots, bnd; };
    std::vector<Chunk> chunks;
    int64_t used = 0, arena_max = 0, slots = 0, bnd = 0, slots_max = 0, bnd_max = 0;
    int first = 0;
    for (size_t k = 0; k < n; ++k) {
        PairInfo& pi = pairs[k];
        int K = 8;
        for (int x : kK) if (32 * x >= pi.N) { K = x; break; }
        pi.K = (int16_t)K;
        pi.nblk = (pi.M + 31 + FSTEPS - 1) / FSTEPS;
        const int64_t nstrips = ((int64_t)pi.N + 32 * K - 1) / (32 * K);
        const int64_t need = nstrips * pi.nblk * 32;
        if (!run.score_only && need > (int64_t)(free_b * 0.9) / 32)
            return fail(GOTOH_B200_ENOMEM, "pair %lld needs %lld bytes of tie-bit arena", (long long)k, (long long)need * 32);
        if (used > 0 && (used + need > budget || slots + nstrips > 0x3fffffff)) {
            chunks.push_back({first, (int)(k - first), slots, bnd});
            first = (int)k; used = 0; slots = 0; bnd = 0;
        }
        pi.dir_off = used;
        used += need;
        extra[k].slot0 = (int32_t)slots;
        extra[k].nstrips = (int32_t)nstrips;
        extra[k].bnd_off = bnd;
        slots += nstrips;
        bnd += (nstrips - 1) * (int64_t)(pi.M + 2);
        arena_max = std::max(arena_max, used);
        slots_max = std::max(slots_max, slots);
        bnd_max = std::max(bnd_max, bnd);
    }
    chunks.push_back({first, (int)(n - first), slots, bnd});
    if (!run.score_only) { CU(b.lo.ensure((size_t)arena_max)); CU(b.hi.ensure((size_t)arena_max)); }
    CU(b.extra.ensure(n));
    CU(b.tasks.ensure((size_t)slots_max * 2 + 8));
    CU(b.prog.ensure((size_t)slots_max * 2 + 2));
    CU(b.part.ensure((size_t)slots_max * 2));
    CU(b.fbnd.ensure((size_t)bnd_max + 1));
    CU(b.rbnd.ensure((size_t)bnd_max + 1));
    CU(b.counters.ensure(40));     // task schedulers: forward [g], reverse [8+g], int16x2 forward [16+g], reverse remainder [24+g]
    CU(cudaMemcpy(b.pairs.p, pairs.data(), n * sizeof(PairInfo), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(b.extra.p, extra.data(), n * sizeof(Extra), cudaMemcpyHostToDevice));
    run.arena_bytes = run.score_only ? 0 : arena_max * 32;
    run.chunks = (int)chunks.size();

    int dmax = 0, dmin = 0;
    for (int x = 0; x < run.l * run.l; ++x) { dmax = std::max(dmax, (int)run.h_dmat[x]); dmin = std::min(dmin, (int)run.h_dmat[x]); }
    std::vector<StripTask> tasks;
    for (const auto& ch : chunks) {
        // groups: (K, single strip) and (8, multi strip); tasks of one group are contiguous in b.tasks.  The reverse sweep
        // takes every (pair, strip) of the group; the forward sweep takes the same list (int32) unless some pairs of a
        // single-strip group run in int16x2, in which case it gets its own int32 list and a list of pair couples.
        struct Group { int K; bool multi; int first, count, f32_first, f32_count, x2_first, x2_count; };
        std::vector<Group> groups;
        tasks.clear();
        std::vector<StripTask> f32, fx2;
        std::unordered_map<int64_t, int> open_couple;          // ref_pos -> index into fx2 of a task that still lacks its second pair
        for (int gi = 0; gi < 6; ++gi) {
            const int K = gi < 5 ? kK[gi] : 8;
            const bool multi = (gi == 5);
            const int t0 = (int)tasks.size();
            // strip-major order (strip 0 of every pair, then strip 1, ...): a warp that claims (pair, s) finds (pair, s-1)
            // far ahead instead of spinning right behind a producer that has just started; a producer still precedes its
            // consumer in claim order (forward) and in reversed order (reverse sweep), which rules out deadlock
            int max_ns = 0;
            for (int k = ch.first; k < ch.first + ch.count; ++k)
                if (pairs[(size_t)k].K == K && (extra[(size_t)k].nstrips > 1) == multi) max_ns = std::max(max_ns, extra[(size_t)k].nstrips);
            for (int s = 0; s < max_ns; ++s)
                for (int k = ch.first; k < ch.first + ch.count; ++k) {
                    if (pairs[(size_t)k].K != K || (extra[(size_t)k].nstrips > 1) != multi || s >= extra[(size_t)k].nstrips) continue;
                    tasks.push_back({k, s});
                }
            if ((int)tasks.size() == t0) continue;
            Group g = {K, multi, t0, (int)tasks.size() - t0, t0, (int)tasks.size() - t0, 0, 0};
            if (use_x2 && !multi) {
                f32.clear(); fx2.clear(); open_couple.clear();
                for (int x = t0; x < (int)tasks.size(); ++x) {
                    const PairInfo& pi = pairs[(size_t)tasks[(size_t)x].pair];
                    if (!g2_fits_int16(pi.M, pi.N, run.gop, run.gep, dmax, dmin)) { f32.push_back(tasks[(size_t)x]); continue; }
                    auto it = open_couple.find(pi.ref_pos);
                    if (it != open_couple.end()) { fx2[(size_t)it->second].strip = tasks[(size_t)x].pair; open_couple.erase(it); }
                    else { open_couple[pi.ref_pos] = (int)fx2.size(); fx2.push_back({tasks[(size_t)x].pair, -1}); }
                }
                if (!fx2.empty()) {
                    g.f32_first = (int)tasks.size(); g.f32_count = (int)f32.size();
                    tasks.insert(tasks.end(), f32.begin(), f32.end());
                    g.x2_first = (int)tasks.size(); g.x2_count = (int)fx2.size();
                    tasks.insert(tasks.end(), fx2.begin(), fx2.end());
                }
            }
            groups.push_back(g);
        }
        CU(cudaMemcpy(b.tasks.p, tasks.data(), tasks.size() * sizeof(StripTask), cudaMemcpyHostToDevice));
        // strip boundaries and last-row partials are self-validating: preset them to "empty" (BND_EMPTY, 0xff, PART_EMPTY2)
        if (ch.bnd > 0) {
            CU(cudaMemset(b.fbnd.p, 0x80, (size_t)ch.bnd * sizeof(int2)));
            CU(cudaMemset(b.rbnd.p, 0xff, (size_t)ch.bnd));
        }
        CU(cudaMemset(b.part.p, 0x80, (size_t)ch.slots * sizeof(int32_t)));
        CU(cudaMemset(b.counters.p, 0, 40 * sizeof(uint32_t)));
        Params p;
        memset(&p, 0, sizeof(p));
        p.pairs = b.pairs.p; p.extra = b.extra.p;
        p.s1_idx = b.idx1.p; p.s2_idx = b.idx2.p; p.dmat = b.dmat.p;
        p.l = run.l; p.v = run.gop; p.u = run.gep; p.is_global = run.is_global;
        p.two = 2; p.four = 4; p.neg1 = 0xffffffffu;
        p.inf16 = g2_inf16(run.gop, dmin); p.shift16 = g2_shift16(run.gop, dmin);
        p.lo = b.lo.p; p.hi = b.hi.p; p.bnd = b.fbnd.p; p.rbnd = b.rbnd.p;
        p.prog_f = b.prog.p; p.prog_r = b.prog.p + slots_max;
        p.part_min = b.part.p; p.part_j = b.part.p + slots_max;
        p.best = b.best.p; p.start_i = b.si.p; p.start_j = b.sj.p;
        CU(cudaEventRecord(run.ev.e[0], 0));
        for (size_t g = 0; g < groups.size(); ++g) {
            for (int pass = 0; pass < 2; ++pass) {             // pass 0: int16x2 couples, pass 1: int32 tasks
                const int first = pass == 0 ? groups[g].x2_first : groups[g].f32_first;
                const int count = pass == 0 ? groups[g].x2_count : groups[g].f32_count;
                if (count == 0) continue;
                p.tasks = b.tasks.p + first; p.task_count = count;
                p.counter_f = b.counters.p + (pass == 0 ? 16 : 0) + g;
                const int phase = pass == 0 ? 2 : 0;
                int rc = 0;
                switch (groups[g].K) {
                    case 2: rc = g2_launch_group<2>(run, p, count, groups[g].multi, phase); break;
                    case 3: rc = g2_launch_group<3>(run, p, count, groups[g].multi, phase); break;
                    case 4: rc = g2_launch_group<4>(run, p, count, groups[g].multi, phase); break;
                    case 6: rc = g2_launch_group<6>(run, p, count, groups[g].multi, phase); break;
                    default: rc = g2_launch_group<8>(run, p, count, groups[g].multi, phase); break;
                }
                if (rc) return rc;
                ++run.launches;
                if (pass == 0) run.pairs_x2 += count;
            }
        }
        CU(cudaEventRecord(run.ev.e[1], 0));
        if (!run.score_only) {
            for (size_t g = 0; g < groups.size(); ++g) {
                // K <= 3: the couples of the int16x2 forward kernel are swept two per warp (k2r_x2), the other pairs of the
                // group one per warp; GOTOH_B200_GOTOH2=r1 pins the one-pair kernel (tests)
                const bool r2 = groups[g].x2_count > 0 && groups[g].K <= 3 && !(g2sel && !strcmp(g2sel, "r1"));
                if (r2) {
                    p.tasks = b.tasks.p + groups[g].x2_first; p.task_count = groups[g].x2_count;
                    p.counter_r = b.counters.p + 8 + g;
                    const int rc = groups[g].K == 2 ? g2_launch_rev_x2<2>(run, p, p.task_count) : g2_launch_rev_x2<3>(run, p, p.task_count);
                    if (rc) return rc;
                    ++run.launches;
                    if (groups[g].f32_count == 0) continue;
                    p.tasks = b.tasks.p + groups[g].f32_first; p.task_count = groups[g].f32_count;
                    p.counter_r = b.counters.p + 24 + g;
                } else {
                    p.tasks = b.tasks.p + groups[g].first; p.task_count = groups[g].count;
                    p.counter_r = b.counters.p + 8 + g;
                }
                int rc = 0;
                switch (groups[g].K) {
                    case 2: rc = g2_launch_group<2>(run, p, p.task_count, groups[g].multi, 1); break;
                    case 3: rc = g2_launch_group<3>(run, p, p.task_count, groups[g].multi, 1); break;
                    case 4: rc = g2_launch_group<4>(run, p, p.task_count, groups[g].multi, 1); break;
                    case 6: rc = g2_launch_group<6>(run, p, p.task_count, groups[g].multi, 1); break;
                    default: rc = g2_launch_group<8>(run, p, p.task_count, groups[g].multi, 1); break;
                }
                if (rc) return rc;
                ++run.launches;
            }
        }
        CU(cudaEventRecord(run.ev.e[2], 0));
        if (run.score_only) {
            CU(cudaEventSynchronize(run.ev.e[2]));
            float a = 0;
            CU(cudaEventElapsedTime(&a, run.ev.e[0], run.ev.e[1]));
            run.ms_f += a;
            continue;
        }
        gotoh::g2f::WalkParams wp;
        memset(&wp, 0, sizeof(wp));
        wp.pairs = b.pairs.p; wp.pair_first = ch.first; wp.pair_count = ch.count;
        wp.hi = reinterpret_cast<const uint32_t*>(b.hi.p);
        wp.best = b.best.p; wp.start_i = b.si.p; wp.start_j = b.sj.p;
        wp.ops = b.ops.p; wp.nops = b.nops.p; wp.i0 = b.i0.p; wp.j0 = b.j0.p; wp.out_len = b.lenp.p; wp.score = b.score.p;
        // long paths or few pairs: one warp per pair with a prefetched tile; many short pairs: one thread per pair
        long long path = 0;
        for (int k = ch.first; k < ch.first + ch.count; ++k) path += pairs[(size_t)k].M + pairs[(size_t)k].N;
        const char* wsel = getenv("GOTOH_B200_WALK");
        const bool warp_walk = wsel ? !strcmp(wsel, "warp") : (path / ch.count >= 1024 || ch.count < 8192);
        if (warp_walk) GOTOH_LAUNCH(k2f_walk_warp, dim3((ch.count + 3) / 4), dim3(128), 0, (cudaStream_t)0, wp);
        else GOTOH_LAUNCH(k2f_walk, dim3((ch.count + 127) / 128), dim3(128), 0, (cudaStream_t)0, wp);
        CU(cudaGetLastError());
        const int rc = g2_emit(b, run, ch.first, ch.count);
        if (rc) return rc;
    }
    return 0;
}

// Byte -> (cleaned byte, class index) tables of one side of the grid.  gotoh2.py:70-72: ASCII upper-case, then every
// byte that is not in the alphabet becomes '?'; cls < 0 marks a byte the alphabet cannot express at all.
struct G2Map {
    uint8_t clean[256];
    int16_t cls[256];
};

int g2_align_batch(int device, const uint8_t* s1_bytes, const int64_t* s1_off, int64_t n_s1, const int32_t* s1_idx,
                   const uint8_t* s2_bytes, const int64_t* s2_off, int64_t n_pairs, int gop, int gep, int is_global,
                   int l, const G2Map& map1, const G2Map& map2, const int32_t* matrix, uint8_t* out1, uint8_t* out2,
                   const int64_t* out_off, int32_t* out_len, int32_t* out_score, bool score_only) {
    using namespace gotoh::g2;

    // ---- clean + index every used first sequence once, every second sequence ---------------------
    std::vector<int32_t> local;
    std::vector<int64_t> used;
    if (s1_idx) {
        local.assign((size_t)n_s1, -1);
        for (int64_t k = 0; k < n_pairs; ++k) {
            const int64_t r = s1_idx[k];
            if (r < 0 || r >= n_s1) return fail(GOTOH_B200_EINVAL, "pair %lld: seq1 index out of range", (long long)k);
            if (local[(size_t)r] < 0) { local[(size_t)r] = (int32_t)used.size(); used.push_back(r); }
        }
    } else {
        used.resize((size_t)n_pairs);
        for (int64_t k = 0; k < n_pairs; ++k) used[(size_t)k] = k;
    }
    std::vector<int64_t> pos1(used.size());
    std::vector<uint8_t> h_raw1, h_idx1, h_raw2, h_idx2;
    auto add_seq = [&](const uint8_t* s, int64_t len, const G2Map& mp, std::vector<uint8_t>& raw, std::vector<uint8_t>& idx,
                       const char* what, long long id) -> int {
        if (len <= 0) return fail(GOTOH_B200_EEMPTY, "%s %lld is empty (gotoh2.py:84-85 asserts non-empty)", what, id);
        if (len >= (1 << 24)) return fail(GOTOH_B200_ERANGE, "%s %lld too long", what, id);
        const size_t at = raw.size();
        raw.resize(at + (size_t)len);
        idx.resize(at + (size_t)len);
        for (int64_t x = 0; x < len; ++x) {
            const int c = mp.cls[s[x]];
            if (c < 0) return fail(GOTOH_B200_EDOMAIN, "%s %lld: byte 0x%02x is not in the alphabet and the alphabet has no '?'", what, id, s[x]);
            raw[at + (size_t)x] = mp.clean[s[x]];
            idx[at + (size_t)x] = (uint8_t)c;
        }
        return 0;
    };
    // every first sequence carries REF_PAD zero bytes on both sides: the wavefront reads the row class a few
    // rows before row 1 and after row l1 without a bounds check
    h_raw1.assign(gotoh::REF_PAD, 0); h_idx1.assign(gotoh::REF_PAD, 0);
    for (size_t x = 0; x < used.size(); ++x) {
        const int64_t r = used[x];
        pos1[x] = (int64_t)h_raw1.size();
        const int rc = add_seq(s1_bytes + s1_off[r], s1_off[r + 1] - s1_off[r], map1, h_raw1, h_idx1, "seq1", (long long)r);
        if (rc) return rc;
        h_raw1.insert(h_raw1.end(), gotoh::REF_PAD, 0); h_idx1.insert(h_idx1.end(), gotoh::REF_PAD, 0);
    }
    std::vector<PairInfo> pairs((size_t)n_pairs);
    int64_t ops_words = 0, max_rows = 0, q_total = 0;
    long long cells = 0;
    for (int64_t k = 0; k < n_pairs; ++k) {
        PairInfo& pi = pairs[(size_t)k];
        memset(&pi, 0, sizeof(pi));
        const int64_t r = s1_idx ? local[(size_t)s1_idx[k]] : k;
        const int64_t u1 = used[(size_t)r];
        const int64_t len2 = s2_off[k + 1] - s2_off[k];
        if (len2 <= 0) return fail(GOTOH_B200_EEMPTY, "seq2 %lld is empty (gotoh2.py:84-85 asserts non-empty)", (long long)k);
        if (len2 >= (1 << 24)) return fail(GOTOH_B200_ERANGE, "seq2 %lld too long", (long long)k);
        pi.ref_pos = pos1[(size_t)r];
        pi.qry_pos = q_total;
        q_total += len2;
        pi.M = (int32_t)(s1_off[u1 + 1] - s1_off[u1]);
        pi.N = (int32_t)len2;
        pi.nblk = pi.M + 1 + 31;                     // T: steps of the wavefront over the (l1+1)-row grid
        pi.K = G2K;
        pi.orig = (int32_t)k;
        if (!score_only) {
            pi.out_off = out_off[k] - out_off[0];
            const int64_t cap = out_off[k + 1] - out_off[k];
            if (cap < (int64_t)pi.M + pi.N || cap > 0x7fffffffLL) return fail(GOTOH_B200_ERANGE, "pair %lld: output stride < l1+l2", (long long)k);
            pi.out_cap = (int32_t)cap;
        }
        if (ops_words + (pi.M + pi.N + 15) / 16 > 0x7fffffffLL) return fail(GOTOH_B200_ERANGE, "op-script arena too large; split the batch");
        pi.ops_off = (int32_t)ops_words;
        ops_words += (pi.M + pi.N + 15) / 16;
        max_rows = std::max<int64_t>(max_rows, pi.M + 1);
        cells += (long long)(pi.M + 1) * (pi.N + 1);
        // costs stay far below the int32 "infinity" of the kernels
        if ((long long)(pi.M + pi.N + 2) * (std::abs(gop) + std::abs(gep) + 64) > (1LL << 27))
            return fail(GOTOH_B200_ERANGE, "pair %lld: penalties x lengths exceed the int32 cost range", (long long)k);
    }
    // second sequences: clean + index in parallel into their final positions
    h_raw2.resize((size_t)q_total);
    h_idx2.resize((size_t)q_total);
    {
        const int nthreads = host_threads(q_total);
        std::vector<int64_t> bad_pair((size_t)nthreads, -1);
        std::vector<int> bad_byte((size_t)nthreads, 0);
        parallel_for(n_pairs, nthreads, [&](int64_t lo_k, int64_t hi_k, int tid) {
            for (int64_t k = lo_k; k < hi_k; ++k) {
                const uint8_t* src = s2_bytes + s2_off[k];
                const int64_t len = s2_off[k + 1] - s2_off[k], at = pairs[(size_t)k].qry_pos;
                for (int64_t x = 0; x < len; ++x) {
                    const int c = map2.cls[src[x]];
                    if (c < 0) { if (bad_pair[(size_t)tid] < 0) { bad_pair[(size_t)tid] = k; bad_byte[(size_t)tid] = src[x]; } return; }
                    h_raw2[(size_t)(at + x)] = map2.clean[src[x]];
                    h_idx2[(size_t)(at + x)] = (uint8_t)c;
                }
            }
        });
        for (int t = 0; t < nthreads; ++t)
            if (bad_pair[(size_t)t] >= 0)
                return fail(GOTOH_B200_EDOMAIN, "seq2 %lld: byte 0x%02x is not in the alphabet and the alphabet has no '?'",
                            (long long)bad_pair[(size_t)t], bad_byte[(size_t)t]);
    }
    for (int x = 0; x < l * l; ++x)
        if (std::abs(matrix[x]) > 1000000) return fail(GOTOH_B200_ERANGE, "substitution score out of range");

    // ---- device ---------------------------------------------------------------------------------
    CU(cudaSetDevice(device));
    // grow-only device buffers, cached per device between calls (gotoh_b200_release_cache frees them): a call on a
    // few hundred pairs must not pay for cudaMalloc/cudaFree of the tie-bit arena
    G2Cache* cache = g2_cache_for(device);
    if (!cache) return fail(GOTOH_B200_ENOMEM, "out of host memory");
    std::lock_guard<std::mutex> cache_lock(cache->mu);
    if (cache->sm_count <= 0) {
        cudaDeviceProp prop;
        CU(cudaGetDeviceProperties(&prop, device));
        cache->sm_count = prop.multiProcessorCount;
    }
    G2Bufs& b = cache->bufs;
    const size_t n = (size_t)n_pairs;
    const int64_t out_bytes = score_only ? 0 : out_off[n_pairs] - out_off[0];
    CU(b.raw1.ensure(h_raw1.size())); CU(b.idx1.ensure(h_idx1.size()));
    CU(b.raw2.ensure(h_raw2.size())); CU(b.idx2.ensure(h_idx2.size()));
    CU(b.pairs.ensure(n)); CU(b.dmat.ensure((size_t)l * l));
    CU(b.best.ensure(n)); CU(b.si.ensure(n)); CU(b.sj.ensure(n)); CU(b.nops.ensure(n)); CU(b.i0.ensure(n)); CU(b.j0.ensure(n));
    CU(b.lenp.ensure(n)); CU(b.score.ensure(n)); CU(b.olen.ensure(n)); CU(b.oscore.ensure(n));
    CU(b.ops.ensure((size_t)ops_words));
    if (!score_only) { CU(b.o1.ensure((size_t)out_bytes)); CU(b.o2.ensure((size_t)out_bytes)); }
    CU(cudaMemcpy(b.raw1.p, h_raw1.data(), h_raw1.size(), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(b.idx1.p, h_idx1.data(), h_idx1.size(), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(b.raw2.p, h_raw2.data(), h_raw2.size(), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(b.idx2.p, h_idx2.data(), h_idx2.size(), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(b.dmat.p, matrix, (size_t)l * l * sizeof(int32_t), cudaMemcpyHostToDevice));

    G2Run run;
    run.n = n_pairs; run.l = l; run.gop = gop; run.gep = gep; run.is_global = is_global ? 1 : 0;
    run.score_only = score_only; run.sm_count = cache->sm_count; run.max_rows = max_rows; run.h_dmat = matrix;
    if (!cache->ev.create()) return fail(GOTOH_B200_ECUDA, "cudaEventCreate failed");
    run.ev = cache->ev;
    const char* force = getenv("GOTOH_B200_GOTOH2");          // tests: "general" pins the un-tuned kernels (still a GPU path)
    const bool fast = gop >= 0 && gep >= 0 && !(force && !strcmp(force, "general"));
    if (score_only && !fast) return fail(GOTOH_B200_ERANGE, "score-only mode needs non-negative penalties");
    const int rc = fast ? g2_run_fast(b, run, pairs) : g2_run_general(b, run, pairs);
    if (rc) return rc;

    g2_stats[0] = (double)cells; g2_stats[1] = run.ms_f + run.ms_r + run.ms_w; g2_stats[2] = run.ms_f; g2_stats[3] = run.ms_r;
    g2_stats[4] = run.ms_w; g2_stats[5] = run.launches; g2_stats[6] = (double)run.arena_bytes; g2_stats[7] = (double)run.chunks;
    g2_stats[8] = (double)(h_raw1.size() * 2 + h_raw2.size() * 2 + n * sizeof(PairInfo) + (size_t)l * l * 4);
    g2_stats[9] = (double)((score_only ? 0 : 2 * out_bytes) + 8 * (int64_t)n);
    g2_stats[10] = (double)run.pairs_x2;
    if (score_only) {
        CU(cudaMemcpy(out_score, b.best.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost));
        for (int64_t k = 0; k < n_pairs; ++k) out_score[k] = -out_score[k];
        return GOTOH_B200_OK;
    }
    CU(cudaMemcpy(out1 + out_off[0], b.o1.p, (size_t)out_bytes, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(out2 + out_off[0], b.o2.p, (size_t)out_bytes, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(out_len, b.olen.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(out_score, b.oscore.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost));
    for (int64_t k = 0; k < n_pairs; ++k)
        if (out_score[k] == (int32_t)0x80000000)
            return fail(GOTOH_B200_ETRACEBACK, "pair %lld: Traceback failed, try local alignment", (long long)k);
    return GOTOH_B200_OK;
}

}  // namespace

extern "C" int32_t gotoh_b200_gotoh2_last_stats(double* out, int32_t n) {
    if (!out || n < 0) return fail(GOTOH_B200_EINVAL, "NULL argument");
    for (int x = 0; x < n && x < 11; ++x) out[x] = g2_stats[x];
    return n < 11 ? n : 11;
}

extern "C" int32_t gotoh_b200_gotoh2_align_batch(const uint8_t* s1_bytes, const int64_t* s1_off, int64_t n_s1,
                                                 const int32_t* s1_idx, const uint8_t* s2_bytes, const int64_t* s2_off,
                                                 int64_t n_pairs, int32_t gop, int32_t gep, int32_t is_global,
                                                 const char* alphabet, const int32_t* matrix, uint8_t* out1,
                                                 uint8_t* out2, const int64_t* out_off, int32_t* out_len,
                                                 int32_t* out_score, int32_t device) {
    if (!s1_bytes || !s1_off || !s2_bytes || !s2_off || !alphabet || !matrix || !out1 || !out2 || !out_off || !out_len || !out_score)
        return fail(GOTOH_B200_EINVAL, "NULL pointer argument");
    if (n_pairs < 0 || n_s1 < 0 || n_pairs > 0x7fffffffLL) return fail(GOTOH_B200_EINVAL, "bad count");
    if (!s1_idx && n_s1 != n_pairs) return fail(GOTOH_B200_EINVAL, "s1_idx is NULL but n_s1 != n_pairs");
    const int ndev = gotoh_b200_device_count();
    if (ndev <= 0) return fail(GOTOH_B200_ENODEVICE, "no CUDA device is visible; libgotoh_b200 has no CPU path");
    if (device < 0 || device >= ndev) return fail(GOTOH_B200_ENODEVICE, "device %d not present (%d visible)", device, ndev);
    if (n_pairs == 0) return GOTOH_B200_OK;
    const int l = (int)strlen(alphabet);
    if (l < 1 || l > 32) return fail(GOTOH_B200_EINVAL, "alphabet length %d not in 1..32", l);
    G2Map mp;
    int cls_of[256];
    for (int c = 0; c < 256; ++c) cls_of[c] = -1;
    for (int x = 0; x < l; ++x) cls_of[(uint8_t)alphabet[x]] = x;                       // _gotoh2.c:68-77
    for (int c = 0; c < 256; ++c) {
        int up = (c >= 'a' && c <= 'z') ? c - 32 : c;                                    // gotoh2.py:72 seq.upper()
        if (cls_of[up] < 0) up = '?';                                                    // gotoh2.py:72 [^alphabet] -> '?'
        mp.clean[c] = (uint8_t)up;
        mp.cls[c] = (int16_t)(c == 0 ? -1 : cls_of[up]);                                 // NUL cannot occur in a Python str argument
    }
    try {
        return g2_align_batch(device, s1_bytes, s1_off, n_s1, s1_idx, s2_bytes, s2_off, n_pairs, gop, gep, is_global,
                              l, mp, mp, matrix, out1, out2, out_off, out_len, out_score, false);
    } catch (const std::bad_alloc&) {
        return fail(GOTOH_B200_ENOMEM, "out of host memory while packing");
    }
}

// Levenshtein.distance(a, b) of remap.py:250 (third-party python-Levenshtein, unpinned in INSTALL.md:8,22): unit-cost
// edit distance = the min-cost global alignment with open 0, extend 1, substitution cost [x != y], i.e. the score-only
// forward kernel with v = 0, u = 1 and a 0/-1 matrix over the bytes that occur on both sides.
extern "C" int32_t gotoh_b200_edit_distance_batch(const uint8_t* a_bytes, const int64_t* a_off, const uint8_t* b_bytes,
                                                  const int64_t* b_off, int64_t n_pairs, int32_t* out_dist, int32_t device) {
    if (!a_bytes || !a_off || !b_bytes || !b_off || !out_dist) return fail(GOTOH_B200_EINVAL, "NULL pointer argument");
    if (n_pairs < 0 || n_pairs > 0x7fffffffLL) return fail(GOTOH_B200_EINVAL, "bad count");
    const int ndev = gotoh_b200_device_count();
    if (ndev <= 0) return fail(GOTOH_B200_ENODEVICE, "no CUDA device is visible; libgotoh_b200 has no CPU path");
    if (device < 0 || device >= ndev) return fail(GOTOH_B200_ENODEVICE, "device %d not present (%d visible)", device, ndev);
    if (n_pairs == 0) return GOTOH_B200_OK;
    try {
        // pairs with an empty side need no DP: the distance is the other length
        std::vector<int64_t> keep;
        bool in_a[256] = {false}, in_b[256] = {false};
        for (int64_t k = 0; k < n_pairs; ++k) {
            const int64_t la = a_off[k + 1] - a_off[k], lb = b_off[k + 1] - b_off[k];
            if (la < 0 || lb < 0) return fail(GOTOH_B200_EINVAL, "pair %lld: negative length", (long long)k);
            if (la == 0 || lb == 0) { out_dist[k] = (int32_t)(la + lb); continue; }
            keep.push_back(k);
            for (int64_t x = a_off[k]; x < a_off[k + 1]; ++x) in_a[a_bytes[x]] = true;
            for (int64_t x = b_off[k]; x < b_off[k + 1]; ++x) in_b[b_bytes[x]] = true;
        }
        if (keep.empty()) return GOTOH_B200_OK;
        // classes: one per byte present on both sides, plus "only in a" (l-2) and "only in b" (l-1), which match nothing
        G2Map ma, mb;
        int ncommon = 0;
        for (int c = 0; c < 256; ++c) if (in_a[c] && in_b[c]) ++ncommon;
        if (ncommon > 30) return fail(GOTOH_B200_ERANGE, "%d distinct bytes occur on both sides; at most 30 are supported", ncommon);
        const int l = ncommon + 2;
        int next = 0;
        for (int c = 0; c < 256; ++c) {
            ma.clean[c] = mb.clean[c] = (uint8_t)c;
            if (in_a[c] && in_b[c]) { ma.cls[c] = mb.cls[c] = (int16_t)next++; }
            else { ma.cls[c] = (int16_t)(l - 2); mb.cls[c] = (int16_t)(l - 1); }
        }
        std::vector<int32_t> matrix((size_t)l * l, -1);
        for (int x = 0; x < ncommon; ++x) matrix[(size_t)x * l + x] = 0;
        // compact the kept pairs' offsets (bytes stay where they are: offsets need not be contiguous)
        const size_t m = keep.size();
        std::vector<int64_t> ao(m + 1), bo(m + 1);
        std::vector<uint8_t> ab, bb;
        for (size_t x = 0; x < m; ++x) {
            const int64_t k = keep[x];
            ao[x] = (int64_t)ab.size(); bo[x] = (int64_t)bb.size();
            ab.insert(ab.end(), a_bytes + a_off[k], a_bytes + a_off[k + 1]);
            bb.insert(bb.end(), b_bytes + b_off[k], b_bytes + b_off[k + 1]);
        }
        ao[m] = (int64_t)ab.size(); bo[m] = (int64_t)bb.size();
        std::vector<int32_t> sc(m);
        const int rc = g2_align_batch(device, ab.data(), ao.data(), (int64_t)m, nullptr, bb.data(), bo.data(), (int64_t)m, 0, 1, 1,
                                      l, ma, mb, matrix.data(), nullptr, nullptr, nullptr, nullptr, sc.data(), true);
        if (rc) return rc;
        for (size_t x = 0; x < m; ++x) out_dist[keep[x]] = -sc[x];
        return GOTOH_B200_OK;
    } catch (const std::bad_alloc&) {
        return fail(GOTOH_B200_ENOMEM, "out of host memory while packing");
    }
}

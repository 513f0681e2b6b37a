// gotoh2_host.cuh - host side of gotoh_b200_gotoh2_align_batch (the live aligner).
// Included by gotoh_b200.cu after its helpers (fail(), CU(), DevBuf, trim-free: gotoh2.py does not trim).
#pragma once

#include "gotoh2_kernels.cuh"

namespace {

// gotoh2.py:70-72 on bytes: ASCII upper-case, then every byte that is not in the alphabet becomes '?'.
inline uint8_t g2_clean(uint8_t c, const bool* in_alpha) {
    if (c >= 'a' && c <= 'z') c = (uint8_t)(c - 32);
    return in_alpha[c] ? c : (uint8_t)'?';
}

int g2_align_batch(int device, const uint8_t* s1_bytes, const int64_t* s1_off, int64_t n_s1, const int32_t* s1_idx,
                   const uint8_t* s2_bytes, const int64_t* s2_off, int64_t n_pairs, int gop, int gep, int is_global,
                   const char* alphabet, const int32_t* matrix, uint8_t* out1, uint8_t* out2, const int64_t* out_off,
                   int32_t* out_len, int32_t* out_score) {
    using namespace gotoh::g2;
    const int l = (int)strlen(alphabet);
    if (l < 1 || l > 32) return fail(GOTOH_B200_EINVAL, "alphabet length %d not in 1..32", l);
    bool in_alpha[256] = {false};
    int map[256];
    for (int c = 0; c < 256; ++c) map[c] = -1;
    for (int x = 0; x < l; ++x) { in_alpha[(uint8_t)alphabet[x]] = true; map[(uint8_t)alphabet[x]] = x; }   // _gotoh2.c:68-77

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
    auto add_seq = [&](const uint8_t* s, int64_t len, std::vector<uint8_t>& raw, std::vector<uint8_t>& idx, const char* what,
                       long long id) -> int {
        if (len <= 0) return fail(GOTOH_B200_EEMPTY, "%s %lld is empty (gotoh2.py:84-85 asserts non-empty)", what, id);
        if (len >= (1 << 24)) return fail(GOTOH_B200_ERANGE, "%s %lld too long", what, id);
        for (int64_t x = 0; x < len; ++x) {
            if (s[x] == 0) return fail(GOTOH_B200_EDOMAIN, "%s %lld contains a NUL byte", what, id);
            const uint8_t c = g2_clean(s[x], in_alpha);
            if (map[c] < 0) return fail(GOTOH_B200_EDOMAIN, "%s %lld: byte 0x%02x is not in the alphabet and the alphabet has no '?'", what, id, s[x]);
            raw.push_back(c);
            idx.push_back((uint8_t)map[c]);
        }
        return 0;
    };
    for (size_t x = 0; x < used.size(); ++x) {
        const int64_t r = used[x];
        pos1[x] = (int64_t)h_raw1.size();
        const int rc = add_seq(s1_bytes + s1_off[r], s1_off[r + 1] - s1_off[r], h_raw1, h_idx1, "seq1", (long long)r);
        if (rc) return rc;
    }
    std::vector<PairInfo> pairs((size_t)n_pairs);
    int64_t ops_words = 0, max_rows = 0;
    long long cells = 0;
    for (int64_t k = 0; k < n_pairs; ++k) {
        PairInfo& pi = pairs[(size_t)k];
        memset(&pi, 0, sizeof(pi));
        const int64_t r = s1_idx ? local[(size_t)s1_idx[k]] : k;
        const int64_t u1 = used[(size_t)r];
        pi.ref_pos = pos1[(size_t)r];
        pi.qry_pos = (int64_t)h_raw2.size();
        const int rc = add_seq(s2_bytes + s2_off[k], s2_off[k + 1] - s2_off[k], h_raw2, h_idx2, "seq2", (long long)k);
        if (rc) return rc;
        pi.M = (int32_t)(s1_off[u1 + 1] - s1_off[u1]);
        pi.N = (int32_t)(s2_off[k + 1] - s2_off[k]);
        pi.nblk = pi.M + 1 + 31;                     // T: steps of the wavefront over the (l1+1)-row grid
        pi.K = G2K;
        pi.orig = (int32_t)k;
        pi.out_off = out_off[k] - out_off[0];
        const int64_t cap = out_off[k + 1] - out_off[k];
        if (cap < (int64_t)pi.M + pi.N || cap > 0x7fffffffLL) return fail(GOTOH_B200_ERANGE, "pair %lld: output stride < l1+l2", (long long)k);
        pi.out_cap = (int32_t)cap;
        if (ops_words + (pi.M + pi.N + 15) / 16 > 0x7fffffffLL) return fail(GOTOH_B200_ERANGE, "op-script arena too large; split the batch");
        pi.ops_off = (int32_t)ops_words;
        ops_words += (pi.M + pi.N + 15) / 16;
        max_rows = std::max<int64_t>(max_rows, pi.M + 1);
        cells += (long long)(pi.M + 1) * (pi.N + 1);
        // costs stay far below the int32 "infinity" of the kernels
        if ((long long)(pi.M + pi.N + 2) * (std::abs(gop) + std::abs(gep) + 64) > (1LL << 27))
            return fail(GOTOH_B200_ERANGE, "pair %lld: penalties x lengths exceed the int32 cost range", (long long)k);
    }
    for (int x = 0; x < l * l; ++x)
        if (std::abs(matrix[x]) > 1000000) return fail(GOTOH_B200_ERANGE, "substitution score out of range");

    // ---- device ---------------------------------------------------------------------------------
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    struct Bufs {
        DevBuf<uint8_t> raw1, idx1, raw2, idx2, o1, o2, rbnd;
        DevBuf<PairInfo> pairs;
        DevBuf<int32_t> dmat, best, si, sj, nops, i0, j0, lenp, score, olen, oscore;
        DevBuf<uint32_t> ops;
        DevBuf<uint2> arena;
        DevBuf<int2> fbnd;
        ~Bufs() {
            raw1.release(); idx1.release(); raw2.release(); idx2.release(); o1.release(); o2.release(); rbnd.release();
            pairs.release(); dmat.release(); best.release(); si.release(); sj.release(); nops.release(); i0.release();
            j0.release(); lenp.release(); score.release(); olen.release(); oscore.release(); ops.release(); arena.release();
            fbnd.release();
        }
    } b;
    const size_t n = (size_t)n_pairs;
    const int64_t out_bytes = out_off[n_pairs] - out_off[0];
    CU(b.raw1.ensure(h_raw1.size())); CU(b.idx1.ensure(h_idx1.size()));
    CU(b.raw2.ensure(h_raw2.size())); CU(b.idx2.ensure(h_idx2.size()));
    CU(b.pairs.ensure(n)); CU(b.dmat.ensure((size_t)l * l));
    CU(b.best.ensure(n)); CU(b.si.ensure(n)); CU(b.sj.ensure(n)); CU(b.nops.ensure(n)); CU(b.i0.ensure(n)); CU(b.j0.ensure(n));
    CU(b.lenp.ensure(n)); CU(b.score.ensure(n)); CU(b.olen.ensure(n)); CU(b.oscore.ensure(n));
    CU(b.ops.ensure((size_t)ops_words));
    CU(b.o1.ensure((size_t)out_bytes)); CU(b.o2.ensure((size_t)out_bytes));
    const int grid = prop.multiProcessorCount * 4, nwarps = grid * 4;
    const int64_t bstride = (max_rows + 2 + 15) & ~15LL;
    CU(b.fbnd.ensure((size_t)(nwarps * bstride)));
    CU(b.rbnd.ensure((size_t)(nwarps * bstride)));

    // arena chunks: one byte per grid cell, [strip][step][lane] x 8 bytes
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    int64_t budget = (int64_t)(free_b * 0.7) / 8;            // in uint2
    if (getenv("GOTOH_B200_ARENA_MB")) budget = ((int64_t)atoll(getenv("GOTOH_B200_ARENA_MB")) << 20) / 8;
    std::vector<std::pair<int, int>> chunks;               // (first, count)
    int64_t used_u2 = 0, arena_max = 0;
    int first = 0;
    for (int64_t k = 0; k < n_pairs; ++k) {
        PairInfo& pi = pairs[(size_t)k];
        const int64_t nstrips = ((int64_t)pi.N + 1 + 32 * G2K - 1) / (32 * G2K);
        const int64_t need = nstrips * pi.nblk * 32;
        if (need > (int64_t)(free_b * 0.9) / 8) return fail(GOTOH_B200_ENOMEM, "pair %lld needs %lld bytes of tie-bit arena", (long long)k, (long long)need * 8);
        if (used_u2 > 0 && used_u2 + need > budget) { chunks.push_back({first, (int)(k - first)}); first = (int)k; used_u2 = 0; }
        pi.dir_off = used_u2;
        used_u2 += need;
        arena_max = std::max(arena_max, used_u2);
    }
    chunks.push_back({first, (int)(n_pairs - first)});
    CU(b.arena.ensure((size_t)arena_max));

    CU(cudaMemcpy(b.raw1.p, h_raw1.data(), h_raw1.size(), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(b.idx1.p, h_idx1.data(), h_idx1.size(), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(b.raw2.p, h_raw2.data(), h_raw2.size(), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(b.idx2.p, h_idx2.data(), h_idx2.size(), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(b.pairs.p, pairs.data(), n * sizeof(PairInfo), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(b.dmat.p, matrix, (size_t)l * l * sizeof(int32_t), cudaMemcpyHostToDevice));

    for (const auto& ch : chunks) {
        Params p;
        memset(&p, 0, sizeof(p));
        p.pairs = b.pairs.p; p.pair_first = ch.first; p.pair_count = ch.second;
        p.s1_idx = b.idx1.p; p.s2_idx = b.idx2.p; p.dmat = b.dmat.p;
        p.l = l; p.v = gop; p.u = gep; p.is_global = is_global ? 1 : 0;
        p.arena = b.arena.p; p.fbnd = b.fbnd.p; p.rbnd = b.rbnd.p; p.bnd_stride = bstride;
        p.best = b.best.p; p.start_i = b.si.p; p.start_j = b.sj.p;
        const int g = std::max(1, std::min(grid, (ch.second + 3) / 4));
        GOTOH_LAUNCH((k2_forward<0>), dim3(g), dim3(128), 0, (cudaStream_t)0, p);
        CU(cudaGetLastError());
        GOTOH_LAUNCH((k2_reverse<0>), dim3(g), dim3(128), 0, (cudaStream_t)0, p);
        CU(cudaGetLastError());
        WalkParams2 wp;
        memset(&wp, 0, sizeof(wp));
        wp.pairs = b.pairs.p; wp.pair_first = ch.first; wp.pair_count = ch.second;
        wp.arena = reinterpret_cast<const uint8_t*>(b.arena.p);
        wp.best = b.best.p; wp.start_i = b.si.p; wp.start_j = b.sj.p;
        wp.ops = b.ops.p; wp.nops = b.nops.p; wp.i0 = b.i0.p; wp.j0 = b.j0.p; wp.out_len = b.lenp.p; wp.score = b.score.p;
        GOTOH_LAUNCH(k2_walk, dim3((ch.second + 127) / 128), dim3(128), 0, (cudaStream_t)0, wp);
        CU(cudaGetLastError());
        EmitParams ep;
        memset(&ep, 0, sizeof(ep));
        ep.pairs = b.pairs.p; ep.pair_first = ch.first; ep.pair_count = ch.second;
        ep.ref_raw = b.raw1.p; ep.qry = b.raw2.p; ep.ops = b.ops.p; ep.nops = b.nops.p;
        ep.i0 = b.i0.p; ep.j0 = b.j0.p; ep.end_i = b.si.p; ep.end_j = b.sj.p;
        ep.out_len_plan = b.lenp.p; ep.score_plan = b.score.p;
        ep.out_ref = b.o1.p; ep.out_qry = b.o2.p; ep.out_len = b.olen.p; ep.out_score = b.oscore.p;
        GOTOH_LAUNCH(k_emit, dim3((ch.second + 3) / 4), dim3(128), 0, (cudaStream_t)0, ep);
        CU(cudaGetLastError());
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
    try {
        return g2_align_batch(device, s1_bytes, s1_off, n_s1, s1_idx, s2_bytes, s2_off, n_pairs, gop, gep, is_global,
                              alphabet, matrix, out1, out2, out_off, out_len, out_score);
    } catch (const std::bad_alloc&) {
        return fail(GOTOH_B200_ENOMEM, "out of host memory while packing");
    }
}

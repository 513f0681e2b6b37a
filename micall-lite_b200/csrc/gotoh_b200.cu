// gotoh_b200.cu - host side of libgotoh_b200.so: C ABI (include/gotoh_b200.h), input
// validation + trim/degap (gotoh.cpp:529-559), bucketing into warp tasks, HBM layout,
// kernel launches, multi-GPU static sharding.  The kernels live in gotoh_kernels.cuh.
//
// There is no CPU compute path in this file: every entry point that aligns anything
// requires a CUDA device and fails with GOTOH_B200_ENODEVICE otherwise.
#include "../../include/gotoh_b200.h"

#include "gotoh_kernels.cuh"
#include "gotoh_tables.h"
#include "gotoh_intpeak.cuh"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

using namespace gotoh;

// ---------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t _e = (call);                                                              \
        if (_e != cudaSuccess)                                                                \
            return fail(_e == cudaErrorMemoryAllocation ? GOTOH_B200_ENOMEM : GOTOH_B200_ECUDA, \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

extern "C" int32_t gotoh_b200_version(void) { return GOTOH_B200_VERSION; }
extern "C" const char* gotoh_b200_last_error(void) { return g_err; }

extern "C" int32_t gotoh_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    return n;
}

extern "C" int32_t gotoh_b200_pairscore_table(int32_t matrix_id, int32_t* out) {
    if (matrix_id < 0 || matrix_id > 2 || !out) return fail(GOTOH_B200_EINVAL, "pairscore_table: bad argument");
    const ScoreTable& t = score_table(matrix_id);
    for (int a = 0; a < 127; ++a)
        for (int b = 0; b < 127; ++b) out[a * 127 + b] = t.v[a][b];
    return GOTOH_B200_OK;
}

extern "C" void* gotoh_b200_host_alloc(int64_t bytes) {
    void* p = nullptr;
    if (bytes <= 0) return nullptr;
    if (cudaMallocHost(&p, (size_t)bytes) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
    return p;
}
extern "C" void gotoh_b200_host_free(void* p) { if (p) cudaFreeHost(p); }

// ---------------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------------
namespace {

const int kSupportedK[] = {2, 3, 4, 6, 8};
const int kMaxK = 8;

struct Launch {
    int x2;            // 0: Vec32, 1: Vec16
    int K;
    int task_first, task_count;
    int rebase_mask;
    int multi_strip;
};
struct Chunk {
    std::vector<Launch> launches;
    int pair_first, pair_count;
};

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaError_t alloc(size_t count) { n = count; return cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T)); }
    void release() { if (p) cudaFree(p); p = nullptr; }
};

}  // namespace

struct gotoh_b200_plan {
    int device = 0;
    cudaStream_t stream = 0;
    cudaEvent_t ev[4] = {0, 0, 0, 0};
    int64_t n_pairs = 0;
    int gip = 0, gep = 0, term = 1, matrix = 0;
    int ncls = 1;
    int smin_m1 = 0;
    int sm_count = 1;
    std::vector<PairInfo> pairs;
    std::vector<Task> tasks;
    std::vector<Chunk> chunks;
    int n_launches = 0;
    int64_t out_base = 0, out_bytes = 0;   // caller's out_off range covered by this plan
    int64_t pair_base = 0;                 // first caller pair index
    // device
    DevBuf<uint8_t> d_ref_raw, d_ref_cls, d_qry, d_out_ref, d_out_qry;
    DevBuf<PairInfo> d_pairs;
    DevBuf<Task> d_tasks;
    DevBuf<int32_t> d_table4, d_score, d_end_i, d_end_j, d_nops, d_i0, d_j0, d_len_plan, d_out_len, d_out_score;
    DevBuf<uint4> d_dir;
    DevBuf<int2> d_bnd;
    DevBuf<uint32_t> d_ops, d_counter;
    int64_t bnd_stride = 0;
    // stats
    int64_t cells = 0, h2d_bytes = 0, d2h_bytes = 0, arena_bytes = 0, pairs_x2 = 0, pairs_x1 = 0;

    ~gotoh_b200_plan() {
        cudaSetDevice(device);
        d_ref_raw.release(); d_ref_cls.release(); d_qry.release(); d_out_ref.release(); d_out_qry.release();
        d_pairs.release(); d_tasks.release(); d_table4.release(); d_score.release(); d_end_i.release();
        d_end_j.release(); d_nops.release(); d_i0.release(); d_j0.release(); d_len_plan.release();
        d_out_len.release(); d_out_score.release(); d_dir.release(); d_bnd.release(); d_ops.release();
        d_counter.release();
        for (int i = 0; i < 4; ++i) if (ev[i]) cudaEventDestroy(ev[i]);
        if (stream) cudaStreamDestroy(stream);
    }
};

namespace {

inline bool is_ws(uint8_t c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r'; }

// trim(): gotoh.cpp:545-559 (both ends, " \t\n\r").
inline void trim_span(const uint8_t* s, int64_t n, int64_t* lo, int64_t* hi) {
    int64_t a = 0, b = n;
    while (a < b && is_ws(s[a])) ++a;
    while (b > a && is_ws(s[b - 1])) --b;
    *lo = a; *hi = b;
}

int pick_K(int n) {
    for (int k : kSupportedK) if (32 * k >= n) return k;
    return kMaxK;
}

// "range proof" for the int16x2 path: with rebase period R every value the Vec16 kernel
// forms for real cells stays inside int16.  All quantities in stored units.
bool fits_int16(int M, int N, int K, int R, int gip, int gep, int minT, int maxT) {
    const long long mn = std::min(M, N);
    const long long smax = (long long)std::max(maxT, 0) * mn;
    const long long smin = (long long)std::min(minT, 0) * mn;
    const long long g = gep;
    const long long vmax = 4 * (smax + (R + 32LL * K + 2) * g) + 8;
    const long long vmin = 4 * (smin - 2LL * gip - gep) - 8;
    const long long add_hi = 4 * (std::max(maxT, 0) + 2 * g) + 4;
    const long long add_lo = 4LL * std::max<long long>(gip, -(long long)std::min(minT, 0)) + 8;
    if (vmax + add_hi > 32000) return false;
    if (vmin - add_lo < -32000) return false;
    if (4LL * R * g > 30000) return false;   // the rebase delta itself must be an int16
    return true;
}

struct HostPair {
    int64_t ref;      // reference index
    int M, N;
    int64_t qpos;     // position in packed query buffer
    int32_t orig;
};

template <class V, int K, bool MULTI>
int launch_forward_k(const gotoh_b200_plan* pl, const FwdParams& fp, int ntasks) {
    const size_t per_warp = FwdSmem<V, K>::per_warp(pl->ncls);
    const size_t smem = per_warp * FWD_WARPS;
    if (smem > 200 * 1024) return fail(GOTOH_B200_ERANGE, "profile needs %zu bytes of shared memory", smem);
    CU(cudaFuncSetAttribute(k_forward<V, K, MULTI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // persistent grid: enough CTAs to fill every SM, tasks are pulled from a counter
    int ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (200 * 1024) / std::max<size_t>(smem, 1)));
    int grid = std::min((ntasks + FWD_WARPS - 1) / FWD_WARPS, pl->sm_count * ctas_per_sm);
    if (MULTI) grid = std::min<long long>(grid, (long long)pl->d_bnd.n / (2 * pl->bnd_stride * FWD_WARPS));
    grid = std::max(grid, 1);
    GOTOH_LAUNCH((k_forward<V, K, MULTI>), dim3(grid), dim3(FWD_WARPS * 32), smem, pl->stream, fp);
    CU(cudaGetLastError());
    return GOTOH_B200_OK;
}

template <class V, bool MULTI>
int launch_forward(const gotoh_b200_plan* pl, const FwdParams& fp, int K, int ntasks) {
    switch (K) {
        case 2: return launch_forward_k<V, 2, MULTI>(pl, fp, ntasks);
        case 3: return launch_forward_k<V, 3, MULTI>(pl, fp, ntasks);
        case 4: return launch_forward_k<V, 4, MULTI>(pl, fp, ntasks);
        case 6: return launch_forward_k<V, 6, MULTI>(pl, fp, ntasks);
        case 8: return launch_forward_k<V, 8, MULTI>(pl, fp, ntasks);
    }
    return fail(GOTOH_B200_EINVAL, "unsupported K=%d", K);
}

int plan_build(gotoh_b200_plan* pl, const uint8_t* ref_bytes, const int64_t* ref_off, int64_t n_refs,
               const int32_t* ref_idx, const uint8_t* qry_bytes, const int64_t* qry_off,
               int64_t pair_begin, int64_t pair_end, const int64_t* out_off) {
    const int64_t n = pair_end - pair_begin;
    const ScoreTable& tab = score_table(pl->matrix);
    const bool degap = (pl->matrix == GOTOH_B200_AA_RB);
    if (degap) pl->term = 0;   // gotoh.cpp:718

    // ---- references used by this shard: trim, degap, validate, classes ------------------
    std::vector<int64_t> ref_local(n_refs, -1);
    std::vector<int64_t> used_refs;
    for (int64_t k = pair_begin; k < pair_end; ++k) {
        const int64_t r = ref_idx ? ref_idx[k] : k;
        if (r < 0 || r >= n_refs) return fail(GOTOH_B200_EINVAL, "pair %lld: ref_idx %lld out of range", (long long)k, (long long)r);
        if (ref_local[r] < 0) { ref_local[r] = (int64_t)used_refs.size(); used_refs.push_back(r); }
    }
    std::vector<uint8_t> h_ref_raw;
    std::vector<int64_t> ref_pos(used_refs.size());
    std::vector<int> ref_len(used_refs.size());
    bool ref_present[128] = {false}, qry_present[128] = {false};
    for (size_t u = 0; u < used_refs.size(); ++u) {
        const int64_t r = used_refs[u];
        const uint8_t* s = ref_bytes + ref_off[r];
        int64_t lo, hi;
        if (ref_off[r + 1] < ref_off[r]) return fail(GOTOH_B200_EINVAL, "ref_off not monotone at %lld", (long long)r);
        trim_span(s, ref_off[r + 1] - ref_off[r], &lo, &hi);
        h_ref_raw.insert(h_ref_raw.end(), REF_PAD, 0);
        ref_pos[u] = (int64_t)h_ref_raw.size();
        int m = 0, dollars = 0;
        for (int64_t x = lo; x < hi; ++x) {
            const uint8_t c = s[x];
            if (degap && c == '-') continue;                       // degap(): gotoh.cpp:529-543
            if (c < 1 || c > 126) return fail(GOTOH_B200_EDOMAIN, "reference %lld: byte 0x%02x outside 1..126", (long long)r, c);
            dollars = (c == '$') ? dollars + 1 : 0;
            if (dollars >= 3 && pl->matrix == GOTOH_B200_NT)
                return fail(GOTOH_B200_EDOMAIN, "reference %lld contains \"$$$\": the stop-codon bonus rule "
                            "(gotoh.cpp:324-344) is not implemented on the device yet", (long long)r);
            ref_present[c] = true;
            h_ref_raw.push_back(c);
            ++m;
        }
        if (m == 0) return fail(GOTOH_B200_EEMPTY, "reference %lld is empty after trim", (long long)r);
        if ((int64_t)m > (1 << 24)) return fail(GOTOH_B200_ERANGE, "reference %lld too long", (long long)r);
        ref_len[u] = m;
    }
    h_ref_raw.insert(h_ref_raw.end(), REF_PAD, 0);

    // ---- queries: trim, degap, validate, pack ---------------------------------------------
    std::vector<uint8_t> h_qry;
    std::vector<HostPair> hp((size_t)n);
    h_qry.reserve((size_t)(qry_off[pair_end] - qry_off[pair_begin]) + 64);
    for (int64_t k = pair_begin; k < pair_end; ++k) {
        const uint8_t* s = qry_bytes + qry_off[k];
        int64_t lo, hi;
        if (qry_off[k + 1] < qry_off[k]) return fail(GOTOH_B200_EINVAL, "qry_off not monotone at %lld", (long long)k);
        trim_span(s, qry_off[k + 1] - qry_off[k], &lo, &hi);
        HostPair& h = hp[(size_t)(k - pair_begin)];
        h.qpos = (int64_t)h_qry.size();
        int nn = 0;
        for (int64_t x = lo; x < hi; ++x) {
            const uint8_t c = s[x];
            if (degap && c == '-') continue;
            if (c < 1 || c > 126) return fail(GOTOH_B200_EDOMAIN, "query %lld: byte 0x%02x outside 1..126", (long long)k, c);
            qry_present[c] = true;
            h_qry.push_back(c);
            ++nn;
        }
        if (nn == 0) return fail(GOTOH_B200_EEMPTY, "query %lld is empty after trim", (long long)k);
        if (nn > (1 << 24)) return fail(GOTOH_B200_ERANGE, "query %lld too long", (long long)k);
        h.ref = ref_local[ref_idx ? ref_idx[k] : k];
        h.M = ref_len[(size_t)h.ref];
        h.N = nn;
        h.orig = (int32_t)(k - pair_begin);
        // the -100000 sentinel (gotoh.cpp:284-286): outside this bound the reference may read
        // uninitialised end indices
        const long long worst = 2LL * pl->gip + ((long long)std::max(h.M, h.N) + 1) * pl->gep;
        if (worst >= 100000 || pl->gip < 0 || pl->gep < 0 || pl->gip > 100000 || pl->gep > 100000)
            return fail(GOTOH_B200_ESENTINEL, "pair %lld: 2*gip+(max(M,N)+1)*gep = %lld is outside [0,100000)", (long long)k, worst);
        if (out_off[k + 1] - out_off[k] < (int64_t)h.M + h.N)
            return fail(GOTOH_B200_ERANGE, "pair %lld: output stride %lld < M+N = %d", (long long)k,
                        (long long)(out_off[k + 1] - out_off[k]), h.M + h.N);
        pl->cells += (int64_t)h.M * h.N;
    }
    h_qry.insert(h_qry.end(), 64, 0);

    // ---- classes of reference bytes, compact table, score range ----------------------------
    int cls_of[128];
    std::vector<int> rep(1, 0);
    for (int c = 0; c < 128; ++c) { cls_of[c] = 0; if (ref_present[c]) { cls_of[c] = (int)rep.size(); rep.push_back(c); } }
    pl->ncls = (int)rep.size();
    std::vector<int32_t> h_table4((size_t)pl->ncls * 128, 0);
    int minT = 0, maxT = 0;
    for (int c = 1; c < pl->ncls; ++c)
        for (int b = 1; b < 127; ++b) {
            const int t = tab.v[rep[c]][b];
            h_table4[(size_t)c * 128 + b] = 4 * (t + 2 * pl->gep);
            if (qry_present[b]) { minT = std::min(minT, t); maxT = std::max(maxT, t); }
        }
    std::vector<uint8_t> h_ref_cls(h_ref_raw.size());
    for (size_t x = 0; x < h_ref_raw.size(); ++x) h_ref_cls[x] = (uint8_t)cls_of[h_ref_raw[x] & 127];

    // ---- choose the path per pair, form warp tasks -------------------------------------------
    // Vec16 (two alignments per warp) needs: single strip (N <= 32*Kmax), same reference for both
    // halves, and the int16 range proof.  Everything else runs Vec32.
    int maxM = 0;
    long long smin_all = 0;
    for (const HostPair& h : hp) {
        maxM = std::max(maxM, h.M);
        smin_all = std::min(smin_all, (long long)std::min(minT, 0) * std::min(h.M, h.N));
    }
    pl->smin_m1 = (int)(smin_all - 2LL * pl->gip - pl->gep - 2);
    const int force = getenv("GOTOH_B200_FORCE_PATH") ? atoi(getenv("GOTOH_B200_FORCE_PATH")) : 0;  // tests: 32 or 16
    std::vector<int> elig;   // indices into hp
    std::vector<int> wide;
    for (size_t x = 0; x < hp.size(); ++x) {
        const HostPair& h = hp[x];
        const bool ok = h.N <= 32 * kMaxK && force != 32 &&
                        fits_int16(h.M, h.N, pick_K(h.N), 32, pl->gip, pl->gep, minT, maxT);
        (ok ? elig : wide).push_back((int)x);
    }
    int R = 32;
    if (!elig.empty()) {
        for (int cand = 4096; cand >= 32; cand >>= 1) {
            bool all = true;
            for (int x : elig)
                if (!fits_int16(hp[x].M, hp[x].N, pick_K(hp[x].N), cand, pl->gip, pl->gep, minT, maxT)) { all = false; break; }
            if (all) { R = cand; break; }
        }
    }
    // order: Vec16 pairs grouped by (K, ref, N) so that partners share the reference and have similar
    // width; Vec32 pairs by (K, cells desc)
    std::sort(elig.begin(), elig.end(), [&](int a, int b) {
        const int ka = pick_K(hp[a].N), kb = pick_K(hp[b].N);
        if (ka != kb) return ka < kb;
        if (hp[a].ref != hp[b].ref) return hp[a].M != hp[b].M ? hp[a].M > hp[b].M : hp[a].ref < hp[b].ref;
        if (hp[a].N != hp[b].N) return hp[a].N > hp[b].N;
        return a < b;
    });
    std::sort(wide.begin(), wide.end(), [&](int a, int b) {
        const int ka = pick_K(hp[a].N), kb = pick_K(hp[b].N);
        if (ka != kb) return ka < kb;
        const bool ma = hp[a].N > 32 * ka, mb = hp[b].N > 32 * kb;   // multi-strip tasks form their own launch
        if (ma != mb) return mb;
        const long long ca = (long long)hp[a].M * hp[a].N, cb = (long long)hp[b].M * hp[b].N;
        if (ca != cb) return ca > cb;
        return a < b;
    });

    pl->pairs.resize((size_t)n);
    pl->tasks.clear();
    struct TaskMeta { int x2, K; int64_t arena; int multi; };
    std::vector<TaskMeta> tmeta;
    int64_t ops_words = 0;
    int next_pair = 0;
    auto add_pair = [&](int hx, int K, int x2, int half) -> int {
        const HostPair& h = hp[hx];
        PairInfo& pi = pl->pairs[(size_t)next_pair];
        memset(&pi, 0, sizeof(pi));
        pi.ref_pos = ref_pos[(size_t)h.ref];
        pi.qry_pos = h.qpos;
        pi.out_off = out_off[pair_begin + h.orig] - out_off[pair_begin];
        pi.M = h.M; pi.N = h.N;
        pi.K = (int16_t)K; pi.x2 = (int8_t)x2; pi.half = (int8_t)half;
        pi.orig = h.orig;
        pi.nblk = x2 ? (h.M + 31 + 3) / 4 : (h.M + 31 + 7) / 8;
        if (ops_words + (h.M + h.N + 15) / 16 > 0x7fffffffLL) return -1;
        pi.ops_off = (int32_t)ops_words;
        ops_words += (h.M + h.N + 15) / 16;
        return next_pair++;
    };
    for (size_t x = 0; x < elig.size();) {
        const int a = elig[x];
        const int K = pick_K(hp[a].N);
        int b = -1;
        if (x + 1 < elig.size() && hp[elig[x + 1]].ref == hp[a].ref && pick_K(hp[elig[x + 1]].N) == K) b = elig[x + 1];
        Task t;
        t.pair_a = add_pair(a, K, 1, 0);
        t.pair_b = b >= 0 ? add_pair(b, K, 1, 1) : -1;
        if (t.pair_a < 0 || (b >= 0 && t.pair_b < 0)) return fail(GOTOH_B200_ERANGE, "op-script arena exceeds 2^31 words; split the batch");
        pl->tasks.push_back(t);
        tmeta.push_back({1, K, (int64_t)pl->pairs[(size_t)t.pair_a].nblk * 32, 0});
        pl->pairs_x2 += (b >= 0) ? 2 : 1;
        x += (b >= 0) ? 2 : 1;
    }
    for (int a : wide) {
        const int K = pick_K(hp[a].N);
        Task t;
        t.pair_a = add_pair(a, K, 0, 0);
        t.pair_b = -1;
        if (t.pair_a < 0) return fail(GOTOH_B200_ERANGE, "op-script arena exceeds 2^31 words; split the batch");
        const int nstrips = (hp[a].N + 32 * K - 1) / (32 * K);
        pl->tasks.push_back(t);
        tmeta.push_back({0, K, (int64_t)nstrips * pl->pairs[(size_t)t.pair_a].nblk * 32, nstrips > 1});
        pl->pairs_x1 += 1;
    }

    // ---- device setup ---------------------------------------------------------------------------
    CU(cudaSetDevice(pl->device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, pl->device));
    pl->sm_count = prop.multiProcessorCount;
    CU(cudaStreamCreateWithFlags(&pl->stream, cudaStreamNonBlocking));
    for (int i = 0; i < 4; ++i) CU(cudaEventCreate(&pl->ev[i]));

    pl->out_base = out_off[pair_begin];
    pl->out_bytes = out_off[pair_end] - out_off[pair_begin];
    pl->pair_base = pair_begin;

    CU(pl->d_ref_raw.alloc(h_ref_raw.size()));
    CU(pl->d_ref_cls.alloc(h_ref_cls.size()));
    CU(pl->d_qry.alloc(h_qry.size()));
    CU(pl->d_table4.alloc(h_table4.size()));
    CU(pl->d_pairs.alloc((size_t)n));
    CU(pl->d_tasks.alloc(pl->tasks.size()));
    CU(pl->d_score.alloc((size_t)n)); CU(pl->d_end_i.alloc((size_t)n)); CU(pl->d_end_j.alloc((size_t)n));
    CU(pl->d_nops.alloc((size_t)n)); CU(pl->d_i0.alloc((size_t)n)); CU(pl->d_j0.alloc((size_t)n));
    CU(pl->d_len_plan.alloc((size_t)n)); CU(pl->d_out_len.alloc((size_t)n)); CU(pl->d_out_score.alloc((size_t)n));
    CU(pl->d_ops.alloc((size_t)ops_words));
    CU(pl->d_out_ref.alloc((size_t)pl->out_bytes));
    CU(pl->d_out_qry.alloc((size_t)pl->out_bytes));
    // bytes between out_len[k] and the pair's stride are never written by k_emit: define them as 0
    CU(cudaMemsetAsync(pl->d_out_ref.p, 0, (size_t)pl->out_bytes, 0));
    CU(cudaMemsetAsync(pl->d_out_qry.p, 0, (size_t)pl->out_bytes, 0));
    CU(cudaDeviceSynchronize());

    // ---- arena budget and chunking ----------------------------------------------------------------
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    int64_t budget = (int64_t)(free_b * 0.80);
    if (getenv("GOTOH_B200_ARENA_MB")) budget = (int64_t)atoll(getenv("GOTOH_B200_ARENA_MB")) << 20;  // tests: force chunking
    int64_t biggest = 0;
    for (const TaskMeta& m : tmeta) biggest = std::max(biggest, m.arena);
    // multi-strip boundary columns: two int2 columns of maxM+1 rows per resident warp
    bool any_multi = false;
    for (const TaskMeta& m : tmeta) any_multi |= (m.multi != 0);
    if (any_multi) {
        pl->bnd_stride = ((int64_t)maxM + 2 + 15) & ~15LL;
        const int64_t slots = (int64_t)pl->sm_count * 4 * FWD_WARPS;
        CU(pl->d_bnd.alloc((size_t)(slots * 2 * pl->bnd_stride)));
        budget -= slots * 2 * pl->bnd_stride * (int64_t)sizeof(int2);
    }
    const int64_t budget_u4 = std::max<int64_t>(budget / 16, biggest);
    if (biggest * 16 > (int64_t)free_b)
        return fail(GOTOH_B200_ENOMEM, "one alignment needs %lld bytes of direction arena, device has %zu free",
                    (long long)biggest * 16, free_b);

    pl->chunks.clear();
    int64_t used = 0, arena_max = 0;
    Chunk cur; cur.pair_first = 0; cur.pair_count = 0;
    auto flush = [&]() {
        if (!cur.launches.empty()) { pl->chunks.push_back(cur); arena_max = std::max(arena_max, used); }
        cur.launches.clear();
        used = 0;
    };
    int pair_cursor = 0;
    for (size_t t = 0; t < pl->tasks.size(); ++t) {
        const TaskMeta& m = tmeta[t];
        if (used + m.arena > budget_u4) { flush(); cur.pair_first = pair_cursor; cur.pair_count = 0; }
        if (cur.launches.empty() || cur.launches.back().x2 != m.x2 || cur.launches.back().K != m.K ||
            cur.launches.back().multi_strip != m.multi) {
            Launch L; L.x2 = m.x2; L.K = m.K; L.task_first = (int)t; L.task_count = 0;
            L.rebase_mask = R - 1; L.multi_strip = m.multi;
            cur.launches.push_back(L);
        }
        Launch& L = cur.launches.back();
        L.task_count++;
        const Task& tk = pl->tasks[t];
        pl->pairs[(size_t)tk.pair_a].dir_off = used;
        if (tk.pair_b >= 0) pl->pairs[(size_t)tk.pair_b].dir_off = used;
        used += m.arena;
        const int np = tk.pair_b >= 0 ? 2 : 1;
        cur.pair_count += np;
        pair_cursor += np;
    }
    flush();
    pl->n_launches = 0;
    for (const Chunk& c : pl->chunks) pl->n_launches += (int)c.launches.size() + 2;
    CU(pl->d_dir.alloc((size_t)arena_max));
    pl->arena_bytes = arena_max * 16;
    CU(pl->d_counter.alloc((size_t)std::max(pl->n_launches, 1)));

    // ---- H2D ----------------------------------------------------------------------------------------
    auto h2d = [&](void* d, const void* h, size_t bytes) -> cudaError_t {
        pl->h2d_bytes += (int64_t)bytes;
        return cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, pl->stream);
    };
    CU(h2d(pl->d_ref_raw.p, h_ref_raw.data(), h_ref_raw.size()));
    CU(h2d(pl->d_ref_cls.p, h_ref_cls.data(), h_ref_cls.size()));
    CU(h2d(pl->d_qry.p, h_qry.data(), h_qry.size()));
    CU(h2d(pl->d_table4.p, h_table4.data(), h_table4.size() * sizeof(int32_t)));
    CU(h2d(pl->d_pairs.p, pl->pairs.data(), pl->pairs.size() * sizeof(PairInfo)));
    CU(h2d(pl->d_tasks.p, pl->tasks.data(), pl->tasks.size() * sizeof(Task)));
    CU(cudaStreamSynchronize(pl->stream));   // host staging vectors die at return
    return GOTOH_B200_OK;
}

int plan_run(gotoh_b200_plan* pl, float* device_ms, float* forward_ms) {
    CU(cudaSetDevice(pl->device));
    CU(cudaMemsetAsync(pl->d_counter.p, 0, pl->d_counter.n * sizeof(uint32_t), pl->stream));
    CU(cudaEventRecord(pl->ev[0], pl->stream));
    float fwd_total = 0.f;
    int launch_no = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> dummy;
    for (size_t ci = 0; ci < pl->chunks.size(); ++ci) {
        const Chunk& c = pl->chunks[ci];
        if (forward_ms) CU(cudaEventRecord(pl->ev[2], pl->stream));
        for (const Launch& L : c.launches) {
            FwdParams fp;
            memset(&fp, 0, sizeof(fp));
            fp.pairs = pl->d_pairs.p; fp.tasks = pl->d_tasks.p;
            fp.task_first = L.task_first; fp.task_count = L.task_count;
            fp.ref_cls = pl->d_ref_cls.p; fp.qry = pl->d_qry.p; fp.table4 = pl->d_table4.p;
            fp.ncls = pl->ncls; fp.gip = pl->gip; fp.gep = pl->gep;
            fp.rebase_mask = L.rebase_mask; fp.smin_m1 = pl->smin_m1;
            fp.dir = pl->d_dir.p;
            fp.bnd = L.multi_strip ? pl->d_bnd.p : nullptr; fp.bnd_stride = pl->bnd_stride;
            fp.score = pl->d_score.p; fp.end_i = pl->d_end_i.p; fp.end_j = pl->d_end_j.p;
            fp.work_counter = pl->d_counter.p + launch_no++;
            fp.four = 4u;
            // multi-strip tasks only exist with K = 8 (pick_K), and only on the int32 path
            const int rc = L.x2 ? launch_forward<Vec16, false>(pl, fp, L.K, L.task_count)
                         : (L.multi_strip ? launch_forward_k<Vec32, 8, true>(pl, fp, L.task_count)
                                          : launch_forward<Vec32, false>(pl, fp, L.K, L.task_count));
            if (rc) return rc;
        }
        if (forward_ms) CU(cudaEventRecord(pl->ev[3], pl->stream));
        WalkParams wp;
        memset(&wp, 0, sizeof(wp));
        wp.pairs = pl->d_pairs.p; wp.pair_first = c.pair_first; wp.pair_count = c.pair_count;
        wp.dir = reinterpret_cast<const uint32_t*>(pl->d_dir.p);
        wp.end_i = pl->d_end_i.p; wp.end_j = pl->d_end_j.p; wp.score = pl->d_score.p;
        wp.ops = pl->d_ops.p; wp.nops = pl->d_nops.p; wp.i0 = pl->d_i0.p; wp.j0 = pl->d_j0.p;
        wp.out_len = pl->d_len_plan.p; wp.gip = pl->gip; wp.gep = pl->gep; wp.term = pl->term;
        GOTOH_LAUNCH(k_walk, dim3((c.pair_count + 127) / 128), dim3(128), 0, pl->stream, wp);
        CU(cudaGetLastError());
        EmitParams ep;
        memset(&ep, 0, sizeof(ep));
        ep.pairs = pl->d_pairs.p; ep.pair_first = c.pair_first; ep.pair_count = c.pair_count;
        ep.ref_raw = pl->d_ref_raw.p; ep.qry = pl->d_qry.p; ep.ops = pl->d_ops.p; ep.nops = pl->d_nops.p;
        ep.i0 = pl->d_i0.p; ep.j0 = pl->d_j0.p; ep.end_i = pl->d_end_i.p; ep.end_j = pl->d_end_j.p;
        ep.out_len_plan = pl->d_len_plan.p; ep.score_plan = pl->d_score.p;
        ep.out_ref = pl->d_out_ref.p; ep.out_qry = pl->d_out_qry.p;
        ep.out_len = pl->d_out_len.p; ep.out_score = pl->d_out_score.p;
        GOTOH_LAUNCH(k_emit, dim3((c.pair_count + 3) / 4), dim3(128), 0, pl->stream, ep);
        CU(cudaGetLastError());
        if (forward_ms) {
            CU(cudaEventSynchronize(pl->ev[3]));
            float ms = 0.f;
            CU(cudaEventElapsedTime(&ms, pl->ev[2], pl->ev[3]));
            fwd_total += ms;
        }
    }
    CU(cudaEventRecord(pl->ev[1], pl->stream));
    CU(cudaEventSynchronize(pl->ev[1]));
    if (device_ms) CU(cudaEventElapsedTime(device_ms, pl->ev[0], pl->ev[1]));
    if (forward_ms) *forward_ms = fwd_total;
    return GOTOH_B200_OK;
}

int plan_fetch(gotoh_b200_plan* pl, uint8_t* out_ref, uint8_t* out_qry, int32_t* out_len, int32_t* out_score) {
    CU(cudaSetDevice(pl->device));
    pl->d2h_bytes = 0;
    auto d2h = [&](void* h, const void* d, size_t bytes) -> cudaError_t {
        pl->d2h_bytes += (int64_t)bytes;
        return cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, pl->stream);
    };
    CU(d2h(out_ref + pl->out_base, pl->d_out_ref.p, (size_t)pl->out_bytes));
    CU(d2h(out_qry + pl->out_base, pl->d_out_qry.p, (size_t)pl->out_bytes));
    CU(d2h(out_len + pl->pair_base, pl->d_out_len.p, (size_t)pl->n_pairs * sizeof(int32_t)));
    CU(d2h(out_score + pl->pair_base, pl->d_out_score.p, (size_t)pl->n_pairs * sizeof(int32_t)));
    CU(cudaStreamSynchronize(pl->stream));
    return GOTOH_B200_OK;
}

int check_common(const void* ref_bytes, const int64_t* ref_off, int64_t n_refs, const int32_t* ref_idx,
                 const void* qry_bytes, const int64_t* qry_off, int64_t n_pairs, int32_t matrix_id,
                 const int64_t* out_off) {
    if (!ref_bytes || !ref_off || !qry_bytes || !qry_off || !out_off) return fail(GOTOH_B200_EINVAL, "NULL pointer argument");
    if (n_pairs < 0 || n_refs < 0) return fail(GOTOH_B200_EINVAL, "negative count");
    if (n_pairs > 0x7fffffffLL) return fail(GOTOH_B200_ERANGE, "more than 2^31-1 pairs in one call");
    if (!ref_idx && n_refs != n_pairs) return fail(GOTOH_B200_EINVAL, "ref_idx is NULL but n_refs != n_pairs");
    if (matrix_id < 0 || matrix_id > 2) return fail(GOTOH_B200_EINVAL, "matrix_id %d not in {0,1,2}", matrix_id);
    return GOTOH_B200_OK;
}

}  // namespace

extern "C" int32_t gotoh_b200_plan_create(int32_t device, const uint8_t* ref_bytes, const int64_t* ref_off,
                                          int64_t n_refs, const int32_t* ref_idx, const uint8_t* qry_bytes,
                                          const int64_t* qry_off, int64_t n_pairs, int32_t gip, int32_t gep,
                                          int32_t use_terminal, int32_t matrix_id, const int64_t* out_off,
                                          gotoh_b200_plan** plan_out) {
    if (!plan_out) return fail(GOTOH_B200_EINVAL, "plan_out is NULL");
    *plan_out = nullptr;
    int rc = check_common(ref_bytes, ref_off, n_refs, ref_idx, qry_bytes, qry_off, n_pairs, matrix_id, out_off);
    if (rc) return rc;
    const int ndev = gotoh_b200_device_count();
    if (ndev <= 0) return fail(GOTOH_B200_ENODEVICE, "no CUDA device is visible; libgotoh_b200 has no CPU path");
    if (device < 0 || device >= ndev) return fail(GOTOH_B200_ENODEVICE, "device %d not present (%d visible)", device, ndev);
    gotoh_b200_plan* pl = new (std::nothrow) gotoh_b200_plan();
    if (!pl) return fail(GOTOH_B200_ENOMEM, "out of host memory");
    pl->device = device;
    pl->n_pairs = n_pairs;
    pl->gip = gip; pl->gep = gep; pl->term = use_terminal ? 1 : 0; pl->matrix = matrix_id;
    try {
        rc = plan_build(pl, ref_bytes, ref_off, n_refs, ref_idx, qry_bytes, qry_off, 0, n_pairs, out_off);
    } catch (const std::bad_alloc&) {
        rc = fail(GOTOH_B200_ENOMEM, "out of host memory while packing");
    }
    if (rc) { delete pl; return rc; }
    *plan_out = pl;
    return GOTOH_B200_OK;
}

extern "C" int32_t gotoh_b200_plan_run(gotoh_b200_plan* plan, float* device_ms, float* forward_ms) {
    if (!plan) return fail(GOTOH_B200_EINVAL, "plan is NULL");
    return plan_run(plan, device_ms, forward_ms);
}

extern "C" int32_t gotoh_b200_plan_fetch(gotoh_b200_plan* plan, uint8_t* out_ref, uint8_t* out_qry,
                                         int32_t* out_len, int32_t* out_score) {
    if (!plan || !out_ref || !out_qry || !out_len || !out_score) return fail(GOTOH_B200_EINVAL, "NULL argument");
    return plan_fetch(plan, out_ref, out_qry, out_len, out_score);
}

extern "C" void gotoh_b200_plan_destroy(gotoh_b200_plan* plan) { delete plan; }

extern "C" int64_t gotoh_b200_plan_stat(const gotoh_b200_plan* pl, int32_t what) {
    if (!pl) return -1;
    switch (what) {
        case 0: return pl->cells;
        case 1: return pl->n_launches;
        case 2: return pl->h2d_bytes;
        case 3: return pl->d2h_bytes;
        case 4: return pl->arena_bytes;
        case 5: return pl->pairs_x2;
        case 6: return pl->pairs_x1;
        case 7: return (int64_t)pl->chunks.size();
    }
    return -1;
}

// One-shot form.  Shards contiguous pair ranges of (nearly) equal cell count across the
// devices in device_mask; one host thread per device; no inter-device traffic.
extern "C" int32_t gotoh_b200_align_batch(const uint8_t* ref_bytes, const int64_t* ref_off, int64_t n_refs,
                                          const int32_t* ref_idx, const uint8_t* qry_bytes,
                                          const int64_t* qry_off, int64_t n_pairs, int32_t gip, int32_t gep,
                                          int32_t use_terminal, int32_t matrix_id, uint8_t* out_ref,
                                          uint8_t* out_qry, const int64_t* out_off, int32_t* out_len,
                                          int32_t* out_score, uint32_t device_mask) {
    int rc = check_common(ref_bytes, ref_off, n_refs, ref_idx, qry_bytes, qry_off, n_pairs, matrix_id, out_off);
    if (rc) return rc;
    if (!out_ref || !out_qry || !out_len || !out_score) return fail(GOTOH_B200_EINVAL, "NULL output pointer");
    const int ndev = gotoh_b200_device_count();
    if (ndev <= 0) return fail(GOTOH_B200_ENODEVICE, "no CUDA device is visible; libgotoh_b200 has no CPU path");
    if (n_pairs == 0) return GOTOH_B200_OK;
    std::vector<int> devs;
    if (device_mask == 0) device_mask = 1;
    for (int d = 0; d < 32; ++d)
        if (device_mask & (1u << d)) {
            if (d >= ndev) return fail(GOTOH_B200_ENODEVICE, "device %d in device_mask not present (%d visible)", d, ndev);
            devs.push_back(d);
        }
    // contiguous split balanced by an O(n) cell estimate (untrimmed lengths)
    const int D = (int)std::min<int64_t>((int64_t)devs.size(), n_pairs);
    std::vector<int64_t> cut(D + 1, 0);
    {
        std::vector<double> pre((size_t)n_pairs + 1, 0.0);
        for (int64_t k = 0; k < n_pairs; ++k) {
            const int64_t r = ref_idx ? ref_idx[k] : k;
            const double m = (r >= 0 && r < n_refs) ? (double)(ref_off[r + 1] - ref_off[r]) : 1.0;
            pre[(size_t)k + 1] = pre[(size_t)k] + std::max(1.0, m) * std::max<double>(1.0, (double)(qry_off[k + 1] - qry_off[k]));
        }
        for (int d = 1; d < D; ++d) {
            const double target = pre[(size_t)n_pairs] * d / D;
            cut[d] = std::lower_bound(pre.begin(), pre.end(), target) - pre.begin();
            cut[d] = std::max(cut[d], cut[d - 1] + 1);
            cut[d] = std::min<int64_t>(cut[d], n_pairs - (D - d));
        }
        cut[D] = n_pairs;
    }
    std::vector<int> rcs(D, 0);
    std::vector<std::string> msgs(D);
    auto work = [&](int d) {
        gotoh_b200_plan* pl = new (std::nothrow) gotoh_b200_plan();
        if (!pl) { rcs[d] = GOTOH_B200_ENOMEM; msgs[d] = "out of host memory"; return; }
        pl->device = devs[d];
        pl->n_pairs = cut[d + 1] - cut[d];
        pl->gip = gip; pl->gep = gep; pl->term = use_terminal ? 1 : 0; pl->matrix = matrix_id;
        int r;
        try {
            r = plan_build(pl, ref_bytes, ref_off, n_refs, ref_idx, qry_bytes, qry_off, cut[d], cut[d + 1], out_off);
            if (!r) r = plan_run(pl, nullptr, nullptr);
            if (!r) r = plan_fetch(pl, out_ref, out_qry, out_len, out_score);
        } catch (const std::bad_alloc&) {
            r = fail(GOTOH_B200_ENOMEM, "out of host memory while packing");
        }
        rcs[d] = r;
        if (r) msgs[d] = g_err;
        delete pl;
    };
    if (D == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int d = 0; d < D; ++d) th.emplace_back(work, d);
        for (auto& t : th) t.join();
    }
    for (int d = 0; d < D; ++d)
        if (rcs[d]) return fail(rcs[d], "device %d: %s", devs[d], msgs[d].c_str());
    return GOTOH_B200_OK;
}

extern "C" int32_t gotoh_b200_int_peak(int32_t device, int32_t which, double* ginstr_per_s) {
    if (!ginstr_per_s) return fail(GOTOH_B200_EINVAL, "NULL argument");
    const int ndev = gotoh_b200_device_count();
    if (ndev <= 0 || device < 0 || device >= ndev) return fail(GOTOH_B200_ENODEVICE, "device %d not present", device);
    CU(cudaSetDevice(device));
    return intpeak::run(which, ginstr_per_s, g_err, sizeof(g_err));
}

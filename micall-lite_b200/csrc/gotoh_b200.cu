// gotoh_b200.cu - host side of libgotoh_b200.so: C ABI (include/gotoh_b200.h), input
// validation + trim/degap (gotoh.cpp:529-559), bucketing into warp tasks, HBM layout,
// kernel launches, the slab pipeline of the one-shot call, multi-GPU static sharding.
// The kernels live in gotoh_kernels.cuh.
//
// There is no CPU compute path in this file: every entry point that aligns anything
// requires a CUDA device and fails with GOTOH_B200_ENODEVICE otherwise.
#include "../../include/gotoh_b200.h"

#include "gotoh_kernels.cuh"
#include "gotoh_prep.cuh"
#include "gotoh_tables.h"
#include "gotoh_intpeak.cuh"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <unordered_map>
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
// workspaces: grow-only device buffers + pinned staging + one stream
// ---------------------------------------------------------------------------------------
namespace {

const int kSupportedK[] = {2, 3, 4, 6, 8};
const int kMaxK = 8;

template <class T>
struct DevBuf {            // grow-only device array
    T* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t count) {
        if (count <= cap && p) return cudaSuccess;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        const size_t want = std::max<size_t>(count + count / 8, 16);
        cudaError_t e = cudaMalloc(&p, want * sizeof(T));
        if (e != cudaSuccess) {          // retry without head-room
            (void)cudaGetLastError();
            e = cudaMalloc(&p, std::max<size_t>(count, 16) * sizeof(T));
            if (e != cudaSuccess) { p = nullptr; return e; }
            cap = std::max<size_t>(count, 16);
            return cudaSuccess;
        }
        cap = want;
        return cudaSuccess;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

template <class T>
struct PinBuf {            // grow-only page-locked host array (H2D staging)
    T* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t count) {
        if (count <= cap && p) return cudaSuccess;
        if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
        const size_t want = std::max<size_t>(count + count / 8, 64);
        cudaError_t e = cudaMallocHost(&p, want * sizeof(T));
        if (e != cudaSuccess) { p = nullptr; return e; }
        cap = want;
        return cudaSuccess;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct Workspace {
    int device = 0;
    bool ready = false;
    cudaStream_t stream = 0;
    cudaEvent_t ev[4] = {0, 0, 0, 0};
    cudaEvent_t ev_fwd = 0;    // one-shot pipeline: recorded after the forward launches of the slab this workspace holds
    int trace_slab = -1;       // GOTOH_B200_TRACE: slab whose events ev[0..2] are pending
    int sm_count = 1;
    // device
    DevBuf<uint8_t> d_ref_raw, d_ref_cls, d_qry, d_out_ref, d_out_qry;
    DevBuf<PairInfo> d_pairs;
    DevBuf<Task> d_tasks;
    DevBuf<int32_t> d_table4, d_score, d_end_i, d_end_j, d_nops, d_i0, d_j0, d_len_plan, d_out_len, d_out_score;
    DevBuf<uint4> d_dir;
    DevBuf<int2> d_bnd;
    DevBuf<Task> d_stasks;          // strip tasks of the dataflow kernel
    DevBuf<int32_t> d_flow;         // [3][slots]: progress, last-row partial score, its column
    PinBuf<Task> h_stasks;
    DevBuf<uint32_t> d_ops, d_counter;
    // tight / compact result forms: caller-order result sizes, their prefix sums, compact records and packed op scripts
    DevBuf<int32_t> d_scan_in, d_rec;
    DevBuf<int64_t> d_scan_out;
    DevBuf<uint32_t> d_cops;
    // device-side plan builder (gotoh_prep.cuh): one carved scratch block, per-reference tables, summary in mapped pinned memory
    DevBuf<uint8_t> d_prep;
    PinBuf<uint8_t> h_prep;
    PinBuf<PrepSummary> h_prep_sum;
    cudaEvent_t ev_p = 0;
    cudaEvent_t ev_done = 0;        // one-shot pipeline: recorded after a slab's last enqueued operation; the workspace is free once it fires
    bool ev_done_pending = false;
    PinBuf<int64_t> h_total;        // the slab's total (bytes / words), written by k_scan / k_publish into mapped pinned memory
    cudaEvent_t ev_a = 0;           // recorded after phase A: the host may read h_total and enqueue the result copy (phase B)
    // pinned staging
    PinBuf<uint8_t> h_ref_raw, h_ref_cls, h_qry;
    PinBuf<PairInfo> h_pairs;
    PinBuf<Task> h_tasks;
    PinBuf<int32_t> h_table4;

    int init(int dev) {
        if (ready) return GOTOH_B200_OK;
        device = dev;
        CU(cudaSetDevice(dev));
        cudaDeviceProp prop;
        CU(cudaGetDeviceProperties(&prop, dev));
        sm_count = prop.multiProcessorCount;
        CU(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        for (int i = 0; i < 4; ++i) CU(cudaEventCreate(&ev[i]));
        CU(cudaEventCreateWithFlags(&ev_fwd, cudaEventDisableTiming));
        // events the host waits on.  Spin-waiting is the default: with cudaEventBlockingSync every wake-up cost ~0.2 ms on
        // B200 hosts, which a pipeline of 1.2 ms slabs cannot afford (C2 one-shot call 153 -> 177 ms).  GOTOH_B200_BLOCKING_SYNC=1
        // trades that latency for idle cores on hosts with fewer cores than waiting threads.
        const unsigned evf = cudaEventDisableTiming | ((getenv("GOTOH_B200_BLOCKING_SYNC") && atoi(getenv("GOTOH_B200_BLOCKING_SYNC"))) ? cudaEventBlockingSync : 0);
        CU(cudaEventCreateWithFlags(&ev_a, evf));
        CU(cudaEventCreateWithFlags(&ev_p, evf));
        CU(cudaEventCreateWithFlags(&ev_done, evf));
        CU(h_total.ensure(8));
        CU(h_prep_sum.ensure(1));
        ready = true;
        return GOTOH_B200_OK;
    }
    void release() {
        if (!ready) return;
        cudaSetDevice(device);
        if (stream) cudaStreamSynchronize(stream);
        d_ref_raw.release(); d_ref_cls.release(); d_qry.release(); d_out_ref.release(); d_out_qry.release();
        d_pairs.release(); d_tasks.release(); d_table4.release(); d_score.release(); d_end_i.release();
        d_end_j.release(); d_nops.release(); d_i0.release(); d_j0.release(); d_len_plan.release();
        d_out_len.release(); d_out_score.release(); d_dir.release(); d_bnd.release(); d_stasks.release(); d_flow.release(); h_stasks.release(); d_ops.release();
        d_counter.release();
        d_scan_in.release(); d_rec.release(); d_scan_out.release(); d_cops.release(); h_total.release();
        if (ev_a) { cudaEventDestroy(ev_a); ev_a = 0; }
        if (ev_p) { cudaEventDestroy(ev_p); ev_p = 0; }
        if (ev_done) { cudaEventDestroy(ev_done); ev_done = 0; }
        d_prep.release(); h_prep.release(); h_prep_sum.release();
        h_ref_raw.release(); h_ref_cls.release(); h_qry.release(); h_pairs.release(); h_tasks.release(); h_table4.release();
        for (int i = 0; i < 4; ++i) if (ev[i]) { cudaEventDestroy(ev[i]); ev[i] = 0; }
        if (ev_fwd) { cudaEventDestroy(ev_fwd); ev_fwd = 0; }
        if (stream) { cudaStreamDestroy(stream); stream = 0; }
        ready = false;
    }
    ~Workspace() { release(); }
};

struct Launch {
    int x2;            // 0: Vec32, 1: Vec16
    int K;
    int task_first, task_count;
    int rebase_mask;
    int multi_strip;
    int hw = 0;            // Vec16 on half-warp wavefronts: tasks are taken in twos
    int nslots = 0;        // multi-strip launches: (pair, strip) slots = strip tasks of the dataflow kernel
    int stask_first = 0;   // first strip task in d_stasks
};
struct Chunk {
    std::vector<Launch> launches;
    int pair_first, pair_count;
};

}  // namespace

enum { OUT_STRIDED = 0, OUT_TIGHT = 1, OUT_COMPACT = 2 };   // result forms (include/gotoh_b200.h)

struct gotoh_b200_plan {
    Workspace* ws = nullptr;
    int out_mode = OUT_STRIDED;
    bool built_on_device = false;          // the device plan builder (gotoh_prep.cuh) laid this plan out
    int64_t ops_words = 0;                 // op-script words of all pairs (capacity of d_ops / d_cops)
    bool owns_ws = false;
    int64_t n_pairs = 0;
    int gip = 0, gep = 0, term = 1, matrix = 0;
    int ncls = 1;
    int has_dollar = 0;
    int smin_m1 = 0;
    int zshift = 0;        // Vec16 frame shift (fits_int16)
    bool mark_forward_done = false;   // one-shot pipeline: record ws->ev_fwd after the last forward launch
    std::vector<Chunk> chunks;
    int n_launches = 0;
    int64_t out_base = 0, out_bytes = 0;   // caller's out_off range covered by this plan
    int64_t pair_base = 0;                 // first caller pair index
    int64_t bnd_stride = 0;
    int flow_slots = 0;                    // > 0: the strip-dataflow kernel can run (boundary columns allocated)
    int task_limit = 0;                    // one-shot pipeline: tasks per warp before a forward CTA retires (0 = persistent)
    int64_t n_stasks = 0;
    int64_t arena_budget_bytes = 0;        // 0: 80 % of free memory
    // stats
    int64_t cells = 0, h2d_bytes = 0, d2h_bytes = 0, arena_bytes = 0, pairs_x2 = 0, pairs_x1 = 0;

    ~gotoh_b200_plan() {
        if (ws && owns_ws) delete ws;
    }
};

namespace {

inline bool is_ws(uint8_t c) { return plan_is_ws(c); }

// trim(): gotoh.cpp:545-559 (both ends, " \t\n\r").
inline void trim_span(const uint8_t* s, int64_t n, int64_t* lo, int64_t* hi) {
    int64_t a = 0, b = n;
    while (a < b && is_ws(s[a])) ++a;
    while (b > a && is_ws(s[b - 1])) --b;
    *lo = a; *hi = b;
}

// per-pair decisions shared with the device builder: gotoh_plan_math.h
thread_local bool t_half_off = false;    // GOTOH_B200_HALF=0 pins the 32-lane kernels (tests, A/B measurements); read once per plan_build
int pick_K(int n) { return plan_pick_K(n); }
int pick_KH(int n) { return plan_pick_KH(n, t_half_off); }
long long int16_low_need(int M, int N, int gip, int gep, int minT) { return plan_int16_low_need(M, N, gip, gep, minT); }
bool fits_int16(int M, int N, int K, int R, int gip, int gep, int minT, int maxT, long long z4) {
    return plan_fits_int16(M, N, K, R, gip, gep, minT, maxT, z4);
}

struct HostPair {
    int32_t ref;      // local reference index
    int32_t M, N;
    int32_t orig;
    int64_t qpos;     // position in the packed query buffer
    int64_t src_lo;   // trimmed span inside the caller's query bytes
};

struct KeyIdx {
    uint64_t key;
    uint32_t idx;
    bool operator<(const KeyIdx& o) const { return key != o.key ? key < o.key : idx < o.idx; }
};

// Stable LSD radix sort of (key, idx) records by key, 16 bits per pass, skipping passes whose digit is constant
// (std::sort was the largest single item of plan_build on batches of short pairs).  Records enter in idx order, so
// equal keys stay in idx order - the same total order as KeyIdx::operator<.
void radix_sort(std::vector<KeyIdx>& v) {
    const size_t n = v.size();
    if (n < 2048) { std::sort(v.begin(), v.end()); return; }
    std::vector<KeyIdx> tmp(n);
    std::vector<uint32_t> cnt(65536);
    KeyIdx* src = v.data();
    KeyIdx* dst = tmp.data();
    for (int pass = 0; pass < 4; ++pass) {
        const int sh = 16 * pass;
        std::fill(cnt.begin(), cnt.end(), 0u);
        for (size_t i = 0; i < n; ++i) ++cnt[(src[i].key >> sh) & 0xffff];
        if (cnt[(src[0].key >> sh) & 0xffff] == n) continue;        // every record has the same digit
        uint32_t sum = 0;
        for (size_t d = 0; d < 65536; ++d) { const uint32_t c = cnt[d]; cnt[d] = sum; sum += c; }
        for (size_t i = 0; i < n; ++i) dst[cnt[(src[i].key >> sh) & 0xffff]++] = src[i];
        std::swap(src, dst);
    }
    if (src != v.data()) memcpy(v.data(), src, n * sizeof(KeyIdx));
}

// Run fn(lo, hi, tid) over [0, n) on up to `threads` host threads.
template <class F>
void parallel_for(int64_t n, int threads, F fn) {
    threads = (int)std::max<int64_t>(1, std::min<int64_t>(threads, n));
    if (threads == 1) { fn((int64_t)0, n, 0); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t)
        th.emplace_back([=]() { fn(n * t / threads, n * (t + 1) / threads, t); });
    for (auto& x : th) x.join();
}

// GOTOH_B200_TRACE=1: per-slab host timings of the one-shot call on stderr (diagnostics only)
inline bool trace_on() { static const bool on = getenv("GOTOH_B200_TRACE") != nullptr; return on; }
inline double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
thread_local double g_trace_phase[8];

int host_threads(int64_t bytes) {
    // one thread per MB of input, at most 16 (or GOTOH_B200_HOST_THREADS): spawning 16 threads for a 4 MB slab costs
    // more CPU than the packing itself, and under torchrun the ranks share the host's cores
    if (bytes < (1 << 20)) return 1;
    const char* e = getenv("GOTOH_B200_HOST_THREADS");
    int hw = e ? atoi(e) : (int)std::thread::hardware_concurrency();
    hw = std::max(1, std::min(hw, 16));
    return (int)std::max<int64_t>(1, std::min<int64_t>(hw, bytes >> 20));
}

template <class V, int K, bool MULTI, bool HALF = false>
int launch_forward_k(const gotoh_b200_plan* pl, FwdParams fp, int ntasks) {
    const Workspace* ws = pl->ws;
    const size_t per_warp = FwdSmem<V, K, HALF>::per_warp(pl->ncls);
    // warps per CTA: 4 unless the query profile (1 KB per class per warp at K = 8) needs more room
    int warps = FWD_WARPS;
    while (warps > 1 && per_warp * warps > 200 * 1024) warps >>= 1;
    // half-warp kernels: a CTA-wide int16 copy of the score table for the per-task profile build, when it is small
    // (amino-acid tables: 21 classes = 5 KB) and no stop-codon bonus applies
    fp.stab_bytes = (HALF && !pl->has_dollar && pl->ncls <= 32) ? pl->ncls * 256 : 0;
    const size_t smem = per_warp * warps + fp.stab_bytes;
    if (smem > 220 * 1024) return fail(GOTOH_B200_ERANGE, "profile needs %zu bytes of shared memory", smem);
    CU(cudaFuncSetAttribute((k_forward<V, K, MULTI, HALF>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // persistent grid: as many CTAs as are resident at once (asked from the runtime: a 21-class amino-acid profile at K = 6
    // takes 68 KB per CTA and three of them fit the SM's 228 KB, which a fixed 200 KB budget got wrong), tasks are pulled
    // from a counter
    int ctas_per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, (k_forward<V, K, MULTI, HALF>), warps * 32, smem) != cudaSuccess || ctas_per_sm < 1) {
        (void)cudaGetLastError();
        ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(GOTOH_MIN_CTAS, (200 * 1024) / std::max<size_t>(smem, 1)));
    }
    ctas_per_sm = std::min(ctas_per_sm, (int)GOTOH_MIN_CTAS);
    int grid = std::min((ntasks + warps - 1) / warps, ws->sm_count * ctas_per_sm);
    // task_limit > 0 (one-shot pipeline): CTAs retire after task_limit tasks per warp, so the short traceback / emit
    // kernels of the previous slab (other streams) get SM slots while this launch is still running
    if (fp.task_limit > 0 && !MULTI) grid = (ntasks + warps * fp.task_limit - 1) / (warps * fp.task_limit);
    if (MULTI) grid = std::min<long long>(grid, (long long)ws->d_bnd.cap / (2 * pl->bnd_stride * warps));
    grid = std::max(grid, 1);
    GOTOH_LAUNCH((k_forward<V, K, MULTI, HALF>), dim3(grid), dim3(warps * 32), smem, ws->stream, fp);
    CU(cudaGetLastError());
    return GOTOH_B200_OK;
}

int launch_forward_cta(const gotoh_b200_plan* pl, FwdParams fp, int ntasks) {
    const Workspace* ws = pl->ws;
    const size_t per_warp = (size_t)pl->ncls * 2 * 32 * 16 + 256 * sizeof(int2);    // K = 8 profile + boundary ring
    const size_t smem = per_warp * FWD_WARPS;
    if (smem > 200 * 1024) return -1;       // caller falls back to the warp-serial multi-strip kernel
    CU(cudaFuncSetAttribute(k_forward_cta<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(GOTOH_MIN_CTAS, (200 * 1024) / smem));
    int grid = std::max(1, std::min(ntasks, ws->sm_count * ctas_per_sm));
    grid = (int)std::min<long long>(grid, std::max<long long>(1, (long long)ws->d_bnd.cap / (2 * pl->bnd_stride)));
    GOTOH_LAUNCH((k_forward_cta<8>), dim3(grid), dim3(FWD_WARPS * 32), smem, ws->stream, fp);
    CU(cudaGetLastError());
    return GOTOH_B200_OK;
}

// K2 (default): strip dataflow, one warp per (pair, strip)
int launch_forward_flow(const gotoh_b200_plan* pl, const FwdParams& fp, int ntasks) {
    const Workspace* ws = pl->ws;
    const size_t per_warp = FwdSmem<Vec32, 8>::per_warp(pl->ncls);
    if (per_warp * FWD_WARPS > 220 * 1024) return fail(GOTOH_B200_ERANGE, "profile needs %zu bytes of shared memory", per_warp * FWD_WARPS);
    CU(cudaFuncSetAttribute(k_forward_flow<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(per_warp * FWD_WARPS)));
    // The kernel is latency-bound (dependent int32 cell chain), so resident warps per SM are what counts: 16 by registers,
    // but the per-warp query profile (1 KB per reference byte class) usually caps it lower.  Ask the runtime how many CTAs
    // of 4, 3 or 2 warps fit and take the shape with the most warps (HCV genomes, 14 classes: 4 x 3 = 12 -> 3 x 5 = 15).
    int warps = FWD_WARPS, ctas_per_sm = 1, best = 0;
    for (int w = FWD_WARPS; w >= 2; --w) {
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_forward_flow<8>, w * 32, per_warp * w) != cudaSuccess) { (void)cudaGetLastError(); continue; }
        if (nb * w > best) { best = nb * w; warps = w; ctas_per_sm = nb; }
    }
    if (best == 0) ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(GOTOH_MIN_CTAS, (200 * 1024) / std::max<size_t>(per_warp * FWD_WARPS, 1)));
    const size_t smem = per_warp * warps;
    const int grid = std::max(1, std::min((ntasks + warps - 1) / warps, ws->sm_count * ctas_per_sm));
    GOTOH_LAUNCH((k_forward_flow<8>), dim3(grid), dim3(warps * 32), smem, ws->stream, fp);
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

int launch_forward_half(const gotoh_b200_plan* pl, const FwdParams& fp, int K, int nwarp_tasks) {
    FwdParams q = fp;
    q.task_count = nwarp_tasks;          // the kernel counts warps' worth of work: two task entries each
    switch (K) {
        case 2: return launch_forward_k<Vec16, 2, false, true>(pl, q, nwarp_tasks);
        case 3: return launch_forward_k<Vec16, 3, false, true>(pl, q, nwarp_tasks);
        case 4: return launch_forward_k<Vec16, 4, false, true>(pl, q, nwarp_tasks);
        case 6: return launch_forward_k<Vec16, 6, false, true>(pl, q, nwarp_tasks);
        case 8: return launch_forward_k<Vec16, 8, false, true>(pl, q, nwarp_tasks);
    }
    return fail(GOTOH_B200_EINVAL, "unsupported K=%d", K);
}

// The references one plan uses: trimmed (+ degapped), validated, copied once into the pinned staging with REF_PAD zero
// bytes on both sides, their byte classes and the compact score table.  O(reference bytes); shared by the host and the
// device plan builders.
struct RefSet {
    std::vector<int32_t> ref_local;   // caller reference index -> local index (-1: unused); empty when ref_idx == NULL
    std::vector<int64_t> used_refs;   // local index -> caller reference index
    std::vector<int64_t> ref_pos;     // local index -> position of row 1 in h_ref_raw / h_ref_cls
    std::vector<int32_t> ref_len;     // local index -> trimmed (degapped) length M
    size_t ref_total = 0;
    bool any_dollar3 = false;
    std::vector<int> rep, rep_mask;   // class -> reference byte, stop-codon rule mask
};

int build_refs(gotoh_b200_plan* pl, const uint8_t* ref_bytes, const int64_t* ref_off, int64_t n_refs, const int32_t* ref_idx,
               int64_t pair_begin, int64_t pair_end, bool degap, RefSet& rs) {
    Workspace* ws = pl->ws;
    const int64_t n = pair_end - pair_begin;
    const ScoreTable& tab = score_table(pl->matrix);
    std::vector<int32_t>& ref_local = rs.ref_local;
    std::vector<int64_t>& used_refs = rs.used_refs;
    if (ref_idx) {
        ref_local.assign((size_t)n_refs, -1);
        for (int64_t k = pair_begin; k < pair_end; ++k) {
            const int64_t r = ref_idx[k];
            if (r < 0 || r >= n_refs) return fail(GOTOH_B200_EINVAL, "pair %lld: ref_idx %lld out of range", (long long)k, (long long)r);
            if (ref_local[(size_t)r] < 0) { ref_local[(size_t)r] = (int32_t)used_refs.size(); used_refs.push_back(r); }
        }
    } else {
        used_refs.resize((size_t)n);
        for (int64_t k = 0; k < n; ++k) used_refs[(size_t)k] = pair_begin + k;
    }
    const size_t nu = used_refs.size();
    std::vector<int64_t>& ref_pos = rs.ref_pos;
    std::vector<int32_t>& ref_len = rs.ref_len;
    std::vector<int64_t> ref_lo(nu);
    ref_pos.resize(nu); ref_len.resize(nu);
    size_t ref_total = REF_PAD;
    for (size_t u = 0; u < nu; ++u) {
        const int64_t r = used_refs[u];
        if (ref_off[r + 1] < ref_off[r]) return fail(GOTOH_B200_EINVAL, "ref_off not monotone at %lld", (long long)r);
        int64_t lo, hi;
        trim_span(ref_bytes + ref_off[r], ref_off[r + 1] - ref_off[r], &lo, &hi);
        ref_lo[u] = ref_off[r] + lo;
        if (hi - lo >= (1 << 24)) return fail(GOTOH_B200_ERANGE, "reference %lld too long", (long long)r);
        ref_len[u] = (int32_t)(hi - lo);             // upper bound; degap may shrink it below
        ref_pos[u] = (int64_t)ref_total;
        ref_total += (size_t)(hi - lo) + REF_PAD;
    }
    rs.ref_total = ref_total;
    CU(ws->h_ref_raw.ensure(ref_total));
    CU(ws->h_ref_cls.ensure(ref_total));
    uint8_t* h_ref_raw = ws->h_ref_raw.p;
    memset(h_ref_raw, 0, ref_total);
    bool any_dollar3 = false;
    for (size_t u = 0; u < nu; ++u) {
        const uint8_t* s = ref_bytes + ref_lo[u];
        uint8_t* dst = h_ref_raw + ref_pos[u];
        const int len = ref_len[u];
        int m = 0, dollars = 0;
        for (int x = 0; x < len; ++x) {
            const uint8_t c = s[x];
            if (degap && c == '-') continue;                       // degap(): gotoh.cpp:529-543
            if (c < 1 || c > 126) return fail(GOTOH_B200_EDOMAIN, "reference %lld: byte 0x%02x outside 1..126", (long long)used_refs[u], c);
            dollars = (c == '$') ? dollars + 1 : 0;
            if (dollars >= 3) any_dollar3 = true;                  // stop-codon bonus rule applies (gotoh.cpp:324-344)
            dst[m++] = c;
        }
        if (m == 0) return fail(GOTOH_B200_EEMPTY, "reference %lld is empty after trim", (long long)used_refs[u]);
        ref_len[u] = m;
    }
    rs.any_dollar3 = any_dollar3;
    // classes of reference bytes, compact table.  A class is a distinct reference byte - or, in references that contain
    // "$$$", a distinct (byte, rmask) where rmask says which of the three stop-codon bonus rules (gotoh.cpp:324-344)
    // can fire in that row: bit0 a[i-3..i-1], bit1 a[i-2..i], bit2 a[i-1..i+1] == "$$$" (i >= 3).
    // The +6 bonuses then live in the query profile and the kernel needs no extra work per cell.
    uint8_t* h_ref_cls = ws->h_ref_cls.p;
    memset(h_ref_cls, 0, ref_total);
    int cls_of[128][8];
    memset(cls_of, 0, sizeof(cls_of));
    std::vector<int>& rep = rs.rep;
    std::vector<int>& rep_mask = rs.rep_mask;
    rep.assign(1, 0); rep_mask.assign(1, 0);
    for (size_t u = 0; u < nu; ++u) {
        const uint8_t* a = h_ref_raw + ref_pos[u];
        const int M = ref_len[u];
        for (int i = 1; i <= M; ++i) {
            int rm = 0;
            if (any_dollar3 && i >= 3) {
                auto dol = [&](int p0) { return p0 >= 0 && p0 + 2 < M && a[p0] == '$' && a[p0 + 1] == '$' && a[p0 + 2] == '$'; };
                rm = (dol(i - 3) ? 1 : 0) | (dol(i - 2) ? 2 : 0) | (dol(i - 1) ? 4 : 0);
            }
            const int c = a[i - 1];
            if (!cls_of[c][rm]) {
                if (rep.size() >= 250) return fail(GOTOH_B200_ERANGE, "more than 249 reference byte classes");
                cls_of[c][rm] = (int)rep.size(); rep.push_back(c); rep_mask.push_back(rm);
            }
            h_ref_cls[ref_pos[u] + i - 1] = (uint8_t)cls_of[c][rm];
        }
    }
    pl->ncls = (int)rep.size();
    pl->has_dollar = any_dollar3 ? 1 : 0;
    CU(ws->h_table4.ensure((size_t)pl->ncls * 136));
    int32_t* h_table4 = ws->h_table4.p;           // [ncls][128] scores, then [ncls][8] bonuses
    memset(h_table4, 0, (size_t)pl->ncls * 136 * sizeof(int32_t));
    // entry 0 of every class row (no query byte is 0): 4u, what a padding column beyond N scores
    for (int c = 0; c < pl->ncls; ++c) h_table4[(size_t)c * 128] = -4 * pl->gip;
    for (int c = 1; c < pl->ncls; ++c) {
        for (int b = 1; b < 127; ++b) h_table4[(size_t)c * 128 + b] = 4 * (tab.v[rep[(size_t)c]][b] + 2 * pl->gep);
        for (int cm = 0; cm < 8; ++cm) {
            int bits = rep_mask[(size_t)c] & cm, cnt = 0;
            while (bits) { cnt += bits & 1; bits >>= 1; }
            h_table4[(size_t)pl->ncls * 128 + (size_t)c * 8 + cm] = 4 * 6 * cnt;
        }
    }
    return GOTOH_B200_OK;
}

// Validate + trim (+ degap) + pack one contiguous range of caller pairs into the workspace's
// pinned staging, choose the kernel path per pair, form warp tasks and arena chunks, and
// enqueue the H2D copies on the workspace stream (no synchronisation).
int plan_build(gotoh_b200_plan* pl, const uint8_t* ref_bytes, const int64_t* ref_off, int64_t n_refs,
               const int32_t* ref_idx, const uint8_t* qry_bytes, const int64_t* qry_off,
               int64_t pair_begin, int64_t pair_end, const int64_t* out_off) {
    Workspace* ws = pl->ws;
    const int64_t n = pair_end - pair_begin;
    const ScoreTable& tab = score_table(pl->matrix);
    const bool degap = (pl->matrix == GOTOH_B200_AA_RB);
    if (degap) pl->term = 0;   // gotoh.cpp:718
    if (pl->gip < 0 || pl->gep < 0 || pl->gip > 100000 || pl->gep > 100000)
        return fail(GOTOH_B200_ESENTINEL, "gap penalties (%d, %d) outside [0, 100000]", pl->gip, pl->gep);
    CU(cudaSetDevice(ws->device));
    pl->cells = 0; pl->h2d_bytes = 0; pl->pairs_x1 = pl->pairs_x2 = 0;

    double tph = now_ms();
    auto phase = [&](int k) { const double t = now_ms(); g_trace_phase[k] += t - tph; tph = t; };
    // ---- references used by this range: trim, degap, validate, classes, score table ---------
    RefSet rs;
    {
        const int rc = build_refs(pl, ref_bytes, ref_off, n_refs, ref_idx, pair_begin, pair_end, degap, rs);
        if (rc) return rc;
    }
    const std::vector<int32_t>& ref_local = rs.ref_local;
    const std::vector<int64_t>& ref_pos = rs.ref_pos;
    const std::vector<int32_t>& ref_len = rs.ref_len;
    const size_t nu = rs.used_refs.size();
    const size_t ref_total = rs.ref_total;
    const bool any_dollar3 = rs.any_dollar3;
    uint8_t* h_ref_raw = ws->h_ref_raw.p;
    uint8_t* h_ref_cls = ws->h_ref_cls.p;
    int32_t* h_table4 = ws->h_table4.p;
    bool qry_present[128] = {false};

    phase(0);
    // ---- queries, pass 1 (parallel): trim span, validate, length after degap, byte presence ----
    std::vector<HostPair> hp((size_t)n);
    const int64_t qbytes_in = qry_off[pair_end] - qry_off[pair_begin];
    const int nthreads = host_threads(qbytes_in);
    struct ThreadErr { int code = 0; int64_t pair = -1; int byte = 0; uint64_t mask[2] = {0, 0}; int64_t cells = 0; char pad[64]; };   // mask: bytes 0-63, 64-127 seen
    std::vector<ThreadErr> terr((size_t)nthreads);
    parallel_for(n, nthreads, [&](int64_t lo_k, int64_t hi_k, int tid) {
        ThreadErr& te = terr[(size_t)tid];
        uint8_t unk[256];
        memset(unk, 1, sizeof(unk));
        for (int64_t kk = lo_k; kk < hi_k; ++kk) {
            const int64_t k = pair_begin + kk;
            HostPair& h = hp[(size_t)kk];
            if (qry_off[k + 1] < qry_off[k]) { if (!te.code) { te.code = GOTOH_B200_EINVAL; te.pair = k; } return; }
            const uint8_t* s = qry_bytes + qry_off[k];
            int64_t lo, hi;
            trim_span(s, qry_off[k + 1] - qry_off[k], &lo, &hi);
            if (hi - lo >= (1 << 24)) { if (!te.code) { te.code = GOTOH_B200_ERANGE; te.pair = k; } return; }
            // Which bytes occur (the score range of the int16 proof depends on it) and are they all in 1..126?  Each
            // thread keeps the set it has already seen as a 256-entry table `unk` (1 = not seen yet / invalid): a query
            // made of known bytes costs one table load and one OR per byte; only a query that brings a new byte (the
            // first few of a batch) or an invalid one takes the exact loop below.
            unsigned fresh = 0;
            {
                unsigned f0 = 0, f1 = 0, f2 = 0, f3 = 0;          // independent chains: the loop is load-bound, not OR-bound
                int64_t x = lo;
                for (; x + 8 <= hi; x += 8) {
                    f0 |= unk[s[x]] | unk[s[x + 4]]; f1 |= unk[s[x + 1]] | unk[s[x + 5]];
                    f2 |= unk[s[x + 2]] | unk[s[x + 6]]; f3 |= unk[s[x + 3]] | unk[s[x + 7]];
                }
                for (; x < hi; ++x) f0 |= unk[s[x]];
                fresh = f0 | f1 | f2 | f3;
            }
            unsigned bad = 0;
            if (fresh) {
                for (int64_t x = lo; x < hi; ++x) {
                    const unsigned c = s[x];
                    if ((uint8_t)(c - 1) > 125) { bad = 1; continue; }
                    unk[c] = 0;
                    te.mask[c >> 6] |= 1ull << (c & 63);
                }
            }
            int dashes = 0;
            if (degap) for (int64_t x = lo; x < hi; ++x) dashes += (s[x] == '-');
            const int nn = (int)(hi - lo) - (degap ? dashes : 0);
            if (bad) {
                if (!te.code) {
                    te.code = GOTOH_B200_EDOMAIN; te.pair = k;
                    for (int64_t x = lo; x < hi; ++x) if ((uint8_t)(s[x] - 1) > 125) { te.byte = s[x]; break; }
                }
                return;
            }
            if (nn == 0) { if (!te.code) { te.code = GOTOH_B200_EEMPTY; te.pair = k; } return; }
            h.ref = ref_idx ? ref_local[(size_t)ref_idx[k]] : (int32_t)kk;
            h.M = ref_len[(size_t)h.ref];
            h.N = nn;
            h.orig = (int32_t)kk;
            h.src_lo = qry_off[k] + lo;
            h.qpos = hi - lo;     // span length for pass 2; replaced by the packed position below
            // the -100000 sentinel (gotoh.cpp:284-286): outside this bound the reference may read
            // uninitialised end indices
            const long long worst = 2LL * pl->gip + ((long long)std::max(h.M, h.N) + 1) * pl->gep;
            if (worst >= 100000) { if (!te.code) { te.code = GOTOH_B200_ESENTINEL; te.pair = k; } return; }
            if (out_off && (out_off[k + 1] - out_off[k] < (int64_t)h.M + h.N || out_off[k + 1] - out_off[k] > 0x7fffffffLL)) {
                if (!te.code) { te.code = GOTOH_B200_ERANGE; te.pair = k; te.byte = -1; }
                return;
            }
            te.cells += (int64_t)h.M * h.N;
        }
    });
    for (const ThreadErr& te : terr) {
        if (te.code == GOTOH_B200_EDOMAIN) return fail(te.code, "query %lld: byte 0x%02x outside 1..126", (long long)te.pair, te.byte);
        if (te.code == GOTOH_B200_EEMPTY) return fail(te.code, "query %lld is empty after trim", (long long)te.pair);
        if (te.code == GOTOH_B200_ESENTINEL) return fail(te.code, "pair %lld: 2*gip+(max(M,N)+1)*gep >= 100000 (sentinel domain, gotoh.cpp:284)", (long long)te.pair);
        if (te.code == GOTOH_B200_ERANGE) return fail(te.code, te.byte == -1 ? "pair %lld: output stride smaller than M+N (or >= 2^31)" : "query %lld too long", (long long)te.pair);
        if (te.code) return fail(te.code, "qry_off not monotone at %lld", (long long)te.pair);
        for (int c = 0; c < 128; ++c) qry_present[c] |= (bool)((te.mask[c >> 6] >> (c & 63)) & 1);
        pl->cells += te.cells;
    }
    phase(1);
    // ---- pass 2 (parallel): packed positions + copy ------------------------------------------
    int64_t qtotal = 0;
    std::vector<int64_t> span((size_t)n);
    for (int64_t kk = 0; kk < n; ++kk) { span[(size_t)kk] = hp[(size_t)kk].qpos; hp[(size_t)kk].qpos = qtotal; qtotal += hp[(size_t)kk].N; }
    CU(ws->h_qry.ensure((size_t)qtotal + 64));
    uint8_t* h_qry = ws->h_qry.p;
    memset(h_qry + qtotal, 0, 64);
    parallel_for(n, nthreads, [&](int64_t lo_k, int64_t hi_k, int) {
        for (int64_t kk = lo_k; kk < hi_k; ++kk) {
            const HostPair& h = hp[(size_t)kk];
            const uint8_t* s = qry_bytes + h.src_lo;
            uint8_t* d = h_qry + h.qpos;
            if (!degap) memcpy(d, s, (size_t)h.N);
            else { int m = 0; for (int64_t x = 0; x < span[(size_t)kk]; ++x) if (s[x] != '-') d[m++] = s[x]; }
        }
    });

    phase(2);
    // ---- score range over the bytes present in the queries (the int16 proof depends on it) --------
    int minT = 0, maxT = 0;
    for (int c = 1; c < pl->ncls; ++c)
        for (int b = 1; b < 127; ++b)
            if (qry_present[b]) { const int t = tab.v[rs.rep[(size_t)c]][b]; minT = std::min(minT, t); maxT = std::max(maxT, t); }
    if (any_dollar3) maxT += 18;                   // up to three +6 bonuses on one cell

    phase(3);
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
    const int force = getenv("GOTOH_B200_FORCE_PATH") ? atoi(getenv("GOTOH_B200_FORCE_PATH")) : 0;  // tests: 32
    t_half_off = getenv("GOTOH_B200_HALF") && atoi(getenv("GOTOH_B200_HALF")) == 0;                  // tests: 32-lane wavefronts only
    std::vector<KeyIdx> elig, wide;
    elig.reserve((size_t)n);
    int worstM = 0, worstN = 0;   // for picking R: fits_int16 is monotone in min(M,N) and K
    // the plan's frame shift: the largest need among the pairs that would fit with their own shift
    long long z4 = 0;
    if (force != 32)
        for (const HostPair& h : hp) {
            if (h.N > 32 * kMaxK) continue;
            const long long need = int16_low_need(h.M, h.N, pl->gip, pl->gep, minT);
            if (fits_int16(h.M, h.N, pick_K(h.N), 32, pl->gip, pl->gep, minT, maxT, need)) z4 = std::max(z4, need);
        }
    pl->zshift = (int)z4;
    for (size_t x = 0; x < hp.size(); ++x) {
        const HostPair& h = hp[x];
        const int K = pick_K(h.N);
        const bool ok = h.N <= 32 * kMaxK && force != 32 && nu < (1u << 26) &&
                        fits_int16(h.M, h.N, K, 32, pl->gip, pl->gep, minT, maxT, z4);
        KeyIdx ki;
        ki.idx = (uint32_t)x;
        if (ok) {
            // group by (K, M desc, ref, N desc): partners share the reference and have similar width
            ki.key = ((uint64_t)pick_KH(h.N) << 59) | ((uint64_t)(0xffffff - h.M) << 35) | ((uint64_t)h.ref << 9) | (uint64_t)(511 - std::min(h.N, 511));
            elig.push_back(ki);
            if (std::min(h.M, h.N) > std::min(worstM, worstN)) { worstM = h.M; worstN = h.N; }
        } else {
            const bool multi = h.N > 32 * K;   // multi-strip tasks form their own launch
            const uint64_t cells = (uint64_t)h.M * (uint64_t)h.N;
            ki.key = ((uint64_t)K << 59) | ((uint64_t)(multi ? 0 : 1) << 58) | ((((uint64_t)1 << 50) - 1 - std::min<uint64_t>(cells, ((uint64_t)1 << 50) - 1)));
            wide.push_back(ki);
        }
    }
    int R = 32;
    if (!elig.empty()) {
        // fits_int16 depends on the pair through min(M,N) and K only and is monotone in min(M,N): test, per K, the
        // eligible pair with the largest min(M,N) instead of every pair for every candidate
        int worst_mn[kMaxK + 1];
        for (int k = 0; k <= kMaxK; ++k) worst_mn[k] = -1;
        for (const KeyIdx& ki : elig) {
            const HostPair& h = hp[ki.idx];
            const int K = pick_K(h.N);
            worst_mn[K] = std::max(worst_mn[K], std::min(h.M, h.N));
        }
        for (int cand = 4096; cand >= 32; cand >>= 1) {
            bool all = true;
            for (int k = 0; k <= kMaxK && all; ++k)
                if (worst_mn[k] >= 0 && !fits_int16(worst_mn[k], worst_mn[k], k, cand, pl->gip, pl->gep, minT, maxT, z4)) all = false;
            if (all) { R = cand; break; }
        }
    }
    phase(4);
    radix_sort(elig);
    std::sort(wide.begin(), wide.end());
    phase(5);

    CU(ws->h_pairs.ensure((size_t)n));
    CU(ws->h_tasks.ensure((size_t)n * 2 + 2));      // one entry per couple / single, plus at most one filler each (half-warp kernels take entries in twos)
    PairInfo* pairs = ws->h_pairs.p;
    Task* tasks = ws->h_tasks.p;
    size_t n_tasks = 0;
    struct TaskMeta { int x2, K; int64_t arena; int multi; int hw; int share_prev; int dummy; };   // hw: half-warp wavefronts; share_prev: second couple of a warp (same arena slab); dummy: filler entry, owns no pairs
    std::vector<TaskMeta> tmeta;
    tmeta.reserve((size_t)n);
    int64_t ops_words = 0;
    int next_pair = 0;
    auto add_pair = [&](uint32_t hx, int K, int x2, int half) -> int {
        const HostPair& h = hp[hx];
        PairInfo& pi = pairs[next_pair];
        memset(&pi, 0, sizeof(pi));
        pi.ref_pos = ref_pos[(size_t)h.ref];
        pi.qry_pos = h.qpos;
        // (tight / compact forms have no caller offsets: k_scan assigns them on the device)
        pi.out_off = out_off ? out_off[pair_begin + h.orig] - out_off[pair_begin] : 0;
        pi.out_cap = out_off ? (int32_t)(out_off[pair_begin + h.orig + 1] - out_off[pair_begin + h.orig]) : h.M + h.N;
        pi.M = h.M; pi.N = h.N;
        pi.K = (int16_t)K; pi.x2 = (int8_t)x2; pi.half = (int8_t)half;
        pi.orig = h.orig;
        pi.nblk = x2 ? (h.M + ((half & 4) ? 15 : 31) + 3) / 4 : (h.M + 31 + 7) / 8;   // wavefront fill: 31 rows, 15 on half-warp wavefronts
        pi.half = (int8_t)(half & 3);
        if (ops_words + (h.M + h.N + 15) / 16 > 0x7fffffffLL) return -1;
        pi.ops_off = (int32_t)ops_words;
        ops_words += (h.M + h.N + 15) / 16;
        return next_pair++;
    };
    int open_quad = -1;            // task index of a half-warp couple that still lacks its warp partner
    auto close_quad = [&]() {
        // a warp of the HALF kernel always takes two task entries: fill the gap with a copy of the lone couple (it
        // recomputes the same pairs on lanes 16-31 and writes identical results)
        if (open_quad < 0) return;
        tasks[n_tasks++] = tasks[open_quad];
        tmeta.push_back({1, tmeta[(size_t)open_quad].K, 0, 0, 1, 1, 1});
        open_quad = -1;
    };
    for (size_t x = 0; x < elig.size();) {
        const uint32_t a = elig[x].idx;
        const int KH = pick_KH(hp[a].N), K = KH & 15, hw = KH >> 4;
        int64_t b = -1;
        if (x + 1 < elig.size() && hp[elig[x + 1].idx].ref == hp[a].ref && pick_KH(hp[elig[x + 1].idx].N) == KH) b = elig[x + 1].idx;
        // second couple of a warp: same kernel, same reference (same M keeps the block loop warp-uniform)
        bool second = false;
        if (hw && open_quad >= 0) {
            const PairInfo& first = pairs[tasks[open_quad].pair_a];
            second = tmeta[(size_t)open_quad].K == K && first.ref_pos == ref_pos[(size_t)hp[a].ref] && first.M == hp[a].M;
            if (!second) close_quad();
        }
        if (!hw) close_quad();
        const int hbits = (hw ? 4 : 0) | (second ? 2 : 0);
        Task t;
        t.pair_a = add_pair(a, K, 1, hbits | 0);
        t.pair_b = b >= 0 ? add_pair((uint32_t)b, K, 1, hbits | 1) : -1;
        if (t.pair_a < 0 || (b >= 0 && t.pair_b < 0)) return fail(GOTOH_B200_ERANGE, "op-script arena exceeds 2^31 words; split the batch");
        tasks[n_tasks++] = t;
        tmeta.push_back({1, K, second ? 0 : (int64_t)pairs[t.pair_a].nblk * 32, 0, hw, second ? 1 : 0, 0});
        if (hw) open_quad = second ? -1 : (int)n_tasks - 1;
        pl->pairs_x2 += (b >= 0) ? 2 : 1;
        x += (b >= 0) ? 2 : 1;
    }
    close_quad();
    for (const KeyIdx& ki : wide) {
        const uint32_t a = ki.idx;
        const int K = pick_K(hp[a].N);
        Task t;
        t.pair_a = add_pair(a, K, 0, 0);
        t.pair_b = -1;
        if (t.pair_a < 0) return fail(GOTOH_B200_ERANGE, "op-script arena exceeds 2^31 words; split the batch");
        const int nstrips = (hp[a].N + 32 * K - 1) / (32 * K);
        tasks[n_tasks++] = t;
        tmeta.push_back({0, K, (int64_t)nstrips * pairs[t.pair_a].nblk * 32, nstrips > 1, 0, 0, 0});
        pl->pairs_x1 += 1;
    }

    if (out_off) { pl->out_base = out_off[pair_begin]; pl->out_bytes = out_off[pair_end] - out_off[pair_begin]; }
    else {
        pl->out_base = 0; pl->out_bytes = 0;
        for (const HostPair& h : hp) pl->out_bytes += (int64_t)h.M + h.N;      // tight form: worst case of the slab
    }
    pl->ops_words = ops_words;
    pl->pair_base = pair_begin;
    pl->n_pairs = n;

    phase(6);
    // ---- device buffers (grow-only, reused across plans on this workspace) ---------------------
    CU(ws->d_ref_raw.ensure(ref_total));
    CU(ws->d_ref_cls.ensure(ref_total));
    CU(ws->d_qry.ensure((size_t)qtotal + 64));
    CU(ws->d_table4.ensure((size_t)pl->ncls * 136));
    CU(ws->d_pairs.ensure((size_t)n));
    CU(ws->d_tasks.ensure(n_tasks));
    CU(ws->d_score.ensure((size_t)n)); CU(ws->d_end_i.ensure((size_t)n)); CU(ws->d_end_j.ensure((size_t)n));
    CU(ws->d_nops.ensure((size_t)n)); CU(ws->d_i0.ensure((size_t)n)); CU(ws->d_j0.ensure((size_t)n));
    CU(ws->d_len_plan.ensure((size_t)n)); CU(ws->d_out_len.ensure((size_t)n)); CU(ws->d_out_score.ensure((size_t)n));
    CU(ws->d_ops.ensure((size_t)ops_words));
    if (pl->out_mode != OUT_COMPACT) {
        CU(ws->d_out_ref.ensure((size_t)pl->out_bytes));
        CU(ws->d_out_qry.ensure((size_t)pl->out_bytes));
    }
    if (pl->out_mode != OUT_STRIDED) {
        if (pl->out_mode == OUT_TIGHT) CU(ws->d_scan_in.ensure((size_t)n));
        CU(ws->d_scan_out.ensure((size_t)n + 1));
    }
    if (pl->out_mode == OUT_COMPACT) {
        CU(ws->d_rec.ensure((size_t)n * 8));
        CU(ws->d_cops.ensure((size_t)ops_words));
    }

    // ---- arena budget and chunking ----------------------------------------------------------------
    int64_t biggest = 0, total_arena = 0;
    bool any_multi = false;
    for (const TaskMeta& m : tmeta) { biggest = std::max(biggest, m.arena); total_arena += m.arena; any_multi |= (m.multi != 0); }
    if (any_multi) {
        // multi-strip boundary columns: two int2 columns of maxM+2 rows per resident warp
        pl->bnd_stride = ((int64_t)maxM + 2 + 15) & ~15LL;
        const int64_t slots = (int64_t)ws->sm_count * GOTOH_MIN_CTAS * FWD_WARPS;
        CU(ws->d_bnd.ensure((size_t)(slots * 2 * pl->bnd_stride)));
    }
    int64_t budget = pl->arena_budget_bytes;
    if (getenv("GOTOH_B200_ARENA_MB")) budget = (int64_t)atoll(getenv("GOTOH_B200_ARENA_MB")) << 20;  // tests: force chunking
    if (budget <= 0) {
        size_t free_b = 0, total_b = 0;
        CU(cudaMemGetInfo(&free_b, &total_b));
        budget = (int64_t)((free_b + ws->d_dir.cap * sizeof(uint4)) * 0.80);
    }
    const int64_t budget_u4 = std::max<int64_t>(std::min<int64_t>(budget / 16, total_arena), biggest);

    pl->chunks.clear();
    int64_t used = 0, arena_max = 0;
    Chunk cur; cur.pair_first = 0; cur.pair_count = 0;
    auto flush = [&]() {
        if (!cur.launches.empty()) { pl->chunks.push_back(cur); arena_max = std::max(arena_max, used); }
        cur.launches.clear();
        used = 0;
    };
    int pair_cursor = 0;
    int64_t prev_off = 0;
    for (size_t t = 0; t < n_tasks; ++t) {
        const TaskMeta& m = tmeta[t];
        // (the second entry of a half-warp warp shares its partner's slab: it never opens a chunk or a launch)
        if (!m.share_prev && used + m.arena > budget_u4) { flush(); cur.pair_first = pair_cursor; cur.pair_count = 0; }
        if (!m.share_prev && (cur.launches.empty() || cur.launches.back().x2 != m.x2 || cur.launches.back().K != m.K ||
                              cur.launches.back().multi_strip != m.multi || cur.launches.back().hw != m.hw)) {
            Launch L; L.x2 = m.x2; L.K = m.K; L.task_first = (int)t; L.task_count = 0;
            L.rebase_mask = R - 1; L.multi_strip = m.multi; L.hw = m.hw;
            cur.launches.push_back(L);
        }
        cur.launches.back().task_count++;
        const Task& tk = tasks[t];
        if (m.multi) {
            // strip dataflow: this pair's strips take the next slots of its launch
            const PairInfo& pp = pairs[tk.pair_a];
            pairs[tk.pair_a].pad1 = cur.launches.back().nslots;
            cur.launches.back().nslots += (pp.N + 32 * pp.K - 1) / (32 * pp.K);
        }
        if (!m.dummy) {
            const int64_t at = m.share_prev ? prev_off : used;
            pairs[tk.pair_a].dir_off = at;
            if (tk.pair_b >= 0) pairs[tk.pair_b].dir_off = at;
            prev_off = at;
        }
        used += m.arena;
        const int np = m.dummy ? 0 : (tk.pair_b >= 0 ? 2 : 1);
        cur.pair_count += np;
        pair_cursor += np;
    }
    flush();
    // strip tasks of the multi-strip launches: (pair, strip) in pair-major, strip-ascending order
    {
        size_t total = 0;
        int max_slots = 0;
        for (Chunk& c : pl->chunks)
            for (Launch& L : c.launches)
                if (L.multi_strip) { L.stask_first = (int)total; total += (size_t)L.nslots; max_slots = std::max(max_slots, L.nslots); }
        if (total) {
            CU(ws->h_stasks.ensure(total));
            CU(ws->d_stasks.ensure(total));
            CU(ws->d_flow.ensure((size_t)max_slots * 3 + 3));
            size_t at = 0;
            for (const Chunk& c : pl->chunks)
                for (const Launch& L : c.launches) {
                    if (!L.multi_strip) continue;
                    // strip-major: strip 0 of every pair, then strip 1, ... - a warp that claims (pair, s) then finds
                    // (pair, s-1) far ahead (or finished) instead of spinning 64 rows behind a producer that has just
                    // started; a producer still precedes its consumer in claim order, which is what rules out deadlock
                    int max_ns = 0;
                    for (int t = L.task_first; t < L.task_first + L.task_count; ++t) {
                        const PairInfo& pp = pairs[tasks[t].pair_a];
                        max_ns = std::max(max_ns, (pp.N + 32 * pp.K - 1) / (32 * pp.K));
                    }
                    for (int sidx = 0; sidx < max_ns; ++sidx)
                        for (int t = L.task_first; t < L.task_first + L.task_count; ++t) {
                            const PairInfo& pp = pairs[tasks[t].pair_a];
                            if (sidx >= (pp.N + 32 * pp.K - 1) / (32 * pp.K)) continue;
                            ws->h_stasks.p[at].pair_a = tasks[t].pair_a; ws->h_stasks.p[at].pair_b = sidx; ++at;
                        }
                }
            pl->flow_slots = max_slots;
            // one boundary column per slot (d_bnd is also what the warp-serial / CTA kernels use)
            const char* pin = getenv("GOTOH_B200_LONG");
            if (!pin || pin[0] == 'f') {
                cudaError_t e = ws->d_bnd.ensure((size_t)max_slots * (size_t)pl->bnd_stride + 16);
                if (e != cudaSuccess) { (void)cudaGetLastError(); pl->flow_slots = 0; }     // fall back to the warp-serial kernel
            }
            pl->n_stasks = (int64_t)total;
        }
    }
    pl->n_launches = 0;
    for (const Chunk& c : pl->chunks) pl->n_launches += (int)c.launches.size() + 2;
    {
        cudaError_t e = ws->d_dir.ensure((size_t)arena_max);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            return fail(GOTOH_B200_ENOMEM, "direction arena of %lld bytes does not fit in device memory", (long long)arena_max * 16);
        }
    }
    pl->arena_bytes = arena_max * 16;
    CU(ws->d_counter.ensure((size_t)std::max(pl->n_launches, 1)));

    // ---- H2D (async, from pinned staging) ------------------------------------------------------------
    auto h2d = [&](void* d, const void* h, size_t bytes) -> cudaError_t {
        pl->h2d_bytes += (int64_t)bytes;
        return cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, ws->stream);
    };
    CU(h2d(ws->d_ref_raw.p, h_ref_raw, ref_total));
    CU(h2d(ws->d_ref_cls.p, h_ref_cls, ref_total));
    CU(h2d(ws->d_qry.p, h_qry, (size_t)qtotal + 64));
    CU(h2d(ws->d_table4.p, h_table4, (size_t)pl->ncls * 136 * sizeof(int32_t)));
    CU(h2d(ws->d_pairs.p, pairs, (size_t)n * sizeof(PairInfo)));
    CU(h2d(ws->d_tasks.p, tasks, n_tasks * sizeof(Task)));
    if (pl->n_stasks) CU(h2d(ws->d_stasks.p, ws->h_stasks.p, (size_t)pl->n_stasks * sizeof(Task)));
    phase(7);
    if (trace_on() && !pl->owns_ws)
        ;   // the one-shot call prints the phases per slab
    else if (trace_on())
        fprintf(stderr, "[gotoh_b200] plan_build %lld pairs: refs %.2f pass1 %.2f pass2 %.2f cls %.2f path %.2f sort %.2f tasks %.2f alloc+h2d %.2f ms\n",
                (long long)n, g_trace_phase[0], g_trace_phase[1], g_trace_phase[2], g_trace_phase[3], g_trace_phase[4], g_trace_phase[5], g_trace_phase[6], g_trace_phase[7]);
    return GOTOH_B200_OK;
}

// Device-side plan builder (gotoh_prep.cuh) for the common batch shape: many short queries against at most PREP_MAX_REFS
// shared references, every pair admitted to the int16x2 kernels.  The host touches the references only; the queries go
// to the device as they are and five small kernels trim, validate, admit, group and lay them out.  *done = false (and
// nothing else changed that matters) means "not this path": the caller then runs the host builder, which also reports
// every input error.  One event synchronisation per plan (the summary comes back through mapped pinned memory).
int plan_build_device(gotoh_b200_plan* pl, const uint8_t* ref_bytes, const int64_t* ref_off, int64_t n_refs,
                      const int32_t* ref_idx, const uint8_t* qry_bytes, const int64_t* qry_off,
                      int64_t pair_begin, int64_t pair_end, const int64_t* out_off, bool* done) {
    *done = false;
    pl->built_on_device = false;
    Workspace* ws = pl->ws;
    const int64_t n = pair_end - pair_begin;
    const char* sw = getenv("GOTOH_B200_DEVICE_PREP");            // 0: host builder only; 1: also for small batches (tests)
    const int mode = sw ? atoi(sw) : -1;
    if (mode == 0) return GOTOH_B200_OK;
    if (!ref_idx || pl->matrix == GOTOH_B200_AA_RB || n_refs > 4096 || n > 0x3fffffff) return GOTOH_B200_OK;
    if (n < (mode == 1 ? 1 : 4096)) return GOTOH_B200_OK;        // small batches: the host builder costs less than a synchronisation
    if (pl->gip < 0 || pl->gep < 0 || pl->gip > 100000 || pl->gep > 100000) return GOTOH_B200_OK;
    if (getenv("GOTOH_B200_FORCE_PATH") && atoi(getenv("GOTOH_B200_FORCE_PATH")) == 32) return GOTOH_B200_OK;
    const int64_t qbytes = qry_off[pair_end] - qry_off[pair_begin];
    if (qbytes <= 0 || qbytes > ((int64_t)1 << 31)) return GOTOH_B200_OK;
    CU(cudaSetDevice(ws->device));
    double tph = now_ms();
    auto phase = [&](int k) { const double t = now_ms(); g_trace_phase[k] += t - tph; tph = t; };
    RefSet rs;
    if (build_refs(pl, ref_bytes, ref_off, n_refs, ref_idx, pair_begin, pair_end, false, rs)) return GOTOH_B200_OK;
    const int nu = (int)rs.used_refs.size();
    if (nu > PREP_MAX_REFS) return GOTOH_B200_OK;
    phase(0);
    // references ranked in the order of the host builder's sort key: M descending, then index
    std::vector<int> order((size_t)nu);
    for (int u = 0; u < nu; ++u) order[(size_t)u] = u;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return rs.ref_len[(size_t)a] != rs.ref_len[(size_t)b] ? rs.ref_len[(size_t)a] > rs.ref_len[(size_t)b] : a < b; });
    // ---- carve the scratch block ----------------------------------------------------------------------
    const size_t nbins = (size_t)PREP_NKH * nu * PREP_NSLOT, ngrp = (size_t)PREP_NKH * nu;
    size_t at = 0;
    auto take = [&](size_t bytes) { const size_t o = (at + 15) & ~(size_t)15; at = o + bytes; return o; };
    const size_t o_zero0 = take(0);
    const size_t o_present = take(16), o_scal = take(32 * 4), o_acc = take(16), o_hist = take(nbins * 4), o_cursor = take(nbins * 4);
    const size_t o_zero1 = take(0);
    const size_t o_bin_pair = take((nbins + 1) * 8), o_bin_ops = take((nbins + 1) * 8), o_grp = take(ngrp * 3 * 8);
    const size_t o_N = take((size_t)n * 4), o_lo = take((size_t)n * 4);
    const size_t o_qoff = take((size_t)(n + 1) * 8), o_ridx = take((size_t)n * 4), o_ooff = take(out_off ? (size_t)(n + 1) * 8 : 0);
    const size_t o_refM = take((size_t)n_refs * 4), o_refrank = take((size_t)n_refs * 4), o_refpos = take((size_t)n_refs * 8), o_rankM = take((size_t)nu * 4);
    const size_t reftab_bytes = at - o_refM;
    CU(ws->d_prep.ensure(at));
    CU(ws->h_prep.ensure(reftab_bytes));
    uint8_t* d = ws->d_prep.p;
    {
        int32_t* h_refM = reinterpret_cast<int32_t*>(ws->h_prep.p + (o_refM - o_refM));
        int32_t* h_refrank = reinterpret_cast<int32_t*>(ws->h_prep.p + (o_refrank - o_refM));
        int64_t* h_refpos = reinterpret_cast<int64_t*>(ws->h_prep.p + (o_refpos - o_refM));
        int32_t* h_rankM = reinterpret_cast<int32_t*>(ws->h_prep.p + (o_rankM - o_refM));
        memset(ws->h_prep.p, 0, reftab_bytes);
        for (int64_t r = 0; r < n_refs; ++r) h_refrank[r] = -1;
        for (int rank = 0; rank < nu; ++rank) {
            const int u = order[(size_t)rank];
            const int64_t r = rs.used_refs[(size_t)u];
            h_refM[r] = rs.ref_len[(size_t)u]; h_refrank[r] = rank; h_refpos[r] = rs.ref_pos[(size_t)u];
            h_rankM[rank] = rs.ref_len[(size_t)u];
        }
    }
    // ---- H2D: references (as the host builder), the raw queries and the caller's offset arrays ------------------
    CU(ws->d_ref_raw.ensure(rs.ref_total));
    CU(ws->d_ref_cls.ensure(rs.ref_total));
    CU(ws->d_table4.ensure((size_t)pl->ncls * 136));
    CU(ws->d_qry.ensure((size_t)qbytes + 64));
    pl->h2d_bytes = 0;
    auto h2d = [&](void* dd, const void* h, size_t bytes) -> cudaError_t {
        pl->h2d_bytes += (int64_t)bytes;
        return cudaMemcpyAsync(dd, h, bytes, cudaMemcpyHostToDevice, ws->stream);
    };
    CU(h2d(ws->d_ref_raw.p, ws->h_ref_raw.p, rs.ref_total));
    CU(h2d(ws->d_ref_cls.p, ws->h_ref_cls.p, rs.ref_total));
    CU(h2d(ws->d_table4.p, ws->h_table4.p, (size_t)pl->ncls * 136 * sizeof(int32_t)));
    CU(h2d(ws->d_qry.p, qry_bytes + qry_off[pair_begin], (size_t)qbytes));
    CU(cudaMemsetAsync(ws->d_qry.p + qbytes, 0, 64, ws->stream));
    CU(h2d(d + o_qoff, qry_off + pair_begin, (size_t)(n + 1) * 8));
    CU(h2d(d + o_ridx, ref_idx + pair_begin, (size_t)n * 4));
    if (out_off) CU(h2d(d + o_ooff, out_off + pair_begin, (size_t)(n + 1) * 8));
    CU(h2d(d + o_refM, ws->h_prep.p, reftab_bytes));
    CU(cudaMemsetAsync(d + o_zero0, 0, o_zero1 - o_zero0, ws->stream));
    phase(1);
    PrepParams pp;
    memset(&pp, 0, sizeof(pp));
    pp.qry = ws->d_qry.p; pp.qry_off = reinterpret_cast<const int64_t*>(d + o_qoff); pp.ref_idx = reinterpret_cast<const int32_t*>(d + o_ridx);
    pp.out_off = out_off ? reinterpret_cast<const int64_t*>(d + o_ooff) : nullptr;
    pp.n = (int32_t)n; pp.n_refs = (int32_t)n_refs;
    pp.ref_M = reinterpret_cast<const int32_t*>(d + o_refM); pp.ref_pos = reinterpret_cast<const int64_t*>(d + o_refpos);
    pp.ref_rank = reinterpret_cast<const int32_t*>(d + o_refrank); pp.n_used = nu; pp.rank_M = reinterpret_cast<const int32_t*>(d + o_rankM);
    pp.table4 = ws->d_table4.p; pp.ncls = pl->ncls; pp.gip = pl->gip; pp.gep = pl->gep; pp.has_dollar = pl->has_dollar;
    pp.half_off = (getenv("GOTOH_B200_HALF") && atoi(getenv("GOTOH_B200_HALF")) == 0) ? 1 : 0;
    pp.pairN = reinterpret_cast<int32_t*>(d + o_N); pp.pairLo = reinterpret_cast<int32_t*>(d + o_lo);
    pp.present = reinterpret_cast<uint32_t*>(d + o_present); pp.scal = reinterpret_cast<int32_t*>(d + o_scal);
    pp.acc = reinterpret_cast<unsigned long long*>(d + o_acc);
    pp.hist = reinterpret_cast<uint32_t*>(d + o_hist); pp.cursor = reinterpret_cast<uint32_t*>(d + o_cursor);
    pp.bin_pair = reinterpret_cast<int64_t*>(d + o_bin_pair); pp.bin_ops = reinterpret_cast<int64_t*>(d + o_bin_ops);
    pp.grp = reinterpret_cast<int64_t*>(d + o_grp);
    pp.summary = ws->h_prep_sum.p;
    const int nblocks = (int)((n + 255) / 256);
    GOTOH_LAUNCH(k_prep_scan, dim3(nblocks), dim3(256), 0, ws->stream, pp);
    GOTOH_LAUNCH(k_prep_range, dim3(1), dim3(256), 0, ws->stream, pp);
    GOTOH_LAUNCH(k_prep_bins, dim3(nblocks), dim3(256), 0, ws->stream, pp);
    GOTOH_LAUNCH(k_prep_layout, dim3(1), dim3(256), 0, ws->stream, pp);
    CU(cudaGetLastError());
    CU(cudaEventRecord(ws->ev_p, ws->stream));
    CU(cudaEventSynchronize(ws->ev_p));
    phase(2);
    const PrepSummary sum = *ws->h_prep_sum.p;
    if (sum.fallback) {
        if (trace_on()) fprintf(stderr, "[gotoh_b200] device plan builder declined %lld pairs (reason %d): host builder\n", (long long)n, sum.fallback);
        return GOTOH_B200_OK;
    }
    // ---- arena budget: one chunk or the host builder (which cuts chunks) ---------------------------------------------
    int64_t budget = pl->arena_budget_bytes;
    if (getenv("GOTOH_B200_ARENA_MB")) budget = (int64_t)atoll(getenv("GOTOH_B200_ARENA_MB")) << 20;
    if (budget <= 0) {
        size_t free_b = 0, total_b = 0;
        CU(cudaMemGetInfo(&free_b, &total_b));
        budget = (int64_t)((free_b + ws->d_dir.cap * sizeof(uint4)) * 0.80);
    }
    if (sum.arena_u4 * 16 > budget) return GOTOH_B200_OK;
    // ---- device buffers of the plan -----------------------------------------------------------------------------------------
    if (out_off) { pl->out_base = out_off[pair_begin]; pl->out_bytes = out_off[pair_end] - out_off[pair_begin]; }
    else { pl->out_base = 0; pl->out_bytes = sum.sum_mn; }
    CU(ws->d_pairs.ensure((size_t)n));
    CU(ws->d_tasks.ensure((size_t)sum.n_tasks));
    CU(ws->d_score.ensure((size_t)n)); CU(ws->d_end_i.ensure((size_t)n)); CU(ws->d_end_j.ensure((size_t)n));
    CU(ws->d_nops.ensure((size_t)n)); CU(ws->d_i0.ensure((size_t)n)); CU(ws->d_j0.ensure((size_t)n));
    CU(ws->d_len_plan.ensure((size_t)n)); CU(ws->d_out_len.ensure((size_t)n)); CU(ws->d_out_score.ensure((size_t)n));
    CU(ws->d_ops.ensure((size_t)sum.ops_words));
    if (pl->out_mode != OUT_COMPACT) {
        CU(ws->d_out_ref.ensure((size_t)pl->out_bytes));
        CU(ws->d_out_qry.ensure((size_t)pl->out_bytes));
    }
    if (pl->out_mode != OUT_STRIDED) {
        if (pl->out_mode == OUT_TIGHT) CU(ws->d_scan_in.ensure((size_t)n));
        CU(ws->d_scan_out.ensure((size_t)n + 1));
    }
    if (pl->out_mode == OUT_COMPACT) {
        CU(ws->d_rec.ensure((size_t)n * 8));
        CU(ws->d_cops.ensure((size_t)sum.ops_words));
    }
    {
        cudaError_t e = ws->d_dir.ensure((size_t)sum.arena_u4);
        if (e != cudaSuccess) { (void)cudaGetLastError(); return GOTOH_B200_OK; }       // the host builder cuts chunks
    }
    pp.pairs = ws->d_pairs.p; pp.tasks = ws->d_tasks.p;
    GOTOH_LAUNCH(k_prep_scatter, dim3(nblocks), dim3(256), 0, ws->stream, pp);
    CU(cudaGetLastError());
    // ---- the plan ---------------------------------------------------------------------------------------------------------------
    pl->chunks.clear();
    Chunk c;
    c.pair_first = 0; c.pair_count = (int)n;
    for (int l = 0; l < sum.n_launch; ++l) {
        Launch L;
        L.x2 = 1; L.K = sum.launch[l].K; L.hw = sum.launch[l].hw;
        L.task_first = sum.launch[l].task_first; L.task_count = sum.launch[l].task_count;
        L.rebase_mask = sum.R - 1; L.multi_strip = 0;
        c.launches.push_back(L);
    }
    pl->chunks.push_back(c);
    pl->n_launches = (int)c.launches.size() + 2;
    CU(ws->d_counter.ensure((size_t)pl->n_launches));
    pl->zshift = sum.z4;
    pl->smin_m1 = (int)((long long)std::min(sum.minT, 0) * 32 * PLAN_MAX_K - 2LL * pl->gip - pl->gep - 2);
    pl->cells = sum.cells; pl->pairs_x2 = n; pl->pairs_x1 = 0;
    pl->pair_base = pair_begin; pl->n_pairs = n; pl->ops_words = sum.ops_words;
    pl->arena_bytes = sum.arena_u4 * 16;
    pl->bnd_stride = 0; pl->flow_slots = 0; pl->n_stasks = 0;
    phase(7);
    pl->built_on_device = true;
    *done = true;
    return GOTOH_B200_OK;
}

int launch_emit(const gotoh_b200_plan* pl, int pair_first, int pair_count, const int64_t* tight_off) {
    const Workspace* ws = pl->ws;
    EmitParams ep;
    memset(&ep, 0, sizeof(ep));
    ep.pairs = ws->d_pairs.p; ep.pair_first = pair_first; ep.pair_count = pair_count;
    ep.ref_raw = ws->d_ref_raw.p; ep.qry = ws->d_qry.p; ep.ops = ws->d_ops.p; ep.nops = ws->d_nops.p;
    ep.i0 = ws->d_i0.p; ep.j0 = ws->d_j0.p; ep.end_i = ws->d_end_i.p; ep.end_j = ws->d_end_j.p;
    ep.out_len_plan = ws->d_len_plan.p; ep.score_plan = ws->d_score.p;
    ep.out_ref = ws->d_out_ref.p; ep.out_qry = ws->d_out_qry.p;
    ep.out_len = ws->d_out_len.p; ep.out_score = ws->d_out_score.p;
    ep.tight_off = tight_off;
    GOTOH_LAUNCH(k_emit, dim3((pair_count + 3) / 4), dim3(128), 0, ws->stream, ep);
    CU(cudaGetLastError());
    return GOTOH_B200_OK;
}

// Enqueue forward DP + traceback + emit for every chunk.  With `timed`, CUDA events bracket the
// whole run (and each chunk's forward launches) and the call synchronises; otherwise it only enqueues.
int plan_run(gotoh_b200_plan* pl, bool timed, float* device_ms, float* forward_ms) {
    Workspace* ws = pl->ws;
    CU(cudaSetDevice(ws->device));
    CU(cudaMemsetAsync(ws->d_counter.p, 0, (size_t)std::max(pl->n_launches, 1) * sizeof(uint32_t), ws->stream));
    if (timed) CU(cudaEventRecord(ws->ev[0], ws->stream));
    float fwd_total = 0.f;
    int launch_no = 0;
    for (size_t ci = 0; ci < pl->chunks.size(); ++ci) {
        const Chunk& c = pl->chunks[ci];
        if (timed && forward_ms) CU(cudaEventRecord(ws->ev[2], ws->stream));
        for (const Launch& L : c.launches) {
            FwdParams fp;
            memset(&fp, 0, sizeof(fp));
            fp.pairs = ws->d_pairs.p; fp.tasks = ws->d_tasks.p;
            fp.task_first = L.task_first; fp.task_count = L.task_count;
            fp.ref_cls = ws->d_ref_cls.p; fp.qry = ws->d_qry.p; fp.table4 = ws->d_table4.p;
            fp.ncls = pl->ncls; fp.gip = pl->gip; fp.gep = pl->gep;
            fp.has_dollar = pl->has_dollar;
            fp.bonus4 = ws->d_table4.p + (size_t)pl->ncls * 128;
            fp.rebase_mask = L.rebase_mask; fp.smin_m1 = pl->smin_m1; fp.zshift = pl->zshift;
            fp.four = 4u;
            fp.dir = ws->d_dir.p;
            fp.bnd = L.multi_strip ? ws->d_bnd.p : nullptr; fp.bnd_stride = pl->bnd_stride;
            fp.score = ws->d_score.p; fp.end_i = ws->d_end_i.p; fp.end_j = ws->d_end_j.p;
            fp.work_counter = ws->d_counter.p + launch_no++;
            fp.task_limit = (!timed && !L.multi_strip) ? pl->task_limit : 0;
            // multi-strip tasks only exist with K = 8 (pick_K), and only on the int32 path
            int rc;
            if (L.x2 && L.hw) rc = launch_forward_half(pl, fp, L.K, L.task_count / 2);
            else if (L.x2) rc = launch_forward<Vec16, false>(pl, fp, L.K, L.task_count);
            else if (!L.multi_strip) rc = launch_forward<Vec32, false>(pl, fp, L.K, L.task_count);
            else {
                // K2: with few long pairs a CTA per pair (4 warps pipelined over adjacent strips) keeps
                // the SMs busy and cuts single-pair latency ~3x; with many pairs one warp per pair needs
                // no synchronisation and wins (measured on B200, 9.6 kb pairs: 2000 pairs 857 vs 1225
                // GCUPS; below ~2 pairs per resident CTA the CTA form is ahead).  GOTOH_B200_LONG=cta|warp
                // pins the choice (tests, benchmarks).
                const char* pin = getenv("GOTOH_B200_LONG");
                // (with very many reference classes the 4-warp CTA's profiles exceed shared memory: warp-serial kernel then)
                const bool flow = pl->flow_slots > 0 && (!pin || pin[0] == 'f') &&
                                  FwdSmem<Vec32, 8>::per_warp(pl->ncls) * FWD_WARPS <= 200 * 1024;
                if (flow) {
                    // K2 default: strip dataflow - every (pair, strip) is a warp task, any number of pairs fills the GPU
                    // boundary columns and last-row partials are self-validating (see Wave::nextb): preset them to "empty"
                    CU(cudaMemsetAsync(ws->d_bnd.p, 0xff, (size_t)L.nslots * (size_t)pl->bnd_stride * sizeof(int2), ws->stream));
                    CU(cudaMemsetAsync(ws->d_flow.p + pl->flow_slots, 0x80, (size_t)L.nslots * sizeof(int32_t), ws->stream));
                    fp.tasks = ws->d_stasks.p; fp.task_first = L.stask_first; fp.task_count = L.nslots;
                    fp.prog = ws->d_flow.p; fp.part_best = ws->d_flow.p + pl->flow_slots; fp.part_j = ws->d_flow.p + 2 * (size_t)pl->flow_slots;
                    rc = launch_forward_flow(pl, fp, L.nslots);
                } else {
                    bool cta = L.task_count < ws->sm_count * GOTOH_MIN_CTAS * 2;
                    if (pin) cta = (pin[0] == 'c');
                    rc = cta ? launch_forward_cta(pl, fp, L.task_count) : -1;
                    if (rc == -1) rc = launch_forward_k<Vec32, 8, true>(pl, fp, L.task_count);
                }
            }
            if (rc) return rc;
        }
        if (timed && forward_ms) CU(cudaEventRecord(ws->ev[3], ws->stream));
        if (pl->mark_forward_done && ci + 1 == pl->chunks.size()) CU(cudaEventRecord(ws->ev_fwd, ws->stream));
        WalkParams wp;
        memset(&wp, 0, sizeof(wp));
        wp.pairs = ws->d_pairs.p; wp.pair_first = c.pair_first; wp.pair_count = c.pair_count;
        wp.dir = reinterpret_cast<const uint32_t*>(ws->d_dir.p);
        wp.end_i = ws->d_end_i.p; wp.end_j = ws->d_end_j.p; wp.score = ws->d_score.p;
        wp.ops = ws->d_ops.p; wp.nops = ws->d_nops.p; wp.i0 = ws->d_i0.p; wp.j0 = ws->d_j0.p;
        wp.out_len = ws->d_len_plan.p; wp.gip = pl->gip; wp.gep = pl->gep; wp.term = pl->term;
        wp.scan_in = pl->out_mode == OUT_TIGHT ? ws->d_scan_in.p : nullptr;
        wp.scan_words = 0;
        GOTOH_LAUNCH(k_walk, dim3((c.pair_count + 127) / 128), dim3(128), 0, ws->stream, wp);
        CU(cudaGetLastError());
        // strided form: this chunk's strings at the caller's offsets.  The tight / compact forms need every pair's
        // size first (k_scan below), so their emit / pack kernel runs once, after the last chunk.
        if (pl->out_mode == OUT_STRIDED) {
            const int rc = launch_emit(pl, c.pair_first, c.pair_count, nullptr);
            if (rc) return rc;
        }
        if (timed && forward_ms) {
            CU(cudaEventSynchronize(ws->ev[3]));
            float ms = 0.f;
            CU(cudaEventElapsedTime(&ms, ws->ev[2], ws->ev[3]));
            fwd_total += ms;
        }
    }
    if (pl->out_mode != OUT_STRIDED && pl->n_pairs > 0) {
        const int n = (int)pl->n_pairs;
        if (pl->out_mode == OUT_TIGHT) {
            GOTOH_LAUNCH(k_scan, dim3(1), dim3(SCAN_THREADS), 0, ws->stream, ws->d_scan_in.p, ws->d_scan_out.p, n, ws->h_total.p);
            CU(cudaGetLastError());
            const int rc = launch_emit(pl, 0, n, ws->d_scan_out.p);
            if (rc) return rc;
        } else {
            CU(cudaMemsetAsync(ws->d_scan_out.p + n, 0, sizeof(int64_t), ws->stream));
            PackParams pp;
            memset(&pp, 0, sizeof(pp));
            pp.pairs = ws->d_pairs.p; pp.pair_count = n; pp.ops = ws->d_ops.p; pp.nops = ws->d_nops.p;
            pp.i0 = ws->d_i0.p; pp.j0 = ws->d_j0.p; pp.end_i = ws->d_end_i.p; pp.end_j = ws->d_end_j.p;
            pp.out_len_plan = ws->d_len_plan.p; pp.score_plan = ws->d_score.p;
            pp.off = ws->d_scan_out.p; pp.counter = reinterpret_cast<unsigned long long*>(ws->d_scan_out.p + n);
            pp.rec = ws->d_rec.p; pp.cops = ws->d_cops.p;
            GOTOH_LAUNCH(k_pack_ops, dim3((n + 3) / 4), dim3(128), 0, ws->stream, pp);
            CU(cudaGetLastError());
            GOTOH_LAUNCH(k_publish, dim3(1), dim3(1), 0, ws->stream, pp.counter, ws->h_total.p);
            CU(cudaGetLastError());
        }
    }
    if (timed) {
        CU(cudaEventRecord(ws->ev[1], ws->stream));
        CU(cudaEventSynchronize(ws->ev[1]));
        if (device_ms) CU(cudaEventElapsedTime(device_ms, ws->ev[0], ws->ev[1]));
        if (forward_ms) *forward_ms = fwd_total;
    }
    return GOTOH_B200_OK;
}

// Enqueue the D2H copies of the plan's results into the caller's buffers (no synchronisation).
int plan_fetch(gotoh_b200_plan* pl, uint8_t* out_ref, uint8_t* out_qry, int32_t* out_len, int32_t* out_score) {
    Workspace* ws = pl->ws;
    CU(cudaSetDevice(ws->device));
    pl->d2h_bytes = 0;
    auto d2h = [&](void* h, const void* d, size_t bytes) -> cudaError_t {
        pl->d2h_bytes += (int64_t)bytes;
        return cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, ws->stream);
    };
    CU(d2h(out_ref + pl->out_base, ws->d_out_ref.p, (size_t)pl->out_bytes));
    CU(d2h(out_qry + pl->out_base, ws->d_out_qry.p, (size_t)pl->out_bytes));
    CU(d2h(out_len + pl->pair_base, ws->d_out_len.p, (size_t)pl->n_pairs * sizeof(int32_t)));
    CU(d2h(out_score + pl->pair_base, ws->d_out_score.p, (size_t)pl->n_pairs * sizeof(int32_t)));
    return GOTOH_B200_OK;
}

// What a call writes into: the caller's result buffers of one of the three result forms (include/gotoh_b200.h).
struct OutSpec {
    int mode = OUT_STRIDED;
    uint8_t* out_ref = nullptr;          // strided, tight
    uint8_t* out_qry = nullptr;
    const int64_t* out_off_in = nullptr; // strided: caller's offsets (n+1)
    int64_t* out_off = nullptr;          // tight: byte offsets per pair; compact: op-script word offsets per pair (OUTPUT, n)
    int32_t* out_len = nullptr;          // strided, tight
    int32_t* out_score = nullptr;
    int32_t* out_rec = nullptr;          // compact: 8 words per pair
    uint32_t* out_ops = nullptr;
    int64_t cap_lo = 0, cap_hi = 0;      // tight / compact: this device's slice of the caller's capacity (bytes / words)
};

// Tight / compact forms.  Phase A (enqueued with the kernels) is only an event: the last kernel of plan_run has written
// the slab's total size into mapped pinned memory (k_scan / k_publish), no copy-engine work is queued yet.
int plan_fetch_a(gotoh_b200_plan* pl) {
    CU(cudaEventRecord(pl->ws->ev_a, pl->ws->stream));
    return GOTOH_B200_OK;
}

// Phase B (issued by the collector, in slab order, once the slab's kernels are done and its total is known): every
// result copy of the slab - per-pair arrays, slab-local offsets (the caller adds `base` after the copies have landed,
// see run_device_range) and the tightly packed result bytes at position `base` of the caller's buffer.
int plan_fetch_b(gotoh_b200_plan* pl, const OutSpec& o, int64_t base, int64_t total) {
    Workspace* ws = pl->ws;
    pl->d2h_bytes = 0;
    auto d2h = [&](void* h, const void* d, size_t bytes) -> cudaError_t {
        pl->d2h_bytes += (int64_t)bytes;
        return bytes ? cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, ws->stream) : cudaSuccess;
    };
    CU(d2h(o.out_off + pl->pair_base, ws->d_scan_out.p, (size_t)pl->n_pairs * sizeof(int64_t)));
    if (o.mode == OUT_TIGHT) {
        CU(d2h(o.out_len + pl->pair_base, ws->d_out_len.p, (size_t)pl->n_pairs * sizeof(int32_t)));
        CU(d2h(o.out_score + pl->pair_base, ws->d_out_score.p, (size_t)pl->n_pairs * sizeof(int32_t)));
        CU(d2h(o.out_ref + base, ws->d_out_ref.p, (size_t)total));
        CU(d2h(o.out_qry + base, ws->d_out_qry.p, (size_t)total));
    } else {
        CU(d2h(o.out_rec + pl->pair_base * 8, ws->d_rec.p, (size_t)pl->n_pairs * 8 * sizeof(int32_t)));
        CU(d2h(o.out_ops + base, ws->d_cops.p, (size_t)total * sizeof(uint32_t)));
    }
    return GOTOH_B200_OK;
}

int check_common(const void* ref_bytes, const int64_t* ref_off, int64_t n_refs, const int32_t* ref_idx,
                 const void* qry_bytes, const int64_t* qry_off, int64_t n_pairs, int32_t matrix_id,
                 const int64_t* out_off, bool need_out_off = true) {
    if (!ref_bytes || !ref_off || !qry_bytes || !qry_off || (need_out_off && !out_off)) return fail(GOTOH_B200_EINVAL, "NULL pointer argument");
    if (n_pairs < 0 || n_refs < 0) return fail(GOTOH_B200_EINVAL, "negative count");
    if (n_pairs > 0x7fffffffLL) return fail(GOTOH_B200_ERANGE, "more than 2^31-1 pairs in one call");
    if (!ref_idx && n_refs != n_pairs) return fail(GOTOH_B200_EINVAL, "ref_idx is NULL but n_refs != n_pairs");
    if (matrix_id < 0 || matrix_id > 2) return fail(GOTOH_B200_EINVAL, "matrix_id %d not in {0,1,2}", matrix_id);
    return GOTOH_B200_OK;
}

// ---- cached per-device contexts for the one-shot call: four workspaces = four slabs in flight ----
// (two being packed by the two builder threads, one in its kernels, one copying back; with two, the host could not start packing
// slab s+1 before slab s-1 had finished its D2H and the kernels idled ~30 % of the time - measured on B200)
enum { NWS = 8, NBUILD = 4 };
struct DeviceCtx {
    std::mutex mu;
    Workspace ws[NWS];
    int64_t free_plus_cached = 0;   // device memory available to the workspaces, measured on the first call (cudaMemGetInfo
                                    // now and then takes tens of ms; gotoh_b200_release_cache resets it)
};
std::mutex g_ctx_mu;
DeviceCtx* g_ctx[64] = {nullptr};

DeviceCtx* ctx_for(int dev) {
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    if (!g_ctx[dev]) g_ctx[dev] = new (std::nothrow) DeviceCtx();
    return g_ctx[dev];
}

// One device's share [lo, hi) of a one-shot call: cut into slabs of bounded arena size, each slab is a
// self-contained plan on one of three workspaces (streams), so host packing of slab s+1 and the D2H of
// slab s-1 overlap the kernels of slab s.
int run_device_range(int dev, const uint8_t* ref_bytes, const int64_t* ref_off, int64_t n_refs,
                     const int32_t* ref_idx, const uint8_t* qry_bytes, const int64_t* qry_off,
                     int64_t lo, int64_t hi, int32_t gip, int32_t gep, int32_t term, int32_t matrix_id, const OutSpec& out) {
    const double t_entry = now_ms();
    DeviceCtx* ctx = ctx_for(dev);
    if (!ctx) return fail(GOTOH_B200_ENOMEM, "out of host memory");
    std::lock_guard<std::mutex> lk(ctx->mu);
    // workspaces in flight (GOTOH_B200_WORKSPACES, default 8) and builder threads (GOTOH_B200_BUILDERS, default 2).  B200
    // trace with 4 workspaces: each builder spent half its time waiting for the workspace it was about to reuse while the
    // GPU idled 22 % of the call - a builder must be able to run several slabs ahead of the kernels.
    int nws = getenv("GOTOH_B200_WORKSPACES") ? atoi(getenv("GOTOH_B200_WORKSPACES")) : 8;
    nws = std::max(1, std::min(nws, (int)NWS));
    for (int w = 0; w < nws; ++w) { const int rc = ctx->ws[w].init(dev); if (rc) return rc; }
    CU(cudaSetDevice(dev));
    {
        // re-measured until the workspaces hold their arenas; after that a call needs no new device memory
        int64_t cached = 0;
        for (int w = 0; w < NWS; ++w) cached += (int64_t)ctx->ws[w].d_dir.cap * 16;
        if (ctx->free_plus_cached <= 0 || cached == 0) {
            size_t free_b = 0, total_b = 0;
            CU(cudaMemGetInfo(&free_b, &total_b));
            ctx->free_plus_cached = (int64_t)free_b + cached;
        }
    }
    // small slabs keep the pipeline's fill and drain short (first packing, last D2H); 1.5 GB of arena = ~8 k reads.  (With
    // the cheaper host packing and the FIFO of forward kernels, 1-1.5 GB measured ~6 % faster end to end than 3 GB on B200.)
    int64_t slab_budget = std::min<int64_t>(ctx->free_plus_cached / (nws + 2), (int64_t)1536 << 20);
    slab_budget = std::max<int64_t>(slab_budget, (int64_t)256 << 20);
    if (getenv("GOTOH_B200_SLAB_MB")) slab_budget = (int64_t)atoll(getenv("GOTOH_B200_SLAB_MB")) << 20;   // tests
    gotoh_b200_plan plans[NWS];
    for (int w = 0; w < nws; ++w) {
        plans[w].ws = &ctx->ws[w];
        plans[w].gip = gip; plans[w].gep = gep; plans[w].term = term ? 1 : 0; plans[w].matrix = matrix_id;
        plans[w].out_mode = out.mode;
        plans[w].arena_budget_bytes = slab_budget;
        plans[w].task_limit = getenv("GOTOH_B200_TASK_LIMIT") ? atoi(getenv("GOTOH_B200_TASK_LIMIT")) : 0;   // measured: no gain over persistent warps
    }
    // slab = as many consecutive pairs as fit the arena estimate ((M+40)*64 B per strip per pair).  The first two slabs get
    // a quarter of the budget: the result copy - which bounds short-pair batches (C3: 720 MB of strings against 8 ms of
    // kernels) - and the first forward kernel then start after a quarter of the packing / H2D / preparation time.
    // (GOTOH_B200_RAMP=0: equal slabs.)  The per-pair need is summed per block of 4096 pairs by a few threads; the cuts walk
    // the block sums and scan pairs only inside the block a cut falls into (1 M pairs: 1.5 -> 0.3 ms; a full per-pair
    // prefix array cost more in page faults than the serial loop it replaced).
    const int64_t npairs = hi - lo;
    auto pair_need = [&](int64_t e, double* cells) -> int64_t {
        const int64_t r = ref_idx ? ref_idx[lo + e] : lo + e;
        const int64_t m = (r >= 0 && r < n_refs) ? ref_off[r + 1] - ref_off[r] : 0;
        const int64_t nq = qry_off[lo + e + 1] - qry_off[lo + e];
        *cells += (double)m * (double)nq;
        return (std::max<int64_t>(nq, 1) + 255) / 256 * (m + 40) * 64;
    };
    enum { CUT_BLOCK = 4096 };
    const int64_t nblocks = (npairs + CUT_BLOCK - 1) / CUT_BLOCK;
    std::vector<int64_t> blk_need((size_t)nblocks);
    std::vector<double> blk_cells((size_t)nblocks);
    parallel_for(nblocks, (int)std::max<int64_t>(1, std::min<int64_t>(8, nblocks / 16)), [&](int64_t b0, int64_t b1, int) {
        for (int64_t bk = b0; bk < b1; ++bk) {
            int64_t acc = 0;
            double c = 0.0;
            for (int64_t e = bk * CUT_BLOCK; e < std::min<int64_t>(npairs, (bk + 1) * CUT_BLOCK); ++e) acc += pair_need(e, &c);
            blk_need[(size_t)bk] = acc; blk_cells[(size_t)bk] = c;
        }
    });
    const bool ramp = !(getenv("GOTOH_B200_RAMP") && atoi(getenv("GOTOH_B200_RAMP")) == 0);
    std::vector<int64_t> cuts(1, lo);
    std::vector<double> slab_cells;                  // estimated DP cells per slab (untrimmed lengths)
    for (int64_t k = 0; k < npairs;) {
        const int64_t budget_here = (ramp && cuts.size() <= 2) ? std::min<int64_t>(slab_budget, std::max<int64_t>(slab_budget / 4, (int64_t)64 << 20)) : slab_budget;
        int64_t est = 0, e = k;
        double cells = 0.0;
        while (e < npairs && e - k < (1 << 20)) {
            if (e % CUT_BLOCK == 0 && e + CUT_BLOCK <= npairs && e - k + CUT_BLOCK <= (1 << 20) && est + blk_need[(size_t)(e / CUT_BLOCK)] <= budget_here) {
                est += blk_need[(size_t)(e / CUT_BLOCK)]; cells += blk_cells[(size_t)(e / CUT_BLOCK)]; e += CUT_BLOCK;   // a whole block fits
                continue;
            }
            double c1 = 0.0;
            const int64_t need = pair_need(e, &c1);
            if (e > k && est + need > budget_here) break;
            est += need; cells += c1;
            ++e;
        }
        cuts.push_back(lo + e);
        slab_cells.push_back(cells);
        k = e;
    }
    const int nslabs = (int)cuts.size() - 1;
    const double t_cuts = now_ms();
    if (trace_on()) { cudaEventRecord(ctx->ws[0].ev[3], ctx->ws[0].stream); cudaEventSynchronize(ctx->ws[0].ev[3]); for (int w = 0; w < NWS; ++w) ctx->ws[w].trace_slab = -1; }
    // Two builder threads pack alternate slabs (each owns half of the workspaces), so the host's packing rate is
    // not the pipeline's bottleneck: per slab the host needs about as long as the kernels (measured on B200).
    int builders = getenv("GOTOH_B200_BUILDERS") ? atoi(getenv("GOTOH_B200_BUILDERS")) : 2;
    builders = std::max(1, std::min(std::min(builders, (int)NBUILD), std::min(nslabs, nws)));
    std::vector<int> rcs((size_t)builders + 1, 0);        // last entry: the collector
    std::vector<std::string> msgs((size_t)builders + 1);
    // Kernels are ENQUEUED in slab order (the builders pack in parallel, then take turns), and the forward kernel of slab
    // s waits for the forward kernel of slab s - depth: without this all slabs in flight time-slice the SMs, finish
    // together and then queue up on the copy engine while the builders - waiting for a drained workspace - leave the
    // GPU idle (B200 trace: 30 ms of a 173 ms call).  depth = 2 keeps one more forward kernel resident to fill the tail
    // of the one ahead; walk/emit/D2H of a slab overlap the forward kernels of the next ones.
    int depth = getenv("GOTOH_B200_FWD_DEPTH") ? atoi(getenv("GOTOH_B200_FWD_DEPTH")) : 2;
    // Shared state of the builders and the collector.  EVERY change of `failed`, `next_enqueue` or `state` happens under
    // order_mu and is followed by notify_all, so no waiter can miss it (a builder that failed outside the ordered section
    // used to leave the other one asleep in order_cv.wait).
    std::mutex order_mu;
    std::condition_variable order_cv;
    bool failed = false;
    int next_enqueue = 0;
    std::vector<int> slab_ws((size_t)nslabs, -1);
    // tight / compact forms: 0 = not enqueued, 1 = kernels + phase A enqueued, 2 = phase B enqueued (workspace may be recycled once its stream drains)
    const bool two_phase = out.mode != OUT_STRIDED;
    std::vector<int> state((size_t)nslabs, 0);
    auto fail_and_wake = [&](int who, int rc) {
        std::lock_guard<std::mutex> g(order_mu);
        if (!rcs[(size_t)who]) { rcs[(size_t)who] = rc; msgs[(size_t)who] = g_err; }
        failed = true;
        order_cv.notify_all();
    };
    auto is_failed = [&]() { std::lock_guard<std::mutex> g(order_mu); return failed; };
    const char* inject_env = getenv("GOTOH_B200_TEST_FAIL_FETCH");                        // tests: a D2H enqueue that fails,
    const int inject_fetch_slab = inject_env ? atoi(inject_env) : -1;                     // at the slab of this index
    auto builder_body = [&](int b) {
        if (cudaSetDevice(dev) != cudaSuccess) { fail_and_wake(b, fail(GOTOH_B200_ECUDA, "cudaSetDevice failed")); return; }
        int mine = 0;
        for (int slab = b; slab < nslabs && !is_failed(); slab += builders, ++mine) {
            // builder b owns workspaces b, b+builders, ...; its previous slab on that workspace must have drained
            const int per = nws / builders;
            gotoh_b200_plan* pl = &plans[b + builders * (mine % per)];
            int rc = GOTOH_B200_OK;
            const double t_a = now_ms();
            if (two_phase && mine >= per) {
                // the collector must have enqueued the result copy of the slab that used this workspace before
                const int prev = slab - builders * per;
                std::unique_lock<std::mutex> g(order_mu);
                order_cv.wait(g, [&] { return state[(size_t)prev] == 2 || failed; });
                if (failed) return;
            }
            const double t_a2 = now_ms();
            if (pl->ws->ev_done_pending) {
                if (cudaEventSynchronize(pl->ws->ev_done) != cudaSuccess) rc = fail(GOTOH_B200_ECUDA, "a slab's kernels or copies failed: %s", cudaGetErrorString(cudaGetLastError()));
                pl->ws->ev_done_pending = false;
            }
            if (trace_on() && two_phase)
                fprintf(stderr, "[gotoh_b200] dev %d builder %d slab %d: workspace released by the collector at %.2f ms (waited %.2f), stream drained at %.2f ms\n",
                        dev, b, slab, t_a2 - t_entry, t_a2 - t_a, now_ms() - t_entry);
            if (trace_on() && pl->ws->trace_slab >= 0) {
                float k0 = 0, k1 = 0, d1 = 0;
                cudaEventElapsedTime(&k0, ctx->ws[0].ev[3], pl->ws->ev[0]);
                cudaEventElapsedTime(&k1, ctx->ws[0].ev[3], pl->ws->ev[1]);
                cudaEventElapsedTime(&d1, ctx->ws[0].ev[3], pl->ws->ev[2]);
                fprintf(stderr, "[gotoh_b200] gpu timeline slab %d: kernels %.2f .. %.2f ms, d2h done %.2f ms\n", pl->ws->trace_slab, k0, k1, d1);
                pl->ws->trace_slab = -1;
            }
            const double t_b = now_ms();
            for (double& x : g_trace_phase) x = 0;
            if (!rc) {
                try {
                    // Who lays the slab out?  The device builder costs a GPU round trip - its kernels queue behind the persistent
                    // forward CTAs of the slabs ahead, up to one forward-kernel duration (2 ms measured on C2) - the host builder
                    // ~0.2 us of one core per pair.  Pairs with many cells (a 251-nt read against a 3 kb standard: 0.15 us of GPU
                    // time) leave a core time to keep up; short pairs (84-aa windows: 0.009 us) do not.
                    bool on_device = false;
                    const double per_pair = slab_cells[(size_t)slab] / (double)std::max<int64_t>(1, cuts[(size_t)slab + 1] - cuts[(size_t)slab]);
                    const char* sw = getenv("GOTOH_B200_DEVICE_PREP");
                    if (sw ? atoi(sw) != 0 : per_pair < 262144.0)
                        rc = plan_build_device(pl, ref_bytes, ref_off, n_refs, ref_idx, qry_bytes, qry_off, cuts[(size_t)slab], cuts[(size_t)slab + 1], out.out_off_in, &on_device);
                    if (!rc && !on_device)
                        rc = plan_build(pl, ref_bytes, ref_off, n_refs, ref_idx, qry_bytes, qry_off, cuts[(size_t)slab], cuts[(size_t)slab + 1], out.out_off_in);
                } catch (const std::bad_alloc&) {
                    rc = fail(GOTOH_B200_ENOMEM, "out of host memory while packing");
                }
            }
            const double t_c = now_ms();
            {
                std::unique_lock<std::mutex> g(order_mu);
                order_cv.wait(g, [&] { return next_enqueue == slab || failed; });
                if (!rc && !failed) {
                    slab_ws[(size_t)slab] = (int)(pl->ws - &ctx->ws[0]);
                    if (depth > 0 && slab - depth >= 0 && slab_ws[(size_t)(slab - depth)] >= 0 &&
                        cudaStreamWaitEvent(pl->ws->stream, ctx->ws[slab_ws[(size_t)(slab - depth)]].ev_fwd, 0) != cudaSuccess)
                        rc = fail(GOTOH_B200_ECUDA, "cudaStreamWaitEvent failed");
                    pl->mark_forward_done = true;
                    if (trace_on() && !rc) cudaEventRecord(pl->ws->ev[0], pl->ws->stream);
                    if (!rc) rc = plan_run(pl, false, nullptr, nullptr);
                }
                next_enqueue = slab + 1;
                if (rc && !rcs[(size_t)b]) { rcs[(size_t)b] = rc; msgs[(size_t)b] = g_err; failed = true; }
                order_cv.notify_all();
                if (failed) return;
            }
            if (trace_on()) cudaEventRecord(pl->ws->ev[1], pl->ws->stream);
            if (slab == inject_fetch_slab) rc = fail(GOTOH_B200_ECUDA, "injected result-copy failure (GOTOH_B200_TEST_FAIL_FETCH)");
            if (!rc) rc = two_phase ? plan_fetch_a(pl) : plan_fetch(pl, out.out_ref, out.out_qry, out.out_len, out.out_score);
            if (!rc && !two_phase) {
                if (cudaEventRecord(pl->ws->ev_done, pl->ws->stream) != cudaSuccess) rc = fail(GOTOH_B200_ECUDA, "cudaEventRecord failed");
                pl->ws->ev_done_pending = true;
            }
            if (trace_on() && !rc) { cudaEventRecord(pl->ws->ev[2], pl->ws->stream); pl->ws->trace_slab = slab; }
            if (trace_on())
                fprintf(stderr, "[gotoh_b200] dev %d builder %d slab %d pairs %lld: wait %.1f ms, build %.1f ms (refs %.1f pass1 %.1f pass2 %.1f cls %.1f path %.1f sort %.1f tasks %.1f alloc+h2d %.1f), enqueue %.1f ms\n",
                        dev, b, slab, (long long)(cuts[(size_t)slab + 1] - cuts[(size_t)slab]), t_b - t_a, t_c - t_b, g_trace_phase[0], g_trace_phase[1],
                        g_trace_phase[2], g_trace_phase[3], g_trace_phase[4], g_trace_phase[5], g_trace_phase[6], g_trace_phase[7], now_ms() - t_c);
            if (rc) { fail_and_wake(b, rc); return; }
            if (two_phase) {
                std::lock_guard<std::mutex> g(order_mu);
                state[(size_t)slab] = 1;
                order_cv.notify_all();
            }
        }
    };
    // a builder thread must not die on an exception (std::terminate) nor leave the others waiting
    auto builder = [&](int b) {
        try {
            builder_body(b);
        } catch (const std::bad_alloc&) {
            fail_and_wake(b, fail(GOTOH_B200_ENOMEM, "out of host memory in a builder thread"));
        } catch (const std::exception& e) {
            fail_and_wake(b, fail(GOTOH_B200_ECUDA, "internal error in a builder thread: %s", e.what()));
        }
    };
    // The collector (tight / compact forms; the calling thread): in slab order, waits for a slab's phase A, learns its
    // total, and enqueues the result copy to the next free position of the caller's buffer.
    struct SlabBase { int64_t pair_lo, n, base; };
    std::vector<SlabBase> bases;                      // per slab: where its results went; offsets are slab-local until the end
    auto collector = [&]() {
        int64_t base = out.cap_lo;
        for (int slab = 0; slab < nslabs; ++slab) {
            {
                std::unique_lock<std::mutex> g(order_mu);
                order_cv.wait(g, [&] { return state[(size_t)slab] >= 1 || failed; });
                if (failed) return;
            }
            Workspace* ws = &ctx->ws[slab_ws[(size_t)slab]];
            gotoh_b200_plan* pl = &plans[slab_ws[(size_t)slab]];
            int rc = GOTOH_B200_OK;
            const double t_seen = now_ms();
            if (cudaEventSynchronize(ws->ev_a) != cudaSuccess) rc = fail(GOTOH_B200_ECUDA, "kernels of slab %d failed: %s", slab, cudaGetErrorString(cudaGetLastError()));
            const int64_t total = rc ? 0 : *ws->h_total.p;
            if (!rc && base + total > out.cap_hi)
                rc = fail(GOTOH_B200_ECAPACITY, "result buffer too small: pairs %lld..%lld need %lld %s at offset %lld of a %lld-%s slice (the worst-case bound of gotoh_b200.h always suffices)",
                          (long long)cuts[(size_t)slab], (long long)cuts[(size_t)slab + 1], (long long)total, out.mode == OUT_TIGHT ? "bytes" : "words",
                          (long long)(base - out.cap_lo), (long long)(out.cap_hi - out.cap_lo), out.mode == OUT_TIGHT ? "byte" : "word");
            const double t_ev = now_ms();
            if (!rc) rc = plan_fetch_b(pl, out, base, total);
            if (!rc) {
                if (cudaEventRecord(ws->ev_done, ws->stream) != cudaSuccess) rc = fail(GOTOH_B200_ECUDA, "cudaEventRecord failed");
                ws->ev_done_pending = true;
            }
            if (rc) { fail_and_wake(builders, rc); return; }
            if (trace_on())
                fprintf(stderr, "[gotoh_b200] dev %d collector slab %d: phase A seen at %.2f ms, done at %.2f ms, phase B enqueued at %.2f ms (total %lld)\n",
                        dev, slab, t_seen - t_entry, t_ev - t_entry, now_ms() - t_entry, (long long)total);
            bases.push_back({pl->pair_base, pl->n_pairs, base});
            base += total;
            std::lock_guard<std::mutex> g(order_mu);
            state[(size_t)slab] = 2;
            order_cv.notify_all();
        }
    };
    if (builders == 1 && !two_phase) builder(0);
    else if (nslabs == 1) {
        // a small call (one slab): no threads - the builder enqueues, then this thread collects
        builder(0);
        if (cudaSetDevice(dev) != cudaSuccess) fail_and_wake(builders, fail(GOTOH_B200_ECUDA, "cudaSetDevice failed"));
        else collector();
    } else {
        std::vector<std::thread> th;
        for (int b = 0; b < builders; ++b) th.emplace_back(builder, b);
        if (two_phase) {
            if (cudaSetDevice(dev) != cudaSuccess) fail_and_wake(builders, fail(GOTOH_B200_ECUDA, "cudaSetDevice failed"));
            else collector();
        }
        for (auto& t : th) t.join();
    }
    int rc = GOTOH_B200_OK;
    const double t_built = now_ms();
    for (int b = 0; b <= builders; ++b)
        if (rcs[(size_t)b] && !rc) rc = fail(rcs[(size_t)b], "%s", msgs[(size_t)b].c_str());
    for (int w = 0; w < nws; ++w) {
        ctx->ws[w].ev_done_pending = false;
        const cudaError_t e = cudaStreamSynchronize(ctx->ws[w].stream);
        if (e != cudaSuccess && !rc) rc = fail(GOTOH_B200_ECUDA, "stream synchronize failed: %s", cudaGetErrorString(e));
    }
    // the copies have landed: turn the slab-local offsets into positions in the caller's buffer
    if (!rc)
        for (const SlabBase& sb : bases)
            if (sb.base) { int64_t* off = out.out_off + sb.pair_lo; for (int64_t k = 0; k < sb.n; ++k) off[k] += sb.base; }
    if (trace_on())
        fprintf(stderr, "[gotoh_b200] dev %d call: setup+cuts %.2f ms (%d slabs), builders %.2f ms, drain %.2f ms\n", dev, t_cuts - t_entry, nslabs,
                t_built - t_cuts, now_ms() - t_built);
    return rc;
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
    Workspace* ws = new (std::nothrow) Workspace();
    if (!pl || !ws) { delete pl; delete ws; return fail(GOTOH_B200_ENOMEM, "out of host memory"); }
    pl->ws = ws;
    pl->owns_ws = true;
    pl->gip = gip; pl->gep = gep; pl->term = use_terminal ? 1 : 0; pl->matrix = matrix_id;
    rc = ws->init(device);
    for (double& x : g_trace_phase) x = 0;
    if (!rc) {
        try {
            bool on_device = false;
            rc = plan_build_device(pl, ref_bytes, ref_off, n_refs, ref_idx, qry_bytes, qry_off, 0, n_pairs, out_off, &on_device);
            if (!rc && !on_device) rc = plan_build(pl, ref_bytes, ref_off, n_refs, ref_idx, qry_bytes, qry_off, 0, n_pairs, out_off);
        } catch (const std::bad_alloc&) {
            rc = fail(GOTOH_B200_ENOMEM, "out of host memory while packing");
        }
    }
    if (!rc && cudaStreamSynchronize(ws->stream) != cudaSuccess) rc = fail(GOTOH_B200_ECUDA, "H2D copy failed");
    if (rc) { delete pl; return rc; }
    *plan_out = pl;
    return GOTOH_B200_OK;
}

extern "C" int32_t gotoh_b200_plan_run(gotoh_b200_plan* plan, float* device_ms, float* forward_ms) {
    if (!plan) return fail(GOTOH_B200_EINVAL, "plan is NULL");
    return plan_run(plan, true, device_ms, forward_ms);
}

extern "C" int32_t gotoh_b200_plan_fetch(gotoh_b200_plan* plan, uint8_t* out_ref, uint8_t* out_qry,
                                         int32_t* out_len, int32_t* out_score) {
    if (!plan || !out_ref || !out_qry || !out_len || !out_score) return fail(GOTOH_B200_EINVAL, "NULL argument");
    const int rc = plan_fetch(plan, out_ref, out_qry, out_len, out_score);
    if (rc) return rc;
    CU(cudaStreamSynchronize(plan->ws->stream));
    return GOTOH_B200_OK;
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
        case 8: return pl->built_on_device ? 1 : 0;
    }
    return -1;
}

namespace { void g2_release_cache(); }   // gotoh2_host.cuh

extern "C" void gotoh_b200_release_cache(void) {
    g2_release_cache();
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    for (int d = 0; d < 64; ++d)
        if (g_ctx[d]) {
            std::lock_guard<std::mutex> lk2(g_ctx[d]->mu);
            for (int w = 0; w < NWS; ++w) g_ctx[d]->ws[w].release();
            g_ctx[d]->free_plus_cached = 0;
        }
}

// One-shot forms.  Shard contiguous pair ranges of (nearly) equal cell count across the
// devices in device_mask; one host thread per device; no inter-device traffic.
namespace {
int align_batch_impl(const uint8_t* ref_bytes, const int64_t* ref_off, int64_t n_refs, const int32_t* ref_idx,
                     const uint8_t* qry_bytes, const int64_t* qry_off, int64_t n_pairs, int32_t gip, int32_t gep,
                     int32_t use_terminal, int32_t matrix_id, const OutSpec& out, int64_t cap, uint32_t device_mask) {
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
    cut[D] = n_pairs;
    std::vector<OutSpec> outs((size_t)D, out);
    outs[0].cap_lo = 0; outs[(size_t)D - 1].cap_hi = cap;
    if (D > 1) {
        // Contiguous split of equal estimated cell count.  Tight / compact forms: each device fills its own slice of the
        // caller's capacity, in proportion to the worst-case size of its pairs (bytes: M+N; words: ceil((M+N)/16)).
        // Both come from per-block sums built by a few threads; a cut scans pairs only inside its block (a per-pair prefix
        // array cost ~6 ms per call of 1 M pairs in one host process driving 8 GPUs).
        auto pair_cost = [&](int64_t k, double* w) -> double {
            const int64_t r = ref_idx ? ref_idx[k] : k;
            const double m = (r >= 0 && r < n_refs) ? (double)(ref_off[r + 1] - ref_off[r]) : 1.0;
            const double nq = (double)(qry_off[k + 1] - qry_off[k]);
            *w += (out.mode == OUT_COMPACT ? std::floor((m + nq + 15) / 16) : m + nq);
            return std::max(1.0, m) * std::max<double>(1.0, nq);
        };
        enum { SPLIT_BLOCK = 4096 };
        const int64_t nblocks = (n_pairs + SPLIT_BLOCK - 1) / SPLIT_BLOCK;
        std::vector<double> blk_c((size_t)nblocks), blk_w((size_t)nblocks);
        parallel_for(nblocks, (int)std::max<int64_t>(1, std::min<int64_t>(8, nblocks / 16)), [&](int64_t b0, int64_t b1, int) {
            for (int64_t bk = b0; bk < b1; ++bk) {
                double c = 0.0, w = 0.0;
                for (int64_t k = bk * SPLIT_BLOCK; k < std::min<int64_t>(n_pairs, (bk + 1) * SPLIT_BLOCK); ++k) c += pair_cost(k, &w);
                blk_c[(size_t)bk] = c; blk_w[(size_t)bk] = w;
            }
        });
        double total_c = 0.0, total_w = 0.0;
        for (int64_t bk = 0; bk < nblocks; ++bk) { total_c += blk_c[(size_t)bk]; total_w += blk_w[(size_t)bk]; }
        int64_t bk = 0;
        double acc_c = 0.0, acc_w = 0.0;                       // sums of the blocks before bk
        for (int d = 1; d < D; ++d) {
            const double target = total_c * d / D;
            while (bk < nblocks && acc_c + blk_c[(size_t)bk] < target) { acc_c += blk_c[(size_t)bk]; acc_w += blk_w[(size_t)bk]; ++bk; }
            int64_t k = std::min(n_pairs, bk * SPLIT_BLOCK);
            double c = acc_c, w = acc_w;
            while (k < std::min<int64_t>(n_pairs, (bk + 1) * SPLIT_BLOCK) && c < target) c += pair_cost(k++, &w);
            // every device gets at least one pair: clamping moves the cut by a few pairs at most, the capacity share follows it
            const int64_t kc = std::min<int64_t>(std::max(k, cut[d - 1] + 1), n_pairs - (D - d));
            for (; k < kc; ++k) (void)pair_cost(k, &w);
            for (; k > kc; --k) { double back = 0.0; (void)pair_cost(k - 1, &back); w -= back; }
            cut[d] = kc;
            const double f = total_w > 0 ? w / total_w : 0.0;
            outs[(size_t)d].cap_lo = outs[(size_t)d - 1].cap_hi = (int64_t)std::floor((double)cap * f);
        }
    }
    std::vector<int> rcs(D, 0);
    std::vector<std::string> msgs(D);
    auto work = [&](int d) {
        rcs[d] = run_device_range(devs[d], ref_bytes, ref_off, n_refs, ref_idx, qry_bytes, qry_off, cut[d], cut[d + 1],
                                  gip, gep, use_terminal, matrix_id, outs[(size_t)d]);
        if (rcs[d]) msgs[d] = g_err;
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
// no C++ exception crosses the C boundary
int align_batch_guarded(const uint8_t* ref_bytes, const int64_t* ref_off, int64_t n_refs, const int32_t* ref_idx,
                        const uint8_t* qry_bytes, const int64_t* qry_off, int64_t n_pairs, int32_t gip, int32_t gep,
                        int32_t use_terminal, int32_t matrix_id, const OutSpec& out, int64_t cap, uint32_t device_mask) {
    try {
        return align_batch_impl(ref_bytes, ref_off, n_refs, ref_idx, qry_bytes, qry_off, n_pairs, gip, gep, use_terminal, matrix_id, out, cap, device_mask);
    } catch (const std::bad_alloc&) {
        return fail(GOTOH_B200_ENOMEM, "out of host memory");
    } catch (const std::exception& e) {
        return fail(GOTOH_B200_ECUDA, "internal error: %s", e.what());
    }
}
}  // namespace

extern "C" int32_t gotoh_b200_align_batch(const uint8_t* ref_bytes, const int64_t* ref_off, int64_t n_refs,
                                          const int32_t* ref_idx, const uint8_t* qry_bytes,
                                          const int64_t* qry_off, int64_t n_pairs, int32_t gip, int32_t gep,
                                          int32_t use_terminal, int32_t matrix_id, uint8_t* out_ref,
                                          uint8_t* out_qry, const int64_t* out_off, int32_t* out_len,
                                          int32_t* out_score, uint32_t device_mask) {
    int rc = check_common(ref_bytes, ref_off, n_refs, ref_idx, qry_bytes, qry_off, n_pairs, matrix_id, out_off);
    if (rc) return rc;
    if (!out_ref || !out_qry || !out_len || !out_score) return fail(GOTOH_B200_EINVAL, "NULL output pointer");
    OutSpec o;
    o.mode = OUT_STRIDED; o.out_ref = out_ref; o.out_qry = out_qry; o.out_off_in = out_off; o.out_len = out_len; o.out_score = out_score;
    return align_batch_guarded(ref_bytes, ref_off, n_refs, ref_idx, qry_bytes, qry_off, n_pairs, gip, gep, use_terminal, matrix_id, o, 0, device_mask);
}

extern "C" int32_t gotoh_b200_align_batch_tight(const uint8_t* ref_bytes, const int64_t* ref_off, int64_t n_refs,
                                                const int32_t* ref_idx, const uint8_t* qry_bytes,
                                                const int64_t* qry_off, int64_t n_pairs, int32_t gip, int32_t gep,
                                                int32_t use_terminal, int32_t matrix_id, uint8_t* out_ref,
                                                uint8_t* out_qry, int64_t out_cap, int64_t* out_off, int32_t* out_len,
                                                int32_t* out_score, uint32_t device_mask) {
    int rc = check_common(ref_bytes, ref_off, n_refs, ref_idx, qry_bytes, qry_off, n_pairs, matrix_id, nullptr, false);
    if (rc) return rc;
    if (!out_ref || !out_qry || !out_off || !out_len || !out_score || out_cap < 0) return fail(GOTOH_B200_EINVAL, "NULL output pointer or negative capacity");
    OutSpec o;
    o.mode = OUT_TIGHT; o.out_ref = out_ref; o.out_qry = out_qry; o.out_off = out_off; o.out_len = out_len; o.out_score = out_score;
    return align_batch_guarded(ref_bytes, ref_off, n_refs, ref_idx, qry_bytes, qry_off, n_pairs, gip, gep, use_terminal, matrix_id, o, out_cap, device_mask);
}

extern "C" int32_t gotoh_b200_align_batch_compact(const uint8_t* ref_bytes, const int64_t* ref_off, int64_t n_refs,
                                                  const int32_t* ref_idx, const uint8_t* qry_bytes,
                                                  const int64_t* qry_off, int64_t n_pairs, int32_t gip, int32_t gep,
                                                  int32_t use_terminal, int32_t matrix_id, int32_t* out_rec,
                                                  uint32_t* out_ops, int64_t out_ops_cap, int64_t* out_ops_off,
                                                  uint32_t device_mask) {
    int rc = check_common(ref_bytes, ref_off, n_refs, ref_idx, qry_bytes, qry_off, n_pairs, matrix_id, nullptr, false);
    if (rc) return rc;
    if (!out_rec || !out_ops || !out_ops_off || out_ops_cap < 0) return fail(GOTOH_B200_EINVAL, "NULL output pointer or negative capacity");
    OutSpec o;
    o.mode = OUT_COMPACT; o.out_rec = out_rec; o.out_ops = out_ops; o.out_off = out_ops_off;
    return align_batch_guarded(ref_bytes, ref_off, n_refs, ref_idx, qry_bytes, qry_off, n_pairs, gip, gep, use_terminal, matrix_id, o, out_ops_cap, device_mask);
}

// Host-ceiling probe: plain device-to-host copies into the caller's (pinned) buffer, CUDA-event timed.
extern "C" int32_t gotoh_b200_d2h_probe(int32_t device, void* host_buf, int64_t bytes, int32_t reps, double* seconds) {
    if (!host_buf || !seconds || bytes <= 0 || reps <= 0) return fail(GOTOH_B200_EINVAL, "d2h_probe: bad argument");
    const int ndev = gotoh_b200_device_count();
    if (ndev <= 0 || device < 0 || device >= ndev) return fail(GOTOH_B200_ENODEVICE, "device %d not present", device);
    CU(cudaSetDevice(device));
    // scratch of at most 1 GB, copied repeatedly to successive positions of the host buffer
    const int64_t piece = std::min<int64_t>(bytes, (int64_t)1 << 30);
    void* d = nullptr;
    cudaStream_t st = 0;
    cudaEvent_t e0 = 0, e1 = 0;
    float ms = 0.f;
    auto run = [&]() -> int {
        CU(cudaMalloc(&d, (size_t)piece));
        CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        CU(cudaEventCreate(&e0));
        CU(cudaEventCreate(&e1));
        CU(cudaMemsetAsync(d, 0x2d, (size_t)piece, st));
        CU(cudaStreamSynchronize(st));
        CU(cudaEventRecord(e0, st));
        for (int r = 0; r < reps; ++r)
            for (int64_t at = 0; at < bytes; at += piece)
                CU(cudaMemcpyAsync((char*)host_buf + at, d, (size_t)std::min<int64_t>(piece, bytes - at), cudaMemcpyDeviceToHost, st));
        CU(cudaEventRecord(e1, st));
        CU(cudaEventSynchronize(e1));
        CU(cudaEventElapsedTime(&ms, e0, e1));
        return GOTOH_B200_OK;
    };
    const int rc = run();
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (st) cudaStreamDestroy(st);
    if (d) cudaFree(d);
    if (rc) return rc;
    *seconds = ms * 1e-3;
    return GOTOH_B200_OK;
}

#include "gotoh2_host.cuh"

extern "C" int32_t gotoh_b200_int_peak(int32_t device, int32_t which, double* ginstr_per_s) {
    if (!ginstr_per_s) return fail(GOTOH_B200_EINVAL, "NULL argument");
    const int ndev = gotoh_b200_device_count();
    if (ndev <= 0 || device < 0 || device >= ndev) return fail(GOTOH_B200_ENODEVICE, "device %d not present", device);
    CU(cudaSetDevice(device));
    return intpeak::run(which, ginstr_per_s, g_err, sizeof(g_err));
}

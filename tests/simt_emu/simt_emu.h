// simt_emu.h - TEST-ONLY CPU emulation of the slice of the CUDA execution model that
// micall-lite_b200/csrc uses, so the exact kernel sources can be exercised against
// the oracle in the GPU-less build container (tests/test_emu_*.py).
//
// THIS IS NOT A PRODUCT PATH.  The emulated library is built into tests/_build/ by
// tests/simt_emu/build_emu.py, is never installed next to the package, and the package
// loader (gotoh_b200/_ffi.py) only ever opens micall-lite_b200/lib/libgotoh_b200.so,
// which requires a CUDA device.
//
// Model: every CUDA thread of a block is a ucontext fiber on one OS thread; blocks run
// one after another.  __syncthreads and the warp collectives are cooperative barriers,
// so lanes observe exactly the lock-step data exchange the hardware gives them.
#pragma once
#define GOTOH_SIMT_EMU 1

#include <ucontext.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

// ---- CUDA keywords -----------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__ static thread_local
#define __constant__ static
#define __align__(n) __attribute__((aligned(n)))

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct int2 { int x, y; };
struct uint2 { unsigned x, y; };
struct __attribute__((aligned(16))) int4 { int x, y, z, w; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }

namespace simt {

struct Warp {
    uint64_t slot[2][32];
    int arrived = 0;
    unsigned gen = 0;
    int live = 0;
};

struct Fiber {
    ucontext_t ctx;
    char* stack = nullptr;
    bool done = false;
    dim3 tid;
    unsigned linear = 0;
};

struct Block {
    std::vector<Fiber> fibers;
    ucontext_t sched;
    int cur = 0;
    int live = 0;
    int bar_count = 0;
    unsigned bar_gen = 0;
    std::vector<Warp> warps;
    dim3 bidx, bdim, gdim;
    unsigned char* dyn_smem = nullptr;
    const std::function<void()>* body = nullptr;
};

extern thread_local Block* B;

inline void yield() {
    Block* b = B;
    swapcontext(&b->fibers[b->cur].ctx, &b->sched);
}

inline void block_barrier() {
    Block* b = B;
    unsigned gen = b->bar_gen;
    if (++b->bar_count >= b->live) { b->bar_count = 0; b->bar_gen++; }
    else while (b->bar_gen == gen) yield();
}

inline Warp& my_warp() { return B->warps[B->fibers[B->cur].linear >> 5]; }
inline int my_lane() { return (int)(B->fibers[B->cur].linear & 31); }

// All live lanes of the warp publish v, wait for each other, and may read any slot.
template <class T>
inline const uint64_t* warp_publish(T v) {
    static_assert(sizeof(T) <= 8, "collective payload");
    Warp& w = my_warp();
    unsigned gen = w.gen;
    int buf = gen & 1;
    uint64_t raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    w.slot[buf][my_lane()] = raw;
    if (++w.arrived >= w.live) { w.arrived = 0; w.gen++; }
    else while (w.gen == gen) yield();
    return w.slot[buf];
}

template <class T>
inline T slot_as(const uint64_t* s, int lane) {
    T r;
    std::memcpy(&r, &s[lane], sizeof(T));
    return r;
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);

}  // namespace simt

#define threadIdx (simt::B->fibers[simt::B->cur].tid)
#define blockIdx (simt::B->bidx)
#define blockDim (simt::B->bdim)
#define gridDim (simt::B->gdim)
#define warpSize 32

// ---- synchronisation and warp collectives -----------------------------------------
static inline void __syncthreads() { simt::block_barrier(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { (void)simt::warp_publish<int>(0); }

template <class T>
static inline T __shfl_sync(unsigned, T v, int src, int = 32) {
    const uint64_t* s = simt::warp_publish(v);
    return simt::slot_as<T>(s, src & 31);
}
template <class T>
static inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) {
    const uint64_t* s = simt::warp_publish(v);
    int lane = simt::my_lane();
    return lane >= (int)d ? simt::slot_as<T>(s, lane - (int)d) : v;
}
template <class T>
static inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) {
    const uint64_t* s = simt::warp_publish(v);
    int lane = simt::my_lane();
    return lane + (int)d < 32 ? simt::slot_as<T>(s, lane + (int)d) : v;
}
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) {
    const uint64_t* s = simt::warp_publish(v);
    return simt::slot_as<T>(s, (simt::my_lane() ^ m) & 31);
}
static inline unsigned __ballot_sync(unsigned, int pred) {
    const uint64_t* s = simt::warp_publish<int>(pred ? 1 : 0);
    unsigned r = 0;
    for (int l = 0; l < 32; ++l) r |= (unsigned)(simt::slot_as<int>(s, l) & 1) << l;
    return r;
}
static inline unsigned __match_any_sync(unsigned, int v) {
    const uint64_t* s = simt::warp_publish<int>(v);
    unsigned r = 0;
    for (int l = 0; l < 32; ++l) r |= (unsigned)(simt::slot_as<int>(s, l) == v) << l;
    return r;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xffffffffu; }

// ---- scalar intrinsics -------------------------------------------------------------
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __ffs(int x) { return __builtin_ffs(x); }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
    uint64_t v = ((uint64_t)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) {
        unsigned sel = (s >> (4 * i)) & 0xf;
        unsigned byte = (unsigned)(v >> (8 * (sel & 7))) & 0xff;
        if (sel & 8) byte = (byte & 0x80) ? 0xff : 0x00;
        r |= byte << (8 * i);
    }
    return r;
}
template <class T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
template <class T> static inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <class T> static inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline void __nanosleep(unsigned) { simt::yield(); }

// DPX (sm_90+/sm_100a: VIADDMNMX, VIMNMX3 and their .S16x2 forms)
static inline int __viaddmax_s32(int a, int b, int c) { return max((int)((unsigned)a + (unsigned)b), c); }
static inline int __vimax3_s32(int a, int b, int c) { return max(max(a, b), c); }
static inline int __viaddmin_s32(int a, int b, int c) { return min((int)((unsigned)a + (unsigned)b), c); }
static inline int __viaddmin_s32_relu(int a, int b, int c) { return max(min((int)((unsigned)a + (unsigned)b), c), 0); }
static inline int __vimin3_s32(int a, int b, int c) { return min(min(a, b), c); }
static inline int16_t simt_lo(unsigned x) { return (int16_t)(x & 0xffff); }
static inline int16_t simt_hi(unsigned x) { return (int16_t)(x >> 16); }
static inline unsigned simt_pack(int lo, int hi) { return ((unsigned)lo & 0xffffu) | ((unsigned)hi << 16); }
static inline unsigned __vadd2(unsigned a, unsigned b) {
    return simt_pack((int16_t)(simt_lo(a) + simt_lo(b)), (int16_t)(simt_hi(a) + simt_hi(b)));
}
static inline unsigned __vsub2(unsigned a, unsigned b) {
    return simt_pack((int16_t)(simt_lo(a) - simt_lo(b)), (int16_t)(simt_hi(a) - simt_hi(b)));
}
static inline unsigned __vmaxs2(unsigned a, unsigned b) {
    return simt_pack(max((int)simt_lo(a), (int)simt_lo(b)), max((int)simt_hi(a), (int)simt_hi(b)));
}
static inline int __vibmax_s32(int a, int b, bool* p) { *p = a >= b; return a >= b ? a : b; }
static inline unsigned __vibmax_s16x2(unsigned a, unsigned b, bool* phi, bool* plo) {
    *plo = simt_lo(a) >= simt_lo(b); *phi = simt_hi(a) >= simt_hi(b);
    return __vmaxs2(a, b);
}
static inline unsigned __viaddmax_s16x2(unsigned a, unsigned b, unsigned c) { return __vmaxs2(__vadd2(a, b), c); }
static inline unsigned __vmins2(unsigned a, unsigned b) {
    return simt_pack(min((int)simt_lo(a), (int)simt_lo(b)), min((int)simt_hi(a), (int)simt_hi(b)));
}
static inline unsigned __viaddmin_s16x2(unsigned a, unsigned b, unsigned c) { return __vmins2(__vadd2(a, b), c); }
static inline unsigned __viaddmin_s16x2_relu(unsigned a, unsigned b, unsigned c) { return __vmaxs2(__vmins2(__vadd2(a, b), c), 0u); }
static inline unsigned __vimin3_s16x2(unsigned a, unsigned b, unsigned c) { return __vmins2(__vmins2(a, b), c); }
static inline unsigned __vibmin_s16x2(unsigned a, unsigned b, bool* phi, bool* plo) {
    *plo = simt_lo(a) <= simt_lo(b); *phi = simt_hi(a) <= simt_hi(b);
    return __vmins2(a, b);
}
static inline unsigned __vimax3_s16x2(unsigned a, unsigned b, unsigned c) { return __vmaxs2(__vmaxs2(a, b), c); }

// ---- kernel launch -------------------------------------------------------------------
#define GOTOH_LAUNCH(kern, grid, block, smem, stream, ...) \
    simt::launch((grid), (block), (smem), [&]() { (kern)(__VA_ARGS__); })
#define GOTOH_DYN_SMEM(name) unsigned char* name = simt::B->dyn_smem

#include "simt_cuda_runtime.h"

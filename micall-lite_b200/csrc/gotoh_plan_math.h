// gotoh_plan_math.h - the per-pair decisions of the plan builder, shared by the host builder (plan_build, gotoh_b200.cu)
// and the device builder (gotoh_prep.cuh) so that both admit exactly the same pairs to the int16x2 kernels.
#pragma once

#include <stdint.h>

#ifndef GOTOH_HD
#if defined(__CUDACC__) && !defined(GOTOH_SIMT_EMU)
#define GOTOH_HD __host__ __device__ __forceinline__
#else
#define GOTOH_HD inline
#endif
#endif

namespace gotoh {

enum { PLAN_MAX_K = 8 };

GOTOH_HD bool plan_is_ws(uint8_t c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r'; }   // trim(): gotoh.cpp:545-559

// query columns per lane: the smallest supported K in {2, 3, 4, 6, 8} with lanes * K >= n
GOTOH_HD int plan_pick_K_lanes(int n, int lanes) {
    if (lanes * 2 >= n) return 2;
    if (lanes * 3 >= n) return 3;
    if (lanes * 4 >= n) return 4;
    if (lanes * 6 >= n) return 6;
    return 8;
}
GOTOH_HD int plan_pick_K(int n) { return plan_pick_K_lanes(n, 32); }
// Queries of at most 128 columns run as two 16-lane wavefronts per warp (k_forward<..., HALF>): K columns per lane with
// 16*K >= n.  Kernel selector of an int16x2 pair: K, plus 16 when it runs on half-warp wavefronts.
GOTOH_HD int plan_pick_KH(int n, bool half_off) {
    return (!half_off && n <= 16 * PLAN_MAX_K) ? (16 | plan_pick_K_lanes(n, 16)) : plan_pick_K(n);
}

// "range proof" for the int16x2 path: with rebase period R every value the Vec16 kernel forms for real cells stays
// inside int16.  All quantities in stored units.  The stored frame is shifted up by z4 (a multiple of
// 4, one per plan) so that no stored S^ and no diagonal candidate D^ is negative: the kernel then adds the packed
// substitution scores with one 32-bit multiply-add.  plan_int16_low_need() is the smallest such shift.
GOTOH_HD long long plan_int16_low_need(int M, int N, int gip, int gep, int minT) {
    const long long mn = M < N ? M : N;
    const long long mt = minT < 0 ? minT : 0;
    const long long smin = mt * mn;
    const long long vmin = 4 * (smin - 2LL * gip - gep) - 8;
    const long long add_lo = 4LL * ((long long)gip > -mt ? (long long)gip : -mt) + 8;
    return ((-(vmin - add_lo)) + 3) & ~3LL;
}
GOTOH_HD bool plan_fits_int16(int M, int N, int K, int R, int gip, int gep, int minT, int maxT, long long z4) {
    const long long mn = M < N ? M : N;
    const long long xt = maxT > 0 ? maxT : 0;
    const long long smax = xt * mn;
    const long long g = gep;
    const long long vmax = 4 * (smax + (R + 32LL * K + 2) * g) + 8;
    const long long add_hi = 4 * (xt + 2 * g) + 4;
    if (vmax + add_hi + z4 > 32000) return false;
    if (plan_int16_low_need(M, N, gip, gep, minT) > z4) return false;
    if (4LL * R * g > 30000) return false;   // the rebase delta itself must be an int16
    return true;
}

}  // namespace gotoh

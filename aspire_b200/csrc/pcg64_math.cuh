// PCG64 (numpy's default bit generator) arithmetic of asp_pcg64_uniform (csrc/rng.cu), host + device so that the CPU
// test suite runs the shipped code against numpy.random.Generator.random bit for bit (tests/native/pcg64_host.cu).
// Draw j is the XSL-RR output of the 128-bit LCG state after j + 1 steps; spec and pin: oracle/pcg64.py.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define ASP_PCG_HD __host__ __device__ __forceinline__
#else
#define ASP_PCG_HD inline
#endif

namespace asp {

typedef unsigned __int128 u128;
ASP_PCG_HD u128 pcg_mult() {
    return ((u128)0x2360ED051FC65DA4ULL << 64) | (u128)0x4385DF649FCCF645ULL;
}
// (A, C) such that state_{t + delta} = A state_t + C
ASP_PCG_HD void pcg_affine_power(uint64_t delta, u128 inc, u128& a_out, u128& c_out) {
    u128 acc_mult = 1, acc_plus = 0, cur_mult = pcg_mult(), cur_plus = inc;
    while (delta > 0) {
        if (delta & 1) {
            acc_mult *= cur_mult;
            acc_plus = acc_plus * cur_mult + cur_plus;
        }
        cur_plus = (cur_mult + 1) * cur_plus;
        cur_mult *= cur_mult;
        delta >>= 1;
    }
    a_out = acc_mult;
    c_out = acc_plus;
}
ASP_PCG_HD double pcg_double(u128 state) {
    const uint64_t hi = (uint64_t)(state >> 64), lo = (uint64_t)state;
    const uint64_t x = hi ^ lo;
    const unsigned rot = (unsigned)(hi >> 58);
    const uint64_t r = (x >> rot) | (x << ((64u - rot) & 63u));
    return (double)(r >> 11) * 0x1p-53;
}

}  // namespace asp

// Integer exponential spacings of the sorted-uniform multinomial resampler (csrc/sorted_resample.cu), host + device so
// that the CPU test suite runs the shipped arithmetic against the numpy spec bit for bit (tests/native/spacing_host.cu,
// oracle/sorted_resample.py).
//
// The order statistics of m iid uniforms are the normalised partial sums of m + 1 iid exponentials.  Element j draws
// the 53-bit integer k_j of the Philox uniform stream (u = k 2^-53, the conversion of asp_philox_uniform) and gets
//     q_j = floor(-ln(1 - u_j) 2^26)          (uint32; -ln(1 - u) <= 53 ln 2)
// Partial sums of the q_j are integers, so any scan order - tiles, blocks, ranks - gives the same S_j and the result
// does not depend on the launch geometry or on how slots are sharded.  ln is evaluated with IEEE add / mul / div only
// (no FMA contraction, no libm), so numpy reproduces it exactly:
//     m = 2^53 - k = f 2^p,  f in (sqrt(1/2), sqrt 2],   s = (f - 1) / (f + 1),  z = s^2
//     ln f = 2 s + 2 s z (1/3 + z (1/5 + z (1/7 + z (1/9 + z (1/11 + z (1/13 + z / 15))))))     (|error| < 3e-14)
//     -ln(1 - u) = (53 - p) ln 2 - ln f
#pragma once

#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define ASP_SP_HD __host__ __device__ __forceinline__
#else
#define ASP_SP_HD inline
#endif

namespace asp {

constexpr int kSpacingBits = 26;
constexpr int kSpacingTile = 2048;  // slots per tile of the integer scan (and CDF entries per search window)

ASP_SP_HD double sp_mul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    volatile double r = a * b;
    return r;
#endif
}
ASP_SP_HD double sp_add(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    volatile double r = a + b;
    return r;
#endif
}
ASP_SP_HD double sp_div(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __ddiv_rn(a, b);
#else
    volatile double r = a / b;
    return r;
#endif
}
ASP_SP_HD double sp_pow2(int e) {  // 2^e, |e| < 1023
    const uint64_t bits = (uint64_t)(1023 + e) << 52;
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)bits);
#else
    double r;
    memcpy(&r, &bits, sizeof r);
    return r;
#endif
}
ASP_SP_HD int sp_floor_log2(uint64_t m) {  // m >= 1
#if defined(__CUDA_ARCH__)
    return 63 - __clzll((long long)m);
#else
    return 63 - __builtin_clzll(m);
#endif
}

// k: the 53-bit integer of the uniform (u = k 2^-53)  ->  floor(-ln(1 - u) 2^26)
ASP_SP_HD uint32_t spacing_q(uint64_t k) {
    const uint64_t m = (1ull << 53) - k;  // [1, 2^53]
    int p = sp_floor_log2(m);
    double f = sp_mul((double)(long long)m, sp_pow2(-p));  // exact: [1, 2)
    if (f > 1.4142135623730951) {
        f = sp_mul(f, 0.5);
        p += 1;
    }
    const double s = sp_div(sp_add(f, -1.0), sp_add(f, 1.0));
    const double z = sp_mul(s, s);
    double poly = 1.0 / 15.0;
    poly = sp_add(sp_mul(poly, z), 1.0 / 13.0);
    poly = sp_add(sp_mul(poly, z), 1.0 / 11.0);
    poly = sp_add(sp_mul(poly, z), 1.0 / 9.0);
    poly = sp_add(sp_mul(poly, z), 1.0 / 7.0);
    poly = sp_add(sp_mul(poly, z), 1.0 / 5.0);
    poly = sp_add(sp_mul(poly, z), 1.0 / 3.0);
    const double s2 = sp_mul(2.0, s);
    const double lnf = sp_add(sp_mul(sp_mul(s2, z), poly), s2);
    const double e = sp_add(sp_mul((double)(53 - p), 0.6931471805599453), -lnf);
    const double scaled = sp_mul(e, 67108864.0);  // 2^26
    return scaled > 0.0 ? (uint32_t)(unsigned long long)scaled : 0u;  // truncation == floor for positive values
}

// threshold of CDF value c against the integer partial sums: rint(c * total)
ASP_SP_HD uint64_t spacing_threshold(double c, double total) {
#if defined(__CUDA_ARCH__)
    return __double2ull_rn(__dmul_rn(c, total));
#else
    const double r = sp_mul(c, total);
    // round half to even, as cvt.rni / numpy.rint: r < 2^63
    const double fl = (double)(unsigned long long)r;
    const double d = r - fl;
    unsigned long long v = (unsigned long long)r;
    if (d > 0.5 || (d == 0.5 && (v & 1ull))) v += 1;
    return v;
#endif
}

}  // namespace asp

// Exact integer arithmetic of the sequential-order-exact prefix sum (csrc/scan_exact.cu; algorithm in
// oracle/exact_scan.py and SURVEY App. E), host + device: the CPU test suite compiles this header for the host
// (tests/native/scan_host.cu) and rebuilds numpy.cumsum from it, bit for bit.
//
// While the running sum s stays in binade e it is m * 2^(e-52) with an integer m in [2^52, 2^53); adding w is
// m += round_half_even(w / 2^(e-52)), and the tie-break depends only on the parity of m: an element is the pair
// (increment if m is even, increment if m is odd) and pairs compose associatively.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define ASP_SCAN_HD __host__ __device__ __forceinline__
#else
#define ASP_SCAN_HD inline
#endif

namespace asp {

constexpr int kEMin = -1022;
constexpr long long kTwo53 = 1LL << 53;
constexpr long long kSatMax = 1LL << 60;

struct Pair {
    long long d0, d1;  // increment if incoming mantissa is even / odd
};

ASP_SCAN_HD long long bits_of(double v) {
#if defined(__CUDA_ARCH__)
    return __double_as_longlong(v);
#else
    long long b;
    memcpy(&b, &v, sizeof b);
    return b;
#endif
}
ASP_SCAN_HD double double_of(long long b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(b);
#else
    double v;
    memcpy(&v, &b, sizeof v);
    return v;
#endif
}
ASP_SCAN_HD long long floor_ll(double q) {
#if defined(__CUDA_ARCH__)
    return __double2ll_rd(q);
#else
    return (long long)floor(q);
#endif
}

ASP_SCAN_HD int binade_of(double v) {
    if (!(v > 0.0)) return kEMin;
    int be = (int)((bits_of(v) >> 52) & 0x7ff);
    return be == 0 ? kEMin : be - 1023;
}
// x * 2^k, exact whenever the true result is representable
ASP_SCAN_HD double scale_pow2(double x, int k) {
    if (k >= -1022 && k <= 1023) return x * double_of((long long)(k + 1023) << 52);
    return ldexp(x, k);
}
ASP_SCAN_HD Pair element_pair(double w, int e) {
    double q = scale_pow2(w, 52 - e);
    Pair r;
    if (!(q < 9007199254740992.0)) {  // >= 2^53, inf or NaN: forces a crossing
        r.d0 = r.d1 = kTwo53;
        return r;
    }
    long long n = floor_ll(q);
    double frac = q - (double)n;
    long long up = frac > 0.5 ? 1 : 0;
    bool tie = frac == 0.5;
    long long odd = n & 1;
    r.d0 = n + up + ((tie && odd) ? 1 : 0);
    r.d1 = n + up + ((tie && !odd) ? 1 : 0);
    return r;
}
ASP_SCAN_HD long long sat(long long a) { return a < kSatMax ? a : kSatMax; }
// a then b
ASP_SCAN_HD Pair compose(Pair a, Pair b) {
    Pair h;
    h.d0 = sat(a.d0 + ((a.d0 & 1) ? b.d1 : b.d0));
    h.d1 = sat(a.d1 + ((a.d1 & 1) ? b.d0 : b.d1));
    return h;
}
ASP_SCAN_HD long long apply(Pair a, long long m0) { return m0 + ((m0 & 1) ? a.d1 : a.d0); }

}  // namespace asp

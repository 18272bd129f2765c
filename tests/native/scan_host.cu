// Test harness (not part of the product library): rebuilds a sequential fp64 cumsum from the SHIPPED integer-pair
// arithmetic (aspire_b200/csrc/scan_math.cuh) the way the kernels use it - per tile, the elements become pairs under
// the binade of the running sum, pairs are composed by a balanced tree (any association must give the same result),
// applied to the running mantissa, and a binade crossing is resolved by one genuine fp64 add - so the CPU suite
// can compare it with numpy.cumsum bit for bit.  Built by tests/test_scan_math_host.py with nvcc (host code only).
#include <stdint.h>

#include <vector>

#include "../../aspire_b200/csrc/scan_math.cuh"

using asp::Pair;

// inclusive scan by recursive halving: a different association than left-to-right on purpose
static void tree_scan(std::vector<Pair>& f, int64_t lo, int64_t hi) {
    if (hi - lo <= 1) return;
    const int64_t mid = lo + (hi - lo) / 2;
    tree_scan(f, lo, mid);
    tree_scan(f, mid, hi);
    const Pair left = f[mid - 1];
    for (int64_t i = mid; i < hi; ++i) f[i] = asp::compose(left, f[i]);
}

extern "C" int scan_host(const double* w, int64_t n, int64_t tile, double* out, int64_t* n_phases) {
    double s = 0.0;
    int64_t phases = 0;
    std::vector<Pair> f;
    for (int64_t t0 = 0; t0 < n; t0 += tile) {
        const int64_t t1 = t0 + tile < n ? t0 + tile : n;
        int64_t pos = t0;
        while (pos < t1) {
            ++phases;
            const int e = asp::binade_of(s);
            const long long m0 = (long long)asp::scale_pow2(s, 52 - e);  // exact: s is a multiple of 2^(e-52)
            f.resize(t1 - pos);
            for (int64_t i = pos; i < t1; ++i) f[i - pos] = asp::element_pair(w[i], e);
            tree_scan(f, 0, t1 - pos);
            int64_t j = 0;
            for (; j < t1 - pos; ++j) {
                const long long m = asp::apply(f[j], m0);
                if (m >= asp::kTwo53) break;  // the running sum left the binade at element pos + j
                out[pos + j] = asp::scale_pow2((double)m, e - 52);
            }
            if (j > 0) s = out[pos + j - 1];
            pos += j;
            if (pos < t1) {  // one genuine add across the crossing
                volatile double r = s + w[pos];
                s = r;
                out[pos] = s;
                ++pos;
            }
        }
    }
    if (n_phases) *n_phases = phases;
    return 0;
}

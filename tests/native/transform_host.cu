// Test harness (not part of the product library): runs the shipped per-element transform arithmetic
// (aspire_b200/csrc/transform_math.cuh) on the HOST, with the kernel's summation order for log |det J| (terms of
// a row added in dimension order), so the CPU test suite can check it against the reference's golden vectors.
// Built by tests/test_transform_math_host.py with nvcc (host code only).
#include <stdint.h>

#include "../../aspire_b200/csrc/transform_math.cuh"

extern "C" int transform_host(int inverse, int dims, int bounded, int affine, double eps, const int32_t* kind,
                              const double* lower, const double* upper, const double* mean, const double* sd,
                              const double* in, int64_t n, double* out, double* log_j) {
    if (dims < 1 || dims > 256) return -1;
    asp::TransformDim c[256];
    for (int j = 0; j < dims; ++j) {
        const int k = kind ? kind[j] : 0;
        c[j] = asp::transform_dim(k, k ? lower[j] : 0.0, k ? upper[j] : 1.0, affine != 0, affine ? mean[j] : 0.0,
                                  affine ? sd[j] : 1.0);
    }
    for (int64_t r = 0; r < n; ++r) {
        double acc = 0.0;
        for (int j = 0; j < dims; ++j) {
            double v = in[r * dims + j];
            acc += inverse ? asp::transform_element<true>(c[j], bounded, affine != 0, eps, v)
                           : asp::transform_element<false>(c[j], bounded, affine != 0, eps, v);
            out[r * dims + j] = v;
        }
        log_j[r] = acc;
    }
    return 0;
}

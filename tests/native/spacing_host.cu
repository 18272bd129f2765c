// Test harness (not part of the product library): the shipped spacing arithmetic of the sorted-uniform resampler
// (aspire_b200/csrc/spacing_math.cuh) compiled for the host, so the CPU suite can compare it with the numpy spec
// (oracle/sorted_resample.py) bit for bit.  Built by tests/test_sorted_resample_oracle.py with nvcc (host code only).
#include <stdint.h>

#include "../../aspire_b200/csrc/spacing_math.cuh"

extern "C" int spacing_host(const uint64_t* k, int64_t n, uint32_t* q) {
    for (int64_t i = 0; i < n; ++i) q[i] = asp::spacing_q(k[i]);
    return 0;
}

extern "C" int threshold_host(const double* c, int64_t n, double total, uint64_t* t) {
    for (int64_t i = 0; i < n; ++i) t[i] = asp::spacing_threshold(c[i], total);
    return 0;
}

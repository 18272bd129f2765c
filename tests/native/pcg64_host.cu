// Test harness (not part of the product library): asp_pcg64_uniform's thread scheme replayed on the host with the
// SHIPPED 128-bit arithmetic (aspire_b200/csrc/pcg64_math.cuh) - "thread" g jumps ahead to draw offset + g with
// the affine power of the LCG and then strides by the grid size T - so the CPU suite can compare the stream with
// numpy.random.Generator.random bit for bit.  Built by tests/test_pcg64_math_host.py with nvcc (host code only).
#include <stdint.h>

#include "../../aspire_b200/csrc/pcg64_math.cuh"

extern "C" int pcg64_host(uint64_t s_hi, uint64_t s_lo, uint64_t i_hi, uint64_t i_lo, int64_t offset, int64_t n,
                          int64_t grid_threads, double* out) {
    using asp::u128;
    const u128 inc = ((u128)i_hi << 64) | i_lo;
    const u128 state0 = ((u128)s_hi << 64) | s_lo;
    u128 a_t, c_t;
    asp::pcg_affine_power((uint64_t)grid_threads, inc, a_t, c_t);
    for (int64_t g = 0; g < grid_threads && g < n; ++g) {
        u128 a, c;
        asp::pcg_affine_power((uint64_t)(offset + g + 1), inc, a, c);
        u128 state = a * state0 + c;
        for (int64_t j = g; j < n; j += grid_threads) {
            out[j] = asp::pcg_double(state);
            state = a_t * state + c_t;
        }
    }
    return 0;
}

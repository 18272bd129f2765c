// K13: Philox4x32-10 fills (layouts in oracle/philox.py).  The reference draws
// from numpy's PCG64 on the host (src/aspire/samples.py:1278); these are the
// device-side, sharding-invariant replacements used in throughput mode.
#include <math.h>

#include "common.cuh"
#include "pcg64_math.cuh"
#include "philox.cuh"

namespace asp {

// element `index` -> Philox block index/2, words (0,1) or (2,3)
__global__ void __launch_bounds__(kBlock) philox_uniform_kernel(uint64_t seed, uint32_t stream, int64_t index0,
                                                                int64_t n, double* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    // one thread per Philox block (two outputs)
    const int64_t first_blk = index0 >> 1, last_blk = (index0 + n - 1) >> 1;
    for (int64_t b = first_blk + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b <= last_blk; b += stride) {
        const uint4 r = philox4x32_10((uint32_t)b, (uint32_t)((uint64_t)b >> 32), stream, 0u, (uint32_t)seed,
                                      (uint32_t)(seed >> 32));
        const int64_t e0 = 2 * b - index0, e1 = e0 + 1;
        if (e0 >= 0 && e0 < n) out[e0] = u53(r.x, r.y);
        if (e1 >= 0 && e1 < n) out[e1] = u53(r.z, r.w);
    }
}

// out[i, j]: dimension j of element i uses block j/2, component j%2 (fp64 Box-Muller)
__global__ void __launch_bounds__(kBlock) philox_normal_kernel(uint64_t seed, uint32_t stream, int64_t index0,
                                                               int64_t n, int dims, const double* __restrict__ mean,
                                                               const double* __restrict__ sigma,
                                                               double* __restrict__ out) {
    const int nblk = (dims + 1) / 2;
    const int64_t total = n * nblk;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += stride) {
        const int64_t i = w / nblk;
        const int b = (int)(w - i * nblk);
        const uint64_t idx = (uint64_t)(index0 + i);
        const uint4 r = philox4x32_10((uint32_t)idx, (uint32_t)(idx >> 32), stream, (uint32_t)b, (uint32_t)seed,
                                      (uint32_t)(seed >> 32));
        const double u1 = ((double)((((uint64_t)r.x << 32) | r.y) >> 11) + 0.5) * 0x1p-53;
        const double u2 = ((double)((((uint64_t)r.z << 32) | r.w) >> 11) + 0.5) * 0x1p-53;
        const double s = sqrt(-2.0 * log(u1));
        double sn, cs;
        sincos(2.0 * 3.141592653589793 * u2, &sn, &cs);
        const int j0 = 2 * b, j1 = j0 + 1;
        double z0 = s * cs, z1 = s * sn;
        if (mean || sigma) {
            z0 = (mean ? mean[j0] : 0.0) + (sigma ? sigma[j0] : 1.0) * z0;
            if (j1 < dims) z1 = (mean ? mean[j1] : 0.0) + (sigma ? sigma[j1] : 1.0) * z1;
        }
        out[i * dims + j0] = z0;
        if (j1 < dims) out[i * dims + j1] = z1;
    }
}

// ---- PCG64 (numpy's default bit generator) -------------------------------------
// Generator.random() of the reference's rng (src/aspire/samples.py:1278 -> Generator.choice ->
// Generator.random) reproduced on the device: draw j is the XSL-RR output of the 128-bit LCG state
// after j + 1 steps.  Thread g jumps to draw offset + g with the O(log) affine power of the LCG, then
// strides by the grid size with the affine map of T steps (computed once per block), so stores are
// coalesced.  Spec + pin against numpy: oracle/pcg64.py, tests/test_pcg64_oracle.py.
// The 128-bit arithmetic lives in pcg64_math.cuh (host + device: tests/native/pcg64_host.cu runs it on the CPU).
__global__ void __launch_bounds__(kBlock) pcg64_uniform_kernel(uint64_t s_hi, uint64_t s_lo, uint64_t i_hi,
                                                               uint64_t i_lo, int64_t offset, int64_t n,
                                                               double* __restrict__ out) {
    __shared__ u128 s_stride[2];
    const u128 inc = ((u128)i_hi << 64) | i_lo;
    const int64_t T = (int64_t)gridDim.x * blockDim.x;
    if (threadIdx.x == 0) pcg_affine_power((uint64_t)T, inc, s_stride[0], s_stride[1]);
    __syncthreads();
    const u128 a_t = s_stride[0], c_t = s_stride[1];
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    u128 a, c;
    pcg_affine_power((uint64_t)(offset + g + 1), inc, a, c);
    u128 state = a * (((u128)s_hi << 64) | s_lo) + c;
    for (int64_t j = g; j < n; j += T) {
        out[j] = pcg_double(state);
        state = a_t * state + c_t;
    }
}

}  // namespace asp

using namespace asp;

extern "C" {

int asp_pcg64_uniform(uint64_t state_hi, uint64_t state_lo, uint64_t inc_hi, uint64_t inc_lo, int64_t offset,
                      int64_t n, double* d_out, void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(d_out, "null pointer");
    ASP_REQUIRE(n > 0 && offset >= 0, "bad sizes");
    ASP_REQUIRE(inc_lo & 1, "PCG64 increment must be odd");
    pcg64_uniform_kernel<<<grid_for(n, 64), kBlock, 0, (cudaStream_t)stream>>>(state_hi, state_lo, inc_hi, inc_lo,
                                                                              offset, n, d_out);
    return check_launch("asp_pcg64_uniform");
}

int asp_philox_uniform(uint64_t seed, uint32_t stream_id, int64_t index0, int64_t n, double* d_out, void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(d_out, "null pointer");
    ASP_REQUIRE(n > 0 && index0 >= 0, "bad sizes");
    philox_uniform_kernel<<<grid_for(n / 2 + 1, 1), kBlock, 0, (cudaStream_t)stream>>>(seed, stream_id, index0, n,
                                                                                       d_out);
    return check_launch("asp_philox_uniform");
}

int asp_philox_normal(uint64_t seed, uint32_t stream_id, int64_t index0, int64_t n, int32_t dims,
                      const double* d_mean, const double* d_sigma, double* d_out, void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(d_out, "null pointer");
    ASP_REQUIRE(n > 0 && dims > 0 && index0 >= 0, "bad sizes");
    philox_normal_kernel<<<grid_for(n * ((dims + 1) / 2), 1), kBlock, 0, (cudaStream_t)stream>>>(
        seed, stream_id, index0, n, dims, d_mean, d_sigma, d_out);
    return check_launch("asp_philox_normal");
}

}  // extern "C"

// tpCN reference measure: EM fit of a multivariate Student-t with diagonal scale, t_nu(mu, diag sigma^2), to (a
// strided sample of) the resampled ensemble - one kernel per EM iteration, nothing leaves the device until the end.
//
// The reference's default step (step_fn = "tpcn", /root/reference/src/aspire/samplers/smc/minipcn.py:48) lives in the
// third-party minipcn package (absent: PARITY UNPINNED, see oracle/pcn.py); spec of this fit:
// oracle/pcn.py::fit_student_t_diag (Liu & Rubin 1995):
//   E step   w_i = (nu + D) / (nu + r2_i),   r2_i = |(x_i - mu) / sigma|^2
//   M step   mu = sum w x / sum w,  sigma^2 = sum w (x - mu_new)^2 / n,
//            nu solves  -psi(nu/2) + log(nu/2) + 1 + mean(log w - w) + psi((nu + D)/2) - log((nu + D)/2) = 0
// One pass over the rows accumulates  sum w, sum (log w - w), A_j = sum w (x_j - mu_j), B_j = sum w (x_j - mu_j)^2
// about the CURRENT mu; then mu_new = mu + A / sum w and sum w (x - mu_new)^2 = B - A^2 / sum w (centred at a point
// next to the new mean: no cancellation).  Per-block partials, summed in block order by the last block
// (deterministic), which also runs the M step and the bisection for nu (60 steps in log nu, as the spec).
#include <math.h>

#include "common.cuh"
#include "density.cuh"
#include "student_t_math.cuh"

namespace asp {

struct TFitWs {
    unsigned int ticket;
    unsigned int pad[63];
    double partial[1];  // [blocks][2 + 2 * G * kDPL]
};

template <int G>
__global__ void __launch_bounds__(kBlock) student_t_em_kernel(const double* __restrict__ x, int dims,
                                                              const int64_t* __restrict__ idx, int64_t stride, int64_t m,
                                                              double* __restrict__ state, double nu_min, double nu_max,
                                                              TFitWs* ws) {
    constexpr int kAcc = 2 + 2 * kDPL;
    extern __shared__ double sm[];  // [warps][G][kAcc]
    __shared__ bool is_last;
    const int lane = threadIdx.x % G;
    const int wlane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const bool vec = (dims % kDPL == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    const int64_t groups = ((int64_t)gridDim.x * blockDim.x) / G;
    const double nu = state[2 * dims];
    double mu[kDPL], sg[kDPL];
#pragma unroll
    for (int k = 0; k < kDPL; ++k) {
        const int j = lane * kDPL + k;
        mu[k] = j < dims ? state[j] : 0.0;
        sg[k] = j < dims ? state[dims + j] : 1.0;
    }
    double acc[kAcc];
#pragma unroll
    for (int k = 0; k < kAcc; ++k) acc[k] = 0.0;
    const int64_t g0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int64_t warp_first = ((int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31)) / G;
    for (int64_t base = warp_first; base < m; base += groups) {  // warp-uniform trip count (group shuffles inside)
        const int64_t i_raw = base + (g0 - warp_first);
        const bool active = i_raw < m;
        const int64_t i = active ? i_raw : m - 1;
        const int64_t src = idx ? idx[i * stride] : i * stride;
        double v[kDPL];
        row_load(x, src, dims, lane, vec, v);
        double r2 = 0.0;
#pragma unroll
        for (int k = 0; k < kDPL; ++k) {
            v[k] -= mu[k];
            const double z = v[k] / sg[k];
            r2 = fma(z, z, r2);
        }
        r2 = group_sum<G>(r2);
        const double w = active ? (nu + (double)dims) / (nu + r2) : 0.0;
        if (lane == 0 && active) {
            acc[0] += w;
            acc[1] += log(w) - w;
        }
#pragma unroll
        for (int k = 0; k < kDPL; ++k) {
            const double wv = w * v[k];
            acc[2 + k] += wv;
            acc[2 + kDPL + k] = fma(wv, v[k], acc[2 + kDPL + k]);
        }
    }
    // lanes of equal rank across the warp's groups -> lanes [0, G)
#pragma unroll
    for (int o = 16; o >= G; o >>= 1)
#pragma unroll
        for (int k = 0; k < kAcc; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    if (wlane < G)
#pragma unroll
        for (int k = 0; k < kAcc; ++k) sm[(wid * G + wlane) * kAcc + k] = acc[k];
    __syncthreads();
    const int width = 2 + 2 * G * kDPL;  // sum w, sum (log w - w), A[G * 4], B[G * 4]
    if (threadIdx.x < G) {
        double t[kAcc];
#pragma unroll
        for (int k = 0; k < kAcc; ++k) t[k] = 0.0;
        for (int w = 0; w < nw; ++w)
#pragma unroll
            for (int k = 0; k < kAcc; ++k) t[k] += sm[(w * G + threadIdx.x) * kAcc + k];
        double* dst = ws->partial + (size_t)blockIdx.x * width;
        if (threadIdx.x == 0) {
            dst[0] = t[0];
            dst[1] = t[1];
        }
#pragma unroll
        for (int k = 0; k < kDPL; ++k) {
            dst[2 + threadIdx.x * kDPL + k] = t[2 + k];
            dst[2 + G * kDPL + threadIdx.x * kDPL + k] = t[2 + kDPL + k];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        is_last = (atomicAdd(&ws->ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // ---- last block: totals in block order, M step, nu ----
    double* tot = sm;  // [width]
    for (int col = threadIdx.x; col < width; col += blockDim.x) {
        double s = 0.0;
        for (unsigned int b = 0; b < gridDim.x; ++b) s += __ldcg(ws->partial + (size_t)b * width + col);
        tot[col] = s;
    }
    __syncthreads();
    const double sw = tot[0];
    for (int j = threadIdx.x; j < dims; j += blockDim.x) {
        double mu_j = state[j], sg_j;
        student_t_mstep(sw, tot[2 + j], tot[2 + G * kDPL + j], (double)m, mu_j, sg_j);
        state[j] = mu_j;
        state[dims + j] = sg_j;
    }
    if (threadIdx.x == 0) {
        state[2 * dims] = student_t_nu(nu, dims, tot[1] / (double)m, nu_min, nu_max);
        ws->ticket = 0;
    }
}

}  // namespace asp

using namespace asp;

extern "C" {

size_t asp_student_t_workspace_bytes(int32_t dims) {
    const int g = group_size_for(dims > 0 ? dims : 1);
    return 256 + sizeof(double) * (size_t)kNumSMs * 4 * (2 + 2 * g * kDPL);
}

int asp_student_t_fit(const double* d_x, int32_t dims, const int64_t* d_idx, int64_t stride, int64_t m, int32_t n_iter,
                      double nu_min, double nu_max, double* d_state, void* ws, void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(d_x && d_state && ws, "null pointer");
    ASP_REQUIRE(m > 0 && stride >= 1 && dims > 0 && dims <= 32 * kDPL, "bad sizes");
    ASP_REQUIRE(n_iter >= 0 && n_iter <= 1000, "n_iter must be in [0, 1000]");
    ASP_REQUIRE(nu_min > 0.0 && nu_max >= nu_min, "bad nu bracket");
    const int g = group_size_for(dims);
    const int grid = grid_for(m * g, 1, kBlock, 4);
    const int width = 2 + 2 * g * kDPL;
    const size_t per_warp = sizeof(double) * g * (2 + 2 * kDPL);
    size_t smem = (size_t)(kBlock / 32) * per_warp;
    if (smem < sizeof(double) * width) smem = sizeof(double) * width;
    cudaStream_t s = (cudaStream_t)stream;
    for (int it = 0; it < n_iter; ++it) {
        switch (g) {
            case 1: student_t_em_kernel<1><<<grid, kBlock, smem, s>>>(d_x, dims, d_idx, stride, m, d_state, nu_min, nu_max, (TFitWs*)ws); break;
            case 2: student_t_em_kernel<2><<<grid, kBlock, smem, s>>>(d_x, dims, d_idx, stride, m, d_state, nu_min, nu_max, (TFitWs*)ws); break;
            case 4: student_t_em_kernel<4><<<grid, kBlock, smem, s>>>(d_x, dims, d_idx, stride, m, d_state, nu_min, nu_max, (TFitWs*)ws); break;
            case 8: student_t_em_kernel<8><<<grid, kBlock, smem, s>>>(d_x, dims, d_idx, stride, m, d_state, nu_min, nu_max, (TFitWs*)ws); break;
            case 16: student_t_em_kernel<16><<<grid, kBlock, smem, s>>>(d_x, dims, d_idx, stride, m, d_state, nu_min, nu_max, (TFitWs*)ws); break;
            default: student_t_em_kernel<32><<<grid, kBlock, smem, s>>>(d_x, dims, d_idx, stride, m, d_state, nu_min, nu_max, (TFitWs*)ws); break;
        }
    }
    return check_launch("asp_student_t_fit");
}

}  // extern "C"

// K1 log-density, K8 pCN mutation sweep, reference-measure moments.
//
// Replaces minipcn.Sampler.sample as driven by MiniPCNSMC.mutate
// (/root/reference/src/aspire/samplers/smc/minipcn.py:69-135) against the
// tempered target of src/aspire/samplers/smc/base.py:507-519
// (log pi = (1-beta) log_q + beta (logL + logP), NaN -> -inf).  Step spec:
// oracle/pcn.py.  The fused sweep reads each particle once (through the
// resampling index, so the gather of samples.py:1279-1287 is never
// materialised), runs all moves in registers and writes it once.  It is
// bound by fp64 density arithmetic + Philox/Box-Muller issue, not by HBM.
#include <math.h>

#include "common.cuh"
#include "density.cuh"
#include "pcn_sweep.cuh"
#include "pcn_sweep_tpp.cuh"
#include "philox.cuh"

#include <stdlib.h>

namespace asp {

// ---- unfused halves (user-supplied Python callables evaluate the densities) ----
template <int G>
__global__ void __launch_bounds__(kBlock) pcn_propose_kernel(const double* __restrict__ xin, int64_t n, int dims,
                                                             const asp_pcn_params_t prm, int move,
                                                             double* __restrict__ xout) {
    const int lane = threadIdx.x % G;
    const double rho = prm.rho, ca = sqrt(1.0 - rho * rho);
    double mu[kDPL], rs[kDPL];
#pragma unroll
    for (int k = 0; k < kDPL; ++k) {
        const int j = lane * kDPL + k;
        mu[k] = j < dims ? __ldg(prm.d_mu + j) : 0.0;
        rs[k] = j < dims ? rho * __ldg(prm.d_sigma + j) : 0.0;
    }
    const bool vec = (dims % kDPL == 0) && ((reinterpret_cast<uintptr_t>(xin) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(xout) & 15) == 0);
    const bool vec_xi = (dims % kDPL == 0) && ((reinterpret_cast<uintptr_t>(prm.d_xi) & 15) == 0);
    const int64_t groups = ((int64_t)gridDim.x * blockDim.x) / G;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G; i < n; i += groups) {
        double x[kDPL], xi[kDPL], xp[kDPL];
        row_load(xin, i, dims, lane, vec, x);
        if (prm.noise_mode == 0) {
            float z[4];
            philox_normal4(prm.seed, prm.stream0 + (uint32_t)move, (uint64_t)(prm.d_slot ? prm.d_slot[i] : prm.index0 + i),
                           (uint32_t)lane, z);
#pragma unroll
            for (int k = 0; k < kDPL; ++k) xi[k] = (double)z[k];
        } else {
            row_load(prm.d_xi + (size_t)move * n * dims, i, dims, lane, vec_xi, xi);
        }
#pragma unroll
        const double sc = prm.d_scale ? prm.d_scale[i] : 1.0;  // tpCN: sqrt(s_i)
#pragma unroll
        for (int k = 0; k < kDPL; ++k)
            xp[k] = mu[k] + (ca * (x[k] - mu[k]) + (prm.d_scale ? rs[k] * sc : rs[k]) * xi[k]);
        row_store(xout, i, dims, lane, vec, xp);
    }
}

template <int G>
__global__ void __launch_bounds__(kBlock) pcn_accept_kernel(double* __restrict__ x, double* __restrict__ ll,
                                                            double* __restrict__ lp, double* __restrict__ lq,
                                                            const double* __restrict__ xp_,
                                                            const double* __restrict__ ll2_,
                                                            const double* __restrict__ lp2_,
                                                            const double* __restrict__ lq2_, int64_t n, int dims,
                                                            const asp_pcn_params_t prm, int move,
                                                            unsigned long long* accepts) {
    __shared__ unsigned int s_acc;
    if (threadIdx.x == 0) s_acc = 0;
    __syncthreads();
    const int lane = threadIdx.x % G;
    double mu[kDPL], isig[kDPL];
#pragma unroll
    for (int k = 0; k < kDPL; ++k) {
        const int j = lane * kDPL + k;
        mu[k] = j < dims ? __ldg(prm.d_mu + j) : 0.0;
        isig[k] = j < dims ? 1.0 / __ldg(prm.d_sigma + j) : 0.0;
    }
    const bool vec = (dims % kDPL == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(xp_) & 15) == 0);
    const int64_t groups = ((int64_t)gridDim.x * blockDim.x) / G;
    const int64_t g0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int64_t warp_first = ((int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31)) / G;
    for (int64_t base = warp_first; base < n; base += groups) {
        const int64_t i_raw = base + (g0 - warp_first);
        const bool active = i_raw < n;
        const int64_t i = active ? i_raw : n - 1;
        double xc[kDPL], xn[kDPL];
        row_load(x, i, dims, lane, vec, xc);
        row_load(xp_, i, dims, lane, vec, xn);
        const double cur = tempered(ll[i], lp[i], lq[i], prm.beta);
        const double nw = tempered(ll2_[i], lp2_[i], lq2_[i], prm.beta);
        const double rx = ref_quad<G>(xc, mu, isig);
        const double rx2 = ref_quad<G>(xn, mu, isig);
        double log_u;
        if (prm.noise_mode == 0) {
            log_u = accept_log_u(accept_words(prm.seed, prm.stream0, move,
                                              (uint64_t)(prm.d_slot ? prm.d_slot[i] : prm.index0 + i)), move);
        } else {
            log_u = prm.d_log_u[(size_t)move * n + i];
        }
        // reference-measure term: Gaussian 1/2 (r2' - r2), or Student-t (nu + D)/2 [log1p(r2'/nu) - log1p(r2/nu)]
        const double dref = prm.nu > 0.0 ? 0.5 * (prm.nu + dims) * (log1p(2.0 * rx2 / prm.nu) - log1p(2.0 * rx / prm.nu))
                                         : rx2 - rx;
        const bool acc = log_u < (nw - cur) + dref;
        // every lane of the group has finished reading row i before any lane overwrites it
        __syncwarp();
        if (acc && active) {
            row_store(x, i, dims, lane, vec, xn);
            if (lane == 0) {
                ll[i] = ll2_[i];
                lp[i] = lp2_[i];
                lq[i] = lq2_[i];
            }
        }
        const unsigned int votes = __ballot_sync(0xffffffffu, acc && active && lane == 0);
        if ((threadIdx.x & 31) == 0 && votes) atomicAdd(&s_acc, __popc(votes));
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_acc) atomicAdd(&accepts[move], (unsigned long long)s_acc);
}

// ---- tpCN proposal scales: sqrt((nu + r2(x)) / (2 g)), g ~ Gamma((D + nu)/2, 1) ----
constexpr uint32_t kGammaBlock0 = 0xFFFFFF00u;

// Marsaglia & Tsang (2000), fp32, counter-based (spec: oracle/pcn.py::gamma_mt_f32)
__device__ __forceinline__ double gamma_mt(uint64_t seed, uint32_t stream, uint64_t index, float a) {
    const float d = a - (1.0f / 3.0f);
    const float c = 1.0f / sqrtf(9.0f * d);
    for (uint32_t it = 0; it < 64; ++it) {
        const uint4 r = philox4x32<kNoiseRounds>((uint32_t)index, (uint32_t)(index >> 32), stream, kGammaBlock0 + it,
                                                 (uint32_t)seed, (uint32_t)(seed >> 32));
        double g;
        if (gamma_mt_try(r, d, c, g)) return g;
    }
    return (double)d;
}

template <int G>
__global__ void __launch_bounds__(kBlock) tpcn_scales_kernel(const double* __restrict__ x, int64_t n, int dims,
                                                             const asp_pcn_params_t prm, int move,
                                                             double* __restrict__ scale) {
    const int lane = threadIdx.x % G;
    double mu[kDPL], isig[kDPL];
#pragma unroll
    for (int k = 0; k < kDPL; ++k) {
        const int j = lane * kDPL + k;
        mu[k] = j < dims ? __ldg(prm.d_mu + j) : 0.0;
        isig[k] = j < dims ? 1.0 / __ldg(prm.d_sigma + j) : 0.0;
    }
    const bool vec = (dims % kDPL == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    const int64_t groups = ((int64_t)gridDim.x * blockDim.x) / G;
    const int64_t g0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int64_t warp_first = ((int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31)) / G;
    const float a = (float)(0.5 * (prm.nu + dims));
    for (int64_t base = warp_first; base < n; base += groups) {  // warp-uniform trip count
        const int64_t i_raw = base + (g0 - warp_first);
        const bool active = i_raw < n;
        const int64_t i = active ? i_raw : n - 1;
        double xc[kDPL];
        row_load(x, i, dims, lane, vec, xc);
        const double r2 = 2.0 * ref_quad<G>(xc, mu, isig);
        if (active && lane == 0) {
            const double g = prm.noise_mode == 0
                                 ? gamma_mt(prm.seed, prm.stream0 + (uint32_t)move,
                                            (uint64_t)(prm.d_slot ? prm.d_slot[i] : prm.index0 + i), a)
                                 : prm.d_gamma[(size_t)move * n + i];
            scale[i] = sqrt((prm.nu + r2) / (2.0 * g));
        }
    }
}

// ---- K1: standalone log-density ----
template <int G, int KIND>
__global__ void __launch_bounds__(kBlock) logdensity_kernel(const asp_density_t D, const double* __restrict__ x,
                                                            int64_t n, double* __restrict__ out) {
    const int dims = D.dims;
    const int lane = threadIdx.x % G;
    Dens<KIND> L;
    L.load(D, lane, dims);
    const bool vec = (dims % kDPL == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    const int64_t groups = ((int64_t)gridDim.x * blockDim.x) / G;
    const int64_t g0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int64_t warp_first = ((int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31)) / G;
    for (int64_t base = warp_first; base < n; base += groups) {
        const int64_t i_raw = base + (g0 - warp_first);
        const bool active = i_raw < n;
        const int64_t i = active ? i_raw : n - 1;
        double v[kDPL];
        row_load(x, i, dims, lane, vec, v);
        const double r = L.template eval<G>(v, lane, dims);
        if (active && lane == 0) out[i] = r;
    }
}

// ---- reference-measure moments: sum w, sum w x_j, sum w x_j^2 ----
struct MomentsWs {
    unsigned int ticket;
    unsigned int pad[63];
    double partial[1];  // [blocks][1 + 2 * dims_padded]
};

template <int G>
__global__ void __launch_bounds__(kBlock) moments_kernel(const double* __restrict__ x, int dims,
                                                         const int64_t* __restrict__ idx, int64_t istride,
                                                         const double* __restrict__ d, int64_t n, double dbeta,
                                                         double c, double* __restrict__ out, MomentsWs* ws) {
    extern __shared__ double sm[];  // [warps][G][9]
    __shared__ bool is_last;
    const int lane = threadIdx.x % G;
    const int wlane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const bool vec = (dims % kDPL == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    const int64_t groups = ((int64_t)gridDim.x * blockDim.x) / G;
    double acc[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] = 0.0;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G; i < n; i += groups) {
        double v[kDPL];
        const int64_t src = idx ? idx[i * istride] : i * istride;
        row_load(x, src, dims, lane, vec, v);
        const double w = d ? exp(dbeta * d[src] - c) : 1.0;
        acc[0] += w;
#pragma unroll
        for (int k = 0; k < kDPL; ++k) {
            acc[1 + k] += w * v[k];
            acc[5 + k] += w * (v[k] * v[k]);
        }
    }
    // groups of one warp -> lanes [0, G)
#pragma unroll
    for (int o = 16; o >= G; o >>= 1)
#pragma unroll
        for (int k = 0; k < 9; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    if (wlane < G)
#pragma unroll
        for (int k = 0; k < 9; ++k) sm[(wid * G + wlane) * 9 + k] = acc[k];
    __syncthreads();
    const int width = 1 + 2 * G * kDPL;  // padded row: sum w, sum wx[G*4], sum wx2[G*4]
    if (threadIdx.x < G) {
        double t[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) t[k] = 0.0;
        for (int w = 0; w < nw; ++w)
#pragma unroll
            for (int k = 0; k < 9; ++k) t[k] += sm[(w * G + threadIdx.x) * 9 + k];
        double* dst = ws->partial + (size_t)blockIdx.x * width;
        if (threadIdx.x == 0) dst[0] = t[0];
#pragma unroll
        for (int k = 0; k < kDPL; ++k) {
            dst[1 + threadIdx.x * kDPL + k] = t[1 + k];
            dst[1 + G * kDPL + threadIdx.x * kDPL + k] = t[5 + k];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        is_last = (atomicAdd(&ws->ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        for (int col = threadIdx.x; col < width; col += blockDim.x) {
            double s = 0.0;
            for (unsigned int b = 0; b < gridDim.x; ++b) s += __ldcg(ws->partial + (size_t)b * width + col);
            // un-pad: col 0 -> out[0]; wx col j -> out[1 + j]; wx2 col j -> out[1 + dims + j]
            if (col == 0) out[0] = s;
            else if (col < 1 + G * kDPL) { if (col - 1 < dims) out[col] = s; }
            else { const int j = col - 1 - G * kDPL; if (j < dims) out[1 + dims + j] = s; }
        }
        if (threadIdx.x == 0) ws->ticket = 0;
    }
}

// mean / std of the reference measure from the moment sums {sum w, sum w x_j, sum w x_j^2}
__global__ void moments_finish_kernel(const double* __restrict__ m, int dims, double floor, double* __restrict__ mu,
                                      double* __restrict__ sigma) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= dims) return;
    const double sw = m[0];
    const double mean = m[1 + j] / sw;
    const double var = m[1 + dims + j] / sw - mean * mean;  // one-pass form, as oracle/pcn.py::weighted_moments
    mu[j] = mean;
    sigma[j] = fmax(sqrt(fmax(var, 0.0)), floor);
}

}  // namespace asp

using namespace asp;

#define ASP_DISPATCH_G(g, CALL)   \
    switch (g) {                  \
        case 1: { constexpr int G = 1; CALL; } break;   \
        case 2: { constexpr int G = 2; CALL; } break;   \
        case 4: { constexpr int G = 4; CALL; } break;   \
        case 8: { constexpr int G = 8; CALL; } break;   \
        case 16: { constexpr int G = 16; CALL; } break; \
        default: { constexpr int G = 32; CALL; } break; \
    }

static int check_density(const asp_density_t* d, int dims, const char* name) {
    if (!d) { set_error("%s density is null", name); return ASP_ERR_ARG; }
    if (d->dims != dims) { set_error("%s density dims %d != %d", name, d->dims, dims); return ASP_ERR_ARG; }
    switch (d->kind) {
        case ASP_DENSITY_GAUSSIAN:
        case ASP_DENSITY_UNIFORM:
            if (!d->p0 || !d->p1) { set_error("%s density: missing parameter arrays", name); return ASP_ERR_ARG; }
            break;
        case ASP_DENSITY_ROSENBROCK:
            if (!(d->s2 != 0.0)) { set_error("%s density: zero scale", name); return ASP_ERR_ARG; }
            break;
        case ASP_DENSITY_GMM:
            if (!d->p0 || !d->p1 || !d->p2 || d->n_comp < 1 || d->n_comp > 32) {
                set_error("%s density: bad mixture parameters", name);
                return ASP_ERR_ARG;
            }
            break;
        default:
            set_error("%s density: unknown kind %d", name, d->kind);
            return ASP_ERR_ARG;
    }
    return ASP_OK;
}

extern "C" {

int asp_logdensity(const asp_density_t* dens, const double* d_x, int64_t n, double* d_out, void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(dens && d_x && d_out, "null pointer");
    ASP_REQUIRE(n > 0, "n must be positive");
    ASP_REQUIRE(dens->dims > 0 && dens->dims <= 32 * kDPL, "dims must be in [1, 128]");
    int rc = check_density(dens, dens->dims, "log");
    if (rc) return rc;
    const int g = group_size_for(dens->dims);
    const int grid = grid_for(n * g, 1);
    cudaStream_t s = (cudaStream_t)stream;
    switch (dens->kind) {
        case ASP_DENSITY_GAUSSIAN:
            ASP_DISPATCH_G(g, (logdensity_kernel<G, ASP_DENSITY_GAUSSIAN><<<grid, kBlock, 0, s>>>(*dens, d_x, n, d_out)));
            break;
        case ASP_DENSITY_ROSENBROCK:
            ASP_DISPATCH_G(g, (logdensity_kernel<G, ASP_DENSITY_ROSENBROCK><<<grid, kBlock, 0, s>>>(*dens, d_x, n, d_out)));
            break;
        case ASP_DENSITY_GMM:
            ASP_DISPATCH_G(g, (logdensity_kernel<G, ASP_DENSITY_GMM><<<grid, kBlock, 0, s>>>(*dens, d_x, n, d_out)));
            break;
        default:
            ASP_DISPATCH_G(g, (logdensity_kernel<G, ASP_DENSITY_UNIFORM><<<grid, kBlock, 0, s>>>(*dens, d_x, n, d_out)));
            break;
    }
    return check_launch("asp_logdensity");
}

int asp_pcn_sweep(const double* d_x, const double* d_log_l, const double* d_log_p, const double* d_log_q,
                  const int64_t* d_idx, int64_t n, int32_t dims, const asp_density_t* like,
                  const asp_density_t* prior, const asp_density_t* q, const asp_pcn_params_t* params,
                  double* d_x_out, double* d_log_l_out, double* d_log_p_out, double* d_log_q_out,
                  unsigned long long* d_accepts, void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(d_x && d_log_l && d_log_p && d_log_q && params && d_x_out && d_log_l_out && d_log_p_out &&
                    d_log_q_out && d_accepts, "null pointer");
    ASP_REQUIRE(n > 0, "n must be positive");
    ASP_REQUIRE(dims > 0 && dims <= 32 * kDPL, "dims must be in [1, 128]");
    ASP_REQUIRE(params->n_moves >= 0 && params->n_moves <= kMaxMovesSmem, "n_moves must be in [0, 1024]");
    ASP_REQUIRE(params->rho > 0.0 && params->rho <= 1.0, "rho must be in (0, 1]");
    ASP_REQUIRE(params->d_mu && params->d_sigma, "reference measure missing");
    ASP_REQUIRE(params->noise_mode == 0 || (params->d_xi && params->d_log_u), "noise arrays missing");
    ASP_REQUIRE(d_x != d_x_out || !d_idx, "in-place sweep cannot gather");
    const bool tpcn = params->nu > 0.0;
    ASP_REQUIRE(!tpcn || params->noise_mode == 0 || params->d_gamma, "gamma variates missing (tpCN, supplied noise)");
    int rc;
    if ((rc = check_density(like, dims, "likelihood"))) return rc;
    if ((rc = check_density(prior, dims, "prior"))) return rc;
    if ((rc = check_density(q, dims, "proposal"))) return rc;
    SweepArgs A;
    A.x = d_x; A.ll = d_log_l; A.lp = d_log_p; A.lq = d_log_q; A.idx = d_idx; A.n = n; A.dims = dims;
    A.like = *like; A.prior = *prior; A.q = *q; A.prm = *params;
    A.keys = philox_keys(params->seed);
    A.xo = d_x_out; A.llo = d_log_l_out; A.lpo = d_log_p_out; A.lqo = d_log_q_out; A.accepts = d_accepts;
    ASP_REQUIRE(q->kind == ASP_DENSITY_GAUSSIAN, "the fused sweep needs a diagonal-Gaussian proposal density");
    ASP_REQUIRE(prior->kind == ASP_DENSITY_GAUSSIAN || prior->kind == ASP_DENSITY_UNIFORM,
                "the fused sweep needs a Gaussian or uniform prior");
    cudaStream_t s = (cudaStream_t)stream;
    // fast path: particle-per-thread kernel (pcn_sweep_tpp.cuh) for the dims it is built for;
    // ASP_SWEEP_LEGACY=1 forces the general lanes-per-particle kernel (A/B checks)
    int dl = 0, gt = 0;
    const bool aligned = ((reinterpret_cast<uintptr_t>(d_x) | reinterpret_cast<uintptr_t>(d_x_out) |
                           reinterpret_cast<uintptr_t>(params->d_xi)) & 15) == 0;
    static const bool legacy = [] { const char* e = getenv("ASP_SWEEP_LEGACY"); return e && e[0] == '1'; }();
    if ((!legacy || tpcn) && aligned && params->rho < 1.0 && tpp_shape(dims, like->kind, dl, gt)) {
        switch (like->kind) {
            case ASP_DENSITY_GAUSSIAN: rc = tpcn ? launch_tpp_t_gaussian(A, s) : launch_tpp_gaussian(A, s); break;
            case ASP_DENSITY_ROSENBROCK: rc = tpcn ? launch_tpp_t_rosenbrock(A, s) : launch_tpp_rosenbrock(A, s); break;
            case ASP_DENSITY_GMM: rc = tpcn ? launch_tpp_t_gmm(A, s) : launch_tpp_gmm(A, s); break;
            default: rc = tpcn ? launch_tpp_t_uniform(A, s) : launch_tpp_uniform(A, s); break;
        }
        if (rc) { set_error("asp_pcn_sweep: unsupported density combination"); return rc; }
        return check_launch("asp_pcn_sweep");
    }
    // the Student-t step is built into the particle-per-thread kernel only (asp_pcn_sweep_tpcn_supported)
    ASP_REQUIRE(!tpcn, "fused tpCN needs dims in {4, 8, 16, 32, 64}, 16-byte aligned rows and rho < 1: use asp_tpcn_scales + the unfused pair");
    const int g = group_size_for(dims);
    const int grid = grid_for(n * g, 1);
    switch (like->kind) {
        case ASP_DENSITY_GAUSSIAN: rc = launch_sweep_gaussian(A, g, grid, s); break;
        case ASP_DENSITY_ROSENBROCK: rc = launch_sweep_rosenbrock(A, g, grid, s); break;
        case ASP_DENSITY_GMM: rc = launch_sweep_gmm(A, g, grid, s); break;
        default: rc = launch_sweep_uniform(A, g, grid, s); break;
    }
    if (rc) { set_error("asp_pcn_sweep: unsupported density combination"); return rc; }
    return check_launch("asp_pcn_sweep");
}

int32_t asp_pcn_sweep_tpcn_supported(int32_t dims) {
    int dl = 0, g = 0;
    return tpp_shape(dims, ASP_DENSITY_GAUSSIAN, dl, g) ? 1 : 0;
}

int asp_tpcn_scales(const double* d_x, int64_t n, int32_t dims, const asp_pcn_params_t* params, int32_t move,
                    double* d_scale, void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(d_x && params && d_scale, "null pointer");
    ASP_REQUIRE(n > 0 && dims > 0 && dims <= 32 * kDPL, "bad sizes");
    ASP_REQUIRE(params->nu > 0.0, "nu must be positive");
    ASP_REQUIRE(params->d_mu && params->d_sigma, "reference measure missing");
    ASP_REQUIRE(params->noise_mode == 0 || params->d_gamma, "gamma variates missing");
    const int g = group_size_for(dims);
    const int grid = grid_for(n * g, 1);
    ASP_DISPATCH_G(g, (tpcn_scales_kernel<G><<<grid, kBlock, 0, (cudaStream_t)stream>>>(d_x, n, dims, *params, move,
                                                                                       d_scale)));
    return check_launch("asp_tpcn_scales");
}

int asp_pcn_propose(const double* d_x, int64_t n, int32_t dims, const asp_pcn_params_t* params, int32_t move,
                    double* d_x_prop, void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(d_x && params && d_x_prop, "null pointer");
    ASP_REQUIRE(n > 0 && dims > 0 && dims <= 32 * kDPL, "bad sizes");
    ASP_REQUIRE(params->d_mu && params->d_sigma, "reference measure missing");
    ASP_REQUIRE(params->noise_mode == 0 || params->d_xi, "noise array missing");
    const int g = group_size_for(dims);
    const int grid = grid_for(n * g, 1);
    ASP_DISPATCH_G(g, (pcn_propose_kernel<G><<<grid, kBlock, 0, (cudaStream_t)stream>>>(d_x, n, dims, *params, move,
                                                                                       d_x_prop)));
    return check_launch("asp_pcn_propose");
}

int asp_pcn_accept(double* d_x, double* d_log_l, double* d_log_p, double* d_log_q, const double* d_x_prop,
                   const double* d_log_l_prop, const double* d_log_p_prop, const double* d_log_q_prop, int64_t n,
                   int32_t dims, const asp_pcn_params_t* params, int32_t move, unsigned long long* d_accepts,
                   void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(d_x && d_log_l && d_log_p && d_log_q && d_x_prop && d_log_l_prop && d_log_p_prop && d_log_q_prop &&
                    params && d_accepts, "null pointer");
    ASP_REQUIRE(n > 0 && dims > 0 && dims <= 32 * kDPL, "bad sizes");
    ASP_REQUIRE(params->d_mu && params->d_sigma, "reference measure missing");
    ASP_REQUIRE(params->noise_mode == 0 || params->d_log_u, "noise array missing");
    const int g = group_size_for(dims);
    const int grid = grid_for(n * g, 1);
    ASP_DISPATCH_G(g, (pcn_accept_kernel<G><<<grid, kBlock, 0, (cudaStream_t)stream>>>(
                          d_x, d_log_l, d_log_p, d_log_q, d_x_prop, d_log_l_prop, d_log_p_prop, d_log_q_prop, n, dims,
                          *params, move, d_accepts)));
    return check_launch("asp_pcn_accept");
}

size_t asp_moments_workspace_bytes(int32_t dims) {
    const int g = group_size_for(dims > 0 ? dims : 1);
    return 256 + sizeof(double) * (size_t)kNumSMs * 4 * (1 + 2 * g * kDPL);
}

int asp_weighted_moments(const double* d_x, int32_t dims, const int64_t* d_idx, int64_t idx_stride, const double* d_d,
                         int64_t n, double dbeta, double c, double* d_moments, void* ws, void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(d_x && d_moments && ws, "null pointer");
    ASP_REQUIRE(n > 0 && dims > 0 && dims <= 32 * kDPL && idx_stride >= 1, "bad sizes");
    const int g = group_size_for(dims);
    int grid = grid_for(n * g, 1, kBlock, 4);
    const size_t smem = sizeof(double) * (kBlock / 32) * g * 9;
    ASP_DISPATCH_G(g, (moments_kernel<G><<<grid, kBlock, smem, (cudaStream_t)stream>>>(d_x, dims, d_idx, idx_stride, d_d, n, dbeta, c,
                                                                                      d_moments, (MomentsWs*)ws)));
    return check_launch("asp_weighted_moments");
}

int asp_moments_finish(const double* d_moments, int32_t dims, double sigma_floor, double* d_mu, double* d_sigma,
                       void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(d_moments && d_mu && d_sigma, "null pointer");
    ASP_REQUIRE(dims > 0, "dims must be positive");
    moments_finish_kernel<<<(dims + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d_moments, dims, sigma_floor, d_mu, d_sigma);
    return check_launch("asp_moments_finish");
}

}  // extern "C"

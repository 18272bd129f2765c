// Semi-fused pCN move for proposals whose density is NOT a built-in (a normalising flow evaluated by
// its own kernel, config C4): per move
//   asp_pcn_propose_eval   x -> x', log P(x'), log L(x'), r(x') - r(x)      (one read of x, one write of x')
//   <caller evaluates log q(x') - e.g. asp_flow_logprob_bf16>
//   asp_pcn_accept_scalar  accept / reject from the scalars, copy the accepted rows
// instead of propose + two density kernels + an accept kernel that re-reads both particle sets
// (asp_pcn_propose / asp_logdensity x2 / asp_pcn_accept).  Same step spec (oracle/pcn.py), same Philox
// stream layout as the fused sweep, so all three routes give the same trajectory.
// Replaces the per-move body of minipcn.Sampler.sample as driven by
// /root/reference/src/aspire/samplers/smc/minipcn.py:89-117 with the target of smc/base.py:507-519.
#include <math.h>

#include "common.cuh"
#include "density.cuh"
#include "pcn_sweep_tpp.cuh"
#include "philox.cuh"

namespace asp {

struct EvalArgs {
    const double* x;
    int64_t n;
    int dims;
    asp_density_t like, prior;
    asp_pcn_params_t prm;
    int move;
    PhiloxKeys keys;
    double *xp, *ll2, *lp2, *dr;
};

constexpr int kEvalBlock = 128;
struct ZeroMu {
    __device__ double operator()(int) const { return 0.0; }
};

template <int DL, int G, int LIKE, int PRIOR>
__global__ void __launch_bounds__(kEvalBlock) pcn_propose_eval_kernel(const EvalArgs A) {
    constexpr int DT = DL * G;
    constexpr int NQ = DL / 4;
    constexpr bool kGmm = LIKE == ASP_DENSITY_GMM;
    extern __shared__ __align__(16) double s_dyn[];  // mixture: means [DT][ncp], {1/2 is^2, logc}[ncp]
    const double ninf = identity(OP_MAX);
    const int n_comp = kGmm ? A.like.n_comp : 0;
    const int ncp = (n_comp + kTppChunk - 1) / kTppChunk * kTppChunk;
    double* s_gm = s_dyn;
    double2* s_gw = reinterpret_cast<double2*>(s_dyn + (size_t)DT * ncp);
    if constexpr (kGmm) {
        for (int e = threadIdx.x; e < DT * ncp; e += blockDim.x) {
            const int k = e / ncp, c = e % ncp;
            s_gm[e] = c < n_comp ? __ldg(A.like.p0 + (size_t)c * DT + k) : 0.0;
        }
        for (int c = threadIdx.x; c < ncp; c += blockDim.x) {
            const double is = c < n_comp ? __ldg(A.like.p1 + c) : 0.0;
            s_gw[c] = make_double2(0.5 * is * is, c < n_comp ? __ldg(A.like.p2 + c) : ninf);
        }
        __syncthreads();
    }
    const int lane = (G == 1) ? 0 : (int)(threadIdx.x % G);
    constexpr int kPerWarp = 32 / G;
    const int sub = (threadIdx.x & 31) / G;
    const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t stride = (((int64_t)gridDim.x * blockDim.x) >> 5) * kPerWarp;
    const double rho = A.prm.rho, ca = sqrt(1.0 - rho * rho);
    const uint32_t stream = A.prm.stream0 + (uint32_t)A.move;
    for (int64_t base = warp_id * kPerWarp; base < A.n; base += stride) {  // warp-uniform trip count
        const int64_t i_raw = base + sub;
        const bool active = i_raw < A.n;
        const int64_t i = active ? i_raw : A.n - 1;
        const double* row = A.x + i * DT + lane * DL;
        const uint64_t pidx = (uint64_t)(A.prm.d_slot ? A.prm.d_slot[i] : A.prm.index0 + i);
        double xp[DL];
        double r0 = 0.0, r1 = 0.0, sp = 0.0, sl = 0.0, ros = 0.0;
        bool pbad = false, lbad = false;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            double xi[4];
            if (A.prm.noise_mode == 0) {
                float z[4];
                philox_normal4(A.keys, stream, pidx, (uint32_t)(lane * NQ + q), z);
#pragma unroll
                for (int j = 0; j < 4; ++j) xi[j] = (double)z[j];
            } else {
                const double* p = A.prm.d_xi + ((size_t)A.move * A.n + i) * DT + lane * DL + q * 4;
                const double2 a = ldg2(p), b = ldg2(p + 2);
                xi[0] = a.x; xi[1] = a.y; xi[2] = b.x; xi[3] = b.y;
            }
            const double2 xa = ldg2(row + q * 4), xb = ldg2(row + q * 4 + 2);
            const double xs[4] = {xa.x, xa.y, xb.x, xb.y};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = q * 4 + j, d = lane * DL + k;
                const double mu = __ldg(A.prm.d_mu + d), sg = __ldg(A.prm.d_sigma + d), isg = 1.0 / sg;
                const double y = xs[j] - mu;
                const double v = mu + (ca * y + (rho * sg) * xi[j]);  // oracle/pcn.py::propose
                xp[k] = v;
                const double z0 = y * isg, z1 = (v - mu) * isg;
                r0 = fma(z0, z0, r0);
                r1 = fma(z1, z1, r1);
                if constexpr (PRIOR == ASP_DENSITY_GAUSSIAN) {
                    const double z = (v - __ldg(A.prior.p0 + d)) * __ldg(A.prior.p1 + d);
                    sp = fma(z, z, sp);
                } else {
                    pbad |= !(v > __ldg(A.prior.p0 + d) && v < __ldg(A.prior.p1 + d));
                }
                if constexpr (LIKE == ASP_DENSITY_GAUSSIAN) {
                    const double z = (v - __ldg(A.like.p0 + d)) * __ldg(A.like.p1 + d);
                    sl = fma(z, z, sl);
                } else if constexpr (LIKE == ASP_DENSITY_UNIFORM) {
                    lbad |= !(v > __ldg(A.like.p0 + d) && v < __ldg(A.like.p1 + d));
                }
            }
        }
        if constexpr (LIKE == ASP_DENSITY_ROSENBROCK) {
#pragma unroll
            for (int k = 0; k < DL; ++k) {
                double x1;
                if (k + 1 < DL) x1 = xp[(k + 1) % DL];
                else x1 = (G > 1) ? __shfl_down_sync(0xffffffffu, xp[0], 1, G) : 0.0;
                const double tt = x1 - xp[k] * xp[k];
                const double u = A.like.s1 - xp[k];
                if (k + 1 < DL || lane < G - 1) ros += A.like.s0 * tt * tt + u * u;
            }
        }
        double lp, ll;
        if constexpr (PRIOR == ASP_DENSITY_GAUSSIAN) lp = A.prior.s0 - 0.5 * group_sum<G>(sp);
        else lp = group_sum<G>(pbad ? 1.0 : 0.0) > 0.0 ? ninf : A.prior.s0;
        if constexpr (LIKE == ASP_DENSITY_GAUSSIAN) ll = A.like.s0 - 0.5 * group_sum<G>(sl);
        else if constexpr (LIKE == ASP_DENSITY_UNIFORM) ll = group_sum<G>(lbad ? 1.0 : 0.0) > 0.0 ? ninf : A.like.s0;
        else if constexpr (LIKE == ASP_DENSITY_ROSENBROCK) ll = -group_sum<G>(ros) / A.like.s2;
        else ll = tpp_gmm<DL, G>(xp, ZeroMu(), lane, s_gm, s_gw, n_comp, ncp);
        const double dr = 0.5 * group_sum<G>(r1) - 0.5 * group_sum<G>(r0);
        if (active) {
            double* dst = A.xp + i * DT + lane * DL;
#pragma unroll
            for (int k = 0; k < DL; k += 2) st_stream2(dst + k, xp[k], xp[k + 1]);
            if (lane == 0) {
                A.ll2[i] = ll;
                A.lp2[i] = lp;
                A.dr[i] = dr;
            }
        }
    }
}

// accept / reject from scalars; one lane decides for one particle, the warp copies the accepted rows
__global__ void __launch_bounds__(kBlock) pcn_accept_scalar_kernel(double* __restrict__ x, double* __restrict__ ll,
                                                                   double* __restrict__ lp, double* __restrict__ lq,
                                                                   const double* __restrict__ xp,
                                                                   const double* __restrict__ ll2,
                                                                   const double* __restrict__ lp2,
                                                                   const double* __restrict__ lq2,
                                                                   const double* __restrict__ dr, int64_t n, int dims,
                                                                   const asp_pcn_params_t prm, int move,
                                                                   unsigned long long* accepts) {
    __shared__ unsigned int s_acc;
    if (threadIdx.x == 0) s_acc = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    unsigned int n_acc = 0;
    for (int64_t base = warp_id * 32; base < n; base += n_warps * 32) {
        const int64_t i = base + lane;
        bool acc = false;
        if (i < n) {
            const double cur = tempered(ll[i], lp[i], lq[i], prm.beta);
            const double l2 = ll2[i], p2 = lp2[i], q2 = lq2[i];
            const double nw = tempered(l2, p2, q2, prm.beta);
            double log_u;
            if (prm.noise_mode == 0) {
                log_u = accept_log_u(accept_words(prm.seed, prm.stream0, move,
                                                  (uint64_t)(prm.d_slot ? prm.d_slot[i] : prm.index0 + i)), move);
            } else {
                log_u = prm.d_log_u[(size_t)move * n + i];
            }
            acc = log_u < (nw - cur) + dr[i];
            if (acc) {
                ll[i] = l2;
                lp[i] = p2;
                lq[i] = q2;
            }
        }
        unsigned int votes = __ballot_sync(0xffffffffu, acc);
        n_acc += __popc(votes);
        while (votes) {
            const int r = __ffs(votes) - 1;
            votes &= votes - 1;
            const double* src = xp + (base + r) * dims;
            double* dst = x + (base + r) * dims;
            for (int k = lane; k < dims; k += 32) dst[k] = src[k];
        }
    }
    if (lane == 0 && n_acc) atomicAdd(&s_acc, n_acc);
    __syncthreads();
    if (threadIdx.x == 0 && s_acc) atomicAdd(&accepts[move], (unsigned long long)s_acc);
}

template <int DL, int G, int LIKE, int PRIOR>
static int launch_eval_n(const EvalArgs& A, cudaStream_t s) {
    const int64_t per_block = (int64_t)kEvalBlock / G;
    const int64_t need = (A.n + per_block - 1) / per_block;
    const int64_t cap = (int64_t)kNumSMs * 4;
    const size_t smem = tpp_gmm_bytes(A.dims, LIKE == ASP_DENSITY_GMM ? A.like.n_comp : 0);
    pcn_propose_eval_kernel<DL, G, LIKE, PRIOR><<<(int)(need < cap ? need : cap), kEvalBlock, smem, s>>>(A);
    return ASP_OK;
}

template <int LIKE, int PRIOR>
static int launch_eval_shape(const EvalArgs& A, cudaStream_t s) {
    int dl = 0, g = 0;
    if (!tpp_shape(A.dims, LIKE, dl, g)) return ASP_ERR_ARG;
    if (g == 2) return launch_eval_n<32, 2, LIKE, PRIOR>(A, s);
    if (dl == 32) return launch_eval_n<32, 1, LIKE, PRIOR>(A, s);
    if (dl == 16) return launch_eval_n<16, 1, LIKE, PRIOR>(A, s);
    if (dl == 8) return launch_eval_n<8, 1, LIKE, PRIOR>(A, s);
    return launch_eval_n<4, 1, LIKE, PRIOR>(A, s);
}

template <int LIKE>
static int launch_eval_like(const EvalArgs& A, cudaStream_t s) {
    if (A.prior.kind == ASP_DENSITY_GAUSSIAN) return launch_eval_shape<LIKE, ASP_DENSITY_GAUSSIAN>(A, s);
    if (A.prior.kind == ASP_DENSITY_UNIFORM) return launch_eval_shape<LIKE, ASP_DENSITY_UNIFORM>(A, s);
    return ASP_ERR_ARG;
}

}  // namespace asp

using namespace asp;

extern "C" {

int32_t asp_pcn_propose_eval_supported(int32_t dims, int32_t like_kind, int32_t prior_kind) {
    int dl = 0, g = 0;
    return tpp_shape(dims, like_kind, dl, g) && like_kind >= ASP_DENSITY_GAUSSIAN && like_kind <= ASP_DENSITY_UNIFORM &&
                   (prior_kind == ASP_DENSITY_GAUSSIAN || prior_kind == ASP_DENSITY_UNIFORM)
               ? 1 : 0;
}

int asp_pcn_propose_eval(const double* d_x, int64_t n, int32_t dims, const asp_density_t* like,
                         const asp_density_t* prior, const asp_pcn_params_t* params, int32_t move, double* d_x_prop,
                         double* d_log_l_prop, double* d_log_p_prop, double* d_dr, void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(d_x && like && prior && params && d_x_prop && d_log_l_prop && d_log_p_prop && d_dr, "null pointer");
    ASP_REQUIRE(n > 0, "n must be positive");
    ASP_REQUIRE(like->dims == dims && prior->dims == dims, "density dims mismatch");
    ASP_REQUIRE(asp_pcn_propose_eval_supported(dims, like->kind, prior->kind), "unsupported dims / density kinds");
    ASP_REQUIRE(params->d_mu && params->d_sigma, "reference measure missing");
    ASP_REQUIRE(params->rho > 0.0 && params->rho <= 1.0, "rho must be in (0, 1]");
    ASP_REQUIRE(params->noise_mode == 0 || params->d_xi, "noise array missing");
    ASP_REQUIRE(!(params->nu > 0.0), "asp_pcn_propose_eval is pCN only (tpCN: asp_tpcn_scales + the unfused pair)");
    ASP_REQUIRE(like->kind != ASP_DENSITY_GMM || (like->p0 && like->p1 && like->p2 && like->n_comp >= 1 &&
                                                  like->n_comp <= 32), "bad mixture parameters");
    ASP_REQUIRE(((reinterpret_cast<uintptr_t>(d_x) | reinterpret_cast<uintptr_t>(d_x_prop) |
                  reinterpret_cast<uintptr_t>(params->d_xi)) & 15) == 0, "rows must be 16-byte aligned");
    EvalArgs A;
    A.x = d_x; A.n = n; A.dims = dims; A.like = *like; A.prior = *prior; A.prm = *params; A.move = move;
    A.keys = philox_keys(params->seed);
    A.xp = d_x_prop; A.ll2 = d_log_l_prop; A.lp2 = d_log_p_prop; A.dr = d_dr;
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    switch (like->kind) {
        case ASP_DENSITY_GAUSSIAN: rc = launch_eval_like<ASP_DENSITY_GAUSSIAN>(A, s); break;
        case ASP_DENSITY_ROSENBROCK: rc = launch_eval_like<ASP_DENSITY_ROSENBROCK>(A, s); break;
        case ASP_DENSITY_GMM: rc = launch_eval_like<ASP_DENSITY_GMM>(A, s); break;
        default: rc = launch_eval_like<ASP_DENSITY_UNIFORM>(A, s); break;
    }
    if (rc) { set_error("asp_pcn_propose_eval: unsupported density combination"); return rc; }
    return check_launch("asp_pcn_propose_eval");
}

int asp_pcn_accept_scalar(double* d_x, double* d_log_l, double* d_log_p, double* d_log_q, const double* d_x_prop,
                          const double* d_log_l_prop, const double* d_log_p_prop, const double* d_log_q_prop,
                          const double* d_dr, int64_t n, int32_t dims, const asp_pcn_params_t* params, int32_t move,
                          unsigned long long* d_accepts, void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(d_x && d_log_l && d_log_p && d_log_q && d_x_prop && d_log_l_prop && d_log_p_prop && d_log_q_prop &&
                    d_dr && params && d_accepts, "null pointer");
    ASP_REQUIRE(n > 0 && dims > 0, "bad sizes");
    ASP_REQUIRE(params->noise_mode == 0 || params->d_log_u, "noise array missing");
    pcn_accept_scalar_kernel<<<grid_for(n, 1), kBlock, 0, (cudaStream_t)stream>>>(
        d_x, d_log_l, d_log_p, d_log_q, d_x_prop, d_log_l_prop, d_log_p_prop, d_log_q_prop, d_dr, n, dims, *params, move,
        d_accepts);
    return check_launch("asp_pcn_accept_scalar");
}

}  // extern "C"

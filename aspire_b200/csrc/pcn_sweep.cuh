// K8: fused pCN mutation sweep, specialised at compile time on
// (lanes per particle G, likelihood kind, prior kind); the proposal density q
// is a diagonal Gaussian.  See pcn.cu for the entry point and oracle/pcn.py for
// the step spec.
#pragma once
#include <math.h>

#include "common.cuh"
#include "density.cuh"
#include "philox.cuh"

namespace asp {

struct SweepArgs {
    const double *x, *ll, *lp, *lq;
    const int64_t* idx;
    int64_t n;
    int dims;
    asp_density_t like, prior, q;
    asp_pcn_params_t prm;
    double *xo, *llo, *lpo, *lqo;
    unsigned long long* accepts;
    PhiloxKeys keys;
};

constexpr int kMaxMovesSmem = 1024;
#ifndef ASP_SWEEP_UNROLL
#define ASP_SWEEP_UNROLL 4
#endif
constexpr int kSweepUnroll = ASP_SWEEP_UNROLL;
#ifndef ASP_SWEEP_MIN_BLOCKS
#define ASP_SWEEP_MIN_BLOCKS 2
#endif

__device__ __forceinline__ double tempered(double ll, double lp, double lq, double beta) {
    const double t = (1.0 - beta) * lq + beta * (ll + lp);
    return (t != t) ? identity(OP_MAX) : t;  // NaN -> -inf (smc/base.py:518)
}

template <int G>
__device__ __forceinline__ double ref_quad(const double (&x)[kDPL], const double (&mu)[kDPL],
                                           const double (&isig)[kDPL]) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < kDPL; ++k) {  // padded dimensions have isig == 0
        const double z = (x[k] - mu[k]) * isig[k];
        s += z * z;
    }
    return 0.5 * group_sum<G>(s);
}

// log(u) of move `move`: word (move & 3) of the accept block of stream stream0 + (move & ~3)
__device__ __forceinline__ uint4 accept_words(uint64_t seed, uint32_t stream0, int move, uint64_t index) {
    return philox4x32<kNoiseRounds>((uint32_t)index, (uint32_t)(index >> 32), stream0 + (uint32_t)(move & ~3),
                                    kAcceptBlock, (uint32_t)seed, (uint32_t)(seed >> 32));
}
__device__ __forceinline__ uint4 accept_words(const PhiloxKeys& K, uint32_t stream0, int move, uint64_t index) {
    return philox4x32<kNoiseRounds>((uint32_t)index, (uint32_t)(index >> 32), stream0 + (uint32_t)(move & ~3),
                                    kAcceptBlock, K);
}
__device__ __forceinline__ double accept_log_u(const uint4& w, int move) {
    const uint32_t r = (move & 3) == 0 ? w.x : (move & 3) == 1 ? w.y : (move & 3) == 2 ? w.z : w.w;
    return (double)sfu_log(u01_f32(r));
}

// The accept/reject test only needs  T(x') - T(x),
//   T(x) = (1-beta) log_q(x) + beta (logL(x) + logP(x)) + r(x),  r = 1/2 |(x - mu)/sigma|^2 .
// Every Gaussian term (q, a Gaussian prior, a Gaussian likelihood, r) is a
// per-dimension quadratic, so they fold into ONE quadratic in the centred state
// y = x - mu:  a2_k y_k^2 + a1_k y_k  (constants cancel in the difference).  That
// is 2 DFMA per dimension and one group reduction per move instead of four
// densities and four reductions; centring on mu keeps the expansion free of
// cancellation (y ~ sigma).  log_q / logP / logL themselves are evaluated once,
// at the end, and only for particles that moved.
template <int G, int LIKE, int PRIOR>
__global__ void __launch_bounds__(kBlock, ASP_SWEEP_MIN_BLOCKS) pcn_sweep_kernel(const SweepArgs A) {
    __shared__ unsigned int s_acc[kMaxMovesSmem];
    const int dims = A.dims;
    const int lane = threadIdx.x % G;
    const int n_moves = A.prm.n_moves;
    for (int t = threadIdx.x; t < max(n_moves, 1); t += blockDim.x) s_acc[t] = 0;
    __syncthreads();

    constexpr bool kLikeQuad = LIKE == ASP_DENSITY_GAUSSIAN;
    constexpr bool kPriorQuad = PRIOR == ASP_DENSITY_GAUSSIAN;
    const double ninf = identity(OP_MAX);
    const double rho = A.prm.rho, beta = A.prm.beta;
    const double ca = sqrt(1.0 - rho * rho);
    double mu[kDPL], rs[kDPL], a2[kDPL], a1[kDPL];
    double plo[kDPL], phi[kDPL], llo[kDPL], lhi[kDPL];  // uniform boxes (dead code otherwise)
#pragma unroll
    for (int k = 0; k < kDPL; ++k) {
        const int j = lane * kDPL + k;
        const bool in = j < dims;
        const double sg = in ? __ldg(A.prm.d_sigma + j) : 1.0;
        mu[k] = in ? __ldg(A.prm.d_mu + j) : 0.0;
        rs[k] = in ? rho * sg : 0.0;
        double c2 = 0.0, c1 = 0.0;
        if (in) {
            const double isr = 1.0 / sg;
            c2 = 0.5 * isr * isr;  // r(x)
            const double isq = __ldg(A.q.p1 + j), wq = -0.5 * (1.0 - beta) * isq * isq;
            c2 += wq;
            c1 += 2.0 * wq * (mu[k] - __ldg(A.q.p0 + j));
            if constexpr (kPriorQuad) {
                const double is = __ldg(A.prior.p1 + j), w = -0.5 * beta * is * is;
                c2 += w;
                c1 += 2.0 * w * (mu[k] - __ldg(A.prior.p0 + j));
            }
            if constexpr (kLikeQuad) {
                const double is = __ldg(A.like.p1 + j), w = -0.5 * beta * is * is;
                c2 += w;
                c1 += 2.0 * w * (mu[k] - __ldg(A.like.p0 + j));
            }
        }
        a2[k] = c2;
        a1[k] = c1;
        plo[k] = phi[k] = llo[k] = lhi[k] = 0.0;
        if constexpr (PRIOR == ASP_DENSITY_UNIFORM) {
            plo[k] = in ? __ldg(A.prior.p0 + j) : ninf;
            phi[k] = in ? __ldg(A.prior.p1 + j) : identity(OP_MIN);
        }
        if constexpr (LIKE == ASP_DENSITY_UNIFORM) {
            llo[k] = in ? __ldg(A.like.p0 + j) : ninf;
            lhi[k] = in ? __ldg(A.like.p1 + j) : identity(OP_MIN);
        }
    }
    Dens<LIKE> like;  // only the non-separable kinds keep per-lane state in the loop
    if constexpr (LIKE == ASP_DENSITY_GMM) like.load(A.like, lane, dims);
    const double ros_a = A.like.s0, ros_b = A.like.s1;
    const double ros_w = LIKE == ASP_DENSITY_ROSENBROCK ? -beta / A.like.s2 : 0.0;  // beta and the scale folded

    // T(y): the lane's partial of the separable part, group-reduced; non-separable likelihoods added after.
    auto target = [&](const double(&y)[kDPL], const double(&x)[kDPL]) -> double {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < kDPL; ++k) s = fma(y[k], fma(a2[k], y[k], a1[k]), s);
        if constexpr (LIKE == ASP_DENSITY_ROSENBROCK) {
            const double xn = __shfl_down_sync(0xffffffffu, x[0], 1, G);
            double r = 0.0;
#pragma unroll
            for (int k = 0; k < kDPL; ++k) {
                const double x1 = k + 1 < kDPL ? x[(k + 1) % kDPL] : xn;
                const double t = x1 - x[k] * x[k];
                const double u = ros_b - x[k];
                r += (lane * kDPL + k + 1 < dims) ? ros_a * t * t + u * u : 0.0;
            }
            s = fma(ros_w, r, s);
        }
        if constexpr (PRIOR == ASP_DENSITY_UNIFORM || LIKE == ASP_DENSITY_UNIFORM) {
            bool bad = false;
#pragma unroll
            for (int k = 0; k < kDPL; ++k) {
                if constexpr (PRIOR == ASP_DENSITY_UNIFORM) bad |= !(x[k] > plo[k] && x[k] < phi[k]);
                if constexpr (LIKE == ASP_DENSITY_UNIFORM) bad |= !(x[k] > llo[k] && x[k] < lhi[k]);
            }
            if (bad) s = ninf;  // -inf target: the move is rejected
        }
        double T = group_sum<G>(s);
        if constexpr (LIKE == ASP_DENSITY_GMM) T += beta * like.template eval<G>(x, lane, dims);
        return T;
    };

    const bool vec = (dims % kDPL == 0) && ((reinterpret_cast<uintptr_t>(A.x) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(A.xo) & 15) == 0);
    const bool vec_xi = (dims % kDPL == 0) && ((reinterpret_cast<uintptr_t>(A.prm.d_xi) & 15) == 0);
    const bool philox = A.prm.noise_mode == 0;
    const bool per_move = A.prm.reserved == 0;  // reserved = 1: only the total lands in accepts[0]
    unsigned int n_acc = 0;
    const int64_t groups = ((int64_t)gridDim.x * blockDim.x) / G;
    const int64_t g0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    // warp-uniform trip count: all 32 lanes stay converged for the shuffles
    const int64_t warp_first = ((int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31)) / G;
    for (int64_t base = warp_first; base < A.n; base += groups) {
        const int64_t i_raw = base + (g0 - warp_first);
        const bool active = i_raw < A.n;
        const int64_t i = active ? i_raw : A.n - 1;
        const int64_t src = A.idx ? A.idx[i] : i;
        // the next particle of this group: its ancestor index is loaded now, its row is
        // prefetched during the moves (the gather is a dependent idx -> row load chain)
        const int64_t i_next = i_raw + groups;
        const int64_t src_next = i_next < A.n ? (A.idx ? A.idx[i_next] : i_next) : src;
        double y[kDPL];
        {
            double x[kDPL];
            row_load(A.x, src, dims, lane, vec, x);
#pragma unroll
            for (int k = 0; k < kDPL; ++k) y[k] = x[k] - mu[k];
        }
        double cur;
        {
            double x[kDPL];
#pragma unroll
            for (int k = 0; k < kDPL; ++k) x[k] = mu[k] + y[k];
            cur = target(y, x);
        }
        if (!active) cur = identity(OP_MIN);  // +inf: a padding group never accepts, so the loop needs no mask
        bool moved = false;
        const uint64_t pidx = (uint64_t)(A.prm.d_slot ? A.prm.d_slot[i] : A.prm.index0 + i);
        uint4 aw = make_uint4(0, 0, 0, 0);
#pragma unroll kSweepUnroll
        for (int t = 0; t < n_moves; ++t) {
            double xi[kDPL], log_u;
            if (t == 1) asm volatile("prefetch.global.L1 [%0];" ::"l"(A.x + src_next * dims + lane * kDPL));
            if (philox) {
                float z[4];
                philox_normal4(A.keys, A.prm.stream0 + (uint32_t)t, pidx, (uint32_t)lane, z);
#pragma unroll
                for (int k = 0; k < kDPL; ++k) xi[k] = (double)z[k];
                if ((t & 3) == 0) aw = accept_words(A.keys, A.prm.stream0, t, pidx);  // every lane, no divergence
                log_u = accept_log_u(aw, t);
            } else {
                row_load(A.prm.d_xi + (size_t)t * A.n * dims, i, dims, lane, vec_xi, xi);
                log_u = A.prm.d_log_u[(size_t)t * A.n + i];
            }
            double xp[kDPL], yp[kDPL];
#pragma unroll
            for (int k = 0; k < kDPL; ++k) {
                yp[k] = ca * y[k] + rs[k] * xi[k];
                xp[k] = mu[k] + yp[k];
            }
            const double nw = target(yp, xp);
            const bool acc = log_u < nw - cur;  // NaN -> reject, like the reference's NaN -> -inf
            if (acc) {
#pragma unroll
                for (int k = 0; k < kDPL; ++k) y[k] = yp[k];
                cur = nw;
                moved = true;
            }
            if (per_move) {
                const unsigned int votes = __ballot_sync(0xffffffffu, acc && lane == 0);
                if ((threadIdx.x & 31) == 0 && votes) atomicAdd(&s_acc[t], __popc(votes));
            } else {
                n_acc += acc ? 1u : 0u;
            }
        }
        // final state and its logs: carried over untouched if the particle never moved, evaluated once
        // otherwise (by every group, so that the shuffles stay converged)
        double x[kDPL];
        if (moved) {
#pragma unroll
            for (int k = 0; k < kDPL; ++k) x[k] = mu[k] + y[k];
        } else {
            row_load(A.x, src, dims, lane, vec, x);
        }
        double lq = eval_once<ASP_DENSITY_GAUSSIAN, G>(A.q, x, lane, dims);
        double lp = eval_once<PRIOR, G>(A.prior, x, lane, dims);
        double ll = eval_once<LIKE, G>(A.like, x, lane, dims);
        if (!moved) {
            ll = A.ll[src];
            lp = A.lp[src];
            lq = A.lq[src];
        }
        if (active) {
            row_store(A.xo, i, dims, lane, vec, x);
            if (lane == 0) {
                A.llo[i] = ll;
                A.lpo[i] = lp;
                A.lqo[i] = lq;
                if (lq != lq) atomicAdd(&A.accepts[per_move ? n_moves : 1], 1ull);  // NaN log q (smc/minipcn.py:133)
            }
        }
    }
    if (!per_move) {
        n_acc = lane == 0 ? n_acc : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) n_acc += __shfl_xor_sync(0xffffffffu, n_acc, o);
        if ((threadIdx.x & 31) == 0 && n_acc) atomicAdd(&s_acc[0], n_acc);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < (per_move ? n_moves : 1); t += blockDim.x)
        if (s_acc[t]) atomicAdd(&A.accepts[t], (unsigned long long)s_acc[t]);
}

template <int LIKE, int PRIOR>
inline void launch_sweep_g(const SweepArgs& A, int g, int grid, cudaStream_t s) {
    switch (g) {
        case 1: pcn_sweep_kernel<1, LIKE, PRIOR><<<grid, kBlock, 0, s>>>(A); break;
        case 2: pcn_sweep_kernel<2, LIKE, PRIOR><<<grid, kBlock, 0, s>>>(A); break;
        case 4: pcn_sweep_kernel<4, LIKE, PRIOR><<<grid, kBlock, 0, s>>>(A); break;
        case 8: pcn_sweep_kernel<8, LIKE, PRIOR><<<grid, kBlock, 0, s>>>(A); break;
        case 16: pcn_sweep_kernel<16, LIKE, PRIOR><<<grid, kBlock, 0, s>>>(A); break;
        default: pcn_sweep_kernel<32, LIKE, PRIOR><<<grid, kBlock, 0, s>>>(A); break;
    }
}

// one translation unit per likelihood kind (parallel compilation)
template <int LIKE>
inline int launch_sweep_like(const SweepArgs& A, int g, int grid, cudaStream_t s) {
    if (A.prior.kind == ASP_DENSITY_GAUSSIAN) launch_sweep_g<LIKE, ASP_DENSITY_GAUSSIAN>(A, g, grid, s);
    else if (A.prior.kind == ASP_DENSITY_UNIFORM) launch_sweep_g<LIKE, ASP_DENSITY_UNIFORM>(A, g, grid, s);
    else return ASP_ERR_ARG;
    return ASP_OK;
}

int launch_sweep_gaussian(const SweepArgs& A, int g, int grid, cudaStream_t s);
int launch_sweep_rosenbrock(const SweepArgs& A, int g, int grid, cudaStream_t s);
int launch_sweep_gmm(const SweepArgs& A, int g, int grid, cudaStream_t s);
int launch_sweep_uniform(const SweepArgs& A, int g, int grid, cudaStream_t s);

}  // namespace asp

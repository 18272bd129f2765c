// K8, particle-per-thread form of the fused pCN sweep (the fast path of
// asp_pcn_sweep; pcn_sweep.cuh keeps the general lanes-per-particle kernel).
//
// One thread (G = 1) or G adjacent lanes own a particle; each lane keeps DL
// dimensions of the centred state y = x - mu and of the proposal in registers
// for the whole sweep.  Compared with the 4-dims-per-lane kernel nothing is
// replicated across lanes (accept test, log u, reductions are per thread), so
// the instruction count per particle-move halves; the price is that the
// per-dimension constants no longer fit in registers and are read from shared
// memory as warp-uniform (broadcast) 128-bit loads.
//
// Step spec: oracle/pcn.py.  Noise layout: dimension j of particle slot p in
// move t is word j % 4 of Philox block j / 4, stream stream0 + t, counter p -
// the same stream as the general kernel, so both produce the same trajectory.
#pragma once
#include "pcn_sweep.cuh"

namespace asp {

constexpr int kTppBlock = 128;
constexpr int kTppChunk = 4;         // mixture components evaluated together
constexpr int kTppRegConstDims = 8;  // lanes with at most this many dims keep their constants in registers
constexpr int kTppMaxDims = 64;
#ifndef ASP_TPP_MINB32
#define ASP_TPP_MINB32 2
#endif

// 32 dims per lane: 2 blocks/SM (255 registers, no spills) measured faster than 3 blocks (168, spills)
template <int DL>
constexpr int tpp_min_blocks() { return DL >= 32 ? ASP_TPP_MINB32 : (DL == 8 ? 3 : 4); }

__device__ __forceinline__ double2 ldg2(const double* p) { return __ldg(reinterpret_cast<const double2*>(p)); }
// Shared-memory constant read that stays where it is written: a plain load is loop-invariant, so the
// compiler hoists all 4 * DL constants out of the move loop into registers and spills the state.
__device__ __forceinline__ double2 lds2(const double2* p) {
    double2 r;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"((uint32_t)__cvta_generic_to_shared(p)));
    return r;
}

// bytes of the mixture descriptor staged in dynamic shared memory (means [dims][ncp] + {h, logc}[ncp]), 16-aligned
__host__ __device__ inline size_t tpp_gmm_bytes(int dims, int n_comp) {
    if (n_comp <= 0) return 0;
    const int ncp = (n_comp + 3) / 4 * 4;
    return (sizeof(double) * ((size_t)dims * ncp + 2 * (size_t)ncp) + 15) & ~(size_t)15;
}

__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gptr) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// branch-free streaming log-sum-exp update with one exp:  (m, acc) <- (m, acc) (+) t
__device__ __forceinline__ void lse_push(double& m, double& acc, double t) {
    const double ninf = identity(OP_MAX);
    const double d = t - m;  // +inf on the first finite term
    const double ad = fabs(d);
    const double e = ad < 746.0 ? exp(-ad) : 0.0;  // exp underflows to exactly 0 below -745.14: far components cost no exp
    const double up = fma(acc, e, 1.0), keep = acc + e;
    if (t > ninf) {
        acc = d > 0.0 ? up : keep;
        m = fmax(m, t);
    }
}

// log of the isotropic mixture at x = mu + v (descriptor staged in shared memory):
//   s_gm [G*DL][ncp] component means, dimension-major;  s_gw [ncp] = {1/2 inv_sigma^2, logc}
template <int DL, int G, class MuOf>
__device__ __forceinline__ double tpp_gmm(const double (&v)[DL], MuOf mu_of, int lane,
                                          const double* __restrict__ s_gm, const double2* __restrict__ s_gw, int n_comp,
                                          int ncp) {
    double m = identity(OP_MAX), acc = 0.0;
    for (int c0 = 0; c0 < n_comp; c0 += kTppChunk) {
        double d2[kTppChunk];
#pragma unroll
        for (int j = 0; j < kTppChunk; ++j) d2[j] = 0.0;
#pragma unroll
        for (int k = 0; k < DL; ++k) {
            const double xk = mu_of(k) + v[k];
            const double* mrow = s_gm + (size_t)(lane * DL + k) * ncp + c0;
            const double2 m01 = lds2(reinterpret_cast<const double2*>(mrow));
            const double2 m23 = lds2(reinterpret_cast<const double2*>(mrow + 2));
            const double e0 = xk - m01.x, e1 = xk - m01.y, e2 = xk - m23.x, e3 = xk - m23.y;
            d2[0] = fma(e0, e0, d2[0]);
            d2[1] = fma(e1, e1, d2[1]);
            d2[2] = fma(e2, e2, d2[2]);
            d2[3] = fma(e3, e3, d2[3]);
        }
#pragma unroll
        for (int j = 0; j < kTppChunk; ++j) {
            const double dd = group_sum<G>(d2[j]);
            if (c0 + j < n_comp) {  // warp-uniform
                const double2 w = lds2(&s_gw[c0 + j]);
                lse_push(m, acc, fma(-w.x, dd, w.y));
            }
        }
    }
    return m + log(acc);
}

// Marsaglia & Tsang (2000) gamma variate of the tpCN scale, fp32, counter-based: the same draws as
// asp_tpcn_scales (pcn.cu::gamma_mt; spec oracle/pcn.py::gamma_mt_f32) from the precomputed round keys
constexpr uint32_t kTppGammaBlock0 = 0xFFFFFF00u;
__device__ __forceinline__ double tpp_gamma_mt(const PhiloxKeys& K, uint32_t stream, uint64_t index, float a) {
    const float d = a - (1.0f / 3.0f);
    const float c = 1.0f / sqrtf(9.0f * d);
    for (uint32_t it = 0; it < 64; ++it) {
        const uint4 r = philox4x32<kNoiseRounds>((uint32_t)index, (uint32_t)(index >> 32), stream, kTppGammaBlock0 + it, K);
        double g;
        if (gamma_mt_try(r, d, c, g)) return g;
    }
    return (double)d;
}

// TPCN: the Student-t preconditioned step (the reference's default step_fn, smc/minipcn.py:48; spec
// oracle/pcn.py::tpcn_sweep).  Per move and particle: g ~ Gamma((D + nu)/2, 1), proposal scale
// sqrt((nu + r2(x)) / (2 g)) on the noise term, and the reference term of the accept test becomes
// (nu + D)/2 log1p(r2 / nu) instead of r2 / 2 - so r2 = |(x - mu)/sigma|^2 is carried beside the folded
// quadratic (3 more fp64 instructions per dimension and move than pCN).
template <int DL, int G, int LIKE, int PRIOR, bool PHILOX, bool TPCN>
__global__ void __launch_bounds__(kTppBlock, tpp_min_blocks<DL>()) pcn_sweep_tpp_kernel(const SweepArgs A) {
    constexpr int DT = DL * G;  // == dims
    constexpr int NQ = DL / 4;
    constexpr bool kLikeQuad = LIKE == ASP_DENSITY_GAUSSIAN;
    constexpr bool kPriorQuad = PRIOR == ASP_DENSITY_GAUSSIAN;
    constexpr bool kRosen = LIKE == ASP_DENSITY_ROSENBROCK;
    constexpr bool kGmm = LIKE == ASP_DENSITY_GMM;
    constexpr bool kPBox = PRIOR == ASP_DENSITY_UNIFORM, kLBox = LIKE == ASP_DENSITY_UNIFORM;
    constexpr bool kNeedX = kRosen || kPBox || kLBox;

    constexpr bool kRC = DL <= kTppRegConstDims;  // constants in registers (4 DL doubles per lane), else shared memory
    __shared__ double2 s_c01[DT];  // {rho sigma_k, mu_k}
    __shared__ double2 s_c23[DT];  // {a2_k, a1_k}: folded quadratic of the accept test
    __shared__ double2 s_pbox[kPBox ? DT : 1], s_lbox[kLBox ? DT : 1];
    __shared__ double s_isr[TPCN ? DT : 1];  // 1 / sigma_k (tpCN: r2 is needed by itself)
    // final logs: a Gaussian density at x = mu + y is  s0 - 1/2 sum_k (is_k y_k + is_k (mu_k - m_k))^2
    __shared__ double2 s_fq[DT], s_fp[kPriorQuad ? DT : 1], s_fl[kLikeQuad ? DT : 1];  // {is_k, is_k (mu_k - m_k)}
    __shared__ unsigned int s_acc[kMaxMovesSmem];
    extern __shared__ __align__(16) double s_dyn[];  // mixture: means [DT][ncp], {1/2 is^2, logc}[ncp]; then row staging

    const int n_moves = A.prm.n_moves;
    const bool per_move = A.prm.reserved == 0;
    const double ninf = identity(OP_MAX);
    const double rho = A.prm.rho, beta = A.prm.beta;
    const double ca = sqrt(1.0 - rho * rho);
    const double nu = TPCN ? A.prm.nu : 1.0, t_half = 0.5 * (nu + (double)DT), inv_nu = 1.0 / nu;
    const float gamma_a = (float)t_half;
    const double ros_w = kRosen ? -beta / A.like.s2 : 0.0;  // beta and the scale folded
    const double ros_wa = ros_w * A.like.s0;
    const int n_comp = kGmm ? A.like.n_comp : 0;
    const int ncp = (n_comp + kTppChunk - 1) / kTppChunk * kTppChunk;
    double* s_gm = s_dyn;
    double2* s_gw = reinterpret_cast<double2*>(s_dyn + (size_t)DT * ncp);

    for (int t = threadIdx.x; t < max(n_moves, 1); t += blockDim.x) s_acc[t] = 0;
    for (int j = threadIdx.x; j < DT; j += blockDim.x) {
        // T(x) = (1-beta) log_q + beta (logL + logP) + r,  r = 1/2 |(x - mu)/sigma|^2: every Gaussian term and
        // Rosenbrock's (b - x_j)^2 are per-dimension quadratics in y = x - mu and fold into a2 y^2 + a1 y
        const double sg = __ldg(A.prm.d_sigma + j), mu = __ldg(A.prm.d_mu + j);
        const double isr = 1.0 / sg;
        double c2 = TPCN ? 0.0 : 0.5 * isr * isr, c1 = 0.0;  // tpCN: the reference term is not a quadratic
        if constexpr (TPCN) s_isr[j] = isr;
        const double isq = __ldg(A.q.p1 + j), wq = -0.5 * (1.0 - beta) * isq * isq;
        c2 += wq;
        c1 += 2.0 * wq * (mu - __ldg(A.q.p0 + j));
        if constexpr (kPriorQuad) {
            const double is = __ldg(A.prior.p1 + j), w = -0.5 * beta * is * is;
            c2 += w;
            c1 += 2.0 * w * (mu - __ldg(A.prior.p0 + j));
        }
        if constexpr (kLikeQuad) {
            const double is = __ldg(A.like.p1 + j), w = -0.5 * beta * is * is;
            c2 += w;
            c1 += 2.0 * w * (mu - __ldg(A.like.p0 + j));
        }
        if constexpr (kRosen) {
            if (j < DT - 1) {  // ros_w (b - mu - y)^2
                c2 += ros_w;
                c1 += 2.0 * ros_w * (mu - A.like.s1);
            }
        }
        s_c01[j] = make_double2(rho * sg, mu);
        s_c23[j] = make_double2(c2, c1);
        s_fq[j] = make_double2(isq, isq * (mu - __ldg(A.q.p0 + j)));
        if constexpr (kPriorQuad) {
            const double is = __ldg(A.prior.p1 + j);
            s_fp[j] = make_double2(is, is * (mu - __ldg(A.prior.p0 + j)));
        }
        if constexpr (kLikeQuad) {
            const double is = __ldg(A.like.p1 + j);
            s_fl[j] = make_double2(is, is * (mu - __ldg(A.like.p0 + j)));
        }
        if constexpr (kPBox) s_pbox[j] = make_double2(__ldg(A.prior.p0 + j), __ldg(A.prior.p1 + j));
        if constexpr (kLBox) s_lbox[j] = make_double2(__ldg(A.like.p0 + j), __ldg(A.like.p1 + j));
    }
    if constexpr (kGmm) {
        for (int e = threadIdx.x; e < DT * ncp; e += blockDim.x) {
            const int k = e / ncp, c = e % ncp;
            s_gm[e] = c < n_comp ? __ldg(A.like.p0 + (size_t)c * DT + k) : 0.0;
        }
        for (int c = threadIdx.x; c < ncp; c += blockDim.x) {
            const double is = c < n_comp ? __ldg(A.like.p1 + c) : 0.0;
            s_gw[c] = make_double2(0.5 * is * is, c < n_comp ? __ldg(A.like.p2 + c) : ninf);
        }
    }
    __syncthreads();

    const int lane = (G == 1) ? 0 : (int)(threadIdx.x % G);
    // A plain shared-memory read of a constant is loop-invariant: the compiler hoists all 4 DL of them out of
    // the move loop into registers and spills the state.  lds2 (volatile) keeps each read next to its use.
    double2 rc01[kRC ? DL : 1] = {}, rc23[kRC ? DL : 1] = {}, rpb[kRC && kPBox ? DL : 1] = {}, rlb[kRC && kLBox ? DL : 1] = {};
    double risr[kRC && TPCN ? DL : 1] = {};
    if constexpr (kRC) {
#pragma unroll
        for (int k = 0; k < DL; ++k) {
            if constexpr (TPCN) risr[k] = s_isr[lane * DL + k];
            rc01[k] = s_c01[lane * DL + k];
            rc23[k] = s_c23[lane * DL + k];
            if constexpr (kPBox) rpb[k] = s_pbox[lane * DL + k];
            if constexpr (kLBox) rlb[k] = s_lbox[lane * DL + k];
        }
    }
    auto C01 = [&](int k) -> double2 { if constexpr (kRC) return rc01[k]; else return lds2(&s_c01[lane * DL + k]); };
    auto C23 = [&](int k) -> double2 { if constexpr (kRC) return rc23[k]; else return lds2(&s_c23[lane * DL + k]); };
    auto PBOX = [&](int k) -> double2 { if constexpr (kRC) return rpb[kPBox ? k : 0]; else return lds2(&s_pbox[kPBox ? lane * DL + k : 0]); };
    auto LBOX = [&](int k) -> double2 { if constexpr (kRC) return rlb[kLBox ? k : 0]; else return lds2(&s_lbox[kLBox ? lane * DL + k : 0]); };
    auto ISR = [&](int k) -> double {
        if constexpr (kRC) return risr[TPCN ? k : 0];
        else {
            double r;
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(r) : "r"((uint32_t)__cvta_generic_to_shared(&s_isr[TPCN ? lane * DL + k : 0])));
            return r;
        }
    };
    auto MU = [&](int k) -> double { return C01(k).y; };
    constexpr int kPerWarp = 32 / G;
    const int sub = (threadIdx.x & 31) / G;
    const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t stride = (((int64_t)gridDim.x * blockDim.x) >> 5) * kPerWarp;
    const uint32_t stream0 = A.prm.stream0;
    unsigned int n_acc = 0;

    // separable part of T at centred state v (+ Rosenbrock's coupling term), group-reduced
    auto target = [&](const double(&v)[DL], double& r2_out) -> double {
        double sq[4] = {0.0, 0.0, 0.0, 0.0}, rr[4] = {0.0, 0.0, 0.0, 0.0};
        if constexpr (TPCN) {
            double r2 = 0.0;
#pragma unroll
            for (int k = 0; k < DL; ++k) {
                const double zz = v[k] * ISR(k);
                r2 = fma(zz, zz, r2);
            }
            r2_out = group_sum<G>(r2);
        }
        double xprev = 0.0, xfirst = 0.0;
        bool bad = false;
#pragma unroll
        for (int k = 0; k < DL; ++k) {
            const double2 e = C23(k);
            sq[k & 3] = fma(v[k], fma(e.x, v[k], e.y), sq[k & 3]);
            if constexpr (kNeedX) {
                const double xk = C01(k).y + v[k];
                if constexpr (kRosen) {
                    if (k > 0) {
                        const double tt = fma(-xprev, xprev, xk);
                        rr[k & 3] = fma(tt, tt, rr[k & 3]);
                    } else {
                        xfirst = xk;
                    }
                    xprev = xk;
                }
                if constexpr (kPBox) { const double2 b = PBOX(k); bad |= !(xk > b.x && xk < b.y); }
                if constexpr (kLBox) { const double2 b = LBOX(k); bad |= !(xk > b.x && xk < b.y); }
            }
        }
        double s = (sq[0] + sq[1]) + (sq[2] + sq[3]);
        if constexpr (kRosen) {
            if constexpr (G > 1) {
                const double xn = __shfl_down_sync(0xffffffffu, xfirst, 1, G);
                const double tt = fma(-xprev, xprev, xn);
                if (lane < G - 1) rr[0] = fma(tt, tt, rr[0]);
            }
            s = fma(ros_wa, (rr[0] + rr[1]) + (rr[2] + rr[3]), s);
        }
        if constexpr (kPBox || kLBox) {
            if (bad) s = ninf;  // -inf target: the move is rejected
        }
        double T = group_sum<G>(s);
        if constexpr (kGmm) T = fma(beta, tpp_gmm<DL, G>(v, MU, lane, s_gm, s_gw, n_comp, ncp), T);
        if constexpr (TPCN) T = fma(t_half, log1p(r2_out * inv_nu), T);
        return T;
    };

    // Row staging.  The resampling gather reads random 8 DT-byte rows and the dependent chain idx -> row
    // costs two DRAM latencies that 8 warps per SM cannot hide (measured: 5.6 ms for ONE move at N = 2e7,
    // 0.8 ms per further move).  Each warp therefore owns two shared-memory buffers of 32 / G rows: the rows
    // of batch b+1 are fetched with cp.async (16-byte chunks, whole rows per instruction, no registers)
    // while batch b moves, and the ancestor indices of batch b+2 are loaded at the same time.  The finished
    // rows go back through the same buffer so that global stores are row-contiguous too.
    constexpr int kChunks = DT / 2;                 // 16-byte chunks per row
    constexpr int kRowStride = DT * 8 + 16;         // +16: rows land in different banks for the per-thread reads
    constexpr int kStageBytes = kPerWarp * kRowStride;
    constexpr int kCopyIters = (kPerWarp * kChunks) / 32;  // = DL / 2
    unsigned char* stage0 = reinterpret_cast<unsigned char*>(s_dyn) + tpp_gmm_bytes(DT, n_comp) +
                            (size_t)(threadIdx.x >> 5) * 2 * kStageBytes;
    const int wl = threadIdx.x & 31;
    auto src_of = [&](int64_t b) -> int64_t {  // ancestor row of this thread's particle in the batch at `b`
        const int64_t ii = b + sub < A.n ? b + sub : A.n - 1;
        return A.idx ? A.idx[ii] : ii;
    };
    auto pid_of = [&](int64_t b) -> uint64_t {  // Philox counter (global output slot) of this thread's particle
        const int64_t ii = b + sub < A.n ? b + sub : A.n - 1;
        return (uint64_t)(A.prm.d_slot ? A.prm.d_slot[ii] : A.prm.index0 + ii);
    };
    auto fetch = [&](int buf, int64_t src_row) {  // all 32 lanes: copy the batch's rows into buffer `buf`
        const uint32_t dst0 = (uint32_t)__cvta_generic_to_shared(stage0 + buf * kStageBytes);
#pragma unroll
        for (int it = 0; it < kCopyIters; ++it) {
            const int e = it * 32 + wl, r = e / kChunks, c = e % kChunks;
            const int64_t sr = __shfl_sync(0xffffffffu, src_row, r * G);
            cp_async16(dst0 + r * kRowStride + c * 16, A.x + sr * DT + c * 2);
        }
        cp_async_commit();
    };
    int64_t base = warp_id * kPerWarp;
    int64_t src = 0, src_next = 0;
    uint64_t pidx_next = 0;
    if (base < A.n) {
        src = src_of(base);
        fetch(0, src);
        src_next = src_of(base + stride);
        pidx_next = pid_of(base);
    }
    for (int bi = 0; base < A.n; base += stride, ++bi) {
        const int64_t i_raw = base + sub;
        const bool active = i_raw < A.n;
        const int64_t i = active ? i_raw : A.n - 1;
        unsigned char* my_rows = stage0 + (bi & 1) * kStageBytes;
        const double* row = reinterpret_cast<const double*>(my_rows + sub * kRowStride) + lane * DL;
        cp_async_wait_all();
        __syncwarp();
        if (base + stride < A.n) fetch((bi + 1) & 1, src_next);  // overlaps this batch's moves
        const int64_t src_next2 = src_of(base + 2 * stride);
        const uint64_t pidx = pidx_next;  // loaded one batch ahead, like the ancestor indices
        pidx_next = pid_of(base + stride);
        double y[DL], yp[DL];
#pragma unroll
        for (int k = 0; k < DL; k += 2) {
            const double2 v = *reinterpret_cast<const double2*>(row + k);
            y[k] = v.x - C01(k).y;
            y[k + 1] = v.y - C01(k + 1).y;
        }
        double r2c = 0.0;  // tpCN: |(x - mu)/sigma|^2 of the current state
        double cur = target(y, r2c);
        if (!active) cur = identity(OP_MIN);  // +inf: a padding particle never accepts
        // from here on y holds ca * (x - mu): the proposal is then one DFMA per dimension and accepting a
        // move is one DMUL per dimension (instead of a DMUL + DFMA and two register moves)
#pragma unroll
        for (int k = 0; k < DL; ++k) y[k] *= ca;
        bool moved = false;
        uint4 aw = make_uint4(0, 0, 0, 0);
        for (int t = 0; t < n_moves; ++t) {
            double log_u;
            if constexpr (PHILOX) {
                if ((t & 3) == 0) aw = accept_words(A.keys, stream0, t, pidx);
                log_u = accept_log_u(aw, t);
            } else {
                log_u = A.prm.d_log_u[(size_t)t * A.n + i];
            }
            double stot = 0.0, rtot = 0.0, r2p = 0.0;
            double xprev = 0.0, xfirst = 0.0;
            bool bad = false;
            double sc = 1.0;
            if constexpr (TPCN) {
                double g;
                if constexpr (PHILOX) g = tpp_gamma_mt(A.keys, stream0 + (uint32_t)t, pidx, gamma_a);
                else g = A.prm.d_gamma[(size_t)t * A.n + i];
                sc = sqrt((nu + r2c) / (2.0 * g));
            }
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                double xi[4];
                if constexpr (PHILOX) {
                    float z[4];
                    philox_normal4(A.keys, stream0 + (uint32_t)t, pidx, (uint32_t)(lane * NQ + q), z);
#pragma unroll
                    for (int j = 0; j < 4; ++j) xi[j] = (double)z[j];
                } else {
                    const double* p = A.prm.d_xi + ((size_t)t * A.n + i) * DT + lane * DL + q * 4;
                    const double2 a = ldg2(p), b = ldg2(p + 2);
                    xi[0] = a.x; xi[1] = a.y; xi[2] = b.x; xi[3] = b.y;
                }
                double sq = 0.0, rr = 0.0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int k = q * 4 + j;
                    const double2 c = C01(k);
                    const double v = TPCN ? fma(c.x * sc, xi[j], y[k]) : fma(c.x, xi[j], y[k]);
                    yp[k] = v;
                    const double2 e = C23(k);
                    sq = fma(v, fma(e.x, v, e.y), sq);
                    if constexpr (TPCN) {
                        const double zz = v * ISR(k);
                        r2p = fma(zz, zz, r2p);
                    }
                    if constexpr (kNeedX) {
                        const double xk = c.y + v;
                        if constexpr (kRosen) {
                            if (k > 0) {
                                const double tt = fma(-xprev, xprev, xk);
                                rr = fma(tt, tt, rr);
                            } else {
                                xfirst = xk;
                            }
                            xprev = xk;
                        }
                        if constexpr (kPBox) { const double2 b = PBOX(k); bad |= !(xk > b.x && xk < b.y); }
                        if constexpr (kLBox) { const double2 b = LBOX(k); bad |= !(xk > b.x && xk < b.y); }
                    }
                }
                stot += sq;
                if constexpr (kRosen) rtot += rr;
            }
            double s = stot;
            if constexpr (kRosen) {
                if constexpr (G > 1) {
                    const double xn = __shfl_down_sync(0xffffffffu, xfirst, 1, G);
                    const double tt = fma(-xprev, xprev, xn);
                    if (lane < G - 1) rtot = fma(tt, tt, rtot);
                }
                s = fma(ros_wa, rtot, s);
            }
            if constexpr (kPBox || kLBox) {
                if (bad) s = ninf;
            }
            double nw = group_sum<G>(s);
            if constexpr (kGmm) nw = fma(beta, tpp_gmm<DL, G>(yp, MU, lane, s_gm, s_gw, n_comp, ncp), nw);
            if constexpr (TPCN) {
                r2p = group_sum<G>(r2p);
                nw = fma(t_half, log1p(r2p * inv_nu), nw);
            }
            const bool acc = log_u < nw - cur;  // NaN -> reject, like the reference's NaN -> -inf
            if (acc) {
#pragma unroll
                for (int k = 0; k < DL; ++k) y[k] = ca * yp[k];
                cur = nw;
                if constexpr (TPCN) r2c = r2p;
                moved = true;
            }
            if (per_move) {
                const unsigned int votes = __ballot_sync(0xffffffffu, acc && lane == 0);
                if ((threadIdx.x & 31) == 0 && votes) atomicAdd(&s_acc[t], __popc(votes));
            } else {
                n_acc += (acc && lane == 0) ? 1u : 0u;
            }
        }
        // final state and its logs: evaluated once (by every lane: the group shuffles stay converged);
        // a particle that never moved keeps its input row and logs bit for bit
        double x[DL];
        const double inv_ca = 1.0 / ca;  // ca > 0: rho = 1 goes to the general kernel
#pragma unroll
        for (int k = 0; k < DL; ++k) {
            y[k] *= inv_ca;
            x[k] = C01(k).y + y[k];
        }
        double sq_ = 0.0, sp_ = 0.0, sl_ = 0.0, r_ = 0.0;
        bool pbad = false, lbad = false;
#pragma unroll
        for (int k = 0; k < DL; ++k) {
            const int j = lane * DL + k;
            const double2 fq = lds2(&s_fq[j]);
            const double zq = fma(y[k], fq.x, fq.y);
            sq_ = fma(zq, zq, sq_);
            if constexpr (kPriorQuad) {
                const double2 fp = lds2(&s_fp[j]);
                const double z = fma(y[k], fp.x, fp.y);
                sp_ = fma(z, z, sp_);
            } else {
                const double2 b = PBOX(k);
                pbad |= !(x[k] > b.x && x[k] < b.y);
            }
            if constexpr (kLikeQuad) {
                const double2 fl = lds2(&s_fl[j]);
                const double z = fma(y[k], fl.x, fl.y);
                sl_ = fma(z, z, sl_);
            } else if constexpr (kLBox) {
                const double2 b = LBOX(k);
                lbad |= !(x[k] > b.x && x[k] < b.y);
            } else if constexpr (kRosen) {
                double x1;
                if (k + 1 < DL) x1 = x[(k + 1) % DL];
                else x1 = (G > 1) ? __shfl_down_sync(0xffffffffu, x[0], 1, G) : 0.0;
                const double tt = fma(-x[k], x[k], x1);
                const double u = A.like.s1 - x[k];
                if (k + 1 < DL || lane < G - 1) r_ = fma(u, u, fma(A.like.s0 * tt, tt, r_));
            }
        }
        double lq = A.q.s0 - 0.5 * group_sum<G>(sq_);
        double lp, ll;
        if constexpr (kPriorQuad) lp = A.prior.s0 - 0.5 * group_sum<G>(sp_);
        else lp = group_sum<G>(pbad ? 1.0 : 0.0) > 0.0 ? ninf : A.prior.s0;
        if constexpr (kLikeQuad) ll = A.like.s0 - 0.5 * group_sum<G>(sl_);
        else if constexpr (kLBox) ll = group_sum<G>(lbad ? 1.0 : 0.0) > 0.0 ? ninf : A.like.s0;
        else if constexpr (kRosen) ll = -group_sum<G>(r_) / A.like.s2;
        else ll = tpp_gmm<DL, G>(y, MU, lane, s_gm, s_gw, n_comp, ncp);
        if (!moved) {
#pragma unroll
            for (int k = 0; k < DL; k += 2) {
                const double2 v = *reinterpret_cast<const double2*>(row + k);
                x[k] = v.x;
                x[k + 1] = v.y;
            }
            ll = A.ll[src];
            lp = A.lp[src];
            lq = A.lq[src];
        }
        // the finished rows go out through the staging buffer: whole rows per store instruction
        {
            double* mine = reinterpret_cast<double*>(my_rows + sub * kRowStride) + lane * DL;
#pragma unroll
            for (int k = 0; k < DL; k += 2) *reinterpret_cast<double2*>(mine + k) = make_double2(x[k], x[k + 1]);
            __syncwarp();
#pragma unroll
            for (int it = 0; it < kCopyIters; ++it) {
                const int e = it * 32 + wl, r = e / kChunks, c = e % kChunks;
                if (base + r < A.n) {
                    const double2 v = *reinterpret_cast<const double2*>(my_rows + r * kRowStride + c * 16);
                    st_stream2(A.xo + (base + r) * DT + c * 2, v.x, v.y);
                }
            }
            __syncwarp();  // the buffer is refilled by the fetch issued at the top of the next batch
        }
        if (active && lane == 0) {
            A.llo[i] = ll;
            A.lpo[i] = lp;
            A.lqo[i] = lq;
            if (lq != lq) atomicAdd(&A.accepts[per_move ? n_moves : 1], 1ull);  // NaN log q (smc/minipcn.py:133)
        }
        src = src_next;
        src_next = src_next2;
    }
    if (!per_move) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) n_acc += __shfl_xor_sync(0xffffffffu, n_acc, o);
        if ((threadIdx.x & 31) == 0 && n_acc) atomicAdd(&s_acc[0], n_acc);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < (per_move ? n_moves : 1); t += blockDim.x)
        if (s_acc[t]) atomicAdd(&A.accepts[t], (unsigned long long)s_acc[t]);
}

// dims the kernel is built for -> (dims per lane DL, lanes per particle G).  Measured on B200 (N = 2e7,
// 10 moves, profiles/r1_sweep_variants.md): one thread per particle up to 32 dims beats the split
// layouts - (32,1) 10.8 ms vs (8,4) 12.9 ms vs (16,2) 12.6 ms at D = 32; (16,1) 4.8 ms vs (8,2) 5.5 ms at D = 16.
inline bool tpp_shape(int dims, int like_kind, int& dl, int& g) {
    (void)like_kind;
    switch (dims) {
        case 4: case 8: case 16: case 32: dl = dims; g = 1; return true;
        case 64: dl = 32; g = 2; return true;
        default: return false;
    }
}

template <int DL, int G>
inline size_t tpp_dyn_smem(const SweepArgs& A) {
    const size_t stage = (size_t)(kTppBlock / 32) * 2 * (32 / G) * (DL * G * 8 + 16);
    return tpp_gmm_bytes(A.dims, A.like.kind == ASP_DENSITY_GMM ? A.like.n_comp : 0) + stage;
}

template <int DL, int G, int LIKE, int PRIOR, bool TPCN>
inline int launch_tpp_n(const SweepArgs& A, cudaStream_t s) {
    const int64_t per_block = (int64_t)kTppBlock / G;
    const int64_t need = (A.n + per_block - 1) / per_block;
    const int64_t cap = (int64_t)kNumSMs * tpp_min_blocks<DL>();
    const int grid = (int)(need < cap ? need : cap);
    const size_t smem = tpp_dyn_smem<DL, G>(A);
    auto launch = [&](auto kernel) -> int {
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return ASP_ERR_CUDA;
        kernel<<<grid, kTppBlock, smem, s>>>(A);
        return ASP_OK;
    };
    if (A.prm.noise_mode == 0) return launch(pcn_sweep_tpp_kernel<DL, G, LIKE, PRIOR, true, TPCN>);
    return launch(pcn_sweep_tpp_kernel<DL, G, LIKE, PRIOR, false, TPCN>);
}

template <int LIKE, int PRIOR, bool TPCN>
inline int launch_tpp_shape(const SweepArgs& A, cudaStream_t s) {
    int dl = 0, g = 0;
    if (!tpp_shape(A.dims, LIKE, dl, g)) return ASP_ERR_ARG;
    if (g == 2) return launch_tpp_n<32, 2, LIKE, PRIOR, TPCN>(A, s);
    if (dl == 32) return launch_tpp_n<32, 1, LIKE, PRIOR, TPCN>(A, s);
    if (dl == 16) return launch_tpp_n<16, 1, LIKE, PRIOR, TPCN>(A, s);
    if (dl == 8) return launch_tpp_n<8, 1, LIKE, PRIOR, TPCN>(A, s);
    return launch_tpp_n<4, 1, LIKE, PRIOR, TPCN>(A, s);
}

template <int LIKE, bool TPCN = false>
inline int launch_tpp_like(const SweepArgs& A, cudaStream_t s) {
    if (A.prior.kind == ASP_DENSITY_GAUSSIAN) return launch_tpp_shape<LIKE, ASP_DENSITY_GAUSSIAN, TPCN>(A, s);
    if (A.prior.kind == ASP_DENSITY_UNIFORM) return launch_tpp_shape<LIKE, ASP_DENSITY_UNIFORM, TPCN>(A, s);
    return ASP_ERR_ARG;
}

int launch_tpp_gaussian(const SweepArgs& A, cudaStream_t s);
int launch_tpp_rosenbrock(const SweepArgs& A, cudaStream_t s);
int launch_tpp_gmm(const SweepArgs& A, cudaStream_t s);
int launch_tpp_uniform(const SweepArgs& A, cudaStream_t s);
// tpCN (params->nu > 0) instantiations: pcn_tpp_t_*.cu
int launch_tpp_t_gaussian(const SweepArgs& A, cudaStream_t s);
int launch_tpp_t_rosenbrock(const SweepArgs& A, cudaStream_t s);
int launch_tpp_t_gmm(const SweepArgs& A, cudaStream_t s);
int launch_tpp_t_uniform(const SweepArgs& A, cudaStream_t s);

}  // namespace asp

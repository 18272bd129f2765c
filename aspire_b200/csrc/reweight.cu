// K2/K3/K4/K11: reweighting, ESS sums, device-driven beta bisection, resampling
// weights + evidence bookkeeping, importance weights.  All HBM-bound streaming
// reductions: coalesced fp64 loads, 4 independent loads in flight per thread,
// warp-shuffle + shared-memory block reduce, deterministic last-block grid reduce.
//
// Reference arithmetic replaced (paths under /root/reference):
//   src/aspire/samples.py:1221-1249  unnormalized_log_weights / log_weights
//   src/aspire/utils.py:248-255,510-512  logsumexp / effective_sample_size
//   src/aspire/samplers/smc/base.py:166-186  determine_beta bisection
//   src/aspire/samples.py:1226-1242  log_evidence_ratio[_variance]
//   src/aspire/samples.py:457-475    Samples.compute_weights
#include <math.h>

#include "beta_math.cuh"
#include "common.cuh"

namespace asp {

constexpr int kUnroll = 4;

// ---------------------------------------------------------------------------
// K2: d = (logL + logP) - logq, with min / max / NaN / +inf statistics
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) reweight_kernel(const double* __restrict__ ll, const double* __restrict__ lp,
                                                          const double* __restrict__ lq, int64_t n,
                                                          double* __restrict__ d, double* __restrict__ stats,
                                                          ReduceWs* ws) {
    __shared__ double smem[4 * 32];
    double v[4] = {identity(OP_MIN), identity(OP_MAX), 0.0, 0.0};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (kUnroll - 1) * stride < n; i += kUnroll * stride) {
        double a[kUnroll], b[kUnroll], c[kUnroll];
#pragma unroll
        for (int k = 0; k < kUnroll; ++k) {
            a[k] = ld_stream(ll + i + k * stride);
            b[k] = ld_stream(lp + i + k * stride);
            c[k] = ld_stream(lq + i + k * stride);
        }
#pragma unroll
        for (int k = 0; k < kUnroll; ++k) {
            double t = (a[k] + b[k]) - c[k];
            d[i + k * stride] = t;
            v[0] = fmin(v[0], t);
            v[1] = fmax(v[1], t);
            v[2] += (t != t) ? 1.0 : 0.0;
            v[3] += (t == identity(OP_MIN)) ? 1.0 : 0.0;
        }
    }
    for (; i < n; i += stride) {
        double t = (ll[i] + lp[i]) - lq[i];
        d[i] = t;
        v[0] = fmin(v[0], t);
        v[1] = fmax(v[1], t);
        v[2] += (t != t) ? 1.0 : 0.0;
        v[3] += (t == identity(OP_MIN)) ? 1.0 : 0.0;
    }
    const int ops[4] = {OP_MIN, OP_MAX, OP_SUM, OP_SUM};
    grid_reduce<4>(v, ops, ws, smem, [&](double(&r)[4]) {
        stats[0] = r[0];
        stats[1] = r[1];
        stats[2] = r[2];
        stats[3] = r[3];
    });
}

// ---------------------------------------------------------------------------
// K3: sums for ESS at NB candidate betas in one pass
// ---------------------------------------------------------------------------
template <int NB>
__global__ void __launch_bounds__(kBlock) ess_sums_kernel(const double* __restrict__ d, int64_t n,
                                                          const double* __restrict__ dbeta,
                                                          const double* __restrict__ stats, double* __restrict__ sums,
                                                          ReduceWs* ws) {
    __shared__ double smem[2 * NB * 32];
    double db[NB], c[NB], v[2 * NB];
    int ops[2 * NB];
#pragma unroll
    for (int k = 0; k < NB; ++k) {
        db[k] = dbeta[k];
        c[k] = db[k] * (db[k] >= 0.0 ? stats[1] : stats[0]);
        v[2 * k] = v[2 * k + 1] = 0.0;
        ops[2 * k] = ops[2 * k + 1] = OP_SUM;
    }
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (kUnroll - 1) * stride < n; i += kUnroll * stride) {
        double x[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) x[u] = ld_stream(d + i + u * stride);
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
#pragma unroll
            for (int k = 0; k < NB; ++k) {
                double e = exp(db[k] * x[u] - c[k]);
                v[2 * k] += e;
                v[2 * k + 1] += e * e;
            }
        }
    }
    for (; i < n; i += stride) {
        double x = d[i];
#pragma unroll
        for (int k = 0; k < NB; ++k) {
            double e = exp(db[k] * x - c[k]);
            v[2 * k] += e;
            v[2 * k + 1] += e * e;
        }
    }
    grid_reduce<2 * NB>(v, ops, ws, smem, [&](double(&r)[2 * NB]) {
#pragma unroll
        for (int k = 0; k < 2 * NB; ++k) sums[k] = r[k];
    });
}

// ---------------------------------------------------------------------------
// All-reduce (sum) of a few doubles across the GPUs of one box INSIDE the reducing kernel, over peer memory
// (NVLink loads / stores into mailboxes the ranks exchanged by CUDA IPC): the last block of every rank's
// pass publishes its partial sums to every mailbox, waits for the others and adds them in rank order, so all
// ranks hold bit-identical totals and take the same bisection decisions - no NCCL call, no extra launch, no
// host round trip between a pass and its decision.  Mailbox layout: [parity][writer rank][16 doubles],
// value 15 = epoch flag; two parities because a rank may start pass e+1 while a slow peer still reads pass e
// (it cannot reach e+2 before every peer has finished reading e).  The epoch is a device-side counter kept
// behind the rank's own mailbox and advanced by the kernel itself, so passes that return early (search
// already converged - the same passes on every rank, their states being bit-identical) do not consume epochs.
// ---------------------------------------------------------------------------
constexpr int kBoxStride = 16;
constexpr unsigned long long kPeerWaitNs = 20000000000ull;  // 20 s of %globaltimer: far above any launch skew
template <int NV>
__device__ __forceinline__ bool p2p_allreduce_sum(double (&v)[NV], const asp_peer_t& P) {
    static_assert(NV < kBoxStride, "mailbox slot too small");
    double* counter = P.box[P.rank] + 2 * ASP_MAX_PEERS * kBoxStride;
    const double flag = *counter + 1.0;
    *counter = flag;
    const size_t par = (size_t)((long long)flag & 1) * ASP_MAX_PEERS * kBoxStride;
    for (int p = 0; p < P.world; ++p) {
        volatile double* slot = P.box[p] + par + (size_t)P.rank * kBoxStride;
#pragma unroll
        for (int k = 0; k < NV; ++k) slot[k] = v[k];
    }
    __threadfence_system();
    for (int p = 0; p < P.world; ++p) P.box[p][par + (size_t)P.rank * kBoxStride + 15] = flag;
    __threadfence_system();
    bool ok = true;
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = 0.0;
    for (int w = 0; w < P.world; ++w) {
        volatile double* slot = P.box[P.rank] + par + (size_t)w * kBoxStride;
        unsigned long long t0 = 0, t1;
        while (slot[15] != flag) {  // bounded: a peer that has not arrived in time is reported, never waited for
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (!t0) t0 = t1;
            if (t1 - t0 > kPeerWaitNs) {
                ok = false;
                break;
            }
        }
        __threadfence_system();
#pragma unroll
        for (int k = 0; k < NV; ++k) acc[k] += slot[k];
    }
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = acc[k];
    return ok;
}

// ---------------------------------------------------------------------------
// device-driven bisection
// ---------------------------------------------------------------------------
__global__ void beta_init_kernel(asp_beta_state_t* st, double beta_prev, double target_eff, double tol, double n_total,
                                 const double* stats) {
    beta_init_state(st, beta_prev, target_eff, tol, n_total, stats[0], stats[1]);
}

__global__ void beta_decide_kernel(asp_beta_state_t* st) { beta_decide(st); }

__global__ void __launch_bounds__(kBlock) beta_pass_kernel(const double* __restrict__ d, int64_t n,
                                                           asp_beta_state_t* st, int fused_decide, ReduceWs* ws,
                                                           const asp_peer_t P) {
    __shared__ double smem[2 * 32];
    if (st->phase == 2 || st->reserved == -1) return;  // uniform across the grid: both only change in the last block's epilogue (-1: a peer timed out, the host raises)
    const double db = st->beta_try - st->beta_prev;
    const double c = db * (db >= 0.0 ? st->d_max : st->d_min);
    double v[2] = {0.0, 0.0};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (kUnroll - 1) * stride < n; i += kUnroll * stride) {
        double x[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) x[u] = ld_stream(d + i + u * stride);
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            double e = exp(db * x[u] - c);
            v[0] += e;
            v[1] += e * e;
        }
    }
    for (; i < n; i += stride) {
        double e = exp(db * d[i] - c);
        v[0] += e;
        v[1] += e * e;
    }
    const int ops[2] = {OP_SUM, OP_SUM};
    grid_reduce<2>(v, ops, ws, smem, [&](double(&r)[2]) {
        if (P.world > 1 && !p2p_allreduce_sum<2>(r, P)) st->reserved = -1;
        st->s1 = r[0];
        st->s2 = r[1];
        if (fused_decide) beta_decide(st);
    });
}

// ---------------------------------------------------------------------------
// Three bisection levels per pass.  The seven betas a 3-level bisection of [beta_min, beta_max] can visit
// are beta_min + k h, h = (beta_max - beta_min) / 8, so their weights are E r^k with
// E = exp((beta_min - beta_prev)(d - d_max)),  r = exp(h (d - d_max)):  two exps and seven multiplies per
// particle instead of three dependent passes (each one exp).  The decisions are the reference's
// (smc/base.py:177-185: beta_try = 0.5 (max + min), ESS/N >= target), replayed on the seven (S1, S2) pairs.
// ---------------------------------------------------------------------------
// kBetaLevels, kBetaCand, beta_decide3: beta_math.cuh
__global__ void beta_decide3_kernel(asp_beta_state_t* st, const double* sums) { beta_decide3(st, sums); }

__global__ void __launch_bounds__(kBlock, 2) beta_pass3_kernel(const double* __restrict__ d, int64_t n,
                                                            asp_beta_state_t* st, double* __restrict__ sums,
                                                            int fused_decide, ReduceWs* ws, const asp_peer_t P) {
    __shared__ double smem[2 * kBetaCand * 32];
    if (st->phase != 1 || st->reserved == -1) return;  // uniform across the grid
    const double d_max = st->d_max;
    const double db0 = st->beta_min - st->beta_prev;
    const double h = (st->beta_max - st->beta_min) * 0.125;
    double v[2 * kBetaCand];
    int ops[2 * kBetaCand];
#pragma unroll
    for (int k = 0; k < 2 * kBetaCand; ++k) {
        v[k] = 0.0;
        ops[k] = OP_SUM;
    }
    auto accumulate = [&](double x) {
        const double t = x - d_max;  // <= 0: every factor below is in [0, 1]
        // db0 == 0 on the first passes: 0 * (-inf) would be NaN for a particle of zero weight (d = -inf)
        const double e0 = db0 > 0.0 ? exp(db0 * t) : 1.0, r = exp(h * t);
        double e = e0;
#pragma unroll
        for (int k = 0; k < kBetaCand; ++k) {
            e *= r;
            v[2 * k] += e;
            v[2 * k + 1] = fma(e, e, v[2 * k + 1]);
        }
    };
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (kUnroll - 1) * stride < n; i += kUnroll * stride) {
        double x[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) x[u] = ld_stream(d + i + u * stride);
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) accumulate(x[u]);
    }
    for (; i < n; i += stride) accumulate(d[i]);
    grid_reduce<2 * kBetaCand>(v, ops, ws, smem, [&](double(&r)[2 * kBetaCand]) {
        if (P.world > 1 && !p2p_allreduce_sum<2 * kBetaCand>(r, P)) st->reserved = -1;
#pragma unroll
        for (int k = 0; k < 2 * kBetaCand; ++k) sums[k] = r[k];
        if (fused_decide) beta_decide3(st, sums);
    });
}

// ---------------------------------------------------------------------------
// Root-guided search (beta_math.cuh): subsample estimate -> Newton passes -> one node pass with the reference's
// formula.  The node pass costs what one three-level pass costs and replaces seven of them.
// ---------------------------------------------------------------------------
constexpr int kSubBlock = 1024;
constexpr int kSubSamples = 32768;  // fp32 (d - d_max) of every (n / kSubSamples)-th particle, in shared memory
constexpr int kSubLevels = 26;

// One block: bisection of ESS/N = target on the subsample, in fp32 (an ESTIMATE: the node pass decides).
__global__ void __launch_bounds__(kSubBlock) beta_sub_kernel(const double* __restrict__ d, int64_t n,
                                                             const asp_beta_state_t* __restrict__ st,
                                                             double* __restrict__ scratch, const asp_peer_t P) {
    extern __shared__ float s_t[];  // [m]
    __shared__ float s_red[2 * 32];
    __shared__ float s_sum[2];
    if (st->phase != 1) return;
    const int64_t stride = n / kSubSamples > 0 ? n / kSubSamples : 1;
    const int m = (int)((n + stride - 1) / stride < kSubSamples ? (n + stride - 1) / stride : kSubSamples);
    const double d_max = st->d_max;
    float t_max = -INFINITY;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const float t = (float)(d[(int64_t)i * stride] - d_max);
        s_t[i] = t;
        t_max = fmaxf(t_max, t);
    }
    // shift by the SAMPLE's maximum: with the global one every sampled weight can underflow in fp32 at the first,
    // large trial steps (0 >= target * 0 would then read as "ESS large enough")
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t_max = fmaxf(t_max, __shfl_xor_sync(0xffffffffu, t_max, o));
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = t_max;
    __syncthreads();
    t_max = s_red[0];
    for (int w = 1; w < kSubBlock / 32; ++w) t_max = fmaxf(t_max, s_red[w]);
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += blockDim.x) s_t[i] -= t_max;  // -inf stays -inf
    __syncthreads();
    const float target = (float)st->target_eff * (float)m;
    float lo = 0.0f, hi = (float)(1.0 - st->beta_prev);  // ESS(lo) >= target > ESS(hi)
    for (int level = 0; level < kSubLevels; ++level) {
        const float mid = 0.5f * (lo + hi);
        if (!(mid > lo && mid < hi)) break;
        float a = 0.0f, b = 0.0f;
        const float c = mid * 1.4426950408889634f;
        for (int i = threadIdx.x; i < m; i += blockDim.x) {
            const float e = exp2f(c * s_t[i]);  // t <= 0; -inf -> 0
            a += e;
            b = fmaf(e, e, b);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
        }
        if ((threadIdx.x & 31) == 0) {
            s_red[threadIdx.x >> 5] = a;
            s_red[32 + (threadIdx.x >> 5)] = b;
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            a = s_red[threadIdx.x];
            b = s_red[32 + threadIdx.x];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                a += __shfl_xor_sync(0xffffffffu, a, o);
                b += __shfl_xor_sync(0xffffffffu, b, o);
            }
            if (threadIdx.x == 0) {
                s_sum[0] = a;
                s_sum[1] = b;
            }
        }
        __syncthreads();
        if (s_sum[1] > 0.0f && s_sum[0] * s_sum[0] >= target * s_sum[1]) lo = mid;
        else hi = mid;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double v[1] = {0.5 * ((double)lo + (double)hi)};
        bool ok = true;
        if (P.world > 1) {  // every rank starts Newton from the same point: the mean of the ranks' estimates
            ok = p2p_allreduce_sum<1>(v, P);
            v[0] /= (double)P.world;
        }
        if (!(v[0] > 0.0)) v[0] = 0.5 * (1.0 - st->beta_prev);
        scratch[kBsRoot] = ok ? st->beta_prev + v[0] : -1.0;
        scratch[kBsStep] = 1.0;
        scratch[kBsTries] = ok ? 0.0 : 3.0;  // a missing peer: skip to the plain bisection (which reports it)
    }
}

// S1, S2 and their derivatives at the current root estimate; the last block (after the optional all-reduce) takes
// the Newton step.  `polish`: skipped when the previous step already left less than a fifth of a tolerance.
__global__ void __launch_bounds__(kBlock, 4) beta_newton_kernel(const double* __restrict__ d, int64_t n,
                                                                const asp_beta_state_t* __restrict__ st,
                                                                double* __restrict__ scratch, int polish, ReduceWs* ws,
                                                                const asp_peer_t P) {
    __shared__ double smem[4 * 32];
    if (st->phase != 1 || scratch[kBsTries] >= 3.0) return;  // uniform across the grid
    const double delta = scratch[kBsRoot] - st->beta_prev;
    if (polish && scratch[kBsStep] * scratch[kBsStep] < 0.2 * st->tol * delta) return;
    const double d_max = st->d_max;
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    auto accumulate = [&](double x) {
        const double t = x - d_max;
        const double e = exp(delta * t);
        const double te = e > 0.0 ? t * e : 0.0;  // d = -inf: zero weight, zero derivative
        v[0] += e;
        v[1] = fma(e, e, v[1]);
        v[2] += te;
        v[3] = fma(te, e, v[3]);
    };
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (kUnroll - 1) * stride < n; i += kUnroll * stride) {
        double x[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) x[u] = ld_stream(d + i + u * stride);
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) accumulate(x[u]);
    }
    for (; i < n; i += stride) accumulate(d[i]);
    const int ops[4] = {OP_SUM, OP_SUM, OP_SUM, OP_SUM};
    grid_reduce<4>(v, ops, ws, smem, [&](double(&r)[4]) {
        if (P.world > 1 && !p2p_allreduce_sum<4>(r, P)) {
            scratch[kBsTries] = 3.0;
            return;
        }
        beta_newton_update(st, scratch, r[0], r[1], r[2], r[3]);
    });
}

// The node pass: sums at a + k h, k = 0..8, of the tree node around the root estimate, weights E r^k as in the
// three-level pass; the last block decides (beta_node_decide).
constexpr int kNodeBlocksPerSm = 4;
__global__ void __launch_bounds__(kBlock, 2) beta_node_kernel(const double* __restrict__ d, int64_t n,
                                                              asp_beta_state_t* st, double* __restrict__ scratch,
                                                              ReduceWs* ws, const asp_peer_t P) {
    __shared__ double smem[2 * kBsNode * 32];
    if (st->phase != 1 || scratch[kBsTries] >= 3.0 || st->reserved == -1) return;  // uniform across the grid
    double a, b;
    int depth;
    beta_node(st, scratch[kBsRoot], a, b, depth);
    const double d_max = st->d_max;
    const double db0 = a - st->beta_prev;
    const double h = (b - a) * 0.125;
    double v[2 * kBsNode];
    int ops[2 * kBsNode];
#pragma unroll
    for (int k = 0; k < 2 * kBsNode; ++k) {
        v[k] = 0.0;
        ops[k] = OP_SUM;
    }
    auto accumulate = [&](double x) {
        const double t = x - d_max;  // <= 0
        double e = db0 > 0.0 ? exp(db0 * t) : 1.0;  // db0 == 0: 0 * (-inf) would be NaN for a zero-weight particle
        const double r = exp(h * t);
        v[0] += e;
        v[1] = fma(e, e, v[1]);
#pragma unroll
        for (int k = 1; k < kBsNode; ++k) {
            e *= r;
            v[2 * k] += e;
            v[2 * k + 1] = fma(e, e, v[2 * k + 1]);
        }
    };
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (kUnroll - 1) * stride < n; i += kUnroll * stride) {
        double x[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) x[u] = ld_stream(d + i + u * stride);
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) accumulate(x[u]);
    }
    for (; i < n; i += stride) accumulate(d[i]);
    grid_reduce<2 * kBsNode>(v, ops, ws, smem, [&](double(&r)[2 * kBsNode]) {
        if (P.world > 1) {  // two mailbox rounds: a slot holds 15 values
            double lo[14], hi[4];
#pragma unroll
            for (int k = 0; k < 14; ++k) lo[k] = r[k];
#pragma unroll
            for (int k = 0; k < 4; ++k) hi[k] = r[14 + k];
            const bool ok = p2p_allreduce_sum<14>(lo, P) && p2p_allreduce_sum<4>(hi, P);
            if (!ok) {
                st->reserved = -1;
                return;
            }
#pragma unroll
            for (int k = 0; k < 14; ++k) r[k] = lo[k];
#pragma unroll
            for (int k = 0; k < 4; ++k) r[14 + k] = hi[k];
        }
#pragma unroll
        for (int k = 0; k < 2 * kBsNode; ++k) scratch[k] = r[k];
        beta_node_decide(st, scratch);
    });
}

// ---------------------------------------------------------------------------
// resampling weights, evidence-variance sum, exact-scan tile statistics
// one block handles one ASP_CDF_TILE (=2048) tile per iteration
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) smc_weights_kernel(const double* __restrict__ d, int64_t n, double dbeta,
                                                             double c, double s1, double n_total,
                                                             double* __restrict__ p, double* __restrict__ result,
                                                             double* __restrict__ tile_stats, ReduceWs* ws) {
    __shared__ double smem[2 * 32];
    constexpr int kItems = ASP_CDF_TILE / kBlock;  // 8
    const double mean = s1 / n_total;
    const int64_t n_tiles = (n + ASP_CDF_TILE - 1) / ASP_CDF_TILE;
    double v[2] = {0.0, 0.0};
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t base = tile * ASP_CDF_TILE;
        double x[kItems];
#pragma unroll
        for (int k = 0; k < kItems; ++k) {
            int64_t i = base + k * kBlock + threadIdx.x;
            x[k] = i < n ? ld_stream(d + i) : identity(OP_MAX);  // -inf -> weight 0
        }
        double t[2] = {0.0, 0.0};  // tile sum, tile max
#pragma unroll
        for (int k = 0; k < kItems; ++k) {
            int64_t i = base + k * kBlock + threadIdx.x;
            double u = exp(dbeta * x[k] - c);
            double w = u / s1;
            if (i < n) {
                if (p) p[i] = w;
                double dv = u - mean;
                v[0] += dv * dv;
                v[1] += w;
                t[0] += w;
                t[1] = fmax(t[1], w);
            }
        }
        if (tile_stats) {
            const int tops[2] = {OP_SUM, OP_MAX};
            block_reduce<2>(t, tops, smem);
            if (threadIdx.x == 0) {
                tile_stats[2 * tile] = t[0];
                tile_stats[2 * tile + 1] = t[1];
            }
        }
    }
    const int ops[2] = {OP_SUM, OP_SUM};
    grid_reduce<2>(v, ops, ws, smem, [&](double(&r)[2]) {
        result[0] = r[0];
        result[1] = r[1];
        result[2] = 0.0;
        result[3] = 0.0;
    });
}

// ---------------------------------------------------------------------------
// K11: importance weights (three stream-ordered passes)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) imp_logw_kernel(const double* __restrict__ ll, const double* __restrict__ lp,
                                                          const double* __restrict__ lq, int64_t n,
                                                          double* __restrict__ log_w, double* __restrict__ result,
                                                          ReduceWs* ws) {
    __shared__ double smem[2 * 32];
    double v[2] = {identity(OP_MAX), 0.0};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double t = ll[i] + lp[i] - lq[i];  // samples.py:459
        log_w[i] = t;
        v[0] = fmax(v[0], t);
        v[1] += (t != t) ? 1.0 : 0.0;
    }
    const int ops[2] = {OP_MAX, OP_SUM};
    grid_reduce<2>(v, ops, ws, smem, [&](double(&r)[2]) {
        result[5] = r[0];
        result[6] = r[1];
        result[7] = 0.0;
    });
}

__global__ void __launch_bounds__(kBlock) imp_sums_kernel(const double* __restrict__ log_w, int64_t n,
                                                          double* __restrict__ weights, double* __restrict__ result,
                                                          ReduceWs* ws) {
    __shared__ double smem[2 * 32];
    const double m = result[5];
    double v[2] = {0.0, 0.0};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double lw = log_w[i];
        double e = exp(lw - m);
        weights[i] = exp(lw);  // samples.py:463 (unshifted)
        v[0] += e;
        v[1] += e * e;
    }
    const int ops[2] = {OP_SUM, OP_SUM};
    grid_reduce<2>(v, ops, ws, smem, [&](double(&r)[2]) {
        double log_z = (m + log(r[0])) - log((double)n);  // samples.py:460-462
        result[0] = log_z;
        result[1] = exp(log_z);
        result[4] = r[0] * r[0] / r[1];  // samples.py:472-475
    });
}

__global__ void __launch_bounds__(kBlock) imp_err_kernel(const double* __restrict__ weights, int64_t n,
                                                         double* __restrict__ result, ReduceWs* ws) {
    __shared__ double smem[32];
    const double z = result[1];
    double v[1] = {0.0};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double t = weights[i] - z;
        v[0] += t * t;
    }
    const int ops[1] = {OP_SUM};
    grid_reduce<1>(v, ops, ws, smem, [&](double(&r)[1]) {
        double nn = (double)n;
        double err = sqrt(r[0] / (nn * (nn - 1.0)));  // samples.py:465-467
        result[2] = err;
        result[3] = fabs(err / z);  // samples.py:468-470
    });
}

}  // namespace asp

using namespace asp;

extern "C" {

int asp_reweight(const double* d_log_l, const double* d_log_p, const double* d_log_q, int64_t n, double* d_d,
                 double* d_stats, void* ws, void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(d_log_l && d_log_p && d_log_q && d_d && d_stats && ws, "null pointer");
    ASP_REQUIRE(n > 0, "n must be positive");
    reweight_kernel<<<grid_for(n, kUnroll), kBlock, 0, (cudaStream_t)stream>>>(d_log_l, d_log_p, d_log_q, n, d_d,
                                                                               d_stats, (ReduceWs*)ws);
    return check_launch("asp_reweight");
}

int asp_ess_sums(const double* d_d, int64_t n, const double* d_dbeta, int n_beta, const double* d_stats,
                 double* d_sums, void* ws, void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(d_d && d_dbeta && d_stats && d_sums && ws, "null pointer");
    ASP_REQUIRE(n > 0, "n must be positive");
    ASP_REQUIRE(n_beta >= 1 && n_beta <= 8, "n_beta must be in [1, 8]");
    const int g = grid_for(n, kUnroll);
    cudaStream_t s = (cudaStream_t)stream;
    ReduceWs* w = (ReduceWs*)ws;
    switch (n_beta) {
        case 1: ess_sums_kernel<1><<<g, kBlock, 0, s>>>(d_d, n, d_dbeta, d_stats, d_sums, w); break;
        case 2: ess_sums_kernel<2><<<g, kBlock, 0, s>>>(d_d, n, d_dbeta, d_stats, d_sums, w); break;
        case 3: ess_sums_kernel<3><<<g, kBlock, 0, s>>>(d_d, n, d_dbeta, d_stats, d_sums, w); break;
        case 4: ess_sums_kernel<4><<<g, kBlock, 0, s>>>(d_d, n, d_dbeta, d_stats, d_sums, w); break;
        case 5: ess_sums_kernel<5><<<g, kBlock, 0, s>>>(d_d, n, d_dbeta, d_stats, d_sums, w); break;
        case 6: ess_sums_kernel<6><<<g, kBlock, 0, s>>>(d_d, n, d_dbeta, d_stats, d_sums, w); break;
        case 7: ess_sums_kernel<7><<<g, kBlock, 0, s>>>(d_d, n, d_dbeta, d_stats, d_sums, w); break;
        default: ess_sums_kernel<8><<<g, kBlock, 0, s>>>(d_d, n, d_dbeta, d_stats, d_sums, w); break;
    }
    return check_launch("asp_ess_sums");
}

int asp_beta_search_max_passes(double beta_prev, double tol) {
    double width = 1.0 - beta_prev;
    int k = 1;  // the beta = 1 evaluation
    if (!(tol > 0.0)) return 1 + 1100;
    while (width > tol && k < 1100) {
        width *= 0.5;
        ++k;
    }
    return k + 2;
}

/* passes of asp_beta_search_pass3 after the single beta = 1 pass: three bisection levels each */
int asp_beta_search_max_passes3(double beta_prev, double tol) {
    const int levels = asp_beta_search_max_passes(beta_prev, tol) - 1;
    return (levels + kBetaLevels - 1) / kBetaLevels;
}

int asp_beta_search_init(asp_beta_state_t* d_state, double beta_prev, double target_eff, double tol, double n_total,
                         const double* d_stats, void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(d_state && d_stats, "null pointer");
    beta_init_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(d_state, beta_prev, target_eff, tol, n_total, d_stats);
    return check_launch("asp_beta_search_init");
}

int asp_beta_search_pass(const double* d_d, int64_t n, asp_beta_state_t* d_state, int fused_decide, void* ws,
                         void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(d_d && d_state && ws, "null pointer");
    ASP_REQUIRE(n > 0, "n must be positive");
    asp_peer_t none = {};
    beta_pass_kernel<<<grid_for(n, kUnroll), kBlock, 0, (cudaStream_t)stream>>>(d_d, n, d_state, fused_decide,
                                                                                (ReduceWs*)ws, none);
    return check_launch("asp_beta_search_pass");
}

static int check_peer(const asp_peer_t* peer) {
    ASP_REQUIRE(peer && peer->world >= 2 && peer->world <= ASP_MAX_PEERS && peer->rank >= 0 && peer->rank < peer->world,
                "bad peer descriptor");
    for (int r = 0; r < peer->world; ++r) ASP_REQUIRE(peer->box[r], "null mailbox pointer");
    return ASP_OK;
}

int asp_beta_search_pass_p2p(const double* d_d, int64_t n, asp_beta_state_t* d_state, int levels3, double* d_sums,
                             const asp_peer_t* peer, void* ws, void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(d_d && d_state && ws && (d_sums || !levels3), "null pointer");
    ASP_REQUIRE(n > 0, "n must be positive");
    int rc = check_peer(peer);
    if (rc) return rc;
    if (levels3)
        beta_pass3_kernel<<<grid_for(n, kUnroll), kBlock, 0, (cudaStream_t)stream>>>(d_d, n, d_state, d_sums, 1,
                                                                                     (ReduceWs*)ws, *peer);
    else
        beta_pass_kernel<<<grid_for(n, kUnroll), kBlock, 0, (cudaStream_t)stream>>>(d_d, n, d_state, 1, (ReduceWs*)ws,
                                                                                    *peer);
    return check_launch("asp_beta_search_pass_p2p");
}

int asp_beta_search_pass3(const double* d_d, int64_t n, asp_beta_state_t* d_state, double* d_sums, int fused_decide,
                          void* ws, void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(d_d && d_state && d_sums && ws, "null pointer");
    ASP_REQUIRE(n > 0, "n must be positive");
    asp_peer_t none = {};
    beta_pass3_kernel<<<grid_for(n, kUnroll), kBlock, 0, (cudaStream_t)stream>>>(d_d, n, d_state, d_sums, fused_decide,
                                                                                 (ReduceWs*)ws, none);
    return check_launch("asp_beta_search_pass3");
}

/* Root-guided search (beta_math.cuh): call after the beta = 1 pass.  d_scratch: asp_beta_search_scratch_words()
 * doubles.  peer may be NULL (one GPU). */
int32_t asp_beta_search_scratch_words(void) { return kBsWords; }

int asp_beta_search_root(const double* d_d, int64_t n, asp_beta_state_t* d_state, double* d_scratch,
                         const asp_peer_t* peer, void* ws, void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(d_d && d_state && d_scratch && ws, "null pointer");
    ASP_REQUIRE(n > 0, "n must be positive");
    asp_peer_t P = {};
    if (peer) {
        int rc = check_peer(peer);
        if (rc) return rc;
        P = *peer;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const size_t smem = sizeof(float) * kSubSamples;
    // per device (context) attribute: set on every call rather than cached per process
    if (cudaFuncSetAttribute(beta_sub_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return ASP_ERR_CUDA;
    beta_sub_kernel<<<1, kSubBlock, smem, s>>>(d_d, n, d_state, d_scratch, P);
    const int g = grid_for(n, kUnroll, kBlock, 4);
    beta_newton_kernel<<<g, kBlock, 0, s>>>(d_d, n, d_state, d_scratch, 0, (ReduceWs*)ws, P);
    beta_newton_kernel<<<g, kBlock, 0, s>>>(d_d, n, d_state, d_scratch, 1, (ReduceWs*)ws, P);
    const int gn = grid_for(n, kUnroll, kBlock, kNodeBlocksPerSm);
    for (int attempt = 0; attempt < 3; ++attempt)
        beta_node_kernel<<<gn, kBlock, 0, s>>>(d_d, n, d_state, d_scratch, (ReduceWs*)ws, P);
    return check_launch("asp_beta_search_root");
}

/* The whole common path in one call: init, the beta = 1 pass, the root-guided search.  The caller reads the state
 * back and launches asp_beta_search_pass3 passes only if phase != 2 (three missed nodes). */
int asp_beta_search_run(const double* d_d, int64_t n, const double* d_stats, double beta_prev, double target_eff,
                        double tol, double n_total, asp_beta_state_t* d_state, double* d_scratch,
                        const asp_peer_t* peer, void* ws, void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(d_d && d_stats && d_state && d_scratch && ws, "null pointer");
    ASP_REQUIRE(n > 0, "n must be positive");
    asp_peer_t P = {};
    if (peer) {
        int rc = check_peer(peer);
        if (rc) return rc;
        P = *peer;
    }
    cudaStream_t s = (cudaStream_t)stream;
    beta_init_kernel<<<1, 1, 0, s>>>(d_state, beta_prev, target_eff, tol, n_total, d_stats);
    beta_pass_kernel<<<grid_for(n, kUnroll), kBlock, 0, s>>>(d_d, n, d_state, 1, (ReduceWs*)ws, P);
    return asp_beta_search_root(d_d, n, d_state, d_scratch, peer, ws, stream);
}

int asp_beta_search_decide3(asp_beta_state_t* d_state, const double* d_sums, void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(d_state && d_sums, "null pointer");
    beta_decide3_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(d_state, d_sums);
    return check_launch("asp_beta_search_decide3");
}

int asp_beta_search_decide(asp_beta_state_t* d_state, void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(d_state, "null pointer");
    beta_decide_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(d_state);
    return check_launch("asp_beta_search_decide");
}

int asp_smc_weights(const double* d_d, int64_t n, double dbeta, double c, double s1, double n_total, double* d_p,
                    double* d_result, double* d_tile_stats, void* ws, void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(d_d && d_result && ws, "null pointer");
    ASP_REQUIRE(n > 0, "n must be positive");
    const int64_t n_tiles = (n + ASP_CDF_TILE - 1) / ASP_CDF_TILE;
    int g = (int)(n_tiles < (int64_t)kNumSMs * 8 ? n_tiles : (int64_t)kNumSMs * 8);
    smc_weights_kernel<<<g, kBlock, 0, (cudaStream_t)stream>>>(d_d, n, dbeta, c, s1, n_total, d_p, d_result,
                                                               d_tile_stats, (ReduceWs*)ws);
    return check_launch("asp_smc_weights");
}

size_t asp_importance_workspace_bytes(int64_t n) {
    (void)n;
    return sizeof(ReduceWs);
}

int asp_importance_weights(const double* d_log_l, const double* d_log_p, const double* d_log_q, int64_t n,
                           double* d_log_w, double* d_weights, double* d_result, void* ws, void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(d_log_l && d_log_p && d_log_q && d_log_w && d_weights && d_result && ws, "null pointer");
    ASP_REQUIRE(n > 1, "n must be at least 2");
    cudaStream_t s = (cudaStream_t)stream;
    const int g = grid_for(n, kUnroll);
    imp_logw_kernel<<<g, kBlock, 0, s>>>(d_log_l, d_log_p, d_log_q, n, d_log_w, d_result, (ReduceWs*)ws);
    imp_sums_kernel<<<g, kBlock, 0, s>>>(d_log_w, n, d_weights, d_result, (ReduceWs*)ws);
    imp_err_kernel<<<g, kBlock, 0, s>>>(d_weights, n, d_result, (ReduceWs*)ws);
    return check_launch("asp_importance_weights");
}

}  // extern "C"

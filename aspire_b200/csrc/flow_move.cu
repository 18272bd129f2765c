// C4: the pCN sweep with a coupling flow as proposal density, in ONE kernel - resampling gather, `n_moves` x
// (propose, built-in prior / likelihood, log q(x') through the flow on tcgen05, accept) and the final store.
//
// Replaces the per-move chain of MiniPCNSMC.mutate (/root/reference/src/aspire/samplers/smc/minipcn.py:89-117) against
// the tempered target of smc/base.py:507-519 when `prior_flow` is a learned flow: the reference (and the semi-fused
// route asp_pcn_propose_eval -> asp_flow_logprob_bf16 -> asp_pcn_accept_scalar) streams the particle set through HBM
// three times per move; here a particle is read once and written once per sweep.  Same step spec (oracle/pcn.py),
// same flow arithmetic (flow_core.cuh::coupling_layer, oracle/flow.py) and the same Philox stream layout as the other
// routes, so all of them walk the same trajectory.
//
// One warpgroup = one tile of 128 particles, thread i owns particle i (its TMEM lane, its fp32 flow state in
// registers).  The fp64 state does not fit in registers beside the flow state (512 threads per SM: 128 registers
// each), so the current point lives in a per-warpgroup scratch tile in global memory, chunk-major ([D / 4][128][4]
// doubles: a warp's 32-byte accesses are contiguous) - 38 MB for the whole grid at D = 64, resident in the L2: the
// move loop does not go to HBM (a first version kept the proposal in a second tile: 75 MB of scratch thrashed the L2,
// 17x the algorithmic DRAM traffic - profiles/r2_ncu_flow_move.md).  The proposal is never stored: it is formed chunk
// by chunk in a rolled loop (densities accumulated on the fly), its standardised image is parked in the thread's own
// TMEM lane (tcgen05.st into the idle accumulator columns, read back with tcgen05.ld as the flow state), and an
// accepted proposal is formed again - from its normals, parked in 64 further TMEM columns - by the NEXT pass over the
// scratch tile (the following move's proposal loop, or the final store), on the chunk that pass reads anyway.
// Four warpgroups per SM share one copy of the weights; while one waits on its MMAs the others run their fp64
// proposal / density phase, so the tensor pipe, the XU (Philox / Box-Muller) and the fp64 pipe overlap.
#include "density.cuh"
#include "flow_core.cuh"
#include "pcn_sweep_tpp.cuh"
#include "philox.cuh"

namespace asp {

constexpr int kMoveWG = 4;          // warpgroups per CTA (one CTA per SM)
constexpr int kMoveMaxMoves = 256;  // per-move accept counters in shared memory
constexpr int kMoveComp = 8;        // mixture components accumulated while the proposal is generated

struct FlowMoveArgs {
    const double *x, *ll, *lp, *lq;
    const int64_t* idx;
    int64_t n;
    asp_density_t like, prior;
    asp_pcn_params_t prm;
    PhiloxKeys keys;
    const uint16_t* weights;
    const float* bias;
    const double* mean;
    const double* inv_std;
    double log_det0;
    int n_layers;
    unsigned int* error;
    double *xo, *llo, *lpo, *lqo;
    unsigned long long* accepts;
    double* scratch;
};

__device__ __forceinline__ void ld_g32(const double* p, double (&v)[4]) {
    asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p) : "memory");
}
__device__ __forceinline__ void st_g32(double* p, const double (&v)[4]) {
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
}

template <int D>
__host__ __device__ constexpr size_t move_scratch_doubles_per_wg() { return (size_t)kFlowM * D; }

// four fp32 values -> columns [taddr, taddr + 4) of this thread's TMEM lane
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const float (&v)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(__float_as_uint(v[0])),
                 "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3]))
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
    uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}

// bytes of the per-dimension constants behind the A tiles
template <int D>
__host__ __device__ inline size_t move_const_bytes(int like_kind, int n_comp) {
    const int ncp = (n_comp + kMoveComp - 1) / kMoveComp * kMoveComp;
    size_t b = (size_t)D * 16 * 3;                                               // s_ref, s_q, s_pr
    b += like_kind == ASP_DENSITY_GMM ? (size_t)D * ncp * 8 + (size_t)ncp * 16  // means [D][ncp], {1/2 is^2, logc}[ncp]
                                      : (size_t)D * 16;                         // Gaussian likelihood {mean, 1/sigma}
    return b;
}

template <int D, int LIKE, bool PHILOX>
__global__ void __launch_bounds__(kFlowThreads* kMoveWG, 1) flow_move_kernel(const FlowMoveArgs A) {
    using S = FlowSizes<D>;
    constexpr int NWG = kMoveWG;
    constexpr int NQ = D / 4;
    constexpr int kTmemAll = 512;  // per warpgroup 64 accumulator columns (+ the parked flow input) and 64 for the noise
    constexpr bool kGmm = LIKE == ASP_DENSITY_GMM;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* s_w = smem;
    unsigned char* s_bb = smem + A.n_layers * S::kLayer;
    unsigned char* s_ones = s_bb + A.n_layers * S::kBiasLayer;
    unsigned char* s_a_all = s_ones + S::kOnes;
    double2* s_ref = reinterpret_cast<double2*>(s_a_all + NWG * kFlowM * kFlowH * 2);  // {mu_k, rho sigma_k}
    double2* s_q = s_ref + D;   // {a2_k, a1_k}: the Gaussian terms of the accept test folded into one quadratic
    double2* s_pf = s_q + D;    // {prior mean_k, prior 1/sigma_k}: the final log prior
    const int n_comp = kGmm ? A.like.n_comp : 0;
    const int ncp = (n_comp + kMoveComp - 1) / kMoveComp * kMoveComp;
    double* s_gm = reinterpret_cast<double*>(s_pf + D);                                // mixture means [D][ncp]
    double2* s_gw = reinterpret_cast<double2*>(s_gm + (size_t)D * ncp);                // {is^2, logc - 1/2 is^2 |m|^2}[ncp]
    double2* s_lk = reinterpret_cast<double2*>(s_pf + D);                              // Gaussian likelihood {mean, 1/sigma}
    unsigned char* tail = reinterpret_cast<unsigned char*>(s_pf + D) + (move_const_bytes<D>(LIKE, n_comp) - (size_t)D * 48);
    double* s_fmean = reinterpret_cast<double*>(tail);                                  // flow mean
    float* s_scale = reinterpret_cast<float*>(s_fmean + D);                             // flow 1 / std (fp32, as flow_kernel)
    unsigned int* s_acc = reinterpret_cast<unsigned int*>(s_scale + D);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_acc + kMoveMaxMoves);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 1 + NWG);

    const int wg = threadIdx.x / kFlowThreads, tid = threadIdx.x % kFlowThreads, warp = tid >> 5;
    unsigned char* s_a = s_a_all + wg * (kFlowM * kFlowH * 2);
    const int wbytes = A.n_layers * S::kLayer;
    const int n_moves = A.prm.n_moves;
    const bool per_move = A.prm.reserved == 0;
    const double rho = A.prm.rho, beta = A.prm.beta, ca = sqrt(1.0 - rho * rho);
    const double ninf = identity(OP_MAX);
    if (threadIdx.x == 0) {
        mbar_init(&s_bar[0], 1);
        for (int w = 0; w < NWG; ++w) mbar_init(&s_bar[1 + w], 1);
        fence_mbar_init();
    }
    if (threadIdx.x < 32) tmem_alloc(s_tmem, kTmemAll);
    for (int i = threadIdx.x; i < A.n_layers * S::kBiasLayer / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(s_bb)[i] = 0u;
    for (int i = threadIdx.x; i < S::kOnes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(s_ones)[i] = 0u;
    for (int t = threadIdx.x; t < kMoveMaxMoves; t += blockDim.x) s_acc[t] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < A.n_layers * 3 * 64; i += blockDim.x) {  // bias blocks: see flow_kernel
        const int l = i / 192, g = (i % 192) / 64, n = i % 64;
        const int N = g < 2 ? kFlowH : D;
        if (n < N) {
            const float b = A.bias[i];
            const __nv_bfloat16 hi = __float2bfloat16_rn(b);
            const __nv_bfloat16 lo = __float2bfloat16_rn(b - __bfloat162float(hi));
            unsigned char* blk = s_bb + l * S::kBiasLayer + (g == 0 ? 0 : (g == 1 ? S::kB12 : 2 * S::kB12));
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(blk + (n >> 3) * 128 + (n & 7) * 16);
            dst[0] = hi;
            dst[1] = lo;
        }
    }
    for (int r = threadIdx.x; r < kFlowM; r += blockDim.x) {
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(s_ones + (r >> 3) * 128 + (r & 7) * 16);
        dst[0] = __float2bfloat16_rn(1.0f);
        dst[1] = __float2bfloat16_rn(1.0f);
    }
    for (int j = threadIdx.x; j < D; j += blockDim.x) {
        const double mu = A.prm.d_mu[j], sg = A.prm.d_sigma[j];
        // T(x) = (1-beta) log q + beta (logL + logP) + r,  r = 1/2 |(x - mu)/sigma|^2: the Gaussian prior, a Gaussian
        // likelihood and r are per-dimension quadratics in y = x - mu and fold into a2 y^2 + a1 y (constants cancel
        // in the accept difference) - as in the particle-per-thread sweep (pcn_sweep_tpp.cuh)
        const double isr = 1.0 / sg;
        double c2 = 0.5 * isr * isr, c1 = 0.0;
        {
            const double is = A.prior.p1[j], w = -0.5 * beta * is * is;
            c2 += w;
            c1 += 2.0 * w * (mu - A.prior.p0[j]);
        }
        if constexpr (!kGmm) {
            const double is = A.like.p1[j], w = -0.5 * beta * is * is;
            c2 += w;
            c1 += 2.0 * w * (mu - A.like.p0[j]);
            s_lk[j] = make_double2(A.like.p0[j], A.like.p1[j]);
        }
        s_ref[j] = make_double2(mu, rho * sg);
        s_q[j] = make_double2(c2, c1);
        s_pf[j] = make_double2(A.prior.p0[j], A.prior.p1[j]);
        s_fmean[j] = A.mean[j];
        s_scale[j] = (float)A.inv_std[j];
    }
    if constexpr (kGmm) {
        for (int e = threadIdx.x; e < D * ncp; e += blockDim.x) {
            const int k = e / ncp, c = e % ncp;
            s_gm[e] = c < n_comp ? A.like.p0[(size_t)c * D + k] : 0.0;
        }
        for (int c = threadIdx.x; c < ncp; c += blockDim.x) {
            const double is = c < n_comp ? A.like.p1[c] : 0.0;
            // log w_c N(v; m_c, s_c^2) = logc - 1/2 is^2 |v - m_c|^2 = (logc - 1/2 is^2 |m_c|^2) + is^2 (v.m_c - 1/2 |v|^2)
            double mm = 0.0;
            for (int k = 0; k < D; ++k) {
                const double mk = c < n_comp ? A.like.p0[(size_t)c * D + k] : 0.0;
                mm = fma(mk, mk, mm);
            }
            s_gw[c] = make_double2(is * is, c < n_comp ? fma(-0.5 * is * is, mm, A.like.p2[c]) : ninf);
        }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem + (uint32_t)(wg * S::kTmemCols);
    if (threadIdx.x == 0) {
        mbar_expect_tx(&s_bar[0], (uint32_t)wbytes);
        for (int l = 0; l < A.n_layers; ++l)
            tma_bulk_load(s_w + l * S::kLayer, reinterpret_cast<const unsigned char*>(A.weights) + (size_t)l * S::kLayer,
                          S::kLayer, &s_bar[0]);
    }
    bool ok = mbar_wait(&s_bar[0], 0);
    uint32_t phase = 0;
    uint64_t* mma_bar = &s_bar[1 + wg];
    const uint32_t lane_taddr = tmem + ((uint32_t)(warp * 32) << 16);
    const uint32_t xi_taddr = lane_taddr + (uint32_t)(NWG * S::kTmemCols);  // this thread's 64 noise columns
    const uint32_t stream0 = A.prm.stream0;
    // this thread's slots of the warpgroup's scratch tile: element (chunk q, thread) at q * 128 * 4 doubles
    double* cur = A.scratch + ((size_t)blockIdx.x * NWG + wg) * move_scratch_doubles_per_wg<D>() + (size_t)tid * 4;
    constexpr size_t kChunk = (size_t)kFlowM * 4;
    unsigned int n_acc = 0;

    // x' = mu + (ca (x - mu) + rho sigma xi) with fixed roundings: the proposal is formed twice (for the densities,
    // and again when it is accepted) and both must give the same bits
    auto propose = [&](double xc, double xi, const double2& rf) -> double {
        return __dadd_rn(rf.x, __fma_rn(ca, __dsub_rn(xc, rf.x), __dmul_rn(rf.y, xi)));
    };
    // the normals of chunk q: drawn once per move (Philox + Box-Muller) and parked in TMEM, recalled from there when
    // the proposal is formed again (further mixture components, accepted move); supplied noise is simply re-read
    auto noise_draw = [&](int t, int64_t row, uint64_t pidx, int q, double (&xi)[4]) {
        if constexpr (PHILOX) {
            float zn[4];
            philox_normal4(A.keys, stream0 + (uint32_t)t, pidx, (uint32_t)q, zn);
            tmem_st4(xi_taddr + (uint32_t)(q * 4), zn);
#pragma unroll
            for (int j = 0; j < 4; ++j) xi[j] = (double)zn[j];
        } else {
            ld_row32(A.prm.d_xi + ((size_t)t * A.n + row) * D + q * 4, xi);
        }
    };
    auto noise_recall = [&](int t, int64_t row, int q, double (&xi)[4]) {  // warp-uniform call sites only
        if constexpr (PHILOX) {
            float zn[4];
            tmem_ld4(xi_taddr + (uint32_t)(q * 4), zn);
#pragma unroll
            for (int j = 0; j < 4; ++j) xi[j] = (double)zn[j];
        } else {
            ld_row32(A.prm.d_xi + ((size_t)t * A.n + row) * D + q * 4, xi);
        }
    };

    const int64_t n_tiles = (A.n + kFlowM - 1) / kFlowM;
    for (int64_t tile = (int64_t)blockIdx.x * NWG + wg; tile < n_tiles; tile += (int64_t)gridDim.x * NWG) {
        const int64_t row_raw = tile * kFlowM + tid;
        const bool live = row_raw < A.n;
        const int64_t row = live ? row_raw : A.n - 1;
        const int64_t src = A.idx ? A.idx[row] : row;
        const uint64_t pidx = (uint64_t)(A.prm.d_slot ? A.prm.d_slot[row] : A.prm.index0 + row);
        double ll = A.ll[src], lp = A.lp[src], lq = A.lq[src];
        // the ancestor's row -> scratch, with the folded quadratic of the current point
        double q_cur = 0.0;
        {
            const double* xs = A.x + src * D;
#pragma unroll 4
            for (int q = 0; q < NQ; ++q) {
                double v[4];
                ld_row32(xs + q * 4, v);
                st_g32(cur + q * kChunk, v);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const double y = v[j] - lds2(&s_ref[q * 4 + j]).x;
                    const double2 e = lds2(&s_q[q * 4 + j]);
                    q_cur = fma(y, fma(e.x, y, e.y), q_cur);
                }
            }
        }
        // current value of the target up to the constants that cancel; NaN -> -inf (smc/base.py:518)
        double t_cur = fma(1.0 - beta, lq, q_cur);
        if constexpr (kGmm) t_cur = fma(beta, ll, t_cur);
        if (t_cur != t_cur) t_cur = ninf;
        bool moved = false;
        uint4 aw = make_uint4(0, 0, 0, 0);
        bool pend = false;  // the last move was accepted and its proposal is not yet in the scratch tile
        for (int t = 0; t < n_moves; ++t) {
            double log_u;
            if constexpr (PHILOX) {
                if ((t & 3) == 0) aw = accept_words(A.keys, stream0, t, pidx);
                log_u = accept_log_u(aw, t);
            } else {
                log_u = A.prm.d_log_u[(size_t)t * A.n + row];
            }
            // ---- proposal, prior, likelihood, reference term, chunk by chunk (rolled: small code, few registers);
            //      the standardised point is parked in this thread's TMEM lane (the accumulator columns are idle) ----
            double qn = 0.0, vv = 0.0;
            double d2[kMoveComp];  // mixture: v . m_c
#pragma unroll
            for (int c = 0; c < kMoveComp; ++c) d2[c] = 0.0;
            double xc[4];
            ld_g32(cur, xc);
            const bool any_pend = __any_sync(0xffffffffu, pend);  // warp-uniform: the TMEM reads are .sync.aligned
#pragma unroll 2
            for (int q = 0; q < NQ; ++q) {
                double xi[4], nxt[4];
                ld_g32(cur + (q + 1 < NQ ? q + 1 : q) * kChunk, nxt);  // the next chunk is in flight during this one
                if (any_pend) {  // the previous move was accepted: its proposal is formed again here, on the chunk
                                 // this move reads anyway, from the normals still parked in TMEM
                    noise_recall(t - 1, row, q, xi);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const double xa = propose(xc[j], xi[j], lds2(&s_ref[q * 4 + j]));
                        xc[j] = pend ? xa : xc[j];
                    }
                    if (pend) st_g32(cur + q * kChunk, xc);
                }
                noise_draw(t, row, pidx, q, xi);
                float zf[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int d = q * 4 + j;
                    const double2 rf = lds2(&s_ref[d]), e = lds2(&s_q[d]);
                    const double v = propose(xc[j], xi[j], rf);
                    const double y = v - rf.x;
                    qn = fma(y, fma(e.x, y, e.y), qn);
                    if constexpr (kGmm) {
                        const double* mrow = s_gm + (size_t)d * ncp;
                        vv = fma(v, v, vv);
#pragma unroll
                        for (int c = 0; c < kMoveComp; c += 2) {
                            const double2 m = lds2(reinterpret_cast<const double2*>(mrow + c));
                            d2[c] = fma(v, m.x, d2[c]);
                            d2[c + 1] = fma(v, m.y, d2[c + 1]);
                        }
                    }
                    zf[j] = (float)(v - s_fmean[d]) * s_scale[d];
                }
                tmem_st4(lane_taddr + (uint32_t)(q * 4), zf);
#pragma unroll
                for (int j = 0; j < 4; ++j) xc[j] = nxt[j];
            }
            double ll2 = 0.0;
            if constexpr (kGmm) {
                double m = ninf, acc = 0.0;
#pragma unroll
                for (int c = 0; c < kMoveComp; ++c) {
                    if (c < n_comp) {
                        const double2 w = lds2(&s_gw[c]);
                        lse_push(m, acc, fma(w.x, fma(-0.5, vv, d2[c]), w.y));
                    }
                }
                if (n_comp > kMoveComp) tmem_wait_st();
                for (int c0 = kMoveComp; c0 < n_comp; c0 += kMoveComp) {  // further components: the proposal is re-formed
#pragma unroll
                    for (int c = 0; c < kMoveComp; ++c) d2[c] = 0.0;
#pragma unroll 1
                    for (int q = 0; q < NQ; ++q) {
                        double xi[4];
                        ld_g32(cur + q * kChunk, xc);
                        noise_recall(t, row, q, xi);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const double v = propose(xc[j], xi[j], lds2(&s_ref[q * 4 + j]));
                            const double* mrow = s_gm + (size_t)(q * 4 + j) * ncp + c0;
#pragma unroll
                            for (int c = 0; c < kMoveComp; c += 2) {
                                const double2 mm = lds2(reinterpret_cast<const double2*>(mrow + c));
                                d2[c] = fma(v, mm.x, d2[c]);
                                d2[c + 1] = fma(v, mm.y, d2[c + 1]);
                            }
                        }
                    }
#pragma unroll
                    for (int c = 0; c < kMoveComp; ++c) {
                        if (c0 + c < n_comp) {
                            const double2 w = lds2(&s_gw[c0 + c]);
                            lse_push(m, acc, fma(w.x, fma(-0.5, vv, d2[c]), w.y));
                        }
                    }
                }
                ll2 = m + log(acc);
            }
            // ---- log q(x') : the coupling layers on the tensor cores (same chain as flow_kernel's log_prob) ----
            float z[D];
            tmem_wait_st();
#pragma unroll
            for (int h = 0; h < D / 16; ++h) {
                float v16[16];
                tmem_ld16(lane_taddr + h * 16, v16);
#pragma unroll
                for (int k = 0; k < 16; ++k) z[h * 16 + k] = v16[k];
            }
            float ladj = 0.f;
            for (int layer = 0; layer < A.n_layers; ++layer) {
                if (layer & 1)
                    ok &= coupling_layer<D, false, true>(z, ladj, layer, s_w, s_a, s_ones, s_bb, mma_bar, tmem, lane_taddr,
                                                         phase, tid, wg);
                else
                    ok &= coupling_layer<D, false, false>(z, ladj, layer, s_w, s_a, s_ones, s_bb, mma_bar, tmem, lane_taddr,
                                                          phase, tid, wg);
            }
            float qq = 0.f;
#pragma unroll
            for (int j = 0; j < D; ++j) qq = fmaf(z[j], z[j], qq);
            const double lq2 = (double)(-0.5f * qq - 0.5f * D * kLog2Pi + ladj) + A.log_det0;
            // ---- accept / reject (asp_pcn_accept_scalar); an accepted proposal is formed again, in place ----
            double t_new = fma(1.0 - beta, lq2, qn);
            if constexpr (kGmm) t_new = fma(beta, ll2, t_new);
            const bool acc = live && (log_u < t_new - t_cur);  // NaN -> reject, like the reference's NaN -> -inf
            if (acc) {
                if constexpr (kGmm) ll = ll2;
                lq = lq2;
                t_cur = t_new;
                moved = true;
            }
            pend = acc;  // the accepted point itself is written by the next pass over the scratch tile
            if (per_move) {
                const unsigned int votes = __ballot_sync(0xffffffffu, acc);
                if ((tid & 31) == 0 && votes) atomicAdd(&s_acc[t], __popc(votes));
            } else {
                n_acc += acc ? 1u : 0u;
            }
        }
        {  // final point and its logs (a particle that never moved keeps its input logs bit for bit)
            double* dst = A.xo + row * D;
            double sp = 0.0, sl = 0.0;
            const bool any_pend = __any_sync(0xffffffffu, pend);
#pragma unroll 2
            for (int q = 0; q < NQ; ++q) {
                double v[4];
                ld_g32(cur + q * kChunk, v);
                if (any_pend) {  // the last move's accepted proposal
                    double xi[4];
                    noise_recall(n_moves - 1, row, q, xi);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const double xa = propose(v[j], xi[j], lds2(&s_ref[q * 4 + j]));
                        v[j] = pend ? xa : v[j];
                    }
                }
                if (live) st_row32(dst + q * 4, v[0], v[1], v[2], v[3]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const double2 pf = lds2(&s_pf[q * 4 + j]);
                    const double zp = (v[j] - pf.x) * pf.y;
                    sp = fma(zp, zp, sp);
                    if constexpr (!kGmm) {
                        const double2 lk = lds2(&s_lk[q * 4 + j]);
                        const double zl = (v[j] - lk.x) * lk.y;
                        sl = fma(zl, zl, sl);
                    }
                }
            }
            if (moved) {
                lp = A.prior.s0 - 0.5 * sp;
                if constexpr (!kGmm) ll = A.like.s0 - 0.5 * sl;
            }
            if (live) {
                A.llo[row] = ll;
                A.lpo[row] = lp;
                A.lqo[row] = lq;
                if (lq != lq) atomicAdd(&A.accepts[per_move ? n_moves : 1], 1ull);  // NaN log q (smc/minipcn.py:133)
            }
        }
    }
    if (!per_move) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) n_acc += __shfl_xor_sync(0xffffffffu, n_acc, o);
        if ((threadIdx.x & 31) == 0 && n_acc) atomicAdd(&s_acc[0], n_acc);
    }
    if (!ok && tid == 0) atomicOr(A.error, 1u);
    tc_fence_before();
    __syncthreads();
    for (int t = threadIdx.x; t < (per_move ? n_moves : 1); t += blockDim.x)
        if (s_acc[t]) atomicAdd(&A.accepts[t], (unsigned long long)s_acc[t]);
    if (threadIdx.x < 32) tmem_dealloc(*s_tmem, kTmemAll);
}

template <int D>
static size_t move_smem_bytes(int n_layers, int like_kind, int n_comp) {
    using S = FlowSizes<D>;
    return (size_t)n_layers * (S::kLayer + S::kBiasLayer) + S::kOnes + (size_t)kMoveWG * kFlowM * kFlowH * 2 +
           move_const_bytes<D>(like_kind, n_comp) + D * 12 + kMoveMaxMoves * 4 + (1 + kMoveWG) * 8 + 16 + 128;
}

static int move_grid(int64_t n) {
    const int64_t n_tiles = (n + kFlowM - 1) / kFlowM;
    const int64_t need = (n_tiles + kMoveWG - 1) / kMoveWG;
    return (int)(need < kNumSMs ? need : kNumSMs);
}

template <int D, int LIKE>
static int launch_move(const FlowMoveArgs& A, cudaStream_t s) {
    const size_t smem = move_smem_bytes<D>(A.n_layers, LIKE, LIKE == ASP_DENSITY_GMM ? A.like.n_comp : 0);
    ASP_REQUIRE(smem <= 227 * 1024, "flow sweep: the flow and the mixture do not fit in shared memory");
    auto launch = [&](auto kernel) -> int {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("flow sweep kernel: %s", cudaGetErrorString(e));
            return ASP_ERR_CUDA;
        }
        kernel<<<move_grid(A.n), kFlowThreads * kMoveWG, smem, s>>>(A);
        return check_launch("asp_pcn_sweep_flow");
    };
    if (A.prm.noise_mode == 0) return launch(flow_move_kernel<D, LIKE, true>);
    return launch(flow_move_kernel<D, LIKE, false>);
}

template <int D>
static int launch_move_like(const FlowMoveArgs& A, cudaStream_t s) {
    if (A.like.kind == ASP_DENSITY_GMM) return launch_move<D, ASP_DENSITY_GMM>(A, s);
    return launch_move<D, ASP_DENSITY_GAUSSIAN>(A, s);
}

}  // namespace asp

using namespace asp;

extern "C" {

int32_t asp_pcn_sweep_flow_supported(int32_t dims, int32_t like_kind, int32_t prior_kind) {
    return (dims == 16 || dims == 32 || dims == 64) && (like_kind == ASP_DENSITY_GMM || like_kind == ASP_DENSITY_GAUSSIAN) &&
                   prior_kind == ASP_DENSITY_GAUSSIAN
               ? 1 : 0;
}

size_t asp_pcn_sweep_flow_workspace_bytes(int32_t dims) {
    return (size_t)kNumSMs * kMoveWG * kFlowM * (size_t)dims * sizeof(double);
}

int asp_pcn_sweep_flow(const double* d_x, const double* d_log_l, const double* d_log_p, const double* d_log_q,
                       const int64_t* d_idx, int64_t n, int32_t dims, const asp_density_t* like,
                       const asp_density_t* prior, const asp_flow_t* flow, const asp_pcn_params_t* params,
                       double* d_x_out, double* d_log_l_out, double* d_log_p_out, double* d_log_q_out,
                       unsigned long long* d_accepts, void* d_ws, void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(d_x && d_log_l && d_log_p && d_log_q && like && prior && flow && params && d_x_out && d_log_l_out &&
                    d_log_p_out && d_log_q_out && d_accepts && d_ws, "null pointer");
    ASP_REQUIRE(n > 0, "n must be positive");
    ASP_REQUIRE(like->dims == dims && prior->dims == dims && flow->dims == dims, "dims mismatch");
    ASP_REQUIRE(asp_pcn_sweep_flow_supported(dims, like->kind, prior->kind), "unsupported dims / density kinds");
    ASP_REQUIRE(flow->d_weights && flow->d_bias && flow->d_mean && flow->d_inv_std && flow->d_error, "flow descriptor incomplete");
    ASP_REQUIRE(flow->hidden == kFlowH && flow->n_layers >= 1 && flow->n_layers <= kFlowMaxLayers, "bad flow shape");
    ASP_REQUIRE(params->n_moves >= 0 && params->n_moves <= kMoveMaxMoves, "n_moves must be in [0, 256]");
    ASP_REQUIRE(params->rho > 0.0 && params->rho <= 1.0, "rho must be in (0, 1]");
    ASP_REQUIRE(params->d_mu && params->d_sigma, "reference measure missing");
    ASP_REQUIRE(params->noise_mode == 0 || (params->d_xi && params->d_log_u), "noise arrays missing");
    ASP_REQUIRE(!(params->nu > 0.0), "asp_pcn_sweep_flow is pCN only");
    ASP_REQUIRE(d_x != d_x_out, "the flow sweep is not in place");
    ASP_REQUIRE(like->kind != ASP_DENSITY_GMM || (like->p0 && like->p1 && like->p2 && like->n_comp >= 1 && like->n_comp <= 32),
                "bad mixture parameters");
    ASP_REQUIRE(((reinterpret_cast<uintptr_t>(d_x) | reinterpret_cast<uintptr_t>(d_x_out) | reinterpret_cast<uintptr_t>(d_ws) |
                  reinterpret_cast<uintptr_t>(params->d_xi)) & 31) == 0, "rows must be 32-byte aligned");
    FlowMoveArgs A;
    A.x = d_x; A.ll = d_log_l; A.lp = d_log_p; A.lq = d_log_q; A.idx = d_idx; A.n = n;
    A.like = *like; A.prior = *prior; A.prm = *params; A.keys = philox_keys(params->seed);
    A.weights = (const uint16_t*)flow->d_weights; A.bias = flow->d_bias; A.mean = flow->d_mean; A.inv_std = flow->d_inv_std;
    A.log_det0 = flow->log_det0; A.n_layers = flow->n_layers; A.error = flow->d_error;
    A.xo = d_x_out; A.llo = d_log_l_out; A.lpo = d_log_p_out; A.lqo = d_log_q_out; A.accepts = d_accepts;
    A.scratch = (double*)d_ws;
    cudaStream_t s = (cudaStream_t)stream;
    switch (dims) {
        case 16: return launch_move_like<16>(A, s);
        case 32: return launch_move_like<32>(A, s);
        default: return launch_move_like<64>(A, s);
    }
}

}  // extern "C"

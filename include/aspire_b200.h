/*
 * aspire_b200 - C-ABI of the B200-native SMC particle engine.
 *
 * This is the drop-in boundary for the ONE hot path of mj-will/aspire this
 * repository rebuilds: importance reweighting, log-sum-exp /
 * ESS, adaptive-beta bisection, exact multinomial / systematic resampling, the
 * pCN mutation sweep and the built-in benchmark log-densities.  The reference
 * is pure Python and has no FFI of its own; each entry point below names the
 * reference code (path:line under /root/reference) whose arithmetic it
 * replaces, and INTEGRATION.md shows the ctypes binding a maintainer of the
 * reference would add.
 *
 * Conventions
 *  - plain C types only; every pointer named d_* is a DEVICE pointer (fp64 /
 *    int64, contiguous, row-major, 16-byte aligned), every h_* a host pointer;
 *  - the caller owns all memory; scratch comes from the caller through
 *    `ws` (size from the matching *_workspace_bytes, 256-byte aligned);
 *  - `stream` is a cudaStream_t passed as void*; all launches are
 *    asynchronous on it, nothing synchronises unless stated;
 *  - every function returns 0 on success, a negative asp_status otherwise,
 *    and never throws; asp_last_error() is thread-local;
 *  - there is NO CPU fallback: without an sm_100 device every compute entry
 *    point fails with ASP_ERR_DEVICE.
 */
#ifndef ASPIRE_B200_H
#define ASPIRE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* the library is built with -fvisibility=hidden: only these entry points are exported */
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define ASP_VERSION 100 /* 0.1.0 */

enum asp_status {
    ASP_OK = 0,
    ASP_ERR_ARG = -1,     /* bad argument (null pointer, size, alignment)      */
    ASP_ERR_DEVICE = -2,  /* no sm_100 device / wrong architecture             */
    ASP_ERR_CUDA = -3,    /* CUDA runtime error, see asp_last_error()          */
    ASP_ERR_NUMERIC = -4  /* reserved: numeric guards are reported in-band     */
};

int asp_version(void);
const char* asp_last_error(void);
/* 0 when the current device is compute capability 10.x, else ASP_ERR_DEVICE. */
int asp_device_check(void);

/* ---- built-in log-densities (K1) ---------------------------------------
 * The reference takes user callables (src/aspire/samplers/base.py:81-91); these
 * are the synthetic benchmark targets of SURVEY 8d, following the reference's
 * fixtures README.md:47-52, tests/integration_tests/conftest.py:84-108,
 * examples/smc_example.py:40-58.
 *   GAUSSIAN   s0 - 1/2 sum_j ((x_j - p0_j) * p1_j)^2        p0=mean[D] p1=1/sigma[D]
 *   ROSENBROCK -sum_{j<D-1}[s0 (x_{j+1}-x_j^2)^2 + (s1-x_j)^2] / s2
 *   GMM        logsumexp_k[p2_k - 1/2 p1_k^2 |x-p0_k|^2]     p0=means[K,D] p1=1/sigma[K] p2=logc[K]
 *   UNIFORM    s0 inside the open box (p0=lo[D], p1=hi[D]), -inf outside
 */
enum asp_density_kind {
    ASP_DENSITY_GAUSSIAN = 0,
    ASP_DENSITY_ROSENBROCK = 1,
    ASP_DENSITY_GMM = 2,
    ASP_DENSITY_UNIFORM = 3
};

typedef struct asp_density {
    int32_t kind;
    int32_t dims;
    int32_t n_comp; /* GMM only (<= 32) */
    int32_t reserved;
    const double* p0; /* device */
    const double* p1; /* device */
    const double* p2; /* device */
    double s0, s1, s2;
} asp_density_t;

/* out[i] = log density(x[i, :]);  x is [n, dims] row-major. */
int asp_logdensity(const asp_density_t* dens, const double* d_x, int64_t n, double* d_out, void* stream);

/* ---- importance weights (K11) -------------------------------------------
 * Samples.compute_weights, src/aspire/samples.py:457-475.
 * d_log_w = logL + logP - logq;  d_weights = exp(log_w) (unshifted, as the
 * reference);  d_result[8] = {log_evidence, evidence, evidence_error,
 * log_evidence_error, effective_sample_size, max log_w, n_nan, reserved}. */
size_t asp_importance_workspace_bytes(int64_t n);
int asp_importance_weights(const double* d_log_l, const double* d_log_p, const double* d_log_q, int64_t n,
                           double* d_log_w, double* d_weights, double* d_result, void* ws, void* stream);

/* ---- SMC reweighting (K2) -----------------------------------------------
 * d[i] = (logL + logP) - logq, the per-particle slope of
 * SMCSamples.unnormalized_log_weights, src/aspire/samples.py:1221-1224
 * ((beta-beta0)*d == (beta0-beta)*logq + (beta-beta0)*(logL+logP)).
 * d_stats[4] = {min d, max d, number of NaN, number of +inf}, over the LOCAL
 * slice; multi-GPU callers all-reduce it (min, max, sum, sum). */
size_t asp_reduce_workspace_bytes(void);
int asp_reweight(const double* d_log_l, const double* d_log_p, const double* d_log_q, int64_t n, double* d_d,
                 double* d_stats, void* ws, void* stream);

/* ---- ESS(beta) sums (K3) ------------------------------------------------
 * effective_sample_size(log_weights(beta)), src/aspire/utils.py:248-255,510-512:
 * for each k < n_beta:  c_k = dbeta_k * (dbeta_k >= 0 ? d_max : d_min),
 * d_sums[2k] = sum_i exp(dbeta_k d_i - c_k), d_sums[2k+1] = sum_i exp(2(dbeta_k d_i - c_k)).
 * ESS = sums[2k]^2 / sums[2k+1].  d_dbeta / d_stats are device arrays so the
 * call can sit in a stream / CUDA graph with no host round trip. n_beta <= 8. */
int asp_ess_sums(const double* d_d, int64_t n, const double* d_dbeta, int n_beta, const double* d_stats,
                 double* d_sums, void* ws, void* stream);

/* ---- adaptive-beta bisection, device-driven (a5) ------------------------
 * SMCSampler.determine_beta, src/aspire/samplers/smc/base.py:166-186.  The state
 * lives on the device; asp_beta_search_init seeds it, each asp_beta_search_pass
 * evaluates ESS at state.beta_try over the LOCAL slice (sums land in
 * state.s1/s2), asp_beta_search_decide applies the reference's `>=` decision
 * and picks the next midpoint.  Single GPU: pass(fused_decide=1) does both.
 * Multi GPU: pass(0) -> all-reduce(state.s1,s2; sum) -> decide.  Passes after
 * convergence are no-ops, so the host launches a fixed count
 * (asp_beta_search_max_passes) without synchronising. */
typedef struct asp_beta_state {
    double beta_prev, beta_min, beta_max; /* bisection bracket                     */
    double target_eff, tol, n_total;      /* target efficiency, tolerance, global N */
    double beta_try;                      /* beta the next pass evaluates           */
    double s1, s2;                        /* sums of the last pass (all-reduce these) */
    double eff_beta_max;                  /* ESS(1)/N (history.ess_target / N)      */
    double s1_min, s2_min;                /* sums at beta_min (ESS(beta*) for free)  */
    double d_min, d_max;                  /* global min / max of d                  */
    int32_t phase;                        /* 0: evaluating beta=1, 1: bisecting, 2: done */
    int32_t n_evals;
    int32_t min_valid;                    /* 1 when s1_min/s2_min belong to beta_min */
    int32_t reserved;
} asp_beta_state_t;

int asp_beta_search_max_passes(double beta_prev, double tol);
int asp_beta_search_init(asp_beta_state_t* d_state, double beta_prev, double target_eff, double tol,
                         double n_total, const double* d_stats, void* stream);
int asp_beta_search_pass(const double* d_d, int64_t n, asp_beta_state_t* d_state, int fused_decide, void* ws,
                         void* stream);
int asp_beta_search_decide(asp_beta_state_t* d_state, void* stream);
/* Three bisection levels per pass over d: the seven candidate betas of a 3-level bisection are equally
 * spaced, so their weights are E r^k (two exps + seven multiplies per particle).  d_sums[14] receives
 * {S1_k, S2_k}, k = 1..7 (all-reduce these across ranks, then asp_beta_search_decide3); decisions and the
 * resulting beta are those of the one-level-per-pass search.  Run the beta = 1 evaluation first with one
 * asp_beta_search_pass, then asp_beta_search_max_passes3 passes of these. */
/* Multi-GPU: the all-reduce of the pass sums done INSIDE the pass kernel over peer memory (no NCCL call, no
 * separate decide launch).  box[r] is rank r's mailbox (ASP_PEER_BOX_BYTES of zero-initialised device memory,
 * exchanged between the processes with CUDA IPC) as addressable from this device; every rank must make the
 * same sequence of calls.  levels3 = 0: the single-beta pass, 1: three levels.
 * A peer that never arrives is reported in asp_beta_state_t.reserved = -1 after a bounded wait. */
#define ASP_MAX_PEERS 8
#define ASP_PEER_BOX_BYTES (2 * ASP_MAX_PEERS * 16 * 8 + 64 + 128)
typedef struct asp_peer {
    double* box[ASP_MAX_PEERS];
    int32_t rank, world;
} asp_peer_t;
int asp_beta_search_pass_p2p(const double* d_d, int64_t n, asp_beta_state_t* d_state, int levels3, double* d_sums,
                             const asp_peer_t* peer, void* ws, void* stream);
int asp_beta_search_max_passes3(double beta_prev, double tol);
int asp_beta_search_pass3(const double* d_d, int64_t n, asp_beta_state_t* d_state, double* d_sums, int fused_decide,
                          void* ws, void* stream);
int asp_beta_search_decide3(asp_beta_state_t* d_state, const double* d_sums, void* stream);
/* Root-guided search (csrc/beta_math.cuh): ESS(beta) is non-increasing, so every decision of the reference's
 * bisection is "mid <= beta*".  After the beta = 1 pass: a subsample bisection and up to two Newton passes
 * estimate the root, the reference's recursion is replayed towards it down to the tree node of width <= 8 tol,
 * and ONE pass evaluates that node's end points (which prove the root is inside: every earlier decision follows
 * by monotonicity) and its seven interior candidates (the last three decisions, as asp_beta_search_pass3).
 * Same beta, bracket, sums and n_evals as the plain search; a miss is retried on the neighbouring node, and
 * after three misses the state is still that of the beta = 1 pass, so the asp_beta_search_pass3 passes that the
 * caller launches afterwards (no-ops once the search is done) finish it.  d_scratch:
 * asp_beta_search_scratch_words() doubles; peer (multi-GPU mailboxes, asp_beta_search_pass_p2p) may be NULL. */
int32_t asp_beta_search_scratch_words(void);
/* asp_beta_search_init + the beta = 1 pass + asp_beta_search_root in one call (one host -> driver round trip). */
int asp_beta_search_run(const double* d_d, int64_t n, const double* d_stats, double beta_prev, double target_eff,
                        double tol, double n_total, asp_beta_state_t* d_state, double* d_scratch,
                        const asp_peer_t* peer, void* ws, void* stream);
int asp_beta_search_root(const double* d_d, int64_t n, asp_beta_state_t* d_state, double* d_scratch,
                         const asp_peer_t* peer, void* ws, void* stream);

/* ---- resampling weights + evidence bookkeeping (K4/K5 front end) --------
 * p_i = exp(dbeta d_i - c - log S1) (== exp(lw - lse(lw)), src/aspire/samples.py:1277),
 * log_evidence_ratio[_variance] (src/aspire/samples.py:1226-1242).
 * h_s1/h_s2/h_c are the GLOBAL sums and shift for this dbeta; n_total the
 * global particle count.  d_p may be NULL (bookkeeping only).
 * d_result[4] = {sum_i (u_i - mean u)^2 over the local slice, sum p, 0, 0}
 * (variance = result[0] / n_total / (n_total * mean^2)); also fills the exact
 * scan's per-tile sums / maxima (d_tile_stats, may be NULL). */
int asp_smc_weights(const double* d_d, int64_t n, double dbeta, double c, double s1, double n_total,
                    double* d_p, double* d_result, double* d_tile_stats, void* ws, void* stream);

/* ---- exact CDF (K5) -----------------------------------------------------
 * numpy Generator.choice(p=...) as called at src/aspire/samples.py:1278:
 * cdf = cumsum(p) with SEQUENTIAL fp64 rounding, then cdf /= cdf[-1].
 * Bit-identical to numpy for any non-negative finite p.
 * Stages (so that ranks can interleave their collectives):
 *   asp_cdf_tile_stats   per-tile sum/max of p          (skipped if asp_smc_weights filled them)
 *   asp_cdf_plan         approximate prefixes -> per-tile binade / dirty flag
 *   asp_cdf_aggregate    integer increment pairs of the clean tiles
 *   asp_cdf_resolve      exact running sum at every tile start; dirty tiles done here
 *   asp_cdf_write        cdf (normalised by d_total[0] when normalise != 0)
 * d_carry[4]: in  {approx offset, running max, exact incoming sum, n_total}
 *             out (after resolve) d_carry_out[2] = {exact outgoing sum, running max}.
 * asp_cdf_exact runs all five on one GPU. */
#define ASP_CDF_TILE 2048
size_t asp_cdf_workspace_bytes(int64_t n);
int asp_cdf_tile_stats(const double* d_p, int64_t n, void* ws, void* stream);
int asp_cdf_set_tile_stats(const double* d_tile_stats, int64_t n, void* ws, void* stream);
int asp_cdf_plan(int64_t n, const double* d_carry, void* ws, void* stream);
int asp_cdf_aggregate(const double* d_p, int64_t n, void* ws, void* stream);
int asp_cdf_resolve(const double* d_p, int64_t n, const double* d_carry, double* d_cdf, double* d_carry_out,
                    void* ws, void* stream);
/* Multi-GPU carry chain over peer memory instead of send / recv / broadcast (mailboxes: asp_peer_t above).
 * Rank r's resolve stage waits for the exact running sum at the end of rank r-1's slice to land in its mailbox,
 * resolves its slice and stores its own running sum into rank r+1's mailbox; the last rank stores the exact
 * total into every mailbox, where each rank's write stage picks it up.  Every rank calls resolve then write on
 * one stream, once per CDF.  d_chain_out[2] = {running sum before this slice, exact total}; a rank that waited
 * more than 20 s gets NaN there (and in its cdf) instead of hanging. */
int asp_cdf_resolve_p2p(const double* d_p, int64_t n, double* d_cdf, double* d_chain_out, const asp_peer_t* peer,
                        void* ws, void* stream);
int asp_cdf_write_p2p(const double* d_p, int64_t n, int normalise, double* d_cdf, double* d_chain_out,
                      const asp_peer_t* peer, void* ws, void* stream);
int asp_cdf_write(const double* d_p, int64_t n, const double* d_total, int normalise, double* d_cdf, void* ws,
                  void* stream);
int asp_cdf_exact(const double* d_p, int64_t n, int normalise, double* d_cdf, void* ws, void* stream);
/* d_info[4] (int64) = {n_tiles, n_dirty, n_phases, error_flags} of the last resolve. */
int asp_cdf_info(void* ws, int64_t* h_info, void* stream);

/* ---- ancestor search (K6) -----------------------------------------------
 * idx[j] = #{k : cdf[k] <= u[j]}  == cdf.searchsorted(u, side='right'), int64
 * (numpy Generator.choice; src/aspire/samples.py:1278).  Systematic variant
 * (extension): u_j = (j0 + j + u0) / m_total in fp64 (one add, one
 * IEEE division), result clamped to n_cdf - 1.  A global offset is added to every
 * index by the caller when the CDF is a shard of a larger one. */
int asp_search_multinomial(const double* d_cdf, int64_t n_cdf, const double* d_u, int64_t m, int64_t* d_idx,
                           void* stream);
/* Same result through a guide table (one pass over the CDF to build it, then ~3
 * random sectors per ancestor instead of a log2(n)-deep dependent chain); worth it
 * when m is comparable to n_cdf.  ws: asp_search_workspace_bytes(n_cdf). */
size_t asp_search_workspace_bytes(int64_t n_cdf);
int asp_search_multinomial_guided(const double* d_cdf, int64_t n_cdf, const double* d_u, int64_t m, int64_t* d_idx,
                                  void* ws, void* stream);
/* Systematic resampling: d_idx[j - j0] = #{k : cdf[k] <= (j + u0) / m_total} (clamped to n_cdf - 1) for the
 * slots j0 <= j < j0 + m of a sweep over m_total slots.  The queries are sorted, so the CDF is streamed
 * once and every entry writes the slots it owns.  ws: asp_search_systematic_workspace_bytes(m). */
size_t asp_search_systematic_workspace_bytes(int64_t m);
int asp_search_systematic(const double* d_cdf, int64_t n_cdf, double u0, int64_t j0, int64_t m,
                          int64_t m_total, int64_t* d_idx, void* ws, void* stream);

/* ---- multinomial resampling with sorted uniforms (throughput mode of K6) ------------------
 * SMCSamples.resample, src/aspire/samples.py:1251-1287 (idx = searchsorted(cumsum(w), rng.random(M), 'right')):
 * only the multiset of ancestors matters, and the order statistics of M iid uniforms are the normalised
 * partial sums of M + 1 iid exponentials.  Element j = 0 .. m_total draws q_j = floor(-ln(1 - u_j) 2^26) from
 * the Philox uniform stream (seed, stream_id) - asp_philox_uniform's stream; S_j = q_0 + ... + q_j are integers,
 * so tiles, blocks and ranks may be summed in any order; total = S_{m_total};
 *     idx_j = min(#{k : rint(cdf_k * total) <= S_j}, n_cdf - 1)         (sorted: idx_j <= idx_{j+1})
 * Spec: oracle/sorted_resample.py.  Tiles are 2048 elements, aligned to the global element index.
 *   asp_sorted_tiles(m_total)   tiles covering elements 0 .. m_total
 *   asp_sorted_spacings         d_q[n_tiles * 2048] (uint32) and d_tile_sums[n_tiles] (uint64) for the elements
 *                               j0 .. j0 + n_tiles * 2048 - 1 (j0 a multiple of 2048; elements > m_total give 0)
 *   asp_sorted_tile_scan        exclusive prefix of ALL tile sums (after an all-gather when sharded) + total
 *   asp_search_sorted           ancestors of the slots j_lo <= j < j_hi within the CDF (slice) d_cdf[n_cdf];
 *                               d_q[0] is element q_j0; d_tile_prefix is indexed by the global tile;
 *                               ws: asp_search_sorted_workspace_bytes(j_hi - j_lo)
 *   asp_sorted_lower_bound      d_out[i] = #{j < m_total : S_j < d_thresholds[i]}: the shard boundaries - the rank
 *                               that owns the CDF slice [k0, k1) owns the slots from the bound of rint(cdf_{k0-1}
 *                               total) to that of rint(cdf_{k1-1} total) (the tile is regenerated in the kernel) */
int64_t asp_sorted_tiles(int64_t m_total);
int asp_sorted_spacings(uint64_t seed, uint32_t stream_id, int64_t j0, int64_t n_tiles, int64_t m_total, uint32_t* d_q,
                        unsigned long long* d_tile_sums, void* stream);
int asp_sorted_tile_scan(const unsigned long long* d_tile_sums, int64_t n_tiles, unsigned long long* d_tile_prefix,
                         unsigned long long* d_total, void* stream);
size_t asp_search_sorted_workspace_bytes(int64_t n_slots);
int asp_search_sorted(const double* d_cdf, int64_t n_cdf, const uint32_t* d_q, int64_t q_j0,
                      const unsigned long long* d_tile_prefix, const unsigned long long* d_total, int64_t j_lo,
                      int64_t j_hi, int64_t* d_idx, void* ws, void* stream);
int asp_sorted_lower_bound(uint64_t seed, uint32_t stream_id, const unsigned long long* d_tile_prefix, int64_t n_tiles,
                           int64_t m_total, const unsigned long long* d_thresholds, int32_t n_thresholds,
                           long long* d_out, void* stream);

/* ---- row gather (K7) ----------------------------------------------------
 * x[idx], logL[idx], logP[idx], logq[idx] - src/aspire/samples.py:1279-1287.
 * Any of the three log pointers (and its output) may be NULL. */
int asp_gather_rows(const double* d_x, int32_t dims, const int64_t* d_idx, int64_t m, double* d_x_out,
                    const double* d_l, const double* d_p, const double* d_q, double* d_l_out, double* d_p_out,
                    double* d_q_out, void* stream);

/* ---- pCN mutation (K8) --------------------------------------------------
 * Replaces minipcn.Sampler.sample as driven by MiniPCNSMC.mutate
 * (src/aspire/samplers/smc/minipcn.py:69-135) against the tempered target of
 * src/aspire/samplers/smc/base.py:507-519.  Step spec: oracle/pcn.py (parity
 * with minipcn itself is UNPINNED - the package is absent).
 * Reference measure moments: d_moments[1 + 2*dims] = {sum w, sum w x_j, sum w x_j^2}
 * over rows r_i = d_idx ? d_idx[i idx_stride] : i idx_stride, i < n, with w = d_d ? exp(dbeta d_d[r_i] - c) : 1
 * (d_idx given, d_d NULL: plain moments of the resampled ensemble, never materialised; idx_stride > 1: every
 * idx_stride-th offspring - the strided subsample the sampler fits the reference measure to, read in place). */
size_t asp_moments_workspace_bytes(int32_t dims);
int asp_weighted_moments(const double* d_x, int32_t dims, const int64_t* d_idx, int64_t idx_stride, const double* d_d,
                         int64_t n, double dbeta, double c, double* d_moments, void* ws, void* stream);

/* mu_j = sum w x_j / sum w,  sigma_j = max(sqrt(max(sum w x_j^2 / sum w - mu_j^2, 0)), sigma_floor) */
int asp_moments_finish(const double* d_moments, int32_t dims, double sigma_floor, double* d_mu, double* d_sigma,
                       void* stream);

typedef struct asp_pcn_params {
    double beta;          /* inverse temperature of the target                  */
    double rho;           /* pCN step, (0, 1]                                   */
    const double* d_mu;   /* reference measure mean [dims]                      */
    const double* d_sigma;/* reference measure std  [dims]                      */
    int32_t n_moves;
    int32_t noise_mode;   /* 0: Philox in-kernel, 1: supplied d_xi / d_log_u    */
    uint64_t seed;        /* Philox key                                         */
    uint32_t stream0;     /* Philox stream of move 0 (move t uses stream0 + t)  */
    uint32_t reserved;    /* 0: d_accepts[move] per move; 1: only the total, in d_accepts[0] */
    int64_t index0;       /* global index of local particle 0 (Philox counter)  */
    const double* d_xi;   /* [n_moves, n, dims] normals (noise_mode 1)          */
    const double* d_log_u;/* [n_moves, n] log-uniforms   (noise_mode 1)         */
    const int64_t* d_slot;/* optional per-particle Philox counter (global output
                             slot); NULL: index0 + i                            */
    /* tpCN (Student-t preconditioned CN, the reference's default step_fn, smc/minipcn.py:48) */
    double nu;            /* 0: pCN (Gaussian reference); > 0: degrees of freedom of the
                             Student-t reference t_nu(mu, diag sigma^2)          */
    const double* d_scale;/* [n] per-particle proposal scale sqrt(s_i) of the current
                             move (written by asp_tpcn_scales); NULL: 1          */
    const double* d_gamma;/* [n_moves, n] Gamma((dims + nu)/2, 1) variates (noise_mode 1) */
} asp_pcn_params_t;

/* Fused sweep: particle i starts from row d_idx[i] of the inputs (d_idx NULL
 * = identity, i.e. no resampling gather), runs n_moves accept/reject moves
 * with built-in densities and writes the final state.  d_accepts[n_moves + 1]
 * (uint64, caller zeroes it) is incremented by the per-move accept counts; the entry behind them - d_accepts[n_moves],
 * or d_accepts[1] when params->reserved = 1 keeps only the total in d_accepts[0] - counts the particles whose final
 * log q is NaN (the reference raises on them, smc/minipcn.py:133-134: no separate pass over the logs is needed).
 * params->nu > 0 runs tpCN moves in the same kernel (gamma variates drawn in-kernel, or params->d_gamma with
 * supplied noise; d_scale is ignored): same trajectory as asp_tpcn_scales + asp_pcn_propose + asp_pcn_accept.
 * Built for the dims asp_pcn_sweep_tpcn_supported reports (4, 8, 16, 32, 64); other dims return ASP_ERR_ARG. */
int32_t asp_pcn_sweep_tpcn_supported(int32_t dims);
int asp_pcn_sweep(const double* d_x, const double* d_log_l, const double* d_log_p, const double* d_log_q,
                  const int64_t* d_idx, int64_t n, int32_t dims, const asp_density_t* like,
                  const asp_density_t* prior, const asp_density_t* q, const asp_pcn_params_t* params,
                  double* d_x_out, double* d_log_l_out, double* d_log_p_out, double* d_log_q_out,
                  unsigned long long* d_accepts, void* stream);

/* Unfused halves for user-supplied Python callables: propose x' for move
 * `move`, then (after the caller evaluated the three logs at x') accept/reject
 * in place. */
/* tpCN: sqrt(s_i) with s_i = (nu + r2(x_i)) / (2 g_i), g_i ~ Gamma((dims + nu)/2, 1) drawn in-kernel
 * (Marsaglia-Tsang on Philox, noise_mode 0) or taken from params->d_gamma (noise_mode 1).  Step spec:
 * oracle/pcn.py::tpcn_sweep.  Pass the result as params->d_scale to asp_pcn_propose; asp_pcn_accept then
 * uses the Student-t reference term when params->nu > 0. */
int asp_tpcn_scales(const double* d_x, int64_t n, int32_t dims, const asp_pcn_params_t* params, int32_t move,
                    double* d_scale, void* stream);
int asp_pcn_propose(const double* d_x, int64_t n, int32_t dims, const asp_pcn_params_t* params, int32_t move,
                    double* d_x_prop, void* stream);
int asp_pcn_accept(double* d_x, double* d_log_l, double* d_log_p, double* d_log_q, const double* d_x_prop,
                   const double* d_log_l_prop, const double* d_log_p_prop, const double* d_log_q_prop, int64_t n,
                   int32_t dims, const asp_pcn_params_t* params, int32_t move, unsigned long long* d_accepts,
                   void* stream);

/* tpCN reference measure: EM fit (Liu & Rubin 1995) of t_nu(mu, diag sigma^2) to m rows of d_x - row i is
 * d_x[(d_idx ? d_idx[i * stride] : i * stride)] - entirely on the device: n_iter kernels, no host round trip.
 * d_state: [2 dims + 1] doubles = mu | sigma | nu, holding the starting point on entry (the ensemble's mean / std and
 * nu0) and the fit on return.  Spec: oracle/pcn.py::fit_student_t_diag (minipcn absent: parity unpinned).
 * ws: asp_student_t_workspace_bytes(dims) bytes, zero-initialised once. */
size_t asp_student_t_workspace_bytes(int32_t dims);
int asp_student_t_fit(const double* d_x, int32_t dims, const int64_t* d_idx, int64_t stride, int64_t m, int32_t n_iter,
                      double nu_min, double nu_max, double* d_state, void* ws, void* stream);

/* Semi-fused move for a proposal density that is not a built-in (e.g. a flow evaluated by
 * asp_flow_logprob_bf16): asp_pcn_propose_eval proposes x' for move `move` and evaluates the built-in
 * prior / likelihood at x' plus the reference-measure term d_dr = r(x') - r(x) in ONE pass (x read once,
 * x' written once); after the caller evaluated log q(x'), asp_pcn_accept_scalar decides from the scalars
 * and copies the accepted rows.  Same step and noise stream as asp_pcn_sweep / asp_pcn_propose + asp_pcn_accept.
 * asp_pcn_propose_eval_supported: 1 if (dims, kinds) is built (dims 4, 8, 16, 32, 64), else use the unfused pair. */
int32_t asp_pcn_propose_eval_supported(int32_t dims, int32_t like_kind, int32_t prior_kind);
int asp_pcn_propose_eval(const double* d_x, int64_t n, int32_t dims, const asp_density_t* like,
                         const asp_density_t* prior, const asp_pcn_params_t* params, int32_t move, double* d_x_prop,
                         double* d_log_l_prop, double* d_log_p_prop, double* d_dr, void* stream);
int asp_pcn_accept_scalar(double* d_x, double* d_log_l, double* d_log_p, double* d_log_q, const double* d_x_prop,
                          const double* d_log_l_prop, const double* d_log_p_prop, const double* d_log_q_prop,
                          const double* d_dr, int64_t n, int32_t dims, const asp_pcn_params_t* params, int32_t move,
                          unsigned long long* d_accepts, void* stream);

/* ---- coupling-flow log_prob / sample on tcgen05 (K9/K10) ---------------------
 * ZukoFlow.log_prob / sample_and_log_prob, src/aspire/flows/torch/flows.py:327-387, with
 * the affine data transform of src/aspire/transforms.py:618-628 folded in.  Spec and
 * weight layout: oracle/flow.py (T coupling layers, conditioner MLP(D/2 -> 64 -> 64 -> D),
 * bf16 x bf16 -> fp32 tensor-core GEMMs, fp32 epilogue; stated tolerance in the tests).
 *   d_weights  bf16 blob, per layer W1[64, max(16, D/2)] | W2[64, 64] | W3[D, 64], each in the
 *              tcgen05 no-swizzle K-major core-matrix layout (asp_flow_weights_bytes)
 *   d_bias     fp32 [n_layers][3][64]   (b1 | b2 | b3 padded to 64)
 *   d_mean / d_inv_std / d_std  fp64 [dims];  log_det0 = -sum log std
 *   d_error    device uint32, bit 0 set if a tensor-core barrier timed out
 * dims in {32, 64}, hidden = 64, n_layers <= 8. */
typedef struct asp_flow {
    int32_t dims, hidden, n_layers, reserved;
    const void* d_weights;
    const float* d_bias;
    const double* d_mean;
    const double* d_inv_std;
    const double* d_std;
    double log_det0;
    unsigned int* d_error;
} asp_flow_t;

size_t asp_flow_weights_bytes(int32_t dims, int32_t n_layers);
/* d_logq[i] = log q(x_i); d_z (optional) receives the base-space point. */
int asp_flow_logprob_bf16(const asp_flow_t* flow, const double* d_x, int64_t n, double* d_logq, double* d_z, void* stream);
/* d_u: base normals [n, dims] -> d_x [n, dims], d_logq[i] = log q(x_i). */
int asp_flow_sample_bf16(const asp_flow_t* flow, const double* d_u, int64_t n, double* d_x, double* d_logq, void* stream);

/* ---- pCN sweep with the coupling flow as proposal density (C4) -----------------
 * MiniPCNSMC.mutate (src/aspire/samplers/smc/minipcn.py:89-117) with a learned flow as prior_flow: the resampling
 * gather, n_moves x (propose, built-in prior / likelihood at x', log q(x') through the flow on tcgen05, accept)
 * and the final store in ONE kernel - a particle is read once and written once per sweep instead of three passes
 * over the particle set per move (asp_pcn_propose_eval + asp_flow_logprob_bf16 + asp_pcn_accept_scalar, which walk
 * the same trajectory).  Arguments as asp_pcn_sweep, with the flow descriptor in place of the proposal density;
 * n_moves <= 256; rows 32-byte aligned; d_ws: asp_pcn_sweep_flow_workspace_bytes(dims) bytes of scratch (the
 * current points and proposals of the tiles in flight; L2-resident).  Built for dims 16, 32, 64, a Gaussian prior
 * and a Gaussian or mixture likelihood (asp_pcn_sweep_flow_supported); a tensor-core barrier time-out sets
 * flow->d_error bit 0. */
int32_t asp_pcn_sweep_flow_supported(int32_t dims, int32_t like_kind, int32_t prior_kind);
size_t asp_pcn_sweep_flow_workspace_bytes(int32_t dims);
int asp_pcn_sweep_flow(const double* d_x, const double* d_log_l, const double* d_log_p, const double* d_log_q,
                       const int64_t* d_idx, int64_t n, int32_t dims, const asp_density_t* like,
                       const asp_density_t* prior, const asp_flow_t* flow, const asp_pcn_params_t* params,
                       double* d_x_out, double* d_log_l_out, double* d_log_p_out, double* d_log_q_out,
                       unsigned long long* d_accepts, void* d_ws, void* stream);

/* ---- coupling-flow fit on the device (8f-1) ----------------------------------
 * The optimisation loop of ZukoFlow.fit (src/aspire/flows/torch/flows.py:170-325: Adam on the mean negative
 * log-likelihood of mini-batches): forward, backward and Adam kernels over the flow's fp32 master weights.
 * One layer's parameters are laid out  W1[64, K/2] | b1[64] | W2[64, 64] | b2[64] | W3[K, 64] | b3[K]  (row-major,
 * K = kdims); d_w / d_m / d_v / d_g hold n_layers such blocks (asp_flow_fit_params floats each).  d_out_mask
 * (optional) is [n_layers][K/2] bytes: 0 marks a padding coordinate of the half the layer transforms - its
 * shift / log-scale rows get zero gradient, so they stay zero.  d_ws: asp_flow_fit_workspace_floats floats.
 * asp_flow_fit_step: d_loss[0] = mean_s(-log q(x_s)) of the batch BEFORE the update (log_const = K/2 log 2 pi +
 * sum log std is added), then one Adam step (torch.optim.Adam semantics; step = 1, 2, ... for the bias correction;
 * clip_grad > 0: global-norm clipping as torch.nn.utils.clip_grad_norm_).  asp_flow_fit_loss: forward only. */
typedef struct asp_flow_train {
    int32_t kdims, hidden, n_layers, batch_cap;
    float *d_w, *d_m, *d_v, *d_g;
    const unsigned char* d_out_mask;
    const float* d_mean;
    const float* d_inv_std;
    float* d_ws;
} asp_flow_train_t;
size_t asp_flow_fit_params(int32_t kdims, int32_t n_layers);
size_t asp_flow_fit_workspace_floats(int32_t kdims, int32_t n_layers, int32_t batch_cap);
int asp_flow_fit_step(const asp_flow_train_t* t, const float* d_x, int32_t b, float lr, float beta1, float beta2,
                      float eps, int32_t step, float clip_grad, float log_const, float* d_loss, void* stream);
int asp_flow_fit_loss(const asp_flow_train_t* t, const float* d_x, int32_t b, float log_const, float* d_loss,
                      void* stream);

/* ---- owner-computes multinomial resampling across GPUs (K12) ---------------
 * (no reference site: the reference is single-process.)  Every rank walks the same
 * global stream of n_total uniforms - Philox (seed, stream_id), identical to
 * asp_philox_uniform, or the supplied device array d_u_all (parity mode) - and
 * keeps, in slot order, the (slot j, u_j) with lo <= u_j < hi, i.e. the output slots
 * whose ancestor lives in this rank's CDF slice.  d_count receives how many; at most
 * `capacity` are written. */
size_t asp_select_workspace_bytes(void);
int asp_select_uniforms(uint64_t seed, uint32_t stream_id, const double* d_u_all, int64_t n_total, double lo, double hi,
                        int64_t capacity, int64_t* d_slot, double* d_u, long long* d_count, void* ws, void* stream);

/* ---- counter-based RNG (K13) --------------------------------------------
 * Philox4x32-10; layouts in oracle/philox.py.  uniform: fp64 in [0,1) with
 * numpy's 53-bit conversion; normal: fp64 Box-Muller, out[n, dims] =
 * mean_j + sigma_j * z (d_mean / d_sigma may be NULL for standard normals). */
int asp_philox_uniform(uint64_t seed, uint32_t stream_id, int64_t index0, int64_t n, double* d_out, void* stream);
int asp_philox_normal(uint64_t seed, uint32_t stream_id, int64_t index0, int64_t n, int32_t dims,
                      const double* d_mean, const double* d_sigma, double* d_out, void* stream);

/* PCG64 (numpy's default bit generator): d_out[j] = draw (offset + j) of Generator.random() for a
 * generator whose bit_generator.state is {state, inc} (128-bit, given as hi/lo words).  Replaces
 * the host draw + upload of numpy.random.Generator.choice's uniforms (src/aspire/samples.py:1278):
 * the same seed gives bit-identical ancestors.  The caller advances the host generator by n
 * (bit_generator.advance) to keep it in step. */
int asp_pcg64_uniform(uint64_t state_hi, uint64_t state_lo, uint64_t inc_hi, uint64_t inc_lo, int64_t offset,
                      int64_t n, double* d_out, void* stream);

/* ---- preconditioning / bounded-parameter transforms (scope row 8f-4) ------
 * CompositeTransform.forward / inverse, src/aspire/transforms.py:271-316: per dimension a periodic wrap
 * (:425-431) or a bounded -> unbounded map (logit :591-601 with utils.py:196-245, or probit :548-565; the
 * unit-interval value is clamped to [eps, 1 - eps]), then an affine standardisation of every dimension
 * (:624-634).  forward applies periodic, bounded, affine; inverse undoes them in the opposite order.  One pass:
 * d_out[n, dims] and d_log_j[n] = log |det J| of the applied direction (d_log_j may be NULL).  d_kind[dims]:
 * 0 untouched, 1 periodic, 2 bounded (NULL = all untouched); lower / upper are read for kinds 1 and 2; mean /
 * std (fitted by the caller: CompositeTransform.fit, :250-269) when affine != 0.  dims <= 256. */
enum asp_bounded_kind { ASP_BOUNDED_NONE = 0, ASP_BOUNDED_LOGIT = 1, ASP_BOUNDED_PROBIT = 2 };
typedef struct asp_transform {
    int32_t dims;
    int32_t bounded; /* asp_bounded_kind used by the dimensions of kind 2 */
    int32_t affine;
    int32_t reserved;
    double eps;
    const int32_t* d_kind;  /* device */
    const double* d_lower;  /* device */
    const double* d_upper;  /* device */
    const double* d_mean;   /* device */
    const double* d_std;    /* device */
} asp_transform_t;
int asp_transform_forward(const asp_transform_t* t, const double* d_x, int64_t n, double* d_y, double* d_log_j,
                          void* stream);
int asp_transform_inverse(const asp_transform_t* t, const double* d_y, int64_t n, double* d_x, double* d_log_j,
                          void* stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif

#ifdef __cplusplus
}
#endif
#endif /* ASPIRE_B200_H */

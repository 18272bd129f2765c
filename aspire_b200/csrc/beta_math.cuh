// Decision logic of the device-driven beta bisection (csrc/reweight.cu), host + device so that the CPU test suite can
// drive the shipped state machine with numpy-computed sums and compare the beta it lands on with the reference's
// SMCSampler.determine_beta (/root/reference/src/aspire/samplers/smc/base.py:166-186) - tests/native/beta_host.cu.
#pragma once

#include "../../include/aspire_b200.h"

#if defined(__CUDACC__)
#define ASP_BETA_HD __host__ __device__ __forceinline__
#else
#define ASP_BETA_HD inline
#endif

namespace asp {

ASP_BETA_HD void beta_init_state(asp_beta_state_t* st, double beta_prev, double target_eff, double tol, double n_total,
                                 double d_min, double d_max) {
    st->beta_prev = beta_prev;
    st->beta_min = beta_prev;
    st->beta_max = 1.0;
    st->target_eff = target_eff;
    st->tol = tol;
    st->n_total = n_total;
    st->beta_try = 1.0;
    st->s1 = st->s2 = 0.0;
    st->eff_beta_max = 0.0;
    st->s1_min = st->s2_min = 0.0;
    st->d_min = d_min;
    st->d_max = d_max;
    st->phase = 0;
    st->n_evals = 0;
    st->min_valid = 0;
    st->reserved = 0;
}

// one level: (s1, s2) = sums of w and w^2 at beta_try
ASP_BETA_HD void beta_decide(asp_beta_state_t* st) {
    if (st->phase == 2) return;
    const double eff = (st->s1 * st->s1 / st->s2) / st->n_total;
    if (st->phase == 0) {
        st->eff_beta_max = eff;
        if (eff >= st->target_eff) {  // smc/base.py:174-175
            st->beta_min = 1.0;
            st->s1_min = st->s1;
            st->s2_min = st->s2;
            st->min_valid = 1;
        }
        st->phase = 1;
    } else {
        if (eff >= st->target_eff) {  // smc/base.py:182-185
            st->beta_min = st->beta_try;
            st->s1_min = st->s1;
            st->s2_min = st->s2;
            st->min_valid = 1;
        } else {
            st->beta_max = st->beta_try;
        }
    }
    st->n_evals += 1;
    if (st->beta_max - st->beta_min > st->tol) {  // smc/base.py:177-178
        st->beta_try = 0.5 * (st->beta_max + st->beta_min);
    } else {
        st->phase = 2;
    }
}

// Three bisection levels per pass.  The seven betas a 3-level bisection of [beta_min, beta_max] can visit
// are beta_min + k h, h = (beta_max - beta_min) / 8, so their weights are E r^k with
// E = exp((beta_min - beta_prev)(d - d_max)),  r = exp(h (d - d_max)):  two exps and seven multiplies per
// particle instead of three dependent passes (each one exp).  The decisions are the reference's
// (smc/base.py:177-185: beta_try = 0.5 (max + min), ESS/N >= target), replayed on the seven (S1, S2) pairs.
constexpr int kBetaLevels = 3;
constexpr int kBetaCand = (1 << kBetaLevels) - 1;  // 7

ASP_BETA_HD void beta_decide3(asp_beta_state_t* st, const double* sums) {
    if (st->phase != 1) return;
    double lo = st->beta_min, hi = st->beta_max;
    int k_lo = 0, k_hi = 1 << kBetaLevels;  // grid indices of lo / hi
    for (int level = 0; level < kBetaLevels; ++level) {
        if (!(hi - lo > st->tol)) break;
        const double mid = 0.5 * (hi + lo);  // the reference's beta_try
        const int k = (k_lo + k_hi) >> 1;
        const double s1 = sums[2 * (k - 1)], s2 = sums[2 * (k - 1) + 1];
        const double eff = (s1 * s1 / s2) / st->n_total;
        if (eff >= st->target_eff) {
            lo = mid;
            k_lo = k;
            st->s1_min = s1;
            st->s2_min = s2;
            st->min_valid = 1;
        } else {
            hi = mid;
            k_hi = k;
        }
        st->s1 = s1;
        st->s2 = s2;
        st->n_evals += 1;
    }
    st->beta_min = lo;
    st->beta_max = hi;
    if (hi - lo > st->tol) st->beta_try = 0.5 * (hi + lo);
    else st->phase = 2;
}

// ---------------------------------------------------------------------------------------------------------
// Root-guided search.  ESS(beta) is non-increasing in beta - beta_prev (d/dt log ESS = 2 (E_t[d] - E_2t[d]) <= 0:
// a tilted mean grows with the tilt), so every decision "ESS(mid)/N >= target" of the reference's bisection is
// "mid <= beta*" for the root beta* of ESS/N = target.  The search therefore (1) estimates beta* - a bisection
// on a subsample, polished by Newton passes over all particles - (2) replays the reference's recursion
// (mid = 0.5 (max + min), smc/base.py:177-185) towards the estimate down to the tree node of width <= 8 tol
// that contains it, and (3) evaluates, in ONE pass, the node's end points and its seven interior candidates
// with the reference's formula: the end points prove that the root is inside (eff(a) >= target > eff(b)), the
// interior ones take the last three decisions exactly as before (beta_decide3).  Every decision the reference
// takes on the way down to the node is implied by those two verified end points and monotonicity, so beta, the
// bracket and n_evals are the reference's.  If the root is not inside, the neighbouring node is tried; after
// three misses the plain three-level bisection runs from the top (the state is untouched until a node succeeds).
//
// scratch (device doubles): [0, 18) = {S1_k, S2_k}, k = 0..8, of the node pass; then the root estimate, the last
// Newton step and the number of missed nodes.
// ---------------------------------------------------------------------------------------------------------
constexpr int kBsRoot = 20, kBsStep = 21, kBsTries = 22, kBsWords = 24;
constexpr int kBsNode = 9;  // candidates of the node pass: a + k h, k = 0..8

// the node [a, b] of the reference's bisection tree (width <= 8 tol) that contains `root`
ASP_BETA_HD void beta_node(const asp_beta_state_t* st, double root, double& a, double& b, int& depth) {
    double lo = st->beta_prev, hi = 1.0;
    depth = 0;
    while (hi - lo > 8.0 * st->tol && depth < 1100) {
        const double mid = 0.5 * (hi + lo);
        if (mid <= root) lo = mid;
        else hi = mid;
        ++depth;
    }
    a = lo;
    b = hi;
}

// Newton step on log ESS as a function of log(beta - beta_prev); (S1, S2) = sums of w, w^2; (D1, D2) = sums of
// t w, t w^2 with t = d - d_max, w = exp((root - beta_prev) t)
ASP_BETA_HD void beta_newton_update(const asp_beta_state_t* st, double* scratch, double s1, double s2, double d1,
                                    double d2) {
    const double delta = scratch[kBsRoot] - st->beta_prev;
    const double f = log((s1 * s1 / s2) / st->n_total) - log(st->target_eff);
    const double fp = 2.0 * (d1 / s1 - d2 / s2);  // d log ESS / d delta <= 0
    double nd = delta;
    if (fp < 0.0 && f == f) {
        double g = -f / (fp * delta);  // step in log delta
        g = g > 1.3862943611198906 ? 1.3862943611198906 : (g < -1.3862943611198906 ? -1.3862943611198906 : g);
        nd = delta * exp(g);           // within a factor 4 of the current estimate
    }
    if (nd > 1.0 - st->beta_prev) nd = 1.0 - st->beta_prev;
    if (!(nd > 0.0)) nd = delta;
    scratch[kBsStep] = nd > delta ? nd - delta : delta - nd;
    scratch[kBsRoot] = st->beta_prev + nd;
}

// decision of the node pass: scratch[0, 18) hold the sums at a + k h.  -> 1 when the search is finished
ASP_BETA_HD int beta_node_decide(asp_beta_state_t* st, double* scratch) {
    if (st->phase != 1) return 1;
    double a, b;
    int depth;
    beta_node(st, scratch[kBsRoot], a, b, depth);
    const double eff_a = (scratch[0] * scratch[0] / scratch[1]) / st->n_total;
    const double eff_b = (scratch[16] * scratch[16] / scratch[17]) / st->n_total;
    const bool a_ok = a == st->beta_prev || eff_a >= st->target_eff;  // beta_prev itself: ESS = N
    const bool b_ok = b == 1.0 || !(eff_b >= st->target_eff);         // beta = 1 failed in phase 0
    if (a_ok && b_ok) {
        st->beta_min = a;
        st->beta_max = b;
        if (a > st->beta_prev) {  // an evaluated beta of the reference's search
            st->s1_min = scratch[0];
            st->s2_min = scratch[1];
            st->min_valid = 1;
        }
        st->n_evals = 1 + depth;  // beta = 1, then one evaluation per level down to the node
        if (b - a > st->tol) st->beta_try = 0.5 * (b + a);
        beta_decide3(st, scratch + 2);
        if (!(st->beta_max - st->beta_min > st->tol)) st->phase = 2;
        return 1;
    }
    const double h = (b - a) * 0.125;
    scratch[kBsRoot] = a_ok ? b + 0.5 * h : a - 0.5 * h;  // the neighbouring node
    scratch[kBsTries] += 1.0;
    return 0;
}

}  // namespace asp

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

}  // namespace asp

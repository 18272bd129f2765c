// Scalar arithmetic of the Student-t EM fit (csrc/student_t.cu), host + device so that the CPU suite can check it
// against scipy and against oracle/pcn.py::fit_student_t_diag (tests/native/student_t_host.cu,
// tests/test_student_t_math_host.py): digamma, the nu equation of the M step, and the M step of one dimension from the
// sums one pass over the rows accumulates about the CURRENT mean.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define ASP_T_HD __host__ __device__ __forceinline__
#else
#define ASP_T_HD inline
#endif

namespace asp {

// digamma for x > 0: recurrence up to x >= 10, then the asymptotic series (error < 1e-15)
ASP_T_HD double digamma_pos(double x) {
    double r = 0.0;
    while (x < 10.0) {
        r -= 1.0 / x;
        x += 1.0;
    }
    const double i = 1.0 / x, i2 = i * i;
    const double series = i2 * (1.0 / 12.0 - i2 * (1.0 / 120.0 - i2 * (1.0 / 252.0 - i2 * (1.0 / 240.0 - i2 * (1.0 / 132.0)))));
    return r + log(x) - 0.5 * i - series;
}

// nu solves  -psi(nu/2) + log(nu/2) + 1 + mean(log w - w) + psi((nu_old + D)/2) - log((nu_old + D)/2) = 0  by 60
// bisection steps in log nu over [nu_min, nu_max] (clamped to the ends when the root is outside), as the spec
ASP_T_HD double student_t_nu(double nu_old, int dims, double mean_logw_minus_w, double nu_min, double nu_max) {
    const double hd = 0.5 * (nu_old + (double)dims);
    const double konst = 1.0 + mean_logw_minus_w + digamma_pos(hd) - log(hd);
    if (-digamma_pos(0.5 * nu_max) + log(0.5 * nu_max) + konst > 0.0) return nu_max;  // Gaussian-like data
    if (-digamma_pos(0.5 * nu_min) + log(0.5 * nu_min) + konst < 0.0) return nu_min;
    double lo = nu_min, hi = nu_max;
    for (int it = 0; it < 60; ++it) {
        const double mid = sqrt(lo * hi);
        if (-digamma_pos(0.5 * mid) + log(0.5 * mid) + konst > 0.0) lo = mid;
        else hi = mid;
    }
    return sqrt(lo * hi);
}

// M step of one dimension from  sw = sum w,  a = sum w (x - mu),  b = sum w (x - mu)^2  over m rows:
//   mu_new = mu + a / sw,   sigma_new^2 = sum w (x - mu_new)^2 / m = (b - a^2 / sw) / m
ASP_T_HD void student_t_mstep(double sw, double a, double b, double m, double& mu, double& sigma) {
    mu = mu + a / sw;
    sigma = sqrt(fmax(b - a * a / sw, 0.0) / m);
}

}  // namespace asp

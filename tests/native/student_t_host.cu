// Test harness (not part of the product library): exposes the SHIPPED scalar arithmetic of the Student-t EM fit
// (aspire_b200/csrc/student_t_math.cuh: digamma, the nu equation, the per-dimension M step) to the CPU tests, which
// compare it with scipy and replay oracle/pcn.py::fit_student_t_diag with it.  Built by
// tests/test_student_t_math_host.py with nvcc (host only).
#include "../../aspire_b200/csrc/student_t_math.cuh"

extern "C" {
double t_host_digamma(double x) { return asp::digamma_pos(x); }
double t_host_nu(double nu_old, int dims, double mean_logw_minus_w, double nu_min, double nu_max) {
    return asp::student_t_nu(nu_old, dims, mean_logw_minus_w, nu_min, nu_max);
}
void t_host_mstep(double sw, double a, double b, double m, double* mu, double* sigma) {
    asp::student_t_mstep(sw, a, b, m, *mu, *sigma);
}
}

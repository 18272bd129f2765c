// Per-element arithmetic of the preconditioning transforms (csrc/transform.cu), host + device so that the CPU test
// suite can run the very same code against the reference's golden vectors (tests/native/transform_host.cu).
// Reference: /root/reference/src/aspire/transforms.py:425-431 (periodic), :476-512 (unit interval), :548-565
// (probit), :591-601 + utils.py:196-245 (logit / sigmoid with the eps clamp), :624-634 (affine).
#pragma once

#include <math.h>

#include "../../include/aspire_b200.h"

#if defined(__CUDACC__)
#define ASP_HD __host__ __device__ __forceinline__
#else
#define ASP_HD inline
#endif

namespace asp {

// products and sums the reference rounds separately (a * b + c in numpy is two roundings): never contracted
ASP_HD double mul_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    volatile double p = a * b;
    return p;
#endif
}
ASP_HD double add_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    volatile double s = a + b;
    return s;
#endif
}

ASP_HD double wrap_periodic(double v, double lo, double w) {
    // numpy's floor-mod for w > 0: fmod, then shifted into [0, w)
    double m = fmod(add_rn(v, -lo), w);
    if (m != 0.0 && m < 0.0) m = add_rn(m, w);
    return add_rn(lo, m);
}

ASP_HD double clamp_unit(double u, double eps) {
    // np.clip(u, eps, 1 - eps); NaN propagates
    if (u < eps) u = eps;
    if (u > 1.0 - eps) u = 1.0 - eps;
    return u;
}

struct TransformDim {  // constants of one dimension
    double lo, width, mean, sd;
    double cj;  // constant Jacobian terms of the forward direction: -log(width) if bounded, -log|sd| if affine
    int kind;   // 0 untouched, 1 periodic, 2 bounded
};

ASP_HD TransformDim transform_dim(int kind, double lower, double upper, bool affine, double mean, double sd) {
    TransformDim c;
    c.kind = kind;
    c.lo = kind != 0 ? lower : 0.0;
    c.width = kind != 0 ? upper - lower : 1.0;
    c.mean = affine ? mean : 0.0;
    c.sd = affine ? sd : 1.0;
    c.cj = 0.0;
    if (kind == 2) c.cj -= log(c.width);
    if (affine) c.cj -= log(fabs(sd));
    return c;
}

// v: value in / out; returns this element's term of log |det J|
template <bool INVERSE>
ASP_HD double transform_element(const TransformDim& c, int bounded, bool affine, double eps, double& v) {
    const double kLog2Pi = 1.8378770664093453;
    const double kSqrt2 = 1.4142135623730951;
    double term;
    if (!INVERSE) {
        term = c.cj;
        if (c.kind == 1) {
            v = wrap_periodic(v, c.lo, c.width);
        } else if (c.kind == 2) {
            const double u = clamp_unit((v - c.lo) / c.width, eps);
            if (bounded == ASP_BOUNDED_LOGIT) {
                const double lu = log(u), l1 = log1p(-u);
                v = lu - l1;
                term += -lu - l1;
            } else {
                v = mul_rn(erfinv(add_rn(mul_rn(2.0, u), -1.0)), kSqrt2);
                term += 0.5 * add_rn(kLog2Pi, mul_rn(v, v));
            }
        }
        if (affine) v = add_rn(v, -c.mean) / c.sd;
    } else {
        term = -c.cj;
        if (affine) v = add_rn(mul_rn(v, c.sd), c.mean);
        if (c.kind == 2) {
            double s;
            if (bounded == ASP_BOUNDED_LOGIT) {
                s = clamp_unit(1.0 / (1.0 + exp(-v)), eps);
                term += log(s) + log1p(-s);
            } else {
                term -= 0.5 * add_rn(kLog2Pi, mul_rn(v, v));
                s = 0.5 * (1.0 + erf(v / kSqrt2));
            }
            v = add_rn(mul_rn(c.width, s), c.lo);
        } else if (c.kind == 1) {
            v = wrap_periodic(v, c.lo, c.width);
        }
    }
    return term;
}

}  // namespace asp

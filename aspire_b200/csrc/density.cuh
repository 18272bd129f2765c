// Built-in log-densities evaluated by a group of G lanes per particle, four
// dimensions per lane (row of D doubles = one contiguous 128-bit-vectorised
// burst per group).  Spec: oracle/densities.py; parameterisation: asp_density_t.
//
// Everything is specialised on the density KIND at compile time: a runtime
// switch around the group shuffles makes ptxas wrap every SHFL in a
// WARPSYNC/ENDCOLLECTIVE convergence sequence and blows the register budget
// (seen in the SASS of the first sweep kernel of round 1).
#pragma once
#include "common.cuh"

namespace asp {

constexpr int kDPL = 4;  // dimensions per lane

template <int G>
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, G);
    return v;
}
template <int G>
__device__ __forceinline__ double group_max(double v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o, G));
    return v;
}

// Per-lane view of one density of compile-time kind KIND.
template <int KIND>
struct Dens {
    double a[kDPL], b[kDPL];  // GAUSSIAN: mean, 1/sigma;  UNIFORM: lo, hi  (unused otherwise)
    double s0, s1, s2;
    const double *p0, *p1, *p2;
    int n_comp;

    __device__ __forceinline__ void load(const asp_density_t& d, int lane, int dims) {
        s0 = d.s0;
        s1 = d.s1;
        s2 = d.s2;
        p0 = d.p0;
        p1 = d.p1;
        p2 = d.p2;
        n_comp = d.n_comp;
        if constexpr (KIND == ASP_DENSITY_GAUSSIAN || KIND == ASP_DENSITY_UNIFORM) {
#pragma unroll
            for (int k = 0; k < kDPL; ++k) {
                const int j = lane * kDPL + k;
                a[k] = j < dims ? __ldg(d.p0 + j) : 0.0;
                b[k] = j < dims ? __ldg(d.p1 + j) : 0.0;
            }
        }
    }

    template <int G>
    __device__ __forceinline__ double eval(const double (&x)[kDPL], int lane, int dims) const {
        const double ninf = identity(OP_MAX);
        if constexpr (KIND == ASP_DENSITY_GAUSSIAN) {
            // padded dimensions have b == 0, so they contribute exactly 0
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < kDPL; ++k) {
                const double z = (x[k] - a[k]) * b[k];
                s += z * z;
            }
            return s0 - 0.5 * group_sum<G>(s);
        } else if constexpr (KIND == ASP_DENSITY_ROSENBROCK) {
            const double xn = __shfl_down_sync(0xffffffffu, x[0], 1, G);  // first dim of the next lane
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < kDPL; ++k) {
                const double x1 = k + 1 < kDPL ? x[(k + 1) % kDPL] : xn;
                const double t = x1 - x[k] * x[k];
                const double u = s1 - x[k];
                s += (lane * kDPL + k + 1 < dims) ? s0 * t * t + u * u : 0.0;
            }
            return -group_sum<G>(s) / s2;
        } else if constexpr (KIND == ASP_DENSITY_GMM) {
            // lane owns components c with c % G == lane; streaming log-sum-exp, no arrays
            double m = ninf, acc = 0.0;
            for (int c = 0; c < n_comp; ++c) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < kDPL; ++k) {
                    const int j = lane * kDPL + k;
                    const double dlt = x[k] - (j < dims ? __ldg(p0 + (size_t)c * dims + j) : x[k]);
                    s += dlt * dlt;
                }
                const double d2 = group_sum<G>(s);
                if (c % G == lane) {
                    const double is = __ldg(p1 + c);
                    const double t = __ldg(p2 + c) - 0.5 * (is * is) * d2;
                    if (t > m) {
                        acc = acc * exp(m - t) + 1.0;  // exp(-inf) = 0 on the first component
                        m = t;
                    } else {
                        acc += exp(t - m);
                    }
                }
            }
            const double mm = group_max<G>(m);
            const double e = (m > ninf) ? acc * exp(m - mm) : 0.0;
            return mm + log(group_sum<G>(e));
        } else {  // UNIFORM
            double bad = 0.0;
#pragma unroll
            for (int k = 0; k < kDPL; ++k)
                if (lane * kDPL + k < dims) bad += (x[k] > a[k] && x[k] < b[k]) ? 0.0 : 1.0;
            return group_sum<G>(bad) > 0.0 ? ninf : s0;
        }
    }
};

// One-off evaluation straight from the descriptor (parameters through L1, no
// per-lane register state): used where a density is evaluated once per particle.
template <int KIND, int G>
__device__ __forceinline__ double eval_once(const asp_density_t& d, const double (&x)[kDPL], int lane, int dims) {
    if constexpr (KIND == ASP_DENSITY_GAUSSIAN) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < kDPL; ++k) {
            const int j = lane * kDPL + k;
            if (j < dims) {
                const double z = (x[k] - __ldg(d.p0 + j)) * __ldg(d.p1 + j);
                s += z * z;
            }
        }
        return d.s0 - 0.5 * group_sum<G>(s);
    } else if constexpr (KIND == ASP_DENSITY_UNIFORM) {
        double bad = 0.0;
#pragma unroll
        for (int k = 0; k < kDPL; ++k) {
            const int j = lane * kDPL + k;
            if (j < dims) bad += (x[k] > __ldg(d.p0 + j) && x[k] < __ldg(d.p1 + j)) ? 0.0 : 1.0;
        }
        return group_sum<G>(bad) > 0.0 ? identity(OP_MAX) : d.s0;
    } else {
        Dens<KIND> t;
        t.load(d, lane, dims);
        return t.template eval<G>(x, lane, dims);
    }
}

// load / store the lane's slice of row `row` ([n, dims] row-major)
__device__ __forceinline__ void row_load(const double* __restrict__ x, int64_t row, int dims, int lane, bool vec,
                                         double (&v)[kDPL]) {
    const double* src = x + row * dims + lane * kDPL;
    if (vec && lane * kDPL + kDPL <= dims) {
        const double2 a = *reinterpret_cast<const double2*>(src);
        const double2 b = *reinterpret_cast<const double2*>(src + 2);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    } else {
#pragma unroll
        for (int k = 0; k < kDPL; ++k) v[k] = (lane * kDPL + k < dims) ? src[k] : 0.0;
    }
}
__device__ __forceinline__ void row_store(double* __restrict__ x, int64_t row, int dims, int lane, bool vec,
                                          const double (&v)[kDPL]) {
    double* dst = x + row * dims + lane * kDPL;
    if (vec && lane * kDPL + kDPL <= dims) {
        st_stream2(dst, v[0], v[1]);
        st_stream2(dst + 2, v[2], v[3]);
    } else {
#pragma unroll
        for (int k = 0; k < kDPL; ++k)
            if (lane * kDPL + k < dims) dst[k] = v[k];
    }
}

inline int group_size_for(int dims) {
    int g = 1;
    while (g < 32 && g * kDPL < dims) g <<= 1;
    return g;
}

}  // namespace asp

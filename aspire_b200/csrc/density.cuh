// Built-in log-densities evaluated by a group of G lanes per particle, four
// dimensions per lane (row of D doubles = one contiguous 128-bit-vectorised
// burst per group).  Spec: oracle/densities.py; parameterisation: asp_density_t.
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

// per-lane view of one density: the lane's slice of the per-dimension parameters
struct DensityLane {
    int kind, n_comp;
    double a[kDPL], b[kDPL];  // GAUSSIAN: mean, 1/sigma;  UNIFORM: lo, hi
    double s0, s1, s2;
    const double *p0, *p1, *p2;
};

__device__ __forceinline__ void density_load(DensityLane& L, const asp_density_t& d, int lane, int dims) {
    L.kind = d.kind;
    L.n_comp = d.n_comp;
    L.s0 = d.s0;
    L.s1 = d.s1;
    L.s2 = d.s2;
    L.p0 = d.p0;
    L.p1 = d.p1;
    L.p2 = d.p2;
#pragma unroll
    for (int k = 0; k < kDPL; ++k) {
        const int j = lane * kDPL + k;
        L.a[k] = 0.0;
        L.b[k] = 0.0;
        if ((d.kind == ASP_DENSITY_GAUSSIAN || d.kind == ASP_DENSITY_UNIFORM) && j < dims) {
            L.a[k] = __ldg(d.p0 + j);
            L.b[k] = __ldg(d.p1 + j);
        }
    }
}

template <int G>
__device__ __forceinline__ double density_eval(const DensityLane& L, const double (&x)[kDPL], int lane, int dims) {
    const double ninf = identity(OP_MAX);
    switch (L.kind) {
        case ASP_DENSITY_GAUSSIAN: {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < kDPL; ++k) {
                const double z = (x[k] - L.a[k]) * L.b[k];
                s += (lane * kDPL + k < dims) ? z * z : 0.0;
            }
            return L.s0 - 0.5 * group_sum<G>(s);
        }
        case ASP_DENSITY_ROSENBROCK: {
            const double xn = __shfl_down_sync(0xffffffffu, x[0], 1, G);  // first dim of the next lane
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < kDPL; ++k) {
                const double x1 = k + 1 < kDPL ? x[(k + 1) % kDPL] : xn;
                const double t = x1 - x[k] * x[k];
                const double u = L.s1 - x[k];
                s += (lane * kDPL + k + 1 < dims) ? L.s0 * t * t + u * u : 0.0;
            }
            return -group_sum<G>(s) / L.s2;
        }
        case ASP_DENSITY_GMM: {
            double t_own = ninf;   // this lane owns components k with k % G == lane
            double tk[(32 + G - 1) / G];
#pragma unroll
            for (int o = 0; o < (32 + G - 1) / G; ++o) tk[o] = ninf;
            for (int c = 0; c < L.n_comp; ++c) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < kDPL; ++k) {
                    const int j = lane * kDPL + k;
                    if (j < dims) {
                        const double dlt = x[k] - __ldg(L.p0 + (size_t)c * dims + j);
                        s += dlt * dlt;
                    }
                }
                const double d2 = group_sum<G>(s);
                if (c % G == lane) {
                    const double is = __ldg(L.p1 + c);
                    const double t = __ldg(L.p2 + c) - 0.5 * (is * is) * d2;
#pragma unroll
                    for (int o = 0; o < (32 + G - 1) / G; ++o)
                        if (o == c / G) tk[o] = t;
                    t_own = fmax(t_own, t);
                }
            }
            const double m = group_max<G>(t_own);
            double e = 0.0;
#pragma unroll
            for (int o = 0; o < (32 + G - 1) / G; ++o)
                if (tk[o] > ninf) e += exp(tk[o] - m);
            const double se = group_sum<G>(e);
            return m + log(se);
        }
        case ASP_DENSITY_UNIFORM: {
            double bad = 0.0;
#pragma unroll
            for (int k = 0; k < kDPL; ++k)
                if (lane * kDPL + k < dims) bad += (x[k] > L.a[k] && x[k] < L.b[k]) ? 0.0 : 1.0;
            return group_sum<G>(bad) > 0.0 ? ninf : L.s0;
        }
    }
    return __longlong_as_double(0x7ff8000000000000LL);
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

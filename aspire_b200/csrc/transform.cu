// Preconditioning / bounded-parameter transforms: CompositeTransform.forward / inverse of
// /root/reference/src/aspire/transforms.py:271-316 as ONE pass over the particle rows - periodic wrap (:425-431),
// bounded -> unbounded by logit (:591-601, utils.py:196-245) or probit (:548-565) with the eps clamp, affine
// standardisation (:624-634) - writing the transformed row and its log |det J|.  Spec: oracle/transforms.py.
//
// HBM-bound elementwise work: 8 D B read + (8 D + 8) B written per row.  Elements are processed in flat,
// coalesced order; each element's Jacobian term goes to a shared-memory tile (odd row stride: conflict-free)
// and one thread per row adds its D terms in dimension order, so log |det J| does not depend on the launch shape.
#include <math.h>

#include "common.cuh"
#include "transform_math.cuh"

namespace asp {

constexpr int kTransformMaxDims = 256;
constexpr int kTransformTile = 2048;  // doubles of Jacobian terms staged per block iteration

template <bool INVERSE>
__global__ void __launch_bounds__(kBlock) transform_kernel(const asp_transform_t t, const double* __restrict__ in,
                                                           int64_t n, double* __restrict__ out,
                                                           double* __restrict__ log_j, int rows_per_tile, int stride) {
    extern __shared__ double s_mem[];
    const int D = t.dims;
    double* s_terms = s_mem;  // [rows_per_tile][stride]
    TransformDim* s_dim = reinterpret_cast<TransformDim*>(s_mem + (size_t)rows_per_tile * stride);
    for (int j = threadIdx.x; j < D; j += blockDim.x) {
        const int kind = t.d_kind ? t.d_kind[j] : 0;
        s_dim[j] = transform_dim(kind, kind ? t.d_lower[j] : 0.0, kind ? t.d_upper[j] : 1.0, t.affine != 0,
                                 t.affine ? t.d_mean[j] : 0.0, t.affine ? t.d_std[j] : 1.0);
    }
    __syncthreads();
    const int64_t n_tiles = (n + rows_per_tile - 1) / rows_per_tile;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * rows_per_tile;
        const int rows = (int)((n - row0) < rows_per_tile ? (n - row0) : rows_per_tile);
        const int elems = rows * D;
        const double* src = in + row0 * D;
        double* dst = out + row0 * D;
        for (int e = threadIdx.x; e < elems; e += blockDim.x) {
            const int r = e / D, j = e - r * D;
            double v = src[e];
            const double term = transform_element<INVERSE>(s_dim[j], t.bounded, t.affine != 0, t.eps, v);
            dst[e] = v;
            s_terms[r * stride + j] = term;
        }
        __syncthreads();
        if (log_j) {
            for (int r = threadIdx.x; r < rows; r += blockDim.x) {
                double acc = 0.0;
                for (int j = 0; j < D; ++j) acc += s_terms[r * stride + j];
                log_j[row0 + r] = acc;
            }
        }
        __syncthreads();
    }
}

static int transform_launch(bool inverse, const asp_transform_t* t, const double* d_in, int64_t n, double* d_out,
                            double* d_log_j, void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(t && d_in && d_out, "null pointer");
    ASP_REQUIRE(n >= 0, "n must not be negative");
    ASP_REQUIRE(t->dims >= 1 && t->dims <= kTransformMaxDims, "dims must be in [1, 256]");
    ASP_REQUIRE(t->bounded == ASP_BOUNDED_NONE || t->bounded == ASP_BOUNDED_LOGIT || t->bounded == ASP_BOUNDED_PROBIT,
                "unknown bounded transform");
    ASP_REQUIRE(!t->affine || (t->d_mean && t->d_std), "affine stage needs mean and std");
    ASP_REQUIRE(!t->d_kind || (t->d_lower && t->d_upper), "periodic / bounded dimensions need lower and upper");
    ASP_REQUIRE(t->eps >= 0.0 && t->eps < 0.5, "eps must be in [0, 0.5)");
    if (n == 0) return ASP_OK;
    const int D = t->dims;
    const int stride = D | 1;
    int rows = kTransformTile / stride;
    if (rows < 1) rows = 1;
    const size_t smem = (size_t)rows * stride * sizeof(double) + (size_t)D * sizeof(TransformDim);
    const int64_t n_tiles = (n + rows - 1) / rows;
    const int64_t cap = (int64_t)kNumSMs * 8;
    const int grid = (int)(n_tiles < cap ? n_tiles : cap);
    if (inverse)
        transform_kernel<true><<<grid, kBlock, smem, (cudaStream_t)stream>>>(*t, d_in, n, d_out, d_log_j, rows, stride);
    else
        transform_kernel<false><<<grid, kBlock, smem, (cudaStream_t)stream>>>(*t, d_in, n, d_out, d_log_j, rows, stride);
    return check_launch(inverse ? "asp_transform_inverse" : "asp_transform_forward");
}

}  // namespace asp

extern "C" {

int asp_transform_forward(const asp_transform_t* t, const double* d_x, int64_t n, double* d_y, double* d_log_j,
                          void* stream) {
    return asp::transform_launch(false, t, d_x, n, d_y, d_log_j, stream);
}

int asp_transform_inverse(const asp_transform_t* t, const double* d_y, int64_t n, double* d_x, double* d_log_j,
                          void* stream) {
    return asp::transform_launch(true, t, d_y, n, d_x, d_log_j, stream);
}

}  // extern "C"

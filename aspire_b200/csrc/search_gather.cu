// K6 ancestor search and K7 row gather.
//
// numpy Generator.choice (called at /root/reference/src/aspire/samples.py:1278):
//   idx = cdf.searchsorted(u, side='right')     -> asp_search_multinomial
// and the fancy-index copies of samples.py:1279-1287 -> asp_gather_rows.
// Both are HBM/L2 bound: the search is a branch-free binary search whose upper
// levels stay L2-resident; the gather moves whole rows with 128-bit accesses,
// one sub-warp group per row so every row is a single contiguous burst.
#include <stdlib.h>

#include "common.cuh"

namespace asp {

// count of cdf[k] <= u  (upper bound), branch-free
__device__ __forceinline__ int64_t upper_bound(const double* __restrict__ cdf, int64_t n, double u) {
    int64_t lo = 0, len = n;
    while (len > 0) {
        int64_t half = len >> 1;
        double v = __ldg(cdf + lo + half);
        bool right = v <= u;
        lo = right ? lo + half + 1 : lo;
        len = right ? len - half - 1 : half;
    }
    return lo;
}

__global__ void __launch_bounds__(kBlock) search_multinomial_kernel(const double* __restrict__ cdf, int64_t n,
                                                                    const double* __restrict__ u, int64_t m,
                                                                    int64_t* __restrict__ idx) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += stride)
        idx[j] = upper_bound(cdf, n, ld_stream(u + j));
}

// Systematic resampling, u_j = fl(fl(j + u0) / M): the queries are sorted, so instead of one 27-level binary
// search per slot the CDF is streamed once - entry k owns the slots j with cdf[k-1] <= u_j < cdf[k] and
// writes idx[j] = k there (same construction as the guide table below, with the offset u0).  Identical
// to upper_bound(cdf, u_j) by construction: the slot boundaries are fixed up with the same IEEE operations.
__device__ __forceinline__ double systematic_u(int64_t j, double u0, double m_total) {
    return __ddiv_rn(__dadd_rn(__ll2double_rn(j), u0), m_total);  // one fp64 add, one IEEE division, no contraction
}
__device__ __forceinline__ int64_t first_slot_geq(double c, double u0, int64_t M, double Md) {
    // smallest j in [0, M] with u_j >= c  (M if none)
    if (!(c > 0.0)) return 0;
    double e = ceil(c * Md - u0);
    int64_t j = e > (double)M ? M : (e < 0.0 ? 0 : (int64_t)e);
    while (j > 0 && systematic_u(j - 1, u0, Md) >= c) --j;
    while (j < M && systematic_u(j, u0, Md) < c) ++j;
    return j;
}

// ---- guide-table search ------------------------------------------------------
// guide[b] = #{k : cdf[k] <= b / B}, b = 0..B, built with sorted (hence coalesced,
// cache-friendly) queries.  A uniform u then needs guide[b], guide[b+1] for its
// bucket b (q_b <= u < q_{b+1}, fixed up with the same IEEE division) and a
// search over the ~1 CDF entries in between: ~3 random 32-byte sectors per
// ancestor instead of a 27-level dependent chain.  Result is identical to
// upper_bound(cdf, u) by construction.
template <typename GT>
__global__ void __launch_bounds__(kBlock) build_guide_kernel(const double* __restrict__ cdf, int64_t n, int64_t B,
                                                             GT* __restrict__ guide) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const double Bd = (double)B;
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b <= B; b += stride)
        guide[b] = (GT)upper_bound(cdf, n, __ddiv_rn(__ll2double_rn(b), Bd));
}

// Streaming build: CDF entry k owns the buckets b with cdf[k-1] <= q_b < cdf[k] (their
// guide value is k), so one coalesced pass over the CDF writes the whole table: O(n)
// instead of a binary search per bucket.  Short ranges are written by their thread,
// medium ones by the warp, ranges longer than 2^16 buckets (collapsed weights) are
// deferred to a second, full-grid pass.
constexpr int kLongRange = 1 << 16;
constexpr int kGuideShift = 0;  // n_cdf buckets; micro-benchmark at N = 1e8: 6.23 (0), 5.94 (2), 7.12 (3), 9.12 ms (4), but 2 was slower inside the tempering step
struct GuideDeferred {
    unsigned int count;
    unsigned int pad;
    long long entry[1][3];  // {b_lo, b_hi, k}; capacity = B / kLongRange + 2
};

__device__ __forceinline__ int64_t first_bucket_geq(double c, int64_t B, double Bd) {
    // smallest b in [0, B + 1] with q_b = b / B >= c   (B + 1 if none)
    if (!(c > 0.0)) return 0;
    double e = ceil(c * Bd);
    int64_t b = e > (double)(B + 1) ? B + 1 : (int64_t)e;
    while (b > 0 && __ddiv_rn(__ll2double_rn(b - 1), Bd) >= c) --b;
    while (b <= B && __ddiv_rn(__ll2double_rn(b), Bd) < c) ++b;
    return b;
}

template <typename GT>
__global__ void __launch_bounds__(kBlock) build_guide_stream_kernel(const double* __restrict__ cdf, int64_t n,
                                                                    int64_t B, GT* __restrict__ guide,
                                                                    GuideDeferred* __restrict__ deferred) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const double Bd = (double)B;
    const int lane = threadIdx.x & 31;
    // k = n is the virtual tail entry (cdf = +inf): buckets at or beyond cdf[n-1] count all n entries
    const int64_t first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t base = first - lane; base <= n; base += stride) {  // warp-uniform trip count
        const int64_t k = base + lane;
        int64_t b_lo = 0, b_hi = 0;
        if (k <= n) {
            b_lo = k > 0 ? first_bucket_geq(__ldg(cdf + k - 1), B, Bd) : 0;
            b_hi = k < n ? first_bucket_geq(__ldg(cdf + k), B, Bd) : B + 1;
        }
        const int64_t len = b_hi - b_lo;
        if (len > 0 && len <= 16) {
            for (int64_t b = b_lo; b < b_hi; ++b) guide[b] = (GT)k;
        } else if (len > kLongRange) {
            const unsigned int slot = atomicAdd(&deferred->count, 1u);
            deferred->entry[slot][0] = b_lo;
            deferred->entry[slot][1] = b_hi;
            deferred->entry[slot][2] = k;
        }
        unsigned int medium = __ballot_sync(0xffffffffu, len > 16 && len <= kLongRange);
        while (medium) {
            const int src = __ffs(medium) - 1;
            medium &= medium - 1;
            const int64_t lo = __shfl_sync(0xffffffffu, b_lo, src);
            const int64_t hi = __shfl_sync(0xffffffffu, b_hi, src);
            const int64_t kk = __shfl_sync(0xffffffffu, k, src);
            for (int64_t b = lo + lane; b < hi; b += 32) guide[b] = (GT)kk;
        }
    }
}

// slots [j0, j0 + m) of a sweep over m_total slots; idx is indexed by j - j0
__global__ void __launch_bounds__(kBlock) search_systematic_stream_kernel(const double* __restrict__ cdf, int64_t n,
                                                                          double u0, int64_t j0, int64_t m,
                                                                          int64_t m_total, int64_t* __restrict__ idx,
                                                                          GuideDeferred* __restrict__ deferred) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const double Md = (double)m_total;
    const int lane = threadIdx.x & 31;
    const int64_t first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t base = first - lane; base <= n; base += stride) {  // warp-uniform trip count; k = n: virtual tail
        const int64_t k = base + lane;
        int64_t lo = 0, hi = 0;
        if (k <= n) {
            lo = k > 0 ? first_slot_geq(__ldg(cdf + k - 1), u0, m_total, Md) : 0;
            hi = k < n ? first_slot_geq(__ldg(cdf + k), u0, m_total, Md) : m_total;
            lo = (lo < j0 ? j0 : lo) - j0;
            hi = (hi > j0 + m ? j0 + m : hi) - j0;
        }
        const int64_t val = k < n - 1 ? k : n - 1;
        const int64_t len = hi - lo;
        if (len > 0 && len <= 16) {
            for (int64_t j = lo; j < hi; ++j) idx[j] = val;
        } else if (len > kLongRange) {
            const unsigned int slot = atomicAdd(&deferred->count, 1u);
            deferred->entry[slot][0] = lo;
            deferred->entry[slot][1] = hi;
            deferred->entry[slot][2] = val;
        }
        unsigned int medium = __ballot_sync(0xffffffffu, len > 16 && len <= kLongRange);
        while (medium) {
            const int src = __ffs(medium) - 1;
            medium &= medium - 1;
            const int64_t l = __shfl_sync(0xffffffffu, lo, src);
            const int64_t h = __shfl_sync(0xffffffffu, hi, src);
            const int64_t v = __shfl_sync(0xffffffffu, val, src);
            for (int64_t j = l + lane; j < h; j += 32) idx[j] = v;
        }
    }
}

template <typename GT>
__global__ void __launch_bounds__(kBlock) build_guide_deferred_kernel(GT* __restrict__ guide,
                                                                      GuideDeferred* __restrict__ deferred) {
    const unsigned int count = deferred->count;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (unsigned int e = 0; e < count; ++e) {
        const int64_t lo = deferred->entry[e][0], hi = deferred->entry[e][1];
        const GT kk = (GT)deferred->entry[e][2];
        for (int64_t b = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < hi; b += stride) guide[b] = kk;
    }
}

__global__ void guide_reset_kernel(GuideDeferred* deferred) { deferred->count = 0; }

template <typename GT>
__global__ void __launch_bounds__(kBlock) search_guided_kernel(const double* __restrict__ cdf, int64_t n,
                                                               const GT* __restrict__ guide, int64_t B,
                                                               const double* __restrict__ u, int64_t m,
                                                               int64_t* __restrict__ idx) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const double Bd = (double)B;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += stride) {
        const double uj = ld_stream(u + j);
        int64_t b = (int64_t)(uj * Bd);
        b = b < 0 ? 0 : (b > B - 1 ? B - 1 : b);
        while (b > 0 && __ddiv_rn(__ll2double_rn(b), Bd) > uj) --b;
        while (b + 1 < B && __ddiv_rn(__ll2double_rn(b + 1), Bd) <= uj) ++b;
        const int64_t lo = (int64_t)__ldg(guide + b);
        const int64_t hi = (int64_t)__ldg(guide + b + 1);
        idx[j] = lo + upper_bound(cdf + lo, hi - lo, uj);
    }
}

// One group of G lanes moves one row of `dims` doubles (G * 2 doubles per step).
template <int G>
__global__ void __launch_bounds__(kBlock) gather_rows_kernel(const double* __restrict__ x, int dims,
                                                             const int64_t* __restrict__ idx, int64_t m,
                                                             double* __restrict__ xo, const double* __restrict__ l,
                                                             const double* __restrict__ p, const double* __restrict__ q,
                                                             double* __restrict__ lo, double* __restrict__ po,
                                                             double* __restrict__ qo) {
    const int lane = threadIdx.x % G;
    const int64_t groups = ((int64_t)gridDim.x * blockDim.x) / G;
    const bool vec = (dims % 2 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(xo) & 15) == 0);
    for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G; r < m; r += groups) {
        const int64_t a = idx[r];
        const double* src = x + a * dims;
        double* dst = xo + r * dims;
        if (vec) {
            for (int c = 2 * lane; c < dims; c += 2 * G) {
                double2 v = *reinterpret_cast<const double2*>(src + c);
                st_stream2(dst + c, v.x, v.y);
            }
        } else {
            for (int c = lane; c < dims; c += G) dst[c] = src[c];
        }
        if (lane == 0) {
            if (l) lo[r] = l[a];
            if (p) po[r] = p[a];
            if (q) qo[r] = q[a];
        }
    }
}

}  // namespace asp

using namespace asp;

extern "C" {

int asp_search_multinomial(const double* d_cdf, int64_t n_cdf, const double* d_u, int64_t m, int64_t* d_idx,
                           void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(d_cdf && d_u && d_idx, "null pointer");
    ASP_REQUIRE(n_cdf > 0 && m > 0, "sizes must be positive");
    search_multinomial_kernel<<<grid_for(m, 1), kBlock, 0, (cudaStream_t)stream>>>(d_cdf, n_cdf, d_u, m, d_idx);
    return check_launch("asp_search_multinomial");
}

static inline size_t guide_bytes(int64_t n_cdf) {
    return (((size_t)(n_cdf + 1) * (n_cdf < 0xFFFFFFFFLL ? 4 : 8)) + 255) & ~(size_t)255;
}
size_t asp_search_workspace_bytes(int64_t n_cdf) {
    if (n_cdf < 1) n_cdf = 1;
    return guide_bytes(n_cdf) + 256 + sizeof(long long) * 3 * (size_t)(n_cdf / kLongRange + 4);
}

int asp_search_multinomial_guided(const double* d_cdf, int64_t n_cdf, const double* d_u, int64_t m, int64_t* d_idx,
                                  void* ws, void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(d_cdf && d_u && d_idx && ws, "null pointer");
    ASP_REQUIRE(n_cdf > 0 && m > 0, "sizes must be positive");
    cudaStream_t s = (cudaStream_t)stream;
    // buckets: n_cdf >> kGuideShift.  Fewer buckets than CDF entries keep the table small enough to stay in
    // L2 while each query still resolves inside one or two 64-byte DRAM bursts of the CDF.
    static const int shift = [] { const char* e = getenv("ASP_GUIDE_SHIFT"); return e ? atoi(e) : kGuideShift; }();
    const int64_t B = (n_cdf >> shift) > 0 ? (n_cdf >> shift) : 1;
    GuideDeferred* def = (GuideDeferred*)((char*)ws + guide_bytes(n_cdf));
    const int gb = grid_for(n_cdf + 1, 1);
    guide_reset_kernel<<<1, 1, 0, s>>>(def);
    if (n_cdf < 0xFFFFFFFFLL) {
        build_guide_stream_kernel<uint32_t><<<gb, kBlock, 0, s>>>(d_cdf, n_cdf, B, (uint32_t*)ws, def);
        build_guide_deferred_kernel<uint32_t><<<kNumSMs * 4, kBlock, 0, s>>>((uint32_t*)ws, def);
        search_guided_kernel<uint32_t><<<grid_for(m, 1), kBlock, 0, s>>>(d_cdf, n_cdf, (const uint32_t*)ws, B, d_u, m,
                                                                         d_idx);
    } else {
        build_guide_stream_kernel<int64_t><<<gb, kBlock, 0, s>>>(d_cdf, n_cdf, B, (int64_t*)ws, def);
        build_guide_deferred_kernel<int64_t><<<kNumSMs * 4, kBlock, 0, s>>>((int64_t*)ws, def);
        search_guided_kernel<int64_t><<<grid_for(m, 1), kBlock, 0, s>>>(d_cdf, n_cdf, (const int64_t*)ws, B, d_u, m,
                                                                        d_idx);
    }
    return check_launch("asp_search_multinomial_guided");
}

size_t asp_search_systematic_workspace_bytes(int64_t m) {
    if (m < 1) m = 1;
    return 256 + sizeof(long long) * 3 * (size_t)(m / kLongRange + 4);
}

int asp_search_systematic(const double* d_cdf, int64_t n_cdf, double u0, int64_t j0, int64_t m, int64_t m_total,
                          int64_t* d_idx, void* ws, void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(d_cdf && d_idx, "null pointer");
    ASP_REQUIRE(n_cdf > 0 && m > 0 && m_total > 0, "sizes must be positive");
    ASP_REQUIRE(u0 >= 0.0 && u0 < 1.0, "u0 must be in [0, 1)");
    ASP_REQUIRE(ws, "null workspace");
    ASP_REQUIRE(j0 >= 0 && j0 + m <= m_total, "slot window outside the sweep");
    cudaStream_t s = (cudaStream_t)stream;
    GuideDeferred* def = (GuideDeferred*)ws;  // at most m / 2^16 + 2 long ranges
    guide_reset_kernel<<<1, 1, 0, s>>>(def);
    search_systematic_stream_kernel<<<grid_for(n_cdf + 1, 1), kBlock, 0, s>>>(d_cdf, n_cdf, u0, j0, m, m_total, d_idx,
                                                                             def);
    build_guide_deferred_kernel<int64_t><<<kNumSMs * 4, kBlock, 0, s>>>(d_idx, def);
    return check_launch("asp_search_systematic");
}

int asp_gather_rows(const double* d_x, int32_t dims, const int64_t* d_idx, int64_t m, double* d_x_out,
                    const double* d_l, const double* d_p, const double* d_q, double* d_l_out, double* d_p_out,
                    double* d_q_out, void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(d_x && d_idx && d_x_out, "null pointer");
    ASP_REQUIRE(dims > 0 && m > 0, "sizes must be positive");
    ASP_REQUIRE((!d_l || d_l_out) && (!d_p || d_p_out) && (!d_q || d_q_out), "log output missing");
    cudaStream_t s = (cudaStream_t)stream;
    // lanes per row: each lane moves 16 B per step
    int g = 1;
    while (g < 32 && g * 2 < dims) g <<= 1;
    const int grid = grid_for(m * g, 1);
#define ASP_GATHER(G) \
    gather_rows_kernel<G><<<grid, kBlock, 0, s>>>(d_x, dims, d_idx, m, d_x_out, d_l, d_p, d_q, d_l_out, d_p_out, d_q_out)
    switch (g) {
        case 1: ASP_GATHER(1); break;
        case 2: ASP_GATHER(2); break;
        case 4: ASP_GATHER(4); break;
        case 8: ASP_GATHER(8); break;
        case 16: ASP_GATHER(16); break;
        default: ASP_GATHER(32); break;
    }
#undef ASP_GATHER
    return check_launch("asp_gather_rows");
}

}  // extern "C"

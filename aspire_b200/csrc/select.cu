// K12 (multi-GPU glue, no reference site): owner-computes multinomial resampling.
// Rank r owns the CDF slice (lo, hi].  Every rank walks the SAME global stream of
// uniforms u_j, j in [0, n_total) - regenerated from Philox on the fly, or read from
// a supplied array in parity mode - and keeps, in slot order, the pairs (j, u_j)
// with lo <= u_j < hi: exactly the output slots whose ancestor lives on this rank
// (searchsorted-right semantics).  The ancestor's row therefore never crosses
// NVLink; only scalars do.  Two passes (count, ordered write) keep the order
// deterministic.
#include "common.cuh"
#include "philox.cuh"

namespace asp {

constexpr int kSelBlocks = kNumSMs * 4;

struct SelectWs {
    long long counts[kSelBlocks];
    long long offsets[kSelBlocks];
    long long total;
};

__device__ __forceinline__ void uniform_pair(bool use_philox, const PhiloxKeys& K, uint32_t stream,
                                             const double* __restrict__ u_all, int64_t pair, int64_t n_total,
                                             double& a, double& b) {
    if (use_philox) {
        const uint4 r = philox4x32_10((uint32_t)pair, (uint32_t)((uint64_t)pair >> 32), stream, 0u, K);
        a = u53(r.x, r.y);
        b = u53(r.z, r.w);
    } else {
        const int64_t j = 2 * pair;
        a = u_all[j];
        b = j + 1 < n_total ? u_all[j + 1] : 2.0;
    }
}

// blocks own contiguous ranges of pairs
__device__ __forceinline__ void block_range(int64_t n_pairs, int64_t& p0, int64_t& p1) {
    const int64_t chunk = (n_pairs + gridDim.x - 1) / gridDim.x;
    p0 = (int64_t)blockIdx.x * chunk;
    p1 = p0 + chunk < n_pairs ? p0 + chunk : n_pairs;
}

__global__ void __launch_bounds__(kBlock) select_count_kernel(bool use_philox, PhiloxKeys K, uint32_t stream,
                                                              const double* __restrict__ u_all, int64_t n_total,
                                                              double lo, double hi, SelectWs* ws) {
    __shared__ double smem[32];
    const int64_t n_pairs = (n_total + 1) / 2;
    int64_t p0, p1;
    block_range(n_pairs, p0, p1);
    double cnt[1] = {0.0};
    for (int64_t p = p0 + threadIdx.x; p < p1; p += blockDim.x) {
        double a, b;
        uniform_pair(use_philox, K, stream, u_all, p, n_total, a, b);
        cnt[0] += (a >= lo && a < hi) ? 1.0 : 0.0;
        cnt[0] += (2 * p + 1 < n_total && b >= lo && b < hi) ? 1.0 : 0.0;
    }
    const int ops[1] = {OP_SUM};
    block_reduce<1>(cnt, ops, smem);
    if (threadIdx.x == 0) ws->counts[blockIdx.x] = (long long)cnt[0];
}

__global__ void select_scan_kernel(SelectWs* ws, int nblocks, long long* d_count) {
    long long run = 0;
    for (int b = 0; b < nblocks; ++b) {
        ws->offsets[b] = run;
        run += ws->counts[b];
    }
    ws->total = run;
    *d_count = run;
}

__global__ void __launch_bounds__(kBlock) select_write_kernel(bool use_philox, PhiloxKeys K, uint32_t stream,
                                                              const double* __restrict__ u_all, int64_t n_total,
                                                              double lo, double hi, const SelectWs* ws,
                                                              int64_t capacity, int64_t* __restrict__ slot,
                                                              double* __restrict__ u_out) {
    __shared__ int s_warp[kBlock / 32];
    __shared__ long long s_base;
    const int64_t n_pairs = (n_total + 1) / 2;
    int64_t p0, p1;
    block_range(n_pairs, p0, p1);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_base = ws->offsets[blockIdx.x];
    __syncthreads();
    for (int64_t pb = p0; pb < p1; pb += blockDim.x) {
        const int64_t p = pb + threadIdx.x;
        double a = 2.0, b = 2.0;
        if (p < p1) uniform_pair(use_philox, K, stream, u_all, p, n_total, a, b);
        const bool ka = p < p1 && a >= lo && a < hi;
        const bool kb = p < p1 && 2 * p + 1 < n_total && b >= lo && b < hi;
        const int mine = (ka ? 1 : 0) + (kb ? 1 : 0);
        int incl = mine;  // warp inclusive scan
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_warp[wid] = incl;
        __syncthreads();
        int wbase = 0, btotal = 0;
#pragma unroll
        for (int w = 0; w < kBlock / 32; ++w) {
            const int c = s_warp[w];
            if (w < wid) wbase += c;
            btotal += c;
        }
        long long pos = s_base + wbase + incl - mine;
        if (ka) {
            if (pos < capacity) { slot[pos] = 2 * p; u_out[pos] = a; }
            ++pos;
        }
        if (kb && pos < capacity) { slot[pos] = 2 * p + 1; u_out[pos] = b; }
        __syncthreads();
        if (threadIdx.x == 0) s_base += btotal;
        __syncthreads();
    }
}

}  // namespace asp

using namespace asp;

extern "C" {

size_t asp_select_workspace_bytes(void) { return sizeof(SelectWs); }

int asp_select_uniforms(uint64_t seed, uint32_t stream_id, const double* d_u_all, int64_t n_total, double lo, double hi,
                        int64_t capacity, int64_t* d_slot, double* d_u, long long* d_count, void* ws, void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(d_slot && d_u && d_count && ws, "null pointer");
    ASP_REQUIRE(n_total > 0 && capacity >= 0, "bad sizes");
    cudaStream_t s = (cudaStream_t)stream;
    const bool use_philox = d_u_all == nullptr;
    const PhiloxKeys K = philox_keys(seed);
    const int64_t n_pairs = (n_total + 1) / 2;
    int nblocks = (int)((n_pairs + kBlock - 1) / kBlock);
    if (nblocks > kSelBlocks) nblocks = kSelBlocks;
    if (nblocks < 1) nblocks = 1;
    select_count_kernel<<<nblocks, kBlock, 0, s>>>(use_philox, K, stream_id, d_u_all, n_total, lo, hi, (SelectWs*)ws);
    select_scan_kernel<<<1, 1, 0, s>>>((SelectWs*)ws, nblocks, d_count);
    select_write_kernel<<<nblocks, kBlock, 0, s>>>(use_philox, K, stream_id, d_u_all, n_total, lo, hi,
                                                   (const SelectWs*)ws, capacity, d_slot, d_u);
    return check_launch("asp_select_uniforms");
}

}  // extern "C"

// Shared device/host helpers for the aspire_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/aspire_b200.h"

namespace asp {

// ---- error plumbing --------------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);
int require_device();

#define ASP_REQUIRE(cond, msg)                  \
    do {                                        \
        if (!(cond)) {                          \
            asp::set_error("%s: %s", __func__, msg); \
            return ASP_ERR_ARG;                 \
        }                                       \
    } while (0)

#define ASP_DEVICE_OR_RETURN()                  \
    do {                                        \
        int _rc = asp::require_device();        \
        if (_rc != ASP_OK) return _rc;          \
    } while (0)

// B200: 2 dies x 74 SMs.  A sizing constant only: grids are capped at a multiple of it and the per-block workspace
// slots are dimensioned by that cap, no kernel relies on all its blocks being resident at once (the cross-block steps
// are last-block tickets or single-block chains), so a part with another SM count runs correctly, merely with grids
// sized for 148 SMs.
constexpr int kNumSMs = 148;
constexpr int kBlock = 256;
constexpr int kMaxPartialBlocks = kNumSMs * 8;  // persistent grids never exceed this

// Workspace header shared by every reducing kernel: a ticket counter plus
// per-block partial slots (up to 16 doubles per block).
struct ReduceWs {
    unsigned int ticket;
    unsigned int pad[63];
    double partial[kMaxPartialBlocks * 16];
};

inline int grid_for(int64_t n, int items_per_thread, int block = kBlock, int blocks_per_sm = 8) {
    int64_t need = (n + (int64_t)block * items_per_thread - 1) / ((int64_t)block * items_per_thread);
    int64_t cap = (int64_t)kNumSMs * blocks_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

// ---- device helpers --------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

enum ReduceOp { OP_SUM = 0, OP_MIN = 1, OP_MAX = 2 };

__device__ __forceinline__ double warp_reduce(double v, int op) {
    if (op == OP_SUM) return warp_sum(v);
    if (op == OP_MIN) return warp_min(v);
    return warp_max(v);
}
__device__ __forceinline__ double combine(double a, double b, int op) {
    if (op == OP_SUM) return a + b;
    if (op == OP_MIN) return fmin(a, b);
    return fmax(a, b);
}
__device__ __forceinline__ double identity(int op) {
    if (op == OP_SUM) return 0.0;
    if (op == OP_MIN) return __longlong_as_double(0x7ff0000000000000LL);
    return __longlong_as_double(0xfff0000000000000LL);
}

// Block-wide reduction of NV values (ops given per value); result valid in
// thread 0.  Deterministic for a fixed block size.
template <int NV>
__device__ __forceinline__ void block_reduce(double (&v)[NV], const int (&ops)[NV], double* smem /*[NV*32]*/) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = warp_reduce(v[k], ops[k]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) smem[k * 32 + wid] = v[k];
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double t = lane < nw ? smem[k * 32 + lane] : identity(ops[k]);
            v[k] = warp_reduce(t, ops[k]);
        }
    }
}

// Grid-wide deterministic reduction: every block stores its NV partials, the
// last block to arrive (ticket) combines them in block order and calls
// `fin(values)` from thread 0.  The ticket is reset for the next launch.
template <int NV, typename Fin>
__device__ __forceinline__ void grid_reduce(double (&v)[NV], const int (&ops)[NV], ReduceWs* ws, double* smem, Fin fin) {
    __shared__ bool is_last;
    block_reduce<NV>(v, ops, smem);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) ws->partial[(size_t)blockIdx.x * NV + k] = v[k];
        __threadfence();
        unsigned int t = atomicAdd(&ws->ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double acc[NV];
#pragma unroll
        for (int k = 0; k < NV; ++k) acc[k] = identity(ops[k]);
        for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
#pragma unroll
            for (int k = 0; k < NV; ++k)
                acc[k] = combine(acc[k], __ldcg(&ws->partial[(size_t)b * NV + k]), ops[k]);
        }
        block_reduce<NV>(acc, ops, smem);
        if (threadIdx.x == 0) {
            ws->ticket = 0;
            fin(acc);
        }
    }
}

// streaming (read-once) 128-bit load of two doubles
__device__ __forceinline__ double2 ld_stream2(const double* p) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ double ld_stream(const double* p) {
    double r;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream2(double* p, double a, double b) {
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b));
}

#endif  // __CUDACC__

}  // namespace asp

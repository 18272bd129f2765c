// tcgen05 / TMEM / TMA building blocks of the coupling flow (flow.cu: log_prob / sample; flow_move.cu: the pCN
// sweep with the flow as proposal density): raw PTX wrappers, the shared-memory operand layouts and one coupling
// layer for the 128 particles of a warpgroup.  See flow.cu for the design notes.
#pragma once
#include <cuda_bf16.h>
#include <math.h>

#include "common.cuh"

namespace asp {

constexpr int kFlowH = 64;        // hidden width
constexpr int kFlowM = 128;       // particles per tile (UMMA M)
constexpr int kFlowThreads = 128;
constexpr int kFlowMaxLayers = 8;
constexpr float kFlowLimit = 6.907755278982137f;  // log(1e3): zuko's soft clip
constexpr float kLog2Pi = 1.8378770664093453f;

// ---- raw PTX wrappers -------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded wait: returns false if the barrier never flipped (reported, never hangs)
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
#pragma unroll 1
    for (int it = 0; it < (1 << 24); ++it) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (ok) return true;
    }
    return false;
}
__device__ __forceinline__ void tma_bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]   (one K = 16 step)
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// 32 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
        "[%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor: no swizzle, K-major; LBO = stride between core
// matrices along K, SBO = stride between 8-row groups; version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// instruction descriptor: D = f32, A = B = bf16, both K-major, M = 128, N
__host__ __device__ constexpr uint32_t umma_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kFlowM >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// relu + round to bf16 + pack: one instruction per two activations
__device__ __forceinline__ uint32_t pack_relu_bf16(float a, float b) {
    uint32_t r;
    asm("cvt.rn.relu.bf16x2.f32 %0, %2, %1;" : "=r"(r) : "f"(a), "f"(b));  // d = {hi: %2... low half = second source
    return r;
}
__device__ __forceinline__ float fast_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
// 32 bytes of a row in one load: a whole sector per instruction, no reliance on L1 between instructions
__device__ __forceinline__ void ld_row32(const double* p, double (&v)[4]) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0, %1, %2, %3}, [%4];"
                 : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}

__device__ __forceinline__ void st_row32(double* p, double a, double b, double c, double d) {
    asm volatile("st.global.L1::no_allocate.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

// write 8 consecutive K elements (one core-matrix row) of this thread's A-tile row
__device__ __forceinline__ void a_tile_store8(unsigned char* a_tile, int row, int kchunk, const float* v) {
    uint4 q = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    *reinterpret_cast<uint4*>(a_tile + kchunk * (kFlowM * 16) + (row >> 3) * 128 + (row & 7) * 16) = q;
}

__device__ __forceinline__ void a_tile_store8_relu(unsigned char* a_tile, int row, int kchunk, const float* v) {
    uint4 q = make_uint4(pack_relu_bf16(v[0], v[1]), pack_relu_bf16(v[2], v[3]), pack_relu_bf16(v[4], v[5]),
                         pack_relu_bf16(v[6], v[7]));
    *reinterpret_cast<uint4*>(a_tile + kchunk * (kFlowM * 16) + (row >> 3) * 128 + (row & 7) * 16) = q;
}

template <int D>
struct FlowSizes {
    static constexpr int kHalf = D / 2;
    static constexpr int kKA = kHalf < 16 ? 16 : kHalf;       // K of the first GEMM (padded to 16)
    static constexpr int kW1 = kFlowH * kKA * 2;              // bytes
    static constexpr int kW2 = kFlowH * kFlowH * 2;
    static constexpr int kW3 = D * kFlowH * 2;
    static constexpr int kLayer = kW1 + kW2 + kW3;
    static constexpr int kTmemCols = 64;                      // >= max(64, D) accumulators, power of two
    // bias operands of the extra K = 16 step, built in the prologue: [N x 16] bf16 K-major core matrices
    static constexpr int kB12 = kFlowH * 16 * 2;              // bias blocks of GEMM 1 / 2 (N = 64)
    static constexpr int kB3 = D * 16 * 2;                    // GEMM 3 (N = D)
    static constexpr int kBiasLayer = 2 * kB12 + kB3;
    static constexpr int kOnes = kFlowM * 16 * 2;             // the constant A operand: columns 0, 1 = 1
};

// issue one GEMM [128 x K] . [K x N] -> TMEM (thread 0 only), then commit to `bar`
template <int K, int N>
__device__ __forceinline__ void issue_gemm(uint32_t tmem, const unsigned char* a_tile, const unsigned char* w,
                                           const unsigned char* ones, const unsigned char* bias_blk, uint64_t* bar) {
    const uint32_t a0 = smem_u32(a_tile), b0 = smem_u32(w);
    constexpr uint32_t idesc = umma_idesc(N);
    // D = ones . bias  (the accumulator starts at the bias), then D += A . W
    umma_f16(tmem, umma_desc(smem_u32(ones), kFlowM * 16, 128), umma_desc(smem_u32(bias_blk), N * 16, 128), idesc, 0u);
#pragma unroll
    for (int s = 0; s < K / 16; ++s) {
        const uint64_t ad = umma_desc(a0 + s * 2 * (kFlowM * 16), kFlowM * 16, 128);
        const uint64_t bd = umma_desc(b0 + s * 2 * (N * 16), N * 16, 128);
        umma_f16(tmem, ad, bd, idesc, 1u);
    }
    umma_commit(bar);
}

// barrier over the 128 threads of one warpgroup (named barriers 1..; 0 is __syncthreads)
__device__ __forceinline__ void wg_sync(int wg) { asm volatile("bar.sync %0, 128;" ::"r"(wg + 1) : "memory"); }

// One coupling layer for this thread's particle; ODD selects which half conditions which
// (compile time, so z[] keeps static indices and stays in registers).
template <int D, bool INVERSE, bool ODD>
__device__ __forceinline__ bool coupling_layer(float (&z)[D], float& ladj, int layer, const unsigned char* s_w,
                                               unsigned char* s_a, const unsigned char* s_ones,
                                               const unsigned char* s_bb, uint64_t* mma_bar, uint32_t tmem,
                                               uint32_t lane_taddr, uint32_t& phase, int tid, int wg) {
    using S = FlowSizes<D>;
    constexpr int HALF = S::kHalf;
    constexpr bool odd = ODD;
    bool ok = true;
    const unsigned char* w1 = s_w + layer * S::kLayer;
    const unsigned char* w2 = w1 + S::kW1;
    const unsigned char* w3 = w2 + S::kW2;
    const unsigned char* bb1 = s_bb + layer * S::kBiasLayer;
    const unsigned char* bb2 = bb1 + S::kB12;
    const unsigned char* bb3 = bb2 + S::kB12;
    // ---- GEMM 1: A = bf16(z_a) ----
#pragma unroll
    for (int c = 0; c < S::kKA / 8; ++c) {
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            constexpr int last = HALF - 1;
            const int j = c * 8 + k < HALF ? c * 8 + k : last;   // (K padding beyond HALF is zero)
            v[k] = c * 8 + k < HALF ? (odd ? z[HALF + j] : z[j]) : 0.f;
        }
        a_tile_store8(s_a, tid, c, v);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    wg_sync(wg);
    if (tid == 0) {
        tc_fence_after();
        issue_gemm<S::kKA, kFlowH>(tmem, s_a, w1, s_ones, bb1, mma_bar);
    }
    ok &= mbar_wait(mma_bar, phase);
    phase ^= 1;
    tc_fence_after();
    // ---- epilogue 1 -> A = bf16(relu(acc)) (the bias is already in the accumulator); GEMM 2 ----
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        float acc[32];
        tmem_ld32(lane_taddr + half * 32, acc);
#pragma unroll
        for (int c = 0; c < 4; ++c) a_tile_store8_relu(s_a, tid, half * 4 + c, acc + c * 8);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    wg_sync(wg);
    if (tid == 0) {
        tc_fence_after();
        issue_gemm<kFlowH, kFlowH>(tmem, s_a, w2, s_ones, bb2, mma_bar);
    }
    ok &= mbar_wait(mma_bar, phase);
    phase ^= 1;
    tc_fence_after();
    // ---- epilogue 2 -> A = bf16(relu(acc)); GEMM 3 (N = D) ----
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        float acc[32];
        tmem_ld32(lane_taddr + half * 32, acc);
#pragma unroll
        for (int c = 0; c < 4; ++c) a_tile_store8_relu(s_a, tid, half * 4 + c, acc + c * 8);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    wg_sync(wg);
    if (tid == 0) {
        tc_fence_after();
        issue_gemm<kFlowH, D>(tmem, s_a, w3, s_ones, bb3, mma_bar);
    }
    ok &= mbar_wait(mma_bar, phase);
    phase ^= 1;
    tc_fence_after();
    // ---- epilogue 3: columns [0, HALF) = shift, [HALF, D) = raw log-scale ----
    auto update = [&](int j, float shift, float r) {
        const float log_s = r * fast_rcp(fmaf(fabsf(r), 1.0f / kFlowLimit, 1.0f));  // soft clip: |log_s| < log(1e3)
        ladj += log_s;
        // z_b is the half NOT fed to the conditioner
        float& zb = odd ? z[j] : z[HALF + j];
        const float t = log_s * 1.4426950408889634f;
        zb = INVERSE ? (zb - shift) * fast_ex2(-t) : fmaf(zb, fast_ex2(t), shift);
    };
    if constexpr (HALF >= 16) {
#pragma unroll
        for (int c = 0; c < HALF / 16; ++c) {
            float sh[16], raw[16];
            tmem_ld16(lane_taddr + c * 16, sh);
            tmem_ld16(lane_taddr + HALF + c * 16, raw);
#pragma unroll
            for (int k = 0; k < 16; ++k) update(c * 16 + k, sh[k], raw[k]);
        }
    } else {  // D = 16: both halves sit in one 16-column load
        static_assert(HALF == 8, "flow kernel: D must be 16, 32 or 64");
        float v[16];
        tmem_ld16(lane_taddr, v);
#pragma unroll
        for (int k = 0; k < 8; ++k) update(k, v[k], v[8 + k]);
    }
    return ok;
}

}  // namespace asp

// K9/K10: affine-coupling flow log_prob / sample on the 5th-gen tensor cores.
//
// Replaces the dense contraction of ZukoFlow.log_prob / sample_and_log_prob
// (/root/reference/src/aspire/flows/torch/flows.py:327-387; data transform
// src/aspire/transforms.py:618-628).  Spec: oracle/flow.py.
//
// One CTA = 4 warps = one tile of 128 particles; thread i owns particle i: its state
// z[D] lives in fp32 registers, its TMEM lane i receives the accumulators.  Per coupling
// layer three GEMMs  [128 x K] . [K x N]  run as tcgen05.mma.kind::f16 (bf16 x bf16 ->
// fp32 in TMEM), issued by one thread:
//   A = activations, written by their owner threads as bf16 into shared memory in the
//       no-swizzle K-major core-matrix layout (8 rows x 16 B per core matrix);
//   B = the layer's weights, resident in shared memory for the whole kernel, brought in
//       once by a TMA bulk copy (cp.async.bulk -> UBLKCP) of the host-packed blob;
//   D = 64 (or D) fp32 TMEM columns, read back with tcgen05.ld.32x32b for the fp32
//       epilogue (ReLU + bf16 pack in one cvt.rn.relu.bf16x2, or soft-clip / exp / affine update).
// The bias rides on the tensor pipe: every GEMM starts with one extra K = 16 step  ones[128 x 16] . bias[16 x N]
// (bias = bf16 hi + bf16 lo, relative error 2^-17), so the epilogues add nothing.  The chain MMA -> wait -> epilogue
// is serial per tile, so several warpgroups per SM keep independent tiles in flight over one copy of the weights.
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

struct FlowArgs {
    const double* x;        // log_prob: particles [n, D];  sample: base normals [n, D]
    double* x_out;          // sample: particles [n, D]  (log_prob: optional z [n, D], may be null)
    double* logq;           // [n]
    int64_t n;
    const uint16_t* weights;  // packed bf16 blob (all layers)
    const float* bias;        // [L][3][64]  (b3 padded to 64)
    const double* mean;       // [D]
    const double* inv_std;    // [D]   (sample: std)
    double log_det0;          // -sum log std
    int n_layers;
    unsigned int* error;
};

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

// The tile chain MMA -> wait -> epilogue is serial for the 128 threads that own a tile, so throughput comes from
// tiles in flight per SM.  D = 64: the weights blob (61 KB) allows two CTAs per SM, so one CTA runs kFlowWG = 3
// independent warpgroups (own A tile, own 64 TMEM columns, own MMA barrier) over one copy of the weights;
// D = 32 / 16: three / four single-warpgroup CTAs fit.
template <int D>
constexpr int flow_wg() { return D >= 64 ? 4 : 6; }
template <int D>
constexpr int flow_ctas_per_sm() { return 1; }

template <int D, bool INVERSE>
__global__ void __launch_bounds__(kFlowThreads * flow_wg<D>(), flow_ctas_per_sm<D>()) flow_kernel(const FlowArgs A) {
    using S = FlowSizes<D>;
    constexpr int HALF = S::kHalf;
    constexpr int NWG = flow_wg<D>();
    constexpr int kTmemAll = NWG == 1 ? 64 : (NWG == 2 ? 128 : (NWG <= 4 ? 256 : 512));  // power of two >= 64 NWG
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* s_w = smem;                                             // weights blob
    unsigned char* s_bb = smem + A.n_layers * S::kLayer;                   // bias blocks (bf16 hi / lo rows)
    unsigned char* s_ones = s_bb + A.n_layers * S::kBiasLayer;             // constant A operand of the bias step
    unsigned char* s_a_all = s_ones + S::kOnes;                            // NWG A tiles (16 KB each)
    double* s_mean = reinterpret_cast<double*>(s_a_all + NWG * kFlowM * kFlowH * 2);
    float* s_scale = reinterpret_cast<float*>(s_mean + D);                 // log_prob: 1 / std;  sample: std
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_scale + D);           // [0]: weights TMA, [1 + wg]: MMA
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 1 + NWG);

    const int wg = threadIdx.x / kFlowThreads, tid = threadIdx.x % kFlowThreads, warp = tid >> 5;
    unsigned char* s_a = s_a_all + wg * (kFlowM * kFlowH * 2);
    const int wbytes = A.n_layers * S::kLayer;
    if (threadIdx.x == 0) {
        mbar_init(&s_bar[0], 1);
        for (int w = 0; w < NWG; ++w) mbar_init(&s_bar[1 + w], 1);
        fence_mbar_init();
    }
    if (threadIdx.x < 32) tmem_alloc(s_tmem, kTmemAll);
    // bias blocks: element (n, k) of an [N x 16] K-major block sits at (k / 8) N 16 + (n / 8) 128 + (n % 8) 16 + (k % 8) 2;
    // k = 0 holds bf16(b_n), k = 1 holds bf16(b_n - bf16(b_n)), the other K rows are zero
    for (int i = threadIdx.x; i < A.n_layers * S::kBiasLayer / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(s_bb)[i] = 0u;
    for (int i = threadIdx.x; i < S::kOnes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(s_ones)[i] = 0u;
    __syncthreads();
    for (int i = threadIdx.x; i < A.n_layers * 3 * 64; i += blockDim.x) {
        const int l = i / 192, g = (i % 192) / 64, n = i % 64;
        const int N = g < 2 ? kFlowH : D;
        if (n < N) {
            const float b = A.bias[i];
            const __nv_bfloat16 hi = __float2bfloat16_rn(b);
            const __nv_bfloat16 lo = __float2bfloat16_rn(b - __bfloat162float(hi));
            unsigned char* blk = s_bb + l * S::kBiasLayer + (g == 0 ? 0 : (g == 1 ? S::kB12 : 2 * S::kB12));
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(blk + (n >> 3) * 128 + (n & 7) * 16);
            dst[0] = hi;
            dst[1] = lo;
        }
    }
    for (int r = threadIdx.x; r < kFlowM; r += blockDim.x) {
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(s_ones + (r >> 3) * 128 + (r & 7) * 16);
        dst[0] = __float2bfloat16_rn(1.0f);
        dst[1] = __float2bfloat16_rn(1.0f);
    }
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
        s_mean[i] = A.mean[i];
        s_scale[i] = (float)A.inv_std[i];
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem + (uint32_t)(wg * S::kTmemCols);  // this warpgroup's accumulator columns
    if (threadIdx.x == 0) {  // weights: one TMA bulk copy per layer blob, all on one mbarrier
        mbar_expect_tx(&s_bar[0], (uint32_t)wbytes);
        for (int l = 0; l < A.n_layers; ++l)
            tma_bulk_load(s_w + l * S::kLayer, reinterpret_cast<const unsigned char*>(A.weights) + (size_t)l * S::kLayer,
                          S::kLayer, &s_bar[0]);
    }
    bool ok = mbar_wait(&s_bar[0], 0);
    uint32_t phase = 0;
    uint64_t* mma_bar = &s_bar[1 + wg];
    const uint32_t lane_taddr = tmem + ((uint32_t)(warp * 32) << 16);  // warp % 4 selects the 32-lane TMEM slab

    const int64_t n_tiles = (A.n + kFlowM - 1) / kFlowM;
    for (int64_t tile = (int64_t)blockIdx.x * NWG + wg; tile < n_tiles; tile += (int64_t)gridDim.x * NWG) {
        const int64_t row = tile * kFlowM + tid;
        const bool live = row < A.n;
        float z[D];
        {
            // a row is 8 D bytes, read 32 bytes (one sector) per instruction; the centring stays in fp64 (a parameter
            // can sit far from zero compared with its spread), the scaling is fp32
            const double* src = A.x + (live ? row : A.n - 1) * D;
#pragma unroll
            for (int j = 0; j < D; j += 4) {
                double v[4];
                ld_row32(src + j, v);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    z[j + k] = INVERSE ? (float)v[k] : (float)(v[k] - s_mean[j + k]) * s_scale[j + k];
            }
        }
        float base = 0.f;
        if (INVERSE) {
#pragma unroll
            for (int j = 0; j < D; ++j) base = fmaf(z[j], z[j], base);
        }
        float ladj = 0.f;
        for (int li = 0; li < A.n_layers; ++li) {
            const int layer = INVERSE ? A.n_layers - 1 - li : li;
            if (layer & 1)
                ok &= coupling_layer<D, INVERSE, true>(z, ladj, layer, s_w, s_a, s_ones, s_bb, mma_bar, tmem, lane_taddr,
                                                       phase, tid, wg);
            else
                ok &= coupling_layer<D, INVERSE, false>(z, ladj, layer, s_w, s_a, s_ones, s_bb, mma_bar, tmem, lane_taddr,
                                                        phase, tid, wg);
        }
        // ---- outputs ----
        if (live) {
            if (INVERSE) {
                double* dst = A.x_out + row * D;
#pragma unroll
                for (int j = 0; j < D; j += 4) {  // 32 bytes (one sector) per store
                    double o[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) o[k] = fma((double)z[j + k], (double)s_scale[j + k], s_mean[j + k]);
                    st_row32(dst + j, o[0], o[1], o[2], o[3]);
                }
                A.logq[row] = (double)(-0.5f * base - 0.5f * D * kLog2Pi + ladj) + A.log_det0;
            } else {
                float q = 0.f;
#pragma unroll
                for (int j = 0; j < D; ++j) q = fmaf(z[j], z[j], q);
                A.logq[row] = (double)(-0.5f * q - 0.5f * D * kLog2Pi + ladj) + A.log_det0;
                if (A.x_out) {
                    double* dst = A.x_out + row * D;
#pragma unroll
                    for (int j = 0; j < D; j += 4) st_row32(dst + j, z[j], z[j + 1], z[j + 2], z[j + 3]);
                }
            }
        }
    }
    if (!ok && tid == 0) atomicOr(A.error, 1u);
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(*s_tmem, kTmemAll);
}

template <int D>
static size_t flow_smem_bytes(int n_layers) {
    using S = FlowSizes<D>;
    return (size_t)n_layers * (S::kLayer + S::kBiasLayer) + S::kOnes + (size_t)flow_wg<D>() * kFlowM * kFlowH * 2 +
           D * 8 + D * 4 + (1 + flow_wg<D>()) * 8 + 16 + 128;
}

template <int D, bool INV>
static int launch_flow(const FlowArgs& A, cudaStream_t s) {
    const size_t smem = flow_smem_bytes<D>(A.n_layers);
    cudaError_t e = cudaFuncSetAttribute(flow_kernel<D, INV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error("flow kernel: %s", cudaGetErrorString(e));
        return ASP_ERR_CUDA;
    }
    const int64_t n_tiles = (A.n + kFlowM - 1) / kFlowM;
    constexpr int NWG = flow_wg<D>();
    constexpr int64_t cap = (int64_t)flow_ctas_per_sm<D>() * kNumSMs;
    const int64_t need = (n_tiles + NWG - 1) / NWG;
    const int grid = (int)(need < cap ? need : cap);
    flow_kernel<D, INV><<<grid, kFlowThreads * NWG, smem, s>>>(A);
    return check_launch("asp_flow");
}

}  // namespace asp

using namespace asp;

extern "C" {

size_t asp_flow_weights_bytes(int32_t dims, int32_t n_layers) {
    switch (dims) {
        case 16: return (size_t)n_layers * FlowSizes<16>::kLayer;
        case 32: return (size_t)n_layers * FlowSizes<32>::kLayer;
        case 64: return (size_t)n_layers * FlowSizes<64>::kLayer;
        default: return 0;
    }
}

static int flow_dispatch(const asp_flow_t* f, const double* d_in, int64_t n, double* d_x_out, double* d_logq,
                         bool inverse, void* stream) {
    ASP_DEVICE_OR_RETURN();
    ASP_REQUIRE(f && d_in && d_logq, "null pointer");
    ASP_REQUIRE(f->d_weights && f->d_bias && f->d_mean && f->d_inv_std && f->d_std && f->d_error,
                "flow descriptor incomplete");
    ASP_REQUIRE(n > 0, "n must be positive");
    ASP_REQUIRE(f->hidden == kFlowH, "hidden width must be 64");
    ASP_REQUIRE(f->n_layers >= 1 && f->n_layers <= kFlowMaxLayers, "n_layers must be in [1, 8]");
    ASP_REQUIRE(!inverse || d_x_out, "sample needs an output buffer");
    FlowArgs A;
    A.x = d_in; A.x_out = d_x_out; A.logq = d_logq; A.n = n;
    A.weights = (const uint16_t*)f->d_weights; A.bias = f->d_bias; A.mean = f->d_mean; A.inv_std = inverse ? f->d_std : f->d_inv_std;
    A.log_det0 = f->log_det0; A.n_layers = f->n_layers; A.error = f->d_error;
    cudaStream_t s = (cudaStream_t)stream;
    switch (f->dims) {
        case 16: return inverse ? launch_flow<16, true>(A, s) : launch_flow<16, false>(A, s);
        case 32: return inverse ? launch_flow<32, true>(A, s) : launch_flow<32, false>(A, s);
        case 64: return inverse ? launch_flow<64, true>(A, s) : launch_flow<64, false>(A, s);
        default: set_error("asp_flow: dims must be 16, 32 or 64"); return ASP_ERR_ARG;
    }
}

int asp_flow_logprob_bf16(const asp_flow_t* flow, const double* d_x, int64_t n, double* d_logq, double* d_z, void* stream) {
    return flow_dispatch(flow, d_x, n, d_z, d_logq, false, stream);
}

int asp_flow_sample_bf16(const asp_flow_t* flow, const double* d_u, int64_t n, double* d_x, double* d_logq, void* stream) {
    return flow_dispatch(flow, d_u, n, d_x, d_logq, true, stream);
}

}  // extern "C"

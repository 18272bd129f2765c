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
#include "flow_core.cuh"

namespace asp {

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

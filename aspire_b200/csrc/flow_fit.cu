// 8f-1: maximum-likelihood fit of the coupling flow on the device - forward, backward and Adam kernels over the
// flow's fp32 master weights (the bf16 blob asp_flow_*_bf16 reads is re-packed from them after the fit).
//
// Replaces the optimisation loop of ZukoFlow.fit (/root/reference/src/aspire/flows/torch/flows.py:170-325: Adam,
// mini-batches of 500, loss = mean negative log-likelihood); the protocol around it (data transform, permutation,
// train / validation split, best-validation restore, early stopping, annealing) stays in aspire_b200/flows.py.
// Setup work, not the hot path: a mini-batch is a few hundred rows through nine 64-wide layers, so the
// kernels are plain fp32 CUDA - one thread per output element - and the step is bound by its ~40 launches.
// Model (oracle/flow.py, flows.py::_torch_forward):  z = (x - mean) / std;  per layer t (alternating halves a | b):
//   h1 = relu(W1 z_a + b1), h2 = relu(W2 h1 + b2), o = W3 h2 + b3 = (shift | raw),
//   log_s = raw / (1 + |raw| / L),  z_b <- z_b exp(log_s) + shift;   -log q = 1/2 |z|^2 + K/2 log 2pi - sum log_s + sum log std
#include <math.h>

#include "common.cuh"

namespace asp {

constexpr float kFitLimit = 6.907755278982137f;  // log(1e3): the soft clip of the log-scale
constexpr int kFitH = 64;

struct FitLayer {
    size_t w1, b1, w2, b2, w3, b3, end;  // offsets (floats) inside one layer's parameter block
};
__host__ __device__ inline FitLayer fit_layout(int kdims) {
    const size_t half = kdims / 2;
    FitLayer L;
    L.w1 = 0;
    L.b1 = L.w1 + kFitH * half;
    L.w2 = L.b1 + kFitH;
    L.b2 = L.w2 + (size_t)kFitH * kFitH;
    L.w3 = L.b2 + kFitH;
    L.b3 = L.w3 + (size_t)kdims * kFitH;
    L.end = L.b3 + kdims;
    return L;
}

// C[m][n] = (bias[n] +) sum_k A(m, k) B(k, n), optionally relu or masked by (mask[m][n] > 0), optionally accumulated.
//   A(m, k) = TA ? A[k lda + m] : A[m lda + k];   B(k, n) = TB ? B[n ldb + k] : B[k ldb + n]
template <bool TA, bool TB>
__global__ void fit_gemm_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                                float* __restrict__ C, int ldc, int M, int N, int K, const float* __restrict__ bias,
                                int relu, const float* __restrict__ mask, int ldm, int accumulate) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * N) return;
    const int m = idx / N, n = idx % N;
    float acc = bias ? bias[n] : 0.f;
    for (int k = 0; k < K; ++k) {
        const float a = TA ? A[(size_t)k * lda + m] : A[(size_t)m * lda + k];
        const float b = TB ? B[(size_t)n * ldb + k] : B[(size_t)k * ldb + n];
        acc = fmaf(a, b, acc);
    }
    if (relu) acc = fmaxf(acc, 0.f);
    if (mask && !(mask[(size_t)m * ldm + n] > 0.f)) acc = 0.f;
    float* c = C + (size_t)m * ldc + n;
    *c = accumulate ? *c + acc : acc;
}

__global__ void fit_standardise_kernel(const float* __restrict__ x, int b, int K, const float* __restrict__ mean,
                                       const float* __restrict__ inv_std, float* __restrict__ z,
                                       float* __restrict__ ladj) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < b * K) z[idx] = (x[idx] - mean[idx % K]) * inv_std[idx % K];
    if (idx < b) ladj[idx] = 0.f;
}

// z_out = z_in with the b half transformed; ladj += sum log_s.  One thread per (sample, j < half).
__global__ void fit_coupling_fwd_kernel(const float* __restrict__ z_in, const float* __restrict__ o, int b, int K, int odd,
                                        float* __restrict__ z_out, float* __restrict__ ladj) {
    const int half = K / 2;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= b * half) return;
    const int s = idx / half, j = idx % half;
    const int ja = odd ? half + j : j, jb = odd ? j : half + j;
    const float shift = o[(size_t)s * K + j], raw = o[(size_t)s * K + half + j];
    const float log_s = raw / (1.f + fabsf(raw) * (1.f / kFitLimit));
    z_out[(size_t)s * K + ja] = z_in[(size_t)s * K + ja];
    z_out[(size_t)s * K + jb] = fmaf(z_in[(size_t)s * K + jb], expf(log_s), shift);
    atomicAdd(&ladj[s], log_s);
}

// loss[0] = mean_s (1/2 |z_s|^2 - ladj_s) + const;  g = z / b (the gradient of the mean with respect to the final z)
__global__ void fit_loss_kernel(const float* __restrict__ z, const float* __restrict__ ladj, int b, int K, float konst,
                                float* __restrict__ g, float* __restrict__ loss) {
    __shared__ float s_part[32];
    float acc = 0.f;
    for (int s = threadIdx.x; s < b; s += blockDim.x) {
        float q = 0.f;
        for (int j = 0; j < K; ++j) {
            const float v = z[(size_t)s * K + j];
            q = fmaf(v, v, q);
            if (g) g[(size_t)s * K + j] = v / (float)b;
        }
        acc += 0.5f * q - ladj[s];
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < (int)(blockDim.x + 31) / 32; ++w) t += s_part[w];
        loss[0] = t / (float)b + konst;
    }
}

// Through the coupling transform, backwards: given g = dL/dz_out, writes dO = (dshift | draw) (rows of padding
// coordinates zeroed) and turns g into dL/dz_in EXCEPT the conditioner's contribution to the a half (added by the
// following gemm).  One thread per (sample, j < half).
__global__ void fit_coupling_bwd_kernel(const float* __restrict__ z_in, const float* __restrict__ o, int b, int K, int odd,
                                        const unsigned char* __restrict__ out_mask, float* __restrict__ g,
                                        float* __restrict__ d_o) {
    const int half = K / 2;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= b * half) return;
    const int s = idx / half, j = idx % half;
    const int jb = odd ? j : half + j;
    const float raw = o[(size_t)s * K + half + j];
    const float den = 1.f + fabsf(raw) * (1.f / kFitLimit);
    const float log_s = raw / den, e = expf(log_s);
    const float gb = g[(size_t)s * K + jb];
    const float d_logs = gb * z_in[(size_t)s * K + jb] * e - 1.f / (float)b;  // -ladj enters the loss with 1 / b
    const bool real = !out_mask || out_mask[j];
    d_o[(size_t)s * K + j] = real ? gb : 0.f;
    d_o[(size_t)s * K + half + j] = real ? d_logs / (den * den) : 0.f;
    g[(size_t)s * K + jb] = gb * e;
}

__global__ void fit_colsum_kernel(const float* __restrict__ A, int lda, int M, int N, float* __restrict__ out) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float acc = 0.f;
    for (int m = 0; m < M; ++m) acc += A[(size_t)m * lda + n];
    out[n] = acc;
}

// g[:, a half] += d_za
__global__ void fit_add_half_kernel(float* __restrict__ g, const float* __restrict__ d_za, int b, int K, int odd) {
    const int half = K / 2;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= b * half) return;
    const int s = idx / half, j = idx % half;
    g[(size_t)s * K + (odd ? half + j : j)] += d_za[(size_t)s * half + j];
}

__global__ void fit_sumsq_kernel(const float* __restrict__ g, size_t n, float* __restrict__ out) {
    __shared__ float s_part[32];
    float acc = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        acc = fmaf(g[i], g[i], acc);
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < (int)(blockDim.x + 31) / 32; ++w) t += s_part[w];
        atomicAdd(out, t);
    }
}

// Adam (torch.optim.Adam defaults: no weight decay, no amsgrad) with the optional global-norm gradient clipping of
// torch.nn.utils.clip_grad_norm_ (scale = min(1, clip / (norm + 1e-6)))
__global__ void fit_adam_kernel(float* __restrict__ w, float* __restrict__ m, float* __restrict__ v,
                                const float* __restrict__ g, size_t n, float lr, float beta1, float beta2, float eps,
                                float bc1, float bc2, float clip, const float* __restrict__ sumsq) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float scale = 1.f;
    if (clip > 0.f) {
        const float norm = sqrtf(sumsq[0]);
        scale = fminf(1.f, clip / (norm + 1e-6f));
    }
    const float gi = g[i] * scale;
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    w[i] -= lr * (mi / bc1) / (sqrtf(vi / bc2) + eps);
}

struct FitWs {  // workspace layout (floats), per layer: z_in [b K] | h1 [b 64] | h2 [b 64] | o [b K]
    float *z_final, *ladj, *g, *d_o, *d_h, *d_h1, *d_za, *sumsq, *layers;
    size_t per_layer;
};
inline FitWs fit_ws(float* ws, int kdims, int n_layers, int cap) {
    FitWs W;
    const size_t bk = (size_t)cap * kdims, bh = (size_t)cap * kFitH;
    W.z_final = ws;
    W.ladj = W.z_final + bk;
    W.g = W.ladj + cap;
    W.d_o = W.g + bk;
    W.d_h = W.d_o + bk;
    W.d_h1 = W.d_h + bh;
    W.d_za = W.d_h1 + bh;
    W.sumsq = W.d_za + bk;
    W.layers = W.sumsq + 32;
    W.per_layer = 2 * bk + 2 * bh;
    (void)n_layers;
    return W;
}

static inline int blocks_for(size_t n) { return (int)((n + 255) / 256); }

template <bool TA, bool TB>
static void gemm(cudaStream_t s, const float* A, int lda, const float* B, int ldb, float* C, int ldc, int M, int N, int K,
                 const float* bias = nullptr, int relu = 0, const float* mask = nullptr, int ldm = 0, int accumulate = 0) {
    fit_gemm_kernel<TA, TB><<<blocks_for((size_t)M * N), 256, 0, s>>>(A, lda, B, ldb, C, ldc, M, N, K, bias, relu, mask,
                                                                     ldm, accumulate);
}

// forward pass over the stored layers; returns the pointer to the final z
static float* fit_forward(const asp_flow_train_t& T, const float* x, int b, const FitWs& W, cudaStream_t s) {
    const int K = T.kdims, half = K / 2;
    const FitLayer L = fit_layout(K);
    float* z = W.layers;  // layer 0's z_in
    fit_standardise_kernel<<<blocks_for((size_t)b * K), 256, 0, s>>>(x, b, K, T.d_mean, T.d_inv_std, z, W.ladj);
    for (int t = 0; t < T.n_layers; ++t) {
        const int odd = t & 1;
        float* base = W.layers + (size_t)t * W.per_layer;
        float* z_in = base;
        float* h1 = z_in + (size_t)T.batch_cap * K;
        float* h2 = h1 + (size_t)T.batch_cap * kFitH;
        float* o = h2 + (size_t)T.batch_cap * kFitH;
        float* z_out = t + 1 < T.n_layers ? base + W.per_layer : W.z_final;
        const float* w = T.d_w + (size_t)t * L.end;
        const float* za = z_in + (odd ? half : 0);
        gemm<false, true>(s, za, K, w + L.w1, half, h1, kFitH, b, kFitH, half, w + L.b1, 1);
        gemm<false, true>(s, h1, kFitH, w + L.w2, kFitH, h2, kFitH, b, kFitH, kFitH, w + L.b2, 1);
        gemm<false, true>(s, h2, kFitH, w + L.w3, kFitH, o, K, b, K, kFitH, w + L.b3, 0);
        fit_coupling_fwd_kernel<<<blocks_for((size_t)b * half), 256, 0, s>>>(z_in, o, b, K, odd, z_out, W.ladj);
    }
    return W.z_final;
}

}  // namespace asp

using namespace asp;

extern "C" {

size_t asp_flow_fit_params(int32_t kdims, int32_t n_layers) { return (size_t)n_layers * fit_layout(kdims).end; }

size_t asp_flow_fit_workspace_floats(int32_t kdims, int32_t n_layers, int32_t batch_cap) {
    const size_t bk = (size_t)batch_cap * kdims, bh = (size_t)batch_cap * kFitH;
    return 4 * bk + batch_cap + 2 * bh + 32 + (size_t)n_layers * (2 * bk + 2 * bh) + 64;
}

static int check_train(const asp_flow_train_t* t, const float* d_x, int32_t b) {
    ASP_REQUIRE(t && d_x, "null pointer");
    ASP_REQUIRE(t->kdims == 16 || t->kdims == 32 || t->kdims == 64, "kdims must be 16, 32 or 64");
    ASP_REQUIRE(t->hidden == kFitH, "hidden width must be 64");
    ASP_REQUIRE(t->n_layers >= 1 && t->n_layers <= 8, "n_layers must be in [1, 8]");
    ASP_REQUIRE(b >= 1 && b <= t->batch_cap, "batch size must be in [1, batch_cap]");
    ASP_REQUIRE(t->d_w && t->d_m && t->d_v && t->d_g && t->d_mean && t->d_inv_std && t->d_ws, "training descriptor incomplete");
    return ASP_OK;
}

int asp_flow_fit_loss(const asp_flow_train_t* t, const float* d_x, int32_t b, float log_const, float* d_loss,
                      void* stream) {
    ASP_DEVICE_OR_RETURN();
    int rc = check_train(t, d_x, b);
    if (rc) return rc;
    ASP_REQUIRE(d_loss, "null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    const FitWs W = fit_ws(t->d_ws, t->kdims, t->n_layers, t->batch_cap);
    const float* z = fit_forward(*t, d_x, b, W, s);
    fit_loss_kernel<<<1, 256, 0, s>>>(z, W.ladj, b, t->kdims, log_const, nullptr, d_loss);
    return check_launch("asp_flow_fit_loss");
}

int asp_flow_fit_step(const asp_flow_train_t* t, const float* d_x, int32_t b, float lr, float beta1, float beta2,
                      float eps, int32_t step, float clip_grad, float log_const, float* d_loss, void* stream) {
    ASP_DEVICE_OR_RETURN();
    int rc = check_train(t, d_x, b);
    if (rc) return rc;
    ASP_REQUIRE(d_loss && step >= 1, "bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    const int K = t->kdims, half = K / 2;
    const FitLayer L = fit_layout(K);
    const FitWs W = fit_ws(t->d_ws, K, t->n_layers, t->batch_cap);
    const float* z = fit_forward(*t, d_x, b, W, s);
    fit_loss_kernel<<<1, 256, 0, s>>>(z, W.ladj, b, K, log_const, W.g, d_loss);
    for (int l = t->n_layers - 1; l >= 0; --l) {
        const int odd = l & 1;
        float* base = W.layers + (size_t)l * W.per_layer;
        const float* z_in = base;
        const float* h1 = z_in + (size_t)t->batch_cap * K;
        const float* h2 = h1 + (size_t)t->batch_cap * kFitH;
        const float* o = h2 + (size_t)t->batch_cap * kFitH;
        const float* w = t->d_w + (size_t)l * L.end;
        float* g = t->d_g + (size_t)l * L.end;
        const unsigned char* mask = t->d_out_mask ? t->d_out_mask + (size_t)l * half : nullptr;
        fit_coupling_bwd_kernel<<<blocks_for((size_t)b * half), 256, 0, s>>>(z_in, o, b, K, odd, mask, W.g, W.d_o);
        // layer 3: o = h2 W3^T + b3
        gemm<true, false>(s, W.d_o, K, h2, kFitH, g + L.w3, kFitH, K, kFitH, b);                  // dW3 = dO^T h2
        fit_colsum_kernel<<<blocks_for(K), 256, 0, s>>>(W.d_o, K, b, K, g + L.b3);
        gemm<false, false>(s, W.d_o, K, w + L.w3, kFitH, W.d_h, kFitH, b, kFitH, K, nullptr, 0, h2, kFitH);  // dH2
        // layer 2
        gemm<true, false>(s, W.d_h, kFitH, h1, kFitH, g + L.w2, kFitH, kFitH, kFitH, b);
        fit_colsum_kernel<<<blocks_for(kFitH), 256, 0, s>>>(W.d_h, kFitH, b, kFitH, g + L.b2);
        gemm<false, false>(s, W.d_h, kFitH, w + L.w2, kFitH, W.d_h1, kFitH, b, kFitH, kFitH, nullptr, 0, h1, kFitH);  // dH1
        // layer 1: h1 = relu(z_a W1^T + b1)
        const float* za = z_in + (odd ? half : 0);
        gemm<true, false>(s, W.d_h1, kFitH, za, K, g + L.w1, half, kFitH, half, b);
        fit_colsum_kernel<<<blocks_for(kFitH), 256, 0, s>>>(W.d_h1, kFitH, b, kFitH, g + L.b1);
        gemm<false, false>(s, W.d_h1, kFitH, w + L.w1, half, W.d_za, half, b, half, kFitH);       // dZa
        fit_add_half_kernel<<<blocks_for((size_t)b * half), 256, 0, s>>>(W.g, W.d_za, b, K, odd);
    }
    const size_t n = (size_t)t->n_layers * L.end;
    if (clip_grad > 0.f) {
        cudaMemsetAsync(W.sumsq, 0, sizeof(float), s);
        fit_sumsq_kernel<<<32, 256, 0, s>>>(t->d_g, n, W.sumsq);
    }
    const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
    fit_adam_kernel<<<blocks_for(n), 256, 0, s>>>(t->d_w, t->d_m, t->d_v, t->d_g, n, lr, beta1, beta2, eps, bc1, bc2,
                                                 clip_grad, W.sumsq);
    return check_launch("asp_flow_fit_step");
}

}  // extern "C"

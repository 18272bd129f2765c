// Issue-rate probe for the instruction classes of the pCN sweep (sm_100a): cycles per warp-instruction
// per SM sub-partition for each class alone and for pairs of classes issued together.  Tells which
// classes share a dispatch path (pair time = sum) and which overlap (pair time = max).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int ITER = 4096;
constexpr int NCH = 8;  // independent chains per thread

enum Op { IMADW, DFMA_, MUFU_, I2FP_, F2F_, FFMA_, LOP3_, IMAD_, NONE };

template <int OP>
__device__ __forceinline__ void step(uint32_t (&a)[NCH], uint64_t (&w)[NCH], double (&d)[NCH], float (&f)[NCH]) {
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        if constexpr (OP == IMADW) asm volatile("mul.wide.u32 %0, %1, 0xD2511F53;" : "=l"(w[c]) : "r"((uint32_t)w[c]));
        if constexpr (OP == DFMA_) asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d[c]) : "d"(d[(c + 1) % NCH]));
        if constexpr (OP == MUFU_) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(f[c]));
        if constexpr (OP == I2FP_) asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(f[c]) : "r"(a[c]));
        if constexpr (OP == F2F_) asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d[c]) : "f"(f[c]));
        if constexpr (OP == FFMA_) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[c]) : "f"(f[(c + 1) % NCH]));
        if constexpr (OP == LOP3_) asm volatile("lop3.b32 %0, %0, %1, 0x9E3779B9, 0x96;" : "+r"(a[c]) : "r"(a[(c + 1) % NCH]));
        if constexpr (OP == IMAD_) asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(a[c]) : "r"(a[(c + 1) % NCH]));
    }
}

template <int A, int B>
__global__ void __launch_bounds__(512, 1) probe(long long* cycles, double* sink) {
    uint32_t a[NCH], a2[NCH];
    uint64_t w[NCH], w2[NCH];
    double d[NCH], d2[NCH];
    float f[NCH], f2[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        a[c] = threadIdx.x * 7 + c; a2[c] = a[c] + 3; w[c] = a[c] + 11; w2[c] = a[c] + 13;
        d[c] = 1.0 + 1e-9 * a[c]; d2[c] = d[c] + 1e-3; f[c] = 1.5f + c + threadIdx.x; f2[c] = f[c] + 0.25f;
    }
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
        step<A>(a, w, d, f);
        step<B>(a2, w2, d2, f2);
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < NCH; ++c) s += d[c] + d2[c] + f[c] + f2[c] + a[c] + a2[c] + (double)w[c] + (double)w2[c];
    if (s == 12345.678) sink[0] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int A, int B>
double run(const char* name, int na, int nb) {
    long long* cyc; double* sink;
    cudaMalloc(&cyc, 148 * sizeof(long long)); cudaMalloc(&sink, 8);
    probe<A, B><<<148, 512>>>(cyc, sink);
    probe<A, B><<<148, 512>>>(cyc, sink);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    // 512 threads = 16 warps = 4 warps per sub-partition; per iteration each warp issues NCH ops of each class
    const double per_iter = avg / ITER;            // cycles per iteration (4 warps interleaved)
    const double warp_instr = 4.0 * NCH;           // warp-instructions of ONE class per iteration per sub-partition
    printf("%-18s cycles/iter %8.1f  -> %6.2f cycles per warp-instr pair-slot (A x%d, B x%d)\n", name, per_iter,
           per_iter / warp_instr, na, nb);
    cudaFree(cyc); cudaFree(sink);
    return per_iter / warp_instr;
}

int main() {
    printf("single classes (cycles per warp-instruction per sub-partition):\n");
    run<IMADW, NONE>("IMAD.WIDE", 1, 0);
    run<DFMA_, NONE>("DFMA", 1, 0);
    run<MUFU_, NONE>("MUFU.LG2", 1, 0);
    run<I2FP_, NONE>("I2FP.F32.U32", 1, 0);
    run<F2F_, NONE>("F2F.F64.F32", 1, 0);
    run<FFMA_, NONE>("FFMA", 1, 0);
    run<LOP3_, NONE>("LOP3", 1, 0);
    run<IMAD_, NONE>("IMAD", 1, 0);
    printf("pairs (one of each per slot; sum of singles = serialised, max = overlapped):\n");
    run<IMADW, DFMA_>("IMAD.WIDE+DFMA", 1, 1);
    run<IMADW, MUFU_>("IMAD.WIDE+MUFU", 1, 1);
    run<DFMA_, MUFU_>("DFMA+MUFU", 1, 1);
    run<DFMA_, F2F_>("DFMA+F2F", 1, 1);
    run<MUFU_, I2FP_>("MUFU+I2FP", 1, 1);
    run<MUFU_, F2F_>("MUFU+F2F", 1, 1);
    run<IMADW, FFMA_>("IMAD.WIDE+FFMA", 1, 1);
    run<IMADW, LOP3_>("IMAD.WIDE+LOP3", 1, 1);
    run<IMADW, IMAD_>("IMAD.WIDE+IMAD", 1, 1);
    run<DFMA_, FFMA_>("DFMA+FFMA", 1, 1);
    run<DFMA_, LOP3_>("DFMA+LOP3", 1, 1);
    run<MUFU_, FFMA_>("MUFU+FFMA", 1, 1);
    run<MUFU_, LOP3_>("MUFU+LOP3", 1, 1);
    run<FFMA_, LOP3_>("FFMA+LOP3", 1, 1);
    return 0;
}

// Dependent-chain latency probe (sm_100a): one warp, one chain, cycles per instruction.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o latency latency.cu && ./latency
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
constexpr int ITER = 2048;
constexpr int UNR = 16;
enum Op { IMADW, DFMA_, DADD_, MUFU_, I2FP_, F2F_, FFMA_, LOP3_, IMAD_, LDSB, SHFL_ };
template <int OP>
__global__ void probe(long long* cycles, double* sink, int zero) {
    __shared__ double2 sm[64];
    if (threadIdx.x < 64) sm[threadIdx.x] = make_double2(1.0 + zero, zero);
    __syncthreads();
    uint32_t a = threadIdx.x + 5, off = (uint32_t)__cvta_generic_to_shared(sm);
    uint64_t w = a;
    double d = 1.0 + 1e-9 * a, e = 1e-12;
    float f = 1.5f + a;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            if constexpr (OP == IMADW) asm volatile("mul.wide.u32 %0, %1, 0xD2511F53;" : "=l"(w) : "r"((uint32_t)w));
            if constexpr (OP == DFMA_) asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d) : "d"(e));
            if constexpr (OP == DADD_) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d) : "d"(e));
            if constexpr (OP == MUFU_) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(f));
            if constexpr (OP == I2FP_) { asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(f) : "r"(a)); asm volatile("mov.b32 %0, %1;" : "=r"(a) : "f"(f)); }
            if constexpr (OP == F2F_) { asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d) : "f"(f)); asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f) : "d"(d)); }
            if constexpr (OP == FFMA_) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f) : "f"(1.0001f));
            if constexpr (OP == LOP3_) asm volatile("lop3.b32 %0, %0, %1, 0x9E3779B9, 0x96;" : "+r"(a) : "r"(zero));
            if constexpr (OP == IMAD_) asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(a) : "r"(zero + 3));
            if constexpr (OP == LDSB) { double2 r; asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"(off)); off += (uint32_t)__double2loint(r.y); }
            if constexpr (OP == SHFL_) a = __shfl_xor_sync(0xffffffffu, a, 1);
        }
    }
    const long long t1 = clock64();
    if (d + f + a + (double)w + off == 12345.678) sink[0] = d;
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
}
template <int OP>
void run(const char* name, int per) {
    long long* cyc; double* sink;
    cudaMalloc(&cyc, 8); cudaMalloc(&sink, 8);
    probe<OP><<<1, 32>>>(cyc, sink, 0);
    probe<OP><<<1, 32>>>(cyc, sink, 0);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-22s %6.2f cycles per dependent step (%d instr per step)\n", name, (double)h / ITER / UNR, per);
    cudaFree(cyc); cudaFree(sink);
}
int main() {
    run<IMADW>("IMAD.WIDE", 1); run<DFMA_>("DFMA", 1); run<DADD_>("DADD", 1); run<MUFU_>("MUFU.LG2", 1);
    run<I2FP_>("I2FP+MOV", 2); run<F2F_>("F2F.F64.F32+F2F.F32.F64", 2); run<FFMA_>("FFMA", 1); run<LOP3_>("LOP3", 1);
    run<IMAD_>("IMAD", 1); run<LDSB>("LDS.128 (+IADD)", 2); run<SHFL_>("SHFL", 1);
    return 0;
}

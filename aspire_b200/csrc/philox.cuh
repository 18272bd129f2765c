// Philox4x32 (Salmon et al., SC'11; 10 rounds, 7 for the sweep noise) and the uniform / normal conversions.
// Bit layout and conversions are specified in oracle/philox.py; the generator is
// pinned there against the Random123 known-answer vectors.
#pragma once
#include <stdint.h>

namespace asp {

constexpr uint32_t kPhiloxM0 = 0xD2511F53u;
constexpr uint32_t kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u;
constexpr uint32_t kPhiloxW1 = 0xBB67AE85u;
constexpr uint32_t kAcceptBlock = 0xFFFFFFFFu;
// Rounds of the sweep's noise streams (proposal normals, accept uniforms, tpCN gamma variates): Philox4x32-7,
// the smallest round count reported crush-resistant by Salmon et al.; the uniform / fp64-normal fills keep 10.
// Measured on B200 (profiles/r2_sweep_experiments.md): 10 -> 7 rounds takes the 10-move sweep from 10.44 to
// 9.26 ms per 2e7 x 32-D particles.  Spec: oracle/philox.py::NOISE_ROUNDS.
#ifndef ASP_NOISE_ROUNDS
#define ASP_NOISE_ROUNDS 7
#endif
constexpr int kNoiseRounds = ASP_NOISE_ROUNDS;

template <int R>
__device__ __forceinline__ uint4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                            uint32_t k1) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const uint64_t p0 = (uint64_t)kPhiloxM0 * c0, p1 = (uint64_t)kPhiloxM1 * c2;  // one IMAD.WIDE each
        const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0;
        c1 = lo1;
        c2 = n2;
        c3 = lo0;
        k0 += kPhiloxW0;
        k1 += kPhiloxW1;
    }
    return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                               uint32_t k1) {
    return philox4x32<10>(c0, c1, c2, c3, k0, k1);
}

// Round keys precomputed on the host and passed by value in the kernel parameters:
// constant-bank operands fold straight into the LOP3 of each round.
struct PhiloxKeys {
    uint32_t k0[10], k1[10];
};
inline PhiloxKeys philox_keys(uint64_t seed) {
    PhiloxKeys K;
    uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        K.k0[r] = a;
        K.k1[r] = b;
        a += kPhiloxW0;
        b += kPhiloxW1;
    }
    return K;
}
template <int R>
__device__ __forceinline__ uint4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKeys& K) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const uint64_t p0 = (uint64_t)kPhiloxM0 * c0, p1 = (uint64_t)kPhiloxM1 * c2;  // one IMAD.WIDE each
        const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        const uint32_t n0 = hi1 ^ c1 ^ K.k0[r], n2 = hi0 ^ c3 ^ K.k1[r];
        c0 = n0;
        c1 = lo1;
        c2 = n2;
        c3 = lo0;
    }
    return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               const PhiloxKeys& K) {
    return philox4x32<10>(c0, c1, c2, c3, K);
}

__device__ __forceinline__ float sfu_lg2(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sfu_rsqrt(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sfu_sin(float x) { float r; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sfu_cos(float x) { float r; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
// natural log on the SFU (absolute error ~1e-7 for u in (0, 1])
__device__ __forceinline__ float sfu_log(float x) { return 0.6931471805599453f * sfu_lg2(x); }

// fp32 uniform in (0, 1]
__device__ __forceinline__ float u01_f32(uint32_t r) { return fmaf(__uint2float_rn(r), 0x1p-32f, 0x1p-33f); }

// two fp32 normals from two words: Box-Muller on the SFU (MUFU.LG2 / SQRT / SIN / COS).
//   u = fma(ra, 2^-32, 2^-33) in (0, 1],  s = sqrt(-2 ln2 * log2 u)
//   theta = fma(rb, pi 2^-31, -pi) in [-pi, pi]   (the SFU's accurate range)
// Absolute error of a normal ~1e-6 (spec: oracle/philox.py::box_muller_f32).
__device__ __forceinline__ float sfu_sqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ void box_muller_f32(uint32_t ra, uint32_t rb, float& z0, float& z1) {
    const float u = u01_f32(ra);
    const float th = fmaf(__uint2float_rn(rb), 1.4629180792671596e-09f, -3.14159265358979f);
    const float s = sfu_sqrt(-1.3862943611198906f * sfu_lg2(u));
    z0 = s * sfu_cos(th);
    z1 = s * sfu_sin(th);
}

// four fp32 normals for (element index, stream, block)
__device__ __forceinline__ void philox_normal4(uint64_t seed, uint32_t stream, uint64_t index, uint32_t block,
                                               float (&z)[4]) {
    const uint4 r = philox4x32<kNoiseRounds>((uint32_t)index, (uint32_t)(index >> 32), stream, block, (uint32_t)seed,
                                             (uint32_t)(seed >> 32));
    box_muller_f32(r.x, r.y, z[0], z[1]);
    box_muller_f32(r.z, r.w, z[2], z[3]);
}

__device__ __forceinline__ void philox_normal4(const PhiloxKeys& K, uint32_t stream, uint64_t index, uint32_t block,
                                               float (&z)[4]) {
    const uint4 r = philox4x32<kNoiseRounds>((uint32_t)index, (uint32_t)(index >> 32), stream, block, K);
    box_muller_f32(r.x, r.y, z[0], z[1]);
    box_muller_f32(r.z, r.w, z[2], z[3]);
}

// One iteration of Marsaglia & Tsang's (2000) Gamma(a, 1) sampler, a >= 1, in fp32 from one Philox block (words 0, 1
// -> the first Box-Muller normal, word 2 -> the uniform); d = a - 1/3, c = 1 / sqrt(9 d).  Spec:
// oracle/pcn.py::gamma_mt_f32.  Shared by asp_tpcn_scales and the fused tpCN sweep so both draw the same variates.
__device__ __forceinline__ bool gamma_mt_try(const uint4& r, float d, float c, double& g) {
    float z, z1;
    box_muller_f32(r.x, r.y, z, z1);
    const float u = u01_f32(r.z);
    const float t = 1.0f + c * z;
    const float v = t * t * t;
    g = (double)(d * v);
    return t > 0.0f && sfu_log(u) < 0.5f * z * z + d - d * v + d * sfu_log(fmaxf(v, 1e-30f));
}

__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {
    return (double)((((uint64_t)hi << 32) | lo) >> 11) * 0x1p-53;
}

}  // namespace asp

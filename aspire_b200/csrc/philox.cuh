// Philox4x32-10 (Salmon et al., SC'11) and the uniform / normal conversions.
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

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                               uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(kPhiloxM0, c0), lo0 = kPhiloxM0 * c0;
        const uint32_t hi1 = __umulhi(kPhiloxM1, c2), lo1 = kPhiloxM1 * c2;
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

// fp32 uniform in (0, 1]
__device__ __forceinline__ float u01_f32(uint32_t r) { return fmaf(__uint2float_rn(r), 0x1p-32f, 0x1p-33f); }

// two fp32 normals from two words
__device__ __forceinline__ void box_muller_f32(uint32_t ra, uint32_t rb, float& z0, float& z1) {
    const float u = u01_f32(ra);
    const float t = fmaf(__uint2float_rn(rb), 0x1p-31f, 0x1p-32f);
    const float s = sqrtf(-2.0f * logf(u));
    float sn, cs;
    sincospif(t, &sn, &cs);
    z0 = s * cs;
    z1 = s * sn;
}

// four fp32 normals for (element index, stream, block)
__device__ __forceinline__ void philox_normal4(uint64_t seed, uint32_t stream, uint64_t index, uint32_t block,
                                               float (&z)[4]) {
    const uint4 r = philox4x32_10((uint32_t)index, (uint32_t)(index >> 32), stream, block, (uint32_t)seed,
                                  (uint32_t)(seed >> 32));
    box_muller_f32(r.x, r.y, z[0], z[1]);
    box_muller_f32(r.z, r.w, z[2], z[3]);
}

// log(u) for the accept/reject test, fp32 logf promoted to fp64
__device__ __forceinline__ double philox_accept_log_u(uint64_t seed, uint32_t stream, uint64_t index) {
    const uint4 r = philox4x32_10((uint32_t)index, (uint32_t)(index >> 32), stream, kAcceptBlock, (uint32_t)seed,
                                  (uint32_t)(seed >> 32));
    return (double)logf(u01_f32(r.x));
}

__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {
    return (double)((((uint64_t)hi << 32) | lo) >> 11) * 0x1p-53;
}

}  // namespace asp

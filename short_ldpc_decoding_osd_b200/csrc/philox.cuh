// Philox4x32-10 and the Box-Muller pairing of the frame generator, shared by framegen.cu (gen_frames_kernel) and
// nms_qc.cu (the Monte-Carlo kernel that generates its frames in the decoder's prologue).  One definition, so both
// produce the same floats bit for bit (tests/test_gpu_pipeline.py::test_simulate_equals_decode_of_generated_frames).
#pragma once
#include "common.cuh"

namespace ldpcb {

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
    const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        const unsigned hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += W0;
        k.y += W1;
    }
    return c;
}

// u in (0,1): 23 random bits + 1/2, exactly representable in fp32
__device__ __forceinline__ float u01(unsigned x) { return ((float)(x >> 9) + 0.5f) * 1.1920928955078125e-07f; }

// radius uniform from all 32 bits: u = fl32(fl32((float)x + 0.5) * 2^-32), clamped below 1.  The smallest value is 2^-33, so the
// Gaussian tail reaches sqrt(66 ln 2) = 6.76 sigma (5.77 with the 23-bit form above: ADVICE r1; at 1e8 frames x 128 samples the
// 5.77-sigma cut, probability 8e-9 per sample, was reached about a hundred times per point)
__device__ __forceinline__ float u01_32(unsigned x) {
    return fminf(__fmul_rn(__fadd_rn(__uint2float_rn(x), 0.5f), 2.3283064365386963e-10f), 0.99999994f);
}

__device__ __forceinline__ void box_muller(unsigned a, unsigned b, float& z0, float& z1) {
    const float r = sqrtf(-2.0f * logf(u01_32(a)));
    float s, c;
    sincospif(2.0f * u01(b), &s, &c);
    z0 = r * c;
    z1 = r * s;
}

// the four channel values of Philox block `blk` (code positions 4*blk .. 4*blk+3) of frame `fr`, before the BPSK sign
__device__ __forceinline__ void gen_block(uint2 key, uint64_t fr, unsigned blk, float sigma, float (&y)[4]) {
    const uint4 x = philox4x32_10(make_uint4((unsigned)fr, (unsigned)(fr >> 32), blk, 0u), key);
    float z[4];
    box_muller(x.x, x.y, z[0], z[1]);
    box_muller(x.z, x.w, z[2], z[3]);
#pragma unroll
    for (int i = 0; i < 4; ++i) y[i] = __fmaf_rn(sigma, z[i], 1.0f);
}
// the 64 message bits of frame `fr`
__device__ __forceinline__ unsigned long long gen_message(uint2 key, uint64_t fr) {
    const uint4 mw = philox4x32_10(make_uint4((unsigned)fr, (unsigned)(fr >> 32), 0u, 1u), key);
    return (unsigned long long)mw.x | ((unsigned long long)mw.y << 32);
}
inline float ebn0_to_sigma(float ebn0_db) {
    const double rate = (double)K / (double)N;
    return (float)sqrt(1.0 / (2.0 * rate * pow(10.0, (double)ebn0_db / 10.0)));
}

}  // namespace ldpcb

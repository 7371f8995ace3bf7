// Counter-based BPSK/AWGN frame generator, one warp per frame.
//
// Replaces testing_data_generating (reference LDPC_128/Testing_data_gen_128/data_generating.py:13-51,
// AWGN branch with random codewords): sigma (:17), channel = N(1, sigma) (:40), message bits (:42),
// cw = msg.G mod 2 (:43), y = +channel for bit 0 and -channel for bit 1 (:44), labels = cw (:45).
// The reference draws from NumPy's global MT19937 stream, which cannot be sharded; here frame f is
// a pure function of (seed, f) through Philox4x32-10 (Salmon et al., SC'11; the same generator as
// cuRAND's curand_philox4x32_x.h), so any rank can produce any slice of a run.
//   counter = (f_lo, f_hi, block, stream), key = (seed_lo, seed_hi)
//   stream 0, block b = 0..31: words (x0,x1) and (x2,x3) -> Box-Muller pairs -> normals 4b..4b+3
//   stream 1, block 0: words x0 | x1<<32 -> the 64 message bits
#include "common.cuh"
#include "internal.cuh"
#include "philox.cuh"

namespace ldpcb {

constexpr int GEN_WARPS = 8;

__global__ void __launch_bounds__(GEN_WARPS * 32) gen_frames_kernel(uint2 key, uint64_t first_frame, int64_t B, float sigma,
                                                                     const uint64_t* __restrict__ gcol, float* llr, uint32_t* cw_bits) {
    const int lane = threadIdx.x & 31;
    const int64_t gw = (int64_t)blockIdx.x * GEN_WARPS + (threadIdx.x >> 5);
    const int64_t nw = (int64_t)gridDim.x * GEN_WARPS;
    unsigned long long g[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) g[i] = gcol[4 * lane + i];
    for (int64_t f = gw; f < B; f += nw) {
        const uint64_t fr = first_frame + (uint64_t)f;
        // message bits (computed by every lane: cheaper than a broadcast of two words + divergence)
        const unsigned long long msg = gen_message(key, fr);
        unsigned nib = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) nib |= (unsigned)(__popcll(msg & g[i]) & 1) << i;
        if (llr) {
            float yy[4];
            gen_block(key, fr, (unsigned)lane, sigma, yy);
            float4 y = make_float4(yy[0], yy[1], yy[2], yy[3]);
            if (nib & 1u) y.x = -y.x;
            if (nib & 2u) y.y = -y.y;
            if (nib & 4u) y.z = -y.z;
            if (nib & 8u) y.w = -y.w;
            reinterpret_cast<float4*>(llr + f * N)[lane] = y;
        }
        if (cw_bits) {
            unsigned v = nib << (4 * (lane & 7));
            v |= __shfl_xor_sync(0xffffffffu, v, 1);
            v |= __shfl_xor_sync(0xffffffffu, v, 2);
            v |= __shfl_xor_sync(0xffffffffu, v, 4);
            if ((lane & 7) == 0) cw_bits[f * 4 + (lane >> 3)] = v;
        }
    }
}

int launch_gen(ldpcb_handle* h, uint64_t seed, uint64_t first_frame, int64_t B, float ebn0_db, float* llr,
               uint32_t* cw_bits, cudaStream_t st) {
    if (B == 0) return LDPCB_OK;
    const float sigma = ebn0_to_sigma(ebn0_db);
    int64_t want = (B + GEN_WARPS - 1) / GEN_WARPS;
    int64_t cap = (int64_t)h->sm_count * 8;
    const uint2 key = make_uint2((unsigned)seed, (unsigned)(seed >> 32));
    gen_frames_kernel<<<(int)(want < cap ? want : cap), GEN_WARPS * 32, 0, st>>>(key, first_frame, B, sigma, h->gcol_dev, llr, cw_bits);
    LDPCB_LAUNCH_CHECK(h, "gen_frames_kernel");
    return LDPCB_OK;
}

}  // namespace ldpcb

using namespace ldpcb;

extern "C" int ldpcb_gen_frames(ldpcb_t* h, uint64_t seed, uint64_t first_frame, int64_t B, float ebn0_db,
                                float* llr_dev, uint32_t* cw_bits_dev, void* stream) {
    LDPCB_ENTER(h);
    if (B < 0 || (!llr_dev && !cw_bits_dev)) return set_error(h, LDPCB_ERR_ARG, "ldpcb_gen_frames: bad arguments");
    if (llr_dev && ((uintptr_t)llr_dev & 15)) return set_error(h, LDPCB_ERR_ALIGN, "ldpcb_gen_frames: llr must be 16-byte aligned");
    return launch_gen(h, seed, first_frame, B, ebn0_db, llr_dev, cw_bits_dev, (cudaStream_t)stream);
}

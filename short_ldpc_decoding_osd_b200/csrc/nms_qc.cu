// Normalized min-sum BP for the CCSDS (128,64) code, specialised to its quasi-cyclic structure: HALF a warp per frame,
// every message exchange is a 16-lane rotation done by one segmented warp shuffle -- no shared memory at all.
//
// Replaces the same reference code as nms.cu (Decoder_Layer.belief_propagation_op and its per-iteration ops,
// LDPC_128/Ldpc_128_testing/ms_test.py:106-137,180-242; hard decision + syndrome of get_eval :38-44,51) with the same
// arithmetic, operation for operation, so every message is bit-identical to nms.cu's and to the oracle's.
//
// H is a 4 x 8 array of 16 x 16 circulants: check 16R+i is connected to variable 16C + (i+s) mod 16 for every
// "class" (R, C, s) of the table below (32 classes: 8 per block row; diagonal blocks hold two, blocks (R, R+4) none).
// Lane i of a half-warp owns the four checks {16R + (i + rho_R) mod 16} and the eight variables
// {16C + (i + sigma_C) mod 16}: its 32 check->variable messages, 8 channel values and 8 variable totals live in
// registers.  A class then connects check-owner lane a to variable-owner lane a + delta, delta = s + rho_R - sigma_C
// (mod 16): one __shfl_sync(.., width 16) moves the 16 messages of a class of BOTH frames of the warp, and the 14
// classes with delta = 0 (rho, sigma were chosen to maximise them) need no exchange at all.  Per iteration and frame
// that is 18 shuffle operations for 512 messages in each direction, against 52 shared-memory operations in nms.cu,
// and 176 instead of 236 issued instructions (2252 instead of 2833 per frame, ncu).
//
// Variable totals are summed in ascending check order like tf.reduce_sum over the check axis (oracle/nms_oracle.py):
// block rows ascending; the two checks a variable of a diagonal block has in ONE block row are ordered by a per-lane
// predicate (they commute when they are the first two terms).
#include <type_traits>
#include <utility>

#include "common.cuh"
#include "internal.cuh"
#include "philox.cuh"

namespace ldpcb {
namespace qc {

// ---- the code, as compile-time functions (usable as constant expressions inside the unrolled kernel) -----------------
struct CS { int C, s; };
__host__ __device__ constexpr CS edge(int R, int e) {
    // CCSDS 131.1-O-2 (128,64): [block row][edge] = {variable block, circulant shift}, ascending variable block
    constexpr int T[4][8][2] = {{{0, 0}, {0, 7}, {1, 2}, {2, 14}, {3, 6}, {5, 0}, {6, 13}, {7, 0}},
                                {{0, 6}, {1, 0}, {1, 15}, {2, 0}, {3, 1}, {4, 0}, {6, 0}, {7, 7}},
                                {{0, 4}, {1, 1}, {2, 0}, {2, 15}, {3, 14}, {4, 11}, {5, 0}, {7, 3}},
                                {{0, 0}, {1, 1}, {2, 9}, {3, 0}, {3, 13}, {4, 14}, {5, 1}, {6, 0}}};
    return CS{T[R][e][0], T[R][e][1]};
}
// lane relabelling that makes 14 of the 32 classes exchange-free and leaves 10 distinct rotation amounts
// (exhaustive search over rho with the best sigma per variable block, scripts in DESIGN.md 4.1)
__host__ __device__ constexpr int rho(int R) {
    constexpr int T[4] = {0, 13, 15, 15};
    return T[R];
}
__host__ __device__ constexpr int sig(int C) {
    constexpr int T[8] = {3, 0, 14, 12, 13, 0, 13, 0};
    return T[C];
}
// rotation of class (R, e): the variable's owner is `delta` lanes above the check's owner (mod 16)
__host__ __device__ constexpr int delta(int R, int e) { return (edge(R, e).s + rho(R) - sig(edge(R, e).C) + 32) % 16; }
__host__ __device__ constexpr int var_degree(int C) { return C < 4 ? 5 : 3; }
// k-th incoming edge of the variables of block C in ascending block-row order -> R * 8 + e
__host__ __device__ constexpr int incoming(int C, int k) {
    int n = 0;
    for (int R = 0; R < 4; ++R)
        for (int e = 0; e < 8; ++e)
            if (edge(R, e).C == C) {
                if (n == k) return R * 8 + e;
                ++n;
            }
    return -1;
}
// position of the first of the two same-row edges of a diagonal block in the incoming list (C < 4), else -1
__host__ __device__ constexpr int twin_pos(int C) {
    for (int k = 0; k + 1 < var_degree(C); ++k)
        if (incoming(C, k) / 8 == incoming(C, k + 1) / 8) return k;
    return -1;
}

template <int... I, class F>
__device__ __forceinline__ void static_for_impl(std::integer_sequence<int, I...>, F&& f) {
    (f(std::integral_constant<int, I>{}), ...);
}
template <int N_, class F>
__device__ __forceinline__ void static_for(F&& f) {
    static_for_impl(std::make_integer_sequence<int, N_>{}, static_cast<F&&>(f));
}

// expansion of the table into a dense H (host): the launcher only takes this kernel when the handle's H equals it
inline bool matches(const uint8_t* H) {
    uint8_t E[M * N] = {};
    for (int R = 0; R < 4; ++R)
        for (int e = 0; e < 8; ++e)
            for (int i = 0; i < 16; ++i) E[(16 * R + i) * N + 16 * edge(R, e).C + (i + edge(R, e).s) % 16] = 1;
    for (int i = 0; i < M * N; ++i)
        if ((H[i] & 1) != E[i]) return false;
    return true;
}

}  // namespace qc

constexpr int QC_WARPS = 8;
constexpr int QC_THREADS = QC_WARPS * 32;

__device__ __forceinline__ float rot16(float v, int src) { return __shfl_sync(0xffffffffu, v, src, 16); }
__device__ __forceinline__ unsigned rot16u(unsigned v, int src) { return __shfl_sync(0xffffffffu, v, src, 16); }

// TRAJ: write the iters+1 posteriors; FIR: accumulate the DIA reliability in registers (see nms.cu);
// FUSE: tallies against truth_bits, append of the frames with a non-zero syndrome to (fail_idx, fail_count) and, when
//       the frames are generated here (GEN), their channel LLRs + transmitted codewords to fail_llr / fail_truth;
// GEN:  generate the frame in the prologue (Philox, same floats as gen_frames_kernel) instead of loading it
// MINB: resident CTAs per SM the register allocation is held to (2: 110 registers, 3: 80 registers, no spills either way)
template <bool TRAJ, bool FIR, bool FUSE, bool GEN, int MINB>
__global__ void __launch_bounds__(QC_THREADS, MINB) nms_qc_kernel(NmsArgs a, NmsFuse z) {
    using namespace qc;
    __shared__ unsigned long long sh_cnt[LDPCB_NUM_COUNTERS];
    __shared__ __align__(16) float sh_gen[GEN ? QC_WARPS * 2 * N : 4];
    const int lane = threadIdx.x & 31, li = lane & 15, half = lane >> 4, warp = threadIdx.x >> 5;
    if (FUSE) {
        if (threadIdx.x < LDPCB_NUM_COUNTERS) sh_cnt[threadIdx.x] = 0ull;
        __syncthreads();
    }
    const int64_t pairs = (a.B + 1) >> 1;
    const int64_t gw = (int64_t)blockIdx.x * QC_WARPS + warp, nw = (int64_t)gridDim.x * QC_WARPS;
    const int rows = a.iters + 1;
    const float w = a.w_vc;  // == w_marg (the launcher routes w_vc != w_marg to nms.cu)
    // per-lane ordering of the two same-row checks of the diagonal blocks C = 1, 2, 3 (C = 0: they are the first two
    // terms of the sum and commute): the check of edge e1 precedes the check of edge e2 iff its index is smaller
    bool first_lo[4];
    static_for<4>([&](auto Cc) {
        constexpr int C = decltype(Cc)::value;
        constexpr int k = twin_pos(C);
        constexpr int s1 = edge(incoming(C, k) / 8, incoming(C, k) % 8).s, s2 = edge(incoming(C, k + 1) / 8, incoming(C, k + 1) % 8).s;
        const int t = (li + sig(C)) & 15;
        first_lo[C] = ((t - s1) & 15) < ((t - s2) & 15);
    });
    // tallies of this thread (only lanes 0 and 16 count)
    unsigned c_frames = 0, c_fe = 0, c_be = 0, c_det = 0, c_und = 0, c_ffe = 0, c_fbe = 0;

    for (int64_t pair = gw; pair < pairs; pair += nw) {
        const int64_t f = 2 * pair + half;
        const bool active = f < a.B;
        const int64_t fr = active ? f : a.B - 1;  // the idle half of an odd batch decodes the last frame again, stores nothing
        const int64_t row = a.idx ? (int64_t)a.idx[fr] : fr;
        float y[8];
        unsigned truth_w[4] = {0u, 0u, 0u, 0u};
        if (GEN) {
            // lane i generates Philox blocks 2i, 2i+1 (code positions 8i..8i+7), applies the BPSK sign of its 8 codeword
            // bits, and the half-warp transposes through 512 B of shared memory into the owner layout
            const uint64_t gf = z.first_frame + (uint64_t)fr;
            const unsigned long long msg = gen_message(z.key, gf);
            float* G = sh_gen + (warp * 2 + half) * N;
            unsigned cwb = 0;
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                float v[4];
                gen_block(z.key, gf, (unsigned)(2 * li + b), z.sigma, v);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const unsigned bit = (unsigned)(__popcll(msg & __ldg(z.gcol + 8 * li + 4 * b + j)) & 1);
                    cwb |= bit << (4 * b + j);
                    if (bit) v[j] = -v[j];
                }
                reinterpret_cast<float4*>(G)[2 * li + b] = make_float4(v[0], v[1], v[2], v[3]);
            }
            // transmitted codeword as four words: byte (li & 3) of word li >> 2
            unsigned tw = cwb << (8 * (li & 3));
            tw |= __shfl_xor_sync(0xffffffffu, tw, 1);
            tw |= __shfl_xor_sync(0xffffffffu, tw, 2);
#pragma unroll
            for (int k = 0; k < 4; ++k) truth_w[k] = rot16u(tw, 4 * k);
            __syncwarp();
            static_for<8>([&](auto Cc) {
                constexpr int C = decltype(Cc)::value;
                y[C] = G[16 * C + ((li + sig(C)) & 15)];
            });
            __syncwarp();
        } else {
            const float* src = a.llr + row * N;
            static_for<8>([&](auto Cc) {
                constexpr int C = decltype(Cc)::value;
                y[C] = __ldg(src + 16 * C + ((li + sig(C)) & 15));
            });
        }
        if (TRAJ && active) {
            static_for<8>([&](auto Cc) {
                constexpr int C = decltype(Cc)::value;
                a.soft_traj[(f * rows) * N + 16 * C + ((li + sig(C)) & 15)] = y[C];
            });
        }
        float T[8], wy[8], fir[8];
#pragma unroll
        for (int C = 0; C < 8; ++C) {
            wy[C] = __fmul_rn(w, y[C]);
            T[C] = wy[C];  // cv = 0: total = 0 + w*y; (ms_test.py:118,131)
            if (FIR) fir[C] = __fmaf_rn(a.fir_taps[0], y[C], 0.0f);
        }
        float cv[4][8];
#pragma unroll
        for (int R = 0; R < 4; ++R)
#pragma unroll
            for (int e = 0; e < 8; ++e) cv[R][e] = 0.0f;

        // two iterations per trip save the register moves of the loop-carried messages: plain kernel 5.60 -> 5.48 ms per 2^21
        // frames, generator-fused kernel 4.58 -> 4.52 ms per 2^20 (with OSD); the fused decode kernel alone is slower unrolled
        // (5.56 -> 5.63 ms, measured twice) and keeps one iteration per trip
#pragma unroll((FUSE && !GEN) ? 1 : 2)
        for (int it = 0; it < a.iters; ++it) {
            // ---- check phase: vc = total - cv_old (ms_test.py:133-136), min1/min2/sign (:184-206), new cv (:207-209) ----
            static_for<4>([&](auto Rc) {
                constexpr int R = decltype(Rc)::value;
                float x[8], ax[8];
                static_for<8>([&](auto ec) {
                    constexpr int e = decltype(ec)::value;
                    constexpr int C = edge(R, e).C, d = delta(R, e);
                    const float t = d == 0 ? T[C] : rot16(T[C], li + d);
                    x[e] = __fsub_rn(t, cv[R][e]);
                    ax[e] = fabsf(x[e]);
                });
                float lo[4], hi[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    lo[i] = fminf(ax[2 * i], ax[2 * i + 1]);
                    hi[i] = fmaxf(ax[2 * i], ax[2 * i + 1]);
                }
                const float l01 = fminf(lo[0], lo[1]), h01 = fminf(fmaxf(lo[0], lo[1]), fminf(hi[0], hi[1]));
                const float l23 = fminf(lo[2], lo[3]), h23 = fminf(fmaxf(lo[2], lo[3]), fminf(hi[2], hi[3]));
                float m1 = fminf(l01, l23);
                float m2 = fminf(fmaxf(l01, l23), fminf(h01, h23));
                m1 = fminf(m1, 1e30f);  // tf.clip_by_value(|vc|, 0, 1e30), ms_test.py:196
                m2 = fminf(m2, 1e30f);
                // tf.sign(0) = 0 zeroes every message of the check (:187-191): some |vc| is 0 iff min1 is 0; alpha*min1 is then 0
                // by itself, and alpha*min2 (finite: min2 <= 1e30) is multiplied by the 0/1 flag -- one FSET on the busy
                // ALU pipe and one FMUL instead of a compare and two selects
                const float nzf = (m1 != 0.0f) ? 1.0f : 0.0f;
                const float a1 = __fmul_rn(a.alpha, m1);
                const float a2 = __fmul_rn(__fmul_rn(a.alpha, m2), nzf);
                unsigned sx = __float_as_uint(x[0]);
#pragma unroll
                for (int e = 1; e < 8; ++e) sx ^= __float_as_uint(x[e]);
                sx &= 0x80000000u;
                const unsigned u1 = __float_as_uint(a1) ^ sx, u2 = __float_as_uint(a2) ^ sx;
                const unsigned du = u1 - u2;
                // strict |x| > min1 select without a compare (see nms.cu): min(|x|, nextafter(min1)) is min1 or its successor
                const unsigned m1b = __float_as_uint(m1);
                const float m1p = __uint_as_float(m1b + 1u);
                const unsigned kc = u2 - m1b * du;
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const unsigned tb = __float_as_uint(fminf(ax[e], m1p));
                    unsigned u;
                    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(u) : "r"(tb), "r"(du), "r"(kc));
                    cv[R][e] = __uint_as_float(u ^ (__float_as_uint(x[e]) & 0x80000000u));
                }
            });
            // ---- variable phase: posterior = sum of incoming cv (ascending check order) + w*y (:226-227) ----
            static_for<8>([&](auto Cc) {
                constexpr int C = decltype(Cc)::value;
                constexpr int deg = var_degree(C);
                float m[deg];
                static_for<deg>([&](auto kc_) {
                    constexpr int k = decltype(kc_)::value;
                    constexpr int R = incoming(C, k) / 8, e = incoming(C, k) % 8, d = delta(R, e);
                    m[k] = d == 0 ? cv[R][e] : rot16(cv[R][e], li + 16 - d);
                });
                if constexpr (C >= 1 && C < 4) {
                    constexpr int k = twin_pos(C);
                    const float p = m[k], q = m[k + 1];
                    m[k] = first_lo[C & 3] ? p : q;
                    m[k + 1] = first_lo[C & 3] ? q : p;
                }
                float S = m[0];
#pragma unroll
                for (int k = 1; k < deg; ++k) S = __fadd_rn(S, m[k]);
                T[C] = __fadd_rn(S, wy[C]);
                if (FIR) fir[C] = __fmaf_rn(a.fir_taps[it + 1], T[C], fir[C]);
                if (TRAJ && active) a.soft_traj[(f * rows + it + 1) * N + 16 * C + ((li + sig(C)) & 15)] = T[C];
            });
        }
        // ---- hard decision (tf.where(x > 0, 0, 1), :39), syndrome (:44), outputs ----
        unsigned hbyte = 0;
#pragma unroll
        for (int C = 0; C < 8; ++C) hbyte |= (unsigned)(!(T[C] > 0.0f)) << C;
        unsigned par = 0;
        static_for<4>([&](auto Rc) {
            constexpr int R = decltype(Rc)::value;
            unsigned acc = 0;
            static_for<8>([&](auto ec) {
                constexpr int e = decltype(ec)::value;
                constexpr int C = edge(R, e).C, d = delta(R, e);
                const unsigned hv = d == 0 ? hbyte : rot16u(hbyte, li + d);
                acc ^= hv >> C;
            });
            par |= acc & 1u;
        });
        const unsigned pb = __ballot_sync(0xffffffffu, par != 0u);
        const bool nz = ((pb >> (16 * half)) & 0xffffu) != 0u;
        unsigned hw[4] = {0u, 0u, 0u, 0u};
        static_for<8>([&](auto Cc) {
            constexpr int C = decltype(Cc)::value;
            const unsigned b = (__ballot_sync(0xffffffffu, (hbyte >> C) & 1u) >> (16 * half)) & 0xffffu;
            // lane i holds the variable at offset (i + sigma_C) mod 16 of its block: rotate the field left by sigma_C
            const unsigned fld = ((b << sig(C)) | (b >> ((16 - sig(C)) & 15))) & 0xffffu;
            hw[C >> 1] |= fld << (16 * (C & 1));
        });
        if (active) {
            if (li < 4 && a.hard_bits) a.hard_bits[f * 4 + li] = li == 0 ? hw[0] : li == 1 ? hw[1] : li == 2 ? hw[2] : hw[3];
            if (li == 0) {
                if (a.iters_used) a.iters_used[f] = (uint8_t)a.iters;
                if (a.syndrome_nz) a.syndrome_nz[f] = nz ? 1 : 0;
            }
            if (FIR) {
                static_for<8>([&](auto Cc) {
                    constexpr int C = decltype(Cc)::value;
                    a.fir_out[f * N + 16 * C + ((li + sig(C)) & 15)] = fir[C] + a.fir_bias;
                });
            }
        }
        if (FUSE) {
            int pos = -1;
            if (active && li == 0) {
                if (!GEN && z.truth) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) truth_w[k] = __ldg(z.truth + row * 4 + k);
                }
                if (GEN || z.truth) {
                    const int d = __popc(hw[0] ^ truth_w[0]) + __popc(hw[1] ^ truth_w[1]) + __popc(hw[2] ^ truth_w[2]) + __popc(hw[3] ^ truth_w[3]);
                    c_frames += 1;
                    c_fe += d != 0;
                    c_be += d;
                    c_det += nz;
                    c_und += (!nz && d != 0);
                    if (!nz || !z.osd_follows) { c_ffe += d != 0; c_fbe += d; }  // frames OSD will not touch are final
                }
                if (nz && z.fail_count) {
                    pos = atomicAdd(z.fail_count, 1);
                    LDPCB_ASSERT(pos >= 0 && pos < a.B);
                    if (z.fail_idx) z.fail_idx[pos] = (int32_t)f;
                    if (GEN && z.fail_truth) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) z.fail_truth[(int64_t)pos * 4 + k] = truth_w[k];
                    }
                }
            }
            if (GEN && z.fail_llr) {
                pos = __shfl_sync(0xffffffffu, pos, 0, 16);
                if (pos >= 0) {
                    static_for<8>([&](auto Cc) {
                        constexpr int C = decltype(Cc)::value;
                        z.fail_llr[(int64_t)pos * N + 16 * C + ((li + sig(C)) & 15)] = y[C];
                    });
                }
            }
        }
    }
    if (FUSE && z.counters) {
        auto add = [&](int slot, unsigned v) {
            v = __reduce_add_sync(0xffffffffu, v);
            if (lane == 0 && v) atomicAdd(&sh_cnt[slot], (unsigned long long)v);
        };
        add(LDPCB_CNT_FRAMES, c_frames);
        add(LDPCB_CNT_NMS_FRAME_ERR, c_fe);
        add(LDPCB_CNT_NMS_BIT_ERR, c_be);
        add(LDPCB_CNT_NMS_DETECTED, c_det);
        add(LDPCB_CNT_NMS_UNDETECTED, c_und);
        add(LDPCB_CNT_NMS_ITERS, c_frames * (unsigned)a.iters);
        add(LDPCB_CNT_FINAL_FRAME_ERR, c_ffe);
        add(LDPCB_CNT_FINAL_BIT_ERR, c_fbe);
        __syncthreads();
        if (threadIdx.x < LDPCB_NUM_COUNTERS && sh_cnt[threadIdx.x]) atomicAdd(reinterpret_cast<unsigned long long*>(z.counters) + threadIdx.x, sh_cnt[threadIdx.x]);
    }
}

template <bool TRAJ, bool FIR, bool FUSE, bool GEN, int MINB>
static int launch_qc_minb(ldpcb_handle* h, const NmsArgs& a, const NmsFuse& z, cudaStream_t st) {
    auto kern = nms_qc_kernel<TRAJ, FIR, FUSE, GEN, MINB>;
    int& occ = h->occ[OCC_NMS_QC + (FUSE ? (GEN ? 4 : 3) : FIR ? 2 : TRAJ ? 1 : 0)];
    if (occ == 0) {
        LDPCB_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, QC_THREADS, 0));
        if (occ < 1) occ = 1;
    }
    const int64_t pairs = (a.B + 1) / 2;
    int64_t want = (pairs + QC_WARPS - 1) / QC_WARPS;
    int64_t cap = (int64_t)h->sm_count * occ;
    int grid = (int)(want < cap ? want : cap);
    if (grid < 1) grid = 1;
    kern<<<grid, QC_THREADS, 0, st>>>(a, z);
    LDPCB_LAUNCH_CHECK(h, "nms_qc_kernel");
    return LDPCB_OK;
}

template <bool TRAJ, bool FIR, bool FUSE, bool GEN>
static int launch_qc_variant(ldpcb_handle* h, const NmsArgs& a, const NmsFuse& z, cudaStream_t st) {
    if (h->qc_minb == 2) return launch_qc_minb<TRAJ, FIR, FUSE, GEN, 2>(h, a, z, st);
    return launch_qc_minb<TRAJ, FIR, FUSE, GEN, 3>(h, a, z, st);
}

bool nms_qc_applies(const ldpcb_handle* h, const NmsArgs& a) {
    return h->qc_ccsds && a.early_stop == 0 && a.iters >= 1 && a.w_vc == a.w_marg && !(a.soft_traj && a.fir_out);
}

// Plain / TRAJ / FIR launches (ldpcb_nms_decode, ldpcb_nms_decode_fir) and the fused pipeline launches
int launch_nms_qc(ldpcb_handle* h, const NmsArgs& a, const NmsFuse* fuse, cudaStream_t st) {
    if (a.B == 0) return LDPCB_OK;
    NmsFuse z = fuse ? *fuse : NmsFuse{};
    if (fuse) {
        if (fuse->gen) return launch_qc_variant<false, false, true, true>(h, a, z, st);
        return launch_qc_variant<false, false, true, false>(h, a, z, st);
    }
    if (a.fir_out) return launch_qc_variant<false, true, false, false>(h, a, z, st);
    if (a.soft_traj) return launch_qc_variant<true, false, false, false>(h, a, z, st);
    return launch_qc_variant<false, false, false, false>(h, a, z, st);
}

bool nms_qc_matches_code(const uint8_t* H) { return qc::matches(H); }

}  // namespace ldpcb

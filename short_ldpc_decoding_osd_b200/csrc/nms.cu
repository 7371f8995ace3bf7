// Normalized min-sum BP, flooding schedule, one warp per frame.
//
// Replaces Decoder_Layer.belief_propagation_op and its per-iteration ops
// (reference LDPC_128/Ldpc_128_testing/ms_test.py:106-137,180-242) plus the hard decision and
// syndrome of Decoding_model.get_eval (:38-44,51).  The reference materialises dense [B,64,128]
// tensors of which 512 entries per frame are live; here the 512 edge messages of a frame live in
// the registers of one warp (lane l owns checks l and l+32, 8 edges each) and cross to the
// variable side through 2.5 KB of shared memory per warp.
//
// Arithmetic is restated operation by operation so that every message matches the fp32 result of
// the reference graph: vc = total - cv (ms_test.py:133-136), sign product with sign(0) = 0
// (:184-191), min1/min2 of the clipped magnitudes and the strict '>' select (:193-206),
// cv = alpha * mag * sign (:207-209), posterior = sum(cv) + w*y with the column sum taken in
// ascending check order (:226-227).
#include "common.cuh"

namespace ldpcb {

constexpr int NMS_WARPS = 8;
constexpr int NMS_THREADS = NMS_WARPS * 32;

__device__ __forceinline__ float4 ldg_nc_f4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

// REG8: every check has degree 8 (CCSDS): the edge labelled e of check c is stored at CV[e*96+rot[e]+c], no padding slots
// FIR: accumulate fir_taps[i] * posterior_i in registers while decoding (the DIA reliability of the DL scheme, aux.cu
// dia_fir_kernel's arithmetic: one fused multiply-add per row in row order, bias added last) instead of writing the
// 13-row trajectory to HBM and reading it back
template <int DVA, int DVB, bool REG8, bool TRAJ, bool EARLY, bool FIR = false>
__global__ void __launch_bounds__(NMS_THREADS, 3) nms_kernel(NmsArgs a, const NmsTables* __restrict__ tab) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    float* T = smem + warp * NMS_FRAME_FLOATS;
    float* CV = T + NMS_CV_OFF;

    // ---- per-lane graph tables (registers) ----
    int tv[2][DC];  // T index of edge e of check lane+32q
    int cs[2][REG8 ? 1 : DC];  // CV index where that edge's message is stored (computed when REG8)
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int c = lane + 32 * q;
#pragma unroll
        for (int e = 0; e < DC; ++e) {
            const int v = tab->chk_var[c][e];
            tv[q][e] = v;
            if (!REG8) cs[q][e] = (v < N) ? (e * NMS_CV_STRIDE + tab->rot[e] + c) : NMS_CV_DUMP;
        }
    }
    int vs[4][DVA > DVB ? DVA : DVB];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int v = lane + 32 * k;
#pragma unroll
        for (int d = 0; d < (k < 2 ? DVA : DVB); ++d) vs[k][d] = tab->var_slot[v][d];
    }
    uint32_t mk[2][4];
    if (EARLY) {  // syndrome masks stay in registers only when they are needed every iteration
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int w = 0; w < 4; ++w) mk[q][w] = tab->chk_mask[lane + 32 * q][w];
    }

    int rot[DC];  // uniform: bank rotation of each edge label
#pragma unroll
    for (int e = 0; e < DC; ++e) rot[e] = tab->rot[e];
    if (lane == 0) {
        T[128] = __int_as_float(0x7f800000);  // +inf feeds padded check edges
        CV[NMS_CV_ZERO] = 0.0f;               // padded variable edges add zero
        CV[NMS_CV_DUMP] = 0.0f;
    }
    const bool same_w = (a.w_vc == a.w_marg);
    const int64_t gw = (int64_t)blockIdx.x * NMS_WARPS + warp;
    const int64_t nw = (int64_t)gridDim.x * NMS_WARPS;
    const int rows = a.iters + 1;

    for (int64_t f = gw; f < a.B; f += nw) {
        const int64_t row = a.idx ? (int64_t)a.idx[f] : f;
        __syncwarp();
        const float4 yv = ldg_nc_f4(reinterpret_cast<const float4*>(a.llr + row * N) + lane);
        reinterpret_cast<float4*>(T)[lane] = yv;
        __syncwarp();
        float y[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) y[k] = T[lane + 32 * k];
        if (TRAJ) {
#pragma unroll
            for (int k = 0; k < 4; ++k) a.soft_traj[(f * rows) * N + lane + 32 * k] = y[k];
        }
        if (a.w_vc != 1.0f) {
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 4; ++k) T[lane + 32 * k] = __fmul_rn(a.w_vc, y[k]);
            __syncwarp();
        }
        float cvo[2][DC];
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int e = 0; e < DC; ++e) cvo[q][e] = 0.0f;

        uint32_t hw[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) hw[k] = __ballot_sync(0xffffffffu, !(y[k] > 0.0f));
        int it_used = 0;
        float soft[4] = {y[0], y[1], y[2], y[3]};
        float fir[4] = {0.f, 0.f, 0.f, 0.f};
        if (FIR) {
#pragma unroll
            for (int k = 0; k < 4; ++k) fir[k] = __fmaf_rn(a.fir_taps[0], y[k], 0.0f);
        }

        for (int it = 0; it < a.iters; ++it) {
            // ---- check phase: vc = T - cv_old, min1/min2/sign, new cv ----
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                float x[DC], ax[DC];
#pragma unroll
                for (int e = 0; e < DC; ++e) {
                    x[e] = __fsub_rn(T[tv[q][e]], cvo[q][e]);
                    ax[e] = fabsf(x[e]);
                }
                // two smallest magnitudes (duplicates kept, like tf.nn.top_k) by a tournament of sorted pairs
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
                // tf.sign(0) = 0 zeroes every message of the check (ms_test.py:187-191); some |vc| is 0 iff min1 is 0
                const bool z = (m1 == 0.0f);
                const float a1 = z ? 0.0f : __fmul_rn(a.alpha, m1);
                const float a2 = z ? 0.0f : __fmul_rn(a.alpha, m2);
                unsigned sx = __float_as_uint(x[0]);
#pragma unroll
                for (int e = 1; e < DC; ++e) sx ^= __float_as_uint(x[e]);
                sx &= 0x80000000u;
                const unsigned u1 = __float_as_uint(a1) ^ sx, u2 = __float_as_uint(a2) ^ sx;
                const unsigned du = u1 - u2;
                // select u1 where |x| > m1 (strict, ms_test.py:206) else u2 without a compare: |x| >= m1 always, so
                // min(|x|, nextafter(m1)) is m1 or its successor, whose bit patterns differ by exactly one, and
                // u = u2 + (bits(min) - bits(m1)) * du is one integer multiply-add on the FMA pipe (the ALU pipe is the busy one)
                const unsigned m1b = __float_as_uint(m1);
                const float m1p = __uint_as_float(m1b + 1u);
                const unsigned kc = u2 - m1b * du;
#pragma unroll
                for (int e = 0; e < DC; ++e) {
                    const unsigned tb = __float_as_uint(fminf(ax[e], m1p));
                    unsigned u;
                    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(u) : "r"(tb), "r"(du), "r"(kc));
                    const float o = __uint_as_float(u ^ (__float_as_uint(x[e]) & 0x80000000u));
                    cvo[q][e] = o;
                    CV[REG8 ? (e * NMS_CV_STRIDE + rot[e] + lane + 32 * q) : cs[q][e]] = o;
                }
            }
            __syncwarp();
            // ---- variable phase: posterior and next totals ----
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float S = CV[vs[k][0]];
#pragma unroll
                for (int d = 1; d < (k < 2 ? DVA : DVB); ++d) S = __fadd_rn(S, CV[vs[k][d]]);
                soft[k] = __fadd_rn(S, __fmul_rn(a.w_marg, y[k]));
                T[lane + 32 * k] = same_w ? soft[k] : __fadd_rn(S, __fmul_rn(a.w_vc, y[k]));
                if (TRAJ) a.soft_traj[(f * rows + it + 1) * N + lane + 32 * k] = soft[k];
                if (FIR) fir[k] = __fmaf_rn(a.fir_taps[it + 1], soft[k], fir[k]);
                if (EARLY) hw[k] = __ballot_sync(0xffffffffu, !(soft[k] > 0.0f));
            }
            it_used = it + 1;
            __syncwarp();
            if (EARLY) {
                uint32_t par = 0;
#pragma unroll
                for (int q = 0; q < 2; ++q)
                    par |= __popc((hw[0] & mk[q][0]) ^ (hw[1] & mk[q][1]) ^ (hw[2] & mk[q][2]) ^ (hw[3] & mk[q][3])) & 1;
                if (!__any_sync(0xffffffffu, par)) break;
            }
        }
        if (TRAJ && EARLY) {
            for (int it = it_used; it < a.iters; ++it)
#pragma unroll
                for (int k = 0; k < 4; ++k) a.soft_traj[(f * rows + it + 1) * N + lane + 32 * k] = soft[k];
        }
        if (!EARLY) {
#pragma unroll
            for (int k = 0; k < 4; ++k) hw[k] = __ballot_sync(0xffffffffu, !(soft[k] > 0.0f));
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int w = 0; w < 4; ++w) mk[q][w] = __ldg(&tab->chk_mask[lane + 32 * q][w]);
        }
        uint32_t par = 0;
#pragma unroll
        for (int q = 0; q < 2; ++q)
            par |= __popc((hw[0] & mk[q][0]) ^ (hw[1] & mk[q][1]) ^ (hw[2] & mk[q][2]) ^ (hw[3] & mk[q][3])) & 1;
        const bool nz = __any_sync(0xffffffffu, par);
        if (lane < 4) {
            const uint32_t w = lane == 0 ? hw[0] : lane == 1 ? hw[1] : lane == 2 ? hw[2] : hw[3];
            a.hard_bits[f * 4 + lane] = w;
        }
        if (FIR) {
#pragma unroll
            for (int k = 0; k < 4; ++k) a.fir_out[f * N + lane + 32 * k] = fir[k] + a.fir_bias;
        }
        if (lane == 0) {
            if (a.iters_used) a.iters_used[f] = (uint8_t)it_used;
            if (a.syndrome_nz) a.syndrome_nz[f] = nz ? 1 : 0;
        }
    }
}

template <int DVA, int DVB, bool REG8, bool TRAJ, bool EARLY, bool FIR = false>
static int launch_variant(ldpcb_handle* h, const NmsArgs& a, cudaStream_t st) {
    auto kern = nms_kernel<DVA, DVB, REG8, TRAJ, EARLY, FIR>;
    const int smem = NMS_WARPS * NMS_FRAME_FLOATS * (int)sizeof(float);
    int& occ = h->occ[OCC_NMS + (TRAJ ? 1 : 0) + (EARLY ? 2 : 0) + (FIR ? 4 : 0)];
    if (occ == 0) {
        LDPCB_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NMS_THREADS, smem));
        if (occ < 1) occ = 1;
    }
    int64_t want = (a.B + NMS_WARPS - 1) / NMS_WARPS;
    int64_t cap = (int64_t)h->sm_count * occ;
    int grid = (int)(want < cap ? want : cap);
    if (grid < 1) grid = 1;
    kern<<<grid, NMS_THREADS, smem, st>>>(a, h->nms_dev);
    LDPCB_LAUNCH_CHECK(h, "nms_kernel");
    return LDPCB_OK;
}

template <int DVA, int DVB, bool REG8>
static int launch_deg(ldpcb_handle* h, const NmsArgs& a, cudaStream_t st) {
    const bool traj = a.soft_traj != nullptr, early = a.early_stop != 0;
    if (a.fir_out) return launch_variant<DVA, DVB, REG8, false, false, true>(h, a, st);
    if (traj) return early ? launch_variant<DVA, DVB, REG8, true, true>(h, a, st) : launch_variant<DVA, DVB, REG8, true, false>(h, a, st);
    return early ? launch_variant<DVA, DVB, REG8, false, true>(h, a, st) : launch_variant<DVA, DVB, REG8, false, false>(h, a, st);
}

int launch_nms(ldpcb_handle* h, const NmsArgs& a, cudaStream_t st) {
    if (a.B == 0) return LDPCB_OK;
    if (nms_qc_applies(h, a)) return launch_nms_qc(h, a, nullptr, st);  // CCSDS (128,64): half-warp-per-frame shuffle kernel (nms_qc.cu)
    bool reg8 = true;
    for (int c = 0; c < M; ++c) reg8 = reg8 && h->nms_host.chk_var[c][DC - 1] < N;
    if (reg8 && h->nms_host.max_var_deg_lo <= 5 && h->nms_host.max_var_deg_hi <= 3) return launch_deg<5, 3, true>(h, a, st);
    return launch_deg<DV, DV, false>(h, a, st);
}

}  // namespace ldpcb

using namespace ldpcb;

extern "C" int ldpcb_nms_decode(ldpcb_t* h, const float* llr_dev, int64_t B, int iters, float alpha_check,
                                float w_vc, float w_marg, int early_stop, uint32_t* hard_bits_dev,
                                uint8_t* iters_used_dev, uint8_t* syndrome_nz_dev, float* soft_traj_dev,
                                void* stream) {
    LDPCB_ENTER(h);
    if (B < 0 || iters < 0 || iters > LDPCB_MAX_ITERS) return set_error(h, LDPCB_ERR_ARG, "ldpcb_nms_decode: B=%lld iters=%d out of range", (long long)B, iters);
    if (B == 0) return LDPCB_OK;
    if (!llr_dev || !hard_bits_dev) return set_error(h, LDPCB_ERR_ARG, "ldpcb_nms_decode: NULL llr or hard_bits");
    if (((uintptr_t)llr_dev & 15) != 0) return set_error(h, LDPCB_ERR_ALIGN, "ldpcb_nms_decode: llr must be 16-byte aligned");
    NmsArgs a;
    a.llr = llr_dev; a.idx = nullptr; a.B = B; a.iters = iters;
    a.alpha = alpha_check; a.w_vc = w_vc; a.w_marg = w_marg; a.early_stop = early_stop;
    a.hard_bits = hard_bits_dev; a.iters_used = iters_used_dev; a.syndrome_nz = syndrome_nz_dev;
    a.soft_traj = soft_traj_dev;
    a.fir_out = nullptr;
    return launch_nms(h, a, (cudaStream_t)stream);
}

extern "C" int ldpcb_nms_decode_fir(ldpcb_t* h, const float* llr_dev, int64_t B, int iters, float alpha_check, float w_vc,
                                    float w_marg, const float* taps_host, float bias, uint32_t* hard_bits_dev,
                                    uint8_t* syndrome_nz_dev, float* metric_dev, void* stream) {
    LDPCB_ENTER(h);
    if (B < 0 || iters < 0 || iters > LDPCB_MAX_ITERS) return set_error(h, LDPCB_ERR_ARG, "ldpcb_nms_decode_fir: B=%lld iters=%d out of range", (long long)B, iters);
    if (B == 0) return LDPCB_OK;
    if (!llr_dev || !hard_bits_dev || !taps_host || !metric_dev) return set_error(h, LDPCB_ERR_ARG, "ldpcb_nms_decode_fir: NULL llr, hard_bits, taps or metric");
    if (((uintptr_t)llr_dev & 15) != 0) return set_error(h, LDPCB_ERR_ALIGN, "ldpcb_nms_decode_fir: llr must be 16-byte aligned");
    NmsArgs a;
    a.llr = llr_dev; a.idx = nullptr; a.B = B; a.iters = iters;
    a.alpha = alpha_check; a.w_vc = w_vc; a.w_marg = w_marg; a.early_stop = 0;
    a.hard_bits = hard_bits_dev; a.iters_used = nullptr; a.syndrome_nz = syndrome_nz_dev; a.soft_traj = nullptr;
    a.fir_out = metric_dev; a.fir_bias = bias;
    for (int i = 0; i <= LDPCB_MAX_ITERS; ++i) a.fir_taps[i] = i <= iters ? taps_host[i] : 0.0f;
    return launch_nms(h, a, (cudaStream_t)stream);
}

// PB-OSD: probability-based OSD (Yue et al. 2021 as re-implemented by the reference), one warp per frame.
//
// Replaces the per-frame body of pb_osd(snr, selected_ds), reference LDPC_128/PB_OSD/pb_testing.py:100-149:
//   best-first TEP generation optimal_tep_sequence (:366-397): pop the live TEP with the smallest MRB weight
//   (first minimum), push "append position k-1" when the last index is < k-1 and the weight is below
//   order_limit, and push "move the last index left by one";
//   promising-probability stop p_e^pro < p_t^pro (:128-132, acquire_prob_promising :431-447, beta_acquire
//   :399-405) and success-probability stop p_e^suc > p_t^suc after an improvement (:137-149, :410-423);
//   thresholds p_t^suc = 0.99*nu, p_t^pro = 0.002*sqrt((1-nu)/N_max), nu = BinCDF(order; k, mean MRB error
//   probability) (:485-500).
// TEP order and weighted distances use the exact integer reliabilities; probabilities are evaluated in fp32 in
// the order the reference evaluates them (oracle/pb_oracle.py restates it and matches the reference's per-frame
// S/F, TEP counts and improvement counters on the golden frames); binomial CDFs in fp64.
#include <cstdlib>

#include "osd_prepare.cuh"

namespace ldpcb {

// list capacity of a warp when the lists live in global memory (L2 in practice): every non-zero TEP is pushed once
__host__ __device__ constexpr int pb_glist_cap(int order) { return order >= 3 ? 43776 : (order == 2 ? 2112 : 96); }

struct __align__(16) PbHead {
    FrameSm fr;
    double cdf_p1[66];                 // BinCDF(b; 64, p1)
};

__constant__ double c_binom64[65];  // C(64, i)

__device__ __forceinline__ float sigmoid32(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }

// numpy's pairwise float32 sum of 64 values followed by /64 (np.mean(dtype=float32) in the oracle):
// 8 interleaved accumulators, then ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7))
__device__ __forceinline__ float mean64_pairwise(const float* v, int lane) {
    float r = 0.0f;
    if (lane < 8) {
        r = v[lane];
#pragma unroll
        for (int m = 1; m < 8; ++m) r = __fadd_rn(r, v[lane + 8 * m]);
    }
    r = __fadd_rn(r, __shfl_down_sync(0xffffffffu, r, 1));  // lanes 0,2,4,6: r0+r1, ...
    r = __fadd_rn(r, __shfl_down_sync(0xffffffffu, r, 2));  // lanes 0,4
    r = __fadd_rn(r, __shfl_down_sync(0xffffffffu, r, 4));  // lane 0
    return __shfl_sync(0xffffffffu, r, 0) * 0.015625f;
}

// The TEP list of each warp is a slice of pp.glist_sum / pp.glist_tep (global memory, L2-resident in practice): shared
// memory only holds the prepared frame, so 28 warps per SM are resident instead of the 4 a 25 KB list per warp allowed.
// MINB: resident CTAs per SM the register allocation is held to: 6 -> 80 registers, 8 -> 64, 10 -> 48 (60 B spilled).  Measured
// (scripts/pb_ab.py, order 2 at 3.0 dB, 1e5 failures): 1.455e7 / 1.421e7 / 1.390e7 frames/s -- more resident warps do not help,
// the 24-warp configuration stays the default (env LDPCB_PB_MINB selects the others for A/B timing).
template <int MINB>
__global__ void __launch_bounds__(OSD_THREADS, MINB) osd_pb_kernel(OsdArgs a, PbParams pp, const uint64_t* __restrict__ gcol) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    PbHead& W = reinterpret_cast<PbHead*>(smem_raw)[warp];
    const int gcap = pb_glist_cap(pp.order);
    const int64_t slice = ((int64_t)blockIdx.x * OSD_FPB + warp) * gcap;
    long long* const lsum = pp.glist_sum + slice;
    unsigned* const ltep = pp.glist_tep + slice;
    // minimum of every 32 consecutive list entries, so that a pop scans nslots/32 values instead of nslots
    long long* const bmin = pp.glist_bmin + ((int64_t)blockIdx.x * OSD_FPB + warp) * (gcap / 32);
    auto refresh_block = [&](int b, int n) {  // n = current list length; all lanes
        const int i = b * 32 + lane;
        const long long s = warp_min_ll(i < n ? lsum[i] : 0x7fffffffffffffffll);
        if (lane == 0) bmin[b] = s;
    };
    FrameSm& F = W.fr;
    const int64_t nframes = a.count ? (int64_t)*a.count : a.B;
    const int n_max = pp.order <= 0 ? 1 : (pp.order == 1 ? 65 : (pp.order == 2 ? 2081 : 43745));

    // frames are handed out one at a time from a device counter: a frame takes between a handful and N_max - 1 TEPs
    // (mean ~100-200, maximum 43,744 at order 3), so a static split leaves most warps idle behind the slowest one
    for (;;) {
        int fq = 0;
        if (lane == 0) fq = atomicAdd(pp.queue, 1);
        const int64_t f = (int64_t)__shfl_sync(0xffffffffu, fq, 0);
        if (f >= nframes) break;
        const int64_t row = a.idx ? (int64_t)a.idx[f] : f;
        __syncwarp();
        Prep P = prepare_frame<false>(a, F, gcol, row, f, lane, false, false);
        __syncwarp();
        // |y| and sigmoid(-4 nv |y|) of the permuted positions (overwrite yo/ys, dead after prepare)
        float ay[4], sg[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            ay[k] = fabsf(F.yo[P.pm[k]]);
            sg[k] = sigmoid32(__fmul_rn(pp.c4, ay[k]));
        }
        __syncwarp();
        float* absy = F.yo;
        float* sig = F.ys;
#pragma unroll
        for (int k = 0; k < 4; ++k) { absy[lane + 32 * k] = ay[k]; sig[lane + 32 * k] = sg[k]; }
        __syncwarp();
        const float p1 = mean64_pairwise(sig + K, lane);
        const float pt = mean64_pairwise(sig, lane);
        const float lrb_mean = mean64_pairwise(absy + K, lane);
        // binomial CDF table for p1, and nu = BinCDF(order; 64, pt)
        {
            const double p = (double)p1, qv = 1.0 - (double)p1;
            for (int i = lane; i <= 64; i += 32) W.cdf_p1[i] = c_binom64[i] * pow(p, (double)i) * pow(qv, (double)(64 - i));
            __syncwarp();
            if (lane == 0) {
                double acc = 0.0;
                for (int i = 0; i <= 64; ++i) { acc += W.cdf_p1[i]; W.cdf_p1[i] = fmin(acc, 1.0); }
            }
        }
        double niu = 0.0;
        {
            const double p = (double)pt, qv = 1.0 - (double)pt;
            for (int i = 0; i <= pp.order; ++i) niu += c_binom64[i] * pow(p, (double)i) * pow(qv, (double)(64 - i));
            niu = fmin(niu, 1.0);
        }
        const float p_t_suc = (float)(0.99 * niu);
        const double p_t_pro = 0.002 * sqrt((1.0 - niu) / (double)n_max);
        float spl = 1.0f;  // com_mrb_prob: sequential fp32 product (pb_testing.py:35-41)
        if (lane == 0) {
            for (int i = 0; i < K; ++i) spl = __fmul_rn(spl, __fsub_rn(1.0f, sig[i]));
        }
        spl = __shfl_sync(0xffffffffu, spl, 0);
        const double scale = __hiloint2double((1023 + P.E - 54) << 20, 0);  // 2^(E-54)
        const long long q_l0 = (long long)F.qlrb[lane], q_l1 = (long long)F.qlrb[lane + 32];
        __syncwarp();

        auto weighted = [&](unsigned long long D) -> long long {
            const long long s = (((D >> lane) & 1ull) ? q_l0 : 0ll) + (((D >> (lane + 32)) & 1ull) ? q_l1 : 0ll);
            return warp_sum_ll(s);
        };
        const unsigned long long d0 = P.d0;
        long long w_dmin = weighted(d0);
        unsigned long long opt_D = d0, opt_flip = 0ull;
        // list initialised with the TEP {k-1}
        int nslots = 1, live = 1;
        if (lane == 0) {
            lsum[0] = F.qd[K - 1];
            ltep[0] = 0xffffff00u | (unsigned)(K - 1);
            bmin[0] = F.qd[K - 1];
        }
        __syncwarp();
        int cost = 0, early = 0, suc1 = 0, suc2 = 0, list_cmp = 0;
        for (int j = 0; j < n_max - 1; ++j) {
            list_cmp += (live == 1) ? 1 : 2;
            // pop the first minimum
            long long bs = 0x7fffffffffffffffll;
            int bi = 0x7fffffff;
            {
                // the first block holding the minimum holds the first minimum
                const int nb = (nslots + 31) >> 5;
                for (int b = lane; b < nb; b += 32) {
                    const long long s = bmin[b];
                    if (s < bs) { bs = s; bi = b; }
                }
                warp_argmin(bs, bi);
                const int i = bi * 32 + lane;
                bs = i < nslots ? lsum[i] : 0x7fffffffffffffffll;
                bi = i;
                warp_argmin(bs, bi);
            }
            const unsigned tw = ltep[bi];
            const long long wsum = bs;
            // successors (uniform), written by lane 0
            unsigned pos[3];
            int w = 0;
#pragma unroll
            for (int x = 0; x < 3; ++x) {
                pos[x] = (tw >> (8 * x)) & 0xffu;
                w += pos[x] < 64u;
            }
            const unsigned last = pos[w - 1];
            int add = 0;
            long long s_ext = 0, s_adj = 0;
            unsigned t_ext = 0, t_adj = 0;
            const bool do_ext = (last < (unsigned)(K - 1)) && (w < pp.order);
            if (do_ext) {
                s_ext = wsum + F.qd[K - 1];
                t_ext = (tw & ~(0xffu << (8 * w))) | ((unsigned)(K - 1) << (8 * w));
            }
            bool do_adj;
            if (w > 1) do_adj = (last - pos[w - 2]) > 1u; else do_adj = last >= 1u;
            if (do_adj) {
                s_adj = wsum - F.qd[last] + F.qd[last - 1];
                t_adj = (tw & ~(0xffu << (8 * (w - 1)))) | ((last - 1u) << (8 * (w - 1)));
            }
            __syncwarp();
            if (lane == 0) {
                lsum[bi] = 0x7fffffffffffffffll;
                if (do_ext) { lsum[nslots] = s_ext; ltep[nslots] = t_ext; }
                if (do_adj) { lsum[nslots + (do_ext ? 1 : 0)] = s_adj; ltep[nslots + (do_ext ? 1 : 0)] = t_adj; }
            }
            add = (do_ext ? 1 : 0) + (do_adj ? 1 : 0);
            const int nslots_old = nslots;
            nslots += add;
            live += add - 1;
            __syncwarp();
            {
                const int b0 = bi >> 5, b1 = nslots_old >> 5, b2 = (nslots - 1) >> 5;
                refresh_block(b0, nslots);
                if (add > 0 && b1 != b0) refresh_block(b1, nslots);
                if (add > 0 && b2 != b1 && b2 != b0) refresh_block(b2, nslots);
                __syncwarp();
            }
            // candidate
            unsigned long long D = d0, flip = 0ull;
            float rel = 0.0f;
#pragma unroll
            for (int x = 0; x < 3; ++x)
                if (pos[x] < 64u) { D ^= F.prow[pos[x]]; flip |= 1ull << pos[x]; rel = __fadd_rn(rel, absy[pos[x]]); }
            const long long w_de = wsum + weighted(D);
            // acquire_prob_promising
            const float w1 = __fmul_rn(expf(__fmul_rn(pp.c4, rel)), spl);
            const float w2 = __fsub_rn(1.0f, w1);
            const float wdf = (float)((double)w_dmin * scale);
            const float tmp = floorf(__fdiv_rn(__fsub_rn(wdf, rel), lrb_mean));
            int beta = tmp > 0.0f ? (tmp >= 64.0f ? 64 : (int)tmp) : 0;
            const float pe = __fadd_rn(__fadd_rn(0.0f, __fmul_rn(w1, (float)W.cdf_p1[beta])), __fmul_rn(w2, (float)pp.cdf_half[beta]));
            if ((double)pe < p_t_pro) { early = 1; cost = j + 1; break; }
            ++suc1;
            if (w_de < w_dmin) {
                w_dmin = w_de;
                opt_D = D;
                opt_flip = flip;
                ++suc2;
                // acquire_p_e_suc: sequential fp32 product over the LRB
                float pes = 0.0f;
                if (lane == 0) {
                    const float ratio = __fdiv_rn(__fsub_rn(1.0f, w1), w1);
                    float prod = 1.0f;
                    for (int i = 0; i < K; ++i) {
                        const float s = sig[K + i];
                        const float fct = __fmul_rn(2.0f, ((D >> i) & 1ull) ? s : __fsub_rn(1.0f, s));
                        prod = __fmul_rn(prod, fct);
                    }
                    pes = __fdiv_rn(1.0f, __fadd_rn(1.0f, __fdiv_rn(ratio, prod)));
                }
                pes = __shfl_sync(0xffffffffu, pes, 0);
                if (pes > p_t_suc) { early = 1; cost = j + 1; break; }
            }
        }
        const int num = early ? cost : n_max;
        // outputs
        const unsigned long long c_lrb = opt_D ^ P.hd_lrb;
        const unsigned long long c_mrb = P.ho_mrb ^ opt_flip;
        __syncwarp();
        F.pos[P.pm[0]] = (unsigned char)((c_mrb >> lane) & 1ull);
        F.pos[P.pm[1]] = (unsigned char)((c_mrb >> (lane + 32)) & 1ull);
        F.pos[P.pm[2]] = (unsigned char)((c_lrb >> lane) & 1ull);
        F.pos[P.pm[3]] = (unsigned char)((c_lrb >> (lane + 32)) & 1ull);
        __syncwarp();
        unsigned wout[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) wout[k] = __ballot_sync(0xffffffffu, F.pos[lane + 32 * k]);
        const int64_t orow = a.idx ? row : f;
        if (lane < 4 && a.cw_bits) {
            const unsigned wv = lane == 0 ? wout[0] : lane == 1 ? wout[1] : lane == 2 ? wout[2] : wout[3];
            a.cw_bits[orow * 4 + lane] = wv;
        }
        if (lane == 0) {
            if (a.best_score_q) a.best_score_q[orow] = w_dmin;
            if (a.score_exp) a.score_exp[f] = P.E;
            if (pp.stats) {
                int32_t* st = pp.stats + orow * 4;
                st[0] = num;
                st[1] = suc1;
                st[2] = suc2;
                st[3] = list_cmp;
            }
        }
    }
}

constexpr int PB_DEFAULT_MINB = 6;

template <int MINB>
static int launch_pb_minb(ldpcb_handle* h, const OsdArgs& a, const PbParams& pp, bool cap_order3, cudaStream_t st) {
    // lists in global memory: one slice per resident warp
    const int smem = OSD_FPB * (int)sizeof(PbHead);
    int& occ = h->occ_pb[MINB == 6 ? 0 : MINB == 8 ? 1 : 2];
    if (occ == 0) {
        LDPCB_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, osd_pb_kernel<MINB>, OSD_THREADS, smem));
        if (occ < 1) occ = 1;
    }
    int64_t want = (a.B + OSD_FPB - 1) / OSD_FPB;
    // order 3: a warp's list slice is 43,776 entries (0.5 MB), 1.9 GB of lists at 24 warps per SM; capping the grid at 16 warps
    // per SM to keep them smaller measured 3-5 % slower, so the cap is opt-in (env LDPCB_PB_CAP3)
    int64_t cap = (int64_t)h->sm_count * (pp.order >= 3 && cap_order3 && occ > 4 ? 4 : occ);
    int grid = (int)(want < cap ? want : cap);
    if (grid < 1) grid = 1;
    const size_t per = (size_t)grid * OSD_FPB * pb_glist_cap(pp.order);
    if (h->pb_list_cap < per) {
        if (h->pb_list) { LDPCB_CUDA(h, cudaDeviceSynchronize()); LDPCB_CUDA(h, cudaFree(h->pb_list)); h->pb_list = nullptr; h->pb_list_cap = 0; }
        LDPCB_CUDA(h, cudaMalloc(&h->pb_list, per * (sizeof(long long) + sizeof(unsigned)) + per / 32 * sizeof(long long)));
        h->pb_list_cap = per;
    }
    PbParams q = pp;
    q.queue = h->pb_queue;
    LDPCB_CUDA(h, cudaMemsetAsync(h->pb_queue, 0, sizeof(int), st));
    q.glist_sum = reinterpret_cast<long long*>(h->pb_list);
    q.glist_bmin = reinterpret_cast<long long*>(h->pb_list + per * sizeof(long long));
    q.glist_tep = reinterpret_cast<unsigned*>(h->pb_list + per * sizeof(long long) + per / 32 * sizeof(long long));
    osd_pb_kernel<MINB><<<grid, OSD_THREADS, smem, st>>>(a, q, h->gcol_dev);
    LDPCB_LAUNCH_CHECK(h, "osd_pb_kernel");
    return LDPCB_OK;
}

int launch_osd_pb(ldpcb_handle* h, const OsdArgs& a, const PbParams& pp, cudaStream_t st) {
    if (a.B == 0) return LDPCB_OK;
    if (!h->pb_consts_ready) {
        double c[65];
        unsigned __int128 e = 1;  // exact integers, rounded once to fp64 like Python's float(math.comb(64, i))
        c[0] = 1.0;
        for (int i = 1; i <= 64; ++i) {
            e = e * (unsigned)(64 - i + 1) / (unsigned)i;
            c[i] = (double)(unsigned long long)e;
        }
        LDPCB_CUDA(h, cudaMemcpyToSymbol(c_binom64, c, sizeof c));
        h->pb_consts_ready = true;
    }
    const char* e = getenv("LDPCB_PB_MINB");  // A/B timing of the register budget
    const int minb = e ? atoi(e) : PB_DEFAULT_MINB;
    const bool cap3 = getenv("LDPCB_PB_CAP3") != nullptr;
    if (minb == 6) return launch_pb_minb<6>(h, a, pp, cap3, st);
    if (minb == 10) return launch_pb_minb<10>(h, a, pp, cap3, st);
    return launch_pb_minb<8>(h, a, pp, cap3, st);
}

}  // namespace ldpcb

using namespace ldpcb;

extern "C" int ldpcb_osd_pb_decode(ldpcb_t* h, const float* llr_dev, int64_t B, int order_limit, float snr_db,
                                   uint32_t* cw_bits_dev, int32_t* stats_dev, int64_t* best_score_q_dev,
                                   int32_t* score_exp_dev, void* stream) {
    LDPCB_ENTER(h);
    if (B < 0 || order_limit < 0 || order_limit > 3)
        return set_error(h, LDPCB_ERR_ARG, "ldpcb_osd_pb_decode: B=%lld order_limit=%d out of range (0..3)", (long long)B, order_limit);
    if (B == 0) return LDPCB_OK;
    if (!llr_dev || !cw_bits_dev) return set_error(h, LDPCB_ERR_ARG, "ldpcb_osd_pb_decode: NULL llr or cw_bits");
    if ((uintptr_t)llr_dev & 15) return set_error(h, LDPCB_ERR_ALIGN, "ldpcb_osd_pb_decode: llr must be 16-byte aligned");
    OsdArgs a = {};
    a.order_llr = llr_dev; a.score_llr = llr_dev; a.B = B;
    a.cw_bits = cw_bits_dev; a.best_score_q = best_score_q_dev; a.score_exp = score_exp_dev;
    PbParams pp;
    const double nv = 1.0 / pow(10.0, (double)snr_db / 10.0);
    pp.c4 = (float)(-4.0 * nv);
    pp.order = order_limit;
    pp.stats = stats_dev;
    pp.glist_sum = nullptr;
    pp.glist_bmin = nullptr;
    pp.glist_tep = nullptr;
    {   // BinCDF(b; 64, 1/2) in fp64, cumulative in index order
        unsigned __int128 e = 1;
        double acc = 0.0;
        const double half64 = ldexp(1.0, -64);
        for (int i = 0; i <= 64; ++i) {
            if (i > 0) e = e * (unsigned)(64 - i + 1) / (unsigned)i;
            const double c = (double)(unsigned long long)e;
            acc += c * half64;
            pp.cdf_half[i] = acc < 1.0 ? acc : 1.0;
        }
    }
    return launch_osd_pb(h, a, pp, (cudaStream_t)stream);
}

extern "C" int ldpcb_osd_pb_decode_host(ldpcb_t* h, const float* llr_host, int64_t B, int order_limit, float snr_db,
                                        uint32_t* cw_bits_host, int32_t* stats_host) {
    LDPCB_ENTER(h);
    if (B < 0 || !llr_host || !cw_bits_host) return set_error(h, LDPCB_ERR_ARG, "ldpcb_osd_pb_decode_host: bad arguments");
    if (B == 0) return LDPCB_OK;
    cudaStream_t st = h->streams[0];
    const int64_t chunk = 1 << 16;
    const size_t need = 256 * 4 + (size_t)chunk * (N * sizeof(float) + 16 + 16);
    int s = ensure_ws(h, 1, need);
    if (s != LDPCB_OK) return s;
    char* base = h->ws[1].buf;
    float* llr = reinterpret_cast<float*>(base);
    uint32_t* bits = reinterpret_cast<uint32_t*>(base + (size_t)chunk * N * sizeof(float));
    int32_t* stats = reinterpret_cast<int32_t*>(base + (size_t)chunk * (N * sizeof(float) + 16));
    for (int64_t b0 = 0; b0 < B; b0 += chunk) {
        const int64_t nb = (B - b0) < chunk ? (B - b0) : chunk;
        LDPCB_CUDA(h, cudaMemcpyAsync(llr, llr_host + b0 * N, sizeof(float) * nb * N, cudaMemcpyHostToDevice, st));
        s = ldpcb_osd_pb_decode(h, llr, nb, order_limit, snr_db, bits, stats, nullptr, nullptr, st);
        if (s != LDPCB_OK) return s;
        LDPCB_CUDA(h, cudaMemcpyAsync(cw_bits_host + b0 * 4, bits, sizeof(uint32_t) * nb * 4, cudaMemcpyDeviceToHost, st));
        if (stats_host) LDPCB_CUDA(h, cudaMemcpyAsync(stats_host + b0 * 4, stats, sizeof(int32_t) * nb * 4, cudaMemcpyDeviceToHost, st));
        LDPCB_CUDA(h, cudaStreamSynchronize(st));
    }
    return LDPCB_OK;
}

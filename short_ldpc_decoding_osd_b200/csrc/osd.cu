// Ordered-statistics decoding, one warp per frame, everything on chip between the LLR load and the
// 16-byte codeword store.
//
// Replaces, per frame (reference paths relative to LDPC_128/):
//   swapped_info           PB_OSD/pb_testing.py:306-320  (reliability sort, pi1)
//   identify_mrb           PB_OSD/pb_testing.py:268-304  (pi2, systematic reduced_G = [I | P'])
//   full_gf2elim           PB_OSD/pb_testing.py:231-266  (GF(2) Gauss-Jordan)
//   convention_osd_main    FS_OSD/convention_osd.py:49-77 (TEP sweep, re-encode, discrepancy, argmin)
//   osd.acquire_min        DL_OSD_Testing_serial/ordered_statistics_decoding.py:153-162 (block minima)
//
// Phases of one frame:
//   1. rank sort of the 128 keys (|y| bits, index) -> pi1; ties as tf.argsort (stable)
//   2. column-major GF(2) elimination of G[:, pi1]: lane l holds sorted columns l, l+32, l+64, l+96
//      as 64-bit words (bit r = row r); columns are scanned most reliable first, a column with a 1
//      in a row not yet used becomes the next pivot (greedy most-reliable basis).  The reference's
//      rule (row swap / column swap with the first 1 of row i) selects the same basis and its
//      outputs depend only on that basis because identify_mrb re-sorts both halves
//      (pb_testing.py:284-300); DESIGN.md gives the argument, tests compare with the lifted reference.
//   3. P' rows by ballot transposition of the 64 non-pivot columns
//   4. exact integer reliabilities q = rint(|y| * 2^(54-E)), byte LUTs of the 64 LRB weights
//   5. TEP sweep: D = d0 ^ XOR_{t in TEP} P'_t, score = base + sum delta_t + sum_b LUT_b[byte_b(D)]
//   6. lexicographic (score, index) minimum = first minimum in enumeration order (tf.argmin)
#include "common.cuh"

namespace ldpcb {

constexpr int OSD_WARPS = 2;  // warps (frames in flight) per CTA
constexpr int OSD_THREADS = OSD_WARPS * 32;

// per-warp shared memory (bytes)
struct __align__(16) OsdSmem {
    unsigned long long lut[8][256];  // 16 KB: weighted-popcount tables of the LRB (aliased by `cols` earlier)
    unsigned long long prow[68];     // P' rows by MRB position (logical), [64] = 0 for padded TEP slots
    long long qd[68];                // signed score delta of flipping MRB position t, [64] = 0
    unsigned long long qlrb[64];     // q of LRB positions
    float yo[N];                     // ordering metric
    float ys[N];                     // scoring metric
    unsigned int key[N];             // |yo| bits
    unsigned char pi1[N];            // sorted position -> original index
    unsigned char perm[N];           // permuted position (MRB then LRB) -> original index
    unsigned char pos[N];            // permuted position -> sorted position
    unsigned char prow_of[64];       // pivot row of MRB position t
    unsigned char tmp[N];
};

__device__ __forceinline__ unsigned long long shfl64(unsigned long long v, int src) {
    unsigned lo = __shfl_sync(0xffffffffu, (unsigned)v, src);
    unsigned hi = __shfl_sync(0xffffffffu, (unsigned)(v >> 32), src);
    return ((unsigned long long)hi << 32) | lo;
}
__device__ __forceinline__ unsigned long long shfl_xor64(unsigned long long v, int m) {
    unsigned lo = __shfl_xor_sync(0xffffffffu, (unsigned)v, m);
    unsigned hi = __shfl_xor_sync(0xffffffffu, (unsigned)(v >> 32), m);
    return ((unsigned long long)hi << 32) | lo;
}
__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
    for (int m = 16; m; m >>= 1) v += (long long)shfl_xor64((unsigned long long)v, m);
    return v;
}
__device__ __forceinline__ unsigned long long warp_xor_ull(unsigned long long v) {
#pragma unroll
    for (int m = 16; m; m >>= 1) v ^= shfl_xor64(v, m);
    return v;
}

// exact integer reliability: q = rint(a * 2^(54-E)); a finite >= 0, E = frexp exponent of the frame max
__device__ __forceinline__ long long quantize(float a, int E) {
    const double scale = __hiloint2double((1023 + 54 - E) << 20, 0);
    return __double2ll_rn((double)a * scale);
}

// |y| used for scoring: NaN -> 0, inf -> FLT_MAX
__device__ __forceinline__ float score_abs(float y) {
    float a = fabsf(y);
    if (!(a == a)) a = 0.0f;
    return fminf(a, 3.402823466e38f);
}

template <int MAXW, bool BLOCKS>
__global__ void __launch_bounds__(OSD_THREADS) osd_kernel(OsdArgs a, const uint64_t* __restrict__ gcol) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    OsdSmem& S = reinterpret_cast<OsdSmem*>(smem_raw)[warp];
    unsigned long long* cols = &S.lut[0][0];  // [128] columns after elimination (before the LUT is built)

    const int64_t nframes = a.count ? (int64_t)*a.count : a.B;
    const int64_t gw = (int64_t)blockIdx.x * OSD_WARPS + warp;
    const int64_t nw = (int64_t)gridDim.x * OSD_WARPS;
    const bool ties_high = (a.flags & LDPCB_OSD_TIES_HIGH_INDEX_FIRST) != 0;
    const bool disc_from_score = (a.flags & LDPCB_OSD_DISC_HARD_FROM_SCORE) != 0;

    for (int64_t f = gw; f < nframes; f += nw) {
        const int64_t row = a.idx ? (int64_t)a.idx[f] : f;
        __syncwarp();
        // ---- load -----------------------------------------------------------------------------
        {
            const float4 v = reinterpret_cast<const float4*>(a.order_llr + row * N)[lane];
            reinterpret_cast<float4*>(S.yo)[lane] = v;
            const float4 w = reinterpret_cast<const float4*>(a.score_llr + row * N)[lane];
            reinterpret_cast<float4*>(S.ys)[lane] = w;
            uint4 kb;
            kb.x = __float_as_uint(v.x) & 0x7fffffffu;
            kb.y = __float_as_uint(v.y) & 0x7fffffffu;
            kb.z = __float_as_uint(v.z) & 0x7fffffffu;
            kb.w = __float_as_uint(v.w) & 0x7fffffffu;
            reinterpret_cast<uint4*>(S.key)[lane] = kb;
            if (lane == 0) { S.prow[64] = 0ull; S.qd[64] = 0ll; }
        }
        __syncwarp();
        // ---- 1. rank sort, descending |y|, stable (ties: lower index first unless ties_high) ----
        {
            unsigned mykey[4];
            int rank[4] = {0, 0, 0, 0};
#pragma unroll
            for (int k = 0; k < 4; ++k) mykey[k] = S.key[lane + 32 * k];
            for (int i4 = 0; i4 < N / 4; ++i4) {
                const uint4 kk = reinterpret_cast<const uint4*>(S.key)[i4];
                const unsigned ki[4] = {kk.x, kk.y, kk.z, kk.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = 4 * i4 + u;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int j = lane + 32 * k;
                        const bool first = ties_high ? (i > j) : (i < j);
                        rank[k] += (ki[u] > mykey[k]) || (ki[u] == mykey[k] && first);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) S.pi1[rank[k]] = (unsigned char)(lane + 32 * k);
        }
        __syncwarp();
        // ---- 2. GF(2) elimination, column-major ------------------------------------------------
        unsigned long long col[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) col[k] = gcol[S.pi1[lane + 32 * k]];
        {
            unsigned long long used = 0ull;
            int npiv = 0, nlrb = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                for (int l = 0; l < 32; ++l) {
                    const int c = 32 * k + l;
                    if (npiv == K) {  // basis complete: everything left is LRB
                        if (lane == 0) S.pos[K + nlrb] = (unsigned char)c;
                        ++nlrb;
                        continue;
                    }
                    const unsigned long long cc = shfl64(col[k], l);
                    const unsigned long long cand = cc & ~used;
                    if (cand == 0ull) {  // dependent on more reliable columns
                        if (lane == 0) S.pos[K + nlrb] = (unsigned char)c;
                        ++nlrb;
                        continue;
                    }
                    const int p = __ffsll((long long)cand) - 1;
                    used |= 1ull << p;
                    if (lane == 0) { S.pos[npiv] = (unsigned char)c; S.prow_of[npiv] = (unsigned char)p; }
                    ++npiv;
                    const unsigned long long m = cc ^ (1ull << p);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        if ((col[kk] >> p) & 1ull) col[kk] ^= m;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) cols[lane + 32 * k] = col[k];
        __syncwarp();
        // ---- permutation pi2 o pi1 and the permuted metrics ------------------------------------
        float yo[4], ys[4];
        unsigned char pm[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int t = lane + 32 * k;
            pm[k] = S.pi1[S.pos[t]];
            yo[k] = S.yo[pm[k]];
            ys[k] = S.ys[pm[k]];
        }
        // ---- 3. P' rows: transpose the 64 LRB columns by ballots --------------------------------
        unsigned long long myprow[2] = {0ull, 0ull};
        {
            const unsigned long long ca = cols[S.pos[K + lane]];
            const unsigned long long cb = cols[S.pos[K + 32 + lane]];
            for (int t = 0; t < K; ++t) {
                const int p = S.prow_of[t];
                const unsigned lo = __ballot_sync(0xffffffffu, (ca >> p) & 1ull);
                const unsigned hi = __ballot_sync(0xffffffffu, (cb >> p) & 1ull);
                if ((t & 31) == lane) myprow[t >> 5] = ((unsigned long long)hi << 32) | lo;
            }
        }
        __syncwarp();  // everyone is done reading cols (aliases lut)
        S.prow[lane] = myprow[0];
        S.prow[lane + 32] = myprow[1];
#pragma unroll
        for (int k = 0; k < 4; ++k) S.perm[lane + 32 * k] = pm[k];
        // ---- 4. exact reliabilities ------------------------------------------------------------
        float as[4];
        unsigned amax_bits = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            as[k] = score_abs(ys[k]);
            amax_bits = max(amax_bits, __float_as_uint(as[k]));
        }
#pragma unroll
        for (int m = 16; m; m >>= 1) amax_bits = max(amax_bits, __shfl_xor_sync(0xffffffffu, amax_bits, m));
        int E = 0;
        frexpf(__uint_as_float(amax_bits), &E);
        long long q[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) q[k] = quantize(as[k], E);
        // hard decisions: 1 iff !(y > 0)   (convention_osd.py:54)
        unsigned ho[4], hd[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            ho[k] = !(yo[k] > 0.0f);
            hd[k] = disc_from_score ? (unsigned)!(ys[k] > 0.0f) : ho[k];
        }
        // MRB: delta of flipping position t, and the base discrepancy of the order-0 MRB part
        long long base = 0;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const unsigned d0m = ho[k] ^ hd[k];
            S.qd[lane + 32 * k] = d0m ? -q[k] : q[k];
            base += d0m ? q[k] : 0ll;
        }
        base = warp_sum_ll(base);
        S.qlrb[lane] = (unsigned long long)q[2];
        S.qlrb[lane + 32] = (unsigned long long)q[3];
        // order-0 codeword: c0_lrb = XOR of P' rows of the MRB positions whose hard decision is 1
        unsigned long long c0 = (ho[0] ? myprow[0] : 0ull) ^ (ho[1] ? myprow[1] : 0ull);
        c0 = warp_xor_ull(c0);
        const unsigned long long hd_lrb =
            (unsigned long long)__ballot_sync(0xffffffffu, hd[2]) | ((unsigned long long)__ballot_sync(0xffffffffu, hd[3]) << 32);
        const unsigned long long ho_mrb =
            (unsigned long long)__ballot_sync(0xffffffffu, ho[0]) | ((unsigned long long)__ballot_sync(0xffffffffu, ho[1]) << 32);
        const unsigned long long d0 = c0 ^ hd_lrb;
        __syncwarp();
        // ---- byte LUTs: lut[b][x] = sum of q_lrb[8b+i] over the set bits i of x -------------------
        {
#pragma unroll 1
            for (int b = 0; b < 8; ++b) {
                unsigned long long w[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) w[i] = S.qlrb[8 * b + i];
                unsigned long long hsum = 0ull;
#pragma unroll
                for (int i = 0; i < 5; ++i) hsum += ((lane >> i) & 1) ? w[3 + i] : 0ull;
                unsigned long long e[8];
                e[0] = hsum;
                e[1] = hsum + w[0];
                e[2] = hsum + w[1];
                e[3] = e[2] + w[0];
                e[4] = hsum + w[2];
                e[5] = e[4] + w[0];
                e[6] = e[4] + w[1];
                e[7] = e[6] + w[0];
                ulonglong2* dst = reinterpret_cast<ulonglong2*>(&S.lut[b][lane * 8]);
#pragma unroll
                for (int i = 0; i < 4; ++i) dst[i] = make_ulonglong2(e[2 * i], e[2 * i + 1]);
            }
        }
        __syncwarp();
        // ---- truth score (DL success test) -------------------------------------------------------
        if (BLOCKS && a.truth_bits && a.truth_score_q) {
            long long ts = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const unsigned tb = (a.truth_bits[row * 4 + (pm[k] >> 5)] >> (pm[k] & 31)) & 1u;
                ts += (tb ^ hd[k]) ? q[k] : 0ll;
            }
            ts = warp_sum_ll(ts);
            if (lane == 0) a.truth_score_q[f] = ts;
        }
        // ---- 5./6. TEP sweep ---------------------------------------------------------------------
        const int nblk = BLOCKS ? a.n_blocks : 1;
        long long best_s = 0x7fffffffffffffffll;
        int best_i = 0x7fffffff;
        for (int blk = 0; blk < nblk; ++blk) {
            const int i0 = BLOCKS ? a.block_start[blk] : 0;
            const int i1 = BLOCKS ? a.block_start[blk + 1] : a.n_teps;
            long long bs = 0x7fffffffffffffffll;
            int bi = 0x7fffffff;
            for (int i = i0 + lane; i < i1; i += 32) {
                const unsigned w = __ldg(a.teps + i);
                unsigned long long D = d0;
                long long s = base;
#pragma unroll
                for (int j = 0; j < MAXW; ++j) {
                    const unsigned t = min((w >> (8 * j)) & 0xffu, 64u);
                    D ^= S.prow[t];
                    s += S.qd[t];
                }
#pragma unroll
                for (int b = 0; b < 8; ++b) s += (long long)S.lut[b][(unsigned)(D >> (8 * b)) & 0xffu];
                if (s < bs) { bs = s; bi = i; }
            }
            // warp argmin, lexicographic (score, index)
#pragma unroll
            for (int m = 16; m; m >>= 1) {
                const long long os = (long long)shfl_xor64((unsigned long long)bs, m);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, m);
                if (os < bs || (os == bs && oi < bi)) { bs = os; bi = oi; }
            }
            if (BLOCKS) {
                if (lane == 0) {
                    a.block_min_q[f * nblk + blk] = bs;
                    if (a.block_arg) a.block_arg[f * nblk + blk] = bi;
                }
            }
            if (bs < best_s || (bs == best_s && bi < best_i)) { best_s = bs; best_i = bi; }
        }
        // ---- outputs -----------------------------------------------------------------------------
        if (!BLOCKS || a.cw_bits) {
            // re-encode the winner and un-permute
            unsigned long long D = d0, flip = 0ull;
            if (best_i != 0x7fffffff) {
                const unsigned w = __ldg(a.teps + best_i);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const unsigned t = (w >> (8 * j)) & 0xffu;
                    if (t < 64u) { D ^= S.prow[t]; flip ^= 1ull << t; }
                }
            }
            const unsigned long long c_lrb = D ^ hd_lrb;
            const unsigned long long c_mrb = ho_mrb ^ flip;
            S.tmp[pm[0]] = (unsigned char)((c_mrb >> lane) & 1ull);
            S.tmp[pm[1]] = (unsigned char)((c_mrb >> (lane + 32)) & 1ull);
            S.tmp[pm[2]] = (unsigned char)((c_lrb >> lane) & 1ull);
            S.tmp[pm[3]] = (unsigned char)((c_lrb >> (lane + 32)) & 1ull);
            __syncwarp();
            unsigned wout[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) wout[k] = __ballot_sync(0xffffffffu, S.tmp[lane + 32 * k]);
            const int64_t orow = a.idx ? row : f;
            if (lane < 4 && a.cw_bits) {
                const unsigned wv = lane == 0 ? wout[0] : lane == 1 ? wout[1] : lane == 2 ? wout[2] : wout[3];
                a.cw_bits[orow * 4 + lane] = wv;
            }
            if (lane == 0) {
                if (a.best_tep) a.best_tep[orow] = best_i;
                if (a.best_score_q) a.best_score_q[orow] = best_s;
            }
        }
        if (lane == 0 && a.score_exp) a.score_exp[f] = E;
        if (a.perm) {
#pragma unroll
            for (int k = 0; k < 4; ++k) a.perm[f * N + lane + 32 * k] = pm[k];
        }
        if (a.redG) {
            a.redG[f * K + lane] = myprow[0];
            a.redG[f * K + lane + 32] = myprow[1];
        }
    }
}

template <int MAXW, bool BLOCKS>
static int launch_variant(ldpcb_handle* h, const OsdArgs& a, cudaStream_t st) {
    auto kern = osd_kernel<MAXW, BLOCKS>;
    const int smem = OSD_WARPS * (int)sizeof(OsdSmem);
    static thread_local int occ_cache[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int& occ = occ_cache[h->device & 7];
    if (occ == 0) {
        LDPCB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        LDPCB_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, OSD_THREADS, smem));
        if (occ < 1) occ = 1;
    }
    int64_t want = (a.B + OSD_WARPS - 1) / OSD_WARPS;
    int64_t cap = (int64_t)h->sm_count * occ;
    int grid = (int)(want < cap ? want : cap);
    if (grid < 1) grid = 1;
    kern<<<grid, OSD_THREADS, smem, st>>>(a, h->gcol_dev);
    LDPCB_LAUNCH_CHECK(h, "osd_kernel");
    return LDPCB_OK;
}

int launch_osd(ldpcb_handle* h, const OsdArgs& a, cudaStream_t st) {
    if (a.B == 0) return LDPCB_OK;
    const bool blocks = a.block_start != nullptr;
    switch (a.maxw) {
        case 1: return blocks ? launch_variant<1, true>(h, a, st) : launch_variant<1, false>(h, a, st);
        case 2: return blocks ? launch_variant<2, true>(h, a, st) : launch_variant<2, false>(h, a, st);
        case 3: return blocks ? launch_variant<3, true>(h, a, st) : launch_variant<3, false>(h, a, st);
        default: return blocks ? launch_variant<4, true>(h, a, st) : launch_variant<4, false>(h, a, st);
    }
}

}  // namespace ldpcb

using namespace ldpcb;

static int check_llr(ldpcb_handle* h, const char* fn, const float* a, const float* b) {
    if (!a || !b) return set_error(h, LDPCB_ERR_ARG, "%s: NULL llr", fn);
    if ((((uintptr_t)a) | ((uintptr_t)b)) & 15) return set_error(h, LDPCB_ERR_ALIGN, "%s: llr must be 16-byte aligned", fn);
    return LDPCB_OK;
}

extern "C" int ldpcb_osd_decode(ldpcb_t* h, const float* order_llr_dev, const float* score_llr_dev, int64_t B,
                                int order, int tep_order, int flags, uint32_t* cw_bits_dev, int32_t* best_tep_dev,
                                int64_t* best_score_q_dev, int32_t* score_exp_dev, uint8_t* perm_dev,
                                uint64_t* redG_dev, void* stream) {
    if (!h) return LDPCB_ERR_ARG;
    if (B < 0 || order < 0 || order > 3 || tep_order < 0 || tep_order > 1 || (flags & ~3))
        return set_error(h, LDPCB_ERR_ARG, "ldpcb_osd_decode: B=%lld order=%d tep_order=%d flags=%d out of range", (long long)B, order, tep_order, flags);
    if (B == 0) return LDPCB_OK;
    int st = check_llr(h, "ldpcb_osd_decode", order_llr_dev, score_llr_dev);
    if (st != LDPCB_OK) return st;
    if (!cw_bits_dev) return set_error(h, LDPCB_ERR_ARG, "ldpcb_osd_decode: NULL cw_bits");
    const TepTable& t = h->tep[order][tep_order];
    OsdArgs a = {};
    a.order_llr = order_llr_dev; a.score_llr = score_llr_dev; a.B = B;
    a.teps = t.dev; a.n_teps = t.n; a.maxw = t.maxw; a.flags = flags;
    a.cw_bits = cw_bits_dev; a.best_tep = best_tep_dev; a.best_score_q = best_score_q_dev;
    a.score_exp = score_exp_dev; a.perm = perm_dev; a.redG = redG_dev;
    return launch_osd(h, a, (cudaStream_t)stream);
}

extern "C" int ldpcb_osd_block_minima(ldpcb_t* h, const float* order_llr_dev, const float* score_llr_dev, int64_t B,
                                      const uint32_t* teps_dev, int32_t n_teps, const int32_t* block_start_dev,
                                      int32_t n_blocks, int flags, int64_t* block_min_q_dev, int32_t* block_arg_dev,
                                      int32_t* score_exp_dev, const uint32_t* truth_bits_dev,
                                      int64_t* truth_score_q_dev, uint8_t* perm_dev, void* stream) {
    if (!h) return LDPCB_ERR_ARG;
    if (B < 0 || n_teps < 0 || n_blocks < 1 || (flags & ~3))
        return set_error(h, LDPCB_ERR_ARG, "ldpcb_osd_block_minima: B=%lld n_teps=%d n_blocks=%d flags=%d out of range", (long long)B, n_teps, n_blocks, flags);
    if (B == 0) return LDPCB_OK;
    int st = check_llr(h, "ldpcb_osd_block_minima", order_llr_dev, score_llr_dev);
    if (st != LDPCB_OK) return st;
    if (!teps_dev || !block_start_dev || !block_min_q_dev)
        return set_error(h, LDPCB_ERR_ARG, "ldpcb_osd_block_minima: NULL teps, block_start or block_min_q");
    OsdArgs a = {};
    a.order_llr = order_llr_dev; a.score_llr = score_llr_dev; a.B = B;
    a.teps = teps_dev; a.n_teps = n_teps; a.maxw = 4; a.flags = flags;
    a.block_start = block_start_dev; a.n_blocks = n_blocks;
    a.block_min_q = block_min_q_dev; a.block_arg = block_arg_dev; a.score_exp = score_exp_dev;
    a.truth_bits = truth_bits_dev; a.truth_score_q = truth_score_q_dev; a.perm = perm_dev;
    return launch_osd(h, a, (cudaStream_t)stream);
}

// Order-2 OSD on the full TEP lists (conventional or FS order, 2081 TEPs): every warp prepares AND sweeps its own frame.
//
// The sweep is the tensor-core pair sweep described in osd_sweep.cuh / DESIGN.md 4.2 (score of the pair TEP {i,j} =
// R_i + C_j - 2 M[i][j], M a 64x64x64 u8 matrix product on IMMA.16832.U8.U8), organised per warp: the 20 16x8 tiles
// that hold a pair i < j are visited in turn, four A-fragment builds per frame, the eight B fragments built once into
// shared memory (the space of the fallback LUT).  An element's accumulator starts at -(R_i + C_j) and ends at -S.  A CTA of
// four warps shares nothing but the 16 KB byte LUT of the rare exact fallback (a frame whose truncated scores leave
// too many candidates, e.g. quantised inputs), so a round costs two CTA barriers instead of thirteen.
// Replaces the same reference code as osd.cu (swapped_info / identify_mrb / full_gf2elim / convention_osd_main).
#include "common.cuh"
#include "osd_prepare.cuh"
#include "osd_sweep.cuh"

namespace ldpcb {

struct __align__(16) PairSmem {
    union {
        unsigned long long lut[8][256];  // exact fallback (after barrier (A): the sweeps are over)
        uint4 bfrag[OSD_FPB][8][32];     // sweep: [warp][column block nj][lane] B fragment registers, built once per frame
    };
    FrameSm fr[OSD_FPB];
    long long red_s[OSD_FPB];        // fallback: per-warp partial minima of the frame in turn
    int red_i[OSD_FPB];
    int fb[OSD_FPB];                 // frame of warp w needs the exact fallback
};

constexpr int PW_CODE_SINGLE = 80, PW_CODE_EMPTY = 82;  // codes 0..79: (tile << 2) | element
constexpr int PW_PEN = 1 << 29;  // an excluded element's accumulator (-S) starts this much lower: it never reaches a gate (S < 2^23)
// packed scores (S << 7 | code) that can still be inside the truncation window when the minimum is m
__device__ __forceinline__ int pw_gate(int m) { return m > 0x7fffffff - ((OSD_WIN + 1) << 7) ? 0x7fffffff : ((((m >> 7) + OSD_WIN) << 7) | 127); }
// the tiles hold -S: packed score <= gate  <=>  S <= gate >> 7  <=>  -S >= -(gate >> 7)   (the low seven bits of a gate are ones)
__device__ __forceinline__ int pw_gate_neg(int gate) { return -(gate >> 7); }

__global__ void __launch_bounds__(OSD_THREADS, 7) osd_pair_kernel(OsdArgs a, const uint64_t* __restrict__ gcol) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PairSmem& S = *reinterpret_cast<PairSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    FrameSm& F = S.fr[warp];
    const int64_t nframes = a.count ? (int64_t)*a.count : a.B;
    const bool ties_high = (a.flags & LDPCB_OSD_TIES_HIGH_INDEX_FIRST) != 0;
    const bool disc_from_score = (a.flags & LDPCB_OSD_DISC_HARD_FROM_SCORE) != 0;
    const int g = lane >> 2, t = lane & 3;
    OsdTally tally;
    // element e of a tile = (row g + 8*(e>>1), column 2t + (e&1)); in a tile whose column block starts dlt = 0 or 8 positions
    // right of its row block it is a pair with i >= j iff (g - 2t) + 8*(e>>1) - (e&1) >= dlt
    int pen_d0[4], pen_d8[4];
    const int pen_none[4] = {0, 0, 0, 0};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        pen_d0[e] = ((g - 2 * t) + 8 * (e >> 1) - (e & 1) >= 0) ? -PW_PEN : 0;
        pen_d8[e] = ((g - 2 * t) + 8 * (e >> 1) - (e & 1) >= 8) ? -PW_PEN : 0;
    }

    for (int64_t f0 = (int64_t)blockIdx.x * OSD_FPB; f0 < nframes; f0 += (int64_t)gridDim.x * OSD_FPB) {
        // A warp past the end of the list (last round only) redoes the last frame and writes the same results again: no
        // per-warp condition around the body, so the compiler sees converged code (no BSSY / BRA.DIV around every
        // shuffle and vote)
        const bool active = f0 + warp < nframes;
        const int64_t f = active ? f0 + warp : nframes - 1;
        const int64_t row = a.idx ? (int64_t)a.idx[f] : f;
        Prep P = {};
        long long best_s = 0x7fffffffffffffffll;
        int best_i = 0x7fffffff;
        bool fallback = false;
        {
            P = prepare_frame<false, PAIR1_SH>(a, F, gcol, row, f, lane, ties_high, disc_from_score);
            const unsigned long long d0 = P.d0;
            __syncwarp();
            // ---- R_i, C_j and the empty TEP through the shuffle tables; weight byte planes; B fragments ----------
            int tb[13];
            build_shfl_tables(F, lane, tb);
            int s0 = 0x7fffffff, s1 = 0x7fffffff;
            {
                const int qa = F.qd32[lane], qb = F.qd32[lane + 32], b32 = F.base32;
                const int z = b32 + wpop_shfl(tb, d0);
                const int r0 = b32 + qa + wpop_shfl(tb, d0 ^ P.myprow[0]);
                const int r1 = b32 + qb + wpop_shfl(tb, d0 ^ P.myprow[1]);
                const int c0 = qa + wpop_shfl(tb, P.myprow[0]);
                const int c1 = qb + wpop_shfl(tb, P.myprow[1]);
                const unsigned wa = F.w32[lane], wb = F.w32[lane + 32];
                __syncwarp();  // every lane is done with yo/ys and w32: reuse yo/ys
                int* RCw = reinterpret_cast<int*>(F.yo);  // -R_i, -C_j: the accumulators start at -(R_i + C_j)
                RCw[lane] = -r0; RCw[lane + 32] = -r1;
                RCw[64 + lane] = -c0; RCw[96 + lane] = -c1;
                unsigned char* wq = reinterpret_cast<unsigned char*>(F.ys);  // [plane][LRB position]; w < 2^14
                wq[lane] = (unsigned char)(wa & 0xffu); wq[lane + 32] = (unsigned char)(wb & 0xffu);
                wq[64 + lane] = (unsigned char)((wa >> 8) << 2); wq[96 + lane] = (unsigned char)((wb >> 8) << 2);
                track2(s0, s1, (r0 << 7) | PW_CODE_SINGLE);
                track2(s0, s1, (r1 << 7) | (PW_CODE_SINGLE + 1));
                if (lane == 0) track2(s0, s1, (z << 7) | PW_CODE_EMPTY);
            }
            uint4* Bc = &S.bfrag[warp][0][0];
#pragma unroll
            for (int x = 0; x < 8; ++x) {  // column j = 8x + g, this thread's nibbles of P'_j
                const unsigned long long cj = F.prow[8 * x + g];
                Bc[32 * x + lane] = make_uint4(spread4x2((unsigned)cj, 4 * t), spread4x2((unsigned)cj, 4 * t + 16),
                                               spread4x2((unsigned)(cj >> 32), 4 * t), spread4x2((unsigned)(cj >> 32), 4 * t + 16));
            }
            __syncwarp();
            // ---- the 20 tiles ---------------------------------------------------------------------------
            const unsigned* wqw = reinterpret_cast<const unsigned*>(F.ys);
            const int* RC = reinterpret_cast<const int*>(F.yo);
            unsigned wr[2][2][2];  // [plane][32-bit half of the LRB][16-bit half]: the weight bytes of this thread's k columns
#pragma unroll
            for (int p = 0; p < 2; ++p)
#pragma unroll
                for (int kk = 0; kk < 2; ++kk)
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) wr[p][kk][hh] = wqw[16 * p + 8 * kk + 4 * hh + t];
            int code = 0;
            int gate = pw_gate(__reduce_min_sync(0xffffffffu, s0));  // the singles and the empty TEP are tracked already
            int gn = pw_gate_neg(gate);
#pragma unroll 1
            for (int mi = 0; mi < 4; ++mi) {
                // masked weights of rows 16mi+g and +8
                unsigned afr[2][2][4];  // [k half][plane][fragment register]
                const int i0 = 16 * mi + g;
                const unsigned long long u0 = d0 ^ F.prow[i0], u1 = d0 ^ F.prow[i0 + 8];
                const int nr0 = RC[i0], nr1 = RC[i0 + 8];
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                    const unsigned w0 = kk ? (unsigned)(u0 >> 32) : (unsigned)u0;
                    const unsigned w1 = kk ? (unsigned)(u1 >> 32) : (unsigned)u1;
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const unsigned m0 = mask4(w0, 4 * t + 16 * hh);
                        const unsigned m1 = mask4(w1, 4 * t + 16 * hh);
#pragma unroll
                        for (int p = 0; p < 2; ++p) {
                            afr[kk][p][2 * hh] = wr[p][kk][hh] & m0;
                            afr[kk][p][2 * hh + 1] = wr[p][kk][hh] & m1;
                        }
                    }
                }
                // One 16x8 tile, ONE accumulator per element: it starts at -(R_i + C_j) and the four IMMAs add 2 M[i][j] --
                // the low plane against B bytes of 2, the high plane (bytes (w >> 8) << 2) against B bytes of 128 -- so it
                // ends as -S(i,j) with no arithmetic after the product.  The triangle i < j costs nothing either: the two
                // tiles of a row block that touch the diagonal start 2^29 lower on the elements with i >= j.  A thread's
                // two smallest scores are tracked behind a warp-uniform gate (running warp minimum + truncation window):
                // every score that can still be a candidate at the end passes it.
                auto tile = [&](int nj, const int (&pen)[4]) {
                    const uint4 bf = Bc[32 * nj + lane];
                    const unsigned bl0[2] = {bf.x, bf.y}, bl1[2] = {bf.z, bf.w};
                    const unsigned bh0[2] = {bf.x << 6, bf.y << 6}, bh1[2] = {bf.z << 6, bf.w << 6};
                    const int2 cc = *reinterpret_cast<const int2*>(RC + 64 + 8 * nj + 2 * t);
                    int acc[4] = {nr0 + cc.x + pen[0], nr0 + cc.y + pen[1], nr1 + cc.x + pen[2], nr1 + cc.y + pen[3]};
                    imma_u8(acc, afr[0][0], bl0);
                    imma_u8(acc, afr[0][1], bh0);
                    imma_u8(acc, afr[1][0], bl1);
                    imma_u8(acc, afr[1][1], bh1);
                    const int m4 = max(max(acc[0], acc[1]), max(acc[2], acc[3]));
                    if (__any_sync(0xffffffffu, m4 >= gn)) {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (acc[e] >= gn) track2(s0, s1, ((-acc[e]) << 7) | (code + e));
                        gate = pw_gate(__reduce_min_sync(0xffffffffu, s0));
                        gn = pw_gate_neg(gate);
                    }
                    code += 4;
                };
                int nj = 2 * mi;
                tile(nj, pen_d0);
                tile(nj + 1, pen_d8);
#pragma unroll 1
                for (nj += 2; nj < 8; ++nj) tile(nj, pen_none);
            }
            // ---- candidates inside the truncation window, exact scores ----------------------------------------------
            int m = s0;
            m = __reduce_min_sync(0xffffffffu, m);
            const int lim = (((m >> 7) + OSD_WIN) << 7) | 127;
            unsigned cm = __ballot_sync(0xffffffffu, s0 <= lim);
            const int nc = __popc(cm);
            fallback = __any_sync(0xffffffffu, s1 <= lim) || nc > OSD_CAND_CAP;  // a thread holds two candidates: exact path
            if (!fallback) {
                const bool exact = nc > 1 || a.best_score_q != nullptr;
                const long long q_l0 = (long long)F.qlrb[lane], q_l1 = (long long)F.qlrb[lane + 32];
                while (cm) {
                    const int src = __ffs(cm) - 1;
                    cm &= cm - 1;
                    const int cd = __shfl_sync(0xffffffffu, s0, src) & 127;
                    int pi;
                    if (cd < PW_CODE_SINGLE) {
                        const int tile = cd >> 2, e = cd & 3;
                        const int mi = tile < 8 ? 0 : (tile < 14 ? 1 : (tile < 18 ? 2 : 3));
                        const int nj = tile < 8 ? tile : (tile < 14 ? tile - 6 : (tile < 18 ? tile - 10 : tile - 12));
                        pi = (16 * mi + (src >> 2) + 8 * (e >> 1)) * K + 8 * nj + 2 * (src & 3) + (e & 1);
                    } else {
                        pi = K * K + (cd == PW_CODE_EMPTY ? K : src + 32 * (cd - PW_CODE_SINGLE));
                    }
                    LDPCB_ASSERT(pi >= 0 && pi < OSD_PAIR_TABLE);
                    const int ci = (int)a.pair_index[pi];
                    LDPCB_ASSERT(ci >= 0 && ci < a.n_teps);
                    long long sc = 0;
                    if (exact) {
                        const unsigned tw = __ldg(a.teps + ci);
                        unsigned long long D = d0;
                        long long sm = F.base;
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const unsigned tt = (tw >> (8 * j)) & 0xffu;
                            if (tt < 64u) { D ^= F.prow[tt]; sm += F.qd[tt]; }
                        }
                        const long long sl = (((D >> lane) & 1ull) ? q_l0 : 0ll) + (((D >> (lane + 32)) & 1ull) ? q_l1 : 0ll);
                        sc = sm + warp_sum_ll(sl);
                    }
                    if (sc < best_s || (sc == best_s && ci < best_i)) { best_s = sc; best_i = ci; }
                }
            }
        }
        if (lane == 0) S.fb[warp] = fallback ? 1 : 0;
        __syncthreads();  // (A) every frame of the round is prepared and swept; fallback flags visible
        // ---- exact fallback, CTA-wide, frame by frame (rare) --------------------------------------------------------
        const int nfr = (int)((nframes - f0) < OSD_FPB ? (nframes - f0) : OSD_FPB);
        for (int w = 0; w < nfr; ++w) {
            if (!S.fb[w]) continue;  // CTA-uniform
            const FrameSm& G = S.fr[w];
            build_lut64(S.lut, G, tid);
            __syncthreads();
            long long bs = 0x7fffffffffffffffll;
            int bi = 0x7fffffff;
            for (int i = tid; i < a.n_teps; i += OSD_THREADS) {
                const long long s = score64<2>(S.lut, G, __ldg(a.teps + i));
                if (s < bs) { bs = s; bi = i; }
            }
            warp_argmin(bs, bi);
            if (lane == 0) { S.red_s[warp] = bs; S.red_i[warp] = bi; }
            __syncthreads();
            if (warp == w) {
                best_s = S.red_s[0];
                best_i = S.red_i[0];
#pragma unroll
                for (int v = 1; v < OSD_FPB; ++v) {
                    const long long os = S.red_s[v];
                    const int oi = S.red_i[v];
                    if (os < best_s || (os == best_s && oi < best_i)) { best_s = os; best_i = oi; }
                }
            }
            __syncthreads();  // red_* and the LUT are free for the next fallback frame
        }
        // ---- outputs (each warp finishes its own frame; a warp past the end has nothing to write) ----------------------
        if (active) {
            unsigned long long D = P.d0, flip = 0ull;
            if (best_i != 0x7fffffff) {
                const unsigned tw = __ldg(a.teps + best_i);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const unsigned tt = (tw >> (8 * j)) & 0xffu;
                    if (tt < 64u) { D ^= F.prow[tt]; flip ^= 1ull << tt; }
                }
            }
            const unsigned long long c_lrb = D ^ P.hd_lrb;
            const unsigned long long c_mrb = P.ho_mrb ^ flip;
            F.pos[P.pm[0]] = (unsigned char)((c_mrb >> lane) & 1ull);
            F.pos[P.pm[1]] = (unsigned char)((c_mrb >> (lane + 32)) & 1ull);
            F.pos[P.pm[2]] = (unsigned char)((c_lrb >> lane) & 1ull);
            F.pos[P.pm[3]] = (unsigned char)((c_lrb >> (lane + 32)) & 1ull);
            __syncwarp();
            unsigned wout[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) wout[k] = __ballot_sync(0xffffffffu, F.pos[lane + 32 * k]);
            const int64_t orow = a.idx ? row : f;
            const unsigned wv = lane == 0 ? wout[0] : lane == 1 ? wout[1] : lane == 2 ? wout[2] : wout[3];
            if (lane < 4 && a.cw_bits) a.cw_bits[orow * 4 + lane] = wv;
            if (a.tally_truth) osd_tally_frame(tally, a, orow, wv, best_i, lane);
            if (lane == 0) {
                if (a.best_tep) a.best_tep[orow] = best_i;
                if (a.best_score_q) a.best_score_q[orow] = best_s;
                if (a.score_exp) a.score_exp[f] = P.E;
            }
            if (a.perm) {
#pragma unroll
                for (int k = 0; k < 4; ++k) a.perm[f * N + lane + 32 * k] = P.pm[k];
            }
            if (a.redG) {
                a.redG[f * K + lane] = P.myprow[0];
                a.redG[f * K + lane + 32] = P.myprow[1];
            }
        }
        __syncthreads();  // (B) the next round's prepare overwrites fr[] (read by the other warps in a fallback) and fb[]
    }
    if (a.tally_truth) osd_tally_flush(tally, a, lane);
}

int launch_osd_pair(ldpcb_handle* h, const OsdArgs& a, cudaStream_t st) {
    if (a.B == 0) return LDPCB_OK;
    const int smem = (int)sizeof(PairSmem);
    int& occ = h->occ[OCC_OSD_PAIR];
    if (occ == 0) {
        LDPCB_CUDA(h, cudaFuncSetAttribute(osd_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        LDPCB_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, osd_pair_kernel, OSD_THREADS, smem));
        if (occ < 1) occ = 1;
    }
    int64_t want = (a.B + OSD_FPB - 1) / OSD_FPB;
    int64_t cap = (int64_t)h->sm_count * occ;
    int grid = (int)(want < cap ? want : cap);
    if (grid < 1) grid = 1;
    osd_pair_kernel<<<grid, OSD_THREADS, smem, st>>>(a, h->gcol_dev);
    LDPCB_LAUNCH_CHECK(h, "osd_pair_kernel");
    return LDPCB_OK;
}

}  // namespace ldpcb

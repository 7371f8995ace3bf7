// Ordered-statistics decoding.  One CTA of four warps works on four frames at a time: each warp
// prepares one frame (sort, elimination, P' rows, exact reliabilities), then the four warps sweep the
// TEP list of each prepared frame together.  Everything stays on chip between the 512-byte LLR load and
// the 16-byte codeword store.
//
// Replaces, per frame (reference paths relative to LDPC_128/):
//   swapped_info           PB_OSD/pb_testing.py:306-320  (reliability sort, pi1)
//   identify_mrb           PB_OSD/pb_testing.py:268-304  (pi2, systematic reduced_G = [I | P'])
//   full_gf2elim           PB_OSD/pb_testing.py:231-266  (GF(2) Gauss-Jordan)
//   convention_osd_main    FS_OSD/convention_osd.py:49-77 (TEP sweep, re-encode, discrepancy, argmin)
//   osd.acquire_min        DL_OSD_Testing_serial/ordered_statistics_decoding.py:153-162 (block minima)
//
// Prepare (one warp, one frame; osd_prepare.cuh):
//   1. bitonic sort of the 128 keys |y|, four per lane, on single words (key bits with the index in the low 7
//      bits); a frame in which two keys agree in their upper 24 bits is re-sorted with (key, index) pairs, and
//      one with two equal keys (2.7e-4 of AWGN frames, every frame of a quantised input) is re-ranked by an
//      exact rank sort with tf.argsort's tie rule.
//   2. column-major GF(2) elimination of G[:, pi1]: lane l holds sorted columns l, l+32, l+64, l+96 as 64-bit
//      words (bit r = row r); columns are taken most reliable first, a column with a 1 in a row not
//      yet used becomes the next pivot (greedy most-reliable basis).  Unit columns of G among the 64 most
//      reliable positions join the basis on their own row without a visit; only the other columns are visited
//      (osd_prepare.cuh step 2 says how the pivot rows are chosen so that this is exact).  The reference's rule (row swap /
//      column swap with the first 1 of row i) selects the same basis and its outputs depend only on
//      that basis because identify_mrb re-sorts both halves (pb_testing.py:284-300); DESIGN.md gives
//      the argument, tests compare with the reference's own full_gf2elim.
//   3. P' rows: 64x64 bit transpose of the non-pivot columns (four 32x32 warp butterflies)
//   4. exact integer reliabilities q = rint(|y| * 2^(54-E)); order-0 codeword; per-position deltas
// Sweep (four warps, one frame at a time): for TEP i, D = d0 ^ XOR_{t in TEP} P'_t and
// score = base + sum delta_t + W(D), W the weighted popcount over the 64 LRB reliabilities.
//   5a. orders 0, 1, 3: truncated 32-bit scores through thirteen 5-bit tables held one entry per lane (shuffles)
//   5b. order 2: the tensor-core pair sweep below (IMMA over masked weight byte planes)
//   5c. block minima (DL path): osd_blocks.cu (truncated scores per block + exact re-scoring); FS policy and the
//       near-tie fallbacks: exact 64-bit scores through byte LUTs
//   6. the TEPs within the truncation window of the minimum are re-scored exactly; lexicographic
//      (score, index) minimum = first minimum in enumeration order (tf.argmin)
#include <cmath>
#include <cstddef>
#include <cstdlib>

#include "common.cuh"
#include "osd_prepare.cuh"
#include "osd_sweep.cuh"

namespace ldpcb {

// SOLO (full order-0/1 lists): every warp sweeps its own frame -- 65 TEPs are three per lane -- so the CTA needs no
// barrier, no LUT and no table in shared memory (the dynamic shared memory then holds the four FrameSm only).
template <int MAXW, bool BLOCKS, bool SOLO>
__global__ void __launch_bounds__(OSD_THREADS, SOLO ? 8 : 6) osd_kernel(OsdArgs a, const uint64_t* __restrict__ gcol) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    OsdSmem& S = *reinterpret_cast<OsdSmem*>(smem_raw);  // not touched when SOLO
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    FrameSm& F = SOLO ? reinterpret_cast<FrameSm*>(smem_raw)[warp] : S.fr[warp];
    const int64_t nframes = a.count ? (int64_t)*a.count : a.B;
    const bool ties_high = (a.flags & LDPCB_OSD_TIES_HIGH_INDEX_FIRST) != 0;
    const bool disc_from_score = (a.flags & LDPCB_OSD_DISC_HARD_FROM_SCORE) != 0;
    OsdTally tally;

    for (int64_t f0 = (int64_t)blockIdx.x * OSD_FPB; f0 < nframes; f0 += (int64_t)gridDim.x * OSD_FPB) {
        // A warp past the end of the list (last round only) prepares the last frame again and writes nothing: without a
        // per-warp condition around prepare and sweep the compiler sees converged code (no BSSY / BRA.DIV around the
        // shuffles and votes)
        const bool active = f0 + warp < nframes;
        const int64_t f = active ? f0 + warp : nframes - 1;
        const int64_t row = a.idx ? (int64_t)a.idx[f] : f;
        const Prep P = prepare_frame<BLOCKS>(a, F, gcol, row, f, lane, ties_high, disc_from_score);
        const unsigned char* pm = P.pm;
        const unsigned long long* myprow = P.myprow;
        const unsigned long long hd_lrb = P.hd_lrb, ho_mrb = P.ho_mrb, d0 = P.d0;
        const int E = P.E;
        // ---- 5./6. sweep: the four warps take the prepared frames in turn -------------------------------------
        long long solo_s = 0x7fffffffffffffffll;
        int solo_i = 0x7fffffff;
        if (!BLOCKS) {
            if (!SOLO && lane == 0) { S.cand_n[warp] = 0; S.cand_ovf[warp] = 0; }
            __syncwarp();
            // 5-bit chunk tables of the 32-bit LRB weights: tabs[j][e] = sum of w32[5j+i] over the set bits i of e
            int tb[13];
            build_shfl_tables(F, lane, tb);
            {
                if (!SOLO) {
#pragma unroll
                    for (int j = 0; j < 13; ++j) S.tabs[warp][j][lane] = tb[j];
                }
            }
            if (SOLO) {
                // truncated scores of the empty TEP and of this lane's two single TEPs, warp minimum, then the exact
                // score of everything inside the truncation window (almost always one TEP)
                const int b32 = F.base32;
                const int s_z = b32 + wpop_shfl(tb, d0);
                int s_a = b32 + F.qd32[lane] + wpop_shfl(tb, d0 ^ myprow[0]);
                int s_b = b32 + F.qd32[lane + 32] + wpop_shfl(tb, d0 ^ myprow[1]);
                if (a.n_teps == 1) s_a = s_b = 0x7fffffff;  // order 0
                int m = min(min(s_a, s_b), s_z);
                m = __reduce_min_sync(0xffffffffu, m);
                const int lim = m + OSD_WIN;
                const bool cz = s_z <= lim;
                unsigned ma = __ballot_sync(0xffffffffu, s_a <= lim), mb = __ballot_sync(0xffffffffu, s_b <= lim);
                const int nc = (cz ? 1 : 0) + __popc(ma) + __popc(mb);
                const bool exact = nc > 1 || a.best_score_q != nullptr;
                const long long q_l0 = (long long)F.qlrb[lane], q_l1 = (long long)F.qlrb[lane + 32];
                auto consider = [&](int t) {  // t = MRB position of the single TEP, K = the empty TEP
                    const int ci = (int)a.pair_index[K * K + t];
                    long long sc = 0;
                    if (exact) {
                        const unsigned long long D = d0 ^ F.prow[t];  // prow[K] = 0
                        const long long sl = (((D >> lane) & 1ull) ? q_l0 : 0ll) + (((D >> (lane + 32)) & 1ull) ? q_l1 : 0ll);
                        sc = F.base + F.qd[t] + warp_sum_ll(sl);       // qd[K] = 0
                    }
                    if (sc < solo_s || (sc == solo_s && ci < solo_i)) { solo_s = sc; solo_i = ci; }
                };
                if (cz) consider(K);
                while (ma) { const int t = __ffs(ma) - 1; ma &= ma - 1; consider(t); }
                while (mb) { const int t = __ffs(mb) - 1; mb &= mb - 1; consider(t + 32); }
            }
        }
        const int nfr = SOLO ? 0 : (int)((nframes - f0) < OSD_FPB ? (nframes - f0) : OSD_FPB);
        for (int w = 0; w < nfr; ++w) {
            const FrameSm& G = S.fr[w];
            __syncthreads();  // (A) frame w prepared; reduction slots and LUT free
            if (!BLOCKS) {
                // Fast sweep on 32-bit truncated scores S32 = sum floor(term / 2^30): the 64 LRB weights are folded
                // into thirteen 32-entry tables (5 bits of D each) that live in registers, one entry per lane, and
                // are looked up with warp shuffles -- no shared-memory bank conflicts.  The exact score satisfies
                // S32 * 2^30 <= S < (S32 + 69) * 2^30, so the exact first minimum is among the TEPs with
                // S32 <= min S32 + OSD_WIN; those few are re-scored exactly at output.
                int tabs[13];
#pragma unroll
                for (int j = 0; j < 13; ++j) tabs[j] = S.tabs[w][j][lane];  // built by the frame's owner warp
                const unsigned long long gd0 = G.d0;
                const int gb32 = G.base32;
                int s0 = 0x7fffffff, s1 = 0x7fffffff, s2 = 0x7fffffff, i0 = 0x7fffffff, i1 = 0x7fffffff, i2 = 0x7fffffff;
                int wm = 0x7fffffff - OSD_WIN;  // warp-wide running minimum: only TEPs within the window of it can matter
                const uint32_t* tp = a.teps + warp * 32 + lane;  // the table is padded to whole 128-TEP tiles
                for (int ib = warp * 32; ib < a.n_teps; ib += OSD_THREADS, tp += OSD_THREADS) {
                    const int i = ib + lane;
                    const unsigned tw = __ldg(tp);
                    unsigned long long D = gd0;
                    int s = gb32;
#pragma unroll
                    for (int j = 0; j < MAXW; ++j) {
                        const unsigned t = min((tw >> (8 * j)) & 0xffu, 64u);
                        D ^= G.prow[t];
                        s += G.qd32[t];
                    }
                    const unsigned lo = (unsigned)D, hi = (unsigned)(D >> 32);
                    s += __shfl_sync(0xffffffffu, tabs[0], lo);  // the source lane is taken modulo 32
                    s += __shfl_sync(0xffffffffu, tabs[1], lo >> 5);
                    s += __shfl_sync(0xffffffffu, tabs[2], lo >> 10);
                    s += __shfl_sync(0xffffffffu, tabs[3], lo >> 15);
                    s += __shfl_sync(0xffffffffu, tabs[4], lo >> 20);
                    s += __shfl_sync(0xffffffffu, tabs[5], lo >> 25);
                    s += __shfl_sync(0xffffffffu, tabs[6], (unsigned)(D >> 30));
                    s += __shfl_sync(0xffffffffu, tabs[7], hi >> 3);
                    s += __shfl_sync(0xffffffffu, tabs[8], hi >> 8);
                    s += __shfl_sync(0xffffffffu, tabs[9], hi >> 13);
                    s += __shfl_sync(0xffffffffu, tabs[10], hi >> 18);
                    s += __shfl_sync(0xffffffffu, tabs[11], hi >> 23);
                    s += __shfl_sync(0xffffffffu, tabs[12], hi >> 28);
                    if (i >= a.n_teps) s = 0x7fffffff;
                    wm = min(wm, __reduce_min_sync(0xffffffffu, s));
                    if (s <= wm + OSD_WIN && s <= s2) {  // thread-local three smallest, earlier index first on equal scores
                        if (s < s0) { s2 = s1; i2 = i1; s1 = s0; i1 = i0; s0 = s; i0 = i; }
                        else if (s < s1) { s2 = s1; i2 = i1; s1 = s; i1 = i; }
                        else if (s < s2) { s2 = s; i2 = i; }
                    }
                }
                int m = s0;
                m = __reduce_min_sync(0xffffffffu, m);
                if (lane == 0) S.red32[warp] = m;
                __syncthreads();  // (B)
                m = min(min(S.red32[0], S.red32[1]), min(S.red32[2], S.red32[3]));
                const int lim = m + OSD_WIN;
                if (s0 <= lim) { const int p = atomicAdd(&S.cand_n[w], 1); if (p < OSD_CAND_CAP) S.cand_i[w][p] = i0; }
                if (s1 <= lim) { const int p = atomicAdd(&S.cand_n[w], 1); if (p < OSD_CAND_CAP) S.cand_i[w][p] = i1; }
                if (s2 <= lim) {  // a fourth one may have been dropped by this thread: take the exact path
                    const int p = atomicAdd(&S.cand_n[w], 1); if (p < OSD_CAND_CAP) S.cand_i[w][p] = i2;
                    S.cand_ovf[w] = 1;
                }
                __syncthreads();  // (C)
                if (S.cand_ovf[w] || S.cand_n[w] > OSD_CAND_CAP) {
                    // too many near-ties (e.g. quantised inputs): exact 64-bit sweep through the byte LUTs
                    build_lut64(S.lut, G, tid);
                    __syncthreads();
                    long long bs = 0x7fffffffffffffffll;
                    int bi = 0x7fffffff;
                    for (int i = tid; i < a.n_teps; i += OSD_THREADS) {
                        const long long s = score64<MAXW>(S.lut, G, __ldg(a.teps + i));
                        if (s < bs) { bs = s; bi = i; }
                    }
                    warp_argmin(bs, bi);
                    if (lane == 0) { S.red_s[w][warp] = bs; S.red_i[w][warp] = bi; }
                }
            } else {
                build_lut64(S.lut, G, tid);
                __syncthreads();
                // block minima: warp v takes blocks v, v+4, ... of this frame
                const int64_t fw = f0 + w;
                for (int blk = warp; blk < a.n_blocks; blk += OSD_FPB) {
                    const int i0 = a.block_start[blk], i1 = a.block_start[blk + 1];
                    long long bs = 0x7fffffffffffffffll;
                    int bi = 0x7fffffff;
                    for (int i = i0 + lane; i < i1; i += 32) {
                        const long long s = score64<MAXW>(S.lut, G, __ldg(a.teps + i));
                        if (s < bs) { bs = s; bi = i; }
                    }
                    warp_argmin(bs, bi);
                    if (lane == 0) {
                        a.block_min_q[fw * a.n_blocks + blk] = bs;
                        if (a.block_arg) a.block_arg[fw * a.n_blocks + blk] = bi;
                    }
                }
            }
        }
        if (!SOLO) __syncthreads();  // all partial minima written; all LUT reads done
        // ---- outputs (each warp finishes its own frame) ---------------------------------------------------
        // SOLO: a warp past the end has the last frame's results and stores them again (the loop body stays free of
        // per-warp conditions); otherwise its frame was not swept and it has nothing to write
        if (SOLO || active) {
            if (!BLOCKS) {
                long long best_s = 0x7fffffffffffffffll;
                int best_i = 0x7fffffff;
                const int nc = SOLO ? 0 : S.cand_n[warp];
                if (SOLO) {
                    best_s = solo_s;
                    best_i = solo_i;
                } else if (S.cand_ovf[warp] || nc > OSD_CAND_CAP) {
                    best_s = S.red_s[warp][0];
                    best_i = S.red_i[warp][0];
#pragma unroll
                    for (int v = 1; v < OSD_FPB; ++v) {
                        const long long os = S.red_s[warp][v];
                        const int oi = S.red_i[warp][v];
                        if (os < best_s || (os == best_s && oi < best_i)) { best_s = os; best_i = oi; }
                    }
                } else if (nc == 1 && !a.best_score_q) {
                    best_i = S.cand_i[warp][0];  // a single candidate is the exact minimum
                } else {
                    // exact scores of the few candidates, warp-cooperative weighted popcount
                    const long long q_l0 = (long long)F.qlrb[lane], q_l1 = (long long)F.qlrb[lane + 32];
                    for (int c = 0; c < nc; ++c) {
                        const int ci = S.cand_i[warp][c];
                        const unsigned tw = __ldg(a.teps + ci);
                        unsigned long long D = d0;
                        long long sm = F.base;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const unsigned t = (tw >> (8 * j)) & 0xffu;
                            if (t < 64u) { D ^= F.prow[t]; sm += F.qd[t]; }
                        }
                        const long long sl = (((D >> lane) & 1ull) ? q_l0 : 0ll) + (((D >> (lane + 32)) & 1ull) ? q_l1 : 0ll);
                        const long long sc = sm + warp_sum_ll(sl);
                        if (sc < best_s || (sc == best_s && ci < best_i)) { best_s = sc; best_i = ci; }
                    }
                }
                // re-encode the winner and un-permute
                unsigned long long D = d0, flip = 0ull;
                if (best_i != 0x7fffffff) {
                    const unsigned tw = __ldg(a.teps + best_i);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const unsigned t = (tw >> (8 * j)) & 0xffu;
                        if (t < 64u) { D ^= F.prow[t]; flip ^= 1ull << t; }
                    }
                }
                const unsigned long long c_lrb = D ^ hd_lrb;
                const unsigned long long c_mrb = ho_mrb ^ flip;
                F.pos[pm[0]] = (unsigned char)((c_mrb >> lane) & 1ull);
                F.pos[pm[1]] = (unsigned char)((c_mrb >> (lane + 32)) & 1ull);
                F.pos[pm[2]] = (unsigned char)((c_lrb >> lane) & 1ull);
                F.pos[pm[3]] = (unsigned char)((c_lrb >> (lane + 32)) & 1ull);
                __syncwarp();
                unsigned wout[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) wout[k] = __ballot_sync(0xffffffffu, F.pos[lane + 32 * k]);
                const int64_t orow = a.idx ? row : f;
                const unsigned wv = lane == 0 ? wout[0] : lane == 1 ? wout[1] : lane == 2 ? wout[2] : wout[3];
                if (lane < 4 && a.cw_bits) a.cw_bits[orow * 4 + lane] = wv;
                if (a.tally_truth && active) osd_tally_frame(tally, a, orow, wv, best_i, lane);
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
        // the next round's prepare overwrites fr[] and red_*: every warp has passed the barrier above and
        // only touches its own FrameSm until the next barrier
        if (SOLO) __syncthreads();  // not needed for the data: a CTA barrier in the loop is what lets the compiler treat the body as converged code (25 BRA.DIV and half of the BSSY/BSYNC pairs go away)
    }
    if (!BLOCKS && a.tally_truth) osd_tally_flush(tally, a, lane);
}


// ---- FS-OSD policy (Choi & Jeong 2019 as re-implemented by the reference) ------------------------------------
// fs_osd, FS_OSD/fs_testing.py:129-165, per frame:
//   order 0: accept at once if the Hamming distance of the order-0 codeword to the hard decision is < tau_e
//   (:131-135); order j+1 is entered only if  sum of the j+1 least reliable MRB |y|  + beta*(n-k) < w_dmin
//   (:22-30,137-139); inside an order TEPs are taken in FS order; a TEP whose codeword is closer than tau_e
//   stops everything WITHOUT becoming the decision (the reference's quirk, :143-147,162); a TEP closer than
//   tau_psc with a strictly smaller weighted distance becomes the decision (:148-152).
// The sequential loop is restated as, per order: find the first stopping TEP by a parallel sweep; the decision
// is the first minimum among the eligible TEPs before it.  Weighted distances are the exact integer scores.

__device__ __forceinline__ void fs_eval(const OsdSmem& S, const FrameSm& G, unsigned tw, unsigned long long gd0, long long gbase,
                                        unsigned long long d0m, long long& s, int& hd) {
    unsigned long long D = gd0, flip = 0ull;
    s = gbase;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const unsigned t = (tw >> (8 * j)) & 0xffu;
        if (t < 64u) { D ^= G.prow[t]; s += G.qd[t]; flip |= 1ull << t; }
    }
#pragma unroll
    for (int b = 0; b < 8; ++b) s += (long long)S.lut[b][(unsigned)(D >> (8 * b)) & 0xffu];
    hd = __popcll(D) + __popcll(flip ^ d0m);
}

constexpr int FS_CHUNK = 8 * OSD_THREADS;  // 1024 TEPs between two "has anybody stopped?" barriers

__global__ void __launch_bounds__(OSD_THREADS) osd_fs_kernel(OsdArgs a, FsParams fp, const uint64_t* __restrict__ gcol) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    OsdSmem& S = *reinterpret_cast<OsdSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    FrameSm& F = S.fr[warp];
    const int64_t nframes = a.count ? (int64_t)*a.count : a.B;
    const int cls_start[5] = {0, 1, 65, 2081, 43745};

    for (int64_t f0 = (int64_t)blockIdx.x * OSD_FPB; f0 < nframes; f0 += (int64_t)gridDim.x * OSD_FPB) {
        const bool active = f0 + warp < nframes;  // a warp past the end prepares the last frame again (see osd_kernel)
        const int64_t f = active ? f0 + warp : nframes - 1;
        const int64_t row = a.idx ? (int64_t)a.idx[f] : f;
        const Prep P = prepare_frame<false>(a, F, gcol, row, f, lane, false, false);
        if (lane == 0) {
            const double sh = (double)fp.beta_shift * __hiloint2double((1023 + 54 - P.E) << 20, 0);
            S.fs_score[warp] = sh >= 4.6e18 ? (1ll << 62) : __double2ll_rn(sh);
        }
        const int nfr = (int)((nframes - f0) < OSD_FPB ? (nframes - f0) : OSD_FPB);
        for (int w = 0; w < nfr; ++w) {
            const FrameSm& G = S.fr[w];
            __syncthreads();
            {
                const int b = tid >> 4, lo = tid & 15;
                unsigned long long wv[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) wv[i] = G.qlrb[8 * b + i];
                unsigned long long lsum = 0ull;
#pragma unroll
                for (int i = 0; i < 4; ++i) lsum += ((lo >> i) & 1) ? wv[i] : 0ull;
                unsigned long long e[16];
                e[0] = lsum;
#pragma unroll
                for (int x = 1; x < 16; ++x) e[x] = e[x & (x - 1)] + wv[4 + (31 - __clz(x & -x))];
#pragma unroll
                for (int x = 0; x < 16; ++x) S.lut[b][x * 16 + lo] = e[x];
            }
            __syncthreads();
            const unsigned long long gd0 = G.d0, d0m = G.d0m;
            const long long gbase = G.base;
            // order 0 (every thread computes it: uniform state without a broadcast)
            long long w_dmin;
            int hd0;
            fs_eval(S, G, 0xffffffffu, gd0, gbase, d0m, w_dmin, hd0);
            int opt = 0, num = 1, kind = 3;
            if ((float)hd0 < fp.tau_e) {
                kind = 0;
            } else {
                // beta*(n-k) in this frame's integer score units; the owner warp left it in fs_score[w]
                const long long shift_q = S.fs_score[w];
                long long bnd = 0;
                for (int j = 0; j < fp.order; ++j) {
                    const long long qv = G.qd[63 - j];
                    bnd += qv < 0 ? -qv : qv;
                    if (!(bnd + shift_q < w_dmin)) { kind = 2; break; }
                    if (fp.defer3 && j == 2) { kind = 4; break; }  // the weight-3 class goes to the tensor-core sweep (osd3.cu)
                    const int r0 = cls_start[j + 1], r1 = cls_start[j + 2];
                    // The class is swept in chunks of FS_CHUNK TEPs: the sequential loop of the reference stops at the first
                    // TEP closer than tau_e, on average after ~600 (order 2) / ~3000 (order 3) of the 2016 / 41664 TEPs of
                    // the class, so sweeping whole classes to find that TEP wasted most of the work.  A chunk costs one
                    // barrier (__syncthreads_or: did any thread see a stop?); only the chunk that holds the stop is swept
                    // twice (the decision must not see the TEPs behind it).
                    long long pbs = 0x7fffffffffffffffll;  // this thread's best over the completed chunks
                    int pbi = 0x7fffffff;
                    int first_stop = 0x7fffffff;
                    for (int c0 = r0; c0 < r1; c0 += FS_CHUNK) {
                        const int c1 = c0 + FS_CHUNK < r1 ? c0 + FS_CHUNK : r1;
                        long long cbs = 0x7fffffffffffffffll;
                        int cbi = 0x7fffffff, fst = 0x7fffffff;
                        for (int i = c0 + tid; i < c1; i += OSD_THREADS) {
                            long long s;
                            int hd;
                            fs_eval(S, G, __ldg(a.teps + i), gd0, gbase, d0m, s, hd);
                            if ((float)hd < fp.tau_e) fst = min(fst, i);
                            if (hd < fp.tau_psc && s < cbs) { cbs = s; cbi = i; }
                        }
                        if (__syncthreads_or(fst != 0x7fffffff)) {
                            fst = __reduce_min_sync(0xffffffffu, fst);
                            if (lane == 0) S.red_stop[warp] = fst;
                            __syncthreads();
                            first_stop = min(min(S.red_stop[0], S.red_stop[1]), min(S.red_stop[2], S.red_stop[3]));
                            cbs = 0x7fffffffffffffffll; cbi = 0x7fffffff;
                            for (int i = c0 + tid; i < first_stop; i += OSD_THREADS) {
                                long long s;
                                int hd;
                                fs_eval(S, G, __ldg(a.teps + i), gd0, gbase, d0m, s, hd);
                                if (hd < fp.tau_psc && s < cbs) { cbs = s; cbi = i; }
                            }
                            if (cbs < pbs) { pbs = cbs; pbi = cbi; }
                            break;
                        }
                        if (cbs < pbs) { pbs = cbs; pbi = cbi; }
                    }
                    long long bs = pbs;
                    int bi = pbi;
                    warp_argmin(bs, bi);
                    __syncthreads();  // previous use of the reduction slots is over
                    if (lane == 0) { S.red_s[w][warp] = bs; S.red_i[w][warp] = bi; }
                    __syncthreads();
                    bs = S.red_s[w][0]; bi = S.red_i[w][0];
#pragma unroll
                    for (int v = 1; v < OSD_FPB; ++v) {
                        const long long os = S.red_s[w][v];
                        const int oi = S.red_i[w][v];
                        if (os < bs || (os == bs && oi < bi)) { bs = os; bi = oi; }
                    }
                    if (bs < w_dmin) { w_dmin = bs; opt = bi; }
                    if (first_stop == 0x7fffffff) {
                        num += r1 - r0;
                    } else {
                        num += first_stop - r0 + 1;
                        kind = 1;
                        break;
                    }
                }
            }
            __syncthreads();
            if (tid == 0) { S.fs_score[w] = w_dmin; S.fs_opt[w] = opt; S.fs_num[w] = num; S.fs_kind[w] = kind; }
        }
        __syncthreads();
        if (active) {
            const int best_i = S.fs_opt[warp];
            unsigned long long D = P.d0, flip = 0ull;
            const unsigned tw = __ldg(a.teps + best_i);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const unsigned t = (tw >> (8 * j)) & 0xffu;
                if (t < 64u) { D ^= F.prow[t]; flip ^= 1ull << t; }
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
            if (lane < 4 && a.cw_bits) {
                const unsigned wv = lane == 0 ? wout[0] : lane == 1 ? wout[1] : lane == 2 ? wout[2] : wout[3];
                a.cw_bits[orow * 4 + lane] = wv;
            }
            if (lane == 0) {
                if (a.best_tep) a.best_tep[orow] = best_i;
                if (a.best_score_q) a.best_score_q[orow] = S.fs_score[warp];
                if (a.score_exp) a.score_exp[f] = P.E;
                if (fp.num_teps) fp.num_teps[orow] = S.fs_num[warp];
                if (fp.stop_kind) fp.stop_kind[orow] = (uint8_t)S.fs_kind[warp];
                if (S.fs_kind[warp] == 4) {
                    const int p = atomicAdd(fp.d3_count, 1);
                    LDPCB_ASSERT(p >= 0 && p < nframes);
                    fp.d3_list[p] = (int32_t)row;
                    fp.d3_wdmin[p] = S.fs_score[warp];
                    fp.d3_opt[p] = S.fs_opt[warp];
                    fp.d3_num[p] = S.fs_num[warp];
                }
            }
            if (a.perm) {
#pragma unroll
                for (int k = 0; k < 4; ++k) a.perm[f * N + lane + 32 * k] = P.pm[k];
            }
        }
        __syncthreads();  // fs_* consumed before the next round's sweeps rewrite them
    }
}

int launch_osd_fs(ldpcb_handle* h, const OsdArgs& a, const FsParams& fp, cudaStream_t st) {
    if (a.B == 0) return LDPCB_OK;
    const int smem = (int)offsetof(OsdSmem, tabs);  // the FS sweeps score through the byte LUT: the shuffle-table copies (last member, 6.5 KB) are not touched -> 7 instead of 5 CTAs per SM
    int& occ = h->occ[OCC_OSD_FS];
    if (occ == 0) {
        LDPCB_CUDA(h, cudaFuncSetAttribute(osd_fs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        LDPCB_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, osd_fs_kernel, OSD_THREADS, smem));
        if (occ < 1) occ = 1;
    }
    int64_t want = (a.B + OSD_FPB - 1) / OSD_FPB;
    int64_t cap = (int64_t)h->sm_count * occ;
    int grid = (int)(want < cap ? want : cap);
    if (grid < 1) grid = 1;
    osd_fs_kernel<<<grid, OSD_THREADS, smem, st>>>(a, fp, h->gcol_dev);
    LDPCB_LAUNCH_CHECK(h, "osd_fs_kernel");
    return LDPCB_OK;
}

// FS policy at order_limit 3 (the reference's default, FS_OSD/globalmap.py:44) in three launches:
//   1. osd_fs_kernel, classes 0..2; a frame that reaches class 3 is appended, with its decision so far, to a list
//   2. osd3_kernel<FS> sweeps class 3 of those frames on the tensor cores (scores + Hamming distances)
//   3. osd_fs_kernel again, full policy, on the few frames step 2 could not decide (a tau_e stop inside class 3)
// About 7 % of the NMS failures at 2.5 dB reach class 3 and nearly all of them sweep it completely: 41,664 TEPs through the
// 64-bit byte LUT was 70 % of the order-3 time.
int launch_osd_fs3(ldpcb_handle* h, const OsdArgs& a, const FsParams& fp, cudaStream_t st) {
    if (a.B == 0) return LDPCB_OK;
    const size_t n = (size_t)a.B;
    const size_t need = 256 + n * (4 + 4 + 8 + 4 + 4) + 1024;
    Workspace& w = h->fb_ws[st];
    if (w.cap < need) {
        if (w.buf) { LDPCB_CUDA(h, cudaDeviceSynchronize()); LDPCB_CUDA(h, cudaFree(w.buf)); w.buf = nullptr; w.cap = 0; }
        LDPCB_CUDA(h, cudaMalloc(&w.buf, need + need / 4));
        w.cap = need + need / 4;
    }
    int32_t* counts = reinterpret_cast<int32_t*>(w.buf);            // [0]: deferred, [32]: undecided
    long long* wdmin = reinterpret_cast<long long*>(w.buf + 256);
    int32_t* d3_list = reinterpret_cast<int32_t*>(w.buf + 256 + n * 8);
    int32_t* l3_list = d3_list + n;
    int32_t* opt = l3_list + n;
    int32_t* num = opt + n;
    LDPCB_CUDA(h, cudaMemsetAsync(counts, 0, 256, st));
    FsParams f1 = fp;
    f1.defer3 = 1; f1.d3_list = d3_list; f1.d3_count = counts; f1.d3_wdmin = wdmin; f1.d3_opt = opt; f1.d3_num = num;
    int s = launch_osd_fs(h, a, f1, st);
    if (s != LDPCB_OK) return s;
    const int64_t bound = a.B < 32768 ? a.B : 32768;  // grid bound of the follow-up launches: they stride over device-side counts
    OsdArgs a2 = a;
    a2.idx = d3_list; a2.count = counts; a2.B = bound;
    a2.score_exp = nullptr; a2.perm = nullptr; a2.redG = nullptr;  // written by launch 1 for every frame
    Fs3Args f2;
    f2.wdmin = wdmin; f2.opt = opt; f2.num = num;
    f2.he = fp.tau_psc - 3;
    f2.hs = (int)ceilf(fp.tau_e) - 3;
    f2.num_teps = fp.num_teps; f2.stop_kind = fp.stop_kind;
    if ((s = launch_osd3_fs(h, a2, f2, l3_list, counts + 32, st)) != LDPCB_OK) return s;
    OsdArgs a3 = a2;
    a3.idx = l3_list; a3.count = counts + 32;
    return launch_osd_fs(h, a3, fp, st);
}

template <int MAXW, bool BLOCKS, bool SOLO = false>
static int launch_variant(ldpcb_handle* h, const OsdArgs& a, cudaStream_t st) {
    auto kern = osd_kernel<MAXW, BLOCKS, SOLO>;
    const int smem = SOLO ? OSD_FPB * (int)sizeof(FrameSm) : (int)sizeof(OsdSmem);
    int& occ = h->occ[OCC_OSD + (SOLO ? 8 : (MAXW - 1) * 2 + (BLOCKS ? 1 : 0))];
    if (occ == 0) {
        LDPCB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        LDPCB_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, OSD_THREADS, smem));
        if (occ < 1) occ = 1;
    }
    int64_t want = (a.B + OSD_FPB - 1) / OSD_FPB;
    int64_t cap = (int64_t)h->sm_count * occ;
    int grid = (int)(want < cap ? want : cap);
    if (grid < 1) grid = 1;
    kern<<<grid, OSD_THREADS, smem, st>>>(a, h->gcol_dev);
    LDPCB_LAUNCH_CHECK(h, "osd_kernel");
    return LDPCB_OK;
}

int launch_osd_generic3(ldpcb_handle* h, const OsdArgs& a, cudaStream_t st) { return launch_variant<3, false>(h, a, st); }

int launch_osd(ldpcb_handle* h, const OsdArgs& a, cudaStream_t st) {
    if (a.B == 0) return LDPCB_OK;
    const bool blocks = a.block_start != nullptr;
    if (blocks && !getenv("LDPCB_BLOCKS_LUT")) return launch_osd_blocks(h, a, st);  // osd_blocks.cu; the variable keeps the byte-LUT sweep below for A/B runs
    switch (a.maxw) {
        case 1:
            if (blocks) return launch_variant<1, true>(h, a, st);
            return (a.pair_index && (a.n_teps == 1 || a.n_teps == 65)) ? launch_variant<1, false, true>(h, a, st) : launch_variant<1, false>(h, a, st);
        case 2:
            if (blocks) return launch_variant<2, true>(h, a, st);
            if (a.pair_index && a.n_teps == 2081) return launch_osd_pair(h, a, st);  // osd_pair.cu
            return launch_variant<2, false>(h, a, st);
        case 3:
            if (blocks) return launch_variant<3, true>(h, a, st);
            if (a.pair_index && a.triple_index && a.n_teps == 43745 && !a.redG_in && !getenv("LDPCB_OSD3_GENERIC")) return launch_osd3(h, a, st);  // osd3.cu
            return launch_variant<3, false>(h, a, st);
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
    LDPCB_ENTER(h);
    if (B < 0 || order < 0 || order > 3 || tep_order < 0 || tep_order > 1 || (flags & ~3))
        return set_error(h, LDPCB_ERR_ARG, "ldpcb_osd_decode: B=%lld order=%d tep_order=%d flags=%d out of range", (long long)B, order, tep_order, flags);
    if (B == 0) return LDPCB_OK;
    int st = check_llr(h, "ldpcb_osd_decode", order_llr_dev, score_llr_dev);
    if (st != LDPCB_OK) return st;
    if (!cw_bits_dev) return set_error(h, LDPCB_ERR_ARG, "ldpcb_osd_decode: NULL cw_bits");
    const TepTable& t = h->tep[order][tep_order];
    OsdArgs a = {};
    a.order_llr = order_llr_dev; a.score_llr = score_llr_dev; a.B = B;
    a.teps = t.dev; a.n_teps = t.n; a.maxw = t.maxw; a.pair_index = t.pair_dev; a.triple_index = t.triple_dev; a.flags = flags;
    a.cw_bits = cw_bits_dev; a.best_tep = best_tep_dev; a.best_score_q = best_score_q_dev;
    a.score_exp = score_exp_dev; a.perm = perm_dev; a.redG = redG_dev;
    return launch_osd(h, a, (cudaStream_t)stream);
}

extern "C" int ldpcb_osd_block_minima(ldpcb_t* h, const float* order_llr_dev, const float* score_llr_dev, int64_t B,
                                      const uint32_t* teps_dev, int32_t n_teps, const int32_t* block_start_dev,
                                      int32_t n_blocks, int flags, int64_t* block_min_q_dev, int32_t* block_arg_dev,
                                      int32_t* score_exp_dev, const uint32_t* truth_bits_dev,
                                      int64_t* truth_score_q_dev, uint8_t* perm_dev, void* stream) {
    LDPCB_ENTER(h);
    const int maxw_hint = (flags >> LDPCB_OSD_MAXW_SHIFT) & 7;
    flags &= ~(7 << LDPCB_OSD_MAXW_SHIFT);
    if (B < 0 || n_teps < 0 || n_blocks < 1 || (flags & ~3) || maxw_hint > 4)
        return set_error(h, LDPCB_ERR_ARG, "ldpcb_osd_block_minima: B=%lld n_teps=%d n_blocks=%d flags=%d out of range", (long long)B, n_teps, n_blocks, flags);
    if (B == 0) return LDPCB_OK;
    int st = check_llr(h, "ldpcb_osd_block_minima", order_llr_dev, score_llr_dev);
    if (st != LDPCB_OK) return st;
    if (!teps_dev || !block_start_dev || !block_min_q_dev)
        return set_error(h, LDPCB_ERR_ARG, "ldpcb_osd_block_minima: NULL teps, block_start or block_min_q");
    OsdArgs a = {};
    a.order_llr = order_llr_dev; a.score_llr = score_llr_dev; a.B = B;
    a.teps = teps_dev; a.n_teps = n_teps; a.maxw = maxw_hint ? maxw_hint : 4; a.flags = flags;
    a.block_start = block_start_dev; a.n_blocks = n_blocks;
    a.block_min_q = block_min_q_dev; a.block_arg = block_arg_dev; a.score_exp = score_exp_dev;
    a.truth_bits = truth_bits_dev; a.truth_score_q = truth_score_q_dev; a.perm = perm_dev;
    return launch_osd(h, a, (cudaStream_t)stream);
}

extern "C" int ldpcb_osd_fs_decode(ldpcb_t* h, const float* llr_dev, int64_t B, int order_limit, float tau_e, int tau_psc,
                                   float beta_shift, uint32_t* cw_bits_dev, int32_t* best_tep_dev, int32_t* num_teps_dev,
                                   uint8_t* stop_kind_dev, int64_t* best_score_q_dev, int32_t* score_exp_dev,
                                   uint8_t* perm_dev, void* stream) {
    LDPCB_ENTER(h);
    if (B < 0 || order_limit < 0 || order_limit > 3)
        return set_error(h, LDPCB_ERR_ARG, "ldpcb_osd_fs_decode: B=%lld order_limit=%d out of range", (long long)B, order_limit);
    if (B == 0) return LDPCB_OK;
    int st = check_llr(h, "ldpcb_osd_fs_decode", llr_dev, llr_dev);
    if (st != LDPCB_OK) return st;
    if (!cw_bits_dev) return set_error(h, LDPCB_ERR_ARG, "ldpcb_osd_fs_decode: NULL cw_bits");
    const TepTable& t = h->tep[order_limit][LDPCB_TEP_FS];
    OsdArgs a = {};
    a.order_llr = llr_dev; a.score_llr = llr_dev; a.B = B; a.teps = t.dev; a.n_teps = t.n; a.maxw = t.maxw;
    a.cw_bits = cw_bits_dev; a.best_tep = best_tep_dev; a.best_score_q = best_score_q_dev; a.score_exp = score_exp_dev;
    a.perm = perm_dev;
    FsParams fp;
    fp.tau_e = tau_e; fp.tau_psc = tau_psc; fp.beta_shift = beta_shift; fp.order = order_limit;
    fp.num_teps = num_teps_dev; fp.stop_kind = stop_kind_dev;
    a.pair_index = t.pair_dev; a.triple_index = t.triple_dev;
    if (order_limit == 3 && !getenv("LDPCB_FS3_EXACT")) return launch_osd_fs3(h, a, fp, (cudaStream_t)stream);
    return launch_osd_fs(h, a, fp, (cudaStream_t)stream);
}

// Order-3 OSD on the full TEP lists (conventional or FS order, 43,745 TEPs): every warp prepares AND sweeps its own frame,
// the weight-3 class on the tensor cores.
//
// Replaces the same reference code as osd.cu (swapped_info / identify_mrb / full_gf2elim, PB_OSD/pb_testing.py:231-320,
// and convention_osd_main, FS_OSD/convention_osd.py:49-77, with order_limit = 3 -- the reference's default,
// PB_OSD/globalmap.py:42, FS_OSD/globalmap.py:44).
//
// The pair identity of osd_pair.cu, one level up.  For a fixed third position k, d_k = d0 ^ P'_k is the discrepancy word
// of the single TEP {k}, and the triple {i, j, k} (i < j < k) scores
//     S(i,j,k) = R^k_i + C_j - 2 M^k[i][j],   R^k_i = base + qd_k + qd_i + W(d_k ^ P'_i)  (= score of the pair {i, k}),
//     C_j = qd_j + W(P'_j),                    M^k[i][j] = sum_l w_l (d_k ^ P'_i)[l] P'_j[l]
// -- 62 pair problems of shrinking size (k = 2..63), plus the pass "no third position" (d = d0) that yields the pairs, the
// singles and the empty TEP.  M^k is an integer product on IMMA.16832.U8.U8 over two byte planes of the truncated
// weights (w = floor(q / 2^38) < 2^16), 556 16x8 tiles per frame.  What makes a pass cheap:
//   * B fragments (bits of P'_j as bytes) do not depend on k: built once per frame, 4 KB of shared memory per warp;
//   * A fragments are w & mask(d_k ^ P'_i) = w & (mask(d_k) ^ mask(P'_i)): the byte masks of the 64 rows are built once
//     per frame (4 KB), mask(d_k) once per pass, and an A register is ONE LOP3 from there;
//   * R^k comes from the warp's 5-bit shuffle tables (13 shuffles per value), C once per frame;
//   * a thread keeps its three smallest scores behind a gate on the tile minimum, so the common tile costs two
//     multiply-adds and one add per element and a 3-input minimum per pair of elements.
// The candidates within the truncation window of the minimum are re-scored exactly in int64 (lexicographic (score,
// enumeration index) minimum = tf.argmin's first minimum).  A frame whose window holds more candidates than the
// threads track (quantised inputs) is appended to a list and swept exactly by the generic kernel afterwards.
#include "common.cuh"
#include "osd_prepare.cuh"
#include "osd_sweep.cuh"

namespace ldpcb {

struct __align__(16) Osd3Warp {
    FrameSm fr;
    uint4 bfrag[8][32];   // [column block nj][lane]: B fragment registers {lo bits 4t.., lo 4t+16.., hi 4t.., hi 4t+16..}
    uint4 maskp[64][4];   // [row i][t]: byte masks (0x00 / 0xFF) of the same four nibbles of P'_i
    int C[64];
    int R[64];
    int HC[64];           // FS: |P'_j| (Hamming weights for the tau_e / tau_psc tests), 2^20 for j >= kl
    int HR[64];           // FS, per pass: |d_k ^ P'_i|
};

constexpr int O3_NONE = 64;  // pass without a third position
constexpr int O3_INF = 0x7fffffff;
constexpr int O3_PEN = 1 << 29;   // an element outside the triangle i < j < kl starts its accumulator (-S) this much lower
constexpr int O3_BIG = 1 << 29;
// candidate ids: pass << 8 | mi << 6 | nj << 3 | e << 1 ... kept simple: fields below
__device__ __forceinline__ int o3_id(int pass, int mi, int nj, int e) { return (pass << 7) | (mi << 5) | (nj << 2) | e; }
constexpr int O3_ID_SINGLE = 1 << 20, O3_ID_EMPTY = 1 << 21;

__device__ __forceinline__ void o3_track(int (&s)[3], int (&id)[3], int p, int pid) {
    if (p < s[2]) {
        if (p < s[0]) { s[2] = s[1]; id[2] = id[1]; s[1] = s[0]; id[1] = id[0]; s[0] = p; id[0] = pid; }
        else if (p < s[1]) { s[2] = s[1]; id[2] = id[1]; s[1] = p; id[1] = pid; }
        else { s[2] = p; id[2] = pid; }
    }
}

// FS = the weight-3 class of the FS policy (FS_OSD/fs_testing.py:137-152) for the frames osd_fs_kernel deferred: only the
// 62 passes with a third position run; a third, unit-weight IMMA plane gives every triple's Hamming distance to the hard
// decision (|u ^ P_j| = |u| + |P_j| - 2 sum u.P_j), TEPs at distance >= tau_psc are not eligible, and a frame in which any
// triple is closer than tau_e (the sequential loop would stop there) or whose window overflows goes to the exact kernel.
template <bool FS>
__global__ void __launch_bounds__(OSD_THREADS, FS ? 3 : 4) osd3_kernel(OsdArgs a, const uint64_t* __restrict__ gcol, const uint16_t* __restrict__ triple_index,
                                                                       int32_t* fb_list, int32_t* fb_count, Fs3Args fs) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    Osd3Warp& W = reinterpret_cast<Osd3Warp*>(smem_raw)[warp];
    FrameSm& F = W.fr;
    const int64_t nframes = a.count ? (int64_t)*a.count : a.B;
    const bool ties_high = (a.flags & LDPCB_OSD_TIES_HIGH_INDEX_FIRST) != 0;
    const bool disc_from_score = (a.flags & LDPCB_OSD_DISC_HARD_FROM_SCORE) != 0;
    const int g = lane >> 2, t = lane & 3;
    const int vb = g - 2 * t;  // i - j of element 0 of a tile whose row and column blocks start at the same index
    OsdTally tally;
    // high-plane accumulator initialisers of the tiles on the diagonal: element e = (row g + 8*(e>>1), column 2t + (e&1)) of a
    // tile whose column block starts dlt = 0 or 8 positions right of its row block is a pair with i >= j iff vb + 8rs - cs >= dlt
    int pen_d0[4], pen_d8[4];
    const int pen_none[4] = {0, 0, 0, 0};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        pen_d0[e] = (vb + 8 * (e >> 1) - (e & 1) >= 0) ? -O3_PEN : 0;
        pen_d8[e] = (vb + 8 * (e >> 1) - (e & 1) >= 8) ? -O3_PEN : 0;
    }
    const int64_t gw = (int64_t)blockIdx.x * OSD_FPB + warp, nw = (int64_t)gridDim.x * OSD_FPB;

    for (int64_t f = gw; f < nframes; f += nw) {
        const int64_t row = a.idx ? (int64_t)a.idx[f] : f;
        __syncwarp();
        Prep P = prepare_frame<false, PAIR_SH>(a, F, gcol, row, f, lane, ties_high, disc_from_score);
        const unsigned long long d0 = P.d0;
        __syncwarp();
        // ---- per-frame tables ----------------------------------------------------------------------------------
        int tb[13];
        build_shfl_tables(F, lane, tb);
        {
            const unsigned wa = F.w32[lane], wb = F.w32[lane + 32];
            W.C[lane] = -(F.qd32[lane] + wpop_shfl(tb, P.myprow[0]));  // -C_j: the low-plane accumulators start at -(R_i + C_j)
            W.C[lane + 32] = -(F.qd32[lane + 32] + wpop_shfl(tb, P.myprow[1]));
            if (FS) { W.HC[lane] = __popcll(P.myprow[0]); W.HC[lane + 32] = __popcll(P.myprow[1]); }
            __syncwarp();  // every lane is done with w32 as words: the byte planes go to ys
            unsigned char* wq = reinterpret_cast<unsigned char*>(F.ys);  // [plane][LRB position]
            wq[lane] = (unsigned char)(wa & 0xffu); wq[lane + 32] = (unsigned char)(wb & 0xffu);
            wq[64 + lane] = (unsigned char)(wa >> 8); wq[96 + lane] = (unsigned char)(wb >> 8);
        }
#pragma unroll
        for (int x = 0; x < 8; ++x) {
            // B fragments: column j = 8x + g, this thread's nibbles of P'_j
            const unsigned long long cj = F.prow[8 * x + g];
            W.bfrag[x][lane] = make_uint4(spread4x2((unsigned)cj, 4 * t), spread4x2((unsigned)cj, 4 * t + 16),
                                          spread4x2((unsigned)(cj >> 32), 4 * t), spread4x2((unsigned)(cj >> 32), 4 * t + 16));
            // masks of rows 8x + g (the same nibbles)
            W.maskp[8 * x + g][t] = make_uint4(mask4((unsigned)cj, 4 * t), mask4((unsigned)cj, 4 * t + 16),
                                              mask4((unsigned)(cj >> 32), 4 * t), mask4((unsigned)(cj >> 32), 4 * t + 16));
        }
        __syncwarp();
        const unsigned* wqw = reinterpret_cast<const unsigned*>(F.ys);
        unsigned wr[2][4];  // [plane][{lo 4t, lo 4t+16, hi 4t, hi 4t+16}]: the weight bytes of this thread's k columns
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int q = 0; q < 4; ++q) wr[p][q] = wqw[16 * p + 8 * (q >> 1) + 4 * (q & 1) + t];

        int s[3] = {O3_INF, O3_INF, O3_INF}, sid[3] = {0, 0, 0};
        int gate = O3_INF;  // running warp minimum + OSD_WIN
        int gn = -O3_INF;   // -gate: the tiles hold -S
        const int b32 = F.base32;
        // ---- passes: no third position (pairs, singles, the empty TEP), then k = 2..63 ------------------------
        bool any_stop = false;
#pragma unroll 1
        for (int pass = FS ? 63 : O3_NONE; pass != 1; pass = (pass == O3_NONE ? 63 : pass - 1)) {
            const int kl = pass;  // pairs i < j < kl
            const unsigned long long dk = pass == O3_NONE ? d0 : d0 ^ F.prow[pass];
            const int bk = b32 + (pass == O3_NONE ? 0 : F.qd32[pass]);
            __syncwarp();  // the previous pass is done with R
            if (pass != O3_NONE && lane == 0) {  // j < kl: passes run downwards, every j > kl is already out
                W.C[pass] = -O3_BIG;
                if (FS) W.HC[pass] = O3_PEN;
            }
            {
                const int r0 = bk + F.qd32[lane] + wpop_shfl(tb, dk ^ P.myprow[0]);
                W.R[lane] = -r0;
                if (FS) { W.HR[lane] = __popcll(dk ^ P.myprow[0]); W.HR[lane + 32] = __popcll(dk ^ P.myprow[1]); }
                int r1 = O3_INF;
                if (kl > 32) {  // warp-uniform
                    r1 = bk + F.qd32[lane + 32] + wpop_shfl(tb, dk ^ P.myprow[1]);
                    W.R[lane + 32] = -r1;
                }
                if (pass == O3_NONE) {  // warp-uniform
                    o3_track(s, sid, r0, O3_ID_SINGLE | lane);
                    o3_track(s, sid, r1, O3_ID_SINGLE | (lane + 32));
                    const int ez = bk + wpop_shfl(tb, dk);  // the empty TEP (shuffles are warp-wide: every lane computes it)
                    if (lane == 0) o3_track(s, sid, ez, O3_ID_EMPTY);
                    gate = __reduce_min_sync(0xffffffffu, s[0]) + OSD_WIN;
                    gn = -gate;
                }
            }
            const uint4 md = make_uint4(mask4((unsigned)dk, 4 * t), mask4((unsigned)dk, 4 * t + 16),
                                        mask4((unsigned)(dk >> 32), 4 * t), mask4((unsigned)(dk >> 32), 4 * t + 16));
            __syncwarp();
            const int n_mi = (kl + 14) >> 4;   // row blocks with some i <= kl - 2
            const int n_nj = (kl + 7) >> 3;    // column blocks with some j < kl
#pragma unroll 1
            for (int mi = 0; mi < n_mi; ++mi) {
                const int i0 = 16 * mi + g;
                const uint4 m0 = W.maskp[i0][t], m1 = W.maskp[i0 + 8][t];
                unsigned afr[2][2][4];  // [k half][plane][fragment register]: rows i0 (regs 0, 2) and i0 + 8 (regs 1, 3)
                const unsigned x0[4] = {m0.x ^ md.x, m0.y ^ md.y, m0.z ^ md.z, m0.w ^ md.w};
                const unsigned x1[4] = {m1.x ^ md.x, m1.y ^ md.y, m1.z ^ md.z, m1.w ^ md.w};
#pragma unroll
                for (int kk = 0; kk < 2; ++kk)
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh)
#pragma unroll
                        for (int p = 0; p < 2; ++p) {
                            afr[kk][p][2 * hh] = wr[p][2 * kk + hh] & x0[2 * kk + hh];
                            afr[kk][p][2 * hh + 1] = wr[p][2 * kk + hh] & x1[2 * kk + hh];
                        }
                unsigned afh[2][4];  // FS: unit-weight plane (the mask bits as bytes)
                if (FS) {
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk)
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            afh[kk][2 * hh] = x0[2 * kk + hh] & 0x01010101u;
                            afh[kk][2 * hh + 1] = x1[2 * kk + hh] & 0x01010101u;
                        }
                }
                const int hr0 = FS ? W.HR[i0] : 0, hr1 = FS ? W.HR[i0 + 8] : 0;
                const int rr0 = W.R[i0], rr1 = W.R[i0 + 8];
                // One 16x8 tile.  The B bytes are 0/2, so the two planes accumulate 2 M_lo and 2 M_hi; the low-plane
                // accumulator starts at -(R_i + C_j) -- `pen` lower on the elements with i >= j of the two tiles that touch
                // the diagonal, and columns j >= kl carry the same penalty in C, so the triangle costs no instruction -- and
                // one multiply-add per element, acc0 + 256 acc1, gives -S(i,j).
                auto tile_scores = [&](int nj, const int (&pen)[4], int (&p4)[4]) {
                    const uint4 bf = W.bfrag[nj][lane];
                    const unsigned b0[2] = {bf.x, bf.y}, b1[2] = {bf.z, bf.w};
                    const int2 cc = *reinterpret_cast<const int2*>(W.C + 8 * nj + 2 * t);
                    int acc0[4] = {rr0 + cc.x + pen[0], rr0 + cc.y + pen[1], rr1 + cc.x + pen[2], rr1 + cc.y + pen[3]};
                    int acc1[4] = {0, 0, 0, 0};
                    imma_u8(acc0, afr[0][0], b0);
                    imma_u8(acc1, afr[0][1], b0);
                    imma_u8(acc0, afr[1][0], b1);
                    imma_u8(acc1, afr[1][1], b1);
#pragma unroll
                    for (int e = 0; e < 4; ++e) p4[e] = acc0[e] + 256 * acc1[e];
                    if (FS) {
                        int acch[4] = {pen[0], pen[1], pen[2], pen[3]};  // excluded elements come out 2^29 too far
                        imma_u8(acch, afh[0], b0);  // B bytes are 2: acch = 2 sum u.P_j
                        imma_u8(acch, afh[1], b1);
                        const int2 hc = *reinterpret_cast<const int2*>(W.HC + 8 * nj + 2 * t);
                        const int hrc[4] = {hr0 + hc.x, hr0 + hc.y, hr1 + hc.x, hr1 + hc.y};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int hd = hrc[e] - acch[e];
                            any_stop |= hd < fs.hs;
                            if (hd >= fs.he) p4[e] = -O3_INF;
                        }
                    }
                };
                // warp-uniform gate: only scores within the truncation window of the running warp minimum can matter
                auto offer = [&](int nj, const int (&p4)[4]) {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (p4[e] >= gn) o3_track(s, sid, -p4[e], o3_id(pass, mi, nj, e));
                };
                auto regate = [&]() {
                    const int wm = __reduce_min_sync(0xffffffffu, s[0]);
                    gate = wm > O3_INF - OSD_WIN ? O3_INF : wm + OSD_WIN;
                    gn = -gate;
                };
                auto tile = [&](int nj, const int (&pen)[4]) {
                    int p4[4];
                    tile_scores(nj, pen, p4);
                    const int m4 = max(max(p4[0], p4[1]), max(p4[2], p4[3]));
                    if (__any_sync(0xffffffffu, m4 >= gn)) {
                        offer(nj, p4);
                        regate();
                    }
                };
                // two interior tiles at once: eight independent IMMAs in flight, one gate test for both
                auto tile2 = [&](int nj) {
                    int pa[4], pb[4];
                    tile_scores(nj, pen_none, pa);
                    tile_scores(nj + 1, pen_none, pb);
                    const int m8 = max(max(max(pa[0], pa[1]), max(pa[2], pa[3])), max(max(pb[0], pb[1]), max(pb[2], pb[3])));
                    if (__any_sync(0xffffffffu, m8 >= gn)) {
                        offer(nj, pa);
                        offer(nj + 1, pb);
                        regate();
                    }
                };
                int nj = 2 * mi;
                if (nj < n_nj) tile(nj, pen_d0);
                if (nj + 1 < n_nj) tile(nj + 1, pen_d8);
#pragma unroll 1
                for (nj += 2; nj + 1 < n_nj; nj += 2) tile2(nj);
                if (nj < n_nj) tile(nj, pen_none);
            }
        }
        // ---- candidates inside the truncation window, exact scores ---------------------------------------------------
        int m = __reduce_min_sync(0xffffffffu, s[0]);
        const int lim = m > O3_INF - OSD_WIN ? O3_INF - 1 : m + OSD_WIN;  // m = INF: FS class without an eligible TEP
        // a thread may have dropped a fourth candidate; FS: the sequential loop would have stopped inside this class
        const bool fallback = __any_sync(0xffffffffu, s[2] <= lim) || (FS && __any_sync(0xffffffffu, any_stop));
        long long best_s = 0x7fffffffffffffffll;
        int best_i = 0x7fffffff;
        unsigned best_pos = 0xffffffffu;
        if (!fallback) {
            const long long q_l0 = (long long)F.qlrb[lane], q_l1 = (long long)F.qlrb[lane + 32];
#pragma unroll 1
            for (int r = 0; r < 3; ++r) {
                unsigned cm = __ballot_sync(0xffffffffu, s[r] <= lim);
                while (cm) {
                    const int src = __ffs(cm) - 1;
                    cm &= cm - 1;
                    const int cid = __shfl_sync(0xffffffffu, sid[r], src);
                    unsigned pos;  // up to three MRB positions, one per byte, ascending, 0xFF = unused
                    int ci;
                    if (cid & O3_ID_EMPTY) {
                        pos = 0xffffffffu;
                        ci = (int)a.pair_index[K * K + K];
                    } else if (cid & O3_ID_SINGLE) {
                        const int i = cid & 63;
                        pos = 0xffffff00u | (unsigned)i;
                        ci = (int)a.pair_index[K * K + i];
                    } else {
                        const int pass = cid >> 7, mi = (cid >> 5) & 3, nj = (cid >> 2) & 7, e = cid & 3;
                        const int i = 16 * mi + (src >> 2) + 8 * (e >> 1), j = 8 * nj + 2 * (src & 3) + (e & 1);
                        LDPCB_ASSERT(i >= 0 && i < j && j < K && (pass == O3_NONE || (pass >= 2 && pass < K && j < pass)));
                        if (pass == O3_NONE) {
                            pos = 0xffff0000u | ((unsigned)j << 8) | (unsigned)i;
                            ci = (int)a.pair_index[i * K + j];
                        } else {
                            const int k = pass;
                            pos = 0xff000000u | ((unsigned)k << 16) | ((unsigned)j << 8) | (unsigned)i;
                            ci = (int)triple_index[k * (k - 1) * (k - 2) / 6 + j * (j - 1) / 2 + i];
                        }
                    }
                    LDPCB_ASSERT(ci >= 0 && ci < a.n_teps && __ldg(a.teps + ci) == pos);  // the inverse tables and the enumeration agree
                    unsigned long long D = d0;
                    long long sm = F.base;
#pragma unroll
                    for (int x = 0; x < 3; ++x) {
                        const unsigned tt = (pos >> (8 * x)) & 0xffu;
                        if (tt < 64u) { D ^= F.prow[tt]; sm += F.qd[tt]; }
                    }
                    const long long sl = (((D >> lane) & 1ull) ? q_l0 : 0ll) + (((D >> (lane + 32)) & 1ull) ? q_l1 : 0ll);
                    const long long sc = sm + warp_sum_ll(sl);
                    if (sc < best_s || (sc == best_s && ci < best_i)) { best_s = sc; best_i = ci; best_pos = pos; }
                }
            }
        } else if (lane == 0) {
            const int fp_ = atomicAdd(fb_count, 1);
            LDPCB_ASSERT(fp_ >= 0 && fp_ < a.B + 64);
            fb_list[fp_] = (int32_t)row;
        }
        if (FS && !fallback) {
            // the decision so far (classes 0..2) stands unless the class holds a strictly smaller eligible score (:148-152)
            const long long w0 = fs.wdmin[f];
            if (!(best_s < w0)) {
                best_s = w0;
                best_i = fs.opt[f];
                best_pos = __ldg(a.teps + best_i);
            }
            if (lane == 0) {
                if (fs.num_teps) fs.num_teps[row] = fs.num[f] + 41664;
                if (fs.stop_kind) fs.stop_kind[row] = 3;
            }
        }
        // ---- outputs -----------------------------------------------------------------------------------------------
        if (!fallback) {
            unsigned long long D = d0, flip = 0ull;
#pragma unroll
            for (int x = 0; x < 3; ++x) {
                const unsigned tt = (best_pos >> (8 * x)) & 0xffu;
                if (tt < 64u) { D ^= F.prow[tt]; flip ^= 1ull << tt; }
            }
            const unsigned long long c_lrb = D ^ P.hd_lrb;
            const unsigned long long c_mrb = P.ho_mrb ^ flip;
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
            const unsigned wv = lane == 0 ? wout[0] : lane == 1 ? wout[1] : lane == 2 ? wout[2] : wout[3];
            if (lane < 4 && a.cw_bits) a.cw_bits[orow * 4 + lane] = wv;
            if (a.tally_truth) osd_tally_frame(tally, a, orow, wv, best_i, lane);
            if (lane == 0) {
                if (a.best_tep) a.best_tep[orow] = best_i;
                if (a.best_score_q) a.best_score_q[orow] = best_s;
            }
        }
        if (lane == 0 && a.score_exp) a.score_exp[f] = P.E;
        if (a.perm) {
#pragma unroll
            for (int k = 0; k < 4; ++k) a.perm[f * N + lane + 32 * k] = P.pm[k];
        }
        if (a.redG) {
            a.redG[f * K + lane] = P.myprow[0];
            a.redG[f * K + lane + 32] = P.myprow[1];
        }
    }
    if (a.tally_truth) osd_tally_flush(tally, a, lane);
}

// Exact sweep of the frames osd3_kernel could not decide, by the generic kernel (osd.cu) on the list it left.
int launch_osd_generic3(ldpcb_handle* h, const OsdArgs& a, cudaStream_t st);

int launch_osd3(ldpcb_handle* h, const OsdArgs& a, cudaStream_t st) {
    const uint16_t* triple_index = a.triple_index;
    if (a.B == 0) return LDPCB_OK;
    const int smem = OSD_FPB * (int)sizeof(Osd3Warp);
    int& occ = h->occ[OCC_OSD3];
    if (occ == 0) {
        LDPCB_CUDA(h, cudaFuncSetAttribute(osd3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        LDPCB_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, osd3_kernel<false>, OSD_THREADS, smem));
        if (occ < 1) occ = 1;
    }
    // the list of frames left to the exact sweep: a buffer of its own per caller stream
    const size_t need = sizeof(int32_t) * ((size_t)a.B + 64);
    Workspace& fbw = h->fb_ws[st];
    if (fbw.cap < need) {
        if (fbw.buf) { LDPCB_CUDA(h, cudaDeviceSynchronize()); LDPCB_CUDA(h, cudaFree(fbw.buf)); fbw.buf = nullptr; fbw.cap = 0; }
        LDPCB_CUDA(h, cudaMalloc(&fbw.buf, need + need / 4));
        fbw.cap = need + need / 4;
    }
    int32_t* fb_count = reinterpret_cast<int32_t*>(fbw.buf);
    int32_t* fb_list = fb_count + 64;
    LDPCB_CUDA(h, cudaMemsetAsync(fb_count, 0, sizeof(int32_t), st));
    int64_t want = (a.B + OSD_FPB - 1) / OSD_FPB;
    int64_t cap = (int64_t)h->sm_count * occ;
    int grid = (int)(want < cap ? want : cap);
    if (grid < 1) grid = 1;
    osd3_kernel<false><<<grid, OSD_THREADS, smem, st>>>(a, h->gcol_dev, triple_index, fb_list, fb_count, Fs3Args{});
    LDPCB_LAUNCH_CHECK(h, "osd3_kernel");
    // exact sweep of the undecided frames: they are addressed by their ORIGINAL row, results go to the same places
    OsdArgs b = a;
    b.idx = fb_list; b.count = fb_count;
    b.score_exp = nullptr; b.perm = nullptr; b.redG = nullptr;  // written by osd3_kernel for every frame (indexed by batch position)
    b.pair_index = nullptr;
    b.B = a.B < 4096 ? a.B : 4096;  // grid bound only: the kernel strides over the device-side count
    return launch_osd_generic3(h, b, st);
}

// The weight-3 class of the FS policy for the frames osd_fs_kernel deferred (a.idx / a.count = that list).  Frames the
// sweep cannot decide (a tau_e stop inside the class, window overflow) are appended to (fb_list, fb_count).
int launch_osd3_fs(ldpcb_handle* h, const OsdArgs& a, const Fs3Args& fs, int32_t* fb_list, int32_t* fb_count, cudaStream_t st) {
    const int smem = OSD_FPB * (int)sizeof(Osd3Warp);
    int& occ = h->occ[OCC_OSD3 + 1];
    if (occ == 0) {
        LDPCB_CUDA(h, cudaFuncSetAttribute(osd3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        LDPCB_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, osd3_kernel<true>, OSD_THREADS, smem));
        if (occ < 1) occ = 1;
    }
    int64_t want = (a.B + OSD_FPB - 1) / OSD_FPB;
    int64_t cap = (int64_t)h->sm_count * occ;
    int grid = (int)(want < cap ? want : cap);
    if (grid < 1) grid = 1;
    osd3_kernel<true><<<grid, OSD_THREADS, smem, st>>>(a, h->gcol_dev, a.triple_index, fb_list, fb_count, fs);
    LDPCB_LAUNCH_CHECK(h, "osd3_kernel<FS>");
    return LDPCB_OK;
}

}  // namespace ldpcb

// Block minima of the DL scheme (ldpcb_osd_block_minima): the exact minimum discrepancy of every TEP block ("order
// pattern") of every frame -- acquire_min, DL_OSD_Testing_serial/ordered_statistics_decoding.py:153-162, over the
// blocks generate_teps builds (nn_testing.py:144-157).
//
// Every warp prepares its own frame and walks that frame's blocks itself (no CTA barrier on the common path).  A block
// is swept on the truncated 32-bit scores of the generic sweep (osd.cu): thirteen 5-bit chunk tables of the LRB weights
// held one entry per lane and looked up with shuffles, S32 * 2^30 <= S < (S32 + 69) * 2^30.  The TEPs whose S32 lies
// within OSD_WIN of the block's smallest S32 (almost always one) are re-scored exactly, so the block minimum that is
// written is the exact int64 score, first index on ties, as the byte-LUT sweep of round 1 produced.  Lane b % 32 keeps
// the result of block b: a single candidate is only noted there and scored exactly by that lane when the warp flushes
// 32 blocks (64 broadcast loads of the LRB weights, all lanes at once), several candidates are scored by the whole warp.
// A thread that would have to remember two TEPs inside the window of one block (near-ties: quantised inputs) marks
// the block instead; marked blocks are redone through the exact 64-bit byte LUT, CTA-wide, after the frame round.
#include "common.cuh"
#include "osd_prepare.cuh"
#include "osd_sweep.cuh"

namespace ldpcb {

struct __align__(16) BlocksSmem {
    union {
        unsigned long long lut[8][256];  // exact redo only (after barrier (A): the sweeps are over)
        uint4 comb[OSD_FPB][K + 1];      // sweep: {P' row, floor(qd / 2^30)} of MRB position t in one 16-byte load, [64] = 0
    };
    FrameSm fr[OSD_FPB];
    int fb[OSD_FPB];                 // frame of warp w has marked blocks
};

constexpr long long BLK_MARK = -1;  // exact scores are >= 0

template <int MAXW>
__global__ void __launch_bounds__(OSD_THREADS, 7) osd_blocks_kernel(OsdArgs a, const uint64_t* __restrict__ gcol) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BlocksSmem& S = *reinterpret_cast<BlocksSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    FrameSm& F = S.fr[warp];
    const int64_t nframes = a.count ? (int64_t)*a.count : a.B;
    const bool ties_high = (a.flags & LDPCB_OSD_TIES_HIGH_INDEX_FIRST) != 0;
    const bool disc_from_score = (a.flags & LDPCB_OSD_DISC_HARD_FROM_SCORE) != 0;
    const int nb = a.n_blocks;

    for (int64_t f0 = (int64_t)blockIdx.x * OSD_FPB; f0 < nframes; f0 += (int64_t)gridDim.x * OSD_FPB) {
        // a warp past the end of the list (last round only) redoes the last frame and stores the same values again: no
        // per-warp condition around the body, so the compiler sees converged code
        const int64_t f = f0 + warp < nframes ? f0 + warp : nframes - 1;
        const int64_t row = a.idx ? (int64_t)a.idx[f] : f;
        bool marked = false;
        {
            const Prep P = prepare_frame<true>(a, F, gcol, row, f, lane, ties_high, disc_from_score);
            __syncwarp();
            int tb[13];
            build_shfl_tables(F, lane, tb);
            const unsigned long long d0 = P.d0;
            const int b32 = F.base32;
            uint4* comb = S.comb[warp];
            comb[lane] = make_uint4((unsigned)P.myprow[0], (unsigned)(P.myprow[0] >> 32), (unsigned)F.qd32[lane], 0u);
            comb[lane + 32] = make_uint4((unsigned)P.myprow[1], (unsigned)(P.myprow[1] >> 32), (unsigned)F.qd32[lane + 32], 0u);
            if (lane == 0) comb[K] = make_uint4(0u, 0u, 0u, 0u);
            __syncwarp();
            long long res_s = 0;  // block (blk & ~31) + lane of the current group of 32
            int res_i = 0, cand = -1;
            // exact score of one TEP, every lane its own: the LRB weights are read as broadcasts
            auto exact_own = [&](int ci) {
                const unsigned tw = __ldg(a.teps + ci);
                unsigned long long D = d0;
                long long sm = F.base;
#pragma unroll
                for (int j = 0; j < MAXW; ++j) {
                    const unsigned t = min((tw >> (8 * j)) & 0xffu, 64u);
                    D ^= F.prow[t];
                    sm += F.qd[t];
                }
#pragma unroll
                for (int l2 = 0; l2 < K; l2 += 2) {
                    const ulonglong2 q2 = *reinterpret_cast<const ulonglong2*>(&F.qlrb[l2]);
                    sm += ((D >> l2) & 1ull) ? (long long)q2.x : 0ll;
                    sm += ((D >> (l2 + 1)) & 1ull) ? (long long)q2.y : 0ll;
                }
                return sm;
            };
            int i1 = __ldg(a.block_start);
            for (int blk = 0; blk < nb; ++blk) {
                const int i0 = i1;
                i1 = __ldg(a.block_start + blk + 1);
                LDPCB_ASSERT(i0 >= 0 && i0 <= i1 && i1 <= a.n_teps);
                int s0 = 0x7fffffff, s1 = 0x7fffffff, c0 = 0x7fffffff;
                const uint32_t* tp = a.teps + i0 + lane;
                for (int i = i0 + lane; i - lane < i1; i += 32, tp += 32) {
                    const bool valid = i < i1;
                    const unsigned tw = valid ? __ldg(tp) : 0xffffffffu;
                    unsigned long long D = d0;
                    int s = b32;
#pragma unroll
                    for (int j = 0; j < MAXW; ++j) {
                        const uint4 e = comb[min((tw >> (8 * j)) & 0xffu, 64u)];
                        D ^= ((unsigned long long)e.y << 32) | e.x;
                        s += (int)e.z;
                    }
                    s += wpop_shfl(tb, D);
                    if (!valid) s = 0x7fffffff;
                    const bool lt = s < s0;  // an equal score stays behind the earlier index and counts as a second one
                    s1 = min(s1, lt ? s0 : s);
                    c0 = lt ? i : c0;
                    s0 = min(s0, s);
                }
                long long bs = 0x7fffffffffffffffll;  // warp-uniform: the block's result unless one lane is to score `cd`
                int bi = 0x7fffffff, cd = -1;
                if (i0 < i1) {
                    const int m = __reduce_min_sync(0xffffffffu, s0);
                    const int lim = m + OSD_WIN;  // m <= base32 + 68 * 2^24 < 2^31 - OSD_WIN
                    if (__any_sync(0xffffffffu, s1 <= lim)) {
                        bs = BLK_MARK;
                        marked = true;
                    } else {
                        unsigned cm = __ballot_sync(0xffffffffu, s0 <= lim);
                        if ((cm & (cm - 1)) == 0u) {
                            cd = __shfl_sync(0xffffffffu, c0, __ffs(cm) - 1);
                        } else {
                            const long long q_l0 = (long long)F.qlrb[lane], q_l1 = (long long)F.qlrb[lane + 32];
                            while (cm) {
                                const int src = __ffs(cm) - 1;
                                cm &= cm - 1;
                                const int ci = __shfl_sync(0xffffffffu, c0, src);
                                const unsigned tw = __ldg(a.teps + ci);
                                unsigned long long D = d0;
                                long long sm = F.base;
#pragma unroll
                                for (int j = 0; j < MAXW; ++j) {
                                    const unsigned t = min((tw >> (8 * j)) & 0xffu, 64u);
                                    D ^= F.prow[t];
                                    sm += F.qd[t];
                                }
                                const long long sl = (((D >> lane) & 1ull) ? q_l0 : 0ll) + (((D >> (lane + 32)) & 1ull) ? q_l1 : 0ll);
                                const long long sc = sm + warp_sum_ll(sl);
                                if (sc < bs || (sc == bs && ci < bi)) { bs = sc; bi = ci; }
                            }
                        }
                    }
                }
                if (lane == (blk & 31)) { res_s = bs; res_i = bi; cand = cd; }
                if ((blk & 31) == 31 || blk == nb - 1) {  // flush a group of 32 blocks: coalesced stores
                    if (cand >= 0) { res_s = exact_own(cand); res_i = cand; cand = -1; }
                    const int b = (blk & ~31) + lane;
                    if (b <= blk) {
                        a.block_min_q[f * nb + b] = res_s;
                        if (a.block_arg) a.block_arg[f * nb + b] = res_i;
                    }
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
        if (lane == 0) S.fb[warp] = marked ? 1 : 0;
        __syncthreads();  // (A) the round's frames are swept; marks (global, written by lane 0 of each warp) and flags visible
        // ---- exact redo of the marked blocks, CTA-wide, frame by frame (rare) --------------------------------------
        const int nfr = (int)((nframes - f0) < OSD_FPB ? (nframes - f0) : OSD_FPB);
        for (int w = 0; w < nfr; ++w) {
            if (!S.fb[w]) continue;  // CTA-uniform
            const FrameSm& G = S.fr[w];
            const int64_t fw = f0 + w;
            build_lut64(S.lut, G, tid);
            __syncthreads();
            for (int blk = warp; blk < nb; blk += OSD_FPB) {
                if (a.block_min_q[fw * nb + blk] != BLK_MARK) continue;  // warp-uniform
                const int i0 = a.block_start[blk], i1 = a.block_start[blk + 1];
                long long bs = 0x7fffffffffffffffll;
                int bi = 0x7fffffff;
                for (int i = i0 + lane; i < i1; i += 32) {
                    const long long s = score64<MAXW>(S.lut, G, __ldg(a.teps + i));
                    if (s < bs) { bs = s; bi = i; }
                }
                warp_argmin(bs, bi);
                __syncwarp();  // every lane has read the mark before lane 0 replaces it
                if (lane == 0) {
                    a.block_min_q[fw * nb + blk] = bs;
                    if (a.block_arg) a.block_arg[fw * nb + blk] = bi;
                }
            }
            __syncthreads();  // the LUT is free for the next marked frame
        }
        __syncthreads();  // (B) the next round's prepare overwrites fr[] (read by the other warps in a redo) and fb[]
    }
}

template <int MAXW>
static int launch_blocks_variant(ldpcb_handle* h, const OsdArgs& a, cudaStream_t st) {
    auto kern = osd_blocks_kernel<MAXW>;
    const int smem = (int)sizeof(BlocksSmem);
    int& occ = h->occ[OCC_OSD_BLOCKS + MAXW - 1];
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
    LDPCB_LAUNCH_CHECK(h, "osd_blocks_kernel");
    return LDPCB_OK;
}

int launch_osd_blocks(ldpcb_handle* h, const OsdArgs& a, cudaStream_t st) {
    if (a.B == 0) return LDPCB_OK;
    switch (a.maxw) {
        case 1: return launch_blocks_variant<1>(h, a, st);
        case 2: return launch_blocks_variant<2>(h, a, st);
        case 3: return launch_blocks_variant<3>(h, a, st);
        default: return launch_blocks_variant<4>(h, a, st);
    }
}

}  // namespace ldpcb

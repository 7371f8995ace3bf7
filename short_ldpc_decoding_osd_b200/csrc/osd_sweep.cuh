// Device helpers shared by the OSD sweeps (osd.cu, osd_pair.cu): exact byte-LUT scoring, the 5-bit shuffle tables of
// the truncated scores, and the pieces of the tensor-core pair sweep.
#pragma once
#include "osd_prepare.cuh"

namespace ldpcb {

constexpr int OSD_WIN = 72;       // >= 64 LRB terms + 4 MRB terms + 1 base term
constexpr int OSD_CAND_CAP = 16;

// ---- fused tallies (convention_osd.py:67-75 success test and phase; get_eval-style error counts) ----------------------
struct OsdTally {
    unsigned frames = 0, fe = 0, be = 0, ph[4] = {0u, 0u, 0u, 0u};
};
// wv: this lane's word of the decided codeword (lanes 0..3), trow: row of the transmitted codeword
__device__ __forceinline__ void osd_tally_frame(OsdTally& t, const OsdArgs& a, int64_t trow, unsigned wv, int best_i, int lane) {
    unsigned d = 0;
    if (lane < 4) d = __popc(wv ^ __ldg(a.tally_truth + trow * 4 + lane));
    d = __reduce_add_sync(0xffffffffu, d);
    t.frames += 1;
    t.fe += d != 0;
    t.be += d;
    if (d == 0) {  // weight class of the winning TEP: boundaries 1, 65, 2081 in both enumerations (convention_osd.py:39-47)
        const int w = best_i < 1 ? 0 : best_i < 65 ? 1 : best_i < 2081 ? 2 : 3;
#pragma unroll
        for (int i = 0; i < 4; ++i) t.ph[i] += (w == i);
    }
}
__device__ __forceinline__ void osd_tally_flush(const OsdTally& t, const OsdArgs& a, int lane) {
    if (lane != 0 || t.frames == 0) return;
    unsigned long long* c = reinterpret_cast<unsigned long long*>(a.tally_counters);
    atomicAdd(c + LDPCB_CNT_OSD_FRAMES, (unsigned long long)t.frames);
    atomicAdd(c + LDPCB_CNT_TEPS, (unsigned long long)t.frames * (unsigned long long)a.n_teps);
    if (t.fe) { atomicAdd(c + LDPCB_CNT_OSD_FRAME_ERR, (unsigned long long)t.fe); atomicAdd(c + LDPCB_CNT_FINAL_FRAME_ERR, (unsigned long long)t.fe); }
    if (t.be) { atomicAdd(c + LDPCB_CNT_OSD_BIT_ERR, (unsigned long long)t.be); atomicAdd(c + LDPCB_CNT_FINAL_BIT_ERR, (unsigned long long)t.be); }
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (t.ph[i]) atomicAdd(c + LDPCB_CNT_PHASE0 + i, (unsigned long long)t.ph[i]);
}

// lut[b][x] = sum of q_lrb[8b+i] over the set bits i of x; thread: table b, low nibble fixed
__device__ __forceinline__ void build_lut64(unsigned long long (*lut)[256], const FrameSm& G, int tid) {
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
    for (int x = 0; x < 16; ++x) lut[b][x * 16 + lo] = e[x];
}

// exact score of one TEP through the byte LUTs
template <int MAXW>
__device__ __forceinline__ long long score64(const unsigned long long (*lut)[256], const FrameSm& G, unsigned tw) {
    unsigned long long D = G.d0;
    long long s = G.base;
#pragma unroll
    for (int j = 0; j < MAXW; ++j) {
        const unsigned t = min((tw >> (8 * j)) & 0xffu, 64u);
        D ^= G.prow[t];
        s += G.qd[t];
    }
#pragma unroll
    for (int b = 0; b < 8; ++b) s += (long long)lut[b][(unsigned)(D >> (8 * b)) & 0xffu];
    return s;
}

// ---- tensor-core pair sweep (full order-2 lists) ------------------------------------------------------------------
// With u_i = d0 ^ P'_i the truncated score of the pair TEP {i, j} is
//     S(i,j) = R_i + C_j - 2 * M[i][j],   R_i = base + qd_i + W(u_i)  (= score of the single TEP {i}),
//     C_j = qd_j + W(P'_j),               M[i][j] = sum_l w_l * u_i[l] * P'_j[l]
// (W(x ^ y) = W(x) + W(y) - 2 W(x & y) for a weighted popcount W).  M is a 64x64x64 integer matrix product per
// frame: A[i][l] = w_l masked by bit l of u_i, split into two byte planes (w < 2^16), B[l][j] = bit l of P'_j, both
// u8, accumulated in s32 by mma.sync m16n8k32 (IMMA.16832.U8.U8).  Only the 20 of the 32 16x8 tiles that contain a
// pair i < j are computed.  The 129 values R, C and the empty TEP's score come from the warp's 5-bit shuffle
// tables.  Scores are packed as (S << 7) | code (code = tile and element, or a single / the empty TEP), so a thread
// tracks its minimum and second minimum with three integer min/max per element; the candidates within the
// truncation window of the minimum are re-scored exactly as in the generic sweep (kernel: osd_pair.cu).
constexpr int PAIR_SH = 38;  // osd3.cu: w = floor(q / 2^38) < 2^16: two byte planes; window = 72 * 2^38 ~ 2^-9 of the largest |y|
// osd_pair.cu: w = floor(q / 2^40) < 2^14 so that BOTH planes accumulate into one s32 per element: 2 M = sum (w & 255) * (2 b)
// + sum ((w >> 8) << 2) * (128 b), every operand a u8; window = 72 * 2^40 ~ 2^-7.8 of the largest |y|
constexpr int PAIR1_SH = 40;
__device__ __forceinline__ void imma_u8(int (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// four bits -> four bytes of 0/1 (bit k of the nibble in byte k)
__device__ __forceinline__ unsigned spread4(unsigned word, int sh) { return (((word >> sh) & 0xFu) * 0x00204081u) & 0x01010101u; }
// four bits -> four bytes of 0/2 (bit k of the nibble in byte k): B operand of the low weight plane, << 6 = of the high one
__device__ __forceinline__ unsigned spread4x2(unsigned word, int sh) { return (((word >> sh) & 0xFu) * 0x00408102u) & 0x02020202u; }
// four bits -> four bytes of 0x00/0xFF: the bits are moved to the byte sign positions and replicated by PRMT
__device__ __forceinline__ unsigned mask4(unsigned word, int sh) {
    const unsigned x = ((word >> sh) & 0xFu) * 0x10204080u;
    unsigned r;
    asm("prmt.b32 %0, %1, 0, 0xba98;" : "=r"(r) : "r"(x));  // selector nibble 8+k: replicate the sign of byte k
    return r;
}
// weighted popcount of D through the thirteen 5-bit tables held one entry per lane
__device__ __forceinline__ int wpop_shfl(const int (&tb)[13], unsigned long long D) {
    const unsigned lo = (unsigned)D, hi = (unsigned)(D >> 32);
    int s = __shfl_sync(0xffffffffu, tb[0], lo);  // the source lane is taken modulo 32
    s += __shfl_sync(0xffffffffu, tb[1], lo >> 5);
    s += __shfl_sync(0xffffffffu, tb[2], lo >> 10);
    s += __shfl_sync(0xffffffffu, tb[3], lo >> 15);
    s += __shfl_sync(0xffffffffu, tb[4], lo >> 20);
    s += __shfl_sync(0xffffffffu, tb[5], lo >> 25);
    s += __shfl_sync(0xffffffffu, tb[6], (unsigned)(D >> 30));
    s += __shfl_sync(0xffffffffu, tb[7], hi >> 3);
    s += __shfl_sync(0xffffffffu, tb[8], hi >> 8);
    s += __shfl_sync(0xffffffffu, tb[9], hi >> 13);
    s += __shfl_sync(0xffffffffu, tb[10], hi >> 18);
    s += __shfl_sync(0xffffffffu, tb[11], hi >> 23);
    s += __shfl_sync(0xffffffffu, tb[12], hi >> 28);
    return s;
}
__device__ __forceinline__ void track2(int& s0, int& s1, int p) {
    s1 = min(s1, max(p, s0));
    s0 = min(s0, p);
}


// 5-bit chunk tables of the truncated LRB weights, one entry per lane: tb[j] (lane e) = sum of w32[5j+i] over the set
// bits i of e.  Four weights per (broadcast) load, one multiply-add per term.
__device__ __forceinline__ void build_shfl_tables(const FrameSm& F, int lane, int (&tb)[13]) {
    unsigned lb[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) lb[i] = (lane >> i) & 1u;
#pragma unroll
    for (int j = 0; j < 13; ++j) tb[j] = 0;
#pragma unroll
    for (int v4 = 0; v4 < 16; ++v4) {
        const uint4 wv = reinterpret_cast<const uint4*>(F.w32)[v4];
        const unsigned ww[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int pos = 4 * v4 + u;
            tb[pos / 5] += (int)(lb[pos % 5] * ww[u]);
        }
    }
}

}  // namespace ldpcb

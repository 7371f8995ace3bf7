// Device code shared by the OSD kernels (osd.cu, osd_pb.cu): per-CTA shared-memory layout and the per-frame
// "prepare" steps (reliability sort, GF(2) elimination, P' rows, exact reliabilities).  See osd.cu for the
// reference citations.
#pragma once
#include "common.cuh"

namespace ldpcb {

constexpr int OSD_FPB = 4;  // frames (= warps) per CTA
constexpr int OSD_THREADS = OSD_FPB * 32;

struct __align__(16) FrameSm {
    unsigned long long prow[66];  // P' rows by MRB position, [64] = 0 for padded TEP slots
    long long qd[66];             // signed score delta of flipping MRB position t, [64] = 0
    unsigned long long qlrb[64];  // q of the LRB positions          (qd..qlrb are reused as cols[128])
    unsigned w32[64];             // floor(q_lrb / 2^SH): 32-bit weights of the fast sweeps (SH = 30, pair sweep 38); 16-byte aligned
    float yo[N];                  // ordering metric (original positions); pair sweep: reused as R32[64], C32[64]
    float ys[N];                  // scoring metric; pair sweep: reused as the two byte planes of the weights
    unsigned long long d0;        // order-0 discrepancy on the LRB
    long long base;               // order-0 discrepancy weight on the MRB
    unsigned long long d0m;       // order-0 discrepancy bits on the MRB (0 when both metrics agree)
    int qd32[66];                 // floor(qd / 2^SH), [64] = 0
    int base32;                   // floor(base / 2^SH)
    int pad32;
    unsigned char pi1[N];         // sorted position -> original index
    unsigned char pos[N];         // permuted position (MRB then LRB) -> sorted position; dead after prepare and reused
                                  // as the scatter buffer of the output step
    unsigned char prow_of[K];     // pivot row of MRB position t
};

struct __align__(16) OsdSmem {
    unsigned long long lut[8][256];  // 16 KB, shared by the four frames in turn
    FrameSm fr[OSD_FPB];
    long long red_s[OSD_FPB][OSD_FPB];  // [frame][warp] partial minima
    int red_i[OSD_FPB][OSD_FPB];
    int red32[OSD_FPB];              // fast sweep: per-warp minima of the 32-bit scores
    int cand_n[OSD_FPB];             // fast sweep: candidates whose exact score can still be the minimum
    int cand_ovf[OSD_FPB];
    int cand_i[OSD_FPB][16];
    int red_stop[OSD_FPB];           // FS: per-warp first stopping TEP index
    long long fs_score[OSD_FPB];     // FS results per frame
    int fs_opt[OSD_FPB], fs_num[OSD_FPB], fs_kind[OSD_FPB];
    int tabs[OSD_FPB][13][32];       // fast sweep: 5-bit chunk tables of each frame (entry = lane)
};

__device__ __forceinline__ unsigned long long shfl64(unsigned long long v, int src) {
    unsigned lo = __shfl_sync(0xffffffffu, (unsigned)v, src);
    unsigned hi = __shfl_sync(0xffffffffu, (unsigned)(v >> 32), src);
    return ((unsigned long long)hi << 32) | lo;
}
// Warp reductions on the warp-reduce unit (REDUX: one instruction per 32-bit word instead of five shuffle rounds).
// sum over the warp of values in [0, 2^60): three 20-bit digit sums
__device__ __forceinline__ long long warp_sum_ll(long long v) {
    const unsigned a = (unsigned)v & 0xfffffu, b = (unsigned)(v >> 20) & 0xfffffu, c = (unsigned)(v >> 40);
    const unsigned long long A = __reduce_add_sync(0xffffffffu, a), B = __reduce_add_sync(0xffffffffu, b), C = __reduce_add_sync(0xffffffffu, c);
    return (long long)(A + (B << 20) + (C << 40));
}
__device__ __forceinline__ unsigned long long warp_xor_ull(unsigned long long v) {
    const unsigned lo = __reduce_xor_sync(0xffffffffu, (unsigned)v), hi = __reduce_xor_sync(0xffffffffu, (unsigned)(v >> 32));
    return ((unsigned long long)hi << 32) | lo;
}
// Lexicographic (value, index) minimum over the warp for non-negative values and indices, on the warp-reduce unit
// (three REDUX instead of five rounds of three shuffles): minimum of the high words, of the low words among the lanes
// that hold it, of the indices among the lanes that hold both.
__device__ __forceinline__ void warp_argmin(long long& s, int& i) {
    const unsigned hi = (unsigned)((unsigned long long)s >> 32), lo = (unsigned)s;
    const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
    const unsigned mi = __reduce_min_sync(0xffffffffu, (hi == mh && lo == ml) ? (unsigned)i : 0xffffffffu);
    s = (long long)(((unsigned long long)mh << 32) | ml);
    i = (int)mi;
}
// minimum of non-negative 64-bit values over the warp
__device__ __forceinline__ long long warp_min_ll(long long s) {
    const unsigned hi = (unsigned)((unsigned long long)s >> 32), lo = (unsigned)s;
    const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
    return (long long)(((unsigned long long)mh << 32) | ml);
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

// if (test & bit) { lo ^= xl; hi ^= xh; } as one LOP3-to-predicate and two predicated XORs (the compiler's own choice
// for the C++ form is four selects and two XORs)
__device__ __forceinline__ void xor_if_bit(unsigned& lo, unsigned& hi, unsigned test, unsigned bit, unsigned xl, unsigned xh) {
    asm("{\n\t.reg .pred q;\n\t.reg .b32 t;\n\tand.b32 t, %2, %3;\n\tsetp.ne.u32 q, t, 0;\n\t@q xor.b32 %0, %0, %4;\n\t@q xor.b32 %1, %1, %5;\n\t}"
        : "+r"(lo), "+r"(hi)
        : "r"(test), "r"(bit), "r"(xl), "r"(xh));
}

// 32x32 bit-matrix transpose across the warp: in: lane i holds word x_i; out: bit j of lane i = bit i of x_j
__device__ __forceinline__ unsigned transpose32(unsigned x, int lane) {
    unsigned m = 0x0000ffffu;
#pragma unroll
    for (int j = 16; j; j >>= 1) {
        const unsigned y = __shfl_xor_sync(0xffffffffu, x, j);
        x = (lane & j) ? ((x & ~m) | ((y >> j) & m)) : ((x & m) | ((y & m) << j));
        m ^= m << (j >> 1);
    }
    return x;
}

// ---- 1. sort: descending key, payload = original index; lane holds sorted positions 4*lane + k ------
__device__ __forceinline__ void ce_lane(unsigned& ka, unsigned& ia, unsigned& kb, unsigned& ib, bool asc) {
    const unsigned mn = min(ka, kb), mx = max(ka, kb);
    const unsigned na = asc ? mn : mx, nb = asc ? mx : mn;
    const bool sw = (na != ka);  // equal keys never swap
    const unsigned ta = ia;
    ia = sw ? ib : ia;
    ib = sw ? ta : ib;
    ka = na;
    kb = nb;
}

__device__ __forceinline__ void bitonic_sort_desc(unsigned (&key)[4], unsigned (&idx)[4], int lane) {
#pragma unroll
    for (int kk = 2; kk <= N; kk <<= 1) {
#pragma unroll
        for (int j = kk >> 1; j > 0; j >>= 1) {
            if (j >= 4) {
                const int lm = j >> 2;
                const bool lower = (lane & lm) == 0;
                const bool asc = (kk < N) && ((lane & (kk >> 2)) != 0);  // final merge is descending
                const bool keep_min = (lower == asc);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const unsigned ok = __shfl_xor_sync(0xffffffffu, key[k], lm);
                    const unsigned oi = __shfl_xor_sync(0xffffffffu, idx[k], lm);
                    const unsigned nk = keep_min ? min(key[k], ok) : max(key[k], ok);
                    idx[k] = (nk != key[k]) ? oi : idx[k];  // equal keys stay put on both sides
                    key[k] = nk;
                }
            } else {
                // element index i = 4*lane + k
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if ((k & j) == 0) {
                        const int i_bit = (kk == 2) ? (k & 2) : (kk == 4 ? (lane & 1) : (lane & (kk >> 2)));
                        const bool asc = (kk < N) && (i_bit != 0);
                        ce_lane(key[k], idx[k], key[k | j], idx[k | j], asc);
                    }
                }
            }
        }
    }
}

// The same network on single words (|y| bits with the low 7 bits replaced by the index): one shuffle and one min/max per
// compare-exchange.  The order is exact unless two keys agree in their upper 24 bits, which the caller detects on
// the sorted sequence (such keys end up adjacent) and repairs with the (key, index) sort above.
__device__ __forceinline__ void bitonic_sort_desc_packed(unsigned (&v)[4], int lane) {
#pragma unroll
    for (int kk = 2; kk <= N; kk <<= 1) {
#pragma unroll
        for (int j = kk >> 1; j > 0; j >>= 1) {
            if (j >= 4) {
                const int lm = j >> 2;
                const bool lower = (lane & lm) == 0;
                const bool asc = (kk < N) && ((lane & (kk >> 2)) != 0);
                const bool keep_min = (lower == asc);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const unsigned o = __shfl_xor_sync(0xffffffffu, v[k], lm);
                    v[k] = keep_min ? min(v[k], o) : max(v[k], o);
                }
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if ((k & j) == 0) {
                        const int i_bit = (kk == 2) ? (k & 2) : (kk == 4 ? (lane & 1) : (lane & (kk >> 2)));
                        const bool asc = (kk < N) && (i_bit != 0);
                        const unsigned a = v[k], b = v[k | j];
                        v[k] = asc ? min(a, b) : max(a, b);
                        v[k | j] = asc ? max(a, b) : min(a, b);
                    }
                }
            }
        }
    }
}

// Registers a warp keeps about its frame between prepare and output.
struct Prep {
    unsigned char pm[4];            // original index of permuted positions lane, lane+32 (MRB), lane+64, lane+96 (LRB)
    unsigned long long myprow[2];   // P' rows of MRB positions lane, lane+32
    unsigned long long hd_lrb;      // hard decisions the discrepancy is measured against, LRB part
    unsigned long long hd_mrb;      // same, MRB part
    unsigned long long ho_mrb;      // MRB hard decisions of the ordering metric (order-0 information bits)
    unsigned long long d0;          // order-0 discrepancy on the LRB
    int E;                          // score exponent
};

// Steps 1-4 for one frame by one warp; fills F (prow, qd, qlrb, d0, base) and returns the registers above.
// SH: truncation of the fast sweeps' 32-bit weights, w32 = floor(q / 2^SH)
template <bool TRUTH, int SH = 30>
__device__ __forceinline__ Prep prepare_frame(const OsdArgs& a, FrameSm& F, const uint64_t* __restrict__ gcol, int64_t row,
                                              int64_t f, int lane, bool ties_high, bool disc_from_score) {
    unsigned long long* cols = reinterpret_cast<unsigned long long*>(F.qd);  // [128], dead before qd/qlrb are written
    unsigned char pm[4] = {0, 0, 0, 0};
    unsigned long long myprow[2] = {0ull, 0ull};
    unsigned long long hd_lrb = 0ull, ho_mrb = 0ull, d0 = 0ull;
    int E = 0;
    LDPCB_ASSERT(row >= 0 && f >= 0);
    // ---- load ---------------------------------------------------------------------------
    const float4 v = reinterpret_cast<const float4*>(a.order_llr + row * N)[lane];
    reinterpret_cast<float4*>(F.yo)[lane] = v;
    reinterpret_cast<float4*>(F.ys)[lane] = reinterpret_cast<const float4*>(a.score_llr + row * N)[lane];
    if (a.redG_in) {
        // pre-permuted frame with its systematic generator rows (convention_osd_main's inputs)
        myprow[0] = a.redG_in[row * K + lane];
        myprow[1] = a.redG_in[row * K + lane + 32];
#pragma unroll
        for (int k = 0; k < 4; ++k) pm[k] = (unsigned char)(lane + 32 * k);
        __syncwarp();
    } else {
        // ---- 1. sort --------------------------------------------------------------------
        unsigned idx[4];
        bool ties = false;  // warp-uniform
        {
            unsigned pk[4] = {(__float_as_uint(v.x) & 0x7fffff80u) | (4u * lane), (__float_as_uint(v.y) & 0x7fffff80u) | (4u * lane + 1),
                              (__float_as_uint(v.z) & 0x7fffff80u) | (4u * lane + 2), (__float_as_uint(v.w) & 0x7fffff80u) | (4u * lane + 3)};
            bitonic_sort_desc_packed(pk, lane);
            const unsigned nx = __shfl_down_sync(0xffffffffu, pk[0], 1);
            const bool close = ((pk[0] ^ pk[1]) < 128u) || ((pk[1] ^ pk[2]) < 128u) || ((pk[2] ^ pk[3]) < 128u) || (lane < 31 && (pk[3] ^ nx) < 128u);
            if (__any_sync(0xffffffffu, close)) {  // two keys share their upper 24 bits (~4 % of AWGN frames): exact (key, index) sort
                unsigned key[4] = {__float_as_uint(v.x) & 0x7fffffffu, __float_as_uint(v.y) & 0x7fffffffu,
                                   __float_as_uint(v.z) & 0x7fffffffu, __float_as_uint(v.w) & 0x7fffffffu};
#pragma unroll
                for (int k = 0; k < 4; ++k) idx[k] = 4u * lane + k;
                bitonic_sort_desc(key, idx, lane);
                const unsigned nxt = __shfl_down_sync(0xffffffffu, key[0], 1);
                const bool tie = (key[0] == key[1]) || (key[1] == key[2]) || (key[2] == key[3]) || (lane < 31 && key[3] == nxt);
                ties = __any_sync(0xffffffffu, tie);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) idx[k] = pk[k] & 0x7fu;
            }
        }
        if (ties) {
            // exact rank sort with the tf.argsort tie rule (stable: lower index first; reversed-ascending: higher first)
            __syncwarp();
            unsigned mykey[4];
            int rank[4] = {0, 0, 0, 0};
#pragma unroll
            for (int k = 0; k < 4; ++k) mykey[k] = __float_as_uint(F.yo[4 * lane + k]) & 0x7fffffffu;
            for (int i = 0; i < N; ++i) {
                const unsigned ki = __float_as_uint(F.yo[i]) & 0x7fffffffu;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int j = 4 * lane + k;
                    const bool first = ties_high ? (i > j) : (i < j);
                    rank[k] += (ki > mykey[k]) || (ki == mykey[k] && first);
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) F.pi1[rank[k]] = (unsigned char)(4 * lane + k);
        } else {
            reinterpret_cast<unsigned*>(F.pi1)[lane] = idx[0] | (idx[1] << 8) | (idx[2] << 16) | (idx[3] << 24);
        }
        __syncwarp();
        // ---- 2. GF(2) elimination, column-major; from here on a lane holds the sorted positions lane + 32k --------------
        // Columns as two 32-bit halves (every step works on one half with 32-bit ops).  `info` marks the unit columns of
        // G (gcol[N + j] != 0: one per generator row, the handle picks them): such a column stays a unit vector until
        // another column takes its row as pivot, and while it is one it joins the basis without any row operation.  The
        // scan therefore VISITS only the other ("parity") columns of the 64 most reliable positions -- about half -- and
        // lets them choose their pivot among the rows no unit column of those 64 positions owns (any unused row with a 1
        // gives the same basis and, the reduced matrix being unique for a basis, the same P').  Only a column without a 1
        // there takes the exact set of free rows and, if it takes the row of a later unit column, puts that column on the
        // visit list.  Positions 64.. are scanned one by one as before (a handful until the basis is complete).
        unsigned clo[4], chi[4];
        bool info[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned o = F.pi1[lane + 32 * k];
            const unsigned long long g = gcol[o];
            info[k] = gcol[N + o] != 0ull;
            clo[k] = (unsigned)g;
            chi[k] = (unsigned)(g >> 32);
        }
        unsigned pvb[4] = {0u, 0u, 0u, 0u};  // pivot-row bit (within its 32-bit half) of my column k, 0 = not a pivot
        unsigned pvh = 0u;                   // bit k: that row lies in the high half
        {
            // rows owned by the unit columns of positions 0..63
            const unsigned own_lo = __reduce_or_sync(0xffffffffu, (info[0] ? clo[0] : 0u) | (info[1] ? clo[1] : 0u));
            const unsigned own_hi = __reduce_or_sync(0xffffffffu, (info[0] ? chi[0] : 0u) | (info[1] ? chi[1] : 0u));
            unsigned used_lo = 0u, used_hi = 0u;        // pivot rows of the visited columns
            unsigned av_lo = ~own_lo, av_hi = ~own_hi;  // rows no unit column of positions 0..63 owns, minus the used ones
            unsigned extra1 = 0u;                       // unit columns of positions 32..63 that lost their row
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                unsigned todo = __ballot_sync(0xffffffffu, !info[k]) | (k == 1 ? extra1 : 0u);
                while (todo != 0u) {  // warp-uniform
                    const int ll = __ffs(todo) - 1;
                    todo &= todo - 1u;
                    const unsigned cl = __shfl_sync(0xffffffffu, clo[k], ll);
                    const unsigned ch = __shfl_sync(0xffffffffu, chi[k], ll);
                    unsigned al = cl & av_lo, ah = ch & av_hi;
                    bool exact = false;
                    if ((al | ah) == 0u) {
                        // no 1 in the rows above: the exact free rows = not used and not owned by a unit column at an
                        // earlier position (that column is in the basis on its own row, or lost the row to a used one)
                        unsigned rl = 0u, rh = 0u;
#pragma unroll
                        for (int kk = 0; kk <= k; ++kk) {
                            const bool before = info[kk] && (kk < k || lane < ll);
                            rl |= before ? clo[kk] : 0u;
                            rh |= before ? chi[kk] : 0u;
                        }
                        al = cl & ~(used_lo | __reduce_or_sync(0xffffffffu, rl));
                        ah = ch & ~(used_hi | __reduce_or_sync(0xffffffffu, rh));
                        if ((al | ah) == 0u) continue;  // dependent on more reliable columns
                        exact = true;
                    }
                    unsigned bit_lo = 0u, bit_hi = 0u;
                    if (al != 0u) {
                        bit_lo = al & (0u - al);
                        used_lo |= bit_lo;
                        av_lo &= ~bit_lo;
                        const unsigned ml = cl ^ bit_lo;
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) xor_if_bit(clo[kk], chi[kk], clo[kk], bit_lo, ml, ch);
                        if (lane == ll) pvb[k] = bit_lo;
                    } else {
                        bit_hi = ah & (0u - ah);
                        used_hi |= bit_hi;
                        av_hi &= ~bit_hi;
                        const unsigned mh = ch ^ bit_hi;
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) xor_if_bit(clo[kk], chi[kk], chi[kk], bit_hi, cl, mh);
                        if (lane == ll) { pvb[k] = bit_hi; pvh |= 1u << k; }
                    }
                    if (exact) {
                        // the row may belong to a unit column further down positions 0..63 (the row operation above has
                        // just changed it): it is an ordinary column from now on and gets its visit
#pragma unroll
                        for (int kk = k; kk < 2; ++kk) {
                            const bool lost = info[kk] && (((clo[kk] & bit_lo) | (chi[kk] & bit_hi)) != 0u);
                            const unsigned m = __ballot_sync(0xffffffffu, lost);
                            if (lost) info[kk] = false;
                            if (kk == k) todo |= m; else extra1 |= m;
                        }
                    }
                }
            }
            // the unit columns of positions 0..63 that kept their row: pivots on it
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                if (info[k]) {
                    pvb[k] = clo[k] | chi[k];
                    if (clo[k] == 0u) pvh |= 1u << k;
                }
            }
            // positions 64..127 one by one; every row a unit column above owns is used (by it or by the column that took it)
            used_lo |= own_lo;
            used_hi |= own_hi;
            int npiv = __popc(used_lo) + __popc(used_hi);
#pragma unroll
            for (int k = 2; k < 4; ++k) {
                for (int ll = 0; ll < 32 && npiv < K; ++ll) {
                    const unsigned cl = __shfl_sync(0xffffffffu, clo[k], ll);
                    const unsigned ch = __shfl_sync(0xffffffffu, chi[k], ll);
                    const unsigned al = cl & ~used_lo, ah = ch & ~used_hi;
                    if ((al | ah) == 0u) continue;  // dependent on more reliable columns
                    if (al != 0u) {
                        const unsigned bit = al & (0u - al);
                        used_lo |= bit;
                        const unsigned ml = cl ^ bit;
                        if ((ml | ch) != 0u) {  // a unit column needs no row operation
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) xor_if_bit(clo[kk], chi[kk], clo[kk], bit, ml, ch);
                        }
                        if (lane == ll) pvb[k] = bit;
                    } else {
                        const unsigned bit = ah & (0u - ah);
                        used_hi |= bit;
                        const unsigned mh = ch ^ bit;
                        if ((cl | mh) != 0u) {
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) xor_if_bit(clo[kk], chi[kk], chi[kk], bit, cl, mh);
                        }
                        if (lane == ll) { pvb[k] = bit; pvh |= 1u << k; }
                    }
                    ++npiv;
                }
            }
            // position lists: MRB = pivot columns, LRB = the others, both in scan (reliability) order
            const unsigned lt = (1u << lane) - 1u;
            int base = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const unsigned bal = __ballot_sync(0xffffffffu, pvb[k] != 0u);
                const int r = base + __popc(bal & lt);  // pivots before this position
                const int p = 32 * k + lane;
                if (pvb[k] != 0u) {
                    F.pos[r] = (unsigned char)p;
                    F.prow_of[r] = (unsigned char)(((pvh >> k) & 1u) * 32u + 31u - __clz(pvb[k]));
                } else {
                    F.pos[K + p - r] = (unsigned char)p;
                }
                base += __popc(bal);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) cols[lane + 32 * k] = ((unsigned long long)chi[k] << 32) | clo[k];
        __syncwarp();
        // ---- permutation pi2 o pi1 -----------------------------------------------------------
#pragma unroll
        for (int k = 0; k < 4; ++k) pm[k] = F.pi1[F.pos[lane + 32 * k]];
        // ---- 3. P' rows: 64x64 bit transpose of the LRB columns -------------------------------
        const unsigned long long ca = cols[F.pos[K + lane]];
        const unsigned long long cb = cols[F.pos[K + 32 + lane]];
        const unsigned tA = transpose32((unsigned)ca, lane);          // rows 0..31,  LRB cols 0..31
        const unsigned tB = transpose32((unsigned)(ca >> 32), lane);  // rows 32..63, LRB cols 0..31
        const unsigned tC = transpose32((unsigned)cb, lane);          // rows 0..31,  LRB cols 32..63
        const unsigned tD = transpose32((unsigned)(cb >> 32), lane);  // rows 32..63, LRB cols 32..63
        __syncwarp();  // all reads of cols are done; reuse it for the physical rows
        unsigned long long* rowsP = cols;  // rows in physical order, first 512 B of the free cols area
        rowsP[lane] = ((unsigned long long)tC << 32) | tA;
        rowsP[lane + 32] = ((unsigned long long)tD << 32) | tB;
        __syncwarp();
        myprow[0] = rowsP[F.prow_of[lane]];
        myprow[1] = rowsP[F.prow_of[lane + 32]];
        __syncwarp();  // rowsP dead before qd/qlrb are written below
    }
    // ---- 4. permuted metrics, exact reliabilities, order-0 codeword ----------------------------
    float yo[4], ys[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { yo[k] = F.yo[pm[k]]; ys[k] = F.ys[pm[k]]; }
    float as[4];
    unsigned amax_bits = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        as[k] = score_abs(ys[k]);
        amax_bits = max(amax_bits, __float_as_uint(as[k]));
    }
    amax_bits = __reduce_max_sync(0xffffffffu, amax_bits);
    frexpf(__uint_as_float(amax_bits), &E);
    long long q[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) q[k] = quantize(as[k], E);
    unsigned ho[4], hd[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        ho[k] = !(yo[k] > 0.0f);  // hard decision: 1 iff !(y > 0)   (convention_osd.py:54)
        hd[k] = disc_from_score ? (unsigned)!(ys[k] > 0.0f) : ho[k];
    }
    long long base = 0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const unsigned d0m = ho[k] ^ hd[k];
        const long long qdv = d0m ? -q[k] : q[k];
        F.qd[lane + 32 * k] = qdv;
        F.qd32[lane + 32 * k] = (int)(qdv >> SH);
        base += d0m ? q[k] : 0ll;
    }
    base = warp_sum_ll(base);
    F.qlrb[lane] = (unsigned long long)q[2];
    F.qlrb[lane + 32] = (unsigned long long)q[3];
    F.w32[lane] = (unsigned)(q[2] >> SH);
    F.w32[lane + 32] = (unsigned)(q[3] >> SH);
    F.prow[lane] = myprow[0];
    F.prow[lane + 32] = myprow[1];
    unsigned long long c0 = (ho[0] ? myprow[0] : 0ull) ^ (ho[1] ? myprow[1] : 0ull);
    c0 = warp_xor_ull(c0);
    hd_lrb = (unsigned long long)__ballot_sync(0xffffffffu, hd[2]) | ((unsigned long long)__ballot_sync(0xffffffffu, hd[3]) << 32);
    ho_mrb = (unsigned long long)__ballot_sync(0xffffffffu, ho[0]) | ((unsigned long long)__ballot_sync(0xffffffffu, ho[1]) << 32);
    const unsigned long long hd_mrb_out = (unsigned long long)__ballot_sync(0xffffffffu, hd[0]) | ((unsigned long long)__ballot_sync(0xffffffffu, hd[1]) << 32);
    d0 = c0 ^ hd_lrb;
    if (lane == 0) { F.d0 = d0; F.base = base; F.d0m = ho_mrb ^ hd_mrb_out; F.prow[64] = 0ull; F.qd[64] = 0ll; F.qd32[64] = 0; F.base32 = (int)(base >> SH); }  // [64]: padded TEP slots (qd aliases cols until here)
    if (TRUTH && a.truth_bits && a.truth_score_q) {
        long long ts = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned tb = (a.truth_bits[row * 4 + (pm[k] >> 5)] >> (pm[k] & 31)) & 1u;
            ts += (tb ^ hd[k]) ? q[k] : 0ll;
        }
        ts = warp_sum_ll(ts);
        if (lane == 0) a.truth_score_q[f] = ts;
    }
    Prep r;
#pragma unroll
    for (int k = 0; k < 4; ++k) r.pm[k] = pm[k];
    r.myprow[0] = myprow[0];
    r.myprow[1] = myprow[1];
    r.hd_lrb = hd_lrb;
    r.hd_mrb = hd_mrb_out;
    r.ho_mrb = ho_mrb;
    r.d0 = d0;
    r.E = E;
    return r;
}

}  // namespace ldpcb

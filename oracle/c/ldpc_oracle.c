/*
 * CPU oracle in plain C (TEST INFRASTRUCTURE ONLY) -- a sparse, bit-packed restatement of the
 * reference's NMS and OSD arithmetic, used (a) as the fast checker for large parity batches and
 * (b) as the CPU baseline ("port") that bench.py times on the host cores.  Nothing under
 * short_ldpc_decoding_osd_b200/ links or calls this file.
 *
 * It computes exactly what oracle/nms_oracle.py and oracle/osd_oracle.py compute (tests assert
 * bit-identical outputs), which in turn are pinned to the reference's own source by tests/golden/.
 *
 *   NMS: LDPC_128/Ldpc_128_testing/ms_test.py:106-137 (compute_vc), :180-210 (compute_cv2),
 *        :220-228 (marginalize), :36-54 (hard decision, syndrome)
 *   OSD: LDPC_128/PB_OSD/pb_testing.py:231-266 (full_gf2elim, literal pivot rule), :268-304
 *        (identify_mrb), :306-320 (swapped_info); LDPC_128/FS_OSD/convention_osd.py:49-77
 *
 * Build: make -C oracle/c   (gcc -O2 -fopenmp -ffp-contract=off: no FMA contraction, so every fp32
 * operation rounds exactly like the NumPy restatement).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define N 128
#define M 64
#define K 64

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---------------------------------------------------------------- NMS ---------------------------------------- */
typedef struct {
    int deg[M];
    int var[M][N]; /* variables of check c, ascending */
} graph_t;

static void build_graph(const uint8_t* H, graph_t* g) {
    for (int c = 0; c < M; ++c) {
        g->deg[c] = 0;
        for (int v = 0; v < N; ++v)
            if (H[c * N + v]) g->var[c][g->deg[c]++] = v;
    }
}

/* y [B,128]; hard [B,128] bytes; syn, iters_used [B]; traj [B,iters+1,128] or NULL */
void oracle_nms(const float* y, int64_t B, const uint8_t* H, int iters, float alpha, float w_vc, float w_marg,
                int early_stop, uint8_t* hard, uint8_t* syn, uint8_t* iters_used, float* traj, int threads) {
    graph_t g;
    build_graph(H, &g);
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(static)
    for (int64_t f = 0; f < B; ++f) {
        const float* yf = y + f * N;
        float cv[M][N]; /* only the edges are touched */
        float tot[N], soft[N];
        for (int c = 0; c < M; ++c)
            for (int e = 0; e < g.deg[c]; ++e) cv[c][g.var[c][e]] = 0.0f;
        for (int v = 0; v < N; ++v) soft[v] = yf[v];
        if (traj) memcpy(traj + (f * (iters + 1)) * N, yf, sizeof(float) * N);
        int used = 0;
        for (int it = 0; it < iters; ++it) {
            /* compute_vc: total = sum over checks (ascending) of cv, then + y*w_vc */
            for (int v = 0; v < N; ++v) tot[v] = 0.0f;
            for (int c = 0; c < M; ++c)
                for (int e = 0; e < g.deg[c]; ++e) tot[g.var[c][e]] = tot[g.var[c][e]] + cv[c][g.var[c][e]];
            for (int v = 0; v < N; ++v) tot[v] = tot[v] + yf[v] * w_vc;
            /* compute_cv2 */
            for (int c = 0; c < M; ++c) {
                float x[N];
                float m1 = INFINITY, m2 = INFINITY, sp = 1.0f;
                const int d = g.deg[c];
                for (int e = 0; e < d; ++e) {
                    const int v = g.var[c][e];
                    x[e] = tot[v] - cv[c][v];
                    float a = fabsf(x[e]);
                    if (a > 1e30f) a = 1e30f;
                    if (a < m1) { m2 = m1; m1 = a; } else if (a < m2) m2 = a;
                    sp = sp * (x[e] > 0.0f ? 1.0f : (x[e] < 0.0f ? -1.0f : 0.0f));
                }
                for (int e = 0; e < d; ++e) {
                    const int v = g.var[c][e];
                    float a = fabsf(x[e]);
                    if (a > 1e30f) a = 1e30f;
                    const float mag = (a > m1) ? m1 : m2;
                    const float sg = sp * (x[e] > 0.0f ? 1.0f : (x[e] < 0.0f ? -1.0f : 0.0f));
                    cv[c][v] = (alpha * mag) * sg;
                }
            }
            /* marginalize */
            for (int v = 0; v < N; ++v) soft[v] = 0.0f;
            for (int c = 0; c < M; ++c)
                for (int e = 0; e < g.deg[c]; ++e) soft[g.var[c][e]] = soft[g.var[c][e]] + cv[c][g.var[c][e]];
            for (int v = 0; v < N; ++v) soft[v] = soft[v] + w_marg * yf[v];
            if (traj) memcpy(traj + (f * (iters + 1) + it + 1) * N, soft, sizeof(float) * N);
            used = it + 1;
            if (early_stop) {
                int bad = 0;
                for (int c = 0; c < M && !bad; ++c) {
                    int p = 0;
                    for (int e = 0; e < g.deg[c]; ++e) p ^= !(soft[g.var[c][e]] > 0.0f);
                    bad |= p;
                }
                if (!bad) break;
            }
        }
        if (traj)
            for (int it = used; it < iters; ++it) memcpy(traj + (f * (iters + 1) + it + 1) * N, soft, sizeof(float) * N);
        int bad = 0;
        for (int v = 0; v < N; ++v) hard[f * N + v] = !(soft[v] > 0.0f);
        for (int c = 0; c < M; ++c) {
            int p = 0;
            for (int e = 0; e < g.deg[c]; ++e) p ^= hard[f * N + g.var[c][e]];
            bad |= p;
        }
        if (syn) syn[f] = (uint8_t)bad;
        if (iters_used) iters_used[f] = (uint8_t)used;
    }
}

/* ---------------------------------------------------------------- OSD ---------------------------------------- */
typedef struct { uint64_t w[2]; } row_t; /* 128 packed columns, bit c of w[c>>6] */

static inline int row_get(const row_t* r, int c) { return (int)((r->w[c >> 6] >> (c & 63)) & 1u); }
static inline void row_put(row_t* r, int c, int b) {
    r->w[c >> 6] = (r->w[c >> 6] & ~(1ull << (c & 63))) | ((uint64_t)(b & 1) << (c & 63));
}

typedef struct { uint32_t key; int idx; } sk_t;
static int cmp_desc_low(const void* a, const void* b) { /* descending key, lower index first */
    const sk_t *x = a, *y = b;
    if (x->key != y->key) return x->key > y->key ? -1 : 1;
    return x->idx - y->idx;
}
static int cmp_desc_high(const void* a, const void* b) { /* reverse of ascending-stable: higher index first */
    const sk_t *x = a, *y = b;
    if (x->key != y->key) return x->key > y->key ? -1 : 1;
    return y->idx - x->idx;
}

static int cmp_int(const void* a, const void* b) { return *(const int*)a - *(const int*)b; }

#define FLAG_TIES_HIGH 1
#define FLAG_DISC_FROM_SCORE 2

/*
 * yo/ys [B,128]; G [64,128] bytes; teps packed uint32 (byte i = MRB position, 0xFF unused);
 * block_start (n_blocks+1) or NULL; outputs may be NULL: cw [B,128] bytes, best_tep, best_q, score_exp [B],
 * perm [B,128], redG [B,64] (P' words), block_min [B,n_blocks], block_arg, truth [B,128] bytes in, truth_q [B].
 */
void oracle_osd(const float* yo_all, const float* ys_all, int64_t B, const uint8_t* G, const uint32_t* teps, int n_teps,
                const int32_t* block_start, int n_blocks, int flags, uint8_t* cw, int32_t* best_tep, int64_t* best_q,
                int32_t* score_exp, uint8_t* perm_out, uint64_t* redG, int64_t* block_min, int32_t* block_arg,
                const uint8_t* truth, int64_t* truth_q, int threads) {
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t f = 0; f < B; ++f) {
        const float* yo = yo_all + f * N;
        const float* ys = ys_all + f * N;
        /* swapped_info: pi1 */
        sk_t sk[N];
        for (int j = 0; j < N; ++j) {
            uint32_t bits;
            float a = fabsf(yo[j]);
            memcpy(&bits, &a, 4);
            sk[j].key = bits;
            sk[j].idx = j;
        }
        qsort(sk, N, sizeof(sk_t), (flags & FLAG_TIES_HIGH) ? cmp_desc_high : cmp_desc_low);
        int pi1[N];
        for (int c = 0; c < N; ++c) pi1[c] = sk[c].idx;
        /* order_G, bit-packed rows */
        row_t A[K];
        for (int r = 0; r < K; ++r) {
            A[r].w[0] = A[r].w[1] = 0;
            for (int c = 0; c < N; ++c) row_put(&A[r], c, G[r * N + pi1[c]]);
        }
        /* full_gf2elim, literal rule (G has full rank, so no row is ever deleted) */
        int index_order[N];
        for (int c = 0; c < N; ++c) index_order[c] = c;
        for (int i = 0; i < K; ++i) {
            const int j = i;
            int k = -1;
            for (int r = i; r < K; ++r)
                if (row_get(&A[r], j)) { k = r; break; }
            if (k >= 0) {
                if (k != i) { row_t t = A[k]; A[k] = A[i]; A[i] = t; }
            } else {
                int ck = -1;
                for (int c = j; c < N; ++c)
                    if (row_get(&A[i], c)) { ck = c; break; }
                /* ck >= 0 because rank(G) = K */
                for (int r = 0; r < K; ++r) {
                    const int a = row_get(&A[r], j), b = row_get(&A[r], ck);
                    row_put(&A[r], j, b);
                    row_put(&A[r], ck, a);
                }
                const int t = index_order[j]; index_order[j] = index_order[ck]; index_order[ck] = t;
            }
            for (int r = 0; r < K; ++r)
                if (r != i && row_get(&A[r], j)) { A[r].w[0] ^= A[i].w[0]; A[r].w[1] ^= A[i].w[1]; }
        }
        /* identify_mrb: sort both halves ascending, permute rows and columns accordingly */
        int mrb[K], lrb[K], mrb_sorted[K], lrb_sorted[K];
        for (int t = 0; t < K; ++t) { mrb[t] = index_order[t]; lrb[t] = index_order[K + t]; }
        memcpy(mrb_sorted, mrb, sizeof mrb);
        memcpy(lrb_sorted, lrb, sizeof lrb);
        qsort(mrb_sorted, K, sizeof(int), cmp_int);
        qsort(lrb_sorted, K, sizeof(int), cmp_int);
        int row_of[K]; /* logical MRB position t (ascending) -> row of A whose pivot is that column */
        for (int t = 0; t < K; ++t)
            for (int r = 0; r < K; ++r)
                if (mrb[r] == mrb_sorted[t]) row_of[t] = r;
        int lcol_of[K]; /* logical LRB position l -> column of A (K + place in lrb[]) */
        for (int l = 0; l < K; ++l)
            for (int c = 0; c < K; ++c)
                if (lrb[c] == lrb_sorted[l]) lcol_of[l] = K + c;
        uint64_t prow[K];
        for (int t = 0; t < K; ++t) {
            uint64_t w = 0;
            for (int l = 0; l < K; ++l) w |= (uint64_t)row_get(&A[row_of[t]], lcol_of[l]) << l;
            prow[t] = w;
        }
        int perm[N];
        for (int t = 0; t < K; ++t) { perm[t] = pi1[mrb_sorted[t]]; perm[K + t] = pi1[lrb_sorted[t]]; }
        /* exact integer reliabilities */
        float amax = 0.0f, as[N];
        for (int t = 0; t < N; ++t) {
            float a = fabsf(ys[perm[t]]);
            if (a != a) a = 0.0f;
            if (a > 3.402823466e38f) a = 3.402823466e38f;
            as[t] = a;
            if (a > amax) amax = a;
        }
        int E = 0;
        frexpf(amax, &E);
        int64_t q[N];
        for (int t = 0; t < N; ++t) q[t] = (int64_t)rint(ldexp((double)as[t], 54 - E));
        uint64_t ho_mrb = 0, hd_mrb = 0, hd_lrb = 0;
        for (int t = 0; t < K; ++t) {
            const uint64_t ho = !(yo[perm[t]] > 0.0f);
            const uint64_t hd = (flags & FLAG_DISC_FROM_SCORE) ? (uint64_t)!(ys[perm[t]] > 0.0f) : ho;
            ho_mrb |= ho << t;
            hd_mrb |= hd << t;
            const uint64_t hl = (flags & FLAG_DISC_FROM_SCORE) ? (uint64_t)!(ys[perm[K + t]] > 0.0f) : (uint64_t)!(yo[perm[K + t]] > 0.0f);
            hd_lrb |= hl << t;
        }
        uint64_t c0 = 0;
        for (int t = 0; t < K; ++t)
            if ((ho_mrb >> t) & 1) c0 ^= prow[t];
        /* sweep */
        int64_t best = INT64_MAX;
        int besti = -1;
        const int nb = block_start ? n_blocks : 1;
        for (int b = 0; b < nb; ++b) {
            const int i0 = block_start ? block_start[b] : 0, i1 = block_start ? block_start[b + 1] : n_teps;
            int64_t bm = INT64_MAX;
            int ba = -1;
            for (int i = i0; i < i1; ++i) {
                uint64_t lrbw = c0, mrbw = ho_mrb;
                for (int jx = 0; jx < 4; ++jx) {
                    const unsigned t = (teps[i] >> (8 * jx)) & 0xffu;
                    if (t < K) { lrbw ^= prow[t]; mrbw ^= 1ull << t; }
                }
                uint64_t dl = lrbw ^ hd_lrb, dm = mrbw ^ hd_mrb;
                int64_t s = 0;
                while (dm) { s += q[__builtin_ctzll(dm)]; dm &= dm - 1; }
                while (dl) { s += q[K + __builtin_ctzll(dl)]; dl &= dl - 1; }
                if (s < bm) { bm = s; ba = i; }
            }
            if (block_min) block_min[f * nb + b] = bm;
            if (block_arg) block_arg[f * nb + b] = ba;
            if (bm < best) { best = bm; besti = ba; }
        }
        /* outputs */
        if (cw) {
            uint64_t lrbw = c0, mrbw = ho_mrb;
            if (besti >= 0)
                for (int jx = 0; jx < 4; ++jx) {
                    const unsigned t = (teps[besti] >> (8 * jx)) & 0xffu;
                    if (t < K) { lrbw ^= prow[t]; mrbw ^= 1ull << t; }
                }
            for (int t = 0; t < K; ++t) {
                cw[f * N + perm[t]] = (uint8_t)((mrbw >> t) & 1);
                cw[f * N + perm[K + t]] = (uint8_t)((lrbw >> t) & 1);
            }
        }
        if (best_tep) best_tep[f] = besti;
        if (best_q) best_q[f] = best;
        if (score_exp) score_exp[f] = E;
        if (perm_out) for (int t = 0; t < N; ++t) perm_out[f * N + t] = (uint8_t)perm[t];
        if (redG) memcpy(redG + f * K, prow, sizeof prow);
        if (truth && truth_q) {
            int64_t s = 0;
            for (int t = 0; t < N; ++t) {
                const int hd = t < K ? (int)((hd_mrb >> t) & 1) : (int)((hd_lrb >> (t - K)) & 1);
                if (truth[f * N + perm[t]] ^ hd) s += q[t];
            }
            truth_q[f] = s;
        }
    }
}

// Small data-parallel kernels around the two decoders: FER/BER tallies, stable compaction of the
// detected failures, row gather, DIA FIR.
#include <cstring>

#include "common.cuh"
#include "internal.cuh"

namespace ldpcb {

// ---- tallies -------------------------------------------------------------------------------------
// get_eval (Ldpc_128_testing/ms_test.py:36-54): frame/bit errors vs labels, detected (syndrome != 0),
// undetected (syndrome == 0 but wrong).
constexpr int TALLY_THREADS = 256;

__device__ __forceinline__ void block_add(unsigned long long* sh, int slot, unsigned long long v) {
#pragma unroll
    for (int m = 16; m; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&sh[slot], v);
}

__device__ __forceinline__ int bit_diff(const uint4& a, const uint4& b) {
    return __popc(a.x ^ b.x) + __popc(a.y ^ b.y) + __popc(a.z ^ b.z) + __popc(a.w ^ b.w);
}

__global__ void __launch_bounds__(TALLY_THREADS) tally_nms_kernel(const uint4* __restrict__ bits, const uint8_t* __restrict__ syn,
                                                                   const uint8_t* __restrict__ iters, const uint4* __restrict__ truth,
                                                                   int64_t B, unsigned long long* counters) {
    __shared__ unsigned long long sh[LDPCB_NUM_COUNTERS];
    if (threadIdx.x < LDPCB_NUM_COUNTERS) sh[threadIdx.x] = 0ull;
    __syncthreads();
    unsigned long long fe = 0, be = 0, det = 0, und = 0, its = 0, frames = 0;
    for (int64_t f = (int64_t)blockIdx.x * TALLY_THREADS + threadIdx.x; f < B; f += (int64_t)gridDim.x * TALLY_THREADS) {
        const int d = bit_diff(bits[f], truth[f]);
        const int nz = syn ? syn[f] : 0;
        frames += 1;
        fe += d != 0;
        be += d;
        det += nz != 0;
        und += (nz == 0 && d != 0);
        its += iters ? iters[f] : 0;
    }
    block_add(sh, LDPCB_CNT_FRAMES, frames);
    block_add(sh, LDPCB_CNT_NMS_FRAME_ERR, fe);
    block_add(sh, LDPCB_CNT_NMS_BIT_ERR, be);
    block_add(sh, LDPCB_CNT_NMS_DETECTED, det);
    block_add(sh, LDPCB_CNT_NMS_UNDETECTED, und);
    block_add(sh, LDPCB_CNT_NMS_ITERS, its);
    __syncthreads();
    if (threadIdx.x < LDPCB_NUM_COUNTERS && sh[threadIdx.x]) atomicAdd(&counters[threadIdx.x], sh[threadIdx.x]);
}

// After OSD: frames with a non-zero NMS syndrome went through OSD (convention_osd.py:67-75).
__global__ void __launch_bounds__(TALLY_THREADS) tally_final_kernel(const uint4* __restrict__ bits, const uint8_t* __restrict__ syn,
                                                                     const int32_t* __restrict__ best_tep, int n_teps, int b1, int b2, int b3,
                                                                     const uint4* __restrict__ truth, int64_t B, unsigned long long* counters) {
    __shared__ unsigned long long sh[LDPCB_NUM_COUNTERS];
    if (threadIdx.x < LDPCB_NUM_COUNTERS) sh[threadIdx.x] = 0ull;
    __syncthreads();
    unsigned long long of = 0, ofe = 0, obe = 0, ffe = 0, fbe = 0, ph[4] = {0, 0, 0, 0};
    for (int64_t f = (int64_t)blockIdx.x * TALLY_THREADS + threadIdx.x; f < B; f += (int64_t)gridDim.x * TALLY_THREADS) {
        const int d = bit_diff(bits[f], truth[f]);
        const int nz = syn ? syn[f] : 0;
        ffe += d != 0;
        fbe += d;
        if (nz) {
            of += 1;
            ofe += d != 0;
            obe += d;
            if (d == 0 && best_tep) {
                const int i = best_tep[f];
                const int w = i < b1 ? 0 : i < b2 ? 1 : i < b3 ? 2 : 3;
                ph[w] += 1;
            }
        }
    }
    block_add(sh, LDPCB_CNT_OSD_FRAMES, of);
    block_add(sh, LDPCB_CNT_OSD_FRAME_ERR, ofe);
    block_add(sh, LDPCB_CNT_OSD_BIT_ERR, obe);
    block_add(sh, LDPCB_CNT_FINAL_FRAME_ERR, ffe);
    block_add(sh, LDPCB_CNT_FINAL_BIT_ERR, fbe);
    block_add(sh, LDPCB_CNT_TEPS, of * (unsigned long long)n_teps);
#pragma unroll
    for (int w = 0; w < 4; ++w) block_add(sh, LDPCB_CNT_PHASE0 + w, ph[w]);
    __syncthreads();
    if (threadIdx.x < LDPCB_NUM_COUNTERS && sh[threadIdx.x]) atomicAdd(&counters[threadIdx.x], sh[threadIdx.x]);
}

static int tally_grid(ldpcb_handle* h, int64_t B) {
    int64_t want = (B + TALLY_THREADS - 1) / TALLY_THREADS;
    int64_t cap = (int64_t)h->sm_count * 8;
    return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

int launch_tally_nms(ldpcb_handle* h, const uint32_t* bits, const uint8_t* syn, const uint8_t* iters,
                     const uint32_t* truth, int64_t B, uint64_t* counters, cudaStream_t st) {
    if (B == 0) return LDPCB_OK;
    tally_nms_kernel<<<tally_grid(h, B), TALLY_THREADS, 0, st>>>((const uint4*)bits, syn, iters, (const uint4*)truth, B,
                                                                 (unsigned long long*)counters);
    LDPCB_LAUNCH_CHECK(h, "tally_nms_kernel");
    return LDPCB_OK;
}

int launch_tally_final(ldpcb_handle* h, const uint32_t* bits, const uint8_t* syn, const int32_t* best_tep,
                       int osd_order, int tep_order, const uint32_t* truth, int64_t B, uint64_t* counters,
                       cudaStream_t st) {
    if (B == 0) return LDPCB_OK;
    const int n_teps = osd_order >= 0 ? h->tep[osd_order][tep_order].n : 0;
    // weight-class boundaries are the same for both enumerations: 1, 65, 2081 (convention_osd.py:39-47)
    tally_final_kernel<<<tally_grid(h, B), TALLY_THREADS, 0, st>>>((const uint4*)bits, syn, best_tep, n_teps, 1, 65, 2081,
                                                                   (const uint4*)truth, B, (unsigned long long*)counters);
    LDPCB_LAUNCH_CHECK(h, "tally_final_kernel");
    return LDPCB_OK;
}

// ---- stable compaction of flagged frames -----------------------------------------------------------
// index = tf.where(syndrome != 0) (ms_test.py:51) keeps frame order; three passes: per-tile counts,
// one-block exclusive scan of the tile counts, per-tile scatter.
constexpr int SEL_THREADS = 256;
constexpr int SEL_PER_THREAD = 8;
constexpr int SEL_TILE = SEL_THREADS * SEL_PER_THREAD;

__device__ __forceinline__ int block_excl_scan(int v, int* total) {
    __shared__ int wsum[SEL_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    int woff = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < SEL_THREADS / 32; ++w) {
        const int s = wsum[w];
        if (w < warp) woff += s;
        tot += s;
    }
    __syncthreads();
    *total = tot;
    return woff + incl - v;
}

__device__ __forceinline__ int tile_count(const uint8_t* flags, int64_t B, int64_t base, unsigned* mask) {
    int c = 0;
    unsigned m = 0;
#pragma unroll
    for (int i = 0; i < SEL_PER_THREAD; ++i) {
        const int64_t f = base + i;
        const int on = (f < B) && flags[f] != 0;
        m |= (unsigned)on << i;
        c += on;
    }
    *mask = m;
    return c;
}

__global__ void __launch_bounds__(SEL_THREADS) sel_count_kernel(const uint8_t* __restrict__ flags, int64_t B, int* tile_counts) {
    const int64_t base = ((int64_t)blockIdx.x * SEL_THREADS + threadIdx.x) * SEL_PER_THREAD;
    unsigned m;
    int c = tile_count(flags, B, base, &m), tot;
    block_excl_scan(c, &tot);
    if (threadIdx.x == 0) tile_counts[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(SEL_THREADS) sel_scan_kernel(int* tile_counts, int ntiles, int32_t* count_out) {
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < ntiles; base += SEL_THREADS) {
        const int i = base + threadIdx.x;
        const int v = i < ntiles ? tile_counts[i] : 0;
        int tot;
        const int ex = block_excl_scan(v, &tot);
        const int c0 = carry;
        if (i < ntiles) tile_counts[i] = c0 + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry = c0 + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *count_out = carry;
}

__global__ void __launch_bounds__(SEL_THREADS) sel_scatter_kernel(const uint8_t* __restrict__ flags, int64_t B,
                                                                   const int* __restrict__ tile_offsets, int32_t* idx) {
    const int64_t base = ((int64_t)blockIdx.x * SEL_THREADS + threadIdx.x) * SEL_PER_THREAD;
    unsigned m;
    int c = tile_count(flags, B, base, &m), tot;
    int o = tile_offsets[blockIdx.x] + block_excl_scan(c, &tot);
#pragma unroll
    for (int i = 0; i < SEL_PER_THREAD; ++i)
        if ((m >> i) & 1u) idx[o++] = (int32_t)(base + i);
}

size_t select_temp_bytes(int64_t B) { return sizeof(int) * (size_t)((B + SEL_TILE - 1) / SEL_TILE + 1); }

int launch_select(ldpcb_handle* h, const uint8_t* flags, int64_t B, int32_t* idx, int32_t* count, void* temp, cudaStream_t st) {
    if (B == 0) {
        LDPCB_CUDA(h, cudaMemsetAsync(count, 0, sizeof(int32_t), st));
        return LDPCB_OK;
    }
    const int ntiles = (int)((B + SEL_TILE - 1) / SEL_TILE);
    int* tc = (int*)temp;
    sel_count_kernel<<<ntiles, SEL_THREADS, 0, st>>>(flags, B, tc);
    LDPCB_LAUNCH_CHECK(h, "sel_count_kernel");
    sel_scan_kernel<<<1, SEL_THREADS, 0, st>>>(tc, ntiles, count);
    LDPCB_LAUNCH_CHECK(h, "sel_scan_kernel");
    sel_scatter_kernel<<<ntiles, SEL_THREADS, 0, st>>>(flags, B, tc, idx);
    LDPCB_LAUNCH_CHECK(h, "sel_scatter_kernel");
    return LDPCB_OK;
}

// ---- row gather ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gather_rows_kernel(const float4* __restrict__ src, const int32_t* __restrict__ idx,
                                                          const int32_t* __restrict__ count, int64_t max_rows, int row_f4, float4* dst) {
    const int64_t n = count ? (int64_t)*count : max_rows;
    const int64_t total = n * row_f4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / row_f4;
        const int c = (int)(i - r * row_f4);
        dst[i] = src[(int64_t)idx[r] * row_f4 + c];
    }
}

// ---- DIA FIR ---------------------------------------------------------------------------------------
// conv_bitwise (DL_OSD_Testing_serial/nn_net.py:174-197) folded into taps + bias on the host.
constexpr int FIR_MAX_ROWS = LDPCB_MAX_ITERS + 1;
struct FirTaps { float t[FIR_MAX_ROWS]; };

__global__ void __launch_bounds__(256) dia_fir_kernel(const float4* __restrict__ traj, int64_t B, int n_rows, FirTaps taps, float bias, float4* out) {
    const int64_t total = B * (N / 4);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / (N / 4);
        const int c = (int)(i - b * (N / 4));
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < n_rows; ++r) {
            const float4 v = traj[(b * n_rows + r) * (N / 4) + c];
            const float t = taps.t[r];
            acc.x = __fmaf_rn(t, v.x, acc.x);
            acc.y = __fmaf_rn(t, v.y, acc.y);
            acc.z = __fmaf_rn(t, v.z, acc.z);
            acc.w = __fmaf_rn(t, v.w, acc.w);
        }
        acc.x += bias; acc.y += bias; acc.z += bias; acc.w += bias;
        out[i] = acc;
    }
}


// ---- DL sliding-window early termination -----------------------------------------------------------------
// osd.sliding_osd's window bookkeeping (DL_OSD_Testing_serial/ordered_statistics_decoding.py:186-219) with
// sliding_window_ops (:141-151) and the 6 -> 6 -> 2 classifier Predict_outlier_light (nn_net.py:136-149), one
// thread per frame over the <= 128 block minima the sweep kernel produced.  Minima are compared as exact integers;
// the classifier sees them as fp32 (q * 2^(E-54)).
constexpr int DLW_MAX_WIDTH = 8;
constexpr int DLW_MAX_BLOCKS = 128;
struct DlwParams {
    float W1[(DLW_MAX_WIDTH + 1) * (DLW_MAX_WIDTH + 1)];
    float W2[(DLW_MAX_WIDTH + 1) * 2];
    int acc[DLW_MAX_BLOCKS + 1];
    float soft_margin;
    int width, n_blocks;
};

__global__ void __launch_bounds__(256) dl_window_kernel(const long long* __restrict__ bm, const int* __restrict__ ex,
                                                        const long long* __restrict__ truth, int64_t B, DlwParams p,
                                                        uint8_t* success, int* windows, int* complexity,
                                                        unsigned long long* counters) {
    __shared__ unsigned long long sh[4];
    if (threadIdx.x < 4) sh[threadIdx.x] = 0ull;
    __syncthreads();
    unsigned long long c_s = 0, c_f = 0, c_w = 0, c_c = 0;
    const int W = p.width, nb = p.n_blocks, in_w = W + 1;
    for (int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; f < B; f += (int64_t)gridDim.x * blockDim.x) {
        const long long* m = bm + f * nb;
        const double scale = __hiloint2double((1023 + ex[f] - 54) << 20, 0);
        long long win[DLW_MAX_WIDTH];
        long long gmin = 0x7fffffffffffffffll;
        for (int j = 0; j < W; ++j) { win[j] = m[j]; gmin = win[j] < gmin ? win[j] : gmin; }
        int deep = W;
        for (int k = 0; k < nb - W + 1; ++k) {
            deep = k + W;
            if (k != 0) {
                const long long ms = m[W + k - 1];
                for (int j = 0; j + 1 < W; ++j) win[j] = win[j + 1];
                win[W - 1] = ms;
                if (ms > gmin) continue;
            }
            // sorted window (ascending) + position k -> classifier
            float x[DLW_MAX_WIDTH + 1];
            long long srt[DLW_MAX_WIDTH];
            for (int j = 0; j < W; ++j) {
                long long v = win[j];
                int q = j;
                while (q > 0 && srt[q - 1] > v) { srt[q] = srt[q - 1]; --q; }
                srt[q] = v;
            }
            for (int j = 0; j < W; ++j) x[j] = (float)((double)srt[j] * scale);
            x[W] = (float)k;
            float o0 = 0.0f, o1 = 0.0f;
            for (int u = 0; u < in_w; ++u) {
                float hsum = 0.0f;
                for (int j = 0; j < in_w; ++j) hsum = __fmaf_rn(x[j], p.W1[j * in_w + u], hsum);
                o0 = __fmaf_rn(hsum, p.W2[u * 2 + 0], o0);
                o1 = __fmaf_rn(hsum, p.W2[u * 2 + 1], o1);
            }
            const float mx = fmaxf(o0, o1);
            const float e0 = expf(o0 - mx), e1 = expf(o1 - mx);
            const float p1 = e1 / (e0 + e1);
            if (srt[0] < gmin) gmin = srt[0];
            if (p1 > p.soft_margin) break;
        }
        const int nwin = deep - W + 1;
        const int cx = p.acc[deep];
        const bool ok = truth ? (gmin == truth[f]) : false;
        if (success) success[f] = ok ? 1 : 0;
        if (windows) windows[f] = nwin;
        if (complexity) complexity[f] = cx;
        c_s += ok; c_f += !ok; c_w += nwin; c_c += cx;
    }
    if (counters) {
        block_add(sh, 0, c_s); block_add(sh, 1, c_f); block_add(sh, 2, c_w); block_add(sh, 3, c_c);
        __syncthreads();
        if (threadIdx.x < 4 && sh[threadIdx.x]) atomicAdd(&counters[threadIdx.x], sh[threadIdx.x]);
    }
}

}  // namespace ldpcb

using namespace ldpcb;

extern "C" int ldpcb_tally(ldpcb_t* h, const uint32_t* nms_bits_dev, const uint8_t* syndrome_nz_dev,
                           const uint8_t* iters_used_dev, const uint32_t* final_bits_dev, const int32_t* best_tep_dev,
                           int osd_order, int tep_order, const uint32_t* truth_bits_dev, int64_t B,
                           uint64_t* counters_dev, void* stream) {
    LDPCB_ENTER(h);
    if (B < 0 || !truth_bits_dev || !counters_dev || (!nms_bits_dev && !final_bits_dev))
        return set_error(h, LDPCB_ERR_ARG, "ldpcb_tally: bad arguments");
    if (final_bits_dev && (osd_order < -1 || osd_order > 3 || tep_order < 0 || tep_order > 1))
        return set_error(h, LDPCB_ERR_ARG, "ldpcb_tally: osd_order %d / tep_order %d out of range", osd_order, tep_order);
    int st = LDPCB_OK;
    if (nms_bits_dev) st = launch_tally_nms(h, nms_bits_dev, syndrome_nz_dev, iters_used_dev, truth_bits_dev, B, counters_dev, (cudaStream_t)stream);
    if (st == LDPCB_OK && final_bits_dev)
        st = launch_tally_final(h, final_bits_dev, syndrome_nz_dev, best_tep_dev, osd_order, tep_order, truth_bits_dev, B, counters_dev, (cudaStream_t)stream);
    return st;
}

extern "C" int ldpcb_select_flagged(ldpcb_t* h, const uint8_t* flags_dev, int64_t B, int32_t* idx_dev,
                                    int32_t* count_dev, void* stream) {
    LDPCB_ENTER(h);
    if (B < 0 || B > 0x7fffffff || !count_dev || (B > 0 && (!flags_dev || !idx_dev)))
        return set_error(h, LDPCB_ERR_ARG, "ldpcb_select_flagged: bad arguments");
    char* temp = nullptr;
    int st = ensure_stream_ws(h, (cudaStream_t)stream, select_temp_bytes(B), &temp);
    if (st != LDPCB_OK) return st;
    return launch_select(h, flags_dev, B, idx_dev, count_dev, temp, (cudaStream_t)stream);
}

extern "C" int ldpcb_gather_rows(ldpcb_t* h, const float* src_dev, const int32_t* idx_dev, const int32_t* count_dev,
                                 int64_t max_rows, int row_floats, float* dst_dev, void* stream) {
    LDPCB_ENTER(h);
    if (max_rows < 0 || row_floats <= 0 || (row_floats & 3) || !src_dev || !idx_dev || !dst_dev)
        return set_error(h, LDPCB_ERR_ARG, "ldpcb_gather_rows: bad arguments");
    if ((((uintptr_t)src_dev) | ((uintptr_t)dst_dev)) & 15) return set_error(h, LDPCB_ERR_ALIGN, "ldpcb_gather_rows: 16-byte alignment required");
    if (max_rows == 0) return LDPCB_OK;
    int64_t want = (max_rows * (row_floats / 4) + 255) / 256;
    int64_t cap = (int64_t)h->sm_count * 8;
    gather_rows_kernel<<<(int)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(
        (const float4*)src_dev, idx_dev, count_dev, max_rows, row_floats / 4, (float4*)dst_dev);
    LDPCB_LAUNCH_CHECK(h, "gather_rows_kernel");
    return LDPCB_OK;
}

extern "C" int ldpcb_dia_fir(ldpcb_t* h, const float* traj_dev, int64_t B, int n_rows, const float* taps_host,
                             float bias, float* out_dev, void* stream) {
    LDPCB_ENTER(h);
    if (B < 0 || n_rows < 1 || n_rows > FIR_MAX_ROWS || !traj_dev || !taps_host || !out_dev)
        return set_error(h, LDPCB_ERR_ARG, "ldpcb_dia_fir: bad arguments");
    if ((((uintptr_t)traj_dev) | ((uintptr_t)out_dev)) & 15) return set_error(h, LDPCB_ERR_ALIGN, "ldpcb_dia_fir: 16-byte alignment required");
    if (B == 0) return LDPCB_OK;
    FirTaps taps;
    for (int i = 0; i < FIR_MAX_ROWS; ++i) taps.t[i] = i < n_rows ? taps_host[i] : 0.0f;
    int64_t want = (B * (N / 4) + 255) / 256;
    int64_t cap = (int64_t)h->sm_count * 8;
    dia_fir_kernel<<<(int)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>((const float4*)traj_dev, B, n_rows, taps, bias, (float4*)out_dev);
    LDPCB_LAUNCH_CHECK(h, "dia_fir_kernel");
    return LDPCB_OK;
}

extern "C" int ldpcb_dl_window_policy(ldpcb_t* h, const int64_t* block_min_q_dev, const int32_t* score_exp_dev,
                                      const int64_t* truth_score_q_dev, int64_t B, int n_blocks, int win_width,
                                      const float* W1_host, const float* W2_host, float soft_margin,
                                      const int32_t* acc_block_size_host, uint8_t* success_dev, int32_t* windows_dev,
                                      int32_t* complexity_dev, uint64_t* counters_dev, void* stream) {
    LDPCB_ENTER(h);
    if (B < 0 || win_width < 1 || win_width > DLW_MAX_WIDTH || n_blocks < win_width || n_blocks > DLW_MAX_BLOCKS)
        return set_error(h, LDPCB_ERR_ARG, "ldpcb_dl_window_policy: B=%lld n_blocks=%d win_width=%d out of range", (long long)B, n_blocks, win_width);
    if (!block_min_q_dev || !score_exp_dev || !W1_host || !W2_host || !acc_block_size_host)
        return set_error(h, LDPCB_ERR_ARG, "ldpcb_dl_window_policy: NULL argument");
    if (B == 0) return LDPCB_OK;
    DlwParams p;
    memset(&p, 0, sizeof p);
    const int in_w = win_width + 1;
    for (int i = 0; i < in_w * in_w; ++i) p.W1[i] = W1_host[i];
    for (int i = 0; i < in_w * 2; ++i) p.W2[i] = W2_host[i];
    for (int i = 0; i <= n_blocks; ++i) p.acc[i] = acc_block_size_host[i];
    p.soft_margin = soft_margin; p.width = win_width; p.n_blocks = n_blocks;
    int64_t want = (B + 255) / 256;
    int64_t cap = (int64_t)h->sm_count * 8;
    dl_window_kernel<<<(int)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(
        (const long long*)block_min_q_dev, score_exp_dev, (const long long*)truth_score_q_dev, B, p, success_dev, windows_dev,
        complexity_dev, (unsigned long long*)counters_dev);
    LDPCB_LAUNCH_CHECK(h, "dl_window_kernel");
    return LDPCB_OK;
}

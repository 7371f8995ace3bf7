// The whole hot path as one stream-ordered sequence of launches, and the host-buffer entry points.
//
// Device pipeline (no host synchronisation anywhere):
//   NMS on all frames -> stable compaction of the frames with a non-zero syndrome -> OSD on those
//   frames' channel LLRs (row 0 of the reference's 13-row retest record, PB_OSD/pb_testing.py:71-72),
//   the OSD kernel reads the failure count from device memory -> decisions merged in place -> tallies.
// This is the Ldpc_128_testing -> ldpc-nonzero-retest.tfrecord -> PB_OSD/FS_OSD chain of the reference
// ("Training and Testing recipe.txt":14-18) without the files in between.
#include <algorithm>
#include <cstring>

#include "common.cuh"
#include "internal.cuh"
#include "philox.cuh"

namespace ldpcb {

struct Carver {
    char* base;
    size_t off = 0;
    explicit Carver(char* b) : base(b) {}
    template <typename T>
    T* take(size_t n) {
        off = (off + 255) & ~(size_t)255;
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += n * sizeof(T);
        return p;
    }
};

struct DecodeWs {
    uint8_t* syn;
    uint8_t* iters;
    int32_t* idx;
    int32_t* count;
    void* sel_temp;
    size_t bytes;
};

static DecodeWs carve_decode(char* base, size_t start, int64_t B) {
    Carver c(base);
    c.off = start;
    DecodeWs w;
    w.syn = c.take<uint8_t>((size_t)B);
    w.iters = c.take<uint8_t>((size_t)B);
    w.idx = c.take<int32_t>((size_t)B);
    w.count = c.take<int32_t>(1);
    w.sel_temp = c.take<char>(select_temp_bytes(B));
    w.bytes = c.off + 256;
    return w;
}

struct DecodeParams {
    int iters;
    float alpha, w_vc, w_marg;
    int early_stop, osd_order, tep_order;
};

// llr/final_bits/... are device pointers; `w` is workspace carved by the caller for this call.
static int decode_on_device(ldpcb_handle* h, const DecodeWs& w, const float* llr, int64_t B, const DecodeParams& p,
                            uint32_t* final_bits, uint8_t* syn_out, int32_t* best_tep, const uint32_t* truth,
                            uint64_t* counters, cudaStream_t st) {
    uint8_t* syn = syn_out ? syn_out : w.syn;
    NmsArgs n;
    n.llr = llr; n.idx = nullptr; n.B = B; n.iters = p.iters;
    n.alpha = p.alpha; n.w_vc = p.w_vc; n.w_marg = p.w_marg; n.early_stop = p.early_stop;
    n.hard_bits = final_bits; n.iters_used = w.iters; n.syndrome_nz = syn; n.soft_traj = nullptr;
    int s = launch_nms(h, n, st);
    if (s != LDPCB_OK) return s;
    if (truth && counters) {
        s = launch_tally_nms(h, final_bits, syn, w.iters, truth, B, counters, st);
        if (s != LDPCB_OK) return s;
    }
    if (best_tep) LDPCB_CUDA(h, cudaMemsetAsync(best_tep, 0xFF, sizeof(int32_t) * (size_t)B, st));
    if (p.osd_order >= 0) {
        s = launch_select(h, syn, B, w.idx, w.count, w.sel_temp, st);
        if (s != LDPCB_OK) return s;
        const TepTable& t = h->tep[p.osd_order][p.tep_order];
        OsdArgs a = {};
        a.order_llr = llr; a.score_llr = llr; a.idx = w.idx; a.count = w.count; a.B = B;
        a.teps = t.dev; a.n_teps = t.n; a.maxw = t.maxw; a.pair_index = t.pair_dev; a.triple_index = t.triple_dev; a.flags = 0;
        a.cw_bits = final_bits; a.best_tep = best_tep;
        s = launch_osd(h, a, st);
        if (s != LDPCB_OK) return s;
    }
    if (truth && counters) {
        s = launch_tally_final(h, final_bits, p.osd_order >= 0 ? syn : nullptr, best_tep, p.osd_order, p.tep_order, truth, B, counters, st);
        if (s != LDPCB_OK) return s;
    }
    return LDPCB_OK;
}

// ---- fused pipeline (CCSDS code, fixed iteration count): two kernels ----------------------------------------------
// nms_qc_kernel<FUSE> decodes, tallies the NMS stage and appends the frames with a non-zero syndrome to a list (in
// completion order: OSD results are written by frame index, so the order is immaterial); the OSD kernel reads the
// list and its length from device memory and tallies its own decisions.  No select / tally kernels, no host sync.
static bool fused_applies(const ldpcb_handle* h, const DecodeParams& p) {
    return h->qc_ccsds && p.early_stop == 0 && p.iters >= 1 && p.w_vc == p.w_marg;
}

struct FusedWs {
    int32_t* count;
    int32_t* idx;
    float* fail_llr;       // simulate only
    uint32_t* fail_truth;  // simulate only
    size_t bytes;
};
static FusedWs carve_fused(char* base, int64_t B, bool gen) {
    Carver c(base);
    FusedWs w;
    w.count = c.take<int32_t>(64);
    w.idx = c.take<int32_t>((size_t)B);
    w.fail_llr = gen ? c.take<float>((size_t)B * N) : nullptr;
    w.fail_truth = gen ? c.take<uint32_t>((size_t)B * 4) : nullptr;
    w.bytes = c.off + 256;
    return w;
}

static int osd_after_nms(ldpcb_handle* h, const FusedWs& w, const float* llr, bool compact, int64_t B, const DecodeParams& p,
                         uint32_t* final_bits, int32_t* best_tep, const uint32_t* truth, uint64_t* counters, cudaStream_t st) {
    const TepTable& t = h->tep[p.osd_order][p.tep_order];
    OsdArgs a = {};
    a.order_llr = llr; a.score_llr = llr; a.idx = compact ? nullptr : w.idx; a.count = w.count; a.B = B;
    a.teps = t.dev; a.n_teps = t.n; a.maxw = t.maxw; a.pair_index = t.pair_dev; a.triple_index = t.triple_dev; a.flags = 0;
    a.cw_bits = final_bits; a.best_tep = best_tep;
    a.tally_truth = (truth && counters) ? truth : nullptr; a.tally_counters = counters;
    return launch_osd(h, a, st);
}

static int decode_fused(ldpcb_handle* h, const FusedWs& w, const float* llr, int64_t B, const DecodeParams& p,
                        uint32_t* final_bits, uint8_t* syn_out, int32_t* best_tep, const uint32_t* truth,
                        uint64_t* counters, cudaStream_t st) {
    const bool tally = truth && counters;
    LDPCB_CUDA(h, cudaMemsetAsync(w.count, 0, sizeof(int32_t), st));
    if (best_tep) LDPCB_CUDA(h, cudaMemsetAsync(best_tep, 0xFF, sizeof(int32_t) * (size_t)B, st));
    NmsArgs n;
    n.llr = llr; n.idx = nullptr; n.B = B; n.iters = p.iters;
    n.alpha = p.alpha; n.w_vc = p.w_vc; n.w_marg = p.w_marg; n.early_stop = 0;
    n.hard_bits = final_bits; n.iters_used = nullptr; n.syndrome_nz = syn_out; n.soft_traj = nullptr;
    NmsFuse z;
    z.truth = tally ? truth : nullptr; z.counters = tally ? counters : nullptr;
    z.osd_follows = p.osd_order >= 0;
    if (p.osd_order >= 0) { z.fail_idx = w.idx; z.fail_count = w.count; }
    int s = launch_nms_qc(h, n, &z, st);
    if (s != LDPCB_OK) return s;
    if (p.osd_order >= 0) s = osd_after_nms(h, w, llr, false, B, p, final_bits, best_tep, truth, counters, st);
    return s;
}

static int check_decode_params(ldpcb_handle* h, const char* fn, int64_t B, const DecodeParams& p) {
    if (B < 0 || B > 0x7fffffff) return set_error(h, LDPCB_ERR_ARG, "%s: B=%lld out of range", fn, (long long)B);
    if (p.iters < 0 || p.iters > LDPCB_MAX_ITERS) return set_error(h, LDPCB_ERR_ARG, "%s: iters=%d out of range", fn, p.iters);
    if (p.osd_order < -1 || p.osd_order > 3 || p.tep_order < 0 || p.tep_order > 1)
        return set_error(h, LDPCB_ERR_ARG, "%s: osd_order=%d tep_order=%d out of range", fn, p.osd_order, p.tep_order);
    return LDPCB_OK;
}

}  // namespace ldpcb

using namespace ldpcb;

extern "C" int ldpcb_decode(ldpcb_t* h, const float* llr_dev, int64_t B, int iters, float alpha_check, float w_vc,
                            float w_marg, int early_stop, int osd_order, int tep_order, uint32_t* final_bits_dev,
                            uint8_t* syndrome_nz_dev, int32_t* best_tep_dev, const uint32_t* truth_bits_dev,
                            uint64_t* counters_dev, void* stream) {
    LDPCB_ENTER(h);
    DecodeParams p{iters, alpha_check, w_vc, w_marg, early_stop, osd_order, tep_order};
    int s = check_decode_params(h, "ldpcb_decode", B, p);
    if (s != LDPCB_OK) return s;
    if (B == 0) return LDPCB_OK;
    if (!llr_dev || !final_bits_dev) return set_error(h, LDPCB_ERR_ARG, "ldpcb_decode: NULL llr or final_bits");
    if ((uintptr_t)llr_dev & 15) return set_error(h, LDPCB_ERR_ALIGN, "ldpcb_decode: llr must be 16-byte aligned");
    char* wsbuf = nullptr;
    if (fused_applies(h, p)) {
        if ((s = ensure_stream_ws(h, (cudaStream_t)stream, carve_fused(nullptr, B, false).bytes, &wsbuf)) != LDPCB_OK) return s;
        return decode_fused(h, carve_fused(wsbuf, B, false), llr_dev, B, p, final_bits_dev, syndrome_nz_dev, best_tep_dev, truth_bits_dev,
                            counters_dev, (cudaStream_t)stream);
    }
    DecodeWs probe = carve_decode(nullptr, 0, B);
    if ((s = ensure_stream_ws(h, (cudaStream_t)stream, probe.bytes, &wsbuf)) != LDPCB_OK) return s;
    DecodeWs w = carve_decode(wsbuf, 0, B);
    return decode_on_device(h, w, llr_dev, B, p, final_bits_dev, syndrome_nz_dev, best_tep_dev, truth_bits_dev,
                            counters_dev, (cudaStream_t)stream);
}

extern "C" int ldpcb_simulate(ldpcb_t* h, uint64_t seed, uint64_t first_frame, int64_t B, float ebn0_db, int iters,
                              float alpha_check, float w_vc, float w_marg, int early_stop, int osd_order,
                              int tep_order, uint64_t* counters_dev, void* stream) {
    LDPCB_ENTER(h);
    DecodeParams p{iters, alpha_check, w_vc, w_marg, early_stop, osd_order, tep_order};
    int s = check_decode_params(h, "ldpcb_simulate", B, p);
    if (s != LDPCB_OK) return s;
    if (B == 0) return LDPCB_OK;
    if (!counters_dev) return set_error(h, LDPCB_ERR_ARG, "ldpcb_simulate: NULL counters");
    if (fused_applies(h, p)) {
        // generator fused into the decoder's prologue: no LLR ever reaches HBM except the channel values of the frames
        // NMS could not fix (512 B each), which the kernel appends for the OSD stage together with their codewords
        cudaStream_t st = (cudaStream_t)stream;
        char* wsbuf = nullptr;
        if ((s = ensure_stream_ws(h, st, carve_fused(nullptr, B, true).bytes, &wsbuf)) != LDPCB_OK) return s;
        const FusedWs w = carve_fused(wsbuf, B, true);
        LDPCB_CUDA(h, cudaMemsetAsync(w.count, 0, sizeof(int32_t), st));
        NmsArgs n;
        n.llr = nullptr; n.idx = nullptr; n.B = B; n.iters = p.iters;
        n.alpha = p.alpha; n.w_vc = p.w_vc; n.w_marg = p.w_marg; n.early_stop = 0;
        n.hard_bits = nullptr; n.iters_used = nullptr; n.syndrome_nz = nullptr; n.soft_traj = nullptr;
        NmsFuse z;
        z.counters = counters_dev; z.osd_follows = p.osd_order >= 0; z.gen = 1;
        z.key = make_uint2((unsigned)seed, (unsigned)(seed >> 32)); z.first_frame = first_frame; z.sigma = ebn0_to_sigma(ebn0_db);
        z.gcol = h->gcol_dev;
        if (p.osd_order >= 0) { z.fail_count = w.count; z.fail_llr = w.fail_llr; z.fail_truth = w.fail_truth; }
        if ((s = launch_nms_qc(h, n, &z, st)) != LDPCB_OK) return s;
        if (p.osd_order >= 0) s = osd_after_nms(h, w, w.fail_llr, true, B, p, nullptr, nullptr, w.fail_truth, counters_dev, st);
        return s;
    }
    Carver c(nullptr);
    c.take<float>((size_t)B * N);
    c.take<uint32_t>((size_t)B * 4);
    c.take<uint32_t>((size_t)B * 4);
    c.take<int32_t>((size_t)B);
    const size_t head = c.off;
    DecodeWs probe = carve_decode(nullptr, head, B);
    char* wsbuf = nullptr;
    if ((s = ensure_stream_ws(h, (cudaStream_t)stream, probe.bytes, &wsbuf)) != LDPCB_OK) return s;
    Carver d(wsbuf);
    float* llr = d.take<float>((size_t)B * N);
    uint32_t* truth = d.take<uint32_t>((size_t)B * 4);
    uint32_t* bits = d.take<uint32_t>((size_t)B * 4);
    int32_t* best = d.take<int32_t>((size_t)B);
    DecodeWs w = carve_decode(wsbuf, head, B);
    cudaStream_t st = (cudaStream_t)stream;
    if ((s = launch_gen(h, seed, first_frame, B, ebn0_db, llr, truth, st)) != LDPCB_OK) return s;
    return decode_on_device(h, w, llr, B, p, bits, nullptr, best, truth, counters_dev, st);
}

// ---- host-buffer entry points ------------------------------------------------------------------------
// Chunks of HOST_CHUNK frames round-robin over three streams, each with its own workspace slot, so
// chunk i+1's H2D copy and chunk i-1's D2H copy overlap chunk i's kernels.
namespace ldpcb {
constexpr int64_t HOST_CHUNK = 1 << 16;

static int sync_streams(ldpcb_handle* h) {
    for (int i = 0; i < 3; ++i) LDPCB_CUDA(h, cudaStreamSynchronize(h->streams[i]));
    return LDPCB_OK;
}
}  // namespace ldpcb

extern "C" int ldpcb_nms_decode_host(ldpcb_t* h, const float* llr_host, int64_t B, int iters, float alpha_check,
                                     float w_vc, float w_marg, int early_stop, uint32_t* hard_bits_host,
                                     uint8_t* iters_used_host, uint8_t* syndrome_nz_host, float* soft_traj_host) {
    LDPCB_ENTER(h);
    if (B < 0 || iters < 0 || iters > LDPCB_MAX_ITERS) return set_error(h, LDPCB_ERR_ARG, "ldpcb_nms_decode_host: B=%lld iters=%d out of range", (long long)B, iters);
    if (B == 0) return LDPCB_OK;
    if (!llr_host || !hard_bits_host) return set_error(h, LDPCB_ERR_ARG, "ldpcb_nms_decode_host: NULL llr or hard_bits");
    const int rows = iters + 1;
    const int64_t chunk = soft_traj_host ? std::min<int64_t>(HOST_CHUNK, 1 << 14) : HOST_CHUNK;
    int ci = 0;
    for (int64_t b0 = 0; b0 < B; b0 += chunk, ++ci) {
        const int64_t nb = std::min(chunk, B - b0);
        const int slot = 1 + ci % 3;
        cudaStream_t st = h->streams[ci % 3];
        Carver probe(nullptr);
        probe.take<float>((size_t)nb * N); probe.take<uint32_t>((size_t)nb * 4); probe.take<uint8_t>((size_t)nb);
        probe.take<uint8_t>((size_t)nb); if (soft_traj_host) probe.take<float>((size_t)nb * rows * N);
        int s = ensure_ws(h, slot, probe.off + 256);
        if (s != LDPCB_OK) return s;
        Carver c(h->ws[slot].buf);
        float* llr = c.take<float>((size_t)nb * N);
        uint32_t* bits = c.take<uint32_t>((size_t)nb * 4);
        uint8_t* it = c.take<uint8_t>((size_t)nb);
        uint8_t* syn = c.take<uint8_t>((size_t)nb);
        float* traj = soft_traj_host ? c.take<float>((size_t)nb * rows * N) : nullptr;
        LDPCB_CUDA(h, cudaMemcpyAsync(llr, llr_host + b0 * N, sizeof(float) * nb * N, cudaMemcpyHostToDevice, st));
        NmsArgs a;
        a.llr = llr; a.idx = nullptr; a.B = nb; a.iters = iters; a.alpha = alpha_check; a.w_vc = w_vc; a.w_marg = w_marg;
        a.early_stop = early_stop; a.hard_bits = bits; a.iters_used = it; a.syndrome_nz = syn; a.soft_traj = traj;
        if ((s = launch_nms(h, a, st)) != LDPCB_OK) return s;
        LDPCB_CUDA(h, cudaMemcpyAsync(hard_bits_host + b0 * 4, bits, sizeof(uint32_t) * nb * 4, cudaMemcpyDeviceToHost, st));
        if (iters_used_host) LDPCB_CUDA(h, cudaMemcpyAsync(iters_used_host + b0, it, (size_t)nb, cudaMemcpyDeviceToHost, st));
        if (syndrome_nz_host) LDPCB_CUDA(h, cudaMemcpyAsync(syndrome_nz_host + b0, syn, (size_t)nb, cudaMemcpyDeviceToHost, st));
        if (traj) LDPCB_CUDA(h, cudaMemcpyAsync(soft_traj_host + b0 * rows * N, traj, sizeof(float) * nb * rows * N, cudaMemcpyDeviceToHost, st));
    }
    return sync_streams(h);
}

extern "C" int ldpcb_osd_decode_host(ldpcb_t* h, const float* order_llr_host, const float* score_llr_host, int64_t B,
                                     int order, int tep_order, int flags, uint32_t* cw_bits_host,
                                     int32_t* best_tep_host, int64_t* best_score_q_host, int32_t* score_exp_host,
                                     uint8_t* perm_host, uint64_t* redG_host) {
    LDPCB_ENTER(h);
    if (B < 0 || order < 0 || order > 3 || tep_order < 0 || tep_order > 1 || (flags & ~3))
        return set_error(h, LDPCB_ERR_ARG, "ldpcb_osd_decode_host: B=%lld order=%d tep_order=%d flags=%d out of range", (long long)B, order, tep_order, flags);
    if (B == 0) return LDPCB_OK;
    if (!order_llr_host || !score_llr_host || !cw_bits_host) return set_error(h, LDPCB_ERR_ARG, "ldpcb_osd_decode_host: NULL llr or cw_bits");
    const bool same = (order_llr_host == score_llr_host);
    const TepTable& t = h->tep[order][tep_order];
    int ci = 0;
    for (int64_t b0 = 0; b0 < B; b0 += HOST_CHUNK, ++ci) {
        const int64_t nb = std::min(HOST_CHUNK, B - b0);
        const int slot = 1 + ci % 3;
        cudaStream_t st = h->streams[ci % 3];
        Carver probe(nullptr);
        probe.take<float>((size_t)nb * N); probe.take<float>((size_t)nb * N); probe.take<uint32_t>((size_t)nb * 4);
        probe.take<int32_t>((size_t)nb); probe.take<int64_t>((size_t)nb); probe.take<int32_t>((size_t)nb);
        probe.take<uint8_t>((size_t)nb * N); probe.take<uint64_t>((size_t)nb * K);
        int s = ensure_ws(h, slot, probe.off + 256);
        if (s != LDPCB_OK) return s;
        Carver c(h->ws[slot].buf);
        float* ol = c.take<float>((size_t)nb * N);
        float* sl = c.take<float>((size_t)nb * N);
        uint32_t* bits = c.take<uint32_t>((size_t)nb * 4);
        int32_t* bt = c.take<int32_t>((size_t)nb);
        int64_t* bq = c.take<int64_t>((size_t)nb);
        int32_t* ex = c.take<int32_t>((size_t)nb);
        uint8_t* pm = c.take<uint8_t>((size_t)nb * N);
        uint64_t* rg = c.take<uint64_t>((size_t)nb * K);
        LDPCB_CUDA(h, cudaMemcpyAsync(ol, order_llr_host + b0 * N, sizeof(float) * nb * N, cudaMemcpyHostToDevice, st));
        if (!same) LDPCB_CUDA(h, cudaMemcpyAsync(sl, score_llr_host + b0 * N, sizeof(float) * nb * N, cudaMemcpyHostToDevice, st));
        OsdArgs a = {};
        a.order_llr = ol; a.score_llr = same ? ol : sl; a.B = nb; a.teps = t.dev; a.n_teps = t.n; a.maxw = t.maxw; a.pair_index = t.pair_dev; a.triple_index = t.triple_dev; a.flags = flags;
        a.cw_bits = bits; a.best_tep = bt; a.best_score_q = bq; a.score_exp = ex;
        a.perm = perm_host ? pm : nullptr; a.redG = redG_host ? rg : nullptr;
        if ((s = launch_osd(h, a, st)) != LDPCB_OK) return s;
        LDPCB_CUDA(h, cudaMemcpyAsync(cw_bits_host + b0 * 4, bits, sizeof(uint32_t) * nb * 4, cudaMemcpyDeviceToHost, st));
        if (best_tep_host) LDPCB_CUDA(h, cudaMemcpyAsync(best_tep_host + b0, bt, sizeof(int32_t) * nb, cudaMemcpyDeviceToHost, st));
        if (best_score_q_host) LDPCB_CUDA(h, cudaMemcpyAsync(best_score_q_host + b0, bq, sizeof(int64_t) * nb, cudaMemcpyDeviceToHost, st));
        if (score_exp_host) LDPCB_CUDA(h, cudaMemcpyAsync(score_exp_host + b0, ex, sizeof(int32_t) * nb, cudaMemcpyDeviceToHost, st));
        if (perm_host) LDPCB_CUDA(h, cudaMemcpyAsync(perm_host + b0 * N, pm, (size_t)nb * N, cudaMemcpyDeviceToHost, st));
        if (redG_host) LDPCB_CUDA(h, cudaMemcpyAsync(redG_host + b0 * K, rg, sizeof(uint64_t) * nb * K, cudaMemcpyDeviceToHost, st));
    }
    return sync_streams(h);
}

extern "C" int ldpcb_osd_sweep_host(ldpcb_t* h, const float* upd_order_llr_host, const float* upd_score_llr_host,
                                    const uint64_t* redG_host, int64_t B, const uint32_t* teps_host, int32_t n_teps, int flags,
                                    uint32_t* cw_bits_host, int32_t* best_tep_host, int64_t* best_score_q_host,
                                    int32_t* score_exp_host) {
    LDPCB_ENTER(h);
    if (B < 0 || n_teps < 1 || (flags & ~3)) return set_error(h, LDPCB_ERR_ARG, "ldpcb_osd_sweep_host: B=%lld n_teps=%d flags=%d out of range", (long long)B, n_teps, flags);
    if (B == 0) return LDPCB_OK;
    if (!upd_order_llr_host || !upd_score_llr_host || !redG_host || !teps_host || !cw_bits_host)
        return set_error(h, LDPCB_ERR_ARG, "ldpcb_osd_sweep_host: NULL argument");
    const bool same = (upd_order_llr_host == upd_score_llr_host);
    int maxw = 1;
    for (int i = 0; i < n_teps; ++i) {
        int w = 0;
        for (int j = 0; j < 4; ++j) {
            const unsigned t = (teps_host[i] >> (8 * j)) & 0xffu;
            if (t < 64u) ++w; else if (t != 0xffu) return set_error(h, LDPCB_ERR_ARG, "ldpcb_osd_sweep_host: TEP %d has position %u", i, t);
        }
        maxw = std::max(maxw, w);
    }
    cudaStream_t st = h->streams[0];
    Carver probe(nullptr);
    probe.take<uint32_t>((size_t)n_teps + 128);
    const int64_t chunk = std::min<int64_t>(B, HOST_CHUNK);
    probe.take<float>((size_t)chunk * N); probe.take<float>((size_t)chunk * N); probe.take<uint64_t>((size_t)chunk * K);
    probe.take<uint32_t>((size_t)chunk * 4); probe.take<int32_t>((size_t)chunk); probe.take<int64_t>((size_t)chunk); probe.take<int32_t>((size_t)chunk);
    int s = ensure_ws(h, 1, probe.off + 256);
    if (s != LDPCB_OK) return s;
    Carver c(h->ws[1].buf);
    uint32_t* teps = c.take<uint32_t>((size_t)n_teps + 128);  // padded like the built-in tables
    float* ol = c.take<float>((size_t)chunk * N);
    float* sl = c.take<float>((size_t)chunk * N);
    uint64_t* rg = c.take<uint64_t>((size_t)chunk * K);
    uint32_t* bits = c.take<uint32_t>((size_t)chunk * 4);
    int32_t* bt = c.take<int32_t>((size_t)chunk);
    int64_t* bq = c.take<int64_t>((size_t)chunk);
    int32_t* ex = c.take<int32_t>((size_t)chunk);
    LDPCB_CUDA(h, cudaMemsetAsync(teps + n_teps, 0xFF, sizeof(uint32_t) * 128, st));
    LDPCB_CUDA(h, cudaMemcpyAsync(teps, teps_host, sizeof(uint32_t) * n_teps, cudaMemcpyHostToDevice, st));
    for (int64_t b0 = 0; b0 < B; b0 += chunk) {
        const int64_t nb = std::min(chunk, B - b0);
        LDPCB_CUDA(h, cudaMemcpyAsync(ol, upd_order_llr_host + b0 * N, sizeof(float) * nb * N, cudaMemcpyHostToDevice, st));
        if (!same) LDPCB_CUDA(h, cudaMemcpyAsync(sl, upd_score_llr_host + b0 * N, sizeof(float) * nb * N, cudaMemcpyHostToDevice, st));
        LDPCB_CUDA(h, cudaMemcpyAsync(rg, redG_host + b0 * K, sizeof(uint64_t) * nb * K, cudaMemcpyHostToDevice, st));
        OsdArgs a = {};
        a.order_llr = ol; a.score_llr = same ? ol : sl; a.redG_in = rg; a.B = nb; a.teps = teps; a.n_teps = n_teps; a.maxw = maxw; a.flags = flags;
        a.cw_bits = bits; a.best_tep = bt; a.best_score_q = bq; a.score_exp = ex;
        if ((s = launch_osd(h, a, st)) != LDPCB_OK) return s;
        LDPCB_CUDA(h, cudaMemcpyAsync(cw_bits_host + b0 * 4, bits, sizeof(uint32_t) * nb * 4, cudaMemcpyDeviceToHost, st));
        if (best_tep_host) LDPCB_CUDA(h, cudaMemcpyAsync(best_tep_host + b0, bt, sizeof(int32_t) * nb, cudaMemcpyDeviceToHost, st));
        if (best_score_q_host) LDPCB_CUDA(h, cudaMemcpyAsync(best_score_q_host + b0, bq, sizeof(int64_t) * nb, cudaMemcpyDeviceToHost, st));
        if (score_exp_host) LDPCB_CUDA(h, cudaMemcpyAsync(score_exp_host + b0, ex, sizeof(int32_t) * nb, cudaMemcpyDeviceToHost, st));
        LDPCB_CUDA(h, cudaStreamSynchronize(st));
    }
    return LDPCB_OK;
}

extern "C" int ldpcb_osd_fs_decode_host(ldpcb_t* h, const float* llr_host, int64_t B, int order_limit, float tau_e, int tau_psc,
                                        float beta_shift, uint32_t* cw_bits_host, int32_t* best_tep_host,
                                        int32_t* num_teps_host, uint8_t* stop_kind_host) {
    LDPCB_ENTER(h);
    if (B < 0 || order_limit < 0 || order_limit > 3) return set_error(h, LDPCB_ERR_ARG, "ldpcb_osd_fs_decode_host: bad arguments");
    if (B == 0) return LDPCB_OK;
    if (!llr_host || !cw_bits_host) return set_error(h, LDPCB_ERR_ARG, "ldpcb_osd_fs_decode_host: NULL llr or cw_bits");
    int ci = 0;
    for (int64_t b0 = 0; b0 < B; b0 += HOST_CHUNK, ++ci) {
        const int64_t nb = std::min(HOST_CHUNK, B - b0);
        const int slot = 1 + ci % 3;
        cudaStream_t st = h->streams[ci % 3];
        Carver probe(nullptr);
        probe.take<float>((size_t)nb * N); probe.take<uint32_t>((size_t)nb * 4); probe.take<int32_t>((size_t)nb);
        probe.take<int32_t>((size_t)nb); probe.take<uint8_t>((size_t)nb);
        int s = ensure_ws(h, slot, probe.off + 256);
        if (s != LDPCB_OK) return s;
        Carver c(h->ws[slot].buf);
        float* llr = c.take<float>((size_t)nb * N);
        uint32_t* bits = c.take<uint32_t>((size_t)nb * 4);
        int32_t* bt = c.take<int32_t>((size_t)nb);
        int32_t* nt = c.take<int32_t>((size_t)nb);
        uint8_t* sk = c.take<uint8_t>((size_t)nb);
        LDPCB_CUDA(h, cudaMemcpyAsync(llr, llr_host + b0 * N, sizeof(float) * nb * N, cudaMemcpyHostToDevice, st));
        s = ldpcb_osd_fs_decode(h, llr, nb, order_limit, tau_e, tau_psc, beta_shift, bits, bt, nt, sk, nullptr, nullptr, nullptr, st);
        if (s != LDPCB_OK) return s;
        LDPCB_CUDA(h, cudaMemcpyAsync(cw_bits_host + b0 * 4, bits, sizeof(uint32_t) * nb * 4, cudaMemcpyDeviceToHost, st));
        if (best_tep_host) LDPCB_CUDA(h, cudaMemcpyAsync(best_tep_host + b0, bt, sizeof(int32_t) * nb, cudaMemcpyDeviceToHost, st));
        if (num_teps_host) LDPCB_CUDA(h, cudaMemcpyAsync(num_teps_host + b0, nt, sizeof(int32_t) * nb, cudaMemcpyDeviceToHost, st));
        if (stop_kind_host) LDPCB_CUDA(h, cudaMemcpyAsync(stop_kind_host + b0, sk, (size_t)nb, cudaMemcpyDeviceToHost, st));
    }
    return sync_streams(h);
}

extern "C" int ldpcb_decode_host(ldpcb_t* h, const float* llr_host, int64_t B, int iters, float alpha_check, float w_vc,
                                 float w_marg, int early_stop, int osd_order, int tep_order,
                                 uint32_t* final_bits_host, uint8_t* syndrome_nz_host, int32_t* best_tep_host,
                                 const uint32_t* truth_bits_host, uint64_t* counters_host) {
    LDPCB_ENTER(h);
    DecodeParams p{iters, alpha_check, w_vc, w_marg, early_stop, osd_order, tep_order};
    int s = check_decode_params(h, "ldpcb_decode_host", B, p);
    if (s != LDPCB_OK) return s;
    if (B == 0) return LDPCB_OK;
    if (!llr_host || !final_bits_host) return set_error(h, LDPCB_ERR_ARG, "ldpcb_decode_host: NULL llr or final_bits");
    const bool tally = truth_bits_host && counters_host;
    // device counters live at the head of workspace slot 0
    if ((s = ensure_ws(h, 0, 4096)) != LDPCB_OK) return s;
    uint64_t* counters = reinterpret_cast<uint64_t*>(h->ws[0].buf);
    if (tally) {
        LDPCB_CUDA(h, cudaMemsetAsync(counters, 0, sizeof(uint64_t) * LDPCB_NUM_COUNTERS, h->streams[0]));
        LDPCB_CUDA(h, cudaEventRecord(h->events[0], h->streams[0]));
        LDPCB_CUDA(h, cudaStreamWaitEvent(h->streams[1], h->events[0], 0));
        LDPCB_CUDA(h, cudaStreamWaitEvent(h->streams[2], h->events[0], 0));
    }
    int ci = 0;
    for (int64_t b0 = 0; b0 < B; b0 += HOST_CHUNK, ++ci) {
        const int64_t nb = std::min(HOST_CHUNK, B - b0);
        const int slot = 1 + ci % 3;
        cudaStream_t st = h->streams[ci % 3];
        Carver probe(nullptr);
        probe.take<float>((size_t)nb * N); probe.take<uint32_t>((size_t)nb * 4); probe.take<uint32_t>((size_t)nb * 4);
        probe.take<uint8_t>((size_t)nb); probe.take<int32_t>((size_t)nb);
        const size_t head = probe.off;
        DecodeWs pw = carve_decode(nullptr, head, nb);
        if ((s = ensure_ws(h, slot, pw.bytes)) != LDPCB_OK) return s;
        Carver c(h->ws[slot].buf);
        float* llr = c.take<float>((size_t)nb * N);
        uint32_t* bits = c.take<uint32_t>((size_t)nb * 4);
        uint32_t* truth = c.take<uint32_t>((size_t)nb * 4);
        uint8_t* syn = c.take<uint8_t>((size_t)nb);
        int32_t* bt = c.take<int32_t>((size_t)nb);
        DecodeWs w = carve_decode(h->ws[slot].buf, head, nb);
        LDPCB_CUDA(h, cudaMemcpyAsync(llr, llr_host + b0 * N, sizeof(float) * nb * N, cudaMemcpyHostToDevice, st));
        if (tally) LDPCB_CUDA(h, cudaMemcpyAsync(truth, truth_bits_host + b0 * 4, sizeof(uint32_t) * nb * 4, cudaMemcpyHostToDevice, st));
        if (fused_applies(h, p)) {
            const FusedWs fw{w.count, w.idx, nullptr, nullptr, 0};
            s = decode_fused(h, fw, llr, nb, p, bits, syn, best_tep_host ? bt : nullptr, tally ? truth : nullptr, tally ? counters : nullptr, st);
        } else {
            s = decode_on_device(h, w, llr, nb, p, bits, syn, (best_tep_host || tally) ? bt : nullptr, tally ? truth : nullptr,
                                 tally ? counters : nullptr, st);
        }
        if (s != LDPCB_OK) return s;
        LDPCB_CUDA(h, cudaMemcpyAsync(final_bits_host + b0 * 4, bits, sizeof(uint32_t) * nb * 4, cudaMemcpyDeviceToHost, st));
        if (syndrome_nz_host) LDPCB_CUDA(h, cudaMemcpyAsync(syndrome_nz_host + b0, syn, (size_t)nb, cudaMemcpyDeviceToHost, st));
        if (best_tep_host) LDPCB_CUDA(h, cudaMemcpyAsync(best_tep_host + b0, bt, sizeof(int32_t) * nb, cudaMemcpyDeviceToHost, st));
    }
    if ((s = sync_streams(h)) != LDPCB_OK) return s;
    if (tally) {
        uint64_t tmp[LDPCB_NUM_COUNTERS];
        LDPCB_CUDA(h, cudaMemcpy(tmp, counters, sizeof tmp, cudaMemcpyDeviceToHost));
        for (int i = 0; i < LDPCB_NUM_COUNTERS; ++i) counters_host[i] += tmp[i];
    }
    return LDPCB_OK;
}

// Decoding_model.call in one host call (ms_test.py:30-34): NMS on every frame with the get_eval tallies (:36-54), then the
// iters+1 posteriors of the frames with a non-zero syndrome, in ascending frame order -- the 13-row retest records of
// collect_failed_output_selective (:55-64).  The failures are compacted on the device and re-decoded there with the
// trajectory output (a few hundred frames), so the host sees one H2D copy, one small synchronisation for the failure
// count, and D2H copies sized by that count.
extern "C" int ldpcb_nms_retest_host(ldpcb_t* h, const float* llr_host, int64_t B, int iters, float alpha_check, float w_vc,
                                     float w_marg, const uint32_t* truth_bits_host, uint32_t* hard_bits_host,
                                     uint8_t* syndrome_nz_host, uint64_t* counters_host, int64_t max_fail,
                                     int32_t* fail_idx_host, float* fail_traj_host, int64_t* n_fail_host) {
    LDPCB_ENTER(h);
    if (B < 0 || B > 0x7fffffff || iters < 0 || iters > LDPCB_MAX_ITERS || max_fail < 0)
        return set_error(h, LDPCB_ERR_ARG, "ldpcb_nms_retest_host: B=%lld iters=%d max_fail=%lld out of range", (long long)B, iters, (long long)max_fail);
    if (!n_fail_host) return set_error(h, LDPCB_ERR_ARG, "ldpcb_nms_retest_host: NULL n_fail");
    *n_fail_host = 0;
    if (B == 0) return LDPCB_OK;
    if (!llr_host || !hard_bits_host || (max_fail > 0 && (!fail_idx_host || !fail_traj_host)))
        return set_error(h, LDPCB_ERR_ARG, "ldpcb_nms_retest_host: NULL llr, hard_bits, fail_idx or fail_traj");
    const bool tally = truth_bits_host && counters_host;
    const int rows = iters + 1;
    const int64_t chunk = std::min<int64_t>(B, 1 << 14);
    cudaStream_t st = h->streams[0];
    int s;
    Carver probe(nullptr);
    probe.take<uint64_t>(LDPCB_NUM_COUNTERS); probe.take<float>((size_t)chunk * N); probe.take<uint32_t>((size_t)chunk * 4);
    probe.take<uint32_t>((size_t)chunk * 4); probe.take<uint32_t>((size_t)chunk * 4); probe.take<float>((size_t)chunk * rows * N);
    const size_t head = probe.off;
    DecodeWs pw = carve_decode(nullptr, head, chunk);
    if ((s = ensure_ws(h, 1, pw.bytes)) != LDPCB_OK) return s;
    Carver c(h->ws[1].buf);
    uint64_t* counters = c.take<uint64_t>(LDPCB_NUM_COUNTERS);
    float* llr = c.take<float>((size_t)chunk * N);
    uint32_t* bits = c.take<uint32_t>((size_t)chunk * 4);
    uint32_t* truth = c.take<uint32_t>((size_t)chunk * 4);
    uint32_t* bits2 = c.take<uint32_t>((size_t)chunk * 4);
    float* traj = c.take<float>((size_t)chunk * rows * N);
    DecodeWs w = carve_decode(h->ws[1].buf, head, chunk);
    if (tally) LDPCB_CUDA(h, cudaMemsetAsync(counters, 0, sizeof(uint64_t) * LDPCB_NUM_COUNTERS, st));
    int64_t total_fail = 0, stored = 0;
    for (int64_t b0 = 0; b0 < B; b0 += chunk) {
        const int64_t nb = std::min(chunk, B - b0);
        LDPCB_CUDA(h, cudaMemcpyAsync(llr, llr_host + b0 * N, sizeof(float) * nb * N, cudaMemcpyHostToDevice, st));
        if (tally) LDPCB_CUDA(h, cudaMemcpyAsync(truth, truth_bits_host + b0 * 4, sizeof(uint32_t) * nb * 4, cudaMemcpyHostToDevice, st));
        NmsArgs n;
        n.llr = llr; n.idx = nullptr; n.B = nb; n.iters = iters; n.alpha = alpha_check; n.w_vc = w_vc; n.w_marg = w_marg; n.early_stop = 0;
        n.hard_bits = bits; n.iters_used = w.iters; n.syndrome_nz = w.syn; n.soft_traj = nullptr;
        if (nms_qc_applies(h, n)) {
            NmsFuse z;
            z.truth = tally ? truth : nullptr; z.counters = tally ? counters : nullptr;
            n.iters_used = nullptr;
            s = launch_nms_qc(h, n, &z, st);
        } else {
            s = launch_nms(h, n, st);
            if (s == LDPCB_OK && tally) s = launch_tally_nms(h, bits, w.syn, w.iters, truth, nb, counters, st);
            if (s == LDPCB_OK && tally) s = launch_tally_final(h, bits, nullptr, nullptr, -1, 0, truth, nb, counters, st);
        }
        if (s != LDPCB_OK) return s;
        if ((s = launch_select(h, w.syn, nb, w.idx, w.count, w.sel_temp, st)) != LDPCB_OK) return s;
        int32_t nf = 0;
        LDPCB_CUDA(h, cudaMemcpyAsync(&nf, w.count, sizeof nf, cudaMemcpyDeviceToHost, st));
        LDPCB_CUDA(h, cudaMemcpyAsync(hard_bits_host + b0 * 4, bits, sizeof(uint32_t) * nb * 4, cudaMemcpyDeviceToHost, st));
        if (syndrome_nz_host) LDPCB_CUDA(h, cudaMemcpyAsync(syndrome_nz_host + b0, w.syn, (size_t)nb, cudaMemcpyDeviceToHost, st));
        LDPCB_CUDA(h, cudaStreamSynchronize(st));
        total_fail += nf;
        const int64_t take = std::min<int64_t>(nf, max_fail - stored);
        if (take > 0) {
            NmsArgs t = n;
            t.idx = w.idx; t.B = take; t.hard_bits = bits2; t.iters_used = nullptr; t.syndrome_nz = nullptr; t.soft_traj = traj;
            if ((s = launch_nms(h, t, st)) != LDPCB_OK) return s;
            LDPCB_CUDA(h, cudaMemcpyAsync(fail_idx_host + stored, w.idx, sizeof(int32_t) * take, cudaMemcpyDeviceToHost, st));
            LDPCB_CUDA(h, cudaMemcpyAsync(fail_traj_host + stored * rows * N, traj, sizeof(float) * take * rows * N, cudaMemcpyDeviceToHost, st));
            LDPCB_CUDA(h, cudaStreamSynchronize(st));
            for (int64_t i = 0; i < take; ++i) fail_idx_host[stored + i] += (int32_t)b0;
            stored += take;
        }
    }
    if (tally) {
        uint64_t tmp[LDPCB_NUM_COUNTERS];
        LDPCB_CUDA(h, cudaMemcpyAsync(tmp, counters, sizeof tmp, cudaMemcpyDeviceToHost, st));
        LDPCB_CUDA(h, cudaStreamSynchronize(st));
        for (int i = 0; i < LDPCB_NUM_COUNTERS; ++i) counters_host[i] += tmp[i];
    }
    *n_fail_host = total_fail;
    return LDPCB_OK;
}

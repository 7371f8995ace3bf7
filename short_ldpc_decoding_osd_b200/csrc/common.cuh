// Internal declarations shared by the translation units of libldpc_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "ldpc_b200.h"

namespace ldpcb {

constexpr int N = LDPCB_N;      // 128 code bits
constexpr int M = LDPCB_M;      // 64 checks
constexpr int K = LDPCB_K;      // 64 information bits
constexpr int DC = LDPCB_MAX_CHK_DEG;  // 8
constexpr int DV = LDPCB_MAX_VAR_DEG;  // 8
constexpr int NUM_WS = 4;       // workspace slots (0: device pipeline, 1..3: host pipeline streams)

// ---- NMS tables (device, built once per handle; reference: the dense H of ms_test.py:126,182) ----
// Shared-memory layout of one frame in the NMS kernel, in floats:
//   T[0..127]   per-variable total (sum of incoming cv + w_vc*y), T[128] = +inf (padding edges)
//   CV[e*96 + rot[e] + c]  check->variable message of the edge labelled e (0..7) of check c, at float offset
//                NMS_CV_OFF.  A lane writes its two checks at consecutive addresses (conflict-free); the labels of
//                a check's edges and the rotations rot[e] (< 32) are chosen at ldpcb_create so that the variable
//                side, which gathers its incoming messages in ascending check order, is bank-conflict free too
//   CV[768] = 0 (padding for variables of low degree), CV[769] = dump slot for padded check edges
constexpr int NMS_T_FLOATS = 132;
constexpr int NMS_CV_OFF = NMS_T_FLOATS;
constexpr int NMS_CV_STRIDE = 96;
constexpr int NMS_CV_ZERO = DC * NMS_CV_STRIDE;      // 768
constexpr int NMS_CV_DUMP = NMS_CV_ZERO + 1;
constexpr int NMS_CV_FLOATS = NMS_CV_ZERO + 4;
constexpr int NMS_FRAME_FLOATS = NMS_T_FLOATS + NMS_CV_FLOATS;  // 904 floats = 3616 B

struct NmsTables {
    // chk_var[c][e]: variable index of the edge labelled e of check c, 128 = padding
    uint8_t chk_var[M][DC];
    // rot[e]: bank rotation of label e
    int rot[DC];
    // var_slot[v][d]: CV index (e*96+rot[e]+c) of the d-th incoming edge of variable v, checks ascending
    // (the summation order of tf.reduce_sum(cv_matrix, 1) restated in oracle/nms_oracle.py); 768 = padding
    uint16_t var_slot[N][DV];
    // chk_mask[c][w]: bit-packed row c of H (syndrome)
    uint32_t chk_mask[M][4];
    int max_var_deg_lo;  // max variable degree over columns 0..63
    int max_var_deg_hi;  // max variable degree over columns 64..127
};

// ---- OSD ----
constexpr int OSD_LUT_BYTES = 8 * 256 * 8;
struct TepTable {
    uint32_t* dev = nullptr;      // packed TEPs
    std::vector<uint32_t> host;
    int n = 0;
    int maxw = 0;
    // orders 0..2: inverse of the enumeration, [i*64+j] (i<j) -> index of the pair TEP, [4096+i] -> index of the
    // single TEP {i}, [4096+64] -> index of the empty TEP (the tensor-core pair sweep visits TEPs in its own order)
    uint16_t* pair_dev = nullptr;
    uint16_t* triple_dev = nullptr;  // order 3: inverse of the enumeration for the weight-3 TEPs, [C(k,3)+C(j,2)+i] for i<j<k
};
constexpr int OSD_PAIR_TABLE = 64 * 64 + 65;

struct Workspace {
    char* buf = nullptr;
    size_t cap = 0;
};
// slots of ldpcb_handle::occ (occupancy is a property of (kernel, device), so it lives in the handle)
enum { OCC_NMS = 0 /* +0..7: template variants */, OCC_NMS_QC = 8 /* +0..4 */, OCC_OSD = 13 /* +0..8 */, OCC_OSD_FS = 23, OCC_OSD_PAIR = 24,
       OCC_OSD_PB = 25, OCC_OSD3 = 26 /* +0..1 */, OCC_OSD_BLOCKS = 28 /* +0..3 */, OCC_SLOTS = 32 };

}  // namespace ldpcb

struct ldpcb_handle {
    int device = 0;
    int sm_count = 0;
    std::string err;
    uint8_t H[ldpcb::M * ldpcb::N];
    uint8_t G[ldpcb::K * ldpcb::N];
    ldpcb::NmsTables nms_host;
    ldpcb::NmsTables* nms_dev = nullptr;
    uint64_t* gcol_dev = nullptr;  // [0,128): column j of G, bit r = G[r][j]; [128,256): unit-column flags (handle.cu)
    uint64_t gcol_host[ldpcb::N];
    ldpcb::TepTable tep[4][2];      // [order][tep_order]
    int32_t* one_block_dev = nullptr;  // {0, n} scratch for single-block calls
    ldpcb::Workspace ws[ldpcb::NUM_WS];   // [1..3]: the *_host pipelines' private slots, [0]: their shared counters
    std::map<cudaStream_t, ldpcb::Workspace> fb_ws;      // osd3.cu: frames left to the exact sweep, one list per caller stream
    std::map<cudaStream_t, ldpcb::Workspace> stream_ws;  // scratch of the device-pointer calls, one per caller stream
    int occ[ldpcb::OCC_SLOTS] = {};       // resident CTAs per SM of each kernel variant on THIS handle's device (0 = not queried)
    int qc_minb = 2;                      // nms_qc.cu register budget: 2 CTAs of 8 warps per SM at 110 registers (default: measured
                                          // faster, 5.61 vs 5.73 ms per 2^21 frames), or 3 at 80 (env LDPCB_QC_MINB=3, A/B timing)
    bool qc_ccsds = false;                // H is the CCSDS (128,64) matrix nms_qc.cu is specialised to
    int occ_pb[3] = {0, 0, 0};            // osd_pb_kernel at 6 / 8 / 10 CTAs per SM
    bool pb_consts_ready = false;         // __constant__ tables of osd_pb.cu uploaded to this device
    char* pb_list = nullptr;       // PB-OSD order 3: TEP lists of the resident warps
    size_t pb_list_cap = 0;        // in list entries
    int* pb_queue = nullptr;       // PB-OSD: work counter of the running launch
    cudaStream_t streams[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t events[3] = {nullptr, nullptr, nullptr};
    uint64_t launches = 0;
};

namespace ldpcb {

int set_error(ldpcb_handle* h, int code, const char* fmt, ...);
int check_cuda(ldpcb_handle* h, cudaError_t e, const char* what);
// Grow workspace slot `slot` to at least `bytes` (cudaMalloc on growth only).
int ensure_ws(ldpcb_handle* h, int slot, size_t bytes);
// Scratch of a device-pointer call, private to the caller's stream (two calls on different streams never share it).
// First use on a stream and growth allocate, and growth synchronises the device (documented in ldpc_b200.h).
int ensure_stream_ws(ldpcb_handle* h, cudaStream_t st, size_t bytes, char** buf);

// Makes the handle's device current for the duration of an entry point and restores the caller's device on return,
// so handles of several devices can be used from one thread.
struct DeviceGuard {
    int prev = -1, want = -1;
    explicit DeviceGuard(const ldpcb_handle* h) {
        if (!h) return;
        want = h->device;
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
        if (prev != want) cudaSetDevice(want);
    }
    ~DeviceGuard() {
        if (prev >= 0 && prev != want) cudaSetDevice(prev);
    }
};
#define LDPCB_ENTER(h)                   \
    if (!(h)) return LDPCB_ERR_ARG;      \
    ldpcb::DeviceGuard _ldpcb_guard(h)

#define LDPCB_CUDA(h, call)                                         \
    do {                                                            \
        int _st = ldpcb::check_cuda((h), (call), #call);            \
        if (_st != LDPCB_OK) return _st;                            \
    } while (0)

// Device-side bounds checks of every data-dependent index (list appends, inverse-table lookups, decoded candidate
// positions).  Compiled in only by `python -m short_ldpc_decoding_osd_b200.build --bounds` (a separate
// libldpc_b200_bounds.so): compute-sanitizer is closed on the GPU pool, so tests/test_gpu_pipeline.py runs
// scripts/sanitize_case.py -- every kernel once, ragged sizes -- against that build; a violated check traps the kernel.
#ifdef LDPCB_BOUNDS
#include <cassert>
#define LDPCB_ASSERT(cond) assert(cond)
#else
#define LDPCB_ASSERT(cond) ((void)0)
#endif

#define LDPCB_LAUNCH_CHECK(h, name)                                 \
    do {                                                            \
        (h)->launches++;                                            \
        int _st = ldpcb::check_cuda((h), cudaGetLastError(), name); \
        if (_st != LDPCB_OK) return _st;                            \
    } while (0)

// kernels' host launchers (device pointers, asynchronous)
struct NmsArgs {
    const float* llr;
    const int32_t* idx;  // optional frame indirection (NULL = identity)
    int64_t B;
    int iters;
    float alpha, w_vc, w_marg;
    int early_stop;
    uint32_t* hard_bits;
    uint8_t* iters_used;
    uint8_t* syndrome_nz;
    float* soft_traj;
    // optional: DIA reliability fused into the decoder, fir_out[f,j] = fir_bias + sum_i fir_taps[i] * posterior_i[f,j]
    // (posterior_0 = the input); excludes soft_traj and early_stop
    float* fir_out = nullptr;
    float fir_bias = 0.0f;
    float fir_taps[LDPCB_MAX_ITERS + 1] = {};
};
int launch_nms(ldpcb_handle* h, const NmsArgs& a, cudaStream_t st);

// Work fused into the quasi-cyclic NMS kernel (nms_qc.cu) by the device pipelines: tallies, the list of frames that go
// to OSD, and -- for the Monte-Carlo step -- the frame generator in front of the decoder.
struct NmsFuse {
    const uint32_t* truth = nullptr;   // [B,4] transmitted codewords (NULL and !gen: no tally)
    uint64_t* counters = nullptr;      // LDPCB_CNT_* block, accumulated
    int32_t* fail_idx = nullptr;       // frames with a non-zero syndrome, in completion order (not sorted)
    int32_t* fail_count = nullptr;     // zeroed by the caller before the launch
    int osd_follows = 0;               // 1: detected failures are tallied as final by the OSD kernel, not here
    int gen = 0;                       // 1: generate frame first_frame + f instead of loading a.llr
    uint2 key = {0u, 0u};
    uint64_t first_frame = 0;
    float sigma = 0.0f;
    const uint64_t* gcol = nullptr;
    float* fail_llr = nullptr;         // gen: [*,128] channel LLRs of the appended frames, row = position in fail_idx
    uint32_t* fail_truth = nullptr;    // gen: [*,4] their transmitted codewords
};
bool nms_qc_matches_code(const uint8_t* H);
bool nms_qc_applies(const ldpcb_handle* h, const NmsArgs& a);
int launch_nms_qc(ldpcb_handle* h, const NmsArgs& a, const NmsFuse* fuse, cudaStream_t st);

struct OsdArgs {
    const float* order_llr;
    const float* score_llr;
    const int32_t* idx;    // optional: frame i reads/writes row idx[i]
    const int32_t* count;  // optional device count (number of frames), else B
    const uint64_t* redG_in;  // optional [B,64] P' rows: frames are already permuted, skip sort + elimination
    int64_t B;             // upper bound on frames (grid sizing)
    const uint32_t* teps;
    int n_teps;
    int maxw;
    const uint16_t* triple_index;  // order-3 tables: [C(k,3) + C(j,2) + i] (i<j<k) -> index of the triple TEP (osd3.cu)
    const uint16_t* pair_index;  // optional (full order-0/1/2 tables): selects the warp-local sweep (orders 0, 1) or the tensor-core pair sweep (order 2)
    const int32_t* block_start;  // NULL => one block [0,n_teps)
    int n_blocks;
    int flags;
    uint32_t* cw_bits;
    int32_t* best_tep;
    int64_t* best_score_q;
    int32_t* score_exp;
    uint8_t* perm;
    uint64_t* redG;
    int64_t* block_min_q;
    int32_t* block_arg;
    const uint32_t* truth_bits;
    int64_t* truth_score_q;
    // fused tallies of the device pipelines (ldpcb_decode / ldpcb_simulate): decisions are compared with the transmitted
    // codewords in the output step and the LDPCB_CNT_OSD_* / FINAL_* / TEPS / PHASE counters accumulated by the kernel
    const uint32_t* tally_truth;   // [.,4]: row = the frame's original index (idx given) or its position in the batch
    uint64_t* tally_counters;
};
int launch_osd(ldpcb_handle* h, const OsdArgs& a, cudaStream_t st);
int launch_osd_pair(ldpcb_handle* h, const OsdArgs& a, cudaStream_t st);
int launch_osd_blocks(ldpcb_handle* h, const OsdArgs& a, cudaStream_t st);  // block minima: per-block truncated sweep + exact re-scoring (osd_blocks.cu)
int launch_osd3(ldpcb_handle* h, const OsdArgs& a, cudaStream_t st);      // order 3, full lists: tensor-core sweep of the 62 pair problems (osd3.cu)  // order 2, full lists: warp-local tensor-core pair sweep

// FS-OSD policy parameters (FS_OSD/fs_testing.py:92,129-160)
struct FsParams {
    float tau_e;
    int tau_psc;
    float beta_shift;  // beta * (n - k) as the reference adds it to an fp32 sum
    int order;
    int32_t* num_teps;
    uint8_t* stop_kind;  // 0: order-0 accepted, 1: tau_e stop inside a sweep, 2: skip rule, 3: all orders swept
    // order 3 in three launches (launch_osd_fs3): a frame that reaches the weight-3 class is not swept here but appended,
    // with the decision so far, to a list that the tensor-core sweep of osd3.cu works off
    int defer3 = 0;
    int32_t* d3_list = nullptr;    // original rows
    int32_t* d3_count = nullptr;
    long long* d3_wdmin = nullptr; // [list position]
    int32_t* d3_opt = nullptr;
    int32_t* d3_num = nullptr;
};
struct Fs3Args {
    const long long* wdmin;   // [list position] exact score of the decision after classes 0..2
    const int32_t* opt;       // its TEP index (FS enumeration)
    const int32_t* num;       // TEPs visited so far
    int he, hs;               // eligible iff |D| < he (= tau_psc - 3); stop iff |D| < hs (= ceil(tau_e) - 3)
    int32_t* num_teps;        // outputs by original row (may be NULL)
    uint8_t* stop_kind;
};
int launch_osd3_fs(ldpcb_handle* h, const OsdArgs& a, const Fs3Args& fs, int32_t* fb_list, int32_t* fb_count, cudaStream_t st);
int launch_osd_fs3(ldpcb_handle* h, const OsdArgs& a, const FsParams& fp, cudaStream_t st);  // FS policy, order_limit 3
int launch_osd_fs(ldpcb_handle* h, const OsdArgs& a, const FsParams& fp, cudaStream_t st);

// PB-OSD policy parameters (PB_OSD/pb_testing.py:44-52,100-149)
struct PbParams {
    float c4;             // -4 * noise_variance, noise_variance = 10^(-snr/10) (pb_testing.py:51-52)
    int order;            // order_limit (0..2)
    int32_t* stats;       // [B,4] out: TEPs visited, p_e^pro passes, improvements, list comparisons
    double cdf_half[65];  // BinCDF(b; 64, 1/2)
    long long* glist_sum; // order 3: per-warp TEP lists in global memory (set by the launcher)
    long long* glist_bmin; // order 3: minima of every 32 list entries
    unsigned* glist_tep;
    int* queue;           // next frame to decode (dynamic work distribution; zeroed by the launcher)
};
int launch_osd_pb(ldpcb_handle* h, const OsdArgs& a, const PbParams& pp, cudaStream_t st);

int build_tep_tables(ldpcb_handle* h);

}  // namespace ldpcb

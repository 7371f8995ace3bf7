// Handle lifetime, code tables, TEP enumerations, error reporting.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>

#include "common.cuh"

namespace ldpcb {

static thread_local std::string g_create_error;

int set_error(ldpcb_handle* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf; else g_create_error = buf;
    return code;
}

int check_cuda(ldpcb_handle* h, cudaError_t e, const char* what) {
    if (e == cudaSuccess) return LDPCB_OK;
    return set_error(h, LDPCB_ERR_CUDA, "CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
}

static int grow(ldpcb_handle* h, Workspace& w, size_t bytes);
int ensure_ws(ldpcb_handle* h, int slot, size_t bytes) { return grow(h, h->ws[slot], bytes); }

static int grow(ldpcb_handle* h, Workspace& w, size_t bytes) {
    if (w.cap >= bytes) return LDPCB_OK;
    if (w.buf) {
        // the buffer may still be in use by queued work of this handle
        LDPCB_CUDA(h, cudaDeviceSynchronize());
        LDPCB_CUDA(h, cudaFree(w.buf));
        w.buf = nullptr;
        w.cap = 0;
    }
    size_t cap = bytes + (bytes >> 2) + 4096;
    LDPCB_CUDA(h, cudaMalloc(&w.buf, cap));
    w.cap = cap;
    return LDPCB_OK;
}

int ensure_stream_ws(ldpcb_handle* h, cudaStream_t st, size_t bytes, char** buf) {
    Workspace& w = h->stream_ws[st];
    int s = grow(h, w, bytes);
    *buf = w.buf;
    return s;
}

// ---- TEP enumerations -------------------------------------------------------------------------
// Packed TEP: byte i = i-th MRB position of the support (ascending), 0xFF = unused.
static uint32_t pack_tep(const int* idx, int w) {
    uint32_t v = 0xFFFFFFFFu;
    for (int i = 0; i < w; ++i) v = (v & ~(0xFFu << (8 * i))) | ((uint32_t)idx[i] << (8 * i));
    return v;
}

// all combinations of {0..63} of weight w in lexicographic order (itertools.combinations)
static void combos(int w, std::vector<std::vector<int>>& out) {
    std::vector<int> c(w);
    for (int i = 0; i < w; ++i) c[i] = i;
    if (w == 0) { out.push_back(c); return; }
    while (true) {
        out.push_back(c);
        int i = w - 1;
        while (i >= 0 && c[i] == K - w + i) --i;
        if (i < 0) break;
        ++c[i];
        for (int j = i + 1; j < w; ++j) c[j] = c[j - 1] + 1;
    }
}

// conventional order: PB_OSD/convention_osd.py:13-38 (descending index sum, stable)
static void build_conv(int order, std::vector<uint32_t>& out) {
    for (int w = 0; w <= order; ++w) {
        std::vector<std::vector<int>> cs;
        combos(w, cs);
        std::stable_sort(cs.begin(), cs.end(), [](const std::vector<int>& a, const std::vector<int>& b) {
            int sa = 0, sb = 0;
            for (int x : a) sa += x;
            for (int x : b) sb += x;
            return sa > sb;
        });
        for (auto& c : cs) out.push_back(pack_tep(c.data(), w));
    }
}

// FS order: FS_OSD/fs_testing.py:32-49 (vector reversed => position 63-c), all-zero TEP first (:131-132)
static void build_fs(int order, std::vector<uint32_t>& out) {
    int none = 0;
    out.push_back(pack_tep(&none, 0));
    for (int w = 1; w <= order; ++w) {
        std::vector<std::vector<int>> cs;
        combos(w, cs);
        for (auto& c : cs) {
            int r[4];
            for (int i = 0; i < w; ++i) r[i] = K - 1 - c[w - 1 - i];  // ascending after reversal
            out.push_back(pack_tep(r, w));
        }
    }
}

int build_tep_tables(ldpcb_handle* h) {
    for (int order = 0; order <= 3; ++order) {
        for (int kind = 0; kind < 2; ++kind) {
            TepTable& t = h->tep[order][kind];
            t.host.clear();
            if (kind == LDPCB_TEP_CONV) build_conv(order, t.host); else build_fs(order, t.host);
            t.n = (int)t.host.size();
            t.maxw = order < 1 ? 1 : order;
            // the device copy is padded with 128 "no position" words: the sweep loads whole 128-TEP tiles unconditionally
            std::vector<uint32_t> padded(t.host);
            padded.resize(t.n + 128, 0xFFFFFFFFu);
            LDPCB_CUDA(h, cudaMalloc(&t.dev, sizeof(uint32_t) * padded.size()));
            LDPCB_CUDA(h, cudaMemcpy(t.dev, padded.data(), sizeof(uint32_t) * padded.size(), cudaMemcpyHostToDevice));
            {
                std::vector<uint16_t> inv(OSD_PAIR_TABLE, 0xFFFFu);
                std::vector<uint16_t> inv3(K * (K - 1) * (K - 2) / 6, 0xFFFFu);
                for (int i = 0; i < t.n; ++i) {
                    const uint32_t v = t.host[i];
                    const unsigned a = v & 0xFFu, b = (v >> 8) & 0xFFu, c = (v >> 16) & 0xFFu;
                    if (a >= (unsigned)K) inv[K * K + K] = (uint16_t)i;
                    else if (b >= (unsigned)K) inv[K * K + a] = (uint16_t)i;
                    else if (c >= (unsigned)K) inv[a * K + b] = (uint16_t)i;
                    else inv3[c * (c - 1) * (c - 2) / 6 + b * (b - 1) / 2 + a] = (uint16_t)i;  // positions ascend: a < b < c
                }
                if (order == 3) {
                    LDPCB_CUDA(h, cudaMalloc(&t.triple_dev, sizeof(uint16_t) * inv3.size()));
                    LDPCB_CUDA(h, cudaMemcpy(t.triple_dev, inv3.data(), sizeof(uint16_t) * inv3.size(), cudaMemcpyHostToDevice));
                }
                LDPCB_CUDA(h, cudaMalloc(&t.pair_dev, sizeof(uint16_t) * inv.size()));
                LDPCB_CUDA(h, cudaMemcpy(t.pair_dev, inv.data(), sizeof(uint16_t) * inv.size(), cudaMemcpyHostToDevice));
            }
        }
    }
    return LDPCB_OK;
}

// ---- code tables ------------------------------------------------------------------------------
// Choose the edge labels of every check and the per-label bank rotations so that both gathers of the NMS kernel are
// (nearly) free of shared-memory bank conflicts:
//   check side    row (q, e): lane l reads T[variable of the edge labelled e of check l + 32q], bank = variable mod 32
//   variable side row (k, d): lane l reads the d-th incoming message (ascending check order) of variable l + 32k,
//                 stored at bank (check + rot[label]) mod 32
// Labels are permuted per "class" = (block row of 16 checks, block column of 16 variables, circulant shift), which
// keeps the regularity a quasi-cyclic H gives the check side; for an unstructured H every edge is its own class.
// Simulated annealing on the number of colliding lanes from the natural labelling, deterministic, ~0.2 s, best kept.
static void optimise_nms_layout(const uint8_t* Hm, int label_of[M][N], int rot[DC]) {
    std::vector<std::vector<int>> chk(M), var(N);
    for (int c = 0; c < M; ++c)
        for (int v = 0; v < N; ++v)
            if (Hm[c * N + v]) { chk[c].push_back(v); var[v].push_back(c); }
    // classes per block row
    struct Cls { int C, s; std::vector<std::pair<int, int>> edges; };
    std::vector<std::vector<Cls>> cls(M / 16);
    bool structured = true;
    for (int R = 0; R < M / 16; ++R) {
        for (int c = 16 * R; c < 16 * R + 16; ++c)
            for (int v : chk[c]) {
                const int C = v / 16, sft = ((v % 16) - (c % 16) + 16) % 16;
                Cls* f = nullptr;
                for (auto& x : cls[R]) if (x.C == C && x.s == sft) f = &x;
                if (!f) { cls[R].push_back({C, sft, {}}); f = &cls[R].back(); }
                f->edges.push_back({c, v});
            }
        if ((int)cls[R].size() > DC) structured = false;
        for (auto& x : cls[R]) if (x.edges.size() != 16) structured = false;
    }
    // natural labelling
    for (int c = 0; c < M; ++c)
        for (size_t e = 0; e < chk[c].size(); ++e) label_of[c][chk[c][e]] = (int)e;
    for (int e = 0; e < DC; ++e) rot[e] = 0;
    if (structured) {  // start from one label per class (ascending block column, shift)
        for (int R = 0; R < M / 16; ++R) {
            std::sort(cls[R].begin(), cls[R].end(), [](const Cls& a, const Cls& b) { return a.C != b.C ? a.C < b.C : a.s < b.s; });
            for (size_t i = 0; i < cls[R].size(); ++i)
                for (auto& e : cls[R][i].edges) label_of[e.first][e.second] = (int)i;
        }
    }
    auto cost = [&]() {
        int tot = 0;
        for (int k = 0; k < 4; ++k)
            for (int d = 0; d < DV; ++d) {
                int cnt[32] = {0};
                for (int l = 0; l < 32; ++l) {
                    const int v = l + 32 * k;
                    if (d < (int)var[v].size()) { const int c = var[v][d]; ++cnt[(c + rot[label_of[c][v]]) & 31]; }
                }
                for (int b = 0; b < 32; ++b) tot += cnt[b] > 1 ? cnt[b] - 1 : 0;
            }
        for (int q = 0; q < 2; ++q) {
            int cnt[DC][32];
            memset(cnt, 0, sizeof cnt);
            for (int l = 0; l < 32; ++l) {
                const int c = l + 32 * q;
                for (int v : chk[c]) ++cnt[label_of[c][v]][v & 31];
            }
            for (int e = 0; e < DC; ++e)
                for (int b = 0; b < 32; ++b) tot += cnt[e][b] > 1 ? cnt[e][b] - 1 : 0;
        }
        return tot;
    };
    uint64_t rng = 0x9E3779B97F4A7C15ull;
    auto next = [&]() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return rng; };
    auto unif = [&]() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); };
    int cur = cost(), best = cur, best_rot[DC];
    static thread_local int best_label[M][N];
    memcpy(best_label, label_of, sizeof best_label);
    memcpy(best_rot, rot, sizeof best_rot);
    double T = 2.0;
    for (int it = 0; it < 60000 && best > 0; ++it) {
        T = std::max(0.05, T * 0.9999);
        if (it % 15000 == 14999) {  // restart from the best so far with a warm temperature
            memcpy(label_of, best_label, sizeof best_label);
            memcpy(rot, best_rot, sizeof best_rot);
            cur = best;
            T = 1.0;
        }
        if (next() % 5 < 2) {
            const int e = (int)(next() % DC), old = rot[e];
            rot[e] = (int)(next() % 32);
            const int nw = cost();
            if (nw <= cur || unif() < exp((cur - nw) / T)) cur = nw; else rot[e] = old;
        } else if (structured) {
            const int R = (int)(next() % (M / 16));
            const int a = (int)(next() % cls[R].size()), b = (int)(next() % cls[R].size());
            if (a == b) continue;
            const int la = label_of[cls[R][a].edges[0].first][cls[R][a].edges[0].second];
            const int lb = label_of[cls[R][b].edges[0].first][cls[R][b].edges[0].second];
            auto apply = [&](int x, int y) {
                for (auto& e : cls[R][a].edges) label_of[e.first][e.second] = x;
                for (auto& e : cls[R][b].edges) label_of[e.first][e.second] = y;
            };
            apply(lb, la);
            const int nw = cost();
            if (nw <= cur || unif() < exp((cur - nw) / T)) cur = nw; else apply(la, lb);
        } else {
            const int c = (int)(next() % M), d = (int)chk[c].size();
            const int v1 = chk[c][next() % d], v2 = chk[c][next() % d];
            if (v1 == v2) continue;
            std::swap(label_of[c][v1], label_of[c][v2]);
            const int nw = cost();
            if (nw <= cur || unif() < exp((cur - nw) / T)) cur = nw; else std::swap(label_of[c][v1], label_of[c][v2]);
        }
        if (cur < best) {
            best = cur;
            memcpy(best_label, label_of, sizeof best_label);
            memcpy(best_rot, rot, sizeof best_rot);
        }
    }
    memcpy(label_of, best_label, sizeof best_label);
    memcpy(rot, best_rot, sizeof best_rot);
}

static int build_code_tables(ldpcb_handle* h) {
    NmsTables& t = h->nms_host;
    memset(&t, 0, sizeof t);
    static thread_local int label_of[M][N];
    for (int c = 0; c < M; ++c) {
        int e = 0;
        for (int v = 0; v < N; ++v) {
            if (h->H[c * N + v]) {
                if (e >= DC) return set_error(h, LDPCB_ERR_SHAPE, "check %d has degree > %d", c, DC);
                t.chk_mask[c][v >> 5] |= 1u << (v & 31);
                ++e;
            }
        }
        if (e < 2) return set_error(h, LDPCB_ERR_SHAPE, "check %d has degree < 2", c);
    }
    for (int v = 0; v < N; ++v) {
        int d = 0;
        for (int c = 0; c < M; ++c) d += h->H[c * N + v];
        if (d > DV) return set_error(h, LDPCB_ERR_SHAPE, "variable %d has degree > %d", v, DV);
    }
    optimise_nms_layout(h->H, label_of, t.rot);
    for (int c = 0; c < M; ++c) {
        for (int e = 0; e < DC; ++e) t.chk_var[c][e] = 128;
        for (int v = 0; v < N; ++v)
            if (h->H[c * N + v]) t.chk_var[c][label_of[c][v]] = (uint8_t)v;
    }
    t.max_var_deg_lo = t.max_var_deg_hi = 0;
    for (int v = 0; v < N; ++v) {
        int d = 0;
        for (int c = 0; c < M; ++c) {
            if (h->H[c * N + v]) {
                const int e = label_of[c][v];
                t.var_slot[v][d++] = (uint16_t)(e * NMS_CV_STRIDE + t.rot[e] + c);
            }
        }
        int& mx = (v < 64) ? t.max_var_deg_lo : t.max_var_deg_hi;
        mx = std::max(mx, d);
        for (; d < DV; ++d) t.var_slot[v][d] = NMS_CV_ZERO;
    }
    // G columns, and H.G^T = 0, rank(G) = K
    for (int j = 0; j < N; ++j) {
        uint64_t c = 0;
        for (int r = 0; r < K; ++r) c |= (uint64_t)(h->G[r * N + j] & 1) << r;
        h->gcol_host[j] = c;
    }
    for (int c = 0; c < M; ++c)
        for (int r = 0; r < K; ++r) {
            int s = 0;
            for (int v = 0; v < N; ++v) s ^= h->H[c * N + v] & h->G[r * N + v];
            if (s) return set_error(h, LDPCB_ERR_CODE, "H.G^T != 0 at check %d, generator row %d", c, r);
        }
    {
        uint64_t cols[N];
        memcpy(cols, h->gcol_host, sizeof cols);
        uint64_t used = 0;
        int rank = 0;
        for (int j = 0; j < N && rank < K; ++j) {
            uint64_t cand = cols[j] & ~used;
            if (!cand) continue;
            int p = __builtin_ctzll(cand);
            used |= 1ull << p;
            ++rank;
            uint64_t m = cols[j] ^ (1ull << p);
            for (int x = 0; x < N; ++x)
                if ((cols[x] >> p) & 1) cols[x] ^= m;
        }
        if (rank != K) return set_error(h, LDPCB_ERR_CODE, "G has rank %d, expected %d", rank, K);
    }
    LDPCB_CUDA(h, cudaMalloc(&h->nms_dev, sizeof(NmsTables)));
    LDPCB_CUDA(h, cudaMemcpy(h->nms_dev, &t, sizeof(NmsTables), cudaMemcpyHostToDevice));
    {
        // [0, N): the columns; [N, 2N): 1 for the unit columns the OSD elimination may treat as such -- one per generator
        // row (the first in index order), see osd_prepare.cuh step 2
        uint64_t tab[2 * N];
        memcpy(tab, h->gcol_host, sizeof(uint64_t) * N);
        uint64_t rows_seen = 0;
        for (int j = 0; j < N; ++j) {
            const uint64_t c = h->gcol_host[j];
            const bool unit = c != 0 && (c & (c - 1)) == 0 && !(rows_seen & c);
            if (unit) rows_seen |= c;
            tab[N + j] = unit ? 1 : 0;
        }
        LDPCB_CUDA(h, cudaMalloc(&h->gcol_dev, sizeof tab));
        LDPCB_CUDA(h, cudaMemcpy(h->gcol_dev, tab, sizeof tab, cudaMemcpyHostToDevice));
    }
    return LDPCB_OK;
}

}  // namespace ldpcb

using namespace ldpcb;

extern "C" {

int ldpcb_abi_version(void) { return LDPCB_ABI_VERSION; }

int ldpcb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int ldpcb_create(ldpcb_t** out, const uint8_t* H_host, const uint8_t* G_host, int n, int m, int k, int device) {
    if (!out || !H_host || !G_host) return set_error(nullptr, LDPCB_ERR_ARG, "ldpcb_create: NULL argument");
    *out = nullptr;
    if (n != N || m != M || k != K)
        return set_error(nullptr, LDPCB_ERR_SHAPE, "ldpcb_create: only n=%d m=%d k=%d is supported (got %d,%d,%d)", N, M, K, n, m, k);
    int ndev = ldpcb_device_count();
    if (ndev <= 0)
        return set_error(nullptr, LDPCB_ERR_NO_DEVICE, "ldpcb_create: no CUDA device (this library has no CPU fallback)");
    if (device < 0 || device >= ndev)
        return set_error(nullptr, LDPCB_ERR_ARG, "ldpcb_create: device %d out of range (have %d)", device, ndev);
    ldpcb_handle* h = new ldpcb_handle();
    h->device = device;
    int st = LDPCB_OK;
    DeviceGuard guard(h);  // the caller's current device is restored when ldpcb_create returns
    auto fail = [&](int code) {
        g_create_error = h->err;
        ldpcb_destroy(h);
        return code;
    };
    if ((st = check_cuda(h, cudaSetDevice(device), "cudaSetDevice")) != LDPCB_OK) return fail(st);
    cudaDeviceProp prop;
    if ((st = check_cuda(h, cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties")) != LDPCB_OK) return fail(st);
    if (prop.major < 10) {
        set_error(h, LDPCB_ERR_NO_DEVICE, "ldpcb_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
        return fail(LDPCB_ERR_NO_DEVICE);
    }
    h->sm_count = prop.multiProcessorCount;
    for (int i = 0; i < m * n; ++i) h->H[i] = H_host[i] & 1;
    for (int i = 0; i < k * n; ++i) h->G[i] = G_host[i] & 1;
    if ((st = build_code_tables(h)) != LDPCB_OK) return fail(st);
    if ((st = build_tep_tables(h)) != LDPCB_OK) return fail(st);
    if (const char* e = getenv("LDPCB_QC_MINB")) h->qc_minb = atoi(e) == 3 ? 3 : 2;
    h->qc_ccsds = nms_qc_matches_code(h->H) && !getenv("LDPCB_NMS_GENERIC");  // env: force the table-driven kernel (tests, A/B timing)
    if ((st = check_cuda(h, cudaMalloc(&h->one_block_dev, 2 * sizeof(int32_t)), "cudaMalloc")) != LDPCB_OK) return fail(st);
    if ((st = check_cuda(h, cudaMalloc(&h->pb_queue, 256), "cudaMalloc")) != LDPCB_OK) return fail(st);
    for (int i = 0; i < 3; ++i) {
        if ((st = check_cuda(h, cudaStreamCreateWithFlags(&h->streams[i], cudaStreamNonBlocking), "cudaStreamCreate")) != LDPCB_OK) return fail(st);
        if ((st = check_cuda(h, cudaEventCreateWithFlags(&h->events[i], cudaEventDisableTiming), "cudaEventCreate")) != LDPCB_OK) return fail(st);
    }
    *out = h;
    return LDPCB_OK;
}

void ldpcb_destroy(ldpcb_t* h) {
    if (!h) return;
    DeviceGuard guard(h);
    cudaDeviceSynchronize();
    for (int o = 0; o < 4; ++o)
        for (int kd = 0; kd < 2; ++kd) {
            if (h->tep[o][kd].dev) cudaFree(h->tep[o][kd].dev);
            if (h->tep[o][kd].pair_dev) cudaFree(h->tep[o][kd].pair_dev);
            if (h->tep[o][kd].triple_dev) cudaFree(h->tep[o][kd].triple_dev);
        }
    for (int i = 0; i < NUM_WS; ++i)
        if (h->ws[i].buf) cudaFree(h->ws[i].buf);
    for (auto& kv : h->stream_ws)
        if (kv.second.buf) cudaFree(kv.second.buf);
    for (auto& kv : h->fb_ws)
        if (kv.second.buf) cudaFree(kv.second.buf);
    if (h->pb_list) cudaFree(h->pb_list);
    if (h->pb_queue) cudaFree(h->pb_queue);
    for (int i = 0; i < 3; ++i) {
        if (h->streams[i]) cudaStreamDestroy(h->streams[i]);
        if (h->events[i]) cudaEventDestroy(h->events[i]);
    }
    if (h->nms_dev) cudaFree(h->nms_dev);
    if (h->gcol_dev) cudaFree(h->gcol_dev);
    if (h->one_block_dev) cudaFree(h->one_block_dev);
    cudaGetLastError();
    delete h;
}

const char* ldpcb_last_error(ldpcb_t* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int ldpcb_sm_count(ldpcb_t* h) { return h ? h->sm_count : LDPCB_ERR_ARG; }

uint64_t ldpcb_launch_count(ldpcb_t* h) { return h ? h->launches : 0; }

int ldpcb_tep_count(ldpcb_t* h, int order, int tep_order) {
    if (!h) return LDPCB_ERR_ARG;
    if (order < 0 || order > 3 || tep_order < 0 || tep_order > 1)
        return set_error(h, LDPCB_ERR_ARG, "ldpcb_tep_count: order %d / tep_order %d out of range", order, tep_order);
    return h->tep[order][tep_order].n;
}

int ldpcb_tep_table(ldpcb_t* h, int order, int tep_order, uint32_t* teps_host) {
    int n = ldpcb_tep_count(h, order, tep_order);
    if (n < 0) return n;
    if (!teps_host) return set_error(h, LDPCB_ERR_ARG, "ldpcb_tep_table: NULL output");
    memcpy(teps_host, h->tep[order][tep_order].host.data(), sizeof(uint32_t) * n);
    return LDPCB_OK;
}

int ldpcb_device_pci_bus_id(int device, char* buf, int len) {
    if (!buf || len < 13) return LDPCB_ERR_ARG;
    cudaError_t e = cudaDeviceGetPCIBusId(buf, len, device);
    if (e != cudaSuccess) { cudaGetLastError(); buf[0] = 0; return set_error(nullptr, LDPCB_ERR_CUDA, "cudaDeviceGetPCIBusId(%d): %s", device, cudaGetErrorString(e)); }
    return LDPCB_OK;
}

int ldpcb_host_alloc(void** p, uint64_t bytes) {
    if (!p) return LDPCB_ERR_ARG;
    cudaError_t e = cudaMallocHost(p, bytes);
    if (e != cudaSuccess) { *p = nullptr; return set_error(nullptr, LDPCB_ERR_CUDA, "cudaMallocHost(%llu): %s", (unsigned long long)bytes, cudaGetErrorString(e)); }
    return LDPCB_OK;
}

int ldpcb_host_free(void* p) {
    if (!p) return LDPCB_OK;
    return cudaFreeHost(p) == cudaSuccess ? LDPCB_OK : LDPCB_ERR_CUDA;
}

}  // extern "C"

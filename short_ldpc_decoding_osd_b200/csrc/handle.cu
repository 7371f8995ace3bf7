// Handle lifetime, code tables, TEP enumerations, error reporting.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
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

int ensure_ws(ldpcb_handle* h, int slot, size_t bytes) {
    Workspace& w = h->ws[slot];
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
            LDPCB_CUDA(h, cudaMalloc(&t.dev, sizeof(uint32_t) * t.n));
            LDPCB_CUDA(h, cudaMemcpy(t.dev, t.host.data(), sizeof(uint32_t) * t.n, cudaMemcpyHostToDevice));
        }
    }
    return LDPCB_OK;
}

// ---- code tables ------------------------------------------------------------------------------
static int build_code_tables(ldpcb_handle* h) {
    NmsTables& t = h->nms_host;
    memset(&t, 0, sizeof t);
    int pos_in_check[M][N];
    for (int c = 0; c < M; ++c) {
        int e = 0;
        for (int v = 0; v < N; ++v) {
            pos_in_check[c][v] = -1;
            if (h->H[c * N + v]) {
                if (e >= DC) return set_error(h, LDPCB_ERR_SHAPE, "check %d has degree > %d", c, DC);
                t.chk_var[c][e] = (uint8_t)v;
                pos_in_check[c][v] = e;
                t.chk_mask[c][v >> 5] |= 1u << (v & 31);
                ++e;
            }
        }
        if (e < 2) return set_error(h, LDPCB_ERR_SHAPE, "check %d has degree < 2", c);
        for (; e < DC; ++e) t.chk_var[c][e] = 128;
    }
    t.max_var_deg_lo = t.max_var_deg_hi = 0;
    for (int v = 0; v < N; ++v) {
        int d = 0;
        for (int c = 0; c < M; ++c) {
            if (h->H[c * N + v]) {
                if (d >= DV) return set_error(h, LDPCB_ERR_SHAPE, "variable %d has degree > %d", v, DV);
                t.var_slot[v][d++] = (uint16_t)(pos_in_check[c][v] * M + c);
            }
        }
        int& mx = (v < 64) ? t.max_var_deg_lo : t.max_var_deg_hi;
        mx = std::max(mx, d);
        for (; d < DV; ++d) t.var_slot[v][d] = 512;
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
    LDPCB_CUDA(h, cudaMalloc(&h->gcol_dev, sizeof(uint64_t) * N));
    LDPCB_CUDA(h, cudaMemcpy(h->gcol_dev, h->gcol_host, sizeof(uint64_t) * N, cudaMemcpyHostToDevice));
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
    if ((st = check_cuda(h, cudaMalloc(&h->one_block_dev, 2 * sizeof(int32_t)), "cudaMalloc")) != LDPCB_OK) return fail(st);
    for (int i = 0; i < 3; ++i) {
        if ((st = check_cuda(h, cudaStreamCreateWithFlags(&h->streams[i], cudaStreamNonBlocking), "cudaStreamCreate")) != LDPCB_OK) return fail(st);
        if ((st = check_cuda(h, cudaEventCreateWithFlags(&h->events[i], cudaEventDisableTiming), "cudaEventCreate")) != LDPCB_OK) return fail(st);
    }
    *out = h;
    return LDPCB_OK;
}

void ldpcb_destroy(ldpcb_t* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (int o = 0; o < 4; ++o)
        for (int kd = 0; kd < 2; ++kd)
            if (h->tep[o][kd].dev) cudaFree(h->tep[o][kd].dev);
    for (int i = 0; i < NUM_WS; ++i)
        if (h->ws[i].buf) cudaFree(h->ws[i].buf);
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

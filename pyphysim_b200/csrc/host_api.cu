// host_api.cu — host-buffer entry points of the C ABI (what a non-CUDA caller binds; bench `e2e`).
// Units are processed in chunks that ping-pong between two internal streams, each with its own
// device staging buffers, so the H2D copy of chunk c+1 overlaps the kernel of chunk c and the D2H
// of chunk c-1.  Pinned host memory makes the copies truly asynchronous; pageable memory still works.
#include <mutex>
#include <vector>

#include "common.cuh"

namespace b200phy {

struct Slot {
    cudaStream_t st = nullptr;
    void *buf[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t cap[5] = {0, 0, 0, 0, 0};
    long long *counters = nullptr;
};

struct HostCtx {
    std::mutex mu;
    int dev = -1;
    Slot slot[2];
    void *table = nullptr;
};

static HostCtx g_ctx;

static int ctx_init(HostCtx &c) {
    int dev = 0;
    int e = check_cuda(cudaGetDevice(&dev), "cudaGetDevice");
    if (e) return e;
    if (c.dev == dev) return B200PHY_OK;
    if (c.dev >= 0) {
        // the caller switched devices (one process per GPU is the normal deployment): release what lives on
        // the previous device before building the context on the current one
        cudaSetDevice(c.dev);
        for (auto &s : c.slot) {
            if (s.st) { cudaStreamSynchronize(s.st); cudaStreamDestroy(s.st); s.st = nullptr; }
            for (int i = 0; i < 5; ++i) { if (s.buf[i]) cudaFree(s.buf[i]); s.buf[i] = nullptr; s.cap[i] = 0; }
            if (s.counters) { cudaFree(s.counters); s.counters = nullptr; }
        }
        if (c.table) { cudaFree(c.table); c.table = nullptr; }
        c.dev = -1;
        if ((e = check_cuda(cudaSetDevice(dev), "cudaSetDevice"))) return e;
    }
    for (auto &s : c.slot) {
        if ((e = check_cuda(cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking), "cudaStreamCreate"))) return e;
        if ((e = check_cuda(cudaMalloc((void **)&s.counters, 4 * sizeof(long long)), "cudaMalloc(counters)"))) return e;
        for (int i = 0; i < 5; ++i) { s.buf[i] = nullptr; s.cap[i] = 0; }
    }
    if ((e = check_cuda(cudaMalloc(&c.table, 256 * 16), "cudaMalloc(table)"))) return e;
    c.dev = dev;
    return B200PHY_OK;
}

static int ensure(Slot &s, int i, size_t bytes) {
    if (bytes <= s.cap[i]) return B200PHY_OK;
    if (s.buf[i]) cudaFree(s.buf[i]);
    s.buf[i] = nullptr; s.cap[i] = 0;
    int e = check_cuda(cudaMalloc(&s.buf[i], bytes), "cudaMalloc(staging)");
    if (!e) s.cap[i] = bytes;
    return e;
}

static int upload_table(HostCtx &c, int dtype, int kind, int M, const double *table_re_im, b200phy_modem *out) {
    out->kind = kind; out->M = M; out->table = c.table;
    if (kind == B200PHY_MODEM_BPSK) return B200PHY_OK;
    if (!table_re_im) { set_error("table_re_im is NULL"); return B200PHY_ERR_INVALID; }
    if (M < 2 || M > 256) { set_error("M out of range"); return B200PHY_ERR_INVALID; }
    if (dtype == B200PHY_F32) {
        std::vector<float> t(2 * M);
        for (int i = 0; i < 2 * M; ++i) t[i] = float(table_re_im[i]);
        return check_cuda(cudaMemcpy(c.table, t.data(), sizeof(float) * 2 * M, cudaMemcpyHostToDevice), "cudaMemcpy(table)");
    }
    return check_cuda(cudaMemcpy(c.table, table_re_im, sizeof(double) * 2 * M, cudaMemcpyHostToDevice), "cudaMemcpy(table)");
}

static int finish(HostCtx &c, int64_t *counters) {
    int e;
    long long h[2][4];
    for (int s = 0; s < 2; ++s) {
        if ((e = check_cuda(cudaStreamSynchronize(c.slot[s].st), "cudaStreamSynchronize"))) return e;
        if ((e = check_cuda(cudaMemcpy(h[s], c.slot[s].counters, sizeof(h[s]), cudaMemcpyDeviceToHost), "cudaMemcpy(counters)"))) return e;
    }
    for (int i = 0; i < 4; ++i) counters[i] += h[0][i] + h[1][i];
    return B200PHY_OK;
}

}  // namespace b200phy

using namespace b200phy;

#define B200_TRY(x) do { int _e = (x); if (_e) return _e; } while (0)
#define B200_CU(x, what) B200_TRY(check_cuda((x), what))

extern "C" {

int b200phy_link_siso_flat_host(int dtype, int modem_kind, int M, const double *table_re_im,
                                int rayleigh, double noise_var, uint64_t seed, uint64_t first_unit,
                                int64_t n_units, const uint8_t *idx, const void *h, const void *noise,
                                uint8_t *idx_hat, int64_t *counters) {
    if (!counters) { set_error("counters is NULL"); return B200PHY_ERR_INVALID; }
    if (dtype != B200PHY_F32 && dtype != B200PHY_F64) { set_error("bad dtype"); return B200PHY_ERR_INVALID; }
    std::lock_guard<std::mutex> lk(g_ctx.mu);
    B200_TRY(ctx_init(g_ctx));
    b200phy_modem modem;
    B200_TRY(upload_table(g_ctx, dtype, modem_kind, M, table_re_im, &modem));
    const size_t csz = dtype == B200PHY_F32 ? 8 : 16;
    const int64_t chunk = int64_t(1) << 22;
    for (auto &s : g_ctx.slot) B200_CU(cudaMemsetAsync(s.counters, 0, 4 * sizeof(long long), s.st), "memset");
    int ci = 0;
    for (int64_t off = 0; off < n_units; off += chunk, ++ci) {
        Slot &s = g_ctx.slot[ci & 1];
        const int64_t n = n_units - off < chunk ? n_units - off : chunk;
        const uint8_t *d_idx = nullptr;
        const void *d_h = nullptr, *d_n = nullptr;
        if (idx) {
            B200_TRY(ensure(s, 0, n));
            B200_TRY(ensure(s, 2, n * csz));
            B200_CU(cudaMemcpyAsync(s.buf[0], idx + off, n, cudaMemcpyHostToDevice, s.st), "H2D idx");
            B200_CU(cudaMemcpyAsync(s.buf[2], (const char *)noise + off * csz, n * csz, cudaMemcpyHostToDevice, s.st), "H2D noise");
            d_idx = (const uint8_t *)s.buf[0]; d_n = s.buf[2];
            if (rayleigh) {
                B200_TRY(ensure(s, 1, n * csz));
                B200_CU(cudaMemcpyAsync(s.buf[1], (const char *)h + off * csz, n * csz, cudaMemcpyHostToDevice, s.st), "H2D h");
                d_h = s.buf[1];
            }
        }
        uint8_t *d_hat = nullptr;
        if (idx_hat) { B200_TRY(ensure(s, 3, n)); d_hat = (uint8_t *)s.buf[3]; }
        B200_TRY(b200phy_link_siso_flat(dtype, &modem, rayleigh, noise_var, seed, first_unit + off, n, d_idx, d_h, d_n, d_hat, nullptr, (int64_t *)s.counters, s.st));
        if (idx_hat) B200_CU(cudaMemcpyAsync(idx_hat + off, d_hat, n, cudaMemcpyDeviceToHost, s.st), "D2H idx_hat");
    }
    return finish(g_ctx, counters);
}

int b200phy_link_ofdm_tdl_host(const b200phy_ofdm_tdl_params *p, int modem_kind, int M,
                               const double *table_re_im, uint64_t first_unit, int64_t n_units,
                               const uint8_t *idx, const void *phi, const void *psi,
                               const void *noise, uint8_t *idx_hat, int64_t *counters) {
    if (!counters) { set_error("counters is NULL"); return B200PHY_ERR_INVALID; }
    if (!p) { set_error("params is NULL"); return B200PHY_ERR_INVALID; }
    if (p->dtype != B200PHY_F32 && p->dtype != B200PHY_F64) { set_error("bad dtype"); return B200PHY_ERR_INVALID; }
    if (p->n_taps < 1 || p->n_taps > B200PHY_MAX_TAPS) { set_error("bad n_taps"); return B200PHY_ERR_INVALID; }
    std::lock_guard<std::mutex> lk(g_ctx.mu);
    B200_TRY(ctx_init(g_ctx));
    b200phy_modem modem;
    B200_TRY(upload_table(g_ctx, p->dtype, modem_kind, M, table_re_im, &modem));
    const size_t rsz = p->dtype == B200PHY_F32 ? 4 : 8, csz = 2 * rsz;
    const size_t n_data = size_t(p->Nt) * p->n_sym * p->used;
    const size_t P = size_t(p->L) * p->n_taps * p->Nr * p->Nt;
    const size_t nrow = size_t(p->Nr) * (size_t(p->n_sym) * (p->fft + p->cp) + p->delays[p->n_taps - 1]);
    const size_t per_frame = n_data + 2 * P * rsz + nrow * csz;
    // ~64 MiB of draws per chunk: long enough to run PCIe at full rate and fill the GPU (>= 2000 frames),
    // short enough that the un-overlapped head (first H2D) and tail (last kernel + D2H) stay small
    int64_t chunk = int64_t((size_t(64) << 20) / per_frame);
    if (chunk < 1) chunk = 1;
    for (auto &s : g_ctx.slot) B200_CU(cudaMemsetAsync(s.counters, 0, 4 * sizeof(long long), s.st), "memset");
    int ci = 0;
    for (int64_t off = 0; off < n_units; off += chunk, ++ci) {
        Slot &s = g_ctx.slot[ci & 1];
        const int64_t n = n_units - off < chunk ? n_units - off : chunk;
        const uint8_t *d_idx = nullptr;
        const void *d_phi = nullptr, *d_psi = nullptr, *d_n = nullptr;
        if (idx) {
            B200_TRY(ensure(s, 0, n * n_data));
            B200_TRY(ensure(s, 1, n * P * rsz));
            B200_TRY(ensure(s, 2, n * P * rsz));
            B200_TRY(ensure(s, 3, n * nrow * csz));
            B200_CU(cudaMemcpyAsync(s.buf[0], idx + off * n_data, n * n_data, cudaMemcpyHostToDevice, s.st), "H2D idx");
            B200_CU(cudaMemcpyAsync(s.buf[1], (const char *)phi + off * P * rsz, n * P * rsz, cudaMemcpyHostToDevice, s.st), "H2D phi");
            B200_CU(cudaMemcpyAsync(s.buf[2], (const char *)psi + off * P * rsz, n * P * rsz, cudaMemcpyHostToDevice, s.st), "H2D psi");
            B200_CU(cudaMemcpyAsync(s.buf[3], (const char *)noise + off * nrow * csz, n * nrow * csz, cudaMemcpyHostToDevice, s.st), "H2D noise");
            d_idx = (const uint8_t *)s.buf[0]; d_phi = s.buf[1]; d_psi = s.buf[2]; d_n = s.buf[3];
        }
        uint8_t *d_hat = nullptr;
        if (idx_hat) { B200_TRY(ensure(s, 4, n * n_data)); d_hat = (uint8_t *)s.buf[4]; }
        B200_TRY(b200phy_link_ofdm_tdl(p, &modem, first_unit + off, n, d_idx, d_phi, d_psi, d_n, d_hat, nullptr, (int64_t *)s.counters, s.st));
        if (idx_hat) B200_CU(cudaMemcpyAsync(idx_hat + off * n_data, d_hat, n * n_data, cudaMemcpyDeviceToHost, s.st), "D2H idx_hat");
    }
    return finish(g_ctx, counters);
}

}  // extern "C"

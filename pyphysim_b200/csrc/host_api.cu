// host_api.cu — host-buffer entry points of the C ABI (what a non-CUDA caller binds; bench `e2e`).
//
// Monte Carlo mode (no draw arrays, no idx_hat): the only things that cross PCIe are the parameters going in
// and the 32 bytes of counters coming back — one kernel launch over all units on one internal stream.
// Stream mode (draw arrays and / or idx_hat given): units are processed in chunks that ping-pong between two
// internal streams, each with its own device staging buffers, so the H2D copy of chunk c+1 overlaps the kernel
// of chunk c and the D2H of chunk c-1.  Pinned host memory makes the copies truly asynchronous; pageable
// memory still works.  Arguments are validated BEFORE anything is copied, and every error path drains the
// internal streams before it returns (queued copies still touch the caller's buffers).
#include <mutex>
#include <vector>

#include "common.cuh"

namespace b200phy {

constexpr int kMaxIn = 4;

struct Slot {
    cudaStream_t st = nullptr;
    void *buf[kMaxIn + 1] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t cap[kMaxIn + 1] = {0, 0, 0, 0, 0};
    long long *counters = nullptr;
};

struct HostCtx {
    std::mutex mu;
    int dev = -1;
    Slot slot[2];
    void *table = nullptr;
    long long *pinned = nullptr;      // 2 x 4 counters, page-locked: the D2H of the result is asynchronous
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
            for (int i = 0; i <= kMaxIn; ++i) { if (s.buf[i]) cudaFree(s.buf[i]); s.buf[i] = nullptr; s.cap[i] = 0; }
            if (s.counters) { cudaFree(s.counters); s.counters = nullptr; }
        }
        if (c.table) { cudaFree(c.table); c.table = nullptr; }
        c.dev = -1;
        if ((e = check_cuda(cudaSetDevice(dev), "cudaSetDevice"))) return e;
    }
    for (auto &s : c.slot) {
        if ((e = check_cuda(cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking), "cudaStreamCreate"))) return e;
        if ((e = check_cuda(cudaMalloc((void **)&s.counters, 4 * sizeof(long long)), "cudaMalloc(counters)"))) return e;
        for (int i = 0; i <= kMaxIn; ++i) { s.buf[i] = nullptr; s.cap[i] = 0; }
    }
    if ((e = check_cuda(cudaMalloc(&c.table, 256 * 16), "cudaMalloc(table)"))) return e;
    if (!c.pinned && (e = check_cuda(cudaMallocHost((void **)&c.pinned, 8 * sizeof(long long)), "cudaMallocHost(counters)"))) return e;
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

static int check_table(int kind, int M, const double *table_re_im) {
    if (kind < B200PHY_MODEM_TABLE || kind > B200PHY_MODEM_QPSK) { set_error("bad modem kind %d", kind); return B200PHY_ERR_INVALID; }
    if (kind == B200PHY_MODEM_BPSK) return B200PHY_OK;
    if (!table_re_im) { set_error("table_re_im is NULL"); return B200PHY_ERR_INVALID; }
    if (M < 2 || M > 256 || (M & (M - 1))) { set_error("M=%d must be a power of two in [2, 256]", M); return B200PHY_ERR_INVALID; }
    return B200PHY_OK;
}

static int upload_table(HostCtx &c, int dtype, int kind, int M, const double *table_re_im, b200phy_modem *out) {
    out->kind = kind; out->M = M; out->table = c.table;
    if (kind == B200PHY_MODEM_BPSK) return B200PHY_OK;
    if (dtype == B200PHY_F32) {
        std::vector<float> t(2 * M);
        for (int i = 0; i < 2 * M; ++i) t[i] = float(table_re_im[i]);
        return check_cuda(cudaMemcpy(c.table, t.data(), sizeof(float) * 2 * M, cudaMemcpyHostToDevice), "cudaMemcpy(table)");
    }
    return check_cuda(cudaMemcpy(c.table, table_re_im, sizeof(double) * 2 * M, cudaMemcpyHostToDevice), "cudaMemcpy(table)");
}

static int drain(HostCtx &c, int e) {
    for (auto &s : c.slot)
        if (s.st) cudaStreamSynchronize(s.st);
    return e;
}

// One input array of a link in stream mode: host base pointer and bytes per unit.
struct HostIn { const void *host; size_t bytes_per_unit; };

// Runs `launch(off, n, d_in[], d_hat, d_counters, stream)` over all units.
//   n_in == 0 and no idx_hat: one launch (Monte Carlo mode).  Otherwise chunks of ~64 MiB of staged bytes:
//   long enough to run PCIe at full rate and to fill the GPU, short enough that the un-overlapped head
//   (first H2D) and tail (last kernel + D2H) stay small.
template <typename Launch>
static int run_host(HostCtx &c, int64_t n_units, const HostIn *in, int n_in, uint8_t *idx_hat, size_t hat_per_unit,
                    int64_t *counters, Launch launch) {
    int e;
    const void *d_in[kMaxIn] = {nullptr, nullptr, nullptr, nullptr};
    if (n_in == 0 && !idx_hat) {
        Slot &s = c.slot[0];
        if ((e = check_cuda(cudaMemsetAsync(s.counters, 0, 4 * sizeof(long long), s.st), "memset"))) return e;
        if (n_units > 0 && (e = launch(0, n_units, d_in, nullptr, (int64_t *)s.counters, s.st))) return drain(c, e);
        if ((e = check_cuda(cudaMemcpyAsync(c.pinned, s.counters, 4 * sizeof(long long), cudaMemcpyDeviceToHost, s.st), "D2H counters"))) return drain(c, e);
        if ((e = check_cuda(cudaStreamSynchronize(s.st), "cudaStreamSynchronize"))) return e;
        for (int i = 0; i < 4; ++i) counters[i] += c.pinned[i];
        return B200PHY_OK;
    }
    size_t per_unit = idx_hat ? hat_per_unit : 0;
    for (int i = 0; i < n_in; ++i) per_unit += in[i].bytes_per_unit;
    int64_t chunk = int64_t((size_t(64) << 20) / (per_unit ? per_unit : 1));
    if (chunk < 1) chunk = 1;
    for (auto &s : c.slot)
        if ((e = check_cuda(cudaMemsetAsync(s.counters, 0, 4 * sizeof(long long), s.st), "memset"))) return drain(c, e);
    int ci = 0;
    for (int64_t off = 0; off < n_units; off += chunk, ++ci) {
        Slot &s = c.slot[ci & 1];
        const int64_t n = n_units - off < chunk ? n_units - off : chunk;
        for (int i = 0; i < n_in; ++i) {
            const size_t b = in[i].bytes_per_unit;
            if ((e = ensure(s, i, size_t(n) * b))) return drain(c, e);
            if ((e = check_cuda(cudaMemcpyAsync(s.buf[i], (const char *)in[i].host + size_t(off) * b, size_t(n) * b,
                                                cudaMemcpyHostToDevice, s.st), "H2D draws")))
                return drain(c, e);
            d_in[i] = s.buf[i];
        }
        uint8_t *d_hat = nullptr;
        if (idx_hat) {
            if ((e = ensure(s, kMaxIn, size_t(n) * hat_per_unit))) return drain(c, e);
            d_hat = (uint8_t *)s.buf[kMaxIn];
        }
        if ((e = launch(off, n, d_in, d_hat, (int64_t *)s.counters, s.st))) return drain(c, e);
        if (idx_hat &&
            (e = check_cuda(cudaMemcpyAsync(idx_hat + size_t(off) * hat_per_unit, d_hat, size_t(n) * hat_per_unit,
                                            cudaMemcpyDeviceToHost, s.st), "D2H idx_hat")))
            return drain(c, e);
    }
    for (int s = 0; s < 2; ++s)
        if ((e = check_cuda(cudaMemcpyAsync(c.pinned + 4 * s, c.slot[s].counters, 4 * sizeof(long long), cudaMemcpyDeviceToHost,
                                            c.slot[s].st), "D2H counters")))
            return drain(c, e);
    for (int s = 0; s < 2; ++s)
        if ((e = check_cuda(cudaStreamSynchronize(c.slot[s].st), "cudaStreamSynchronize"))) return drain(c, e);
    for (int i = 0; i < 4; ++i) counters[i] += c.pinned[i] + c.pinned[4 + i];
    return B200PHY_OK;
}

static int check_head(int dtype, int64_t n_units, double noise_var, const int64_t *counters) {
    if (!counters) { set_error("counters is NULL"); return B200PHY_ERR_INVALID; }
    if (dtype != B200PHY_F32 && dtype != B200PHY_F64) { set_error("dtype must be B200PHY_F32 or B200PHY_F64"); return B200PHY_ERR_INVALID; }
    if (n_units < 0) { set_error("n_units must be non-negative"); return B200PHY_ERR_INVALID; }
    if (!(noise_var >= 0.0)) { set_error("Noise variance must be a non-negative value."); return B200PHY_ERR_INVALID; }
    return B200PHY_OK;
}

// all-or-nothing draw arrays (need_b: the middle array is part of the set)
static int check_draws(const void *a, const void *b, const void *c, bool need_b, const char *names) {
    const bool any = a || c || (need_b && b), all = a && c && (!need_b || b);
    if (any && !all) { set_error("stream mode needs %s together; Monte Carlo mode needs all of them NULL", names); return B200PHY_ERR_INVALID; }
    return B200PHY_OK;
}

}  // namespace b200phy

using namespace b200phy;

#define B200_TRY(x) do { int _e = (x); if (_e) return _e; } while (0)

extern "C" {

int b200phy_link_siso_flat_host(int dtype, int modem_kind, int M, const double *table_re_im,
                                int rayleigh, double noise_var, uint64_t seed, uint64_t first_unit,
                                int64_t n_units, const uint8_t *idx, const void *h, const void *noise,
                                uint8_t *idx_hat, int64_t *counters) {
    B200_TRY(check_head(dtype, n_units, noise_var, counters));
    B200_TRY(check_table(modem_kind, M, table_re_im));
    B200_TRY(check_draws(idx, h, noise, rayleigh != 0, rayleigh ? "idx, h, noise" : "idx, noise"));
    std::lock_guard<std::mutex> lk(g_ctx.mu);
    B200_TRY(ctx_init(g_ctx));
    b200phy_modem modem;
    B200_TRY(upload_table(g_ctx, dtype, modem_kind, M, table_re_im, &modem));
    const size_t csz = dtype == B200PHY_F32 ? 8 : 16;
    HostIn in[3];
    int n_in = 0;
    if (idx) {
        in[n_in++] = {idx, 1};
        in[n_in++] = {noise, csz};
        if (rayleigh) in[n_in++] = {h, csz};
    }
    return run_host(g_ctx, n_units, in, n_in, idx_hat, 1, counters,
                    [&](int64_t off, int64_t n, const void *const *d, uint8_t *d_hat, int64_t *d_cnt, cudaStream_t st) {
                        return b200phy_link_siso_flat(dtype, &modem, rayleigh, noise_var, seed, first_unit + off, n,
                                                      (const uint8_t *)d[0], rayleigh ? d[2] : nullptr, d[1], d_hat, nullptr, d_cnt, st);
                    });
}

int b200phy_link_alamouti_host(int dtype, int modem_kind, int M, const double *table_re_im, int Nr, int S,
                               double noise_var, uint64_t seed, uint64_t first_unit, int64_t n_units,
                               const uint8_t *idx, const void *H, const void *noise, uint8_t *idx_hat,
                               int64_t *counters) {
    B200_TRY(check_head(dtype, n_units, noise_var, counters));
    B200_TRY(check_table(modem_kind, M, table_re_im));
    if (Nr < 1 || Nr > B200PHY_MAX_ANT) { set_error("Alamouti: Nr=%d must be in [1, %d]", Nr, B200PHY_MAX_ANT); return B200PHY_ERR_UNSUPPORTED; }
    if (S < 2 || (S & 1)) { set_error("Alamouti: number of symbols S=%d must be even", S); return B200PHY_ERR_INVALID; }
    B200_TRY(check_draws(idx, H, noise, true, "idx, H, noise"));
    std::lock_guard<std::mutex> lk(g_ctx.mu);
    B200_TRY(ctx_init(g_ctx));
    b200phy_modem modem;
    B200_TRY(upload_table(g_ctx, dtype, modem_kind, M, table_re_im, &modem));
    const size_t csz = dtype == B200PHY_F32 ? 8 : 16;
    HostIn in[3] = {{idx, size_t(S)}, {H, size_t(Nr) * 2 * csz}, {noise, size_t(Nr) * S * csz}};
    return run_host(g_ctx, n_units, in, idx ? 3 : 0, idx_hat, size_t(S), counters,
                    [&](int64_t off, int64_t n, const void *const *d, uint8_t *d_hat, int64_t *d_cnt, cudaStream_t st) {
                        return b200phy_link_alamouti(dtype, &modem, Nr, S, noise_var, seed, first_unit + off, n,
                                                     (const uint8_t *)d[0], d[1], d[2], d_hat, nullptr, d_cnt, st);
                    });
}

int b200phy_link_blast_host(int dtype, int modem_kind, int M, const double *table_re_im, int Nr, int Nt, int S,
                            double noise_var, double filter_noise_var, uint64_t seed, uint64_t first_unit,
                            int64_t n_units, const uint8_t *idx, const void *H, const void *noise,
                            uint8_t *idx_hat, int64_t *counters) {
    B200_TRY(check_head(dtype, n_units, noise_var, counters));
    B200_TRY(check_table(modem_kind, M, table_re_im));
    if (Nr < 1 || Nr > B200PHY_MAX_ANT || Nt < 1 || Nt > B200PHY_MAX_ANT) {
        set_error("Blast: Nr=%d, Nt=%d must be in [1, %d]", Nr, Nt, B200PHY_MAX_ANT);
        return B200PHY_ERR_UNSUPPORTED;
    }
    if (!(filter_noise_var >= 0.0)) { set_error("Noise variance must be a non-negative value."); return B200PHY_ERR_INVALID; }
    if (filter_noise_var == 0.0 && Nt > Nr) { set_error("Blast ZF needs Nt <= Nr (got %dx%d)", Nr, Nt); return B200PHY_ERR_UNSUPPORTED; }
    if (S < 1) { set_error("S must be positive"); return B200PHY_ERR_INVALID; }
    B200_TRY(check_draws(idx, H, noise, true, "idx, H, noise"));
    std::lock_guard<std::mutex> lk(g_ctx.mu);
    B200_TRY(ctx_init(g_ctx));
    b200phy_modem modem;
    B200_TRY(upload_table(g_ctx, dtype, modem_kind, M, table_re_im, &modem));
    const size_t csz = dtype == B200PHY_F32 ? 8 : 16;
    HostIn in[3] = {{idx, size_t(S) * Nt}, {H, size_t(Nr) * Nt * csz}, {noise, size_t(Nr) * S * csz}};
    return run_host(g_ctx, n_units, in, idx ? 3 : 0, idx_hat, size_t(S) * Nt, counters,
                    [&](int64_t off, int64_t n, const void *const *d, uint8_t *d_hat, int64_t *d_cnt, cudaStream_t st) {
                        return b200phy_link_blast(dtype, &modem, Nr, Nt, S, noise_var, filter_noise_var, seed,
                                                  first_unit + off, n, (const uint8_t *)d[0], d[1], d[2], d_hat, nullptr, d_cnt, st);
                    });
}

int b200phy_link_precoded_host(int dtype, int modem_kind, int M, const double *table_re_im, int scheme, int Nr,
                               int Nt, int S, double noise_var, double filter_noise_var, uint64_t seed,
                               uint64_t first_unit, int64_t n_units, const uint8_t *idx, const void *H,
                               const void *noise, uint8_t *idx_hat, int64_t *counters) {
    B200_TRY(check_head(dtype, n_units, noise_var, counters));
    B200_TRY(check_table(modem_kind, M, table_re_im));
    if (scheme != B200PHY_MIMO_SVD && scheme != B200PHY_MIMO_GMD && scheme != B200PHY_MIMO_MRT) { set_error("bad MIMO scheme %d", scheme); return B200PHY_ERR_INVALID; }
    if (Nr < 1 || Nr > B200PHY_MAX_ANT || Nt < 1 || Nt > B200PHY_MAX_ANT || S < 1) {
        set_error("precoded link: Nr=%d, Nt=%d must be in [1, %d] and S=%d positive", Nr, Nt, B200PHY_MAX_ANT, S);
        return B200PHY_ERR_UNSUPPORTED;
    }
    if (!(filter_noise_var >= 0.0)) { set_error("Noise variance must be a non-negative value."); return B200PHY_ERR_INVALID; }
    B200_TRY(check_draws(idx, H, noise, true, "idx, H, noise"));
    std::lock_guard<std::mutex> lk(g_ctx.mu);
    B200_TRY(ctx_init(g_ctx));
    b200phy_modem modem;
    B200_TRY(upload_table(g_ctx, dtype, modem_kind, M, table_re_im, &modem));
    const size_t csz = dtype == B200PHY_F32 ? 8 : 16;
    const size_t layers = scheme == B200PHY_MIMO_MRT ? 1 : size_t(Nt);
    HostIn in[3] = {{idx, size_t(S) * layers}, {H, size_t(Nr) * Nt * csz}, {noise, size_t(Nr) * S * csz}};
    return run_host(g_ctx, n_units, in, idx ? 3 : 0, idx_hat, size_t(S) * layers, counters,
                    [&](int64_t off, int64_t n, const void *const *d, uint8_t *d_hat, int64_t *d_cnt, cudaStream_t st) {
                        return b200phy_link_precoded(dtype, &modem, scheme, Nr, Nt, S, noise_var, filter_noise_var, seed,
                                                     first_unit + off, n, (const uint8_t *)d[0], d[1], d[2], d_hat, nullptr, d_cnt, st);
                    });
}

int b200phy_link_ofdm_tdl_host(const b200phy_ofdm_tdl_params *p, int modem_kind, int M,
                               const double *table_re_im, uint64_t first_unit, int64_t n_units,
                               const uint8_t *idx, const void *phi, const void *psi,
                               const void *noise, uint8_t *idx_hat, int64_t *counters) {
    // the struct sizes every allocation and copy below: validate it (size, dtype, shape ranges) first
    B200_TRY(b200phy_ofdm_tdl_check_params(p));
    B200_TRY(check_head(p->dtype, n_units, p->noise_var, counters));
    B200_TRY(check_table(modem_kind, M, table_re_im));
    {
        const bool any = idx || phi || psi || noise, all = idx && phi && psi && noise;
        if (any && !all) { set_error("stream mode needs idx, phi, psi and noise together; Monte Carlo mode needs all NULL"); return B200PHY_ERR_INVALID; }
    }
    std::lock_guard<std::mutex> lk(g_ctx.mu);
    B200_TRY(ctx_init(g_ctx));
    b200phy_modem modem;
    B200_TRY(upload_table(g_ctx, p->dtype, modem_kind, M, table_re_im, &modem));
    const size_t rsz = p->dtype == B200PHY_F32 ? 4 : 8, csz = 2 * rsz;
    const size_t n_data = size_t(p->Nt) * p->n_sym * p->used;
    const size_t P = size_t(p->L) * p->n_taps * p->Nr * p->Nt;
    const size_t nrow = size_t(p->Nr) * (size_t(p->n_sym) * (p->fft + p->cp) + p->delays[p->n_taps - 1]);
    HostIn in[4] = {{idx, n_data}, {phi, P * rsz}, {psi, P * rsz}, {noise, nrow * csz}};
    return run_host(g_ctx, n_units, in, idx ? 4 : 0, idx_hat, n_data, counters,
                    [&](int64_t off, int64_t n, const void *const *d, uint8_t *d_hat, int64_t *d_cnt, cudaStream_t st) {
                        return b200phy_link_ofdm_tdl(p, &modem, first_unit + off, n, (const uint8_t *)d[0], d[1], d[2], d[3],
                                                     d_hat, nullptr, nullptr, d_cnt, st);
                    });
}

}  // extern "C"

// lib.cu — library-level plumbing of libb200phy: error reporting, launch accounting.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace b200phy {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_cuda(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return B200PHY_OK;
    set_error("%s: %s", what, cudaGetErrorString(e));
    return B200PHY_ERR_CUDA;
}

void count_launch(int n) { g_launches.fetch_add(uint64_t(n), std::memory_order_relaxed); }

static char g_last_kernel[160] = "";
void note_kernel(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_kernel, sizeof(g_last_kernel), fmt, ap);
    va_end(ap);
}

int check_modem(const b200phy_modem *m, Modem *out) {
    if (!m) { set_error("modem is NULL"); return B200PHY_ERR_INVALID; }
    const int M = m->M;
    if (M < 2 || M > 256 || (M & (M - 1))) {
        set_error("constellation size M=%d must be a power of two in [2, 256]", M);
        return B200PHY_ERR_INVALID;
    }
    if (m->kind == B200PHY_MODEM_QAM) {
        const int b = ilog2(M);
        if (b & 1) { set_error("M must be a square power of 2"); return B200PHY_ERR_INVALID; }
    } else if (m->kind == B200PHY_MODEM_BPSK) {
        if (M != 2) { set_error("BPSK requires M=2"); return B200PHY_ERR_INVALID; }
    } else if (m->kind == B200PHY_MODEM_QPSK) {
        if (M != 4) { set_error("QPSK requires M=4"); return B200PHY_ERR_INVALID; }
    } else if (m->kind != B200PHY_MODEM_TABLE) {
        set_error("unknown modem kind %d", m->kind);
        return B200PHY_ERR_INVALID;
    }
    if (m->kind != B200PHY_MODEM_BPSK && !m->table) {
        set_error("modem table pointer is NULL");
        return B200PHY_ERR_INVALID;
    }
    *out = make_modem(m->kind, M);
    return B200PHY_OK;
}

}  // namespace b200phy

extern "C" {
int b200phy_version(void) { return B200PHY_VERSION; }
const char *b200phy_last_error(void) { return b200phy::g_err; }
uint64_t b200phy_launch_count(void) { return b200phy::g_launches.load(); }
const char *b200phy_last_kernel(void) { return b200phy::g_last_kernel; }
}

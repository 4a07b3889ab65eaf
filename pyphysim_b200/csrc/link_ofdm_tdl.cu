// link_ofdm_tdl.cu — host side of the fused OFDM/TDL link: parameter checks, Jakes-mode choice,
// dispatch to the per-antenna-configuration kernel instantiations, and the draw dump kernel.
#include <cmath>

#include "ofdm_tdl.cuh"

namespace b200phy {

// Error model of the POLY mode.  Around the expansion point the kernel evaluates the order-P Taylor
// polynomial of exp(j w tau) for |tau| <= tau_max samples; the remainder of each unit ray is
// (w tau)^(P+1)/(P+1)! and a tap sums L rays of amplitude sqrt(P_l/L), so
// |error| <= sqrt(L P_l) x^(P+1)/(P+1)! <= sqrt(L) x^(P+1)/(P+1)!  with x = 2 pi Fd Ts tau_max.
static double poly_error_bound(const b200phy_ofdm_tdl_params *q, int order, double tau_max) {
    const double x = 2.0 * M_PI * fabs(q->Fd) * q->Ts * 1.0000000001 * tau_max;
    double r = sqrt(double(q->L));
    for (int i = 1; i <= order + 1; ++i) r *= x / i;
    return r;
}

static int fill_params(const b200phy_ofdm_tdl_params *q, const Modem &m, OfdmP *out) {
    if (!q) { set_error("params is NULL"); return B200PHY_ERR_INVALID; }
    if (q->struct_size != (int)sizeof(b200phy_ofdm_tdl_params)) {
        set_error("b200phy_ofdm_tdl_params size mismatch (%d vs %zu): header/library out of sync",
                  q->struct_size, sizeof(b200phy_ofdm_tdl_params));
        return B200PHY_ERR_INVALID;
    }
    if (q->dtype != B200PHY_F32 && q->dtype != B200PHY_F64) { set_error("bad dtype"); return B200PHY_ERR_INVALID; }
    // OFDM.set_parameters (modulators/ofdm.py:75-90)
    if (q->cp < 0 || q->cp > q->fft) { set_error("cp_size must be nonnegative and cannot be greater than fft_size"); return B200PHY_ERR_INVALID; }
    if (q->used > q->fft) { set_error("Number of used subcarriers cannot be greater than the fft_size"); return B200PHY_ERR_INVALID; }
    if ((q->used % 2) != 0 || q->used < 2) { set_error("Number of used subcarriers must be a multiple of 2"); return B200PHY_ERR_INVALID; }
    if (q->fft < 8 || q->fft > 4096 || (q->fft & (q->fft - 1))) { set_error("fft_size=%d must be a power of two in [8, 4096]", q->fft); return B200PHY_ERR_UNSUPPORTED; }
    if (q->n_sym < 1) { set_error("n_sym must be positive"); return B200PHY_ERR_INVALID; }
    if (q->n_taps < 1 || q->n_taps > B200PHY_MAX_TAPS) { set_error("n_taps=%d must be in [1, %d]", q->n_taps, B200PHY_MAX_TAPS); return B200PHY_ERR_UNSUPPORTED; }
    if (q->L < 1 || q->L > B200PHY_MAX_RAYS) { set_error("L=%d must be in [1, %d]", q->L, B200PHY_MAX_RAYS); return B200PHY_ERR_UNSUPPORTED; }
    if (!(q->noise_var >= 0.0) || !(q->filter_noise_var >= 0.0)) { set_error("Noise variance must be a non-negative value."); return B200PHY_ERR_INVALID; }
    if (!(q->Ts > 0.0)) { set_error("Ts must be positive"); return B200PHY_ERR_INVALID; }
    if (q->filter_noise_var == 0.0 && q->Nt > q->Nr) { set_error("ZF needs Nt <= Nr (got %dx%d)", q->Nr, q->Nt); return B200PHY_ERR_UNSUPPORTED; }
    for (int l = 0; l < q->n_taps; ++l) {
        if (q->delays[l] < 0 || (l > 0 && q->delays[l] <= q->delays[l - 1])) { set_error("tap delays must be non-negative and strictly increasing"); return B200PHY_ERR_INVALID; }
        if (!(q->tap_powers[l] >= 0.0)) { set_error("tap powers must be non-negative"); return B200PHY_ERR_INVALID; }
    }
    OfdmP &p = *out;
    p.fft = q->fft; p.lg = ilog2(q->fft); p.cp = q->cp; p.used = q->used; p.half = q->used / 2;
    p.n_sym = q->n_sym; p.S = q->fft + q->cp; p.N = p.n_sym * p.S;
    p.mem = q->delays[q->n_taps - 1];
    if (p.mem > q->fft) { set_error("channel memory %d exceeds fft_size %d", p.mem, q->fft); return B200PHY_ERR_UNSUPPORTED; }
    p.n_taps = q->n_taps; p.L = q->L;
    p.n_data = q->Nt * q->n_sym * q->used;
    p.row = 2 * ((p.N + p.mem + 1) / 2);
    p.P = q->L * q->n_taps * q->Nr * q->Nt;
    p.P4 = (p.P + 3) / 4 * 4;
    p.ifft_in_w = (((p.lg >> 1) + (p.lg & 1)) & 1);
    for (int l = 0; l < q->n_taps; ++l) {
        p.delays[l] = q->delays[l];
        p.amp[l] = sqrt(q->tap_powers[l] / double(q->L));
    }
    {
        int j = 0;
        const int order[4] = {0, 2, 1, 3};           // runs of d mod 4; the classes mod 2 stay contiguous
        for (int i = 0; i < 4; ++i) {
            p.cls_start[i] = j;
            for (int l = 0; l < q->n_taps; ++l)
                if ((q->delays[l] & 3) == order[i]) { p.cls_pos[l] = j; p.cls_delay[j] = q->delays[l]; ++j; }
        }
        p.cls_start[4] = j;
    }
    p.w0 = 2.0 * M_PI * q->Fd;
    p.Ts1 = q->Ts * 1.0000000001;
    p.t0 = q->t0;
    p.sigma = sqrt(q->noise_var);
    p.fnv = q->filter_noise_var;
    const double power_scale = double(q->fft) * double(q->fft) / (double(q->used) + double(q->cp));   // ofdm.py:370-392
    p.tx_scale = sqrt(power_scale) / double(q->fft) / sqrt(double(q->Nt));   // ifft 1/N and Blast 1/sqrt(Nt)
    p.rx_scale = 1.0 / sqrt(power_scale);
    p.snt = sqrt(double(q->Nt));
    p.seed = q->seed;
    p.rx_out = nullptr;

    // Jakes evaluation mode
    const double tol = q->dtype == B200PHY_F32 ? 2e-8 : 2e-14;
    int seg = q->fft;
    const int min_seg = q->fft < kOT ? q->fft : kOT;
    while (seg > min_seg && poly_error_bound(q, 3, 0.5 * seg) > tol) seg >>= 1;
    const bool poly_ok = poly_error_bound(q, 3, 0.5 * seg) <= tol;
    if (q->jakes_mode == B200PHY_JAKES_POLY && !poly_ok) {
        set_error("JAKES_POLY requested but its error bound %.3g exceeds %.3g (Fd*Ts too large)",
                  poly_error_bound(q, 3, 0.5 * seg), tol);
        return B200PHY_ERR_UNSUPPORTED;
    }
    p.poly = (q->jakes_mode == B200PHY_JAKES_POLY) || (q->jakes_mode == B200PHY_JAKES_AUTO && poly_ok);
    p.seg_len = seg; p.seg_lg = ilog2(seg); p.nseg = q->fft / seg;
    p.porder = poly_error_bound(q, 2, 0.5 * seg) <= tol ? 2 : 3;
    p.gbar_poly = 0;
    if (p.poly && p.nseg == 1) {
        // the symbol-mean taps (CP included) from the same polynomial, extrapolated over CP / delay
        const double tg = 0.5 * q->fft + (q->cp > p.mem ? q->cp : p.mem);
        if (poly_error_bound(q, p.porder, tg) <= tol) p.gbar_poly = 1;
        else if (poly_error_bound(q, 3, tg) <= tol) { p.porder = 3; p.gbar_poly = 1; }
    }
    for (int l = 0; l < q->n_taps; ++l) {
        double m1 = 0, m2 = 0, m3 = 0;
        const double c = q->cp + 0.5 * (q->fft - 1) - q->delays[l];
        for (int n = 0; n < p.S; ++n) { const double t = n - c; m1 += t; m2 += t * t; m3 += t * t * t; }
        p.mu[0][l] = m1 / p.S; p.mu[1][l] = m2 / p.S; p.mu[2][l] = m3 / p.S;
    }
    // float cos(phi) perturbs every ray frequency by ~1e-7 relative: harmless while the total phase
    // advance w0 * t_end stays small
    p.no_pair = (q->reserved & 1);
    p.tc_hk = (q->reserved & 2) ? 1 : 0;
    p.cos_f32 = (q->dtype == B200PHY_F32) && (p.w0 * (fabs(q->t0) + p.N * p.Ts1) < 0.05);
    {
        const int items = q->n_taps * q->Nr;
        int g = 1;
        while (g < 16 && (g * 2) * items <= kOT) g *= 2;
        p.cgrp = g;
    }
    if (q->jakes_mode != B200PHY_JAKES_AUTO && q->jakes_mode != B200PHY_JAKES_POLY &&
        q->jakes_mode != B200PHY_JAKES_RECURRENCE) { set_error("bad jakes_mode"); return B200PHY_ERR_INVALID; }
    (void)m;
    return B200PHY_OK;
}

// ---------------------------------------------------------------- draw dump
template <typename T>
__global__ void draw_ofdm_tdl_kernel(const __grid_constant__ OfdmP p, int bits, int NR, uint64_t first_unit, long long n_units,
                                     uint8_t *idx, T *phi, T *psi, cx<T> *noise) {
    const long long rowlen = p.N + p.mem;
    for (long long frame = blockIdx.x; frame < n_units; frame += gridDim.x) {
        const uint64_t unit = first_unit + uint64_t(frame);
        if (idx)
            for (int w = threadIdx.x; w < p.n_data; w += blockDim.x)
                idx[frame * p.n_data + w] =
                    uint8_t(lane_of(rng_block(p.seed, STREAM_DATA, unit, uint64_t(w >> 2)), w & 3) >> (32 - bits));
        if (phi)
            for (int i = threadIdx.x; i < p.P; i += blockDim.x)
                phi[frame * p.P + i] = phase_from_word<T>(lane_of(rng_block(p.seed, STREAM_CHANNEL, unit, uint64_t(i >> 2)), i & 3));
        if (psi)
            for (int i = threadIdx.x; i < p.P; i += blockDim.x) {
                const int i2 = p.P4 + i;
                psi[frame * p.P + i] = phase_from_word<T>(lane_of(rng_block(p.seed, STREAM_CHANNEL, unit, uint64_t(i2 >> 2)), i2 & 3));
            }
        if (noise)
            for (long long it = threadIdx.x; it < NR * rowlen; it += blockDim.x) {
                const int r = int(it / rowlen);
                const long long mm = it % rowlen;
                noise[(frame * NR + r) * rowlen + mm] = cnormal_at<T>(p.seed, STREAM_NOISE, unit, uint64_t(r) * p.row + mm);
            }
    }
}

}  // namespace b200phy

using namespace b200phy;

extern "C" {

int b200phy_link_ofdm_tdl(const b200phy_ofdm_tdl_params *q, const b200phy_modem *modem,
                          uint64_t first_unit, int64_t n_units, const uint8_t *idx, const void *phi,
                          const void *psi, const void *noise, uint8_t *idx_hat, void *eq_out,
                          void *rx_out, int64_t *counters, void *stream) {
    Modem m;
    int e = check_modem(modem, &m);
    if (e) return e;
    OfdmP p;
    if ((e = fill_params(q, m, &p))) return e;
    p.rx_out = rx_out;
    if (!counters) { set_error("counters is NULL"); return B200PHY_ERR_INVALID; }
    if (n_units < 0) { set_error("n_units must be non-negative"); return B200PHY_ERR_INVALID; }
    const bool any = idx || phi || psi || noise, all = idx && phi && psi && noise;
    if (any && !all) { set_error("stream mode needs idx, phi, psi and noise together; fused mode needs all NULL"); return B200PHY_ERR_INVALID; }
    {
        const size_t rs = q->dtype == B200PHY_F32 ? 4 : 8;
        if ((e = require_aligned(idx, 4, "idx")) || (e = require_aligned(phi, rs, "phi")) || (e = require_aligned(psi, rs, "psi")) ||
            (e = require_aligned(noise, 2 * rs, "noise")) || (e = require_aligned(eq_out, 2 * rs, "eq_out")) ||
            (e = require_aligned(rx_out, 2 * rs, "rx_out")))
            return e;
    }
    if (n_units == 0) return B200PHY_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int code = q->Nr * 10 + q->Nt;
#define B200_DISPATCH(R, T_) case R * 10 + T_: return launch_ofdm_tdl<R, T_>(q->dtype, p, m, modem->table, first_unit, n_units, idx, phi, psi, noise, idx_hat, eq_out, counters, st)
    switch (code) {
        B200_DISPATCH(1, 1);
        B200_DISPATCH(2, 1);
        B200_DISPATCH(2, 2);
        B200_DISPATCH(3, 1);
        B200_DISPATCH(3, 2);
        B200_DISPATCH(3, 3);
        B200_DISPATCH(4, 1);
        B200_DISPATCH(4, 2);
        B200_DISPATCH(4, 3);
        B200_DISPATCH(4, 4);
        default:
            set_error("OFDM/TDL link is built for Nt <= Nr <= 4; got %dx%d", q->Nr, q->Nt);
            return B200PHY_ERR_UNSUPPORTED;
    }
#undef B200_DISPATCH
}

int b200phy_ofdm_tdl_check_params(const b200phy_ofdm_tdl_params *q) {
    OfdmP p;
    Modem m = make_modem(B200PHY_MODEM_TABLE, 2);
    int e = fill_params(q, m, &p);
    if (e) return e;
    switch (q->Nr * 10 + q->Nt) {
        case 11: case 21: case 22: case 31: case 32: case 33: case 41: case 42: case 43: case 44: return B200PHY_OK;
        default:
            set_error("OFDM/TDL link is built for Nt <= Nr <= 4; got %dx%d", q->Nr, q->Nt);
            return B200PHY_ERR_UNSUPPORTED;
    }
}

int b200phy_draw_ofdm_tdl(const b200phy_ofdm_tdl_params *q, int bits, uint64_t first_unit,
                          int64_t n_units, uint8_t *idx, void *phi, void *psi, void *noise,
                          void *stream) {
    OfdmP p;
    Modem m = make_modem(B200PHY_MODEM_TABLE, 1 << bits);
    int e = fill_params(q, m, &p);
    if (e) return e;
    if (n_units <= 0) return B200PHY_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = int(n_units < 1184 ? n_units : 1184);
    if (q->dtype == B200PHY_F32)
        draw_ofdm_tdl_kernel<float><<<grid, 256, 0, st>>>(p, bits, q->Nr, first_unit, n_units, idx, (float *)phi, (float *)psi, (cx<float> *)noise);
    else
        draw_ofdm_tdl_kernel<double><<<grid, 256, 0, st>>>(p, bits, q->Nr, first_unit, n_units, idx, (double *)phi, (double *)psi, (cx<double> *)noise);
    B200_CHECK_LAUNCH("draw_ofdm_tdl_kernel");
    return B200PHY_OK;
}

}  // extern "C"

// stages_refsig.cu — SURVEY.md §8f row next-4: reference-signal sequences (Zadoff-Chu / SRS / DMRS,
// reference_signals/{zadoffchu,root_sequence,srs,dmrs}.py) and the pilot-based channel estimators
// (reference_signals/channel_estimation.py, channel_estimation/estimators.py), batched over independent
// received vectors / realizations.  Phases are reduced in integer arithmetic before sincospi, so the
// sequences are accurate to the last ulp for any length; small solves run in double like linalg.cuh.
#include <cmath>
#include <cstdlib>
#include <utility>

#include "common.cuh"

namespace b200phy {

struct PhiTable { signed char v[24]; };

// r[n] = scale * base[n mod Nzc] * exp(j 2 pi n_cs n / denom)
//   base = exp(-j pi u m (m + 1 + 2q) / Nzc)   (calcBaseZC, zadoffchu.py:11-36), or
//   base = exp(j pi phi[m] / 4)                (RootSequence for 12 / 24 elements, root_sequence.py:273-283)
// cyclic extension = the n mod Nzc (get_extended_ZF, zadoffchu.py:75-113); shift = get_shifted_root_seq (:39-72)
template <typename T>
__global__ void __launch_bounds__(256)
refsig_sequence_kernel(int Nzc, int u, double q_re, double q_im, int use_table, PhiTable tab, int size, int n_cs,
                       int denom, double s_re, double s_im, cx<T> *__restrict__ out) {
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < size; n += gridDim.x * blockDim.x) {
        const int m = n % Nzc;
        double half_turns, amp = 1.0;                   // angle / pi
        if (use_table) {
            half_turns = 0.25 * double(tab.v[m]);
        } else {
            const long long k = ((long long)u * m % (2LL * Nzc)) * (m + 1) % (2LL * Nzc);   // u m (m+1) mod 2 Nzc
            half_turns = -double(k) / double(Nzc);
            if (q_re != 0.0 || q_im != 0.0) {
                const double f = 2.0 * double(u) * double(m) / double(Nzc);
                half_turns -= f * q_re;                 // exp(-j pi f (q_re + j q_im))
                amp = exp(3.14159265358979323846 * f * q_im);
            }
        }
        const long long sh = ((long long)n_cs * n) % denom;
        half_turns += 2.0 * double(sh) / double(denom);
        double sn, cs;
        sincospi(half_turns, &sn, &cs);
        cs *= amp; sn *= amp;
        out[n] = {T(cs * s_re - sn * s_im), T(cs * s_im + sn * s_re)};
    }
}

// CazacBasedChannelEstimator.estimate_channel_freq_domain (reference_signals/channel_estimation.py:69-131) with the
// cover-code average of CazacBasedWithOCCChannelEstimator (:163-251) folded in.  One CTA per received vector:
//   z[k] = conj(r[k]) * mean_c(cc[c] * y[c][k]);  h[l] = (1/Nsc) sum_k z[k] e^{+2 pi j k l / Nsc}, l < n_keep;
//   out[m] = scale * sum_l h[l] e^{-2 pi j m l / Nout},  Nout = mult * Nsc.
// Both transforms are evaluated directly (n_keep is a handful of taps), twiddles from one table of Nout roots.
struct Cover { double re[4], im[4]; };

template <typename T>
__global__ void __launch_bounds__(256)
cazac_estimate_kernel(const cx<T> *__restrict__ ref, const cx<T> *__restrict__ y, Cover cc, int n_cover, int Nsc,
                      int n_keep, int mult, double scale, cx<T> *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int Nout = mult * Nsc;
    cx<T> *tw = reinterpret_cast<cx<T> *>(smem_raw);           // e^{-2 pi j i / Nout}
    cx<T> *z = tw + Nout;
    cx<T> *h = z + Nsc;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    for (int i = tid; i < Nout; i += blockDim.x) {
        double sn, cs;
        sincospi(-2.0 * double(i) / double(Nout), &sn, &cs);
        tw[i] = {T(cs), T(sn)};
    }
    const cx<T> *yb = y + size_t(blockIdx.x) * n_cover * Nsc;
    const T inv_c = T(1.0 / double(n_cover));
    for (int k = tid; k < Nsc; k += blockDim.x) {
        cx<T> acc = {T(0), T(0)};
        for (int c = 0; c < n_cover; ++c) cmac(acc, mk<T>(T(cc.re[c]), T(cc.im[c])), yb[size_t(c) * Nsc + k]);
        acc = inv_c * acc;
        cx<T> v = {T(0), T(0)};
        cmac_conj(v, ref[k], acc);
        z[k] = v;
    }
    __syncthreads();
    const T inv_n = T(1.0 / double(Nsc));
    for (int l = warp; l < n_keep; l += nwarp) {
        cx<T> acc = {T(0), T(0)};
        for (int k = lane; k < Nsc; k += 32) {
            const int e = int(((long long)k * l) % Nsc) * mult;
            cmac_conj(acc, tw[e], z[k]);                         // conj(tw) = e^{+2 pi j k l / Nsc}
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            acc.re += __shfl_xor_sync(0xffffffffu, acc.re, o);
            acc.im += __shfl_xor_sync(0xffffffffu, acc.im, o);
        }
        if (lane == 0) h[l] = inv_n * acc;
    }
    __syncthreads();
    cx<T> *ob = out + size_t(blockIdx.x) * Nout;
    for (int mo = tid; mo < Nout; mo += blockDim.x) {
        cx<T> acc = {T(0), T(0)};
        int e = 0;                                               // (mo * l) mod Nout, updated incrementally
        for (int l = 0; l < n_keep; ++l) {
            cmac(acc, h[l], tw[e]);
            e += mo;
            if (e >= Nout) e -= Nout;
        }
        ob[mo] = T(scale) * acc;
    }
}

// warp-wide sum of a complex double
__device__ __forceinline__ cx<double> warp_sum(cx<double> v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v.re += __shfl_xor_sync(0xffffffffu, v.re, o);
        v.im += __shfl_xor_sync(0xffffffffu, v.im, o);
    }
    return v;
}

constexpr int kEstMaxNt = 4, kEstMaxNr = 8;

// compute_ls_estimation (channel_estimation/estimators.py:12-61): H = Y s^H (s s^H)^-1.  One warp per realization:
// lanes stride the pilots, B = s s^H (Hermitian positive definite) is Cholesky-factored in double by every lane,
// then row r of H is the conjugate of B^-1 (Y[r] s^H)^H.
template <typename T>
__global__ void __launch_bounds__(128)
ls_estimate_kernel(const cx<T> *__restrict__ Y, const cx<T> *__restrict__ s, int s_per_unit, long long batch, int Nr,
                   int Nt, int P, cx<T> *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long w0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long i = w0; i < batch; i += nw) {
        const cx<T> *si = s + (s_per_unit ? size_t(i) * Nt * P : 0);
        const cx<T> *Yi = Y + size_t(i) * Nr * P;
        cx<double> B[kEstMaxNt][kEstMaxNt];
        for (int a = 0; a < Nt; ++a)
            for (int b = 0; b <= a; ++b) {
                cx<double> acc = {0.0, 0.0};
                for (int p = lane; p < P; p += 32) cmac_conj(acc, cvt<double>(si[b * P + p]), cvt<double>(si[a * P + p]));
                B[a][b] = warp_sum(acc);                         // B[a][b] = sum_p s[a][p] conj(s[b][p])
            }
        // Cholesky B = L L^H (lower), diagonal kept as reciprocal
        cx<double> L[kEstMaxNt][kEstMaxNt];
        double invd[kEstMaxNt];
        for (int j = 0; j < Nt; ++j) {
            double d = B[j][j].re;
            for (int k = 0; k < j; ++k) d -= norm2(L[j][k]);
            const double inv = 1.0 / sqrt(d);
            invd[j] = inv;
            for (int a = j + 1; a < Nt; ++a) {
                cx<double> v = B[a][j];
                for (int k = 0; k < j; ++k) {
                    const cx<double> x = L[a][k], y = L[j][k];
                    v.re -= x.re * y.re + x.im * y.im;
                    v.im -= x.im * y.re - x.re * y.im;
                }
                L[a][j] = {v.re * inv, v.im * inv};
            }
        }
        for (int r = 0; r < Nr; ++r) {
            cx<double> b[kEstMaxNt];
            for (int t = 0; t < Nt; ++t) {
                cx<double> acc = {0.0, 0.0};                    // conj(A[r][t]) = sum_p conj(Y[r][p]) s[t][p]
                for (int p = lane; p < P; p += 32) cmac_conj(acc, cvt<double>(Yi[r * P + p]), cvt<double>(si[t * P + p]));
                b[t] = warp_sum(acc);
            }
            for (int a = 0; a < Nt; ++a) {                       // forward: L v = b
                cx<double> v = b[a];
                for (int k = 0; k < a; ++k) { const cx<double> x = L[a][k], wv = b[k]; v.re -= x.re * wv.re - x.im * wv.im; v.im -= x.re * wv.im + x.im * wv.re; }
                b[a] = {v.re * invd[a], v.im * invd[a]};
            }
            for (int a = Nt - 1; a >= 0; --a) {                  // backward: L^H x = v
                cx<double> v = b[a];
                for (int k = a + 1; k < Nt; ++k) { const cx<double> x = L[k][a], wv = b[k]; v.re -= x.re * wv.re + x.im * wv.im; v.im -= x.re * wv.im - x.im * wv.re; }
                b[a] = {v.re * invd[a], v.im * invd[a]};
            }
            if (lane == 0)
                for (int t = 0; t < Nt; ++t) out[(size_t(i) * Nr + r) * Nt + t] = {T(b[t].re), T(-b[t].im)};
        }
    }
}

// compute_mmse_estimation (estimators.py:100-174), one tx antenna: h = W (Y s^H) P / (s s^H) with
// W = (noise I + P C)^-1 C prepared on the host side of the entry point.
struct EstW { double re[kEstMaxNr][kEstMaxNr], im[kEstMaxNr][kEstMaxNr]; };

template <typename T>
__global__ void __launch_bounds__(128)
mmse_estimate_kernel(const cx<T> *__restrict__ Y, const cx<T> *__restrict__ s, int s_per_unit, long long batch, int Nr,
                     int P, EstW W, cx<T> *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long w0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long i = w0; i < batch; i += nw) {
        const cx<T> *si = s + (s_per_unit ? size_t(i) * P : 0);
        const cx<T> *Yi = Y + size_t(i) * Nr * P;
        cx<double> ss = {0.0, 0.0};
        for (int p = lane; p < P; p += 32) ss.re += norm2(cvt<double>(si[p]));
        ss = warp_sum(ss);
        cx<double> v[kEstMaxNr];
        for (int r = 0; r < Nr; ++r) {
            cx<double> acc = {0.0, 0.0};                        // (Y s^H)[r] = sum_p Y[r][p] conj(s[p])
            for (int p = lane; p < P; p += 32) cmac_conj(acc, cvt<double>(si[p]), cvt<double>(Yi[r * P + p]));
            v[r] = warp_sum(acc);
        }
        const double g = double(P) / ss.re;
        if (lane < Nr) {
            cx<double> acc = {0.0, 0.0};
            for (int c = 0; c < Nr; ++c) cmac(acc, mk<double>(W.re[lane][c], W.im[lane][c]), v[c]);
            out[size_t(i) * Nr + lane] = {T(acc.re * g), T(acc.im * g)};
        }
    }
}

// host: W = (noise I + P C)^-1 C by Gauss-Jordan with partial pivoting (Nr <= 8, one-off per call)
static bool mmse_weight(const double *C, int Nr, double noise, int P, EstW *W) {
    double ar[kEstMaxNr][2 * kEstMaxNr], ai[kEstMaxNr][2 * kEstMaxNr];
    for (int r = 0; r < Nr; ++r)
        for (int c = 0; c < Nr; ++c) {
            const double cr = C[2 * (r * Nr + c)], ci = C[2 * (r * Nr + c) + 1];
            ar[r][c] = P * cr + (r == c ? noise : 0.0); ai[r][c] = P * ci;
            ar[r][Nr + c] = cr; ai[r][Nr + c] = ci;
        }
    for (int j = 0; j < Nr; ++j) {
        int piv = j;
        double best = -1.0;
        for (int r = j; r < Nr; ++r) { const double mag = ar[r][j] * ar[r][j] + ai[r][j] * ai[r][j]; if (mag > best) { best = mag; piv = r; } }
        if (!(best > 0.0)) return false;
        for (int c = 0; c < 2 * Nr; ++c) { std::swap(ar[j][c], ar[piv][c]); std::swap(ai[j][c], ai[piv][c]); }
        const double pr = ar[j][j] / best, pi = -ai[j][j] / best;       // 1 / pivot
        for (int c = 0; c < 2 * Nr; ++c) { const double x = ar[j][c], y = ai[j][c]; ar[j][c] = x * pr - y * pi; ai[j][c] = x * pi + y * pr; }
        for (int r = 0; r < Nr; ++r) {
            if (r == j) continue;
            const double fr = ar[r][j], fi = ai[r][j];
            for (int c = 0; c < 2 * Nr; ++c) { ar[r][c] -= fr * ar[j][c] - fi * ai[j][c]; ai[r][c] -= fr * ai[j][c] + fi * ar[j][c]; }
        }
    }
    for (int r = 0; r < Nr; ++r)
        for (int c = 0; c < Nr; ++c) { W->re[r][c] = ar[r][Nr + c]; W->im[r][c] = ai[r][Nr + c]; }
    return true;
}

static int warp_grid(long long batch, int threads) {
    const long long per = threads / 32;
    long long b = (batch + per - 1) / per;
    return int(b < 1 ? 1 : (b < 148 * 16 ? b : 148 * 16));
}

}  // namespace b200phy

using namespace b200phy;

extern "C" {

int b200phy_refsig_sequence(int dtype, int Nzc, int u, double q_re, double q_im, const int8_t *phi_table, int size,
                            int n_cs, int denominator, double scale_re, double scale_im, void *out, void *stream) {
    if (Nzc < 1 || size < 1) { set_error("sequence sizes must be positive (Nzc=%d, size=%d)", Nzc, size); return B200PHY_ERR_INVALID; }
    if (!phi_table && !(u < Nzc)) { set_error("the root index u=%d must be lower than Nzc=%d", u, Nzc); return B200PHY_ERR_INVALID; }
    if (phi_table && Nzc != 12 && Nzc != 24) { set_error("Invalid root sequence size"); return B200PHY_ERR_INVALID; }
    if (denominator < 1 || !(abs(n_cs) < denominator)) { set_error("cyclic shift n_cs=%d must be in [0, %d)", n_cs, denominator); return B200PHY_ERR_INVALID; }
    PhiTable tab = {};
    if (phi_table) for (int i = 0; i < Nzc; ++i) tab.v[i] = phi_table[i];
    const int ncs = ((n_cs % denominator) + denominator) % denominator;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = (size + 255) / 256;
    if (dtype == B200PHY_F32)
        refsig_sequence_kernel<float><<<grid, 256, 0, st>>>(Nzc, u, q_re, q_im, phi_table != nullptr, tab, size, ncs, denominator, scale_re, scale_im, (cx<float> *)out);
    else
        refsig_sequence_kernel<double><<<grid, 256, 0, st>>>(Nzc, u, q_re, q_im, phi_table != nullptr, tab, size, ncs, denominator, scale_re, scale_im, (cx<double> *)out);
    B200_CHECK_LAUNCH("refsig_sequence_kernel");
    return B200PHY_OK;
}

int b200phy_cazac_estimate(int dtype, const void *ref_seq, const void *y, const double *cover_re_im, int n_cover,
                           int64_t batch, int Nsc, int num_taps_to_keep, int size_multiplier, double scale, void *out,
                           void *stream) {
    if (Nsc < 1 || size_multiplier < 1 || num_taps_to_keep < 0) { set_error("CAZAC estimator: Nsc=%d, size_multiplier=%d, num_taps_to_keep=%d", Nsc, size_multiplier, num_taps_to_keep); return B200PHY_ERR_INVALID; }
    if (n_cover < 1 || n_cover > 4) { set_error("cover code length %d must be in [1, 4]", n_cover); return B200PHY_ERR_UNSUPPORTED; }
    const long long Nout = (long long)size_multiplier * Nsc;
    const size_t es = dtype == B200PHY_F32 ? 8 : 16;
    int n_keep = num_taps_to_keep + 1;
    if (n_keep > Nsc) n_keep = Nsc;                               // y[0:num_taps_to_keep + 1] of an Nsc-point IFFT
    const size_t smem = es * size_t(Nout + Nsc + n_keep);
    if (smem > 200 * 1024) { set_error("CAZAC estimator: %lld output subcarriers exceed the shared-memory tile", Nout); return B200PHY_ERR_UNSUPPORTED; }
    if (batch <= 0) return B200PHY_OK;
    Cover cc = {};
    for (int c = 0; c < n_cover; ++c) { cc.re[c] = cover_re_im ? cover_re_im[2 * c] : 1.0; cc.im[c] = cover_re_im ? cover_re_im[2 * c + 1] : 0.0; }
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == B200PHY_F32) {
        auto k = cazac_estimate_kernel<float>;
        if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        k<<<unsigned(batch), 256, smem, st>>>((const cx<float> *)ref_seq, (const cx<float> *)y, cc, n_cover, Nsc, n_keep, size_multiplier, scale, (cx<float> *)out);
    } else {
        auto k = cazac_estimate_kernel<double>;
        if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        k<<<unsigned(batch), 256, smem, st>>>((const cx<double> *)ref_seq, (const cx<double> *)y, cc, n_cover, Nsc, n_keep, size_multiplier, scale, (cx<double> *)out);
    }
    B200_CHECK_LAUNCH("cazac_estimate_kernel");
    return B200PHY_OK;
}

int b200phy_ls_estimate(int dtype, const void *Y, const void *s, int s_per_unit, int64_t batch, int Nr, int Nt, int P,
                        void *out, void *stream) {
    if (Nr < 1 || Nt < 1 || P < 1) { set_error("LS estimator: Nr=%d, Nt=%d, pilots=%d must be positive", Nr, Nt, P); return B200PHY_ERR_INVALID; }
    if (Nt > kEstMaxNt) { set_error("LS estimator: Nt=%d exceeds %d", Nt, kEstMaxNt); return B200PHY_ERR_UNSUPPORTED; }
    if (P < Nt) { set_error("LS estimator: %d pilots cannot separate %d transmit antennas (s s^H is singular)", P, Nt); return B200PHY_ERR_INVALID; }
    if (batch <= 0) return B200PHY_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = warp_grid(batch, 128);
    if (dtype == B200PHY_F32) ls_estimate_kernel<float><<<grid, 128, 0, st>>>((const cx<float> *)Y, (const cx<float> *)s, s_per_unit, batch, Nr, Nt, P, (cx<float> *)out);
    else ls_estimate_kernel<double><<<grid, 128, 0, st>>>((const cx<double> *)Y, (const cx<double> *)s, s_per_unit, batch, Nr, Nt, P, (cx<double> *)out);
    B200_CHECK_LAUNCH("ls_estimate_kernel");
    return B200PHY_OK;
}

int b200phy_mmse_estimate(int dtype, const void *Y, const void *s, int s_per_unit, int64_t batch, int Nr, int P,
                          double noise_power, const double *C_re_im, void *out, void *stream) {
    if (Nr < 1 || P < 1) { set_error("MMSE estimator: Nr=%d, pilots=%d must be positive", Nr, P); return B200PHY_ERR_INVALID; }
    if (Nr > kEstMaxNr) { set_error("MMSE estimator: Nr=%d exceeds %d", Nr, kEstMaxNr); return B200PHY_ERR_UNSUPPORTED; }
    if (!C_re_im) { set_error("MMSE estimator: covariance matrix is NULL"); return B200PHY_ERR_INVALID; }
    EstW W = {};
    if (!mmse_weight(C_re_im, Nr, noise_power, P, &W)) { set_error("MMSE estimator: noise_power I + num_pilots C is singular"); return B200PHY_ERR_INVALID; }
    if (batch <= 0) return B200PHY_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = warp_grid(batch, 128);
    if (dtype == B200PHY_F32) mmse_estimate_kernel<float><<<grid, 128, 0, st>>>((const cx<float> *)Y, (const cx<float> *)s, s_per_unit, batch, Nr, P, W, (cx<float> *)out);
    else mmse_estimate_kernel<double><<<grid, 128, 0, st>>>((const cx<double> *)Y, (const cx<double> *)s, s_per_unit, batch, Nr, P, W, (cx<double> *)out);
    B200_CHECK_LAUNCH("mmse_estimate_kernel");
    return B200PHY_OK;
}

}  // extern "C"

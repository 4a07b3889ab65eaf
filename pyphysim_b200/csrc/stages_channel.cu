// stages_channel.cu — stage ops behind JakesSampleGenerator.generate_more_samples,
// TdlChannel.corrupt_data and OfdmOneTapEqualizer.equalize_data / per-subcarrier Blast.decode.
// These are the API-parity path (arrays in HBM between stages); the throughput path is the fused
// kernel in ofdm_tdl.cuh.
#include <vector>

#include "ofdm_tdl.cuh"

namespace b200phy {

struct TapTable {
    int n_taps;
    int delays[B200PHY_MAX_TAPS];
    double amp[B200PHY_MAX_TAPS];       // sqrt(P_l)
};

// h[p][n] = L^-1/2 sum_o exp(j(2 pi Fd cos(phi[o][p]) t_n + psi[o][p])), t_n = t0 + n Ts (1+1e-10)
// (channels/fading_generators.py:459-467, 519-523).  Phase is formed in double and reduced before
// the (T-precision) sincos so that long clocks (large t0) keep their accuracy.
template <typename T>
__global__ void __launch_bounds__(256)
jakes_kernel(const T *__restrict__ phi, const T *__restrict__ psi, int L, long long P, long long N,
             double w0, double Ts1, double t0, cx<T> *__restrict__ h) {
    const T norm = T(1.0 / sqrt(double(L)));
    for (long long it = blockIdx.x * (long long)blockDim.x + threadIdx.x; it < P * N;
         it += (long long)gridDim.x * blockDim.x) {
        const long long pi = it / N, n = it % N;
        const double t = fma(double(n), Ts1, t0);
        cx<T> acc = {T(0), T(0)};
        for (int o = 0; o < L; ++o) {
            const double w = w0 * cos(double(phi[o * P + pi]));
            T s, c;
            sincos_t(T(reduce_2pi(fma(w, t, double(psi[o * P + pi])))), &s, &c);
            acc.re += c;
            acc.im += s;
        }
        h[it] = norm * acc;
    }
}

// y[r][m] = sum_l sum_t sqrt(P_l) fading[l][r][t][m-d_l] x[t][m-d_l]   (fading.py:1089-1117)
template <typename T>
__global__ void __launch_bounds__(256)
tdl_apply_kernel(const cx<T> *__restrict__ x, const cx<T> *__restrict__ fading, TapTable tt, int Nr,
                 int Nt, long long N, cx<T> *__restrict__ y) {
    const long long M = N + tt.delays[tt.n_taps - 1];
    for (long long it = blockIdx.x * (long long)blockDim.x + threadIdx.x; it < Nr * M;
         it += (long long)gridDim.x * blockDim.x) {
        const int r = int(it / M);
        const long long mm = it % M;
        cx<T> acc = {T(0), T(0)};
        for (int l = 0; l < tt.n_taps; ++l) {
            const long long n = mm - tt.delays[l];
            if (n < 0 || n >= N) continue;
            const T a = T(tt.amp[l]);
            for (int t = 0; t < Nt; ++t)
                cmac(acc, a * fading[((size_t(l) * Nr + r) * Nt + t) * N + n], x[size_t(t) * N + n]);
        }
        y[it] = acc;
    }
}

// mean of sqrt(P_l) fading[l][r][t][.] over the S samples of each OFDM symbol -> gbar[sym][l][r][t]
template <typename T>
__global__ void __launch_bounds__(256)
tap_mean_kernel(const cx<T> *__restrict__ fading, TapTable tt, int NrNt, int n_sym, int S,
                cx<T> *__restrict__ gbar) {
    __shared__ T sre[256], sim[256];
    const int item = blockIdx.x;                      // (l * NrNt + a) * n_sym + sym
    const int sym = item % n_sym, la = item / n_sym, l = la / NrNt;
    const cx<T> *src = fading + (size_t(la) * n_sym + sym) * S;
    T re = T(0), im = T(0);
    for (int i = threadIdx.x; i < S; i += 256) { re += src[i].re; im += src[i].im; }
    sre[threadIdx.x] = re; sim[threadIdx.x] = im;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { sre[threadIdx.x] += sre[threadIdx.x + o]; sim[threadIdx.x] += sim[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const T a = T(tt.amp[l]) / T(S);
        gbar[(size_t(sym) * tt.n_taps * NrNt) + la] = {a * sre[0], a * sim[0]};
    }
}

// per used subcarrier: H_k = sum_l gbar_l exp(-2 pi i k d_l / fft), then one-tap divide (1x1) or the
// Blast receive filter (mimo.py:590-660)
template <typename T, int NT>
__global__ void __launch_bounds__(128)
equalize_kernel(const cx<T> *__restrict__ y, const cx<T> *__restrict__ gbar, TapTable tt, int Nr,
                int n_sym, int fft, int used, double fnv, cx<T> *__restrict__ out) {
    const int half = used / 2;
    const double snt = sqrt(double(NT));
    for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < n_sym * used; it += gridDim.x * blockDim.x) {
        const int sym = it / used, q = it % used;
        const int k = bin_of(q, fft, used, half);
        cx<T> H[B200PHY_MAX_ANT][NT];
        for (int r = 0; r < B200PHY_MAX_ANT; ++r)
            for (int t = 0; t < NT; ++t) H[r][t] = {T(0), T(0)};
        for (int l = 0; l < tt.n_taps; ++l) {
            double s, c;
            sincospi(-2.0 * double((k * tt.delays[l]) & (fft - 1)) / double(fft), &s, &c);
            const cx<T> w = {T(c), T(s)};
            for (int r = 0; r < B200PHY_MAX_ANT; ++r)
                if (r < Nr)
#pragma unroll
                    for (int t = 0; t < NT; ++t)
                        cmac(H[r][t], gbar[((size_t(sym) * tt.n_taps + l) * Nr + r) * NT + t], w);
        }
        cx<T> yy[B200PHY_MAX_ANT];
        for (int r = 0; r < B200PHY_MAX_ANT; ++r)
            if (r < Nr) yy[r] = y[(size_t(r) * n_sym + sym) * used + q];
        if (NT == 1 && Nr == 1) {
            out[it] = cdiv(yy[0], H[0][0]);
        } else {
            HermSolver<NT> sol;
            sol.factor_from_channel(H, Nr, fnv);
            cx<double> b[NT];
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                b[t] = {0.0, 0.0};
                for (int r = 0; r < B200PHY_MAX_ANT; ++r)
                    if (r < Nr) cmac_conj(b[t], cvt<double>(H[r][t]), cvt<double>(yy[r]));
            }
            sol.solve(b);
#pragma unroll
            for (int t = 0; t < NT; ++t) out[size_t(it) * NT + t] = {T(b[t].re * snt), T(b[t].im * snt)};
        }
    }
}

// out[k][a][n] = sum_l taps[l][a][n] W^(k d_l)
template <typename T>
__global__ void __launch_bounds__(256)
freq_response_kernel(const cx<T> *__restrict__ taps, TapTable tt, long long AN, int fft, cx<T> *__restrict__ out) {
    for (long long it = blockIdx.x * (long long)blockDim.x + threadIdx.x; it < AN * fft;
         it += (long long)gridDim.x * blockDim.x) {
        const int k = int(it / AN);
        const long long an = it % AN;
        cx<T> acc = {T(0), T(0)};
        for (int l = 0; l < tt.n_taps; ++l) {
            double s, c;
            sincospi(-2.0 * double((long long)k * tt.delays[l] % fft) / double(fft), &s, &c);
            cmac(acc, taps[size_t(l) * AN + an], mk<T>(T(c), T(s)));
        }
        out[it] = acc;
    }
}

// y[r][b*bs+i] = sum_t H[car[i]][r][t][b] * x[t][b*bs+i]   (fading.py:1212-1270)
template <typename T>
__global__ void __launch_bounds__(256)
freq_apply_kernel(const cx<T> *__restrict__ H, const cx<T> *__restrict__ x, const int *__restrict__ car,
                  int bs, int Nr, int Nt, long long B, cx<T> *__restrict__ y) {
    const long long N = B * bs;
    for (long long it = blockIdx.x * (long long)blockDim.x + threadIdx.x; it < Nr * N;
         it += (long long)gridDim.x * blockDim.x) {
        const int r = int(it / N);
        const long long n = it % N, b = n / bs;
        const int i = int(n % bs);
        const long long k = car ? car[i] : i;
        cx<T> acc = {T(0), T(0)};
        for (int t = 0; t < Nt; ++t)
            cmac(acc, H[((k * Nr + r) * Nt + t) * B + b], x[size_t(t) * N + n]);
        y[it] = acc;
    }
}

struct RowScales { int rows; double s[64]; };

template <typename T>
__global__ void __launch_bounds__(256)
scale_rows_kernel(cx<T> *x, RowScales rs, long long cols) {
    for (long long it = blockIdx.x * (long long)blockDim.x + threadIdx.x; it < rs.rows * cols;
         it += (long long)gridDim.x * blockDim.x) {
        const T a = T(rs.s[it / cols]);
        x[it] = a * x[it];
    }
}

static int fill_taps(const double *tap_powers, const int32_t *delays, int n_taps, TapTable *tt) {
    if (!tap_powers || !delays) { set_error("tap_powers/delays is NULL"); return B200PHY_ERR_INVALID; }
    if (n_taps < 1 || n_taps > B200PHY_MAX_TAPS) { set_error("n_taps=%d must be in [1, %d]", n_taps, B200PHY_MAX_TAPS); return B200PHY_ERR_UNSUPPORTED; }
    tt->n_taps = n_taps;
    for (int l = 0; l < n_taps; ++l) {
        if (delays[l] < 0 || (l && delays[l] <= delays[l - 1])) { set_error("tap delays must be non-negative and strictly increasing"); return B200PHY_ERR_INVALID; }
        tt->delays[l] = delays[l];
        tt->amp[l] = sqrt(tap_powers[l]);
    }
    return B200PHY_OK;
}

static int blocks_for(long long n, int threads) {
    long long b = (n + threads - 1) / threads;
    return int(b < 1 ? 1 : (b < 148 * 8 ? b : 148 * 8));
}

template <typename T>
static int run_equalize(const void *y, const void *fading, const TapTable &tt, int Nr, int Nt, int n_sym,
                        int fft, int cp, int used, double fnv, void *out, cudaStream_t st) {
    cx<T> *gbar = nullptr;
    const size_t ng = size_t(n_sym) * tt.n_taps * Nr * Nt;
    int e = check_cuda(cudaMallocAsync((void **)&gbar, sizeof(cx<T>) * ng, st), "cudaMallocAsync(gbar)");
    if (e) return e;
    tap_mean_kernel<T><<<int(ng), 256, 0, st>>>((const cx<T> *)fading, tt, Nr * Nt, n_sym, fft + cp, gbar);
    count_launch();
    const int grid = blocks_for((long long)n_sym * used, 128);
#define B200_EQ(NT_) equalize_kernel<T, NT_><<<grid, 128, 0, st>>>((const cx<T> *)y, gbar, tt, Nr, n_sym, fft, used, fnv, (cx<T> *)out)
    switch (Nt) { case 1: B200_EQ(1); break; case 2: B200_EQ(2); break; case 3: B200_EQ(3); break; default: B200_EQ(4); }
#undef B200_EQ
    count_launch();
    e = check_cuda(cudaGetLastError(), "equalize kernels");
    cudaFreeAsync(gbar, st);
    return e;
}

}  // namespace b200phy

using namespace b200phy;

extern "C" {

int b200phy_jakes(int dtype, const void *phi, const void *psi, int L, int64_t P, int64_t N, double Fd,
                  double Ts, double t0, void *h, void *stream) {
    if (L < 1) { set_error("L must be positive"); return B200PHY_ERR_INVALID; }
    if (P <= 0 || N <= 0) return B200PHY_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = blocks_for(P * N, 256);
    const double w0 = 2.0 * M_PI * Fd, Ts1 = Ts * 1.0000000001;
    if (dtype == B200PHY_F32)
        jakes_kernel<float><<<grid, 256, 0, st>>>((const float *)phi, (const float *)psi, L, P, N, w0, Ts1, t0, (cx<float> *)h);
    else
        jakes_kernel<double><<<grid, 256, 0, st>>>((const double *)phi, (const double *)psi, L, P, N, w0, Ts1, t0, (cx<double> *)h);
    B200_CHECK_LAUNCH("jakes_kernel");
    return B200PHY_OK;
}

int b200phy_tdl_apply(int dtype, const void *x, const void *fading, const double *tap_powers,
                      const int32_t *delays, int n_taps, int Nr, int Nt, int64_t N, void *y,
                      void *stream) {
    TapTable tt;
    int e = fill_taps(tap_powers, delays, n_taps, &tt);
    if (e) return e;
    if (Nr < 1 || Nt < 1) { set_error("Nr, Nt must be positive"); return B200PHY_ERR_INVALID; }
    if (N <= 0) return B200PHY_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = blocks_for((long long)Nr * (N + tt.delays[n_taps - 1]), 256);
    if (dtype == B200PHY_F32)
        tdl_apply_kernel<float><<<grid, 256, 0, st>>>((const cx<float> *)x, (const cx<float> *)fading, tt, Nr, Nt, N, (cx<float> *)y);
    else
        tdl_apply_kernel<double><<<grid, 256, 0, st>>>((const cx<double> *)x, (const cx<double> *)fading, tt, Nr, Nt, N, (cx<double> *)y);
    B200_CHECK_LAUNCH("tdl_apply_kernel");
    return B200PHY_OK;
}

int b200phy_tdl_freq_response(int dtype, const void *taps, const int32_t *delays, int n_taps, int64_t A,
                              int64_t N, int fft, void *out, void *stream) {
    std::vector<double> ones(n_taps > 0 ? n_taps : 1, 1.0);
    TapTable tt;
    int e = fill_taps(ones.data(), delays, n_taps, &tt);
    if (e) return e;
    if (fft < 1) { set_error("fft_size must be positive"); return B200PHY_ERR_INVALID; }
    if (A * N <= 0) return B200PHY_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = blocks_for(A * N * fft, 256);
    if (dtype == B200PHY_F32) freq_response_kernel<float><<<grid, 256, 0, st>>>((const cx<float> *)taps, tt, A * N, fft, (cx<float> *)out);
    else freq_response_kernel<double><<<grid, 256, 0, st>>>((const cx<double> *)taps, tt, A * N, fft, (cx<double> *)out);
    B200_CHECK_LAUNCH("freq_response_kernel");
    return B200PHY_OK;
}

int b200phy_freq_apply(int dtype, const void *H, const void *x, const int32_t *carriers, int fft, int bs,
                       int Nr, int Nt, int64_t B, void *y, void *stream) {
    if (Nr < 1 || Nt < 1 || bs < 1 || fft < 1) { set_error("freq_apply: bad dimensions"); return B200PHY_ERR_INVALID; }
    if (!carriers && bs != fft) { set_error("freq_apply: without carrier indexes the block size must equal fft_size"); return B200PHY_ERR_INVALID; }
    if (B <= 0) return B200PHY_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = blocks_for((long long)Nr * B * bs, 256);
    if (dtype == B200PHY_F32) freq_apply_kernel<float><<<grid, 256, 0, st>>>((const cx<float> *)H, (const cx<float> *)x, carriers, bs, Nr, Nt, B, (cx<float> *)y);
    else freq_apply_kernel<double><<<grid, 256, 0, st>>>((const cx<double> *)H, (const cx<double> *)x, carriers, bs, Nr, Nt, B, (cx<double> *)y);
    B200_CHECK_LAUNCH("freq_apply_kernel");
    return B200PHY_OK;
}

int b200phy_scale_rows(int dtype, void *x, int rows, int64_t cols, const double *scales, void *stream) {
    if (rows < 1 || rows > 64 || !scales) { set_error("scale_rows: rows=%d must be in [1, 64]", rows); return B200PHY_ERR_INVALID; }
    if (cols <= 0) return B200PHY_OK;
    RowScales rs;
    rs.rows = rows;
    for (int i = 0; i < rows; ++i) rs.s[i] = scales[i];
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = blocks_for(rows * cols, 256);
    if (dtype == B200PHY_F32) scale_rows_kernel<float><<<grid, 256, 0, st>>>((cx<float> *)x, rs, cols);
    else scale_rows_kernel<double><<<grid, 256, 0, st>>>((cx<double> *)x, rs, cols);
    B200_CHECK_LAUNCH("scale_rows_kernel");
    return B200PHY_OK;
}

int b200phy_ofdm_equalize(int dtype, const void *y, const void *fading, const double *tap_powers,
                          const int32_t *delays, int n_taps, int Nr, int Nt, int n_sym, int fft,
                          int cp, int used, double filter_noise_var, void *out, void *stream) {
    TapTable tt;
    int e = fill_taps(tap_powers, delays, n_taps, &tt);
    if (e) return e;
    if (Nr < 1 || Nr > B200PHY_MAX_ANT || Nt < 1 || Nt > B200PHY_MAX_ANT) { set_error("Nr=%d, Nt=%d must be in [1, %d]", Nr, Nt, B200PHY_MAX_ANT); return B200PHY_ERR_UNSUPPORTED; }
    if (!(filter_noise_var >= 0.0)) { set_error("Noise variance must be a non-negative value."); return B200PHY_ERR_INVALID; }
    if (filter_noise_var == 0.0 && Nt > Nr) { set_error("ZF needs Nt <= Nr"); return B200PHY_ERR_UNSUPPORTED; }
    if (fft < 2 || (fft & (fft - 1))) { set_error("fft_size must be a power of two"); return B200PHY_ERR_UNSUPPORTED; }
    if (n_sym <= 0) return B200PHY_OK;
    cudaStream_t st = (cudaStream_t)stream;
    return dtype == B200PHY_F32 ? run_equalize<float>(y, fading, tt, Nr, Nt, n_sym, fft, cp, used, filter_noise_var, out, st)
                                : run_equalize<double>(y, fading, tt, Nr, Nt, n_sym, fft, cp, used, filter_noise_var, out, st);
}

}  // extern "C"

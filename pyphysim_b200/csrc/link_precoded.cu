// link_precoded.cu — SURVEY.md §8f row next-3: the MIMO schemes with a channel-dependent precoder.
//
//   b200phy_svd / b200phy_gmd   batched decompositions of small channel matrices (csrc/svd.cuh): what
//                               SVDMimo / GMDMimo take from np.linalg.svd and util.misc.gmd
//                               (pyphysim/mimo/mimo.py:855-898, 974-1019; util/misc.py:18-159)
//   b200phy_mat_apply           Y = A X for one small matrix and a block of symbol vectors: the
//                               W.dot(X) / G_H.dot(Y) of encode / decode (mimo.py:900-948, 1021-1067)
//   b200phy_link_precoded       fused Monte Carlo link, one channel realization per thread
//                               (apps/mimo/simulate_mimo.py:68-142 with mimo.SVDMimo / GMDMimo / MRT):
//                               draw H, decompose it, then S symbol vectors: modulate -> precode ->
//                               H x + n -> receive filter -> demap -> count.
#include "common.cuh"
#include "linalg.cuh"
#include "rng.cuh"
#include "svd.cuh"

namespace b200phy {

constexpr int kPT = 128;        // threads per CTA: the decompositions are register-heavy

static int grid_units(long long n, int threads) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long blocks = (n + threads - 1) / threads;
    const long long cap = (long long)sms * 8;
    return int(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

template <typename T> __device__ __forceinline__ cx<T> ldc(const cx<T> *__restrict__ p) {
    if constexpr (sizeof(T) == 4) {
        const float2 v = __ldg(reinterpret_cast<const float2 *>(p));
        return {v.x, v.y};
    } else {
        const double2 v = __ldg(reinterpret_cast<const double2 *>(p));
        return {v.x, v.y};
    }
}

// ================================================================= batched decompositions
template <int NR, int NT>
__global__ void __launch_bounds__(kPT)
svd_kernel(const cx<double> *__restrict__ H, long long batch, cx<double> *__restrict__ U,
           double *__restrict__ S, cx<double> *__restrict__ V) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < batch;
         i += (long long)gridDim.x * blockDim.x) {
        cx<double> h[NR][NT];
#pragma unroll
        for (int r = 0; r < NR; ++r)
#pragma unroll
            for (int t = 0; t < NT; ++t) h[r][t] = ldc(H + (i * NR + r) * NT + t);
        SmallSvd<NR, NT> d;
        d.compute(h);
#pragma unroll
        for (int r = 0; r < NR; ++r)
#pragma unroll
            for (int t = 0; t < NT; ++t) U[(i * NR + r) * NT + t] = d.U[r][t];
#pragma unroll
        for (int t = 0; t < NT; ++t) S[i * NT + t] = d.S[t];
#pragma unroll
        for (int r = 0; r < NT; ++r)
#pragma unroll
            for (int t = 0; t < NT; ++t) V[(i * NT + r) * NT + t] = d.V[r][t];
    }
}

template <int NR, int NT>
__global__ void __launch_bounds__(kPT)
gmd_kernel(const cx<double> *__restrict__ U, const double *__restrict__ S, const cx<double> *__restrict__ V,
           long long batch, cx<double> *__restrict__ Q, double *__restrict__ R, cx<double> *__restrict__ P) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < batch;
         i += (long long)gridDim.x * blockDim.x) {
        cx<double> q[NR][NT], p[NT][NT];
        double s[NT], r[NT][NT];
#pragma unroll
        for (int a = 0; a < NR; ++a)
#pragma unroll
            for (int b = 0; b < NT; ++b) q[a][b] = ldc(U + (i * NR + a) * NT + b);
#pragma unroll
        for (int b = 0; b < NT; ++b) s[b] = S[i * NT + b];
#pragma unroll
        for (int a = 0; a < NT; ++a)
#pragma unroll
            for (int b = 0; b < NT; ++b) p[a][b] = ldc(V + (i * NT + a) * NT + b);
        small_gmd<NR, NT>(q, s, p, r);
#pragma unroll
        for (int a = 0; a < NR; ++a)
#pragma unroll
            for (int b = 0; b < NT; ++b) Q[(i * NR + a) * NT + b] = q[a][b];
#pragma unroll
        for (int a = 0; a < NT; ++a)
#pragma unroll
            for (int b = 0; b < NT; ++b) { R[(i * NT + a) * NT + b] = r[a][b]; P[(i * NT + a) * NT + b] = p[a][b]; }
    }
}

template <template <int, int> class Launch, typename... Args>
static int dispatch_shape(int Nr, int Nt, Args... args) {
#define B200_SHAPE(R_, T_) if (Nr == R_ && Nt == T_) return Launch<R_, T_>::go(args...)
    B200_SHAPE(1, 1); B200_SHAPE(2, 1); B200_SHAPE(3, 1); B200_SHAPE(4, 1);
    B200_SHAPE(2, 2); B200_SHAPE(3, 2); B200_SHAPE(4, 2);
    B200_SHAPE(3, 3); B200_SHAPE(4, 3); B200_SHAPE(4, 4);
#undef B200_SHAPE
    set_error("decomposition supports 1 <= Nt <= Nr <= %d (got Nr=%d, Nt=%d)", B200PHY_MAX_ANT, Nr, Nt);
    return B200PHY_ERR_UNSUPPORTED;
}

template <int NR, int NT> struct LaunchSvd {
    static int go(const void *H, int64_t batch, void *U, double *S, void *V, cudaStream_t st) {
        svd_kernel<NR, NT><<<grid_units(batch, kPT), kPT, 0, st>>>((const cx<double> *)H, (long long)batch,
                                                                  (cx<double> *)U, S, (cx<double> *)V);
        B200_CHECK_LAUNCH("svd_kernel");
        return B200PHY_OK;
    }
};
template <int NR, int NT> struct LaunchGmd {
    static int go(const void *U, const double *S, const void *V, int64_t batch, void *Q, double *R, void *P,
                  cudaStream_t st) {
        gmd_kernel<NR, NT><<<grid_units(batch, kPT), kPT, 0, st>>>((const cx<double> *)U, S, (const cx<double> *)V,
                                                                  (long long)batch, (cx<double> *)Q, R, (cx<double> *)P);
        B200_CHECK_LAUNCH("gmd_kernel");
        return B200PHY_OK;
    }
};

// ================================================================= Y = A X
template <typename T>
__global__ void __launch_bounds__(256)
mat_apply_kernel(const cx<T> *__restrict__ A, int rows, int cols, const cx<T> *__restrict__ X, long long n,
                 cx<T> *__restrict__ Y) {
    __shared__ cx<T> a[B200PHY_MAX_ANT * 2 * B200PHY_MAX_ANT * 2];
    for (int k = threadIdx.x; k < rows * cols; k += blockDim.x) a[k] = A[k];
    __syncthreads();
    for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x) {
        cx<T> x[B200PHY_MAX_ANT * 2];
#pragma unroll
        for (int c = 0; c < B200PHY_MAX_ANT * 2; ++c)
            if (c < cols) x[c] = ldc(X + (long long)c * n + j);
        for (int r = 0; r < rows; ++r) {
            cx<T> acc = {T(0), T(0)};
#pragma unroll
            for (int c = 0; c < B200PHY_MAX_ANT * 2; ++c)
                if (c < cols) cmac(acc, a[r * cols + c], x[c]);
            Y[(long long)r * n + j] = acc;
        }
    }
}

// ================================================================= fused link
// SCHEME: B200PHY_MIMO_SVD or B200PHY_MIMO_GMD; square N x N channel, N layers.
// Symbol p = l * S + s of a realization is layer l at time s (X = transmit_data.reshape(Nt, -1)).
template <typename T, bool FUSED, int N, int SCHEME>
__global__ void __launch_bounds__(kPT)
precoded_kernel(Modem m, const cx<T> *__restrict__ tab_g, int S, T sigma, double fnv, const __grid_constant__ PhiloxKey seed,
                uint64_t first_unit, long long n, const uint8_t *__restrict__ idx,
                const cx<T> *__restrict__ Hg, const cx<T> *__restrict__ noise,
                uint8_t *__restrict__ idx_hat, cx<T> *__restrict__ dec_out, unsigned long long *counters) {
    __shared__ cx<T> tab[256];
    if (m.kind != B200PHY_MODEM_BPSK)
        for (int k = threadIdx.x; k < m.M; k += blockDim.x) tab[k] = tab_g[k];
    __syncthreads();
    unsigned sym_err = 0, bit_err = 0;
    const double snt = sqrt(double(N)), rsnt = 1.0 / snt;
    const int row = 2 * ((S + 1) / 2);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const uint64_t unit = first_unit + uint64_t(i);
        cx<T> H[N][N];
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int t = 0; t < N; ++t) {
                if constexpr (FUSED) H[r][t] = cnormal_at<T>(seed, STREAM_CHANNEL, unit, r * N + t);
                else H[r][t] = ldc(Hg + (i * N + r) * N + t);
            }
        cx<T> W[N][N], G[N][N];               // precoder; SVD: receive filter, GMD: equivalent channel Q R
        HermSolver<N> sol;
        {
            SmallSvd<N, N> d;
            d.compute(H);
            if constexpr (SCHEME == B200PHY_MIMO_SVD) {
#pragma unroll
                for (int a = 0; a < N; ++a)
#pragma unroll
                    for (int b = 0; b < N; ++b) {
                        W[a][b] = {T(d.V[a][b].re * rsnt), T(d.V[a][b].im * rsnt)};
                        const double g = snt / d.S[a];                                  // diag(1/S) U^H sqrt(Nt)
                        G[a][b] = {T(d.U[b][a].re * g), T(-d.U[b][a].im * g)};
                    }
            } else {
                double R[N][N];
                small_gmd<N, N>(d.U, d.S, d.V, R);                                       // U -> Q, V -> P
                cx<double> Heq[N][N];
#pragma unroll
                for (int a = 0; a < N; ++a)
#pragma unroll
                    for (int b = 0; b < N; ++b) {
                        W[a][b] = {T(d.V[a][b].re * rsnt), T(d.V[a][b].im * rsnt)};
                        cx<double> acc = {0.0, 0.0};
#pragma unroll
                        for (int k = 0; k <= b; ++k) { acc.re += d.U[a][k].re * R[k][b]; acc.im += d.U[a][k].im * R[k][b]; }
                        Heq[a][b] = acc;
                        G[a][b] = cvt<T>(acc);
                    }
                sol.factor_from_channel(Heq, N, fnv);                                    // Blast filter of Q R
            }
        }
        for (int s = 0; s < S; ++s) {
            int a[N];
            cx<T> sym[N], x[N], y[N];
#pragma unroll
            for (int l = 0; l < N; ++l) {
                const int p = l * S + s;
                if constexpr (FUSED)
                    a[l] = int(lane_of(rng_block(seed, STREAM_DATA, unit, uint64_t(p >> 2)), p & 3) >> (32 - m.bits));
                else
                    a[l] = idx[i * S * N + p];
                sym[l] = map_symbol<T>(m, tab, a[l]);
            }
#pragma unroll
            for (int t = 0; t < N; ++t) {
                x[t] = {T(0), T(0)};
#pragma unroll
                for (int l = 0; l < N; ++l) cmac(x[t], W[t][l], sym[l]);
            }
#pragma unroll
            for (int r = 0; r < N; ++r) {
                cx<T> nz;
                if constexpr (FUSED) nz = cnormal_at<T>(seed, STREAM_NOISE, unit, uint64_t(r) * row + s);
                else nz = ldc(noise + (i * N + r) * S + s);
                y[r] = sigma * nz;
#pragma unroll
                for (int t = 0; t < N; ++t) cmac(y[r], H[r][t], x[t]);
            }
            cx<T> z[N];
            if constexpr (SCHEME == B200PHY_MIMO_SVD) {
#pragma unroll
                for (int l = 0; l < N; ++l) {
                    z[l] = {T(0), T(0)};
#pragma unroll
                    for (int r = 0; r < N; ++r) cmac(z[l], G[l][r], y[r]);
                }
            } else {
                cx<double> b[N];
#pragma unroll
                for (int l = 0; l < N; ++l) {
                    b[l] = {0.0, 0.0};
#pragma unroll
                    for (int r = 0; r < N; ++r) cmac_conj(b[l], cvt<double>(G[r][l]), cvt<double>(y[r]));
                }
                sol.solve(b);
#pragma unroll
                for (int l = 0; l < N; ++l) z[l] = {T(b[l].re * snt), T(b[l].im * snt)};
            }
#pragma unroll
            for (int l = 0; l < N; ++l) {
                const int e = demap_symbol<T>(m, tab, z[l]);
                sym_err += (e != a[l]);
                bit_err += __popc(e ^ a[l]);
                const long long o = i * S * N + (long long)l * S + s;
                if (idx_hat) idx_hat[o] = (uint8_t)e;
                if (dec_out) dec_out[o] = z[l];
            }
        }
    }
    flush_counters(sym_err, bit_err, counters);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        atomicAdd(&counters[2], (unsigned long long)n * S * N);
        atomicAdd(&counters[3], (unsigned long long)n * S * N * m.bits);
    }
}

// MRT (mimo.py:666-783): one receive antenna, one layer; W = conj(h) / |h| / sqrt(Nt), G = sqrt(Nt) / sum |h|
template <typename T, bool FUSED, int NT>
__global__ void __launch_bounds__(256)
mrt_kernel(Modem m, const cx<T> *__restrict__ tab_g, int S, T sigma, const __grid_constant__ PhiloxKey seed, uint64_t first_unit,
           long long n, const uint8_t *__restrict__ idx, const cx<T> *__restrict__ Hg,
           const cx<T> *__restrict__ noise, uint8_t *__restrict__ idx_hat, cx<T> *__restrict__ dec_out,
           unsigned long long *counters) {
    __shared__ cx<T> tab[256];
    if (m.kind != B200PHY_MODEM_BPSK)
        for (int k = threadIdx.x; k < m.M; k += blockDim.x) tab[k] = tab_g[k];
    __syncthreads();
    unsigned sym_err = 0, bit_err = 0;
    const double snt = sqrt(double(NT));
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const uint64_t unit = first_unit + uint64_t(i);
        cx<T> h[NT], W[NT];
        double sum_abs = 0.0;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            if constexpr (FUSED) h[t] = cnormal_at<T>(seed, STREAM_CHANNEL, unit, t);
            else h[t] = ldc(Hg + i * NT + t);
            const double mag = sqrt(double(h[t].re) * h[t].re + double(h[t].im) * h[t].im);
            sum_abs += mag;
            const double w = mag > 0.0 ? 1.0 / (mag * snt) : 0.0;
            W[t] = {T(h[t].re * w), T(-h[t].im * w)};
            if (!(mag > 0.0)) W[t] = {T(1.0 / snt), T(0)};          // angle(0) = 0
        }
        const T G = T(snt / sum_abs);
        for (int s = 0; s < S; ++s) {
            int a;
            if constexpr (FUSED)
                a = int(lane_of(rng_block(seed, STREAM_DATA, unit, uint64_t(s >> 2)), s & 3) >> (32 - m.bits));
            else
                a = idx[i * S + s];
            const cx<T> sym = map_symbol<T>(m, tab, a);
            cx<T> nz;
            if constexpr (FUSED) nz = cnormal_at<T>(seed, STREAM_NOISE, unit, uint64_t(s));
            else nz = ldc(noise + i * S + s);
            cx<T> y = sigma * nz;
#pragma unroll
            for (int t = 0; t < NT; ++t) cmac(y, h[t], W[t] * sym);
            const cx<T> z = G * y;
            const int e = demap_symbol<T>(m, tab, z);
            sym_err += (e != a);
            bit_err += __popc(e ^ a);
            if (idx_hat) idx_hat[i * S + s] = (uint8_t)e;
            if (dec_out) dec_out[i * S + s] = z;
        }
    }
    flush_counters(sym_err, bit_err, counters);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        atomicAdd(&counters[2], (unsigned long long)n * S);
        atomicAdd(&counters[3], (unsigned long long)n * S * m.bits);
    }
}

template <typename T, int N, int SCHEME>
static int launch_precoded(const Modem &m, const void *table, int S, double noise_var, double fnv, uint64_t seed,
                           uint64_t first, int64_t n, const uint8_t *idx, const void *H, const void *noise,
                           uint8_t *idx_hat, void *dec, int64_t *counters, cudaStream_t st) {
    const int grid = grid_units(n, kPT);
    auto go = [&](auto kern) {
        kern<<<grid, kPT, 0, st>>>(m, (const cx<T> *)table, S, T(sqrt(noise_var)), fnv, seed, first, (long long)n,
                                   idx, (const cx<T> *)H, (const cx<T> *)noise, idx_hat, (cx<T> *)dec,
                                   (unsigned long long *)counters);
    };
    if (!idx) go(precoded_kernel<T, true, N, SCHEME>); else go(precoded_kernel<T, false, N, SCHEME>);
    B200_CHECK_LAUNCH("precoded_kernel");
    return B200PHY_OK;
}

template <typename T, int NT>
static int launch_mrt(const Modem &m, const void *table, int S, double noise_var, uint64_t seed, uint64_t first,
                      int64_t n, const uint8_t *idx, const void *H, const void *noise, uint8_t *idx_hat, void *dec,
                      int64_t *counters, cudaStream_t st) {
    const int grid = grid_units(n, 256);
    auto go = [&](auto kern) {
        kern<<<grid, 256, 0, st>>>(m, (const cx<T> *)table, S, T(sqrt(noise_var)), seed, first, (long long)n, idx,
                                   (const cx<T> *)H, (const cx<T> *)noise, idx_hat, (cx<T> *)dec,
                                   (unsigned long long *)counters);
    };
    if (!idx) go(mrt_kernel<T, true, NT>); else go(mrt_kernel<T, false, NT>);
    B200_CHECK_LAUNCH("mrt_kernel");
    return B200PHY_OK;
}

template <typename T>
static int launch_scheme(int scheme, const Modem &m, const void *table, int Nt, int S, double noise_var, double fnv,
                         uint64_t seed, uint64_t first, int64_t n, const uint8_t *idx, const void *H,
                         const void *noise, uint8_t *idx_hat, void *dec, int64_t *counters, cudaStream_t st) {
#define B200_PRE(N_, SC_) launch_precoded<T, N_, SC_>(m, table, S, noise_var, fnv, seed, first, n, idx, H, noise, idx_hat, dec, counters, st)
#define B200_MRT(N_) launch_mrt<T, N_>(m, table, S, noise_var, seed, first, n, idx, H, noise, idx_hat, dec, counters, st)
    if (scheme == B200PHY_MIMO_SVD) {
        switch (Nt) { case 2: return B200_PRE(2, B200PHY_MIMO_SVD); case 3: return B200_PRE(3, B200PHY_MIMO_SVD); default: return B200_PRE(4, B200PHY_MIMO_SVD); }
    }
    if (scheme == B200PHY_MIMO_GMD) {
        switch (Nt) { case 2: return B200_PRE(2, B200PHY_MIMO_GMD); case 3: return B200_PRE(3, B200PHY_MIMO_GMD); default: return B200_PRE(4, B200PHY_MIMO_GMD); }
    }
    switch (Nt) { case 1: return B200_MRT(1); case 2: return B200_MRT(2); case 3: return B200_MRT(3); default: return B200_MRT(4); }
#undef B200_PRE
#undef B200_MRT
}

}  // namespace b200phy

using namespace b200phy;

extern "C" {

int b200phy_svd(const void *H, int64_t batch, int Nr, int Nt, void *U, double *S, void *V, void *stream) {
    if (batch < 0) { set_error("batch must be non-negative"); return B200PHY_ERR_INVALID; }
    if (batch == 0) return B200PHY_OK;
    if (!H || !U || !S || !V) { set_error("svd: NULL argument"); return B200PHY_ERR_INVALID; }
    return dispatch_shape<LaunchSvd>(Nr, Nt, H, batch, U, S, V, (cudaStream_t)stream);
}

int b200phy_gmd(const void *U, const double *S, const void *V, int64_t batch, int Nr, int Nt, void *Q, double *R,
                void *P, void *stream) {
    if (batch < 0) { set_error("batch must be non-negative"); return B200PHY_ERR_INVALID; }
    if (batch == 0) return B200PHY_OK;
    if (!U || !S || !V || !Q || !R || !P) { set_error("gmd: NULL argument"); return B200PHY_ERR_INVALID; }
    return dispatch_shape<LaunchGmd>(Nr, Nt, U, S, V, batch, Q, R, P, (cudaStream_t)stream);
}

int b200phy_mat_apply(int dtype, const void *A, int rows, int cols, const void *X, int64_t n, void *Y,
                      void *stream) {
    if (dtype != B200PHY_F32 && dtype != B200PHY_F64) { set_error("dtype must be B200PHY_F32 or B200PHY_F64"); return B200PHY_ERR_INVALID; }
    if (rows < 1 || cols < 1 || rows > 2 * B200PHY_MAX_ANT || cols > 2 * B200PHY_MAX_ANT) {
        set_error("mat_apply: matrix must be at most %d x %d (got %d x %d)", 2 * B200PHY_MAX_ANT, 2 * B200PHY_MAX_ANT, rows, cols);
        return B200PHY_ERR_UNSUPPORTED;
    }
    if (n < 0) { set_error("n must be non-negative"); return B200PHY_ERR_INVALID; }
    if (n == 0) return B200PHY_OK;
    if (!A || !X || !Y) { set_error("mat_apply: NULL argument"); return B200PHY_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = grid_units(n, 256);
    if (dtype == B200PHY_F32)
        mat_apply_kernel<float><<<grid, 256, 0, st>>>((const cx<float> *)A, rows, cols, (const cx<float> *)X, (long long)n, (cx<float> *)Y);
    else
        mat_apply_kernel<double><<<grid, 256, 0, st>>>((const cx<double> *)A, rows, cols, (const cx<double> *)X, (long long)n, (cx<double> *)Y);
    B200_CHECK_LAUNCH("mat_apply_kernel");
    return B200PHY_OK;
}

int b200phy_link_precoded(int dtype, const b200phy_modem *modem, int scheme, int Nr, int Nt, int S,
                          double noise_var, double filter_noise_var, uint64_t seed, uint64_t first_unit,
                          int64_t n_units, const uint8_t *idx, const void *H, const void *noise,
                          uint8_t *idx_hat, void *dec_out, int64_t *counters, void *stream) {
    Modem m;
    int e = check_modem(modem, &m);
    if (e) return e;
    if (dtype != B200PHY_F32 && dtype != B200PHY_F64) { set_error("dtype must be B200PHY_F32 or B200PHY_F64"); return B200PHY_ERR_INVALID; }
    if (n_units < 0) { set_error("n_units must be non-negative"); return B200PHY_ERR_INVALID; }
    if (!(noise_var >= 0.0) || !(filter_noise_var >= 0.0)) { set_error("Noise variance must be a non-negative value."); return B200PHY_ERR_INVALID; }
    if (!counters) { set_error("counters is NULL"); return B200PHY_ERR_INVALID; }
    if (S < 1) { set_error("S must be positive"); return B200PHY_ERR_INVALID; }
    if (scheme == B200PHY_MIMO_SVD || scheme == B200PHY_MIMO_GMD) {
        if (Nr != Nt || Nt < 2 || Nt > B200PHY_MAX_ANT) {
            set_error("SVD / GMD links need a square channel with 2 <= Nt <= %d (got %d x %d)", B200PHY_MAX_ANT, Nr, Nt);
            return B200PHY_ERR_UNSUPPORTED;
        }
    } else if (scheme == B200PHY_MIMO_MRT) {
        if (Nr != 1) { set_error("The MRT scheme is only defined for the scenario with a single receive antenna"); return B200PHY_ERR_INVALID; }
        if (Nt < 1 || Nt > B200PHY_MAX_ANT) { set_error("MRT: Nt=%d must be in [1, %d]", Nt, B200PHY_MAX_ANT); return B200PHY_ERR_UNSUPPORTED; }
    } else {
        set_error("unknown MIMO scheme %d", scheme);
        return B200PHY_ERR_INVALID;
    }
    const bool any = idx || H || noise, all = idx && H && noise;
    if (any && !all) { set_error("stream mode needs idx, H, noise together; fused mode needs all NULL"); return B200PHY_ERR_INVALID; }
    if (n_units == 0) return B200PHY_OK;
    cudaStream_t st = (cudaStream_t)stream;
    return dtype == B200PHY_F32
               ? launch_scheme<float>(scheme, m, modem->table, Nt, S, noise_var, filter_noise_var, seed, first_unit, n_units, idx, H, noise, idx_hat, dec_out, counters, st)
               : launch_scheme<double>(scheme, m, modem->table, Nt, S, noise_var, filter_noise_var, seed, first_unit, n_units, idx, H, noise, idx_hat, dec_out, counters, st);
}

}  // extern "C"

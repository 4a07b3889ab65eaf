// stages_mimo.cu — stage ops behind Blast.decode and Alamouti.encode/decode (mimo/mimo.py), batched
// over independent channel matrices: one thread per channel, matrices held in registers.
#include "common.cuh"
#include "linalg.cuh"

namespace b200phy {

static int blocks_for(long long n, int threads) {
    long long b = (n + threads - 1) / threads;
    return int(b < 1 ? 1 : (b < 148 * 8 ? b : 148 * 8));
}

// Blast.decode (mimo.py:642-660): out = (G y).reshape(-1, order='F'),
// G = sqrt(Nt) * (noise_var > 0 ? solve(H^H H + s2 I, H^H) : pinv(H))  (:590-607)
template <typename T, int NT>
__global__ void __launch_bounds__(128)
blast_decode_kernel(const cx<T> *__restrict__ Hg, const cx<T> *__restrict__ y, long long batch, int Nr,
                    int Tn, double fnv, cx<T> *__restrict__ out) {
    const double snt = sqrt(double(NT));
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < batch;
         i += (long long)gridDim.x * blockDim.x) {
        cx<T> H[B200PHY_MAX_ANT][NT];
        for (int r = 0; r < B200PHY_MAX_ANT; ++r)
#pragma unroll
            for (int t = 0; t < NT; ++t)
                if (r < Nr) H[r][t] = Hg[(i * Nr + r) * NT + t];
        HermSolver<NT> sol;
        sol.factor_from_channel(H, Nr, fnv);
        for (int c = 0; c < Tn; ++c) {
            cx<double> b[NT];
#pragma unroll
            for (int t = 0; t < NT; ++t) b[t] = {0.0, 0.0};
            for (int r = 0; r < B200PHY_MAX_ANT; ++r)
                if (r < Nr) {
                    const cx<double> yy = cvt<double>(y[(i * Nr + r) * Tn + c]);
#pragma unroll
                    for (int t = 0; t < NT; ++t) cmac_conj(b[t], cvt<double>(H[r][t]), yy);
                }
            sol.solve(b);
#pragma unroll
            for (int t = 0; t < NT; ++t)
                out[(i * Tn + c) * NT + t] = {T(b[t].re * snt), T(b[t].im * snt)};
        }
    }
}

// Alamouti.encode (mimo.py:1166-1214)
template <typename T>
__global__ void __launch_bounds__(256)
alamouti_encode_kernel(const cx<T> *__restrict__ s, long long batch, int Tn, cx<T> *__restrict__ x) {
    const T rs2 = T(0.70710678118654752440);
    const long long pairs = batch * (Tn / 2);
    for (long long it = blockIdx.x * (long long)blockDim.x + threadIdx.x; it < pairs;
         it += (long long)gridDim.x * blockDim.x) {
        const long long i = it / (Tn / 2);
        const int c = int(it % (Tn / 2));
        const cx<T> s0 = s[i * Tn + 2 * c], s1 = s[i * Tn + 2 * c + 1];
        cx<T> *x0 = x + (i * 2 + 0) * Tn + 2 * c, *x1 = x + (i * 2 + 1) * Tn + 2 * c;
        x0[0] = rs2 * s0;
        x0[1] = rs2 * mk<T>(-s1.re, s1.im);
        x1[0] = rs2 * s1;
        x1[1] = rs2 * conj(s0);
    }
}

// Alamouti.decode (mimo.py:1216-1287)
template <typename T>
__global__ void __launch_bounds__(256)
alamouti_decode_kernel(const cx<T> *__restrict__ Hg, const cx<T> *__restrict__ y, long long batch,
                       int Nr, int Tn, cx<T> *__restrict__ out) {
    const long long pairs = batch * (Tn / 2);
    for (long long it = blockIdx.x * (long long)blockDim.x + threadIdx.x; it < pairs;
         it += (long long)gridDim.x * blockDim.x) {
        const long long i = it / (Tn / 2);
        const int c = int(it % (Tn / 2));
        cx<T> d0 = {T(0), T(0)}, d1 = {T(0), T(0)};
        T fro = T(0);
        for (int r = 0; r < Nr; ++r) {
            const cx<T> h0 = Hg[(i * Nr + r) * 2], h1 = Hg[(i * Nr + r) * 2 + 1];
            const cx<T> y0 = y[(i * Nr + r) * Tn + 2 * c], y1 = y[(i * Nr + r) * Tn + 2 * c + 1];
            fro += norm2(h0) + norm2(h1);
            cmac_conj(d0, h0, y0); cmac(d0, h1, conj(y1));
            cmac_conj(d1, h1, y0); cmac(d1, mk<T>(-h0.re, -h0.im), conj(y1));
        }
        const T gain = T(1.41421356237309504880) / fro;
        out[i * Tn + 2 * c] = gain * d0;
        out[i * Tn + 2 * c + 1] = gain * d1;
    }
}

}  // namespace b200phy

using namespace b200phy;

extern "C" {

int b200phy_blast_decode(int dtype, const void *H, const void *y, int64_t batch, int Nr, int Nt, int T,
                         double filter_noise_var, void *out, void *stream) {
    if (Nr < 1 || Nr > B200PHY_MAX_ANT || Nt < 1 || Nt > B200PHY_MAX_ANT) { set_error("Blast: Nr=%d, Nt=%d must be in [1, %d]", Nr, Nt, B200PHY_MAX_ANT); return B200PHY_ERR_UNSUPPORTED; }
    if (!(filter_noise_var >= 0.0)) { set_error("Noise variance must be a non-negative value."); return B200PHY_ERR_INVALID; }
    if (filter_noise_var == 0.0 && Nt > Nr) { set_error("Blast ZF needs Nt <= Nr (got %dx%d)", Nr, Nt); return B200PHY_ERR_UNSUPPORTED; }
    if (batch <= 0 || T <= 0) return B200PHY_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = blocks_for(batch, 128);
#define B200_BD(TT, NT_) blast_decode_kernel<TT, NT_><<<grid, 128, 0, st>>>((const cx<TT> *)H, (const cx<TT> *)y, batch, Nr, T, filter_noise_var, (cx<TT> *)out)
    if (dtype == B200PHY_F32) {
        switch (Nt) { case 1: B200_BD(float, 1); break; case 2: B200_BD(float, 2); break; case 3: B200_BD(float, 3); break; default: B200_BD(float, 4); }
    } else {
        switch (Nt) { case 1: B200_BD(double, 1); break; case 2: B200_BD(double, 2); break; case 3: B200_BD(double, 3); break; default: B200_BD(double, 4); }
    }
#undef B200_BD
    B200_CHECK_LAUNCH("blast_decode_kernel");
    return B200PHY_OK;
}

int b200phy_alamouti_encode(int dtype, const void *s, int64_t batch, int T, void *x, void *stream) {
    if (T < 2 || (T & 1)) { set_error("Alamouti: number of symbols T=%d must be even", T); return B200PHY_ERR_INVALID; }
    if (batch <= 0) return B200PHY_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = blocks_for(batch * (T / 2), 256);
    if (dtype == B200PHY_F32) alamouti_encode_kernel<float><<<grid, 256, 0, st>>>((const cx<float> *)s, batch, T, (cx<float> *)x);
    else alamouti_encode_kernel<double><<<grid, 256, 0, st>>>((const cx<double> *)s, batch, T, (cx<double> *)x);
    B200_CHECK_LAUNCH("alamouti_encode_kernel");
    return B200PHY_OK;
}

int b200phy_alamouti_decode(int dtype, const void *H, const void *y, int64_t batch, int Nr, int T,
                            void *out, void *stream) {
    if (T < 2 || (T & 1)) { set_error("Alamouti: number of symbols T=%d must be even", T); return B200PHY_ERR_INVALID; }
    if (Nr < 1) { set_error("Nr must be positive"); return B200PHY_ERR_INVALID; }
    if (batch <= 0) return B200PHY_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = blocks_for(batch * (T / 2), 256);
    if (dtype == B200PHY_F32) alamouti_decode_kernel<float><<<grid, 256, 0, st>>>((const cx<float> *)H, (const cx<float> *)y, batch, Nr, T, (cx<float> *)out);
    else alamouti_decode_kernel<double><<<grid, 256, 0, st>>>((const cx<double> *)H, (const cx<double> *)y, batch, Nr, T, (cx<double> *)out);
    B200_CHECK_LAUNCH("alamouti_decode_kernel");
    return B200PHY_OK;
}

}  // extern "C"

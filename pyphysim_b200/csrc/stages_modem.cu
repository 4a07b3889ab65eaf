// stages_modem.cu — stage ops behind Modulator.modulate / demodulate, count_bit_errors, randn_c.
// Elementwise, HBM-bound: one element per thread iteration, coalesced, grid = SMs x 8.
#include "common.cuh"
#include "rng.cuh"

namespace b200phy {

constexpr int kT = 256;

static int grid_for(long long n) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long b = (n + kT - 1) / kT;
    const long long cap = (long long)sms * 8;
    return int(b < 1 ? 1 : (b < cap ? b : cap));
}

template <typename T>
__global__ void __launch_bounds__(kT)
map_kernel(Modem m, const cx<T> *__restrict__ tab_g, const long long *__restrict__ idx, long long n,
           cx<T> *__restrict__ out, int *err_flag) {
    __shared__ cx<T> tab[256];
    if (m.kind != B200PHY_MODEM_BPSK)
        for (int k = threadIdx.x; k < m.M; k += blockDim.x) tab[k] = tab_g[k];
    __syncthreads();
    bool bad = false;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        long long a = idx[i];
        if (m.kind == B200PHY_MODEM_BPSK) {
            // BPSK.modulate: 1 - 2*idx, ValueError if any idx > 1 (fundamental.py:626-630)
            if (a > 1) bad = true;
            out[i] = {T(1 - 2 * a), T(0)};
        } else {
            if (a < 0) a += m.M;                     // NumPy negative-index wrap (fundamental.py:193-194)
            if (a < 0 || a >= m.M) { bad = true; a = 0; }
            out[i] = tab[a];
        }
    }
    if (bad) atomicExch(err_flag, 1);
}

template <typename T>
__global__ void __launch_bounds__(kT)
demap_kernel(Modem m, const cx<T> *__restrict__ tab_g, const cx<T> *__restrict__ r, long long n,
             long long *__restrict__ out) {
    __shared__ cx<T> tab[256];
    if (m.kind != B200PHY_MODEM_BPSK)
        for (int k = threadIdx.x; k < m.M; k += blockDim.x) tab[k] = tab_g[k];
    __syncthreads();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        out[i] = demap_symbol<T>(m, tab, r[i]);
}

__global__ void __launch_bounds__(kT)
count_errors_kernel(const long long *__restrict__ a, const long long *__restrict__ b, long long n,
                    unsigned long long *out) {
    unsigned long long se = 0, be = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const long long x = a[i], y = b[i];
        se += (x != y);
        be += __popcll((unsigned long long)(x ^ y));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        se += __shfl_xor_sync(0xffffffffu, se, o);
        be += __shfl_xor_sync(0xffffffffu, be, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (se) atomicAdd(&out[0], se);
        if (be) atomicAdd(&out[1], be);
    }
}

__global__ void __launch_bounds__(kT)
count_bits_kernel(const long long *__restrict__ a, long long n, long long *__restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const long long v = a[i];
        out[i] = v > 0 ? __popcll((unsigned long long)v) : 0;     // the reference loops while n > 0
    }
}

template <typename T>
__global__ void __launch_bounds__(kT)
awgn_kernel(cx<T> *x, long long n, T sigma, uint64_t seed, uint32_t stream_id, uint64_t unit,
            uint64_t first) {
    // thread handles the pair of normals of one Philox block
    const long long pairs = (n + 1) / 2 + 1;
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < pairs;
         g += (long long)gridDim.x * blockDim.x) {
        const uint64_t slot = (first >> 1) + uint64_t(g);
        const uint4 b = rng_block(seed, stream_id, unit, slot);
        const long long j0 = (long long)(2 * slot) - (long long)first;    // element of normal 2*slot
        if (j0 >= 0 && j0 < n) { const cx<T> c = cnormal<T>(b.x, b.y); x[j0].re += sigma * c.re; x[j0].im += sigma * c.im; }
        if (j0 + 1 >= 0 && j0 + 1 < n) { const cx<T> c = cnormal<T>(b.z, b.w); x[j0 + 1].re += sigma * c.re; x[j0 + 1].im += sigma * c.im; }
    }
}

}  // namespace b200phy

using namespace b200phy;

extern "C" {

int b200phy_map(int dtype, const b200phy_modem *modem, const int64_t *idx, int64_t n, void *out,
                int32_t *err_flag, void *stream) {
    Modem m;
    int e = check_modem(modem, &m);
    if (e) return e;
    if (!err_flag) { set_error("err_flag is NULL"); return B200PHY_ERR_INVALID; }
    if (n <= 0) return B200PHY_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == B200PHY_F32)
        map_kernel<float><<<grid_for(n), kT, 0, st>>>(m, (const cx<float> *)modem->table, (const long long *)idx, n, (cx<float> *)out, err_flag);
    else
        map_kernel<double><<<grid_for(n), kT, 0, st>>>(m, (const cx<double> *)modem->table, (const long long *)idx, n, (cx<double> *)out, err_flag);
    B200_CHECK_LAUNCH("map_kernel");
    return B200PHY_OK;
}

int b200phy_demap(int dtype, const b200phy_modem *modem, const void *r, int64_t n, int64_t *idx_hat,
                  void *stream) {
    Modem m;
    int e = check_modem(modem, &m);
    if (e) return e;
    if (n <= 0) return B200PHY_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == B200PHY_F32)
        demap_kernel<float><<<grid_for(n), kT, 0, st>>>(m, (const cx<float> *)modem->table, (const cx<float> *)r, n, (long long *)idx_hat);
    else
        demap_kernel<double><<<grid_for(n), kT, 0, st>>>(m, (const cx<double> *)modem->table, (const cx<double> *)r, n, (long long *)idx_hat);
    B200_CHECK_LAUNCH("demap_kernel");
    return B200PHY_OK;
}

int b200phy_count_errors(const int64_t *a, const int64_t *b, int64_t n, int64_t *out, void *stream) {
    if (!out) { set_error("out is NULL"); return B200PHY_ERR_INVALID; }
    if (n <= 0) return B200PHY_OK;
    count_errors_kernel<<<grid_for(n), kT, 0, (cudaStream_t)stream>>>((const long long *)a, (const long long *)b, n, (unsigned long long *)out);
    B200_CHECK_LAUNCH("count_errors_kernel");
    return B200PHY_OK;
}

int b200phy_count_bits(const int64_t *a, int64_t n, int64_t *out, void *stream) {
    if (n <= 0) return B200PHY_OK;
    count_bits_kernel<<<grid_for(n), kT, 0, (cudaStream_t)stream>>>((const long long *)a, n, (long long *)out);
    B200_CHECK_LAUNCH("count_bits_kernel");
    return B200PHY_OK;
}

int b200phy_awgn(int dtype, void *x, int64_t n, double noise_var, uint64_t seed, uint32_t stream_id,
                 uint64_t unit, uint64_t first, void *stream) {
    if (!(noise_var >= 0.0)) { set_error("Noise variance must be a non-negative value."); return B200PHY_ERR_INVALID; }
    if (n <= 0) return B200PHY_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == B200PHY_F32)
        awgn_kernel<float><<<grid_for((n + 1) / 2 + 1), kT, 0, st>>>((cx<float> *)x, n, float(sqrt(noise_var)), seed, stream_id, unit, first);
    else
        awgn_kernel<double><<<grid_for((n + 1) / 2 + 1), kT, 0, st>>>((cx<double> *)x, n, sqrt(noise_var), seed, stream_id, unit, first);
    B200_CHECK_LAUNCH("awgn_kernel");
    return B200PHY_OK;
}

}  // extern "C"

// stages_ofdm.cu — stage ops behind OFDM.modulate / OFDM.demodulate (modulators/ofdm.py:394-466):
// one CTA per OFDM symbol, subcarrier scatter/gather fused with the shared-memory Stockham FFT and
// with cyclic-prefix insertion/removal.  HBM traffic = the input and output samples, once.
#include "ofdm_tdl.cuh"

namespace b200phy {

template <typename T>
__global__ void __launch_bounds__(kOT)
ofdm_mod_kernel(const cx<T> *__restrict__ x, cx<T> *__restrict__ out, long long n_symbols_total,
                int fft, int lg, int cp, int used, T scale) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cx<T> *tw = (cx<T> *)smem_raw, *a = tw + fft, *b = a + fft;
    const int half = used / 2;
    for (int i = threadIdx.x; i < fft; i += kOT) {
        double s, c;
        sincospi(-2.0 * double(i) / double(fft), &s, &c);
        tw[i] = {T(c), T(s)};
    }
    for (long long sy = blockIdx.x; sy < n_symbols_total; sy += gridDim.x) {
        __syncthreads();
        const cx<T> *src = x + sy * used;
        for (int k = threadIdx.x; k < fft; k += kOT) {
            const int q = pos_of(k, fft, used, half);
            a[k] = q >= 0 ? scale * src[q] : mk<T>(T(0), T(0));
        }
        cx<T> *res = fft_stockham<T, true>(a, b, tw, fft, lg);
        cx<T> *dst = out + sy * (fft + cp);
        for (int i = threadIdx.x; i < fft + cp; i += kOT)
            dst[i] = res[i < cp ? fft - cp + i : i - cp];       // _add_CP (ofdm.py:320-341)
    }
}

template <typename T>
__global__ void __launch_bounds__(kOT)
ofdm_demod_kernel(const cx<T> *__restrict__ r, cx<T> *__restrict__ y, long long n_symbols_total,
                  int fft, int lg, int cp, int used, T scale) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cx<T> *tw = (cx<T> *)smem_raw, *a = tw + fft, *b = a + fft;
    const int half = used / 2;
    for (int i = threadIdx.x; i < fft; i += kOT) {
        double s, c;
        sincospi(-2.0 * double(i) / double(fft), &s, &c);
        tw[i] = {T(c), T(s)};
    }
    for (long long sy = blockIdx.x; sy < n_symbols_total; sy += gridDim.x) {
        __syncthreads();
        const cx<T> *src = r + sy * (fft + cp) + cp;            // _remove_CP (ofdm.py:343-368)
        for (int i = threadIdx.x; i < fft; i += kOT) a[i] = src[i];
        cx<T> *res = fft_stockham<T, false>(a, b, tw, fft, lg);
        cx<T> *dst = y + sy * used;
        for (int q = threadIdx.x; q < used; q += kOT) dst[q] = scale * res[bin_of(q, fft, used, half)];
    }
}

static int check_ofdm(int fft, int cp, int used) {
    if (cp < 0 || cp > fft) { set_error("cp_size must be nonnegative and cannot be greater than fft_size"); return B200PHY_ERR_INVALID; }
    if (used > fft) { set_error("Number of used subcarriers cannot be greater than the fft_size"); return B200PHY_ERR_INVALID; }
    if ((used % 2) != 0 || used < 2) { set_error("Number of used subcarriers must be a multiple of 2"); return B200PHY_ERR_INVALID; }
    if (fft < 8 || fft > 4096 || (fft & (fft - 1))) { set_error("fft_size=%d must be a power of two in [8, 4096]", fft); return B200PHY_ERR_UNSUPPORTED; }
    return B200PHY_OK;
}

template <typename T, bool MOD>
static int launch_ofdm(const void *in, void *out, int64_t batch, int n_sym, int fft, int cp, int used,
                       cudaStream_t st) {
    const long long total = (long long)batch * n_sym;
    if (total <= 0) return B200PHY_OK;
    const size_t smem = sizeof(cx<T>) * 3 * fft;
    const double ps = double(fft) * double(fft) / (double(used) + double(cp));
    auto kern = MOD ? ofdm_mod_kernel<T> : ofdm_demod_kernel<T>;
    int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)), "cudaFuncSetAttribute(ofdm)");
    if (e) return e;
    const int grid = int(total < 148 * 4 ? total : 148 * 4);
    const T scale = MOD ? T(sqrt(ps) / double(fft)) : T(1.0 / sqrt(ps));
    kern<<<grid, kOT, smem, st>>>((const cx<T> *)in, (cx<T> *)out, total, fft, ilog2(fft), cp, used, scale);
    B200_CHECK_LAUNCH("ofdm_mod/demod_kernel");
    return B200PHY_OK;
}

}  // namespace b200phy

using namespace b200phy;

extern "C" {

int b200phy_ofdm_mod(int dtype, const void *x, void *out, int64_t batch, int n_sym, int fft, int cp,
                     int used, void *stream) {
    int e = check_ofdm(fft, cp, used);
    if (e) return e;
    return dtype == B200PHY_F32 ? launch_ofdm<float, true>(x, out, batch, n_sym, fft, cp, used, (cudaStream_t)stream)
                                : launch_ofdm<double, true>(x, out, batch, n_sym, fft, cp, used, (cudaStream_t)stream);
}

int b200phy_ofdm_demod(int dtype, const void *r, void *y, int64_t batch, int n_sym, int fft, int cp,
                       int used, void *stream) {
    int e = check_ofdm(fft, cp, used);
    if (e) return e;
    return dtype == B200PHY_F32 ? launch_ofdm<float, false>(r, y, batch, n_sym, fft, cp, used, (cudaStream_t)stream)
                                : launch_ofdm<double, false>(r, y, batch, n_sym, fft, cp, used, (cudaStream_t)stream);
}

}  // extern "C"

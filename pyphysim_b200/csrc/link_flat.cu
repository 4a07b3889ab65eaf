// link_flat.cu — fused flat-fading links: one realization per thread (SISO: four per thread),
// stream mode (draws are tensors in HBM) or fused mode (in-kernel Philox).
//
//   siso_flat  notebooks/Transmission_with_Rayleigh_and_AWGN_channels.ipynb cell 8
//              (AWGN: apps/awgn_modulators/simulate_psk.py:51-115)
//   alamouti   apps/mimo/simulate_mimo.py:68-142 + mimo/mimo.py:1166-1287
//   blast      apps/mimo/simulate_mimo.py:68-142 + mimo/mimo.py:590-660
//
// These links are HBM-bound in stream mode (18 B, 68 B, ... per realization), so the
// kernels use 16-byte loads and keep everything between load and the 1-byte store in registers.
#include "common.cuh"
#include "linalg.cuh"
#include "rng.cuh"

namespace b200phy {

constexpr int kThreads = 256;

// ---------------------------------------------------------------- vector load helpers
template <typename T> struct Vec16;
template <> struct Vec16<float> { using type = float4; static constexpr int n = 2; };
template <> struct Vec16<double> { using type = double2; static constexpr int n = 1; };

// load `count` complex values starting at a 16-byte aligned address (count*sizeof(cx<T>) % 16 == 0)
template <typename T, int MAXC>
__device__ __forceinline__ void load_cx(const cx<T> *__restrict__ p, int count, cx<T> (&out)[MAXC]) {
    if constexpr (sizeof(T) == 4) {
        const float4 *q = reinterpret_cast<const float4 *>(p);
#pragma unroll
        for (int i = 0; i < MAXC / 2; ++i)
            if (2 * i < count) {
                const float4 v = __ldg(q + i);
                out[2 * i] = {v.x, v.y};
                out[2 * i + 1] = {v.z, v.w};
            }
    } else {
        const double2 *q = reinterpret_cast<const double2 *>(p);
#pragma unroll
        for (int i = 0; i < MAXC; ++i)
            if (i < count) {
                const double2 v = __ldg(q + i);
                out[i] = {v.x, v.y};
            }
    }
}

template <typename T> __device__ __forceinline__ cx<T> load1(const cx<T> *__restrict__ p) {
    if constexpr (sizeof(T) == 4) {
        const float2 v = __ldg(reinterpret_cast<const float2 *>(p));
        return {v.x, v.y};
    } else {
        const double2 v = __ldg(reinterpret_cast<const double2 *>(p));
        return {v.x, v.y};
    }
}

template <typename T> __device__ __forceinline__ void stage_table(const Modem &m, const cx<T> *tab_g, cx<T> *tab_s) {
    if (m.kind != B200PHY_MODEM_BPSK)
        for (int k = threadIdx.x; k < m.M; k += blockDim.x) tab_s[k] = tab_g[k];
    __syncthreads();
}

// ================================================================= SISO flat
// One realization: r = h*x + sigma*n; r /= h; demap; count.  (notebook Rayleigh cell 8)
template <typename T, bool RAYLEIGH, bool DEC>
__device__ __forceinline__ int siso_one(const Modem &m, const cx<T> *tab, int a, cx<T> hh, cx<T> nn, T sigma,
                                        unsigned &sym_err, unsigned &bit_err, cx<T> *dec_out, long long i) {
    const cx<T> x = map_symbol<T>(m, tab, a);
    cx<T> r;
    if (RAYLEIGH) {
        r = hh * x + sigma * nn;      // received_data = h * x + n
        r = cdiv(r, hh);              // received_data /= h
    } else {
        r = x + sigma * nn;
    }
    const int d = demap_symbol<T>(m, tab, r);
    sym_err += (d != a);
    bit_err += __popc(d ^ a);
    if (DEC) dec_out[i] = r;
    return d;
}

template <typename T, bool FUSED, bool RAYLEIGH, bool QAM, bool DEC>
__global__ void __launch_bounds__(kThreads, sizeof(T) == 4 ? 5 : 2)
siso_flat_kernel(Modem m_in, const cx<T> *__restrict__ tab_g, T sigma, const __grid_constant__ PhiloxKey seed,
                 uint64_t first_unit, long long n, const uint8_t *__restrict__ idx,
                 const cx<T> *__restrict__ h, const cx<T> *__restrict__ noise,
                 uint8_t *__restrict__ idx_hat, cx<T> *__restrict__ dec_out,
                 unsigned long long *counters) {
    __shared__ cx<T> tab[256];
    Modem m = m_in;
    if (QAM) m.kind = B200PHY_MODEM_QAM;      // compile-time kind: the slicer is inlined without branches
    stage_table(m, tab_g, tab);
    unsigned sym_err = 0, bit_err = 0;
    const long long n4 = n / 4;               // full groups of 4: branch-free body
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < n4;
         g += (long long)gridDim.x * blockDim.x) {
        const long long i0 = g * 4;
        int a[4];
        cx<T> hh[4], nn[4];
        if constexpr (FUSED) {
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const uint64_t unit = first_unit + uint64_t(i0 + v);
                a[v] = int(rng_block(seed, STREAM_DATA, unit, 0).x >> (32 - m.bits));
                const uint4 bn = rng_block(seed, STREAM_NOISE, unit, 0);
                nn[v] = cnormal<T>(bn.x, bn.y);
                if (RAYLEIGH) {
                    const uint4 bh = rng_block(seed, STREAM_CHANNEL, unit, 0);
                    hh[v] = cnormal<T>(bh.x, bh.y);
                }
            }
        } else {
            const uchar4 b = __ldg(reinterpret_cast<const uchar4 *>(idx + i0));
            a[0] = b.x; a[1] = b.y; a[2] = b.z; a[3] = b.w;
            load_cx<T, 4>(noise + i0, 4, nn);
            if (RAYLEIGH) load_cx<T, 4>(h + i0, 4, hh);
        }
        int d[4];
#pragma unroll
        for (int v = 0; v < 4; ++v)
            d[v] = siso_one<T, RAYLEIGH, DEC>(m, tab, a[v], hh[v], nn[v], sigma, sym_err, bit_err, dec_out, i0 + v);
        if (idx_hat)
            *reinterpret_cast<uchar4 *>(idx_hat + i0) =
                make_uchar4((unsigned char)d[0], (unsigned char)d[1], (unsigned char)d[2], (unsigned char)d[3]);
    }
    // ragged tail (n % 4 realizations): one thread each
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const long long i = n4 * 4 + threadIdx.x;
        int a;
        cx<T> hh = {T(1), T(0)}, nn;
        if constexpr (FUSED) {
            const uint64_t unit = first_unit + uint64_t(i);
            a = int(rng_block(seed, STREAM_DATA, unit, 0).x >> (32 - m.bits));
            const uint4 bn = rng_block(seed, STREAM_NOISE, unit, 0);
            nn = cnormal<T>(bn.x, bn.y);
            if (RAYLEIGH) { const uint4 bh = rng_block(seed, STREAM_CHANNEL, unit, 0); hh = cnormal<T>(bh.x, bh.y); }
        } else {
            a = idx[i];
            nn = load1(noise + i);
            if (RAYLEIGH) hh = load1(h + i);
        }
        const int d = siso_one<T, RAYLEIGH, DEC>(m, tab, a, hh, nn, sigma, sym_err, bit_err, dec_out, i);
        if (idx_hat) idx_hat[i] = (uint8_t)d;
    }
    flush_counters(sym_err, bit_err, counters);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        atomicAdd(&counters[2], (unsigned long long)n);
        atomicAdd(&counters[3], (unsigned long long)n * m.bits);
    }
}

// ================================================================= Alamouti flat
// One realization per thread: H[Nr][2], S symbols (S/2 codewords), noise[Nr][S].
template <typename T, bool FUSED, int NR, bool DEC, bool QPSK>
__global__ void __launch_bounds__(kThreads, sizeof(T) == 4 ? 5 : 2)
alamouti_kernel(Modem m_in, const cx<T> *__restrict__ tab_g, int S, T sigma, const __grid_constant__ PhiloxKey seed,
                uint64_t first_unit, long long n, const uint8_t *__restrict__ idx,
                const cx<T> *__restrict__ Hg, const cx<T> *__restrict__ noise,
                uint8_t *__restrict__ idx_hat, cx<T> *__restrict__ dec_out,
                unsigned long long *counters) {
    __shared__ cx<T> tab[256];
    Modem m = m_in;
    if (QPSK) m.kind = B200PHY_MODEM_QPSK;    // compile-time kind: the quadrant slicer is inlined without branches
    stage_table(m, tab_g, tab);
    unsigned sym_err = 0, bit_err = 0;
    const T rs2 = T(0.70710678118654752440);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const uint64_t unit = first_unit + uint64_t(i);
        cx<T> H[2 * NR];                     // H[r][t] at 2r+t
        if constexpr (FUSED) {
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                const uint4 b = rng_block(seed, STREAM_CHANNEL, unit, r);
                H[2 * r] = cnormal<T>(b.x, b.y);
                H[2 * r + 1] = cnormal<T>(b.z, b.w);
            }
        } else {
            load_cx<T, 2 * NR>(Hg + i * NR * 2, NR * 2, H);
        }
        T fro = T(0);
#pragma unroll
        for (int r = 0; r < NR; ++r) fro += norm2(H[2 * r]) + norm2(H[2 * r + 1]);
        const T gain = T(1.41421356237309504880) / fro;    // (1/||H||_F^2) * sqrt(2)
        for (int c = 0; c < S / 2; ++c) {
            int a0, a1;
            if constexpr (FUSED) {
                const uint4 b = rng_block(seed, STREAM_DATA, unit, uint64_t(c >> 1));
                a0 = int(((c & 1) ? b.z : b.x) >> (32 - m.bits));
                a1 = int(((c & 1) ? b.w : b.y) >> (32 - m.bits));
            } else {
                const uchar2 b = *reinterpret_cast<const uchar2 *>(idx + i * S + 2 * c);
                a0 = b.x; a1 = b.y;
            }
            const cx<T> s0 = map_symbol<T>(m, tab, a0), s1 = map_symbol<T>(m, tab, a1);
            // encode: [[s0, -s1*], [s1, s0*]] / sqrt(2)   (mimo.py:1193-1214)
            const cx<T> x00 = rs2 * s0, x01 = rs2 * mk<T>(-s1.re, s1.im);
            const cx<T> x10 = rs2 * s1, x11 = rs2 * conj(s0);
            cx<T> d0 = {T(0), T(0)}, d1 = {T(0), T(0)};
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                    cx<T> n0, n1;
                    if constexpr (FUSED) {
                        // noise row r holds S normals; every row starts on a fresh slot
                        const uint4 b = rng_block(seed, STREAM_NOISE, unit, uint64_t(r) * (S / 2) + c);
                        n0 = cnormal<T>(b.x, b.y);
                        n1 = cnormal<T>(b.z, b.w);
                    } else {
                        const cx<T> *p = noise + (i * NR + r) * S + 2 * c;
                        cx<T> t2[2];
                        load_cx<T, 2>(p, 2, t2);
                        n0 = t2[0]; n1 = t2[1];
                    }
                    const cx<T> h0 = H[2 * r], h1 = H[2 * r + 1];
                    cx<T> y0 = sigma * n0, y1 = sigma * n1;     // y = H x + n
                    cmac(y0, h0, x00); cmac(y0, h1, x10);
                    cmac(y1, h0, x01); cmac(y1, h1, x11);
                    // decode (mimo.py:1258-1264): d0 += h0* y0 + h1 y1*; d1 += h1* y0 - h0 y1*
                    cmac_conj(d0, h0, y0); cmac(d0, h1, conj(y1));
                    cmac_conj(d1, h1, y0); cmac(d1, mk<T>(-h0.re, -h0.im), conj(y1));
                }
            d0 = gain * d0;
            d1 = gain * d1;
            const int e0 = demap_symbol<T>(m, tab, d0), e1 = demap_symbol<T>(m, tab, d1);
            sym_err += (e0 != a0) + (e1 != a1);
            bit_err += __popc(e0 ^ a0) + __popc(e1 ^ a1);
            if (idx_hat)
                *reinterpret_cast<uchar2 *>(idx_hat + i * S + 2 * c) =
                    make_uchar2((unsigned char)e0, (unsigned char)e1);
            if (DEC) { dec_out[i * S + 2 * c] = d0; dec_out[i * S + 2 * c + 1] = d1; }
        }
    }
    flush_counters(sym_err, bit_err, counters);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        atomicAdd(&counters[2], (unsigned long long)n * S);
        atomicAdd(&counters[3], (unsigned long long)n * S * m.bits);
    }
}


// ---- C4 shape (float, Nr = 2, one codeword of S = 2 symbols per realization) with everything the generic kernel
// decides at run time fixed at compile time: no codeword loop, no per-antenna switch, four unconditional LDG.128
// and one 16-bit load per realization.  The generic kernel spends 295 instructions per realization (ncu: 60 %
// issue-active, FMA pipe only 29 % — most of it is not arithmetic) and is co-limited by instruction issue at 0.82
// of the HBM peak.  QPSK without a sample output also skips the 1 / ||H||_F^2 scaling: the quadrant slicer only
// looks at signs and the gain is positive.
// (A packed-FP32 variant with lane = rx antenna was built first: the operands of a lane pair come from two
// different 128-bit loads, so every pair costs two register moves and the instruction count did not drop.)
template <bool FUSED, bool DEC, bool QPSK>
__global__ void __launch_bounds__(kThreads, 5)
alamouti22_kernel(Modem m_in, const cx<float> *__restrict__ tab_g, float sigma, const __grid_constant__ PhiloxKey seed, uint64_t first_unit,
                  long long n, const uint8_t *__restrict__ idx, const float4 *__restrict__ Hg,
                  const float4 *__restrict__ noise, uint8_t *__restrict__ idx_hat, cx<float> *__restrict__ dec_out,
                  unsigned long long *counters) {
    __shared__ cx<float> tab[256];
    Modem m = m_in;
    if (QPSK) m.kind = B200PHY_MODEM_QPSK;
    stage_table(m, tab_g, tab);
    unsigned sym_err = 0, bit_err = 0;
    const float rs2 = 0.70710678118654752440f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        cx<float> h[2][2], w[2][2];          // h[r][t]; w[r][time]
        int a0, a1;
        if constexpr (FUSED) {
            const uint64_t unit = first_unit + uint64_t(i);
            const uint4 bd = rng_block(seed, STREAM_DATA, unit, 0);
            a0 = int(bd.x >> (32 - m.bits));
            a1 = int(bd.y >> (32 - m.bits));
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const uint4 c = rng_block(seed, STREAM_CHANNEL, unit, r), e = rng_block(seed, STREAM_NOISE, unit, r);
                h[r][0] = cnormal<float>(c.x, c.y); h[r][1] = cnormal<float>(c.z, c.w);
                w[r][0] = cnormal<float>(e.x, e.y); w[r][1] = cnormal<float>(e.z, e.w);
            }
        } else {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const float4 hv = __ldg(Hg + 2 * i + r), nv = __ldg(noise + 2 * i + r);
                h[r][0] = {hv.x, hv.y}; h[r][1] = {hv.z, hv.w};
                w[r][0] = {nv.x, nv.y}; w[r][1] = {nv.z, nv.w};
            }
            const uchar2 b = __ldg(reinterpret_cast<const uchar2 *>(idx) + i);
            a0 = b.x;
            a1 = b.y;
        }
        const cx<float> s0 = map_symbol<float>(m, tab, a0), s1 = map_symbol<float>(m, tab, a1);
        // encode: [[s0, -s1*], [s1, s0*]] / sqrt(2)   (mimo.py:1193-1214)
        const cx<float> x00 = rs2 * s0, x01 = rs2 * mk<float>(-s1.re, s1.im);
        const cx<float> x10 = rs2 * s1, x11 = rs2 * conj(s0);
        cx<float> d0 = {0.f, 0.f}, d1 = {0.f, 0.f};
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const cx<float> h0 = h[r][0], h1 = h[r][1];
            cx<float> y0 = sigma * w[r][0], y1 = sigma * w[r][1];           // y = H x + n (simulate_mimo.py:96-98)
            cmac(y0, h0, x00); cmac(y0, h1, x10);
            cmac(y1, h0, x01); cmac(y1, h1, x11);
            // decode (mimo.py:1258-1264): d0 += h0* y0 + h1 y1*; d1 += h1* y0 - h0 y1*
            cmac_conj(d0, h0, y0); cmac(d0, h1, conj(y1));
            cmac_conj(d1, h1, y0); cmac(d1, mk<float>(-h0.re, -h0.im), conj(y1));
        }
        if constexpr (DEC || !QPSK) {
            const float fro = norm2(h[0][0]) + norm2(h[0][1]) + norm2(h[1][0]) + norm2(h[1][1]);
            const float gain = 1.41421356237309504880f / fro;                  // sqrt(2) / ||H||_F^2
            d0 = gain * d0;
            d1 = gain * d1;
        }
        const int e0 = demap_symbol<float>(m, tab, d0), e1 = demap_symbol<float>(m, tab, d1);
        sym_err += (e0 != a0) + (e1 != a1);
        bit_err += __popc(e0 ^ a0) + __popc(e1 ^ a1);
        if (idx_hat) reinterpret_cast<uchar2 *>(idx_hat)[i] = make_uchar2((unsigned char)e0, (unsigned char)e1);
        if constexpr (DEC) { dec_out[2 * i] = d0; dec_out[2 * i + 1] = d1; }
    }
    flush_counters(sym_err, bit_err, counters);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        atomicAdd(&counters[2], (unsigned long long)n * 2);
        atomicAdd(&counters[3], (unsigned long long)n * 2 * m.bits);
    }
}

// ================================================================= Blast flat (ZF / MMSE)
template <typename T, bool FUSED, int NT>
__global__ void __launch_bounds__(kThreads)
blast_kernel(Modem m, const cx<T> *__restrict__ tab_g, int Nr, int S, T sigma, double fnv,
             const __grid_constant__ PhiloxKey seed, uint64_t first_unit, long long n, const uint8_t *__restrict__ idx,
             const cx<T> *__restrict__ Hg, const cx<T> *__restrict__ noise,
             uint8_t *__restrict__ idx_hat, cx<T> *__restrict__ dec_out,
             unsigned long long *counters) {
    __shared__ cx<T> tab[256];
    stage_table(m, tab_g, tab);
    unsigned sym_err = 0, bit_err = 0;
    const T rsnt = T(1.0 / sqrt(double(NT)));
    const double snt = sqrt(double(NT));
    const int row = 2 * ((S + 1) / 2);      // normals per noise row in the Philox layout
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const uint64_t unit = first_unit + uint64_t(i);
        cx<T> H[B200PHY_MAX_ANT][NT];
#pragma unroll
        for (int r = 0; r < B200PHY_MAX_ANT; ++r)
#pragma unroll
            for (int t = 0; t < NT; ++t)
                if (r < Nr) {
                    if constexpr (FUSED) H[r][t] = cnormal_at<T>(seed, STREAM_CHANNEL, unit, r * NT + t);
                    else H[r][t] = load1(Hg + (i * Nr + r) * NT + t);
                }
        HermSolver<NT> sol;
        sol.factor_from_channel(H, Nr, fnv);
        for (int s = 0; s < S; ++s) {
            int a[NT];
            cx<T> x[NT];
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                const int p = s * NT + t;          // symbol p -> antenna p % NT (order='F')
                if constexpr (FUSED)
                    a[t] = int(lane_of(rng_block(seed, STREAM_DATA, unit, uint64_t(p >> 2)), p & 3) >>
                               (32 - m.bits));
                else
                    a[t] = idx[i * S * NT + p];
                x[t] = rsnt * map_symbol<T>(m, tab, a[t]);
            }
            cx<double> b[NT];
#pragma unroll
            for (int t = 0; t < NT; ++t) b[t] = {0.0, 0.0};
#pragma unroll
            for (int r = 0; r < B200PHY_MAX_ANT; ++r)
                if (r < Nr) {
                    cx<T> nz;
                    if constexpr (FUSED) nz = cnormal_at<T>(seed, STREAM_NOISE, unit, uint64_t(r) * row + s);
                    else nz = load1(noise + (i * Nr + r) * S + s);
                    cx<T> y = sigma * nz;
#pragma unroll
                    for (int t = 0; t < NT; ++t) cmac(y, H[r][t], x[t]);
#pragma unroll
                    for (int t = 0; t < NT; ++t) cmac_conj(b[t], cvt<double>(H[r][t]), cvt<double>(y));
                }
            sol.solve(b);
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                const cx<T> z = {T(b[t].re * snt), T(b[t].im * snt)};
                const int e = demap_symbol<T>(m, tab, z);
                sym_err += (e != a[t]);
                bit_err += __popc(e ^ a[t]);
                const long long o = i * S * NT + s * NT + t;
                if (idx_hat) idx_hat[o] = (uint8_t)e;
                if (dec_out) dec_out[o] = z;
            }
        }
    }
    flush_counters(sym_err, bit_err, counters);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        atomicAdd(&counters[2], (unsigned long long)n * S * NT);
        atomicAdd(&counters[3], (unsigned long long)n * S * NT * m.bits);
    }
}

// ================================================================= draw dumps
template <typename T>
__global__ void draw_siso_flat_kernel(int bits, const __grid_constant__ PhiloxKey seed, uint64_t first_unit, long long n,
                                      uint8_t *idx, cx<T> *h, cx<T> *noise) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const uint64_t unit = first_unit + uint64_t(i);
        if (idx) idx[i] = uint8_t(rng_block(seed, STREAM_DATA, unit, 0).x >> (32 - bits));
        if (h) { const uint4 b = rng_block(seed, STREAM_CHANNEL, unit, 0); h[i] = cnormal<T>(b.x, b.y); }
        if (noise) { const uint4 b = rng_block(seed, STREAM_NOISE, unit, 0); noise[i] = cnormal<T>(b.x, b.y); }
    }
}

template <typename T>
__global__ void draw_flat_mimo_kernel(int bits, int Nr, int Nt, int S, int n_data, const __grid_constant__ PhiloxKey seed,
                                      uint64_t first_unit, long long n, uint8_t *idx, cx<T> *H,
                                      cx<T> *noise) {
    const int row = 2 * ((S + 1) / 2);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const uint64_t unit = first_unit + uint64_t(i);
        if (idx)
            for (int p = 0; p < n_data; ++p)
                idx[i * n_data + p] =
                    uint8_t(lane_of(rng_block(seed, STREAM_DATA, unit, uint64_t(p >> 2)), p & 3) >> (32 - bits));
        if (H)
            for (int j = 0; j < Nr * Nt; ++j) H[i * Nr * Nt + j] = cnormal_at<T>(seed, STREAM_CHANNEL, unit, j);
        if (noise)
            for (int r = 0; r < Nr; ++r)
                for (int s = 0; s < S; ++s)
                    noise[(i * Nr + r) * S + s] = cnormal_at<T>(seed, STREAM_NOISE, unit, uint64_t(r) * row + s);
    }
}

// ---------------------------------------------------------------- launch helpers
static int grid_for(long long work_items) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long blocks = (work_items + kThreads - 1) / kThreads;
    const long long cap = (long long)sms * 8;        // 8 resident 256-thread CTAs per SM
    return int(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

static bool draws_consistent(const void *a, const void *b, const void *c, bool need_b) {
    const bool any = a || c || (need_b && b);
    const bool all = a && c && (!need_b || b);
    return !any || all;
}

template <typename T>
static int launch_siso_flat(const Modem &m, const void *table, int rayleigh, double noise_var,
                            uint64_t seed, uint64_t first, int64_t n, const uint8_t *idx,
                            const void *h, const void *noise, uint8_t *idx_hat, void *dec,
                            int64_t *counters, cudaStream_t st) {
    const bool fused = !idx;
    const int grid = grid_for(n / 4 + 1);
    auto args = [&](auto kern) {
        kern<<<grid, kThreads, 0, st>>>(m, (const cx<T> *)table, T(sqrt(noise_var)), seed, first,
                                        (long long)n, idx, (const cx<T> *)h, (const cx<T> *)noise,
                                        idx_hat, (cx<T> *)dec, (unsigned long long *)counters);
    };
    const bool qam = m.kind == B200PHY_MODEM_QAM;
#define B200_SISO3(F, R, Q) do { if (dec) args(siso_flat_kernel<T, F, R, Q, true>); else args(siso_flat_kernel<T, F, R, Q, false>); } while (0)
#define B200_SISO2(F, R) do { if (qam) B200_SISO3(F, R, true); else B200_SISO3(F, R, false); } while (0)
    if (fused) { if (rayleigh) B200_SISO2(true, true); else B200_SISO2(true, false); }
    else { if (rayleigh) B200_SISO2(false, true); else B200_SISO2(false, false); }
#undef B200_SISO2
#undef B200_SISO3
    note_kernel("siso_flat_kernel<%s,%d,%d,%d,%d>", sizeof(T) == 4 ? "float" : "double", int(fused), int(rayleigh != 0), int(qam), int(dec != nullptr));
    B200_CHECK_LAUNCH("siso_flat_kernel");
    return B200PHY_OK;
}

template <typename T>
static int launch_alamouti(const Modem &m, const void *table, int Nr, int S, double noise_var,
                           uint64_t seed, uint64_t first, int64_t n, const uint8_t *idx,
                           const void *H, const void *noise, uint8_t *idx_hat, void *dec,
                           int64_t *counters, cudaStream_t st) {
    const bool fused = !idx;
    const bool qpsk = m.kind == B200PHY_MODEM_QPSK;
    if constexpr (sizeof(T) == 4) {
        if (Nr == 2 && S == 2) {                 // the C4 shape: packed rx-antenna lanes
            const int grid = grid_for(n);
            auto go = [&](auto kern) {
                kern<<<grid, kThreads, 0, st>>>(m, (const cx<float> *)table, float(sqrt(noise_var)), seed, first,
                                                (long long)n, idx, (const float4 *)H, (const float4 *)noise, idx_hat,
                                                (cx<float> *)dec, (unsigned long long *)counters);
            };
#define B200_A22(F, D) do { if (qpsk) go(alamouti22_kernel<F, D, true>); else go(alamouti22_kernel<F, D, false>); } while (0)
            if (fused) { if (dec) B200_A22(true, true); else B200_A22(true, false); }
            else { if (dec) B200_A22(false, true); else B200_A22(false, false); }
#undef B200_A22
            note_kernel("alamouti22_kernel<%d,%d,%d>", int(fused), int(dec != nullptr), int(qpsk));
            B200_CHECK_LAUNCH("alamouti22_kernel");
            return B200PHY_OK;
        }
    }
    const int grid = grid_for(n);
    auto args = [&](auto kern) {
        kern<<<grid, kThreads, 0, st>>>(m, (const cx<T> *)table, S, T(sqrt(noise_var)), seed, first,
                                        (long long)n, idx, (const cx<T> *)H, (const cx<T> *)noise,
                                        idx_hat, (cx<T> *)dec, (unsigned long long *)counters);
    };
#define B200_ALA3(F, N_, D_) do { if (qpsk) args(alamouti_kernel<T, F, N_, D_, true>); else args(alamouti_kernel<T, F, N_, D_, false>); } while (0)
#define B200_ALA2(F, N_) do { if (dec) B200_ALA3(F, N_, true); else B200_ALA3(F, N_, false); } while (0)
#define B200_ALA(F) do { switch (Nr) { case 1: B200_ALA2(F, 1); break; case 2: B200_ALA2(F, 2); break; \
                                        case 3: B200_ALA2(F, 3); break; default: B200_ALA2(F, 4); } } while (0)
    if (fused) B200_ALA(true); else B200_ALA(false);
#undef B200_ALA
#undef B200_ALA2
#undef B200_ALA3
    note_kernel("alamouti_kernel<%s,%d,%d,%d,%d>", sizeof(T) == 4 ? "float" : "double", int(fused), Nr > 3 ? 4 : Nr, int(dec != nullptr), int(qpsk));
    B200_CHECK_LAUNCH("alamouti_kernel");
    return B200PHY_OK;
}

template <typename T, int NT>
static int launch_blast_nt(const Modem &m, const void *table, int Nr, int S, double noise_var,
                           double fnv, uint64_t seed, uint64_t first, int64_t n, const uint8_t *idx,
                           const void *H, const void *noise, uint8_t *idx_hat, void *dec,
                           int64_t *counters, cudaStream_t st) {
    const bool fused = !idx;
    const int grid = grid_for(n);
    auto args = [&](auto kern) {
        kern<<<grid, kThreads, 0, st>>>(m, (const cx<T> *)table, Nr, S, T(sqrt(noise_var)), fnv, seed, first,
                                        (long long)n, idx, (const cx<T> *)H, (const cx<T> *)noise,
                                        idx_hat, (cx<T> *)dec, (unsigned long long *)counters);
    };
    if (fused) args(blast_kernel<T, true, NT>); else args(blast_kernel<T, false, NT>);
    B200_CHECK_LAUNCH("blast_kernel");
    return B200PHY_OK;
}

template <typename T>
static int launch_blast(const Modem &m, const void *table, int Nr, int Nt, int S, double noise_var,
                        double fnv, uint64_t seed, uint64_t first, int64_t n, const uint8_t *idx,
                        const void *H, const void *noise, uint8_t *idx_hat, void *dec,
                        int64_t *counters, cudaStream_t st) {
    switch (Nt) {
        case 1: return launch_blast_nt<T, 1>(m, table, Nr, S, noise_var, fnv, seed, first, n, idx, H, noise, idx_hat, dec, counters, st);
        case 2: return launch_blast_nt<T, 2>(m, table, Nr, S, noise_var, fnv, seed, first, n, idx, H, noise, idx_hat, dec, counters, st);
        case 3: return launch_blast_nt<T, 3>(m, table, Nr, S, noise_var, fnv, seed, first, n, idx, H, noise, idx_hat, dec, counters, st);
        default: return launch_blast_nt<T, 4>(m, table, Nr, S, noise_var, fnv, seed, first, n, idx, H, noise, idx_hat, dec, counters, st);
    }
}

static int check_common(int dtype, int64_t n, double noise_var, const int64_t *counters) {
    if (dtype != B200PHY_F32 && dtype != B200PHY_F64) { set_error("dtype must be B200PHY_F32 or B200PHY_F64"); return B200PHY_ERR_INVALID; }
    if (n < 0) { set_error("n_units must be non-negative"); return B200PHY_ERR_INVALID; }
    if (!(noise_var >= 0.0)) { set_error("Noise variance must be a non-negative value."); return B200PHY_ERR_INVALID; }
    if (!counters) { set_error("counters is NULL"); return B200PHY_ERR_INVALID; }
    return B200PHY_OK;
}

}  // namespace b200phy

using namespace b200phy;

extern "C" {

int b200phy_link_siso_flat(int dtype, const b200phy_modem *modem, int rayleigh, double noise_var,
                           uint64_t seed, uint64_t first_unit, int64_t n_units, const uint8_t *idx,
                           const void *h, const void *noise, uint8_t *idx_hat, void *dec_out,
                           int64_t *counters, void *stream) {
    Modem m;
    int e = check_modem(modem, &m);
    if (e) return e;
    if ((e = check_common(dtype, n_units, noise_var, counters))) return e;
    if (!draws_consistent(idx, h, noise, rayleigh != 0)) {
        set_error("stream mode needs idx, noise%s together; fused mode needs all NULL", rayleigh ? ", h" : "");
        return B200PHY_ERR_INVALID;
    }
    if ((e = require_aligned(idx, 4, "idx")) || (e = require_aligned(h, 16, "h")) || (e = require_aligned(noise, 16, "noise")) ||
        (e = require_aligned(idx_hat, 4, "idx_hat")) || (e = require_aligned(dec_out, dtype == B200PHY_F32 ? 8 : 16, "dec_out")))
        return e;
    if (n_units == 0) return B200PHY_OK;
    cudaStream_t st = (cudaStream_t)stream;
    return dtype == B200PHY_F32
               ? launch_siso_flat<float>(m, modem->table, rayleigh, noise_var, seed, first_unit, n_units, idx, h, noise, idx_hat, dec_out, counters, st)
               : launch_siso_flat<double>(m, modem->table, rayleigh, noise_var, seed, first_unit, n_units, idx, h, noise, idx_hat, dec_out, counters, st);
}

int b200phy_link_alamouti(int dtype, const b200phy_modem *modem, int Nr, int S, double noise_var,
                          uint64_t seed, uint64_t first_unit, int64_t n_units, const uint8_t *idx,
                          const void *H, const void *noise, uint8_t *idx_hat, void *dec_out,
                          int64_t *counters, void *stream) {
    Modem m;
    int e = check_modem(modem, &m);
    if (e) return e;
    if ((e = check_common(dtype, n_units, noise_var, counters))) return e;
    if (Nr < 1 || Nr > B200PHY_MAX_ANT) { set_error("Alamouti: Nr=%d must be in [1, %d]", Nr, B200PHY_MAX_ANT); return B200PHY_ERR_UNSUPPORTED; }
    if (S < 2 || (S & 1)) { set_error("Alamouti: number of symbols S=%d must be even", S); return B200PHY_ERR_INVALID; }
    if (!draws_consistent(idx, H, noise, true)) { set_error("stream mode needs idx, H, noise together"); return B200PHY_ERR_INVALID; }
    if ((e = require_aligned(idx, 2, "idx")) || (e = require_aligned(H, 16, "H")) || (e = require_aligned(noise, 16, "noise")) ||
        (e = require_aligned(idx_hat, 2, "idx_hat")) || (e = require_aligned(dec_out, dtype == B200PHY_F32 ? 8 : 16, "dec_out")))
        return e;
    if (n_units == 0) return B200PHY_OK;
    cudaStream_t st = (cudaStream_t)stream;
    return dtype == B200PHY_F32
               ? launch_alamouti<float>(m, modem->table, Nr, S, noise_var, seed, first_unit, n_units, idx, H, noise, idx_hat, dec_out, counters, st)
               : launch_alamouti<double>(m, modem->table, Nr, S, noise_var, seed, first_unit, n_units, idx, H, noise, idx_hat, dec_out, counters, st);
}

int b200phy_link_blast(int dtype, const b200phy_modem *modem, int Nr, int Nt, int S, double noise_var,
                       double filter_noise_var, uint64_t seed, uint64_t first_unit, int64_t n_units,
                       const uint8_t *idx, const void *H, const void *noise, uint8_t *idx_hat,
                       void *dec_out, int64_t *counters, void *stream) {
    Modem m;
    int e = check_modem(modem, &m);
    if (e) return e;
    if ((e = check_common(dtype, n_units, noise_var, counters))) return e;
    if (Nr < 1 || Nr > B200PHY_MAX_ANT || Nt < 1 || Nt > B200PHY_MAX_ANT) {
        set_error("Blast: Nr=%d, Nt=%d must be in [1, %d]", Nr, Nt, B200PHY_MAX_ANT);
        return B200PHY_ERR_UNSUPPORTED;
    }
    if (!(filter_noise_var >= 0.0)) { set_error("Noise variance must be a non-negative value."); return B200PHY_ERR_INVALID; }
    if (filter_noise_var == 0.0 && Nt > Nr) { set_error("Blast ZF needs Nt <= Nr (got %dx%d)", Nr, Nt); return B200PHY_ERR_UNSUPPORTED; }
    if (S < 1) { set_error("S must be positive"); return B200PHY_ERR_INVALID; }
    if (!draws_consistent(idx, H, noise, true)) { set_error("stream mode needs idx, H, noise together"); return B200PHY_ERR_INVALID; }
    {
        const size_t ca = dtype == B200PHY_F32 ? 8 : 16;
        if ((e = require_aligned(H, ca, "H")) || (e = require_aligned(noise, ca, "noise")) || (e = require_aligned(dec_out, ca, "dec_out")))
            return e;
    }
    if (n_units == 0) return B200PHY_OK;
    cudaStream_t st = (cudaStream_t)stream;
    return dtype == B200PHY_F32
               ? launch_blast<float>(m, modem->table, Nr, Nt, S, noise_var, filter_noise_var, seed, first_unit, n_units, idx, H, noise, idx_hat, dec_out, counters, st)
               : launch_blast<double>(m, modem->table, Nr, Nt, S, noise_var, filter_noise_var, seed, first_unit, n_units, idx, H, noise, idx_hat, dec_out, counters, st);
}

int b200phy_draw_siso_flat(int dtype, int bits, uint64_t seed, uint64_t first_unit, int64_t n_units,
                           uint8_t *idx, void *h, void *noise, void *stream) {
    if (n_units <= 0) return B200PHY_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = grid_for(n_units);
    if (dtype == B200PHY_F32)
        draw_siso_flat_kernel<float><<<grid, kThreads, 0, st>>>(bits, seed, first_unit, n_units, idx, (cx<float> *)h, (cx<float> *)noise);
    else
        draw_siso_flat_kernel<double><<<grid, kThreads, 0, st>>>(bits, seed, first_unit, n_units, idx, (cx<double> *)h, (cx<double> *)noise);
    B200_CHECK_LAUNCH("draw_siso_flat_kernel");
    return B200PHY_OK;
}

int b200phy_draw_flat_mimo(int dtype, int bits, int Nr, int Nt, int S, int n_data, uint64_t seed,
                           uint64_t first_unit, int64_t n_units, uint8_t *idx, void *H, void *noise,
                           void *stream) {
    if (n_units <= 0) return B200PHY_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = grid_for(n_units);
    if (dtype == B200PHY_F32)
        draw_flat_mimo_kernel<float><<<grid, kThreads, 0, st>>>(bits, Nr, Nt, S, n_data, seed, first_unit, n_units, idx, (cx<float> *)H, (cx<float> *)noise);
    else
        draw_flat_mimo_kernel<double><<<grid, kThreads, 0, st>>>(bits, Nr, Nt, S, n_data, seed, first_unit, n_units, idx, (cx<double> *)H, (cx<double> *)noise);
    B200_CHECK_LAUNCH("draw_flat_mimo_kernel");
    return B200PHY_OK;
}

}  // extern "C"

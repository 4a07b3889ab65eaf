// ofdm_tdl.cuh — the fused OFDM-over-Jakes/TDL link kernel (one CTA per frame, persistent).
//
// Per frame (notebooks/TDL_and_OFDM.ipynb cell 32; MIMO: SURVEY.md §8d C5):
//   data symbols -> QAM map -> [Blast encode] -> per tx antenna: subcarrier scatter + IFFT + CP
//   -> time-varying sparse FIR with Jakes taps (TdlChannel.corrupt_data, fading.py:1046-1124)
//   -> + AWGN -> strip CP + FFT per rx antenna -> H_k = DFT(mean taps) (ofdm.py:515-552 restated)
//   -> one-tap divide or per-subcarrier ZF/MMSE (mimo.py:264-309) -> demap -> error counters.
// Nothing but the draws (stream mode) / nothing at all (fused mode) is read from HBM and only the
// demapped indices + 4 counters are written; every intermediate lives in shared memory/registers.
//
// Jakes taps h(n) = L^-1/2 sum_o exp(j(theta_o + n*Delta_o)) (fading_generators.py:519-523) are
// evaluated one of two ways (DESIGN.md "Jakes evaluation"):
//   RECURRENCE  chunks of 8 samples: exact sincos at the chunk start, then z += z*d rotations.
//   POLY        3rd-order Taylor polynomial in the sample offset around a segment centre, used when
//               sqrt(L)*(w*dmax)^4/24 is below the dtype's resolution (slow fading: all BASELINE
//               configs); 4 FMAs per tap sample instead of 6*L.
#pragma once
#include "common.cuh"
#include "linalg.cuh"
#include "rng.cuh"

namespace b200phy {

constexpr int kOT = 256;       // threads per CTA
constexpr int kJBC = 4;        // outputs per thread per register block in the channel apply
constexpr int kCH = 8;         // recurrence chunk length

struct OfdmP {
    int fft, lg, cp, used, half, n_sym, S, N, mem, n_taps, L, n_data;
    int poly, nseg, seg_len, seg_lg;
    int row;        // noise normals per rx row in the Philox layout: 2*ceil((N+mem)/2)
    int P, P4;      // phases per frame, rounded up to a multiple of 4
    int ifft_in_w;  // 1: scatter into W so that the ping-pong IFFT ends in E.body
    int delays[B200PHY_MAX_TAPS];
    double amp[B200PHY_MAX_TAPS];   // sqrt(P_l / L)
    double w0;      // 2*pi*Fd
    double Ts1;     // Ts * 1.0000000001 (fading_generators.py:462)
    double t0;
    double sigma, fnv, tx_scale, rx_scale, snt;
    uint64_t seed;
};

template <typename T> struct OscRec { double th0, dl; cx<T> d; };

// ---------------------------------------------------------------- shared-memory Stockham FFT
// Radix-4 passes (+ one radix-2 pass when lg is odd), natural order in and out, ping-pong between
// a and b.  tw[m] = exp(-2 pi i m / N), m in [0, N).  INV uses conjugate twiddles (unnormalised).
// Every thread of the CTA must call it; returns the buffer holding the result (synchronised).
template <typename T, bool INV>
__device__ cx<T> *fft_stockham(cx<T> *a, cx<T> *b, const cx<T> *tw, int N, int lg) {
    cx<T> *src = a, *dst = b;
    int Ns = 1;
    for (int st = 0; st < (lg >> 1); ++st) {
        __syncthreads();
        const int q = N >> 2;
        for (int j = threadIdx.x; j < q; j += blockDim.x) {
            const int k = j & (Ns - 1);
            cx<T> v0 = src[j], v1 = src[j + q], v2 = src[j + 2 * q], v3 = src[j + 3 * q];
            if (Ns > 1) {
                const int ts = k * (q / Ns);
                cx<T> w1 = tw[ts], w2 = tw[2 * ts], w3 = tw[3 * ts];
                if (INV) { w1.im = -w1.im; w2.im = -w2.im; w3.im = -w3.im; }
                v1 = v1 * w1; v2 = v2 * w2; v3 = v3 * w3;
            }
            const cx<T> a0 = v0 + v2, a1 = v0 - v2, a2 = v1 + v3, a3 = v1 - v3;
            // forward: -j*a3 = (a3.im, -a3.re); inverse: +j*a3 = (-a3.im, a3.re)
            const cx<T> rot = INV ? mk<T>(-a3.im, a3.re) : mk<T>(a3.im, -a3.re);
            const int j0 = ((j - k) << 2) + k;
            dst[j0] = a0 + a2;
            dst[j0 + Ns] = a1 + rot;
            dst[j0 + 2 * Ns] = a0 - a2;
            dst[j0 + 3 * Ns] = a1 - rot;
        }
        Ns <<= 2;
        cx<T> *t = src; src = dst; dst = t;
    }
    if (lg & 1) {
        __syncthreads();
        const int h = N >> 1;
        for (int j = threadIdx.x; j < h; j += blockDim.x) {
            const int k = j & (Ns - 1);
            cx<T> w = tw[k * (h / Ns)];
            if (INV) w.im = -w.im;
            const cx<T> v0 = src[j], v1 = src[j + h] * w;
            const int j0 = ((j - k) << 1) + k;
            dst[j0] = v0 + v1;
            dst[j0 + Ns] = v0 - v1;
        }
        cx<T> *t = src; src = dst; dst = t;
    }
    __syncthreads();
    return src;
}

__device__ __forceinline__ int fft_passes(int lg) { return (lg >> 1) + (lg & 1); }

// OFDM.get_used_subcarrier_indexes (modulators/ofdm.py:188-224): data position q -> FFT bin
__device__ __forceinline__ int bin_of(int q, int fft, int used, int half) {
    if (used == fft) return (q + (fft >> 1)) & (fft - 1);
    return q < half ? fft - half + q : q - half + 1;
}
// inverse: FFT bin -> data position or -1 (unused bin)
__device__ __forceinline__ int pos_of(int k, int fft, int used, int half) {
    if (used == fft) return (k + (fft >> 1)) & (fft - 1);
    if (k >= 1 && k <= half) return half + k - 1;
    if (k >= fft - half) return k - (fft - half);
    return -1;
}

// theta reduced to [-pi, pi] in double, then cast
__device__ __forceinline__ double reduce_2pi(double th) {
    return fma(-6.283185307179586476925, rint(th * 0.15915494309189533577), th);
}

__device__ __forceinline__ void sincos_t(float x, float *s, float *c) { sincosf(x, s, c); }
__device__ __forceinline__ void sincos_t(double x, double *s, double *c) { sincos(x, s, c); }

template <typename P, int N> __device__ __forceinline__ P pick(P const (&arr)[N], int r) {
    P v = arr[0];
#pragma unroll
    for (int i = 1; i < N; ++i) v = (r == i) ? arr[i] : v;
    return v;
}

template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// mean over S consecutive samples of exp(j(th + n*dl)) = exp(j(th + (S-1)dl/2)) * sin(S dl/2)/(S sin(dl/2))
template <typename T> __device__ __forceinline__ T dirichlet(double dl, int S) {
    const T h = T(0.5 * dl);
    const T den = sin(h);
    if (fabs(den) < T(1e-30)) return T(1);
    return sin(T(S) * h) / (T(S) * den);
}

// ================================================================= the kernel
template <typename T, bool FUSED, int NR, int NT, bool WSG>
__global__ void __launch_bounds__(kOT, (sizeof(T) == 4 && NR * NT <= 4) ? 3 : 1)
ofdm_tdl_kernel(const __grid_constant__ OfdmP p, const Modem m, const cx<T> *__restrict__ tab_g,
                uint64_t first_unit, long long n_units, const uint8_t *__restrict__ idx_g,
                const T *__restrict__ phi_g, const T *__restrict__ psi_g,
                const cx<T> *__restrict__ noise_g, uint8_t *__restrict__ idx_hat,
                cx<T> *__restrict__ eq_out, cx<T> *__restrict__ ws_g, unsigned long long *counters) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fft = p.fft, S = p.S, mem = p.mem, cp = p.cp;

    // ---- carve shared memory
    unsigned char *sp = smem_raw;
    auto take = [&](size_t bytes) { unsigned char *r = sp; sp += (bytes + 15) & ~size_t(15); return r; };
    cx<T> *tw = (cx<T> *)take(sizeof(cx<T>) * fft);
    cx<T> *E = (cx<T> *)take(sizeof(cx<T>) * (mem + S));      // [tail | cp | body]
    cx<T> *body = E + mem + cp;
    cx<T> *gbar = (cx<T> *)take(sizeof(cx<T>) * p.n_taps * NR * NT);
    cx<T> *tails = (cx<T> *)take(p.n_sym > 1 ? sizeof(cx<T>) * NT * mem : 0);
    cx<T> *coef = (cx<T> *)take(p.poly ? sizeof(cx<T>) * p.n_taps * NR * p.nseg * 4 : 0);
    OscRec<T> *osc = (OscRec<T> *)take(p.poly ? 0 : sizeof(OscRec<T>) * p.n_taps * NR * p.L);
    cx<T> *tab = (cx<T> *)take(sizeof(cx<T>) * m.M);
    uint8_t *dsym = (uint8_t *)take(NT * p.used);
    cx<T> *pool;                                              // NR + 1 buffers of fft samples
    if constexpr (WSG) pool = ws_g + (size_t)blockIdx.x * (NR + 1) * fft;
    else pool = (cx<T> *)take(sizeof(cx<T>) * (NR + 1) * fft);

    for (int i = tid; i < fft; i += kOT) {
        double s, c;
        sincospi(-2.0 * double(i) / double(fft), &s, &c);
        tw[i] = {T(c), T(s)};
    }
    if (m.kind != B200PHY_MODEM_BPSK)
        for (int k = tid; k < m.M; k += kOT) tab[k] = tab_g[k];
    __syncthreads();

    unsigned sym_err = 0, bit_err = 0;
    const T sigma = T(p.sigma), tx_scale = T(p.tx_scale), rx_scale = T(p.rx_scale);
    const int n_items = p.n_taps * NR;

    for (long long frame = blockIdx.x; frame < n_units; frame += gridDim.x) {
        const uint64_t unit = first_unit + uint64_t(frame);
        cx<T> *Yp[NR];
#pragma unroll
        for (int r = 0; r < NR; ++r) Yp[r] = pool + r * fft;
        cx<T> *W = pool + NR * fft;

        for (int s = 0; s < p.n_sym; ++s) {
            const int n_s = s * S;
            // ---------------- P0: data symbols of this OFDM symbol, noise into Y
            {
                const int w0 = s * p.used * NT, cnt = p.used * NT;
                if constexpr (FUSED) {
                    const int b0 = w0 >> 2, b1 = (w0 + cnt - 1) >> 2;
                    for (int b = b0 + tid; b <= b1; b += kOT) {
                        const uint4 blk = rng_block(p.seed, STREAM_DATA, unit, uint64_t(b));
#pragma unroll
                        for (int l = 0; l < 4; ++l) {
                            const int w = 4 * b + l - w0;
                            if (w >= 0 && w < cnt) dsym[w] = uint8_t(lane_of(blk, l) >> (32 - m.bits));
                        }
                    }
                } else {
                    const uint8_t *src = idx_g + frame * p.n_data + w0;
                    for (int i = tid; i < cnt; i += kOT) dsym[i] = src[i];
                }
                const int m0 = n_s + cp;          // first needed rx sample of this symbol
                if constexpr (FUSED) {
                    const int pr0 = m0 >> 1, npr = ((m0 + fft - 1) >> 1) - pr0 + 1;
                    for (int it = tid; it < NR * npr; it += kOT) {
                        const int r = it / npr, pr = pr0 + it % npr;
                        const uint4 blk = rng_block(p.seed, STREAM_NOISE, unit, uint64_t(r) * (p.row >> 1) + pr);
                        const int j = 2 * pr - m0;
                        cx<T> *yr = pick(Yp, r);
                        if (j >= 0 && j < fft) yr[j] = sigma * cnormal<T>(blk.x, blk.y);
                        if (j + 1 >= 0 && j + 1 < fft) yr[j + 1] = sigma * cnormal<T>(blk.z, blk.w);
                    }
                } else {
                    const size_t rowlen = size_t(p.N + mem);
#pragma unroll
                    for (int r = 0; r < NR; ++r) {
                        const cx<T> *src = noise_g + (size_t(frame) * NR + r) * rowlen + m0;
                        for (int j = tid; j < fft; j += kOT) Yp[r][j] = sigma * src[j];
                    }
                }
            }
            __syncthreads();

            for (int t = 0; t < NT; ++t) {
                // ---------------- A: QAM map + subcarrier scatter (input of the IFFT)
                cx<T> *in = p.ifft_in_w ? W : body;
                cx<T> *other = p.ifft_in_w ? body : W;
                for (int k = tid; k < fft; k += kOT) {
                    const int q = pos_of(k, fft, p.used, p.half);
                    cx<T> v = {T(0), T(0)};
                    if (q >= 0) v = tx_scale * map_symbol<T>(m, tab, dsym[q * NT + t]);
                    in[k] = v;
                }
                // ---------------- C: per-oscillator setup for tx antenna t (lanes = rays)
                for (int it = warp; it < n_items; it += kOT / 32) {
                    const int l = it / NR, r = it % NR;
                    const T amp = T(p.amp[l]);
                    cx<T> gsum = {T(0), T(0)};
                    cx<T> a[4][4];      // [seg][order]; nseg <= 4 handled in registers per pass
                    for (int sg0 = 0; sg0 < (p.poly ? p.nseg : 1); sg0 += 4) {
#pragma unroll
                        for (int u = 0; u < 4; ++u)
#pragma unroll
                            for (int o = 0; o < 4; ++o) a[u][o] = {T(0), T(0)};
                        for (int o = lane; o < p.L; o += 32) {
                            const int i = ((o * p.n_taps + l) * NR + r) * NT + t;
                            T phi, psi;
                            if constexpr (FUSED) {
                                phi = phase_from_word<T>(lane_of(rng_block(p.seed, STREAM_CHANNEL, unit, uint64_t(i >> 2)), i & 3));
                                const int i2 = p.P4 + i;
                                psi = phase_from_word<T>(lane_of(rng_block(p.seed, STREAM_CHANNEL, unit, uint64_t(i2 >> 2)), i2 & 3));
                            } else {
                                phi = phi_g[size_t(frame) * p.P + i];
                                psi = psi_g[size_t(frame) * p.P + i];
                            }
                            const double cphi = cos(double(phi));
                            const double dl = p.w0 * p.Ts1 * cphi;             // phase step per sample
                            const double th0 = double(psi) + p.w0 * cphi * p.t0;
                            if (sg0 == 0) {
                                // mean tap over the S samples of this symbol (CP included, ofdm.py:541-548)
                                const double mid = reduce_2pi(fma(dl, double(n_s) + 0.5 * double(S - 1), th0));
                                T sn, cs;
                                sincos_t(T(mid), &sn, &cs);
                                const T g = amp * dirichlet<T>(dl, S);
                                gsum.re += g * cs;
                                gsum.im += g * sn;
                                if (!p.poly) {
                                    OscRec<T> rec;
                                    rec.th0 = th0; rec.dl = dl;
                                    T sh, ch, sd, cd;
                                    sincos_t(T(0.5 * dl), &sh, &ch);
                                    sincos_t(T(dl), &sd, &cd);
                                    rec.d = {T(-2) * sh * sh, sd};              // exp(j dl) - 1
                                    osc[(l * NR + r) * p.L + o] = rec;
                                }
                            }
                            if (p.poly) {
                                const T d1 = T(dl), d2 = T(-0.5 * dl * dl), d3 = T(dl * dl * dl * (1.0 / 6.0));
#pragma unroll
                                for (int u = 0; u < 4; ++u)
                                    if (sg0 + u < p.nseg) {
                                        const double c = double(n_s + cp + (sg0 + u) * p.seg_len) + 0.5 * double(p.seg_len - 1);
                                        const double thc = reduce_2pi(fma(dl, c, th0));
                                        T sn, cs;
                                        sincos_t(T(thc), &sn, &cs);
                                        const cx<T> e = {amp * cs, amp * sn};
                                        a[u][0].re += e.re;        a[u][0].im += e.im;
                                        a[u][1].re -= d1 * e.im;   a[u][1].im += d1 * e.re;     // (j dl) e
                                        a[u][2].re += d2 * e.re;   a[u][2].im += d2 * e.im;     // -(dl^2/2) e
                                        a[u][3].re += d3 * e.im;   a[u][3].im -= d3 * e.re;     // -j(dl^3/6) e
                                    }
                            }
                        }
                        if (p.poly) {
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                if (sg0 + u < p.nseg) {
#pragma unroll
                                    for (int o = 0; o < 4; ++o) {
                                        const T re = warp_sum(a[u][o].re), im = warp_sum(a[u][o].im);
                                        if (lane == 0) coef[((l * NR + r) * p.nseg + sg0 + u) * 4 + o] = {re, im};
                                    }
                                }
                        }
                    }
                    const T gre = warp_sum(gsum.re), gim = warp_sum(gsum.im);
                    if (lane == 0) gbar[(l * NR + r) * NT + t] = {gre, gim};
                }
                // ---------------- B: IFFT (result lands in E.body), cyclic prefix, ISI tail
                fft_stockham<T, true>(in, other, tw, fft, p.lg);
                for (int i = tid; i < cp; i += kOT) E[mem + i] = body[fft - cp + i];
                for (int i = tid; i < mem; i += kOT)
                    E[i] = (s > 0) ? tails[t * mem + i] : mk<T>(T(0), T(0));
                __syncthreads();

                // ---------------- D: time-varying sparse FIR, accumulate into Y[r]
                if (p.poly) {
                    const T half_seg = T(0.5) * T(p.seg_len - 1);
                    for (int jb0 = 0; jb0 * kOT < fft; jb0 += kJBC) {
                        cx<T> acc[kJBC][NR];
#pragma unroll
                        for (int jb = 0; jb < kJBC; ++jb)
#pragma unroll
                            for (int r = 0; r < NR; ++r) acc[jb][r] = {T(0), T(0)};
                        int seg_prev = -1;
                        cx<T> cf[NR][4];
                        for (int l = 0; l < p.n_taps; ++l) {
                            const int d = p.delays[l];
                            seg_prev = -1;
#pragma unroll
                            for (int jb = 0; jb < kJBC; ++jb) {
                                const int j = tid + (jb0 + jb) * kOT;
                                if (j < fft) {
                                    const int seg = j >> p.seg_lg;         // CTA-uniform per jb: seg_len % 256 == 0 or nseg == 1
                                    if (seg != seg_prev) {
#pragma unroll
                                        for (int r = 0; r < NR; ++r)
#pragma unroll
                                            for (int o = 0; o < 4; ++o)
                                                cf[r][o] = coef[((l * NR + r) * p.nseg + seg) * 4 + o];
                                        seg_prev = seg;
                                    }
                                    const cx<T> x = E[mem + cp + j - d];
                                    const T dt = T(j - seg * p.seg_len - d) - half_seg;
#pragma unroll
                                    for (int r = 0; r < NR; ++r) {
                                        cx<T> g;
                                        g.re = fma(cf[r][3].re, dt, cf[r][2].re);
                                        g.im = fma(cf[r][3].im, dt, cf[r][2].im);
                                        g.re = fma(g.re, dt, cf[r][1].re);
                                        g.im = fma(g.im, dt, cf[r][1].im);
                                        g.re = fma(g.re, dt, cf[r][0].re);
                                        g.im = fma(g.im, dt, cf[r][0].im);
                                        cmac(acc[jb][r], g, x);
                                    }
                                }
                            }
                        }
#pragma unroll
                        for (int jb = 0; jb < kJBC; ++jb) {
                            const int j = tid + (jb0 + jb) * kOT;
                            if (j < fft)
#pragma unroll
                                for (int r = 0; r < NR; ++r) Yp[r][j] = Yp[r][j] + acc[jb][r];
                        }
                    }
                } else {
                    const int nch = (fft + kCH - 1) / kCH;
                    for (int it = tid; it < NR * nch; it += kOT) {
                        const int r = it / nch, j0 = (it % nch) * kCH;
                        cx<T> acc[kCH];
#pragma unroll
                        for (int i = 0; i < kCH; ++i) acc[i] = {T(0), T(0)};
                        for (int l = 0; l < p.n_taps; ++l) {
                            const int d = p.delays[l];
                            const double n0 = double(n_s + cp + j0 - d);
                            cx<T> g[kCH];
#pragma unroll
                            for (int i = 0; i < kCH; ++i) g[i] = {T(0), T(0)};
                            const OscRec<T> *orec = osc + (l * NR + r) * p.L;
                            for (int o = 0; o < p.L; ++o) {
                                const OscRec<T> rec = orec[o];
                                T sn, cs;
                                sincos_t(T(reduce_2pi(fma(rec.dl, n0, rec.th0))), &sn, &cs);
                                cx<T> z = {cs, sn};
#pragma unroll
                                for (int i = 0; i < kCH; ++i) {
                                    g[i].re += z.re; g[i].im += z.im;
                                    const cx<T> zn = {fma(z.re, rec.d.re, fma(-z.im, rec.d.im, z.re)),
                                                      fma(z.re, rec.d.im, fma(z.im, rec.d.re, z.im))};
                                    z = zn;
                                }
                            }
                            const T amp = T(p.amp[l]);
#pragma unroll
                            for (int i = 0; i < kCH; ++i)
                                if (j0 + i < fft) cmac(acc[i], amp * g[i], E[mem + cp + j0 + i - d]);
                        }
                        cx<T> *yr = pick(Yp, r);
#pragma unroll
                        for (int i = 0; i < kCH; ++i)
                            if (j0 + i < fft) yr[j0 + i] = yr[j0 + i] + acc[i];
                    }
                }
                __syncthreads();
                // ---------------- E: keep the last `mem` tx samples for the next symbol's ISI
                if (p.n_sym > 1) {
                    for (int i = tid; i < mem; i += kOT) tails[t * mem + i] = E[S + i];
                    __syncthreads();
                }
            }   // tx antennas

            // ---------------- F: FFT of every rx antenna (rotating buffer pool)
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                cx<T> *res = fft_stockham<T, false>(Yp[r], W, tw, fft, p.lg);
                if (res != Yp[r]) { W = Yp[r]; Yp[r] = res; }
            }

            // ---------------- G: H_k, equalise / detect, demap, count
            for (int q = tid; q < p.used; q += kOT) {
                const int k = bin_of(q, fft, p.used, p.half);
                cx<T> H[NR][NT];
#pragma unroll
                for (int r = 0; r < NR; ++r)
#pragma unroll
                    for (int t = 0; t < NT; ++t) H[r][t] = {T(0), T(0)};
                for (int l = 0; l < p.n_taps; ++l) {
                    const cx<T> w = tw[(k * p.delays[l]) & (fft - 1)];
#pragma unroll
                    for (int r = 0; r < NR; ++r)
#pragma unroll
                        for (int t = 0; t < NT; ++t) cmac(H[r][t], gbar[(l * NR + r) * NT + t], w);
                }
                cx<T> y[NR];
#pragma unroll
                for (int r = 0; r < NR; ++r) y[r] = rx_scale * Yp[r][k];
                cx<T> z[NT];
                if constexpr (NR == 1 && NT == 1) {
                    z[0] = cdiv(y[0], H[0][0]);                 // OfdmOneTapEqualizer (ofdm.py:510-511)
                } else {
                    HermSolver<NT> sol;
                    sol.template factor_from_channel<cx<T>, NR>(H, NR, p.fnv);
                    cx<double> b[NT];
#pragma unroll
                    for (int t = 0; t < NT; ++t) {
                        b[t] = {0.0, 0.0};
#pragma unroll
                        for (int r = 0; r < NR; ++r) cmac_conj(b[t], cvt<double>(H[r][t]), cvt<double>(y[r]));
                    }
                    sol.solve(b);
#pragma unroll
                    for (int t = 0; t < NT; ++t) z[t] = {T(b[t].re * p.snt), T(b[t].im * p.snt)};
                }
#pragma unroll
                for (int t = 0; t < NT; ++t) {
                    const int a = dsym[q * NT + t];
                    const int e = demap_symbol<T>(m, tab, z[t]);
                    sym_err += (e != a);
                    bit_err += __popc(e ^ a);
                    const size_t o = size_t(frame) * p.n_data + size_t(s * p.used + q) * NT + t;
                    if (idx_hat) idx_hat[o] = uint8_t(e);
                    if (eq_out) eq_out[o] = z[t];
                }
            }
            __syncthreads();
        }   // OFDM symbols
    }       // frames
    flush_counters(sym_err, bit_err, counters);
    if (blockIdx.x == 0 && tid == 0) {
        atomicAdd(&counters[2], (unsigned long long)n_units * p.n_data);
        atomicAdd(&counters[3], (unsigned long long)n_units * p.n_data * m.bits);
    }
}

// dynamic shared memory the kernel carves (must mirror the take() sequence above)
template <typename T> size_t ofdm_tdl_smem(const OfdmP &p, int M, int NR, int NT, bool wsg) {
    auto al = [](size_t b) { return (b + 15) & ~size_t(15); };
    size_t s = 0;
    s += al(sizeof(cx<T>) * p.fft);
    s += al(sizeof(cx<T>) * (p.mem + p.S));
    s += al(sizeof(cx<T>) * p.n_taps * NR * NT);
    s += al(p.n_sym > 1 ? sizeof(cx<T>) * NT * p.mem : 0);
    s += al(p.poly ? sizeof(cx<T>) * p.n_taps * NR * p.nseg * 4 : 0);
    s += al(p.poly ? 0 : sizeof(OscRec<T>) * p.n_taps * NR * p.L);
    s += al(sizeof(cx<T>) * M);
    s += al(size_t(NT) * p.used);
    if (!wsg) s += al(sizeof(cx<T>) * (NR + 1) * p.fft);
    return s;
}

// host-side launcher implemented once per (NR, NT) in link_ofdm_tdl_inst_*.cu
template <int NR, int NT>
int launch_ofdm_tdl(int dtype, const OfdmP &p, const Modem &m, const void *table, uint64_t first_unit,
                    int64_t n_units, const uint8_t *idx, const void *phi, const void *psi,
                    const void *noise, uint8_t *idx_hat, void *eq_out, int64_t *counters,
                    cudaStream_t st);

}  // namespace b200phy

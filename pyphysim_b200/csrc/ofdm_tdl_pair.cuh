// ofdm_tdl_pair.cuh — FFMA2 variant of the fused OFDM/TDL link for float, even Nr and even Nt, in the
// slow-fading regime (POLY, one segment, fft a multiple of 1024): antennas are processed in PAIRS that
// occupy the two lanes of Blackwell's packed FP32 instructions (fma/add/mul .f32x2).
//
// Layout: a "pair sample" is a float4 (a.re, b.re, a.im, b.im) for antennas (a, b) = (2p, 2p+1).
//   * the IFFT of both tx antennas of a pair and the FFT of both rx antennas of a pair are ONE Stockham
//     transform on float4 elements: every butterfly add is an FADD2, every twiddle product two FMUL2 +
//     two FFMA2 with the (scalar) twiddle as a broadcast operand — half the instructions per antenna;
//   * the FIR reads one float4 per (tap, output) for both tx antennas and accumulates rx pairs with
//     FFMA2 (coefficients stored as rx pairs), then adds into the rx pair buffer with two FADD2.
// Semantics, draws, restatements and error bounds are exactly those of ofdm_tdl.cuh (same OfdmP);
// tests run both kernels against the oracle and against each other.
#pragma once
#include <type_traits>

#include "ofdm_tdl.cuh"

namespace b200phy {


// pair sample: re = (a.re, b.re), im = (a.im, b.im)
struct ps { u64 re, im; };
__device__ __forceinline__ ps ld_ps(const float4 *p) {
    const float4 v = *p;
    return {pk2(v.x, v.y), pk2(v.z, v.w)};
}
__device__ __forceinline__ void st_ps(float4 *p, ps v) {
    float a, b, c, d;
    upk2(v.re, a, b);
    upk2(v.im, c, d);
    *p = make_float4(a, b, c, d);
}
__device__ __forceinline__ ps operator+(ps a, ps b) { return {add2(a.re, b.re), add2(a.im, b.im)}; }
__device__ __forceinline__ ps operator-(ps a, ps b) { return {sub2(a.re, b.re), sub2(a.im, b.im)}; }
// v * w for a scalar complex w (both lanes)
__device__ __forceinline__ ps mul_w(ps v, float wre, float wim) {
    const u64 WR = pk2(wre, wre), WI = pk2(wim, wim), NWI = pk2(-wim, -wim);
    return {fma2(v.im, NWI, mul2(v.re, WR)), fma2(v.im, WR, mul2(v.re, WI))};
}

// The twiddle table tw[N] (w^i = e^{-2 pi j i / N}) is followed by kTwc compact per-stage tables for the
// radix-4 stages with Ns = 4, 16, 64: stage Ns keeps (w^{k s}, w^{2 k s}, w^{3 k s}), s = N / (4 Ns), k < Ns,
// as three contiguous runs at offset Ns - 4.  Reading them from the full table costs 4- to 16-way shared
// memory bank conflicts in exactly these stages (stride s * 8 B between lanes; ncu: 13 M of the kernel's
// 23 M excess wavefronts).
constexpr int kTwc = 252;

__device__ inline void fill_compact_twiddles(cx<float> *tw, int N) {
    cx<float> *twc = tw + N;
    for (int e = threadIdx.x; e < kTwc; e += blockDim.x) {
        const int Ns = e < 12 ? 4 : (e < 60 ? 16 : 64);
        const int r = e - (Ns - 4), m = r / Ns + 1, k = r - (m - 1) * Ns;
        cx<float> w = {1.f, 0.f};
        if (4 * Ns <= N) {
            double s, c;
            sincospi(-2.0 * double(m * k * (N / (4 * Ns))) / double(N), &s, &c);
            w = {float(c), float(s)};
        }
        twc[e] = w;
    }
}

// ---- Stockham radix-4 (+ radix-2) FFT on pair samples; same structure as fft_stockham, split into pieces so
// that the kernels can keep the first stage's inputs / the last stage's outputs in registers (the thread that
// maps the symbols of bins j + i N/4 owns butterfly j of stage 0; the thread that detects bins k0 + u N/NU owns
// butterfly k0 of the last stage), which saves a shared-memory round trip and a barrier each.

// radix-4 butterfly; forward: y1 = a1 - j a3, y3 = a1 + j a3; inverse: signs swapped
template <bool INV>
__device__ __forceinline__ void bfly4(ps v0, ps v1, ps v2, ps v3, ps &y0, ps &y1, ps &y2, ps &y3) {
    const ps a0 = v0 + v2, a1 = v0 - v2, a2 = v1 + v3, a3 = v1 - v3;
    y0 = a0 + a2;
    y2 = a0 - a2;
    if (INV) { y1 = {sub2(a1.re, a3.im), add2(a1.im, a3.re)}; y3 = {add2(a1.re, a3.im), sub2(a1.im, a3.re)}; }
    else     { y1 = {add2(a1.re, a3.im), sub2(a1.im, a3.re)}; y3 = {sub2(a1.re, a3.im), add2(a1.im, a3.re)}; }
}

__device__ __forceinline__ int fft_swz(int x) { return (x & 3) | (((x >> 1) & 1) << 2); }
// the first stage's output buffer is XOR-swizzled when there are at least two radix-4 stages (see below)
__device__ __forceinline__ bool fft_sw(int N, int lg) { return (lg >> 1) >= 2 && N >= 64; }

// Stage 0 writes 4 consecutive elements per lane (64 B stride between lanes: a 4-way bank conflict on every
// STS.128).  Its output buffer is therefore XOR-swizzled in the low 3 index bits, a -> a ^ swz(a >> 3), which
// spreads a quarter-warp over all 8 bank groups; stage 1 reads it back through the same swizzle (a constant
// XOR per aligned group of 8 lanes: still conflict-free).
__device__ __forceinline__ void fft_store_stage0(float4 *dst, int j, bool sw, ps y0, ps y1, ps y2, ps y3) {
    if (sw) {
        const int f = fft_swz(j >> 1);                     // (4 j + m) >> 3 == j >> 1 for m < 4
        float4 *d4 = dst + ((4 * j) ^ (f & 4));
        const int c = f & 3;
        st_ps(d4 + c, y0);                                 // element m goes to slot m ^ c
        st_ps(d4 + (1 ^ c), y1);
        st_ps(d4 + (2 ^ c), y2);
        st_ps(d4 + (3 ^ c), y3);
    } else {
        st_ps(dst + 4 * j, y0);
        st_ps(dst + 4 * j + 1, y1);
        st_ps(dst + 4 * j + 2, y2);
        st_ps(dst + 4 * j + 3, y3);
    }
}

// twiddled inputs of butterfly j of radix-4 stage st >= 1 (Ns = 4^st)
template <bool INV>
__device__ __forceinline__ void fft_load_stage(const float4 *src, const cx<float> *tw, int N, int lg, int st, int Ns,
                                               int j, bool sw, ps &v0, ps &v1, ps &v2, ps &v3) {
    const int q = N >> 2, k = j & (Ns - 1);
    const int jl = (sw && st == 1) ? (j ^ fft_swz(j >> 3)) : j;
    v0 = ld_ps(src + jl); v1 = ld_ps(src + jl + q); v2 = ld_ps(src + jl + 2 * q); v3 = ld_ps(src + jl + 3 * q);
    cx<float> w1, w2, w3;
    if (Ns <= 64) {
        const cx<float> *t = tw + N + (Ns - 4) + k;
        w1 = t[0]; w2 = t[Ns]; w3 = t[2 * Ns];
    } else {
        const int ts = k << (lg - 2 - 2 * st);
        w1 = tw[ts]; w2 = tw[2 * ts]; w3 = tw[3 * ts];
    }
    v1 = mul_w(v1, w1.re, INV ? -w1.im : w1.im);
    v2 = mul_w(v2, w2.re, INV ? -w2.im : w2.im);
    v3 = mul_w(v3, w3.re, INV ? -w3.im : w3.im);
}

// Runs the radix-4 stages [first_stage, lg/2) and the radix-2 tail, except the LAST pass when skip_last
// (the caller then applies it from registers: fft_last_pass).  Data starts in `a` (natural order for stage 0,
// stage-0 output layout for first_stage == 1) and ping-pongs with `b`; returns the buffer holding the result,
// synchronised.  cp > 0 (never with skip_last): the last pass also writes the cyclic prefix, i.e. output
// element o >= N - cp goes to index o - N of the result buffer as well.
// NTHR: CTA size at compile time (0 = blockDim.x).  With N, lg and NTHR known the stage loop unrolls into
// straight-line stages (one butterfly per thread for N = 4 NTHR, constant strides, no stage-kind branches).
template <bool INV, int NTHR = 0>
__device__ __forceinline__ float4 *fft_stockham_pair(float4 *a, float4 *b, const cx<float> *tw, int N, int lg,
                                                     int first_stage = 0, bool skip_last = false, int cp = 0) {
    float4 *src = a, *dst = b;
    const bool sw = fft_sw(N, lg);
    const int nst = lg >> 1, q = N >> 2;
    const int nthr = NTHR ? NTHR : int(blockDim.x);
    const int last4 = (lg & 1) ? nst : (skip_last ? nst - 1 : nst);       // radix-4 stages to run: [first, last4)
    int Ns = 1 << (2 * first_stage);
#pragma unroll
    for (int st = first_stage; st < last4; ++st) {
        __syncthreads();
        const bool wcp = cp > 0 && !(lg & 1) && st == nst - 1;
        for (int j = threadIdx.x; j < q; j += nthr) {
            ps y0, y1, y2, y3;
            if (st == 0) {
                bfly4<INV>(ld_ps(src + j), ld_ps(src + j + q), ld_ps(src + j + 2 * q), ld_ps(src + j + 3 * q), y0, y1, y2, y3);
                fft_store_stage0(dst, j, sw, y0, y1, y2, y3);
            } else {
                ps v0, v1, v2, v3;
                fft_load_stage<INV>(src, tw, N, lg, st, Ns, j, sw, v0, v1, v2, v3);
                bfly4<INV>(v0, v1, v2, v3, y0, y1, y2, y3);
                const int k = j & (Ns - 1), j0 = ((j - k) << 2) + k;
                st_ps(dst + j0, y0);
                st_ps(dst + j0 + Ns, y1);
                st_ps(dst + j0 + 2 * Ns, y2);
                st_ps(dst + j0 + 3 * Ns, y3);
                if (wcp) {
                    if (j0 >= N - cp) st_ps(dst + j0 - N, y0);
                    if (j0 + Ns >= N - cp) st_ps(dst + j0 + Ns - N, y1);
                    if (j0 + 2 * Ns >= N - cp) st_ps(dst + j0 + 2 * Ns - N, y2);
                    if (j0 + 3 * Ns >= N - cp) st_ps(dst + j0 + 3 * Ns - N, y3);
                }
            }
        }
        Ns <<= 2;
        float4 *t = src; src = dst; dst = t;
    }
    if ((lg & 1) && !skip_last) {
        __syncthreads();
        const int h = N >> 1;
        for (int j = threadIdx.x; j < h; j += nthr) {
            const int k = j & (Ns - 1);
            const cx<float> w = tw[k];
            const ps v0 = ld_ps(src + j), v1 = mul_w(ld_ps(src + j + h), w.re, INV ? -w.im : w.im);
            const int j0 = ((j - k) << 1) + k;
            const ps y0 = v0 + v1, y1 = v0 - v1;
            st_ps(dst + j0, y0);
            st_ps(dst + j0 + Ns, y1);
            if (cp > 0) {
                if (j0 >= N - cp) st_ps(dst + j0 - N, y0);
                if (j0 + Ns >= N - cp) st_ps(dst + j0 + Ns - N, y1);
            }
        }
        float4 *t = src; src = dst; dst = t;
    }
    __syncthreads();
    return src;
}

// The last pass of a forward transform whose earlier passes left their result in `src` (skip_last): outputs
// of butterfly j, i.e. the bins j + u N/NU.  NU = 4: radix-4 pass (lg even), NU = 2: radix-2 pass (lg odd).
template <int NU>
__device__ __forceinline__ void fft_last_pass(const float4 *src, const cx<float> *tw, int N, int lg, int j, ps (&y)[NU]) {
    if constexpr (NU == 4) {
        const int st = (lg >> 1) - 1;
        ps v0, v1, v2, v3;
        fft_load_stage<false>(src, tw, N, lg, st, N >> 2, j, fft_sw(N, lg), v0, v1, v2, v3);
        bfly4<false>(v0, v1, v2, v3, y[0], y[1], y[2], y[3]);
    } else {
        const cx<float> w = tw[j];
        const ps v0 = ld_ps(src + j), v1 = mul_w(ld_ps(src + j + (N >> 1)), w.re, w.im);
        y[0] = v0 + v1;
        y[1] = v0 - v1;
    }
}
// whether the last pass of an N = 2^lg transform produces exactly the NU bins of a detection thread
__device__ __forceinline__ bool fft_last_fusable(int lg, int NU) {
    return lg >= 4 && ((NU == 4 && !(lg & 1)) || (NU == 2 && (lg & 1)));
}

// Unscaled Taylor moments of one (tap, rx, tx) item's rays about the expansion centre,
//   a_o = sum_rays cis(theta_ray) (j d_ray)^o / o!,   theta = psi + eps cos(phi),  d = wts cos(phi),
// entirely in float.  Valid in the slow-fading regime the host flags with OfdmP::cos_f32 (total phase
// advance |eps| < 0.05 rad, so theta needs no extended precision); phases are drawn in [0, 2 pi).
template <bool O3>
__device__ __forceinline__ void ray_moments_f32(cx<float> (&a)[4], const float *pphi, const float *ppsi, int first,
                                                int L, int step, int stride, float eps, float wts) {
    // psi / pi as a two-float product (hi + lo of 1/pi), reduced by the nearest half-integer k / 2 (exactly)
    // before the small terms are added: the phase reaches the polynomials with ~6e-8 half-turn error, like
    // the double path
    constexpr float kInvPi = 0.31830987334251404f, kInvPiLo = 1.284127663345183e-08f;
    const float eps_pi = eps * kInvPi;
#pragma unroll 2
    for (int o = first; o < L; o += step, pphi += step * stride, ppsi += step * stride) {
        // the Doppler factor cos(phi) only scales phase increments of < 0.05 rad over the whole frame, so the
        // SFU's 5e-7 absolute error moves a ray's phase by < 3e-8 rad: MUFU.COS on phi - pi in [-pi, pi)
        const float cphi = -__cosf(*pphi - 3.14159274f);
        const float d1 = wts * cphi, d2 = -0.5f * d1 * d1;
        const float psi = *ppsi, t = psi * kInvPi;
        float e = fmaf(psi, kInvPi, -t);
        e = fmaf(psi, kInvPiLo, e);
        e = fmaf(eps_pi, cphi, e);
        float sn, cs;
#ifdef B200_LIB_SINCOSPI
        sincospif((t - 2.0f * rintf(0.5f * t)) + e, &sn, &cs);
#else
        const float k = rintf(t + t);            // t in [0, 2): k in 0..4
        sincospi_kf(k, fmaf(k, -0.5f, t) + e, &sn, &cs);
#endif
        a[0].re += cs;                      a[0].im += sn;
        a[1].re = fmaf(-d1, sn, a[1].re);   a[1].im = fmaf(d1, cs, a[1].im);
        a[2].re = fmaf(d2, cs, a[2].re);    a[2].im = fmaf(d2, sn, a[2].im);
        if constexpr (O3) {
            const float d3 = (-1.0f / 3.0f) * d1 * d2;
            a[3].re = fmaf(d3, sn, a[3].re);
            a[3].im = fmaf(-d3, cs, a[3].im);
        }
    }
}

// QAMK: the modem kind is square Gray QAM at compile time (the slicer is inlined, the table-search / PSK paths
// are not even in the binary: ~10 % less code in a kernel whose hot path does not fit the instruction cache)
// LGF: log2(fft) at compile time for the full-band shapes the BASELINE configs use (used == fft; 10 or 11),
// 0 = run-time shape.  With the size known every FFT stage loop, stage-kind branch and bin -> position map
// folds to constants (one trip per loop, immediate offsets).
template <bool FUSED, int NR, int NT, bool QAMK, int KT = kOT, int LGF = 0, bool TC = false>
__global__ void __launch_bounds__(KT, (NR * NT <= 4) ? 3 : 1)
ofdm_tdl_pair_kernel(const __grid_constant__ OfdmP p, const Modem m_in, const cx<float> *__restrict__ tab_g,
                     uint64_t first_unit, long long n_units, const uint8_t *__restrict__ idx_g,
                     const float *__restrict__ phi_g, const float *__restrict__ psi_g,
                     const cx<float> *__restrict__ noise_g, uint8_t *__restrict__ idx_hat,
                     cx<float> *__restrict__ eq_out, unsigned long long *counters) {
    using T = float;
    constexpr int NP = NR / 2, TP = NT / 2;
    Modem m = m_in;
    if (QAMK) m.kind = B200PHY_MODEM_QAM;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int fft = LGF ? (1 << LGF) : p.fft, lg = LGF ? LGF : p.lg;
    const int used = LGF ? fft : p.used, half = LGF ? (fft >> 1) : p.half;
    const int S = p.S, mem = p.mem, cp = p.cp;

    unsigned char *sp = smem_raw;
    auto take = [&](size_t bytes) { unsigned char *r = sp; sp += (bytes + 15) & ~size_t(15); return r; };
    cx<T> *tw = (cx<T> *)take(sizeof(cx<T>) * (fft + kTwc));
    float4 *E2 = (float4 *)take(sizeof(float4) * (mem + S));            // one tx pair: [tail | cp | body]
    float4 *body = E2 + mem + cp;
    take(32);                                                           // TMA landing pad: a row that starts 8 B off lands 16 B early
    float4 *pool = (float4 *)take(sizeof(float4) * (NP + 1) * fft);     // rx pair buffers + 1 scratch
    float4 *gb2 = (float4 *)take(sizeof(float4) * p.n_taps * NT * NP);  // mean taps [class-sorted tap][t][rx pair]
    float4 *tails = (float4 *)take(p.n_sym > 1 ? sizeof(float4) * TP * mem : 0);
    u64 *coef = (u64 *)take(sizeof(u64) * p.n_taps * NP * 2 * 4 * 2);   // [tap][rx pair][t in pair][order][re|im]
    cx<T> *tab = (cx<T> *)take(sizeof(cx<T>) * m.M);
    uint8_t *dsym = (uint8_t *)take(NT * used);
    T *ph_phi = (T *)take(sizeof(T) * p.P4);
    T *ph_psi = (T *)take(sizeof(T) * p.P4);

    for (int i = tid; i < fft; i += KT) {
        double s, c;
        sincospi(-2.0 * double(i) / double(fft), &s, &c);
        tw[i] = {T(c), T(s)};
    }
    fill_compact_twiddles(tw, fft);
    if (m.kind != B200PHY_MODEM_BPSK)
        for (int k = tid; k < m.M; k += KT) tab[k] = tab_g[k];
    __shared__ __align__(8) unsigned long long mbar[3];      // [0] noise rows of a symbol, [1] phases of the next frame, [2] H_k MMAs
    __shared__ unsigned tmem_base_s;
    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        mbar_init(&mbar[2], 1);
        fence_mbar_init();
    }
    // Tensor-core H_k (2x2, fft 1024, <= 16 taps, one OFDM symbol per frame): the per-subcarrier channel matrices
    // H_k = sum_j gbar_j W^(k d_j) of a frame are ONE tcgen05 tile.  W^((k0 + off) d) = W^(k0 d) W^(off d) with
    // off = 128 v + 256 u (8 offsets): all 1024 bins share the DFT operand of the bins k0 < 128, the coefficients are
    // rotated by the 8th roots W^(off d) instead.  D[128, 64] = A[128, 32] B[32, 64]: A = (cos, sin) of W^(k0 d_j)
    // (constant: tf32 hi / lo parts resident in TMEM, 64 columns), B = rotated coefficients (per frame, hi / lo, K-major
    // in the tx sample buffer once the FIR has consumed it), D in TMEM (64 columns); 3xTF32 = 12 MMAs per frame issued
    // by one thread, completion on mbar[2]; thread tid reads back exactly the bins tid + 256 u it detects.
    // (template parameter TC; the launcher selects it: ofdm_tdl_pair_tc_ok)
    static_assert(!TC || (LGF == 10 && NR == 2 && NT == 2 && KT == 256), "tensor-core H_k: 2x2, fft 1024, 256 threads");
    constexpr bool kTcShape = TC;
    constexpr bool tc = TC;
    constexpr unsigned kTmemCols = 128;
    if constexpr (kTcShape) {
        if (tc && tid < 32) tmem_alloc(&tmem_base_s, kTmemCols);
        if (tc) tc_fence_before();
    }
    __syncthreads();
    if constexpr (kTcShape) {
        if (tc) tc_fence_after();
    }
    const unsigned tmem = tc ? tmem_base_s : 0u;
    unsigned par_h = 0;
    // B operand lives in the tx sample buffer (free between the FIR and the next frame's IFFT), 128-byte aligned;
    // its descriptor is frame-invariant: K-major, no swizzle, 128 B between the K chunks, 256 B between 8-row groups
    float *bop = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(E2) + 127) & ~uintptr_t(127));
    const unsigned long long bdesc0 = umma_smem_desc(bop, 128, 256);
    if constexpr (kTcShape) {
        if (tc) {
            if (tid < 128) {                                  // row k0 = tid of A: columns (2 j, 2 j + 1) = (cos, sin) of W^(k0 d_j)
                unsigned hi[32], lo[32];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    cx<T> w = {0.f, 0.f};
                    if (j < p.n_taps) w = tw[(tid * p.cls_delay[j]) & (fft - 1)];
                    split_tf32(w.re, hi[2 * j], lo[2 * j]);
                    split_tf32(w.im, hi[2 * j + 1], lo[2 * j + 1]);
                }
                const unsigned lane_base = unsigned(tid & ~31) << 16;
                tmem_st32(tmem + lane_base, hi);
                tmem_st32(tmem + 32 + lane_base, lo);
            }
            tc_fence_before();
            __syncthreads();
            tc_fence_after();
        }
    }

    unsigned sym_err = 0, bit_err = 0;
    const T sigma = T(p.sigma), tx_scale = T(p.tx_scale), rx_scale = T(p.rx_scale);
    const int n_items = p.n_taps * NR * 2;           // (tap, rx, t in pair)
    int G = 1;
    while (G < 16 && (G * 2) * n_items <= KT) G *= 2;
    const int sub = tid & (G - 1);
    const int ostride = p.n_taps * NR * NT;
    const double wts = p.w0 * p.Ts1, wt0 = p.w0 * p.t0;
    const bool in_w = p.ifft_in_w != 0;

    // Stream mode input pipeline (LDGSTS).  `pf`: the phases and data symbols of the NEXT frame are copied
    // into shared memory while this frame is in the FIR (one OFDM symbol per frame only: then the phase
    // buffers are free after the ray setup).  `apipe`: the raw noise rows of an rx pair are copied into that
    // pair's (still unused) accumulator buffer at frame start and merged in the FIR epilogue.
    // The data symbols of the next frame (8 B per thread and 2048 symbols) wait in registers instead.
    constexpr int kPre = (NR * NT <= 4) ? 1 : NT;
    const bool pf = !FUSED && p.n_sym == 1 && (p.n_data & 7) == 0 && p.n_data <= 8 * KT * kPre &&
                    (reinterpret_cast<uintptr_t>(idx_g) & 7) == 0;
    const bool pf16 = pf && (p.P & 3) == 0 && aligned16(phi_g) && aligned16(psi_g);
    const bool apipe = !FUSED;
    // rx FFT stage 0 straight from the FIR accumulators (one rx pair, one tx pair, one output block per frame)
    constexpr bool kFuseRx0 = (NP == 1 && TP == 1 && kJBC == 4);
    const bool rx0_fused = kFuseRx0 && fft == KT * kJBC;
    // TMA input pipeline (stream mode, the shapes whose FIR output leaves the landing zone alone): the two noise
    // rows of a symbol are two bulk copies into the rx pair buffer (rows back to back, not interleaved), the phases
    // of the next frame two more; one thread issues them, completion is counted on an mbarrier.
    // (cp >= 1 or a 16-byte aligned base: the copy of a row that starts 8 bytes off begins one element before it)
    const bool tma = !FUSED && rx0_fused && mem >= 1 && p.n_sym == 1 && (cp >= 1 || aligned16(noise_g));
    const bool tma_ph = tma && pf16;
    unsigned par_noise = 0, par_phase = 0;
    const cx<T> *nrow0 = nullptr, *nrow1 = nullptr;
    uint2 idx_pre[kPre];
    auto prefetch = [&](long long f) {
        const T *gp = phi_g + size_t(f) * p.P, *gq = psi_g + size_t(f) * p.P;
        if (tma_ph) {
            if (tid == 0) {
                mbar_expect_tx(&mbar[1], 2u * unsigned(p.P) * sizeof(T));
                bulk_g2s(ph_phi, gp, unsigned(p.P) * sizeof(T), &mbar[1]);
                bulk_g2s(ph_psi, gq, unsigned(p.P) * sizeof(T), &mbar[1]);
            }
        } else if (pf16) {
            for (int i = tid; i < (p.P >> 2); i += KT) {
                cp_async<16>(ph_phi + 4 * i, gp + 4 * i);
                cp_async<16>(ph_psi + 4 * i, gq + 4 * i);
            }
        } else {
            for (int i = tid; i < p.P; i += KT) {
                cp_async<4>(ph_phi + i, gp + i);
                cp_async<4>(ph_psi + i, gq + i);
            }
        }
#pragma unroll
        for (int u = 0; u < kPre; ++u)
            if (tid + u * KT < (p.n_data >> 3))
                idx_pre[u] = __ldg(reinterpret_cast<const uint2 *>(idx_g + size_t(f) * p.n_data) + tid + u * KT);
    };
    if constexpr (!FUSED) {
        if (pf && blockIdx.x < n_units) prefetch(blockIdx.x);
        cp_async_commit();
    }

    for (long long frame = blockIdx.x; frame < n_units; frame += gridDim.x) {
        const uint64_t unit = first_unit + uint64_t(frame);
        float4 *Yp[NP];
#pragma unroll
        for (int q = 0; q < NP; ++q) Yp[q] = pool + q * fft;
        float4 *W = pool + NP * fft;

        // ---- phases of all rays of this frame -> shared memory
        if constexpr (FUSED) {
            for (int b = tid; b < (p.P4 >> 2); b += KT) {
                const uint4 b1 = rng_block(p.seed, STREAM_CHANNEL, unit, uint64_t(b));
                const uint4 b2 = rng_block(p.seed, STREAM_CHANNEL, unit, uint64_t((p.P4 >> 2) + b));
                reinterpret_cast<float4 *>(ph_phi)[b] = make_float4(phase_from_word<T>(b1.x), phase_from_word<T>(b1.y),
                                                                    phase_from_word<T>(b1.z), phase_from_word<T>(b1.w));
                reinterpret_cast<float4 *>(ph_psi)[b] = make_float4(phase_from_word<T>(b2.x), phase_from_word<T>(b2.y),
                                                                    phase_from_word<T>(b2.z), phase_from_word<T>(b2.w));
            }
        } else if (pf) {
            if (tma_ph) { mbar_wait(&mbar[1], par_phase); par_phase ^= 1u; }
            else cp_async_wait<0>();                 // prefetched during the previous frame (visible after the barrier below)
#pragma unroll
            for (int u = 0; u < kPre; ++u)
                if (tid + u * KT < (p.n_data >> 3)) reinterpret_cast<uint2 *>(dsym)[tid + u * KT] = idx_pre[u];
        } else {
            const T *gp = phi_g + size_t(frame) * p.P, *gq = psi_g + size_t(frame) * p.P;
            for (int i0 = tid; i0 < p.P; i0 += 4 * KT) {
                T a[4], b[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (i0 + u * KT < p.P) { a[u] = __ldg(gp + i0 + u * KT); b[u] = __ldg(gq + i0 + u * KT); }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (i0 + u * KT < p.P) { ph_phi[i0 + u * KT] = a[u]; ph_psi[i0 + u * KT] = b[u]; }
            }
        }

        for (int s = 0; s < p.n_sym; ++s) {
            const int n_s = s * S;
            // ---------------- P0: data symbols, noise into the rx pair buffers
            {
                const int w0 = s * used * NT, cnt = used * NT;
                if constexpr (FUSED) {
                    const int b0 = w0 >> 2, b1 = (w0 + cnt - 1) >> 2;
                    if (((w0 | cnt) & 3) == 0) {
                        // whole blocks: the four symbols of a Philox block are one 32-bit store
                        const int sh = 32 - m.bits;
                        for (int b = b0 + tid; b <= b1; b += KT) {
                            const uint4 blk = rng_block(p.seed, STREAM_DATA, unit, uint64_t(b));
                            reinterpret_cast<uint32_t *>(dsym)[b - b0] =
                                (blk.x >> sh) | ((blk.y >> sh) << 8) | ((blk.z >> sh) << 16) | ((blk.w >> sh) << 24);
                        }
                    } else {
                        for (int b = b0 + tid; b <= b1; b += KT) {
                            const uint4 blk = rng_block(p.seed, STREAM_DATA, unit, uint64_t(b));
#pragma unroll
                            for (int l = 0; l < 4; ++l) {
                                const int w = 4 * b + l - w0;
                                if (w >= 0 && w < cnt) dsym[w] = uint8_t(lane_of(blk, l) >> (32 - m.bits));
                            }
                        }
                    }
                } else if (!pf) {
                    const uint8_t *src = idx_g + frame * p.n_data + w0;
                    if ((cnt & 3) == 0 && ((frame * p.n_data + w0) & 3) == 0) {
                        const uint32_t *s4 = reinterpret_cast<const uint32_t *>(src);
                        uint32_t *d4 = reinterpret_cast<uint32_t *>(dsym);
                        for (int i = tid; i < (cnt >> 2); i += KT) d4[i] = __ldg(s4 + i);
                    } else {
                        for (int i = tid; i < cnt; i += KT) dsym[i] = src[i];
                    }
                }
                const int m0 = n_s + cp;
                if constexpr (FUSED) {
                    // one thread draws sample pair (j, j + 1) of BOTH antennas of an rx pair: two independent Philox
                    // chains per iteration and whole pair samples (one 16-byte store each) instead of scalar stores
                    const int pr0 = m0 >> 1, npr = ((m0 + fft - 1) >> 1) - pr0 + 1;
#pragma unroll
                    for (int q = 0; q < NP; ++q)
                        for (int i = tid; i < npr; i += KT) {
                            const int pr = pr0 + i, j = 2 * pr - m0;
                            const uint4 b0 = rng_block(p.seed, STREAM_NOISE, unit, uint64_t(2 * q) * (p.row >> 1) + pr);
                            const uint4 b1 = rng_block(p.seed, STREAM_NOISE, unit, uint64_t(2 * q + 1) * (p.row >> 1) + pr);
                            // unit-variance normals, exactly what the draw kernel stores: sigma is applied in the FIR
                            // epilogue by the same single FFMA2 as in stream mode (bit-identical results in both modes)
                            if (j >= 0 && j < fft) {
                                const cx<T> c0 = cnormal<T>(b0.x, b0.y), c1 = cnormal<T>(b1.x, b1.y);
                                Yp[q][j] = make_float4(c0.re, c1.re, c0.im, c1.im);
                            }
                            if (j + 1 >= 0 && j + 1 < fft) {
                                const cx<T> c0 = cnormal<T>(b0.z, b0.w), c1 = cnormal<T>(b1.z, b1.w);
                                Yp[q][j + 1] = make_float4(c0.re, c1.re, c0.im, c1.im);
                            }
                        }
                } else if (tma) {
                    // rows (2 q, 2 q + 1) of this symbol as two bulk copies.  A row that starts 8 bytes off a 16-byte
                    // boundary is copied from one element earlier; sizes are rounded up to 16 bytes (the extra element
                    // is still inside the row: mem >= 1).  The zone starts 32 bytes before the pair buffer.
                    const size_t rowlen = size_t(p.N + mem);
                    const cx<T> *s0 = noise_g + (size_t(frame) * NR) * rowlen + m0, *s1 = s0 + rowlen;
                    const int sh0 = int((reinterpret_cast<uintptr_t>(s0) >> 3) & 1), sh1 = int((reinterpret_cast<uintptr_t>(s1) >> 3) & 1);
                    const int c0 = (fft + sh0 + 1) & ~1, c1 = (fft + sh1 + 1) & ~1;
                    cx<T> *land = reinterpret_cast<cx<T> *>(Yp[0]) - 4;
                    nrow0 = land + sh0;
                    nrow1 = land + c0 + sh1;
                    if (tid == 0) {
                        fence_proxy_async();             // the zone was last written by ordinary stores (previous FFT)
                        mbar_expect_tx(&mbar[0], unsigned(c0 + c1) * sizeof(cx<T>));
                        bulk_g2s(land, s0 - sh0, unsigned(c0) * sizeof(cx<T>), &mbar[0]);
                        bulk_g2s(land + c0, s1 - sh1, unsigned(c1) * sizeof(cx<T>), &mbar[0]);
                    }
                } else if (apipe) {
                    // raw noise of rx 2q / 2q+1 at sample j lands in the two halves of the float4 slot j of the pair
                    // buffer (n0.re, n0.im, n1.re, n1.im); the FIR epilogue turns each slot into pair layout in
                    // place.  A thread consumes only slots it copied itself, so its own wait_group is enough.
                    const size_t rowlen = size_t(p.N + mem);
#pragma unroll
                    for (int q = 0; q < NP; ++q) {
                        const cx<T> *s0 = noise_g + (size_t(frame) * NR + 2 * q) * rowlen + m0, *s1 = s0 + rowlen;
                        cx<T> *raw = reinterpret_cast<cx<T> *>(Yp[q]);
                        for (int j = tid; j < fft; j += KT) {
                            cp_async<8>(raw + 2 * j, s0 + j);
                            cp_async<8>(raw + 2 * j + 1, s1 + j);
                        }
                    }
                    cp_async_commit();
                } else {
                    const size_t rowlen = size_t(p.N + mem);
#pragma unroll
                    for (int q = 0; q < NP; ++q) {
                        const cx<T> *s0 = noise_g + (size_t(frame) * NR + 2 * q) * rowlen + m0, *s1 = s0 + rowlen;
                        for (int j0 = tid; j0 < fft; j0 += 4 * KT) {
                            cx<T> v0[4], v1[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                if (j0 + u * KT < fft) { v0[u] = load_stream(s0 + j0 + u * KT); v1[u] = load_stream(s1 + j0 + u * KT); }
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                if (j0 + u * KT < fft)
                                    Yp[q][j0 + u * KT] = make_float4(sigma * v0[u].re, sigma * v1[u].re, sigma * v0[u].im, sigma * v1[u].im);
                        }
                    }
                }
            }
            __syncthreads();

            for (int tp = 0; tp < TP; ++tp) {
                // ---------------- A: map + scatter for both antennas of the tx pair, fused with IFFT stage 0: the
                // thread that maps bins j + i fft/4 owns butterfly j of the first (twiddle-free) radix-4 stage, so
                // the mapped symbols never touch shared memory.  Also the ISI tail of the previous symbol.
                float4 *in = in_w ? W : body;
                float4 *other = in_w ? body : W;
                for (int j = tid; j < (fft >> 2); j += KT) {
                    ps v[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int q = pos_of(j + i * (fft >> 2), fft, used, half);
                        v[i] = {0ull, 0ull};
                        if (q >= 0) {
                            // the two antennas' symbols of a subcarrier are adjacent (and NT is even): one 16-bit load
                            const unsigned two = *reinterpret_cast<const uint16_t *>(dsym + q * NT + 2 * tp);
                            const cx<T> s0 = map_symbol<T>(m, tab, int(two & 0xffu));
                            const cx<T> s1 = map_symbol<T>(m, tab, int(two >> 8));
                            v[i] = {pk2(tx_scale * s0.re, tx_scale * s1.re), pk2(tx_scale * s0.im, tx_scale * s1.im)};
                        }
                    }
                    ps y0, y1, y2, y3;
                    bfly4<true>(v[0], v[1], v[2], v[3], y0, y1, y2, y3);
                    fft_store_stage0(other, j, true, y0, y1, y2, y3);
                }
                for (int i = tid; i < mem; i += KT)
                    E2[i] = (s > 0) ? tails[tp * mem + i] : make_float4(0.f, 0.f, 0.f, 0.f);
                // ---------------- C: ray setup, items (tap, rx, t in pair), G lanes per item
                for (int it0 = 0; it0 < n_items; it0 += KT / G) {
                    const int it = it0 + tid / G;
                    const bool act = it < n_items;
                    const int l = act ? it / (NR * 2) : 0;
                    const int rem = act ? it - l * NR * 2 : 0;
                    const int r = rem >> 1, tt = rem & 1, t = 2 * tp + tt;
                    const T amp = T(p.amp[l]);
                    cx<T> a0 = {0.f, 0.f}, a1 = a0, a2 = a0, a3 = a0;
                    const double cseg = double(n_s + cp - p.delays[l]) + 0.5 * double(fft - 1);
                    if (act && p.cos_f32) {
                        const T *pphi = ph_phi + ((l * NR + r) * NT + t) + sub * ostride;
                        const T *ppsi = ph_psi + ((l * NR + r) * NT + t) + sub * ostride;
                        cx<T> a[4] = {a0, a0, a0, a0};
                        const float eps = float(fma(wts, cseg, wt0)), wtsf = float(wts);
                        if (p.porder == 3) ray_moments_f32<true>(a, pphi, ppsi, sub, p.L, G, ostride, eps, wtsf);
                        else ray_moments_f32<false>(a, pphi, ppsi, sub, p.L, G, ostride, eps, wtsf);
                        a0 = amp * a[0]; a1 = amp * a[1]; a2 = amp * a[2]; a3 = amp * a[3];
                    } else if (act) {
                        const T *pphi = ph_phi + ((l * NR + r) * NT + t) + sub * ostride;
                        const T *ppsi = ph_psi + ((l * NR + r) * NT + t) + sub * ostride;
                        for (int o = sub; o < p.L; o += G, pphi += G * ostride, ppsi += G * ostride) {
                            const double cphi = p.cos_f32 ? double(cosf(*pphi)) : cos(double(*pphi));
                            const double dl = wts * cphi;
                            T sn, cs;
                            cis_phase<T>(fma(dl, cseg, fma(wt0, cphi, double(*ppsi))), &sn, &cs);
                            const cx<T> e = {amp * cs, amp * sn};
                            const T d1 = T(dl), d2 = -0.5f * d1 * d1, d3 = (-1.0f / 3.0f) * d1 * d2;
                            a0.re += e.re;        a0.im += e.im;
                            a1.re -= d1 * e.im;   a1.im += d1 * e.re;
                            a2.re += d2 * e.re;   a2.im += d2 * e.im;
                            a3.re += d3 * e.im;   a3.im -= d3 * e.re;
                        }
                    }
                    for (int o = G >> 1; o > 0; o >>= 1) {
                        a0.re += __shfl_xor_sync(0xffffffffu, a0.re, o); a0.im += __shfl_xor_sync(0xffffffffu, a0.im, o);
                        a1.re += __shfl_xor_sync(0xffffffffu, a1.re, o); a1.im += __shfl_xor_sync(0xffffffffu, a1.im, o);
                        a2.re += __shfl_xor_sync(0xffffffffu, a2.re, o); a2.im += __shfl_xor_sync(0xffffffffu, a2.im, o);
                        a3.re += __shfl_xor_sync(0xffffffffu, a3.re, o); a3.im += __shfl_xor_sync(0xffffffffu, a3.im, o);
                    }
                    if (act && sub == 0) {
                        if (p.porder != 3) a3 = {0.f, 0.f};
                        // [(tap, rx pair, tt)][order][re|im][lane = rx & 1]
                        T *cq = reinterpret_cast<T *>(coef) + (((l * NP + (r >> 1)) * 2 + tt) * 4) * 4 + (r & 1);
                        cq[0] = a0.re; cq[2] = a0.im; cq[4] = a1.re; cq[6] = a1.im;
                        cq[8] = a2.re; cq[10] = a2.im; cq[12] = a3.re; cq[14] = a3.im;
                        const T m1 = T(p.mu[0][l]), m2 = T(p.mu[1][l]), m3 = T(p.mu[2][l]);
                        T *gq = reinterpret_cast<T *>(gb2 + (p.cls_pos[l] * NT + t) * NP + (r >> 1)) + (r & 1);
                        gq[0] = a0.re + m1 * a1.re + m2 * a2.re + m3 * a3.re;
                        gq[2] = a0.im + m1 * a1.im + m2 * a2.im + m3 * a3.im;
                    }
                }
                // ---------------- B: remaining IFFT passes (end in E2.body); the last one also writes the cyclic prefix
                fft_stockham_pair<true, LGF ? KT : 0>(other, in, tw, fft, lg, 1, false, cp);
                if constexpr (!FUSED) {
                    if (tp == TP - 1) {              // last ray setup done: the phase buffers are free
                        if (pf && frame + gridDim.x < n_units) prefetch(frame + gridDim.x);
                        cp_async_commit();
                    }
                }

                // ---------------- D: FIR for both tx antennas of the pair into the rx pair buffers
                {
                    const float tau0 = float(tid) - 0.5f * float(fft - 1);
                    const float4 *xb = E2 + mem + cp + tid;
                    for (int jo0 = 0; jo0 < fft; jo0 += KT * kJBC) {
                        u64 aRe[kJBC][NP], aIm[kJBC][NP];
                        float tauv[kJBC];
#pragma unroll
                        for (int jb = 0; jb < kJBC; ++jb) {
                            tauv[jb] = tau0 + float(jo0 + jb * KT);
#pragma unroll
                            for (int q = 0; q < NP; ++q) { aRe[jb][q] = 0ull; aIm[jb][q] = 0ull; }
                        }
                        // the whole tap loop is specialised on the polynomial order (one uniform branch per
                        // block of outputs instead of one per tap, which cost a register shuffle at every join)
                        auto taps = [&](auto order3) {
                            constexpr bool O3 = decltype(order3)::value;
                            constexpr int NO = O3 ? 4 : 3;
                            for (int l = 0; l < p.n_taps; ++l) {
                                const float4 *xl = xb + (jo0 - p.delays[l]);
                                float4 x4[kJBC];
#pragma unroll
                                for (int jb = 0; jb < kJBC; ++jb) x4[jb] = xl[jb * KT];
#pragma unroll
                                for (int tt = 0; tt < 2; ++tt) {
                                    u64 cR[NP][NO], cI[NP][NO];
#pragma unroll
                                    for (int q = 0; q < NP; ++q)
#pragma unroll
                                        for (int o = 0; o < NO; ++o) {
                                            const ulonglong2 c = reinterpret_cast<const ulonglong2 *>(coef)[((l * NP + q) * 2 + tt) * 4 + o];
                                            cR[q][o] = c.x; cI[q][o] = c.y;
                                        }
#pragma unroll
                                    for (int jb = 0; jb < kJBC; ++jb) {
                                        const u64 tt2 = pk2(tauv[jb], tauv[jb]);
                                        const float xr = tt ? x4[jb].y : x4[jb].x, xi = tt ? x4[jb].w : x4[jb].z;
                                        const u64 xrr = pk2(xr, xr), xii = pk2(xi, xi), nxii = pk2(-xi, -xi);
#pragma unroll
                                        for (int q = 0; q < NP; ++q) {
                                            u64 gR = cR[q][NO - 1], gI = cI[q][NO - 1];
#pragma unroll
                                            for (int o = NO - 2; o >= 0; --o) {
                                                gR = fma2(gR, tt2, cR[q][o]);
                                                gI = fma2(gI, tt2, cI[q][o]);
                                            }
                                            aRe[jb][q] = fma2(gR, xrr, aRe[jb][q]);
                                            aRe[jb][q] = fma2(gI, nxii, aRe[jb][q]);
                                            aIm[jb][q] = fma2(gR, xii, aIm[jb][q]);
                                            aIm[jb][q] = fma2(gI, xrr, aIm[jb][q]);
                                        }
                                    }
                                }
                            }
                        };
                        if (p.porder == 3) taps(std::true_type{}); else taps(std::false_type{});
                        ps yv[kJBC][NP];
                        if (!FUSED && tma) {
                            // the two rows have landed (byte count complete on the barrier): y = sigma * noise + FIR
                            mbar_wait(&mbar[0], par_noise);
                            par_noise ^= 1u;
                            const u64 sg = pk2(sigma, sigma);
#pragma unroll
                            for (int jb = 0; jb < kJBC; ++jb) {
                                const cx<T> n0 = nrow0[tid + jo0 + jb * KT], n1 = nrow1[tid + jo0 + jb * KT];
                                yv[jb][0].re = fma2(pk2(n0.re, n1.re), sg, aRe[jb][0]);
                                yv[jb][0].im = fma2(pk2(n0.im, n1.im), sg, aIm[jb][0]);
                            }
                        } else if (!FUSED && apipe && tp == 0) {
                            // the raw noise has landed in this thread's slots: y = sigma * noise + FIR, re-laid as pairs
                            if (TP == 1) cp_async_wait<1>(); else cp_async_wait<0>();
                            const u64 sg = pk2(sigma, sigma);
#pragma unroll
                            for (int jb = 0; jb < kJBC; ++jb)
#pragma unroll
                                for (int q = 0; q < NP; ++q) {
                                    const float4 v = Yp[q][tid + jo0 + jb * KT];       // (n0.re, n0.im, n1.re, n1.im)
                                    // one fused multiply-add, as in every other mode: bit-identical results
                                    yv[jb][q].re = fma2(pk2(v.x, v.z), sg, aRe[jb][q]);
                                    yv[jb][q].im = fma2(pk2(v.y, v.w), sg, aIm[jb][q]);
                                }
                        } else if (FUSED && tp == 0) {
                            // fused RNG: the buffer holds this symbol's unit-variance normals (pair layout)
                            const u64 sg = pk2(sigma, sigma);
#pragma unroll
                            for (int jb = 0; jb < kJBC; ++jb)
#pragma unroll
                                for (int q = 0; q < NP; ++q) {
                                    const ps y = ld_ps(Yp[q] + tid + jo0 + jb * KT);
                                    yv[jb][q].re = fma2(y.re, sg, aRe[jb][q]);
                                    yv[jb][q].im = fma2(y.im, sg, aIm[jb][q]);
                                }
                        } else {
#pragma unroll
                            for (int jb = 0; jb < kJBC; ++jb)
#pragma unroll
                                for (int q = 0; q < NP; ++q) {
                                    const ps y = ld_ps(Yp[q] + tid + jo0 + jb * KT);
                                    yv[jb][q].re = add2(y.re, aRe[jb][q]);
                                    yv[jb][q].im = add2(y.im, aIm[jb][q]);
                                }
                        }
                        if (rx0_fused) {
                            // one output block per frame and one rx pair: this thread holds the four inputs
                            // tid + i fft/4 of butterfly `tid` of the rx FFT's first stage -> straight into W
                            if constexpr (kFuseRx0) {
                                ps y0, y1, y2, y3;
                                bfly4<false>(yv[0][0], yv[1][0], yv[2][0], yv[3][0], y0, y1, y2, y3);
                                fft_store_stage0(W, tid, true, y0, y1, y2, y3);
                            }
                        } else {
#pragma unroll
                            for (int jb = 0; jb < kJBC; ++jb)
#pragma unroll
                                for (int q = 0; q < NP; ++q) st_ps(Yp[q] + tid + jo0 + jb * KT, yv[jb][q]);
                        }
                    }
                }
                __syncthreads();
                if (p.n_sym > 1) {
                    for (int i = tid; i < mem; i += KT) tails[tp * mem + i] = E2[S + i];
                    __syncthreads();
                }
            }   // tx pairs

            if constexpr (kTcShape) {
                // B operand, element (kk = 2 j + p, n = 8 r8 + 2 e + q), e = 2 t + rx, g' = gbar_j[e] W^(off_r8 d_j):
                //   p = 0: (g'.re, g'.im)[q];  p = 1: (-g'.im, g'.re)[q].   Layout [hi|lo][k-step][n / 8][k-chunk][n % 8][4 k].
                // One item per thread: (r8, e, chunk c8) = rows n, n + 1 of one 16-byte K chunk (taps 2 c8, 2 c8 + 1).
                {
                    const int e = tid & 3, c8 = (tid >> 2) & 7, r8 = tid >> 5;     // a quarter-warp stores 4 rows x 2 chunks: 2-way conflicts at worst
                    const int off = 128 * (r8 >> 2) + 256 * (r8 & 3);
                    cx<T> g[2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int j = 2 * c8 + h;
                        g[h] = {0.f, 0.f};
                        if (j < p.n_taps) {
                            const float4 gg = gb2[(j * NT + (e >> 1)) * NP];            // (rx0.re, rx1.re, rx0.im, rx1.im)
                            const cx<T> c = (e & 1) ? cx<T>{gg.y, gg.w} : cx<T>{gg.x, gg.z};
                            const cx<T> w = tw[(off * p.cls_delay[j]) & (fft - 1)];
                            // explicit roundings: the fused-RNG and the stream instantiation must agree bit for bit
                            g[h] = {__fmaf_rn(c.re, w.re, -__fmul_rn(c.im, w.im)), __fmaf_rn(c.re, w.im, __fmul_rn(c.im, w.re))};
                        }
                    }
                    unsigned r0h, r0l, i0h, i0l, r1h, r1l, i1h, i1l;
                    split_tf32(g[0].re, r0h, r0l); split_tf32(g[0].im, i0h, i0l);
                    split_tf32(g[1].re, r1h, r1l); split_tf32(g[1].im, i1h, i1l);
                    constexpr unsigned kNeg = 0x80000000u;
                    // float4 index of (k-step c8 / 2, n-group r8, chunk c8 % 2, row 2 e) in one part; im row = + 1; lo part = + 512
                    uint4 *b4 = reinterpret_cast<uint4 *>(bop) + (((c8 >> 1) * 8 + r8) * 2 + (c8 & 1)) * 8 + 2 * e;
                    b4[0] = make_uint4(r0h, i0h ^ kNeg, r1h, i1h ^ kNeg);             // re row: (g'.re, -g'.im) of taps j, j + 1
                    b4[1] = make_uint4(i0h, r0h, i1h, r1h);                           // im row: (g'.im,  g'.re)
                    b4[512] = make_uint4(r0l, i0l ^ kNeg, r1l, i1l ^ kNeg);
                    b4[513] = make_uint4(i0l, r0l, i1l, r1l);
                }
                fence_proxy_async();
                tc_fence_before();
                __syncthreads();
                if (tid == 0) {
                    tc_fence_after();
                    constexpr unsigned idesc = umma_idesc_tf32(128, 64);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const unsigned long long bhi = bdesc0 + (unsigned long long)(ks * 128), blo = bhi + 512ull;     // + bytes / 16
                        umma_tf32_ts(tmem + 64, tmem + 32 + 8 * ks, bhi, idesc, ks ? 1u : 0u);       // A_lo B_hi
                        umma_tf32_ts(tmem + 64, tmem + 8 * ks, blo, idesc, 1u);                      // A_hi B_lo
                        umma_tf32_ts(tmem + 64, tmem + 8 * ks, bhi, idesc, 1u);                      // A_hi B_hi
                    }
                    umma_commit(&mbar[2]);
                }
            }

            // ---------------- F: paired FFT of every rx pair (rotating pool).  When the last pass produces exactly
            // the bins a detection thread owns, it is left to the detection phase (fft_last_pass, from registers).
#ifndef B200_PAIR_BIG_NU
#define B200_PAIR_BIG_NU 2
#endif
            constexpr int NU = (NR * NT <= 4) ? 4 : B200_PAIR_BIG_NU;
            const bool fuse_last = fft_last_fusable(lg, NU);
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                float4 *res = rx0_fused ? fft_stockham_pair<false, LGF ? KT : 0>(W, Yp[q], tw, fft, lg, 1, fuse_last)
                                        : fft_stockham_pair<false, LGF ? KT : 0>(Yp[q], W, tw, fft, lg, 0, fuse_last);
                if (res != Yp[q]) { W = Yp[q]; Yp[q] = res; }
            }

            // ---------------- G: H_k, detect, demap, count.  H_k = sum_l gbar_l W^{k d_l} on rx-pair lanes (FFMA2):
            // a thread owns the NU bins k0 + u fft/NU; W^{(k0 + u fft/NU) d} = W^{k0 d} e^{-2 pi j u d / NU}, so the
            // taps are summed per residue class d mod NU (host-sorted runs in the class order 0, 2, 1, 3:
            // OfdmP::cls_*, so that the classes mod 2 are contiguous too) and the class sums are combined by an
            // NU-point DFT.  NU = 4 when the 4 x NT x NP packed accumulators fit the register budget, else 2.
            const int kstride = fft / NU;
            uint8_t *hat_fs = idx_hat ? idx_hat + size_t(frame) * p.n_data + size_t(s) * used * NT : nullptr;
            cx<T> *eq_fs = eq_out ? eq_out + size_t(frame) * p.n_data + size_t(s) * used * NT : nullptr;
            const bool hat_vec = (reinterpret_cast<uintptr_t>(hat_fs) & (NT - 1)) == 0;
            for (int k0 = tid; k0 < kstride; k0 += KT) {
                ps Hc[NU][NT][NP];
#pragma unroll
                for (int u = 0; u < NU; ++u)
#pragma unroll
                    for (int t = 0; t < NT; ++t)
#pragma unroll
                        for (int q = 0; q < NP; ++q) Hc[u][t][q] = {0ull, 0ull};
                auto tap_sum = [&](ps (&acc)[NT][NP], int j0, int j1) {
#if defined(B200_HK_UNROLL1)
#pragma unroll 1
#elif defined(B200_HK_UNROLL2)
#pragma unroll 2
#elif defined(B200_HK_UNROLL4)
#pragma unroll 4
#endif
                    for (int j = j0; j < j1; ++j) {
                        const cx<T> w = tw[(k0 * p.cls_delay[j]) & (fft - 1)];
                        const u64 WR = pk2(w.re, w.re), WI = pk2(w.im, w.im), NWI = pk2(-w.im, -w.im);
#pragma unroll
                        for (int t = 0; t < NT; ++t)
#pragma unroll
                            for (int q = 0; q < NP; ++q) {
                                const ps g = ld_ps(gb2 + (j * NT + t) * NP + q);
                                acc[t][q].re = fma2(g.im, NWI, fma2(g.re, WR, acc[t][q].re));
                                acc[t][q].im = fma2(g.im, WR, fma2(g.re, WI, acc[t][q].im));
                            }
                    }
                };
                bool hk_done = false;
                if constexpr (kTcShape) {
                    if (tc) {
                        mbar_wait(&mbar[2], par_h);          // the 12 MMAs of this frame have completed
                        par_h ^= 1u;
                        tc_fence_after();
                        unsigned v[32];                      // bins tid + 256 u: columns 32 (tid >> 7) + 8 u + 4 t + 2 rx + (re|im)
                        tmem_ld32(tmem + 64 + (unsigned(tid & 96) << 16) + 32 * (tid >> 7), v);
#pragma unroll
                        for (int u = 0; u < 4; ++u)
#pragma unroll
                            for (int t = 0; t < 2; ++t)
                                Hc[u][t][0] = {pk2(__uint_as_float(v[8 * u + 4 * t]), __uint_as_float(v[8 * u + 4 * t + 2])),
                                               pk2(__uint_as_float(v[8 * u + 4 * t + 1]), __uint_as_float(v[8 * u + 4 * t + 3]))};
                        hk_done = true;
                    }
                }
                if (hk_done) {
                } else if constexpr (NU == 4) {
                    tap_sum(Hc[0], p.cls_start[0], p.cls_start[1]);
                    tap_sum(Hc[2], p.cls_start[1], p.cls_start[2]);
                    tap_sum(Hc[1], p.cls_start[2], p.cls_start[3]);
                    tap_sum(Hc[3], p.cls_start[3], p.cls_start[4]);
#pragma unroll
                    for (int t = 0; t < NT; ++t)
#pragma unroll
                        for (int q = 0; q < NP; ++q) {
                            const ps a0 = Hc[0][t][q] + Hc[2][t][q], a1 = Hc[0][t][q] - Hc[2][t][q];
                            const ps a2 = Hc[1][t][q] + Hc[3][t][q], a3 = Hc[1][t][q] - Hc[3][t][q];
                            Hc[0][t][q] = a0 + a2;
                            Hc[1][t][q] = {add2(a1.re, a3.im), sub2(a1.im, a3.re)};      // a1 - j a3
                            Hc[2][t][q] = a0 - a2;
                            Hc[3][t][q] = {sub2(a1.re, a3.im), add2(a1.im, a3.re)};      // a1 + j a3
                        }
                } else if constexpr (NU == 1) {
                    tap_sum(Hc[0], p.cls_start[0], p.cls_start[4]);
                } else {
                    tap_sum(Hc[0], p.cls_start[0], p.cls_start[2]);          // even delays
                    tap_sum(Hc[1], p.cls_start[2], p.cls_start[4]);          // odd delays
#pragma unroll
                    for (int t = 0; t < NT; ++t)
#pragma unroll
                        for (int q = 0; q < NP; ++q) {
                            const ps e = Hc[0][t][q], o = Hc[1][t][q];
                            Hc[0][t][q] = e + o;
                            Hc[1][t][q] = e - o;
                        }
                }
                // received bins of the NU subcarriers: the last FFT pass from registers, or plain loads
                ps Yv[NP][NU];
#pragma unroll
                for (int qq = 0; qq < NP; ++qq) {
                    if (NU >= 2 && fuse_last) {
                        if constexpr (NU >= 2) fft_last_pass<NU>(Yp[qq], tw, fft, lg, k0, Yv[qq]);
                    } else {
#pragma unroll
                        for (int u = 0; u < NU; ++u) Yv[qq][u] = ld_ps(Yp[qq] + k0 + u * kstride);
                    }
                }
                // detection of the NU bins: branch-free arithmetic first (independent double chains of the bins
                // overlap), then demap / count / store under the bin's validity predicate
                cx<T> z[NU][NT];
                int qv[NU];
#pragma unroll
                for (int u = 0; u < NU; ++u) {
                    const int k = k0 + u * kstride;
                    qv[u] = pos_of(k, fft, used, half);
                    cx<T> H[NR][NT], y[NR];
#pragma unroll
                    for (int qq = 0; qq < NP; ++qq) {
#pragma unroll
                        for (int t = 0; t < NT; ++t) {
                            upk2(Hc[u][t][qq].re, H[2 * qq][t].re, H[2 * qq + 1][t].re);
                            upk2(Hc[u][t][qq].im, H[2 * qq][t].im, H[2 * qq + 1][t].im);
                        }
                        float yr0, yr1, yi0, yi1;
                        upk2(Yv[qq][u].re, yr0, yr1);
                        upk2(Yv[qq][u].im, yi0, yi1);
                        y[2 * qq] = {rx_scale * yr0, rx_scale * yi0};
                        y[2 * qq + 1] = {rx_scale * yr1, rx_scale * yi1};
                    }
                    if constexpr (NT == 2) {
                        // closed-form 2x2 (H^H H + s2 I)^-1 H^H y in double; 1/det by one Newton step on the float
                        // reciprocal (relative error 4e-15) instead of a double division
                        double a = p.fnv, bb = p.fnv;
                        cx<double> c = {0.0, 0.0}, v0 = {0.0, 0.0}, v1 = {0.0, 0.0};
#pragma unroll
                        for (int r = 0; r < NR; ++r) {
                            const cx<double> h0 = cvt<double>(H[r][0]), h1 = cvt<double>(H[r][1]), yr = cvt<double>(y[r]);
                            a += norm2(h0);
                            bb += norm2(h1);
                            cmac_conj(c, h1, h0);
                            cmac_conj(v0, h0, yr);
                            cmac_conj(v1, h1, yr);
                        }
                        const double det = a * bb - norm2(c);
                        // MUFU.RCP seed (2^-23 relative) + one Newton step in double: 1e-14 relative
                        float rdf;
                        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rdf) : "f"(float(det)));
                        double rd = double(rdf);
                        rd = rd * fma(-det, rd, 2.0);
                        const double g = rd * p.snt;
                        z[u][0] = {T(g * (bb * v0.re - (c.re * v1.re + c.im * v1.im))), T(g * (bb * v0.im - (c.re * v1.im - c.im * v1.re)))};
                        z[u][1] = {T(g * (a * v1.re - (c.re * v0.re - c.im * v0.im))), T(g * (a * v1.im - (c.re * v0.im + c.im * v0.re)))};
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
                        for (int t = 0; t < NT; ++t) z[u][t] = {T(b[t].re * p.snt), T(b[t].im * p.snt)};
                    }
                }
#pragma unroll
                for (int u = 0; u < NU; ++u) {
                    const int q = qv[u];
                    if (q < 0) continue;
                    // the NT symbols of this subcarrier are adjacent (Blast layer interleave): one packed compare / store
                    unsigned av, ev = 0;
                    if constexpr (NT == 2) av = *reinterpret_cast<const uint16_t *>(dsym + q * NT);
                    else av = *reinterpret_cast<const uint32_t *>(dsym + q * NT);
#pragma unroll
                    for (int t = 0; t < NT; ++t) ev |= unsigned(demap_symbol<T>(m, tab, z[u][t])) << (8 * t);
                    const unsigned x = ev ^ av;
                    bit_err += __popc(x);
#pragma unroll
                    for (int t = 0; t < NT; ++t) sym_err += ((x >> (8 * t)) & 0xffu) != 0u;
                    if (hat_fs) {
                        if (hat_vec) {
                            if constexpr (NT == 2) *reinterpret_cast<uint16_t *>(hat_fs + q * NT) = uint16_t(ev);
                            else *reinterpret_cast<uint32_t *>(hat_fs + q * NT) = ev;
                        } else {
#pragma unroll
                            for (int t = 0; t < NT; ++t) hat_fs[q * NT + t] = uint8_t(ev >> (8 * t));
                        }
                    }
                    if (eq_fs) {
#pragma unroll
                        for (int t = 0; t < NT; ++t) eq_fs[q * NT + t] = z[u][t];
                    }
                    if (p.rx_out) {
                        // slow path (parity tests): the demodulated rx samples of this bin, re-derived from the packed bins
#pragma unroll
                        for (int qq = 0; qq < NP; ++qq) {
                            float yr0, yr1, yi0, yi1;
                            upk2(Yv[qq][u].re, yr0, yr1);
                            upk2(Yv[qq][u].im, yi0, yi1);
                            cx<T> *ro = static_cast<cx<T> *>(p.rx_out) + (size_t(frame) * NR + 2 * qq) * (size_t(p.n_sym) * used) + size_t(s) * used + q;
                            ro[0] = {rx_scale * yr0, rx_scale * yi0};
                            ro[size_t(p.n_sym) * used] = {rx_scale * yr1, rx_scale * yi1};
                        }
                    }
                }
            }
            if (tma) fence_proxy_async();        // this frame's ordinary stores before the next frame's bulk copies
            if constexpr (kTcShape) {
                if (tc) tc_fence_before();       // this frame's TMEM reads before the next frame's MMAs
            }
            __syncthreads();
        }   // OFDM symbols
    }       // frames
    if constexpr (kTcShape) {
        if (tc) {
            tc_fence_before();
            __syncthreads();
            if (tid < 32) tmem_dealloc(tmem, kTmemCols);
        }
    }
    flush_counters(sym_err, bit_err, counters);
    if (blockIdx.x == 0 && tid == 0) {
        atomicAdd(&counters[2], (unsigned long long)n_units * p.n_data);
        atomicAdd(&counters[3], (unsigned long long)n_units * p.n_data * m.bits);
    }
}

inline size_t ofdm_tdl_pair_smem(const OfdmP &p, int M, int NR, int NT) {
    auto al = [](size_t b) { return (b + 15) & ~size_t(15); };
    const int NP = NR / 2, TP = NT / 2;
    size_t s = 0;
    s += al(sizeof(cx<float>) * (p.fft + kTwc));
    s += al(sizeof(float4) * (p.mem + p.S));
    s += 32 + al(sizeof(float4) * (NP + 1) * p.fft);
    s += al(sizeof(float4) * p.n_taps * NT * NP);
    s += al(p.n_sym > 1 ? sizeof(float4) * TP * p.mem : 0);
    s += al(sizeof(u64) * p.n_taps * NP * 2 * 4 * 2);
    s += al(sizeof(cx<float>) * M);
    s += al(size_t(NT) * p.used);
    s += 2 * al(sizeof(float) * p.P4);
    return s;
}

// tensor-core H_k (template parameter TC of the 2x2 / fft 1024 instantiation): at most 16 taps (K = 32), one OFDM symbol
// per frame (the tx sample buffer is free after the FIR) and room for the 16 KB coefficient operand in it
inline bool ofdm_tdl_pair_tc_ok(const OfdmP &p) {
    return p.tc_hk && p.n_taps <= 16 && p.n_sym == 1 && size_t(p.mem + p.S) * sizeof(float4) >= 16384 + 128;
}

// the pair kernel applies when: float, even Nr/Nt, POLY with one segment, fft % 1024 == 0, mean taps
// from the polynomial, and the caller has not disabled it (b200phy_ofdm_tdl_params.reserved bit 0)
inline bool ofdm_tdl_pair_ok(const OfdmP &p, int NR, int NT) {
    return (NR % 2 == 0) && (NT % 2 == 0) && p.poly && p.nseg == 1 && (p.fft & (kOT * kJBC - 1)) == 0 &&
           p.gbar_poly && !p.no_pair;
}

}  // namespace b200phy

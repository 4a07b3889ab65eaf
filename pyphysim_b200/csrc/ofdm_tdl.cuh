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
//   POLY        2nd/3rd-order Taylor polynomial in the output-sample offset tau around a segment
//               centre (per tap the expansion point is shifted by its delay so tau is common to all
//               taps), used when sqrt(L)*(w*tau_max)^(P+1)/(P+1)! is below the dtype's resolution
//               (slow fading: all BASELINE configs); 4-6 FMAs per tap sample instead of 6*L.
#pragma once
#include "common.cuh"
#include "linalg.cuh"
#include "rng.cuh"

namespace b200phy {

#ifndef B200_OFDM_MINB
#define B200_OFDM_MINB 4       /* A/B on B200: 4 CTAs/SM (64 regs) is 2-3 % ahead of 3 CTAs (80 regs) */
#endif
constexpr int kOT = 256;       // threads per CTA
#ifndef B200_OFDM_JBC
#define B200_OFDM_JBC 4
#endif
constexpr int kJBC = B200_OFDM_JBC;   // outputs per thread per register block in the channel apply
constexpr int kCH = 8;         // recurrence chunk length

struct OfdmP {
    int fft, lg, cp, used, half, n_sym, S, N, mem, n_taps, L, n_data;
    int poly, nseg, seg_len, seg_lg;
    int porder;     // Taylor order of the POLY mode (2 or 3)
    int gbar_poly;  // 1: symbol-mean taps from the polynomial moments mu (nseg == 1 only)
    int cos_f32;    // 1: cos(phi) may be evaluated in float (total phase advance is small)
    int cgrp;       // lanes cooperating on one (tap, rx) item in the oscillator setup (power of 2)
    int no_pair;    // 1: do not use the antenna-pair FFMA2 kernel (A/B and tests)
    int tc_hk;      // 1: per-subcarrier channel matrices of the 2x2 / fft 1024 pair kernel through tcgen05 (3xTF32)
    int row;        // noise normals per rx row in the Philox layout: 2*ceil((N+mem)/2)
    int P, P4;      // phases per frame, rounded up to a multiple of 4
    int ifft_in_w;  // 1: scatter into W so that the ping-pong IFFT ends in E.body
    int delays[B200PHY_MAX_TAPS];
    // taps grouped by delay residue class d mod 4 (the H_k evaluation of the pair kernels sums each class
    // separately and combines the sums with a 4- or 2-point DFT): run i = class (0, 2, 1, 3)[i] occupies sorted
    // positions [cls_start[i], cls_start[i+1]); cls_pos[l] = sorted position of tap l, cls_delay[j] = its delay
    int cls_start[5], cls_pos[B200PHY_MAX_TAPS], cls_delay[B200PHY_MAX_TAPS];
    double amp[B200PHY_MAX_TAPS];   // sqrt(P_l / L)
    double mu[3][B200PHY_MAX_TAPS]; // mean of tau^p (p = 1..3) over the S samples of a symbol, per tap
    double w0;      // 2*pi*Fd
    double Ts1;     // Ts * 1.0000000001 (fading_generators.py:462)
    double t0;
    double sigma, fnv, tx_scale, rx_scale, snt;
    PhiloxKey seed;              // round keys of the 64-bit seed (rng.cuh)
    void *rx_out;   // optional: demodulated rx samples before detection, complex[n][Nr][n_sym*used] (parity tests)
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
                const int ts = k << (lg - 2 - 2 * st);          // k * (N/4) / Ns
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
            cx<T> w = tw[k];                                     // last pass: Ns == N/2
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

// cis(th) for a phase formed in double: reduce in turns (exact rint), then sincospi in T precision
__device__ __forceinline__ void sincospi_t(float x, float *s, float *c) { sincospif(x, s, c); }
__device__ __forceinline__ void sincospi_t(double x, double *s, double *c) { sincospi(x, s, c); }
template <typename T> __device__ __forceinline__ void cis_phase(double th, T *s, T *c) {
    const double t = th * 0.15915494309189533577;
    sincospi_t(T(2.0 * (t - rint(t))), s, c);
}

__device__ __forceinline__ void sincos_t(float x, float *s, float *c) { sincosf(x, s, c); }
__device__ __forceinline__ void sincos_t(double x, double *s, double *c) { sincos(x, s, c); }

// read-once stream data: 8/16-byte loads through the read-only path
template <typename T> __device__ __forceinline__ cx<T> load_stream(const cx<T> *p) {
    if constexpr (sizeof(T) == 4) {
        const float2 v = __ldg(reinterpret_cast<const float2 *>(p));
        return {v.x, v.y};
    } else {
        const double2 v = __ldg(reinterpret_cast<const double2 *>(p));
        return {v.x, v.y};
    }
}

// asynchronous global -> shared copies (LDGSTS): the input pipeline of the float pair kernels.  Groups
// complete in commit order; cp_async_wait<N>() returns when at most N groups are still in flight.
template <int BYTES> __device__ __forceinline__ void cp_async(void *smem, const void *gmem) {
    const unsigned s = unsigned(__cvta_generic_to_shared(smem));
    if constexpr (BYTES == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(s), "l"(gmem), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// TMA bulk copies (cp.async.bulk, SASS UBLKCP) global -> shared, completion counted in bytes on an mbarrier: one
// elected thread arms the barrier with the byte count and issues the copies, every consumer polls the barrier's
// phase parity.  Source, destination and size must be multiples of 16 bytes.
__device__ __forceinline__ unsigned smem_u32(const void *p) { return unsigned(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(void *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(void *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "B200_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra B200_DONE_%=;\n"
        "bra B200_WAIT_%=;\n"
        "B200_DONE_%=:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// ---- tcgen05 (5th-generation tensor cores): TMEM allocation, operand staging, MMA issue, read-back.
// Shared-memory operand descriptor, K-major, no swizzle: 8-row x 16-byte core matrices; lbo = bytes between the two
// 16-byte K chunks of one MMA, sbo = bytes between 8-row groups (cute::UMMA::SmemDescriptor, version 1).
__device__ __forceinline__ unsigned long long umma_smem_desc(const void *p, unsigned lbo, unsigned sbo) {
    return (unsigned long long)((smem_u32(p) >> 4) & 0x3fffu) | ((unsigned long long)((lbo >> 4) & 0x3fffu) << 16) |
           ((unsigned long long)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, dense, M x N
__host__ __device__ constexpr unsigned umma_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (unsigned(N >> 3) << 17) | (unsigned(M >> 4) << 24);
}
// D[tmem] (+)= A[tmem] * B[smem]; issued by ONE thread
__device__ __forceinline__ void umma_tf32_ts(unsigned d_tmem, unsigned a_tmem, unsigned long long b_desc, unsigned idesc,
                                             unsigned accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit(void *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(unsigned *smem_result, unsigned cols) {          // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(unsigned addr, unsigned cols) {               // the same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// 32 consecutive 32-bit columns of this thread's TMEM lane (a warp reaches lanes 32 (warp % 4) .. + 31)
__device__ __forceinline__ void tmem_st32(unsigned addr, const unsigned (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
        "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
          "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
          "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
          "r"(v[30]), "r"(v[31]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld32(unsigned addr, unsigned (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,"
        "%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(addr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// x = hi + lo, hi representable in tf32 (the MMA ignores the low 13 bits of lo)
__device__ __forceinline__ void split_tf32(float x, unsigned &hi, unsigned &lo) {
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
    lo = __float_as_uint(x - __uint_as_float(hi));
}

__device__ __forceinline__ void bulk_g2s(void *smem, const void *gmem, unsigned bytes, void *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem)),
                 "l"(gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

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
// Horner evaluation of the tap polynomial at tau and accumulation of g*x, P = 2 or 3
template <typename T, int P>
__device__ __forceinline__ void tap_mac(cx<T> &acc, const cx<T> (&cf)[4], T tau, cx<T> x) {
    cx<T> g;
    if (P == 3) {
        g.re = fma(cf[3].re, tau, cf[2].re);
        g.im = fma(cf[3].im, tau, cf[2].im);
        g.re = fma(g.re, tau, cf[1].re);
        g.im = fma(g.im, tau, cf[1].im);
    } else {
        g.re = fma(cf[2].re, tau, cf[1].re);
        g.im = fma(cf[2].im, tau, cf[1].im);
    }
    g.re = fma(g.re, tau, cf[0].re);
    g.im = fma(g.im, tau, cf[0].im);
    cmac(acc, g, x);
}

// (packed FP32 pair helpers pk2 / upk2 / fma2 / add2 / sub2 / mul2: common.cuh)

// S[C] += gbar_l[r][t] * w for the residue class C = d_l mod 4 of a tap
template <typename T, int NR, int NT, int C>
__device__ __forceinline__ void hk_class(cx<T> (&S)[4][NR][NT], const cx<T> *__restrict__ gl, cx<T> w) {
#pragma unroll
    for (int r = 0; r < NR; ++r)
#pragma unroll
        for (int t = 0; t < NT; ++t) cmac(S[C][r][t], gl[r * NT + t], w);
}

template <typename T, bool FUSED, int NR, int NT, bool WSG>
__global__ void __launch_bounds__(kOT, (sizeof(T) == 4 && NR * NT <= 4) ? B200_OFDM_MINB : 1)
ofdm_tdl_kernel(const __grid_constant__ OfdmP p, const Modem m, const cx<T> *__restrict__ tab_g,
                uint64_t first_unit, long long n_units, const uint8_t *__restrict__ idx_g,
                const T *__restrict__ phi_g, const T *__restrict__ psi_g,
                const cx<T> *__restrict__ noise_g, uint8_t *__restrict__ idx_hat,
                cx<T> *__restrict__ eq_out, cx<T> *__restrict__ ws_g, unsigned long long *counters) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int fft = p.fft, S = p.S, mem = p.mem, cp = p.cp;

    // ---- carve shared memory (mirrored by ofdm_tdl_smem below)
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
    T *ph_phi = (T *)take(sizeof(T) * p.P4);                  // Jakes phases of the current frame
    T *ph_psi = (T *)take(sizeof(T) * p.P4);
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
    const int G = p.cgrp, sub = tid & (G - 1);
    // FIR variants: fastD = one segment and fft a multiple of 1024 (branch-free); in float the inner
    // loop is issued as FFMA2 pairs over two rx antennas (kPackR) or over (re, im) (kPackC, Nr = 1)
    constexpr bool kPackR = (sizeof(T) == 4) && (NR % 2 == 0);
    constexpr bool kPackC = (sizeof(T) == 4) && (NR == 1);
    const bool fastD = p.poly && p.nseg == 1 && (fft & (kOT * kJBC - 1)) == 0;

    for (long long frame = blockIdx.x; frame < n_units; frame += gridDim.x) {
        const uint64_t unit = first_unit + uint64_t(frame);
        cx<T> *Yp[NR];
#pragma unroll
        for (int r = 0; r < NR; ++r) Yp[r] = pool + r * fft;
        cx<T> *W = pool + NR * fft;

        // ---------------- phases of all rays of this frame -> shared memory (coalesced / one Philox
        // block per 4 phases); ordered before their first use by the barrier after P0
        if constexpr (FUSED) {
            for (int b = tid; b < (p.P4 >> 2); b += kOT) {
                const uint4 b1 = rng_block(p.seed, STREAM_CHANNEL, unit, uint64_t(b));
                const uint4 b2 = rng_block(p.seed, STREAM_CHANNEL, unit, uint64_t((p.P4 >> 2) + b));
#pragma unroll
                for (int l = 0; l < 4; ++l) {
                    ph_phi[4 * b + l] = phase_from_word<T>(lane_of(b1, l));
                    ph_psi[4 * b + l] = phase_from_word<T>(lane_of(b2, l));
                }
            }
        } else {
            const T *gp = phi_g + size_t(frame) * p.P, *gq = psi_g + size_t(frame) * p.P;
            for (int i0 = tid; i0 < p.P; i0 += 4 * kOT) {
                T a[4], b[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (i0 + u * kOT < p.P) { a[u] = __ldg(gp + i0 + u * kOT); b[u] = __ldg(gq + i0 + u * kOT); }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (i0 + u * kOT < p.P) { ph_phi[i0 + u * kOT] = a[u]; ph_psi[i0 + u * kOT] = b[u]; }
            }
            // pull the next frame of this CTA towards L2 while this one is being computed
            const long long nf = frame + gridDim.x;
            if (nf < n_units) {
                const char *q0 = reinterpret_cast<const char *>(noise_g + size_t(nf) * NR * size_t(p.N + mem));
                const size_t nb = sizeof(cx<T>) * NR * size_t(p.N + mem);
                for (size_t o = size_t(tid) * 128; o < nb; o += size_t(kOT) * 128)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(q0 + o));
                const char *q1 = reinterpret_cast<const char *>(phi_g + size_t(nf) * p.P);
                const char *q2 = reinterpret_cast<const char *>(psi_g + size_t(nf) * p.P);
                for (size_t o = size_t(tid) * 128; o < sizeof(T) * p.P; o += size_t(kOT) * 128) {
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(q1 + o));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(q2 + o));
                }
                const char *q3 = reinterpret_cast<const char *>(idx_g + size_t(nf) * p.n_data);
                for (size_t o = size_t(tid) * 128; o < size_t(p.n_data); o += size_t(kOT) * 128)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(q3 + o));
            }
        }

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
                    if ((cnt & 3) == 0 && ((frame * p.n_data + w0) & 3) == 0) {
                        const uint32_t *s4 = reinterpret_cast<const uint32_t *>(src);
                        uint32_t *d4 = reinterpret_cast<uint32_t *>(dsym);
                        for (int i = tid; i < (cnt >> 2); i += kOT) d4[i] = __ldg(s4 + i);
                    } else {
                        for (int i = tid; i < cnt; i += kOT) dsym[i] = src[i];
                    }
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
                        // batches of 4 independent loads per thread so one DRAM latency covers them all
                        for (int j0 = tid; j0 < fft; j0 += 4 * kOT) {
                            cx<T> v[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                if (j0 + u * kOT < fft) v[u] = load_stream(src + j0 + u * kOT);
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                if (j0 + u * kOT < fft) Yp[r][j0 + u * kOT] = sigma * v[u];
                        }
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
                // ---------------- C: per-ray setup for tx antenna t.  G lanes share one (tap, rx) item,
                // each summing a subset of the L rays; a G-wide shuffle reduction finishes the item.
                if (fastD && p.gbar_poly) {
                    // common case (slow fading, one segment): flag-free loop, strength-reduced indices
                    const int ostride = p.n_taps * NR * NT;
                    const double wts = p.w0 * p.Ts1, wt0 = p.w0 * p.t0;
                    for (int it0 = 0; it0 < n_items; it0 += kOT / G) {
                        const int it = it0 + tid / G;
                        const bool act = it < n_items;
                        const int l = act ? it / NR : 0, r = act ? it - l * NR : 0;
                        const T amp = T(p.amp[l]);
                        cx<T> a0 = {T(0), T(0)}, a1 = a0, a2 = a0, a3 = a0;
                        const double cseg = double(n_s + cp - p.delays[l]) + 0.5 * double(fft - 1);
                        if (act) {
                            const T *pphi = ph_phi + ((l * NR + r) * NT + t) + sub * ostride;
                            const T *ppsi = ph_psi + ((l * NR + r) * NT + t) + sub * ostride;
                            for (int o = sub; o < p.L; o += G, pphi += G * ostride, ppsi += G * ostride) {
                                const double cphi = (sizeof(T) == 4 && p.cos_f32) ? double(cosf(float(*pphi)))
                                                                                  : cos(double(*pphi));
                                const double dl = wts * cphi;
                                T sn, cs;
                                cis_phase<T>(fma(dl, cseg, fma(wt0, cphi, double(*ppsi))), &sn, &cs);
                                const cx<T> e = {amp * cs, amp * sn};
                                const T d1 = T(dl), d2 = T(-0.5) * d1 * d1, d3 = T(-1.0 / 3.0) * d1 * d2;
                                a0.re += e.re;        a0.im += e.im;
                                a1.re -= d1 * e.im;   a1.im += d1 * e.re;     // (j dl) e
                                a2.re += d2 * e.re;   a2.im += d2 * e.im;     // -(dl^2/2) e
                                a3.re += d3 * e.im;   a3.im -= d3 * e.re;     // -j(dl^3/6) e
                            }
                        }
                        for (int o = G >> 1; o > 0; o >>= 1) {
                            a0.re += __shfl_xor_sync(0xffffffffu, a0.re, o); a0.im += __shfl_xor_sync(0xffffffffu, a0.im, o);
                            a1.re += __shfl_xor_sync(0xffffffffu, a1.re, o); a1.im += __shfl_xor_sync(0xffffffffu, a1.im, o);
                            a2.re += __shfl_xor_sync(0xffffffffu, a2.re, o); a2.im += __shfl_xor_sync(0xffffffffu, a2.im, o);
                            a3.re += __shfl_xor_sync(0xffffffffu, a3.re, o); a3.im += __shfl_xor_sync(0xffffffffu, a3.im, o);
                        }
                        if (act && sub == 0) {
                            if (p.porder != 3) a3 = {T(0), T(0)};
                            if constexpr (kPackR) {
                                T *cq = reinterpret_cast<T *>(coef) + ((l * (NR / 2) + (r >> 1)) * 4) * 4 + (r & 1);
                                cq[0] = a0.re; cq[2] = a0.im; cq[4] = a1.re; cq[6] = a1.im;
                                cq[8] = a2.re; cq[10] = a2.im; cq[12] = a3.re; cq[14] = a3.im;
                            } else {
                                cx<T> *c4 = coef + (l * NR + r) * 4;
                                c4[0] = a0; c4[1] = a1; c4[2] = a2; c4[3] = a3;
                            }
                            const T m1 = T(p.mu[0][l]), m2 = T(p.mu[1][l]), m3 = T(p.mu[2][l]);
                            gbar[(l * NR + r) * NT + t] = {a0.re + m1 * a1.re + m2 * a2.re + m3 * a3.re,
                                                           a0.im + m1 * a1.im + m2 * a2.im + m3 * a3.im};
                        }
                    }
                } else
                for (int sg = 0; sg < (p.poly ? p.nseg : 1); ++sg) {
                    for (int it0 = 0; it0 < n_items; it0 += kOT / G) {
                        const int it = it0 + tid / G;
                        const bool act = it < n_items;
                        const int l = act ? it / NR : 0, r = act ? it % NR : 0;
                        const T amp = T(p.amp[l]);
                        cx<T> a0 = {T(0), T(0)}, a1 = a0, a2 = a0, a3 = a0, gs = a0;
                        const double cseg = double(n_s + cp + sg * p.seg_len - p.delays[l]) + 0.5 * double(p.seg_len - 1);
                        if (act)
                            for (int o = sub; o < p.L; o += G) {
                                const int i = ((o * p.n_taps + l) * NR + r) * NT + t;
                                const T phi = ph_phi[i], psi = ph_psi[i];
                                const double cphi = (sizeof(T) == 4 && p.cos_f32) ? double(cosf(float(phi))) : cos(double(phi));
                                const double dl = p.w0 * p.Ts1 * cphi;             // phase step per sample
                                const double th0 = fma(p.w0 * cphi, p.t0, double(psi));
                                if (p.poly) {
                                    T sn, cs;
                                    cis_phase<T>(fma(dl, cseg, th0), &sn, &cs);
                                    const cx<T> e = {amp * cs, amp * sn};
                                    const T d1 = T(dl), d2 = T(-0.5 * dl * dl);
                                    a0.re += e.re;        a0.im += e.im;
                                    a1.re -= d1 * e.im;   a1.im += d1 * e.re;     // (j dl) e
                                    a2.re += d2 * e.re;   a2.im += d2 * e.im;     // -(dl^2/2) e
                                    if (p.porder == 3) {
                                        const T d3 = T(dl * dl * dl * (1.0 / 6.0));
                                        a3.re += d3 * e.im;   a3.im -= d3 * e.re; // -j(dl^3/6) e
                                    }
                                }
                                if (sg == 0 && !p.gbar_poly) {
                                    // mean tap over the S samples of this symbol (CP included, ofdm.py:541-548)
                                    T sn, cs;
                                    cis_phase<T>(fma(dl, double(n_s) + 0.5 * double(S - 1), th0), &sn, &cs);
                                    const T g = amp * dirichlet<T>(dl, S);
                                    gs.re += g * cs;
                                    gs.im += g * sn;
                                }
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
                        for (int o = G >> 1; o > 0; o >>= 1) {
                            if (p.poly) {
                                a0.re += __shfl_xor_sync(0xffffffffu, a0.re, o); a0.im += __shfl_xor_sync(0xffffffffu, a0.im, o);
                                a1.re += __shfl_xor_sync(0xffffffffu, a1.re, o); a1.im += __shfl_xor_sync(0xffffffffu, a1.im, o);
                                a2.re += __shfl_xor_sync(0xffffffffu, a2.re, o); a2.im += __shfl_xor_sync(0xffffffffu, a2.im, o);
                                if (p.porder == 3) { a3.re += __shfl_xor_sync(0xffffffffu, a3.re, o); a3.im += __shfl_xor_sync(0xffffffffu, a3.im, o); }
                            }
                            if (sg == 0 && !p.gbar_poly) { gs.re += __shfl_xor_sync(0xffffffffu, gs.re, o); gs.im += __shfl_xor_sync(0xffffffffu, gs.im, o); }
                        }
                        if (act && sub == 0) {
                            if (p.poly) {
                                if constexpr (kPackR) {
                                    if (fastD) {
                                        // rx-pair layout for the FFMA2 FIR: [(tap, pair)][order][re|im][lane]
                                        T *cq = reinterpret_cast<T *>(coef) + ((l * (NR / 2) + (r >> 1)) * 4) * 4 + (r & 1);
                                        cq[0] = a0.re; cq[2] = a0.im; cq[4] = a1.re; cq[6] = a1.im;
                                        cq[8] = a2.re; cq[10] = a2.im; cq[12] = a3.re; cq[14] = a3.im;
                                    } else {
                                        cx<T> *c4 = coef + ((l * NR + r) * p.nseg + sg) * 4;
                                        c4[0] = a0; c4[1] = a1; c4[2] = a2; c4[3] = a3;
                                    }
                                } else {
                                    cx<T> *c4 = coef + ((l * NR + r) * p.nseg + sg) * 4;
                                    c4[0] = a0; c4[1] = a1; c4[2] = a2; c4[3] = a3;
                                }
                            }
                            if (sg == 0) {
                                if (p.gbar_poly) {
                                    const T m1 = T(p.mu[0][l]), m2 = T(p.mu[1][l]), m3 = T(p.mu[2][l]);
                                    gs.re = a0.re + m1 * a1.re + m2 * a2.re + m3 * a3.re;
                                    gs.im = a0.im + m1 * a1.im + m2 * a2.im + m3 * a3.im;
                                }
                                gbar[(l * NR + r) * NT + t] = gs;
                            }
                        }
                    }
                }
                // ---------------- B: IFFT (result lands in E.body), cyclic prefix, ISI tail
                fft_stockham<T, true>(in, other, tw, fft, p.lg);
                for (int i = tid; i < cp; i += kOT) E[mem + i] = body[fft - cp + i];
                for (int i = tid; i < mem; i += kOT)
                    E[i] = (s > 0) ? tails[t * mem + i] : mk<T>(T(0), T(0));
                __syncthreads();

                // ---------------- D: time-varying sparse FIR, accumulate into Y[r]
                if (fastD && kPackR) {
                    if constexpr (kPackR) {
                        // FFMA2 over rx pairs: re/im accumulators hold (rx 2p, rx 2p+1)
                        constexpr int NP = NR / 2;
                        const float tau0 = float(tid) - 0.5f * float(fft - 1);
                        const float2 *xb = reinterpret_cast<const float2 *>(E) + mem + cp + tid;
                        const u64 *cq0 = reinterpret_cast<const u64 *>(coef);
                        for (int jo0 = 0; jo0 < fft; jo0 += kOT * kJBC) {
                            u64 aRe[kJBC][NP], aIm[kJBC][NP], tt[kJBC];
#pragma unroll
                            for (int jb = 0; jb < kJBC; ++jb) {
                                const float tau = tau0 + float(jo0 + jb * kOT);
                                tt[jb] = pk2(tau, tau);
#pragma unroll
                                for (int q = 0; q < NP; ++q) { aRe[jb][q] = 0ull; aIm[jb][q] = 0ull; }
                            }
#pragma unroll 3
                            for (int l = 0; l < p.n_taps; ++l) {
                                const float2 *xl = xb + (jo0 - p.delays[l]);
                                u64 cR[NP][4], cI[NP][4];
#pragma unroll
                                for (int q = 0; q < NP; ++q)
#pragma unroll
                                    for (int o = 0; o < 4; ++o) {
                                        cR[q][o] = cq0[((l * NP + q) * 4 + o) * 2];
                                        cI[q][o] = cq0[((l * NP + q) * 4 + o) * 2 + 1];
                                    }
#pragma unroll
                                for (int jb = 0; jb < kJBC; ++jb) {
                                    const float2 x = xl[jb * kOT];
                                    const u64 xrr = pk2(x.x, x.x), xii = pk2(x.y, x.y), nxii = pk2(-x.y, -x.y);
#pragma unroll
                                    for (int q = 0; q < NP; ++q) {
                                        u64 gR, gI;
                                        if (p.porder == 3) {
                                            gR = fma2(cR[q][3], tt[jb], cR[q][2]); gI = fma2(cI[q][3], tt[jb], cI[q][2]);
                                            gR = fma2(gR, tt[jb], cR[q][1]);       gI = fma2(gI, tt[jb], cI[q][1]);
                                        } else {
                                            gR = fma2(cR[q][2], tt[jb], cR[q][1]); gI = fma2(cI[q][2], tt[jb], cI[q][1]);
                                        }
                                        gR = fma2(gR, tt[jb], cR[q][0]);
                                        gI = fma2(gI, tt[jb], cI[q][0]);
                                        aRe[jb][q] = fma2(gR, xrr, aRe[jb][q]);
                                        aRe[jb][q] = fma2(gI, nxii, aRe[jb][q]);
                                        aIm[jb][q] = fma2(gR, xii, aIm[jb][q]);
                                        aIm[jb][q] = fma2(gI, xrr, aIm[jb][q]);
                                    }
                                }
                            }
#pragma unroll
                            for (int jb = 0; jb < kJBC; ++jb) {
                                const int j = tid + jo0 + jb * kOT;
#pragma unroll
                                for (int q = 0; q < NP; ++q) {
                                    float r0, r1, i0, i1;
                                    upk2(aRe[jb][q], r0, r1);
                                    upk2(aIm[jb][q], i0, i1);
                                    cx<T> *y0 = Yp[2 * q], *y1 = Yp[2 * q + 1];
                                    y0[j] = y0[j] + mk<T>(T(r0), T(i0));
                                    y1[j] = y1[j] + mk<T>(T(r1), T(i1));
                                }
                            }
                        }
                    }
                } else if (fastD && kPackC) {
                    if constexpr (kPackC) {
                        // Nr = 1: FFMA2 over (re, im); g.re / g.im enter as broadcast scalars
                        const float tau0 = float(tid) - 0.5f * float(fft - 1);
                        const float2 *xb = reinterpret_cast<const float2 *>(E) + mem + cp + tid;
                        const u64 *cq0 = reinterpret_cast<const u64 *>(coef);
                        for (int jo0 = 0; jo0 < fft; jo0 += kOT * kJBC) {
                            u64 acc2[kJBC], tt[kJBC];
#pragma unroll
                            for (int jb = 0; jb < kJBC; ++jb) {
                                const float tau = tau0 + float(jo0 + jb * kOT);
                                tt[jb] = pk2(tau, tau);
                                acc2[jb] = 0ull;
                            }
#pragma unroll 3
                            for (int l = 0; l < p.n_taps; ++l) {
                                const float2 *xl = xb + (jo0 - p.delays[l]);
                                const u64 c0 = cq0[l * 4], c1 = cq0[l * 4 + 1], c2 = cq0[l * 4 + 2], c3 = cq0[l * 4 + 3];
#pragma unroll
                                for (int jb = 0; jb < kJBC; ++jb) {
                                    const float2 x = xl[jb * kOT];
                                    u64 g = (p.porder == 3) ? fma2(fma2(c3, tt[jb], c2), tt[jb], c1) : fma2(c2, tt[jb], c1);
                                    g = fma2(g, tt[jb], c0);
                                    float gre, gim;
                                    upk2(g, gre, gim);
                                    acc2[jb] = fma2(pk2(x.x, x.y), pk2(gre, gre), acc2[jb]);
                                    acc2[jb] = fma2(pk2(-x.y, x.x), pk2(gim, gim), acc2[jb]);
                                }
                            }
#pragma unroll
                            for (int jb = 0; jb < kJBC; ++jb) {
                                const int j = tid + jo0 + jb * kOT;
                                float re, im;
                                upk2(acc2[jb], re, im);
                                Yp[0][j] = Yp[0][j] + mk<T>(T(re), T(im));
                            }
                        }
                    }
                } else if (fastD) {
                    // fast path (one segment, fft a multiple of 1024): no per-sample branches
                    const T tau0 = T(tid) - T(0.5) * T(fft - 1);
                    const cx<T> *xb = E + mem + cp + tid;
                    for (int jo0 = 0; jo0 < fft; jo0 += kOT * kJBC) {
                        cx<T> acc[kJBC][NR];
                        T tau[kJBC];
#pragma unroll
                        for (int jb = 0; jb < kJBC; ++jb) {
                            tau[jb] = tau0 + T(jo0 + jb * kOT);
#pragma unroll
                            for (int r = 0; r < NR; ++r) acc[jb][r] = {T(0), T(0)};
                        }
                        if (p.porder == 3) {
                            for (int l = 0; l < p.n_taps; ++l) {
                                const cx<T> *xl = xb + (jo0 - p.delays[l]);
                                cx<T> cf[NR][4];
#pragma unroll
                                for (int r = 0; r < NR; ++r)
#pragma unroll
                                    for (int o = 0; o < 4; ++o) cf[r][o] = coef[(l * NR + r) * 4 + o];
#pragma unroll
                                for (int jb = 0; jb < kJBC; ++jb) {
                                    const cx<T> x = xl[jb * kOT];
#pragma unroll
                                    for (int r = 0; r < NR; ++r) tap_mac<T, 3>(acc[jb][r], cf[r], tau[jb], x);
                                }
                            }
                        } else {
#pragma unroll 3
                            for (int l = 0; l < p.n_taps; ++l) {
                                const cx<T> *xl = xb + (jo0 - p.delays[l]);
                                cx<T> cf[NR][4];
#pragma unroll
                                for (int r = 0; r < NR; ++r)
#pragma unroll
                                    for (int o = 0; o < 3; ++o) cf[r][o] = coef[(l * NR + r) * 4 + o];
#pragma unroll
                                for (int jb = 0; jb < kJBC; ++jb) {
                                    const cx<T> x = xl[jb * kOT];
#pragma unroll
                                    for (int r = 0; r < NR; ++r) tap_mac<T, 2>(acc[jb][r], cf[r], tau[jb], x);
                                }
                            }
                        }
#pragma unroll
                        for (int jb = 0; jb < kJBC; ++jb) {
                            const int j = tid + jo0 + jb * kOT;
#pragma unroll
                            for (int r = 0; r < NR; ++r) Yp[r][j] = Yp[r][j] + acc[jb][r];
                        }
                    }
                } else if (p.poly) {
                    const T tau0 = T(tid) - T(0.5) * T(p.seg_len - 1);
                    const cx<T> *xb = E + mem + cp + tid;
                    for (int jb0 = 0; jb0 * kOT < fft; jb0 += kJBC) {
                        cx<T> acc[kJBC][NR];
                        T tau[kJBC];
#pragma unroll
                        for (int jb = 0; jb < kJBC; ++jb) {
                            const int j = tid + (jb0 + jb) * kOT;
                            tau[jb] = tau0 + T(((jb0 + jb) * kOT) & (p.seg_len - 1));
                            (void)j;
#pragma unroll
                            for (int r = 0; r < NR; ++r) acc[jb][r] = {T(0), T(0)};
                        }
                        for (int l = 0; l < p.n_taps; ++l) {
                            const cx<T> *xl = xb - p.delays[l];
                            cx<T> cf[NR][4];
                            int seg_prev = -1;
#pragma unroll
                            for (int jb = 0; jb < kJBC; ++jb) {
                                const int jo = (jb0 + jb) * kOT;
                                if (jo + tid < fft) {
                                    const int seg = jo >> p.seg_lg;        // CTA-uniform (seg_len % 256 == 0 or nseg == 1)
                                    if (seg != seg_prev) {
                                        const cx<T> *c4 = coef + ((l * NR) * p.nseg + seg) * 4;
#pragma unroll
                                        for (int r = 0; r < NR; ++r)
#pragma unroll
                                            for (int o = 0; o < 4; ++o) cf[r][o] = c4[r * p.nseg * 4 + o];
                                        seg_prev = seg;
                                    }
                                    const cx<T> x = xl[jo];
                                    if (p.porder == 3) {
#pragma unroll
                                        for (int r = 0; r < NR; ++r) tap_mac<T, 3>(acc[jb][r], cf[r], tau[jb], x);
                                    } else {
#pragma unroll
                                        for (int r = 0; r < NR; ++r) tap_mac<T, 2>(acc[jb][r], cf[r], tau[jb], x);
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
                                cis_phase<T>(fma(rec.dl, n0, rec.th0), &sn, &cs);
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

            // ---------------- G: H_k = sum_l gbar_l W^(k d_l), equalise / detect, demap, count.
            // One thread owns the 4 bins k0 + u*fft/4: W^((k0+u fft/4) d) = W^(k0 d) * (-j)^(u d),
            // so the twiddle product is shared and the other three bins cost only sign swaps.
            constexpr int NU = (NR * NT <= 4) ? 4 : 1;
            const int kstride = fft / NU;
            for (int k0 = tid; k0 < kstride; k0 += kOT) {
                // W^((k0 + u fft/4) d) = W^(k0 d) (-j)^(u d): accumulate the taps into 4 residue classes
                // S_c (c = d mod 4), then H[u] = sum_c (-j)^(u c) S_c is a 4-point DFT per matrix entry.
                cx<T> H[NU][NR][NT];
#pragma unroll
                for (int u = 0; u < NU; ++u)
#pragma unroll
                    for (int r = 0; r < NR; ++r)
#pragma unroll
                        for (int t = 0; t < NT; ++t) H[u][r][t] = {T(0), T(0)};
                for (int l = 0; l < p.n_taps; ++l) {
                    const int d = p.delays[l];
                    const cx<T> w = tw[(k0 * d) & (fft - 1)];
                    const cx<T> *gl = gbar + l * NR * NT;
                    if constexpr (NU == 4) {
                        switch (d & 3) {                         // CTA-uniform
                            case 0: hk_class<T, NR, NT, 0>(H, gl, w); break;
                            case 1: hk_class<T, NR, NT, 1>(H, gl, w); break;
                            case 2: hk_class<T, NR, NT, 2>(H, gl, w); break;
                            default: hk_class<T, NR, NT, 3>(H, gl, w); break;
                        }
                    } else {
#pragma unroll
                        for (int r = 0; r < NR; ++r)
#pragma unroll
                            for (int t = 0; t < NT; ++t) cmac(H[0][r][t], gl[r * NT + t], w);
                    }
                }
                if constexpr (NU == 4) {
#pragma unroll
                    for (int r = 0; r < NR; ++r)
#pragma unroll
                        for (int t = 0; t < NT; ++t) {
                            const cx<T> a0 = H[0][r][t] + H[2][r][t], a1 = H[0][r][t] - H[2][r][t];
                            const cx<T> a2 = H[1][r][t] + H[3][r][t], a3 = H[1][r][t] - H[3][r][t];
                            const cx<T> rot = {a3.im, -a3.re};           // -j * a3
                            H[0][r][t] = a0 + a2;
                            H[1][r][t] = a1 + rot;
                            H[2][r][t] = a0 - a2;
                            H[3][r][t] = a1 - rot;
                        }
                }
#pragma unroll
                for (int u = 0; u < NU; ++u) {
                    const int k = k0 + u * kstride;
                    const int q = pos_of(k, fft, p.used, p.half);
                    if (q < 0) continue;
                    cx<T> y[NR];
#pragma unroll
                    for (int r = 0; r < NR; ++r) y[r] = rx_scale * Yp[r][k];
                    if (p.rx_out) {
#pragma unroll
                        for (int r = 0; r < NR; ++r)
                            static_cast<cx<T> *>(p.rx_out)[(size_t(frame) * NR + r) * (size_t(p.n_sym) * p.used) + size_t(s) * p.used + q] = y[r];
                    }
                    cx<T> z[NT];
                    if constexpr (NR == 1 && NT == 1) {
                        z[0] = cdiv(y[0], H[u][0][0]);              // OfdmOneTapEqualizer (ofdm.py:510-511)
                    } else {
                        HermSolver<NT> sol;
                        sol.template factor_from_channel<cx<T>, NR>(H[u], NR, p.fnv);
                        cx<double> b[NT];
#pragma unroll
                        for (int t = 0; t < NT; ++t) {
                            b[t] = {0.0, 0.0};
#pragma unroll
                            for (int r = 0; r < NR; ++r) cmac_conj(b[t], cvt<double>(H[u][r][t]), cvt<double>(y[r]));
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
    s += 2 * al(sizeof(T) * p.P4);
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

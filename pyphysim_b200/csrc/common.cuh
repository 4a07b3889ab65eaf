// common.cuh — complex arithmetic, constellation map / demap, counter reduction.
// Shared by every kernel of libb200phy (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200phy.h"

namespace b200phy {

// ---------------------------------------------------------------- complex numbers
template <typename T> struct cx { T re, im; };

template <typename T> __host__ __device__ __forceinline__ cx<T> mk(T a, T b) { return cx<T>{a, b}; }
template <typename T> __device__ __forceinline__ cx<T> operator+(cx<T> a, cx<T> b) { return {a.re + b.re, a.im + b.im}; }
template <typename T> __device__ __forceinline__ cx<T> operator-(cx<T> a, cx<T> b) { return {a.re - b.re, a.im - b.im}; }
template <typename T> __device__ __forceinline__ cx<T> operator*(cx<T> a, cx<T> b) {
    return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
template <typename T> __device__ __forceinline__ cx<T> operator*(T s, cx<T> a) { return {s * a.re, s * a.im}; }
template <typename T> __device__ __forceinline__ cx<T> conj(cx<T> a) { return {a.re, -a.im}; }
template <typename T> __device__ __forceinline__ T norm2(cx<T> a) { return a.re * a.re + a.im * a.im; }
// acc += a * b
template <typename T> __device__ __forceinline__ void cmac(cx<T> &acc, cx<T> a, cx<T> b) {
    acc.re = fma(a.re, b.re, acc.re);
    acc.re = fma(-a.im, b.im, acc.re);
    acc.im = fma(a.re, b.im, acc.im);
    acc.im = fma(a.im, b.re, acc.im);
}
// acc += conj(a) * b
template <typename T> __device__ __forceinline__ void cmac_conj(cx<T> &acc, cx<T> a, cx<T> b) {
    acc.re = fma(a.re, b.re, acc.re);
    acc.re = fma(a.im, b.im, acc.re);
    acc.im = fma(a.re, b.im, acc.im);
    acc.im = fma(-a.im, b.re, acc.im);
}
// a / b (plain formula; operands here are O(1) so no scaling is needed)
__device__ __forceinline__ float recip(float x) { return __frcp_rn(x); }     // correctly rounded, no div sequence
__device__ __forceinline__ double recip(double x) { return 1.0 / x; }
template <typename T> __device__ __forceinline__ cx<T> cdiv(cx<T> a, cx<T> b) {
    T d = recip(b.re * b.re + b.im * b.im);
    return {(a.re * b.re + a.im * b.im) * d, (a.im * b.re - a.re * b.im) * d};
}
template <typename T, typename S> __device__ __forceinline__ cx<T> cvt(cx<S> a) { return {T(a.re), T(a.im)}; }

// ---- packed FP32 pairs (Blackwell FFMA2: fma.rn.f32x2).  One instruction issues two FMAs; a scalar
// operand written as pk2(s, s) is folded by ptxas into a broadcast operand (no MOVs), so an issue-bound
// loop halves its issue slots.  Same rounding as two scalar fmaf.
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(u64 v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// ---------------------------------------------------------------- constellations
// Device view of a modulator (Modulator.symbols + how to slice it).
struct Modem {
    int kind;     // B200PHY_MODEM_*
    int M;        // constellation size
    int bits;     // log2(M) (level2bits, util/misc.py:392-414)
    int side;     // QAM: L = sqrt(M)
    int hbits;    // QAM: log2(L)
    double scale; // QAM: sqrt(2(M-1)/3) (fundamental.py:712-716)
    float hscale, hside;   // 0.5*scale and 0.5*L in float, for the slicer
    unsigned m1, m2, m4;   // masks that keep the Gray -> binary shifts inside each half of the index
};

__host__ inline int ilog2(int v) { int b = 0; while ((1 << b) < v) ++b; return b; }

__host__ inline Modem make_modem(int kind, int M) {
    Modem m;
    m.kind = kind; m.M = M; m.bits = M <= 2 ? 1 : ilog2(M);
    m.side = 1; m.hbits = 0; m.scale = 1.0;
    if (kind == B200PHY_MODEM_QAM) {
        m.hbits = m.bits / 2; m.side = 1 << m.hbits;
        m.scale = sqrt((M - 1) * 2.0 / 3.0);
    }
    m.hscale = float(0.5 * m.scale); m.hside = float(0.5 * m.side);
    // bit b of each half may only receive bits b+1.. of the SAME half
    const unsigned half = (1u << m.hbits) - 1u;
    m.m1 = ((half >> 1) << m.hbits) | (half >> 1);
    m.m2 = ((half >> 2) << m.hbits) | (half >> 2);
    m.m4 = ((half >> 4) << m.hbits) | (half >> 4);
    return m;
}

// gray2binary restricted to 8-bit values (util/conversion.py:252-279)
__device__ __forceinline__ int ungray8(int g) { g ^= g >> 4; g ^= g >> 2; g ^= g >> 1; return g; }

// Modulator.modulate: table gather (BPSK: 1 - 2*idx, fundamental.py:630)
template <typename T>
__device__ __forceinline__ cx<T> map_symbol(const Modem &m, const cx<T> *__restrict__ tab, int idx) {
    if (m.kind == B200PHY_MODEM_BPSK) return {T(1 - 2 * idx), T(0)};
    return tab[idx];
}

// Modulator.demodulate.  TABLE: argmin over squared distance, first minimum wins.  QAM: per-axis
// slicer (jj from Re, ii from Im, data index = ungray(ii)*L + ungray(jj)), which returns the same
// index as the M-way search for the Gray-mapped square constellation of fundamental.py:689-777.
template <typename T>
__device__ __forceinline__ int demap_symbol(const Modem &m, const cx<T> *__restrict__ tab, cx<T> r) {
    if (m.kind == B200PHY_MODEM_QAM) {
        int jj, ii;
        if constexpr (sizeof(T) == 4) {
            jj = __float2int_rd(fmaf(r.re, m.hscale, m.hside));
            ii = __float2int_rd(fmaf(-r.im, m.hscale, m.hside));
        } else {
            const T sc = T(m.scale), L = T(m.side);
            jj = int(floor((r.re * sc + L) * T(0.5)));
            ii = int(floor((L - r.im * sc) * T(0.5)));
        }
        jj = min(max(jj, 0), m.side - 1);
        ii = min(max(ii, 0), m.side - 1);
        // Gray -> binary of both halves at once (gray2binary, util/conversion.py:252-279)
        unsigned g = (unsigned(ii) << m.hbits) | unsigned(jj);
        g ^= (g >> 4) & m.m4;
        g ^= (g >> 2) & m.m2;
        g ^= (g >> 1) & m.m1;
        return int(g);
    }
    if (m.kind == B200PHY_MODEM_QPSK) return int(r.re < T(0)) | (int(r.im < T(0)) << 1);
    if (m.kind == B200PHY_MODEM_BPSK) {
        // NumPy orders complex numbers lexicographically: (r < 0) == re<0 or (re==0 and im<0)
        return (r.re < T(0) || (r.re == T(0) && r.im < T(0))) ? 1 : 0;
    }
    int best = 0;
    T bd = norm2(tab[0] - r);
    for (int k = 1; k < m.M; ++k) {
        T d = norm2(tab[k] - r);
        if (d < bd) { bd = d; best = k; }
    }
    return best;
}

// ---------------------------------------------------------------- counters
// Block-wide sum of two per-thread error counts, one int64 atomic per counter per block.
// counters = {symbol_errors, bit_errors, num_symbols, num_bits}
__device__ __forceinline__ void flush_counters(unsigned sym_err, unsigned bit_err,
                                               unsigned long long *counters) {
    __shared__ unsigned s_cnt[2];
    if (threadIdx.x == 0) { s_cnt[0] = 0; s_cnt[1] = 0; }
    __syncthreads();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sym_err += __shfl_xor_sync(0xffffffffu, sym_err, o);
        bit_err += __shfl_xor_sync(0xffffffffu, bit_err, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (sym_err) atomicAdd(&s_cnt[0], sym_err);
        if (bit_err) atomicAdd(&s_cnt[1], bit_err);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_cnt[0]) atomicAdd(&counters[0], (unsigned long long)s_cnt[0]);
        if (s_cnt[1]) atomicAdd(&counters[1], (unsigned long long)s_cnt[1]);
    }
}

// ---------------------------------------------------------------- host-side helpers
void set_error(const char *fmt, ...);
int check_cuda(cudaError_t e, const char *what);
void count_launch(int n = 1);
void note_kernel(const char *fmt, ...);   // name<template arguments> of the fused-link kernel just launched
int check_modem(const b200phy_modem *m, Modem *out);

// Stream-mode kernels use vector loads / stores: a base pointer that is not aligned to `a` bytes (a sliced
// tensor, an odd element offset) must be refused here, not fault inside the kernel.  NULL passes.
inline int require_aligned(const void *ptr, size_t a, const char *name) {
    if (ptr && (reinterpret_cast<uintptr_t>(ptr) & (a - 1))) {
        set_error("%s must be %zu-byte aligned (got %p): pass a contiguous tensor from its first element", name, a, ptr);
        return B200PHY_ERR_INVALID;
    }
    return B200PHY_OK;
}

#define B200_CHECK_LAUNCH(what)                                            \
    do {                                                                   \
        count_launch();                                                    \
        int _e = check_cuda(cudaGetLastError(), what);                     \
        if (_e) return _e;                                                 \
    } while (0)

}  // namespace b200phy

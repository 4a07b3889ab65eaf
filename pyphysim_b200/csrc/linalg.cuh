// linalg.cuh — per-thread small Hermitian solves in registers (always double precision).
// Used for Blast ZF / MMSE receive filters (MimoBase._calcZeroForceFilter / _calcMMSEFilter,
// mimo/mimo.py:264-309): G y = (H^H H + s2 I)^-1 H^H y, evaluated as a Cholesky solve of the
// normal equations.  s2 = 0 gives the pseudo-inverse for full-column-rank H (ZF).
#pragma once
#include "common.cuh"

namespace b200phy {

template <int NT> struct HermSolver {
    cx<double> L[NT][NT];  // lower triangle; diagonal stored as reciprocal in invd
    double invd[NT];

    // A = Hh H + s2 I from H[Nr][NT] (rows r < Nr valid)
    template <typename HT, int NRMAX>
    __device__ __forceinline__ void factor_from_channel(const HT (&H)[NRMAX][NT], int Nr, double s2) {
        cx<double> A[NT][NT];
#pragma unroll
        for (int i = 0; i < NT; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                cx<double> acc = {i == j ? s2 : 0.0, 0.0};
#pragma unroll
                for (int r = 0; r < NRMAX; ++r)
                    if (r < Nr) cmac_conj(acc, cvt<double>(H[r][i]), cvt<double>(H[r][j]));
                A[i][j] = acc;
            }
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            double d = A[j][j].re;
#pragma unroll
            for (int k = 0; k < j; ++k) d -= norm2(L[j][k]);
            const double inv = rsqrt(d);
            invd[j] = inv;
            L[j][j] = {d * inv, 0.0};
#pragma unroll
            for (int i = j + 1; i < NT; ++i) {
                cx<double> v = A[i][j];
#pragma unroll
                for (int k = 0; k < j; ++k) {
                    // v -= L[i][k] * conj(L[j][k])
                    const cx<double> a = L[i][k], b = L[j][k];
                    v.re -= a.re * b.re + a.im * b.im;
                    v.im -= a.im * b.re - a.re * b.im;
                }
                L[i][j] = {v.re * inv, v.im * inv};
            }
        }
    }

    // z = A^-1 b (in place)
    __device__ __forceinline__ void solve(cx<double> (&b)[NT]) const {
#pragma unroll
        for (int i = 0; i < NT; ++i) {
            cx<double> v = b[i];
#pragma unroll
            for (int k = 0; k < i; ++k) {
                const cx<double> a = L[i][k], w = b[k];
                v.re -= a.re * w.re - a.im * w.im;
                v.im -= a.re * w.im + a.im * w.re;
            }
            b[i] = {v.re * invd[i], v.im * invd[i]};
        }
#pragma unroll
        for (int i = NT - 1; i >= 0; --i) {
            cx<double> v = b[i];
#pragma unroll
            for (int k = i + 1; k < NT; ++k) {
                // v -= conj(L[k][i]) * b[k]
                const cx<double> a = L[k][i], w = b[k];
                v.re -= a.re * w.re + a.im * w.im;
                v.im -= a.re * w.im - a.im * w.re;
            }
            b[i] = {v.re * invd[i], v.im * invd[i]};
        }
    }
};

// 2x2: closed-form inverse of A = [[a, conj(c)], [c, b]] (no square roots)
template <> struct HermSolver<2> {
    double a, b, idet;
    cx<double> c;

    template <typename HT, int NRMAX>
    __device__ __forceinline__ void factor_from_channel(const HT (&H)[NRMAX][2], int Nr, double s2) {
        a = s2; b = s2; c = {0.0, 0.0};
#pragma unroll
        for (int r = 0; r < NRMAX; ++r)
            if (r < Nr) {
                const cx<double> h0 = cvt<double>(H[r][0]), h1 = cvt<double>(H[r][1]);
                a += norm2(h0);
                b += norm2(h1);
                cmac_conj(c, h1, h0);                    // A[1][0] = sum conj(H[r][1]) H[r][0]
            }
        idet = 1.0 / (a * b - norm2(c));
    }

    __device__ __forceinline__ void solve(cx<double> (&v)[2]) const {
        const cx<double> v0 = v[0], v1 = v[1];
        // A^-1 = [[b, -conj(c)], [-c, a]] / det
        v[0] = {idet * (b * v0.re - (c.re * v1.re + c.im * v1.im)), idet * (b * v0.im - (c.re * v1.im - c.im * v1.re))};
        v[1] = {idet * (a * v1.re - (c.re * v0.re - c.im * v0.im)), idet * (a * v1.im - (c.re * v0.im + c.im * v0.re))};
    }
};

}  // namespace b200phy

// svd.cuh — per-thread decompositions of one small complex channel matrix:
//   * SmallSvd<NR, NT>: thin SVD H = U diag(S) V^H, NR >= NT <= 4, by one-sided (Hestenes) Jacobi in double:
//     pairs of columns of A = H are rotated until mutually orthogonal while V accumulates the same
//     rotations; then S_i = |a_i|, u_i = a_i / S_i.  Sorted descending and put in a FIXED gauge: each pair
//     (u_i, v_i) is multiplied by the unit phase that makes the largest-magnitude entry of v_i real positive
//     (oracle/mimo.py svd_canonical; numpy's own gauge is LAPACK's and not reproducible — the two differ by
//     one unit phase per singular pair, which cancels in G_H H W).  What the reference takes from
//     np.linalg.svd in SVDMimo / GMDMimo (pyphysim/mimo/mimo.py:855-898, 974-1019).
//   * small_gmd<NR, NT>: the geometric mean decomposition of Jiang, Hager and Li that util.misc.gmd
//     implements (pyphysim/util/misc.py:18-159), on an SVD given in registers: H = Q R P^H, R upper
//     triangular with constant diagonal (the geometric mean of the singular values).
// Everything is unrolled over the compile-time sizes, so the matrices live in registers.
#pragma once
#include "common.cuh"

namespace b200phy {

template <int N> __device__ __forceinline__ int sel_get(const int (&a)[N], int i) {
    int v = a[0];
#pragma unroll
    for (int k = 1; k < N; ++k) v = (i == k) ? a[k] : v;
    return v;
}
template <int N> __device__ __forceinline__ void sel_set(int (&a)[N], int i, int v) {
#pragma unroll
    for (int k = 0; k < N; ++k) a[k] = (i == k) ? v : a[k];
}

template <int NR, int NT> struct SmallSvd {
    cx<double> U[NR][NT], V[NT][NT];
    double S[NT];

    template <typename HT> __device__ void compute(const HT (&H)[NR][NT]) {
#pragma unroll
        for (int r = 0; r < NR; ++r)
#pragma unroll
            for (int t = 0; t < NT; ++t) U[r][t] = cvt<double>(H[r][t]);
#pragma unroll
        for (int a = 0; a < NT; ++a)
#pragma unroll
            for (int b = 0; b < NT; ++b) V[a][b] = {a == b ? 1.0 : 0.0, 0.0};

        // cyclic sweeps over the column pairs; quadratic convergence, 4x4 needs ~6 sweeps in double
        for (int sweep = 0; sweep < 16; ++sweep) {
            bool rotated = false;
#pragma unroll
            for (int i = 0; i < NT - 1; ++i)
#pragma unroll
                for (int j = i + 1; j < NT; ++j) {
                    double al = 0.0, be = 0.0;
                    cx<double> ga = {0.0, 0.0};
#pragma unroll
                    for (int r = 0; r < NR; ++r) {
                        al += norm2(U[r][i]);
                        be += norm2(U[r][j]);
                        cmac_conj(ga, U[r][i], U[r][j]);          // a_i^H a_j
                    }
                    const double g2 = norm2(ga);
                    if (g2 > 1e-30 * al * be && g2 > 0.0) {
                        rotated = rotated || g2 > 1e-28 * al * be;
                        const double g = sqrt(g2);
                        const cx<double> ph = {ga.re / g, -ga.im / g};     // e^{-j arg(gamma)}
                        const double zeta = (be - al) / (2.0 * g);
                        const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                        const double c = rsqrt(1.0 + t * t), s = c * t;
                        // (a_i, a_j) <- (c a_i - s ph a_j, s a_i + c ph a_j); same for the columns of V
#pragma unroll
                        for (int r = 0; r < NR; ++r) {
                            const cx<double> x = U[r][i], y = ph * U[r][j];
                            U[r][i] = {c * x.re - s * y.re, c * x.im - s * y.im};
                            U[r][j] = {s * x.re + c * y.re, s * x.im + c * y.im};
                        }
#pragma unroll
                        for (int r = 0; r < NT; ++r) {
                            const cx<double> x = V[r][i], y = ph * V[r][j];
                            V[r][i] = {c * x.re - s * y.re, c * x.im - s * y.im};
                            V[r][j] = {s * x.re + c * y.re, s * x.im + c * y.im};
                        }
                    }
                }
            if (!rotated) break;
        }
#pragma unroll
        for (int i = 0; i < NT; ++i) {
            double n2 = 0.0;
#pragma unroll
            for (int r = 0; r < NR; ++r) n2 += norm2(U[r][i]);
            S[i] = sqrt(n2);
        }
        // descending order: bubble network with static indices
#pragma unroll
        for (int a = 0; a < NT - 1; ++a)
#pragma unroll
            for (int b = 0; b < NT - 1 - a; ++b)
                if (S[b] < S[b + 1]) {
                    const double ts = S[b]; S[b] = S[b + 1]; S[b + 1] = ts;
#pragma unroll
                    for (int r = 0; r < NR; ++r) { const cx<double> tu = U[r][b]; U[r][b] = U[r][b + 1]; U[r][b + 1] = tu; }
#pragma unroll
                    for (int r = 0; r < NT; ++r) { const cx<double> tv = V[r][b]; V[r][b] = V[r][b + 1]; V[r][b + 1] = tv; }
                }
        // normalise u_i and fix the gauge of every singular pair
#pragma unroll
        for (int i = 0; i < NT; ++i) {
            cx<double> big = V[0][i];
            double bm = norm2(big);
#pragma unroll
            for (int r = 1; r < NT; ++r) {
                const double m2 = norm2(V[r][i]);
                if (m2 > bm) { bm = m2; big = V[r][i]; }
            }
            const double ib = rsqrt(bm);
            const cx<double> ph = {big.re * ib, -big.im * ib};             // conj(big) / |big|
            const double is = S[i] > 0.0 ? 1.0 / S[i] : 0.0;
#pragma unroll
            for (int r = 0; r < NT; ++r) V[r][i] = ph * V[r][i];
#pragma unroll
            for (int r = 0; r < NR; ++r) { const cx<double> u = ph * U[r][i]; U[r][i] = {u.re * is, u.im * is}; }
        }
    }
};

// Q (NR x NT), R (NT x NT real upper triangular), P (NT x NT) from U, S, V (all NT singular values kept:
// the reference's tol = 0).  Q and P start as U and V and are updated in place.
template <int NR, int NT>
__device__ void small_gmd(cx<double> (&Q)[NR][NT], const double (&S)[NT], cx<double> (&P)[NT][NT], double (&R)[NT][NT]) {
    double d[NT], carry[NT];
    int where[NT], rank[NT];
    double lg = 0.0;
#pragma unroll
    for (int i = 0; i < NT; ++i) {
        d[i] = S[i]; carry[i] = 0.0; where[i] = i; rank[i] = i;
        lg += log(S[i]);
#pragma unroll
        for (int j = 0; j < NT; ++j) R[i][j] = 0.0;
    }
    const double mean = exp(lg / NT);
    int next_big = 1, next_small = NT - 1;
#pragma unroll
    for (int k = 0; k < NT - 1; ++k) {
        int i;
        bool trivial;
        if (d[k] >= mean) { i = sel_get(where, next_small); --next_small; }
        else              { i = sel_get(where, next_big); ++next_big; }
        double di = d[k + 1];
#pragma unroll
        for (int c = k + 2; c < NT; ++c) di = (i == c) ? d[c] : di;
        trivial = (d[k] >= mean) ? (di >= mean) : (di <= mean);
        // bring slot i into slot k+1 (columns of Q and P follow)
#pragma unroll
        for (int c = k + 2; c < NT; ++c)
            if (i == c) {
                const double td = d[k + 1]; d[k + 1] = d[c]; d[c] = td;
                const int r = rank[k + 1];
                sel_set(where, r, c);
                rank[c] = r;
#pragma unroll
                for (int r2 = 0; r2 < NR; ++r2) { const cx<double> tq = Q[r2][k + 1]; Q[r2][k + 1] = Q[r2][c]; Q[r2][c] = tq; }
#pragma unroll
                for (int r2 = 0; r2 < NT; ++r2) { const cx<double> tp = P[r2][k + 1]; P[r2][k + 1] = P[r2][c]; P[r2][c] = tp; }
            }
        const double a = d[k], b = d[k + 1];
        double c = 1.0, s = 0.0;
        if (!trivial) {
            c = sqrt((mean * mean - b * b) / (a * a - b * b));
            s = sqrt(1.0 - c * c);
        }
        d[k + 1] = a * b / mean;
        carry[k] = s * c * (b * b - a * a) / mean;
        R[k][k] = mean;
#pragma unroll
        for (int r = 0; r < k; ++r) { R[r][k] = carry[r] * c; carry[r] = -carry[r] * s; }
        // P[:, (k, k+1)] <- P[:, (k, k+1)] [[c, -s], [s, c]]
#pragma unroll
        for (int r = 0; r < NT; ++r) {
            const cx<double> x = P[r][k], y = P[r][k + 1];
            P[r][k] = {c * x.re + s * y.re, c * x.im + s * y.im};
            P[r][k + 1] = {-s * x.re + c * y.re, -s * x.im + c * y.im};
        }
        // Q[:, (k, k+1)] <- Q[:, (k, k+1)] [[c a, -s b], [s b, c a]] / mean
        const double ca = c * a / mean, sb = s * b / mean;
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const cx<double> x = Q[r][k], y = Q[r][k + 1];
            Q[r][k] = {ca * x.re + sb * y.re, ca * x.im + sb * y.im};
            Q[r][k + 1] = {-sb * x.re + ca * y.re, -sb * x.im + ca * y.im};
        }
    }
    R[NT - 1][NT - 1] = mean;
#pragma unroll
    for (int r = 0; r < NT - 1; ++r) R[r][NT - 1] = carry[r];
}

}  // namespace b200phy

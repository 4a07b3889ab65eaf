"""Blast (ZF / MMSE) and Alamouti — NumPy restatement.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Reference: ``pyphysim/mimo/mimo.py``.
"""
import math

import numpy as np


def zf_filter(H):
    """MimoBase._calcZeroForceFilter (mimo.py:264-285): pinv(H)."""
    return np.linalg.pinv(H)


def mmse_filter(H, noise_var):
    """MimoBase._calcMMSEFilter (mimo.py:287-309): solve(H^H H + s2 I, H^H)."""
    Hh = H.conj().T
    return np.linalg.solve(np.dot(Hh, H) + noise_var * np.eye(H.shape[1]), Hh)


def blast_receive_filter(H, noise_var=0.0):
    """Blast._calc_receive_filter (mimo.py:590-607): MMSE iff noise_var > 0, x sqrt(Nt)."""
    G = mmse_filter(H, noise_var) if noise_var > 0 else zf_filter(H)
    return G * math.sqrt(H.shape[1])


def blast_encode(x, Nt):
    """Blast.encode (mimo.py:609-640): symbol k -> antenna k mod Nt, / sqrt(Nt)."""
    if x.size % Nt != 0:
        raise ValueError("Input array number of elements must be a multiple of the"
                         " number of transmit antennas")
    return x.reshape((Nt, -1), order='F') / math.sqrt(Nt)


def blast_decode(y, H, noise_var=0.0):
    """Blast.decode (mimo.py:642-660)."""
    return blast_receive_filter(H, noise_var).dot(y).reshape(-1, order='F')


def blast_receive_filter_batched(H, noise_var=0.0):
    """blast_receive_filter for a stack H[K, Nr, Nt] (one matrix per subcarrier).  np.linalg.solve / pinv on a
    stack run the same LAPACK routine once per matrix, so every G[k] equals the per-matrix call
    (tests/test_oracle_kat.py::test_blast_batched_equals_loop)."""
    Hh = np.conj(np.swapaxes(H, -1, -2))
    if noise_var > 0:
        G = np.linalg.solve(Hh @ H + noise_var * np.eye(H.shape[-1]), Hh)
    else:
        G = np.linalg.pinv(H)
    return G * math.sqrt(H.shape[-1])


def alamouti_encode(s):
    """Alamouti.encode/_encode (mimo.py:1166-1214): [[s0, -s1*], [s1, s0*]] / sqrt(2)."""
    Ns = s.size
    out = np.empty((2, Ns), dtype=complex)
    out[0, 0::2] = s[0::2]
    out[0, 1::2] = -np.conj(s[1::2])
    out[1, 0::2] = s[1::2]
    out[1, 1::2] = np.conj(s[0::2])
    return out / math.sqrt(2)


def alamouti_decode(y, H):
    """Alamouti.decode/_decode (mimo.py:1216-1287); H is [Nr, 2]."""
    H = np.atleast_2d(H)
    h0, h1 = H[:, 0], H[:, 1]
    y0, y1 = y[:, 0::2], y[:, 1::2]
    out = np.empty(y.shape[1], dtype=complex)
    out[0::2] = h0.conj() @ y0 + h1 @ y1.conj()
    out[1::2] = h1.conj() @ y0 + (-h0) @ y1.conj()
    out /= np.linalg.norm(H, 'fro') ** 2
    return out * math.sqrt(2)


# ---------------------------------------------------------------------------------------------------
# MRT / MRC / SVDMimo / GMDMimo and the post-processing SINRs
# ---------------------------------------------------------------------------------------------------
def post_processing_linear_sinrs(H, W, G_H, noise_var=0.0):
    """calc_post_processing_linear_SINRs (mimo.py:63-114): per stream |diag|^2 / (|row sum - diag|^2 +
    noise_var * ||row of G_H||^2); a scalar G_H amplifies the single stream's noise by |G_H|^2."""
    Heq = np.dot(G_H, np.dot(H, W))
    Heq = np.atleast_2d(Heq)
    diag = np.diag(Heq)
    leak = Heq.sum(axis=1) - diag
    if isinstance(G_H, np.ndarray):
        gain = np.linalg.norm(np.atleast_2d(G_H), axis=1) ** 2
    else:
        gain = abs(G_H) ** 2
    return np.abs(diag) ** 2 / (np.abs(leak) ** 2 + noise_var * gain)


def mrt_precoder(h):
    """MRT._calc_precoder (mimo.py:688-712): co-phasing weights exp(-j angle(h))^T / sqrt(Nt); h is [1, Nt]."""
    h = np.atleast_2d(h)
    return np.exp(-1j * np.angle(h)).T / math.sqrt(h.shape[1])


def mrt_receive_filter(h):
    """MRT._calc_receive_filter (mimo.py:714-735): the scalar sqrt(Nt) / sum |h|."""
    h = np.atleast_2d(h)
    return math.sqrt(h.shape[1]) / np.sum(np.abs(h))


def mrt_encode(x, h):
    """MRT.encode (mimo.py:737-761): [Nt, n] = W * x."""
    return mrt_precoder(h) * x[np.newaxis, :]


def mrt_decode(y, h):
    """MRT.decode (mimo.py:763-783): G_H * y flattened."""
    return (mrt_receive_filter(h) * y).reshape(-1)


def svd_precoder(H):
    """SVDMimo._calc_precoder (mimo.py:855-874): V / sqrt(Nt) with numpy's (LAPACK's) SVD gauge."""
    Vh = np.linalg.svd(H)[2]
    return Vh.conj().T / math.sqrt(H.shape[1])


def svd_receive_filter(H):
    """SVDMimo._calc_receive_filter (mimo.py:876-898): diag(1/S) U^H sqrt(Nt)."""
    U, S, _ = np.linalg.svd(H)
    return np.diag(1.0 / S).dot(U.conj().T) * math.sqrt(H.shape[1])


def gmd(U, S, Vh, tol=0.0):
    """util.misc.gmd (misc.py:18-159), the geometric mean decomposition of Jiang, Hager and Li:
    from A = U diag(S) Vh build Q, R, P with A = Q R P^H, Q and P with orthonormal columns, R upper
    triangular with every diagonal entry equal to the geometric mean of the singular values >= tol.

    Step k makes R[k, k] the geometric mean: the pair (d[k], d[k+1]) must bracket it, so the smallest
    unused singular value is swapped into slot k+1 when d[k] is above the mean and the largest unused
    one when it is below (with the matching column swaps in Q and P); then one plane rotation of the two
    P columns and one scaled rotation of the two Q columns put the mean on the diagonal, leave the product
    d[k] d[k+1] / mean in slot k+1 and push the off-diagonal mass into column k of R.
    """
    m, n = U.shape[0], Vh.shape[0]
    R = np.zeros((m, n))
    P = Vh.conj().T.copy()
    Q = U.copy()
    d = np.copy(S)
    p = int(np.sum(np.asarray(S) >= tol))
    if p < 1:
        raise RuntimeError("This is no singular value greater than the tolerance")
    if p < 2:
        R[0, 0] = d[0]
    carry = np.zeros(max(p - 1, 0))               # off-diagonal mass still travelling right
    where = list(range(p))                        # where[r]: slot holding the r-th largest value
    rank = list(range(p))                         # rank[s]: which rank sits in slot s
    next_big, next_small = 1, p - 1
    mean = np.prod(np.asarray(S)[:p]) ** (1.0 / p)        # np.float64, squared with ** like the reference (libm pow)
    for k in range(p - 1):
        if d[k] >= mean:
            i = where[next_small]
            next_small -= 1
            trivial = d[i] >= mean
        else:
            i = where[next_big]
            next_big += 1
            trivial = d[i] <= mean
        k1 = k + 1
        if i != k1:
            d[k1], d[i] = d[i], d[k1]
            r = rank[k1]
            where[r] = i
            rank[i] = r
            Q[:, [k1, i]] = Q[:, [i, k1]]
            P[:, [k1, i]] = P[:, [i, k1]]
        a, b = d[k], d[k1]
        if trivial:
            c, s = 1.0, 0.0
        else:
            c = math.sqrt((mean ** 2 - b ** 2) / (a ** 2 - b ** 2))
            s = math.sqrt(1 - c ** 2)
        d[k1] = a * b / mean
        carry[k] = s * c * (b ** 2 - a ** 2) / mean
        R[k, k] = mean
        if k > 0:
            R[:k, k] = carry[:k] * c
            carry[:k] = -carry[:k] * s
        P[:, [k, k1]] = P[:, [k, k1]].dot(np.array([[c, -s], [s, c]]))
        Q[:, [k, k1]] = Q[:, [k, k1]].dot((1.0 / mean) * np.array([[c * a, -s * b], [s * b, c * a]]))
    R[p - 1, p - 1] = mean
    R[:p - 1, p - 1] = carry
    return Q, R, P


def gmd_precoder(H):
    """GMDMimo._calc_precoder (mimo.py:974-994): P / sqrt(Nt)."""
    U, S, Vh = np.linalg.svd(H)
    return gmd(U, S, Vh)[2] / math.sqrt(H.shape[1])


def gmd_receive_filter(H, noise_var=0.0):
    """GMDMimo._calc_receive_filter (mimo.py:996-1019): the Blast filter of the equivalent channel Q R."""
    U, S, Vh = np.linalg.svd(H)
    Q, R, _ = gmd(U, S, Vh)
    return blast_receive_filter(Q.dot(R), noise_var)


def precoded_encode(x, W):
    """SVDMimo.encode / GMDMimo.encode (mimo.py:900-928, 1021-1048): W . reshape(x, (Nt, -1))."""
    Nt = W.shape[1]
    if x.size % Nt != 0:
        raise ValueError("Input array number of elements must be a multiple of the"
                         " number of transmit antennas")
    return W.dot(x.reshape(Nt, -1))


def svd_canonical(H):
    """Thin SVD with a FIXED gauge, the convention of the CUDA decomposition (csrc/svd.cuh): singular
    values descending; each pair (u_i, v_i) rotated by the unit phase that makes the largest-magnitude
    entry of v_i real and positive.  numpy's own gauge is LAPACK's and not reproducible elsewhere; the
    two differ by one unit phase per singular pair, which cancels in G_H H W."""
    U, S, Vh = np.linalg.svd(H, full_matrices=False)
    V = Vh.conj().T
    for i in range(V.shape[1]):
        j = int(np.argmax(np.abs(V[:, i])))
        ph = V[j, i] / abs(V[j, i])
        V[:, i] = V[:, i] / ph
        U[:, i] = U[:, i] / ph
    return U, S, V

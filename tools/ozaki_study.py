"""CPU study for DESIGN.md section 11, item 1: would FP64 emulation on the INT8 tensor cores (Ozaki scheme) keep the
likelihood inside the 1e-10 parity bar?

No GPU involved.  The blocked algorithm of the product (gpe_api.cu:potrf_inv_rec -- leaf factor + inverse on 128x128
blocks, every update a GEMM with one triangular operand, then A^-1 = L^-T L^-1) is restated in NumPy with a pluggable
GEMM, and run three ways on the covariance matrices of the bench workload's generator:

  * ``f64``    -- plain float64 GEMMs (what the DMMA kernels compute, up to summation order);
  * ``ozaki-s`` -- every GEMM replaced by its INT8-slice emulation: each operand row (left) / column (right) is scaled
                  by a power of two, cut into ``s`` signed 7-bit slices (exact), the slice products are exact integer
                  GEMMs (here: float64 GEMMs of integer-valued matrices, |sum| < 2^53, so exact as INT32 accumulation
                  would be), grouped by scale and recombined in float64 from the smallest group up.  Products with
                  slice index sum >= s are dropped (the usual triangular truncation): s (s + 1) / 2 INT8 GEMMs.
  * LAPACK     -- numpy.linalg / scipy on the same matrix, as the common reference.

For each it reports the quantities the likelihood and its gradient are made of: log det A, y^T A^-1 y, and
max |A^-1 - A^-1_LAPACK| relative to max |A^-1|, over a range of condition numbers (correlation lengths 0.5 ... 4,
nugget 1e-4 / 1e-6: cond(A) from 1e3 to beyond 1e8).

    python tools/ozaki_study.py [n] [d] [delta:nugget,delta:nugget,...]
"""
import sys
import time

import numpy as np
import scipy.linalg as sl

NB = 128
BITS = 7


def synth(n, d, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.random((n, d))
    w = rng.normal(size=d)
    return X, np.sin(X @ w) + 0.1 * (X ** 2).sum(1)


def cov(X, delta, nugget):
    Z = X / delta
    sq = (Z * Z).sum(1)
    D = np.maximum(sq[:, None] + sq[None, :] - 2.0 * Z @ Z.T, 0.0)
    A = (1.0 - nugget) * np.exp(-D)
    np.fill_diagonal(A, 1.0)
    return A


# ------------------------------------------------------------------------------------------- GEMM flavours
def gemm_f64(A, B):
    return A @ B


def _slices(M, axis, s):
    """Scale each row (axis=1) or column (axis=0) by a power of two to magnitude < 1 and cut it into s signed
    7-bit integer slices.  Returns (list of integer-valued float64 matrices, exponent vector)."""
    amax = np.abs(M).max(axis=axis, keepdims=True)
    e = np.where(amax > 0, np.ceil(np.log2(np.where(amax > 0, amax, 1.0))) + 1.0, 0.0)   # |M| / 2^e <= 1/2
    R = M * np.exp2(-e)
    out = []
    for _ in range(s):
        R = R * float(1 << BITS)
        Q = np.trunc(R)                      # |Q| <= 64: fits a signed byte
        R = R - Q                            # exact
        out.append(Q)
    return out, e


def make_gemm_ozaki(s):
    def gemm(A, B):
        QA, eA = _slices(A, 1, s)
        QB, eB = _slices(B, 0, s)
        groups = [None] * s                  # groups[g]: sum of products with slice indices t + u == g  (exact integers)
        for t in range(s):
            for u in range(s - t):
                P = QA[t] @ QB[u]            # integer-valued, |entries| < 2^12 * k: exact in float64 for k < 2^40
                groups[t + u] = P if groups[t + u] is None else groups[t + u] + P
        C = np.zeros((A.shape[0], B.shape[1]))
        for g in range(s - 1, -1, -1):       # smallest contributions first
            C += groups[g] * float(2.0 ** (-BITS * (g + 2)))
        return C * np.exp2(eA) * np.exp2(eB)
    return gemm


# ------------------------------------------------------------------------------------------- the blocked algorithm
def potrf_inv_rec(A, Li, off, m, gemm, logdet):
    """In place on the lower triangle of A[off:off+m, off:off+m]; writes L^-1 into Li.  Mirrors potrf_inv_rec."""
    if m <= NB:
        blk = A[off:off + m, off:off + m]
        L = np.linalg.cholesky(np.tril(blk) + np.tril(blk, -1).T)
        logdet[0] += 2.0 * np.log(np.diag(L)).sum()
        Li[off:off + m, off:off + m] = sl.solve_triangular(L, np.eye(m), lower=True)
        return
    m1 = (m // 2 + NB - 1) // NB * NB
    m2 = m - m1
    potrf_inv_rec(A, Li, off, m1, gemm, logdet)
    Li11 = Li[off:off + m1, off:off + m1]
    A21 = A[off + m1:off + m, off:off + m1]
    L21 = gemm(A21, Li11.T)                                   # L21 = A21 L11^-T
    A22 = A[off + m1:off + m, off + m1:off + m]
    A22 -= gemm(L21, L21.T)                                   # SYRK
    T = gemm(L21, Li11)                                       # T = L21 L11^-1
    potrf_inv_rec(A, Li, off + m1, m2, gemm, logdet)
    Li22 = Li[off + m1:off + m, off + m1:off + m]
    Li[off + m1:off + m, off:off + m1] = -gemm(Li22, T)       # L^-1_21 = -L22^-1 T


def factor_inverse(A, gemm):
    n = A.shape[0]
    npad = (n + NB - 1) // NB * NB
    Ap = np.eye(npad)
    Ap[:n, :n] = A
    Li = np.zeros((npad, npad))
    logdet = [0.0]
    potrf_inv_rec(Ap, Li, 0, npad, gemm, logdet)
    Ainv = gemm(Li.T, Li)                                     # LAUUM
    return Li[:n, :n], Ainv[:n, :n], logdet[0]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    X, y = synth(n, d)
    print("n = %d, d = %d; relative differences against LAPACK (float64)" % (n, d))
    print("%-6s %-8s %-9s %-10s %-12s %-12s %-12s %-8s" % ("delta", "nugget", "cond(A)", "GEMM", "logdet", "y'A^-1 y", "max|dA^-1|", "seconds"))
    cases = ((0.5, 1e-4), (1.0, 1e-4), (2.0, 1e-4), (2.0, 1e-6), (4.0, 1e-6))
    if len(sys.argv) > 3:                                     # e.g. "1.0:1e-4,2.0:1e-4"
        cases = tuple(tuple(float(v) for v in c.split(":")) for c in sys.argv[3].split(","))
    for delta, nugget in cases:
        A = cov(X, np.full(d, delta), nugget)
        c = np.linalg.cond(A)
        Lr = np.linalg.cholesky(A)
        ld_ref = 2.0 * np.log(np.diag(Lr)).sum()
        Ainv_ref = sl.cho_solve((Lr, True), np.eye(n))
        q_ref = y @ Ainv_ref @ y
        for name, gemm in (("f64", gemm_f64), ("ozaki-7", make_gemm_ozaki(7)), ("ozaki-8", make_gemm_ozaki(8)),
                           ("ozaki-9", make_gemm_ozaki(9)), ("ozaki-10", make_gemm_ozaki(10))):
            t0 = time.time()
            Li, Ainv, ld = factor_inverse(A.copy(), gemm)
            Ainv = np.tril(Ainv) + np.tril(Ainv, -1).T
            q = y @ Ainv @ y
            print("%-6.1f %-8.0e %-9.1e %-10s %-12.2e %-12.2e %-12.2e %-8.1f" % (
                delta, nugget, c, name, abs(ld - ld_ref) / abs(ld_ref), abs(q - q_ref) / abs(q_ref),
                np.abs(Ainv - Ainv_ref).max() / np.abs(Ainv_ref).max(), time.time() - t0), flush=True)


if __name__ == "__main__":
    main()

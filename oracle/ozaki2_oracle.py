"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the INT8 modular FP64-GEMM emulation of csrc/gpe_ozaki.cu.

Nothing in the product imports this file.  There is no reference code for this part (the reference calls NumPy's
float64 matmul, `_emulatoroptimise.py:313-335` and friends); the published algorithm restated here is the
"Ozaki scheme II" (Ozaki, Uchino, Imamura 2025: GEMM emulation by integer modular arithmetic / CRT):

  1. every row of op(A) and column of op(B) is scaled by a power of two and truncated to a `bits`-bit integer,
  2. the integer matrices are reduced modulo `nmod` pairwise coprime moduli p_i <= 256 (unsigned 8-bit residues),
  3. one exact INT8 x INT8 -> INT32 GEMM per modulus (the tensor-core part), reduced mod p_i,
  4. the residues are recombined by the Chinese remainder theorem: V/P = frac(sum_i r_i y_i / p_i), evaluated in
     96-bit fixed point (three 32-bit limbs, exactly like the kernel), centred, scaled back and rounded to float64.

Step 3 is exact, so the only error is the truncation of step 1: |err_ij| <= K 2^(1-bits) rowmax_i colmax_j (about).
The functions below follow the kernel's arithmetic limb for limb, so planes and residues compare bit for bit and the
final doubles compare exactly as well.
"""
import math

import numpy as np

MODULI = [256, 255, 253, 251, 247, 241, 239, 233, 229, 227, 223, 217, 211, 199, 197, 193, 191, 181, 179, 173]


def modulus_product(nmod):
    P = 1
    for p in MODULI[:nmod]:
        P *= p
    return P


def operand_bits(nmod, K):
    """Operand width b <= 63 with K * 2^(2b) < P/2 (the exact product must stay inside (-P/2, P/2)); the kernel's
    rule: 2b + ceil(log2 K) <= bitlength(P) - 2."""
    L = modulus_product(nmod).bit_length()
    lk = 0
    while (1 << lk) < K:
        lk += 1
    return min(63, (L - 2 - lk) // 2)


def crt_fractions(nmod):
    """f_i = floor(2^96 * y_i / p_i) with y_i = (P/p_i)^-1 mod p_i, as three 32-bit limbs (most significant first)."""
    P = modulus_product(nmod)
    out = []
    for p in MODULI[:nmod]:
        Mi = P // p
        y = pow(Mi % p, -1, p) if p > 1 else 0
        f = (y << 96) // p
        out.append(((f >> 64) & 0xFFFFFFFF, (f >> 32) & 0xFFFFFFFF, f & 0xFFFFFFFF))
    return out


def scale_exponents(X, bits):
    """Per row of X [R, K]: s with |x * 2^s| < 2^bits (s = bits - 1 - floor(log2 max|x|)); zero rows get s = 0."""
    mx = np.max(np.abs(X), axis=1)
    s = np.zeros(X.shape[0], dtype=np.int64)
    nz = mx > 0
    _, e = np.frexp(mx[nz])              # mx = m * 2^e, 0.5 <= m < 1  ->  floor(log2 mx) = e - 1
    s[nz] = bits - e
    return s


def to_integers(X, s):
    """trunc(x * 2^s) as Python ints (object array): exact."""
    R, K = X.shape
    out = np.empty((R, K), dtype=object)
    for i in range(R):
        for k in range(K):
            m, e = math.frexp(float(X[i, k]))
            v = int(m * (1 << 53))       # exact 54-bit integer
            sh = e - 53 + int(s[i])
            out[i, k] = (v << sh) if sh >= 0 else (abs(v) >> (-sh)) * (1 if v >= 0 else -1)
    return out


def residues(XI, nmod):
    """Unsigned residues [nmod, R, K] (uint8) of an integer matrix."""
    R, K = XI.shape
    out = np.empty((nmod, R, K), dtype=np.uint8)
    for a, p in enumerate(MODULI[:nmod]):
        out[a] = np.array([[int(v) % p for v in row] for row in XI], dtype=np.int64).astype(np.uint8)
    return out


def residue_gemm(RA, RB):
    """D[a] = (RA[a] @ RB[a]^T) mod p_a, uint8 -- the tensor-core part (exact int32 accumulation)."""
    nmod = RA.shape[0]
    out = np.empty((nmod, RA.shape[1], RB.shape[1]), dtype=np.uint8)
    for a, p in enumerate(MODULI[:nmod]):
        acc = RA[a].astype(np.int64) @ RB[a].astype(np.int64).T
        out[a] = (acc % p).astype(np.uint8)
    return out


def crt_combine(D, sA, sB, nmod):
    """Residues -> float64 exactly as the kernel does it: 96-bit fixed-point fraction, centred, times P 2^-(sA+sB)."""
    fr = crt_fractions(nmod)
    P = modulus_product(nmod)
    Pd = float(P)                         # correctly rounded
    M, N = D.shape[1], D.shape[2]
    out = np.empty((M, N))
    for i in range(M):
        for j in range(N):
            a2 = a1 = a0 = 0
            for a in range(nmod):
                r = int(D[a, i, j])
                a2 += r * fr[a][0]
                a1 += r * fr[a][1]
                a0 += r * fr[a][2]
            mid = a1 + (a0 >> 32)
            top = (a2 + (mid >> 32)) & 0xFFFFFFFF
            hi64 = (top << 32) | (mid & 0xFFFFFFFF)
            if hi64 >= 1 << 63:
                hi64 -= 1 << 64
            lo32 = a0 & 0xFFFFFFFF
            frac = float(hi64) * 2.0 ** -64 + float(lo32) * 2.0 ** -96
            out[i, j] = math.ldexp(frac * Pd, -int(sA[i]) - int(sB[j]))
    return out


def emulated_gemm(A, B, nmod, bits=None):
    """C = A @ B^T for A [M,K], B [N,K] (both 'K-major'), emulated; returns C and the intermediate pieces."""
    K = A.shape[1]
    if bits is None:
        bits = operand_bits(nmod, K)
    sA, sB = scale_exponents(A, bits), scale_exponents(B, bits)
    AI, BI = to_integers(A, sA), to_integers(B, sB)
    RA, RB = residues(AI, nmod), residues(BI, nmod)
    D = residue_gemm(RA, RB)
    return crt_combine(D, sA, sB, nmod), dict(sA=sA, sB=sB, RA=RA, RB=RB, D=D, AI=AI, BI=BI, bits=bits)


def exact_integer_product(AI, BI):
    M, N = AI.shape[0], BI.shape[0]
    out = np.empty((M, N), dtype=object)
    for i in range(M):
        for j in range(N):
            out[i, j] = sum(int(a) * int(b) for a, b in zip(AI[i], BI[j]))
    return out

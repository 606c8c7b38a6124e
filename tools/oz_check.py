"""Piece-by-piece check of the INT8 modular GEMM (csrc/gpe_ozaki.cu) on the GPU box, with diagnostics instead of asserts:
   python tools/oz_check.py [M N K batch layout nmod]
scale exponents, residue planes of both operands and of the product are compared bit for bit with NumPy integer
arithmetic; the final doubles with the CPU restatement (oracle/ozaki2_oracle.py) and with a float64 matmul."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gp_emu_uqsa_b200 import _lib  # noqa: E402
from oracle import ozaki2_oracle as oz  # noqa: E402


def residues_np(X, s, nmod):
    """X [R,K] float64, s [R] -> uint8 [nmod,R,K] with exact integer arithmetic (values < 2^63 as Python ints)."""
    V = np.trunc(np.ldexp(X, s[:, None].astype(np.int32)))
    Vi = np.array([[int(v) for v in row] for row in V], dtype=object)
    return np.stack([(Vi % p).astype(np.uint8) for p in oz.MODULI[:nmod]])


def modes(n, batch, nmod):
    """Every (layout, k range, lower) combination the factorisation uses, INT8 route against the DMMA kernels."""
    dev = _lib.Device(0)
    g = torch.Generator(device="cuda").manual_seed(11)
    tril = torch.tril(torch.ones(n, n, dtype=torch.float64, device="cuda"))
    scale = torch.exp(torch.empty(batch, n, 1, dtype=torch.float64, device="cuda").uniform_(-4, 4, generator=g))
    T = torch.randn(batch, n, n, dtype=torch.float64, device="cuda", generator=g) * tril * scale
    F = torch.randn(batch, n, n, dtype=torch.float64, device="cuda", generator=g) * scale
    sz = n * n
    cases = [("L21 = A21 Li11^T", F, T, 0, _lib.KM_LE_J, 0, 1.0, 0), ("T = L21 Li11", F, T, 1, _lib.KM_GE_J, 0, 1.0, 0),
             ("A22 -= L21 L21^T", F, F, 0, _lib.KM_FULL, 1, -1.0, 1), ("Li21 = -Li22 T", T, F, 1, _lib.KM_LE_I, 0, -1.0, 0),
             ("LAUUM", T, T, 2, _lib.KM_GE_I, 1, 1.0, 0), ("dense NN", F, F, 1, _lib.KM_FULL, 0, 1.0, 0)]
    ok = True
    for name, X, Y, layout, kmode, lower, alpha, acc in cases:
        C0 = torch.randn(batch, n, n, dtype=torch.float64, device="cuda", generator=g)
        C1, C2 = C0.clone(), C0.clone()
        dev.dbg_gemm(X, Y, C1, n, n, n, n, n, n, sz, sz, sz, alpha=alpha, accumulate=acc, kmode=kmode, lower=lower, batch=batch, layout=layout)
        dev.dbg_gemm_oz(X, Y, C2, n, n, n, n, n, n, sz, sz, sz, alpha=alpha, accumulate=acc, kmode=kmode, lower=lower, batch=batch,
                        layout=layout, nmod=nmod)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record(torch.cuda.ExternalStream(dev.stream_ptr))
        dev.dbg_gemm(X, Y, C1, n, n, n, n, n, n, sz, sz, sz, alpha=alpha, accumulate=0, kmode=kmode, lower=lower, batch=batch, layout=layout)
        ev[1].record(torch.cuda.ExternalStream(dev.stream_ptr))
        dev.dbg_gemm_oz(X, Y, C2, n, n, n, n, n, n, sz, sz, sz, alpha=alpha, accumulate=0, kmode=kmode, lower=lower, batch=batch,
                        layout=layout, nmod=nmod)
        ev[2].record(torch.cuda.ExternalStream(dev.stream_ptr))
        torch.cuda.synchronize()
        Xo = X if layout != 2 else X.transpose(1, 2)
        Yo = Y.transpose(1, 2) if layout == 0 else Y
        # the INT8 route's contract: error relative to (row max of X) x (column max of Y) x K 2^(1-bits); the element-wise
        # lower triangle is compared when only lower tiles are computed (the two routes use different tile sizes)
        sc = Xo.abs().amax(2, keepdim=True) * Yo.abs().amax(1, keepdim=True) * n + 1e-300
        sc2 = torch.matmul(Xo.abs(), Yo.abs()) + 1e-300
        d = (C1 - C2).abs()
        if lower:
            d = d * tril
        err, err2 = (d / sc).max().item(), (d / sc2).max().item()
        print(f"{name:20s} n={n} batch={batch}: max |dmma - int8| / (K rowmax colmax) = {err:.2e}, / (|X||Y|) = {err2:.2e};  "
              f"dmma {ev[0].elapsed_time(ev[1]):.3f} ms, int8 route {ev[1].elapsed_time(ev[2]):.3f} ms")
        ok &= err < 1e-15
    print("OK" if ok else "FAILED")
    dev.close()
    return 0 if ok else 1


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "modes":
        a = [int(x) for x in sys.argv[2:]]
        n, batch, nmod = (a + [1024, 2, 18][len(a):])[:3]
        return modes(n, batch, nmod)
    a = [int(x) for x in sys.argv[1:]]
    M, N, K, batch, layout, nmod = (a + [128, 256, 256, 2, 0, 18][len(a):])[:6]
    dev = _lib.Device(0)
    rng = np.random.default_rng(5)
    A = rng.standard_normal((batch, M, K)) * np.exp(rng.uniform(-6, 6, (batch, M, 1))) * np.exp(rng.uniform(-8, 0, (batch, M, K)))
    B = rng.standard_normal((batch, N, K)) * np.exp(rng.uniform(-6, 6, (batch, N, 1)))
    A[0, 3] = 0.0
    tA, tB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    As = tA.contiguous() if layout in (0, 1) else tA.transpose(1, 2).contiguous()
    Bs = tB.contiguous() if layout == 0 else tB.transpose(1, 2).contiguous()
    lda = K if layout in (0, 1) else M
    ldb = K if layout == 0 else N
    C = torch.zeros(batch, M, N, dtype=torch.float64, device="cuda")
    pA = torch.zeros(batch, nmod, M, K, dtype=torch.uint8, device="cuda")
    pB = torch.zeros(batch, nmod, N, K, dtype=torch.uint8, device="cuda")
    pD = torch.zeros(batch, nmod, M, N, dtype=torch.uint8, device="cuda")
    sA = torch.zeros(batch, M, dtype=torch.int32, device="cuda")
    sB = torch.zeros(batch, N, dtype=torch.int32, device="cuda")
    t0 = time.time()
    dev.dbg_gemm_oz(As, Bs, C, M, N, K, lda, ldb, N, sA=M * K, sB=N * K, sC=M * N, batch=batch, layout=layout, nmod=nmod,
                    planesA=pA, planesB=pB, planesD=pD, sexpA=sA, sexpB=sB)
    torch.cuda.synchronize()
    print(f"shape {M}x{N}x{K} batch {batch} layout {layout} nmod {nmod}: call returned in {time.time() - t0:.3f} s")
    bits = oz.operand_bits(nmod, K)
    ok = True
    for b in range(batch):
        esA, esB = oz.scale_exponents(A[b], bits), oz.scale_exponents(B[b], bits)
        gA, gB = sA[b].cpu().numpy(), sB[b].cpu().numpy()
        print(f"item {b}: exponent mismatches A {np.sum(esA != gA)} B {np.sum(esB != gB)}")
        RA, RB = residues_np(A[b], esA, nmod), residues_np(B[b], esB, nmod)
        mA = RA != pA[b].cpu().numpy()
        mB = RB != pB[b].cpu().numpy()
        print(f"  plane mismatches A {mA.sum()} of {mA.size} (per modulus {mA.reshape(nmod, -1).sum(1).tolist()}), B {mB.sum()}")
        D = oz.residue_gemm(RA, RB)
        gD = pD[b].cpu().numpy()
        mD = D != gD
        print(f"  product residue mismatches {mD.sum()} of {mD.size} (per modulus {mD.reshape(nmod, -1).sum(1).tolist()})")
        if mD.any():
            pl = int(np.argmax(mD.reshape(nmod, -1).sum(1)))
            rows = np.where(mD[pl].any(1))[0]
            cols = np.where(mD[pl].any(0))[0]
            print(f"    plane {pl}: bad rows {rows[:16].tolist()} ... ({len(rows)}), bad cols {cols[:16].tolist()} ... ({len(cols)})")
            i, j = rows[0], cols[0]
            print(f"    e.g. ({i},{j}): got {gD[pl, i, j]} want {D[pl, i, j]}")
        sub = slice(0, min(M, 16)), slice(0, min(N, 24))
        Cc = oz.crt_combine(D[:, sub[0], sub[1]], esA[sub[0]], esB[sub[1]], nmod)
        gC = C[b].cpu().numpy()
        ex = gC[sub] == Cc
        print(f"  CRT doubles equal to the CPU restatement on a {Cc.shape} corner: {ex.all()} ({(~ex).sum()} differ)")
        ref = A[b] @ B[b].T
        sc = np.abs(A[b]) @ np.abs(B[b]).T + 1e-300
        err = np.max(np.abs(gC - ref) / sc)
        print(f"  max |C - A B^T| / (|A||B|^T) = {err:.3e}")
        ok &= not (mA.any() or mB.any() or mD.any()) and ex.all() and err < 3e-15
    print("OK" if ok else "FAILED")
    dev.close()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())

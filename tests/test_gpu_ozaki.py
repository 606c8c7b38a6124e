"""GPU parity of the INT8 modular FP64 GEMM (csrc/gpe_ozaki.cu) through the C-ABI debug entry gpe_dbg_gemm_oz:
scale exponents, residue planes and product residues bit for bit against integer arithmetic on the host, the CRT output
exactly against the CPU restatement (oracle/ozaki2_oracle.py), and every (layout, k range, lower) combination of the
factorisation against the FP64 DMMA kernels at the INT8 route's error contract."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
from oracle import ozaki2_oracle as oz  # noqa: E402


@pytest.fixture(scope="module")
def dev():
    from gp_emu_uqsa_b200 import _lib
    d = _lib.Device(0)
    yield d
    d.close()


def _residues(X, s, nmod):
    V = np.trunc(np.ldexp(X, s[:, None].astype(np.int32)))       # integers below 2^63 with at most 53 significant bits: exact
    Vi = np.array([[int(v) for v in row] for row in V], dtype=object)
    return np.stack([(Vi % p).astype(np.uint8) for p in oz.MODULI[:nmod]])


# (M, N, K, batch, layout, nmod): one CTA per tile (M = 128) and the two-CTA multicast clusters (M % 256 == 0)
@pytest.mark.parametrize("M,N,K,batch,layout,nmod", [(128, 256, 256, 2, 0, 18), (256, 512, 384, 1, 2, 17), (256, 256, 256, 1, 1, 16),
                                                     (384, 256, 128, 1, 0, 16)])
def test_pieces_bit_exact(dev, M, N, K, batch, layout, nmod):
    rng = np.random.default_rng(M + N + K + nmod)
    A = rng.standard_normal((batch, M, K)) * np.exp(rng.uniform(-6, 6, (batch, M, 1))) * np.exp(rng.uniform(-8, 0, (batch, M, K)))
    B = rng.standard_normal((batch, N, K)) * np.exp(rng.uniform(-6, 6, (batch, N, 1)))
    A[0, 3] = 0.0                                                # a zero row: exponent 0, residues 0
    tA, tB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    As = tA.contiguous() if layout in (0, 1) else tA.transpose(1, 2).contiguous()
    Bs = tB.contiguous() if layout == 0 else tB.transpose(1, 2).contiguous()
    lda, ldb = (K if layout in (0, 1) else M), (K if layout == 0 else N)
    C = torch.zeros(batch, M, N, dtype=torch.float64, device="cuda")
    pA = torch.zeros(batch, nmod, M, K, dtype=torch.uint8, device="cuda")
    pB = torch.zeros(batch, nmod, N, K, dtype=torch.uint8, device="cuda")
    pD = torch.zeros(batch, nmod, M, N, dtype=torch.uint8, device="cuda")
    sA = torch.zeros(batch, M, dtype=torch.int32, device="cuda")
    sB = torch.zeros(batch, N, dtype=torch.int32, device="cuda")
    dev.dbg_gemm_oz(As, Bs, C, M, N, K, lda, ldb, N, sA=M * K, sB=N * K, sC=M * N, batch=batch, layout=layout, nmod=nmod,
                    planesA=pA, planesB=pB, planesD=pD, sexpA=sA, sexpB=sB)
    torch.cuda.synchronize()
    bits = oz.operand_bits(nmod, K)
    for b in range(batch):
        esA, esB = oz.scale_exponents(A[b], bits), oz.scale_exponents(B[b], bits)
        assert (esA == sA[b].cpu().numpy()).all() and (esB == sB[b].cpu().numpy()).all()
        RA, RB = _residues(A[b], esA, nmod), _residues(B[b], esB, nmod)
        assert (RA == pA[b].cpu().numpy()).all() and (RB == pB[b].cpu().numpy()).all()
        D = oz.residue_gemm(RA, RB)
        assert (D == pD[b].cpu().numpy()).all()                  # TMA + tcgen05.mma kind::i8 + mod p: exact
        gC = C[b].cpu().numpy()
        want = oz.crt_combine(D[:, :16, :24], esA[:16], esB[:24], nmod)
        assert (gC[:16, :24] == want).all()                      # CRT in 96-bit fixed point: the same doubles
        ref = A[b] @ B[b].T
        assert np.max(np.abs(gC - ref) / (np.abs(A[b]) @ np.abs(B[b]).T + 1e-300)) < 3e-15


@pytest.mark.parametrize("n,batch,nmod", [(1024, 2, 16), (2048, 2, 16), (1024, 3, 18)])
def test_every_product_of_the_factorisation_against_dmma(dev, n, batch, nmod):
    from gp_emu_uqsa_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(11)
    tril = torch.tril(torch.ones(n, n, dtype=torch.float64, device="cuda"))
    scale = torch.exp(torch.empty(batch, n, 1, dtype=torch.float64, device="cuda").uniform_(-4, 4, generator=g))
    T = torch.randn(batch, n, n, dtype=torch.float64, device="cuda", generator=g) * tril * scale
    F = torch.randn(batch, n, n, dtype=torch.float64, device="cuda", generator=g) * scale
    sz = n * n
    cases = [(F, T, 0, _lib.KM_LE_J, 0, 1.0, 0), (F, T, 1, _lib.KM_GE_J, 0, 1.0, 0), (F, F, 0, _lib.KM_FULL, 1, -1.0, 1),
             (T, F, 1, _lib.KM_LE_I, 0, -1.0, 0), (T, T, 2, _lib.KM_GE_I, 1, 1.0, 0), (F, F, 1, _lib.KM_FULL, 0, 1.0, 0)]
    for X, Y, layout, kmode, lower, alpha, acc in cases:
        C0 = torch.randn(batch, n, n, dtype=torch.float64, device="cuda", generator=g)
        C1, C2 = C0.clone(), C0.clone()
        dev.dbg_gemm(X, Y, C1, n, n, n, n, n, n, sz, sz, sz, alpha=alpha, accumulate=acc, kmode=kmode, lower=lower, batch=batch,
                     layout=layout)
        dev.dbg_gemm_oz(X, Y, C2, n, n, n, n, n, n, sz, sz, sz, alpha=alpha, accumulate=acc, kmode=kmode, lower=lower, batch=batch,
                        layout=layout, nmod=nmod)
        torch.cuda.synchronize()
        Xo = X if layout != 2 else X.transpose(1, 2)
        Yo = Y.transpose(1, 2) if layout == 0 else Y
        # the contract of the INT8 route: error relative to (row max of X)(column max of Y) K 2^(1-bits)
        sc = Xo.abs().amax(2, keepdim=True) * Yo.abs().amax(1, keepdim=True) * n + 1e-300
        d = (C1 - C2).abs()
        if lower:
            d = d * tril                 # the two routes compute different tiles above the diagonal
        assert (d / sc).max().item() < 1e-15, (layout, kmode, lower)

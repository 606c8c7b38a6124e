"""CPU-only checks: the C-ABI library loads and exports every symbol include/gpe_b200.h declares
(no compute calls without a GPU), the host-side mirror of the reference's plumbing, the lock-step
batched L-BFGS-B driver, and the world-size-2 (gloo) sharding logic."""
import contextlib
import ctypes
import io
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@contextlib.contextmanager
def _quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "gpe_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(gpe_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_header_symbol():
    from gp_emu_uqsa_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "libgpe_b200.so is not built (run __graft_entry__.build())"
    L = ctypes.CDLL(_lib.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), "library does not export " + s
    assert set(_lib.EXPORTS) == set(syms), "gp_emu_uqsa_b200/_lib.py EXPORTS and include/gpe_b200.h disagree"
    _lib.load()                                   # prototypes resolve
    assert L.gpe_version() >= 100


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    from gp_emu_uqsa_b200 import _lib
    with pytest.raises(_lib.GpeError):
        _lib.Device(0)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "gp_emu_uqsa_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S), f + " references oracle/"


def test_config_beliefs_parsing_and_split(golden_dir, tmp_path):
    from gp_emu_uqsa_b200 import _emulatorclasses as C
    src = os.path.join(golden_dir, "toy-sim")
    old = os.getcwd()
    os.chdir(src)
    try:
        with _quiet():
            cfg = C.Config("toy-sim_config")
            bel = C.Beliefs(cfg.beliefs)
            par = C.Hyperparams(bel)
            basis = C.Basis(bel)
            tv = C.TV_config(*cfg.tv_config)
            np.random.seed(0)
            data = C.All_Data(cfg.inputs, cfg.outputs, tv, bel, par, True, True)
    finally:
        os.chdir(old)
    assert cfg.tv_config == [10, 0, 2] and cfg.tries == 10 and cfg.constraints == "none"
    assert bel.mucm == "T" and bel.fix_nugget == "T" and bel.alt_nugget == "F" and bel.basis_str == ["1.0", "x"]
    assert basis.poly == [1] and basis.meanf == "m(x) = b + b0x[0]"
    assert (data.T, data.V) == (48, 6)
    gold = np.load(os.path.join(golden_dir, "toysim.npz"))
    # the reference's final training set is all 60 shuffled points with validation sets folded in front
    assert sorted(map(tuple, data.x_full)) == sorted(map(tuple, gold["X"]))
    xT, yT = data.choose_T()
    xV, yV = data.choose_V()
    assert xT.shape == (48, 2) and xV.shape == (6, 2)
    H = basis.design_matrix(xT)
    assert np.array_equal(H[:, 0], np.ones(48)) and np.array_equal(H[:, 1], xT[:, 0])


def test_make_inputs_grids():
    from gp_emu_uqsa_b200._emulatorplotting import make_inputs
    x = make_inputs(3, 30, 30, [0, 2], [1], [0.25], False, [[0.0, 1.0], [2.0, 3.0]])
    assert x.shape == (900, 3) and np.all(x[:, 1] == 0.25)
    assert x[31, 0] == np.linspace(0, 1, 30)[1] and x[31, 2] == np.linspace(2, 3, 30)[1]
    x1 = make_inputs(3, 30, 30, [1], [0, 2], [0.1, 0.9], True, [[0.0, 1.0]])
    assert x1.shape == (900, 3) and np.all(x1[:, 0] == 0.1) and np.all(x1[:, 2] == 0.9) and x1[-1, 1] == 1.0


def test_latin_hypercube_consumes_rng_like_reference(golden_dir, tmp_path):
    """Same seed -> the design the real reference produced inside imp_plot (golden lhc_*)."""
    import gp_emu_uqsa_b200.design_inputs as d
    gold = np.load(os.path.join(golden_dir, "hmapi_n100_d3.npz"))
    old = os.getcwd()
    os.chdir(tmp_path)
    try:
        with _quiet():
            np.random.seed(77)
            for s_ in ([0, 1], [0, 2], [1, 2]):
                d.optLatinHyperCube(1, 30, 15, [[0.0, 1.0]], "f")
                assert np.array_equal(np.loadtxt("f"), gold["lhc_%d_%d" % tuple(s_)])
    finally:
        os.chdir(old)


def test_lockstep_lbfgsb_matches_sequential_scipy():
    """Every start must follow exactly the path scipy.optimize.minimize takes when run alone."""
    from scipy.optimize import minimize
    from gp_emu_uqsa_b200._lbfgsb_batch import minimize_batch
    rng = np.random.default_rng(0)
    Q = rng.normal(size=(5, 5)); Q = Q @ Q.T + np.eye(5)
    c = rng.normal(size=5)

    def fg(x):
        return float(0.5 * x @ Q @ x - c @ x + 0.1 * np.sum(x ** 4)), Q @ x - c + 0.4 * x ** 3

    calls = []

    def eval_batch(X):
        calls.append(len(X))
        out = [fg(x) for x in X]
        ok = np.array([x[0] < 50.0 for x in X])        # starts that wander to x0 > 50 are "non-PD"
        return np.array([o[0] for o in out]), np.array([o[1] for o in out]), ok

    x0s = rng.normal(size=(12, 5)) * 3
    x0s[4, 0] = 60.0
    bounds = [[-2.0, 2.0]] * 5
    half_open = [[None, 1.0], [0.0, None], [None, None], [-1, 1], [-3, 0.5]]
    for driver in ("rc", "threads", None):             # reverse-communication, threaded, and the default choice
        for bnd in (None, bounds, half_open):
            x0 = x0s.copy()
            if bnd is not None:
                x0[4, 0] = 1.0
            res, rounds, evals = minimize_batch(eval_batch, x0, bnd, driver=driver)
            for i in range(12):
                if bnd is None and i == 4:
                    assert res[i] is None
                    continue
                ref = minimize(fg, list(x0[i]), method="L-BFGS-B", jac=True, **({} if bnd is None else {"bounds": bnd}))
                assert res[i].nfev == ref.nfev and res[i].nit == ref.nit and res[i].status == ref.status
                assert np.array_equal(res[i].x, ref.x) and res[i].fun == ref.fun and np.array_equal(res[i].jac, ref.jac)
            assert evals == sum(calls[-rounds:]) and rounds >= 3


def test_lockstep_lbfgsb_error_reaches_caller():
    from gp_emu_uqsa_b200._lbfgsb_batch import minimize_batch

    def boom(X):
        raise RuntimeError("device error")

    for driver in ("rc", "threads"):
        with pytest.raises(RuntimeError, match="device error"):
            minimize_batch(boom, np.ones((4, 3)), None, driver=driver)


_GLOO_WORKER = r"""
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, %r)
from gp_emu_uqsa_b200 import _dist
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%s" %% sys.argv[1], rank=int(sys.argv[2]), world_size=2)
rank, world = _dist.rank_world()
assert (rank, world) == (int(sys.argv[2]), 2)
# multistart table: every rank fills its block of guesses; all ranks end with the same full table
B, p = 7, 3
lo, hi = _dist.block(B, rank, world)
full = np.arange(B * (p + 2), dtype=float).reshape(B, p + 2) + 0.5
pack = np.full((B, p + 2), np.nan)
pack[lo:hi] = full[lo:hi]
got = _dist.gather_blocks(pack, B)
assert np.array_equal(got, full), got
# the reference's selection rule (first strictly smaller, in guess order) gives the same winner everywhere
fun = np.array([3.0, 1.0, 2.0, 1.0, 5.0, 0.5, 0.5])
best, first = None, True
for C in range(B):
    if first or fun[C] < best[0]:
        best, first = (fun[C], C), False
assert best[1] == 5
# implausibility cell statistics: min / sum over ranks
ncell = 6
c0, c1 = _dist.block(ncell, rank, world)
imp = np.full((ncell, 2), np.inf); cnt = np.zeros((ncell, 2), dtype=np.uint64)
imp[c0:c1] = np.arange(c0, c1)[:, None] + np.array([0.25, 0.75])
cnt[c0:c1] = np.arange(c0, c1)[:, None] + np.array([1, 2], dtype=np.uint64)
imp = _dist.all_reduce(imp, "min"); cnt = _dist.all_reduce(cnt, "sum")
assert np.array_equal(imp, np.arange(ncell)[:, None] + np.array([0.25, 0.75]))
assert cnt.dtype == np.uint64 and np.array_equal(cnt, np.arange(ncell)[:, None] + np.array([1, 2], dtype=np.uint64))
# keep-mask of a row-partitioned point set: concatenation in ascending flat index
n = 11
lo, hi = _dist.block(n, rank, world)
keep = np.zeros(n); keep[lo:hi] = (np.arange(lo, hi) %% 3 == 0)
keep = _dist.gather_blocks(keep, n) > 0.5
assert np.array_equal(np.nonzero(keep)[0], np.array([0, 3, 6, 9]))
# a NaN objective travels through the gather as a NaN (it must not become 0.0, the best possible value)
pack = np.full((B, p + 2), -1.0)
lo, hi = _dist.block(B, rank, world)
pack[lo:hi, 1] = np.where(np.arange(lo, hi) == 4, np.nan, np.arange(lo, hi))
got = _dist.gather_blocks(pack, B)
assert np.isnan(got[4, 1]) and np.array_equal(np.delete(got[:, 1], 4), np.delete(np.arange(B, dtype=float), 4))
# point-sharded ranges: tile-aligned interior boundaries, the same rows back in flat order
n = 1000
lo, hi = _dist.block_aligned(n, rank, world)
assert (lo, hi) == ((0, 512) if rank == 0 else (512, 1000))
keep = np.zeros(n); keep[lo:hi] = (np.arange(lo, hi) %% 7 == 0)
keep = _dist.gather_blocks(keep, n, bounds=_dist.block_aligned) > 0.5
assert np.array_equal(np.nonzero(keep)[0], np.arange(0, n, 7))
# an unseeded run: every rank continues from rank 0's generator state (shuffle, guesses, Latin hypercubes)
np.random.seed(1234 + rank)
_dist.sync_numpy_rng()
draw = np.random.random_sample(3)
both = _dist.gather_blocks(np.tile(draw, (2, 1)), 2)
assert np.array_equal(both[0], both[1])
assert _dist.is_writer() == (rank == 0)
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_world_size_2_gloo_sharding(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER % ROOT)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), port, str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in (0, 1)]
    outs = [p.communicate(timeout=180)[0].decode() for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and ("rank %d ok" % r) in o, o


def test_header_is_plain_c(tmp_path):
    """The C-ABI header must be consumable from C (cgo / JNI / ctypes-style bindings): plain pointers and
    sizes, no C++ or torch types."""
    src = tmp_path / "t.c"
    src.write_text('#include "gpe_b200.h"\nint main(void) { gpe_handle* h = 0; (void)h; return GPE_MODE_MUCM | GPE_MODE_ALT_NUGGET | GPE_MODE_NUGGET_FREE; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_block_aligned_partition_properties():
    """Point ranges of the prediction / implausibility shards: contiguous, exhaustive, interior boundaries on
    128-point tiles, sizes within one tile of each other."""
    from gp_emu_uqsa_b200 import _dist
    rng = np.random.default_rng(0)
    for _ in range(300):
        n, world = int(rng.integers(0, 10 ** 9)), int(rng.integers(1, 65))
        edges = [_dist.block_aligned(n, r, world) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == n
        for (lo, hi), (lo2, _) in zip(edges, edges[1:]):
            assert hi == lo2 and lo <= hi and (hi % 128 == 0 or hi == n)
        sizes = [hi - lo for lo, hi in edges]
        assert max(sizes) - min(sizes) <= 128 + 127
    assert _dist.block_aligned(10 ** 8, 3, 8) == (37499904, 50000000)


def test_data_fingerprint_sees_permutations_and_swaps():
    """The upload key of a data set: moments are not enough -- reordered rows, swapped outputs, a swapped
    pair of r entries must all change it."""
    from gp_emu_uqsa_b200 import _emulatorclasses as C
    rng = np.random.default_rng(1)
    a = rng.random((50, 3))
    base = C._checksum(a)
    assert C._checksum(a.copy()) == base
    perm = a[rng.permutation(50)]
    assert C._checksum(perm) != base and abs(perm.sum() - a.sum()) < 1e-12
    sw = a.copy(); sw[[0, 1]] = sw[[1, 0]]
    assert C._checksum(sw) != base
    neg = a.copy(); neg[3, 1] = -neg[3, 1]
    assert C._checksum(neg) != base


def test_block_partition_properties():
    """Units (guesses, grid cells, rows) are split into contiguous blocks that tile range(n) exactly, differ in size
    by at most one and may be empty when there are fewer units than ranks (the C-ABI treats zero units as a no-op)."""
    from hypothesis import given, settings, strategies as st
    from gp_emu_uqsa_b200 import _dist

    @settings(max_examples=300, deadline=None)
    @given(st.integers(0, 10 ** 9), st.integers(1, 64))
    def check(n, world):
        edges = [_dist.block(n, r, world) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == n
        for (lo, hi), (lo2, _) in zip(edges, edges[1:]):
            assert lo <= hi == lo2
        sizes = [hi - lo for lo, hi in edges]
        assert max(sizes) - min(sizes) <= 1 and sum(sizes) == n

    check()
    # single-process defaults: no process group -> rank 0 of 1, collectives are identities
    assert _dist.rank_world() == (0, 1)
    x = np.arange(6.0).reshape(3, 2)
    assert np.array_equal(_dist.gather_blocks(x.copy(), 3), x)


def test_int8_slice_gemm_emulation_reaches_float64_accuracy():
    """tools/ozaki_study.py (the CPU parity study behind DESIGN section 11): its exact INT8-slice GEMM emulation must
    converge to the float64 product as slices are added -- 2^-7 per slice, float64 level at 9-10 slices -- also for
    rows whose entries span many decades, and the blocked factorisation built on it must reproduce LAPACK."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ozaki_study", os.path.join(ROOT, "tools", "ozaki_study.py"))
    oz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(oz)
    rng = np.random.default_rng(0)
    A = rng.normal(size=(96, 160)) * 10.0 ** rng.integers(-6, 3, size=(96, 1)) * 10.0 ** (-8 * rng.random((96, 160)))
    B = rng.normal(size=(160, 64)) * 10.0 ** rng.integers(-3, 4, size=(1, 64))
    ref = A @ B
    bound = np.abs(A) @ np.abs(B)
    errs = [np.max(np.abs(oz.make_gemm_ozaki(s)(A, B) - ref) / bound) for s in (6, 8, 10)]
    assert errs[0] > errs[1] > errs[2] and errs[1] < 1e-10 and errs[2] < 5e-15, errs
    X, y = oz.synth(300, 4)
    Acov = oz.cov(X, np.full(4, 0.7), 1e-4)
    Li, Ainv, ld = oz.factor_inverse(Acov.copy(), oz.make_gemm_ozaki(10))
    L = np.linalg.cholesky(Acov)
    assert abs(ld - 2 * np.log(np.diag(L)).sum()) <= 1e-12 * abs(ld)
    assert np.abs(Li @ L - np.eye(300)).max() < 1e-9


def test_modular_gemm_restatement_is_exact_up_to_the_operand_truncation():
    """oracle/ozaki2_oracle.py (the CPU restatement the GPU test of csrc/gpe_ozaki.cu compares with bit for bit): the CRT in
    96-bit fixed point returns the exact integer product of the truncated operands, rounded once; against the true product
    the error is at float64 level relative to |A||B| for 16, 17 and 18 moduli."""
    import math
    from fractions import Fraction
    from oracle import ozaki2_oracle as oz
    assert all(math.gcd(p, q) == 1 for i, p in enumerate(oz.MODULI) for q in oz.MODULI[:i])
    assert [oz.operand_bits(m, 4096) for m in (16, 17, 18)] == [56, 59, 63]
    rng = np.random.default_rng(0)
    M, N, K = 10, 9, 48
    A = rng.standard_normal((M, K)) * np.exp(rng.uniform(-8, 8, (M, 1))) * np.exp(rng.uniform(-6, 0, (M, K)))
    B = rng.standard_normal((N, K)) * np.exp(rng.uniform(-8, 8, (N, 1)))
    A[3] = 0.0
    true = np.array([[float(sum(Fraction(float(a)) * Fraction(float(b)) for a, b in zip(A[i], B[j]))) for j in range(N)]
                     for i in range(M)])
    sc = np.abs(A) @ np.abs(B).T + 1e-300
    for nmod in (16, 17, 18):
        C, aux = oz.emulated_gemm(A, B, nmod)
        E = oz.exact_integer_product(aux["AI"], aux["BI"])
        want = np.array([[math.ldexp(float(E[i, j]), -int(aux["sA"][i]) - int(aux["sB"][j])) if E[i, j] else 0.0
                          for j in range(N)] for i in range(M)])
        assert np.max(np.abs(C - want) / np.maximum(np.abs(want), 1e-300)) < 4e-16      # P is rounded to a double once
        assert np.max(np.abs(C - true) / sc) < 3e-16
        assert (C[3] == 0.0).all()

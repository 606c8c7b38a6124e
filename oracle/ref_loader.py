"""TEST INFRASTRUCTURE ONLY -- imports the *real* reference from /root/reference.

Only usable in the build container (the GPU box has no /root/reference): used by
``tests/golden/make_golden.py`` to generate the committed golden vectors and by
``tests/test_oracle_vs_reference.py`` (skipped when the reference is absent) to pin the
NumPy restatement.  Recipe: SURVEY.md Appendix A.1.
"""
import contextlib
import io
import os
import sys
import types

REF_ROOT = os.environ.get("GPE_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "gp_emu_uqsa"))


def load():
    """Return (g, h, s, gn) = the reference's top-level, history_match, sensitivity and
    noise_fit modules, with matplotlib/cycler stubbed and numpy.int restored."""
    import numpy as np
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    if "cycler" not in sys.modules:
        cyc = types.ModuleType("cycler")
        cyc.cycler = lambda *a, **k: None
        sys.modules["cycler"] = cyc
    if not hasattr(np, "int"):
        np.int = int                              # design_inputs.py:55
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    with contextlib.redirect_stdout(io.StringIO()):
        import gp_emu_uqsa as g
        import gp_emu_uqsa.history_match as h
        import gp_emu_uqsa.sensitivity as s
        import gp_emu_uqsa.noise_fit as gn
    return g, h, s, gn


def write_emulator_files(workdir, X, y, *, mucm, fix_nugget, alt_nugget, nugget=1e-4, delta=None,
                         sigma=1.0, tries=1, constraints="bounds", tv_config="10 0 0", name="syn",
                         basis="linear"):
    """Write config/beliefs/inputs/outputs text files for a synthetic emulator."""
    import numpy as np
    n, d = X.shape
    delta = [0.5] * d if delta is None else list(delta)
    os.makedirs(workdir, exist_ok=True)
    p = lambda f: os.path.join(workdir, f)
    np.savetxt(p(name + "_input"), X, fmt="%.17g")
    np.savetxt(p(name + "_output"), np.atleast_2d(y).T if y.ndim == 1 else y, fmt="%.17g")
    if basis == "linear":
        bstr = "1.0 " + " ".join(["x"] * d)
        binf = "NA " + " ".join(str(i) for i in range(d))
        beta = " ".join(["1.0"] * (d + 1))
    else:
        bstr, binf, beta = "1.0", "NA", "1.0"
    with open(p(name + "_beliefs"), "w") as f:
        f.write("active all\noutput 0\n")
        f.write("basis_str " + bstr + "\n")
        f.write("basis_inf " + binf + "\n")
        f.write("beta " + beta + "\n")
        f.write("delta " + " ".join(repr(float(v)) for v in delta) + "\n")
        f.write("sigma %r\nnugget %r\n" % (float(sigma), float(nugget)))
        f.write("fix_nugget %s\nalt_nugget %s\nmucm %s\n" % (fix_nugget, alt_nugget, mucm))
    with open(p(name + "_config"), "w") as f:
        f.write("beliefs %s_beliefs\ninputs %s_input\noutputs %s_output\n" % (name, name, name))
        f.write("tv_config %s\ndelta_bounds [ ]\nsigma_bounds [ ]\nnugget_bounds [ ]\n" % tv_config)
        f.write("tries %d\nconstraints %s\n" % (tries, constraints))
    return name + "_config"


@contextlib.contextmanager
def cwd(path):
    old = os.getcwd()
    os.makedirs(path, exist_ok=True)
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield

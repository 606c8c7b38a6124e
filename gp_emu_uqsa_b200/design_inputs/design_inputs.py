"""Maximin Latin-hypercube design (reference: gp_emu_uqsa/design_inputs/design_inputs.py).

The design is driven by the global NumPy RNG and must consume it exactly like the reference (one
``uniform`` and one ``shuffle`` per dimension per candidate design) so that history matching, which
calls this inside its loops, reproduces the reference's designs under a fixed seed; that part stays on
the host.  The O(N n^2 dim) maximin criterion of the N candidates is what costs time once prediction
is fast: ``device_criterion`` evaluates it for all candidates in one launch
(``gpe_pdist_argmin``, bit-identical to scipy's pdist + argmin) and is what history_match / noise_fit
pass in; called on its own, ``optLatinHyperCube`` uses scipy like the reference (it is a host utility
that works without a GPU and is not part of the hot path)."""
import numpy as _np
import scipy.spatial.distance as _dist

__all__ = ["optLatinHyperCube"]


def _host_criterion(designs, fextra):
    """argmin of the condensed squared-distance vector per candidate design (reference :73)."""
    out = _np.empty(designs.shape[0], dtype=_np.int64)
    for k in range(designs.shape[0]):
        xt = _np.concatenate([designs[k], fextra]) if fextra is not None else designs[k]
        out[k] = _np.argmin(_dist.pdist(xt, 'sqeuclidean'))
    return out


def device_criterion(designs, fextra):
    """The same criterion for all candidates in one device launch (gpe_pdist_argmin)."""
    from .. import _lib
    return _lib.scratch_device().pdist_argmin(designs, fextra)


def optLatinHyperCube(dim=None, n=None, N=None, minmax=None, filename="inputs", fextra=None, _criterion=None, _return=False):
    """Generate N random Latin hypercubes of n points in `dim` dimensions, keep the "best", scale it
    to `minmax` and save it to `filename` ('%.8f' text).  Returns None (with ``_return=True``, the
    package's own callers get the design exactly as the file holds it -- the '%.8f' text parsed back -- so
    that ranks of a multi-GPU job need not re-read a file another rank is writing).

    Multi-rank: the draws start from rank 0's generator state on every rank (``_dist.sync_numpy_rng``) and
    only rank 0 writes the file.

    Selection rule kept from the reference (:73-77, a known quirk): the criterion compared between
    designs is ``argmin(pdist)`` -- the *index* of the closest pair -- not the minimum distance."""
    print('dim:', dim)
    print('n:', n)
    print('N:', N)
    print('minmax:', minmax)
    print('filename:', filename)
    if dim is None or n is None or N is None or minmax is None:
        print("Please supply values for function arguments (default for filename is \"inputs\")")
    if len(minmax) != dim:
        print("WARNING: length of 'minmax' (list of lists) must equal 'dim'")
        raise SystemExit(1)
    if fextra is not None:
        print("\nGenerating", N, "oLHC samples of", n, "points, combining with supplied extra data, and checking maximin "
              "criterion (pick design with maximum minimum distance between design points)...")
    else:
        print("\nGenerating", N, "oLHC samples of", n, "points and checking maximin criterion (pick design with maximum "
              "minimum distance between design points)...")
    if N < 1 or n < 1:      # the reference dies on an unbound name here (e.g. 2-input imp_plot: dim = 0)
        raise UnboundLocalError("optLatinHyperCube: no candidate design was generated (N = 0)")
    from .. import _dist as _d
    _d.sync_numpy_rng()
    # all candidates first (the RNG draws do not depend on the criterion), then the criterion in one go
    designs = _np.empty((N, n, dim))
    for k in range(0, N):
        for i in range(0, dim):
            u = _np.random.uniform(0.0, 1.0, n)
            b = _np.arange(0, n, 1)
            _np.random.shuffle(b)
            designs[k, :, i] = (b + u) / float(n)
    fx = None if fextra is None else _np.ascontiguousarray(fextra, dtype=float).reshape(-1, dim)
    if n + (0 if fx is None else fx.shape[0]) < 2:
        crit = _np.zeros(N, dtype=_np.int64)
    else:
        crit = (_criterion or _host_criterion)(designs, fx)
    best_k, best_maximin = 0, crit[0]
    for k in range(1, N):
        if crit[k] > best_maximin:
            best_k, best_maximin = k, crit[k]
    best_D = _np.copy(designs[best_k])
    D = best_D
    print("Optimal LHC design was no.", best_k)
    print("Saving inputs to file...")
    inputs = _np.array(minmax)
    for i in range(0, dim):
        D[:, i] = D[:, i] * (inputs[i, 1] - inputs[i, 0]) + inputs[i, 0]
    if _d.is_writer():
        _np.savetxt(filename, D, delimiter=" ", fmt='%.8f')
    _d.barrier()
    print("DONE!")
    if _return:       # what np.loadtxt(filename) gives: the values rounded to the file's 8 decimals
        return _np.array([[float('%.8f' % v) for v in row] for row in D]).reshape(n, dim)
    return None

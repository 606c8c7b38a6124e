"""Maximin Latin-hypercube design (reference: gp_emu_uqsa/design_inputs/design_inputs.py).

Host code: the design is driven by the global NumPy RNG and must consume it exactly like the
reference (one ``uniform`` and one ``shuffle`` per dimension per candidate design) so that history
matching, which calls this inside its loops, reproduces the reference's designs under a fixed seed.
It is a "next" row of the scope table, not part of the GPU hot path."""
import numpy as _np
import scipy.spatial.distance as _dist

__all__ = ["optLatinHyperCube"]


def optLatinHyperCube(dim=None, n=None, N=None, minmax=None, filename="inputs", fextra=None):
    """Generate N random Latin hypercubes of n points in `dim` dimensions, keep the "best", scale it
    to `minmax` and save it to `filename` ('%.8f' text).  Returns None.

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
    x = _np.zeros((n, dim))
    best_D, best_k, best_maximin = None, None, None
    for k in range(0, N):
        for i in range(0, dim):
            u = _np.random.uniform(0.0, 1.0, n)
            b = _np.arange(0, n, 1)
            _np.random.shuffle(b)
            x[:, i] = (b + u) / float(n)
        xt = _np.concatenate([x, fextra]) if fextra is not None else x
        maximin = _np.argmin(_dist.pdist(xt, 'sqeuclidean'))
        if k == 0 or maximin > best_maximin:
            best_D, best_k, best_maximin = _np.copy(x), k, maximin
    if best_D is None:      # N == 0: the reference dies on an unbound name here (e.g. 2-input imp_plot)
        raise UnboundLocalError("optLatinHyperCube: no candidate design was generated (N = 0)")
    D = best_D
    print("Optimal LHC design was no.", best_k)
    print("Saving inputs to file...")
    inputs = _np.array(minmax)
    for i in range(0, dim):
        D[:, i] = D[:, i] * (inputs[i, 1] - inputs[i, 0]) + inputs[i, 0]
    _np.savetxt(filename, D, delimiter=" ", fmt='%.8f')
    print("DONE!")
    return None

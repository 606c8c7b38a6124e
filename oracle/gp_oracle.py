"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the GP_emu_UQSA dense-GP hot path.

A functional NumPy/SciPy restatement of the reference algorithm.  Every function cites
the reference file:line (relative to /root/reference/gp_emu_uqsa/) whose arithmetic it
follows, *including* the reference's quirks (MUCM gradient scaled by sigma_hat^2,
``predict`` ignored by Posterior, ``remake()`` adding un-scaled ``r`` ...).  It keeps the
reference's own numerical route (pdist/cdist, ``np.linalg.solve`` on the Cholesky factor,
``scipy.linalg.solve`` for the posterior) so that it is the same arithmetic, not merely the
same mathematics.

Parity status: PINNED against the real reference (see oracle/__init__.py).

``kind``: 0 = ``kernel`` (nugget form), 1 = ``kernel_alt_nug``.
"""
import numpy as np
import scipy.spatial.distance as _dist
from scipy import linalg as _sl

KERNEL, KERNEL_ALT = 0, 1


# ----------------------------------------------------------------------------- a1
def transform(hp):
    """_emulatorkernels.py:31-32 (same :104-105)."""
    return 2.0 * np.log(hp)


def untransform(theta):
    """_emulatorkernels.py:35-36 (same :108-109)."""
    return np.exp(np.asarray(theta, dtype=float) / 2.0)


# ----------------------------------------------------------------------------- a3/a4
def cov_exp_condensed(X, delta):
    """exp_save of _emulatorkernels.py:40-42 / :113-115 (condensed n(n-1)/2 vector)."""
    w = 1.0 / np.asarray(delta, dtype=float)
    return np.exp(-_dist.pdist(X * w, 'sqeuclidean'))


def cov_var(X, delta, nugget, kind=KERNEL, predict=True):
    """kernel.var _emulatorkernels.py:39-50; kernel_alt_nug.var :112-123."""
    E = cov_exp_condensed(X, delta)
    if kind == KERNEL:
        A = _dist.squareform((1.0 - nugget) * E)
        np.fill_diagonal(A, 1.0 if predict else 1.0 - nugget)
    else:
        A = _dist.squareform(E)
        np.fill_diagonal(A, 1.0 + nugget ** 2 if predict else 1.0)
    return A


def make_A(X, delta, nugget, kind=KERNEL, r=0.0, s2=1.0, predict=True):
    """Data.make_A _emulatorclasses.py:572-575 (r added only for alt_nugget)."""
    A = cov_var(X, delta, nugget, kind, predict)
    if kind == KERNEL_ALT:
        np.fill_diagonal(A, A.diagonal() + np.asarray(r, dtype=float) / s2)
    return A


# ----------------------------------------------------------------------------- a6/a7
def grad_delta_A(X, E_cond, delta, nugget, di, s2, kind=KERNEL):
    """grad_delta_A _emulatorkernels.py:53-63 / :126-136."""
    col = (X[:, di] * (1.0 / delta[di])).reshape(-1, 1)
    f = _dist.pdist(col, 'sqeuclidean')
    pref = (1.0 - nugget) * s2 if kind == KERNEL else s2
    return _dist.squareform(pref * f * E_cond)


def grad_nugget_A(X, E_cond, nugget, s2, kind=KERNEL):
    """grad_nugget_A _emulatorkernels.py:66-71 / :139-144."""
    if kind == KERNEL:
        return _dist.squareform((0.5 * (-nugget) * s2) * E_cond)
    f = np.zeros((X.shape[0], X.shape[0]))
    np.fill_diagonal(f, nugget ** 2 * s2)
    return f


# ----------------------------------------------------------------------------- a8
def cov_covar(XT, XV, delta, nugget, kind=KERNEL):
    """kernel.covar _emulatorkernels.py:75-79 / :148-152."""
    w = 1.0 / np.asarray(delta, dtype=float)
    C = np.exp(-_dist.cdist(XT * w, XV * w, 'sqeuclidean'))
    return (1.0 - nugget) * C if kind == KERNEL else C


# ----------------------------------------------------------------------------- a15
def make_H_linear(X, basis_inf=None, powers=None):
    """Data.make_H _emulatorclasses.py:558-566 for polynomial bases:
    column 0 is h_0(1.0)=1, column j is X[:, basis_inf[j-1]] ** powers[j-1]."""
    n, d = X.shape
    basis_inf = list(range(d)) if basis_inf is None else list(basis_inf)
    powers = [1] * len(basis_inf) if powers is None else list(powers)
    H = np.ones((n, 1 + len(basis_inf)))
    for j, (c, pw) in enumerate(zip(basis_inf, powers)):
        H[:, j + 1] = X[:, c] ** pw
    return H


# ----------------------------------------------------------------------------- shared
def _split_hp(x, d, has_sigma):
    """set_params semantics _emulatorkernels.py:20-24: delta = x[:d]; nugget = x[-1] if
    anything is left after delta (sigma stripped first for gp4ml, _emulatoroptimise.py:414)."""
    x = np.asarray(x, dtype=float)
    sigma = None
    if has_sigma:
        sigma, x = x[-1], x[:-1]
    delta = x[:d]
    nugget = x[-1] if x.size > d else None
    return delta, nugget, sigma


def _gls_pieces(A, H, y):
    """The block common to _emulatoroptimise.py:313-323, :390-399, :425-433, :498-504."""
    L = np.linalg.cholesky(A)
    w = np.linalg.solve(L, H)
    Q = w.T.dot(w)
    K = np.linalg.cholesky(Q)
    invA_f = np.linalg.solve(L.T, np.linalg.solve(L, y))
    invA_H = np.linalg.solve(L.T, np.linalg.solve(L, H))
    solve_K_HT = np.linalg.solve(K, H.T)
    B = np.linalg.solve(K.T, solve_K_HT.dot(invA_f))
    return L, Q, K, invA_f, invA_H, solve_K_HT, B


def _grad_term(temp, L, K, y, invA_f, invA_H, invA_H_dot_B, H_dot_B, solve_K_HT, factor):
    """One hyper-parameter's gradient, _emulatoroptimise.py:347-357 (mucm, ``factor``) and
    :452-460 (gp4ml, factor = 1)."""
    invA_gradHP = np.linalg.solve(L.T, np.linalg.solve(L, temp))
    sam = invA_gradHP.dot(invA_H_dot_B)
    return -0.5 * (
        -np.trace(invA_gradHP)
        + factor * (y.T.dot(invA_gradHP).dot(invA_f) + (-2 * y.T + H_dot_B).dot(sam))
        + np.trace(np.linalg.solve(K.T, solve_K_HT.dot(invA_gradHP)).dot(invA_H)))


# ----------------------------------------------------------------------------- a10
def loglikelihood_mucm(theta, X, y, H, kind=KERNEL, nugget_fixed=1e-4, r=0.0):
    """Optimize.loglikelihood_mucm _emulatoroptimise.py:305-378.
    Returns (LLH, grad, sigma_hat) or None when the Cholesky fails (:374-376)."""
    n, d = X.shape
    q = H.shape[1]
    x = untransform(theta)
    delta, nug, _ = _split_hp(x, d, has_sigma=False)
    nugget = nugget_fixed if nug is None else nug
    A = make_A(X, delta, nugget, kind, r, 1.0, True)
    E = cov_exp_condensed(X, delta)
    try:
        L, Q, K, invA_f, invA_H, solve_K_HT, B = _gls_pieces(A, H, y)
        invA_H_dot_B = invA_H.dot(B)
        sig2 = (1.0 / (n - q - 2.0)) * y.T.dot(invA_f - invA_H_dot_B)
        logdetA = 2.0 * np.sum(np.log(np.diag(L)))
        LLH = -0.5 * (-(n - q) * np.log(sig2) - logdetA - np.log(np.linalg.det(Q)))
        grad = np.empty(x.size)
        H_dot_B = H.dot(B).T
        factor = (n - q) / (sig2 * (n - q - 2))
        for i in range(d):
            temp = grad_delta_A(X, E, delta, nugget, i, sig2, kind)
            grad[i] = _grad_term(temp, L, K, y, invA_f, invA_H, invA_H_dot_B, H_dot_B, solve_K_HT, factor)
        if x.size == d + 1:
            temp = grad_nugget_A(X, E, nugget, sig2, kind)
            grad[x.size - 1] = _grad_term(temp, L, K, y, invA_f, invA_H, invA_H_dot_B, H_dot_B, solve_K_HT, factor)
    except np.linalg.LinAlgError:
        return None
    return LLH, grad, np.sqrt(sig2)


# ----------------------------------------------------------------------------- a9
def loglikelihood_gp4ml(theta, X, y, H, kind=KERNEL, nugget_fixed=1e-4, r=0.0):
    """Optimize.loglikelihood_gp4ml _emulatoroptimise.py:412-493.
    Returns (LLH, grad) or None on a non-PD matrix (:489-491)."""
    n, d = X.shape
    q = H.shape[1]
    x = untransform(theta)
    delta, nug, sigma = _split_hp(x, d, has_sigma=True)
    nugget = nugget_fixed if nug is None else nug
    s2 = sigma ** 2
    A = s2 * make_A(X, delta, nugget, kind, r, s2, True)
    E = cov_exp_condensed(X, delta)
    rvec = np.asarray(r, dtype=float)   # Data.r as used at :478, whatever the kernel kind
    try:
        L, Q, K, invA_f, invA_H, solve_K_HT, B = _gls_pieces(A, H, y)
        logdetA = 2.0 * np.sum(np.log(np.diag(L)))
        invA_H_dot_B = invA_H.dot(B)
        longexp = y.T.dot(invA_f - invA_H_dot_B)
        LLH = -0.5 * (-longexp - logdetA - np.log(_sl.det(Q)) - (n - q) * np.log(2.0 * np.pi))
        grad = np.empty(x.size)
        H_dot_B = H.dot(B).T
        for i in range(d):
            temp = grad_delta_A(X, E, delta, nugget, i, s2, kind)
            grad[i] = _grad_term(temp, L, K, y, invA_f, invA_H, invA_H_dot_B, H_dot_B, solve_K_HT, 1.0)
        if x.size == d + 2:
            temp = grad_nugget_A(X, E, nugget, s2, kind)
            grad[x.size - 2] = _grad_term(temp, L, K, y, invA_f, invA_H, invA_H_dot_B, H_dot_B, solve_K_HT, 1.0)
        temp = A.copy()                                      # :476-478 (A already times s2)
        np.fill_diagonal(temp, temp.diagonal() - rvec)
        grad[x.size - 1] = _grad_term(temp, L, K, y, invA_f, invA_H, invA_H_dot_B, H_dot_B, solve_K_HT, 1.0)
    except np.linalg.LinAlgError:
        return None
    return LLH, grad


# ----------------------------------------------------------------------------- a11/a12
def sigma_analytic_mucm(A, H, y):
    """Optimize.sigma_analytic_mucm _emulatoroptimise.py:382-408 (A = correlation matrix)."""
    n, q = H.shape
    L, Q, K, invA_f, invA_H, solve_K_HT, B = _gls_pieces(A, H, y)
    sig2 = (1.0 / (n - q - 2.0)) * y.T.dot(invA_f - invA_H.dot(B))
    return np.sqrt(sig2)


def optimalbeta(A, H, y):
    """Optimize.optimalbeta _emulatoroptimise.py:497-504."""
    L = np.linalg.cholesky(A)
    w = np.linalg.solve(L, H)
    Q = w.T.dot(w)
    K = np.linalg.cholesky(Q)
    invA_f = np.linalg.solve(L.T, np.linalg.solve(L, y))
    return np.linalg.solve(K.T, np.linalg.solve(K, H.T).dot(invA_f))


# ----------------------------------------------------------------------------- a16-a18
def posterior(Xs, Hs, X, y, H, A, beta, sigma, delta, nugget, kind=KERNEL, r_new=0.0):
    """Posterior.make_covar/make_mean/make_var _emulatorclasses.py:607-631.
    ``A`` is Dold.A exactly as left by training.remake() (correlation matrix + un-scaled r
    for alt nugget).  The prior of the new points is Data(x*).A = make_A() with s2=1,
    predict=True (:546-550); the Posterior ``predict`` flag is ignored (:595,:621).
    Returns (mean[m], var[m,m])."""
    covar = cov_covar(X, Xs, delta, nugget, kind)
    mean = Hs.dot(beta) + covar.T.dot(_sl.solve(A, y - H.dot(beta)))
    invA_H = _sl.solve(A, H)
    temp1 = Hs - covar.T.dot(invA_H)
    temp2 = H.T.dot(invA_H)
    Anew = make_A(Xs, delta, nugget, kind, r_new, 1.0, True)
    temp3 = Anew - covar.T.dot(_sl.solve(A, covar))
    var = sigma ** 2 * (temp3 + temp1.dot(_sl.solve(temp2, temp1.T)))
    return mean, var


def posterior_diag_chunked(Xs, Hs, X, y, H, A, beta, sigma, delta, nugget, kind=KERNEL, chunk=1000):
    """g.posterior called in chunks (emulatorfunctions.py:226-252), keeping mean and
    np.diag(var) as every large-m consumer does (history_match.py:117-118)."""
    m = Xs.shape[0]
    mean, var = np.empty(m), np.empty(m)
    for s in range(0, m, chunk):
        mu, V = posterior(Xs[s:s + chunk], Hs[s:s + chunk], X, y, H, A, beta, sigma, delta, nugget, kind)
        mean[s:s + chunk], var[s:s + chunk] = mu, np.diag(V)
    return mean, var


# ----------------------------------------------------------------------------- a21
def implausibility(means, variances, zs, var_extra, cm, maxno=1):
    """history_match.py:121-132 / :237-250 / :317-329.
    means/variances: [n_emul][m].  Returns (Imaxes[m,maxno] ascending, keep[m] bool for the
    flat routines' test Imaxes[r,-maxno] < cm, odp_count[maxno])."""
    means, variances = np.asarray(means), np.asarray(variances)
    n_emul, m = means.shape
    I2 = np.zeros((m, n_emul))
    for o in range(n_emul):
        for r in range(m):
            I2[r, o] = (means[o, r] - zs[o]) ** 2 / (variances[o, r] + var_extra[o])
    I = np.sqrt(I2)
    Imaxes = np.empty((m, maxno))
    odp_count = np.zeros(maxno, dtype=np.uint32)
    for r in range(m):
        Imaxes[r, :] = np.sort(np.partition(I[r, :], -maxno)[-maxno:])[-maxno:]
        for k in range(maxno):
            if Imaxes[r, -(k + 1)] < cm:
                odp_count[k] += 1
    keep = Imaxes[:, -((maxno - 1) + 1)] < cm
    return Imaxes, keep, odp_count


def implausibility_cells(Imaxes_per_cell, cm):
    """history_match.py:134-136: IMP_m = min over the LHC points, ODP_m = count/n."""
    out_imp, out_odp = [], []
    for Imaxes in Imaxes_per_cell:
        n, maxno = Imaxes.shape
        out_imp.append([np.amin(Imaxes[:, -(k + 1)]) for k in range(maxno)])
        out_odp.append([float(np.sum(Imaxes[:, -(k + 1)] < cm)) / float(n) for k in range(maxno)])
    return np.array(out_imp), np.array(out_odp)

"""Prediction-grid builder of the plotting layer (reference: gp_emu_uqsa/_emulatorplotting.py).

``make_inputs`` defines the grid ``g.plot`` predicts on (30x30 map or 900-point line); the drawing
itself is matplotlib and outside the rebuilt hot path: ``plotting`` draws when
matplotlib is importable and otherwise only reports what was computed."""
import numpy as _np


def make_inputs(dim, rows, cols, plot_dims, fixed_dims, fixed_vals, one_d, minmax):
    """Inputs for a 2-D map (rows x cols over plot_dims) or a 1-D line (rows*cols points), other
    inputs held at fixed_vals (reference :10-42)."""
    if dim >= 2 and not one_d:
        X1 = _np.linspace(minmax[0][0], minmax[0][1], rows)
        X2 = _np.linspace(minmax[1][0], minmax[1][1], cols)
        x_all = _np.zeros((rows * cols, dim))
        x_all[:, plot_dims[0]] = _np.repeat(X1, cols)
        x_all[:, plot_dims[1]] = _np.tile(X2, rows)
        if dim > 2:
            for i in range(len(fixed_dims)):
                x_all[:, fixed_dims[i]] = fixed_vals[i]
        return x_all
    npts = rows * cols
    x_all = _np.zeros((npts, max(dim, 1)))
    x_all[:, plot_dims[0] if dim >= 2 else 0] = _np.linspace(minmax[0][0], minmax[0][1], npts)
    if dim > 1:
        for i in range(len(fixed_dims)):
            x_all[:, fixed_dims[i]] = fixed_vals[i]
    return x_all


def plotting(dim, post, rows, cols, one_d, mean_or_var, minmax, x=[], y=[], labels=[]):
    """Draw the posterior mean or variance (reference :46-120) if matplotlib is present."""
    try:
        import matplotlib.pyplot as plt
    except ImportError:
        print("Plotting skipped: matplotlib is not installed (posterior values were computed on the GPU)")
        return
    diag = post.var_diag
    if dim >= 2 and not one_d:
        Z = (post.mean if mean_or_var == "mean" else diag).reshape(rows, cols).T
        fig = plt.figure()
        im = plt.imshow(Z, origin="lower", extent=(minmax[0][0], minmax[0][1], minmax[1][0], minmax[1][1]))
        plt.colorbar(im)
        if len(x):
            plt.scatter(x, y)
    else:
        xs = _np.linspace(minmax[0][0], minmax[0][1], rows * cols)
        post.interval()
        plt.plot(xs, post.mean if mean_or_var == "mean" else diag)
        if mean_or_var == "mean":
            plt.fill_between(xs, post.LI, post.UI, alpha=0.3)
        if len(x):
            plt.scatter(x, y)
    if labels:
        plt.xlabel(labels[0]); plt.ylabel(labels[1])
    plt.show()

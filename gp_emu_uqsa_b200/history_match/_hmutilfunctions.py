"""History-matching helpers (reference: gp_emu_uqsa/history_match/_hmutilfunctions.py): index
bookkeeping between emulators with different active inputs, data-file loading.  The matplotlib
helpers (make_plots / plot_options) are outside the rebuilt hot path.  The printed lines are the
reference's (users parse them); the bookkeeping itself is written with itertools / comprehensions."""
import itertools as _it

import numpy as _np


def make_sets(ai):
    """All pairs [i, j], i < j, of the active indices, in the order the reference's double loop meets them (:6-12)."""
    pairs = []
    for i, j in _it.product(ai, repeat=2):
        if i < j and [i, j] not in pairs:
            pairs.append([i, j])
    return pairs


def emulsetup(emuls):
    """(pairs, scaled minmax, original minmax) from the emulators' updated beliefs (:15-37).  As in the
    reference the pairs are those of the LAST emulator; the minmax dictionaries collect every emulator's inputs."""
    minmax, orig_minmax, sets = {}, {}, []
    for e in emuls:
        ai, mm = getattr(e.beliefs, "active_index", None), getattr(e.beliefs, "input_minmax", None)
        if ai is None or mm is None:
            print("ERROR: Emulator(s) were not previously trained and reconstructed "
                  "using updated beliefs files, "
                  "so they are missing 'active_index' and 'input_minmax'. Exiting.")
            raise SystemExit(1)
        sets = make_sets(ai)
        for idx, (lo, hi) in zip(ai, mm):
            minmax[str(idx)] = list((_np.array([lo, hi]) - lo) / (hi - lo))
            orig_minmax[str(idx)] = list(_np.array([lo, hi]))
    print("\nactive index pairs:", sets)
    print("\nminmax for active inputs:", minmax)
    print("original units minmax for active inputs:", orig_minmax)
    return sets, minmax, orig_minmax


def _numbered(keys):
    return {str(key): pos for pos, key in enumerate(keys)}


def ref_act(minmax):
    """active index -> column number in the combined input array (:40-48)."""
    act_ref = _numbered(sorted(minmax, key=int))
    print("\nrelate active_indices to integers:", act_ref)
    return act_ref


def ref_plt(act):
    plt_ref = _numbered(sorted(act))
    print("\nrelate restricted active_indices to subplot indices:", plt_ref)
    return plt_ref


def check_act(act, sets):
    if type(act) is not list:
        print("ERROR: 'act' argument must be a list, but", act, "was supplied. Exiting.")
        raise SystemExit(1)
    known = set(_it.chain.from_iterable(sets))
    for a in act:
        if a not in known:
            print("ERROR: index", a, "in 'act' is not an active_index of the emulator(s). Exiting.")
            raise SystemExit(1)
    return True


def load_datafiles(datafiles, orig_minmax):
    """Load inputs/outputs files and scale the inputs with the emulators' input_minmax (:126-139)."""
    try:
        sim_x, sim_y = _np.loadtxt(datafiles[0]), _np.loadtxt(datafiles[1])
    except FileNotFoundError:
        print("ERROR: datafile(s)", datafiles, "for inputs and/or outputs not found. Exiting.")
        raise SystemExit(1)
    for key, (lo, hi) in orig_minmax.items():
        col = int(key)
        sim_x[:, col] = (sim_x[:, col] - lo) / (hi - lo)
    return sim_x, sim_y

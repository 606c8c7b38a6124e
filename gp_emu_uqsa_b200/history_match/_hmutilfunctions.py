"""History-matching helpers (reference: gp_emu_uqsa/history_match/_hmutilfunctions.py): index
bookkeeping between emulators with different active inputs, data-file loading.  The matplotlib
helpers (make_plots / plot_options) are outside the rebuilt hot path."""
import numpy as _np


def make_sets(ai):
    """All pairs [i, j], i < j, of the active indices (:6-12)."""
    sets = []
    for i in ai:
        for j in ai:
            if i != j and i < j and [i, j] not in sets:
                sets.append([i, j])
    return sets


def emulsetup(emuls):
    """(pairs, scaled minmax, original minmax) from the emulators' updated beliefs (:15-37)."""
    minmax, orig_minmax = {}, {}
    sets = []
    for e in emuls:
        try:
            ai = e.beliefs.active_index
            mm = e.beliefs.input_minmax
        except AttributeError:
            print("ERROR: Emulator(s) were not previously trained and reconstructed "
                  "using updated beliefs files, "
                  "so they are missing 'active_index' and 'input_minmax'. Exiting.")
            raise SystemExit(1)
        sets = make_sets(ai)
        for i in range(len(ai)):
            minmax[str(ai[i])] = list((_np.array(mm[i]) - mm[i][0]) / (mm[i][1] - mm[i][0]))
            orig_minmax[str(ai[i])] = list((_np.array(mm[i])))
    print("\nactive index pairs:", sets)
    print("\nminmax for active inputs:", minmax)
    print("original units minmax for active inputs:", orig_minmax)
    return sets, minmax, orig_minmax


def ref_act(minmax):
    """active index -> column number in the combined input array (:40-48)."""
    act_ref = {}
    for count, key in enumerate(sorted(minmax.keys(), key=lambda x: int(x))):
        act_ref[key] = count
    print("\nrelate active_indices to integers:", act_ref)
    return act_ref


def ref_plt(act):
    plt_ref = {}
    for count, key in enumerate(sorted(act)):
        plt_ref[str(key)] = count
    print("\nrelate restricted active_indices to subplot indices:", plt_ref)
    return plt_ref


def check_act(act, sets):
    if type(act) is not list:
        print("ERROR: 'act' argument must be a list, but", act, "was supplied. Exiting.")
        raise SystemExit(1)
    for a in act:
        if a not in [item for sublist in sets for item in sublist]:
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
    for key in orig_minmax.keys():
        sim_x[:, int(key)] = (sim_x[:, int(key)] - orig_minmax[key][0]) / (orig_minmax[key][1] - orig_minmax[key][0])
    return sim_x, sim_y

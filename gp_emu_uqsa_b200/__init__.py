"""gp_emu_uqsa_b200 -- B200-native drop-in for the dense Gaussian-process hot path of GP_emu_UQSA.

    import gp_emu_uqsa_b200 as g
    E = g.setup("config"); g.train(E); mean, var = g.posterior(E, x)
    import gp_emu_uqsa_b200.history_match as h, gp_emu_uqsa_b200.sensitivity as s

Same call surface as the reference package (gp_emu_uqsa/__init__.py:6-7); all arithmetic runs in
libgpe_b200.so (include/gpe_b200.h) on an sm_100 GPU -- there is no CPU fallback."""
from .emulatorfunctions import *  # noqa: F401,F403

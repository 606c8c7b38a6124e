from .noise_fit import *  # noqa: F401,F403

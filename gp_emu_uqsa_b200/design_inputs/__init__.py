from .design_inputs import *  # noqa: F401,F403

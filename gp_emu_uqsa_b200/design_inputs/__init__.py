from .design_inputs import *  # noqa: F401,F403
from .design_inputs import device_criterion  # noqa: F401

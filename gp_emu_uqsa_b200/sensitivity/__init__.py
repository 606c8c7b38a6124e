from .sensitivityfunctions import *  # noqa: F401,F403

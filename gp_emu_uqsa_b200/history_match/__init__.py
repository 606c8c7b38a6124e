from .history_match import *  # noqa: F401,F403

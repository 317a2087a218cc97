"""recman.th.BestModelFinder: re-export (the reference keeps it at recman/tf/BestModelFinder.py)."""
from recman_b200.th.BestModelFinder import *  # noqa: F401,F403
from recman_b200.th.BestModelFinder import BestModelFinder  # noqa: F401

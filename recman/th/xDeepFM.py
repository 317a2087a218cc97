"""recman.th.xDeepFM -> recman_b200.th.xDeepFM (drop-in path; the reference's recman/th/ is an empty stub)."""
from recman_b200.th.xDeepFM import *  # noqa: F401,F403
from recman_b200.th.xDeepFM import xDeepFM  # noqa: F401

"""recman.th.DeepFM -> recman_b200.th.DeepFM (drop-in path; the reference's recman/th/ is an empty stub)."""
from recman_b200.th.DeepFM import *  # noqa: F401,F403
from recman_b200.th.DeepFM import DeepFM  # noqa: F401

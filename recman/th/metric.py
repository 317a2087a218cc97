"""recman.th.metric -> recman_b200.th.metric (drop-in path; the reference's recman/th/ is an empty stub)."""
from recman_b200.th.metric import *  # noqa: F401,F403

"""recman.th.layers -> recman_b200.th.layers (drop-in path; the reference's recman/th/ is an empty stub)."""
from recman_b200.th.layers import *  # noqa: F401,F403

"""recman.th.DCN -> recman_b200.th.DCN (drop-in path; the reference's recman/th/ is an empty stub)."""
from recman_b200.th.DCN import *  # noqa: F401,F403
from recman_b200.th.DCN import DCN  # noqa: F401

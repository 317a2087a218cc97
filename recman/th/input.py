"""recman.th.input -> recman_b200.th.input (drop-in path; the reference's recman/th/ is an empty stub)."""
from recman_b200.th.input import *  # noqa: F401,F403

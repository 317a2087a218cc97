"""recman.th.DeepModel -> recman_b200.th.DeepModel (drop-in path; the reference's recman/th/ is an empty stub)."""
from recman_b200.th.DeepModel import *  # noqa: F401,F403
from recman_b200.th.DeepModel import DeepModel  # noqa: F401

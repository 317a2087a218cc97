"""recman.th.hparams -> recman_b200.th.hparams (drop-in path; the reference's recman/th/ is an empty stub)."""
from recman_b200.th.hparams import *  # noqa: F401,F403

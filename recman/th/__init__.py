from recman_b200.th import *  # noqa: F401,F403
from recman_b200.th import DCN, DeepFM, DeepModel, xDeepFM  # noqa: F401

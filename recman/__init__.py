"""Drop-in import path: ``recman.th`` resolves to the B200-native implementation in ``recman_b200.th``."""

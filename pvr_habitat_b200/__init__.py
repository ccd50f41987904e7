"""pvr_habitat_b200 — B200-native (sm_100a) implementation of pvr_habitat's PVR-embed -> BC-train hot path."""
__version__ = "0.1.0"

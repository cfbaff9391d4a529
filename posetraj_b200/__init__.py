"""posetraj_b200 — B200-native (sm_100a) implementation of PoseTraj's denoising hot path."""
__version__ = "0.1.0"

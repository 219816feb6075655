"""pbe_b200 — B200-native (sm_100a) denoising hot path for Paint-by-Example (zhanwenchen/pbe)."""
__version__ = "0.1.0"

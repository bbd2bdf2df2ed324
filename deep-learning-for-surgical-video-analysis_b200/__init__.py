"""surgvid_b200 — B200-native LFB extraction path (MiT-EVP encoder + SegFormer embedding head + MS-TCN)."""
__version__ = "0.1.0"

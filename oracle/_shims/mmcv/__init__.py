"""Minimal stand-in for `mmcv` (absent). TEST INFRASTRUCTURE ONLY."""

"""Shim for `timm.models` (reference import: mix_transformer_evp.py:7)."""


def register_model(fn):
    return fn

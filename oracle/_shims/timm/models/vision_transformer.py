"""Shim for `timm.models.vision_transformer` (reference import: mix_transformer_evp.py:8)."""


def _cfg(url="", **kwargs):
    return {"url": url, **kwargs}

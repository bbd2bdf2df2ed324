"""Minimal stand-in for the `timm` package (absent from this image).

TEST INFRASTRUCTURE ONLY. Exists so the *unmodified* reference modules under
/root/reference can be imported as the parity oracle. Only the
three names `mix_transformer_evp.py:6-8` imports are provided.
"""

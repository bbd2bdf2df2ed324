"""`models` package mirroring the layout the reference drivers import from
(`from models.mix_transformer_evp import mit_b3_evp`, generate_evp_LFB.py:21)."""
from .mix_transformer_evp import (MixVisionTransformerEVP, mit_b0_evp, mit_b1_evp, mit_b2_evp, mit_b3_evp, mit_b4_evp,  # noqa: F401
                                  mit_b5_evp)

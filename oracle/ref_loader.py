"""Import the UNMODIFIED reference modules from /root/reference as the parity oracle.

TEST INFRASTRUCTURE ONLY — never imported by the product package. Only `tests/`, the golden
generator (`tests/golden/make_golden.py`) and `bench.py --impl reference` use it, and only where
`/root/reference` exists (this container; it does not exist on the GPU box).

The GitHub snapshot is flat, but the code expects a `models/` package
(`generate_evp_LFB.py:21`, relative import at `mix_transformer_evp.py:12`) and a top-level
`visualizer` module (`mix_transformer_evp.py:69`).  We recreate that layout with symlinks in a
temporary directory — no reference source is copied into the repo.
"""
from __future__ import annotations

import importlib
import os
import sys
import tempfile

REFERENCE_ROOT = os.environ.get("SURGVID_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_shims")
_cache = {}


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "mix_transformer_evp.py"))


def _layout_dir() -> str:
    if "dir" in _cache:
        return _cache["dir"]
    d = tempfile.mkdtemp(prefix="surgvid_ref_")
    pkg = os.path.join(d, "refpkg_models")
    os.makedirs(pkg)
    open(os.path.join(pkg, "__init__.py"), "w").close()
    for f in ("mix_transformer_evp.py", "segformer_head.py"):
        os.symlink(os.path.join(REFERENCE_ROOT, f), os.path.join(pkg, f))
    os.symlink(os.path.join(REFERENCE_ROOT, "visualizer.py"), os.path.join(d, "visualizer.py"))
    os.symlink(os.path.join(REFERENCE_ROOT, "mstcn.py"), os.path.join(d, "ref_mstcn.py"))
    _cache["dir"] = d
    return d


def load_reference():
    """Returns (mix_transformer_evp module, mstcn module) of the real reference."""
    if "mods" in _cache:
        return _cache["mods"]
    if not reference_available():
        raise FileNotFoundError(f"reference not found at {REFERENCE_ROOT}")
    d = _layout_dir()
    added = []
    for p in (d, _SHIMS):
        if p not in sys.path:
            sys.path.insert(0, p)
            added.append(p)
    # the reference pins a module-level `device` from torch.cuda.is_available() (mix_transformer_evp.py:463);
    # the oracle always runs on CPU, so hide CUDA while importing.
    import torch

    real = torch.cuda.is_available
    torch.cuda.is_available = lambda: False
    try:
        evp = importlib.import_module("refpkg_models.mix_transformer_evp")
        mstcn = importlib.import_module("ref_mstcn")
    finally:
        torch.cuda.is_available = real
    _cache["mods"] = (evp, mstcn)
    return evp, mstcn


def patch_input_size(model, H: int, W: int):
    """`forward_features` hard-codes `view(-1,3,224,224)` (mix_transformer_evp.py:354-355).  For the
    480x854 stress config we wrap it with a generalised view; everything else is size-agnostic (F8)."""
    import types

    def forward_features(self, x, y):
        x = x.view(-1, 3, H, W)
        y = y.view(-1, 3, H, W)
        B = x.shape[0]
        outs = []
        hc = self.prompt_generator.init_prompts(y)
        for s in range(4):
            pe = getattr(self, f"patch_embed{s + 1}")
            x, h, w = pe(x)
            prompt = self.prompt_generator.init_prompt(x, hc[s], s + 1)
            for i, blk in enumerate(getattr(self, f"block{s + 1}")):
                x = self.prompt_generator.get_prompt(x, prompt, s + 1, i)
                x = blk(x, h, w)
            x = getattr(self, f"norm{s + 1}")(x)
            x = x.reshape(B, h, w, -1).permute(0, 3, 1, 2).contiguous()
            outs.append(x)
        return outs

    model.forward_features = types.MethodType(forward_features, model)
    return model

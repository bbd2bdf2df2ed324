"""Pins the oracle (oracle/*.py, a CPU restatement of the reference) against
  (a) the committed golden vectors, which are outputs of the REAL reference (tests/golden/make_golden.py), and
  (b) the real reference itself when /root/reference is present (this container only).
The reference ships no tests or golden vectors of its own."""
import os

import numpy as np
import pytest
import torch

import surgvid_b200  # noqa: F401
from oracle import evp_oracle as EO
from oracle import mstcn_oracle as MO
from oracle.ref_loader import load_reference, reference_available
from surgvid_b200 import synthetic as S

CFG = S.EVP_CONFIGS["mit_b3_evp"]


def _tap_sample(t, n=2048):
    flat = t.reshape(-1)
    idx = torch.linspace(0, flat.numel() - 1, n).long()
    return flat[idx].numpy()


@pytest.mark.parametrize("mode", ["ref_init", "stress"])
def test_evp_oracle_matches_golden(golden_dir, mode):
    g = np.load(os.path.join(golden_dir, f"evp_b3_{mode}_224.npz"))
    sd = S.synth_state_dict(S.evp_key_shapes("mit_b3_evp"), seed=int(g["weight_seed"]), mode=mode)
    x, seg, flow = S.synth_frames(int(g["n"]), seed=int(g["input_seed"]))
    taps = {}
    feats = EO.evp_forward(sd, CFG, x, seg, flow, taps=taps)
    assert np.abs(feats.numpy() - g["feats"]).max() < 5e-5
    for k in ("stage1_tokens", "stage2_tokens", "stage3_tokens", "stage4_tokens", "fused3_tokens", "fused4_tokens"):
        assert np.abs(_tap_sample(taps[k]) - g[k]).max() < 1e-3, k
    assert np.abs(EO.evp_forward(sd, CFG, x, seg, None).numpy() - g["feats_noflow"]).max() < 5e-5
    y, y_ant = EO.evp_forward(sd, CFG, x, seg, flow, return_features=False)
    assert np.abs(y.numpy() - g["y"]).max() < 5e-5 and np.abs(y_ant.numpy() - g["y_ant"]).max() < 5e-5


def test_evp_oracle_matches_golden_480x854(golden_dir):
    g = np.load(os.path.join(golden_dir, "evp_b3_stress_480x854.npz"))
    sd = S.synth_state_dict(S.evp_key_shapes("mit_b3_evp"), seed=int(g["weight_seed"]), mode="stress")
    x, seg, flow = S.synth_frames(1, seed=int(g["input_seed"]), H=480, W=854)
    feats = EO.evp_forward(sd, CFG, x, seg, flow)
    assert np.abs(feats.numpy() - g["feats"]).max() < 1e-4


@pytest.mark.parametrize("f_maps", [32, 64])
@pytest.mark.parametrize("mode", ["ref_init", "stress", "phase"])
def test_mstcn_oracle_matches_golden(golden_dir, f_maps, mode):
    g = np.load(os.path.join(golden_dir, f"mstcn_f{f_maps}_{mode}_T700.npz"))
    sd = S.synth_mstcn_state_dict(2, 8, f_maps, 2048, 14, seed=int(g["weight_seed"]), mode=mode)
    feats = S.synth_lfb_features(int(g["T"]), seed=int(g["feat_seed"]))
    out = MO.mstcn_forward(sd, feats.unsqueeze(0).transpose(2, 1))
    scale = max(1.0, float(np.abs(g["logits"]).max()))
    assert np.abs(out.numpy() - g["logits"]).max() < 2e-5 * scale
    taps = MO.mstcn_forward_taps(sd, feats)  # the causal-tap form the CUDA kernel implements
    assert np.abs(taps.numpy() - g["logits"][:, 0]).max() < 1e-4 * scale


def test_mstcn_phase_weights_give_nondegenerate_argmax(golden_dir):
    g = np.load(os.path.join(golden_dir, "mstcn_f32_phase_T700.npz"))
    hist = np.bincount(g["logits"][-1, 0, :7].argmax(0), minlength=7)
    assert (hist > 0).sum() >= 5 and hist.max() < 0.6 * hist.sum()


def test_mstcn_causality():
    sd = S.synth_mstcn_state_dict(mode="stress")
    a = S.synth_lfb_features(600, seed=5)
    b = a.clone()
    b[400:] = S.synth_lfb_features(200, seed=6)
    oa, ob = MO.mstcn_forward_taps(sd, a), MO.mstcn_forward_taps(sd, b)
    assert torch.equal(oa[..., :400], ob[..., :400])


@pytest.mark.skipif(not reference_available(), reason="/root/reference not present (GPU box)")
def test_oracle_matches_live_reference():
    evp, mstcn = load_reference()
    torch.manual_seed(0)
    m = evp.mit_b3_evp().eval()
    sd = S.synth_state_dict(S.evp_key_shapes("mit_b3_evp"), seed=3, mode="stress")
    assert list(sd.keys()) == list(m.state_dict().keys())
    m.load_state_dict(sd, strict=True)
    x, seg, flow = S.synth_frames(1, seed=21)
    with torch.no_grad():
        ref = m(x, seg, flow, return_features=True)
    assert (EO.evp_forward(sd, CFG, x, seg, flow) - ref).abs().max() < 5e-5
    # survey known answer: the reference module's own init under manual_seed(0)
    torch.manual_seed(0)
    mm = mstcn.MultiStageModel_S(2, 8, 32, 2048, 14, True).eval()
    gen = torch.Generator().manual_seed(1234)
    xx = torch.randn(1, 2300, 2048, generator=gen).transpose(2, 1)
    with torch.no_grad():
        o = mm(xx)
    assert abs(float(o.double().sum()) - (-4348.2007)) < 5e-3
    mine = MO.mstcn_forward({k: v for k, v in mm.state_dict().items()}, xx)
    assert (mine - o).abs().max() < 1e-5

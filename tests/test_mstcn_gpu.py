"""Parity of the CUDA MS-TCN (drop-in MultiStageModel_S -> torch.ops.surgvid.mstcn_forward -> C ABI), fp32:
logits max-abs <= 1e-4 (scaled by max|ref| when logits are large) vs golden / oracle, argmax agreement >= 99.9 % on a
non-degenerate class histogram (SURVEY.md §8d parity gates)."""
import os

import numpy as np
import pytest
import torch

import surgvid_b200  # noqa: F401
from oracle import mstcn_oracle as MO
from surgvid_b200 import lfb
from surgvid_b200 import synthetic as S
from surgvid_b200.mstcn import MultiStageModel_S

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _model(f_maps, mode, seed=1):
    m = MultiStageModel_S(2, 8, f_maps, 2048, 14, True)
    sd = S.synth_mstcn_state_dict(2, 8, f_maps, 2048, 14, seed=seed, mode=mode)
    m.load_state_dict(sd, strict=True)
    return m.to(DEV).eval(), sd


@pytest.mark.parametrize("f_maps", [32, 64])
@pytest.mark.parametrize("mode", ["ref_init", "stress", "phase"])
def test_matches_golden(golden_dir, f_maps, mode):
    g = np.load(os.path.join(golden_dir, f"mstcn_f{f_maps}_{mode}_T700.npz"))
    m, _ = _model(f_maps, mode, int(g["weight_seed"]))
    feats = S.synth_lfb_features(int(g["T"]), seed=int(g["feat_seed"])).to(DEV)
    long_feature = feats.unsqueeze(0)                 # [1, T, 2048]  (trans_SV_output.py:271)
    video_fe = long_feature.transpose(2, 1)           # [1, 2048, T] non-contiguous view (:272)
    with torch.no_grad():
        out = m.forward(video_fe)
    torch.cuda.synchronize()
    assert out.shape == (2, 1, 14, 700)
    ref = torch.from_numpy(g["logits"])
    err = float((out.cpu() - ref).abs().max())
    print(f"[parity] mstcn f{f_maps} {mode}: max-abs {err:.3e} (max|ref| {float(ref.abs().max()):.2f})")
    assert err <= 1e-4 * max(1.0, float(ref.abs().max()))
    last = out[-1].squeeze(1)                          # the reference's `[-1]` + squeeze (:279-280)
    assert last.shape == (1, 14, 700)


def test_batched_videos_equal_per_video_and_argmax_agreement():
    m, sd = _model(32, "phase")
    lengths = [700, 1531, 64, 2300, 1, 3, 255, 511, 513]
    feats = [S.synth_lfb_features(T, seed=100 + i) for i, T in enumerate(lengths)]
    allf = torch.cat(feats).to(DEV)
    with torch.no_grad():
        outs = lfb.run_mstcn_per_video(m, allf, lengths)
    agree = total = 0
    hist = np.zeros(7, dtype=np.int64)
    margins = []
    for f, o in zip(feats, outs):
        ref = MO.mstcn_forward(sd, f.unsqueeze(0).transpose(2, 1))[-1, 0]   # [14, T]
        got = o.cpu()
        assert float((got - ref).abs().max()) <= 1e-4 * max(1.0, float(ref.abs().max()))
        a, b = got[:7].argmax(0), ref[:7].argmax(0)
        agree += int((a == b).sum()); total += a.numel()
        hist += np.bincount(b.numpy(), minlength=7)
        top2 = ref[:7].topk(2, dim=0).values
        margins.append((top2[0] - top2[1]))
        # single-video call gives bit-identical logits to the batched call (videos never see each other)
        with torch.no_grad():
            solo = m.forward(f.to(DEV).unsqueeze(0).transpose(2, 1))[-1, 0]
        assert torch.equal(solo, o)
    margins = torch.cat(margins)
    print(f"[parity] argmax agreement {agree}/{total}; class histogram {hist.tolist()}; top-2 margin median {float(margins.median()):.3f} "
          f"p1 {float(margins.kthvalue(max(1, margins.numel() // 100)).values):.4f}")
    assert (hist > 0).sum() >= 5, "degenerate class histogram makes the agreement metric vacuous"
    assert agree / total >= 0.999


def test_causality_on_device():
    m, _ = _model(32, "stress")
    a = S.synth_lfb_features(900, seed=5).to(DEV)
    b = a.clone()
    b[600:] = S.synth_lfb_features(300, seed=6).to(DEV)
    with torch.no_grad():
        oa, ob = m.forward_videos(a, [900]), m.forward_videos(b, [900])
    assert torch.equal(oa[..., :600], ob[..., :600]) and not torch.equal(oa[..., 600:], ob[..., 600:])


def test_full_size_cholec80_batch_properties():
    """BASELINE configs[3] at full size: the 80 Cholec80-shaped sequences (184 578 frames) in ONE batched call.  Size-independent
    properties: (1) each video's logits are bit-identical to running that video alone (videos are independent, history never crosses
    an offset); (2) causality: truncating a video leaves the logits of the kept prefix unchanged."""
    lengths = S.cholec80_video_lengths()
    assert len(lengths) == 80 and sum(lengths) == 184578
    m = MultiStageModel_S(2, 8, 32, 2048, 14, True)
    m.load_state_dict(S.synth_mstcn_state_dict(2, 8, 32, 2048, 14, seed=7, mode="phase"))
    m = m.to(DEV).eval()
    feats = S.synth_lfb_features(sum(lengths), seed=31).to(DEV)
    with torch.no_grad():
        allv = m.forward_videos(feats, lengths)
        assert allv.shape == (2, 14, sum(lengths)) and bool(torch.isfinite(allv).all())
        offs = [0]
        for n in lengths:
            offs.append(offs[-1] + n)
        for v in (0, 17, 79):
            one = m.forward_videos(feats[offs[v]:offs[v + 1]], [lengths[v]])
            assert torch.equal(one, allv[:, :, offs[v]:offs[v + 1]])
        v = 40
        cut = lengths[v] // 2
        pre = m.forward_videos(feats[offs[v]:offs[v] + cut], [cut])
        assert torch.equal(pre, allv[:, :, offs[v]:offs[v] + cut])

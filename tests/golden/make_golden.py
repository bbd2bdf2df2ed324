"""Generate tests/golden/*.npz by running the REAL reference (/root/reference, via oracle/ref_loader.py)
on deterministic synthetic weights and inputs.  Run in the build container only:

    python tests/golden/make_golden.py

The reference ships no golden vectors of its own, so these outputs of the reference
itself are what pins the oracle (tests/test_oracle_cpu.py) and, through it, the CUDA path.  Inputs and
weights are NOT stored: they are rebuilt from seeds by surgvid_b200.synthetic on any machine.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import surgvid_b200  # noqa: E402,F401
from surgvid_b200 import synthetic as S  # noqa: E402
from oracle.ref_loader import load_reference, patch_input_size  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
TAP_SAMPLES = 2048


def tap_sample(t: torch.Tensor) -> np.ndarray:
    flat = t.reshape(-1)
    idx = torch.linspace(0, flat.numel() - 1, TAP_SAMPLES).long()
    return flat[idx].numpy().astype(np.float32)


def capture_taps(model):
    taps = {}
    hooks = []
    for s in range(4):
        def mk(name):
            def hook(_m, _i, o):
                taps[name] = o.detach()
            return hook
        hooks.append(getattr(model, f"norm{s + 1}").register_forward_hook(mk(f"stage{s + 1}_tokens")))
    hooks.append(model.cross_attn_s3.register_forward_hook(lambda m, i, o: taps.__setitem__("fused3_tokens", o.detach())))
    hooks.append(model.cross_attn_s4.register_forward_hook(lambda m, i, o: taps.__setitem__("fused4_tokens", o.detach())))
    return taps, hooks


def main():
    torch.set_num_threads(os.cpu_count())
    evp, mstcn = load_reference()
    name = "mit_b3_evp"
    model = getattr(evp, name)().eval()
    shapes = S.evp_key_shapes(name)
    for mode in ("ref_init", "stress"):
        sd = S.synth_state_dict(shapes, seed=0, mode=mode)
        model.load_state_dict(sd, strict=True)
        x, seg, flow = S.synth_frames(2, seed=7)
        taps, hooks = capture_taps(model)
        with torch.no_grad():
            feats = model(x, seg, flow, return_features=True)
            tap_arrays = {k: tap_sample(v) for k, v in taps.items()}
            feats_noflow = model(x, seg, None, return_features=True)
            y, y_ant = model(x, seg, flow, return_features=False)
        for h in hooks:
            h.remove()
        np.savez_compressed(os.path.join(OUT, f"evp_b3_{mode}_224.npz"), feats=feats.numpy(), feats_noflow=feats_noflow.numpy(),
                            y=y.numpy(), y_ant=y_ant.numpy(), weight_seed=0, input_seed=7, n=2, **tap_arrays)
        print(mode, "224:", feats.shape, float(feats.abs().mean()))
    # 480x854 stress configuration (BASELINE.json configs[4]); view generalised per F8
    sd = S.synth_state_dict(shapes, seed=0, mode="stress")
    model.load_state_dict(sd, strict=True)
    patch_input_size(model, 480, 854)
    x, seg, flow = S.synth_frames(1, seed=11, H=480, W=854)
    with torch.no_grad():
        feats = model(x, seg, flow, return_features=True)
    np.savez_compressed(os.path.join(OUT, "evp_b3_stress_480x854.npz"), feats=feats.numpy(), weight_seed=0, input_seed=11, n=1)
    print("480x854:", feats.shape, float(feats.abs().mean()))

    # MS-TCN (trans_SV_output.py:197 configuration) + a 64-map variant (tecno.py:105)
    for f_maps in (32, 64):
        for mode in ("ref_init", "stress", "phase"):
            m = mstcn.MultiStageModel_S(2, 8, f_maps, 2048, 14, True).eval()
            sd = S.synth_mstcn_state_dict(2, 8, f_maps, 2048, 14, seed=1, mode=mode)
            m.load_state_dict(sd, strict=True)
            feats = S.synth_lfb_features(700, seed=3)
            with torch.no_grad():
                out = m(feats.unsqueeze(0).transpose(2, 1))
            np.savez_compressed(os.path.join(OUT, f"mstcn_f{f_maps}_{mode}_T700.npz"), logits=out.numpy(), weight_seed=1, feat_seed=3, T=700)
            print("mstcn", f_maps, mode, out.shape, float(out.abs().mean()),
                  "class hist", np.bincount(out[-1, 0, :7].argmax(0).numpy(), minlength=7))
    # known answer quoted in SURVEY.md §8c (reference module's own init under manual_seed(0))
    torch.manual_seed(0)
    m = mstcn.MultiStageModel_S(2, 8, 32, 2048, 14, True).eval()
    g = torch.Generator().manual_seed(1234)
    xx = torch.randn(1, 2300, 2048, generator=g).transpose(2, 1)
    with torch.no_grad():
        o = m(xx)
    print("survey anchor: sum", float(o.double().sum()), "absmean", float(o.abs().mean()))


if __name__ == "__main__":
    main()
